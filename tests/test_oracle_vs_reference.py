"""Oracle vs the unmodified reference imported from /root/reference.  Runs only in the
build container (skipped where the mount is absent, e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import generator_ref as gr
from oracle import refshim, synth

pytestmark = pytest.mark.skipif(not refshim.available(), reason="/root/reference not mounted")


def test_generator_forward_matches_reference_modules(synthetic_sd):
    g = refshim.reference_generator(synthetic_sd)
    assert set(g.state_dict().keys()) == set(synthetic_sd.keys())
    x, mask, cam, ratio = synth.synthetic_slices(2, seed=5, per_sample_masks=True)
    with torch.no_grad():
        ref = g(x, mask, cam, ratio)
        got = gr.generator_forward(synthetic_sd, x, mask, cam, ratio)
    for r, o in zip(ref, got):
        assert float((r - o).abs().max()) <= 2e-6


def test_contextual_attention_dense_form_matches_reference():
    refshim.install()
    from models.inpaint_networks import ContextualAttention
    ca = ContextualAttention(False, ksize=3, stride=1, rate=2, fuse_k=3, softmax_scale=10, fuse=True)
    gen = torch.Generator().manual_seed(0)
    f = torch.relu(torch.randn(2, 64, 64, 64, generator=gen))
    mask = torch.zeros(2, 1, 256, 256)
    mask[0, :, 100:141] = 1
    mask[1, :, 30:71] = 1
    y_ref, flow_ref = ca(f, f, mask)
    y, off = gr.contextual_attention(f, mask)
    assert float((y - y_ref).abs().max()) <= 1e-4
    assert torch.equal(gr.flow_image(off), flow_ref)


def test_train_mode_power_iteration_matches_reference(synthetic_sd):
    sd = {k: v.clone() for k, v in synthetic_sd.items()}
    # perturb u so that the power iteration actually moves it
    gen = torch.Generator().manual_seed(1)
    for k in sd:
        if k.endswith("weight_u"):
            sd[k] = torch.nn.functional.normalize(sd[k] + 0.1 * torch.randn(sd[k].shape, generator=gen), dim=0)
    g = refshim.reference_generator({k: v.clone() for k, v in sd.items()})
    g.train()
    x, mask, cam, ratio = synth.synthetic_slices(1, seed=9)
    with torch.no_grad():
        ref = g(x, mask, cam, ratio)
        got = gr.generator_forward(sd, x, mask, cam, ratio, training=True)
    for r, o in zip(ref, got):
        assert float((r - o).abs().max()) <= 5e-6
    ref_sd = g.state_dict()
    for k in sd:
        assert float((ref_sd[k] - sd[k]).abs().max()) <= 1e-6, k
