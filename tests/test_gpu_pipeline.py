"""GPU parity of the uint8 host interface (hv_pipeline_*, healthivert_gan_b200.pipeline): the CUDA-graph pipeline must return
exactly what the tensor API returns for the tensors the reference driver would build from the same uint8 planes
(eval_3d_sagittal_twostage.py:84-98 in, :103-121 out), in both precisions, with and without the graph, for full and ragged
batches and with several slots in flight."""
import numpy as np
import pytest
import torch

import healthivert_gan_b200 as hv
from oracle import generator_ref as gr

pytestmark = pytest.mark.gpu


def _inputs(n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    ct = rng.integers(0, 256, size=(n, 256, 256), dtype=np.uint8)
    cam = rng.integers(0, 256, size=(n, 256, 256), dtype=np.uint8)
    r0 = rng.integers(60, 150, size=n).astype(np.int32)
    rows = np.stack([r0, r0 + 41], axis=1).astype(np.int32)       # eval convention: 41 rows [min_x, max_x]
    for i in range(n):
        ct[i, rows[i, 0]:rows[i, 1]] = 0                            # the composed CT plane is empty inside the mask rows
    ratio = rng.random(n).astype(np.float32)
    return ct, cam, rows, ratio


def _tensors(ct, cam, rows, ratio):
    """What run_model builds: ToTensor + Normalize(0.5, 0.5) on the CT, ToTensor on mask / CAM, the model gets 1 - CAM."""
    x = (torch.from_numpy(ct).float().div(255) - 0.5) / 0.5
    c = 1 - torch.from_numpy(cam).float().div(255)
    m = torch.zeros(ct.shape, dtype=torch.float32)
    for i, (a, b) in enumerate(rows):
        m[i, a:b] = 1
    return x[:, None], m[:, None], c[:, None], torch.from_numpy(ratio)


def _gen(sd, precision):
    g = hv.Generator({"input_dim": 1, "ngf": 16}, True)
    g.load_state_dict(sd)
    g = g.cuda().eval()
    g.precision = precision
    g.per_sample_mask = True      # the eval driver is batch 1: every slice has its own mask
    g.return_flow = False
    return g


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("use_graph", [True, False])
def test_pipeline_equals_tensor_api(synthetic_sd, precision, use_graph):
    g = _gen(synthetic_sd, precision)
    pipe = hv.SlicePipeline(g, batch=16, depth=3, use_graph=use_graph)
    cases = [(16, 1), (5, 2), (16, 3), (1, 4), (16, 5), (5, 6)]   # ragged batches re-use their own captured graph
    want = []
    for n, seed in cases:
        ct, cam, rows, ratio = _inputs(n, seed)
        with torch.no_grad():
            out = g(*[t.cuda() for t in _tensors(ct, cam, rows, ratio)])
        torch.cuda.synchronize()
        want.append((((out[3] + 1) * 127.5).to(torch.int32).to(torch.uint8).cpu().numpy()[:, 0], (out[1] > 0.5).cpu().numpy()[:, 0],
                     (out[0] > 0.5).cpu().numpy()[:, 0], out[5].cpu().numpy().reshape(-1), out[6].cpu().numpy().reshape(-1)))
    # three slots in flight at a time
    for base in range(0, len(cases), 3):
        for k in range(3):
            n, seed = cases[base + k]
            ct, cam, rows, ratio = _inputs(n, seed)
            s = pipe.slot(k)
            s.ct[:n], s.cam[:n], s.rows[:n], s.ratio[:n] = ct, cam, rows, ratio
            pipe.submit(k, n)
        for k in range(3):
            n, _ = cases[base + k]
            pipe.wait(k)
            s = pipe.slot(k)
            w = want[base + k]
            assert np.array_equal(s.ct_out[:n], w[0]), (base + k, "ct")
            assert np.array_equal(s.fine_mask[:n].astype(bool), w[1]), (base + k, "fine mask")
            assert np.array_equal(s.coarse_mask[:n].astype(bool), w[2]), (base + k, "coarse mask")
            assert np.array_equal(s.heights[0, :n], w[3]) and np.array_equal(s.heights[1, :n], w[4])
    pipe.close()


def test_forward_u8_against_oracle(synthetic_sd):
    """fp32 mode through the uint8 interface vs the CPU oracle on the tensors the reference driver builds: CT within one grey
    level (truncation of a value within 1e-3 * 127.5 of the oracle's), masks equal outside the 1e-5 guard band."""
    g = _gen(synthetic_sd, "fp32")
    ct, cam, rows, ratio = _inputs(3, 11)
    got_ct, got_fine, got_coarse, p1, p2 = g.forward_u8(ct, cam, rows, ratio)
    refs = []
    for i in range(3):   # batch-1 oracle calls = per-sample masks
        with torch.no_grad():
            refs.append(gr.generator_forward(synthetic_sd, *[t[i:i + 1] for t in _tensors(ct, cam, rows, ratio)], flow=False))
    for i, r in enumerate(refs):
        want_ct = ((r[3][0, 0].double() + 1) * 127.5).numpy()
        assert np.abs(got_ct[i].astype(np.float64) - np.floor(want_ct)).max() <= 1
        assert (got_ct[i] != np.floor(want_ct)).mean() <= 1e-3
        for got, ref in ((got_fine[i], r[1][0, 0].numpy()), (got_coarse[i], r[0][0, 0].numpy())):
            guard = np.abs(ref - 0.5) > 1e-5
            assert np.array_equal(got.astype(bool)[guard], (ref > 0.5)[guard])
        assert abs(p1[i] - float(r[5])) <= 5e-5 and abs(p2[i] - float(r[6])) <= 5e-5
    # weights changed -> forward_u8 follows (refresh re-prepares the plan the captured graph reads)
    with torch.no_grad():
        g.coarse_generator.conv1.conv.bias.add_(0.25)
    again = g.forward_u8(ct, cam, rows, ratio)[0]
    assert not np.array_equal(again, got_ct)
