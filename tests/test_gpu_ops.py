"""GPU parity of the stand-alone operators (through the C ABI) vs torch-CPU fp32 / the numpy oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import healthivert_gan_b200 as hv
from healthivert_gan_b200 import _lib, mask_ops
from healthivert_gan_b200._lib import check, ptr
from healthivert_gan_b200.inpaint_networks import conv2d_fused
from oracle import generator_ref as gr
from oracle import mask_ops_ref as mo
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cin,cout,k,stride,pad,dil,h,w,act", [
    (3, 16, 5, 1, 2, 1, 64, 64, "elu"),
    (16, 32, 3, 2, 1, 1, 64, 128, "elu"),
    (64, 64, 3, 1, 16, 16, 64, 64, "elu"),
    (64, 64, 3, 1, 4, 4, 64, 64, "relu"),
    (33, 32, 3, 1, 1, 1, 40, 72, "elu"),       # ragged extents
    (8, 1, 3, 1, 1, 1, 64, 64, "sigmoid"),
    (1, 64, 4, 2, 1, 1, 64, 64, "lrelu"),      # PatchGAN first layer
    (64, 128, 4, 1, 1, 1, 31, 31, "none"),     # PatchGAN stride-1 layer, odd extent
    (20, 70, 3, 1, 1, 1, 17, 33, "none"),      # cout not a multiple of the CTA tile
])
def test_conv2d_against_torch(cin, cout, k, stride, pad, dil, h, w, act):
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    x = torch.randn(2, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x, wt, b, stride=stride, padding=pad, dilation=dil)
    ref = {"elu": F.elu, "relu": F.relu, "sigmoid": torch.sigmoid, "none": lambda t: t,
           "lrelu": lambda t: F.leaky_relu(t, 0.2)}[act](ref)
    y = conv2d_fused([(x.cuda(), 0)], wt.cuda(), b.cuda(), k, stride, pad, dil, act, h, w).cpu()
    assert y.shape == ref.shape
    assert float((y - ref).abs().max()) <= 2e-5


def test_conv2d_fused_sources_upsample_concat_scalar():
    g = torch.Generator().manual_seed(0)
    a = torch.randn(2, 6, 16, 16, generator=g)
    cam = torch.rand(2, 1, 64, 64, generator=g)
    ratio = torch.rand(2, generator=g)
    wt = torch.randn(8, 8, 3, 3, generator=g) * 0.1
    b = torch.randn(8, generator=g)
    up = a.repeat_interleave(2, 2).repeat_interleave(2, 3)
    cat = torch.cat([up, cam[:, :, ::2, ::2], ratio.view(2, 1, 1, 1).expand(-1, -1, 32, 32)], 1)
    ref = F.elu(F.conv2d(cat, wt, b, padding=1))
    y = conv2d_fused([(a.cuda(), 1), (cam.cuda(), 2), (ratio.cuda(), 3)], wt.cuda(), b.cuda(), 3, 1, 1, 1, "elu", 32, 32)
    assert float((y.cpu() - ref).abs().max()) <= 2e-5


def test_conv2d_rejects_bad_arguments():
    x = torch.zeros(1, 3, 8, 8, device="cuda")
    w = torch.zeros(4, 5, 3, 3, device="cuda")
    d = _lib.hv_conv_desc()
    d.n, d.cin, d.cout, d.hin, d.win, d.k, d.stride, d.pad, d.dil, d.act, d.nsrc = 1, 5, 4, 8, 8, 3, 1, 1, 1, 1, 1
    d.src[0].ptr, d.src[0].channels, d.src[0].mode = x.data_ptr(), 3, 0
    with pytest.raises(_lib.HvError, match="cin"):
        _lib.check(_lib.lib().hv_conv2d_fwd(d, w.data_ptr(), None, x.data_ptr(), None, None))
    with pytest.raises(_lib.HvError, match="not built"):
        conv2d_fused([(x, 0)], torch.zeros(4, 3, 7, 7, device="cuda"), None, 7, 1, 3, 1, "elu", 8, 8)
    with pytest.raises(_lib.HvError, match="null"):
        _lib.check(_lib.lib().hv_conv2d_fwd(d, None, None, x.data_ptr(), None, None))


def test_spectral_norm_prepare_eval_and_train():
    g = torch.Generator().manual_seed(2)
    w = torch.randn(64, 32, 3, 3, generator=g)
    u = F.normalize(torch.randn(64, generator=g), dim=0)
    v = F.normalize(torch.randn(288, generator=g), dim=0)
    for training in (False, True):
        uu, vv = u.clone(), v.clone()
        if training:
            gr.sn_power_iteration(w, uu, vv)
        sigma = gr.sn_sigma(w, uu, vv)
        conv = hv.inpaint_networks.SNConv2d(32, 64, 3, 1, 1, 1)
        conv.load_state_dict({"weight_orig": w, "weight_u": u, "weight_v": v, "bias": torch.zeros(64)})
        conv = conv.cuda()
        w_eff, s = conv.effective_weight(training)
        assert abs(float(s) - float(sigma)) <= 1e-5 * abs(float(sigma))
        ref_eff = w / sigma
        assert float((w_eff.cpu() - ref_eff).abs().max()) <= 1e-5 * float(ref_eff.abs().max())
        assert float((conv.weight_u.cpu() - uu).abs().max()) <= 1e-6
        assert float((conv.weight_v.cpu() - vv).abs().max()) <= 1e-6


def test_contextual_attention_against_oracle():
    g = torch.Generator().manual_seed(0)
    f = torch.relu(torch.randn(3, 64, 64, 64, generator=g))
    mask = torch.zeros(3, 1, 256, 256)
    mask[0, :, 100:141] = 1
    mask[1, :, 30:71] = 1
    mask[2, :, 180:221] = 1
    ca = hv.ContextualAttention(True, ksize=3, stride=1, rate=2, fuse_k=3, softmax_scale=10, fuse=True)
    y_ref, off_ref, inter = gr.contextual_attention(f, mask, return_intermediates=True)
    fc = f.cuda()
    y, flow = ca(fc, fc, mask.cuda())
    assert float((y.cpu() - y_ref).abs().max()) <= 1e-4
    agree = (ca.last_offsets.cpu().long() == off_ref).float().mean()
    assert agree >= 0.999   # argmax ties between near-equal scores may break differently
    assert ((flow.cpu() - gr.flow_image(off_ref)).abs() > 1e-6).float().mean() < 0.01
    ca.per_sample_mask = True
    y2, _ = ca(fc, fc, mask.cuda())
    for i in range(3):
        yi, _ = gr.contextual_attention(f[i:i + 1], mask[i:i + 1])
        assert float((y2[i:i + 1].cpu() - yi).abs().max()) <= 1e-4
    # no fuse, empty mask (the reference's mask=None branch)
    ca2 = hv.ContextualAttention(True, ksize=3, stride=1, rate=2, fuse_k=3, softmax_scale=10, fuse=False)
    f1 = f[:1].cuda()
    y3, _ = ca2(f1, f1, None)
    y3_ref, _ = gr.contextual_attention(f[:1], torch.zeros(1, 1, 256, 256), fuse=False)
    assert float((y3.cpu() - y3_ref).abs().max()) <= 1e-4


@pytest.mark.parametrize("fuse", [True, False])
def test_contextual_attention_bf16_tensor_core_path(fuse):
    """tcgen05 similarity/paste GEMMs + fused fuse/softmax vs the fp32 oracle on bf16-representable features."""
    g = torch.Generator().manual_seed(5)
    f = torch.relu(torch.randn(3, 64, 64, 64, generator=g)).to(torch.bfloat16).float()
    mask = torch.zeros(3, 1, 256, 256)
    mask[0, :, 100:141] = 1
    mask[1, :, 30:71] = 1
    mask[2, :, 180:221] = 1
    ca = hv.ContextualAttention(True, ksize=3, stride=1, rate=2, fuse_k=3, softmax_scale=10, fuse=fuse)
    ca.precision = "bf16"
    ca.per_sample_mask = True
    fc = f.cuda()
    y, flow = ca(fc, fc, mask.cuda())
    torch.cuda.synchronize()
    agree = []
    for i in range(3):
        yi, oi = gr.contextual_attention(f[i:i + 1], mask[i:i + 1], fuse=fuse)
        err = float((y[i:i + 1].cpu() - yi).abs().max())
        scale = float(yi.abs().max())
        assert err <= 0.02 * scale + 1e-3, (i, err, scale)   # bf16 attention weights and bf16 output rounding
        agree.append((ca.last_offsets[i:i + 1].cpu().long() == oi).float().mean().item())
    assert min(agree) >= 0.98, agree
    assert flow.shape == (3, 3, 256, 256) and float(flow.min()) >= 0.0 and float(flow.max()) <= 1.0


def test_threshold_and_stitch_bit_exact():
    rng = np.random.Generator(np.random.PCG64(11))
    n = 6
    gen = rng.random((n, 1, 256, 256), dtype=np.float32) * 2 - 1
    real = rng.random((n, 1, 256, 256), dtype=np.float32) * 2 - 1
    x1 = np.array([100, 98, 102, 96, 20, 200])
    height = np.array([28, 30, 26, 33, 25, 31])
    x2 = x1 + height
    pred = np.array([0.5, 0.9, 0.99, 0.2, 0.7251, 0.775], np.float32)  # some > height/40, some below
    out, rows = mask_ops.stitch(torch.from_numpy(gen).cuda(), torch.from_numpy(real).cuda(), torch.from_numpy(pred).cuda(),
                                torch.from_numpy(x1), torch.from_numpy(x2), torch.from_numpy(height), 40, return_rows=True)
    for i in range(n):
        ref = mo.stitch_plane(gen[i, 0], real[i, 0], int(x1[i]), int(x2[i]), int(height[i]), pred[i], 40)
        assert np.array_equal(out[i, 0].cpu().numpy(), ref), i
        assert tuple(rows[i].tolist()) == mo.stitch_rows(pred[i], int(height[i]), int(x1[i]), 40)
    p = rng.random((3, 1, 64, 64), dtype=np.float32)
    p[0, 0, 0, :4] = [0.5, np.nextafter(np.float32(0.5), np.float32(1)), 0.49999997, 1.0]
    t = torch.from_numpy(p).cuda()
    assert np.array_equal(mask_ops.threshold(t).cpu().numpy(), mo.threshold_mask(p))
    assert np.array_equal(mask_ops.threshold(t, 20, as_u8=True).cpu().numpy(), (mo.threshold_mask(p) * 20).astype(np.uint8))


def test_sobel_and_edge_loss_bit_exact_on_masks():
    rng = np.random.Generator(np.random.PCG64(12))
    a = (rng.random((4, 1, 256, 256)) > 0.5).astype(np.float32)
    b = np.zeros_like(a)
    b[:, :, 90:130, 80:150] = 1
    sob = hv.Sobel().cuda()
    assert np.array_equal(sob(torch.from_numpy(a).cuda()).cpu().numpy(), mo.sobel_edges(a))
    assert np.array_equal(sob(torch.from_numpy(b).cuda()).cpu().numpy(), mo.sobel_edges(b))
    soft = rng.random((2, 1, 64, 64), dtype=np.float32)
    assert np.abs(sob(torch.from_numpy(soft).cuda()).cpu().numpy() - mo.sobel_edges(soft)).max() <= 1e-6
    loss, cnt = hv.edge_mse_loss(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    assert int(cnt) == mo.edge_xor_count(a, b)
    assert abs(float(loss) - mo.edge_loss(a, b)) <= 1e-3 * mo.edge_loss(a, b)
    assert float(hv.edge_mse_loss(torch.from_numpy(b).cuda(), torch.from_numpy(b).cuda())[0]) == 0.0


@pytest.mark.parametrize("axis", [2, 1])
def test_column_heights_and_rhlv_against_oracle(axis, golden_dir):
    import json
    import os
    v = np.load(os.path.join(golden_dir, "rhlv_label_0007_20.npz"))["label"]
    lab = (v == 20).astype(np.float64)
    fk = np.maximum(lab, np.roll(lab, -3, axis=0))
    fk[-3:] = lab[-3:]
    known = json.load(open(os.path.join(golden_dir, "rhlv_known.json")))[f"0007_20_axis{axis}"]
    c, ln = known["center"], known["length"]
    tf = torch.from_numpy(fk.astype(np.uint8)).cuda()
    tl = torch.from_numpy(lab.astype(np.uint8)).cuda()
    counts, meta = mask_ops.column_heights(tf, tl, axis, c - ln, c + ln)
    counts, meta = counts.cpu().numpy(), meta.cpu().numpy()
    win_f = np.moveaxis(np.take(fk, range(c - ln, c + ln), axis=axis), axis, 0)   # [s, rows, cols]
    win_l = np.moveaxis(np.take(lab, range(c - ln, c + ln), axis=axis), axis, 0)
    for s in range(2 * ln):
        if meta[s, 0]:
            assert np.array_equal(counts[s, 0], np.count_nonzero(win_f[s], axis=0)), ("raw fake counts", s)
            assert np.array_equal(counts[s, 4], np.count_nonzero(win_l[s], axis=0)), ("raw label counts", s)
    sl = [slice(None)] * 3
    sl[axis] = slice(c - ln, c + ln)
    recs = {r["z"]: r for r in mo.column_heights(fk[tuple(sl)], lab[tuple(sl)], axis)}
    for s in range(2 * ln):
        if s not in recs:
            assert meta[s, 0] == 0
            continue
        r = recs[s]
        assert tuple(meta[s, :5]) == (1, r["t1"], r["t2"], r["center_fake"], r["center_label"])
        ncols = counts.shape[2]
        seg = {"all": slice(0, ncols), "pre": slice(0, r["t1"]), "mid": slice(r["t1"], r["t2"]), "post": slice(r["t2"], ncols)}
        for i, k in enumerate(("all", "pre", "mid", "post")):
            assert np.array_equal(counts[s, i, seg[k]], r[k + "_fake"]), (s, k)
            assert np.array_equal(counts[s, 4 + i, seg[k]], r[k + "_label"]), (s, k)
    got = mask_ops.calculate_rhlv(fk, lab, c, ln, "0007_20", 0.7, axis=axis)
    assert np.allclose(got, known["rhlv"], rtol=0, atol=1e-12)
    # empty window / empty volume edge cases
    z = torch.zeros_like(tl)
    counts, meta = mask_ops.column_heights(z, tl, axis, 0, 4)
    assert int(meta[:, 0].sum()) == 0 and int(counts.sum()) == 0
    counts, meta = mask_ops.column_heights(tf, tl, axis, 5, 5)
    assert counts.shape[0] == 0
    with pytest.raises(_lib.HvError):
        mask_ops.column_heights(tf, tl, axis, 0, 10_000)


@pytest.mark.parametrize("axis", [2, 1])
def test_rhlv_table_from_nifti_files(axis, golden_dir, tmp_path):
    """grading.rhlv_table (process_datasets_to_excel, RHLV_quantification.py:150-195) on NIfTI files == the reference's known answers
    for the shipped 0007_20 label volume (fake := label shifted up by 3 rows, threshold 0.7, divisor 5)."""
    import json
    import os
    from healthivert_gan_b200 import grading, nifti
    v = np.load(os.path.join(golden_dir, "rhlv_label_0007_20.npz"))["label"].astype(np.float64)
    lab = (v == 20).astype(np.float64)
    fk = np.maximum(lab, np.roll(lab, -3, axis=0))
    fk[-3:] = lab[-3:]
    fake = np.where(fk > 0, 20.0, np.where(v == 20, 0.0, v))
    os.makedirs(tmp_path / "label"), os.makedirs(tmp_path / "label_fake")
    nifti.save(str(tmp_path / "label" / "0007_20.nii.gz"), v, np.eye(4))
    nifti.save(str(tmp_path / "label_fake" / "0007_20.nii.gz"), fake, np.eye(4))
    info = {"train": {"0007_19": 0}, "val": {"0007_20": 2}}          # 0007_19 has no files: skipped like the reference does
    rows = grading.rhlv_table(info, str(tmp_path / "label"), str(tmp_path / "label_fake"), length_divisor=5, height_threshold=0.7, axis=axis)
    known = json.load(open(os.path.join(golden_dir, "rhlv_known.json")))[f"0007_20_axis{axis}"]
    assert len(rows) == 1 and rows[0]["Vertebra"] == "0007_20" and rows[0]["Label"] == 2 and rows[0]["Dataset"] == "val"
    got = [rows[0][k] for k in ("All RHLV", "Pre RHLV", "Mid RHLV", "Post RHLV", "Relative Height Label")]
    assert np.allclose(got, known["rhlv"], rtol=0, atol=1e-12)
    grading.write_table(rows, str(tmp_path / "t.csv"))
    assert grading.read_table(str(tmp_path / "t.csv"))["Mid RHLV"][0] == rows[0]["Mid RHLV"]


def test_fused_post_forward_equals_the_separate_kernels():
    """hv_post_forward (one pass over the planes: thresholds, both stitches, centre crops, both Sobel maps, XOR-count edge loss) ==
    hv_threshold / hv_stitch / hv_masked_center / hv_sobel / hv_edge_xor_loss, bit for bit (models/pix2pix_model.py:201-264, :349)."""
    from healthivert_gan_b200 import mask_ops
    from healthivert_gan_b200.edge_operator import Sobel, edge_mse_loss
    g = torch.Generator().manual_seed(5)
    n, h, w = 5, 256, 256
    r = lambda *s: torch.rand(*s, generator=g).cuda()
    fine, coarse = r(n, 1, h, w), r(n, 1, h, w)
    fine[0, 0, 3, 7] = 0.5                                   # exactly 0.5 is NOT above the threshold
    x2s, x1s, real = r(n, 1, h, w) * 2 - 1, r(n, 1, h, w) * 2 - 1, r(n, 1, h, w) * 2 - 1
    real_mask = (r(n, 1, h, w) > 0.6).float()
    mask = torch.zeros(n, 1, h, w, device="cuda")
    mask[:, :, 100:140] = 1
    p1, p2 = r(n, 1), r(n, 1)
    x1 = torch.tensor([100, 98, 102, 96, 5], dtype=torch.int32).cuda()
    hh = torch.tensor([28, 30, 26, 33, 39], dtype=torch.int32).cuda()
    x2 = x1 + hh
    c0, c1 = w // 2 - 35, w // 2 + 35
    new = lambda: torch.empty(n, 1, h, w, device="cuda")
    outs = [new() for _ in range(8)]
    rows_f, rows_c = (torch.empty(n, 4, dtype=torch.int32, device="cuda") for _ in range(2))
    xor, loss = torch.empty(1, dtype=torch.int64, device="cuda"), torch.empty(1, device="cuda")
    check(_lib.lib().hv_post_forward(ptr(fine), ptr(coarse), ptr(x2s), ptr(x1s), ptr(real), ptr(real_mask), ptr(mask), ptr(p2.reshape(-1)),
                                     ptr(p1.reshape(-1)), ptr(x1), ptr(x2), ptr(hh), 40, c0, c1, *[ptr(o) for o in outs], ptr(rows_f),
                                     ptr(rows_c), ptr(xor), ptr(loss), n, h, w, _lib.stream()))
    torch.cuda.synchronize()
    fake_mask, coarse_bin, fake_B, fake_Bc, fake_loc, real_loc, real_e, fake_e = outs
    assert torch.equal(fake_mask, mask_ops.threshold(fine)) and torch.equal(coarse_bin, mask_ops.threshold(coarse))
    wf, rf = mask_ops.stitch(x2s, real, p2, x1, x2, hh, 40, return_rows=True)
    wc, rc = mask_ops.stitch(x1s, real, p1, x1, x2, hh, 40, return_rows=True)
    assert torch.equal(fake_B, wf) and torch.equal(fake_Bc, wc) and torch.equal(rows_f, rf) and torch.equal(rows_c, rc)
    loc = lambda t: torch.where((torch.arange(w, device="cuda") >= c0) & (torch.arange(w, device="cuda") < c1), t * mask, torch.zeros_like(t))
    assert torch.equal(fake_loc, loc(wf)) and torch.equal(real_loc, loc(real))
    sob = Sobel().cuda()
    assert torch.equal(real_e, sob(real_mask)) and torch.equal(fake_e, sob(mask_ops.threshold(fine)))
    want_loss, want_xor = edge_mse_loss(mask_ops.threshold(fine), real_mask)
    assert int(xor) == int(want_xor) and float(loss) == float(want_loss)
