#!/usr/bin/env python
"""Run every bandwidth-bound kernel of the path once at its BASELINE-config size (after one warm-up call each), so that

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/hbm_kernels.csv python tools/hbm_kernels.py

gives duration + DRAM traffic per kernel (ncu flushes the caches before each measured launch: cold, DRAM-bound numbers).
The script itself prints one JSON line {kernel name regex: algorithmic bytes per launch} that tools/hbm_table.py joins with the
ncu CSV into the achieved-GB/s table under profiles/ (north star: "achieved HBM GB/s against B200 peak for the elementwise,
edge and mask-scan kernels").  It also times every kernel WARM with CUDA events (20 back-to-back launches, inputs larger than
nothing: L2-resident) for the in-pipeline view.  usage: python tools/hbm_kernels.py [--events-only]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import healthivert_gan_b200 as hv
from healthivert_gan_b200 import _lib, mask_ops, train_ops as T
from healthivert_gan_b200._lib import check, ptr
from healthivert_gan_b200.edge_operator import Sobel, edge_mse_loss
from oracle import synth

N, H, W = 16, 256, 256
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
L = _lib.lib()
st = _lib.stream
algo = {}     # kernel-name regex -> algorithmic bytes per launch
warm = {}


def timed(name, fn, reps=20):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    warm[name] = e0.elapsed_time(e1) / reps * 1e3


g = torch.Generator().manual_seed(0)
px = N * H * W
seg_a = torch.rand(N, 1, H, W, generator=g).cuda()
seg_b = torch.rand(N, 1, H, W, generator=g).cuda()
mask_a, mask_b = (seg_a > 0.5).float(), (seg_b > 0.5).float()
ct = (torch.rand(N, 1, H, W, generator=g) * 2 - 1).cuda()
real = (torch.rand(N, 1, H, W, generator=g) * 2 - 1).cuda()
pred_h = torch.rand(N, generator=g).cuda()
x1 = torch.full((N,), 100, dtype=torch.int32).cuda()
hh = torch.full((N,), 28, dtype=torch.int32).cuda()
x2 = x1 + hh

# ---- A4 threshold / stitch, A5 Sobel / edge loss
timed("threshold_kernel", lambda: mask_ops.threshold(seg_a))
algo["threshold_kernel"] = px * 8                       # fp32 in, fp32 out
timed("stitch_kernel", lambda: mask_ops.stitch(ct, real, pred_h, x1, x2, hh, 40))
algo["stitch_kernel"] = px * 12                         # generated + real planes in, stitched plane out
sob = Sobel().cuda()
timed("sobel_kernel", lambda: sob(mask_a))
algo["sobel_kernel"] = px * 8
timed("edge_loss_kernel", lambda: edge_mse_loss(mask_a, mask_b))
algo["edge_loss_kernel"] = px * 8                       # two mask planes in, a scalar out

# ---- A4 + A5 fused: the whole tail of Pix2PixModel.forward + the edge loss in one pass (hv_post_forward)
new = lambda: torch.empty(N, 1, H, W, device=dev)
post_out = [new() for _ in range(8)]
rows_f, rows_c = (torch.empty(N, 4, dtype=torch.int32, device=dev) for _ in range(2))
xor_c, loss_c = torch.empty(1, dtype=torch.int64, device=dev), torch.empty(1, device=dev)
mask_rows = torch.zeros(N, 1, H, W, device=dev)
mask_rows[:, :, 100:140] = 1
timed("post_forward_kernel", lambda: check(L.hv_post_forward(
    ptr(seg_a), ptr(seg_b), ptr(ct), ptr(real), ptr(real), ptr(mask_b), ptr(mask_rows), ptr(pred_h), ptr(pred_h), ptr(x1), ptr(x2), ptr(hh),
    40, W // 2 - 35, W // 2 + 35, *[ptr(o) for o in post_out], ptr(rows_f), ptr(rows_c), ptr(xor_c), ptr(loss_c), N, H, W, st())))
algo["post_forward_kernel"] = px * (7 + 8) * 4          # seven fp32 planes in, eight out

# ---- A10 column heights (RHLV): 256^3 uint8 volumes, sagittal and coronal window of 2 * (extent // 5) slices
label, _, _ = synth.synthetic_volume(seed=2, depth=256)
lab = torch.as_tensor((label == 20).astype(np.uint8)).cuda()
fake = torch.as_tensor(np.maximum(label == 20, np.roll(label == 20, -3, axis=0)).astype(np.uint8)).cuda()
for axis in (2, 1):
    loc = np.where(label == 20)[axis]
    c, ln = int(loc.mean()), int((loc.max() - loc.min()) // 5)
    timed(f"column_heights axis {axis}", lambda: mask_ops.column_heights(fake, lab, axis, c - ln, c + ln))
    algo.setdefault("column_count_kernel", []).append(2 * 256 * 256 * 2 * ln)     # two u8 volumes, the window's slices
algo["column_split_kernel"] = None

# ---- A7 BatchNorm + LeakyReLU of the PatchGAN (layer 2: [16, 128, 64, 64]) and A8 Adam (largest D tensor: 512 x 256 x 4 x 4)
bn = torch.nn.BatchNorm2d(128).cuda()
xb = T.Var(torch.randn(N, 128, 64, 64, generator=g).cuda())
gy = torch.randn(xb.data.shape, generator=g).cuda()
el = xb.data.numel()


def bn_fwd_bwd():
    t = T.Tape()
    y = T.bn_lrelu(t, xb, bn)
    y.grad = gy
    t.backward()


timed("bn_lrelu fwd", lambda: T.bn_lrelu(None, xb, bn))
algo["bn_stats_kernel"] = el * 4                        # x read once for mean / variance
algo["bn_lrelu_apply_kernel"] = el * 8                  # x read, y written
timed("bn_lrelu fwd + bwd", bn_fwd_bwd)
algo["bn_bwd_reduce_kernel"] = el * 12                  # x, y, dy read for the two per-channel sums
algo["bn_bwd_apply_kernel"] = el * 16                   # x, y, dy read, dx written
p = torch.nn.Parameter(torch.randn(512 * 256 * 16, generator=g).cuda())
p.grad = torch.randn(512 * 256 * 16, generator=g).cuda()
opt = T.FusedAdam([p])
timed("adam_multi_kernel", lambda: opt.step())
algo["adam_multi_kernel"] = p.numel() * 28              # p, g, m, v read; p, m, v written

# ---- N1 slice preparation (CCL + bounds + plane build), volume -> u8, uint8 pipeline kernels, input packing (inside a forward)
from healthivert_gan_b200.volume import VolumeSynthesizer
sd = synth.synthetic_generator_state_dict()
gen = hv.Generator({"input_dim": 1, "ngf": 16}, True)
gen.load_state_dict(sd)
gen = gen.cuda().eval()
gen.precision = "bf16"
lab64, ct64, cam64 = synth.synthetic_volume(seed=0, depth=64)
vs = VolumeSynthesizer(gen, batch=64)
timed("vol_to_u8", lambda: vs._to_u8_slices(ct64, 2, 1.0), reps=3)
algo["vol_to_u8_kernel"] = 256 * 256 * 64 * 9           # float64 in, uint8 out
timed("volume synthesize (slice_prepare / finish inside)", lambda: vs.synthesize(ct64, lab64, cam64, 20), reps=2)
algo["slice_build_kernel"] = None
algo["slice_finish_kernel"] = None
algo["ccl_bounds_kernel"] = None
pipe = hv.SlicePipeline(gen, batch=N, depth=1, use_graph=False)
timed("pipeline step (pl_unpack / pl_finish inside)", lambda: (pipe.submit(0), pipe.wait(0)), reps=5)
algo["pl_unpack_kernel"] = px * (2 + 12)                # two u8 planes in, three fp32 planes out
algo["pl_finish_kernel"] = px * (12 + 3)                # three fp32 planes in, three u8 planes out
algo["pack_kx_kernel<3, 5>"] = px * (8 + 32)            # 2 fp32 planes (+ a scalar) in, 16 bf16 channels out
algo["pack_kx_kernel<1, 5>"] = px * (4 + 32)            # the coarse mask plane in, 16 bf16 channels out
algo["pack_kx_kernel<1, 3>"] = None
algo["pack_nbhd4_kernel"] = px * 4 + (px // 4) * 32     # the CAM plane in, its 4x4 neighbourhood per low-res position out
pipe.close()

# ---- N4 spine straightening: trilinear gather of a float64 256^3 volume on 300 curve planes of 128 x 128 samples (hv_resample_curve)
from healthivert_gan_b200 import straighten as stn
vol = torch.rand(256, 256, 256, generator=g, dtype=torch.float64).cuda()
curve = np.stack([np.linspace(20, 235, 40), 128 + 20 * np.sin(np.linspace(0, 3, 40)), 128 + 10 * np.cos(np.linspace(0, 2, 40))], axis=1)
inter = stn.Interpolator(curve, step=1, get_local_basis=stn.get_local_basis)
npts = inter.knots.shape[0]
timed("resample_curve (trilinear)", lambda: inter.interpolate_along(vol, (128, 128), order=1, return_device=True), reps=5)
algo["resample_curve_kernel"] = npts * 128 * 128 * (8 * 8 + 8)      # eight float64 corner reads + one float64 write per sample (upper bound: neighbours share corners)
torch.cuda.synchronize()
print(json.dumps({"algorithmic_bytes": algo, "warm_us_cuda_events": warm}))
