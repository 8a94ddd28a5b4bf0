"""Configs 3 / 5 of BASELINE.json: full synthetic straightened volumes through the batched three-stage synthesis in both
orientations (sagittal = the reference driver, coronal = its twin), then the RHLV features of both orientations (3 + 3
features per vertebra = the 2.5D input of SVM_grading_2.5d.py).  Volumes are sharded round-robin over the ranks, no collective.

usage: python tools/bench_volume.py [--volumes 1] [--depth 256] [--precision bf16] [--cpu-slices 0]
       torchrun --nproc-per-node N tools/bench_volume.py --volumes 64 --depth 64
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import healthivert_gan_b200 as hv
from healthivert_gan_b200 import _lib, mask_ops, sharding
from healthivert_gan_b200.volume import VolumeSynthesizer
from oracle import synth


def rhlv_features(label, fake, vid, axis):
    lab = (label == vid).astype(np.float64)
    fk = (fake == vid).astype(np.float64)
    loc = np.where(lab)[axis]
    c, ln = int(np.mean(loc)), int((loc.max() - loc.min()) // 5)
    return mask_ops.calculate_rhlv(fk, lab, c, ln, None, 0.7, axis=axis)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--volumes", type=int, default=1)
    ap.add_argument("--depth", type=int, default=256, help="slices per volume along axis 2 (reference data: 64)")
    ap.add_argument("--precision", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    rank, world, local = sharding.world_from_env()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = hv.Generator({"input_dim": 1, "ngf": 16}, True)
    g.load_state_dict(synth.synthetic_generator_state_dict())
    g = g.cuda().eval()
    g.precision = args.precision
    vs = VolumeSynthesizer(g, batch=args.batch)
    mine = sharding.shard_round_robin(args.volumes, rank, world)
    vols = {v: synth.synthetic_volume(seed=v, depth=args.depth) for v in mine}
    # warm-up on a volume of the same shape (plan creation, kernel attributes, caching-allocator pools)
    lab, ct, cam = synth.synthetic_volume(seed=999, depth=args.depth)
    coronal_synth = args.depth == 256   # the generator is a 256x256 network: coronal planes are 256 x depth
    vs.synthesize(ct, lab, cam, 20)
    if coronal_synth:
        vs.synthesize(ct, lab, cam, 20, axis=1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.launch_count()
    slices = 0
    feats = {}
    dt = t_rhlv = 0.0
    for v in mine:
        label, ct, cam = vols[v]
        t0 = time.perf_counter()
        lab_dev = torch.as_tensor(label).cuda() if coronal_synth else label   # two orientations: the label volume crosses PCIe once
        ct_s, lab_s = vs.synthesize(ct, lab_dev, cam, 20, axis=2)
        lab_c = vs.synthesize(ct, lab_dev, cam, 20, axis=1)[1] if coronal_synth else lab_s
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        slices += int(lab_s.any(axis=(0, 1)).sum()) + (int(lab_c.any(axis=(0, 2)).sum()) if coronal_synth else 0)
        # 2.5D features: sagittal + coronal RHLV (the reference reads the same synthesized volume in both scripts)
        feats[v] = list(rhlv_features(label, lab_s, 20, 2)[:3]) + list(rhlv_features(label, lab_c, 20, 1)[:3])
        dt += t1 - t0
        t_rhlv += time.perf_counter() - t1
    stat = torch.tensor([dt, float(slices)], device="cuda", dtype=torch.float64)
    if world > 1:
        tmax = stat[:1].clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tot = stat[1:].clone()
        dist.all_reduce(tot)
        dt, slices = float(tmax), int(tot)
    if rank == 0:
        first = mine[0]
        print(json.dumps({"workload": "two-stage synthesis of full straightened volumes, sagittal + coronal, + RHLV (BASELINE.json configs[2]/[4])",
                          "volumes": args.volumes, "depth": args.depth, "n_gpus": world, "precision": args.precision,
                          "seconds": dt, "volumes_per_s": args.volumes / dt, "output_slices": slices, "output_slices_per_s": slices / dt,
                          "launches": _lib.launch_count() - l0, "rhlv_tail_seconds_rank0": t_rhlv, "coronal_synthesis": coronal_synth,
                          "timing": "device-synchronised wall clock around the synthesize() calls incl. H2D of the float64 volumes and D2H of the results; max over ranks",
                          "rhlv_2p5d_features_volume0": [round(float(x), 6) for x in feats[first]]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
