#!/usr/bin/env python
"""bench.py — slices/sec of the two-stage generator forward (256x256), BASELINE.json's metric.

Workload (config.workload): BASELINE.json configs[1] = batch-16 256x256 two-stage generator
inference on one B200 (weights: oracle.synth random-init, spectral norm converged; inputs:
oracle.synth.synthetic_slices).  One "step" = one Generator.forward over one batch of 16 slices.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|bf16] [--impl reference]

* value      : slices/s with inputs resident in HBM; every step is timed with its own CUDA-event pair
               on the launching stream and L2 is flushed (256 MiB write) between steps.
* e2e        : the same metric through the public API (healthivert_gan_b200.Generator.forward) from
               pinned HOST buffers: H2D of the step's inputs and D2H of its outputs inside the timed region.
* roofline   : the dominant kernel (the 64->64 3x3 conv family: 19 of 47 layers) timed alone, live,
               with CUDA events; achieved = algorithmic FLOPs per launch / mean launch time.
* cpu_baseline : the oracle port of the reference forward on the host cores (rank 0, N=1 only).
* --impl reference : the reference's CPU implementation (oracle port; /root/reference cannot travel to
               the GPU box) on all host threads, same config / metric.
Multi-GPU (torchrun): weak scaling, every rank runs its own batch-16 stream of slices, no collective
on the data path; barrier + synchronize around the timed region, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 16
FLOP_PER_SLICE = 17.54e9          # 14.187 (47 convs) + 1.208 + 2.147 (attention) GFLOP
METRIC = "slices/sec two-stage gen fwd (256x256)"


def _traffic(precision):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel per launch, from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh).get(precision)
    except (OSError, ValueError):
        return None


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (every 100 ms: measured on B200, polling every 20 ms
    takes driver locks often enough to slow the launch thread - device value -5 %, end-to-end -16 %)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", os.environ.get("HV_CLOCK_MS", "100")], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def _cpu_forward_rate(steps, warmup, batch=BATCH):
    """Oracle port of the reference forward on all host threads: slices/s over `steps` batches."""
    import torch
    from oracle import generator_ref as gr
    from oracle import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.synthetic_generator_state_dict()
    x, mask, cam, ratio = synth.synthetic_slices(batch, seed=123)
    with torch.no_grad():
        for _ in range(warmup):
            gr.generator_forward(sd, x, mask, cam, ratio)
        t0 = time.perf_counter()
        for _ in range(steps):
            gr.generator_forward(sd, x, mask, cam, ratio)
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps = max(1, min(args.steps, 40))
    rate, ms, threads = _cpu_forward_rate(steps, min(args.warmup, 3))
    sample = f"{steps} steps x batch {BATCH} (each step = one full batch-{BATCH} forward), oracle port of the reference"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "slices/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"batch-{BATCH} 256x256 two-stage generator inference (BASELINE.json configs[1])",
                   "batch": BATCH, "precision": "fp32", "device": "host CPU"},
        "cpu_baseline": {"value": rate, "unit": "slices/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import healthivert_gan_b200 as hv
    from healthivert_gan_b200 import _lib
    from healthivert_gan_b200.inpaint_networks import conv2d_fused
    from oracle import synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sd = synth.synthetic_generator_state_dict()
    g = hv.Generator({"input_dim": 1, "ngf": 16}, True)
    g.load_state_dict(sd)
    g = g.to(dev).eval()
    g.precision = args.precision
    g.return_flow = True
    host = [t.pin_memory() for t in synth.synthetic_slices(BATCH, seed=123 + rank)]
    x, mask, cam, ratio = (t.to(dev) for t in host)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step():
        with torch.no_grad():
            return g(x, mask, cam, ratio)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    wall0 = time.perf_counter()
    launches0 = _lib.launch_count()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - wall0
    launches = _lib.launch_count() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)

    # ---------------- end to end through the public API from pinned host buffers
    # Every step copies its own inputs host->device and its results device->host; copies run on two extra streams
    # (double-buffered) so that the H2D of step i+1 and the D2H of step i-1 overlap the forward of step i.  The timed
    # region is one CUDA-event pair around the whole loop, closed after the last D2H has landed.
    copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    dev_in = [[torch.empty_like(t, device=dev) for t in host] for _ in range(2)]
    with torch.no_grad():
        probe = g(x, mask, cam, ratio)
    # pinned result buffers are allocated up front (page-locking is a one-off set-up cost, not part of a step)
    outs_host = [[torch.empty(probe[k].shape, dtype=probe[k].dtype).pin_memory() for k in (0, 1, 2, 3, 5, 6)] for _ in range(2)]
    alive = [None, None]
    torch.cuda.synchronize()
    def e2e_loop(nsteps):
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_comp = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        barrier()
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record(stream)
        copy_in.wait_event(t_start)
        for i in range(nsteps):
            b = i % 2
            with torch.cuda.stream(copy_in):
                if i >= 2:
                    copy_in.wait_event(ev_comp[b])      # the forward that read this input buffer has finished
                for d, h in zip(dev_in[b], host):
                    d.copy_(h, non_blocking=True)
                ev_in[b].record(copy_in)
            stream.wait_event(ev_in[b])
            with torch.no_grad():
                out = g(*dev_in[b])
            ev_comp[b].record(stream)
            keep = [out[k] for k in (0, 1, 2, 3, 5, 6)]
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(ev_comp[b])
                if i >= 2:
                    ev_out[b].synchronize()             # the host consumed (here: may overwrite) step i-2's results
                for h, t in zip(outs_host[b], keep):
                    h.copy_(t, non_blocking=True)
                ev_out[b].record(copy_out)
            alive[b] = keep      # the device results stay referenced until their D2H has been waited for (two steps later)
        stream.wait_event(ev_out[0])
        stream.wait_event(ev_out[1])
        t_end.record(stream)
        t_end.synchronize()
        barrier()
        return t_start.elapsed_time(t_end)

    e2e_loop(max(args.warmup, 3))     # warm-up of the pipelined loop itself (the allocator grows by the two result sets kept alive)
    e2e_ms = e2e_loop(args.steps)
    clocks = sampler.stop() if rank == 0 else None
    h2d = sum(t.numel() * t.element_size() for t in host)
    d2h = sum(t.numel() * t.element_size() for t in outs_host[0])

    # ---------------- dominant kernel alone (roofline): 64->64 3x3 conv on the batch's 64x64 maps
    # (19 of 47 layers; coarse conv5 = layer 4).  Timed live with CUDA events on the launch stream.
    reps = 20
    kev = []
    if args.precision == "bf16":
        # layers 4..10 = coarse conv5, conv6, conv7..10_atrous, conv11: seven consecutive 64->64 3x3 layers of the plan, each
        # reading its predecessor's output exactly as in the forward (L2 flushed before the chain, not inside it)
        chain = list(range(4, 11))
        launch_k = lambda: [g.run_layer(l, BATCH) for l in chain]
    else:
        a = torch.randn(BATCH, 64, 64, 64, device=dev)
        w = torch.randn(64, 64, 3, 3, device=dev) * 0.05
        b = torch.zeros(64, device=dev)
        chain = [0]
        launch_k = lambda: conv2d_fused([(a, 0)], w, b, 3, 1, 1, 1, "elu", 64, 64)
    for _ in range(3):
        launch_k()
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launch_k()
        e1.record(stream)
        kev.append((e0, e1))
    torch.cuda.synchronize()
    k_ms = sum(x0.elapsed_time(x1) for x0, x1 in kev) / (reps * len(chain))
    k_flop = 2.0 * BATCH * 64 * 64 * 64 * 576

    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = t.tolist()
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    if rank == 0:
        peaks = _peaks()
        value = world * BATCH * args.steps / (dev_ms * 1e-3)
        e2e = world * BATCH * args.steps / (e2e_ms * 1e-3)
        achieved = k_flop / (k_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "slices/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": f"batch-{BATCH} 256x256 two-stage generator inference (BASELINE.json configs[1])",
                       "batch": BATCH, "precision": args.precision, "l2": "flushed (256 MiB write) between steps",
                       "timing": "per-step CUDA-event pairs on the launch stream, summed; max over ranks",
                       "e2e_timing": "one CUDA-event pair around all steps; per-step H2D/D2H on copy streams overlap the forward",
                       "parallelism": f"slice-sharded x{world}, no collective"},
            "e2e": {"value": e2e, "unit": "slices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "conv 64->64 3x3 (fp32 SIMT parity kernel)"
                         if args.precision == "fp32" else "conv 64->64 3x3 (tcgen05 bf16)",
                         "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_tflops"], "traffic": _traffic(args.precision),
                         "algorithmic_flop_per_launch": k_flop, "launch_us": k_ms * 1e3,
                         "launches_timed": f"{reps} x {len(chain)} launches (CUDA events around each chain of consecutive 64->64 layers, L2 flushed before each chain)",
                         "peak_source": peaks["source"] + " burst bf16 (kernel timed alone)",
                         "whole_forward_tensor_frac_sustained": value / world * FLOP_PER_SLICE / (peaks["bf16_tflops_sustained"] * 1e12)},
            "wall_s_timed_region": wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, ms, threads = _cpu_forward_rate(10, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "slices/s", "cores": threads, "kind": "port",
                                    "sample": f"10 steps x batch {BATCH} of the same workload ({ms:.0f} ms/step), oracle port of the reference forward"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("HV_PRECISION", "bf16"), choices=["fp32", "bf16"],
                    help="bf16 = tcgen05 tensor-core mode (headline); fp32 = SIMT parity mode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
