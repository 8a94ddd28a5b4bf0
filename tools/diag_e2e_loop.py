"""Variants of bench.py's end-to-end loop to find what serialises it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import healthivert_gan_b200 as hv
from oracle import synth

dev = torch.device("cuda", 0)
n, steps = 16, 40
g = hv.Generator({"input_dim": 1, "ngf": 16}, True); g.load_state_dict(synth.synthetic_generator_state_dict()); g = g.cuda().eval(); g.precision = "bf16"
g.return_flow = True
host = [t.pin_memory() for t in synth.synthetic_slices(n, seed=1)]
x, mask, cam, ratio = (t.to(dev) for t in host)
stream = torch.cuda.current_stream()
with torch.no_grad():
    for _ in range(3): probe = g(x, mask, cam, ratio)
copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
dev_in = [[torch.empty_like(t, device=dev) for t in host] for _ in range(2)]
outs_host = [[torch.empty(probe[k].shape, dtype=probe[k].dtype).pin_memory() for k in (0, 1, 2, 3, 5, 6)] for _ in range(2)]

def run(do_h2d, do_d2h, host_sync, prealloc_out=False):
    alive = [None, None]
    ev_in = [torch.cuda.Event() for _ in range(2)]; ev_comp = [torch.cuda.Event() for _ in range(2)]; ev_out = [torch.cuda.Event() for _ in range(2)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); copy_in.wait_event(e0)
    t0 = time.perf_counter(); cpu_fwd = 0.0
    for i in range(steps):
        b = i % 2
        if do_h2d:
            with torch.cuda.stream(copy_in):
                if i >= 2: copy_in.wait_event(ev_comp[b])
                for d, h in zip(dev_in[b], host): d.copy_(h, non_blocking=True)
                ev_in[b].record(copy_in)
            stream.wait_event(ev_in[b])
        tf = time.perf_counter()
        with torch.no_grad(): out = g(*dev_in[b]) if do_h2d else g(x, mask, cam, ratio)
        cpu_fwd += time.perf_counter() - tf
        ev_comp[b].record(stream)
        keep = [out[k] for k in (0, 1, 2, 3, 5, 6)]
        if do_d2h:
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(ev_comp[b])
                if i >= 2 and host_sync: ev_out[b].synchronize()
                for h, t in zip(outs_host[b], keep): h.copy_(t, non_blocking=True)
                ev_out[b].record(copy_out)
        alive[b] = keep
    cpu = time.perf_counter() - t0
    if do_d2h: stream.wait_event(ev_out[0]); stream.wait_event(ev_out[1])
    e1.record(stream); e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"h2d={do_h2d} d2h={do_d2h} host_sync={host_sync}: {ms:.3f} ms/step ({n / ms * 1e3:.0f} slices/s); host loop {cpu / steps * 1e3:.3f} ms/step, of which forward call {cpu_fwd / steps * 1e3:.3f}")

for cfg in [(False, False, False), (True, False, False), (False, True, False), (False, True, True), (True, True, False), (True, True, True)]:
    run(*cfg); run(*cfg)
