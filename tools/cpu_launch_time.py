"""Host-side cost of one bf16 forward call (launch loop only, no sync) vs its device time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import healthivert_gan_b200 as hv
from oracle import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = hv.Generator({"input_dim": 1, "ngf": 16}, True); g.load_state_dict(synth.synthetic_generator_state_dict()); g = g.cuda().eval(); g.precision = "bf16"
x, mask, cam, ratio = (t.cuda() for t in synth.synthetic_slices(n, seed=1))
with torch.no_grad():
    for _ in range(5):
        g(x, mask, cam, ratio)
    torch.cuda.synchronize()
    for rep in range(3):
        cpu, e0, e1 = [], torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            t0 = time.perf_counter(); g(x, mask, cam, ratio); cpu.append(time.perf_counter() - t0)
        e1.record(); torch.cuda.synchronize()
        print(f"host call {1e6 * sum(cpu) / 20:.0f} us (min {1e6 * min(cpu):.0f}); device {e0.elapsed_time(e1) / 20 * 1e3:.0f} us per forward (back to back, L2 warm)")
        # one forward at a time: host time when the queue is empty
        cpu = []
        for _ in range(10):
            torch.cuda.synchronize(); t0 = time.perf_counter(); g(x, mask, cam, ratio); cpu.append(time.perf_counter() - t0)
        print(f"host call on an empty queue {1e6 * sum(cpu) / 10:.0f} us")
