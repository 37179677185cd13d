"""How much would K2w gain from reading ONE wall plane?  Upper bound without writing the kernel:
step the config-4 batch with actions restricted to one axis (LEFT/RIGHT only touch the rows plane,
UP/DOWN only the columns plane), so DRAM serves 24 B of walls per env-step instead of ~47 B; the
instruction stream is unchanged.  Prints us/step for mixed / horizontal-only / vertical-only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import tiler_slider_b200 as ts  # noqa: E402

N = 4_194_304
env = ts.BatchedTilerSliderEnv.synthetic(N, 12, 8, 36, True, seed=1003, max_steps=100, auto_reset=True, track_flags=False)
g = torch.Generator(device="cuda").manual_seed(1)
for name, lo, hi in (("mixed", 0, 4), ("horizontal only (LEFT/RIGHT)", 2, 4), ("vertical only (UP/DOWN)", 0, 2), ("mixed", 0, 4)):
    acts = torch.randint(lo, hi, (8, env.capacity), dtype=torch.uint8, device="cuda", generator=g)
    graph = env.capture_steps(acts)
    for _ in range(20):
        graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        graph.replay()
    b.record()
    b.synchronize()
    us = a.elapsed_time(b) / 400 * 1e3
    print(f"{name:32s} {us:7.2f} us/step  {N / us * 1e6:.3e} env-steps/s  frac(50 B) {50 * N / us / 1e3 / 6534.1:.3f}")
