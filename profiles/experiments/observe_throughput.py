"""Throughput of K3 ts_observe (float32 [N,S,S,3] written per call) against the HBM write roofline."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import tiler_slider_b200 as ts
for (S, T, W, multi, N) in [(6, 4, 8, True, 4_194_304), (5, 1, 5, False, 4_194_304), (12, 8, 36, True, 1_048_576)]:
    env = ts.BatchedTilerSliderEnv.synthetic(N, S, T, W, multi, seed=1)
    out = torch.empty(N, S, S, 3, dtype=torch.float32, device="cuda")
    for _ in range(3):
        env.observe(out)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        env.observe(out)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 20
    gb = out.numel() * 4 / 1e9
    print(f"S={S} T={T} N={N}: {ms:.3f} ms per call, {gb / ms * 1e3:.0f} GB/s written, {N / ms * 1e3:.3e} obs/s")
