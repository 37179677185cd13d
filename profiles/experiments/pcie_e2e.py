"""What bounds `e2e`: raw pinned-memory PCIe rates of the box beside ts_step_host at several
chunk sizes / stream counts (config 3, 16.7M envs).  Prints one JSON object.

    python profiles/experiments/pcie_e2e.py
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import tiler_slider_b200 as ts  # noqa: E402

dev = torch.device("cuda", 0)
out = {}


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for mb in (16, 64, 256):
    n = mb << 20
    h_a, h_b = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a, d_b = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h2d = timed(lambda: d_a.copy_(h_a, non_blocking=True))
    d2h = timed(lambda: h_b.copy_(d_b, non_blocking=True))

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)
    bi = timed(both)
    out[f"pcie_{mb}MiB"] = {"h2d_GBs": n / h2d / 1e9, "d2h_GBs": n / d2h / 1e9, "bidir_each_GBs": n / bi / 1e9}

N = 1 << 24
env = ts.BatchedTilerSliderEnv.synthetic(N, 6, 4, 4, True, seed=1, max_steps=64, auto_reset=True, device=dev)
h_act = torch.randint(0, 4, (N,), dtype=torch.uint8).pin_memory()
h_rew = torch.empty(N, dtype=torch.float32).pin_memory()
h_done = torch.empty(N, dtype=torch.uint8).pin_memory()
h_flags = torch.empty(N, dtype=torch.uint8).pin_memory()
res = {}
for n_streams in (2, 4, 8):
    for chunk in (1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 24):
        env._host_ctx and env._lib.ts_host_ctx_destroy(env._host_ctx)
        env._host_ctx = None
        full = timed(lambda: env.step_host(h_act, h_rew, h_done, chunk_envs=chunk, n_streams=n_streams), reps=8)
        comp = timed(lambda: env.step_host(h_act, h_flags=h_flags, chunk_envs=chunk, n_streams=n_streams), reps=8)
        res[f"streams{n_streams}_chunk{chunk}"] = {"full_env_steps_per_s": N / full, "full_d2h_GBs": 5 * N / full / 1e9,
                                                   "compact_env_steps_per_s": N / comp, "compact_each_way_GBs": N / comp / 1e9}
out["step_host"] = res
print(json.dumps(out, indent=1))
