"""Host-driven (one read-back per level) against device-driven (frontier sizes stay on the GPU,
16 levels per read-back, ts_bfs_levels) single-GPU BFS, same process, alternating.

    python profiles/experiments/bfs_device_driven.py
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import tiler_slider_b200 as ts  # noqa: E402
from tiler_slider_b200.bfs import BfsSolver  # noqa: E402

dev = torch.device("cuda", 0)
out = []
for P, log2, tiles in ((1, 20, 4), (1, 22, 6), (1024, 24, 4), (65536, 30, 4)):
    table = ts.BatchedTilerSliderEnv.synthetic(P, 6, tiles, 8, True, seed=1004, device=dev)
    solver = BfsSolver(table, table_capacity=1 << log2)
    row = {"puzzles": P, "tiles": tiles, "table_log2": log2, "host_driven_s": [], "device_driven_s": []}
    for rep in range(4):
        for mode in (False, True):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = solver.solve(device_driven=mode)
            torch.cuda.synchronize()
            row["device_driven_s" if mode else "host_driven_s"].append(round(time.perf_counter() - t0, 5))
    row["unique_states"], row["levels"] = res.n_states, len(res.levels)
    out.append(row)
print(json.dumps(out, indent=1))
