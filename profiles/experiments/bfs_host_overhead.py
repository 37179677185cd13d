"""BFS wall clock vs kernel time: where does the host side lose time?  Runs the same batched
search several times per visited-table size and prints wall time, cudaMalloc counts and the
allocator's reserved bytes.

    python profiles/experiments/bfs_host_overhead.py [puzzles]
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import tiler_slider_b200 as ts  # noqa: E402
from tiler_slider_b200.bfs import BfsSolver  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda", 0)
table = ts.BatchedTilerSliderEnv.synthetic(P, 6, 4, 8, True, seed=1004, device=dev)
out = []
for log2 in (29, 29, 29, 33, 33, 29):
    solver = BfsSolver(table, table_capacity=1 << log2)
    s0 = torch.cuda.memory_stats()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = solver.solve()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    s1 = torch.cuda.memory_stats()
    out.append({"table_log2": log2, "seconds": dt, "generated_per_s": res.generated / dt,
                "cudaMalloc_calls": s1["num_device_alloc"] - s0["num_device_alloc"],
                "cudaFree_calls": s1["num_device_free"] - s0["num_device_free"],
                "alloc_retries": s1["num_alloc_retries"] - s0["num_alloc_retries"],
                "reserved_GB": s1["reserved_bytes.all.current"] / 1e9})
    del solver, res
print(json.dumps(out, indent=1))
