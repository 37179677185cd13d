"""Where the time of a batched BFS goes (single GPU): per-phase wall clock with a sync after each."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import tiler_slider_b200 as ts
from tiler_slider_b200 import bfs as B
P = 4096
table = ts.BatchedTilerSliderEnv.synthetic(P, 6, 4, 8, True, seed=1004)
k = B.CudaBfsKernels(table)
acc = {}
def timed(name, fn, *a):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(*a); torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
    return r
for rep in range(2):
    acc.clear()
    tab = timed("new_table", k.new_table, 1 << 29)
    frontier, _ = k.insert(tab, k.seed())
    levels = 0
    while frontier.numel():
        succ = timed("expand", k.expand, frontier & ~B.WON_BIT)
        recv = timed("filter", lambda s: s[s != B.NONE], succ)
        frontier, n_won = timed("insert", k.insert, tab, recv)
        levels += 1
    print(rep, levels, {n: round(v * 1e3, 2) for n, v in acc.items()}, "ms")
