"""Launch each non-headline kernel a few times at a representative size, so that one `ncu -k
regex:<kernel> -s <skip> -c 1` per kernel can capture it (profiles/README.md lists the commands).
    python profiles/experiments/r2_targets.py observe|valid|goal|bfs_local|bfs_hash
Also prints CUDA-event timings of the same launches (the numbers quoted beside the ncu pages)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import tiler_slider_b200 as ts  # noqa: E402
from tiler_slider_b200.bfs import BfsSolver, LocalBfs  # noqa: E402


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / n


def main():
    what = sys.argv[1]
    if what in ("observe", "valid", "goal"):
        for S, T, W, multi, N in ((6, 4, 8, True, 4_194_304), (12, 8, 36, True, 1_048_576)):
            env = ts.BatchedTilerSliderEnv.synthetic(N, S, T, W, multi, seed=1)
            if what == "observe":
                out = torch.empty(N, S, S, 3, dtype=torch.float32, device="cuda")
                ms = timed(lambda: env.observe(out))
                print(f"observe {S}x{S}/{T}: {ms * 1e3:.1f} us, {N * S * S * 12 / ms / 1e6:.0f} GB/s written")
            elif what == "valid":
                ms = timed(env.valid_moves)
                print(f"valid_moves {S}x{S}/{T}: {ms * 1e3:.1f} us, {N / ms / 1e6:.2f} G envs/s")
            else:
                ms = timed(env.goal_check)
                print(f"goal_check {S}x{S}/{T}: {ms * 1e3:.1f} us, {N / ms / 1e6:.2f} G envs/s")
    elif what == "bfs_local":
        table = ts.BatchedTilerSliderEnv.synthetic(65_536, 6, 4, 8, True, seed=1004)
        loc = LocalBfs(table)
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = loc.solve()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"bfs_local 65536 puzzles: {dt * 1e3:.2f} ms, {r.generated / dt:.3e} successors/s, plan {loc.plan()}")
    elif what == "bfs_hash":
        table = ts.BatchedTilerSliderEnv.synthetic(16_384, 6, 4, 8, True, seed=1004)
        s = BfsSolver(table, table_capacity=1 << 28)
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = s.solve(device_driven=False)        # one launch per level and kernel: the launch index selects the level
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"bfs_hash 16384 puzzles: {dt * 1e3:.2f} ms, {r.generated / dt:.3e} successors/s, levels {r.levels[:24]}")


if __name__ == "__main__":
    main()
