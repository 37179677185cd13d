"""Run under torchrun with 2+ ranks (tests/test_gpu_bfs_multi.py): shortest move strings from the
hash-partitioned search when the visited set and its parent links are spread over the ranks.
Every rank runs the collective search over all ranks (both exchanges: a search that records parents
takes the all-to-all path either way) and, on its own, a single-rank search of the same batch;
rank 0 compares them, replays every string through the CPU oracle and prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    solo = [dist.new_group([r]) for r in range(world)][rank]          # a group of this rank alone
    import tiler_slider_b200 as ts
    from tiler_slider_b200.bfs import BfsSolver
    from tests.helpers import random_puzzles, reachable_targets
    from oracle import oracle as orc
    report = {"world": world, "cases": []}
    for S, T, W, multi, n in ((5, 2, 4, True, 96), (6, 3, 9, False, 64), (6, 4, 8, True, 32)):
        rng = np.random.default_rng(S * 10 + T)
        blocked, tiles, _ = random_puzzles(rng, n, S, T, W)
        targets = reachable_targets(orc, rng, S, blocked, tiles, multi, 9)
        table = ts.BatchedTilerSliderEnv.from_arrays(S, blocked, tiles, targets, multi)
        want = BfsSolver(table, table_capacity=1 << 21, group=solo).solve(with_paths=True)
        for exchange in ("nccl", "auto"):
            got = BfsSolver(table, table_capacity=1 << 21, exchange=exchange).solve(with_paths=True)
            plain = BfsSolver(table, table_capacity=1 << 21, exchange=exchange).solve()
            if rank:
                continue
            assert got.levels == want.levels == plain.levels and got.generated == want.generated == plain.generated
            assert got.states_per_puzzle.tolist() == want.states_per_puzzle.tolist()
            assert got.solve_depth_per_puzzle.tolist() == want.solve_depth_per_puzzle.tolist()
            solved = 0
            for e in range(n):
                depth, sol = int(got.solve_depth_per_puzzle[e]), got.solutions[e]
                if depth < 0:
                    assert sol is None and want.solutions[e] is None
                    continue
                solved += 1
                assert len(sol) == depth == len(want.solutions[e]), (e, sol, want.solutions[e])
                if depth == 0:
                    continue
                bl = [(c // S, c % S) for c in np.flatnonzero(blocked[e])]
                st = orc.OracleState(S, bl, tiles[e].tolist(), targets[e].tolist(), multi)
                won_at = [k + 1 for k, ch in enumerate(sol) if st.move("UDLR".index(ch))]
                assert won_at[:1] == [depth], (e, sol, won_at)
            report["cases"].append({"shape": [S, T, W, multi, n], "exchange": exchange, "solved": solved,
                                    "states": got.n_states})
    # one big puzzle, 6 tiles: the key has no puzzle id, the state space is spread over the ranks
    rng = np.random.default_rng(66)
    blocked, tiles, _ = random_puzzles(rng, 1, 6, 6, 5)
    targets = reachable_targets(orc, rng, 6, blocked, tiles, True, 60)
    p = ts.Puzzle(6, [(int(c) // 6, int(c) % 6) for c in np.flatnonzero(blocked[0])], [tuple(map(int, t)) for t in tiles[0]],
                  [tuple(map(int, t)) for t in targets[0]], True)
    want = BfsSolver([p], table_capacity=1 << 22, group=solo).solve(with_paths=True)
    got = BfsSolver([p], table_capacity=1 << 22).solve(with_paths=True)
    if rank == 0:
        assert got.levels == want.levels and got.solve_depth == want.solve_depth
        sol = got.solutions[0]
        if want.solve_depth >= 0:
            st = orc.OracleState(6, p.blocked_locations, p.initial_locations, p.target_locations, True)
            won_at = [k + 1 for k, ch in enumerate(sol) if st.move("UDLR".index(ch))]
            assert len(sol) == want.solve_depth and won_at[:1] == [want.solve_depth]
        else:
            assert sol is None
        report["six_tiles"] = {"states": got.n_states, "solve_depth": got.solve_depth, "solution": sol}
        assert want.solve_depth == 20 and got.n_states == 126729        # the CPU oracle BFS of this puzzle
        report["ok"] = True
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
