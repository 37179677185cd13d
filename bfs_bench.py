"""bfs_bench.py -- BASELINE config 5: batched breadth-first search with hash-partitioned dedup.

    python bfs_bench.py [--puzzles P] [--check C] [--mode hash|local]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bfs_bench.py --puzzles P

P synthetic 6x6 puzzles with 4 coloured tiles and 8 walls (K0, the create_simple_env recipe) are
searched to exhaustion at once.  --mode hash: key = puzzle id || positions, one visited table in
HBM; with N ranks every rank owns the keys that hash to it and the successors travel to their
owners once per depth (NCCL all-to-all, or the expand kernel writing into peer inboxes over NVLink).
--mode local: one CTA per puzzle with the visited set in shared memory (K6); with N ranks the
puzzles are sharded by index and nothing is exchanged (tiler_slider_b200/bfs.py).  Work unit: one
generated successor (state x move), SURVEY 8(d).
The first C puzzles are cross-checked against the CPU oracle's BFS (state count, solve depth).
Prints one JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--puzzles", type=int, default=1024)
    ap.add_argument("--size", type=int, default=6)
    ap.add_argument("--tiles", type=int, default=4)
    ap.add_argument("--walls", type=int, default=8)
    ap.add_argument("--seed", type=int, default=1004)
    ap.add_argument("--table-log2", type=int, default=0, help="log2 of the per-rank visited-table capacity (0 = auto)")
    ap.add_argument("--check", type=int, default=8, help="puzzles cross-checked against the CPU oracle on rank 0")
    ap.add_argument("--exchange", choices=["auto", "nccl", "p2p"], default="auto",
                    help="multi-GPU: NCCL all-to-all, or the expand kernel writing into peer inboxes over NVLink")
    ap.add_argument("--mode", choices=["hash", "local"], default="hash",
                    help="hash: one visited table in HBM, hash-partitioned over the ranks; local: one CTA per puzzle, on chip (K6)")
    ap.add_argument("--profile", action="store_true", help="a third, phase-synchronised search: wall time per phase (rank 0)")
    args = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist
    import tiler_slider_b200 as ts
    from tiler_slider_b200.bfs import BfsSolver, LocalBfs

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    table = ts.BatchedTilerSliderEnv.synthetic(args.puzzles, args.size, args.tiles, args.walls, True, seed=args.seed, device=dev)
    log2 = args.table_log2 or max(16, int(np.ceil(np.log2(args.puzzles * 16384 / world))))   # ~3,400 states per puzzle on average
    solver = BfsSolver(table, table_capacity=1 << log2, exchange=args.exchange) if args.mode == "hash" else LocalBfs(table)
    if world > 1:   # create the NCCL communicator and its all-to-all channels outside the timed region
        w = torch.zeros(world, dtype=torch.int64, device=dev)
        dist.all_to_all_single(torch.empty_like(w), w)
        dist.all_reduce(w)
    BfsSolver(ts.BatchedTilerSliderEnv.synthetic(min(128, args.puzzles), args.size, args.tiles, args.walls, True, seed=1, device=dev),
              table_capacity=1 << 22).solve(max_depth=3)      # warm the kernels / allocator
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times = []
    for _ in range(2):      # the first search also pays the cudaMallocs of table and workspace (reported as cold)
        t0 = time.perf_counter()
        res = solver.solve()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times.append(time.perf_counter() - t0)
    dt_cold, dt = times
    phases = None
    if args.profile and args.mode == "hash":
        solver.profile = True
        res_p = solver.solve()          # host-driven levels (a read-back per level): must find the same search
        phases = {k: round(v, 4) for k, v in solver.phase_seconds.items()}
        phases["same_search"] = bool(res_p.levels == res.levels and res_p.generated == res.generated and
                                     torch.equal(res_p.states_per_puzzle, res.states_per_puzzle))

    ok = phases is None or phases["same_search"]
    if rank == 0 and args.check:
        from oracle import oracle as orc
        S = args.size
        blocked = table.blocked_cells().cpu().numpy()
        tiles = table.positions().cpu().numpy()
        targets = table.target_positions().cpu().numpy()
        for e in range(min(args.check, args.puzzles)):
            b = [(c // S, c % S) for c in np.flatnonzero(blocked[e])]
            n, _, depth, _ = orc.OracleState(S, b, tiles[e].tolist(), targets[e].tolist(), True).bfs(max_states=1 << 22)
            ok &= n == int(res.states_per_puzzle[e]) and depth == int(res.solve_depth_per_puzzle[e])
    if rank == 0:
        solved = int((res.solve_depth_per_puzzle >= 0).sum())
        print(json.dumps({"config": f"BFS {args.puzzles} puzzles {args.size}x{args.size}/{args.tiles} tiles/{args.walls} walls, "
                                    f"{world} GPU(s), " + (f"table 2^{log2} per rank, exchange {solver.exchange if world > 1 else 'none'}"
                                                           if args.mode == "hash" else f"one CTA per puzzle on chip, plan {solver.plan()}"),
                          "mode": args.mode,
                          "n_gpus": world, "unique_states": res.n_states, "generated_successors": res.generated,
                          "depth": len(res.levels) - 1, "seconds": dt, "seconds_cold": dt_cold,
                          "generated_successors_per_s": res.generated / dt, "unique_states_per_s": res.n_states / dt,
                          "puzzles_solved": solved, "max_solve_depth": int(res.solve_depth_per_puzzle.max()),
                          "oracle_check": {"puzzles": min(args.check, args.puzzles), "ok": bool(ok)},
                          **({"phase_seconds_synchronised": phases} if phases else {})}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
