"""BFS driver logic (hash-partitioned frontier exchange, termination, per-puzzle statistics) on
CPU over gloo with world_size 2.  The CUDA kernels are replaced by a stand-in built on the
CPU oracle (test infrastructure); the product driver `tiler_slider_b200.bfs.BfsSolver` is the
code under test.  Known answers: tests/golden/misc.json (plain BFS over the reference's move)."""
import json
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBfsKernels:
    """Same interface as CudaBfsKernels, on CPU tensors, successor function = oracle move."""

    def __init__(self, puzzles):
        from oracle import oracle as orc
        self.device = torch.device("cpu")
        self.puzzles = puzzles
        self.states = [orc.OracleState(p["size"], p["blocked"], p["tiles"], p["targets"], p["multi_color"]) for p in puzzles]

    def _key(self, pid, st, won=False):
        S = st.size
        cells = [r * S + c for r, c in st.current_locations]
        if not st.multi_color:
            cells = sorted(cells)
        k = sum(c << (8 * i) for i, c in enumerate(cells)) | (pid << 32)
        return k - (1 << 63) if won else k

    def seed(self):
        return torch.tensor([self._key(i, st) for i, st in enumerate(self.states)], dtype=torch.int64)

    def expand(self, frontier):
        out = []
        for key in frontier.tolist():
            key &= 0x7FFFFFFFFFFFFFFF                       # the goal bit of an input key is ignored
            pid, st = (key >> 32) & 0x7FFFFFFF, None
            st = self.states[pid]
            cells = [(key >> (8 * i)) & 0xFF for i in range(st.n_tiles)]
            for d in range(4):
                st.set_locations([(c // st.size, c % st.size) for c in cells])
                won = st.move(d)
                k = self._key(pid, st, won)
                out.append(-1 if (k & ~(-(1 << 63))) == key and not won else k)
        return torch.tensor(out, dtype=torch.int64)

    @staticmethod
    def _owner(k, n_ranks):
        return ((k & 0x7FFFFFFFFFFFFFFF) * 2654435761 >> 7) % n_ranks

    def partition(self, keys, n_ranks, parents=None):
        """(bucketed keys, sizes, bucketed parents or None): parents = the frontier `keys` was expanded from"""
        items = [(k, i) for i, k in enumerate(keys.tolist()) if k != -1]
        owner = [self._owner(k, n_ranks) for k, _ in items]
        order = sorted(range(len(items)), key=lambda j: owner[j])
        par = None
        if parents is not None:
            pl = parents.tolist()
            par = torch.tensor([pl[items[j][1] // 4] & 0x7FFFFFFFFFFFFFFF for j in order], dtype=torch.int64)
        return torch.tensor([items[j][0] for j in order], dtype=torch.int64), [owner.count(r) for r in range(n_ranks)], par

    def new_table(self, capacity):
        return {}

    def insert(self, table, keys, parents=None, parent_table=None, stats=None, parent_per_item=False):
        """table: key -> True; parent_table (a second dict): key -> parent key (-1 for roots)"""
        new, n_won = [], 0
        pl = None if parents is None else parents.tolist()
        for i, raw in enumerate(keys.tolist()):
            if raw == -1:
                continue
            n_won += raw < 0
            k = raw & 0x7FFFFFFFFFFFFFFF
            if k not in table:
                table[k] = True
                if parent_table is not None:
                    parent_table[k] = -1 if pl is None else pl[i if parent_per_item else i // 4] & 0x7FFFFFFFFFFFFFFF
                new.append(raw)
        return torch.tensor(new, dtype=torch.int64), n_won

    def _move_between(self, parent, key):
        succ = self.expand(torch.tensor([parent], dtype=torch.int64)).tolist()
        for d in range(4):
            if succ[d] != -1 and succ[d] & 0x7FFFFFFFFFFFFFFF == key:
                return d
        return -1

    def traceback(self, table, parent_table, goals, max_moves):
        """goals = int64[2, n]: goal successor, the state it was generated from"""
        n = goals.shape[1]
        moves, lengths = torch.zeros(n, max_moves, dtype=torch.uint8), torch.full((n,), -1, dtype=torch.int32)
        for i, (g, frm) in enumerate(zip(*goals.tolist())):
            if g == -1:
                continue
            k, path = frm & 0x7FFFFFFFFFFFFFFF, [self._move_between(frm & 0x7FFFFFFFFFFFFFFF, g & 0x7FFFFFFFFFFFFFFF)]
            while parent_table[k] != -1:
                path.append(self._move_between(parent_table[k], k))
                k = parent_table[k]
            lengths[i] = len(path)
            moves[i, : len(path)] = torch.tensor(path[::-1], dtype=torch.uint8)
        return moves, lengths

    def trace_step(self, table, parent_table, keys, rank, n_ranks, given_parents=None):
        """the contract of ts_bfs_trace_step (include/tiler_slider.h)"""
        out = torch.zeros(2, keys.numel(), dtype=torch.int64)
        for i, raw in enumerate(keys.tolist()):
            k = raw & 0x7FFFFFFFFFFFFFFF
            if raw != -1 and given_parents is not None:
                out[0, i] = int(given_parents[i]) & 0x7FFFFFFFFFFFFFFF
                out[1, i] = self._move_between(int(out[0, i]), k)
            elif raw == -1 or (n_ranks > 1 and self._owner(k, n_ranks) != rank):
                out[0, i] = -(1 << 63)
            elif k not in parent_table:
                out[0, i] = -2
            else:
                out[0, i] = parent_table[k]
                if parent_table[k] != -1:
                    out[1, i] = self._move_between(parent_table[k], k)
        return out


def _worker(rank, world, port, puzzles, q, with_paths=False):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tiler_slider_b200.bfs import BfsSolver
    res = BfsSolver(kernels=OracleBfsKernels(puzzles), n_puzzles=len(puzzles), table_capacity=1 << 16).solve(with_paths=with_paths)
    if rank == 0:
        q.put((res.n_states, res.levels, res.solve_depth, res.states_per_puzzle.tolist(),
               res.solve_depth_per_puzzle.tolist(), res.generated) + ((res.solutions,) if with_paths else ()))
    dist.destroy_process_group()


def _golden_bfs():
    with open(os.path.join(ROOT, "tests", "golden", "misc.json")) as f:
        return json.load(f)["bfs"]


@pytest.mark.parametrize("world", [1, 2])
def test_bfs_driver_over_gloo(world):
    gold = {b["name"]: b for b in _golden_bfs()}
    batch = [gold["puzzle_multi_111"], gold["puzzle_multi_180"]]     # same shape: 6x6, 3 ordered tiles
    if world == 1:
        sys.path.insert(0, ROOT)
        from tiler_slider_b200.bfs import BfsSolver
        res = BfsSolver(kernels=OracleBfsKernels(batch), n_puzzles=2, table_capacity=1 << 16).solve()
        got = (res.n_states, res.levels, res.solve_depth, res.states_per_puzzle.tolist(),
               res.solve_depth_per_puzzle.tolist(), res.generated)
    else:
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, world, 29533, batch, q)) for r in range(world)]
        for p in procs:
            p.start()
        got = q.get(timeout=120)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    n_states, levels, solve_depth, spp, dpp, generated = got
    assert spp == [558, 950] and dpp == [8, 13]
    assert n_states == 558 + 950 and solve_depth == 8
    la, lb = batch[0]["levels"], batch[1]["levels"]
    want = [(la[i] if i < len(la) else 0) + (lb[i] if i < len(lb) else 0) for i in range(max(len(la), len(lb)))]
    assert levels == want
    assert generated == 4 * n_states


def _replay(puzzle, solution):
    """does the move string take the puzzle from its initial state to the goal (oracle move)?"""
    from oracle import oracle as orc
    st = orc.OracleState(puzzle["size"], puzzle["blocked"], puzzle["tiles"], puzzle["targets"], puzzle["multi_color"])
    won = False
    for ch in solution:
        won = st.move("UDLR".index(ch))
    return won


def test_bfs_shortest_paths_over_gloo():
    """with_paths on two ranks: the parent of every key travels with it to the key's owner, and the
    chains are walked one collective per move, each rank answering for the keys it owns.  The
    strings must have the known optimal lengths and solve the puzzles when replayed; a one-rank
    run of the same driver gives strings of the same lengths."""
    gold = {b["name"]: b for b in _golden_bfs()}
    # the fourth puzzle starts on its goal: it is won by its first step (the reference evaluates the
    # goal inside step(), environment.py:133), so its shortest solution is one move, not none
    on_goal = dict(size=6, blocked=[[3, 3]], tiles=[[0, 0], [0, 2], [0, 4]], targets=[[0, 0], [0, 2], [0, 4]], multi_color=True)
    batch = [gold["puzzle_multi_111"], gold["puzzle_multi_180"], gold["puzzle_multi_111"], on_goal]
    sys.path.insert(0, ROOT)
    from tiler_slider_b200.bfs import BfsSolver
    one = BfsSolver(kernels=OracleBfsKernels(batch), n_puzzles=4, table_capacity=1 << 16).solve(with_paths=True)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29539, batch, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[3] == one.states_per_puzzle.tolist() and got[3][:3] == [558, 950, 558]
    assert got[4] == one.solve_depth_per_puzzle.tolist() == [8, 13, 8, 1]          # UP moves nothing and wins
    for sols in (one.solutions, got[6]):
        assert [len(s) for s in sols] == got[4]
        assert all(_replay(p, s) for p, s in zip(batch, sols))


def test_bfs_driver_argument_checks():
    """Exchange / device-driven options that need the CUDA kernels are refused, not ignored, when
    the driver runs on stand-in kernels."""
    sys.path.insert(0, ROOT)
    from tiler_slider_b200.bfs import BfsSolver
    gold = {b["name"]: b for b in _golden_bfs()}
    kernels = OracleBfsKernels([gold["puzzle_multi_111"]])
    with pytest.raises(ValueError, match="exchange"):
        BfsSolver(kernels=kernels, n_puzzles=1, exchange="carrier pigeon")
    with pytest.raises(ValueError, match="power of two"):
        BfsSolver(kernels=kernels, n_puzzles=1, table_capacity=1000)
    solver = BfsSolver(kernels=kernels, n_puzzles=1, table_capacity=1 << 12, exchange="p2p")   # one rank: nothing to exchange
    assert solver.exchange == "nccl"
    with pytest.raises(ValueError, match="device-driven"):
        solver.solve(device_driven=True)
    with pytest.raises(ValueError, match="with_paths"):
        solver.solve(with_paths=True, per_puzzle=False)
    assert solver.solve(max_depth=2).levels == gold["puzzle_multi_111"]["levels"][:3]


def test_shard_and_reduce_over_gloo():
    """The step path's only multi-rank logic: disjoint contiguous shards + a MAX reduce of the
    per-rank time, as bench.py does it."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, 29534, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=60) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1:3] == (0, 501) and out[1][1:3] == (501, 1001)
    assert out[0][3] == out[1][3] == 2.5


def _shard_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tiler_slider_b200.batch_env import shard_range
    lo, hi = shard_range(1001, rank, world)
    t = torch.tensor([1.5 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, lo, hi, float(t.item())))
    dist.destroy_process_group()


# ---- LocalBfs (K6 driver): puzzles sharded over the ranks, no exchange, results merged once ------------
class OracleLocalKernels:
    """Same interface as CudaLocalKernels (plan / search / puzzle), on CPU tensors: a plain BFS over
    the oracle's move per puzzle.  `too_big` puzzles report status 1 like a puzzle whose state space
    does not fit the shared-memory bitmap."""
    N_LEVELS, MAX_MOVES = 256, 255

    def __init__(self, puzzles, too_big=()):
        from oracle import oracle as orc
        self.device, self.puzzles, self.too_big, self.orc = torch.device("cpu"), puzzles, set(too_big), orc

    def plan(self, lo, hi):
        return {"ctas_per_sm": 1, "queue_smem": 64}

    def search(self, lo, hi, max_depth, out):
        for pid in range(lo, hi):
            if pid in self.too_big:
                out["status"][pid] = 1
                continue
            p = self.puzzles[pid]
            st = self.orc.OracleState(p["size"], p["blocked"], p["tiles"], p["targets"], p["multi_color"])
            S = st.size

            def key(locs):
                cells = [r * S + c for r, c in locs]
                return tuple(cells if st.multi_color else sorted(cells))
            start = key(st.current_locations)
            seen, frontier, depth, solve, goal = {start: None}, [start], 0, -1, None
            out["levels"][0] += 1
            while frontier and depth < max_depth:
                nxt = []
                for k in frontier:
                    for d in range(4):
                        st.set_locations([(c // S, c % S) for c in k])
                        won = st.move(d)
                        k2 = key(st.current_locations)
                        if won and solve < 0:
                            solve, goal = depth + 1, (k, d)
                        if k2 not in seen:
                            seen[k2] = (k, d)
                            nxt.append(k2)
                out["counters"][1] += 4 * len(frontier)
                depth += 1
                if nxt:
                    out["levels"][depth] += len(nxt)
                frontier = nxt
            out["states"][pid], out["depth"][pid] = len(seen), solve
            if out.get("lengths") is not None:
                moves = []
                link = goal
                while link is not None:
                    moves.append(link[1])
                    link = seen[link[0]]
                out["lengths"][pid] = len(moves) if goal is not None else -1
                for i, m in enumerate(reversed(moves)):
                    out["moves"][pid, i] = m


def _local_worker(rank, world, port, puzzles, too_big, kw, q):
    sys.path.insert(0, ROOT)
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from tiler_slider_b200.bfs import BfsSolver, LocalBfs

    def fallback(rest, max_depth, with_paths):          # the hash-partitioned driver on stand-in kernels, collective over the same ranks
        return BfsSolver(kernels=OracleBfsKernels([puzzles[i] for i in rest]), n_puzzles=len(rest), table_capacity=1 << 16).solve(max_depth=max_depth)
    res = LocalBfs(kernels=OracleLocalKernels(puzzles, too_big), n_puzzles=len(puzzles), fallback=fallback).solve(**kw)
    if rank == 0:
        q.put((res.n_states, res.levels, res.solve_depth, res.states_per_puzzle.tolist(), res.solve_depth_per_puzzle.tolist(),
               res.generated, res.solutions, res.fallback_puzzles))
    if world > 1:
        dist.destroy_process_group()


def _run_local(world, puzzles, too_big=(), port=29535, **kw):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_local_worker, args=(r, world, port, puzzles, tuple(too_big), kw, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


def test_local_bfs_driver_shards_puzzles_over_gloo():
    """K6's driver on two gloo ranks: five puzzles sharded 3 + 2, nothing exchanged during the
    search, per-puzzle results / level histogram / successor count / shortest solutions merged at
    the end -- equal to the one-rank run and to the known answers; then with one puzzle reported as
    not fitting on chip, which both ranks hand to the hash-partitioned driver together."""
    gold = {b["name"]: b for b in _golden_bfs()}
    names = ["puzzle_multi_111", "puzzle_multi_180", "puzzle_multi_111", "puzzle_multi_180", "puzzle_multi_111"]
    batch = [gold[n] for n in names]
    want_states = [gold[n]["n_states"] for n in names]
    want_depth = [gold[n]["solve_depth"] for n in names]
    want_levels = [sum(gold[n]["levels"][i] if i < len(gold[n]["levels"]) else 0 for n in names) for i in range(25)]
    one = _run_local(1, batch, with_paths=True)
    two = _run_local(2, batch, with_paths=True, port=29536)
    for got in (one, two):
        n_states, levels, solve_depth, spp, dpp, generated, solutions, n_fb = got
        assert spp == want_states and dpp == want_depth and n_fb == 0
        assert levels == want_levels and n_states == sum(want_states) and solve_depth == 8 and generated == 4 * n_states
        assert [len(s) for s in solutions] == want_depth
    assert one[:6] == two[:6]
    fb = _run_local(2, batch, too_big=[1, 4], port=29537)
    assert fb[3] == want_states and fb[4] == want_depth and fb[1] == want_levels and fb[7] == 2 and fb[5] == 4 * sum(want_states)
    lim = _run_local(2, batch, max_depth=5, port=29538)
    assert lim[1] == want_levels[:6] and lim[4] == [-1] * 5
