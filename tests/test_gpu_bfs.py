"""BFS kernels on the GPU (ts_bfs_*): known answers recorded from a plain BFS over the
reference's move (tests/golden/misc.json, SURVEY 8(c): 29 / 558 / 950 / 51 states, depths
1 / 8 / 13 / 7) and random puzzle batches against the CPU oracle's BFS."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ts():
    import tiler_slider_b200 as t
    t.lib()
    return t


def puzzle_of(ts, b):
    return ts.Puzzle(b["size"], [tuple(x) for x in b["blocked"]], [tuple(x) for x in b["tiles"]],
                     [tuple(x) for x in b["targets"]], b["multi_color"])


def test_known_answers_single_puzzles(ts, golden_misc):
    from tiler_slider_b200.bfs import solve_puzzle
    for b in golden_misc["bfs"]:
        res = solve_puzzle(puzzle_of(ts, b), table_capacity=1 << 16)
        assert (res.n_states, res.levels, res.solve_depth) == (b["n_states"], b["levels"], b["solve_depth"]), b["name"]
        assert res.states_per_puzzle.tolist() == [b["n_states"]]
        assert res.generated == 4 * b["n_states"]


def test_batched_puzzles_share_one_table(ts, golden_misc):
    from tiler_slider_b200.bfs import BfsSolver
    gold = {b["name"]: b for b in golden_misc["bfs"]}
    batch = [gold["puzzle_multi_111"], gold["puzzle_multi_180"]] * 3
    res = BfsSolver([puzzle_of(ts, b) for b in batch], table_capacity=1 << 16).solve()
    assert res.states_per_puzzle.tolist() == [558, 950] * 3
    assert res.solve_depth_per_puzzle.tolist() == [8, 13] * 3
    assert res.n_states == 3 * (558 + 950) and res.solve_depth == 8


@pytest.mark.parametrize("S,T,W,multi,n", [(4, 2, 3, True, 48), (4, 3, 2, False, 32), (5, 2, 5, False, 32),
                                            (6, 3, 9, True, 24), (3, 4, 1, True, 32), (7, 2, 12, True, 16),
                                            (8, 2, 20, False, 16)])
def test_random_batches_vs_oracle_bfs(ts, S, T, W, multi, n):
    from tiler_slider_b200.bfs import BfsSolver
    from tests.helpers import random_puzzles
    rng = np.random.default_rng(S * 100 + T)
    blocked, tiles, targets = random_puzzles(rng, n, S, T, W)
    table = ts.BatchedTilerSliderEnv.from_arrays(S, blocked, tiles, targets, multi)
    res = BfsSolver(table, table_capacity=1 << 20).solve()
    want_states, want_depth, want_levels = [], [], {}
    for e in range(n):
        b = [(c // S, c % S) for c in np.flatnonzero(blocked[e])]
        st = orc.OracleState(S, b, tiles[e].tolist(), targets[e].tolist(), multi)
        ns, lv, depth, _ = st.bfs(max_states=1 << 20)
        want_states.append(ns)
        want_depth.append(depth)
        for d, c in enumerate(lv):
            want_levels[d] = want_levels.get(d, 0) + c
    assert res.states_per_puzzle.tolist() == want_states
    assert res.solve_depth_per_puzzle.tolist() == want_depth
    assert res.levels == [want_levels[d] for d in range(len(want_levels))]


def _replay(b_or_args, moves):
    size, blocked, tiles, targets, multi = b_or_args
    st = orc.OracleState(size, blocked, tiles, targets, multi)
    won_at = []
    for k, ch in enumerate(moves):
        if st.move("UDLR".index(ch)):
            won_at.append(k + 1)
    return won_at


def test_shortest_solution_strings(ts, golden_misc):
    """parent tracking + traceback.  Every solved puzzle gets a move string of
    exactly its BFS solve depth that wins on its last move (replayed through the oracle) and
    not earlier; unsolvable puzzles get None."""
    from tiler_slider_b200.bfs import BfsSolver
    from tests.helpers import random_puzzles
    for b in golden_misc["bfs"]:
        res = BfsSolver([puzzle_of(ts, b)], table_capacity=1 << 16).solve(with_paths=True)
        sol = res.solutions[0]
        if b["solve_depth"] < 0:
            assert sol is None
            continue
        assert len(sol) == b["solve_depth"], (b["name"], sol)
        assert _replay((b["size"], b["blocked"], b["tiles"], b["targets"], b["multi_color"]), sol) == [len(sol)]
    S, T, W, n = 5, 2, 4, 64
    rng = np.random.default_rng(12)
    blocked, tiles, targets = random_puzzles(rng, n, S, T, W)
    for multi in (True, False):
        table = ts.BatchedTilerSliderEnv.from_arrays(S, blocked, tiles, targets, multi)
        res = BfsSolver(table, table_capacity=1 << 18).solve(with_paths=True)
        n_solved = 0
        for e in range(n):
            depth = int(res.solve_depth_per_puzzle[e])
            sol = res.solutions[e]
            if depth < 0:
                assert sol is None
                continue
            n_solved += 1
            bl = [(c // S, c % S) for c in np.flatnonzero(blocked[e])]
            assert len(sol) == depth
            assert _replay((S, bl, tiles[e].tolist(), targets[e].tolist(), multi), sol)[:1] == [depth]
        assert n_solved > 5


@pytest.mark.parametrize("S,T,W,multi,n", [(5, 2, 4, True, 96), (6, 3, 9, False, 64), (4, 1, 2, False, 48)])
def test_solutions_of_reachable_targets_and_boards_that_start_on_their_goal(ts, S, T, W, multi, n):
    """Targets = where the tiles stand after a random walk, so every puzzle has a solution -- and a
    few walks end where they began: such a board is won by its first step (the goal is evaluated
    inside step(), environment.py:133), so its solution is one move (or a cycle), never the empty
    string.  Depths equal the oracle's BFS; every string wins on its last move, not earlier; the
    hash-partitioned search (device- and host-driven) and the on-chip search agree."""
    from tiler_slider_b200.bfs import BfsSolver, LocalBfs
    from tests.helpers import random_puzzles, reachable_targets
    rng = np.random.default_rng(S * 10 + T)
    blocked, tiles, _ = random_puzzles(rng, n, S, T, W)
    targets = reachable_targets(orc, rng, S, blocked, tiles, multi, 9)
    tiles[0] = targets[0]                                  # at least one board starts on its goal
    table = ts.BatchedTilerSliderEnv.from_arrays(S, blocked, tiles, targets, multi)
    want = []
    for e in range(n):
        bl = [(c // S, c % S) for c in np.flatnonzero(blocked[e])]
        want.append(orc.OracleState(S, bl, tiles[e].tolist(), targets[e].tolist(), multi).bfs(max_states=1 << 20)[2])
    assert sum(d > 0 for d in want) > n // 2
    on_goal = [e for e in range(n) if (tiles[e] == targets[e]).all()]
    assert on_goal and all(want[e] != 0 for e in on_goal)
    for res in (BfsSolver(table, table_capacity=1 << 20).solve(with_paths=True),
                BfsSolver(table, table_capacity=1 << 20).solve(with_paths=True, device_driven=False),
                LocalBfs(table).solve(with_paths=True)):
        assert res.solve_depth_per_puzzle.tolist() == want
        for e in range(n):
            if want[e] < 0:
                assert res.solutions[e] is None
                continue
            bl = [(c // S, c % S) for c in np.flatnonzero(blocked[e])]
            assert len(res.solutions[e]) == want[e]
            assert _replay((S, bl, tiles[e].tolist(), targets[e].tolist(), multi), res.solutions[e])[:1] == [want[e]]


def test_more_than_four_tiles_single_puzzle(ts):
    from tiler_slider_b200.bfs import solve_puzzle
    p = ts.Puzzle(5, [(2, 2), (0, 3)], [(0, 0), (0, 1), (1, 0), (4, 4), (3, 3)], [(4, 0), (4, 1), (4, 2), (4, 3), (0, 4)], False)
    res = solve_puzzle(p, table_capacity=1 << 20)
    st = orc.OracleState(5, p.blocked_locations, p.initial_locations, p.target_locations, False)
    ns, lv, depth, _ = st.bfs(max_states=1 << 20)
    assert (res.n_states, res.levels, res.solve_depth) == (ns, lv, depth)


def test_device_driven_and_host_driven_searches_agree(ts):
    """The single-rank default keeps the frontier sizes on the device and launches 16 levels
    between host read-backs; the host-driven loop (one read-back per level, the multi-rank
    code path) must give the same search: level histogram, per-puzzle tallies, solution lengths."""
    from tiler_slider_b200.bfs import BfsSolver
    table = ts.BatchedTilerSliderEnv.synthetic(512, 6, 4, 8, True, seed=77)
    for kw in (dict(), dict(with_paths=True), dict(max_depth=5), dict(max_depth=16), dict(max_depth=17)):
        a = BfsSolver(table, table_capacity=1 << 22).solve(device_driven=True, **kw)
        b = BfsSolver(table, table_capacity=1 << 22).solve(device_driven=False, **kw)
        assert (a.n_states, a.levels, a.solve_depth, a.generated) == (b.n_states, b.levels, b.solve_depth, b.generated), kw
        assert torch.equal(a.states_per_puzzle, b.states_per_puzzle)
        assert torch.equal(a.solve_depth_per_puzzle, b.solve_depth_per_puzzle)
        if kw.get("with_paths"):
            assert [None if x is None else len(x) for x in a.solutions] == [None if x is None else len(x) for x in b.solutions]
    single = ts.BatchedTilerSliderEnv.synthetic(1, 6, 4, 8, False, seed=3)
    a = BfsSolver(single, table_capacity=1 << 18).solve(device_driven=True)
    b = BfsSolver(single, table_capacity=1 << 18).solve(device_driven=False)
    assert (a.n_states, a.levels, a.solve_depth, a.generated) == (b.n_states, b.levels, b.solve_depth, b.generated)


def test_table_overflow_is_reported(ts, golden_misc):
    from tiler_slider_b200.bfs import solve_puzzle
    b = [x for x in golden_misc["bfs"] if x["name"] == "puzzle_multi_180"][0]
    with pytest.raises(RuntimeError, match="table is full"):
        solve_puzzle(puzzle_of(ts, b), table_capacity=256)


def test_level_corpus_on_gpu(ts):
    """All 400 real levels of the reference (tests/golden/levels_400.txt) through the product
    path: text loader -> batches per shape -> BFS with parent tracking.  State counts and solve
    depths must equal the oracle's BFS, and every returned move string must solve its puzzle on
    its last move when replayed through the oracle."""
    import os
    from tiler_slider_b200.bfs import BfsSolver
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "levels_400.txt")
    puzzles = ts.load_puzzle_file(path)
    assert len(puzzles) == 400
    groups = {}
    for p in puzzles:
        groups.setdefault((p.size, len(p.initial_locations), p.multiple_colors), []).append(p)
    n_checked = 0
    for (S, T, multi), ps in groups.items():
        res = BfsSolver(ps, table_capacity=1 << 20).solve(with_paths=True)
        for i, p in enumerate(ps):
            st = orc.OracleState(S, p.blocked_locations, p.initial_locations, p.target_locations, multi)
            n, _, depth, _ = st.bfs()
            assert int(res.states_per_puzzle[i]) == n and int(res.solve_depth_per_puzzle[i]) == depth
            sol = res.solutions[i]
            assert sol is not None and len(sol) == depth
            assert _replay((S, p.blocked_locations, p.initial_locations, p.target_locations, multi), sol) == [depth]
            n_checked += 1
    assert n_checked == 400


# ---- K6: one CTA per puzzle, visited bitmap in shared memory (LocalBfs) --------------------------
def _oracle_batch(S, multi, blocked, tiles, targets):
    states, depths, levels = [], [], {}
    for e in range(len(blocked)):
        b = [(c // S, c % S) for c in np.flatnonzero(blocked[e])]
        ns, lv, depth, _ = orc.OracleState(S, b, tiles[e].tolist(), targets[e].tolist(), multi).bfs(max_states=1 << 20)
        states.append(ns)
        depths.append(depth)
        for d, c in enumerate(lv):
            levels[d] = levels.get(d, 0) + c
    return states, depths, [levels[d] for d in range(len(levels))]


def test_local_bfs_known_answers(ts, golden_misc):
    """The on-chip search on the known answers (29 / 558 / 950 / 51 states, depths
    1 / 8 / 13 / 7), one puzzle at a time and with paths."""
    from tiler_slider_b200.bfs import LocalBfs
    for b in golden_misc["bfs"]:
        res = LocalBfs([puzzle_of(ts, b)]).solve(with_paths=True)
        assert res.fallback_puzzles == 0
        assert (res.n_states, res.levels, res.solve_depth) == (b["n_states"], b["levels"], b["solve_depth"]), b["name"]
        assert res.states_per_puzzle.tolist() == [b["n_states"]] and res.generated == 4 * b["n_states"]
        sol = res.solutions[0]
        if b["solve_depth"] < 0:
            assert sol is None
        else:
            assert len(sol) == b["solve_depth"]
            assert _replay((b["size"], b["blocked"], b["tiles"], b["targets"], b["multi_color"]), sol) == [len(sol)]


@pytest.mark.parametrize("S,T,W,multi,n", [(4, 2, 3, True, 48), (4, 3, 2, False, 32), (5, 2, 5, False, 32), (6, 3, 9, True, 24),
                                            (3, 4, 1, True, 32), (7, 2, 12, True, 16), (8, 2, 20, False, 16), (5, 1, 3, False, 40),
                                            (6, 4, 10, False, 24), (6, 4, 8, True, 96), (8, 3, 30, True, 12), (2, 2, 0, False, 8),
                                            (1, 1, 0, False, 4)])
def test_local_bfs_random_batches_vs_oracle(ts, S, T, W, multi, n):
    from tiler_slider_b200.bfs import LocalBfs
    from tests.helpers import random_puzzles
    rng = np.random.default_rng(S * 100 + T + 7)
    blocked, tiles, targets = random_puzzles(rng, n, S, T, W)
    table = ts.BatchedTilerSliderEnv.from_arrays(S, blocked, tiles, targets, multi)
    res = LocalBfs(table).solve(with_paths=True)
    want_states, want_depth, want_levels = _oracle_batch(S, multi, blocked, tiles, targets)
    assert res.fallback_puzzles == 0
    assert res.states_per_puzzle.tolist() == want_states
    assert res.solve_depth_per_puzzle.tolist() == want_depth
    assert res.levels == want_levels and res.generated == 4 * sum(want_states)
    for e in range(n):
        sol = res.solutions[e]
        if want_depth[e] < 0:
            assert sol is None
        else:
            bl = [(c // S, c % S) for c in np.flatnonzero(blocked[e])]
            assert len(sol) == want_depth[e]
            assert _replay((S, bl, tiles[e].tolist(), targets[e].tolist(), multi), sol)[:1] == [want_depth[e]]


def test_local_bfs_agrees_with_hash_partitioned_search(ts):
    """4,096 benchmark-shape puzzles (a few of them outgrow the shared-memory part of the queue and
    spill to HBM): per-puzzle state counts, solve depths, the level histogram and the successor
    count must equal the hash-partitioned search; also with a depth limit, and in set-goal mode."""
    from tiler_slider_b200.bfs import BfsSolver, LocalBfs
    for multi, n in ((True, 4096), (False, 1024)):
        table = ts.BatchedTilerSliderEnv.synthetic(n, 6, 4, 8, multi, seed=1004)
        for kw in (dict(), dict(max_depth=7), dict(max_depth=1)):
            a = BfsSolver(table, table_capacity=1 << 26).solve(**kw)
            loc = LocalBfs(table)
            b = loc.solve(**kw)
            assert b.fallback_puzzles == 0 and loc.plan()["ctas_per_sm"] >= 2
            assert (a.n_states, a.levels, a.solve_depth, a.generated) == (b.n_states, b.levels, b.solve_depth, b.generated), (multi, kw)
            assert torch.equal(a.states_per_puzzle, b.states_per_puzzle)
            assert torch.equal(a.solve_depth_per_puzzle.to(torch.int32), b.solve_depth_per_puzzle)
        if multi:
            assert int(b.states_per_puzzle.max()) >= 1  # (depth-limited run)
            full = LocalBfs(table).solve()
            assert int(full.states_per_puzzle.max()) > loc.plan()["queue_smem"]      # the spill path was exercised


def test_local_bfs_queue_spill_and_fallback(ts):
    """With only 64 queue entries in shared memory nearly every state of every puzzle goes through
    the HBM spill slab (paths included).  With two walls the state space (34*33*32*31 bits = 139 KB)
    does not fit the bitmap of two CTAs per SM: the planner drops to one CTA per SM; asked for more
    than fit, the puzzles are searched by the hash-partitioned path instead (fallback), same answers."""
    from tiler_slider_b200.bfs import BfsSolver, LocalBfs
    batch = ts.BatchedTilerSliderEnv.synthetic(300, 6, 4, 8, True, seed=5)
    ref = BfsSolver(batch, table_capacity=1 << 22).solve(with_paths=True)
    tiny = LocalBfs(batch, queue_smem=64)
    res = tiny.solve(with_paths=True)
    assert tiny.plan()["queue_smem"] == 64 and res.fallback_puzzles == 0
    assert torch.equal(res.states_per_puzzle, ref.states_per_puzzle) and res.levels == ref.levels
    assert torch.equal(res.solve_depth_per_puzzle, ref.solve_depth_per_puzzle.to(torch.int32))
    assert [None if x is None else len(x) for x in res.solutions] == [None if x is None else len(x) for x in ref.solutions]
    blocked, tiles, targets = batch.blocked_cells().cpu().numpy(), batch.positions().cpu().numpy(), batch.target_positions().cpu().numpy()
    for e in range(0, 300, 7):
        if res.solutions[e] is not None:
            bl = [(c // 6, c % 6) for c in np.flatnonzero(blocked[e])]
            assert _replay((6, bl, tiles[e].tolist(), targets[e].tolist(), True), res.solutions[e])[:1] == [len(res.solutions[e])]
    few_walls = ts.BatchedTilerSliderEnv.synthetic(64, 6, 4, 2, True, seed=9)      # F = 34: 139 KB of bitmap
    want = BfsSolver(few_walls, table_capacity=1 << 25).solve()
    one = LocalBfs(few_walls)
    got = one.solve()
    assert one.plan()["ctas_per_sm"] == 1 and got.fallback_puzzles == 0
    forced = LocalBfs(few_walls, ctas_per_sm=2)                                     # does not fit twice per SM
    assert forced.plan() is None
    fb = forced.solve()
    assert fb.fallback_puzzles == 64
    for r in (got, fb):
        assert torch.equal(r.states_per_puzzle, want.states_per_puzzle) and r.levels == want.levels
        assert torch.equal(r.solve_depth_per_puzzle.to(torch.int32), want.solve_depth_per_puzzle.to(torch.int32))


def test_local_bfs_level_corpus(ts):
    """The reference's 400 real levels through the on-chip search, with shortest solutions."""
    import os
    from tiler_slider_b200.bfs import LocalBfs
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "levels_400.txt")
    puzzles = ts.load_puzzle_file(path)
    groups = {}
    for p in puzzles:
        groups.setdefault((p.size, len(p.initial_locations), p.multiple_colors), []).append(p)
    n_checked = 0
    for (S, T, multi), ps in groups.items():
        loc = LocalBfs(ps)
        res = loc.solve(with_paths=True)
        assert res.fallback_puzzles == 0 and loc.plan()["ctas_per_sm"] >= 4      # tiny bitmaps: many CTAs per SM
        for i, p in enumerate(ps):
            st = orc.OracleState(S, p.blocked_locations, p.initial_locations, p.target_locations, multi)
            n, _, depth, _ = st.bfs()
            assert int(res.states_per_puzzle[i]) == n and int(res.solve_depth_per_puzzle[i]) == depth
            sol = res.solutions[i]
            assert sol is not None and len(sol) == depth
            assert _replay((S, p.blocked_locations, p.initial_locations, p.target_locations, multi), sol) == [depth]
            n_checked += 1
    assert n_checked == 400


def test_solve_batch_picks_the_search_that_fits(ts, golden_misc):
    """bfs.solve_batch: up to 4 tiles -> the on-chip search; 5..8 tiles -> the hash-partitioned search,
    puzzle by puzzle (its key has no room for a puzzle id).  Same answers as the oracle either way."""
    from tiler_slider_b200.bfs import solve_batch
    gold = {b["name"]: b for b in golden_misc["bfs"]}
    r = solve_batch([puzzle_of(ts, gold["puzzle_multi_111"]), puzzle_of(ts, gold["puzzle_multi_180"])], with_paths=True)
    assert r.states_per_puzzle.tolist() == [558, 950] and r.solve_depth_per_puzzle.tolist() == [8, 13] and [len(s) for s in r.solutions] == [8, 13]
    five = [ts.Puzzle(5, [(2, 2), (0, 3)], [(0, 0), (0, 1), (1, 0), (4, 4), (3, 3)], [(4, 0), (4, 1), (4, 2), (4, 3), (0, 4)], False),
            ts.Puzzle(5, [(1, 1)], [(0, 0), (0, 1), (2, 0), (4, 4), (3, 3)], [(4, 0), (4, 1), (4, 2), (4, 3), (0, 4)], False)]
    r5 = solve_batch(five)
    for i, p in enumerate(five):
        n, lv, depth, _ = orc.OracleState(5, p.blocked_locations, p.initial_locations, p.target_locations, False).bfs(max_states=1 << 20)
        assert int(r5.states_per_puzzle[i]) == n and int(r5.solve_depth_per_puzzle[i]) == depth
    assert r5.n_states == int(r5.states_per_puzzle.sum())
