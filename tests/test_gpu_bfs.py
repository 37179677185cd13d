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
