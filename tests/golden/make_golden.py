"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box):   python tests/golden/make_golden.py
The reference package eagerly imports pygame and matplotlib (explainrl/environment/
__init__.py:13-16), neither of which is installed, so two empty stub modules are
registered first; no reference code is modified or copied.

Outputs (committed):
  scenarios.json   the two scripted games of tests/test_user_scenarios.py replayed through
                   TilerSliderEnv + TextRender, plus SURVEY appendix B.3, with board string,
                   positions, done and info after every step
  rollouts.npz     seeded random puzzles of many shapes x random action strings, reference
                   positions / flags / step_count per step under `step; if done: reset`
  misc.json        slide tables, observations, text-grammar parses, seeded factory puzzles,
                   valid-move lists, win-logic cases, BFS level histograms (BFS is a plain
                   Python loop over the reference's GameState.move)
"""
import json
import os
import sys
import types

import numpy as np

REF = os.environ.get("TS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

for name in ("pygame", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, REF)
from explainrl.environment import GameState, TilerSliderEnv, TilerSliderEnvFactory, TextRender, ImageLoader  # noqa: E402

Move = GameState.Move
F_DONE, F_WON, F_INVALID, F_TIMEOUT = 1, 2, 4, 8


def tolist(locs):
    return [[int(r), int(c)] for r, c in locs]


def play(env, moves, render=True):
    """Replay a move string; record everything the reference exposes after each step."""
    env.reset()
    rend = TextRender(env)
    rec = {"initial_board": rend.render(show_info=False), "steps": []}
    for ch in moves:
        obs, done, info = env.step(Move.from_char(ch))
        rec["steps"].append({
            "move": ch,
            "board": rend.render(show_info=False) if render else None,
            "positions": tolist(env.state.current_locations),
            "done": bool(done),
            "info": {k: (bool(v) if isinstance(v, (bool, np.bool_)) else int(v)) for k, v in info.items()},
            "obs_sum": float(obs.sum()),
        })
        if done:
            break
    return rec


def scenarios():
    lvl = dict(size=4, blocked_locations=[(1, 0), (2, 3)], initial_locations=[(0, 3), (3, 2)],
               target_locations=[(0, 0), (3, 0)], multiple_colors=True)
    out = []
    for moves in ("RDLULD", "DLDRRL"):
        env = TilerSliderEnv.from_level(ImageLoader.ImageProcessed(**lvl))
        rec = play(env, moves)
        rec["puzzle"] = {"size": 4, "blocked": tolist(lvl["blocked_locations"]),
                         "tiles": tolist(lvl["initial_locations"]),
                         "targets": tolist(lvl["target_locations"]), "multi_color": True}
        rec["moves"] = moves
        out.append(rec)
    grid = "a..X.b\n.X....\n..c..X\nX...d.\n.D.X.C\nB..A..\n"
    env = TilerSliderEnvFactory.create_from_string(grid, multi_color=True)
    rec = play(env, "DRULDLUR")
    rec["puzzle"] = {"size": env.size, "blocked": tolist(env.blocked_locations),
                     "tiles": tolist(env.initial_locations), "targets": tolist(env.target_locations),
                     "multi_color": True, "text": grid}
    rec["moves"] = "DRULDLUR"
    out.append(rec)
    # max_steps=1 with a winning move: success and timeout on the same step
    env = TilerSliderEnv(size=3, blocked_locations=[], initial_locations=[(0, 0)],
                         target_locations=[(0, 2)], multi_color=False, max_steps=1)
    rec = play(env, "R")
    rec["puzzle"] = {"size": 3, "blocked": [], "tiles": [[0, 0]], "targets": [[0, 2]],
                     "multi_color": False, "max_steps": 1}
    rec["moves"] = "R"
    out.append(rec)
    return out


def random_puzzle(rng, S, T, W):
    cells = [(i, j) for i in range(S) for j in range(S)]
    perm = rng.permutation(len(cells))
    pick = [cells[k] for k in perm]
    if W + 2 * T <= len(cells):
        return pick[:W], pick[W:W + T], pick[W + T:W + 2 * T]
    # too crowded for disjoint tiles and targets: targets may sit under tiles (legal)
    free = pick[W:]
    tperm = rng.permutation(len(free))
    return pick[:W], free[:T], [free[k] for k in tperm[:T]]


def rollouts():
    """Shapes cover the bench configs (5x5/1, 6x6/4, 12x12/8) and the edges: 1x1, 2x2,
    no walls, crowded boards, both colour modes, small max_steps (timeouts + resets)."""
    rng = np.random.default_rng(20261018)
    shapes = [  # S, T, W, multi, n_envs, K, max_steps
        (1, 1, 0, False, 2, 6, 100), (2, 1, 1, False, 8, 12, 5), (2, 2, 0, True, 8, 12, 7),
        (3, 2, 2, False, 16, 24, 9), (3, 3, 1, True, 16, 24, 100), (4, 2, 2, True, 24, 32, 11),
        (4, 4, 3, False, 24, 32, 100), (5, 1, 5, False, 48, 48, 17), (5, 3, 4, True, 32, 40, 100),
        (6, 4, 8, True, 64, 64, 23), (6, 4, 8, False, 32, 48, 100), (6, 3, 0, True, 16, 32, 100),
        (7, 5, 10, True, 24, 40, 31), (8, 8, 12, True, 24, 40, 100), (8, 6, 20, False, 16, 40, 13),
        (9, 4, 20, True, 12, 32, 100), (12, 8, 36, True, 32, 64, 29), (12, 8, 36, False, 12, 40, 100),
        (16, 8, 60, True, 8, 40, 100), (6, 8, 4, True, 16, 40, 100), (5, 8, 2, False, 16, 40, 19),
    ]
    out = {}
    meta = []
    for idx, (S, T, W, multi, n, K, max_steps) in enumerate(shapes):
        blocked = np.zeros((n, S * S), dtype=np.uint8)
        tiles = np.zeros((n, T, 2), dtype=np.uint8)
        targets = np.zeros((n, T, 2), dtype=np.uint8)
        actions = rng.integers(0, 4, size=(K, n), dtype=np.uint8)
        pos = np.zeros((K, n, T, 2), dtype=np.int16)
        flags = np.zeros((K, n), dtype=np.uint8)
        count = np.zeros((K, n), dtype=np.int32)
        obs_final = np.zeros((n, S, S, 3), dtype=np.float32)
        for e in range(n):
            b, i, t = random_puzzle(rng, S, T, W)
            for r, c in b:
                blocked[e, r * S + c] = 1
            tiles[e] = np.array(i, dtype=np.uint8).reshape(T, 2)
            targets[e] = np.array(t, dtype=np.uint8).reshape(T, 2)
            env = TilerSliderEnv(S, b, i, t, multi_color=multi, max_steps=max_steps)
            obs = env.reset()
            for k in range(K):
                obs, done, info = env.step(Move(int(actions[k, e])))
                pos[k, e] = np.array(tolist(env.state.current_locations), dtype=np.int16).reshape(T, 2)
                f = (F_DONE if done else 0) | (F_WON if info["is_won"] else 0)
                f |= (F_INVALID if info["invalid_move"] else 0) | (F_TIMEOUT if info.get("timeout") else 0)
                assert bool(info.get("success", False)) == bool(info["is_won"])
                flags[k, e] = f
                count[k, e] = info["step_count"]
                if done:
                    obs = env.reset()
            obs_final[e] = obs
        tag = f"r{idx:02d}"
        out.update({f"{tag}_blocked": blocked, f"{tag}_tiles": tiles, f"{tag}_targets": targets,
                    f"{tag}_actions": actions, f"{tag}_pos": pos, f"{tag}_flags": flags,
                    f"{tag}_count": count, f"{tag}_obs_final": obs_final})
        meta.append(dict(tag=tag, S=S, T=T, W=W, multi=bool(multi), n=n, K=K, max_steps=max_steps))
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    return out


def bfs_levels(size, blocked, tiles, targets, multi):
    """Plain BFS over the reference's GameState.move (the reference has no solver)."""
    def key(locs):
        cells = [int(r) * size + int(c) for r, c in locs]
        return tuple(cells if multi else sorted(cells))
    start = GameState(size, blocked, list(tiles), list(targets), multi)
    seen = {key(start.current_locations)}
    frontier = [list(map(tuple, start.current_locations))]
    levels, solve_depth, depth = [1], -1, 0
    while frontier:
        depth += 1
        nxt = []
        for locs in frontier:
            for mv in Move:
                st = GameState(size, blocked, list(locs), list(targets), multi)
                won = st.move(mv)
                if won and solve_depth < 0:
                    solve_depth = depth
                k = key(st.current_locations)
                if k not in seen:
                    seen.add(k)
                    nxt.append([(int(r), int(c)) for r, c in st.current_locations])
        if nxt:
            levels.append(len(nxt))
        frontier = nxt
    return {"n_states": len(seen), "levels": levels, "solve_depth": solve_depth}


def misc():
    out = {}
    # slide tables + observations
    tabs = []
    for S, blocked in [(3, []), (4, [(1, 0), (2, 3)]), (5, [(2, 2), (0, 4), (4, 0)]),
                       (6, [(0, 3), (1, 1), (2, 5), (3, 0), (4, 3)])]:
        st = GameState(S, blocked, [(0, 0)], [(S - 1, S - 1)], False)
        tabs.append({"size": S, "blocked": tolist(blocked), "move_to": st.move_to.tolist()})
    out["slide_tables"] = tabs
    obs = []
    for text, multi in [("ab..\n.X..\n....\n..BA", True), ("ab..\n.X..\n....\n..BA", False),
                        ("a.X\n.b.\nB.A", True)]:
        env = TilerSliderEnvFactory.create_from_string(text, multi_color=multi)
        o = env.reset()
        obs.append({"text": text, "multi_color": multi, "obs": o.tolist()})
    out["observations"] = obs
    # text grammar
    gram = []
    for text in ["a.A", "c.a\n...\nA.C", "ab..\n.X..\n....\n..BA", "  a..  \n\n .X. \n ..A \n",
                 "a.b\n.xX\nB.A", "aB\nA.b\n", "...\n...\n...", "a1?\n#.A\n.-.", "x.X\n...\n..A"]:
        env = TilerSliderEnvFactory.create_from_string(text)
        gram.append({"text": text, "size": env.size, "blocked": tolist(env.blocked_locations),
                     "tiles": tolist(env.initial_locations), "targets": tolist(env.target_locations)})
    out["grammar"] = gram
    # seeded factory
    fac = []
    for kw in [dict(size=5, num_tiles=2, num_obstacles=3, seed=42), dict(size=6, num_tiles=4, num_obstacles=8, seed=7),
               dict(size=4, num_tiles=1, num_obstacles=0, seed=0), dict(size=12, num_tiles=8, num_obstacles=36, seed=123)]:
        env = TilerSliderEnvFactory.create_simple_env(**kw)
        fac.append({"kwargs": kw, "blocked": tolist(env.blocked_locations),
                    "tiles": tolist(env.initial_locations), "targets": tolist(env.target_locations),
                    "multi_color": bool(env.multi_color)})
    out["factory"] = fac
    # collisions (tests/test_state.py:248-311) and win logic (:317-377) -- recorded from the reference
    col = []
    for S, blocked, tiles, mv in [(5, [], [(4, 2), (3, 2)], "U"), (5, [], [(2, 4), (2, 3)], "L"),
                                  (6, [], [(5, 1), (4, 1), (3, 1)], "U"), (5, [(2, 1)], [(4, 1), (3, 1)], "U"),
                                  (3, [(0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1), (2, 2)], [(0, 0)], "R"),
                                  (6, [(2, 2)], [(0, 2), (1, 2), (4, 2), (5, 2)], "D"),
                                  (6, [(3, 3)], [(3, 0), (3, 1), (3, 4), (3, 5)], "R")]:
        st = GameState(S, blocked, list(tiles), [(0, 0)] * len(tiles), False)
        st.move(Move.from_char(mv))
        col.append({"size": S, "blocked": tolist(blocked), "tiles": tolist(tiles), "move": mv,
                    "after": tolist(st.current_locations)})
    out["collisions"] = col
    wins = []
    for tiles, targets, multi in [([(0, 0), (1, 1)], [(1, 1), (0, 0)], False), ([(0, 0), (1, 1)], [(1, 1), (0, 0)], True),
                                  ([(0, 0), (1, 1)], [(0, 0), (1, 1)], True), ([], [], False), ([], [], True),
                                  ([(0, 0)], [(0, 0), (0, 0)], False), ([(0, 0), (1, 1)], [(0, 0)], False),
                                  ([(0, 0)], [(0, 0), (1, 1)], True), ([], [(0, 0)], False)]:
        st = GameState(3, [], list(tiles), list(targets), multi)
        wins.append({"tiles": tolist(tiles), "targets": tolist(targets), "multi_color": multi,
                     "is_won": bool(st.is_won())})
    out["win_logic"] = wins
    # valid moves
    vm = []
    for text, multi in [("a..\n...\n..A", False), ("aX.\nX..\n..A", False), ("ab..\n.X..\n....\n..BA", True)]:
        env = TilerSliderEnvFactory.create_from_string(text, multi_color=multi)
        env.reset()
        vm.append({"text": text, "multi_color": multi, "valid": [m.value for m in env.get_valid_moves()]})
    out["valid_moves"] = vm
    # BFS known answers; puzzles decoded from data/*.jpg by the survey
    bfs = []
    for name, S, blocked, tiles, targets, multi in [
        ("puzzle_multi_001", 4, [(1, 0), (2, 3)], [(0, 3), (3, 2)], [(0, 0), (3, 0)], True),
        ("puzzle_multi_111", 6, [(0, 0), (0, 1), (1, 1), (2, 4), (3, 4), (4, 1), (4, 3), (5, 5)],
         [(3, 5), (0, 5), (2, 1)], [(4, 2), (5, 2), (5, 4)], True),
        ("puzzle_multi_180", 6, [(0, 1), (0, 3), (1, 1), (3, 0), (3, 2), (3, 3), (4, 0), (4, 5), (5, 2)],
         [(4, 1), (3, 4), (1, 2)], [(3, 5), (5, 1), (5, 4)], True),
        ("puzzle_single_151", 6, [(0, 3), (0, 4), (0, 5), (1, 1), (1, 3), (3, 3), (4, 3), (4, 5), (5, 3)],
         [(1, 4), (2, 2), (4, 2)], [(0, 2), (1, 5), (5, 5)], False),
        ("open_5x5_2tiles_multi", 5, [(2, 2)], [(0, 0), (4, 4)], [(4, 0), (0, 4)], True),
        ("open_5x5_2tiles_single", 5, [(2, 2)], [(0, 0), (4, 4)], [(4, 0), (0, 4)], False),
        ("b3_6x6_4tiles", 6, [(0, 3), (1, 1), (2, 5), (3, 0), (4, 3)], [(0, 0), (0, 5), (2, 2), (3, 4)],
         [(5, 3), (5, 0), (4, 5), (4, 1)], True),
    ]:
        r = bfs_levels(S, blocked, tiles, targets, multi)
        r.update(name=name, size=S, blocked=tolist(blocked), tiles=tolist(tiles), targets=tolist(targets),
                 multi_color=multi)
        bfs.append(r)
    out["bfs"] = bfs
    # round 2: slide tables of every board class with blocked START cells in them (the table is
    # defined there too, state.py:85-118), incl. wide boards
    tabs2 = []
    rng = np.random.default_rng(20261018)
    for S, W in [(7, 9), (8, 14), (9, 20), (12, 36), (14, 40), (15, 50), (16, 64)]:
        cells = rng.permutation(S * S)[:W]
        blocked = [(int(c) // S, int(c) % S) for c in cells]
        st = GameState(S, blocked, [], [], False)
        tabs2.append({"size": S, "blocked": tolist(blocked), "move_to": st.move_to.tolist()})
    out["slide_tables_r2"] = tabs2
    # round 2: degenerate boards the reference accepts -- no tiles (won iff no targets,
    # tests/test_state.py:40-52, tests/test_environment.py:569-580) and multi-colour boards whose
    # target count differs from the tile count (never won, state.py:183-184) -- played through
    # TilerSliderEnv.step with everything it exposes recorded
    deg = []
    for S, blocked, tiles, targets, multi, moves, max_steps in [
        (3, [], [], [], False, "UDLR", 100),
        (3, [], [], [], True, "UD", 100),
        (3, [(1, 1)], [], [(0, 0)], False, "UDLRU", 4),
        (3, [(1, 1)], [], [(0, 0), (2, 2)], True, "LR", 100),
        (4, [(1, 2)], [(0, 0), (3, 3)], [(0, 3)], True, "RDLURD", 100),
        (4, [(1, 2)], [(0, 0)], [(0, 3), (3, 3)], True, "RDLU", 3),
        (5, [(2, 2)], [(0, 0), (4, 4), (0, 4)], [], True, "DRUL", 100),
        (5, [(2, 2)], [(0, 0), (4, 4)], [(4, 0)], False, "DRUL", 100),
        (4, [], [(0, 0)], [(3, 0), (3, 0)], False, "DU", 100),
        (12, [(5, 5), (0, 7)], [(0, 0), (11, 11)], [(11, 0), (11, 0)], False, "DRUL", 100),
        (12, [(5, 5)], [(0, 0), (3, 3)], [(11, 0)], True, "DRDL", 100),
        (10, [(5, 5)], [], [(9, 0)], False, "DR", 100),
        (10, [(5, 5)], [], [], False, "D", 100),
    ]:
        env = TilerSliderEnv(size=S, blocked_locations=blocked, initial_locations=list(tiles),
                             target_locations=list(targets), multi_color=multi, max_steps=max_steps)
        obs0 = env.reset()
        rec = {"size": S, "blocked": tolist(blocked), "tiles": tolist(tiles), "targets": tolist(targets),
               "multi_color": multi, "moves": moves, "max_steps": max_steps, "won_at_reset": bool(env.state.is_won()),
               "obs_reset": obs0.tolist(), "valid_at_reset": [m.value for m in env.get_valid_moves()], "steps": []}
        for ch in moves:
            obs, done, info = env.step(Move.from_char(ch))
            rec["steps"].append({"move": ch, "positions": tolist(env.state.current_locations), "done": bool(done),
                                 "state_is_won": bool(env.state.is_won()),
                                 "info": {k: (bool(v) if isinstance(v, (bool, np.bool_)) else int(v)) for k, v in info.items()},
                                 "obs": obs.tolist()})
            if done:
                break
        deg.append(rec)
    out["degenerate"] = deg
    return out


def main():
    with open(os.path.join(HERE, "scenarios.json"), "w") as f:
        json.dump(scenarios(), f, indent=1)
    np.savez_compressed(os.path.join(HERE, "rollouts.npz"), **rollouts())
    with open(os.path.join(HERE, "misc.json"), "w") as f:
        json.dump(misc(), f, indent=None, separators=(",", ":"))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
