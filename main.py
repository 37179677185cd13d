"""Entry point kept from the reference: `python3 main.py -input_file input.txt`.

The reference's main.py (main.py:1-6) is a stub that ignores its arguments; README.md:5 is the
only place the `-input_file` flag is named and explainrl/__main__.py:8 spells it
`--input-file`.  Both spellings are accepted here.  The file holds one or more puzzles in the
text grammar of TilerSliderEnvFactory.create_from_string (explainrl/environment/
environment.py:236-288; see tiler_slider_b200/puzzle.py for the file format).  Each puzzle is
loaded into the CUDA environment, reset, and stepped through its scripted `moves:` line (or
the -moves argument); the board is printed after every step in the reference's TextRender
style (explainrl/environment/display.py:56-79: target > tile > blocked > empty).
"""
from __future__ import annotations

import argparse
import sys


def render_board(env) -> str:
    st = env.state
    tiles = [(int(r), int(c)) for r, c in st.current_locations]
    targets = [(int(r), int(c)) for r, c in st.target_locations]
    rows = []
    for i in range(env.size):
        row = []
        for j in range(env.size):
            if (i, j) in targets:
                row.append(chr((targets.index((i, j)) if env.multi_color else 0) + ord("A")))
            elif (i, j) in tiles:
                row.append(chr((tiles.index((i, j)) if env.multi_color else 0) + ord("a")))
            elif st.is_blocked[i, j]:
                row.append("X")
            else:
                row.append(".")
        rows.append("".join(row))
    return "\n".join(rows)


def run_puzzle(puzzle, moves: str, max_steps: int, quiet: bool = False) -> dict:
    from tiler_slider_b200 import Move, TilerSliderEnv
    env = TilerSliderEnv.from_level(puzzle, max_steps=max_steps)
    env.reset()
    if not quiet:
        print(f"Step: {env.step_count}/{env.max_steps}\nDone: {env.done}\n\n{render_board(env)}\n")
    info, done, total_reward = {}, False, 0.0
    for ch in moves:
        mv = Move.from_char(ch)
        if mv is None:
            raise ValueError(f"unknown move {ch!r} (use U, D, L, R)")
        _, done, info = env.step(mv)
        total_reward += env.last_reward
        if not quiet:
            print(f"Move {env.step_count}: {mv.name}  reward={env.last_reward:+.2f}  info={info}")
            print(f"Step: {env.step_count}/{env.max_steps}\nDone: {env.done}\n\n{render_board(env)}\n")
        if done:
            break
    result = {"steps": env.step_count, "done": bool(done), "is_won": bool(info.get("is_won", False)),
              "timeout": bool(info.get("timeout", False)), "total_reward": total_reward,
              "positions": [(int(r), int(c)) for r, c in env.state.current_locations]}
    if not quiet:
        print("Puzzle solved!" if result["is_won"] else ("Timeout!" if result["timeout"] else "Not solved."))
    return result


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="Tiler-Slider on B200: load puzzles from a text file and play scripted moves")
    ap.add_argument("-input_file", "--input-file", "--input_file", dest="input_file", default=None,
                    help="puzzle text file (grammar of create_from_string)")
    ap.add_argument("-level", "--level", default=None,
                    help="level screenshot by name, e.g. puzzle_multi_001 (the reference's play.py --level)")
    ap.add_argument("-data_dir", "--data-dir", dest="data_dir", default="data", help="directory of the level screenshots")
    ap.add_argument("-moves", "--moves", default=None, help="action string over UDLR; overrides the file's `moves:` lines")
    ap.add_argument("-max_steps", "--max-steps", dest="max_steps", type=int, default=None)
    ap.add_argument("-multi_color", "--multi-color", dest="multi_color", action="store_true",
                    help="force ordered tile/target matching for every puzzle of the file")
    ap.add_argument("-solve", "--solve", action="store_true",
                    help="also run the GPU breadth-first search and print a shortest solution per puzzle")
    ap.add_argument("-quiet", "--quiet", action="store_true")
    args = ap.parse_args(argv)

    from tiler_slider_b200 import load_puzzle_file
    if (args.input_file is None) == (args.level is None):
        ap.error("give exactly one of -input_file and -level")
    if args.level is not None:
        from tiler_slider_b200.levels import load_level
        puzzles = [load_level(args.level, args.data_dir)]
    else:
        puzzles = load_puzzle_file(args.input_file)
    for k, p in enumerate(puzzles):
        if args.multi_color:
            p.multiple_colors = True
        moves = (args.moves if args.moves is not None else p.moves).upper()
        max_steps = args.max_steps or p.max_steps or 100
        if not args.quiet:
            print(f"=== puzzle {k + 1}/{len(puzzles)}: {p.size}x{p.size}, {len(p.initial_locations)} tile(s), "
                  f"{'multi' if p.multiple_colors else 'single'}-colour, moves '{moves}' ===")
        res = run_puzzle(p, moves, max_steps, args.quiet)
        print(f"result[{k}]: {res}")
        if args.solve:
            from tiler_slider_b200.bfs import solve_batch
            if p.size > 8 or not p.initial_locations:
                print(f"solution[{k}]: BFS supports board sizes up to 8 with at least one tile")
            else:
                r = solve_batch([p], with_paths=True)
                print(f"solution[{k}]: {r.solutions[0]!r} (depth {r.solve_depth}, {r.n_states} reachable states)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
