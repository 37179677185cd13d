"""BFS known answers for every 8th level of levels_400.txt, computed with a plain Python BFS
over the UNMODIFIED reference GameState.move (build container only).  Pins the oracle's BFS on
real puzzles: tests/test_oracle_golden.py::test_level_corpus_bfs."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import make_golden as mg                      # registers the stubs and imports the reference  # noqa: E402
from tiler_slider_b200.puzzle import load_puzzle_file  # noqa: E402  (text loader only; no CUDA is touched)

levels = load_puzzle_file(os.path.join(HERE, "levels_400.txt"))
out = []
for i in range(0, len(levels), 8):
    p = levels[i]
    r = mg.bfs_levels(p.size, p.blocked_locations, p.initial_locations, p.target_locations, p.multiple_colors)
    r["index"] = i
    out.append(r)
with open(os.path.join(HERE, "levels_bfs.json"), "w") as f:
    json.dump(out, f, separators=(",", ":"))
print(len(out), "levels;", sum(r["n_states"] for r in out), "states")
