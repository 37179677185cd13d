"""Export the reference's 400 level screenshots (data/*.jpg) as text puzzles.

Run in the build container only (needs /root/reference and cv2):
    python tests/golden/make_levels.py
Every image is decoded by the UNMODIFIED reference parser, ImageLoader.parse_puzzle_image
(explainrl/environment/dataloader.py:44-133), exactly the way its CLI does it
(explainrl/environment/play.py:165-216: crop of ImageLoader.__getitem__, multi-colour iff
'_multi_' is in the file name).  matplotlib is not installed, so `plt.imread` is replaced by
cv2 (BGR -> RGB), the shim SURVEY 8(c) describes.  Output: tests/golden/levels_400.txt in the
`-input_file` format of tiler_slider_b200/puzzle.py -- a real-puzzle corpus (4x4 .. 6x6, 1-3
tiles) for the parity and BFS tests; no reference code or image is copied.
"""
import os
import sys
import types

import cv2
import numpy as np

REF = os.environ.get("TS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

plt = types.ModuleType("matplotlib.pyplot")
plt.imread = lambda p: cv2.cvtColor(cv2.imread(p), cv2.COLOR_BGR2RGB)
mpl = types.ModuleType("matplotlib")
mpl.pyplot = plt
sys.modules.update({"pygame": types.ModuleType("pygame"), "matplotlib": mpl, "matplotlib.pyplot": plt})
sys.path.insert(0, REF)
from explainrl.environment import ImageLoader  # noqa: E402


def main():
    os.chdir(os.path.join(REF, "data"))          # ImageLoader() lists the cwd (dataloader.py:29-30)
    loader = ImageLoader()
    out = ["# 400 levels decoded from the reference's data/*.jpg by its own parser (tests/golden/make_levels.py)"]
    shapes = {}
    for i in range(len(loader)):
        raw = loader[i]
        multi = "_multi_" in raw.name
        lvl = ImageLoader.parse_puzzle_image(raw.puzzle_image, multi)
        S = lvl.size
        grid = [["." for _ in range(S)] for _ in range(S)]
        for r, c in lvl.blocked_locations:
            grid[r][c] = "X"
        for k, (r, c) in enumerate(lvl.target_locations):
            grid[r][c] = chr(ord("A") + k)
        for k, (r, c) in enumerate(lvl.initial_locations):
            assert grid[r][c] == ".", "tile on a target / wall cannot be written in the grammar"
            grid[r][c] = chr(ord("a") + k)
        out.append(f"# {os.path.splitext(raw.name)[0]}")
        out.append(f"multi_color: {'true' if multi else 'false'}")
        out.extend("".join(row) for row in grid)
        out.append("---")
        key = (S, len(lvl.initial_locations), multi)
        shapes[key] = shapes.get(key, 0) + 1
    with open(os.path.join(HERE, "levels_400.txt"), "w") as f:
        f.write("\n".join(out[:-1]) + "\n")
    print(len(loader), "levels;", sorted(shapes.items()))


if __name__ == "__main__":
    main()
