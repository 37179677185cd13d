"""Level ingest: phone screenshots of the game -> Puzzle (host-side, one-off).

Same observable behaviour as the reference's ImageLoader (explainrl/environment/dataloader.py:
8-133) and the level lookup of its CLI (explainrl/environment/play.py:165-216), re-implemented
on cv2 + numpy: the board crop of a 1080x2340 screenshot, grid lines found by the background
colour, every cell classified as empty / tile / target / blocked from three probe windows, tiles
paired with targets by colour in multi-colour levels.  `tests/test_levels.py` checks it against
the 400 levels decoded by the reference's own parser (tests/golden/levels_400.txt).

This is load-time code: it produces Puzzle objects for BatchedTilerSliderEnv / TilerSliderEnv,
nothing here is on the step path.
"""
from __future__ import annotations

import os
from typing import Iterable

import numpy as np

from .puzzle import Puzzle, puzzle_to_text

BACKGROUND_RGB = np.array([0, 172, 194])        # dataloader.py:10
EMPTY_CELL_RGB = np.array([223, 247, 249])      # dataloader.py:11
TOLERANCE = 10                                  # per channel, dataloader.py:12
BOARD_ROWS = slice(665, 1710)                   # crop of ImageLoader.__getitem__, dataloader.py:39
BOARD_COLS = slice(15, -15)
MIN_CELL_GAP = 25                               # grid lines further apart than this bound a cell


def _cv2():
    import cv2
    return cv2


def _near(img: np.ndarray, rgb: np.ndarray) -> np.ndarray:
    """255 where every channel is within TOLERANCE of rgb (cv2.inRange)."""
    return _cv2().inRange(img, rgb - TOLERANCE, rgb + TOLERANCE)


def _cell_spans(line_mask_mean: np.ndarray) -> list[tuple[int, int]]:
    """Pixel spans between grid lines.  The reference adds 1 to every line index
    (`[1,] + r_lines` is an elementwise add, dataloader.py:53-54); kept, so spans match."""
    lines = np.flatnonzero(line_mask_mean > 100.0) + 1
    return [(int(a) + 1, int(b)) for a, b in zip(lines[:-1], lines[1:]) if b > a + MIN_CELL_GAP]


def _all_empty(window: np.ndarray) -> bool:
    return bool(np.all(_near(window, EMPTY_CELL_RGB)))


def parse_level_image(board_rgb: np.ndarray, multiple_colors: bool) -> Puzzle:
    """Board crop (RGB) -> Puzzle, classification rules of dataloader.py:64-104:
    whole cell empty-coloured -> empty; centre probe empty -> tile (a ring; colour from the
    top-left probe); top-left probe empty -> target (a dot; colour from the centre probe);
    anything else -> blocked."""
    grid = _near(board_rgb, BACKGROUND_RGB)
    rows, cols = _cell_spans(grid.mean(axis=1)), _cell_spans(grid.mean(axis=0))
    if len(rows) != len(cols):
        raise ValueError("board should always be a square")
    blocked, tiles, goals = [], [], []
    for r, (r0, r1) in enumerate(rows):
        for c, (c0, c1) in enumerate(cols):
            cell = board_rgb[r0:r1, c0:c1]
            m = int(0.1 * len(cell))
            cell = cell[m:-m, m:-m]
            n = len(cell)
            centre = cell[int(0.45 * n):int(0.55 * n), int(0.45 * n):int(0.55 * n)]
            corner = cell[0:int(0.2 * n), 0:int(0.2 * n)]
            if _all_empty(cell):
                continue
            if _all_empty(centre):
                tiles.append(((r, c), corner.mean(axis=(0, 1))))
            elif _all_empty(corner):
                goals.append(((r, c), centre.mean(axis=(0, 1))))
            else:
                blocked.append((r, c))
    if len(goals) != len(tiles):
        raise ValueError("each tile should have a goal")
    if multiple_colors:                           # tile i belongs to target i: match by colour
        limit = float(np.linalg.norm(np.full(3, TOLERANCE)))
        ordered = []
        for _, gcol in goals:
            match = [pos for pos, tcol in tiles if np.linalg.norm(gcol - tcol) < limit]
            if len(match) != 1:
                raise ValueError("exactly one tile should want to come to this goal")
            ordered.append(match[0])
        tile_cells = ordered
    else:
        tile_cells = [pos for pos, _ in tiles]
    return Puzzle(size=len(rows), blocked_locations=blocked, initial_locations=tile_cells,
                  target_locations=[pos for pos, _ in goals], multiple_colors=bool(multiple_colors))


def load_level_image(path: str, multiple_colors: bool | None = None) -> Puzzle:
    """Screenshot file -> Puzzle.  Multi-colour iff '_multi_' is in the file name
    (play.py:209) unless given."""
    cv2 = _cv2()
    bgr = cv2.imread(path)
    if bgr is None:
        raise FileNotFoundError(path)
    rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
    if multiple_colors is None:
        multiple_colors = "_multi_" in os.path.basename(path)
    return parse_level_image(rgb[BOARD_ROWS, BOARD_COLS], multiple_colors)


def load_level(level: str, data_dir: str = "data") -> Puzzle:
    """Level by name ('puzzle_multi_001', with or without .jpg), as `play.py --level` does."""
    name = level if level.endswith(".jpg") else level + ".jpg"
    return load_level_image(os.path.join(data_dir, name))


def export_levels(paths: Iterable[str]) -> str:
    """Screenshots -> one text file in the `-input_file` format (puzzle.parse_puzzle_file_text)."""
    blocks = []
    for path in paths:
        p = load_level_image(path)
        blocks.append(f"# {os.path.splitext(os.path.basename(path))[0]}\n"
                      f"multi_color: {'true' if p.multiple_colors else 'false'}\n{puzzle_to_text(p)}")
    return "\n---\n".join(blocks) + "\n"
