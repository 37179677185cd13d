"""Puzzle specification, the text grammar and the `-input_file` loader.

Reference behaviour kept (paths relative to the reference checkout):
  * text grammar: TilerSliderEnvFactory.create_from_string, explainrl/environment/
    environment.py:236-288 -- 'X' blocked, lowercase = tile (index = letter - 'a'), uppercase
    other than 'X' = target (index = letter - 'A'), anything else empty; size = number of
    non-blank lines; gaps in the letter sequence are compacted in order.
  * puzzle container: ImageLoader.ImageProcessed, explainrl/environment/dataloader.py:21-27.
The reference has no input-file loader at all (main.py:1-6 is a stub); the file format below
is defined by this repo (parity unpinned) on top of that grammar.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

from ._lib import MAX_SIZE, MAX_TILES

Cell = Tuple[int, int]


@dataclass
class Puzzle:
    """One board: the same five fields as the reference's ImageProcessed
    (dataloader.py:21-27), plus optional metadata read from an input file."""
    size: int
    blocked_locations: List[Cell] = field(default_factory=list)
    initial_locations: List[Cell] = field(default_factory=list)
    target_locations: List[Cell] = field(default_factory=list)
    multiple_colors: bool = False
    moves: str = ""          # optional scripted action string from the input file
    max_steps: int | None = None

    @property
    def multi_color(self) -> bool:
        return self.multiple_colors

    def validate(self) -> "Puzzle":
        """Reject puzzles outside the bit-exact domain: the reference itself is
        erratic / platform dependent on duplicate tiles and on tiles standing on walls."""
        S = self.size
        if not isinstance(S, int) or not 1 <= S <= MAX_SIZE:
            raise ValueError(f"board size {S!r} outside 1..{MAX_SIZE}")
        if len(self.initial_locations) > MAX_TILES:
            raise ValueError(f"{len(self.initial_locations)} tiles exceed the supported maximum of {MAX_TILES}")
        for name, cells in (("blocked", self.blocked_locations), ("tile", self.initial_locations),
                            ("target", self.target_locations)):
            for r, c in cells:
                if not (0 <= int(r) < S and 0 <= int(c) < S):
                    raise ValueError(f"{name} cell {(r, c)} outside the {S}x{S} board")
        tiles = [(int(r), int(c)) for r, c in self.initial_locations]
        if len(set(tiles)) != len(tiles):
            raise ValueError("two tiles share a cell (outside the reference's well-defined domain)")
        blocked = {(int(r), int(c)) for r, c in self.blocked_locations}
        if blocked & set(tiles):
            raise ValueError("a tile stands on a blocked cell (outside the reference's well-defined domain)")
        return self


def parse_board_text(board_str: str, multi_color: bool = False) -> Puzzle:
    """Text grid -> Puzzle with exactly the reference grammar (environment.py:254-281).

    Per-line surrounding whitespace is stripped, blank lines are dropped, the board size is
    the number of remaining lines (columns are not validated by the reference; here a
    character beyond column size-1 fails validation)."""
    rows = [ln.strip() for ln in board_str.strip().split("\n") if ln.strip()]
    blocked: list[Cell] = []
    tiles: dict[int, Cell] = {}
    targets: dict[int, Cell] = {}
    for i, row in enumerate(rows):
        for j, ch in enumerate(row):
            if ch == "X":
                blocked.append((i, j))
            elif ch.islower():
                tiles[ord(ch) - ord("a")] = (i, j)
            elif ch.isupper():
                targets[ord(ch) - ord("A")] = (i, j)
    return Puzzle(size=len(rows), blocked_locations=blocked,
                  initial_locations=[tiles[k] for k in sorted(tiles)],
                  target_locations=[targets[k] for k in sorted(targets)],
                  multiple_colors=bool(multi_color))


_TRUE = {"1", "true", "yes", "on"}


def parse_puzzle_file_text(text: str) -> list[Puzzle]:
    """Input-file format (repo-defined): one or more puzzles separated by a line of dashes
    ('---').  Lines starting with '#' are comments.  Optional `key: value` header lines
    (before or after the grid) set `multi_color` (true/false), `moves` (e.g. RDLULD) and
    `max_steps`; every other non-blank line is a grid row in the reference grammar."""
    puzzles = []
    for block in _split_blocks(text):
        opts = {"multi_color": False, "moves": "", "max_steps": None}
        grid = []
        for ln in block:
            s = ln.strip()
            if not s or s.startswith("#"):
                continue
            if ":" in s:
                k, v = s.split(":", 1)
                k, v = k.strip().lower().replace("-", "_"), v.strip()
                if k in ("multi_color", "multicolor", "multiple_colors"):
                    opts["multi_color"] = v.lower() in _TRUE
                    continue
                if k == "moves":
                    opts["moves"] = "".join(ch for ch in v.upper() if ch in "UDLR")
                    continue
                if k == "max_steps":
                    opts["max_steps"] = int(v)
                    continue
            grid.append(s)
        if not grid:
            continue
        p = parse_board_text("\n".join(grid), opts["multi_color"])
        p.moves, p.max_steps = opts["moves"], opts["max_steps"]
        puzzles.append(p.validate())
    if not puzzles:
        raise ValueError("input file holds no puzzle grid")
    return puzzles


def _split_blocks(text: str):
    block: list[str] = []
    for ln in text.splitlines():
        if ln.strip() and set(ln.strip()) <= {"-"} and len(ln.strip()) >= 3:
            yield block
            block = []
        else:
            block.append(ln)
    yield block


def load_puzzle_file(path: str) -> list[Puzzle]:
    """`main.py -input_file PATH` loader: returns the validated puzzles of the file."""
    with open(path, "r", encoding="utf-8") as f:
        return parse_puzzle_file_text(f.read())


def puzzle_to_text(p: Puzzle) -> str:
    """Inverse of parse_board_text for boards where no tile shares a cell with a target."""
    grid = [["." for _ in range(p.size)] for _ in range(p.size)]
    for r, c in p.blocked_locations:
        grid[r][c] = "X"
    for k, (r, c) in enumerate(p.target_locations):
        ch = chr(ord("A") + k)
        grid[r][c] = ch if ch != "X" else "Y"
    for k, (r, c) in enumerate(p.initial_locations):
        grid[r][c] = chr(ord("a") + k)
    return "\n".join("".join(row) for row in grid)


def as_puzzle(obj) -> Puzzle:
    """Accept a Puzzle or any object with the ImageProcessed field names."""
    if isinstance(obj, Puzzle):
        return obj
    return Puzzle(size=obj.size, blocked_locations=list(obj.blocked_locations),
                  initial_locations=list(obj.initial_locations),
                  target_locations=list(obj.target_locations),
                  multiple_colors=bool(getattr(obj, "multiple_colors", False)))


def uniform_shape(puzzles: Sequence[Puzzle]) -> tuple[int, int, int, bool]:
    """(size, n_tiles, n_targets, multi_color) shared by a batch, or ValueError."""
    p0 = puzzles[0]
    key = (p0.size, len(p0.initial_locations), len(p0.target_locations), bool(p0.multiple_colors))
    for p in puzzles:
        k = (p.size, len(p.initial_locations), len(p.target_locations), bool(p.multiple_colors))
        if k != key:
            raise ValueError(f"a batch needs one board size, tile count, target count and colour mode; got {key} and {k}")
    return key
