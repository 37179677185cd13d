"""CPU-only tests: text grammar / input-file loader, layout arithmetic, C-ABI symbol export,
sharding helper, main.py argument surface.  No CUDA compute is called here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from tests.helpers import parse_text

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """include/tiler_slider.h is the contract: every function it declares must be exported by
    the built library and bound by the ctypes layer."""
    from tiler_slider_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tiler_slider.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char \*)\s*\*?(ts_[a-z_0-9]+)\s*\(", header, flags=re.M))
    assert len(declared) >= 16
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    lib = _lib.lib()
    assert lib.ts_version() == 201


def test_layout_arithmetic():
    from tiler_slider_b200 import _lib
    lib = _lib.lib()
    assert [lib.ts_pos_bytes(t) for t in range(1, 9)] == [1, 2, 4, 4, 8, 8, 8, 8]
    assert [lib.ts_board_bytes(s) for s in (1, 2, 3, 4, 5, 6, 7, 8, 12, 16)] == [1, 1, 2, 3, 4, 6, 7, 8, 18, 32]
    assert [lib.ts_board_stride(s) for s in (1, 5, 6, 7, 8)] == [2, 6, 7, 7, 8]
    assert [lib.ts_pos_stride(s) for s in (1, 5, 6, 7, 8, 12)] == [2, 6, 7, 16, 16, 16]
    for nb in range(1, 33):
        widths = [lib.ts_plane_width(nb, k) for k in range(lib.ts_plane_count(nb))]
        assert sum(widths) == nb and widths == sorted(widths, reverse=True)
        assert all(w in (1, 2, 4, 8, 16) for w in widths)
        assert [lib.ts_plane_offset(nb, k) for k in range(len(widths))] == [sum(widths[:k]) for k in range(len(widths))]
    assert lib.ts_supported(6, 4) == 1 and lib.ts_supported(17, 1) == 0 and lib.ts_supported(6, 9) == 1 and lib.ts_supported(6, 33) == 0
    assert [lib.ts_pos_bytes(t) for t in (0, 9, 16, 17, 32)] == [1, 16, 16, 32, 32]


def test_argument_validation_without_gpu():
    """Argument errors are detected before any CUDA call."""
    from tiler_slider_b200 import _lib
    lib = _lib.lib()
    a = _lib.StepArgs(size=6, n_tiles=4, n_envs=8, capacity=100)
    assert lib.ts_step(ctypes.byref(a), None) == -3
    a = _lib.StepArgs(size=0, n_tiles=4, n_envs=8, capacity=128)
    assert lib.ts_step(ctypes.byref(a), None) == -1
    a = _lib.StepArgs(size=6, n_tiles=4, n_envs=8, capacity=128, count_bytes=1, max_steps=100)
    assert lib.ts_step(ctypes.byref(a), None) == -4
    assert b"null" in lib.ts_last_error_string()


def test_grammar_matches_golden(golden_misc):
    from tiler_slider_b200 import parse_board_text
    for g in golden_misc["grammar"]:
        p = parse_board_text(g["text"])
        assert p.size == g["size"]
        assert [list(x) for x in p.blocked_locations] == g["blocked"]
        assert [list(x) for x in p.initial_locations] == g["tiles"]
        assert [list(x) for x in p.target_locations] == g["targets"]
        assert (p.size, p.blocked_locations, p.initial_locations, p.target_locations) == parse_text(g["text"])


def test_input_file_format(tmp_path):
    from tiler_slider_b200 import load_puzzle_file, parse_puzzle_file_text, puzzle_to_text
    text = """# two puzzles
multi_color: true
moves: R D L U L D
A..a
X...
...X
B.b.
---
max_steps: 7
a....
.X...
.....
...X.
....A
"""
    f = tmp_path / "input.txt"
    f.write_text(text)
    ps = load_puzzle_file(str(f))
    assert len(ps) == 2
    assert ps[0].size == 4 and ps[0].multiple_colors and ps[0].moves == "RDLULD"
    assert ps[0].initial_locations == [(0, 3), (3, 2)] and ps[0].target_locations == [(0, 0), (3, 0)]
    assert ps[1].size == 5 and not ps[1].multiple_colors and ps[1].max_steps == 7
    assert parse_puzzle_file_text(puzzle_to_text(ps[0]))[0].blocked_locations == ps[0].blocked_locations
    with pytest.raises(ValueError):
        parse_puzzle_file_text("# nothing\n")
    with pytest.raises(ValueError):
        parse_puzzle_file_text("a.\n..b\n")            # tile in column 2 of a 2x2 board


def test_validation_rejects_ill_formed():
    from tiler_slider_b200 import Puzzle
    Puzzle(3, [(0, 0)], [(1, 1)], [(1, 1)]).validate()
    for bad in [Puzzle(3, [], [(0, 0), (0, 0)], [(1, 1), (2, 2)]), Puzzle(3, [(1, 1)], [(1, 1)], [(0, 0)]),
                Puzzle(3, [], [(3, 0)], [(0, 0)]), Puzzle(17, [], [(0, 0)], [(1, 1)]),
                Puzzle(6, [], [(i // 6, i % 6) for i in range(33)], [])]:      # 33 tiles: one more than the kernels cover
        with pytest.raises(ValueError):
            bad.validate()


def test_shard_range():
    from tiler_slider_b200 import shard_range
    for n, w in [(16, 4), (17, 4), (3, 8), (16_777_216, 8)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_missing_library_fails_loudly(monkeypatch):
    from tiler_slider_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libtiler_slider.so")
    with pytest.raises(_lib.TilerSliderError, match="no CPU fallback"):
        _lib.lib()


def test_no_gpu_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import tiler_slider_b200 as ts
    with pytest.raises(ts.TilerSliderError):
        ts.BatchedTilerSliderEnv(6, 4, 16, True)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing in the package may reference it."""
    pkg = os.path.join(ROOT, "tiler_slider_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in src.lower(), fn
    assert "oracle" not in open(os.path.join(ROOT, "main.py")).read().lower()


def test_ctypes_mirrors_match_the_c_header(tmp_path):
    """Every POD argument struct of include/tiler_slider.h, compiled as plain C by gcc, has the
    size and the field offsets of its ctypes mirror in _lib.py (the header is the ABI; the
    mirrors are what the Python host passes)."""
    import ctypes as C
    import subprocess
    from tiler_slider_b200 import _lib
    pairs = {"ts_encode_args": _lib.EncodeArgs, "ts_synth_args": _lib.SynthArgs, "ts_step_args": _lib.StepArgs,
             "ts_observe_args": _lib.ObserveArgs, "ts_valid_args": _lib.ValidArgs, "ts_goal_args": _lib.GoalArgs,
             "ts_bfs_args": _lib.BfsArgs, "ts_bfs_local_args": _lib.BfsLocalArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "tiler_slider.h"', 'int main(void) {']
    for cname, mirror in pairs.items():
        lines.append(f'printf("SIZEOF {cname} - %zu\\n", sizeof({cname}));')
        for field, _ in mirror._fields_:
            lines.append(f'printf("OFFSET {cname} {field} %zu\\n", offsetof({cname}, {field}));')
    lines += ['return 0;', '}']
    src, exe = tmp_path / "abi.c", tmp_path / "abi"
    src.write_text("\n".join(lines))
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines()
    assert len(out) > 100
    for line in out:
        kind, cname, field, value = line.split()
        if kind == "SIZEOF":
            assert int(value) == C.sizeof(pairs[cname]), f"sizeof({cname})"
        else:
            assert int(value) == getattr(pairs[cname], field).offset, f"{cname}.{field}"
