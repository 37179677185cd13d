"""Level ingest (tiler_slider_b200/levels.py) against the reference's own image parser.

tests/golden/levels_400.txt holds the 400 levels as decoded by the UNMODIFIED reference
(ImageLoader.parse_puzzle_image, dataloader.py:44-133; tests/golden/make_levels.py).  The
screenshots themselves (99 MB of JPEGs) live only in the reference checkout, so this test runs
where that is mounted (the build container) and is skipped elsewhere."""
import os

import pytest

DATA = os.path.join(os.environ.get("TS_REFERENCE", "/root/reference"), "data")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "levels_400.txt")


@pytest.mark.skipif(not os.path.isdir(DATA), reason="reference screenshots not available here")
def test_every_level_matches_the_reference_parser():
    from tiler_slider_b200.levels import export_levels, load_level, load_level_image
    from tiler_slider_b200.puzzle import load_puzzle_file, parse_puzzle_file_text
    want = load_puzzle_file(GOLD)
    names = sorted(f for f in os.listdir(DATA) if f.endswith(".jpg"))
    assert len(names) == len(want) == 400
    for name, w in zip(names, want):
        got = load_level_image(os.path.join(DATA, name))
        assert (got.size, got.multiple_colors) == (w.size, w.multiple_colors), name
        assert got.blocked_locations == w.blocked_locations, name
        assert got.initial_locations == w.initial_locations, name
        assert got.target_locations == w.target_locations, name
    p = load_level("puzzle_multi_001", DATA)
    assert p.initial_locations == [(0, 3), (3, 2)] and p.target_locations == [(0, 0), (3, 0)]
    text = export_levels(os.path.join(DATA, n) for n in names[:5])
    again = parse_puzzle_file_text(text)
    assert [a.initial_locations for a in again] == [w.initial_locations for w in want[:5]]
