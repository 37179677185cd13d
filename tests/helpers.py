"""Shared test helpers (test infrastructure)."""
import numpy as np


def parse_text(text):
    """Independent restatement of the reference text grammar (environment.py:254-281), used
    only to cross-check the product loader and to feed the oracle."""
    lines = [ln for ln in text.strip().split("\n") if ln.strip()]
    blocked, tiles, targets = [], {}, {}
    for i, ln in enumerate(lines):
        for j, ch in enumerate(ln.strip()):
            if ch == "X":
                blocked.append((i, j))
            elif ch.islower():
                tiles[ord(ch) - ord("a")] = (i, j)
            elif ch.isupper():
                targets[ord(ch) - ord("A")] = (i, j)
    return len(lines), blocked, [tiles[k] for k in sorted(tiles)], [targets[k] for k in sorted(targets)]


def random_puzzles(rng, n, S, T, W):
    """numpy version of the create_simple_env recipe (environment.py:221-226): a random
    permutation of the cells; first W blocked, next T tiles, next T targets."""
    perm = np.argsort(rng.random((n, S * S)), axis=1)
    blocked = np.zeros((n, S * S), np.uint8)
    np.put_along_axis(blocked, perm[:, :W], 1, axis=1)
    tcell = perm[:, W:W + T]
    gcell = perm[:, W + T:W + 2 * T]
    tiles = np.stack([tcell // S, tcell % S], axis=-1).astype(np.uint8)
    targets = np.stack([gcell // S, gcell % S], axis=-1).astype(np.uint8)
    return blocked, tiles, targets


def reachable_targets(orc, rng, S, blocked, tiles, multi, n_moves):
    """targets = where the tiles stand after n_moves random moves (oracle), so that every puzzle of the
    batch has a solution"""
    out = np.zeros_like(tiles)
    for e in range(tiles.shape[0]):
        bl = [(c // S, c % S) for c in np.flatnonzero(blocked[e])]
        st = orc.OracleState(S, bl, tiles[e].tolist(), tiles[e].tolist(), multi)
        for m in rng.integers(0, 4, n_moves):
            st.move(int(m))
        out[e] = np.array(st.current_locations, dtype=tiles.dtype)
    return out
