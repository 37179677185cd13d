"""Two-GPU BFS: the NCCL all-to-all exchange and the peer-memory exchange (ts_bfs_expand_exchange)
must find exactly what a single GPU finds, and the first puzzles must agree with the CPU oracle's
BFS (bfs_bench.py --check).  Needs two GPUs: skipped on a one-GPU box."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")]


def _bench(world: int, exchange: str, puzzles: int = 4096, mode: str = "hash") -> dict:
    cmd = [sys.executable, "bfs_bench.py", "--puzzles", str(puzzles), "--check", "64", "--exchange", exchange, "--mode", mode]
    if world > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", "29541"] + cmd[1:]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_two_gpu_exchanges_match_single_gpu():
    one = _bench(1, "nccl")
    for exchange in ("nccl", "p2p"):
        two = _bench(2, exchange)
        assert exchange in two["config"]
        assert two["oracle_check"]["ok"]
        for key in ("unique_states", "generated_successors", "depth", "puzzles_solved", "max_solve_depth"):
            assert two[key] == one[key], (exchange, key)
    # the on-chip search, puzzles sharded over two ranks with no exchange: same totals again
    local = _bench(2, "nccl", mode="local")
    assert local["oracle_check"]["ok"] and local["mode"] == "local"
    for key in ("unique_states", "generated_successors", "depth", "puzzles_solved", "max_solve_depth"):
        assert local[key] == one[key], ("local", key)


def test_two_gpu_shortest_paths():
    """with_paths over two ranks (parents exchanged with the keys, the chains walked one collective
    per move): same strings lengths as a single-rank search, every string replayed through the CPU
    oracle wins on its last move (tests/bfs_paths_multi_check.py)."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join("tests", "bfs_paths_multi_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    rep = json.loads(out.stdout.strip().splitlines()[-1])
    assert rep["ok"] and len(rep["cases"]) == 6 and all(c["solved"] > 0 for c in rep["cases"])
