"""Run by tests/test_gpu_parity.py::test_pipelined_kernel_matches_direct_kernel_and_oracle in a
subprocess with TS_LIB_PATH = the -DTS_WITH_PIPE variant of the library and TS_STEP_PIPE=1.

Batches of >= 2^18 envs then run the persistent bulk-async kernel (step_kernel_pipe); the same batch
stepped through 65,536-env sub-ranges runs the direct kernel.  Both must agree bit for bit on every
array, and the first 2,048 envs must match the oracle.  N is a whole number of 128-env tiles plus
a ragged remainder handled by the direct / per-env kernels (the pipelined kernel only takes whole
tiles), so both hand-overs are exercised."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tiler_slider_b200 as ts  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = [(6, 4, 8, True, True, 100), (6, 4, 8, False, False, 11), (5, 1, 5, False, True, 7), (4, 2, 2, True, True, 100),
         (8, 8, 12, True, True, 100), (6, 3, 4, True, False, 100)]


def main():
    assert os.environ.get("TS_LIB_PATH") and os.environ.get("TS_STEP_PIPE") == "1"
    lib = ts.lib()
    for S, T, W, multi, auto_reset, max_steps in CASES:
        K = 24
        n_chk = 2048
        for N in ((1 << 18) + 128 * 5, (1 << 18) + 128 * 5 + 37):      # whole tiles (pipelined) / ragged (direct + per-env)
            kw = dict(seed=5, max_steps=max_steps, auto_reset=auto_reset, track_terminal=True)
            a = ts.BatchedTilerSliderEnv.synthetic(N, S, T, W, multi, **kw)
            b = ts.BatchedTilerSliderEnv.synthetic(N, S, T, W, multi, **kw)
            blocked = a.blocked_cells()[:n_chk].cpu().numpy().astype(np.uint8)
            tiles = a.positions()[:n_chk].cpu().numpy()
            if a.goal_mode == ts.GOAL_ORDERED:
                targets = a.target_positions()[:n_chk].cpu().numpy()
            else:
                cells = a.target_positions()[:n_chk].cpu().numpy()
                tc = np.stack([np.flatnonzero(r) for r in cells])
                targets = np.stack([tc // S, tc % S], -1).astype(np.uint8)
            g = torch.Generator(device="cuda").manual_seed(3)
            actions = torch.randint(0, 4, (K, a.capacity), dtype=torch.uint8, device="cuda", generator=g)
            want = orc.rollout(S, multi, blocked, tiles, targets, actions[:, :n_chk].cpu().numpy(), max_steps=max_steps,
                               auto_reset=auto_reset)
            for k in range(K):
                _, r, d = a.step(actions[k])
                args = b._step_args(actions[k].data_ptr())
                for lo in range(0, N, 65536):
                    args.first_env, args.n_envs = lo, min(65536, N - lo)
                    assert lib.ts_step(C.byref(args), torch.cuda.current_stream().cuda_stream) == 0
                for name in ("_pos", "_count", "_reward", "_done", "_flags"):
                    assert torch.equal(getattr(a, name)[:N], getattr(b, name)[:N]), (name, k)
                dd = d[:n_chk]
                post = a.positions()[:n_chk]
                if auto_reset:
                    post = torch.where(dd[:, None, None], a.positions(a.terminal_pos)[:n_chk], post)
                assert np.array_equal(post.cpu().numpy(), want["pos"][k])
                assert np.array_equal(a.flags[:n_chk].cpu().numpy(), want["flags"][k])
                assert np.array_equal(r[:n_chk].cpu().numpy(), want["reward"][k])
    print("pipe variant ok")


if __name__ == "__main__":
    main()
