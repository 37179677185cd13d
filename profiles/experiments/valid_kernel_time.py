"""Kernel-only timing of ts_valid_moves (prebuilt argument block, CUDA events): python profiles/experiments/valid_kernel_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C, torch, tiler_slider_b200 as ts
from tiler_slider_b200._lib import ValidArgs, GoalArgs
lib = ts.lib()
for S, T, W, N in ((6, 4, 8, 4_194_304), (12, 8, 36, 1_048_576), (8, 8, 12, 4_194_304)):
    env = ts.BatchedTilerSliderEnv.synthetic(N, S, T, W, True, seed=1)
    mask = torch.zeros(env.capacity, dtype=torch.uint8, device="cuda")
    a = ValidArgs(size=S, n_tiles=T, first_env=0, n_envs=N, capacity=env.capacity, d_walls=env._walls.data_ptr(), d_pos=env._pos.data_ptr(), d_mask=mask.data_ptr())
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5): lib.ts_valid_moves(C.byref(a), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): lib.ts_valid_moves(C.byref(a), st)
    e1.record(); e1.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    pw = lib.ts_pos_bytes(T); wb = lib.ts_walls_bytes(S) // (2 if S > 8 else 1)
    print(f"ts_valid_moves {S}x{S}/{T}, {N} envs: {us:.1f} us, {N/us/1e3:.1f} G envs/s, {(pw+wb+1)*N/us/1e3:.0f} GB/s of {pw+wb+1} B/env")
