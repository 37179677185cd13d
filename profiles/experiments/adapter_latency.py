"""Latency of the single-env drop-in surface (batch of one, host_io): python profiles/experiments/adapter_latency.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import time, torch
import tiler_slider_b200 as ts
env = ts.TilerSliderEnvFactory.create_simple_env(size=6, num_tiles=4, num_obstacles=8, seed=1, max_steps=10**6)
env.reset(); env.step(ts.Move.UP)
for name, fn, n in (("reset", env.reset, 50), ("get_valid_moves", env.get_valid_moves, 100), ("get_info", env.get_info, 100), ("state.is_won", lambda: env.state.is_won(), 100),
                    ("state.copy+move", lambda: env.state.copy().move(ts.Move.LEFT), 50), ("state.get_state_array", lambda: env.state.get_state_array(), 100)):
    fn()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    print(f"{name:24s} {(time.perf_counter()-t0)/n*1e6:8.1f} us")
