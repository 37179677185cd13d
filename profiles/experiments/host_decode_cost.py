"""What a packed end-to-end protocol (2-bit actions up, 4-bit status down) would cost on the HOST:
unpacking the status nibbles of one 16,777,216-env step into reward f32 + done u8 with numpy on one
core.  Run anywhere (no GPU):  python profiles/experiments/host_decode_cost.py"""
import time

import numpy as np

n = 16_777_216
packed = np.random.default_rng(0).integers(0, 256, n // 2, dtype=np.uint8)
lut = np.zeros(16, np.float32)
lut[[0, 8]], lut[[2, 3, 10, 11]], lut[[4, 12]] = -0.01, 1.0, -0.05      # step / won / invalid (status bits DONE=1 WON=2 INVALID=4 TIMEOUT=8)
best = 1e9
for _ in range(5):
    t0 = time.perf_counter()
    st = np.empty(n, np.uint8)
    st[0::2], st[1::2] = packed & 15, packed >> 4
    reward, done = lut[st], (st & 1)
    best = min(best, time.perf_counter() - t0)
print(f"host decode of one {n}-env step: {best * 1e3:.1f} ms on one core = {n / best:.2e} env-steps/s "
      f"(the 5 B/env PCIe protocol runs at ~1.0e10 env-steps/s without any host work)")
