"""Does cudaLimitMaxL2FetchGranularity (32 / 64 / 128 bytes) change the DRAM over-fetch of the
wide-board step kernel?  Measured on B200: no (99.0 us per 4.2M-env step at every setting)."""
import sys, json, subprocess, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
rt = torch.cuda.cudart()
torch.cuda.init()
import tiler_slider_b200 as ts
def bench(tag):
    env = ts.BatchedTilerSliderEnv.synthetic(4_194_304, 12, 8, 36, True, seed=1003, max_steps=100, auto_reset=True)
    acts = torch.randint(0, 4, (8, env.capacity), dtype=torch.uint8, device="cuda")
    for k in range(10): env.step(acts[k % 8])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for k in range(200): env.step(acts[k % 8])
    e.record(); torch.cuda.synchronize()
    print(tag, s.elapsed_time(e) / 200 * 1e3, "us/step")
val = ctypes.c_size_t()
crt = ctypes.CDLL(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart.so.12")) if os.path.exists(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart.so.12")) else ctypes.CDLL("libcudart.so")
crt.cudaDeviceGetLimit(ctypes.byref(val), 5); print("default granularity", val.value)
bench("default")
for g in (32, 64, 128):
    rc = crt.cudaDeviceSetLimit(5, ctypes.c_size_t(g)); crt.cudaDeviceGetLimit(ctypes.byref(val), 5)
    print("set", g, "rc", rc, "now", val.value)
    bench(f"gran{g}")
