"""Does pinning part of the read-only puzzle data (walls) in L2 pay?  The step re-reads the
walls and targets of every env each step; they are larger than L2, so plain LRU keeps nothing.
A persisting access-policy window on the walls buffer keeps a fraction of its lines resident
across steps.  Stream attribute on the launching stream; no library change needed to try it.

    python profiles/experiments/l2_persist.py c4|c3
"""
import json
import os
import sys

import torch
from cuda.bindings import runtime as rt

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import tiler_slider_b200 as ts  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
S, T, W, N = {"c4": (12, 8, 36, 1 << 22), "c3": (6, 4, 8, 1 << 24), "c2": (5, 1, 5, 1 << 24)}[cfg]
multi = cfg != "c2"
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
env = ts.BatchedTilerSliderEnv.synthetic(N, S, T, W, multi, seed=7, max_steps=100, auto_reset=True, device=dev)
actions = torch.randint(0, 4, (8, env.capacity), dtype=torch.uint8, device=dev)
err, prop = rt.cudaGetDeviceProperties(0)
out = {"config": cfg, "l2_bytes": prop.l2CacheSize, "persisting_max": prop.persistingL2CacheMaxSize,
       "window_max": prop.accessPolicyMaxWindowSize, "walls_bytes": env._walls.numel(), "runs": []}
side = torch.cuda.Stream()          # attributes cannot be set on the legacy default stream
torch.cuda.set_stream(side)
stream = side.cuda_stream


def time_steps(n=200):
    for k in range(16):
        env.step(actions[k % 8])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(n):
        env.step(actions[k % 8])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n


def set_window(ptr, nbytes, ratio):
    attr = rt.cudaStreamAttrValue()
    attr.accessPolicyWindow.base_ptr = ptr
    attr.accessPolicyWindow.num_bytes = nbytes
    attr.accessPolicyWindow.hitRatio = ratio
    attr.accessPolicyWindow.hitProp = rt.cudaAccessProperty.cudaAccessPropertyPersisting
    attr.accessPolicyWindow.missProp = rt.cudaAccessProperty.cudaAccessPropertyStreaming
    (e,) = rt.cudaStreamSetAttribute(stream, rt.cudaStreamAttrID.cudaLaunchAttributeAccessPolicyWindow, attr)
    return int(e)


out["runs"].append({"persist_mb": 0, "us_per_step": time_steps()})
walls_ptr, walls_bytes = env._walls.data_ptr(), env._walls.numel()
for persist_mb in (32, 48, 64, 80):
    set_aside = min(persist_mb << 20, prop.persistingL2CacheMaxSize)
    (e1,) = rt.cudaDeviceSetLimit(rt.cudaLimit.cudaLimitPersistingL2CacheSize, set_aside)
    win = min(walls_bytes, prop.accessPolicyMaxWindowSize)
    for scale in (1.0, 0.8):
        ratio = min(1.0, scale * set_aside / win)
        e2 = set_window(walls_ptr, win, ratio)
        out["runs"].append({"persist_mb": persist_mb, "set_aside": set_aside, "window": win, "hit_ratio": ratio,
                            "rc": [int(e1), e2], "us_per_step": time_steps()})
set_window(0, 0, 0.0)
rt.cudaCtxResetPersistingL2Cache()
out["runs"].append({"persist_mb": "off again", "us_per_step": time_steps()})
print(json.dumps(out, indent=1))
