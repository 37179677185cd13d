"""Why `e2e` does not scale linearly with the GPU count: concurrent pinned-memory PCIe rates of
all ranks, with and without binding each rank to the CPUs local to its GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 profiles/experiments/pcie_multi.py [bind]
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
bind = len(sys.argv) > 1 and sys.argv[1] == "bind"
info = {"rank": rank, "bind": bind}


def local_cpus(index: int):
    import pynvml
    pynvml.nvmlInit()
    bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
    if isinstance(bus, bytes):
        bus = bus.decode()
    dom, rest = bus.split(":", 1)
    path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}"
    info["bus"] = bus
    try:
        info["numa_node"] = int(open(path + "/numa_node").read())
        cpus = set()
        for part in open(path + "/local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus
    except OSError as e:
        info["sysfs_error"] = str(e)
        return set()


cpus = local_cpus(local)
info["n_local_cpus"] = len(cpus)
info["affinity_before"] = len(os.sched_getaffinity(0))
if bind and cpus:
    os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import tiler_slider_b200 as ts  # noqa: E402


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


n = 64 << 20
h_a, h_b = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_a, d_b = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
info["h2d_GBs"] = n / timed(lambda: d_a.copy_(h_a, non_blocking=True)) / 1e9
info["d2h_GBs"] = n / timed(lambda: h_b.copy_(d_b, non_blocking=True)) / 1e9


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


info["bidir_each_GBs"] = n / timed(both) / 1e9

N = 1 << 24
env = ts.BatchedTilerSliderEnv.synthetic(N, 6, 4, 4, True, seed=1, env_index_base=rank * N, max_steps=64, auto_reset=True, device=dev)
h_act = torch.randint(0, 4, (N,), dtype=torch.uint8).pin_memory()
h_rew = torch.empty(N, dtype=torch.float32).pin_memory()
h_done = torch.empty(N, dtype=torch.uint8).pin_memory()
h_flags = torch.empty(N, dtype=torch.uint8).pin_memory()
info["e2e_full"] = N / timed(lambda: env.step_host(h_act, h_rew, h_done), reps=8)
info["e2e_compact"] = N / timed(lambda: env.step_host(h_act, h_flags=h_flags), reps=8)
out = [None] * world
dist.all_gather_object(out, info)
if rank == 0:
    print(json.dumps({"world": world, "bind": bind, "sum_e2e_full": sum(o["e2e_full"] for o in out),
                      "sum_e2e_compact": sum(o["e2e_compact"] for o in out), "ranks": out}))
dist.destroy_process_group()
