"""bench.py -- env-steps/s of the Tiler-Slider step path on B200, with its HBM roofline and the
CPU step loop beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the one its metric is quoted on): 6x6 boards, 4 coloured
tiles and targets (ordered goal), 8 random walls, 16,777,216 envs PER GPU (weak scaling: envs
are independent, sharded by index, no collective on the step path), uniform random actions,
max_steps=100, auto-reset on.  A "step" is one ts_step launch advancing every env by one
action.  `value` = total env-steps/s over all ranks with actions resident in HBM, timed with
CUDA events on the launching stream, max over ranks.  `e2e` = the same step driven from pinned
HOST buffers through ts_step_host (actions uploaded, reward+done downloaded every step).
`roofline` = algorithmic bytes (3T + ceil(S^2/8) + 8 = 25 B per env-step) divided
by the measured launch duration, against the measured copy bandwidth of MEASURED_PEAKS.json.
`cpu_baseline` = the reference's Python step loop restated in oracle/py_port.py, timed on this
box's host cores on a bounded sample (the reference itself is pure Python and is not on this
box).  `--impl reference` times that CPU loop on all host cores and prints the same line.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs: c3 is the one the metric is quoted on (the bench line); c2 / c4 can be
# timed with --config for the record (they are parity-test cases otherwise)
CONFIGS = {"c2": (5, 1, 5, False, 1_048_576, 1001, 2001), "c3": (6, 4, 8, True, 16_777_216, 1002, 2002),
           "c4": (12, 8, 36, True, 4_194_304, 1003, 2003)}
S, T, W_WALLS, MULTI = 6, 4, 8, True
ENVS_PER_GPU = 16_777_216
MAX_STEPS = 100
PUZZLE_SEED, ACTION_SEED = 1002, 2002
ALGO_BYTES = 3 * T + (S * S + 7) // 8 + 8      # 25 B per env-step
METRIC = "env-steps/sec at 1/2/4/8 B200 and % HBM roofline vs reference CPU step loop"
UNIT = "env-steps/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel at the default
# size, from the committed `ncu --set full` capture (profiles/r1_final_step_kernel_full_raw.csv)
NCU_TRAFFIC_BYTES = {"c3": 268.448512e6 + 126.243840e6}
N_ACTION_ROWS = 8                               # distinct pre-generated action vectors, cycled


def workload_config(n_gpus: int, envs_per_gpu: int) -> dict:
    return {"workload": f"{S}x{S} boards, {T} {'coloured ' if MULTI else ''}tile(s)+target(s) ({'ordered' if MULTI else 'set'} goal), "
                        f"{W_WALLS} walls, {envs_per_gpu} envs per GPU "
                        f"sharded by index over {n_gpus} GPU(s), uniform random actions, max_steps=100, auto-reset",
            "size": S, "tiles": T, "walls": W_WALLS, "multi_color": MULTI, "envs_per_gpu": envs_per_gpu,
            "envs_total": envs_per_gpu * n_gpus, "max_steps": MAX_STEPS, "auto_reset": True,
            "algorithmic_bytes_per_env_step": ALGO_BYTES,
            "l2_policy": ("inputs larger than L2 (>=400 MB of state+io per step vs 126 MB L2); no flush needed"
                          if envs_per_gpu * ALGO_BYTES > 2 * 126e6 else "working set fits the 126 MB L2 (not a bench line)"),
            "parallelism": f"env-index sharding x{n_gpus}, no collective on the step path"}


# ------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One process: its own envs, W warm-up + K timed passes of the reference-style loop."""
    puzzles, actions, warmup, steps = args
    from oracle import py_port
    envs = [py_port.PortEnv(s, b, i, t, m, MAX_STEPS) for (s, b, i, t, m) in puzzles]
    for e in envs:
        e.reset()

    def one_pass(k):
        row = actions[k % len(actions)]
        for j, e in enumerate(envs):
            _, done, _ = e.step(int(row[j]))
            if done:
                e.reset()
    for k in range(warmup):
        one_pass(k)
    t0 = time.perf_counter()
    for k in range(steps):
        one_pass(warmup + k)
    return time.perf_counter() - t0, len(envs) * steps


def synth_puzzles_host(n: int, seed: int):
    """Host copy of the synthetic recipe (random permutation prefix; environment.py:221-226)
    for the CPU legs, which must not need a GPU."""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        perm = rng.permutation(S * S)
        cells = [(int(c) // S, int(c) % S) for c in perm]
        out.append((S, cells[:W_WALLS], cells[W_WALLS:W_WALLS + T], cells[W_WALLS + T:W_WALLS + 2 * T], MULTI))
    return out


def cpu_loop_rate(n_procs: int, steps: int, warmup: int, budget_s: float):
    """env-steps/s of the Python step loop on n_procs processes, sized to ~budget_s."""
    import numpy as np
    est_rate = 3.0e4                                   # per core, SURVEY section 6
    envs_per_proc = int(max(1, min(4096, budget_s * est_rate / max(1, steps + warmup))))
    rng = np.random.default_rng(ACTION_SEED)
    jobs = []
    for p in range(n_procs):
        acts = rng.integers(0, 4, size=(min(steps + warmup, 256), envs_per_proc), dtype=np.uint8)
        jobs.append((synth_puzzles_host(envs_per_proc, PUZZLE_SEED + p), acts, warmup, steps))
    if n_procs == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(n_procs) as pool:
            res = pool.map(_cpu_worker, jobs)
    elapsed = max(r[0] for r in res)
    total = sum(r[1] for r in res)
    return total / elapsed, elapsed, envs_per_proc


def c_oracle_rate(n_envs: int = 4096, steps: int = 200):
    import numpy as np
    from oracle import oracle as orc
    rng = np.random.default_rng(PUZZLE_SEED)
    perm = np.argsort(rng.random((n_envs, S * S)), axis=1)
    blocked = np.zeros((n_envs, S * S), np.uint8)
    np.put_along_axis(blocked, perm[:, :W_WALLS], 1, axis=1)
    tc, gc = perm[:, W_WALLS:W_WALLS + T], perm[:, W_WALLS + T:W_WALLS + 2 * T]
    tiles = np.stack([tc // S, tc % S], -1).astype(np.uint8)
    targets = np.stack([gc // S, gc % S], -1).astype(np.uint8)
    actions = rng.integers(0, 4, size=(steps, n_envs), dtype=np.uint8)
    t0 = time.perf_counter()
    orc.rollout(S, MULTI, blocked, tiles, targets, actions, max_steps=MAX_STEPS, auto_reset=True)
    return n_envs * steps / (time.perf_counter() - t0)


def parity_check(ts, dev, n_envs: int = 8192, steps: int = 128) -> dict:
    """The parity run that accompanies the throughput number: the first n_envs
    puzzles of the bench workload, `steps` random actions, every position / flag / reward of the
    CUDA path compared with the C oracle (the checker, part of the cpu_baseline leg)."""
    import numpy as np
    import torch
    from oracle import oracle as orc
    env = ts.BatchedTilerSliderEnv.synthetic(n_envs, S, T, W_WALLS, MULTI, seed=PUZZLE_SEED, max_steps=MAX_STEPS,
                                             auto_reset=True, track_terminal=True, device=dev)
    blocked = env.blocked_cells().cpu().numpy().astype(np.uint8)
    tiles = env.positions().cpu().numpy()
    if env.goal_mode == ts.GOAL_ORDERED:
        targets = env.target_positions().cpu().numpy()
    else:
        cells = np.stack([np.flatnonzero(r) for r in env.target_positions().cpu().numpy()])
        targets = np.stack([cells // S, cells % S], -1).astype(np.uint8)
    gen = torch.Generator(device=dev).manual_seed(ACTION_SEED)
    actions = torch.randint(0, 4, (steps, env.capacity), dtype=torch.uint8, device=dev, generator=gen)
    want = orc.rollout(S, MULTI, blocked, tiles, targets, actions[:, :n_envs].cpu().numpy(), max_steps=MAX_STEPS, auto_reset=True)
    bad = 0
    for k in range(steps):
        _, r, d = env.step(actions[k])
        post = torch.where(d[:, None, None], env.positions(env.terminal_pos), env.positions())
        bad += int((post.cpu().numpy() != want["pos"][k]).any(axis=(1, 2)).sum())
        bad += int((env.flags.cpu().numpy() != want["flags"][k]).sum())
        bad += int((r.cpu().numpy() != want["reward"][k]).sum())
    # K3 / valid-move mask of the states the rollout ended in (first n_obs envs), against the oracle
    n_obs = min(n_envs, 1024)
    obs = env.observe()[:n_obs].cpu().numpy()
    valid = env.valid_moves()[:n_obs].cpu().numpy()
    blocked_rc = [np.argwhere(b.reshape(S, S)) for b in blocked[:n_obs]]
    obs_bad = 0
    for i in range(n_obs):
        st = orc.OracleState(S, blocked_rc[i], want["final_pos"][i], targets[i], MULTI)
        obs_bad += int((st.get_state_array() != obs[i]).any())
        obs_bad += int(sum(1 << d for d in st.valid_moves()) != int(valid[i]))
    return {"env_steps_checked": n_envs * steps, "mismatches": bad,
            "fields": "positions, flags (done/won/invalid/timeout), reward",
            "observations_checked": n_obs, "observation_or_valid_mask_mismatches": obs_bad,
            "checker": "oracle/ts_oracle.c"}


def run_reference(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    rate, elapsed, envs_per_proc = cpu_loop_rate(cores, args.steps, args.warmup, budget_s=25.0)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / max(1, args.steps),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": workload_config(args.gpus, args.envs),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{cores} processes x {envs_per_proc} envs x {args.steps} steps of the same 6x6/4-tile "
                                       "workload; oracle/py_port.py = the reference's Python step loop restated "
                                       "(the pure-Python reference cannot travel to this box)"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------
# clocks sampler (NVML, in-process)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
               0x10: "sync_boost"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.stop_flag = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for b, name in self.REASONS.items():
                    if bits & b:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop_flag.set()
        if self.nv:
            self.thread.join(timeout=2)

    def summary(self) -> dict:
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_ours(args) -> int:
    import torch
    import torch.distributed as dist
    import tiler_slider_b200 as ts

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_local = args.envs
    env = ts.BatchedTilerSliderEnv.synthetic(n_local, S, T, W_WALLS, MULTI, seed=PUZZLE_SEED, env_index_base=rank * n_local,
                                             max_steps=MAX_STEPS, auto_reset=True, device=dev)
    gen = torch.Generator(device=dev).manual_seed(ACTION_SEED + rank)
    actions = torch.randint(0, 4, (N_ACTION_ROWS, env.capacity), dtype=torch.uint8, device=dev, generator=gen)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput --------------------------------------------------------
    # The N_ACTION_ROWS steps of one pass over the action rows are captured once into a CUDA graph
    # (still one ts_step launch per step): +1-2 % at 16.7M envs, 3x for launch-bound batches
    graph = env.capture_steps(actions) if args.graph else None

    def run_steps(n):
        if graph is None:
            for k in range(n):
                env.step(actions[k % N_ACTION_ROWS])
        else:
            for _ in range(n // N_ACTION_ROWS):
                graph.replay()
            for k in range(n % N_ACTION_ROWS):
                env.step(actions[k])
    run_steps(args.warmup)
    barrier()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        start.record()
        run_steps(args.steps)
        stop.record()
        barrier()
    ms = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=dev)
    wins = ((env.flags & ts.F_WON) != 0).sum().to(torch.float64).reshape(1)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(wins, op=dist.ReduceOp.SUM)
    total_ms = float(ms.item())
    value = n_local * world * args.steps / (total_ms * 1e-3)
    per_launch_s = total_ms * 1e-3 / args.steps
    clocks = clk.summary()

    # ---- end to end through host buffers -----------------------------------------------------
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    h_act = torch.randint(0, 4, (n_local,), dtype=torch.uint8).pin_memory()
    h_rew = torch.empty(n_local, dtype=torch.float32).pin_memory()
    h_done = torch.empty(n_local, dtype=torch.uint8).pin_memory()
    for _ in range(3):
        env.step_host(h_act, h_rew, h_done)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        env.step_host(h_act, h_rew, h_done)       # returns after reward/done landed on the host
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n_local * world * e2e_steps / float(e2e_s.item())
    # compact variant: only the status byte travels back (done = bit 0, reward = f(WON, INVALID))
    h_flags = torch.empty(n_local, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        env.step_host(h_act, h_flags=h_flags)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        env.step_host(h_act, h_flags=h_flags)
    barrier()
    e2c_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2c_s, op=dist.ReduceOp.MAX)
    e2e_compact = n_local * world * e2e_steps / float(e2c_s.item())

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = ALGO_BYTES * n_local / per_launch_s / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": workload_config(world, n_local),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": args.traffic_bytes, "kernel": f"ts::step_kernel<{S},{T}>" if S <= 8 else f"ts::wide_step_kernel<{T}>",
                             "algorithmic_bytes_per_launch": ALGO_BYTES * n_local,
                             "avg_launch_ms": per_launch_s * 1e3,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_local * world,
                        "d2h_bytes_per_step": 5 * n_local * world, "steps": e2e_steps,
                        "path": "BatchedTilerSliderEnv.step_host -> ts_step_host (pinned host actions in, reward f32 + done u8 out)",
                        "compact_variant": {"value": e2e_compact, "d2h_bytes_per_step": n_local * world,
                                            "what": "same call, only the 1-byte status word downloaded (done = bit 0, "
                                                    "reward = function of the WON / INVALID bits)"}},
                "gpu_launches": args.steps, "clocks": clocks, "cuda_graph": bool(args.graph),
                "wins_in_last_step": int(wins.item())}
        if world == 1 and not args.no_cpu:
            rate, elapsed, n_cpu = cpu_loop_rate(1, 300, 10, budget_s=12.0)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{n_cpu} envs x 300 steps of the same workload, single process, "
                                              "oracle/py_port.py (reference Python step loop restated)",
                                    "c_oracle_env_steps_per_s_1core": c_oracle_rate()}
            line["parity"] = parity_check(ts, dev)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def select_config(name: str) -> None:
    global S, T, W_WALLS, MULTI, ENVS_PER_GPU, PUZZLE_SEED, ACTION_SEED, ALGO_BYTES
    S, T, W_WALLS, MULTI, ENVS_PER_GPU, PUZZLE_SEED, ACTION_SEED = CONFIGS[name]
    ALGO_BYTES = 3 * T + (S * S + 7) // 8 + 8


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", choices=sorted(CONFIGS), default="c3")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (default: the config's)")
    ap.add_argument("--scaling", choices=["weak", "strong"], default="weak",
                    help="weak (default): the config's env count on EVERY GPU; strong: that count split over the GPUs")
    ap.add_argument("--e2e-steps", type=int, default=50)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch every step from Python instead of replaying a CUDA graph of N_ACTION_ROWS steps")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per launch from the committed ncu --set full capture (profiles/)")
    args = ap.parse_args()
    select_config(args.config)
    if args.envs is None:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        args.envs = ENVS_PER_GPU if args.scaling == "weak" else -(-ENVS_PER_GPU // world)
    if args.traffic_bytes is None and args.envs == ENVS_PER_GPU:
        args.traffic_bytes = NCU_TRAFFIC_BYTES.get(args.config)
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
