"""Condense `ncu --page raw --csv` exports into one JSON of the metrics the roofline discussion uses.
    python profiles/experiments/ncu_summary.py gpurun_out/r2_*_raw.csv > profiles/r2_ncu_summary.json"""
import csv
import json
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_bytes.sum": "l2_bytes",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__inst_executed.sum": "warp_instructions",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__instruction_throughput.avg.pct_of_peak_sustained_active": "instruction_throughput_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "launch__shared_mem_per_block_static": "smem_static",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "global_load_sectors",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum": "global_load_requests",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier_per_issue",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard_per_issue",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle_per_issue",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle_per_issue",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait_per_issue",
}
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    out = {}
    for path in sys.argv[1:]:
        with open(path, newline="") as f:
            rows = list(csv.reader(f))
        hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        names, units = rows[hdr], rows[hdr + 1]
        launches = []
        for r in rows[hdr + 2:]:
            if len(r) != len(names):
                continue
            rec = {"kernel": r[names.index("Kernel Name")][:110]}
            for i, n in enumerate(names):
                k = KEYS.get(n)
                if k is None or k in rec:
                    continue
                v = num(r[i])
                if v is None:
                    continue
                rec[k] = v * UNIT_SCALE.get(units[i], 1.0) if k in ("duration_us", "dram_read_bytes", "dram_write_bytes", "l2_bytes") else v
            if "dram_read_bytes" in rec:
                rec["dram_bytes"] = rec["dram_read_bytes"] + rec.get("dram_write_bytes", 0.0)
                if rec.get("duration_us"):
                    rec["dram_gb_per_s"] = rec["dram_bytes"] / rec["duration_us"] / 1e3
            if rec.get("global_load_requests"):
                rec["sectors_per_load_request"] = rec["global_load_sectors"] / rec["global_load_requests"]
            launches.append(rec)
        out[path.split("/")[-1].replace("_raw.csv", "")] = launches
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
