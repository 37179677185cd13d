#!/bin/bash
# A/B of the three delivery schemes of ts_bfs_expand_exchange (TS_BFS_XCHG) on N GPUs of one box.
# Needs commit 7ceda0d (the alternative schemes were removed afterwards; results: profiles/r2_xchg_modes_n4.txt):
#   bash profiles/experiments/xchg_modes.sh N [puzzles per GPU]
# Hash-partitioned BFS, 6x6 / 4 tiles / 8 walls, two repetitions per mode; one JSON line each in
# gpurun_out/xchg_modes_nN.log.
N=${1:-2}
PER=${2:-65536}
OUT=gpurun_out/xchg_modes_n$N.log
: > $OUT
for rep in 1 2; do
for mode in cursor staged segments; do
    TS_BFS_XCHG=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 \
        --master-port 29551 bfs_bench.py --puzzles $((PER * N)) --check 0 --exchange p2p --mode hash 2>gpurun_out/xchg_err.log | grep '^{' >> $OUT
done
done
python - <<PY
import json
for l in open("$OUT"):
    d = json.loads(l)
    print(d["config"].split("exchange ")[1], round(d["seconds"] * 1e3, 2), "ms", "%.3e" % d["generated_successors_per_s"])
PY
