#!/bin/bash
# Round-2 ncu captures, second batch: the kernels rewritten after the first batch (K3, valid, K6).
set -x
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
T="python profiles/experiments/r2_targets.py"
for w in observe valid; do
  $T $w > $O/r2b_plain_$w.log 2>&1 && $NCU -k regex:${w}_kernel -s 1 -c 1 -o $O/r2b_$w $T $w > $O/r2b_ncu_$w.log 2>&1
done
$NCU -k regex:wide_valid_kernel -s 1 -c 1 -o $O/r2b_valid_wide $T valid > $O/r2b_ncu_valid_wide.log 2>&1
$NCU -k regex:observe_kernel -s 7 -c 1 -o $O/r2b_observe_wide $T observe > $O/r2b_ncu_observe_wide.log 2>&1
$T bfs_local > $O/r2b_plain_bfs_local.log 2>&1 && $NCU -k regex:bfs_local_kernel -s 1 -c 1 -o $O/r2b_bfs_local $T bfs_local > $O/r2b_ncu_bfs_local.log 2>&1
for r in $O/r2b_*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > ${b}_raw.csv 2>/dev/null
  rm -f $r
done
ls -la $O | grep r2b_
