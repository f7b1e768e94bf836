#!/bin/bash
# static / queue schedule x unsorted / sorted-by-rho start order on C2 (tools/c2_probe.cu): throughput, and the
# active threads per executed warp instruction from ncu
N=${1:-8388608}
for sched in 0 1; do for sorted in 0 1; do
  ./tools/bin/c2_probe_v4_mb5 $N 3 1 0 $sched $sorted
done; done
for sched in 0 1; do for sorted in 0 1; do
  echo -n "sched=$sched sorted=$sorted: "
  ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum -k regex:k_probe -s 1 -c 1 ./tools/bin/c2_probe_v4_mb5 4194304 1 1 0 $sched $sorted 2>&1 | grep -E "ratio|inst_executed.sum" | awk '{printf "%s %s  ", $1, $NF}'; echo
done; done
