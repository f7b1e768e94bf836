#!/bin/bash
# A/B run of c2_probe builds on the GPU box: every binary under tools/bin/ whose name starts with c2_
# usage: tools/run_c2_ab.sh [N] [reps] ["retire batches"]  -> gpurun_out/c2_ab.jsonl
N=${1:-8388608}; R=${2:-3}; RBS=${3:-0}
mkdir -p gpurun_out
: > gpurun_out/c2_ab.jsonl
for b in tools/bin/c2_*; do
  for rb in $RBS; do
    echo "== $b retire_batch=$rb" | tee -a gpurun_out/c2_ab.jsonl
    timeout 120 $b $N $R 1 1 1 0 $rb | tee -a gpurun_out/c2_ab.jsonl
  done
done
