#!/bin/bash
# round-2 ncu captures (one GPU): the headline kernel through the standalone harness (fast and strict builds), the lock-step
# fixed-dt kernel, C3 / C4 / C5 through the package.  Build the probes first (tools/experiments/README.md).
set -x
mkdir -p gpurun_out
for v in fast strict; do
  ncu --set full --clock-control none --import-source on -k regex:k_probe -s 1 -c 1 -f -o gpurun_out/r2b_c2_$v ./tools/bin/c2_probe_$v 4194304 1 1 0 > gpurun_out/ncu_c2_$v.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_ode_lockstep -s 1 -c 1 -f -o gpurun_out/r2e_c1_ls_fast_ref python tools/ncu_targets.py c1 fast > gpurun_out/ncu_c1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sde_solve -s 1 -c 1 -f -o gpurun_out/r2_c5_fast python tools/ncu_targets.py c5 fast > gpurun_out/ncu_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ode_asolve2 -s 1 -c 1 -f -o gpurun_out/r2_c4_fast python tools/ncu_targets.py c4 fast > gpurun_out/ncu_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ode_asolve2 -s 1 -c 1 -f -o gpurun_out/r2_c3_fast python tools/ncu_targets.py c3 fast > gpurun_out/ncu_c3.log 2>&1
python tools/gpu_configs.py > gpurun_out/r2_configs_1gpu.log 2>&1
tail -12 gpurun_out/r2_configs_1gpu.log
ls -la gpurun_out/*.ncu-rep
