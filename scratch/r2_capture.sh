#!/bin/bash
# Round-2 evidence capture (run through gpurun): timing marks, launch list of one training iteration, ncu --set full of the hot kernels.
set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
LRNDE_TIMING=1 python scratch/step_timing.py > gpurun_out/r2_step_timing.log 2>&1
python bench.py --steps 2 --warmup 1 --loop-mode 2 --no-cpu-baseline --no-secondary --e2e-steps 0 > gpurun_out/r2_plain_lm2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_launches_bench.csv \
   python bench.py --steps 2 --warmup 1 --loop-mode 2 --no-cpu-baseline --no-secondary --e2e-steps 0 > gpurun_out/r2_ncu_launch.log 2>&1
ITERS=1 python scratch/adj_prof.py > gpurun_out/r2_plain_adj.log 2>&1 || exit 1
ITERS=1 ncu --set full --clock-control none --import-source on -k regex:'chain_kernel|kgemm_kernel|pairacc_kernel|adj_mu_kernel|adj_reduce_kernel' \
   --launch-skip 20 -c 12 -f -o gpurun_out/r2_hot_full python scratch/adj_prof.py > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out | tail
