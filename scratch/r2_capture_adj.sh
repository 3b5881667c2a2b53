#!/bin/bash
# ncu --set full of the adjoint attempt kernels (scratch/adj_prof.py, one iteration at B = 8192)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/t_r2j.log 2>&1; tail -3 gpurun_out/t_r2j.log
ITERS=1 python scratch/adj_prof.py > gpurun_out/r2_plain_adj.log 2>&1 || exit 1
ITERS=1 ncu --set full --clock-control none --import-source on -k regex:'adj_chain_kernel|pairacc_kernel|adj_mu_kernel|adj_reduce_kernel' \
   --launch-skip 9 -c 8 -f -o gpurun_out/r2_adj_full python scratch/adj_prof.py > gpurun_out/r2_ncu_adj.log 2>&1
ITERS=1 ncu --set full --clock-control none -k regex:'kgemm_kernel' --launch-skip 48 -c 4 -f -o gpurun_out/r2_kgemm_adj_full python scratch/adj_prof.py > gpurun_out/r2_ncu_adj2.log 2>&1
tail -3 gpurun_out/r2_ncu_adj.log
