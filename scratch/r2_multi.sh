#!/bin/bash
# multi-GPU record (gpurun --gpus N): parity test of a sharded run against the single-GPU run, weak and strong scaling lines
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=${1:-2}
python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/r2_multi_tests_n$N.log 2>&1; tail -3 gpurun_out/r2_multi_tests_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
$TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_weak_n$N.log 2>&1; tail -c 400 gpurun_out/r2_bench_weak_n$N.log
$TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --scaling strong --global-batch 65536 > gpurun_out/r2_bench_strong65536_n$N.log 2>&1; tail -c 400 gpurun_out/r2_bench_strong65536_n$N.log
$TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --scaling strong --global-batch 4096 > gpurun_out/r2_bench_strong4096_n$N.log 2>&1; tail -c 400 gpurun_out/r2_bench_strong4096_n$N.log
