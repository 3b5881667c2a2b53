#!/bin/bash
# builder-run bench lines of the secondary configs (BASELINE configs[0..3]) at the reference tolerances
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_default.log 2>&1; tail -c 200 gpurun_out/r2_bench_default.log; echo
python bench.py --batch 128 --ref-batch 128 --no-secondary > gpurun_out/r2_bench_b128.log 2>&1; tail -c 200 gpurun_out/r2_bench_b128.log; echo
python bench.py --workload mnist_sde > gpurun_out/r2_bench_sde.log 2>&1; tail -c 200 gpurun_out/r2_bench_sde.log; echo
python bench.py --workload physionet > gpurun_out/r2_bench_physionet.log 2>&1; tail -c 200 gpurun_out/r2_bench_physionet.log; echo
python bench.py --workload cifar10 > gpurun_out/r2_bench_cifar10.log 2>&1; tail -c 200 gpurun_out/r2_bench_cifar10.log; echo
python bench.py --batch 65536 --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_b65536.log 2>&1; tail -c 200 gpurun_out/r2_bench_b65536.log; echo
