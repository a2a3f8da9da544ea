#!/bin/bash
# round 2, call 4 (2 GPUs): full GPU suite with safe-overlap default + new kernels + C driver iterated mode;
# step-time probe with the true iteration variants; bench N=2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_tests.log
tail -5 gpurun_out/r2d_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 opencl-spmv-algorithms_b200/tools/iter_probe.py > gpurun_out/r2d_iter_probe_n2.json 2> gpurun_out/r2d_iter_probe_n2.err; echo "probe rc=$?"
cat gpurun_out/r2d_iter_probe_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; echo "bench n2 rc=$?"
opencl-spmv-algorithms_b200/host/bin/sigma_c --synthetic laplace7:400x400x100 --iters 100 --gpus 2 --json > gpurun_out/r2d_driver_sigma_c_n2.json 2> gpurun_out/r2d_driver_sigma_c_n2.err; echo "driver rc=$?"
cat gpurun_out/r2d_driver_sigma_c_n2.json
opencl-spmv-algorithms_b200/host/bin/csr --synthetic laplace7:400x400x100 --iters 100 --gpus 2 --json > gpurun_out/r2d_driver_csr_n2.json 2>> gpurun_out/r2d_driver_sigma_c_n2.err; echo "driver rc=$?"
cat gpurun_out/r2d_driver_csr_n2.json
timeout 600 python bench.py --workload cant --dtype f64 --steps 50 > gpurun_out/r2d_bench_cant_f64.json 2> gpurun_out/r2d_bench_cant_f64.err; echo "cant rc=$?"
