#!/bin/bash
# round 2, call 6 (2 GPUs): multicast path on real GPUs (tests, driver, bench), one-graph iteration, rmat N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_drivers.py tests/test_gpu_synth.py -m gpu -x -q -s > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2f_tests.log
grep -h "multicast path\|passed\|failed\|rc=" gpurun_out/r2f_tests.log | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 100 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 python bench.py --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench n1 rc=$?"
for sync in nccl mcast; do
  opencl-spmv-algorithms_b200/host/bin/sigma_c --synthetic laplace7:400x400x100 --iters 100 --gpus 2 --sync $sync --json > gpurun_out/r2f_driver_sigma_c_n2_$sync.json 2> gpurun_out/r2f_driver_sigma_c_n2_$sync.err; echo "driver $sync rc=$?"
  cat gpurun_out/r2f_driver_sigma_c_n2_$sync.json
done
opencl-spmv-algorithms_b200/host/bin/csr --synthetic laplace7:400x400x100 --iters 100 --gpus 2 --json > gpurun_out/r2f_driver_csr_n2.json 2> gpurun_out/r2f_driver_csr_n2.err; echo "driver csr rc=$?"
cat gpurun_out/r2f_driver_csr_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --workload rmat --steps 20 > gpurun_out/r2f_rmat24_n2.json 2> gpurun_out/r2f_rmat24_n2.err; echo "rmat n2 rc=$?"
