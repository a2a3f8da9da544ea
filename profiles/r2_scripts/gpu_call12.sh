#!/bin/bash
# round 2, call 12 (8 GPUs): the records after the OMP_PROC_BIND / multicast fixes
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29581 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2l_bench_n8.json 2> gpurun_out/r2l_bench_n8.err; echo "bench n8 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29582 opencl-spmv-algorithms_b200/tools/iter_probe.py > gpurun_out/r2l_iter_probe_n8.json 2> gpurun_out/r2l_iter_probe_n8.err; echo "probe rc=$?"
cat gpurun_out/r2l_iter_probe_n8.json
timeout 400 $TR --nproc-per-node 8 --master-port 29583 bench.py --gpus 8 --workload laplace-iter --steps 200 --iter-extras > gpurun_out/r2l_iter_n8.json 2> gpurun_out/r2l_iter_n8.err; echo "iter n8 rc=$?"
for sync in nccl mcast; do
  timeout 200 opencl-spmv-algorithms_b200/host/bin/sigma_c --synthetic laplace7:400x400x400 --iters 100 --gpus 8 --sync $sync --json > gpurun_out/r2l_driver_sigma_c_n8_$sync.json 2>/dev/null; cat gpurun_out/r2l_driver_sigma_c_n8_$sync.json
done
