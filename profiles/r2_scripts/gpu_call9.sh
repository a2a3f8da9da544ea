#!/bin/bash
# round 2, call 9 (8 GPUs): the multi-GPU records -- tests over 2 and 4 real GPUs, default bench line at N = 8 and 4,
# laplace-iter with extras, the C drivers over 8 devices (NCCL / multicast / all-gather), R-MAT at N = 8, step probe
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2i_topo.txt 2>&1
opencl-spmv-algorithms_b200/tools/probe_fabric > gpurun_out/r2i_fabric_n8.json 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_drivers.py -m gpu -x -q -k "real_gpus or iterated" > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2i_tests.log
tail -3 gpurun_out/r2i_tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/r2i_nccl.%p.log timeout 800 $TR --nproc-per-node 8 --master-port 29551 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2i_bench_n8.json 2> gpurun_out/r2i_bench_n8.err; echo "bench n8 rc=$?"
timeout 600 $TR --nproc-per-node 4 --master-port 29552 bench.py --gpus 4 --steps 100 --warmup 5 > gpurun_out/r2i_bench_n4.json 2> gpurun_out/r2i_bench_n4.err; echo "bench n4 rc=$?"
timeout 600 $TR --nproc-per-node 8 --master-port 29553 bench.py --gpus 8 --workload laplace-iter --steps 200 --iter-extras > gpurun_out/r2i_iter_n8.json 2> gpurun_out/r2i_iter_n8.err; echo "iter n8 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29554 opencl-spmv-algorithms_b200/tools/iter_probe.py > gpurun_out/r2i_iter_probe_n8.json 2> gpurun_out/r2i_iter_probe_n8.err; echo "probe rc=$?"
cat gpurun_out/r2i_iter_probe_n8.json
for sync in nccl mcast; do
  timeout 300 opencl-spmv-algorithms_b200/host/bin/sigma_c --synthetic laplace7:400x400x400 --iters 100 --gpus 8 --sync $sync --json > gpurun_out/r2i_driver_sigma_c_n8_$sync.json 2> gpurun_out/r2i_driver_sigma_c_n8_$sync.err; echo "driver $sync rc=$?"
  cat gpurun_out/r2i_driver_sigma_c_n8_$sync.json
done
timeout 300 opencl-spmv-algorithms_b200/host/bin/csr --synthetic laplace7:400x400x400 --iters 100 --gpus 8 --json > gpurun_out/r2i_driver_csr_n8.json 2> gpurun_out/r2i_driver_csr_n8.err; echo "driver csr rc=$?"
cat gpurun_out/r2i_driver_csr_n8.json
timeout 900 $TR --nproc-per-node 8 --master-port 29555 bench.py --gpus 8 --workload rmat --steps 20 > gpurun_out/r2i_rmat24_n8.json 2> gpurun_out/r2i_rmat24_n8.err; echo "rmat n8 rc=$?"
timeout 900 $TR --nproc-per-node 4 --master-port 29556 bench.py --gpus 4 --workload rmat --steps 20 --rmat-sigmas 65536 > gpurun_out/r2i_rmat24_n4.json 2> gpurun_out/r2i_rmat24_n4.err; echo "rmat n4 rc=$?"
grep -h -i "nvls" gpurun_out/r2i_nccl.*.log | sort | uniq -c | sort -rn | head -5 > gpurun_out/r2i_nccl_nvls.txt
rm -f gpurun_out/r2i_nccl.*.log
