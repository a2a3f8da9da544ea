#!/bin/bash
# round 2, call 2 (2 GPUs): real multi-GPU parity test, fabric probe, default bench line at N=1 and N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2b_topo.txt 2>&1
opencl-spmv-algorithms_b200/tools/probe_fabric > gpurun_out/r2b_fabric_n2.json 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_synth.py -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2b_tests.log
tail -5 gpurun_out/r2b_tests.log
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/r2b_nccl.%p.log timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; echo "bench n2 rc=$?"
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload rmat --rmat-scale 22 --steps 10 --rmat-sigmas 1,4096 > gpurun_out/r2b_rmat22_n2.json 2> gpurun_out/r2b_rmat22_n2.err; echo "rmat n2 rc=$?"
grep -h -i "nvls\|multicast" gpurun_out/r2b_nccl.*.log | head -5
