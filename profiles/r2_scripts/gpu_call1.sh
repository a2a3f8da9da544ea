#!/bin/bash
# round 2, call 1 (1 GPU): whole GPU test suite + default bench line after the bench.py restructure
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.csv
opencl-spmv-algorithms_b200/tools/probe_fabric > gpurun_out/r2a_fabric_n1.json 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 50 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --workload laplace-iter --steps 100 > gpurun_out/r2a_iter_n1.json 2> gpurun_out/r2a_iter_n1.err; echo "iter rc=$?"
tail -3 gpurun_out/r2a_tests.log
