#!/bin/bash
# round 2, call 3 (2 GPUs): new kernel tests, where-does-the-step-go probe, e2e with column blocks, cant sweeps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -5 gpurun_out/r2c_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 opencl-spmv-algorithms_b200/tools/iter_probe.py > gpurun_out/r2c_iter_probe_n2.json 2> gpurun_out/r2c_iter_probe_n2.err; echo "probe rc=$?"
cat gpurun_out/r2c_iter_probe_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err; echo "bench n2 rc=$?"
for dt in f64 f32; do
  timeout 600 python opencl-spmv-algorithms_b200/tools/sweep_variants.py --workload cant --dtype $dt --no-tma --only cmrs,cmrs_packed,csr,coo --out gpurun_out/r2c_sweep_cant_$dt.json 2> gpurun_out/r2c_sweep_cant_$dt.txt; echo "sweep $dt rc=$?"
done
timeout 600 python bench.py --workload rmat --rmat-scale 22 --steps 10 --rmat-sigmas 4096 --no-cpu-baseline > gpurun_out/r2c_rmat22_n1.json 2> gpurun_out/r2c_rmat22_n1.err; echo "rmat rc=$?"
