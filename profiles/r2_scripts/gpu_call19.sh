#!/bin/bash
# round 2, call 19 (1 GPU): ncu --set full of the pipelined kernels (fused and plain) + the CSR stream kernel on the
# 7-point Laplacian with the library's final defaults, and the launch list of the driver's bench command
mkdir -p gpurun_out
T=opencl-spmv-algorithms_b200/tools/ncu_target.py
python $T --workload laplace > gpurun_out/r2s_plain_laplace.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:sell32_|csr_stream_kernel' -c 12 -f -o gpurun_out/r2s_ncu_laplace python $T --workload laplace > gpurun_out/r2s_ncu_laplace.log 2>&1
echo "ncu laplace rc=$?"
ncu -i gpurun_out/r2s_ncu_laplace.ncu-rep --page raw --csv > gpurun_out/r2s_ncu_laplace_raw.csv 2>/dev/null
ncu -i gpurun_out/r2s_ncu_laplace.ncu-rep --page source --csv > gpurun_out/r2s_ncu_laplace_source.csv 2>/dev/null
gzip -f gpurun_out/r2s_ncu_laplace_source.csv
du -sh gpurun_out/r2s_ncu_laplace.ncu-rep
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s_bench_plain.json 2> gpurun_out/r2s_bench_plain.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2s_launches.csv python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2s_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
wc -l gpurun_out/r2s_launches.csv
