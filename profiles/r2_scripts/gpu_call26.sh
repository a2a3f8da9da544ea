#!/bin/bash
# round 2, call 26 (1 GPU): ncu --set full of the five kernels on the cant-shaped matrix in fp32 (BASELINE configs[0] names
# CSR on cant in fp32; the round's earlier capture was fp64), library defaults
mkdir -p gpurun_out
T=opencl-spmv-algorithms_b200/tools/ncu_target.py
K='regex:csr_vector|csr_stream_kernel|csr_long|coo_kernel|cmrs_kernel|cmrs_stream|sell32_|ell_rowmajor'
python $T --workload cant --dtype f32 > gpurun_out/r2z_plain_cant_f32.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "$K" -c 20 -f -o gpurun_out/r2z_ncu_cant_f32 python $T --workload cant --dtype f32 > gpurun_out/r2z_ncu_cant_f32.log 2>&1
echo "ncu cant f32 rc=$?"
ncu -i gpurun_out/r2z_ncu_cant_f32.ncu-rep --page raw --csv > gpurun_out/r2z_ncu_cant_f32_raw.csv 2>/dev/null
rm -f gpurun_out/r2z_ncu_cant_f32.ncu-rep
wc -l gpurun_out/r2z_ncu_cant_f32_raw.csv
