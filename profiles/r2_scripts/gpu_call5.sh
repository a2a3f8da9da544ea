#!/bin/bash
# round 2, call 5 (1 GPU): parity tests of the new layouts, then ncu --set full of every SpMV kernel with the
# library's default variants on the four workloads (R-MAT scale 24 first: it had no capture at all)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2e_tests.log
tail -4 gpurun_out/r2e_tests.log
T=opencl-spmv-algorithms_b200/tools/ncu_target.py
K='regex:csr_vector|csr_stream_kernel|csr_long|coo_kernel|cmrs_kernel|cmrs_stream|sell32_|ell_rowmajor'
for w in rmat banded cant laplace; do
  python $T --workload $w > gpurun_out/r2e_plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "$K" -c 40 -f -o gpurun_out/r2e_ncu_$w python $T --workload $w > gpurun_out/r2e_ncu_$w.log 2>&1
  echo "ncu $w rc=$?"
  if [ -f gpurun_out/r2e_ncu_$w.ncu-rep ]; then
    ncu -i gpurun_out/r2e_ncu_$w.ncu-rep --page raw --csv > gpurun_out/r2e_ncu_${w}_raw.csv 2>/dev/null
    ls -la gpurun_out/r2e_ncu_$w.ncu-rep
  fi
done
# source-level pages of the R-MAT capture (where the stalls are), then drop the big report files
ncu -i gpurun_out/r2e_ncu_rmat.ncu-rep --page source --csv > gpurun_out/r2e_ncu_rmat_source.csv 2>/dev/null
ncu -i gpurun_out/r2e_ncu_laplace.ncu-rep --page source --csv > gpurun_out/r2e_ncu_laplace_source.csv 2>/dev/null
du -sh gpurun_out/*.ncu-rep
rm -f gpurun_out/r2e_ncu_banded.ncu-rep gpurun_out/r2e_ncu_cant.ncu-rep
gzip -f gpurun_out/r2e_ncu_*_source.csv
