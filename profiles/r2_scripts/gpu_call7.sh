#!/bin/bash
# round 2, call 7 (1 GPU): CSR stream G sweep on R-MAT scale 24, record runs (rmat24, cant f64/f32), memcheck
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "csr_stream or sell16 or conversions or advice" > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2g_tests.log
tail -3 gpurun_out/r2g_tests.log
for g in 1 2 4; do
  B200_CSR_STREAM_G=$g timeout 600 python bench.py --workload rmat --steps 20 --rmat-sigmas "" --no-e2e --no-cpu-baseline > gpurun_out/r2g_rmat24_G$g.json 2> gpurun_out/r2g_rmat24_G$g.err; echo "rmat G=$g rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2g_rmat24_G$g.json'));print({k:v['ms'] for k,v in d['formats'].items()})"
done
timeout 900 python bench.py --workload rmat --steps 20 > gpurun_out/r2g_bench_rmat24.json 2> gpurun_out/r2g_bench_rmat24.err; echo "rmat full rc=$?"
timeout 600 python bench.py --workload cant --dtype f64 --steps 50 > gpurun_out/r2g_bench_cant_f64.json 2> gpurun_out/r2g_bench_cant_f64.err; echo "cant f64 rc=$?"
timeout 600 python bench.py --workload cant --dtype f32 --steps 50 > gpurun_out/r2g_bench_cant_f32.json 2> gpurun_out/r2g_bench_cant_f32.err; echo "cant f32 rc=$?"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "formats_random or sell16 or cmrs_stream or sell_narrow or csr_stream" > gpurun_out/r2g_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/r2g_memcheck.log
grep -c "Invalid\|ERROR SUMMARY" gpurun_out/r2g_memcheck.log
