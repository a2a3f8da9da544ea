#!/bin/bash
# round-1e GPU call 4: launch overlap (PDL), halo-limited exchange (emulated ranks), packed CMRS; sweeps with overlap
mkdir -p gpurun_out
S=opencl-spmv-algorithms_b200/tools/sweep_variants.py
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synth.py tests/test_gpu_tma.py -m gpu -q > gpurun_out/c4_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c4_tests.log
timeout 300 python $S --workload cant --dtype f64 --no-tma --overlap --out gpurun_out/c4_sweep_cant_f64_ovl.json 2> gpurun_out/c4_sweep_cant_f64_ovl.log
timeout 300 python $S --workload cant --dtype f32 --no-tma --overlap --out gpurun_out/c4_sweep_cant_f32_ovl.json 2> gpurun_out/c4_sweep_cant_f32_ovl.log
timeout 300 python bench.py --workload cant > gpurun_out/c4_bench_cant.json 2> gpurun_out/c4_bench_cant.err; echo "rc=$?" >> gpurun_out/c4_bench_cant.err
timeout 300 python bench.py --workload cant --dtype f32 --no-cpu-baseline > gpurun_out/c4_bench_cant_f32.json 2> gpurun_out/c4_bench_cant_f32.err
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c4_bench_banded.json 2> gpurun_out/c4_bench_banded.err; echo "rc=$?" >> gpurun_out/c4_bench_banded.err
timeout 300 python bench.py --dtype f64 --no-cpu-baseline --no-e2e > gpurun_out/c4_bench_banded_f64.json 2> gpurun_out/c4_bench_banded_f64.err
timeout 300 python bench.py --workload laplace-iter --iter-format sell --steps 100 > gpurun_out/c4_iter_sell.json 2> gpurun_out/c4_iter_sell.err; echo "rc=$?" >> gpurun_out/c4_iter_sell.err
tail -n 15 gpurun_out/c4_tests.log
tail -n 3 gpurun_out/c4_bench_cant.err gpurun_out/c4_iter_sell.err gpurun_out/c4_bench_banded.err
exit 0
