#!/bin/bash
# round-1e GPU call 2: full GPU test suite, smoke, sustained benches (defaults + A/B), ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c2_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/c2_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c2_smoke.log
timeout 400 python bench.py > gpurun_out/c2_bench_banded.json 2> gpurun_out/c2_bench_banded.err; echo "rc=$?" >> gpurun_out/c2_bench_banded.err
timeout 300 python bench.py --workload cant > gpurun_out/c2_bench_cant.json 2> gpurun_out/c2_bench_cant.err; echo "rc=$?" >> gpurun_out/c2_bench_cant.err
timeout 300 python bench.py --workload cant --dtype f32 --no-cpu-baseline > gpurun_out/c2_bench_cant_f32.json 2> gpurun_out/c2_bench_cant_f32.err
timeout 300 python bench.py --dtype f64 --no-cpu-baseline > gpurun_out/c2_bench_banded_f64.json 2> gpurun_out/c2_bench_banded_f64.err
# sustained A/B of the defaults the burst sweep picked
B="python bench.py --no-cpu-baseline --no-e2e --steps 200"
B200_SELL_TMA=0 timeout 200 $B > gpurun_out/c2_ab_sell_notma.json 2>/dev/null
B200_CMRS_U=2 B200_COO_U=2 timeout 200 $B > gpurun_out/c2_ab_u2.json 2>/dev/null
B200_CSR_UNROLL=1 B200_ELL_UNROLL=1 B200_SELL_TMA_BLOCKS=3 timeout 200 $B > gpurun_out/c2_ab_u1.json 2>/dev/null
B200_CSR_UNROLL=4 B200_ELL_UNROLL=4 B200_SELL_TMA_BLOCKS=1 timeout 200 $B > gpurun_out/c2_ab_u4.json 2>/dev/null
timeout 200 $B --gpus 1 > gpurun_out/c2_ab_default_again.json 2>/dev/null
# ncu: launch list + full-set capture, banded (default bench) and cant
K='regex:csr_vector|sell32|ell_rowmajor|cmrs_kernel|coo_kernel'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c2_launches_banded.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c2_ncu_l_banded.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "$K" -s 15 -c 5 -f -o gpurun_out/c2_prof_banded python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c2_ncu_f_banded.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c2_launches_cant.csv python bench.py --workload cant --steps 7 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c2_ncu_l_cant.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k "$K" -s 70 -c 5 -f -o gpurun_out/c2_prof_cant python bench.py --workload cant --steps 7 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/c2_ncu_f_cant.log 2>&1
for f in gpurun_out/c2_tests.log gpurun_out/c2_smoke.log gpurun_out/c2_bench_banded.err gpurun_out/c2_bench_cant.err; do echo "== $f"; tail -n 3 $f; done
ls -la gpurun_out | tail -n 40
exit 0
