#!/bin/bash
# round-1e GPU call 3: sustained (200-step) A/B of the kernel defaults, rmat + laplace-iter refresh
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --no-e2e --steps 200"
run() { name=$1; shift; env "$@" timeout 200 $B $EXTRA > gpurun_out/c3_$name.json 2> gpurun_out/c3_$name.err; }
EXTRA=""
run f32_default A=1
run f32_tma_poll B200_SELL_TMA_SUSPEND_NS=0
run f32_tma_sleep20us B200_SELL_TMA_SUSPEND_NS=20000
run f32_notma B200_SELL_TMA=0
run f32_notma_u2 B200_SELL_TMA=0 B200_SELL_UNROLL=2 B200_COO_U=2 B200_CMRS_U=1
EXTRA="--dtype f64"
run f64_default A=1
run f64_u2 B200_CSR_UNROLL=2 B200_ELL_UNROLL=2 B200_SELL_TMA=0 B200_SELL_UNROLL=2 B200_COO_U=1 B200_CMRS_U=2
run f64_u4 B200_CSR_UNROLL=4 B200_ELL_UNROLL=4 B200_SELL_TMA=0 B200_SELL_UNROLL=4 B200_COO_U=4
run f64_tma_poll B200_SELL_TMA_SUSPEND_NS=0 B200_SELL_TMA_BLOCKS=1
EXTRA=""
run f32_default_again A=1
timeout 200 python -m pytest tests/test_gpu_tma.py -m gpu -x -q > gpurun_out/c3_tma.log 2>&1; echo "rc=$?" >> gpurun_out/c3_tma.log
timeout 400 python bench.py --workload rmat --steps 20 --warmup 3 > gpurun_out/c3_rmat.json 2> gpurun_out/c3_rmat.err; echo "rc=$?" >> gpurun_out/c3_rmat.err
timeout 300 python bench.py --workload laplace-iter --iter-format sell --steps 100 > gpurun_out/c3_iter_sell.json 2> gpurun_out/c3_iter_sell.err; echo "rc=$?" >> gpurun_out/c3_iter_sell.err
timeout 300 python bench.py --workload laplace-iter --iter-format csr --steps 100 > gpurun_out/c3_iter_csr.json 2> gpurun_out/c3_iter_csr.err; echo "rc=$?" >> gpurun_out/c3_iter_csr.err
tail -n 3 gpurun_out/c3_tma.log
tail -n 2 gpurun_out/c3_rmat.err gpurun_out/c3_iter_sell.err
exit 0
