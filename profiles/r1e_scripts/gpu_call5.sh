#!/bin/bash
# round-1e GPU call 5 (2 GPUs): halo-limited fused exchange at world 2, banded weak scaling, COO overlap
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "overlap or packed or graph" > gpurun_out/c5_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c5_tests.log
timeout 300 $T bench.py --gpus 2 --workload laplace-iter --iter-format sell --steps 100 > gpurun_out/c5_iter_sell_n2.json 2> gpurun_out/c5_iter_sell_n2.err; echo "rc=$?" >> gpurun_out/c5_iter_sell_n2.err
timeout 300 $T bench.py --gpus 2 --no-cpu-baseline > gpurun_out/c5_banded_n2.json 2> gpurun_out/c5_banded_n2.err; echo "rc=$?" >> gpurun_out/c5_banded_n2.err
timeout 300 python bench.py --workload cant --no-cpu-baseline --no-e2e > gpurun_out/c5_bench_cant.json 2> gpurun_out/c5_bench_cant.err; echo "rc=$?" >> gpurun_out/c5_bench_cant.err
tail -n 4 gpurun_out/c5_tests.log; tail -n 3 gpurun_out/c5_iter_sell_n2.err gpurun_out/c5_banded_n2.err gpurun_out/c5_bench_cant.err
exit 0
