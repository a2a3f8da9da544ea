#!/bin/bash
# round-1e GPU call 8 (2 GPUs): ring exchange, reworked (prefetch before the wait, one fence per block)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514"
timeout 300 python -m pytest tests/test_gpu_synth.py -m gpu -q -k "ring or halo or fused" > gpurun_out/c8_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c8_tests.log
timeout 200 python bench.py --workload laplace-iter --iter-format sell --steps 100 > gpurun_out/c8_iter_sell_n1.json 2> gpurun_out/c8_iter_sell_n1.err; echo "rc=$?" >> gpurun_out/c8_iter_sell_n1.err
timeout 300 $T bench.py --gpus 2 --workload laplace-iter --iter-format sell --steps 100 > gpurun_out/c8_iter_sell_n2.json 2> gpurun_out/c8_iter_sell_n2.err; echo "rc=$?" >> gpurun_out/c8_iter_sell_n2.err
B200_RING_NC=1 timeout 200 python bench.py --workload laplace-iter --iter-format sell --steps 100 > gpurun_out/c8_iter_sell_n1_nc.json 2> gpurun_out/c8_iter_sell_n1_nc.err
tail -n 6 gpurun_out/c8_tests.log; tail -n 2 gpurun_out/c8_iter_sell_n2.err gpurun_out/c8_iter_sell_n1.err
exit 0
