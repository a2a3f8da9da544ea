#!/bin/bash
# round-1e GPU call 11 (8 GPUs): iterated mode with the halo-limited exchange, banded weak scaling
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $T bench.py --gpus 8 --workload laplace-iter --iter-format sell --steps 100 > gpurun_out/c11_iter_sell_n8.json 2> gpurun_out/c11_iter_sell_n8.err; echo "rc=$?" >> gpurun_out/c11_iter_sell_n8.err
timeout 300 $T bench.py --gpus 8 --no-cpu-baseline > gpurun_out/c11_banded_n8.json 2> gpurun_out/c11_banded_n8.err; echo "rc=$?" >> gpurun_out/c11_banded_n8.err
tail -n 3 gpurun_out/c11_iter_sell_n8.err gpurun_out/c11_banded_n8.err
exit 0
