#!/bin/bash
# round-1e GPU call 10 (2 GPUs): where does the ring's cross-GPU flag latency come from?
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516"
B="bench.py --gpus 2 --workload laplace-iter --iter-format sell --steps 100"
run() { name=$1; shift; env "$@" timeout 200 $T $B > gpurun_out/c10_$name.json 2> gpurun_out/c10_$name.err; }
run default A=1
run noflush B200_RING_FLUSH=0
run relaxed B200_RING_POLL=1
run nosleep B200_RING_SLEEP_NS=0
run relaxed_nosleep B200_RING_POLL=1 B200_RING_SLEEP_NS=0
exit 0
