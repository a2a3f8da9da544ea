#!/bin/bash
# round 2, call 14 (2 GPUs): real 2-GPU parity with the pipelined fused kernel, drivers (mcast halo fix), default bench at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_drivers.py -m gpu -x -q -k "real_gpus or iterated" > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2n_tests.log
tail -n 5 gpurun_out/r2n_tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29591 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err; echo "bench n2 rc=$?"
for sync in nccl mcast; do
  timeout 200 opencl-spmv-algorithms_b200/host/bin/sigma_c --synthetic laplace7:400x400x100 --iters 100 --gpus 2 --sync $sync --json > gpurun_out/r2n_driver_sigma_c_n2_$sync.json 2>/dev/null; cat gpurun_out/r2n_driver_sigma_c_n2_$sync.json
done
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2n_bench_n2.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", json.dumps({k: d["e2e"][k] for k in ("value", "ms_per_step", "one_queue", "two_queues", "queue_per_format", "link_gbs_each_way")}))
print("frac", d["roofline"]["per_format_frac"], "strong", d["strong"]["value"], d["strong"]["frac_measured_max_rank"])
it = d["iterated"]
print("iter", it["ms_per_step"], it["split_ms"], it["roofline"]["frac"], it["nvswitch_multicast"]["ms_per_step"], it["nccl_allgather_formulation"]["ms_per_step"], it["parity_ok"], it["e2e"]["value"])
PY
