#!/bin/bash
# round 2, call 15 (4 GPUs): real 2- and 4-GPU parity with the pipelined fused kernel, default bench at N = 4, drivers
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2o_tests.log
tail -n 5 gpurun_out/r2o_tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 4 --master-port 29601 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2o_bench_n4.json 2> gpurun_out/r2o_bench_n4.err; echo "bench n4 rc=$?"
for sync in nccl mcast; do
  timeout 200 opencl-spmv-algorithms_b200/host/bin/sigma_c --synthetic laplace7:400x400x200 --iters 100 --gpus 4 --sync $sync --json > gpurun_out/r2o_driver_sigma_c_n4_$sync.json 2>/dev/null; cat gpurun_out/r2o_driver_sigma_c_n4_$sync.json
done
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2o_bench_n4.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", json.dumps({k: d["e2e"][k] for k in ("value", "ms_per_step", "one_queue", "two_queues", "queue_per_format", "link_gbs_each_way")}))
print("frac", d["roofline"]["per_format_frac"], "strong", d["strong"]["value"], d["strong"]["frac_measured_max_rank"])
it = d["iterated"]
print("iter", it["ms_per_step"], it["fused_with_nccl_allreduce"], it["split_ms"], it["roofline"]["frac"], it["nvswitch_multicast"]["ms_per_step"], it["nccl_allgather_formulation"]["ms_per_step"], it["parity_ok"], it["e2e"]["value"])
PY
