#!/bin/bash
# round 2, call 17 (8 GPUs): the driver's own bench command at N = 8 with the pipelined fused kernel and the
# queue-per-format e2e; the C drivers over 8 devices (NCCL / multicast hand-over, halo rows only)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2q_bench_n8.json 2> gpurun_out/r2q_bench_n8.err; echo "bench n8 rc=$?"
for sync in nccl mcast; do
  timeout 200 opencl-spmv-algorithms_b200/host/bin/sigma_c --synthetic laplace7:400x400x400 --iters 100 --gpus 8 --sync $sync --json > gpurun_out/r2q_driver_sigma_c_n8_$sync.json 2>/dev/null; cat gpurun_out/r2q_driver_sigma_c_n8_$sync.json
done
timeout 200 opencl-spmv-algorithms_b200/host/bin/csr --synthetic laplace7:400x400x400 --iters 100 --gpus 8 --json > gpurun_out/r2q_driver_csr_n8.json 2>/dev/null; cat gpurun_out/r2q_driver_csr_n8.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2q_bench_n8.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", json.dumps({k: d["e2e"][k] for k in ("value", "ms_per_step", "one_queue", "two_queues", "queue_per_format", "link_gbs_each_way")}))
print("frac", d["roofline"]["per_format_frac"], "strong", d["strong"]["value"], d["strong"]["frac_measured_max_rank"], "f64", d["f64"]["value"], d["f64"]["e2e"]["value"])
it = d["iterated"]
print("iter", it["ms_per_step"], it["fused_with_nccl_allreduce"], it["split_ms"], it["roofline"]["frac"], it["nvswitch_multicast"]["ms_per_step"], it["nccl_allgather_formulation"]["ms_per_step"], it["parity_ok"], it["e2e"]["value"])
PY
