#!/bin/bash
# round 2, call 13 (1 GPU): full GPU test suite with the pipelined fused kernel, A/B of its variants,
# the driver's own bench command (queue-per-format e2e, x bytes counted over the columns read)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_tests.log
tail -n 4 gpurun_out/r2m_tests.log
timeout 200 python opencl-spmv-algorithms_b200/tools/bcast_probe.py > gpurun_out/r2m_bcast_probe.json 2> gpurun_out/r2m_bcast_probe.err; echo "probe rc=$?"
cat gpurun_out/r2m_bcast_probe.json
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2m_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", json.dumps({k: d["e2e"][k] for k in ("value", "ms_per_step", "one_queue", "two_queues", "queue_per_format", "link_gbs_each_way")}))
print("f64 e2e", d["f64"]["e2e"]["value"], d["f64"]["e2e"]["queue_per_format"])
print("iter", d["iterated"]["ms_per_step"], d["iterated"]["split_ms"], d["iterated"]["roofline"]["frac"], d["iterated"].get("oracle_80cubed", {}).get("ok"))
PY
