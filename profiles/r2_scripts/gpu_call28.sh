#!/bin/bash
# round 2, call 28 (1 GPU): final-library records of the two workloads the driver does not run -- R-MAT scale 24 (full line:
# sigma sweep, L2-persistence A/B, e2e with a queue per format, CPU baseline) and the Laplacian iteration as the headline
mkdir -p gpurun_out
timeout 600 python bench.py --workload rmat --steps 20 > gpurun_out/r2ab_rmat24_n1.json 2> gpurun_out/r2ab_rmat24_n1.err; echo "rmat rc=$?"
timeout 400 python bench.py --workload laplace-iter --steps 200 > gpurun_out/r2ab_iter_n1.json 2> gpurun_out/r2ab_iter_n1.err; echo "iter rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ab_rmat24_n1.json").read().strip().splitlines()[-1])
print("rmat", d["value"], {k: v.get("gflops") for k, v in d["formats"].items()}, "e2e", d["e2e"]["value"], d["e2e"].get("queue_per_format"), "cpu", d["cpu_baseline"]["value"])
d = json.loads(open("gpurun_out/r2ab_iter_n1.json").read().strip().splitlines()[-1])
print("iter", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"])
PY
