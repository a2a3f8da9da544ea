#!/bin/bash
# round 2, call 18 (1 GPU): the pipelined kernel behind the plain SELL SpMV of stencil matrices -- full GPU suite,
# A/B of its variants (fp64 / fp32), cant-shaped records with the round's final library
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2r_tests.log
tail -n 4 gpurun_out/r2r_tests.log
timeout 200 python opencl-spmv-algorithms_b200/tools/bcast_probe.py --variants 1,3 > gpurun_out/r2r_sell_pipe_probe.json 2> gpurun_out/r2r_sell_pipe_probe.err; echo "probe rc=$?"
cat gpurun_out/r2r_sell_pipe_probe.json
for dt in f64 f32; do
  timeout 300 python bench.py --workload cant --dtype $dt --steps 100 > gpurun_out/r2r_bench_cant_$dt.json 2> gpurun_out/r2r_bench_cant_$dt.err; echo "cant $dt rc=$?"
done
python - <<'PY'
import json
for dt in ("f64", "f32"):
    d = json.loads(open(f"gpurun_out/r2r_bench_cant_{dt}.json").read().strip().splitlines()[-1])
    print("cant", dt, d["value"], {k: (v["ms"], v["frac_measured"]) for k, v in d["formats"].items()}, "e2e", d["e2e"]["value"], d["e2e"]["queue_per_format"])
PY
