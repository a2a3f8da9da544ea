#!/bin/bash
# round 2, call 23 (1 GPU): the whole GPU suite at the round's final library (new full-size parity tests), smoke(), the
# driver's bench command and its reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2w_tests.log
tail -n 4 gpurun_out/r2w_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2w_smoke.log 2>&1; tail -n 2 gpurun_out/r2w_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2w_bench_ref.json 2> gpurun_out/r2w_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2w_bench_n1.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/r2w_bench_ref.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ref", r["value"], "ratio", round(d["value"] / r["value"], 1), "e2e ratio", round(d["e2e"]["value"] / r["value"], 1))
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["traffic"], "cpu", d["cpu_baseline"]["value"], "iter", d["iterated"]["ms_per_step"], d["iterated"]["roofline"]["frac"], d["iterated"]["oracle_80cubed"]["ok"])
PY
