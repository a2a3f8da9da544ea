#!/bin/bash
# round 2, call 29 (2 GPUs): final validation -- the whole GPU suite (incl. the 2-GPU parity tests), smoke, and the driver's bench
# command at N = 1 and N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2ac_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2ac_tests.log
tail -n 4 gpurun_out/r2ac_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ac_bench_n1.json 2> gpurun_out/r2ac_bench_n1.err; echo "bench n1 rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29641 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2ac_bench_n2.json 2> gpurun_out/r2ac_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
for n in (1, 2):
    d = json.loads(open(f"gpurun_out/r2ac_bench_n{n}.json").read().strip().splitlines()[-1])
    it = d["iterated"]
    print("N", n, "value", d["value"], "e2e", d["e2e"]["value"], "strong", d["strong"]["value"], "iter", it["ms_per_step"], it["parity_ok"], it["nccl_allgather_formulation"]["ms_per_step"], "roofline", d["roofline"]["frac"])
PY
