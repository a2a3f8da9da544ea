#!/bin/bash
# round 2, call 20 (1 GPU): bulk L2 prefetch (cp.async.bulk.prefetch.L2) in the CSR vector and ELL row-major kernels --
# parity, then A/B on the cant-shaped workload (cold by rotation), fp32 and fp64
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2t_tests.log
tail -n 3 gpurun_out/r2t_tests.log
for dt in f32 f64; do for pf in 0 1; do
  B200_L2_PREFETCH=$pf timeout 300 python bench.py --workload cant --dtype $dt --steps 100 --no-e2e --no-cpu-baseline > gpurun_out/r2t_cant_${dt}_pf$pf.json 2> gpurun_out/r2t_cant_${dt}_pf$pf.err; echo "cant $dt pf=$pf rc=$?"
done; done
python - <<'PY'
import json
for dt in ("f32", "f64"):
    for pf in (0, 1):
        d = json.loads(open(f"gpurun_out/r2t_cant_{dt}_pf{pf}.json").read().strip().splitlines()[-1])
        print("cant", dt, "pf", pf, d["value"], {k: (round(v["ms"] * 1e3, 2), v["frac_measured"]) for k, v in d["formats"].items() if k in ("coo", "csr", "ell", "sell", "cmrs")},
              "overlap off", {k: round(v["ms"] * 1e3, 2) for k, v in d["launch_overlap_off"]["formats"].items()})
PY
