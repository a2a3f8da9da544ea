#!/bin/bash
# round 2, call 22 (1 GPU): persistent row-pipelined CSR / ELL kernel (csr_rows_pipe_kernel) -- parity, then A/B on
# the cant-shaped workload (fp32, fp64; off / 2 / 3 blocks per SM) and on the banded headline (forced)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_drivers.py -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2v_tests.log
tail -n 3 gpurun_out/r2v_tests.log
for dt in f32 f64; do for pf in 0 2 3; do
  B200_ROWS_PIPE=$pf timeout 300 python bench.py --workload cant --dtype $dt --steps 100 --no-e2e --no-cpu-baseline > gpurun_out/r2v_cant_${dt}_rp$pf.json 2> gpurun_out/r2v_cant_${dt}_rp$pf.err; echo "cant $dt rp=$pf rc=$?"
done; done
for pf in 0 2 3; do
  B200_ROWS_PIPE=$pf timeout 300 python bench.py --steps 50 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2v_banded_rp$pf.json 2> gpurun_out/r2v_banded_rp$pf.err; echo "banded rp=$pf rc=$?"
done
python - <<'PY'
import json
for dt in ("f32", "f64"):
    for pf in (0, 2, 3):
        d = json.loads(open(f"gpurun_out/r2v_cant_{dt}_rp{pf}.json").read().strip().splitlines()[-1])
        print("cant", dt, "rp", pf, d["value"], {k: (round(v["ms"] * 1e3, 2), v["frac_measured"]) for k, v in d["formats"].items() if k in ("coo", "csr", "ell", "sell", "cmrs")},
              "overlap off", {k: round(v["ms"] * 1e3, 2) for k, v in d["launch_overlap_off"]["formats"].items()})
for pf in (0, 2, 3):
    d = json.loads(open(f"gpurun_out/r2v_banded_rp{pf}.json").read().strip().splitlines()[-1])
    print("banded rp", pf, d["value"], {k: (v["ms"], v["frac_measured"]) for k, v in d["formats"].items() if k in ("csr", "ell")})
PY
