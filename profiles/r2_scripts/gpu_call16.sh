#!/bin/bash
# round 2, call 16 (1 GPU): persistent pipelined CSR nnz-split kernel -- parity (full GPU suite) and A/B on the
# 7-point Laplacian (fp64) and on R-MAT scale 24 (fp32)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2p_tests.log
tail -n 4 gpurun_out/r2p_tests.log
for pipe in 0 1; do
  B200_CSR_STREAM_PIPE=$pipe timeout 300 python bench.py --workload laplace-iter --steps 100 --no-cpu-baseline > gpurun_out/r2p_iter_n1_pipe$pipe.json 2> gpurun_out/r2p_iter_n1_pipe$pipe.err; echo "iter pipe=$pipe rc=$?"
  B200_CSR_STREAM_PIPE=$pipe timeout 400 python bench.py --workload rmat --steps 20 --rmat-sigmas "" --no-e2e --no-cpu-baseline > gpurun_out/r2p_rmat24_pipe$pipe.json 2> gpurun_out/r2p_rmat24_pipe$pipe.err; echo "rmat pipe=$pipe rc=$?"
done
python - <<'PY'
import json
for pipe in (0, 1):
    d = json.loads(open(f"gpurun_out/r2p_iter_n1_pipe{pipe}.json").read().strip().splitlines()[-1])
    it = d["iterated"]
    print("laplace pipe", pipe, it["ms_per_step"], it["split_ms"], it["nccl_allgather_formulation"]["ms_per_step"])
    d = json.loads(open(f"gpurun_out/r2p_rmat24_pipe{pipe}.json").read().strip().splitlines()[-1])
    print("rmat pipe", pipe, {k: (v["ms"], v["gflops"]) for k, v in d["formats"].items()})
PY
