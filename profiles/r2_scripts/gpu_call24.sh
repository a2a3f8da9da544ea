#!/bin/bash
# round 2, call 24 (2 GPUs): b200_comm_allgather_bytes over real GPUs (multi-GPU parity suite) and the R-MAT e2e with the
# sharded x upload + NVLink all-gather
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2x_tests.log
tail -n 4 gpurun_out/r2x_tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29621 bench.py --gpus 2 --workload rmat --steps 10 --rmat-sigmas "" > gpurun_out/r2x_rmat24_n2.json 2> gpurun_out/r2x_rmat24_n2.err; echo "rmat n2 rc=$?"
tail -n 3 gpurun_out/r2x_rmat24_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2x_rmat24_n2.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("rmat n2 csr", d["value"], "e2e", e["value"], e["h2d_bytes_per_step"], "sharded", json.dumps({k: v for k, v in e["x_sharded_upload_nvlink_allgather"].items() if k != "what"}))
PY
