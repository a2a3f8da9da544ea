#!/bin/bash
# round 2, call 25 (8 GPUs): R-MAT scale 24 over 8 row blocks -- e2e with every rank uploading the whole x vs the
# sharded upload + NVLink all-gather (b200_comm_allgather_bytes)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --workload rmat --steps 10 --rmat-sigmas "" > gpurun_out/r2y_rmat24_n8.json 2> gpurun_out/r2y_rmat24_n8.err; echo "rmat n8 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2y_rmat24_n8.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("rmat n8 csr", d["value"], {k: v["gflops"] for k, v in d["formats"].items()}, "e2e", e["value"], e["h2d_bytes_per_step"], "sharded", json.dumps({k: v for k, v in e["x_sharded_upload_nvlink_allgather"].items() if k != "what"}))
PY
