#!/bin/bash
# round 2, call 31 (4 GPUs): R-MAT scale 24 over 4 row blocks -- the N = 4 point of the e2e table (whole x per rank vs sharded
# upload + NVLink all-gather)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 150 $TR --nproc-per-node 4 --master-port 29651 bench.py --gpus 4 --workload rmat --steps 10 --rmat-sigmas "" > gpurun_out/r2ae_rmat24_n4.json 2> gpurun_out/r2ae_rmat24_n4.err; echo "rmat n4 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2ae_rmat24_n4.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("rmat n4 csr", d["value"], "e2e", e["value"], "sharded", e["x_sharded_upload_nvlink_allgather"]["value"], e["x_sharded_upload_nvlink_allgather"]["gathered_x_equals_host_x"])
PY
