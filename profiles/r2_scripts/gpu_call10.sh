#!/bin/bash
# round 2, call 10 (4 GPUs): why is bench.py's iterated section slower than the probe and the C driver?  bisect.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 4 --master-port 29561 opencl-spmv-algorithms_b200/tools/iter_probe.py > gpurun_out/r2j_iter_probe_n4.json 2> gpurun_out/r2j_iter_probe_n4.err; echo "probe rc=$?"
cat gpurun_out/r2j_iter_probe_n4.json; tail -3 gpurun_out/r2j_iter_probe_n4.err
timeout 600 $TR --nproc-per-node 4 --master-port 29562 bench.py --gpus 4 --workload laplace-iter --steps 100 > gpurun_out/r2j_iter_n4.json 2> gpurun_out/r2j_iter_n4.err; echo "iter rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r2j_iter_n4.json'));i=d['iterated'];print(i['ms_per_step'],i['direct_launches_no_graph'],i['nvswitch_multicast']['ms_per_step'],i['split_ms'])"
timeout 300 opencl-spmv-algorithms_b200/host/bin/sigma_c --synthetic laplace7:400x400x200 --iters 100 --gpus 4 --json > gpurun_out/r2j_driver_sigma_c_n4.json 2>/dev/null; cat gpurun_out/r2j_driver_sigma_c_n4.json
timeout 300 opencl-spmv-algorithms_b200/host/bin/sigma_c --synthetic laplace7:400x400x200 --iters 100 --gpus 4 --sync mcast --json > gpurun_out/r2j_driver_sigma_c_n4_mcast.json 2>/dev/null; cat gpurun_out/r2j_driver_sigma_c_n4_mcast.json
