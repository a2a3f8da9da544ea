#!/bin/bash
# round 2, call 11 (4 GPUs): after removing OMP_PROC_BIND from bench.py and re-doing the multicast hand-over
# (2 multicast operations per rank and step instead of 33): probe, laplace-iter arm, default line, drivers, tests
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_drivers.py -m gpu -x -q -k "real_gpus or iterated" > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2k_tests.log
tail -3 gpurun_out/r2k_tests.log
timeout 600 $TR --nproc-per-node 4 --master-port 29571 opencl-spmv-algorithms_b200/tools/iter_probe.py > gpurun_out/r2k_iter_probe_n4.json 2> gpurun_out/r2k_iter_probe_n4.err; echo "probe rc=$?"
cat gpurun_out/r2k_iter_probe_n4.json
timeout 600 $TR --nproc-per-node 4 --master-port 29572 bench.py --gpus 4 --workload laplace-iter --steps 100 > gpurun_out/r2k_iter_n4.json 2> gpurun_out/r2k_iter_n4.err; echo "iter rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r2k_iter_n4.json'));i=d['iterated'];print(i['ms_per_step'],i['direct_launches_no_graph'],i['nvswitch_multicast']['ms_per_step'],i['split_ms'])"
timeout 600 $TR --nproc-per-node 4 --master-port 29573 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2k_bench_n4.json 2> gpurun_out/r2k_bench_n4.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r2k_bench_n4.json'));i=d['iterated'];print(d['value'],d['e2e']['value'],d['strong']['value'],i['ms_per_step'],i['nvswitch_multicast']['ms_per_step'],i['split_ms'],i['parity_ok'])"
for sync in nccl mcast; do
timeout 300 opencl-spmv-algorithms_b200/host/bin/sigma_c --synthetic laplace7:400x400x200 --iters 100 --gpus 4 --sync $sync --json > gpurun_out/r2k_driver_sigma_c_n4_$sync.json 2>/dev/null; cat gpurun_out/r2k_driver_sigma_c_n4_$sync.json
done
