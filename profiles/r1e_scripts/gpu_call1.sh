#!/bin/bash
# round-1e GPU call 1: parity of the new kernel variants, the TMA kernel in its own process, sweeps
mkdir -p gpurun_out
S=opencl-spmv-algorithms_b200/tools/sweep_variants.py
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c1_parity.log 2>&1; echo "rc=$?" >> gpurun_out/c1_parity.log
timeout 300 python -m pytest tests/test_gpu_tma.py -m gpu -x -q > gpurun_out/c1_tma.log 2>&1; echo "rc=$?" >> gpurun_out/c1_tma.log
timeout 300 python $S --workload cant --dtype f64 --no-tma --out gpurun_out/sweep_cant_f64.json 2> gpurun_out/sweep_cant_f64.log; echo "rc=$?" >> gpurun_out/sweep_cant_f64.log
timeout 300 python $S --workload cant --dtype f32 --no-tma --out gpurun_out/sweep_cant_f32.json 2> gpurun_out/sweep_cant_f32.log; echo "rc=$?" >> gpurun_out/sweep_cant_f32.log
timeout 300 python $S --workload banded --dtype f32 --no-tma --out gpurun_out/sweep_banded_f32.json 2> gpurun_out/sweep_banded_f32.log; echo "rc=$?" >> gpurun_out/sweep_banded_f32.log
timeout 300 python $S --workload banded --dtype f64 --no-tma --out gpurun_out/sweep_banded_f64.json 2> gpurun_out/sweep_banded_f64.log; echo "rc=$?" >> gpurun_out/sweep_banded_f64.log
for w in cant banded; do for d in f64 f32; do
timeout 200 python $S --workload $w --dtype $d --tma-only --out gpurun_out/sweep_tma_${w}_$d.json 2> gpurun_out/sweep_tma_${w}_$d.log; echo "rc=$?" >> gpurun_out/sweep_tma_${w}_$d.log
done; done
timeout 300 python bench.py --workload cant --steps 200 --warmup 5 > gpurun_out/c1_bench_cant.json 2> gpurun_out/c1_bench_cant.err; echo "rc=$?" >> gpurun_out/c1_bench_cant.err
timeout 300 python -m pytest tests/test_gpu_synth.py -m gpu -x -q > gpurun_out/c1_synth.log 2>&1; echo "rc=$?" >> gpurun_out/c1_synth.log
tail -3 gpurun_out/c1_parity.log gpurun_out/c1_tma.log gpurun_out/c1_synth.log gpurun_out/c1_bench_cant.err
