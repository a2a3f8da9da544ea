#!/bin/bash
# round 2, call 21 (1 GPU): what a plain streaming read reaches at cant-sized and large launch sizes
mkdir -p gpurun_out
timeout 300 opencl-spmv-algorithms_b200/tools/stream_probe > gpurun_out/r2u_stream_probe.json 2> gpurun_out/r2u_stream_probe.err; echo "probe rc=$?"
cat gpurun_out/r2u_stream_probe.json
