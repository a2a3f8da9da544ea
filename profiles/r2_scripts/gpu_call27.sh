#!/bin/bash
# round 2, call 27 (1 GPU): what the x gather costs by itself -- a one-shot kernel over the same 1.07 GB of (index, value)
# pairs as the banded bench matrix, with and without the gather
mkdir -p gpurun_out
timeout 200 opencl-spmv-algorithms_b200/tools/stream_probe --gather > gpurun_out/r2aa_gather_probe.json 2> gpurun_out/r2aa_gather_probe.err; echo "probe rc=$?"
cat gpurun_out/r2aa_gather_probe.json
