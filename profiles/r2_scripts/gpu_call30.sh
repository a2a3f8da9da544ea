#!/bin/bash
# round 2, call 30 (1 GPU): the GPU suite after the drivers' new text parse (mapped file, own number scanners), and one
# reference-style run of every driver on the cant-shaped file with the load timing on stderr
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2ad_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2ad_tests.log
tail -n 4 gpurun_out/r2ad_tests.log
make databases > /dev/null 2>&1 && make all > /dev/null 2>&1
for f in csr sigma_c; do
  B200_PARSE_TIMING=1 ./bin/$f > gpurun_out/r2ad_driver_$f.out 2> gpurun_out/r2ad_driver_$f.err; echo "$f rc=$?"
  grep -E "result is|PERFORMANCE|read_entries" gpurun_out/r2ad_driver_$f.out gpurun_out/r2ad_driver_$f.err | head -8
done
nproc
