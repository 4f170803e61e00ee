#!/bin/bash
# the concurrent-stream c3 build (all topologies 20..100 tags, FF + LSTM) under compute-sanitizer memcheck at reduced rows
mkdir -p gpurun_out
( timeout 1500 compute-sanitizer --tool memcheck --print-limit 8 python bench.py --config c3 --machines 32 --rows 6000 --streams 8 --steps 1 --warmup 0 --cpu-seconds 0 --no-extras ) > gpurun_out/r2i_c3_memcheck.log 2>&1
echo "rc=$?"
grep -m1 -B2 -A30 "Invalid\|out of bounds\|misaligned" gpurun_out/r2i_c3_memcheck.log | head -70
grep "ERROR SUMMARY" gpurun_out/r2i_c3_memcheck.log; tail -c 600 gpurun_out/r2i_c3_memcheck.log
