#!/bin/bash
# the GPU checks written after round 2's GPU budget was spent (tests/pending_gpu_checks.py: never run on a B200)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/pending_gpu_checks.py -q -m gpu -p no:cacheprovider 2>&1 | tee gpurun_out/r3_pending.log | tail -15
