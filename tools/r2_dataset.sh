#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dataset.py -x -q -m gpu > gpurun_out/r2k_dataset_tests.log 2>&1
echo "tests rc=$?"; tail -25 gpurun_out/r2k_dataset_tests.log
timeout 600 python tools/bench_resample.py > gpurun_out/r2k_resample.json 2> gpurun_out/r2k_resample.err; tail -c 900 gpurun_out/r2k_resample.json; tail -3 gpurun_out/r2k_resample.err
