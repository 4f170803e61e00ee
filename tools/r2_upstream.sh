#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py::test_kfcv_detector_over_transformed_target_regressor tests/test_gpu_dataset.py -q -m gpu 2>&1 | tail -3
( timeout 900 python bench.py ) > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
python - <<PY
import json
l=json.loads(open("gpurun_out/r2l_bench.json").read().strip().splitlines()[-1])
print("c2", l["value"], l["roofline"]["frac"], "e2e", l["e2e"]["value"])
print(json.dumps(l["other_configs"].get("upstream_of_x"), indent=None)[:1500])
PY
tail -2 gpurun_out/r2l_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resample_kernel -c 1 -f -o gpurun_out/r2l_resample python tools/bench_resample.py --machines 16 > gpurun_out/r2l_ncu.log 2>&1
ncu -i gpurun_out/r2l_resample.ncu-rep --page raw --csv > gpurun_out/r2l_resample_raw.csv 2>/dev/null
ncu -i gpurun_out/r2l_resample.ncu-rep --page details > gpurun_out/r2l_resample_details.txt 2>/dev/null
rm -f gpurun_out/r2l_resample.ncu-rep
grep -E "Duration|Executed Ipc Active|Issue Slots Busy|DRAM Throughput|Memory Throughput|Achieved Occupancy|Theoretical Occupancy|Registers Per|L2 Hit|Stall" gpurun_out/r2l_resample_details.txt | head -20
