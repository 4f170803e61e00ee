#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
# LSTM fit of a c3-sized topology (60 tags, lookback 16): wall time per step vs the sum of kernel durations
( python tools/bench_lstm.py --rows 4000 --tags 60 --lookback 16 --fit-jobs 4 --fit-rows 3216 --cpu-windows 8 ) > gpurun_out/r2g_lstm_small.json 2> gpurun_out/r2g_lstm_small.err
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2g_lstm_small_launches.csv \
   python tools/bench_lstm.py --rows 4000 --tags 60 --lookback 16 --fit-jobs 4 --fit-rows 336 --cpu-windows 8 ) > gpurun_out/r2g_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.DictReader(l for l in open("gpurun_out/r2g_lstm_small_launches.csv") if l.startswith('"')))
agg = collections.OrderedDict()
for r in rows:
    k = r["Kernel Name"].split("(")[0][-60:]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r["Metric Value"])
tot = sum(v[1] for v in agg.values())
with open("gpurun_out/r2g_lstm_small_summary.txt", "w") as f:
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{ns/1e3:10.1f} us {n:5d} x {ns/n/1e3:7.2f} us  {k}\n")
    f.write(f"total {tot/1e3:.1f} us over {len(rows)} launches\n")
print(open("gpurun_out/r2g_lstm_small_summary.txt").read())
PY
cat gpurun_out/r2g_lstm_small.json; tail -3 gpurun_out/r2g_lstm_small.err
