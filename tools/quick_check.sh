#!/bin/bash
# Fast loop on the GPU box: a test subset, a short bench line, per-kernel times of one eager step.
#   gpurun -- 'bash tools/quick_check.sh "<pytest -k expr>" [ncu]'
mkdir -p gpurun_out
if [ -n "$1" ]; then timeout 400 python -m pytest tests -m gpu -x -q -k "$1" 2>&1 | tail -4; fi
timeout 200 python bench.py --no-cpu-baseline --no-parity --batch8 0 2>gpurun_out/quick_bench.err | tail -1 > gpurun_out/quick_bench.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/quick_bench.json"))
print("value ms", round(d["ms_per_step"], 4), "q/s", round(d["value"]), "| e2e ms", round(d["e2e"]["ms_per_step"], 4),
      "| launches", d["gpu_launches"], "| gather stage ms", round(d["roofline"]["launch_ms"], 4), "| clocks", d["clocks"]["sm_mhz"])
PY
if [ "$2" = ncu ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 240 -c 90 --csv --log-file gpurun_out/quick_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-parity --batch8 0 > gpurun_out/quick_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/quick_launches.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    k = r[4].split("(")[0].replace("void ", "")[:56]
    agg.setdefault(k, []).append(float(r[-1]) / 1e3)
for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
    print(f"{k:58s} n={len(v):3d} mean {sum(v)/len(v):8.1f} us")
PY
fi
