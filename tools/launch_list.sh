#!/bin/bash
# Per-kernel device times of the gather stage (ncu, cold-cache, serialised): compare SHARES.
#   gpurun -- 'bash tools/launch_list.sh tag'
tag=${1:-x}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bin_|project_bin|sample_params|gather_" -c 60 --csv \
  --log-file gpurun_out/gather_launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-parity ${LL_ARGS} \
  > gpurun_out/gather_launches_$tag.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/gather_launches_$tag.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4].split("(")[0][:60]].append(float(r[-1]) / 1e3)
for k, v in agg.items():
    print(f"{k:62s} n={len(v):3d} mean {sum(v)/len(v):8.1f} us  min {min(v):8.1f}  max {max(v):8.1f}")
PY
