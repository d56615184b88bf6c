#!/bin/bash
# Per-kernel times of one eager step WITHOUT ncu's cache flush (--cache-control none): their sum against the
# graph-replay step time bounds what launch gaps / dependency latency cost.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 240 -c 70 --csv --log-file gpurun_out/warm_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-parity --batch8 0 > gpurun_out/warm_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/warm_launches.csv")) if len(r) > 5 and r[0].isdigit()]
names = [r[4].split("(")[0].replace("void ", "")[:56] for r in rows]
t = [float(r[-1]) / 1e3 for r in rows]
st = [i for i, n in enumerate(names) if "pack_cameras" in n]
lo, hi = (st[0], st[1]) if len(st) > 1 else (0, len(rows))
agg = collections.OrderedDict()
for n, x in zip(names[lo:hi], t[lo:hi]):
    agg.setdefault(n, []).append(x)
for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
    print(f"{k:58s} n={len(v):3d} mean {sum(v)/len(v):8.1f} us  total {sum(v):8.1f}")
print("one step:", hi - lo, "launches, summed kernel time %.1f us" % sum(t[lo:hi]))
PY
