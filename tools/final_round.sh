#!/bin/bash
# Last evidence pass of the round on a tight GPU budget: bench line, tests (two of the six size configs: the bench's
# own parity block covers configs[1] at full size), launch list, ncu --set full of the gather stage's kernels.
tag=${1:-r2}
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
cut -c1-400 gpurun_out/bench_$tag.json
(timeout 400 python -m pytest tests -m gpu -q -k "not parity_at_baseline_size or c1_q128 or c5_shelf" > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu_$tag.log)
tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 300 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-parity --batch8 0 \
  > gpurun_out/launches_$tag.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:"^(project_bin|bin_scan|bin_scatter|sample_params|gather_tiles|gather_direct)" -s 72 -c 6 -o gpurun_out/step_${tag}g \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-parity --batch8 0 > gpurun_out/step_${tag}g.log 2>&1
ls -la gpurun_out/ | tail -8
