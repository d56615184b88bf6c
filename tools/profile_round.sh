#!/bin/bash
# Round evidence (run on the GPU box): tests, bench line, ncu launch list, ncu --set full of one step.
#   gpurun -- 'bash tools/profile_round.sh r1'
tag=${1:-r1}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu_$tag.log)
tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
cat gpurun_out/bench_$tag.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${LIST_SKIP:-700} -c 400 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-parity \
  > gpurun_out/launches_$tag.log 2>&1
# one decoder layer + the per-call kernels (camera packing, pyramid hand-off, value GEMM): ~20 launches, ~40 replays each.
# Keep the report small: gpurun merges at most 64 MiB back.
timeout 600 ncu --set full --clock-control none -k regex:"^(pack_cameras|pyramid_to_cl|linear_tcgen05|add_cast|project_bin|bin_scan|bin_scatter|sample_params|gather_tiles|gather_direct|masked_view_mean|ffn_chain|class_head|select_pad|offset_chain|offsets_dlt)" -s ${NCU_SKIP:-186} -c ${NCU_COUNT:-18} -o gpurun_out/step_$tag \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-parity > gpurun_out/step_$tag.log 2>&1
tail -2 gpurun_out/step_$tag.log | cut -c1-200
ls -la gpurun_out/
