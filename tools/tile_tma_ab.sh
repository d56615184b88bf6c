#!/bin/bash
# A/B of the tensor-map tile staging (MVG_TILE_TMA=1 default / 0 = one cp.async.bulk per tile row) + gather-stage launch list
timeout 200 python -m pytest tests -m gpu -x -q -k "projection_bit_exact or full_size_properties or teacher_forced or query_sharded or cabi_driver or projattn_module or decoder_forward_api" 2>&1 | tail -2
for v in 1 0; do
MVG_TILE_TMA=$v timeout 100 python bench.py --no-cpu-baseline --no-parity --no-e2e --batch8 0 2>gpurun_out/quick_bench.err | tail -1 > gpurun_out/quick_bench_$v.json
python - <<PY
import json
d = json.load(open("gpurun_out/quick_bench_$v.json"))
print("TILE_TMA=$v value ms", round(d["ms_per_step"], 4), "q/s", round(d["value"]), "| gather stage ms", round(d["roofline"]["launch_ms"], 4))
PY
done
LL_ARGS="--batch8 0" bash tools/launch_list.sh tma | tail -8
