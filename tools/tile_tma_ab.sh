#!/bin/bash
# A/B of the tensor-map tile staging (MVG_TILE_TMA=1 default / 0 = one cp.async.bulk per tile row)
timeout 400 python -m pytest tests -m gpu -x -q -k "projection_bit_exact or full_size_properties or teacher_forced or query_sharded or cabi_driver or projattn_module or layer_parity_other or decoder_forward_api" 2>&1 | tail -4
for q in 1024 128; do for v in 1 0; do
MVG_TILE_TMA=$v timeout 200 python bench.py --no-cpu-baseline --no-parity --no-e2e --batch8 0 --queries $q 2>gpurun_out/quick_bench.err | tail -1 > gpurun_out/quick_bench.json
python - <<PY
import json
d = json.load(open("gpurun_out/quick_bench.json"))
print("Q=$q TILE_TMA=$v value ms", round(d["ms_per_step"], 4), "q/s", round(d["value"]), "| gather stage ms", round(d["roofline"]["launch_ms"], 4))
PY
done; done
