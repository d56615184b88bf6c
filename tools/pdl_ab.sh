timeout 400 python -m pytest tests -m gpu -x -q -k "elementwise or decoder_forward_api or cabi_driver or query_sharded or teacher_forced or tcgen05_matches or select_pad or linear_tcgen05 or ffn_chain or offset_chain or projection_bit_exact or full_size_properties" 2>&1 | tail -4
for v in 1 0 1 0; do
MVG_PDL=$v timeout 200 python bench.py --no-cpu-baseline --no-parity --batch8 0 2>gpurun_out/quick_bench.err | tail -1 > gpurun_out/quick_bench.json
python - <<PY
import json
d = json.load(open("gpurun_out/quick_bench.json"))
print("PDL=$v value ms", round(d["ms_per_step"], 4), "q/s", round(d["value"]), "| e2e ms", round(d["e2e"]["ms_per_step"], 4), "| gather stage ms", round(d["roofline"]["launch_ms"], 4), "| all-selected ms", round(d["all_queries_selected"]["ms_per_step"],4))
PY
done
