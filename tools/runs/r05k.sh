mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py --no-eager --no-cpu > gpurun_out/r05k_bench.json 2> gpurun_out/r05k_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r05k_bench.json")); print("pretrain", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["launch_sequence"]["ms_per_step"], d["roofline"]["achieved"], d["fusion_layer"]["ms_per_step"])
PY
