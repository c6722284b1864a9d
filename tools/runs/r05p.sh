mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r05p_bench.json 2> gpurun_out/r05p_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r05p_bench.json")); print("pretrain", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["launch_sequence"]["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["launches_per_step"], d["fusion_layer"]["ms_per_step"], d["fusion_layer"]["tflops_algorithmic"], d["gpu_launches"], d["clocks"]["sm_mhz"], d["gpu_eager_baseline"]["bf16_autocast"]["value"])
PY
