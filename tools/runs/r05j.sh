mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r05j_bench.json 2> gpurun_out/r05j_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r05j_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r05j_bench.json")); print("pretrain", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["launch_sequence"]["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["fusion_layer"]["ms_per_step"], d["fusion_layer"]["tflops_algorithmic"], d["gpu_launches"], d["clocks"], d["gpu_eager_baseline"]["bf16_autocast"]["value"], d["cpu_baseline"]["value"])
PY
