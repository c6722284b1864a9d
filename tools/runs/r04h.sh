mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r04h_pytest.log; tail -6 gpurun_out/r04h_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-eager --no-cpu > gpurun_out/r04h_bench.json 2> gpurun_out/r04h_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r04h_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r04h_bench.json")); print("pretrain", d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launch_sequence"]["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["launches_per_step"], d["fusion_layer"]["ms_per_step"], d["losses_last_step"])
PY
