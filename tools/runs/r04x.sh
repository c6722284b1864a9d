mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x -k "nlvr or retrieval or vqa" 2>&1 | tail -5
timeout 600 python bench.py --config nlvr --steps 10 --warmup 3 --no-eager --no-cpu > gpurun_out/r04x_bench_nlvr.json 2> gpurun_out/r04x_bench_nlvr.err; echo "rc=$?"; tail -2 gpurun_out/r04x_bench_nlvr.err
python - <<PY
import json
d=json.load(open("gpurun_out/r04x_bench_nlvr.json")); print("nlvr", d["value"], d["unit"], d["ms_per_step"])
PY
