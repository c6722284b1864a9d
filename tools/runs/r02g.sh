mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02g_pytest.log; tail -8 gpurun_out/r02g_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu --no-eager --graph > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; tail -2 gpurun_out/r02g_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02g_bench.json")); print("pretrain", d["ms_per_step"], d["e2e"]["ms_per_step"], d["step_ms"], d.get("cuda_graph"), d["gpu_launches"])
PY
