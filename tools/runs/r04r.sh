mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "vit_attention" 2>&1 | tail -15
timeout 600 python bench.py --config retrieval --steps 10 --warmup 3 --no-eager --no-cpu > gpurun_out/r04r_bench_retrieval.json 2> gpurun_out/r04r_bench_retrieval.err; echo "rc=$?"; tail -2 gpurun_out/r04r_bench_retrieval.err
python - <<PY
import json
d=json.load(open("gpurun_out/r04r_bench_retrieval.json")); print("retrieval", d["value"], d["unit"], d["ms_per_step"])
PY
