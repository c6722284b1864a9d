mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_graph_gpu.py tests/test_optim_gpu.py -m gpu -q 2>&1 | tail -3
for m in 1 0; do
XFM_MERGE_CROSS_KV=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-eager --no-cpu > gpurun_out/r05n_bench_m$m.json 2> gpurun_out/r05n_bench_m$m.err; echo "m=$m rc=$?"
done
python - <<PY
import json
for m in (1,0):
    d=json.load(open(f"gpurun_out/r05n_bench_m{m}.json")); print("merge", m, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launch_sequence"]["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["launches_per_step"], d["fusion_layer"]["ms_per_step"])
PY
