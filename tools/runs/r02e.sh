mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02e_pytest.log; tail -25 gpurun_out/r02e_pytest.log
python bench.py --steps 10 --warmup 3 --no-eager --no-cpu --graph > gpurun_out/r02e_bench_graph.json 2> gpurun_out/r02e_bench_graph.err; tail -3 gpurun_out/r02e_bench_graph.err
for c in retrieval nlvr vqa; do python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r02e_bench_$c.json 2> gpurun_out/r02e_bench_$c.err; tail -2 gpurun_out/r02e_bench_$c.err; done
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02e_bench_graph.json")); print("pretrain", d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("cuda_graph"))
except Exception as e: print("pretrain ERR", e)
for c in ("retrieval","nlvr","vqa"):
    try:
        d=json.load(open("gpurun_out/r02e_bench_%s.json"%c)); print(c, d["ms_per_step"], d["eager"]["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"], d["kernels_per_step"])
    except Exception as e: print(c, "ERR", e)
PY
