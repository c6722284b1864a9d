mkdir -p gpurun_out
for rep in 1 2; do
for tw in 1 0; do
XFM_TWIN_VISION=$tw timeout 600 python bench.py --steps 30 --warmup 5 --no-eager --no-cpu > gpurun_out/r03b_bench_tw${tw}_$rep.json 2> gpurun_out/r03b_bench_tw${tw}_$rep.err; echo "tw=$tw rep=$rep rc=$?"
done; done
python - <<PY
import json
for rep in (1,2):
  for tw in (1,0):
    d=json.load(open(f"gpurun_out/r03b_bench_tw{tw}_{rep}.json")); print("twin", tw, rep, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launch_sequence"]["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["launches_per_step"], d["step_ms"]["graph_resident"])
PY
