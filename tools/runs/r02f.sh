mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-cpu --graph > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -2 gpurun_out/r02f_bench.err
python tools/gemm_sweep.py > gpurun_out/r02f_gemm_sweep.log 2>&1; tail -3 gpurun_out/r02f_gemm_sweep.log
NCU="ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
XFM_PROFILE_CONFIG=retrieval $NCU --log-file gpurun_out/r02f_launches_retrieval.csv python tools/profile_step.py > gpurun_out/r02f_ncu_retrieval.log 2>&1
XFM_PROFILE_CONFIG=pretrain $NCU --log-file gpurun_out/r02f_launches_pretrain.csv python tools/profile_step.py > gpurun_out/r02f_ncu_pretrain.log 2>&1
python tools/summarize_launches.py gpurun_out/r02f_launches_retrieval.csv | head -40
python tools/summarize_launches.py gpurun_out/r02f_launches_pretrain.csv | head -30
python - <<PY
import json
d=json.load(open("gpurun_out/r02f_bench.json")); print("pretrain", d["ms_per_step"], d["e2e"]["ms_per_step"], d["step_ms"], d.get("cuda_graph"))
PY
