mkdir -p gpurun_out
timeout 300 python tools/time_vit577.py 32 | tee gpurun_out/r05d_vit577.jsonl
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r05d_pytest.log; tail -3 gpurun_out/r05d_pytest.log
timeout 900 python bench.py > gpurun_out/r05d_bench.json 2> gpurun_out/r05d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r05d_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r05d_bench.json")); print("pretrain", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["launch_sequence"]["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["fusion_layer"]["ms_per_step"], d["fusion_layer"]["tflops_algorithmic"], d["gpu_launches"], d["clocks"])
PY
NCU="ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv"
XFM_PROFILE_CONFIG=pretrain timeout 600 $NCU --log-file gpurun_out/r05d_launches_pretrain.csv python tools/profile_step.py > gpurun_out/r05d_ncu_pretrain.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r05d_launches_pretrain.csv > gpurun_out/r05d_launches_pretrain_summary.txt; head -30 gpurun_out/r05d_launches_pretrain_summary.txt
