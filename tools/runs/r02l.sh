mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02l_pytest.log; tail -4 gpurun_out/r02l_pytest.log
timeout 300 python tools/dev_kernels.py ln > gpurun_out/r02l_dev_ln.log 2>&1; tail -4 gpurun_out/r02l_dev_ln.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-eager --no-cpu > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"layernorm_bwd|layerscale_bwd" -c 6 -o gpurun_out/r02l_ln python tools/profile_step.py > gpurun_out/r02l_ncu_ln.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02l_launches_pretrain.csv python tools/profile_step.py > gpurun_out/r02l_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r02l_bench.json")); print("pretrain", d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launch_sequence"]["ms_per_step"], d["step_ms"])
PY
