mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02x_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02x_bench.json")); print("pretrain", d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launch_sequence"]["ms_per_step"], d["roofline"]["achieved"], d["fusion_layer"])
PY
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02x_launches_pretrain.csv python tools/profile_step.py > gpurun_out/r02x_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"gemm_tcgen05_pair_kernel" -s 40 -c 14 -o gpurun_out/r02x_gemm_pair python tools/profile_step.py > gpurun_out/r02x_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
