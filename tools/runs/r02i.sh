mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02i_bench.err
timeout 600 ncu --profile-from-start off --clock-control none -k regex:gemm_tcgen05 --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/r02i_gemm_traffic.csv python tools/profile_step.py > gpurun_out/r02i_ncu_traffic.log 2>&1; echo "ncu rc=$?"
for c in retrieval nlvr vqa; do timeout 400 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r02i_bench_$c.json 2> gpurun_out/r02i_bench_$c.err; echo "$c rc=$?"; done
python - <<PY
import json
d=json.load(open("gpurun_out/r02i_bench.json")); print("pretrain", d["value"], d["ms_per_step"], d["e2e"], d["launch_sequence"], d["step_ms"])
for c in ("retrieval","nlvr","vqa"):
    try:
        d=json.load(open("gpurun_out/r02i_bench_%s.json"%c)); print(c, d["ms_per_step"], d["eager"]["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["achieved"], d["kernels_per_step"])
    except Exception as e: print(c, "ERR", e)
PY
