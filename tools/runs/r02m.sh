mkdir -p gpurun_out
XFM_GEMM_F32_DEEP=0 timeout 300 python tools/dev_gemm_f32epi.py ring2 > gpurun_out/r02m_f32epi_ring2.log 2>&1; tail -6 gpurun_out/r02m_f32epi_ring2.log
timeout 300 python tools/dev_gemm_f32epi.py ring4 > gpurun_out/r02m_f32epi_ring4.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r02m_f32epi_ring4.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02m_pytest.log; tail -4 gpurun_out/r02m_pytest.log
timeout 300 python tools/dev_kernels.py ln > gpurun_out/r02m_dev_ln.log 2>&1; tail -4 gpurun_out/r02m_dev_ln.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-eager --no-cpu > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; echo "bench rc=$?"
XFM_GEMM_F32_DEEP=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-eager --no-cpu > gpurun_out/r02m_bench_ring2.json 2> gpurun_out/r02m_bench_ring2.err; echo "bench rc=$?"
python - <<PY
import json
for f in ("r02m_bench", "r02m_bench_ring2"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launch_sequence"]["ms_per_step"], d["roofline"]["achieved"])
PY
