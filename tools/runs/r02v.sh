mkdir -p gpurun_out
cp xfm_b200/libxfm_b200.so /tmp/db1.so
echo "== DB=1 (TMEM double buffer)"; timeout 300 python tools/gemm_case.py plain plain_wide plain_aux gelu_noaux gelu dgelu proj fc2 ffn_out > gpurun_out/r02v_cases_db1.jsonl 2>&1; echo "rc=$?"; cat gpurun_out/r02v_cases_db1.jsonl
cp xfm_b200/libxfm_b200_db0.so xfm_b200/libxfm_b200.so
echo "== DB=0"; timeout 300 python tools/gemm_case.py plain plain_wide plain_aux gelu_noaux gelu dgelu proj fc2 ffn_out > gpurun_out/r02v_cases_db0.jsonl 2>&1; echo "rc=$?"; cat gpurun_out/r02v_cases_db0.jsonl
cp /tmp/db1.so xfm_b200/libxfm_b200.so
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -5
