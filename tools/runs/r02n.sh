mkdir -p gpurun_out
XFM_GEMM_F32_DEEP=0 timeout 300 python tools/dev_gemm_f32epi.py ring2b > gpurun_out/r02n_f32epi_ring2.log 2>&1; tail -9 gpurun_out/r02n_f32epi_ring2.log
timeout 300 python tools/dev_gemm_f32epi.py ring4b > gpurun_out/r02n_f32epi_ring4.log 2>&1; echo "rc=$?"; tail -9 gpurun_out/r02n_f32epi_ring4.log
