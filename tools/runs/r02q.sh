mkdir -p gpurun_out
XFM_GEMM_EPI_INTERLEAVE=0 timeout 300 python tools/dev_gemm_f32epi.py il0 > gpurun_out/r02q_f32epi_il0.log 2>&1; echo "rc=$?"; tail -9 gpurun_out/r02q_f32epi_il0.log
XFM_GEMM_EPI_INTERLEAVE=1 timeout 300 python tools/dev_gemm_f32epi.py il1 > gpurun_out/r02q_f32epi_il1.log 2>&1; echo "rc=$?"; tail -9 gpurun_out/r02q_f32epi_il1.log
XFM_GEMM_EPI_INTERLEAVE=0 timeout 300 python tools/dev_gemm_pair.py > gpurun_out/r02q_pair_il0.log 2>&1; tail -7 gpurun_out/r02q_pair_il0.log
XFM_GEMM_EPI_INTERLEAVE=1 timeout 300 python tools/dev_gemm_pair.py > gpurun_out/r02q_pair_il1.log 2>&1; tail -7 gpurun_out/r02q_pair_il1.log
