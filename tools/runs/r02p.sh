mkdir -p gpurun_out
XFM_GEMM_RASTER=1 timeout 300 python tools/dev_gemm_f32epi.py raster1 > gpurun_out/r02p_f32epi_r1.log 2>&1; echo "rc=$?"; tail -9 gpurun_out/r02p_f32epi_r1.log
XFM_GEMM_RASTER=0 timeout 300 python tools/dev_gemm_f32epi.py raster0 > gpurun_out/r02p_f32epi_r0.log 2>&1; echo "rc=$?"; tail -9 gpurun_out/r02p_f32epi_r0.log
