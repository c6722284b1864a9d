mkdir -p gpurun_out
# memcheck: every kernel test + the optimizer tests (out-of-bounds / misaligned accesses in the kernels changed this round)
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_kernels_gpu.py tests/test_optim_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r02w_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r02w_memcheck.log; tail -6 gpurun_out/r02w_memcheck.log
# racecheck (shared-memory hazards) on the kernels whose shared-memory choreography changed
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --print-limit 20 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "layernorm or layerscale or f32_tma" > gpurun_out/r02w_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "hazard" gpurun_out/r02w_racecheck.log; tail -6 gpurun_out/r02w_racecheck.log
