mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vit_attn_fwd_tc|vit_attn_bwd_fused" -c 2 -o gpurun_out/r04l_attn python tools/attn_case.py fwd 1 > gpurun_out/r04l_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vit_attn_bwd_fused" -c 1 -o gpurun_out/r04l_attn_bwd python tools/attn_case.py bwd 1 > gpurun_out/r04l_ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
ls -la gpurun_out/*.ncu-rep
