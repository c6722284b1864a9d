mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"xattn_bwd_fused_tc|xattn_fwd_tc|sattn_fwd_tc|sattn_bwd_tc|vit_attn_fwd_tc" -s 12 -c 10 -o gpurun_out/r05q_attn python tools/dev_kernels.py attn > gpurun_out/r05q_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r05q_attn.ncu-rep
