for st in 1 0 1 0; do XFM_XATTN_STAGE=$st timeout 300 python tools/dev_kernels.py attn 2>&1 | grep -E "fusion_cross" | sed "s/^/stage=$st /"; done
