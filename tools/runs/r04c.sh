XFM_ATTN_PROF=1 timeout 300 python tools/attn_case.py bwd 1 2>&1 | grep -v "fwd prof" | tail -3
