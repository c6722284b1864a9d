XFM_ATTN_PROF=1 timeout 300 python tools/attn_case.py fwd 1 2>&1 | tail -3
