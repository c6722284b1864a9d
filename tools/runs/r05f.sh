timeout 300 python tools/diag_balance.py 2>&1 | tail -12
