timeout 300 python tools/dbg_nn3.py 2>&1 | tail -20
