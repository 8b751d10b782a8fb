timeout 300 python tools/exp_overlap.py 2>&1 | tail -6
