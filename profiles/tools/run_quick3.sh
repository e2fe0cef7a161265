timeout 300 python profiles/tools/time_fill_pits.py 2>&1 | tail -16
