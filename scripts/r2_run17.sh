#!/bin/bash
echo "modes 1 2 3"; MODES=1,2,3 timeout 300 python scripts/bench_wgrad.py 2>&1 | head -6 | cut -c1-90
echo "rows 4"; PB_WG_RS_ROWS=4 MODES=1 timeout 300 python scripts/bench_wgrad.py 2>&1 | head -4 | cut -c1-90
echo "UC 1"; PB_WG_RS_UC=1 MODES=1,2,3 timeout 300 python scripts/bench_wgrad.py 2>&1 | head -2 | cut -c1-90
