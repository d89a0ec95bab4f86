#!/bin/bash
# sweep of the column-chunked plane-pass geometry (biobank rows) on config 5
for cfg in "16 7168" "16 3584" "8 7168" "8 14336" "8 4608" "4 14336" "4 7168" "4 28672"; do
  set -- $cfg
  FM_CHUNK_WARPS=$1 FM_CHUNK_STEP_BYTES=$2 timeout 300 python tools/bench_configs.py cfg5 --scale 0.5 2>/dev/null | head -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$cfg', round(d['ms_per_pass'],4), round(d['algorithmic_GBps']), round(d['frac_of_measured_peak'],3))"
done
