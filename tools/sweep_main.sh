#!/bin/bash
# sweep of warps-per-CTA x step bytes for the non-chunked plane pass on the bench workload
for cfg in "16 7168" "8 14336" "8 7168" "8 9344" "12 9344" "6 18688" "4 28672"; do
  set -- $cfg
  FM_WARPS=$1 FM_STEP_BYTES=$2 timeout 300 python bench.py --steps 20 --warmup 3 --skip-e2e --skip-cpu 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin); r=d['roofline']; print('$cfg', 'ms/step %.4f'%d['ms_per_step'], [(g['haplotypes'], round(g['ms'],4), round(g['GBps'])) for g in r['per_group']])"
done
for cfg in "8 14336" "6 18688" "8 9344" "10 11392" "12 9344"; do
  set -- $cfg
  FM_CHUNK_WARPS=$1 FM_CHUNK_STEP_BYTES=$2 timeout 300 python tools/bench_configs.py cfg5 --scale 0.5 2>/dev/null | head -1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('chunked $cfg', round(d['ms_per_pass'],4), round(d['algorithmic_GBps']), round(d['frac_of_measured_peak'],3))"
done
