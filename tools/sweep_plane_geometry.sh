#!/bin/bash
# usage: tools_sweep.sh "STEP LG DEBUG" ...   (tuning sweep of the plane-pass geometry; prints per-group GB/s)
for cfg in "$@"; do
  set -- $cfg
  FM_STEP_BYTES=$1 FM_FORCE_LG=$2 FM_DEBUG=${3:-0} timeout 300 python bench.py --steps 20 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/b.json 2> gpurun_out/b.err
  rc=$?
  python - "$cfg" $rc <<'PY'
import json,sys
cfg, rc = sys.argv[1], sys.argv[2]
try:
    d=json.load(open("gpurun_out/b.json")); r=d["roofline"]
    print(cfg, "rc", rc, "ms/step %.4f"%d["ms_per_step"], "value %.3e"%d["value"], r["per_group"] and [(g["haplotypes"], round(g["ms"],4), round(g["GBps"])) for g in r["per_group"]], round(r["achieved"]))
except Exception as e:
    print(cfg, "rc", rc, "FAILED", open("gpurun_out/b.err").read()[-300:].replace("\n"," | "))
PY
done
