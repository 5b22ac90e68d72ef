#!/bin/bash
# 1 GPU: a possibly hanging new kernel, every step under its own short timeout
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -q --timeout 120 -x -k "$2" > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
tail -25 $OUT/${TAG}_tests.log
: > $OUT/${TAG}_bench.jsonl
for cfgline in $3; do
  wl=${cfgline%%:*}; path=${cfgline##*:}
  COMMS_B200_FFT_PATH=$path timeout 120 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu --no-e2e >> $OUT/${TAG}_bench.jsonl 2>> $OUT/${TAG}_bench.err
  echo "$wl $path exit $?"
done
python - <<PY
import json
for l in open("$OUT/${TAG}_bench.jsonl"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["config"]["workload"], round(d["value"]), "ms %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"])
PY
tail -3 $OUT/${TAG}_bench.err 2>/dev/null
