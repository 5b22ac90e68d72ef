#!/bin/bash
# One gpurun call: selected parity tests + selected bench workloads (no ncu).
# usage: scripts/gpu_quick.sh <tag> "<pytest -k expr>" "<workloads>"
TAG=${1:-q}
KEXPR=${2:-}
WLS=${3:-"fir64"}
OUT=gpurun_out
mkdir -p $OUT
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "$KEXPR" > $OUT/${TAG}_tests.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_tests.log
  tail -15 $OUT/${TAG}_tests.log
fi
: > $OUT/${TAG}_bench.jsonl
for wl in $WLS; do
  timeout 600 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu --no-e2e >> $OUT/${TAG}_bench.jsonl 2>> $OUT/${TAG}_bench.err
done
python - <<PY
import json
for l in open("$OUT/${TAG}_bench.jsonl"):
    d = json.loads(l)
    print(d["config"]["workload"], "value %.0f Ms/s" % d["value"], "ms/step %.3f" % d["ms_per_step"],
          "frac %.3f" % d["roofline"]["frac"], d["clocks"])
PY
tail -5 $OUT/${TAG}_bench.err
