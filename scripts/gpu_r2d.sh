#!/bin/bash
# 1 GPU: selected tests (-k "$2"), then one bench line per workload in "$3" (device-resident + e2e, no CPU leg)
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
if [ -n "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -k "$2" > $OUT/${TAG}_tests.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_tests.log
  tail -30 $OUT/${TAG}_tests.log
fi
: > $OUT/${TAG}_bench.jsonl
for wl in $3; do
  timeout 600 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu ${4:-} >> $OUT/${TAG}_bench.jsonl 2>> $OUT/${TAG}_bench.err
done
python - <<PY
import json
for l in open("$OUT/${TAG}_bench.jsonl"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["config"]["workload"], round(d["value"]), "ms %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e", d["e2e"] and round(d["e2e"]["value"]))
PY
tail -3 $OUT/${TAG}_bench.err 2>/dev/null
