#!/bin/bash
# Round-2 single-GPU call: parity tests, smoke, the default bench line (with its `also` block) and the reference arm.
# usage: scripts/gpu_r2.sh <tag> [pytest -k expression]
TAG=${1:-r02a}
KEXPR=${2:-}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.csv 2>&1
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
if [ -n "$KEXPR" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -k "$KEXPR" > $OUT/${TAG}_tests.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $OUT/${TAG}_tests.log 2>&1
fi
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
tail -15 $OUT/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
tail -2 $OUT/${TAG}_smoke.log
SECONDS=0
timeout 900 python bench.py --pageable > $OUT/${TAG}_bench_default.jsonl 2> $OUT/${TAG}_bench_default.err
echo "bench default exit $? after ${SECONDS}s"
tail -3 $OUT/${TAG}_bench_default.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_default.jsonl").read().strip().splitlines()[-1])
    print("fir64", round(d["value"]), "frac %.3f" % d["roofline"]["frac"], "e2e", round(d["e2e"]["value"]), d["e2e"].get("pageable"), "cpu", d.get("cpu_baseline", {}).get("value"))
    for k, v in d.get("also", {}).items():
        print(k, round(v["value"]), "ms %.4f" % v["ms_per_step"], "frac %.3f" % v["roofline"]["frac"], "e2e", round(v["e2e"]["value"]), "cpu %.2f" % v.get("cpu_baseline", {}).get("value", -1), v.get("messages"))
except Exception as e:
    print("parse failed", e)
PY
