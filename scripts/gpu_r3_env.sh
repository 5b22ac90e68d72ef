#!/bin/bash
# 1 GPU: one workload under a list of environment settings ("A=1,B=2 C=3" -> two runs), each under its own timeout
TAG=${1:-r03}
WL=${2:-fft65536}
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/${TAG}_env.jsonl
for envs in $3; do
  E=$(echo $envs | tr ',' ' ')
  env $E timeout 120 python bench.py --steps 20 --warmup 3 --workload $WL --no-cpu --no-e2e > $OUT/${TAG}_one.json 2>> $OUT/${TAG}_env.err
  echo "$envs exit $?"
  python - <<PY
import json
for l in open("$OUT/${TAG}_one.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("   ", d["config"]["workload"], round(d["value"]), "ms %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"])
        d["env"] = "$envs"
        open("$OUT/${TAG}_env.jsonl", "a").write(json.dumps(d) + "\n")
PY
done
tail -3 $OUT/${TAG}_env.err 2>/dev/null
