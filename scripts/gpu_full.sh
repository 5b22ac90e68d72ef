#!/bin/bash
# One gpurun call: full GPU parity suite, smoke, every bench workload, then the ncu launch list
# of the default bench command and `--set full` captures of the named kernels.
# usage: scripts/gpu_full.sh <tag> "<workload:kernel-regex ...>"
TAG=${1:-r1}
CAPS=${2:-"fir64:fir_tc fft4096:fft2_frames chain:chain_kernel"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.csv 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
tail -6 $OUT/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
tail -2 $OUT/${TAG}_smoke.log
: > $OUT/${TAG}_bench.jsonl
for wl in fir64 fir64_real fft1024 fft4096 ifft4096 fft65536 chain pulse4 poly8x1024; do
  extra="--no-cpu"
  [ "$wl" = "fir64" ] && extra=""
  timeout 600 python bench.py --steps 20 --warmup 3 --workload $wl $extra >> $OUT/${TAG}_bench.jsonl 2>> $OUT/${TAG}_bench.err
done
python - <<PY
import json
for l in open("$OUT/${TAG}_bench.jsonl"):
    d = json.loads(l)
    print(d["config"]["workload"], "value %.0f Ms/s" % d["value"], "ms/step %.3f" % d["ms_per_step"],
          "frac %.3f" % d["roofline"]["frac"], "e2e %.0f" % (d["e2e"] or {}).get("value", 0), d["clocks"], d.get("cpu_baseline", {}).get("value"))
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
for cap in $CAPS; do
  wl=${cap%%:*}; rx=${cap##*:}
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --workload $wl"
  $CMD > $OUT/${TAG}_plain_$wl.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o $OUT/${TAG}_prof_$wl $CMD > $OUT/${TAG}_ncu_$wl.log 2>&1
  tail -2 $OUT/${TAG}_ncu_$wl.log
done
ls -la $OUT | tail -20
