#!/bin/bash
# One gpurun call: bench lines for every workload, then ncu launch list + full captures.
# usage: scripts/gpu_prof.sh <tag> "<workloads to bench>" "<workload:kernel-regex ...> to capture"
TAG=${1:-r1}
WLS=${2:-"fir64 fir64_real fft1024 fft4096 fft65536 ifft4096 chain pulse4"}
CAPS=${3:-"fir64:fir_stream fft4096:fft_frames chain:chain_kernel"}
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/${TAG}_bench.jsonl
for wl in $WLS; do
  extra="--no-cpu"
  [ "$wl" = "fir64" ] && extra=""
  timeout 600 python bench.py --steps 20 --warmup 3 --workload $wl $extra >> $OUT/${TAG}_bench.jsonl 2>> $OUT/${TAG}_bench.err
done
python - <<PY
import json
for l in open("$OUT/${TAG}_bench.jsonl"):
    d = json.loads(l)
    print(d["config"]["workload"], "value %.0f Ms/s" % d["value"], "ms/step %.3f" % d["ms_per_step"],
          "frac %.3f" % d["roofline"]["frac"], "e2e %.0f" % (d["e2e"] or {}).get("value", 0), d["clocks"])
PY
# launch list of the default bench command (plain run first, same command line)
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
ls -la $OUT
