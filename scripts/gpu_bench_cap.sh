#!/bin/bash
# One gpurun call: bench lines of every workload, the ncu launch list of the default bench command, then ncu --set full
# captures of the kernels named on the command line as <workload>:<kernel-regex> pairs.
# usage: scripts/gpu_bench_cap.sh <tag> [workload:regex ...]
TAG=${1:-r1}; shift
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/${TAG}_bench.jsonl
for wl in fir64 fir64_real fir1024 fir63d5 fir63d5_real fft1024 fft4096 ifft4096 fft8192 fft16384 fft32768 fft65536 fft262144 fft1048576 freqest timing10x5 mixer fm chain chain5 chain5_u8 pulse4 pulse4_i16 poly8x1024 poly8x1024c; do
  extra="--no-cpu"
  [ "$wl" = "fir64" ] && extra=""
  timeout 600 python bench.py --steps 20 --warmup 3 --workload $wl $extra >> $OUT/${TAG}_bench.jsonl 2>> $OUT/${TAG}_bench.err
done
python - <<P
import json
for l in open("$OUT/${TAG}_bench.jsonl"):
    d = json.loads(l)
    print(d["config"]["workload"], round(d["value"]), round(d["ms_per_step"], 3), round(d["roofline"]["frac"], 3), round(d["e2e"]["value"]) if d.get("e2e") else None)
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_fir64_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_launches.log 2>&1
for pair in "$@"; do
  wl=${pair%%:*}; rx=${pair#*:}
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --workload $wl"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o $OUT/${TAG}_prof_$wl $CMD > $OUT/${TAG}_ncu_$wl.log 2>&1
  tail -1 $OUT/${TAG}_ncu_$wl.log
  ncu -i $OUT/${TAG}_prof_$wl.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_$wl.csv 2>/dev/null
done
