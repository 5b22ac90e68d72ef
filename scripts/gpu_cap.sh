#!/bin/bash
# One ncu --set full capture.  usage: scripts/gpu_cap.sh <tag> <workload> <kernel-regex>
TAG=$1; wl=$2; rx=$3; OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --workload $wl"
$CMD > $OUT/${TAG}_plain_$wl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o $OUT/${TAG}_prof_$wl $CMD > $OUT/${TAG}_ncu_$wl.log 2>&1
tail -2 $OUT/${TAG}_ncu_$wl.log
