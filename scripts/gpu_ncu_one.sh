#!/bin/bash
# One ncu --set full capture of one kernel of one bench workload (plain run first), raw CSV + summary under gpurun_out/
# usage: scripts/gpu_ncu_one.sh <tag> <workload> <kernel regex> [env assignments...]
TAG=$1; WL=$2; RX=$3; shift 3
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --also none --workload $WL"
env "$@" $CMD > $OUT/${TAG}_plain_$WL.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain_$WL.log; exit 1; }
env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k regex:$RX -s 2 -c 1 -f -o $OUT/${TAG}_prof_$WL $CMD > $OUT/${TAG}_ncu_$WL.log 2>&1
tail -2 $OUT/${TAG}_ncu_$WL.log
ncu -i $OUT/${TAG}_prof_$WL.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_$WL.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/${TAG}_ncu_full_$WL.csv | tee $OUT/${TAG}_ncu_summary_$WL.txt | head -80
rm -f $OUT/${TAG}_prof_$WL.ncu-rep
