#!/bin/bash
# Final profiling call of the round: ncu launch list of the default bench command, then one --set full capture of each
# kernel that changed in the second session
TAG=${1:-r04p}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --also none"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_fir64_bench.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
tail -2 $OUT/${TAG}_ncu_launches.log
for cap in "fir64:fir_tc_kernel" "fir64_iq16:fir_tc_kernel" "pulse4_i16:fir_ptc_kernel" "fft4096_iq16:fft2_frames_iq16"; do
  wl=${cap%%:*}; rx=${cap##*:}
  bash scripts/gpu_ncu_one.sh $TAG $wl $rx > $OUT/${TAG}_ncu_one_$wl.log 2>&1
  grep -E "gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |sm__inst_issued.avg.pct_of_peak_sustained_active" $OUT/${TAG}_ncu_summary_$wl.txt | head -5
done
ls $OUT | grep $TAG | head -40
