#!/bin/bash
# Round-2 profiling call: ncu launch list of the default bench command, then one --set full capture per headline kernel
TAG=${1:-r02y}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --also none"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_fir64_bench.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
tail -2 $OUT/${TAG}_ncu_launches.log
for cap in "fir64:fir_tc_kernel" "chain5_u8:chain_tc_kernel" "chain:chain3_kernel" "fft4096:fft2_frames" "fft65536:fft65536_fused" "poly8x1024:fir_ptc_kernel"; do
  wl=${cap%%:*}; rx=${cap##*:}
  bash scripts/gpu_ncu_one.sh $TAG $wl $rx > $OUT/${TAG}_ncu_one_$wl.log 2>&1
  grep -E "gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |sm__pipe_tensor_cycles_active_realtime|sm__issue_active.avg.pct" $OUT/${TAG}_ncu_summary_$wl.txt | head -6
done
COMMS_B200_FFT_PATH=cpipe bash scripts/gpu_ncu_one.sh ${TAG}cpipe fft65536 fft65536_cpipe COMMS_B200_FFT_PATH=cpipe > $OUT/${TAG}_ncu_one_cpipe.log 2>&1
ls $OUT | grep $TAG | head -40
