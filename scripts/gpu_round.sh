#!/bin/bash
# One gpurun call: parity tests, smoke, bench lines, then the ncu launch list of the bench command.
# usage: scripts/gpu_round.sh <tag> [pytest -k expression]
TAG=${1:-r1}
KEXPR=${2:-}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.csv 2>&1
if [ -n "$KEXPR" ]; then
  timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -k "$KEXPR" > $OUT/${TAG}_tests.log 2>&1
else
  timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > $OUT/${TAG}_tests.log 2>&1
fi
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
tail -5 $OUT/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.log
tail -2 $OUT/${TAG}_smoke.log
: > $OUT/${TAG}_bench.jsonl
for wl in fir64 fir64_real fir1024 fir63d5 fir63d5_real fft1024 fft4096 ifft4096 fft8192 fft16384 fft32768 fft65536 fft262144 fft1048576 freqest timing10x5 mixer fm chain chain5 chain5_u8 pulse4 pulse4_i16 poly8x1024 poly8x1024c; do
  extra="--no-cpu"
  [ "$wl" = "fir64" ] && extra=""
  timeout 600 python bench.py --steps 20 --warmup 3 --workload $wl $extra >> $OUT/${TAG}_bench.jsonl 2>> $OUT/${TAG}_bench.err
done
cat $OUT/${TAG}_bench.jsonl | cut -c1-600
