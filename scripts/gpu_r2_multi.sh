#!/bin/bash
# N-GPU call (gpurun --gpus N): the torchrun parity worker, then the default bench line at N GPUs (with `also` and the gather leg)
TAG=${1:-r02c}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 tests/mg_worker.py > $OUT/${TAG}_mg_worker_${N}gpu.log 2>&1
echo "mg_worker exit $?" >> $OUT/${TAG}_mg_worker_${N}gpu.log
grep -v "^\[W\|^W0\|^\*\*\*" $OUT/${TAG}_mg_worker_${N}gpu.log | tail -12
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 800 > $OUT/${TAG}_pytest_multi.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_multi.log
  tail -3 $OUT/${TAG}_pytest_multi.log
fi
SECONDS=0
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $N ${3:-} > $OUT/${TAG}_bench_${N}gpu.jsonl 2> $OUT/${TAG}_bench_${N}gpu.err
echo "bench exit $? after ${SECONDS}s"
grep -v "^\[W\|^W0\|^\*\*\*" $OUT/${TAG}_bench_${N}gpu.err | tail -5
python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_bench_${N}gpu.jsonl") if l.startswith("{")][-1])
    print("fir64", d["n_gpus"], round(d["value"]), "frac %.3f" % d["roofline"]["frac"], "e2e", round(d["e2e"]["value"]))
    print("  gather", json.dumps(d.get("gather"))[:900])
    for k, v in d.get("also", {}).items():
        print(k, round(v["value"]), "ms %.4f" % v["ms_per_step"], "frac %.3f" % v["roofline"]["frac"], "e2e", round(v["e2e"]["value"]))
        if "gather" in v: print("  gather", json.dumps(v["gather"])[:900])
except Exception as e:
    print("parse failed", e)
PY
