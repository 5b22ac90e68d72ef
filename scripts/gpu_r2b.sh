#!/bin/bash
# 1 GPU: new tests of this sub-round, graph message-rate bench, default bench line with pageable e2e
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -k "$2" > $OUT/${TAG}_tests.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_tests.log
tail -25 $OUT/${TAG}_tests.log
make -C comms-rs_b200/host -s all
timeout 300 ./comms-rs_b200/host/bench_graph 8192 2048 > $OUT/${TAG}_bench_graph.jsonl 2>&1; echo "bench_graph exit $?"
cat $OUT/${TAG}_bench_graph.jsonl
timeout 600 python bench.py --pageable --also none --steps 10 > $OUT/${TAG}_bench_pageable.jsonl 2> $OUT/${TAG}_bench_pageable.err; echo "bench exit $?"
python -c "
import json; d=json.loads(open('$OUT/${TAG}_bench_pageable.jsonl').read().strip().splitlines()[-1]); print(d['value'], d['e2e'])"
tail -3 $OUT/${TAG}_bench_pageable.err
