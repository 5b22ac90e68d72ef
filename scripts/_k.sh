timeout 600 python -m pytest tests -m gpu -q -x -k "real_input_fir or fm_radio_example" 2>&1 | tail -12
for w in fir63d5_real; do timeout 600 python bench.py --workload $w --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/est_$w.json; python -c "
import sys,json; d=json.loads(open('gpurun_out/est_$w.json').read()); print(d['config']['workload'][:12], d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d.get('cpu_baseline',{}).get('value'))"; done
