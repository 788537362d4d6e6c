#!/bin/bash
# quick GPU check: parity tests (fail fast) + short bench without the CPU baseline; prints the headline numbers
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err || tail -5 gpurun_out/bench_quick.err
python - <<'P'
import json
d = json.load(open('gpurun_out/bench_quick.json'))
r = d['roofline']
print('value', round(d['value']), 'Mtri/s  fps', round(d['fps']), ' e2e', round(d['e2e']['value']), ' kernel ms',
      {k: round(v['ms_per_launch'], 4) for k, v in r['kernels'].items()}, ' frac', round(r['frac'], 3), ' path', round(r['path_frac'], 3))
P
