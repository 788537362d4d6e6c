set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
timeout 120 python scripts/latency.py --frames 3000 2>&1 | tail -3
timeout 120 python scripts/latency.py --frames 3000 --no-depth 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err || tail -20 gpurun_out/bench_r02b.err
python - <<'P'
import json
d = json.load(open('gpurun_out/bench_r02b.json'))
print('value', round(d['value']), 'fps', round(d['fps']), 'e2e', json.dumps(d['e2e']), 'kernel ms', d['roofline']['kernel_ms_per_launch'])
P
