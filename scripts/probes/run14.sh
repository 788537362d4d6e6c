mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "multigpu or strips or mirror" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode strips --steps 16 --warmup 2 --strips-exchange peer > gpurun_out/strips_r02j_2gpu_peer.json 2> gpurun_out/strips_2.err || tail -5 gpurun_out/strips_2.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_r02j.json 2> gpurun_out/bench_r02j.err || tail -20 gpurun_out/bench_r02j.err
python scripts/latency.py --frames 3000
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/strips_r02j_*.json')):
    d = json.load(open(f))
    print(f, round(d['value']), 'Mtri/s', round(d['ms_per_frame']*1e3,1), 'us/frame fpc', d['frames_per_call'], d['frames_timed'], d['covered_pixels'], d['checksum'], 'timeouts', d['signal_timeouts'], 'speedup', d.get('speedup_vs_single_gpu'))
d = json.load(open('gpurun_out/bench_r02j.json'))
print('value', round(d['value']), 'fps', round(d['fps']), 'e2e', json.dumps(d['e2e']))
P
