mkdir -p gpurun_out
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --mode strips --steps 16 --warmup 2 > gpurun_out/strips_r02k_${n}gpu.json 2> gpurun_out/strips_$n.err || tail -5 gpurun_out/strips_$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --mode strips --steps 16 --warmup 2 --strips-frames-per-call 1 > gpurun_out/strips_r02k_8gpu_f1.json 2> gpurun_out/strips_8f1.err || tail -5 gpurun_out/strips_8f1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --mode strips --steps 16 --warmup 2 --strips-exchange none > gpurun_out/strips_r02k_8gpu_none.json 2> gpurun_out/strips_8no.err || tail -5 gpurun_out/strips_8no.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 > gpurun_out/bench_r02k_8gpu.json 2> gpurun_out/bench_r02k_8gpu.err ) 2>&1 | tail -3
tail -3 gpurun_out/bench_r02k_8gpu.err
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/strips_r02k_*.json')):
    try:
        d = json.load(open(f))
        print(f, round(d['value']), 'Mtri/s', round(d['ms_per_frame']*1e3,1), 'us/frame fpc', d['frames_per_call'], d['covered_pixels'], d['checksum'], 'timeouts', d['signal_timeouts'], 'speedup', d.get('speedup_vs_single_gpu'), d['nvlink_bytes_per_frame'])
    except Exception as e:
        print(f, 'FAILED', e)
try:
    d = json.load(open('gpurun_out/bench_r02k_8gpu.json'))
    print('bench 8gpu value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['d2h_gbs'], d['e2e'].get('d2h_ceiling_gbs'), 'px only', round(d['e2e']['pixels_only']['value']), 'full', round(d['e2e']['full_frame_copies']['value']))
    s = d.get('strips'); print('strips', round(s['value']), s['ms_per_frame'], s.get('speedup_vs_single_gpu'))
except Exception as e:
    print('bench 8 FAILED', e)
P
