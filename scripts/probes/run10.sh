mkdir -p gpurun_out
nvidia-smi topo -m | head -12
for n in 8 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --mode strips --steps 8 --warmup 2 > gpurun_out/strips_r02g_${n}gpu.json 2> gpurun_out/strips_$n.err || tail -5 gpurun_out/strips_$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --mode strips --steps 8 --warmup 2 --strips-frames-per-call 1 > gpurun_out/strips_r02g_8gpu_f1.json 2> gpurun_out/strips_8f1.err || tail -5 gpurun_out/strips_8f1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --mode strips --steps 8 --warmup 2 --strips-exchange nccl > gpurun_out/strips_r02g_8gpu_nccl.json 2> gpurun_out/strips_8n.err || tail -5 gpurun_out/strips_8n.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 > gpurun_out/bench_r02g_8gpu.json 2> gpurun_out/bench_r02g_8gpu.err ) 2>&1 | tail -3
tail -3 gpurun_out/bench_r02g_8gpu.err
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/strips_r02g_*.json')):
    try:
        d = json.load(open(f))
        print(f, round(d['value']), 'Mtri/s', round(d['ms_per_frame']*1e3,1), 'us/frame fpc', d['frames_per_call'], d['rows_per_rank'], d['covered_pixels'], d['checksum'], 'timeouts', d['signal_timeouts'], 'speedup', d.get('speedup_vs_single_gpu'), d.get('single_gpu_same_calls'))
    except Exception as e:
        print(f, 'FAILED', e)
try:
    d = json.load(open('gpurun_out/bench_r02g_8gpu.json'))
    print('bench 8gpu value', round(d['value']), 'e2e', json.dumps(d['e2e']))
    print('strips', json.dumps(d.get('strips')))
    print('latency', json.dumps(d.get('latency')))
except Exception as e:
    print('bench 8 FAILED', e)
P
