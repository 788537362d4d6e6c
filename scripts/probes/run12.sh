mkdir -p gpurun_out
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --mode strips --steps 8 --warmup 2 > gpurun_out/strips_r02h_${n}gpu.json 2> gpurun_out/strips_$n.err || tail -5 gpurun_out/strips_$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --mode strips --steps 8 --warmup 2 --strips-frames-per-call 1 > gpurun_out/strips_r02h_8gpu_f1.json 2> gpurun_out/strips_8f1.err || tail -5 gpurun_out/strips_8f1.err
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/strips_r02h_*.json')):
    d = json.load(open(f))
    print(f, round(d['value']), 'Mtri/s', round(d['ms_per_frame']*1e3,1), 'us/frame fpc', d['frames_per_call'], d['covered_pixels'], d['checksum'], 'timeouts', d['signal_timeouts'], 'speedup', d.get('speedup_vs_single_gpu'), d['nvlink_bytes_per_frame'])
P
