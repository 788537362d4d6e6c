mkdir -p gpurun_out
for ex in none peer; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode strips --steps 8 --warmup 2 --strips-exchange $ex > gpurun_out/strips_r02i_2gpu_$ex.json 2> gpurun_out/strips_2.err || tail -5 gpurun_out/strips_2.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode strips --steps 32 --warmup 2 --strips-exchange peer > gpurun_out/strips_r02i_2gpu_peer_long.json 2> gpurun_out/strips_2.err || tail -5 gpurun_out/strips_2.err
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/strips_r02i_*.json')):
    d = json.load(open(f))
    print(f, round(d['value']), 'Mtri/s', round(d['ms_per_frame']*1e3,1), 'us/frame fpc', d['frames_per_call'], d['frames_timed'], d['covered_pixels'], d['checksum'], 'timeouts', d['signal_timeouts'], 'speedup', d.get('speedup_vs_single_gpu'))
P
