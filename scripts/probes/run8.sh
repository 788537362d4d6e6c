set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python scripts/strips_model.py 2>&1 | tail -10
for fpc in 1 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode strips --steps 8 --warmup 2 --strips-frames-per-call $fpc > gpurun_out/strips_r02f_2gpu_f$fpc.json 2> gpurun_out/strips_2.err; tail -2 gpurun_out/strips_2.err
done
python - <<'P'
import json, glob
for f in sorted(glob.glob('gpurun_out/strips_r02f_*.json')):
    try:
        d = json.load(open(f))
        print(f, round(d['value']), 'Mtri/s', d['ms_per_frame'], 'ms/frame', d['rows_per_rank'], d['covered_pixels'], d['checksum'], d['signal_timeouts'], d.get('speedup_vs_single_gpu'), d.get('single_gpu_same_calls'))
    except Exception as e:
        print(f, 'FAILED', e)
P
