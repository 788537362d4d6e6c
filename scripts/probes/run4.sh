set -x
mkdir -p gpurun_out
( time timeout 600 python bench.py > gpurun_out/bench_r02d_1gpu.json 2> gpurun_out/bench_r02d_1gpu.err ) 2>&1 | tail -4
tail -3 gpurun_out/bench_r02d_1gpu.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_r02d_2gpu.json 2> gpurun_out/bench_r02d_2gpu.err ) 2>&1 | tail -4
tail -3 gpurun_out/bench_r02d_2gpu.err
timeout 300 python bench.py --mode strips --steps 8 --warmup 2 > gpurun_out/strips_r02d_1gpu.json 2> gpurun_out/strips_1.err; tail -2 gpurun_out/strips_1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --mode strips --steps 8 --warmup 2 > gpurun_out/strips_r02d_2gpu.json 2> gpurun_out/strips_2.err; tail -2 gpurun_out/strips_2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --mode strips --steps 8 --warmup 2 --strips-exchange nccl > gpurun_out/strips_r02d_2gpu_nccl.json 2> gpurun_out/strips_2n.err; tail -2 gpurun_out/strips_2n.err
python - <<'P'
import json
for f in ('bench_r02d_1gpu','bench_r02d_2gpu'):
    try:
        d = json.load(open(f'gpurun_out/{f}.json'))
    except Exception as e:
        print(f, 'FAILED', e); continue
    print(f, 'value', round(d['value']), 'ms/step', d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'frac', d['roofline']['frac'], 'path_frac', d['roofline']['path_frac'])
    print('  e2e', json.dumps(d['e2e']))
    print('  latency', json.dumps(d.get('latency')))
    for k, v in (d.get('other_configs') or {}).items():
        print('  ', k, round(v['value']), 'Mtri/s', round(v['fps']), 'fps', v['kernel_ms_per_launch'], 'path_frac', round(v['roofline']['path_frac'],3), 'cpu', v.get('cpu_baseline',{}).get('fps'))
    print('  strips', json.dumps(d.get('strips')))
    print('  cpu', json.dumps(d.get('cpu_baseline')))
for f in ('strips_r02d_1gpu','strips_r02d_2gpu','strips_r02d_2gpu_nccl'):
    try:
        d = json.load(open(f'gpurun_out/{f}.json'))
        print(f, round(d['value']), 'Mtri/s', d['ms_per_frame'], 'ms/frame', d['rows_per_rank'], d['covered_pixels'], d['checksum'], d['signal_timeouts'])
    except Exception as e:
        print(f, 'FAILED', e)
P
