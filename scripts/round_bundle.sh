#!/bin/bash
# On the GPU box: everything profiles/ needs for one round — tests, both bench arms, the ncu launch list of the bench
# command and one `ncu --set full` capture of each kernel.  usage: scripts/round_bundle.sh r01e
set -x
tag="${1:-rXX}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err
timeout 400 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -2 gpurun_out/bench_${tag}.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --frames 128 --no-cpu-baseline > /dev/null 2> gpurun_out/launches_${tag}.err
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:setup_kernel|raster_kernel" -s 6 -c 2 -f \
    -o gpurun_out/prof_${tag} python scripts/kernel_times.py c3 --reps 2 2>&1 | tail -2
