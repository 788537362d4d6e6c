#!/bin/bash
# On the GPU box: everything profiles/ needs for one round — tests, both bench arms, the ncu launch list of the bench
# command and one `ncu --set full` capture of each kernel for C3 / C4 / C2 pose B.  usage: scripts/round_bundle.sh r02
set -x
tag="${1:-rXX}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err
timeout 600 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -2 gpurun_out/bench_${tag}.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --frames 128 --no-cpu-baseline --no-extras > /dev/null 2> gpurun_out/launches_${tag}.err
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:setup_kernel|raster_kernel" -s 6 -c 2 -f \
    -o gpurun_out/prof_${tag}_c3 python scripts/kernel_times.py c3 --reps 2 2>&1 | tail -2
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:setup_kernel|raster_kernel" -s 6 -c 2 -f \
    -o gpurun_out/prof_${tag}_c4 python scripts/kernel_times.py c4 --batch 8 --reps 2 2>&1 | tail -2
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:setup_|raster_kernel" -s 9 -c 3 -f \
    -o gpurun_out/prof_${tag}_c2b python scripts/kernel_times.py c2b --reps 2 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none -k "regex:mirror_update_kernel" -s 2 -c 1 -f \
    -o gpurun_out/prof_${tag}_mirror python scripts/mirror_tune.py 2>&1 | tail -2
bash scripts/sanitize.sh > gpurun_out/sanitize_${tag}.txt 2>&1; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitize_${tag}.txt
