set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
nproc
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
./scripts/probes/pcie_probe > gpurun_out/pcie_probe_r02.txt 2>&1; cat gpurun_out/pcie_probe_r02.txt
timeout 400 ncu --set full --clock-control none --import-source on -k "regex:setup_kernel|raster_kernel" -s 6 -c 2 -f \
    -o gpurun_out/prof_r02a python scripts/kernel_times.py c3 --reps 2 2>&1 | tail -2
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err || tail -5 gpurun_out/bench_r02a.err
cat gpurun_out/bench_r02a.json | head -c 1500
