mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:setup_|raster_kernel|reject_kernel" --csv --log-file gpurun_out/strip_ncu.csv python scripts/probes/strip_ncu.py
python - <<'P'
import csv
rows = list(csv.reader(open('gpurun_out/strip_ncu.csv')))
hdr = None
out = {}
for r in rows:
    if 'Kernel Name' in r:
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    d = dict(zip(hdr, r))
    out.setdefault((d['ID'], d['Kernel Name'][:30]), {})[d['Metric Name']] = d['Metric Value']
for k, v in out.items():
    print(k, v)
P
