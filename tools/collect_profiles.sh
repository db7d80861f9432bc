# Copy the judged evidence from gpurun_out/ (scratch) to profiles/ (tracked): ncu summaries, launch list, bench lines,
# and the per-launch DRAM traffic that bench.py's roofline.traffic reports.
set -e
R=${1:-r01}
python tools/ncu_summary.py gpurun_out/scan_alt_${R}_final.ncu-rep > profiles/scan_alt_grid_${R}_ncu_summary.txt
python tools/ncu_summary.py gpurun_out/scan_null_${R}_final.ncu-rep > profiles/scan_null_grid_${R}_ncu_summary.txt
python tools/ncu_summary.py gpurun_out/scan_exact_${R}.ncu-rep > profiles/scan_null_exact_${R}_ncu_summary.txt
cp gpurun_out/launches_alt_${R}.csv profiles/launches_bench_alt_grid_${R}.csv
python - "$R" <<'PY'
import csv, io, json, subprocess, sys
R = sys.argv[1]
out = {}
for wl, rep in (("alt-grid", f"scan_alt_{R}_final"), ("null-grid", f"scan_null_{R}_final"), ("null-exact", f"scan_exact_{R}")):
    raw = subprocess.run(["ncu", "-i", f"gpurun_out/{rep}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]
    d = {k: (float(x.replace(",", "")), un) for k, un, x in zip(h, u, v) if k in ("dram__bytes_read.sum", "dram__bytes_write.sum")}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    out[wl] = sum(val * scale[un] for val, un in d.values())
json.dump(out, open(f"profiles/scan_traffic_{R}.json", "w"))
print(out)
PY
