"""List the SASS instructions of an .ncu-rep with the most stall samples, in address order with context,
so the hot regions of the kernel (spin loops, epilogue, K loop) can be told apart."""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}
body = [r for r in rows[2:] if len(r) >= len(h)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
cols = ["stall_long_sb", "stall_math", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_mio"]
print("total samples", tot)
print(f"{'idx':>5} {'samples':>8} {'exec':>10}  " + " ".join(c[6:12].rjust(6) for c in cols) + "  sass")
for i, r in enumerate(body):
    s = int(r[ix["# Samples"]] or 0)
    if s >= thr * tot:
        print(f"{i:5d} {s:8d} {int(r[ix['Instructions Executed']] or 0):10d}  " + " ".join((r[ix[c]] or '0').rjust(6) for c in cols) + "  " + r[ix["Source"]].strip()[:90])
