"""Sum stall samples of an .ncu-rep over SASS index ranges: python tools/ncu_regions.py rep a:b a:b ..."""
import csv, io, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}
body = [r for r in rows[2:] if len(r) >= len(h)]
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
for rg in sys.argv[2:]:
    a, b = [int(x) for x in rg.split(":")]
    sub = body[a:b]
    s = sum(int(r[ix["# Samples"]] or 0) for r in sub)
    ex = sum(int(r[ix["Instructions Executed"]] or 0) for r in sub)
    per = collections.Counter()
    for r in sub:
        for c in stalls:
            per[c] += int(r[ix[c]] or 0)
    print(f"[{a}:{b}] samples {s} ({100*s/tot:.1f}%) executed {ex}  " + " ".join(f"{k[6:]}={v}" for k, v in per.most_common(7)))
