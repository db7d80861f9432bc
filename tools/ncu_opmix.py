"""Opcode mix (executed warp instructions) of a SASS index range of an .ncu-rep."""
import csv, io, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}
body = [r for r in rows[2:] if len(r) >= len(h)]
a, b = [int(x) for x in sys.argv[2].split(":")]
ex = collections.Counter(); n = collections.Counter()
for r in body[a:b]:
    toks = r[ix["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "SHF", "IMAD")) else op.split(".")[0]
    ex[op] += int(r[ix["Instructions Executed"]] or 0); n[op] += 1
tot = sum(ex.values())
for k, v in ex.most_common(40): print(f"{k:14s} static {n[k]:5d} executed {v:11d} {100*v/tot:5.1f}%")
if len(sys.argv) > 3:
    for i, r in enumerate(body[a:b]): print(a + i, r[ix["Instructions Executed"]], r[ix["# Samples"]], r[ix["Source"]].strip())
