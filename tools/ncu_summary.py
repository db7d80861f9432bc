"""Summarise an .ncu-rep (one kernel launch): key raw metrics + stall samples / executed instructions by opcode."""
import collections, csv, io, subprocess, sys

def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1]] + args, capture_output=True, text=True).stdout

raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, units, vals = raw[0], raw[1], raw[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "sm__cycles_elapsed.avg",
        "lts__t_bytes.sum", "launch__shared_mem_per_block_dynamic"]
for h, u, v in zip(hdr, units, vals):
    if h in want or h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"):
        print(f"{h:80s} {v} {u}")
src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv"]))))
h = src[1]; ix = {k: i for i, k in enumerate(h)}
samp = collections.Counter(); exe = collections.Counter(); tot = 0
for r in src[2:]:
    if len(r) < len(h): continue
    toks = r[ix["Source"]].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    s = int(r[ix["# Samples"]] or 0); tot += s
    samp[op] += s; exe[op] += int(r[ix["Instructions Executed"]] or 0)
print("--- stall samples by opcode (total %d)" % tot)
for k, v in samp.most_common(14): print(f"{k:10s} {v:9d} {100*v/tot:5.1f}%   executed {exe[k]}")
