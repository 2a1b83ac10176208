"""Fold an .ncu-rep (ncu --set full --import-source on, one kernel) into the two tracked summaries:
key metrics (CSV) and the source page by opcode / stall reason (text).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/<prefix>"""
import csv, io, subprocess, sys, collections

rep, prefix = sys.argv[1], sys.argv[2]
KEYS = ("gpu__time_duration", "launch__", "sm__pipe_fp64_cycles_active", "sm__inst_executed_pipe_fp64", "sm__issue_active", "sm__warps_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "l1tex__t_sector_hit_rate", "lts__t_sector_hit_rate",
        "smsp__thread_inst_executed_per_inst_executed", "smsp__inst_executed.sum", "smsp__average_warp", "sass__inst_executed_local",
        "smsp__warp_issue_stalled", "sm__throughput", "smsp__issue_active", "sm__inst_executed.sum")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
with open(prefix + "_keymetrics.csv", "w") as f:
    f.write("metric,unit,value\n")
    for h, u, v in sorted(zip(hdr, units, vals)):
        if v != "" and (h in ("Kernel Name", "Block Size", "Grid Size") or any(k in h for k in KEYS)):
            f.write(f"{h},{u},{v}\n")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
while rows and (not rows[0] or rows[0][0] != "Address"):      # the page starts with a "Kernel Name" line
    rows.pop(0)
h = rows[0]
def col(name):
    for i, x in enumerate(h):
        if x.strip() == name:
            return i
    return None
ci, cs, cx = col("Source"), col("# Samples"), col("Instructions Executed")
stall_cols = [(i, x) for i, x in enumerate(h) if x.startswith("stall_")]
by_op, by_stall = collections.Counter(), collections.Counter()
n_s = n_x = 0
ops_x = collections.Counter()
for r in rows[1:]:
    if ci is None or len(r) <= ci:
        continue
    parts = r[ci].split()
    op = ""
    for t in parts:
        if t.startswith("@"):
            continue
        op = t.split(".")[0]
        break
    s = int(r[cs] or 0) if cs is not None and r[cs].isdigit() else 0
    x = int(r[cx] or 0) if cx is not None and r[cx].isdigit() else 0
    by_op[op] += s; ops_x[op] += x; n_s += s; n_x += x
    for i, nm in stall_cols:
        if i < len(r) and r[i].isdigit():
            by_stall[nm] += int(r[i])
with open(prefix + "_source_summary.txt", "w") as f:
    f.write(f"{rep}: SASS lines {len(rows) - 1}, warp-instructions executed {n_x:.4g}, stall samples {n_s}\n\nby opcode (share of samples, share of executed warp-instructions):\n")
    for op, s in by_op.most_common(18):
        f.write(f"  {op:10s} samples {100.0 * s / max(n_s, 1):5.1f}%  instructions {100.0 * ops_x[op] / max(n_x, 1):5.1f}%\n")
    fp64 = sum(ops_x[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU", "DMNMX"))
    f.write(f"  FP64-pipe (+MUFU) instructions: {100.0 * fp64 / max(n_x, 1):.1f}% of all warp-instructions\n\nstall reasons (all samples):\n")
    tot = sum(by_stall.values())
    for nm, s in by_stall.most_common(12):
        f.write(f"  {nm:28s} {100.0 * s / max(tot, 1):5.1f}%\n")
print(open(prefix + "_source_summary.txt").read())
