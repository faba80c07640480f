"""Per-opcode instruction mix and top stall sites of one kernel in an .ncu-rep (read here, no GPU).
    python tools/ncu_hot.py gpurun_out/prof.ncu-rep [n_top] [kernel-name regex] [launch-skip among the matching launches]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
flt = []
if len(sys.argv) > 3: flt += ["-k", "regex:" + sys.argv[3]]
flt += ["-s", sys.argv[4] if len(sys.argv) > 4 else "0", "-c", "1"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + flt, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum"]
for h, u, v in zip(hdr, units, vals):
    if h in want or ("issue_stalled" in h and "per_issue_active" in h and float(v or 0) > 0.05):
        print(f"{h:95s} {v:>16s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"] + flt, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]
ia, ie, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break          # next kernel of the report
    if len(r) > max(ia, ie, iss) and r[ie].isdigit(): data.append(r)
tot = sum(int(r[ie]) for r in data); tots = sum(int(r[iss]) for r in data)
print("total inst", tot, "samples", tots)
op = collections.Counter(); ops = collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ia]); o = m.group(2) if m else r[ia][:10]
    op[o] += int(r[ie]); ops[o] += int(r[iss])
for o, c in op.most_common(22):
    print(f"{o:12s} {c:>12d} {100 * c / tot:6.2f}%   stall-samples {100 * ops[o] / tots:6.2f}%")
print("--- top stall instructions")
top = sorted(range(len(data)), key=lambda i: -int(data[i][iss]))[:ntop]
for i in sorted(top):
    print(i, data[i][ia][:80], data[i][ie], data[i][iss])
