"""Summarise an .ncu-rep (read here, no GPU) into profiles/<tag>_summary.md + <tag>_metrics.csv.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep r1a [gpurun_out/launches.csv]
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    launches = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    keep = [k for k in KEEP if k in idx]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"{tag}_metrics.csv"), "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["id", "kernel"] + keep)
        w.writerow(["", ""] + [units[idx[k]] for k in keep])
        for r in rows[2:]:
            w.writerow([r[0], r[idx["Kernel Name"]]] + [r[idx[k]] for k in keep])
    lines = [f"# ncu summary `{tag}`", "",
             f"Source: `{os.path.basename(rep)}` (`ncu --set full --clock-control none --import-source on`, B200, one GPU).",
             "Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.", ""]
    if launches and os.path.isfile(launches):
        lr = list(csv.reader(l for l in open(launches) if not l.startswith("==")))
        ki, vi = lr[0].index("Kernel Name"), lr[0].index("Metric Value")
        agg = collections.OrderedDict()
        for r in lr[1:]:
            if len(r) <= vi:
                continue
            a = agg.setdefault(r[ki], [0, 0.0])
            a[0] += 1
            a[1] += float(r[vi].replace(",", ""))
        tot = sum(a[1] for a in agg.values())
        lines += ["## Launch list (`--metrics gpu__time_duration.sum`)", "", "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            lines.append(f"| `{k[:100]}` | {c} | {t / 1e6:.3f} | {100 * t / tot:.1f}% |")
        lines.append("")
        with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as fh:
            fh.write(open(launches).read())
    lines += ["## Per-kernel metrics (first launch of each distinct kernel / size)", ""]
    seen = set()
    # some launches come back with only the timing pass (every other metric -nan): prefer a fully measured launch
    probe = "smsp__issue_active.avg.pct_of_peak_sustained_active"
    full = [r for r in rows[2:] if probe not in idx or "nan" not in r[idx[probe]]]
    partial = [r for r in rows[2:] if r not in full]
    for r in full + partial:
        key = (r[idx["Kernel Name"]], r[idx["gpu__time_duration.sum"]][:3])
        if key in seen:
            continue
        seen.add(key)
        lines.append(f"### `{r[idx['Kernel Name']][:110]}` (id {r[0]})")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---:|---|")
        for k in keep:
            lines.append(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |")
        lines.append("")
    with open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w") as fh:
        fh.write("\n".join(lines))
    print("wrote profiles/%s_summary.md" % tag)


if __name__ == "__main__":
    main()
