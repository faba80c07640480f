"""Distil an .ncu-rep into profiles/roofline_figures.json: the few ncu numbers bench.py quotes, keyed by kernel, with the
digest of the build they were measured on (monte-carlo-portfolio_b200/build/stamp at profile time).

    python tools/ncu_figures.py gpurun_out/prof.ncu-rep <build digest> <tag> [more.ncu-rep ...]

bench.py reads the file and refuses to quote a figure whose digest differs from the library it is running.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "roofline_figures.json")

FIELDS = {
    "fp32_pipe_busy_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "fma_inst_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "tensor_pipe_busy_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "registers_per_thread": "launch__registers_per_thread",
    "duration_ms": "gpu__time_duration.sum",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "inst_executed": "smsp__inst_executed.sum",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
              "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def short_name(full):
    m = re.match(r"(?:void )?(?:mcp::)?([A-Za-z0-9_]+(?:<[^>]*>)?)", full)
    return m.group(1) if m else full


def main():
    digest, tag, reps = sys.argv[2], sys.argv[3], [sys.argv[1]] + sys.argv[4:]
    data = {"kernels": {}}
    if os.path.isfile(OUT):
        with open(OUT) as fh:
            data = json.load(fh)
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        best = {}
        for r in rows[2:]:
            if "nan" in r[idx[FIELDS["issue_active_pct"]]]:
                continue                                  # timing-only launch
            name = short_name(r[idx["Kernel Name"]])
            dur = float(r[idx[FIELDS["duration_ms"]]].replace(",", "")) * UNIT_SCALE.get(units[idx[FIELDS["duration_ms"]]], 1.0)
            if name in best and best[name][0] >= dur:
                continue                                  # keep the longest launch of each kernel (the bench-sized one)
            rec = {}
            for k, m in FIELDS.items():
                if m not in idx:
                    continue
                v = float(r[idx[m]].replace(",", ""))
                u = units[idx[m]]
                if k.startswith("dram_bytes") or k == "duration_ms":
                    v *= UNIT_SCALE.get(u, 1.0)
                rec[k] = v
            rec["dram_bytes"] = rec.get("dram_bytes_read", 0.0) + rec.get("dram_bytes_write", 0.0)
            rec.update(build_digest=digest, source=f"profiles/{tag}_summary.md", report=os.path.basename(rep),
                       how="ncu --set full --clock-control none --import-source on (one GPU)")
            best[name] = (dur, rec)
        for name, (_, rec) in best.items():
            data["kernels"][name] = rec
    with open(OUT, "w") as fh:
        json.dump(data, fh, indent=1, sort_keys=True)
    for k, v in sorted(data["kernels"].items()):
        print(f"{k:45s} {v['duration_ms']:9.3f} ms  fp32 pipe {v['fp32_pipe_busy_pct']:5.1f}%  issue {v['issue_active_pct']:5.1f}%  "
              f"tensor {v['tensor_pipe_busy_pct']:5.1f}%  dram {v['dram_bytes']:.3g} B  digest {v['build_digest'][:12]}")


if __name__ == "__main__":
    main()
