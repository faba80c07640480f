"""Static SASS opcode counts of the shipped kernels (`cuobjdump -sass libmcp.so`) -> profiles/<tag>_sass_counts.md.
Evidence that the hot kernels are Blackwell-native (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk,
FFMA2 = fma.rn.f32x2).      python tools/sass_counts.py r2"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "monte-carlo-portfolio_b200", "lib", "libmcp.so")
OPS = ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "ELECT", "SYNCS", "FFMA2", "FFMA", "IMAD.WIDE", "LOP3", "MUFU", "FMNMX3", "REDUX", "DFMA")
WANT = ("small_sweep_packedILi16ELi4ELb0ELi10", "small_sweep_packedILi16ELi4ELb0ELi7", "small_sweepIdLi16ELi2ELi0ELb0ELi10", "large_sweep_tcILb1ELi10ELb0", "large_sweep_tcILb1ELi10ELb1",
        "large_sweep_tcILb0ELi10ELb0", "path_kernel_tcILi16ELi4ELi1ELi1ELi10", "path_kernel_tcILi16ELi4ELi1ELi1ELi7", "path_kernel_tcILi32ELi3ELi1ELi1ELi10", "path_kernel_tc_wideILi64ELi2ELi10", "path_kernel_tc_wideILi128ELi1ELi10", "path_kernel_tc16ILi64ELi2ELi10", "path_kernel_tc16ILi128ELi4ELi10",
        "hist_var_fastILi12", "path_kernel_packedILi16ELi10", "large_sweepILi10", "path_kernel_wideIfLi", "moments_cov")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn, counts = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if fn and m:
            for o in OPS:
                if m.group(1) == o or m.group(1).startswith(o + "."):
                    counts[fn][o] += 1
    rows = []
    for fn, c in counts.items():
        if any(w in fn for w in WANT):
            name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
            rows.append((name[:140], c))
    rows.sort()
    with open(os.path.join(ROOT, "monte-carlo-portfolio_b200", "build", "stamp")) as fh:
        digest = fh.read().strip()
    out = ["# SASS opcode counts of the shipped kernels (`cuobjdump -sass libmcp.so`, static instruction counts)", "",
           f"build digest (`monte-carlo-portfolio_b200/build/stamp`): `{digest}`", "",
           "Blackwell-only opcodes: `UTCHMMA` = `tcgen05.mma` (kind::f16 / kind::tf32), `LDTM` / `STTM` = `tcgen05.ld` / `tcgen05.st` (tensor memory), "
           "`UBLKCP` = `cp.async.bulk` (TMA bulk copy), `ELECT` = `elect.sync`, `SYNCS` = mbarrier operations, `FFMA2` = `fma.rn.f32x2`, "
           "`FMNMX3` = three-input `min` / `max`.", "",
           "| kernel | " + " | ".join(OPS) + " |", "|---|" + "---:|" * len(OPS)]
    for name, c in rows:
        out.append(f"| `{name}` | " + " | ".join(str(c.get(k, 0)) for k in OPS) + " |")
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_counts.md")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    print("wrote", path, len(rows), "kernels")


if __name__ == "__main__":
    main()
