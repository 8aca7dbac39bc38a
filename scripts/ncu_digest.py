#!/usr/bin/env python
"""Digest of an .ncu-rep: per kernel the headline metrics, pipe utilisation and the SASS summary.
usage: python scripts/ncu_digest.py X.ncu-rep [--top N]"""
import csv
import io
import subprocess
import sys

KEYS = ("Duration", "Executed Ipc Active", "Issue Slots Busy", "No Eligible", "Eligible Warps Per", "Active Warps Per",
        "Registers Per", "Achieved Occupancy", "Theoretical Occupancy", "Mem Busy", "L1/TEX Hit", "L2 Hit",
        "DRAM Throughput", "Executed Instructions", "Mem Pipes Busy", "Waves Per SM", "Block Limit Reg", "Block Limit Sha")
PIPES = ("alu", "fma", "xu", "lsu", "uniform", "adu", "cbu")


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 10
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    for line in det.splitlines():
        if line.startswith("  ") and not line.startswith("   ") and "(" in line:
            print("\n==", line.strip()[:110])
        elif any(k in line for k in KEYS) and "OPT" not in line and "INF" not in line:
            print("  ", " ".join(line.split()))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    names = [r[hdr.index("Kernel Name")] for r in rows[2:]]
    print("\npipe utilisation (% of peak, active):")
    for pipe in PIPES:
        col = f"sm__inst_executed_pipe_{pipe}.avg.pct_of_peak_sustained_active"
        if col in hdr:
            print(f"  {pipe:8s}", [f"{float(r[hdr.index(col)]):.1f}" for r in rows[2:]])
    for col in ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"):
        if col in hdr:
            print(f"  {col}", [r[hdr.index(col)] for r in rows[2:]])
    print("  kernels:", [n[:40] for n in names])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = src.split("\n")
    idx = [i for i, l in enumerate(lines) if l.startswith('"Address"')]
    seen = set()
    for n, a in enumerate(idx):
        name = lines[a - 1] if a else ""
        if name in seen:
            continue
        seen.add(name)
        end = idx[n + 1] - 1 if n + 1 < len(idx) else len(lines)
        print("\n######", name[:120])
        out = subprocess.run([sys.executable, __file__.replace("scripts/ncu_digest.py", "profiles/ncu_sass_summary.py"), "--top", str(top)],
                             input="\n".join(lines[a:end]), capture_output=True, text=True)
        print(out.stdout)


if __name__ == "__main__":
    main()
