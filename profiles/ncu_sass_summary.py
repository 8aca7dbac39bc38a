#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): executed warp instructions and
stall samples per opcode and per contiguous region, so the hot loop can be read without the GUI.
usage: ncu -i prof.ncu-rep --page source --csv | python profiles/ncu_sass_summary.py [--top N]"""
import csv
import sys
from collections import defaultdict


def main():
    rows = list(csv.reader(sys.stdin))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ops = defaultdict(lambda: [0, 0])
    total_inst = total_samp = 0
    stalls = defaultdict(int)
    body = rows[hi + 1:]
    lines = []
    for r in body:
        if len(r) < len(hdr):
            continue
        sass = r[col["Source"]]
        inst = int(r[col["Instructions Executed"]] or 0)
        samp = int(r[col["# Samples"]] or 0)
        op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
        op = op.split(".")[0] + ("." + sass.split()[0 if not sass.startswith("@") else 1].split(".")[1]
                                 if op.split(".")[0] in ("LDS", "STS", "LDG", "STG") and "." in sass.split()[0 if not sass.startswith("@") else 1] else "")
        ops[op][0] += inst
        ops[op][1] += samp
        total_inst += inst
        total_samp += samp
        for s in stall_cols:
            stalls[s] += int(r[col[s]] or 0)
        lines.append((r[col["Address"]], inst, samp, sass))
    print(f"total warp instructions executed: {total_inst}   stall samples: {total_samp}")
    print("\nby opcode (warp instr, share, samples share):")
    for op, (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:24]:
        print(f"  {op:14s} {n:12d} {100.0 * n / max(total_inst, 1):6.2f}%   samples {100.0 * s / max(total_samp, 1):6.2f}%")
    print("\nstall reasons (share of samples):")
    for s, n in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {s:28s} {100.0 * n / max(total_samp, 1):6.2f}%")
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
    if top:
        print("\nhottest instructions by samples:")
        for a, n, s, sass in sorted(lines, key=lambda t: -t[2])[:top]:
            print(f"  {a} inst={n:10d} samp={s:7d}  {sass}")
    if "--dump" in sys.argv:
        for a, n, s, sass in lines:
            print(f"{a} {n:10d} {s:7d}  {sass}")


if __name__ == "__main__":
    main()
