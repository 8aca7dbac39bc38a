#!/usr/bin/env python
"""Executed warp instructions per CUDA source line of one kernel: joins the per-instruction counts of an
.ncu-rep (`--page source --csv`) with the line table of the cubin (`nvdisasm -g`), instruction by
instruction (both list the kernel's SASS in address order).
usage: python scripts/ncu_lines.py X.ncu-rep <kernel-regex> <cubin> [--top N] [--div D]"""
import csv, io, re, subprocess, sys


def main():
    rep, kre, cubin = sys.argv[1:4]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    div = float(sys.argv[sys.argv.index("--div") + 1]) if "--div" in sys.argv else 1.0
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    ic, sc = hdr.index("Instructions Executed"), hdr.index("Source")
    counts = []
    for r in rows[start + 1:]:
        if len(r) <= ic or not r[0].startswith("0x"):
            break
        counts.append((int(r[ic]), r[sc].strip()))
    name = rows[start - 1][1] if start else ""
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    # locate the function: ".text.<mangled>" section whose mangled name matches the regex
    # every ".text.<mangled>" section whose name matches the regex; the instantiation that was profiled is the
    # one with as many SASS instructions as the report lists
    secs, cur, name_m = {}, None, None
    for l in dis:
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m:
            name_m = m.group(1) if re.search(kre, m.group(1)) else None
            if name_m:
                secs[name_m] = []
            continue
        if not name_m:
            continue
        m = re.search(r"//## File \"([^\"]+)\", line (\d+)(.*)", l)
        if m:
            cur = int(m.group(2))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            secs[name_m].append(cur)
    lines = min(secs.values(), key=lambda v: abs(len(v) - len(counts))) if secs else []
    if len(lines) != len(counts):
        print(f"warning: {len(lines)} SASS instructions in the cubin, {len(counts)} in the report", file=sys.stderr)
    per = {}
    for (n, _), ln in zip(counts, lines):
        per[ln] = per.get(ln, 0) + n
    tot = sum(per.values())
    text = open("/root/repo/gst_b200/csrc/gst_kernels.cu").read().splitlines()
    print(f"{name[:100]}: {tot / div:.1f} warp instructions" + (f" per unit (div {div:g})" if div != 1 else ""))
    for ln, n in sorted(per.items(), key=lambda kv: -kv[1])[:top]:
        t = text[ln - 1].strip()[:100] if ln and ln <= len(text) else "?"
        print(f"{n / div:10.1f} {100.0 * n / tot:5.1f}%  L{ln}: {t}")


if __name__ == "__main__":
    main()
