#!/usr/bin/env python
"""Per-source-line share of executed warp instructions for one kernel.

    python tools/sass_lines.py <ncu source-page csv> <nvdisasm -g -c listing> <kernel substring> [source file]

ncu's `--page source --csv` lists SASS instructions with "Instructions Executed" but no line numbers; `nvdisasm -g -c`
lists the same instructions in the same order with `//## File ..., line N` markers (needs -lineinfo).  Joined by order.
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    src_csv, sass, kern = sys.argv[1:4]
    source = sys.argv[4] if len(sys.argv) > 4 else None
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ie = hdr.index("Instructions Executed")
    execd = [int(r[ie]) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
    lines, cur, inside = [], None, False
    for ln in open(sass):
        if ln.startswith(".text.") and ln.rstrip().endswith(":"):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]*)", line (\d+)', ln)
        if m:
            cur = (m.group(1).rsplit("/", 1)[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append(cur)
    n = min(len(lines), len(execd))
    if len(lines) != len(execd):
        print(f"warning: {len(lines)} SASS instructions in the listing, {len(execd)} in the profile", file=sys.stderr)
    per = defaultdict(int)
    for i in range(n):
        per[lines[i]] += execd[i]
    tot = sum(per.values())
    text = open(source).read().splitlines() if source else None
    print(f"total warp instructions {tot}")
    base = source.rsplit("/", 1)[-1] if source else None
    for key, c in sorted(per.items(), key=lambda kv: kv[0] or ("", 0)):
        if c >= tot * 0.003:
            f, line = key if key else ("?", 0)
            t = text[line - 1].strip()[:100] if text and f == base and line else ""
            print(f"{100.0 * c / tot:5.1f}%  {f}:{line}  {t}")


if __name__ == "__main__":
    main()
