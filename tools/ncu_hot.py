#!/usr/bin/env python
"""Top stall-sample instructions per kernel from an .ncu-rep (source page, SASS view)."""
import csv
import subprocess
import sys


def main(path, top=18, only=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks = out.split('"Kernel Name",')
    seen = set()
    for b in blocks[1:]:
        lines = b.split("\n")
        name = lines[0].strip().strip('",')[:70]
        if only and only not in name:
            continue
        if name in seen:
            continue
        seen.add(name)
        r = list(csv.reader(lines[1:]))
        hdr = r[0]
        si, src = hdr.index("# Samples"), hdr.index("Source")
        rows = [x for x in r[1:] if len(x) > si and x[si].isdigit()]
        tot = sum(int(x[si]) for x in rows) or 1
        print(f"=== {name}  (samples {tot})")
        for idx, x in sorted(enumerate(rows), key=lambda t: -int(t[1][si]))[:top]:
            ctx = rows[idx - 1][src][:50] if idx > 0 else ""
            print(f"  {100 * int(x[si]) / tot:5.1f}%  {x[src][:70]:70s} | prev: {ctx}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 18, sys.argv[3] if len(sys.argv) > 3 else None)
