"""Aggregates an ncu source page (``ncu -i x.ncu-rep --page source --csv``) by code region.

    python tools/ncu_regions.py x.ncu-rep [chunk_bytes]

Regions: the kernel's main body, split into chunks of ``chunk_bytes`` of SASS (default 0x400), and
every out-of-line subroutine (a CALL target).  Per region: static instructions, executed
warp-instructions (and per warp of the launch), average active lanes, stall samples.
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    chunk = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x400
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    print(rows[start - 1][1] if start else "")
    hdr, data = rows[start], [r for r in rows[start + 1:] if len(r) > 5 and r[0].startswith("0x")]
    ia, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    base = int(data[0][0], 16)
    targets = sorted({int(m.group(1), 16) for r in data for m in [re.search(r"CALL\S*\s+(0x[0-9a-f]+)", r[1])] if m})
    first_sub = targets[0] if targets else 1 << 60
    warps = max(int(r[ia]) for r in data[:4]) or 1  # the prologue runs once per warp
    agg = collections.OrderedDict()
    for r in data:
        off = int(r[0], 16) - base
        if off < first_sub:
            key = "main+0x%04x" % (off // chunk * chunk)
        else:
            key = "sub@0x%04x" % [t for t in targets if t <= off][-1]
        a = agg.setdefault(key, [0, 0, 0, 0])
        a[0] += 1
        a[1] += int(r[ia])
        a[2] += int(r[it])
        a[3] += int(r[isamp])
    tot = [sum(a[k] for a in agg.values()) for k in range(4)]
    print(f"warps {warps}  static {tot[0]}  warp-instructions {tot[1]} = {tot[1] / warps:.0f} per warp  "
          f"avg lanes {tot[2] / max(tot[1], 1):.1f}  samples {tot[3]}")
    for k, a in agg.items():
        if a[1] == 0 and a[3] == 0:
            continue
        print(f"{k:16s} static {a[0]:5d}  per-warp {a[1] / warps:7.1f} ({100 * a[1] / tot[1]:5.1f} %)  lanes {a[2] / max(a[1], 1):5.1f}  "
              f"samples {a[3]:6d} ({100 * a[3] / max(tot[3], 1):4.1f} %)")


main()
