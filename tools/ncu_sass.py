"""List the SASS of one kernel from an `ncu --page source --csv --print-source sass` export, in address order, with
stall samples per instruction.

usage: ncu -i rep --page source --csv --print-source sass --kernel-name regex:NAME > x.csv
       python tools/ncu_sass.py x.csv [PATTERN] [before] [after]    # window around the first instruction matching PATTERN
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and ("Source" in r) and ("# Samples" in r))
hdr = rows[hi]
col = {h: i for i, h in reversed(list(enumerate(hdr)))}
sass = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2]
S = col["Source"]


def I(r, k):
    try:
        return int(r[col[k]])
    except (ValueError, KeyError, IndexError):
        return 0


pat = sys.argv[2] if len(sys.argv) > 2 else "DMMA"
before = int(sys.argv[3]) if len(sys.argv) > 3 else 20
after = int(sys.argv[4]) if len(sys.argv) > 4 else 100
idx = [i for i, r in enumerate(sass) if pat in r[S]]
tot = sum(I(r, "# Samples") for r in sass)
print(f"{len(sass)} instructions, {len(idx)} match {pat!r}, total samples {tot}")
lo = max(0, idx[0] - before) if idx else 0
for r in sass[lo: lo + before + after]:
    print(f"{I(r,'# Samples'):6d} L{I(r,'stall_long_sb'):5d} S{I(r,'stall_short_sb'):5d} M{I(r,'stall_math'):5d} "
          f"W{I(r,'stall_wait'):5d} B{I(r,'stall_barrier'):5d} | {r[S][:100]}")
