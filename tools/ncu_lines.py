"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export by CUDA source line.

usage: ncu -i rep --page source --csv --print-source sass,cuda --kernel-name regex:NAME > x.csv; python tools/ncu_lines.py x.csv [top]
"""
import csv
import collections
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hdr_i]
col = {}
for i, h in enumerate(hdr):
    col.setdefault(h, i)
agg = collections.defaultdict(lambda: collections.Counter())
cur_line, cur_src = None, ""
srcs = {}
total = 0
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr) - 5:
        continue
    ln = r[0]
    if ln.strip():
        # a CUDA source line record (SASS rows that follow belong to it) -- format depends on ncu; handle both
        cur_line = ln
        srcs[cur_line] = r[1][:110]
    try:
        samples = int(r[col["# Samples"]] or 0)
    except ValueError:
        samples = 0
    a = agg[cur_line]
    a["samples"] += samples
    total += samples
    for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_lg", "stall_math", "stall_wait",
              "stall_not_selected", "stall_selected", "stall_dispatch", "stall_branch_resolving", "stall_no_inst"):
        try:
            a[k] += int(r[col[k]] or 0)
        except (ValueError, KeyError):
            pass
    for k in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive", "L2 Theoretical Sectors Global", "Instructions Executed"):
        try:
            a[k] += int(r[col[k]] or 0)
        except (ValueError, KeyError):
            pass
print(f"total samples {total}")
tot = collections.Counter()
for a in agg.values():
    tot.update(a)
print("stall totals:", {k: v for k, v in tot.most_common() if k.startswith("stall")})
print("smem wavefronts", tot["L1 Wavefronts Shared"], "excessive", tot["L1 Wavefronts Shared Excessive"], "L2 sectors", tot["L2 Theoretical Sectors Global"])
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    main = sorted(((k, v) for k, v in a.items() if k.startswith("stall")), key=lambda kv: -kv[1])[:3]
    print(f"{100*a['samples']/max(total,1):5.1f}% L{ln:>4s} smemWF {a['L1 Wavefronts Shared']:>9d} {str(main):70s} | {srcs.get(ln,'')}")
