"""Pretty-print a bench.py JSON line."""
import json
import sys

for path in sys.argv[1:]:
    d = json.load(open(path))
    print(f"== {path}: {d['value']:.0f} {d['unit']}  ({d['ms_per_step']:.2f} ms/step)  clocks {d.get('clocks')}")
    if d.get("e2e"):
        print(f"   e2e {d['e2e']['value']:.0f}  cpu {d.get('cpu_baseline')}")
    for k, v in (d.get("stages") or {}).items():
        print(f"   {k:14s} {v['ms_total']:8.2f} ms  avg {v['avg_launch_ms']*1e3:7.1f} us  {v['bound']:6s} "
              f"{v['achieved']:8.1f} {v['unit']:8s} frac {v['frac']:.2f}  (hbm {v['hbm_gbs']:.0f} GB/s, fp64 {v['fp64_tflops']:.1f} TF)")
    r = d.get("roofline")
    if r:
        print(f"   dominant: {r['kernel']} {r['bound']} frac {r['frac']:.2f} share {r['share_of_step']:.2f}")
