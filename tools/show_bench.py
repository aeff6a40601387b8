#!/usr/bin/env python
"""Pretty-print the interesting fields of a bench.py JSON line."""
import json, sys
for p in sys.argv[1:]:
    d = json.loads(open(p).read().strip().splitlines()[-1])
    print(f"== {p}")
    print(f" value {d['value']:.1f} {d['unit']}   e2e {(d.get('e2e') or {}).get('value', float('nan')):.1f}   ms/step {d['ms_per_step']:.3f}   launches {d.get('gpu_launches')}")
    print(" stages_ms", {k: round(v, 4) for k, v in d["stages_ms"].items()})
    r = d.get("roofline")
    if r: print(f" roofline {r['kernel']}: {r['achieved']:.1f} {r['unit']} frac {r['frac']:.4f} launch_ms {r['launch_ms']:.4f} share {r.get('share_of_step', float('nan')):.3f}")
    r = d.get("roofline_spmm")
    if r: print(f" spmm: {r['achieved']:.0f} GB/s (B_min) frac {r['frac']:.3f}  gather-model {r['b_gather_gbs']:.0f} GB/s")
    print(" prop", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["prop"].items()})
    if d.get("cpu_baseline"): print(" cpu", round(d["cpu_baseline"]["value"], 2), d["cpu_baseline"]["unit"], "cores", d["cpu_baseline"]["cores"])
    print(" clocks", d.get("clocks"), " result", d.get("result"))
