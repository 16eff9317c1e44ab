#!/usr/bin/env python
"""Compare shape-sweep files: tools/cmp_shapes.py base.json other.json [...]"""
import json, sys
def load(p): return {json.loads(l)['shape']: json.loads(l) for l in open(p) if l.strip()}
base = load(sys.argv[1])
for p in sys.argv[2:]:
    print('==', p)
    for k, d in load(p).items():
        b = base.get(k)
        if not b: continue
        print(f"{k:26s} fwd {d['fwd_ms']:.4f} ({d['fwd_frac']:.2f}) vs {b['fwd_ms']:.4f} ({b['fwd_frac']:.2f}) | "
              f"bwd {d['bwd_ms']:.4f} ({d['bwd_frac']:.2f}) vs {b['bwd_ms']:.4f} ({b['bwd_frac']:.2f}) | step {d['step_frac']:.2f} vs {b['step_frac']:.2f}")
