#!/usr/bin/env python
"""Print the interesting parts of a bench.py JSON line (file may start with an NCCL banner)."""
import json, sys
for path in sys.argv[1:]:
    for line in open(path):
        if not line.startswith('{'):
            continue
        l = json.loads(line)
        print(f"== {path}: N={l['n_gpus']} value={l['value']:.1f} {l['unit']} ms/step={l['ms_per_step']:.4f} launches={l.get('gpu_launches')}")
        sp = l.get('spread', {})
        print("   spread", {k: round(v, 4) for k, v in sp.get('ms_per_step', {}).items()}, "bracket", round(sp.get('bracket_ms_per_step', 0), 4), "worst", sp.get('worst_step'))
        if 'exchange_wait_ms_median_per_rank' in sp:
            w = sp['exchange_wait_ms_median_per_rank']
            print("   waits totals", [round(x * 1e3, 1) for x in w['totals']], "us; done", [round(x * 1e3, 1) for x in w['done']], "us")
        print("   parity", l.get('parity_ok'), "walk_ms", round(l['roofline']['kernel_ms'], 4) if 'roofline' in l else None)
        for k in ('e2e', 'sponza16M', 'hbm_bound', 'build', 'cpu_baseline'):
            if k in l:
                v = dict(l[k]); v.pop('what', None); v.pop('sample', None)
                print("  ", k, json.dumps(v)[:600])
