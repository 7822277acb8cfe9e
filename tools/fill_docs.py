"""Fills the R2_* placeholders of DESIGN.md / README.md from a bench line: python tools/fill_docs.py profiles/r02_bench_final.json"""
import json, sys, re
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e = d["e2e"]
rep = {"R2_MS": f"{d['ms_per_step']:.1f}", "R2_VALUE": f"{d['value'] / 1e6:.0f}", "R2_E2E_FRAC": f"{e['value'] / d['value']:.2f}", "R2_E2E": f"{e['value'] / 1e6:.0f}",
       "R2_FULL": f"{(e.get('full_cycle') or {}).get('value', 0) / 1e6:.0f}", "R2_HOST": f"{(e.get('host_state_every_step') or {}).get('value', 0) / 1e6:.0f}",
       "R2_CPU": f"{d['cpu_baseline']['value'] / 1e6:.1f}"}
for f in ("DESIGN.md", "README.md"):
    s = open(f).read()
    for k in sorted(rep, key=len, reverse=True):
        s = s.replace(k, rep[k])
    open(f, "w").write(s)
print(rep)
