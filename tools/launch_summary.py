"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py launches.csv [out.md]
Per kernel: launches, total/avg duration and share of the listed GPU time (cold-cache, serialised: shares matter, not absolutes)."""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
kn, mv, mn = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
t = collections.OrderedDict()
for r in rows[1:]:
    if r[mn] != "gpu__time_duration.sum":
        continue
    name = r[kn].split("(")[0]
    a = t.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += float(r[mv].replace(",", ""))
tot = sum(a[1] for a in t.values())
lines = ["| kernel | launches | total ms | avg us | share |", "|---|---:|---:|---:|---:|"]
for n, (c, ns) in sorted(t.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"| `{n}` | {c} | {ns/1e6:.3f} | {ns/c/1e3:.1f} | {100*ns/tot:.1f}% |")
lines.append(f"| total | {sum(a[0] for a in t.values())} | {tot/1e6:.3f} | | |")
txt = "\n".join(lines) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt)
print(txt)
