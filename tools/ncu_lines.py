"""Per-source-line stall samples / instructions from an .ncu-rep captured with --import-source on.
usage: python tools/ncu_lines.py file.ncu-rep kernel_regex [topN]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                      "regex:" + kern, "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, data = "", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        si, ii = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        try:
            data.append((float(r[si]), float(r[ii]), fname, int(r[0]), r[1].strip()[:120]))
        except ValueError:
            pass
tot = sum(d[0] for d in data) or 1.0
toti = sum(d[1] for d in data) or 1.0
print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
for d in sorted(data, reverse=True)[:top]:
    print(f"{100*d[0]/tot:5.1f}% stall {100*d[1]/toti:5.1f}% inst  {d[2]}:{d[3]}  {d[4]}")
