"""Summarise an ncu report per CUDA source line: samples, instructions.  usage: dev_ncu_lines.py rep.ncu-rep [kernel-id]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"] + (["--launch-skip", sys.argv[2], "--launch-count", "1"] if len(sys.argv) > 2 else []), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; hdr = None
agg = collections.OrderedDict()
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0] != "":   # a cuda source line row (aggregated)
        try:
            ln = int(r[0])
        except ValueError:
            continue
        d = dict(zip(hdr[4:], r[-(len(hdr) - 4):]))
        try:
            samples = int(d.get("# Samples", "0") or 0); inst = int(d.get("Instructions Executed", "0") or 0)
        except ValueError:
            continue
        key = (cur_file, ln)
        a = agg.setdefault(key, [0, 0, r[1]])
        a[0] += samples; a[1] += inst
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot} total warp-instructions {toti}")
for (f, ln), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{100*s/tot:5.1f}% samp {100*i/toti:5.1f}% inst  {f}:{ln}: {src.strip()[:110]}")
