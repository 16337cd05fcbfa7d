"""per-source-line instruction counts and stall samples of an ncu report (needs -lineinfo + --import-source on)
usage: ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(f"ncu -i {rep} --page source --csv --print-source sass,cuda", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
fname, hdr, res, tot, tots = None, None, [], 0, 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; iE = hdr.index("Instructions Executed"); iW = hdr.index("Warp Stall Sampling (All Samples)"); continue
    if r[0] in ("Function Name",) or hdr is None or not r[0].isdigit(): continue
    try: n = int(r[iE]); w = int(r[iW])
    except ValueError: continue
    res.append((n, w, fname, int(r[0]), r[1].strip()[:110])); tot += n; tots += w
res.sort(reverse=True)
print(f"total warp instructions {tot}, samples {tots}")
for n, w, f, ln, src in res[:top]:
    print(f"{100*n/tot:5.1f}% inst {100*w/max(tots,1):5.1f}% smp  {f}:{ln}  {src}")
