"""Per-source-line totals from an ncu source-page CSV (cuda,sass view):
   ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv ; python tools/ncu_lines.py f.csv [top]"""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
out = []; fname = ""; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); ss = hdr.index("# Samples"); continue
    if hdr and r[0].strip().isdigit():
        try: out.append((fname, int(r[0]), r[1].strip()[:105], int(r[ie] or 0), int(r[ss] or 0)))
        except ValueError: pass
ti = sum(o[3] for o in out); ts = sum(o[4] for o in out)
print(f"total warp-instr {ti}  samples {ts}")
for f, ln, txt, n, s in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{f[:18]:18s}{ln:5d} {100*n/max(ti,1):5.1f}% inst {100*s/max(ts,1):5.1f}% smp  {txt}")
