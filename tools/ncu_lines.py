"""per-source-line instruction counts of one kernel in an .ncu-rep (needs -lineinfo + --import-source on):
python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname = None; hdr = None; out = []; tot = 0; totsmp = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ie = r.index("Instructions Executed"); te = r.index("Thread Instructions Executed"); sm = r.index("# Samples"); continue
    if hdr is None or r[0] == "": continue          # SASS rows have an empty line number
    try: v = int(r[ie]); t = int(r[te]); s = int(r[sm])
    except ValueError: continue
    if v > 0: out.append((v, t, s, fname, r[0], r[1].strip()[:100])); tot += v; totsmp += s
out.sort(reverse=True)
print(f"total warp instructions {tot:.4g}, samples {totsmp}")
for v, t, s, f, l, src in out[:top]:
    print(f"{100*v/tot:5.1f}% inst {100*s/max(totsmp,1):5.1f}% smp lanes {t/v:5.1f} {f}:{l} {src}")
