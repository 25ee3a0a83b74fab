"""the hot loop of a kernel in an .ncu-rep: SASS instructions executed at least FRAC x the maximum count, in address order.
python tools/ncu_hot_sass.py rep [frac=0.5]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ie = hdr.index("Instructions Executed"); te = hdr.index("Thread Instructions Executed"); sm = hdr.index("# Samples")
ins = []
for r in rows[2:]:
    try: ins.append((r[0], r[1].strip(), int(r[ie]), int(r[te]), int(r[sm])))
    except (ValueError, IndexError): pass
mx = max(i[2] for i in ins); tot = sum(i[2] for i in ins); n = 0
for a, s, v, t, smp in ins:
    if v >= frac * mx:
        n += 1; print(f"{a[-5:]} {v/mx:5.2f} lanes {t/max(v,1):5.1f} smp {smp:5d}  {s[:90]}")
print(f"{n} hot instructions; max count {mx}, total {tot}")
