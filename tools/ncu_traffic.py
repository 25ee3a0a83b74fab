"""sums an ncu --csv metric log per kernel: python tools/ncu_traffic.py <log.csv> -> {kernel: {metric: sum, launches}} as JSON on stdout"""
import csv, io, json, re, sys
text = open(sys.argv[1]).read()
start = text.find('"ID"')
rows = list(csv.DictReader(io.StringIO(text[start:])))
out = {}
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
    name = re.sub(r"^void\s+", "", name).replace("skg::", "")
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = r.get("Metric Unit", "")
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0, "second": 1e9}.get(unit, 1.0)
    d = out.setdefault(name, {"ids": set()})
    d[r["Metric Name"]] = d.get(r["Metric Name"], 0.0) + v * scale
    d["ids"].add(r["ID"])
for d in out.values():
    d["launches"] = len(d.pop("ids"))
print(json.dumps(out, indent=1))
