"""registers / spills per kernel from the ptxas logs of the last build: python tools/regs.py [filter]"""
import glob, re, subprocess, sys
flt = sys.argv[1] if len(sys.argv) > 1 else ""
for f in sorted(glob.glob("skirt_b200/csrc/build/*.ptxas.log")):
    name = None
    for line in open(f):
        m = re.search(r"Function properties for (\S+)", line)
        if m: name = m.group(1); spill = ""; continue
        m = re.search(r"(\d+) bytes spill stores", line)
        if m and int(m.group(1)): spill = f" SPILL {m.group(1)}"
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
            if flt in dem: print(f"{int(m.group(1)):4d}{spill}  {dem}")
            name = None
