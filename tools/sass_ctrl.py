"""SASS of one kernel with decoded scheduling control (stall, yield, write/read barrier, wait mask):
python tools/sass_ctrl.py <lib.so|.o> <mangled-name-substring> [from_hex to_hex]"""
import re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
on = False; pend = None
for line in out:
    if "Function :" in line:
        on = pat in line
        continue
    if not on: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", line)
    if m: pend = (int(m.group(1), 16), m.group(2)); continue
    m = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", line)
    if m and pend:
        w = int(m.group(1), 16); c = w >> 41
        stall = c & 0xf; yld = (c >> 4) & 1; wb = (c >> 5) & 7; rb = (c >> 8) & 7; wait = (c >> 11) & 0x3f
        a, txt = pend; pend = None
        if lo <= a <= hi:
            print(f"{a:5x} st{stall:2d} {'Y' if not yld else ' '} w{wb if wb != 7 else '-'} r{rb if rb != 7 else '-'} wait{wait:06b}  {txt.strip()}")
