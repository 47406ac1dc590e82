"""Static look at a kernel's SASS: loops and opcode histogram of the largest one.
    python tools/sass_loop.py <mangled-name-substring> [lib]"""
import collections
import re
import subprocess
import sys

name = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "rfest_b200/librfest_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(txt) if "Function :" in l and name in l)
end = next((i for i in range(start + 1, len(txt)) if "Function :" in txt[i]), len(txt))
ins = []
for l in txt[start:end]:
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
print(txt[start].strip(), "instructions:", len(ins))
loops = []
for a, t in ins:
    m = re.search(r"BRA.*0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        loops.append((a - int(m.group(1), 16), int(m.group(1), 16), a))
loops.sort(reverse=True)
for sz, lo, hi in loops[:6]:
    print(f"loop {lo:#x}..{hi:#x}: {sz // 16} instrs")
if loops:
    sz, lo, hi = loops[0]
    c = collections.Counter()
    for a, t in ins:
        if lo <= a <= hi:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            c[t.split()[0].split(".")[0]] += 1
    print(c.most_common(45))
if "--dump" in sys.argv:
    sz, lo, hi = loops[0]
    for a, t in ins:
        if lo <= a <= hi:
            print(f"{a:05x}  {t}")
