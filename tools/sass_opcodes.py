"""Per-kernel SASS opcode table of the built library (cuobjdump -sass): which kernels use the tcgen05 tensor path
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit), TMA (UTMALDG, UBLKCP), the warp-level
tensor path (HMMA, LDSM), and how large they are.   python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parents[1] / "openviic_b200" / "lib" / "libopenviic_cap.so"
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "LDSM", "LDGSTS", "SYNCS", "ACQBULK",
         "LDL", "STL"]
rows, cur, k = [], None, 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = collections.Counter()
        rows.append((names[k], cur))
        k += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
        cur["_total"] += 1
print(f"{'kernel':100s} {'instr':>6s} " + " ".join(f"{w:>7s}" for w in WATCH))
for name, c in sorted(rows, key=lambda r: r[0]):
    short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", ""))[:100]
    print(f"{short:100s} {c['_total']:6d} " + " ".join(f"{sum(v for o, v in c.items() if o.startswith(w)):7d}" for w in WATCH))
