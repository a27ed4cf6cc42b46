"""Per-kernel SASS opcode histogram of libstmp.so (cuobjdump -sass): the evidence that the hot kernels are Blackwell-native
(UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP/UTMALDG = TMA bulk/tensor copies, UTCBAR = tcgen05.commit, SYNCS = mbarrier).
   python tools/sass_histogram.py [lib] > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pytorch_geometric_temporal_b200", "lib", "libstmp.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "")
        kern = re.sub(r"\(.*", "", name)
        hist.setdefault(kern, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0]] += 1
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "LDGSTS", "FFMA2", "ATOMS", "REDUX")
print(f"# {os.path.relpath(lib, ROOT)}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a)")
for k in sorted(hist):
    h = hist[k]
    tot = sum(h.values())
    if tot == 0 or "cub::" in k:
        continue
    key = " ".join(f"{o}={h[o]}" for o in KEY if h[o])
    top = " ".join(f"{o}:{c}" for o, c in h.most_common(8))
    print(f"{k}\n    instructions {tot}; Blackwell/async: [{key or '-'}]\n    top: {top}")
