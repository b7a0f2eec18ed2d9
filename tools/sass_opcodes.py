"""Per-kernel SASS opcode counts of libfsnerf_b200.so (the mnemonics that prove a Blackwell-native
kernel: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st, UBLKCP = cp.async.bulk,
UTMALDG / UTMASTG = tensor-map TMA, HMMA = legacy mma.sync).
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fsnerf_b200", "libfsnerf_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "UTCCP", "HMMA", "HGMMA",
         "R2UR.BROADCAST", "ELECT", "SYNCS", "LDGSTS", "RED", "ATOMG", "ATOMS", "MEMBAR", "FENCE", "CCTL", "LDS", "STS", "LDG", "STG", "MUFU", "SHFL"]
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("fs::(anonymous namespace)::", "").replace("(anonymous namespace)::", "")
        kern = re.sub(r"^void ", "", kern)
        kern = re.sub(r"\(.*", "", kern)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[kern][w] += 1
                break
print("SASS opcode counts per kernel of fsnerf_b200/libfsnerf_b200.so (cuobjdump -sass, sm_100a)")
print("no kernel uses tensor-map TMA (UTMALDG / UTMASTG): operand images are pre-swizzled in global memory and moved by")
print("plain bulk copies (cp.async.bulk = UBLKCP), completion on mbarriers (SYNCS); no legacy HMMA / HGMMA anywhere.")
print("R2UR.BROADCAST / ELECT: the per-instruction uniform-operand loops; after the elect.sync issue fix the MLP kernels keep them")
print("only around bulk copies issued by several lanes of one warp (by design), not around UTCHMMA.\n")
for k, c in counts.items():
    print(f"{k}  [{total[k]} instructions]")
    print("   " + "  ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
