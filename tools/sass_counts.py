"""Per-kernel SASS opcode counts of the shipped library: the evidence that the hot path is tcgen05 / TMEM / TMA
code and holds no legacy tensor-core instructions.  Runs here (no GPU): python tools/sass_counts.py > profiles/rNN_sass_opcode_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200", "lib", "libfa_fwd_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "MUFU.EX2", "HMMA"]
cur, counts = None, collections.OrderedDict()
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", ln)
    if m:
        op = m.group(1)
        counts[cur]["total"] += 1
        for o in ops:
            if op.startswith(o):
                counts[cur][o] += 1
                break
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS opcode counts per kernel of lib/libfa_fwd_sm100.so (cuobjdump -sass, sm_100a), tools/sass_counts.py.")
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load /")
print("# store / reduce-add, UTMAPF = TMA prefetch to L2, MUFU.EX2 = ex2.approx, HMMA = legacy mma.sync (must be 0).")
hdr = ("UTCHMMA", ".2CTA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "MUFU.EX2", "HMMA", "total")
print(" ".join(f"{h:>8}" for h in hdr) + "  kernel")
tot = collections.Counter()
for (k, c), name in zip(counts.items(), names):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    row = (c["UTCHMMA"], c["UTCHMMA.2CTA"], c["LDTM"], c["STTM"], c["UTMALDG"], c["UTMASTG"], c["UTMAREDG"], c["UTMAPF"],
           c["MUFU.EX2"], c["HMMA"], c["total"])
    print(" ".join(f"{v:>8}" for v in row) + "  " + name)
    tot.update(c)
row = (tot["UTCHMMA"], tot["UTCHMMA.2CTA"], tot["LDTM"], tot["STTM"], tot["UTMALDG"], tot["UTMASTG"], tot["UTMAREDG"],
       tot["UTMAPF"], tot["MUFU.EX2"], tot["HMMA"], tot["total"])
print(" ".join(f"{v:>8}" for v in row) + "  ALL KERNELS")
