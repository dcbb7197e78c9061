"""cuobjdump -sass of the in-tree library: tensor-core / TMA / TMEM mnemonic counts per tcgen05 kernel (no GPU needed).
    python scripts/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "stemseg_b200", "libstemseg_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
names = {}
try:
    import cxxfilt  # noqa: F401
except ImportError:
    cxxfilt = None
MNEMONICS = ("UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTCATOMSWS", "LDTM", "SYNCS", "ELECT", "UTMACCTL", "HMMA")
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        if op in MNEMONICS:
            counts[cur][op] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts.keys()), stdout=subprocess.PIPE, text=True).stdout.splitlines()
print("# cuobjdump -sass stemseg_b200/libstemseg_b200.so (sm_100a): tensor-core / TMA / TMEM mnemonics per kernel")
print("# UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA load), LDTM = tcgen05.ld,")
print("# UTCATOMSWS = tcgen05.alloc/dealloc, SYNCS = mbarrier ops, ELECT = elect.sync; HMMA would be the legacy mma.sync path")
total = collections.Counter()
for mangled, name in zip(counts.keys(), dem):
    c = counts[mangled]
    total.update(c)
    if not (c["UTCHMMA"] or c["UTMALDG"] or c["LDTM"]):
        continue
    short = name.replace("void ", "").replace("stemseg::(anonymous namespace)::", "").replace("stemseg::<unnamed>::", "")
    short = re.sub(r"\((anonymous namespace|CUtensorMap|stemseg::).*", "", short).rstrip("(")
    print("%-44s %s" % (short, "  ".join("%s=%d" % (k, c[k]) for k in MNEMONICS if c[k])))
print("# whole library: " + "  ".join("%s=%d" % (k, total[k]) for k in MNEMONICS))
print("# kernels in the library: %d" % len(counts))
