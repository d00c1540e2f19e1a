"""Per-kernel counts of the SASS mnemonics that show which hardware paths the library uses (cuobjdump -sass of the built library):
UTCIMMA = tcgen05.mma kind::i8, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk
copy, UTCBAR = tcgen05.commit, DMMA = FP64 tensor MMA, SYNCS = mbarrier operations.
Usage: python profiles/sass_counts.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "picard-ica_b200", "libpicard_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
want = ["UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "DMMA", "SYNCS", "DFMA", "DMUL", "DADD", "LDS", "STS"]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["total"] += 1
        for w in want:
            if op == w or op.startswith(w + ".") or (w in ("LDS", "STS") and op.startswith(w)):
                counts[cur][w] += 1
demangle = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines() if counts else []
names = dict(zip(counts, demangle)) if len(demangle) == len(counts) else {k: k for k in counts}
print("# " + __doc__.strip().replace("\n", "\n# "))
print(f"# library: picard-ica_b200/libpicard_b200.so ({os.path.getsize(lib)} bytes), {len(counts)} kernels")
print(f"{'instr':>7} " + " ".join(f"{w:>7}" for w in want) + "  kernel")
for k, c in sorted(counts.items(), key=lambda kv: (-(kv[1]['UTCIMMA'] + kv[1]['LDTM']), -kv[1]['DMMA'], names[kv[0]])):
    short = re.sub(r"\(.*", "", re.sub(r"\((?:int|bool|unsigned int|long)\)", "", names[k]))
    print(f"{c['total']:7d} " + " ".join(f"{c[w]:7d}" for w in want) + f"  {short}")
