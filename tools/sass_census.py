"""Per-kernel census of the Blackwell-native SASS mnemonics in the built library (no GPU needed):
    python tools/sass_census.py [lib.so] > profiles/r02_sass_census.txt
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG/UBLKCP = TMA loads / stores / reduce-adds / bulk
copies, SYNCS = mbarrier traffic, MUFU.EX2 = exponentials, HMMA would be the legacy mma.sync path."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dit4sr_b200", "libdit4sr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
PAT = [("UTC*MMA", r"\bUTC\w*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UTMALDG", r"\bUTMALDG\b"),
       ("UTMASTG", r"\bUTMASTG\b"), ("UTMAREDG", r"\bUTMAREDG\b"), ("UBLKCP", r"\bUBLKCP\b"), ("SYNCS", r"\bSYNCS\b"),
       ("MUFU.EX2", r"\bMUFU\.EX2\b"), ("FFMA2", r"\bFFMA2\b"), ("HMMA", r"\bHMMA\b"), ("LDL/STL", r"\b(LDL|STL)\b")]
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    counts[cur]["instructions"] += 1 if re.search(r"/\*[0-9a-f]{4}\*/", line) else 0
    for name, pat in PAT:
        if re.search(pat, line):
            counts[cur][name] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS census of {os.path.relpath(lib, ROOT)} (cuobjdump -sass), one row per kernel")
hdr = ["instructions"] + [n for n, _ in PAT]
print("kernel | " + " | ".join(hdr))
tot = collections.Counter()
for (k, c), d in zip(counts.items(), dem):
    d = re.sub(r"\(.*", "", d).replace("void ", "")
    print(d[:110] + " | " + " | ".join(str(c.get(h, 0)) for h in hdr))
    tot.update(c)
print("TOTAL | " + " | ".join(str(tot.get(h, 0)) for h in hdr))
