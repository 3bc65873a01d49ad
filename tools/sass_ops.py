"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libtsd_b200.so (cuobjdump -sass):
UTC*MMA (tcgen05.mma), UTMALDG / UTMASTG (TMA loads / stores), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), MUFU.EX2.  Usage: python tools/sass_ops.py > profiles/rNN_sass_ops.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "stable-diffusion.mojo_b200", "csrc", "libtsd_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
pats = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("UTC*MMA.2CTA", r"\bUTC[A-Z]*MMA\.2CTA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
        ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("MUFU.EX2", r"\bMUFU\.EX2"),
        ("HMMA (legacy)", r"\bHMMA"), ("instructions", r"^\s*/\*[0-9a-f]{4,}\*/")]
cur, counts = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("tsd::", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name)
        cur = counts.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    for key, pat in pats:
        if re.search(pat, ln):
            cur[key] += 1
print(f"# {os.path.basename(lib)}: Blackwell SASS mnemonics per kernel (cuobjdump -sass, sm_100a); no cuBLAS / cuDNN / CUTLASS is linked")
print("# " + " | ".join(k for k, _ in pats))
tot = collections.Counter()
for name, c in counts.items():
    if any(c[k] for k, _ in pats[:9]):
        print(f"{name}: " + " ".join(f"{k}={c[k]}" for k, _ in pats if c[k]))
    tot.update(c)
print("TOTAL: " + " ".join(f"{k}={tot[k]}" for k, _ in pats))
