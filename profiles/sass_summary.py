"""Per-kernel SASS mnemonic summary of the shipped library (cuobjdump -sass), written to profiles/r2_sass_summary.txt.
    python profiles/sass_summary.py [lib.so]"""
import re
import subprocess
import sys
from collections import Counter

lib = sys.argv[1] if len(sys.argv) > 1 else "homophily_marl_b200/libssd_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "STG.E.ENL2.256", "STG.E.128", "LDG.E.128", "FFMA2",
       "ACQBULK", "PREEXIT", "PRMT", "SHF", "LOP3", "MATCH", "VOTE", "REDUX", "CREDUX", "SHFL", "ATOMS", "LDS", "STS", "FFMA", "IMAD", "HMMA", "F2FP", "BAR")
fn, counts, arch = None, {}, None
for line in out.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        counts[fn] = Counter()
        counts[fn]["_arch_" + str(arch)] = 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fn:
        op = m.group(1)
        counts[fn]["_total"] += 1
        for k in KEY:
            if op.startswith(k) and not (k == "FFMA" and op.startswith("FFMA2")):
                counts[fn][k] += 1
print(f"{lib}: {len(counts)} kernels")
for fn, c in counts.items():
    short = re.sub(r"\(anonymous namespace\)::", "", fn)
    short = re.sub(r"\(.*", "", short)
    arch = [k[6:] for k in c if k.startswith("_arch_")][0]
    keys = "  ".join(f"{k} {c[k]}" for k in KEY if c[k])
    print(f"\n{short}  [{arch}]  {c['_total']} instructions\n    {keys}")
