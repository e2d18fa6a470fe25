"""Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMA path (B200_PROFILING.md): UTCHMMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, UBLKCP = bulk copy, REDG = red.global, FADD2 = packed fp32 add.
    python scripts/sass_digest.py > profiles/r02_sass_digest.txt        (cuobjdump on the built library, no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "cmlpl_b200", "libcmlpl_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
ops = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "REDG", "FADD2"]
cnt, order, cur, k = collections.defaultdict(collections.Counter), [], None, 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names[k].split("(")[0]
        k += 1
        order.append(cur)
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        if op in ops:
            cnt[cur][op] += 1
for f in sorted(set(order)):
    if any(cnt[f][o] for o in ops[:5]):
        print(f.ljust(62) + "  ".join(f"{o} {cnt[f][o]:4d}" for o in ops))
