"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA
(tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG / UBLKCP (TMA), HMMA (legacy mma.sync, must
be absent), plus MUFU and barrier counts.  Reads `cuobjdump -sass` of the in-tree library; no GPU needed.

    python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "interactive-vit_b200", "libvitb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMALDG.2CTA", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR",
        "SYNCS", "MUFU", "HMMA", "LDGSTS", "BAR"]
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["instructions"] += 1
        base = op.split(".")[0]
        for k in KEYS:
            if "." in k:
                if op.startswith(k.split(".")[0]) and "." + k.split(".")[1] in op:
                    counts[kern][k] += 1
            elif base == k:
                counts[kern][k] += 1
print(f"# SASS mnemonic counts per kernel of {os.path.relpath(lib, ROOT)} (cuobjdump -sass; sm_100a)")
print("# UTCHMMA = tcgen05.mma kind::f16, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UTMAREDG = cp.async.bulk.tensor load/store/"
      "reduce, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, HMMA = legacy mma.sync (none expected)")
cols = ["instructions"] + KEYS
print("kernel".ljust(92) + "".join(c.rjust(14) for c in cols))
for k, c in counts.items():
    name = demangle(k)
    name = re.sub(r"\(.*", "", name)[:90]
    print(name.ljust(92) + "".join(str(c.get(col, 0)).rjust(14) for col in cols))
    total.update(c)
print("TOTAL".ljust(92) + "".join(str(total.get(col, 0)).rjust(14) for col in cols))
