"""SASS digest of librpg_b200.so: per kernel, the counts of the mnemonics that prove the Blackwell-native path
(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor loads / stores, UBLKCP = bulk copies,
SYNCS = mbarrier, FFMA2 = packed fp32x2) plus the legacy tensor path (HMMA) that must NOT appear.
usage: python tools/sass_digest.py [lib.so] > profiles/rN_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "relpose_gnn_b200", "librpg_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "LDGSTS", "FFMA2", "MUFU.EX2", "HMMA"]
per = collections.OrderedDict()
cur = None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = per.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and cur is not None:
        op = m.group(1)
        cur["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                cur[w] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            cur["UTCHMMA.2CTA"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per), stdout=subprocess.PIPE, text=True).stdout.splitlines()
tot = collections.Counter()
print(f"# {os.path.basename(lib)}: SASS mnemonic counts per kernel (cuobjdump -sass); arch sm_100a")
print("# " + " ".join(f"{w:>12}" for w in WATCH) + "  instr  kernel")
for (name, c), nice in zip(per.items(), demangle):
    tot.update(c)
    if any(c[w] for w in WATCH if w not in ("SYNCS", "FFMA2", "LDGSTS", "MUFU.EX2")) or "attention" in nice or "gemm" in nice:
        short = re.sub(r"\(.*", "", nice.replace("(anonymous namespace)::", "")).replace("rpg::", "")
        print("  " + " ".join(f"{c[w]:12d}" for w in WATCH) + f" {c['_total']:6d}  {short[:90]}")
print("# total " + " ".join(f"{w}={tot[w]}" for w in WATCH))
assert tot["HMMA"] == 0, "legacy mma.sync path found"
