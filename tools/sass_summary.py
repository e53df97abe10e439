"""Per-kernel SASS evidence for profiles/sass_summary.txt: counts of the Blackwell tensor-core / TMEM / TMA mnemonics in every
kernel of libmapf_gpt_b200.so (cuobjdump -sass; runs without a GPU).

  UTCHMMA[.2CTA]  tcgen05.mma (cta_group::1 / ::2)        LDTM / STTM   tcgen05.ld / tcgen05.st (tensor memory)
  UTCBAR          tcgen05.commit -> mbarrier               UBLKCP        cp.async.bulk (global -> shared bulk copies)
  UTMALDG         cp.async.bulk.tensor (TMA tiled loads)   UTMAPF/...    bulk L2 prefetch
  HMMA            legacy mma.sync (must be 0)
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

lib = Path(sys.argv[1] if len(sys.argv) > 1 else Path(__file__).resolve().parents[1] / "mapf_gpt_b200" / "libmapf_gpt_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
keys = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UBLKPF", "MUFU.EX2", "HMMA", "instructions"]
per = OrderedDict()
cur = None
it = iter(names)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next(it)
        cur = re.sub(r"\(.*", "", cur).replace("mg::", "")
        per[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m or cur is None:
        continue
    op = m.group(1)
    per[cur]["instructions"] += 1
    for k in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UBLKPF", "HMMA"):
        if op.startswith(k):
            per[cur][k] += 1
    if op.startswith("UTCHMMA") and ".2CTA" in line:
        per[cur]["UTCHMMA.2CTA"] += 1
    if op.startswith("MUFU.EX2"):
        per[cur]["MUFU.EX2"] += 1
print(f"# {lib.name}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a); tools/sass_summary.py")
print(f"{'kernel':<64}" + "".join(f"{k:>13}" for k in keys))
tot = Counter()
for name, c in per.items():
    tot.update(c)
    print(f"{name[:63]:<64}" + "".join(f"{c[k]:>13}" for k in keys))
print(f"{'TOTAL':<64}" + "".join(f"{tot[k]:>13}" for k in keys))
assert tot["HMMA"] == 0, "legacy mma.sync found"
