"""Per-kernel SASS opcode summary of the in-tree libwjb200.so (run here, no GPU):  python scripts/sass_summary.py r2
Counts the opcodes that prove which hardware path a kernel uses -- UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor load), UBLKCP
(cp.async.bulk), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), SYNCS (mbarrier), HMMA (mma.sync) -- next to plain
LDG / STG / LDS / STS / MUFU, and writes profiles/<tag>_sass_opcodes.md."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
so = ROOT / "whisperjav_b200" / "libwjb200.so"
txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "LDG", "STG", "LDS", "STS", "LDSM",
        "MUFU", "FFMA", "HFMA2", "ATOM", "RED", "BAR", "UCGABAR"]
cur, counts, total = None, collections.OrderedDict(), {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("wjb::", "").replace("void ", "")
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        total[cur] += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
rows = ["# %s: SASS opcode counts per kernel of libwjb200.so (sm_100a, `cuobjdump -sass`)\n" % tag,
        "Columns are instruction counts in the kernel's SASS (static, not executed counts).  UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load,",
        "UBLKCP = cp.async.bulk, LDTM/STTM = tcgen05.ld/st (TMEM), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = mma.sync (none expected).\n",
        "| kernel | instrs | " + " | ".join(KEYS) + " |", "|---|---|" + "---|" * len(KEYS)]
for k, c in counts.items():
    rows.append(f"| `{k[:70]}` | {total[k]} | " + " | ".join(str(c.get(x, 0) or "") for x in KEYS) + " |")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
rows.append("| **all kernels** | %d | " % sum(total.values()) + " | ".join(str(tot.get(x, 0)) for x in KEYS) + " |")
(ROOT / "profiles" / f"{tag}_sass_opcodes.md").write_text("\n".join(rows) + "\n")
print("\n".join(rows[-1:]))
