"""Turn gpurun_out/launches.csv and gpurun_out/prof_*.ncu-rep (scripts/gpu_profile.sh) into the committed
summaries under profiles/ (run here, no GPU needed):  python scripts/summarize_profiles.py r1"""
import collections
import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = ROOT / "profiles"
out.mkdir(exist_ok=True)
G = ROOT / "gpurun_out"


def short(name):
    return re.sub(r"^void ", "", re.sub(r"\(.*", "", name)).replace("wjb::", "")


# ---- launch list
lines = [l for l in open(G / "launches.csv") if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    u = row["Metric Unit"]
    v = v / 1000.0 if u == "ns" else (v * 1000.0 if u == "ms" else v)
    seq.append((short(row["Kernel Name"]), v))
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in seq:
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
md = [f"# {tag}: ncu launch list of one shortened hot-path step (large-v3, batch 64, 4 sampled tokens)\n",
      "`ncu --metrics gpu__time_duration.sum --clock-control none` over `scripts/profile_step.py --batch 64 --tokens 4`, the kernel nodes of the decode graph are",
      "listed individually.  Times are cold-cache and serialised: compare shares, not absolutes.\n",
      f"total {tot / 1000:.2f} ms over {len(seq)} launches\n", "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    md.append(f"| `{k[:80]}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} | {100 * v[1] / tot:.1f}% |")
idx = [i for i, (n, _) in enumerate(seq) if "sample_kernel" in n]
if len(idx) >= 2:
    step = seq[idx[-2] + 1: idx[-1] + 1]
    a2 = collections.defaultdict(lambda: [0, 0.0])
    for n, v in step:
        a2[n][0] += 1
        a2[n][1] += v
    md += ["", f"## one decode step ({len(step)} kernels, {sum(v for _, v in step):.1f} us)\n", "| kernel | launches | total us | avg us |", "|---|---|---|---|"]
    for k, v in sorted(a2.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| `{k[:80]}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} |")
idx = [i for i, (n, _) in enumerate(seq) if "beam_select_kernel" in n]
if len(idx) >= 2:
    step = seq[idx[-2] + 1: idx[-1] + 1]
    a2 = collections.defaultdict(lambda: [0, 0.0])
    for n, v in step:
        a2[n][0] += 1
        a2[n][1] += v
    md += ["", f"## one beam-search step, 64 windows x beam 2 ({len(step)} kernels, {sum(v for _, v in step):.1f} us)\n", "| kernel | launches | total us | avg us |",
           "|---|---|---|---|"]
    for k, v in sorted(a2.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| `{k[:80]}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} |")
(out / f"{tag}_launches.md").write_text("\n".join(md) + "\n")

# ---- full captures
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg"]
reps = sorted(G.glob("prof_*.raw.csv")) + sorted(G.glob("prof_*.ncu-rep"))
for rep in reps:
    if rep.suffix == ".csv":
        raw = rep.read_text()
    else:
        raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    raw = "\n".join(l for l in raw.splitlines() if not l.startswith("=="))
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"kernel": short(r[ix["Kernel Name"]])}
        for k in KEYS:
            if k in ix:
                d[k] = f"{r[ix[k]]} {units[ix[k]]}".strip()
        res.append(d)
    (out / f"{tag}_{rep.name.split('.')[0]}.json").write_text(json.dumps(res, indent=1))
print("wrote", sorted(p.name for p in out.glob(f"{tag}_*")))
