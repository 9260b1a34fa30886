"""Summarise a round's ncu captures (gpurun_out/*.ncu-rep, launches_<tag>.csv) into profiles/ (run in the build container):

    python tools/ncu_summarize2.py r02

  profiles/<tag>_ncu_summary.md   launch list of one forward (per-kernel share) + key raw metrics of every --set full capture
  profiles/<tag>_sass_summary.md  cuobjdump -sass mnemonic counts per kernel of the shipped library (UTCHMMA / LDTM / STTM / UTMALDG ...)
  profiles/ncu_traffic.json       DRAM bytes per launch of the kernels bench.py reports a roofline for (read by bench.py)
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
out = [f"# Round {tag} -- ncu evidence (tf32x3 mode, B=32, 12x288x288)\n",
       "Commands: tools/gpu_ncu_r2.sh (under gpurun, 1 GPU): `ncu --metrics gpu__time_duration.sum --clock-control none` launch list of",
       "`bench.py --steps 1 --warmup 3 --no-graph`, and `ncu --set full --clock-control none -k regex:<kernel>` captures.\n"]


def short(n):
    n = n.split("(")[0].replace("void ", "").replace("smaat::", "")
    return n.strip()


# ---- launch list: steady-state forward = one period of the kernel-name sequence
lf = f"gpurun_out/launches_{tag}.csv"
if os.path.exists(lf):
    rows = list(csv.reader(l for l in open(lf) if not l.startswith("==")))
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    seq = [(r[ik], float(r[iv].replace(",", ""))) for r in rows[1:] if len(r) > iv and r[im] == "gpu__time_duration.sum"]
    names = [short(n) for n, _ in seq]
    per = None
    for P in range(20, 200):
        base = max(0, len(names) - 2 * P - 5)
        if base + 2 * P <= len(names) and names[base:base + P] == names[base + P:base + 2 * P]:
            per = (base + P, base + 2 * P)
            break
    fwd = seq[per[0]:per[1]] if per else seq[-60:]
    agg = collections.OrderedDict()
    for n, v in fwd:
        k = re.sub(r"<.*", "", short(n))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in agg.values())
    out.append(f"## Launch list of one forward: {sum(c for c, _ in agg.values())} launches, per-kernel share (cold-cache, serialised under ncu)\n")
    out.append("| kernel | launches | us | share |\n|---|---|---|---|")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k} | {c} | {v / 1000.0:.1f} | {100 * v / tot:.1f}% |")
    out.append(f"| **total** | {sum(c for c, _ in agg.values())} | {tot / 1000.0:.1f} | 100% |\n")

# ---- full captures
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
traffic = {}
UNIT = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
out.append("## `ncu --set full` captures (reports stay in gpurun_out/; key raw metrics, DRAM bytes per launch)\n")
for f in sorted(os.listdir("gpurun_out")):
    if not (f.startswith(f"prof_{tag}_") and f.endswith(".ncu-rep")):
        continue
    txt = subprocess.run(["ncu", "-i", os.path.join("gpurun_out", f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(txt.splitlines()))
    if len(rr) < 3:
        continue
    h, units = rr[0], rr[1]
    out.append(f"### {f}\n")
    for r in rr[2:]:
        d = dict(zip(h, r))
        u = dict(zip(h, units))
        kn = short(d.get("Kernel Name", ""))[:100]
        vals = "; ".join(f"{w}={d[w]}{(' ' + u[w]) if u.get(w) else ''}" for w in want if w in d and d[w] != "")
        out.append(f"- `{kn}` grid {d.get('Grid Size', '')} block {d.get('Block Size', '')}: {vals}")
        try:
            rd = float(d["dram__bytes_read.sum"].replace(",", "")) * UNIT.get(u["dram__bytes_read.sum"], 1.0)
            wr = float(d["dram__bytes_write.sum"].replace(",", "")) * UNIT.get(u["dram__bytes_write.sum"], 1.0)
            key = "smaat_dsconv_fwd" if "dsconv" in kn else ("smaat_pw1x1_fwd" if "pw1x1_tc" in kn else ("smaat_dw3x3_fwd" if kn.startswith("dw3x3_kernel") else None))
            if key and (key not in traffic or rd + wr > traffic[key]["dram_bytes_per_launch"]):
                traffic[key] = {"dram_bytes_per_launch": rd + wr,
                                "note": f"ncu --set full, dram read+write of the largest captured launch of {kn.split('<')[0]} "
                                        f"(grid {d.get('Grid Size', '')}), profiles/{tag}_ncu_summary.md"}
        except Exception:
            pass
    out.append("")
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")
if traffic:
    json.dump(traffic, open("profiles/ncu_traffic.json", "w"), indent=1, sort_keys=True)

# ---- SASS evidence of the shipped library
sass = subprocess.run(["cuobjdump", "-sass", "smaat_unet_b200/libsmaat_b200.so"], capture_output=True, text=True).stdout
cur, counts = None, collections.OrderedDict()
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|FFMA2|HMMA|SYNCS)\b")
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "").replace("smaat::", "")
        counts[cur] = collections.Counter()
        continue
    if cur:
        for mm in pat.findall(line):
            counts[cur][mm] += 1
so = [f"# Round {tag} -- SASS mnemonics of smaat_unet_b200/libsmaat_b200.so (`cuobjdump -sass`, counted per kernel)\n",
      "UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / tcgen05.st, UTMALDG / UTMAPF = TMA tensor load / L2 prefetch, FFMA2 = fma.rn.f32x2.\n",
      "| kernel | UTCHMMA | LDTM | STTM | UTMALDG | UTMAPF | FFMA2 | HMMA |", "|---|---|---|---|---|---|---|---|"]
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
    if any(c[x] for x in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "FFMA2")):
        so.append(f"| `{k[:90]}` | {c['UTCHMMA']} | {c['LDTM']} | {c['STTM']} | {c['UTMALDG']} | {c['UTMAPF']} | {c['FFMA2']} | {c['HMMA']} |")
so.append(f"| **library total** | {tot['UTCHMMA']} | {tot['LDTM']} | {tot['STTM']} | {tot['UTMALDG']} | {tot['UTMAPF']} | {tot['FFMA2']} | {tot['HMMA']} |")
open(f"profiles/{tag}_sass_summary.md", "w").write("\n".join(so) + "\n")
print("\n".join(out[:30]))
print("\n".join(so[-6:]))
