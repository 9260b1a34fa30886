"""Summarise the round's ncu captures into profiles/<tag>_ncu_summary.md (run in the build container)."""
import csv, subprocess, sys, collections, os
tag = sys.argv[1] if len(sys.argv) > 1 else "r01c"
mode = "tf32x3"
out = [f"# Round 1 (final code) -- ncu evidence, {mode} mode\n",
       "Commands (under gpurun, 1 GPU, tools/gpu_ncu2.sh):\n",
       "    ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv python bench.py --mode tf32x3 --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-alt",
       "    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c <n> ... (same workload; tools/prof_train.py for the training kernels)\n"]
# ---- launch list: one forward = the launches between two consecutive occurrences of the first kernel of a forward
rows = list(csv.reader(l for l in open(f"gpurun_out/launches_{tag}_{mode}.csv") if not l.startswith("==")))
hdr = rows[0]; ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
seq = [(r[ik], float(r[iv].replace(",", ""))) for r in rows[1:] if len(r) > iv and r[im] == "gpu__time_duration.sum"]
names = [n for n, _ in seq]
def short(n):
    n = n.split("(")[0].replace("void ", "").replace("smaat::", "")
    return n.split("<")[0]
# steady state: find the period P of the launch sequence well after the cache-building first forward
per = None
for P in range(40, 200):
    base = 200
    if base + 2 * P < len(names) and names[base:base + P] == names[base + P:base + 2 * P]:
        # align the window to the start of a forward (the Cin=12 DS conv is the first kernel of a forward)
        off = next((j for j in range(base, base + P) if "dsconv_fused_kernel<64, 2, 32" in names[j] and "dsconv" not in names[j - 1] and "cbam" not in names[j - 1]), base)
        per = (off, off + P)
        break
fwd = seq[per[0]:per[1]] if per else seq[:100]
agg = collections.OrderedDict()
for n, v in fwd:
    k = short(n); a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in agg.values())
unit_div = 1000.0  # ns -> us
out.append(f"## Launch list of one forward (B=32, 12x288x288): {sum(c for c, _ in agg.values())} launches, per-kernel share (cold-cache, serialised under ncu)\n")
out.append("| kernel | launches | us | share |\n|---|---|---|---|")
for k, (c, v) in agg.items():
    out.append(f"| {k} | {c} | {v / unit_div:.1f} | {100 * v / tot:.1f}% |")
out.append(f"| **total** | {sum(c for c, _ in agg.values())} | {tot / unit_div:.1f} | 100% |\n")
# ---- full captures
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
out.append("## `ncu --set full` captures (reports kept in gpurun_out/, key raw metrics here; DRAM bytes are per launch)\n")
for rep in ("dsconv_" + mode, "pw_" + mode, "dw_unfused", "cbam_up", "train"):
    f = f"gpurun_out/prof_{tag}_{rep}.ncu-rep"
    if not os.path.exists(f): continue
    txt = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(txt.splitlines()))
    h = rr[0]
    out.append(f"### prof_{tag}_{rep}\n")
    for r in rr[2:]:
        d = dict(zip(h, r))
        kn = d.get("Kernel Name", "")[:90]
        grid, blk = d.get("Grid Size", ""), d.get("Block Size", "")
        vals = "; ".join(f"{w}={d[w]}" for w in want if w in d and d[w] != "")
        out.append(f"- `{kn}` grid {grid} block {blk}: {vals}")
    out.append("")
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:40]))
