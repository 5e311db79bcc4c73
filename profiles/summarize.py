"""Turn the outputs of profiles/run_ncu.sh (gpurun_out/) into the tracked summaries under profiles/.
   python profiles/summarize.py v8 [scans_per_step] [round-dir, default r2]
Writes profiles/<round>/search_accum_<tag>_raw.csv (the ten K3 launches of one step, ncu --set full),
copies the launch list, rewrites profiles/traffic.json and prints the markdown tables for SUMMARY.md."""
import csv, collections, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
sps = int(sys.argv[2]) if len(sys.argv) > 2 else 256
out = os.path.join(ROOT, "gpurun_out")
rnd = sys.argv[3] if len(sys.argv) > 3 else "r2"
dst = os.path.join(ROOT, "profiles", rnd)
os.makedirs(dst, exist_ok=True)
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
           "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_tensor_op_dmma.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "launch__grid_size", "launch__block_size"]

raw = subprocess.run(["ncu", "-i", os.path.join(out, f"search_accum_{tag}.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units, launches = rows[0], rows[1], rows[2:]
table = []
for m in METRICS:
    if m in head:
        i = head.index(m)
        table.append([m, units[i]] + [r[i].replace(",", "") for r in launches])
with open(os.path.join(dst, f"search_accum_{tag}_raw.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch{k + 1}" for k in range(len(launches))])
    w.writerows(table)

def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
rd = next(t for t in table if t[0] == "dram__bytes_read.sum")
wr = next(t for t in table if t[0] == "dram__bytes_write.sum")
per = [to_bytes(a, rd[1]) + to_bytes(b, wr[1]) for a, b in zip(rd[2:], wr[2:])]
json.dump({"workload": "c2", "scans_per_step": sps, "kernel": "search_accum_kernel<GN_P2PLANE, 128>",
           "dram_bytes_per_launch": int(sum(per) / len(per)),
           "source": f"profiles/{rnd}/search_accum_{tag}_raw.csv: mean over the {len(per)} launches of one {sps}-scan step of "
                     "dram__bytes_read.sum + dram__bytes_write.sum (one ncu --set full capture, final build); per launch "
                     + str([round(p / 1e6, 1) for p in per]) + " MB"},
          open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)

print("| metric | unit | " + " | ".join(f"L{k + 1}" for k in range(len(launches))) + " |")
print("|---|---|" + "---|" * len(launches))
for t in table:
    print("| " + " | ".join([t[0], t[1]] + [f"{float(v):.4g}" if v.replace('.', '', 1).replace('e+', '', 1).isdigit() else v for v in t[2:]]) + " |")

# launch list: one step
src = os.path.join(out, f"launches_{tag}_c2.csv")
if not os.path.exists(src):
    src = os.path.join(out, f"launches_{tag}.csv")
shutil.copy(src, os.path.join(dst, f"launches_{tag}_c2.csv"))
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
L = [(r[kn].split("(")[0].replace("void ", "").replace("ssf::", ""),
      float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1.0)) for r in rows[hi + 1:] if len(r) > mv and r[mv]]
idx = [i for i, (n, _) in enumerate(L) if n.startswith("vb_init")]
big = [(a, b) for a, b in zip(idx, idx[1:] + [len(L)]) if sum(1 for n, _ in L[a:b] if n.startswith("search_accum")) >= 10]
a, b = big[-1]
while b > a and not L[b - 1][0].startswith("results"):
    b -= 1
agg = collections.OrderedDict()
for n, t in L[a:b]:
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += t
tot = sum(v[1] for v in agg.values())
print(f"\nLaunch list of one {sps}-scan step ({tot:.0f} us serialised, {b - a} launches):\n")
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    if t / tot >= 0.01:
        print(f"| {n} | {c} | {t:.1f} | {t / c:.1f} | {100 * t / tot:.1f}% |")
