"""Per-source-line totals of an ncu source page (cuda,sass view) of one kernel launch.
   python profiles/tools/src_hot.py REPORT.ncu-rep KERNEL_REGEX LAUNCH_SKIP [TOP]
Prints the lines that issue the most warp instructions with their lane occupancy and stall samples."""
import csv, io, subprocess, sys, collections

rep, rx, skip = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{rx}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname, head = None, None
agg = collections.OrderedDict()
cur = None
for r in rows:
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        head = r
        continue
    if head is None or len(r) < len(head):
        continue
    if r[0]:  # a source line starts a group
        cur = (fname, int(r[0]), r[1].strip()[:90])
        agg.setdefault(cur, [0, 0, 0, 0])
        continue
    if cur is None or r[2] in ("", "..."):
        continue
    def num(name):
        v = r[head.index(name)]
        return int(v) if v.isdigit() else 0
    a = agg[cur]
    a[0] += num("Instructions Executed")
    a[1] += num("Thread Instructions Executed")
    a[2] += num("# Samples")
    a[3] += num("stall_barrier") + num("stall_long_sb")
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[2] for a in agg.values()) or 1
print(f"total warp instructions {tot_i/1e6:.1f} M, thread instructions {sum(a[1] for a in agg.values())/1e6:.1f} M, samples {tot_s}")
byfile = collections.OrderedDict()
for k, a in agg.items():
    b = byfile.setdefault(k[0], [0, 0, 0])
    b[0] += a[0]; b[1] += a[1]; b[2] += a[2]
for f, b in byfile.items():
    if b[0]:
        print(f"  {f:34s} warp-instr {100*b[0]/tot_i:5.1f}%  lanes {b[1]/b[0]:4.1f}  samples {100*b[2]/tot_s:5.1f}%")
# optional line ranges "file:lo-hi,..." as 5th argument: totals per range
if len(sys.argv) > 5:
    for spec in sys.argv[5].split(","):
        f, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        t = [0, 0, 0]
        for k, a in agg.items():
            if k[0] == f and lo <= k[1] <= hi:
                t[0] += a[0]; t[1] += a[1]; t[2] += a[2]
        if t[0]:
            print(f"  {spec:34s} warp-instr {100*t[0]/tot_i:5.1f}%  lanes {t[1]/t[0]:4.1f}  samples {100*t[2]/tot_s:5.1f}%")
print("file:line  warp-instr%  lanes  samples%  (barrier+long_sb samples%)  source")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    if a[0] == 0:
        continue
    print(f"{k[0]}:{k[1]:<5d} {100*a[0]/tot_i:5.1f}%  {a[1]/a[0]:4.1f}  {100*a[2]/tot_s:5.1f}%  {100*a[3]/tot_s:5.1f}%  {k[2]}")
