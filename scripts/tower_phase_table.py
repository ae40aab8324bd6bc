"""Summarises an ncu per-phase launch list of the CLIP tower (scripts/profile_clip_tower.py under
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --csv)."""
import csv
import sys

for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    idx = {h: i for i, h in enumerate(rows[0])}
    data = {}
    for r in rows[1:]:
        data.setdefault(int(r[idx["ID"]]), {})[r[idx["Metric Name"]]] = float(r[idx["Metric Value"]].replace(",", ""))
    ids = sorted(data)
    layers = (len(ids) - 5) // 5
    names = ["im2col", "patch", "tokens"] + ["qkv", "attn", "out", "fc1", "fc2"] * layers + ["head", "l2norm"]
    tot, agg = 0.0, {}
    for i, k in enumerate(ids):
        d = data[k]
        t = d["gpu__time_duration.sum"] / 1e3
        tot += t
        a = agg.setdefault(names[i], [0.0, 0.0, 0])
        a[0] += t
        a[1] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
        a[2] += 1
    print("%s: %d launches, sum of per-launch device time %.1f us (cold cache, serialised; each carries the ~10 us launch + TMEM set-up of a 200 KB-smem kernel)" % (f, len(ids), tot))
    for n, (t, p, c) in agg.items():
        print("  %-7s n=%2d  avg %6.1f us  tensor pipe %5.1f %%  total %7.1f us" % (n, c, t / c, p / c, t))
