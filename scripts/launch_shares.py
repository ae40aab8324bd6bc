"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --csv) into per-kernel totals and shares.

    python scripts/launch_shares.py gpurun_out/launches.csv > profiles/rNN_launch_shares.txt

Only launches from the first render_tc_fwd_kernel<0> on are counted (everything before is one-time set-up: weight
initialisation, CLIP plane packing); ncu times every launch cold-cache and serialised, so read the SHARES."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    start = next((i for i, r in enumerate(rows) if "render_tc_fwd_kernel<(int)0>" in r["Kernel Name"] or
                  "render_tc_fwd_kernel<0>" in r["Kernel Name"]), 0)
    rows = rows[start:]
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1.0)
        a = agg.setdefault(r["Kernel Name"][:90], [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
    steps = sum(n for k, (n, _) in agg.items() if "render_tc_bwd_kernel<(int)0" in k or "render_tc_bwd_kernel<0" in k) / 2.0
    print("# %d launches from the first render launch on, %.1f ms of kernel time, %.1f training steps (2 render backward launches each)"
          % (len(rows), total / 1e6, steps))
    print("# kernel, launches, launches/step, total_ns, share")
    ours = 0.0
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        mine = any(t in k for t in ("sct::", "scr::", "scclip::", "sctc::", "scch::", "chamfer"))
        ours += ns if mine else 0.0
        print("%s, %d, %.1f, %d, %.4f" % (k, n, n / max(steps, 1e-9), ns, ns / total))
    print("# share of kernel time in this library's own kernels: %.3f" % (ours / total))


if __name__ == "__main__":
    main(sys.argv[1])
