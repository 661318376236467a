"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>`):
python tools/launch_summary.py X.csv > profiles/Y.md   — per kernel: launches, total and mean duration, share of the run."""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[h], rows[h + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    total = 0.0
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        name = r[ki].split("(")[0].replace("void ", "").replace("rtp::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
        total += v * scale
    print(f"# kernel launch list of `{sys.argv[1].split('/')[-1]}`\n")
    print("Per-launch `gpu__time_duration.sum` under ncu (cold caches, serialised launches: shares, not absolute times, are what counts).\n")
    print(f"{sum(a[0] for a in agg.values())} launches, {total / 1e3:.2f} ms in kernels.\n")
    print("| kernel | launches | total ms | mean us | share |\n|---|---|---|---|---|")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {n} | {us / 1e3:.3f} | {us / n:.1f} | {100 * us / total:.1f}% |")


if __name__ == "__main__":
    main()
