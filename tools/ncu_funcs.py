"""Per-function instruction and stall-sample shares of an .ncu-rep (source lines of rtp_device.cu grouped by the function that
contains them): python tools/ncu_funcs.py report.ncu-rep"""
import bisect
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, per = None, {}
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r or not r[0]:
            continue
        try:
            line, inst = int(r[0]), int(r[hdr.index("Instructions Executed")])
            thr, smp = int(r[hdr.index("Thread Instructions Executed")]), int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        a = per.setdefault(line, [0, 0, 0])
        a[0] += inst; a[1] += thr; a[2] += smp
    src = open(os.path.join(ROOT, "raytracing-potato_b200", "csrc", "rtp_device.cu")).read().split("\n")
    funcs = []
    for i, l in enumerate(src, 1):
        m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|static|__host__).*?(\w+)\s*\(", l)
        if m and not l.startswith(" "):
            funcs.append((i, m.group(1)))
    starts = [f[0] for f in funcs]
    agg = {}
    for line, (i, t, s) in per.items():
        k = bisect.bisect_right(starts, line) - 1
        name = funcs[k][1] if k >= 0 else "?"
        if name == "__launch_bounds__":
            name = "trace_persistent_kernel (loop control, refill)"
        a = agg.setdefault(name, [0, 0, 0])
        a[0] += i; a[1] += t; a[2] += s
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[2] for v in agg.values()) or 1
    print(f"total warp instructions {ti}, stall samples {ts}")
    print("| function | warp-inst share | avg lanes | sample share |\n|---|---|---|---|")
    for n, (i, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if i * 1000 >= ti:
            print(f"| `{n}` | {100 * i / ti:.1f}% | {t / max(i, 1):.1f} | {100 * s / ts:.1f}% |")


if __name__ == "__main__":
    main()
