"""Per-source-line instruction and stall-sample shares of an .ncu-rep: python tools/ncu_lines.py report.ncu-rep [top]
(needs -lineinfo at compile time and --import-source on at capture time)."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = None
    per_line = {}
    for r in rows:
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 2 or not r[0]:
            continue
        try:
            line = int(r[0])
            inst = int(r[hdr.index("Instructions Executed")])
            thr = int(r[hdr.index("Thread Instructions Executed")])
            samp = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        a = per_line.setdefault(line, [0, 0, 0, r[1]])
        a[0] += inst; a[1] += thr; a[2] += samp
    ti = sum(v[0] for v in per_line.values()) or 1
    ts = sum(v[2] for v in per_line.values()) or 1
    print(f"total warp instructions {ti}, stall samples {ts}")
    print("| line | warp-inst share | avg lanes | sample share | source |\n|---|---|---|---|---|")
    for line, (inst, thr, samp, src) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"| {line} | {100 * inst / ti:.1f}% | {thr / max(inst, 1):.1f} | {100 * samp / ts:.1f}% | `{src.strip()[:110]}` |")


if __name__ == "__main__":
    main()
