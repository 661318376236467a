"""Summarise an .ncu-rep (read here, without a GPU) into markdown: python tools/ncu_summary.py report.ncu-rep > profiles/x.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_bytes.sum.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__cycles_active.avg",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print(f"# ncu summary of `{rep.split('/')[-1]}`\n")
    print("Captured with `ncu --set full --clock-control none --import-source on` under gpurun on one B200; read offline with")
    print("`ncu -i <rep> --page raw --csv`. Per-launch values (one column per captured launch).\n")
    print("| metric | unit | " + " | ".join(r[name_i].split("(")[0][-48:] for r in data) + " |")
    print("|---|---|" + "---|" * len(data))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| `{k}` | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    print("\n## warp stall samples (smsp__pcsamp_warps_issue_stalled_*)\n")
    stalls = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                stalls.append((sum(float(r[i].replace(",", "")) for r in data), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            except ValueError:
                pass
    tot = sum(v for v, _ in stalls) or 1.0
    print("| reason | samples | share |\n|---|---|---|")
    for v, h in sorted(stalls, reverse=True)[:12]:
        print(f"| {h} | {v:.0f} | {100 * v / tot:.1f}% |")


if __name__ == "__main__":
    main()
