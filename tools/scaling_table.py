"""Writes profiles/r02_scaling.md from the bench lines of the four scaling runs:
    python tools/scaling_table.py gpurun_out/r02_bench_{1,2,4,8}gpu.json > profiles/r02_scaling.md
Each file holds the stdout of `bench.py --gpus N --steps 20 --warmup 3` (N > 1: under torchrun, one rank per GPU); the last line is the JSON."""
import json
import sys


def load(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def main():
    runs = {}
    for p in sys.argv[1:]:
        d = load(p)
        runs[d["n_gpus"]] = d
    ns = sorted(runs)
    first, last = ns[0], ns[-1]

    def row(label, unit, get, eff="weak", fmt="{:.1f}"):
        vals = []
        for n in ns:
            try:
                vals.append(get(runs[n]))
            except (KeyError, TypeError):
                vals.append(None)
        cells = [fmt.format(v) if v is not None else "—" for v in vals]
        e = "—"
        if eff and vals[0] and vals[-1]:
            e = "{:.2f}".format(vals[-1] / (last / first * vals[0]) if eff == "weak" else vals[0] / (last / first * vals[-1]))
        print(f"| {label} | {unit} | " + " | ".join(cells) + f" | {e} |")

    print("# Round 2: `bench.py` at " + ", ".join(str(n) for n in ns) + " B200 (one box each, `gpurun --gpus N`, `torchrun --nproc-per-node N`, `--steps 20 --warmup 3`)")
    print()
    print("Made by `tools/scaling_table.py` from the bench lines of the four runs (`gpurun_out/r02_bench_{1,2,4,8}gpu.json`, scratch; the 8-GPU line is kept as")
    print("`profiles/r02_bench_8gpu.json`). Boxes differ from call to call (host cores per run: "
          + ", ".join(f"N = {n}: {runs[n]['cpu_baseline']['cores'] if runs[n].get('cpu_baseline') else '?'}" for n in ns)
          + f"); efficiency = value(N) / (N x value({first})) for weak legs, time({first}) / (N x time(N)) for the strong one.")
    print()
    print("| leg | unit | " + " | ".join(f"N = {n}" for n in ns) + f" | efficiency at {last} |")
    print("|---|---|" + "---|" * len(ns) + "---|")
    row("C2 primary batch, device-resident (`value`, weak)", "Mrays/s", lambda d: d["value"])
    row("C2 through `rtp_trace_closest`, host buffers (`e2e`)", "Mrays/s", lambda d: d["e2e"]["value"])
    row("... its bytes over the host links", "GB/s", lambda d: d["e2e"]["link_gbs"], eff=None)
    row("... plain pinned copies of the same bytes (`link_ceiling_gbs`)", "GB/s", lambda d: d["e2e"]["link_ceiling_gbs"], eff=None)
    row("C2 through `rtp_trace_camera` (hits only travel)", "Mrays/s", lambda d: d["e2e_camera"]["value"])
    row("C3 2^24 incoherent rays per GPU (weak)", "Mrays/s", lambda d: d["incoherent"]["mrays_per_s"])
    row("C1 640x360, 16 spp per GPU (weak, reduce to rank 0)", "Msamples/s", lambda d: d["render"]["samples_per_s"] / 1e6)
    row("... through `rtp_render` with the frame copied to the host (`render.e2e`)", "Msamples/s", lambda d: d["render"]["e2e"]["value"] / 1e6)
    row("C4 1920x1080, 256 spp in ONE frame, rows over ranks (STRONG, gather)", "ms/frame", lambda d: d["render_c4"]["ms_per_frame"], eff="strong")
    row("C4 scene, 32 spp per GPU (weak, reduce)", "Msamples/s", lambda d: d["render_c4_weak"]["samples_per_s"] / 1e6)
    row("C5 10.17 M triangles, 3840x2160, 128 spp per GPU, rows over ranks (1024 spp at N = 8 = the named config)", "Msamples/s",
        lambda d: d["render_c5"]["samples_per_s"] / 1e6)
    row("... C5 frame", "ms/frame", lambda d: d["render_c5"]["ms_per_frame"], eff=None)
    row("... C5 scene build on the device (max over ranks)", "s", lambda d: d["render_c5"]["scene_build_s"], eff=None, fmt="{:.2f}")
    print()
    print("One host call, N devices (`multi_device_abi`: rank 0 drives all GPUs through `rtp_scene_create_multi` + ONE `rtp_render` / `rtp_trace_closest`; "
          "other ranks wait in a gloo barrier):")
    print()
    print("| N | C4 scene 1080p x 64 spp, 1 device | N devices | speed-up | bit-identical | C2 batch through one `rtp_trace_closest`: 1 device | N devices |")
    print("|---|---|---|---|---|---|---|")
    for n in ns:
        m = runs[n].get("multi_device_abi")
        if not m or n == 1:
            continue
        t = m["trace_closest_host_buffers"]
        print(f"| {n} | {m['ms_per_frame_1_device']:.1f} ms | {m[f'ms_per_frame_{n}_devices']:.1f} ms | {m['speedup']:.2f} | {m['bit_identical_to_1_device']} | "
              f"{t['mrays_per_s_1_device']:.1f} Mrays/s | {t[f'mrays_per_s_{n}_devices']:.1f} Mrays/s |")
    print()
    d = runs[first]
    c = d["cpu_baseline"]
    print(f"CPU arms on the N = {first} box ({c['cores']} cores, oracle port): C2 {c['value']:.1f} Mrays/s (4 threads as shipped: {c['as_shipped_4_threads']['value']:.1f}), "
          f"C3 {d['incoherent']['cpu_baseline']['value']:.1f} Mrays/s, C1 {d['render']['cpu_baseline']['value'] / 1e6:.1f} Msamples/s, "
          f"C4 {d['render_c4']['cpu_baseline']['value'] / 1e6:.1f} Msamples/s, C5 {d['render_c5']['cpu_baseline']['value'] / 1e6:.2f} Msamples/s.")
    print("Clocks during the timed regions: " + ", ".join(f"N = {n}: {runs[n]['clocks']['sm_mhz']} MHz, reasons {runs[n]['clocks']['reasons']}" for n in ns) + ".")


if __name__ == "__main__":
    main()
