#!/usr/bin/env python
"""bench.py — headline measurement of the hot path on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus 1 ...            # the reference's CPU path (oracle port) on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU

Workload (BASELINE.json configs[1], "C2"): the 1920x1080 primary-ray closest-hit batch against the bunny BVH
(4,968 triangles + ground sphere, 9,937 nodes). One step = one pass of the closest-hit path over the batch
(2,073,600 rays per GPU). At N > 1 the path shards with no data-path collective: rank r traces sub-sample r of
an N-times supersampled 1920x1080 primary batch (weak scaling, 2,073,600 rays per rank).

Printed JSON (one line, rank 0): `value` = Mrays/s with rays and hits resident in HBM; `e2e` = the same metric
through rtp_trace_closest with pinned HOST buffers (H2D + D2H inside the timed region); `roofline` for the
traversal kernel; `cpu_baseline` = the oracle port on the host cores; `render` = the C1 path-traced frame
(samples/s, Mrays/s) split by sample range with an NCCL sum of the accumulation buffers.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

W, H = 1920, 1080
N_RAYS = W * H
WORKLOAD = "C2: 1920x1080 primary-ray closest-hit batch vs bunny BVH (4968 triangles + ground sphere, 9937 nodes)"
RAY_BYTES, HIT_BYTES = 64, 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-render", action="store_true", help="skip the secondary C1 render measurement")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def bind_to_gpu_local_cores(local_rank):
    """One process per GPU: run this rank (and first-touch its pinned host buffers) on the cores NVML reports as local to its
    GPU, so that the host side of the e2e leg does not cross the socket interconnect. Returns the number of cores, or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, mask in enumerate(words) for b in range(64) if (mask >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def c2_scene_and_camera(scenes, api):
    """bunny scene (example_scenes.rs:309-350 with a Lambert bunny) and its camera at 1920x1080, lens 0"""
    sc = scenes.bunny_lambert()
    cam = api.Camera(W / H, sc.camera.fov, sc.camera.focal_dist, 0.0, sc.camera.transformation)
    return sc, cam


# ----------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    """samples SM clock / throttle reasons through NVML every 20 ms while `active`"""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.active = False
        self._stop = False
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        if nv is None:
            return
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop:
            if self.active:
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.02)

    def report(self):
        self._stop = True
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------------- reference arm

def run_reference(args):
    """The reference's own CPU implementation of the path (oracle port; the Rust crate cannot be built here),
    all host threads, same workload / metric / unit. Each step = one pass over the full C2 batch."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import oracle
    from rtp_b200 import api, scenes

    sc, cam = c2_scene_and_camera(scenes, api)
    o = oracle.Scene(sc)
    rays = oracle.camera_rays(cam, W, H)
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        o.hit_full(rays, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.hit_full(rays, threads=cores)
    dt = time.perf_counter() - t0
    v = args.steps * N_RAYS / dt / 1e6
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": N_RAYS},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                         "sample": f"full C2 batch (2,073,600 rays) x {args.steps} steps, oracle/rtp_oracle.c, {cores} pthreads"},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm

def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return

    os.environ["NCCL_DEBUG"] = os.environ.get("RTP_NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: one JSON line only
    import torch
    import torch.distributed as dist

    from rtp_b200 import _abi as A
    from rtp_b200 import api, scenes
    from rtp_b200 import dist as rdist

    rank, local_rank, world = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    affinity = bind_to_gpu_local_cores(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    api.init(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    sampler = ClockSampler(local_rank) if rank == 0 else None

    sc = scenes.bunny_lambert()
    scene = api.Scene(sc)
    cam = api.Camera(W / H, sc.camera.fov, sc.camera.focal_dist, 0.0, sc.camera.transformation)
    stream = torch.cuda.current_stream().cuda_stream

    # --- inputs: this rank's primary batch, generated on the device by the camera kernel (render.rs:32-52) -----------
    # two distinct ray buffers are rotated so that no step can reuse its rays from L2 (2 x 132.7 MB > 126 MB L2)
    d_rays = [torch.empty((N_RAYS, 8), dtype=torch.float64, device=dev) for _ in range(2)]
    if world == 1:
        api.camera_rays_device(cam, W, H, d_rays[0].data_ptr(), stream)
    else:
        # rank r = sub-sample r of a world-times supersampled frame: shift the pixel centres by (r+0.5)/world - 0.5
        i = torch.arange(W, device=dev, dtype=torch.float64).repeat(H)
        j = torch.arange(H, device=dev, dtype=torch.float64).repeat_interleave(W)
        off = (rank + 0.5) / world
        rays_np = _camera_rays_numpy(cam, (i.cpu().numpy() + off) / W, (j.cpu().numpy() + off) / H)
        d_rays[0].copy_(torch.from_numpy(rays_np))
    d_rays[1].copy_(d_rays[0])
    d_hits = torch.empty((N_RAYS, 2), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # --- value: device-resident batch, CUDA events on the launching stream ------------------------------------------------
    for k in range(args.warmup):
        scene.hit_device(d_rays[k & 1].data_ptr(), N_RAYS, d_hits.data_ptr(), stream)
    barrier()
    if sampler:
        sampler.active = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        scene.hit_device(d_rays[k & 1].data_ptr(), N_RAYS, d_hits.data_ptr(), stream)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if sampler:
        sampler.active = False
    ms_t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_max = float(ms_t.item())
    value = args.steps * N_RAYS * world / (ms_max * 1e-3) / 1e6
    kernel_ms = ms / args.steps  # one kernel per step, back to back on one stream

    # --- e2e: host buffers through rtp_trace_closest (chunked H2D -> kernel -> D2H pipeline inside the call) -----------
    h_rays = api.PinnedBuffer((N_RAYS,), A.RAY_DTYPE)
    h_hits = api.PinnedBuffer((N_RAYS,), A.HIT_DTYPE)
    h_rays.array.view(np.float64).reshape(-1, 8)[:] = d_rays[0].cpu().numpy()
    e2e_launches = 0
    for _ in range(max(args.warmup, 1)):
        _, st = scene.hit(h_rays.array, out=h_hits.array, stats=True)
        e2e_launches = int(st.kernel_launches)
    e2e_steps = max(3, min(args.steps, 50))
    barrier()
    if sampler:
        sampler.active = True
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        scene.hit(h_rays.array, out=h_hits.array)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if sampler:
        sampler.active = False
    dt_t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
    e2e_value = e2e_steps * N_RAYS * world / float(dt_t.item()) / 1e6
    # the e2e result must equal the device-resident result
    same = bool((torch.from_numpy(h_hits.array.view(np.float64).reshape(-1, 2).copy()).to(dev).view(torch.int64) == d_hits_after(scene, d_rays[0], d_hits, stream).view(torch.int64)).all())

    # --- e2e, camera-driven: the same batch through rtp_trace_camera (H2D = the camera, rays made on the device, D2H = hits) ----
    scene.hit_camera(cam, W, H, out=h_hits.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        scene.hit_camera(cam, W, H, out=h_hits.array)
    torch.cuda.synchronize()
    dtc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dtc, op=dist.ReduceOp.MAX)
    e2e_camera = e2e_steps * N_RAYS * world / float(dtc.item()) / 1e6
    same_cam = bool((torch.from_numpy(h_hits.array.view(np.float64).reshape(-1, 2).copy()).to(dev).view(torch.int64) == d_hits.view(torch.int64)).all()) if world == 1 else None

    # --- roofline inputs: work counters of one counted pass (outside any timed region) ---------------------------------
    cst = scene.hit_device_counted(d_rays[0].data_ptr(), N_RAYS, d_hits.data_ptr())

    # --- secondary: C3 incoherent batch (2^22 of the 2^24 rays per rank; secondary-ray divergence stress) ----------------------
    incoherent = None
    if not args.no_render:
        n3 = 1 << 22
        d3 = torch.from_numpy(scenes.incoherent_rays(n3, first=rank * n3).view(np.float64).reshape(-1, 8)).to(dev)
        h3 = torch.empty((n3, 2), dtype=torch.float64, device=dev)
        for _ in range(3):
            scene.hit_device(d3.data_ptr(), n3, h3.data_ptr(), stream)
        barrier()
        e0.record()
        reps3 = max(3, min(args.steps, 20))
        for _ in range(reps3):
            scene.hit_device(d3.data_ptr(), n3, h3.data_ptr(), stream)
        e1.record()
        barrier()
        ms3 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms3, op=dist.ReduceOp.MAX)
        c3 = scene.hit_device_counted(d3.data_ptr(), n3, h3.data_ptr())
        incoherent = {"workload": "C3 subset: 2^22 incoherent rays per GPU vs bunny BVH (268 MB of rays > L2)", "mrays_per_s": reps3 * n3 * world / (float(ms3.item()) * 1e-3) / 1e6,
                      "per_ray": {"node_visits": c3.node_visits / n3, "leaf_gates": c3.leaf_gates / n3, "triangle_tests": c3.triangle_tests / n3, "sphere_tests": c3.sphere_tests / n3}}
        del d3, h3

    # --- secondary: C1 render (640x360, 16 spp per rank, depth 8), sample ranges + NCCL sum -----------------------------
    render = render_c4 = None
    if not args.no_render:
        render = bench_render(args, torch, dist, api, A, scene, sc, dev, rank, world, stream, barrier)
        sc4 = scenes.demo()
        scene4 = api.Scene(sc4)
        render_c4 = bench_render(args, torch, dist, api, A, scene4, sc4, dev, rank, world, stream, barrier, rw=1920, rh=1080, spp_rank=8,
                                 name="C4 subset: demo scene (glass bunny, Lambert bunny, earthmap sphere, light, metal ground, sky)", max_steps=5)
        scene4.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    algo_bytes = N_RAYS * (RAY_BYTES + HIT_BYTES)
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("trace_persistent_kernel_c2_dram_bytes_per_launch")
    except Exception:
        pass
    clocks = sampler.report()
    sm_hz = (clocks["sm_mhz"] or 1965) * 1e6
    # f64 work of the reference algorithm per launch: 24 ops per slab test, 75 per triangle test, 20 per sphere test
    f64_ops = 24 * cst.leaf_gates + 75 * cst.triangle_tests + 20 * cst.sphere_tests  # inner nodes are culled in f32
    fp64_peak, fp64_src = 148 * 64 * sm_hz, "148 SM x 64 FP64 lanes x median SM clock during the run (nominal)"
    try:
        import ctypes

        g = ctypes.c_double(0.0)
        if A.load().rtp_probe_fp64(ctypes.byref(g)) == 0 and g.value > 0:
            fp64_peak, fp64_src = g.value * 1e9, "measured: rtp_probe_fp64 (independent DMUL+DADD chains, best of 3)"
    except Exception:
        pass
    scene_bytes = 128 * cst.node_visits + 128 * (cst.triangle_tests + cst.sphere_tests)  # one 128 B DWide per node visit, one 128 B DPrim per tested leaf

    line = {
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": WORKLOAD,
            "rays_per_step_per_gpu": N_RAYS, "sharding": "rank r traces sub-sample r of a world-times supersampled primary batch; no data-path collective",
            "l2": "inputs larger than L2: two 132.7 MB ray buffers are rotated between steps (265 MB > 126 MB L2); no flush needed",
            "host_cores_bound_per_rank": affinity,
        },
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": N_RAYS * RAY_BYTES, "d2h_bytes_per_step": N_RAYS * HIT_BYTES,
                "steps": e2e_steps, "api": "rtp_trace_closest (pinned host buffers, 256Ki-ray chunks on 3 streams)", "gpu_launches_per_step": e2e_launches,
                "matches_device_result": same},
        "e2e_camera": {"value": e2e_camera, "unit": "Mrays/s", "h2d_bytes_per_step": 168, "d2h_bytes_per_step": N_RAYS * HIT_BYTES, "steps": e2e_steps,
                       "api": "rtp_trace_camera: Camera::shoot (render.rs:32-52) on the device + closest hit; the reference never materialises a ray array, its input is the camera",
                       "matches_device_result": same_cam, "note": "secondary figure; `e2e` above is the strict one with the rays in host memory"},
        "gpu_launches": args.steps,
        "clocks": clocks,
        "roofline": {
            "kernel": "trace_persistent_kernel<COUNT=false, OUT_HIT, LIST=false, ANY=1> (any-order walk, DESIGN.md 4b)", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": kernel_ms,
            "note": "the bunny scene (0.9 MB of culling nodes + primitives) is cache-resident, so HBM carries only the 80 B/ray stream; the kernel is bound by instruction issue and L1 latency (profiles/): see issue/fp64 below",
            "fp64": {"ops_per_launch": f64_ops, "achieved_gops": f64_ops / (kernel_ms * 1e-3) / 1e9, "peak_gops": fp64_peak / 1e9,
                     "frac": f64_ops / (kernel_ms * 1e-3) / fp64_peak, "peak_source": fp64_src},
            "l2": {"bytes_per_launch": scene_bytes, "achieved_gbs": scene_bytes / (kernel_ms * 1e-3) / 1e9},
            "l1_note": "scene bytes are served by L1/L2, not HBM",
            "issue": {"note": "binding resource per ncu (profiles/r01_trace_any_c2.md): smsp__issue_active 61 % of peak at 22.9 of 32 lanes per instruction, 70 warp-instructions per ray, 29 % warp occupancy (96 registers, 5 blocks of 128 per SM), L1 hit rate 68 %; HBM 7 % of peak"},
            "per_ray": {"node_visits": cst.node_visits / N_RAYS, "leaf_gates": cst.leaf_gates / N_RAYS, "triangle_tests": cst.triangle_tests / N_RAYS, "sphere_tests": cst.sphere_tests / N_RAYS,
                        "conservative_violations": int(cst.conservative_violations)},
        },
    }
    if render is not None:
        line["render"] = render
        line["render_c4"] = render_c4
    if incoherent is not None:
        line["incoherent"] = incoherent
    if not args.no_cpu and world == 1:
        line["cpu_baseline"] = cpu_baseline(sc, cam)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def d_hits_after(scene, d_rays, d_hits, stream):
    import torch

    scene.hit_device(d_rays.data_ptr(), N_RAYS, d_hits.data_ptr(), stream)
    torch.cuda.synchronize()
    return d_hits


def _camera_rays_numpy(cam, u, v):
    """render.rs:32-52 in numpy float64 (lens 0), used only to build the shifted multi-GPU input batches"""
    import math

    tan_fov = math.tan(0.5 * cam.fov)
    tx = (2.0 * u - 1.0) * tan_fov * cam.focal_dist * cam.aspect_ratio
    ty = (2.0 * v - 1.0) * tan_fov * cam.focal_dist
    tz = np.full_like(tx, -cam.focal_dist)
    n = np.sqrt(0.0 + ((tx * tx + ty * ty) + tz * tz))
    d = np.stack([tx / n, ty / n, tz / n], axis=1)
    m = np.asarray(cam.transformation.orientation)
    wd = np.stack([(m[r, 0] * d[:, 0] + m[r, 1] * d[:, 1]) + m[r, 2] * d[:, 2] for r in range(3)], axis=1)
    rays = np.empty((len(u), 8), dtype=np.float64)
    rays[:, 0:3] = np.asarray(cam.transformation.position)
    rays[:, 3:6] = wd
    rays[:, 6], rays[:, 7] = 1e-3, np.inf
    return rays


def bench_render(args, torch, dist, api, A, scene, sc, dev, rank, world, stream, barrier, rw=640, rh=360, spp_rank=16, name="C1: bunny Lambert + sky", max_steps=20):
    """C1: bunny Lambert + sky, 640x360, depth 8, 16 spp per rank (weak scaling: world*16 spp in total). Each rank
    renders its sample range into raw sums; one NCCL all-reduce sums the accumulation buffers; rank 0 divides.
    Also used for the C4 scene (demo scene, 1920x1080) at 8 of its 256 spp per rank."""
    depth = 8
    spp = spp_rank * world
    cam = api.Camera(rw / rh, sc.camera.fov, sc.camera.focal_dist, sc.camera.lens_radius, sc.camera.transformation)
    acc = torch.zeros((rh * rw * 4,), dtype=torch.float64, device=dev)  # rgb (3*npix) then foreground (npix)
    from rtp_b200 import dist as rdist

    sb, se = rdist.sample_range(spp, rank, world)
    p = api.render_params(rw, rh, spp, depth, seed=1, sample_begin=sb, sample_end=se, flags=A.RENDER_RAW_SUMS)
    st = scene.render_device(p, cam, acc.data_ptr(), acc.data_ptr() + rh * rw * 3 * 8, stream, stats=True)
    rays_rank = torch.tensor([float(st.rays)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(rays_rank)
    steps = max(3, min(args.steps, max_steps))
    for _ in range(3):  # warm-up includes the collective: NCCL sets up its channels for this message size on first use
        scene.render_device(p, cam, acc.data_ptr(), acc.data_ptr() + rh * rw * 3 * 8, stream)
        rdist.reduce_frame(acc)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        scene.render_device(p, cam, acc.data_ptr(), acc.data_ptr() + rh * rw * 3 * 8, stream)
        rdist.reduce_frame(acc)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    sec = float(ms.item()) * 1e-3 / steps
    paths = rw * rh * spp
    return {
        "workload": f"{name}, {rw}x{rh}, {spp_rank} spp per GPU ({spp} total), max depth {depth}",
        "samples_per_s": paths / sec, "mrays_per_s": float(rays_rank.item()) / sec / 1e6, "ms_per_frame": sec * 1e3, "steps": steps,
        "rays_per_path": float(rays_rank.item()) / paths, "collective": f"NCCL all-reduce(sum, f64) of the {rw}x{rh}x4 accumulation buffer" if world > 1 else "none (1 GPU)",
        "launches_per_frame": int(st.kernel_launches),
    }


def cpu_baseline(sc, cam):
    """oracle port (kind 'port': the Rust reference cannot be compiled here) on all host cores, bounded sample: 6 passes
    over the full C2 batch (~15 core-seconds)."""
    import oracle

    o = oracle.Scene(sc)
    rays = oracle.camera_rays(cam, W, H)
    cores = os.cpu_count() or 1
    o.hit_full(rays, threads=cores)
    reps = 6
    t0 = time.perf_counter()
    for _ in range(reps):
        o.hit_full(rays, threads=cores)
    dt = time.perf_counter() - t0
    out = {"value": reps * N_RAYS / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
           "sample": f"{reps} passes over the full C2 batch (2,073,600 rays each), oracle/rtp_oracle.c with {cores} pthreads, gcc -O3 -ffp-contract=off"}
    # the reference ships with 4 worker threads hard-coded (main.rs:27): the same port at 4 threads, 2 passes
    t0 = time.perf_counter()
    for _ in range(2):
        o.hit_full(rays, threads=4)
    out["as_shipped_4_threads"] = {"value": 2 * N_RAYS / (time.perf_counter() - t0) / 1e6, "unit": "Mrays/s", "cores": 4}
    return out


if __name__ == "__main__":
    main()
