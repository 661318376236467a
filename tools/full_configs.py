"""BASELINE configs C4 and C5 at their full sizes, on 1..8 GPUs (run under torchrun for N > 1):
    C4: demo scene, 1920x1080, 256 spp, depth 8            C5: 64x32 bunny field (10,174,464 triangles), 3840x2160, 1024 spp, depth 8
Frames are split by sample range (rtp_b200.dist), raw sums are all-reduced over NCCL, rank 0 divides, converts to sRGB8 and
(optionally) writes a TGA. Prints one JSON line with samples/s and Mrays/s over the whole job."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NCCL_DEBUG"] = os.environ.get("RTP_NCCL_DEBUG", "WARN")
import numpy as np
import torch
import torch.distributed as dist
from rtp_b200 import api, scenes, _abi as A, dist as rdist

ap = argparse.ArgumentParser()
ap.add_argument("--config", choices=["c4", "c5"], default="c4")
ap.add_argument("--spp", type=int, default=0, help="total samples per pixel (default: the config's 256 / 1024)")
ap.add_argument("--out", default="")
args = ap.parse_args()
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
api.init(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
w, h, spp = (1920, 1080, 256) if args.config == "c4" else (3840, 2160, 1024)
spp = args.spp or spp
t0 = time.time()
sc = scenes.demo() if args.config == "c4" else scenes.bunny_field(64, 32)
scene = api.Scene(sc)
t_build = time.time() - t0
cam = api.Camera(w / h, sc.camera.fov, sc.camera.focal_dist, sc.camera.lens_radius, sc.camera.transformation)
acc = torch.zeros((w * h * 4,), dtype=torch.float64, device=dev)
sb, se = rdist.sample_range(spp, rank, world)
p = api.render_params(w, h, spp, 8, seed=1, sample_begin=sb, sample_end=se, flags=A.RENDER_RAW_SUMS)
stream = torch.cuda.current_stream().cuda_stream
# warm-up: up to 8 spp (enough paths for a full-size launch, so the per-launch buffers are allocated before the timed region) + the collective
pw = api.render_params(w, h, spp, 8, seed=1, sample_begin=sb, sample_end=sb + min(se - sb, 8), flags=A.RENDER_RAW_SUMS)
scene.render_device(pw, cam, acc.data_ptr(), acc.data_ptr() + w * h * 24, stream)
rdist.reduce_frame(acc)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
st = scene.render_device(p, cam, acc.data_ptr(), acc.data_ptr() + w * h * 24, stream, stats=True)
rdist.reduce_frame(acc)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
rays = torch.tensor([float(st.rays)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(rays)
if rank == 0:
    sec = float(ms.item()) * 1e-3
    frame = rdist.finish_frame(acc[: w * h * 3].cpu().numpy().reshape(h, w, 3), spp)
    line = {"config": args.config, "n_gpus": world, "width": w, "height": h, "spp": spp, "paths": w * h * spp, "seconds": sec,
            "samples_per_s": w * h * spp / sec, "mrays_per_s": float(rays.item()) / sec / 1e6, "rays_per_path": float(rays.item()) / (w * h * spp),
            "scene_build_s": t_build, "leaves": int(scene.info().n_leaves), "device_GiB": scene.info().device_bytes / 2**30,
            "mean_rgb": [float(x) for x in frame.reshape(-1, 3).mean(axis=0)]}
    print(json.dumps(line), flush=True)
    if args.out:
        api.tga.save(api.to_srgb_u8(frame), args.out)
if world > 1:
    dist.destroy_process_group()
