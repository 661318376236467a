"""Short single-workload runner for ncu captures: python tools/prof_run.py c2|c3|c1|c4|c5s [reps]
c2 = 1920x1080 primary batch, c3 = 2^22 incoherent rays, c1 = 640x360x16spp bunny render, c4 = demo scene 960x540x8spp,
c5s = 3840x2160 primary batch against a 32x16 bunny field (2,543,617 leaves, scene in HBM)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from rtp_b200 import _abi as A
from rtp_b200 import api, scenes


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "c2"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    api.init(0)
    st = torch.cuda.current_stream().cuda_stream
    sc = scenes.demo() if what == "c4" else (scenes.bunny_field(32, 16) if what == "c5s" else scenes.bunny_lambert())
    scene = api.Scene(sc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if what in ("c2", "c3", "c5s"):
        if what in ("c2", "c5s"):
            W, H = (1920, 1080) if what == "c2" else (3840, 2160)
            cam = api.Camera(W / H, sc.camera.fov, sc.camera.focal_dist, 0.0, sc.camera.transformation)
            rays = torch.empty((W * H, 8), dtype=torch.float64, device="cuda")
            api.camera_rays_device(cam, W, H, rays.data_ptr(), st)
        else:
            rays = torch.from_numpy(scenes.incoherent_rays(1 << 22).view(np.float64).reshape(-1, 8)).cuda()
        hits = torch.empty((rays.shape[0], 2), dtype=torch.float64, device="cuda")
        scene.hit_device(rays.data_ptr(), rays.shape[0], hits.data_ptr(), st)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            scene.hit_device(rays.data_ptr(), rays.shape[0], hits.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        print(f"{what}: {rays.shape[0] * reps / (e0.elapsed_time(e1) * 1e-3) / 1e6:.1f} Mrays/s")
        cst = scene.hit_device_counted(rays.data_ptr(), rays.shape[0], hits.data_ptr())
        n = rays.shape[0]
        print(f"per ray: nodes {cst.node_visits / n:.2f} gates {cst.leaf_gates / n:.2f} tri {cst.triangle_tests / n:.2f} sph {cst.sphere_tests / n:.2f} violations {cst.conservative_violations}")
    else:
        w, h, spp = (640, 360, 16) if what == "c1" else (960, 540, 8)
        cam = api.Camera(w / h, sc.camera.fov, sc.camera.focal_dist, sc.camera.lens_radius, sc.camera.transformation)
        acc = torch.zeros((w * h * 4,), dtype=torch.float64, device="cuda")
        p = api.render_params(w, h, spp, 8, seed=1, flags=A.RENDER_RAW_SUMS)
        s = scene.render_device(p, cam, acc.data_ptr(), acc.data_ptr() + w * h * 24, st, stats=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            scene.render_device(p, cam, acc.data_ptr(), acc.data_ptr() + w * h * 24, st)
        e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3 / reps
        print(f"{what}: {w * h * spp / sec / 1e6:.2f} Msamples/s, {s.rays / sec / 1e6:.1f} Mrays/s, {s.rays / (w * h * spp):.3f} rays/path, {sec * 1e3:.2f} ms/frame")


if __name__ == "__main__":
    main()
