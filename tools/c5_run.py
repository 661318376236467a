"""BASELINE config C5 on one GPU: the 2,048-copy bunny field (10,174,464 triangles + ground sphere), 3840x2160.
python tools/c5_run.py [nx nz] — scene build time, primary-ray batch Mrays/s, work counters, a full-size parity property
(the f32-culled 4-wide walk and the exact f64 pre-order walk return identical bits), and a short render."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from rtp_b200 import _abi as A
from rtp_b200 import api, scenes


def main():
    nx, nz = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 32)
    W, H = 3840, 2160
    api.init(0)
    t0 = time.time()
    sc = scenes.bunny_field(nx, nz)
    t1 = time.time()
    scene = api.Scene(sc)
    t2 = time.time()
    info = scene.info()
    print(f"field {nx}x{nz}: {info.n_leaves} leaves, {info.n_nodes} reference nodes, depth {info.depth}, {info.device_bytes / 2**30:.2f} GiB on device; "
          f"host scene {t1 - t0:.1f} s, flatten+BVH+upload {t2 - t1:.1f} s")
    st = torch.cuda.current_stream().cuda_stream
    cam = api.Camera(W / H, sc.camera.fov, sc.camera.focal_dist, 0.0, sc.camera.transformation)
    rays = torch.empty((W * H, 8), dtype=torch.float64, device="cuda")
    api.camera_rays_device(cam, W, H, rays.data_ptr(), st)
    hits = torch.empty((W * H, 2), dtype=torch.float64, device="cuda")
    n = W * H
    for _ in range(2):
        scene.hit_device(rays.data_ptr(), n, hits.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        scene.hit_device(rays.data_ptr(), n, hits.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    print(f"primary 4K batch: {5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e6:.1f} Mrays/s")
    c = scene.hit_device_counted(rays.data_ptr(), n, hits.data_ptr())
    print(f"per ray: nodes {c.node_visits / n:.2f} gates {c.leaf_gates / n:.2f} tri {c.triangle_tests / n:.2f} sph {c.sphere_tests / n:.2f} violations {c.conservative_violations}")
    fast = hits.clone()
    # incoherent secondary-like rays: shuffle directions between pixels
    perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    rays2 = rays.clone()
    rays2[:, 0:3] = rays[:, 0:3] + 0.0
    rays2[:, 3:6] = rays[perm, 3:6]
    e0.record()
    scene.hit_device(rays2.data_ptr(), n, hits.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    print(f"direction-shuffled batch: {n / (e0.elapsed_time(e1) * 1e-3) / 1e6:.1f} Mrays/s")
    fast2 = hits.clone()
    # exact f64 walk over the same scene (second device scene with the f32 culling switched off)
    os.environ["RTP_F32_CULLING"] = "0"
    exact = api.Scene(sc)
    del os.environ["RTP_F32_CULLING"]
    sub = slice(0, n, 7)
    r_sub = rays[sub].contiguous()
    h_sub = torch.empty((r_sub.shape[0], 2), dtype=torch.float64, device="cuda")
    exact.hit_device(r_sub.data_ptr(), r_sub.shape[0], h_sub.data_ptr(), st)
    torch.cuda.synchronize()
    same = bool((h_sub.view(torch.int64) == fast[sub].view(torch.int64)).all())
    r_sub2 = rays2[sub].contiguous()
    exact.hit_device(r_sub2.data_ptr(), r_sub2.shape[0], h_sub.data_ptr(), st)
    torch.cuda.synchronize()
    same2 = bool((h_sub.view(torch.int64) == fast2[sub].view(torch.int64)).all())
    print(f"f32-culled walk == exact f64 walk on every 7th ray: primary {same}, shuffled {same2}")
    exact.close()
    # short render: 4K, 2 spp, depth 8
    acc = torch.zeros((W * H * 4,), dtype=torch.float64, device="cuda")
    cam2 = api.Camera(W / H, sc.camera.fov, sc.camera.focal_dist, sc.camera.lens_radius, sc.camera.transformation)
    p = api.render_params(W, H, 2, 8, seed=1, flags=A.RENDER_RAW_SUMS)
    s = scene.render_device(p, cam2, acc.data_ptr(), acc.data_ptr() + W * H * 24, st, stats=True)
    print(f"render 4K x 2 spp: {s.device_ms:.1f} ms, {s.paths / s.device_ms / 1e3:.1f} Msamples/s, {s.rays / s.device_ms / 1e3:.1f} Mrays/s, {s.rays / s.paths:.2f} rays/path")
    assert same and same2 and c.conservative_violations == 0


if __name__ == "__main__":
    main()
