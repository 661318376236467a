"""more_balls_optimized (3,786 spheres under Bvh::new): in-order against any-order walk, primary batch and a frame."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rtp_b200 import _abi as A
from rtp_b200 import api, scenes

api.init(0)
st = torch.cuda.current_stream().cuda_stream
sc = scenes.more_balls_optimized()
W, H = 1920, 1920
cam = api.Camera(1.0, sc.camera.fov, sc.camera.focal_dist, 0.0, sc.camera.transformation)
rays = torch.empty((W * H, 8), dtype=torch.float64, device="cuda")
api.camera_rays_device(cam, W, H, rays.data_ptr(), st)
hits = torch.empty((W * H, 2), dtype=torch.float64, device="cuda")
ref = None
for mode in ("inorder", "any"):
    os.environ["RTP_TRAVERSAL"] = mode
    scene = api.Scene(sc)
    n = W * H
    for _ in range(2):
        scene.hit_device(rays.data_ptr(), n, hits.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        scene.hit_device(rays.data_ptr(), n, hits.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    c = scene.hit_device_counted(rays.data_ptr(), n, hits.data_ptr())
    same = "" if ref is None else f" same bits: {bool((ref.view(torch.int64) == hits.view(torch.int64)).all())}"
    ref = hits.clone() if ref is None else ref
    cam2 = api.Camera(1.0, sc.camera.fov, sc.camera.focal_dist, sc.camera.lens_radius, sc.camera.transformation)
    acc = torch.zeros((800 * 800 * 4,), dtype=torch.float64, device="cuda")
    p = api.render_params(800, 800, 8, 8, seed=1, flags=A.RENDER_RAW_SUMS)
    scene.render_device(p, cam2, acc.data_ptr(), acc.data_ptr() + 800 * 800 * 24, st)
    s = scene.render_device(p, cam2, acc.data_ptr(), acc.data_ptr() + 800 * 800 * 24, st, stats=True)
    print(f"{mode:8s} primary {8 * n / (e0.elapsed_time(e1) * 1e-3) / 1e6:8.1f} Mrays/s nodes {c.node_visits / n:5.2f} sphere tests {c.sphere_tests / n:5.2f} rewalks {c.order_rewalks}{same}; "
          f"800x800x8 frame {s.device_ms:.2f} ms, {s.rays / s.device_ms / 1e3:.0f} Mrays/s", flush=True)
    scene.close()
