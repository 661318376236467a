"""Sweep of the traversal kernel's refill / leaf-batch thresholds on a bunny field (scene in HBM, long and uneven walks):
python tools/tune_field.py [nx nz]"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rtp_b200 import _abi as A
from rtp_b200 import api, scenes

nx, nz = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 16)
api.init(0)
st = torch.cuda.current_stream().cuda_stream
sc = scenes.bunny_field(nx, nz)
W, H = 3840, 2160
cam = api.Camera(W / H, sc.camera.fov, sc.camera.focal_dist, 0.0, sc.camera.transformation)
n = W * H
rays = torch.empty((n, 8), dtype=torch.float64, device="cuda")
api.camera_rays_device(cam, W, H, rays.data_ptr(), st)
hits = torch.empty((n, 2), dtype=torch.float64, device="cuda")
W2, H2, spp = 1920, 1080, 2
cam2 = api.Camera(W2 / H2, sc.camera.fov, sc.camera.focal_dist, sc.camera.lens_radius, sc.camera.transformation)
acc = torch.zeros((W2 * H2 * 4,), dtype=torch.float64, device="cuda")
p = api.render_params(W2, H2, spp, 8, seed=1, flags=A.RENDER_RAW_SUMS)
ref = None
for refill, prim in itertools.product((4, 8, 12, 16, 24), (4, 8, 12)):
    os.environ["RTP_REFILL_MIN"], os.environ["RTP_PRIM_BATCH"] = str(refill), str(prim)
    scene = api.Scene(sc)
    for _ in range(2):
        scene.hit_device(rays.data_ptr(), n, hits.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        scene.hit_device(rays.data_ptr(), n, hits.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    mr = 4 * n / (e0.elapsed_time(e1) * 1e-3) / 1e6
    if ref is None:
        ref = hits.clone()
    same = bool((ref.view(torch.int64) == hits.view(torch.int64)).all())
    scene.render_device(p, cam2, acc.data_ptr(), acc.data_ptr() + W2 * H2 * 24, st)
    s = scene.render_device(p, cam2, acc.data_ptr(), acc.data_ptr() + W2 * H2 * 24, st, stats=True)
    print(f"refill {refill:2d} prim {prim:2d}: primary {mr:7.1f} Mrays/s, frame {s.device_ms:6.2f} ms{'' if same else '  MISMATCH'}", flush=True)
    scene.close()
