"""Small workload for a sanitizer or a debug build (compute-sanitizer is closed on the round-1 GPU pool, so it was run plain): closest-hit batches and a tiny render in both visiting orders,
with the any-order stack forced small so that re-walking lanes and any-order lanes share warps.
    compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from rtp_b200 import api, scenes

api.init(0)
sc = scenes.bunny_lambert()
cam = api.Camera(16 / 9, sc.camera.fov, sc.camera.focal_dist, 0.0, sc.camera.transformation)
rays = np.concatenate([api.camera_rays(cam, 96, 54), scenes.incoherent_rays(6000)])
ref = None
for order, cap in (("inorder", "32"), ("any", "32"), ("any", "1")):
    os.environ["RTP_TRAVERSAL"], os.environ["RTP_ANY_CAP"] = order, cap
    g = api.Scene(sc)
    h = g.hit(rays)
    img, fg, st = g.render(48, 32, 2, seed=3)
    tile, _, _ = g.render(48, 32, 2, seed=3, tile=(16, 0, 32, 32))
    if ref is None:
        ref = (h.copy(), img.copy())
    same = h.tobytes() == ref[0].tobytes() and img.tobytes() == ref[1].tobytes() and np.array_equal(tile[:, 16:48], img[:, 16:48])
    print(order, cap, "same bits as the first run:", same, "rays", int(st.rays), flush=True)
    g.close()
