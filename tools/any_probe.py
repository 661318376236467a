"""In-order vs any-order walk (RTP_TRAVERSAL) on the bunny scene and on bunny fields: python tools/any_probe.py [nx nz]..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from rtp_b200 import _abi as A
from rtp_b200 import api, scenes


def timed(scene, rays, hits, st, reps):
    n = rays.shape[0]
    scene.hit_device(rays.data_ptr(), n, hits.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        scene.hit_device(rays.data_ptr(), n, hits.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e6


def main():
    api.init(0)
    st = torch.cuda.current_stream().cuda_stream
    fields = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)] or [(16, 8)]
    cases = [("bunny", scenes.bunny_lambert(), 1920, 1080)] + [(f"field {nx}x{nz}", scenes.bunny_field(nx, nz), 3840, 2160) for nx, nz in fields]
    for name, sc, W, H in cases:
        cam = api.Camera(W / H, sc.camera.fov, sc.camera.focal_dist, 0.0, sc.camera.transformation)
        n = W * H
        rays = torch.empty((n, 8), dtype=torch.float64, device="cuda")
        api.camera_rays_device(cam, W, H, rays.data_ptr(), st)
        if name == "bunny":
            rays2 = torch.from_numpy(scenes.incoherent_rays(1 << 22).view(np.float64).reshape(-1, 8)).cuda()
        else:
            perm = torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
            rays2 = rays.clone()
            rays2[:, 3:6] = rays[perm, 3:6]
        ref = {}
        for mode in ("inorder", "any", "any+free"):
            os.environ["RTP_TRAVERSAL"] = mode.split("+")[0]
            os.environ["RTP_FREE_TREE"] = "1" if mode.endswith("free") else "0"
            scene = api.Scene(sc)
            for tag, r in (("primary", rays), ("incoherent", rays2)):
                hits = torch.empty((r.shape[0], 2), dtype=torch.float64, device="cuda")
                mr = timed(scene, r, hits, st, 6)
                c = scene.hit_device_counted(r.data_ptr(), r.shape[0], hits.data_ptr())
                m = r.shape[0]
                same = ""
                if tag in ref:
                    same = f" same bits as in-order: {bool((ref[tag].view(torch.int64) == hits.view(torch.int64)).all())}"
                else:
                    ref[tag] = hits.clone()
                print(f"{name:12s} {mode:9s} {tag:10s} {mr:8.1f} Mrays/s  nodes {c.node_visits / m:6.2f} gates {c.leaf_gates / m:5.2f} tri {c.triangle_tests / m:5.2f} "
                      f"sph {c.sphere_tests / m:5.2f} viol {c.conservative_violations} rewalks {c.order_rewalks}{same}", flush=True)
            if name != "bunny":
                W2, H2, spp = 1920, 1080, 2
                cam2 = api.Camera(W2 / H2, sc.camera.fov, sc.camera.focal_dist, sc.camera.lens_radius, sc.camera.transformation)
                acc = torch.zeros((W2 * H2 * 4,), dtype=torch.float64, device="cuda")
                p = api.render_params(W2, H2, spp, 8, seed=1, flags=A.RENDER_RAW_SUMS)
                s = scene.render_device(p, cam2, acc.data_ptr(), acc.data_ptr() + W2 * H2 * 24, st, stats=True)
                s = scene.render_device(p, cam2, acc.data_ptr(), acc.data_ptr() + W2 * H2 * 24, st, stats=True)
                print(f"{name:12s} {mode:9s} render {W2}x{H2}x{spp}: {s.device_ms:.1f} ms, {s.rays / s.device_ms / 1e3:.1f} Mrays/s, sum {float(acc.sum()):.10e}", flush=True)
            scene.close()
        del os.environ["RTP_TRAVERSAL"]


if __name__ == "__main__":
    main()
