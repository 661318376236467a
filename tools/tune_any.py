"""Sweeps RTP_REFILL_MIN x RTP_PRIM_BATCH for the traversal kernel on C2 / C3 (bunny) in one process.
    python tools/tune_any.py "8,12,16,20,24" "4,6,8,12" """
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np
import torch

from rtp_b200 import api, scenes
from sweep_env import time_batch


def main():
    refills = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "8,12,16,20,24").split(",")]
    prims = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "4,6,8,12").split(",")]
    api.init(0)
    sc = scenes.bunny_lambert()
    cam = api.Camera(1920 / 1080, sc.camera.fov, sc.camera.focal_dist, 0.0, sc.camera.transformation)
    c2 = [torch.empty((1920 * 1080, 8), dtype=torch.float64, device="cuda") for _ in range(2)]
    api.camera_rays_device(cam, 1920, 1080, c2[0].data_ptr())
    torch.cuda.synchronize()
    c2[1].copy_(c2[0])
    c3 = [torch.from_numpy(scenes.incoherent_rays(1 << 22).view(np.float64).reshape(-1, 8)).cuda()]
    h2 = torch.empty((c2[0].shape[0], 2), dtype=torch.float64, device="cuda")
    h3 = torch.empty((c3[0].shape[0], 2), dtype=torch.float64, device="cuda")
    print(f"{'refill':>6} {'prim':>5} {'C2 Mrays/s':>11} {'C3 Mrays/s':>11}")
    for r in refills:
        for p in prims:
            os.environ["RTP_REFILL_MIN"], os.environ["RTP_PRIM_BATCH"] = str(r), str(p)
            scene = api.Scene(sc)
            m2 = time_batch(scene, c2, h2, 40)
            m3 = time_batch(scene, c3, h3, 8)
            print(f"{r:>6} {p:>5} {m2:>11.1f} {m3:>11.1f}", flush=True)
            scene.close()


if __name__ == "__main__":
    main()
