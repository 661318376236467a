"""Where does the time of a tiny traversal launch go? Host enqueue cost per launch (wall clock of the enqueue loop, no
synchronisation) against device time per launch (CUDA events), for a few batch sizes and RTP_MIN_LANES settings."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rtp_b200 import api, scenes

api.init(0)
sc = scenes.bunny_lambert()
st = torch.cuda.current_stream().cuda_stream
rays = torch.from_numpy(scenes.incoherent_rays(1 << 16).view(np.float64).reshape(-1, 8)).cuda()
hits = torch.empty((rays.shape[0], 2), dtype=torch.float64, device="cuda")
for min_lanes in ("1", "4", "32"):
    os.environ["RTP_MIN_LANES"] = min_lanes
    scene = api.Scene(sc)
    for n in (1, 32, 1024, 8192):
        for _ in range(5):
            scene.hit_device(rays.data_ptr(), n, hits.data_ptr(), st)
        torch.cuda.synchronize()
        reps = 200
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            scene.hit_device(rays.data_ptr(), n, hits.data_ptr(), st)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        print(f"min_lanes={min_lanes:>2s} n={n:6d}: host enqueue {1e6 * (t1 - t0) / reps:6.1f} us/launch, device {1e3 * e0.elapsed_time(e1) / reps:6.1f} us/launch", flush=True)
    scene.close()
