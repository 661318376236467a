"""Launch latency of the traversal kernel against batch size (incoherent rays): python tools/small_batches.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rtp_b200 import api, scenes

api.init(0)
sc = scenes.bunny_lambert()
scene = api.Scene(sc)
st = torch.cuda.current_stream().cuda_stream
rays = torch.from_numpy(scenes.incoherent_rays(1 << 20).view(np.float64).reshape(-1, 8)).cuda()
hits = torch.empty((rays.shape[0], 2), dtype=torch.float64, device="cuda")
for n in (32, 1024, 4096, 16384, 65536, 262144, 1 << 20):
    for _ in range(3):
        scene.hit_device(rays.data_ptr(), n, hits.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        scene.hit_device(rays.data_ptr(), n, hits.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"n={n:8d}  {us:8.1f} us/launch  {n / us:8.1f} Mrays/s")
