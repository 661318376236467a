import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rtp_b200 import api, scenes
api.init(0)
sc = scenes.bunny_lambert()
scene = api.Scene(sc)
st = torch.cuda.current_stream().cuda_stream
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rays = torch.from_numpy(scenes.incoherent_rays(n).view(np.float64).reshape(-1, 8)).cuda()
hits = torch.empty((rays.shape[0], 2), dtype=torch.float64, device="cuda")
for _ in range(4):
    scene.hit_device(rays.data_ptr(), n, hits.data_ptr(), st)
torch.cuda.synchronize()
print("ok")
