import sys, os
sys.path.insert(0, os.getcwd())
import torch
from rtp_b200 import api, scenes, _abi as A
api.init(0)
sc = scenes.bunny_lambert(); scene = api.Scene(sc)
w,h,spp=640,360,16
cam = api.Camera(w / h, sc.camera.fov, sc.camera.focal_dist, sc.camera.lens_radius, sc.camera.transformation)
acc = torch.zeros((w*h*4,), dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for depth in (1,2,3,4,5,6,8):
    p = api.render_params(w,h,spp,depth,seed=1,flags=A.RENDER_RAW_SUMS)
    s = scene.render_device(p, cam, acc.data_ptr(), acc.data_ptr()+w*h*24, st, stats=True)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): scene.render_device(p, cam, acc.data_ptr(), acc.data_ptr()+w*h*24, st)
    e1.record(); torch.cuda.synchronize()
    print(depth, s.rays, f"{e0.elapsed_time(e1)/10:.3f} ms")
