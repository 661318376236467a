"""Multi-GPU render-leg probe (run under torchrun): where does the frame time go when the sample ranges are summed over NCCL?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NCCL_DEBUG"] = "WARN"
import torch
import torch.distributed as dist
from rtp_b200 import api, scenes, _abi as A, dist as rdist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
api.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
sc = scenes.bunny_lambert(); scene = api.Scene(sc)
w, h, spp_rank = 640, 360, 16
spp = spp_rank * world
cam = api.Camera(w / h, sc.camera.fov, sc.camera.focal_dist, sc.camera.lens_radius, sc.camera.transformation)
acc = torch.zeros((w * h * 4,), dtype=torch.float64, device=dev)
sb, se = rdist.sample_range(spp, rank, world)
p = api.render_params(w, h, spp, 8, seed=1, sample_begin=sb, sample_end=se, flags=A.RENDER_RAW_SUMS)
st = torch.cuda.current_stream().cuda_stream

def timed(fn, reps=30):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

def render(): scene.render_device(p, cam, acc.data_ptr(), acc.data_ptr() + w * h * 24, st)
def render_allreduce(): render(); dist.all_reduce(acc)
def render_reduce(): render(); dist.reduce(acc, dst=0)
def allreduce_only(): dist.all_reduce(acc)
accf = acc.view(torch.float64)
res = {"render": timed(render), "render+all_reduce": timed(render_allreduce), "render+reduce": timed(render_reduce), "all_reduce only": timed(allreduce_only)}
if rank == 0:
    print(world, {k: round(v, 3) for k, v in res.items()}, flush=True)
dist.destroy_process_group()
