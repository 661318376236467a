"""End-to-end rtp_trace_closest throughput (pinned host rays in, hits out) against the pipeline chunk size: python tools/e2e_sweep.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time
sys.path.insert(0, %r)
import numpy as np
from rtp_b200 import api, scenes, _abi as A
api.init(0)
sc = scenes.bunny_lambert(); scene = api.Scene(sc)
W, H = 1920, 1080
cam = api.Camera(W / H, sc.camera.fov, sc.camera.focal_dist, 0.0, sc.camera.transformation)
r = api.PinnedBuffer((W * H,), A.RAY_DTYPE); h = api.PinnedBuffer((W * H,), A.HIT_DTYPE)
r.array[:] = api.camera_rays(cam, W, H)
for _ in range(3): scene.hit(r.array, out=h.array)
t0 = time.perf_counter()
for _ in range(30): scene.hit(r.array, out=h.array)
dt = (time.perf_counter() - t0) / 30
print(f"chunk 2^{os.environ.get('RTP_CHUNK_LOG2', '18')}: {W * H / dt / 1e6:.1f} Mrays/s, {W * H * 64 / dt / 1e9:.1f} GB/s H2D, {dt * 1e3:.3f} ms/step")
''' % ROOT
for l in (16, 17, 18, 19, 20, 21):
    env = dict(os.environ, RTP_CHUNK_LOG2=str(l))
    print(subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
