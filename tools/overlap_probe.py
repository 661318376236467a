"""Would two wavefronts in flight help a small frame? Renders C1 (640x360, 16 spp) as ONE 16-spp call, and as TWO concurrent 8-spp
calls on two scene handles / two streams / two host threads (sample ranges 0-8 and 8-16): python tools/overlap_probe.py"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from rtp_b200 import _abi as A
from rtp_b200 import api, scenes


def main():
    api.init(0)
    sc = scenes.bunny_lambert()
    w, h, spp = 640, 360, 16
    cam = api.Camera(w / h, sc.camera.fov, sc.camera.focal_dist, sc.camera.lens_radius, sc.camera.transformation)
    handles = [api.Scene(sc), api.Scene(sc)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    acc = [torch.zeros((w * h * 4,), dtype=torch.float64, device="cuda") for _ in range(2)]

    def render(k, begin, end, reps):
        p = api.render_params(w, h, spp, 8, seed=1, sample_begin=begin, sample_end=end, flags=A.RENDER_RAW_SUMS)
        for _ in range(reps):
            handles[k].render_device(p, cam, acc[k].data_ptr(), acc[k].data_ptr() + w * h * 24, streams[k].cuda_stream)

    for reps in (5, 40):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        render(0, 0, 16, reps)
        torch.cuda.synchronize()
        one = (time.perf_counter() - t0) / reps * 1e3
        t0 = time.perf_counter()
        th = [threading.Thread(target=render, args=(k, 8 * k, 8 * k + 8, reps)) for k in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        two = (time.perf_counter() - t0) / reps * 1e3
        t0 = time.perf_counter()
        render(0, 0, 8, reps)
        render(0, 8, 16, reps)
        torch.cuda.synchronize()
        seq = (time.perf_counter() - t0) / reps * 1e3
        if reps > 5:
            print(f"one 16-spp call {one:.3f} ms; two concurrent 8-spp calls {two:.3f} ms; two 8-spp calls one after the other {seq:.3f} ms")


if __name__ == "__main__":
    main()
