"""Sweeps the traversal kernel's runtime tunables on the C2 (coherent) and C3 (incoherent) batches.
Run on a GPU box: python tools/tune_trace.py [--c3-log2 22]. Prints Mrays/s per configuration and checks that
every configuration returns bit-identical hits."""
import argparse
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from rtp_b200 import api, scenes


def time_batch(scene, d_rays, d_hits, reps=20):
    n = d_rays.shape[0]
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        scene.hit_device(d_rays.data_ptr(), n, d_hits.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        scene.hit_device(d_rays.data_ptr(), n, d_hits.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-log2", type=int, default=22)
    ap.add_argument("--grid", default="refill=4,8,12,16;prim=4,8,12,16;fast=1")
    ap.add_argument("--scene", default="bunny_lambert")
    ap.add_argument("--reps", default="2")
    args = ap.parse_args()
    grid = dict(kv.split("=") for kv in args.grid.split(";"))
    api.init(0)
    sc = getattr(scenes, args.scene)()
    cam = api.Camera(1920 / 1080, sc.camera.fov, sc.camera.focal_dist, 0.0, sc.camera.transformation)
    c2 = torch.empty((1920 * 1080, 8), dtype=torch.float64, device="cuda")
    api.camera_rays_device(cam, 1920, 1080, c2.data_ptr())
    c3 = torch.from_numpy(scenes.incoherent_rays(1 << args.c3_log2).view(np.float64).reshape(-1, 8)).cuda()
    h2 = torch.empty((c2.shape[0], 2), dtype=torch.float64, device="cuda")
    h3 = torch.empty((c3.shape[0], 2), dtype=torch.float64, device="cuda")
    ref2 = ref3 = None
    print(f"{'kernel':>8} {'refill':>6} {'prim':>5} {'reps':>4} {'C2 Mrays/s':>11} {'C3 Mrays/s':>11}")
    combos = [("persist", int(r), int(p), int(f)) for r, p, f in
              itertools.product(grid["refill"].split(","), grid["prim"].split(","), args.reps.split(","))]
    for kern, r, p, f in combos:
        os.environ["RTP_TRACE_KERNEL"] = kern
        f = 1
        os.environ["RTP_REFILL_MIN"], os.environ["RTP_PRIM_BATCH"], os.environ["RTP_FAST_SLAB"] = str(r or 8), str(p or 8), str(f)
        scene = api.Scene(sc)
        m2 = time_batch(scene, c2, h2)
        m3 = time_batch(scene, c3, h3, reps=5)
        a2, a3 = h2.clone(), h3.clone()
        if ref2 is None:
            ref2, ref3 = a2, a3
        ok = bool((a2.view(torch.int64) == ref2.view(torch.int64)).all() and (a3.view(torch.int64) == ref3.view(torch.int64)).all())
        print(f"{kern:>8} {r:>6} {p:>5} {2:>4} {m2:>11.1f} {m3:>11.1f} {'' if ok else 'MISMATCH'}", flush=True)
        scene.close()


if __name__ == "__main__":
    main()
