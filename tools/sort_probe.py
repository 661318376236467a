"""How much would ray reordering buy on the incoherent batch (C3)? Sort 2^22 C3 rays on the host by several candidate
keys and time the unchanged traversal kernel on each order: python tools/sort_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from rtp_b200 import api, scenes


def part1by2(x):
    x = x.astype(np.uint64) & np.uint64(0x3FF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x30000FF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x300F00F)
    x = (x | (x << np.uint64(4))) & np.uint64(0x30C30C3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x9249249)
    return x


def morton3(q):
    return part1by2(q[:, 0]) | (part1by2(q[:, 1]) << np.uint64(1)) | (part1by2(q[:, 2]) << np.uint64(2))


def part1by1(x):
    x = x.astype(np.uint64) & np.uint64(0xFFFF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x00FF00FF)
    x = (x | (x << np.uint64(4))) & np.uint64(0x0F0F0F0F)
    x = (x | (x << np.uint64(2))) & np.uint64(0x33333333)
    x = (x | (x << np.uint64(1))) & np.uint64(0x55555555)
    return x


def quant(x, bits):
    lo, hi = x.min(axis=0), x.max(axis=0)
    return np.minimum(((x - lo) / np.maximum(hi - lo, 1e-300) * (1 << bits)).astype(np.int64), (1 << bits) - 1)


def octa(d, bits):
    n = d / np.abs(d).sum(axis=1, keepdims=True)
    u, v = n[:, 0].copy(), n[:, 1].copy()
    neg = n[:, 2] < 0
    uu = (1 - np.abs(v)) * np.sign(u + 1e-300)
    vv = (1 - np.abs(u)) * np.sign(v + 1e-300)
    u[neg], v[neg] = uu[neg], vv[neg]
    q = np.stack([u, v], axis=1) * 0.5 + 0.5
    return np.minimum((q * (1 << bits)).astype(np.int64), (1 << bits) - 1)


def main():
    n = 1 << 22
    api.init(0)
    st = torch.cuda.current_stream().cuda_stream
    sc = scenes.bunny_lambert()
    scene = api.Scene(sc)
    rays = scenes.incoherent_rays(n)
    o, d = rays["origin"], rays["direction"]
    flat = rays.view(np.float64).reshape(-1, 8)

    def run(order, name):
        r = torch.from_numpy(flat if order is None else flat[order]).cuda()
        hits = torch.empty((n, 2), dtype=torch.float64, device="cuda")
        scene.hit_device(r.data_ptr(), n, hits.data_ptr(), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            scene.hit_device(r.data_ptr(), n, hits.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:46s} {n * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e6:8.1f} Mrays/s", flush=True)
        return hits

    h = run(None, "unsorted")
    leaf = h.cpu().numpy().view(np.uint32).reshape(n, 4)[:, 0]
    run(np.argsort(leaf, kind="stable"), "by hit leaf id (ceiling)")
    octant = (d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << np.uint64(1)) | ((d[:, 2] < 0).astype(np.uint64) << np.uint64(2))
    for pb, db in ((4, 5), (5, 4), (3, 6), (6, 3)):
        mo, od = morton3(quant(o, pb)), octa(d, db)
        md = part1by1(od[:, 0]) | (part1by1(od[:, 1]) << np.uint64(1))
        run(np.argsort((mo << np.uint64(2 * db)) | md, kind="stable"), f"origin morton {pb}b/axis major, dir oct {db}b minor")
        run(np.argsort((md << np.uint64(3 * pb)) | mo, kind="stable"), f"dir oct {db}b major, origin morton {pb}b/axis minor")
    # a point along the ray near the scene content: origin + direction * (distance from the origin to the batch's mean target)
    t_mid = np.linalg.norm(o - (o + d * 3.0).mean(axis=0), axis=1)
    p = o + d * t_mid[:, None]
    for pb, db in ((5, 4), (6, 3), (7, 2)):
        mp, od = morton3(quant(p, pb)), octa(d, db)
        md = part1by1(od[:, 0]) | (part1by1(od[:, 1]) << np.uint64(1))
        run(np.argsort((mp << np.uint64(2 * db)) | md, kind="stable"), f"mid-point morton {pb}b major, dir oct {db}b minor")
        run(np.argsort((md << np.uint64(3 * pb)) | mp, kind="stable"), f"dir oct {db}b major, mid-point morton {pb}b minor")
    run(np.argsort((octant << np.uint64(30)) | morton3(quant(p, 10)), kind="stable"), "octant major, mid-point morton 10b")


if __name__ == "__main__":
    main()
