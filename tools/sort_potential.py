"""What would ray reordering buy on the incoherent batch? Sorts the C3 rays on the HOST by a few candidate keys and times the device
trace of each order (the sort itself is not timed here): python tools/sort_potential.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from rtp_b200 import api, scenes


def spread10(x):
    x = x.astype(np.uint64) & np.uint64(0x3FF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x30000FF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x300F00F)
    x = (x | (x << np.uint64(4))) & np.uint64(0x30C30C3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x9249249)
    return x


def morton(p, bits):
    lo, hi = p.min(axis=0), p.max(axis=0)
    q = np.clip((p - lo) / np.maximum(hi - lo, 1e-30) * (2 ** bits - 1), 0, 2 ** bits - 1).astype(np.uint64)
    return spread10(q[:, 0]) | (spread10(q[:, 1]) << np.uint64(1)) | (spread10(q[:, 2]) << np.uint64(2))


def main():
    api.init(0)
    sc = scenes.bunny_lambert()
    scene = api.Scene(sc)
    n = 1 << 22
    rays = scenes.incoherent_rays(n)
    o, d = rays["origin"], rays["direction"]
    octant = ((d[:, 0] < 0).astype(np.uint64) | ((d[:, 1] < 0).astype(np.uint64) << np.uint64(1)) | ((d[:, 2] < 0).astype(np.uint64) << np.uint64(2)))
    keys = {
        "as given": None,
        "origin morton 10b": morton(o, 10),
        "origin morton 5b, then direction morton 5b": (morton(o, 5) << np.uint64(15)) | morton(d, 5),
        "origin morton 4b, then direction morton 6b": (morton(o, 4) << np.uint64(18)) | morton(d, 6),
        "octant, then origin morton 8b": (octant << np.uint64(24)) | morton(o, 8),
        "target point (o + 3 d) morton 10b": morton(o + 3.0 * d, 10),
        "origin morton 6b, then target morton 4b": (morton(o, 6) << np.uint64(12)) | morton(o + 3.0 * d, 4),
    }
    st = torch.cuda.current_stream().cuda_stream
    hits = torch.empty((n, 2), dtype=torch.float64, device="cuda")
    for name, key in keys.items():
        r = rays if key is None else rays[np.argsort(key, kind="stable")]
        dr = torch.from_numpy(r.view(np.float64).reshape(-1, 8)).cuda()
        for _ in range(3):
            scene.hit_device(dr.data_ptr(), n, hits.data_ptr(), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            scene.hit_device(dr.data_ptr(), n, hits.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        s = scene.hit_device_counted(dr.data_ptr(), n, hits.data_ptr())
        print(f"{name:48s} {n * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e6:9.1f} Mrays/s   nodes/ray {s.node_visits / n:.2f}", flush=True)


if __name__ == "__main__":
    main()
