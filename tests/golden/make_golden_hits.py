"""Writes tests/golden/c2_hits_every97.npz: oracle closest hits (leaf id, t bit pattern) for every 97th ray
of the BASELINE C2 batch (1920x1080 pixel-centre camera rays vs the bunny scene). A regression pin for the
oracle itself; regenerate only when the oracle's defined semantics change (and say why in DESIGN.md).

    python tests/golden/make_golden_hits.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from rtp_b200 import api, scenes  # noqa: E402


def main():
    sc = scenes.bunny_lambert()
    o = oracle.Scene(sc)
    cam = api.Camera(1920 / 1080, sc.camera.fov, 1.0, 0.0, sc.camera.transformation)
    rays = oracle.camera_rays(cam, 1920, 1080)[::97]
    h = o.hit(rays)
    path = os.path.join(HERE, "c2_hits_every97.npz")
    np.savez_compressed(path, leaf=h["leaf"], t_bits=h["t"].view(np.uint64))
    print("wrote", path, len(rays), "rays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
