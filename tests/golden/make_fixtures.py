"""Regenerates raytracing-potato_b200/data/assets.npz (the scene inputs the package ships) from the reference checkout
(run in the build container only).

    python tests/golden/make_fixtures.py

Reads /root/reference/assets/{bunny.obj,bunny_flat.obj,earthmap.tga} through the ORACLE's loaders
(oracle/rtp_oracle.c, restating mesh.rs:112-183 and image.rs:73-114) and stores the parsed arrays, so
that tests, smoke() and bench.py never touch /root/reference at run time (it does not exist on the
GPU box). The fixture is data derived from the reference's assets, not reference source.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402

REF = "/root/reference/assets"


def main():
    out = {}
    for name in ("bunny", "bunny_flat"):
        m = oracle.obj_load(os.path.join(REF, name + ".obj"))
        out[name + "_position"] = m.vertices["position"].copy()
        out[name + "_normal"] = m.vertices["normal"].copy()
        out[name + "_uv"] = m.vertices["uv"].copy()
        out[name + "_indices"] = m.indices.astype(np.uint16 if m.indices.max() < 65536 else np.uint32)
        print(name, len(m.vertices), "vertices", len(m.indices), "indices")
    img = oracle.tga_load(os.path.join(REF, "earthmap.tga"))
    assert (img[..., 3] == 255).all()
    out["earthmap_rgb"] = img[..., :3].copy()
    print("earthmap", img.shape)
    path = os.path.join(os.path.dirname(os.path.dirname(HERE)), "raytracing-potato_b200", "data", "assets.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
