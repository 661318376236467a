"""Scene inputs: the reference's meshes/texture (from the committed fixture) and the synthetic sky.

The GPU box has no /root/reference, so the reference assets travel as package data,
raytracing-potato_b200/data/assets.npz, written by tests/golden/make_fixtures.py from
/root/reference/assets/{bunny.obj,bunny_flat.obj,earthmap.tga}. `assets/sky_panorama.tga` is absent
from the reference checkout (.MISSING_LARGE_BLOBS), so the sky is generated here with integer
arithmetic only — identical bytes on every machine.
"""
from __future__ import annotations

import os

import numpy as np

from . import _abi as A
from .api import Mesh

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "assets.npz")
_cache = {}


def _fixture():
    if "npz" not in _cache:
        if not os.path.exists(FIXTURE):
            raise FileNotFoundError(f"{FIXTURE} missing; run tests/golden/make_fixtures.py where /root/reference exists")
        _cache["npz"] = np.load(FIXTURE)
    return _cache["npz"]


def _mesh(prefix: str) -> Mesh:
    z = _fixture()
    v = np.zeros(len(z[prefix + "_position"]), dtype=A.VERTEX_DTYPE)
    v["position"], v["normal"], v["uv"] = z[prefix + "_position"], z[prefix + "_normal"], z[prefix + "_uv"]
    return Mesh(v, z[prefix + "_indices"].astype(np.uint32), 0)


def bunny() -> Mesh:
    """obj::load("assets/bunny.obj") — 2,503 vertices, 14,904 indices (smooth normals)"""
    return _mesh("bunny")


def bunny_flat() -> Mesh:
    """obj::load("assets/bunny_flat.obj") — 14,902 vertices (flat normals)"""
    return _mesh("bunny_flat")


def earthmap() -> np.ndarray:
    """tga::load("assets/earthmap.tga") — [512, 1024, 4] RGBA8, row 0 = bottom (south)"""
    z = _fixture()
    rgb = z["earthmap_rgb"]
    out = np.full(rgb.shape[:2] + (4,), 255, dtype=np.uint8)
    out[..., :3] = rgb
    return out


def _hash2(ix: np.ndarray, iy: np.ndarray, seed: int) -> np.ndarray:
    """64-bit integer lattice hash → uint64 (wrapping arithmetic)"""
    with np.errstate(over="ignore"):
        h = ix.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + iy.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F) + np.uint64(seed) * np.uint64(0x165667B19E3779F9)
        h ^= h >> np.uint64(29)
        h *= np.uint64(0xBF58476D1CE4E5B9)
        h ^= h >> np.uint64(32)
    return h


def sky_panorama(seed: int = 1, width: int = 2048, height: int = 1024) -> np.ndarray:
    """Synthetic equirectangular sky, [height, width, 4] RGBA8, row 0 = bottom (nadir).

    Upper half: zenith-blue → horizon-white gradient with integer value-noise clouds and a sun disc at
    azimuth column 5*width/8, elevation row 3*height/4. Lower half: grey-brown ground gradient.
    Everything is integer arithmetic (uint64/int64), so the bytes do not depend on the platform libm.
    """
    j, i = np.meshgrid(np.arange(height, dtype=np.int64), np.arange(width, dtype=np.int64), indexing="ij")
    half = height // 2
    up = np.clip(j - half, 0, None)              # 0 at horizon … half at zenith
    r = 235 - (up * 150) // max(half, 1)
    g = 240 - (up * 110) // max(half, 1)
    b = 250 - (up * 30) // max(half, 1)
    # clouds: bilinear value noise on a 64-texel lattice, amplitude ±24, upper half only
    cell = 64
    x0, y0 = i // cell, j // cell
    fx, fy = i % cell, j % cell
    x1 = (x0 + 1) % (width // cell)              # wrap in azimuth
    n00 = (_hash2(x0, y0, seed) >> np.uint64(56)).astype(np.int64)
    n10 = (_hash2(x1, y0, seed) >> np.uint64(56)).astype(np.int64)
    n01 = (_hash2(x0, y0 + 1, seed) >> np.uint64(56)).astype(np.int64)
    n11 = (_hash2(x1, y0 + 1, seed) >> np.uint64(56)).astype(np.int64)
    top = n00 * (cell - fx) + n10 * fx
    bot = n01 * (cell - fx) + n11 * fx
    noise = (top * (cell - fy) + bot * fy) // (cell * cell)      # 0..255
    cloud = ((noise - 128) * 24) // 128
    sky = j >= half
    r = np.where(sky, r + cloud, 0)
    g = np.where(sky, g + cloud, 0)
    b = np.where(sky, b + cloud // 2, 0)
    # ground
    down = np.clip(half - 1 - j, 0, None)
    gr = 120 - (down * 60) // max(half, 1)
    r = np.where(sky, r, gr + 10)
    g = np.where(sky, g, gr)
    b = np.where(sky, b, gr - 15)
    # sun
    si, sj, rad = (5 * width) // 8, (3 * height) // 4, max(width // 64, 2)
    sun = (i - si) ** 2 + (j - sj) ** 2 < rad * rad
    r = np.where(sun, 255, r)
    g = np.where(sun, 250, g)
    b = np.where(sun, 225, b)
    out = np.empty((height, width, 4), dtype=np.uint8)
    out[..., 0] = np.clip(r, 0, 255)
    out[..., 1] = np.clip(g, 0, 255)
    out[..., 2] = np.clip(b, 0, 255)
    out[..., 3] = 255
    return out
