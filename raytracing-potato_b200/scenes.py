"""Scene definitions: the reference's example scenes (example_scenes.rs) and the BASELINE.json configs.

Each function returns an `ExampleScene` built through the same constructors the reference uses.
`more_balls` draws its geometry from `StdRng::from_seed([249; 32])` (example_scenes.rs:98), i.e. rand 0.8's ChaCha12 stream;
`StdRngStream` restates that stream (the published ChaCha algorithm and rand_core's block-buffer reading order; the crates
themselves are not vendored with the reference, so the sequence is pinned by the ChaCha known answers only).
"""
from __future__ import annotations

import math

import numpy as np

from . import _abi as A
from . import assets
from .api import (Absorb, Camera, Emit, ExampleScene, Hittable, Material, Mesh, Scatter, SceneData, Texture, Transformation, rgb)

FRAC_PI_2 = math.pi / 2.0
FRAC_PI_4 = math.pi / 4.0


def three_balls() -> ExampleScene:
    """example_scenes.rs:22-60 — List root, Solid textures, all three scatter models, thin lens"""
    camera = Camera(1.0, FRAC_PI_2, 3.46, 0.1, Transformation.lookat([-2.0, 2.0, 1.0], [0.0, 0.0, -1.0], [0.0, 1.0, 0.0]))
    textures = [Texture.Solid(rgb(0.8, 0.8, 0.0)), Texture.Solid(rgb(0.1, 0.2, 0.5))]
    materials = [
        Material.new(Scatter.Lambert, Absorb.AlbedoMap(0), Emit.NONE),
        Material.new(Scatter.Lambert, Absorb.AlbedoMap(1), Emit.NONE),
        Material.new(Scatter.Dielectric(1.5), Absorb.WhiteBody, Emit.NONE),
        Material.new(Scatter.Metal(0.0), Absorb.Albedo(rgb(0.8, 0.6, 0.2)), Emit.NONE),
    ]
    root = Hittable.concat([
        Hittable.Sphere([0.0, -100.5, -1.0], 100.0, 0),
        Hittable.Sphere([0.0, 0.0, -1.0], 0.5, 1),
        Hittable.Sphere([-1.0, 0.0, -1.0], 0.5, 2),
        Hittable.Sphere([1.0, 0.0, -1.0], 0.5, 3),
    ])
    return ExampleScene(camera, SceneData(materials, textures, []), "list", root, Emit.SkyGradient)


class StdRngStream:
    """`rand::rngs::StdRng` of rand 0.8 (= rand_chacha's ChaCha12Rng) as far as the reference uses it: `from_seed(seed)` and
    `gen::<f64>()`. State: the ChaCha constants, the 32 seed bytes as eight little-endian key words, a 64-bit block counter
    (words 12-13) starting at 0, a 64-bit stream id (words 14-15) of 0; 12 rounds; the output is the keystream in order.
    rand_core's BlockRng reads a u64 as (next word) << 32 | (this word); `Standard` maps it to [0, 1) as (u64 >> 11) * 2^-53."""

    def __init__(self, seed: bytes, rounds: int = 12):
        assert len(seed) == 32
        self.key = [int.from_bytes(seed[4 * k:4 * k + 4], "little") for k in range(8)]
        self.rounds, self.counter, self.buf, self.pos = rounds, 0, [], 0

    def block(self, counter: int):
        m = 0xFFFFFFFF
        init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + self.key + [counter & m, (counter >> 32) & m, 0, 0]
        x = list(init)

        def rotl(v, n):
            return ((v << n) & m) | (v >> (32 - n))

        def quarter(a, b, c, d):
            x[a] = (x[a] + x[b]) & m; x[d] = rotl(x[d] ^ x[a], 16)
            x[c] = (x[c] + x[d]) & m; x[b] = rotl(x[b] ^ x[c], 12)
            x[a] = (x[a] + x[b]) & m; x[d] = rotl(x[d] ^ x[a], 8)
            x[c] = (x[c] + x[d]) & m; x[b] = rotl(x[b] ^ x[c], 7)

        for _ in range(self.rounds // 2):
            quarter(0, 4, 8, 12); quarter(1, 5, 9, 13); quarter(2, 6, 10, 14); quarter(3, 7, 11, 15)
            quarter(0, 5, 10, 15); quarter(1, 6, 11, 12); quarter(2, 7, 8, 13); quarter(3, 4, 9, 14)
        return [(x[k] + init[k]) & m for k in range(16)]

    def next_u32(self) -> int:
        if self.pos == len(self.buf):
            self.buf, self.pos = self.block(self.counter), 0
            self.counter += 1
        v = self.buf[self.pos]
        self.pos += 1
        return v

    def next_u64(self) -> int:
        lo = self.next_u32()
        return (self.next_u32() << 32) | lo

    def gen(self) -> float:
        return float(self.next_u64() >> 11) * 2.0 ** -53


def more_balls(optimized: bool = False) -> ExampleScene:
    """example_scenes.rs:63-138 (List root) and :141-150 (`more_balls_optimized`: the same list under Bvh::new): a checkered
    ground, three big spheres and 3,782 small ones with random materials"""
    camera = Camera(1.0, FRAC_PI_2, 7.5, 0.02, Transformation.lookat([6.0, 2.0, 4.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0]))
    textures = [Texture.Checker(1, 2), Texture.Solid(rgb(0.2, 0.3, 0.1)), Texture.Solid(rgb(0.9, 0.9, 0.9))]
    materials = [
        Material.new(Scatter.Lambert, Absorb.AlbedoMap(0), Emit.NONE),
        Material.new(Scatter.Lambert, Absorb.Albedo(rgb(0.1, 0.2, 0.5)), Emit.NONE),
        Material.new(Scatter.Metal(0.0), Absorb.Albedo(rgb(0.8, 0.6, 0.2)), Emit.NONE),
        Material.new(Scatter.Dielectric(1.5), Absorb.WhiteBody, Emit.NONE),
    ]
    root = [Hittable.Sphere([0.0, -1000.0, -1.0], 1000.0, 0), Hittable.Sphere([-4.0, 1.8, 0.0], 1.8, 1),
            Hittable.Sphere([4.0, 1.8, 0.0], 1.8, 2), Hittable.Sphere([0.0, 1.8, 0.0], 1.8, 3)]
    rng = StdRngStream(bytes([249] * 32))

    def closed_range(lo, hi):  # randomness.rs:12-16
        return lo + rng.gen() * (hi - lo)

    for x in range(-31, 31):
        for z in range(-31, 31):
            if z == 0:
                continue
            radius = closed_range(0.1, 0.3)
            cx = float(x) + closed_range(-0.5 + radius, 0.5 - radius)
            cz = float(z) + closed_range(-0.5 + radius, 0.5 - radius)
            root.append(Hittable.Sphere([cx, radius, cz], radius, len(materials)))
            albedo = rgb(rng.gen(), rng.gen(), rng.gen())
            if rng.gen() < 0.7:
                materials.append(Material.new(Scatter.Lambert, Absorb.Albedo(albedo), Emit.NONE))
            elif rng.gen() < 0.7:
                materials.append(Material.new(Scatter.Metal(rng.gen()), Absorb.Albedo(albedo), Emit.NONE))
            else:
                materials.append(Material.new(Scatter.Dielectric(1.5), Absorb.WhiteBody, Emit.NONE))
    return ExampleScene(camera, SceneData(materials, textures, []), "bvh" if optimized else "list", Hittable.concat(root), Emit.SkyGradient)


def more_balls_optimized() -> ExampleScene:
    """example_scenes.rs:141-150"""
    return more_balls(True)


def two_balls() -> ExampleScene:
    """example_scenes.rs:153-187 — Checker + Perlin textures on a 2-leaf BVH"""
    camera = Camera(1.0, FRAC_PI_2, 7.5, 0.0, Transformation.lookat([6.0, 0.0, 4.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0]))
    textures = [Texture.Solid(rgb(0.2, 0.2, 0.2)), Texture.Solid(rgb(0.9, 0.0, 0.5)), Texture.Checker(0, 1), Texture.Perlin(0)]
    materials = [
        Material.new(Scatter.Lambert, Absorb.AlbedoMap(2), Emit.NONE),
        Material.new(Scatter.Lambert, Absorb.AlbedoMap(3), Emit.NONE),
    ]
    root = Hittable.concat([Hittable.Sphere([0.0, -10.0, 0.0], 10.0, 0), Hittable.Sphere([0.0, 10.0, 0.0], 10.0, 1)])
    return ExampleScene(camera, SceneData(materials, textures, []), "bvh", root, Emit.SkyGradient)


def earth() -> ExampleScene:
    """example_scenes.rs:190-219 — one textured sphere (pins sphere uv + TGA orientation)"""
    camera = Camera(1.0, math.pi / 9.0, 1.0, 0.0, Transformation.lookat([13.0, 7.0, 3.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0]))
    textures = [Texture.Image(assets.earthmap())]
    materials = [Material.new(Scatter.Lambert, Absorb.AlbedoMap(0), Emit.NONE)]
    root = Hittable.Sphere([0.0, 0.0, 0.0], 2.0, 0)
    return ExampleScene(camera, SceneData(materials, textures, []), "bvh", root, Emit.SkyGradient)


def one_triangle() -> ExampleScene:
    """example_scenes.rs:222-262 — analytic pin: plane x+y+z=1, DebugNormals = (1,1,1)/sqrt(3)"""
    normal = [1.0 / math.sqrt(3.0)] * 3  # vector![1,1,1].normalize() = v / sqrt(0 + ((1+1)+1))
    mesh = Mesh.from_arrays([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]], normals=[normal] * 3, uvs=[[0.0, 0.0]] * 3,
                            indices=[0, 1, 2], material=0)
    materials = [
        Material.new(Scatter.NONE, Absorb.BlackBody, Emit.DebugNormals),
        Material.new(Scatter.Lambert, Absorb.Albedo(rgb(0.1, 0.2, 0.5)), Emit.NONE),
    ]
    root = Hittable.concat([Hittable.Triangle(0, 0), Hittable.Sphere([0.0, -1000.0, -1.0], 1000.0, 1)])
    camera = Camera(1.0, FRAC_PI_2, 1.0, 0.0, Transformation.lookat([2.0, 0.5, 1.0], [0.0, 0.0, 0.0], [0.0, 1.0, 0.0]))
    return ExampleScene(camera, SceneData(materials, [], [mesh]), "bvh", root, Emit.SkyGradient)


def _bunny_camera() -> Camera:
    """example_scenes.rs:337-347"""
    return Camera(1.0, FRAC_PI_4, 1.0, 0.0, Transformation.lookat([-1.5, 1.5, 2.5], [0.0, 0.5, 0.0], [0.0, 1.0, 0.0]))


def _bunny_like(mesh: Mesh, bunny_material: Material, sky_seed: int = 1) -> ExampleScene:
    materials = [bunny_material, Material.new(Scatter.Metal(0.05), Absorb.Albedo(rgb(0.8, 0.8, 0.8)), Emit.NONE)]
    textures = [Texture.Image(assets.sky_panorama(sky_seed))]  # stands in for assets/sky_panorama.tga (absent upstream)
    root = Hittable.concat([Hittable.triangles_of(mesh, 0), Hittable.Sphere([0.0, -1000.0, -1.0], 1000.0, 1)])
    return ExampleScene(_bunny_camera(), SceneData(materials, textures, [mesh]), "bvh", root, Emit.SkySphere(0))


def bunny() -> ExampleScene:
    """example_scenes.rs:309-350 — smooth bunny with DebugNormals, metal ground, sky sphere"""
    return _bunny_like(assets.bunny(), Material.new(Scatter.NONE, Absorb.BlackBody, Emit.DebugNormals))


def glass_bunny() -> ExampleScene:
    """example_scenes.rs:265-306 — flat-shaded dielectric bunny (images/demo.png)"""
    return _bunny_like(assets.bunny_flat(), Material.new(Scatter.Dielectric(1.5), Absorb.Albedo(rgb(0.7, 0.8, 0.7)), Emit.NONE))


def bunny_lambert() -> ExampleScene:
    """BASELINE config C1/C2/C3: `bunny()` with material 0 = Lambert albedo 0.8 ("Lambert mesh lit by sky_panorama")."""
    return _bunny_like(assets.bunny(), Material.new(Scatter.Lambert, Absorb.Albedo(rgb(0.8, 0.8, 0.8)), Emit.NONE))


def bunny_triangles_only() -> ExampleScene:
    """C2 variant without the ground sphere."""
    s = bunny_lambert()
    s.hittables = s.hittables[:-1]
    return s


def _translated(mesh: Mesh, offset, material: int) -> Mesh:
    v = mesh.vertices.copy()
    v["position"] = v["position"] + np.asarray(offset, dtype=np.float64)
    return Mesh(v, mesh.indices.copy(), material)


def demo() -> ExampleScene:
    """BASELINE config C4 "full demo scene": the reference has no single function for it, so it is composed from
    its parts through the same constructors — glass flat bunny + metal ground (glass_bunny, :265-306), a Lambert
    earthmap sphere (earth, :203-215), a Lambert smooth bunny, an emissive sphere, sky-panorama background."""
    glass = assets.bunny_flat()
    glass.material = 0
    lambert = _translated(assets.bunny(), [1.7, 0.0, -0.4], 2)
    materials = [
        Material.new(Scatter.Dielectric(1.5), Absorb.Albedo(rgb(0.7, 0.8, 0.7)), Emit.NONE),   # glass bunny
        Material.new(Scatter.Metal(0.05), Absorb.Albedo(rgb(0.8, 0.8, 0.8)), Emit.NONE),        # ground
        Material.new(Scatter.Lambert, Absorb.Albedo(rgb(0.8, 0.45, 0.3)), Emit.NONE),           # lambert bunny
        Material.new(Scatter.Lambert, Absorb.AlbedoMap(1), Emit.NONE),                          # earth
        Material.new(Scatter.NONE, Absorb.BlackBody, Emit.Color(rgb(6.0, 5.5, 4.5))),           # light
    ]
    textures = [Texture.Image(assets.sky_panorama(1)), Texture.Image(assets.earthmap())]
    root = Hittable.concat([
        Hittable.triangles_of(glass, 0),
        Hittable.triangles_of(lambert, 1),
        Hittable.Sphere([0.0, -1000.0, -1.0], 1000.0, 1),
        Hittable.Sphere([-1.7, 0.6, -0.2], 0.6, 3),
        Hittable.Sphere([0.3, 2.6, 0.8], 0.35, 4),
    ])
    camera = Camera(1.0, FRAC_PI_4, 1.0, 0.0, Transformation.lookat([-1.2, 1.7, 4.6], [0.0, 0.65, 0.0], [0.0, 1.0, 0.0]))
    return ExampleScene(camera, SceneData(materials, textures, [glass, lambert]), "bvh", root, Emit.SkySphere(0))


# exact rational rotations (Pythagorean triples): cos/sin need no libm, so every machine bakes the same field
_TRIPLES = [(3, 4, 5), (5, 12, 13), (8, 15, 17), (7, 24, 25), (20, 21, 29), (12, 35, 37), (9, 40, 41), (28, 45, 53)]


def _yaw(k: int):
    a, b, c = _TRIPLES[k % len(_TRIPLES)]
    q = (k // len(_TRIPLES)) % 8
    cs, sn = (a / c, b / c) if q % 2 == 0 else (b / c, a / c)
    if q & 2:
        cs = -cs
    if q & 4:
        sn = -sn
    return cs, sn


def bunny_field(nx: int = 64, nz: int = 32, pitch: float = 2.0, seed: int = 1) -> ExampleScene:
    """BASELINE config C5: nx*nz baked copies of bunny.obj (the reference has no instancing, mesh.rs:4) on a grid,
    per-copy yaw chosen by a hash of (seed, copy), cycling Lambert / Metal / Dielectric / Emissive; one ground sphere."""
    base = assets.bunny()
    materials = [
        Material.new(Scatter.Lambert, Absorb.Albedo(rgb(0.8, 0.8, 0.8)), Emit.NONE),
        Material.new(Scatter.Metal(0.05), Absorb.Albedo(rgb(0.8, 0.6, 0.2)), Emit.NONE),
        Material.new(Scatter.Dielectric(1.5), Absorb.Albedo(rgb(0.7, 0.8, 0.7)), Emit.NONE),
        Material.new(Scatter.NONE, Absorb.BlackBody, Emit.Color(rgb(2.0, 1.6, 1.2))),
        Material.new(Scatter.Metal(0.05), Absorb.Albedo(rgb(0.8, 0.8, 0.8)), Emit.NONE),  # ground
    ]
    textures = [Texture.Image(assets.sky_panorama(1))]
    meshes, parts = [], []
    pos, nrm = base.vertices["position"], base.vertices["normal"]
    for cz in range(nz):
        for cx in range(nx):
            k = cz * nx + cx
            h = (k * 2654435761 + seed * 40503) & 0xFFFFFFFF
            cs, sn = _yaw(h >> 8)
            v = base.vertices.copy()
            v["position"][:, 0] = (cs * pos[:, 0] + sn * pos[:, 2]) + (cx - (nx - 1) / 2.0) * pitch
            v["position"][:, 2] = (cs * pos[:, 2] - sn * pos[:, 0]) + (cz - (nz - 1) / 2.0) * pitch
            v["normal"][:, 0] = cs * nrm[:, 0] + sn * nrm[:, 2]
            v["normal"][:, 2] = cs * nrm[:, 2] - sn * nrm[:, 0]
            m = Mesh(v, base.indices, k % 4)
            meshes.append(m)
            parts.append(Hittable.triangles_of(m, k))
    parts.append(Hittable.Sphere([0.0, -1000.0, -1.0], 1000.0, 4))
    span = max(nx, nz) * pitch
    camera = Camera(1.0, FRAC_PI_4, 1.0, 0.0,
                    Transformation.lookat([-0.35 * span, 0.3 * span + 2.0, 0.75 * span + 2.0], [0.0, 0.5, 0.0], [0.0, 1.0, 0.0]))
    return ExampleScene(camera, SceneData(materials, textures, meshes), "bvh", Hittable.concat(parts), Emit.SkySphere(0))


def incoherent_rays(n: int, seed: int = 0x00C0FFEE, first: int = 0) -> np.ndarray:
    """BASELINE config C3: ray k has its origin uniform on the sphere of radius 3 about the bunny
    AABB centre and points at a uniform target inside the bunny AABB; draws come from stream RTP_RNG_STREAM_RAYS of
    the shared counter-based generator, counter = ray index. Built on the host with numpy (vectorised Philox)."""
    lo = np.array([-0.9438, -0.00078, -0.61679])
    hi = np.array([0.60779, 1.53609, 0.58715])
    centre = np.array([-0.168, 0.768, -0.015])
    idx = np.arange(first, first + n, dtype=np.uint64)
    d = philox_draws(seed, (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32), A.RNG_STREAM_RAYS, 6)
    z = 2.0 * d[:, 0] - 1.0
    phi = 2.0 * math.pi * d[:, 1]
    r = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    origin = centre + 3.0 * np.stack([r * np.cos(phi), z, r * np.sin(phi)], axis=1)
    target = lo + d[:, 2:5] * (hi - lo)
    direction = target - origin
    direction /= np.sqrt((direction * direction).sum(axis=1, keepdims=True))
    rays = np.zeros(n, dtype=A.RAY_DTYPE)
    rays["origin"], rays["direction"], rays["t_min"], rays["t_max"] = origin, direction, 1e-3, np.inf
    return rays


def philox_draws(seed: int, index_lo: np.ndarray, index_hi: np.ndarray, stream: int, n_draws: int) -> np.ndarray:
    """Vectorised host copy of the shared stream (rtp.h rtp_rng_draws): [len(index), n_draws] float64 in [0,1)."""
    n = len(index_lo)
    out = np.empty((n, n_draws), dtype=np.float64)
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for block in range((n_draws + 1) // 2):
        c0 = index_lo.astype(np.uint64)
        c1 = index_hi.astype(np.uint64)
        c2 = np.full(n, block, dtype=np.uint64)
        c3 = np.full(n, stream, dtype=np.uint64)
        k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
        for _ in range(10):
            p0, p1 = M0 * c0, M1 * c2
            c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & mask, p1 & mask, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & mask, p0 & mask
            k0 = (k0 + np.uint64(0x9E3779B9)) & mask
            k1 = (k1 + np.uint64(0xBB67AE85)) & mask
        for pair in range(2):
            k = 2 * block + pair
            if k < n_draws:
                u = ((c3 if pair else c1) << np.uint64(32)) | (c2 if pair else c0)
                out[:, k] = (u >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return out
