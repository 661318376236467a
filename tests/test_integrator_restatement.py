"""A second, independent restatement of the reference's integrator — pure Python floats, written from the Rust sources, sharing no
code with oracle/rtp_oracle.c — run against the oracle path by path. Covers Camera::shoot with a thin lens (render.rs:32-52),
make_uv_jitter (render.rs:74-82), trace_path / _first / _continue (render.rs:94-146), Material::evaluate and the three scatter
models (material.rs:99-180), reflect / refract (utility.rs:106-119), hit_sphere (hittable.rs:39-63), hit_list (hittable.rs:110-120),
SkyGradient (material.rs:54-57), the distributions of randomness.rs:21-82 and the draw order of the shared counter-based stream
(include/rtp.h). Python floats are IEEE doubles and CPython never contracts a*b+c, like rustc. No GPU needed."""
import math

import numpy as np
import pytest

import oracle
from rtp_b200 import _abi as A
from rtp_b200 import api, scenes
from rtp_b200.api import rgb

INF = float("inf")
RAY_EPSILON = 1e-3  # utility.rs:30


def dot(a, b):  # nalgebra Vector3::dot, (x x' + y y') + z z'
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def sub(a, b): return (a[0] - b[0], a[1] - b[1], a[2] - b[2])
def add(a, b): return (a[0] + b[0], a[1] + b[1], a[2] + b[2])
def scale(s, a): return (s * a[0], s * a[1], s * a[2])
def cmul(a, b): return (a[0] * b[0], a[1] * b[1], a[2] * b[2])


def normalize(a):  # unscale by the norm: one sqrt, three divides
    n = math.sqrt(dot(a, a))
    return (a[0] / n, a[1] / n, a[2] / n)


def mat_vec(m, v):  # Matrix3 * Vector3, accumulated column by column
    return tuple((m[r][0] * v[0] + m[r][1] * v[1]) + m[r][2] * v[2] for r in range(3))


def powi(x, n):  # compiler-rt __powidf2: square and multiply
    r = 1.0
    while True:
        if n & 1:
            r *= x
        n //= 2
        if n == 0:
            return r
        x *= x


class Stream:
    """draws of one (pixel, sample): rtp.h RTP_RNG_STREAM_PATH, draw k = word pair k & 1 of Philox block k >> 1"""

    def __init__(self, seed, pixel, sample, first=None):
        self.seed, self.pixel, self.sample = seed, pixel, sample
        self.d = first if first is not None else self.draws(64)
        self.k = 0

    def draws(self, n):
        return scenes.philox_draws(self.seed, np.array([self.pixel], dtype=np.uint32), np.array([self.sample], dtype=np.uint32), A.RNG_STREAM_PATH, n)[0]

    def gen(self):
        if self.k >= len(self.d):
            self.d = self.draws(4 * len(self.d))
        v = float(self.d[self.k])
        self.k += 1
        return v


def unit_disk(rng):  # randomness.rs:21-34
    while True:
        x = 2.0 * rng.gen() - 1.0
        y = 2.0 * rng.gen() - 1.0
        if x * x + y * y < 1.0:
            return x, y


def unit_ball(rng):  # randomness.rs:39-53
    while True:
        v = (2.0 * rng.gen() - 1.0, 2.0 * rng.gen() - 1.0, 2.0 * rng.gen() - 1.0)
        if dot(v, v) < 1.0:
            return v


def unit_sphere(rng):  # randomness.rs:58-73
    while True:
        x = 2.0 * rng.gen() - 1.0
        y = 2.0 * rng.gen() - 1.0
        s = x * x + y * y
        if s < 1.0:
            n = 2.0 * math.sqrt(1.0 - s)
            return (x * n, y * n, 1.0 - 2.0 * s)


def reflect(i, n):  # utility.rs:106-108
    return sub(i, scale(2.0 * dot(i, n), n))


def refract(i, n, eta):  # utility.rs:111-119
    cos_theta = dot(n, i)
    k = 1.0 - eta * eta * (1.0 - cos_theta * cos_theta)
    if k < 0.0:
        return None
    return sub(scale(eta, i), scale(eta * cos_theta + math.sqrt(k), n))


def hit_sphere(center, radius, o, d, t_min, t_max):  # hittable.rs:39-63 (uv not needed by these scenes)
    to_center = sub(o, center)
    a = dot(d, d)
    half_b = dot(d, to_center)
    c = dot(to_center, to_center) - radius * radius
    delta = half_b * half_b - a * c
    if delta <= 0.0:
        return None
    sq = math.sqrt(delta)
    t = (-half_b - sq) / a
    if t < t_min or t > t_max:
        t = (-half_b + sq) / a
        if t < t_min or t > t_max:
            return None
    position = add(o, scale(t, d))
    normal = normalize(sub(position, center))
    uv = (0.5 - math.atan2(normal[2], normal[0]) / math.tau, math.asin(normal[1]) / math.pi + 0.5)  # hittable.rs:61
    return t, position, normal, uv


class Restated:
    def __init__(self, sc, width, height, max_bounce, seed):
        self.sc, self.w, self.h, self.max_bounce, self.seed = sc, width, height, max_bounce, seed
        cam = sc.camera
        self.cam = cam
        self.aspect = width / height  # main.rs:22
        self.m = [[float(cam.transformation.orientation[r][c]) for c in range(3)] for r in range(3)]
        self.pos = tuple(float(x) for x in cam.transformation.position)
        self.rays = 0
        # the first 64 draws of every (pixel, sample) of the frame in one vectorised call
        pix = np.repeat(np.arange(width * height, dtype=np.uint32), 8)
        smp = np.tile(np.arange(8, dtype=np.uint32), width * height)
        self.first = scenes.philox_draws(seed, pix, smp, A.RNG_STREAM_PATH, 64).reshape(width * height, 8, 64)

    def scene_hit(self, o, d, t_min, t_max):  # hittable.rs:110-120 on a List root
        best = None
        for h in self.sc.hittables:
            r = hit_sphere(tuple(float(x) for x in h["center"]), float(h["radius"]), o, d, t_min, t_max)
            if r is not None:
                t_max = r[0]
                best = (r, int(h["material"]))
        return best

    def texture(self, tid, uv):  # texture.rs:20-49
        t = self.sc.scene_data.texture_table[tid]
        if t.kind == A.TEXTURE_SOLID:
            return tuple(float(x) for x in t.color)
        assert t.kind == A.TEXTURE_IMAGE
        h, w = t.image.shape[:2]

        def texel(x, n):  # (x * n).clamp(0, n - 1) as u32: truncation, NaN -> 0
            v = x * float(n)
            if v != v:
                return 0
            return int(min(max(v, 0.0), float(n) - 1.0))

        px = t.image[texel(uv[1], h), texel(uv[0], w)]  # Array2d::get(i, j) = data[i + j * width], row 0 at the bottom
        return (float(px[0]) / 255.0, float(px[1]) / 255.0, float(px[2]) / 255.0)

    def emit(self, e, d, normal, uv):  # material.rs:49-60
        if e.kind == A.EMIT_SKY_SPHERE:
            return self.texture(e.texture, uv)
        if e.kind == A.EMIT_NONE:
            return (0.0, 0.0, 0.0)
        if e.kind == A.EMIT_COLOR:
            return tuple(float(x) for x in e.color)
        if e.kind == A.EMIT_DEBUG_NORMALS:
            return normal
        assert e.kind == A.EMIT_SKY_GRADIENT
        t = 0.5 * (d[1] / math.sqrt(dot(d, d)) + 1.0)
        return add(scale(1.0 - t, (1.0, 1.0, 1.0)), scale(t, (0.5, 0.7, 1.0)))

    def absorb(self, a, uv):  # material.rs:74-81
        if a.kind == A.ABSORB_BLACKBODY:
            return (0.0, 0.0, 0.0)
        if a.kind == A.ABSORB_WHITEBODY:
            return (1.0, 1.0, 1.0)
        if a.kind == A.ABSORB_ALBEDO:
            return tuple(float(x) for x in a.color)
        return self.texture(a.texture, uv)

    def scatter(self, s, d, position, normal, rng):  # material.rs:27-34, 115-180; returns the scattered direction or None
        if s.kind == A.SCATTER_NONE:
            return None
        if s.kind == A.SCATTER_LAMBERT:
            if dot(normal, d) > 0.0:
                return None
            return normalize(add(normal, unit_sphere(rng)))
        if s.kind == A.SCATTER_METAL:
            if dot(normal, d) > 0.0:
                return None
            r = normalize(add(reflect(d, normal), scale(s.param, unit_ball(rng))))
            return None if dot(normal, r) < 0.0 else r
        ior = s.param
        if dot(normal, d) > 0.0:
            eta, n = ior, (-normal[0], -normal[1], -normal[2])
        else:
            eta, n = 1.0 / ior, normal
        r0 = powi((1.0 - eta) / (1.0 + eta), 2)
        reflectance = r0 + (1.0 - r0) * powi(1.0 + dot(n, d), 5)
        if rng.gen() < reflectance:
            return reflect(d, n)
        t = refract(d, n, eta)
        return t if t is not None else reflect(d, n)

    def shade(self, o, d, depth, rng, first):  # render.rs:102-146
        self.rays += 1
        hit = self.scene_hit(o, d, RAY_EPSILON, INF)
        if hit is None:
            # Hit::at_infinity (utility.rs:93-100): the direction stands in for position and normal, equirectangular uv
            uv_inf = (0.5 - math.atan2(d[2], d[0]) / math.tau, math.asin(d[1]) / math.pi + 0.5) if self.sc.background.kind == A.EMIT_SKY_SPHERE else None
            return self.emit(self.sc.background, d, d, uv_inf), False
        (t, position, normal, uv), mid = hit
        mat = self.sc.scene_data.material_table[mid]
        out_dir = self.scatter(mat.scatter, d, position, normal, rng)  # order: scatter, absorb, emit (material.rs:104-110)
        absorb = self.absorb(mat.absorb, uv)
        emit = self.emit(mat.emit, d, normal, uv)
        if out_dir is None:
            return add(emit, (0.0, 0.0, 0.0)), True
        if depth - 1 == 0:
            inner = (0.0, 0.0, 0.0)  # render.rs:128-131
        else:
            inner, _ = self.shade(position, out_dir, depth - 1, rng, False)
        return add(emit, cmul(absorb, inner)), True

    def path(self, i, j, s):
        rng = Stream(self.seed, j * self.w + i, s, self.first[j * self.w + i, s] if s < 8 else None)
        u = (i + rng.gen()) / self.w  # render.rs:78-79
        v = (j + rng.gen()) / self.h
        cam = self.cam
        tan_fov = math.tan(0.5 * cam.fov)
        lx, ly = unit_disk(rng)  # render.rs:36: drawn even for a pinhole
        origin = (cam.lens_radius * lx, cam.lens_radius * ly, 0.0)
        direction = normalize(sub(((2.0 * u - 1.0) * tan_fov * cam.focal_dist * self.aspect, (2.0 * v - 1.0) * tan_fov * cam.focal_dist, -cam.focal_dist), origin))
        d = mat_vec(self.m, direction)
        o = add(mat_vec(self.m, origin), self.pos)
        self.rays = 0
        color, hit = self.shade(o, d, self.max_bounce, rng, True)
        return color, hit, self.rays


def zoo_scene():
    """List root with every scatter / absorb / emit variant that needs no texture lookup by uv"""
    camera = api.Camera(1.0, 1.1, 2.5, 0.05, api.Transformation.lookat([0.3, 1.2, 3.0], [0.0, 0.2, 0.0], [0.0, 1.0, 0.0]))
    textures = [api.Texture.Solid(rgb(0.3, 0.7, 0.4))]
    materials = [
        api.Material.new(api.Scatter.Lambert, api.Absorb.AlbedoMap(0), api.Emit.NONE),
        api.Material.new(api.Scatter.Metal(0.35), api.Absorb.Albedo(rgb(0.9, 0.8, 0.7)), api.Emit.NONE),
        api.Material.new(api.Scatter.Dielectric(1.33), api.Absorb.Albedo(rgb(0.95, 0.97, 1.0)), api.Emit.NONE),
        api.Material.new(api.Scatter.NONE, api.Absorb.BlackBody, api.Emit.Color(rgb(4.0, 3.0, 2.0))),
        api.Material.new(api.Scatter.Lambert, api.Absorb.WhiteBody, api.Emit.DebugNormals),
        api.Material.new(api.Scatter.Dielectric(0.8), api.Absorb.WhiteBody, api.Emit.Color(rgb(0.01, 0.0, 0.02))),
    ]
    root = api.Hittable.concat([
        api.Hittable.Sphere([0.0, -50.0, 0.0], 50.0, 0), api.Hittable.Sphere([-0.9, 0.4, 0.2], 0.4, 1), api.Hittable.Sphere([0.0, 0.45, 0.0], 0.45, 2),
        api.Hittable.Sphere([0.9, 0.35, 0.3], 0.35, 4), api.Hittable.Sphere([0.2, 1.6, -0.5], 0.3, 3), api.Hittable.Sphere([0.0, 0.45, 0.0], 0.25, 5),
        api.Hittable.Sphere([0.0, 0.45, 0.0], 0.45, 2),  # an exact duplicate: the later one must win every tie (hittable.rs:99, 116-117)
    ])
    return api.ExampleScene(camera, api.SceneData(materials, textures, []), "list", root, api.Emit.SkyGradient)


def textured_scene():
    """earthmap.tga on a Lambert sphere and on an emitter, the synthetic sky panorama as SkySphere background (unnormalised
    directions reach asin through the thin lens' non-unit camera basis): sphere uv, Hit::at_infinity uv and sample_image"""
    from rtp_b200 import assets

    sc = zoo_scene()
    sc.scene_data.texture_table += [api.Texture.Image(assets.earthmap()), api.Texture.Image(assets.sky_panorama())]
    sc.scene_data.material_table[4] = api.Material.new(api.Scatter.Lambert, api.Absorb.AlbedoMap(1), api.Emit.NONE)
    sc.scene_data.material_table[1] = api.Material.new(api.Scatter.Metal(0.1), api.Absorb.AlbedoMap(1), api.Emit.NONE)
    sc.background = api.Emit.SkySphere(2)
    return sc


@pytest.mark.parametrize("name,depth", [("three_balls", 8), ("zoo", 8), ("zoo", 2), ("three_balls", 1), ("textured", 8)])
def test_pure_python_integrator_equals_oracle_path_by_path(name, depth):
    sc = scenes.three_balls() if name == "three_balls" else (zoo_scene() if name == "zoo" else textured_scene())
    w, h, spp, seed = 36, 24, 3, 5
    o = oracle.Scene(sc)
    r = Restated(sc, w, h, depth, seed)
    scattered = 0
    for j in range(0, h):
        for i in range(0, w):
            for s in range(spp):
                want_rgb, want_hit, want_rays = o.trace_one(w, h, i, j, s, max_bounce=depth, seed=seed)
                got_rgb, got_hit, got_rays = r.path(i, j, s)
                assert got_hit == want_hit and got_rays == want_rays, (i, j, s, got_rays, want_rays)
                assert np.array(got_rgb, dtype=np.float64).tobytes() == np.asarray(want_rgb, dtype=np.float64).tobytes(), (i, j, s, got_rgb, want_rgb)
                scattered += got_rays > 1
    assert depth == 1 or scattered > w * h * spp // 4
    # and the frame is the per-pixel sum in sample order divided by the sample count (main.rs:78-87)
    img, fg, st = o.render(w, h, spp, max_bounce=depth, seed=seed)
    for (i, j) in ((0, 0), (7, 5), (35, 23)):
        acc, hits = (0.0, 0.0, 0.0), 0.0
        for s in range(spp):
            c, hit, _ = r.path(i, j, s)
            acc = add(acc, c)
            hits += 1.0 if hit else 0.0
        want = tuple(x / float(spp) for x in acc)
        assert np.array(want).tobytes() == img[j, i].tobytes() and fg[j, i] == hits / float(spp)
    o.close()


# ---------------------------------------------------------------------------------------------
# procedural textures (texture.rs:51-119, randomness.rs:86-110) restated with Python integers
# ---------------------------------------------------------------------------------------------

def _wrap(x):  # Wrapping<isize>: two's complement, 64 bits
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >> 63 else x


def noise_integer(x, y, z, seed):  # randomness.rs:91-104; `>>` on isize is arithmetic, Python's `>>` on a negative int is too
    a, b, c, d = 0x369E6D3B899E43CF, 0x53F89E7FFDA3B07D, 0x3B13C1CA4937E629, 0x577C2C6E4019D645
    h = _wrap(_wrap(_wrap(_wrap(a * x) + _wrap(b * y)) + _wrap(c * z)) + _wrap(d * seed))
    h = _wrap((h >> 13) ^ h)
    return _wrap(_wrap(h * _wrap(_wrap(_wrap(h * h) * 60493) + 19990303)) + 1376312589)


def noise_real(x, y, z, seed):  # randomness.rs:107-109: isize as f64 (round to nearest) / isize::MAX as f64 (= 2^63)
    return float(noise_integer(x, y, z, seed)) / float((1 << 63) - 1)


def sample_noise(p, seed):  # texture.rs:62-68
    x = noise_real(math.floor(p[0]), math.floor(p[1]), math.floor(p[2]), seed)
    return 0.5 * x + 0.5


def sample_perlin(p, seed):  # texture.rs:70-119
    fl = [math.floor(v) for v in p]
    cl = [v + 1 for v in fl]

    def grad_dot(cx, cy, cz):
        g = (noise_real(cx, cy, cz, seed + 1), noise_real(cx, cy, cz, seed + 2), noise_real(cx, cy, cz, seed + 3))
        return dot((p[0] - float(cx), p[1] - float(cy), p[2] - float(cz)), g)

    k1, k2 = grad_dot(fl[0], fl[1], fl[2]), grad_dot(cl[0], fl[1], fl[2])
    k3, k4 = grad_dot(fl[0], cl[1], fl[2]), grad_dot(cl[0], cl[1], fl[2])
    k5, k6 = grad_dot(fl[0], fl[1], cl[2]), grad_dot(cl[0], fl[1], cl[2])
    k7, k8 = grad_dot(fl[0], cl[1], cl[2]), grad_dot(cl[0], cl[1], cl[2])
    t = [p[k] - float(fl[k]) for k in range(3)]
    t = [(v * (v * 6.0 - 15.0) + 10.0) * v * v * v for v in t]

    def mix(a, b, w):
        return (b - a) * w + a

    k12, k34, k56, k78 = mix(k1, k2, t[0]), mix(k3, k4, t[0]), mix(k5, k6, t[0]), mix(k7, k8, t[0])
    return 0.5 * mix(mix(k12, k34, t[1]), mix(k56, k78, t[1]), t[2]) + 0.5


def test_procedural_textures_equal_python_restatement():
    sc = scenes.two_balls()
    sc.scene_data.texture_table += [api.Texture.Noise(3), api.Texture.Noise(-77), api.Texture.Perlin(12345), api.Texture.Perlin(-1)]
    o = oracle.Scene(sc)
    rng = np.random.default_rng(9)
    pts = np.concatenate([rng.uniform(-40.0, 40.0, (300, 3)), rng.uniform(-1e6, 1e6, (50, 3)), np.round(rng.uniform(-5, 5, (30, 3))),
                          np.array([[0.0, 0.0, 0.0], [-0.0, 0.5, -0.5], [1e-300, -1e-300, 0.999999999999], [2.5, -3.5, 4.5]])])
    for p in pts:
        p = [float(v) for v in p]
        for tid, seed in ((4, 3), (5, -77)):
            want = sample_noise(p, seed)
            got = o.texture_sample(tid, p, [0.0, 0.0])
            assert np.float64(want).tobytes() == got[0].tobytes() and got[0] == got[1] == got[2], (p, seed, want, got)
        for tid, seed in ((3, 0), (6, 12345), (7, -1)):  # texture 3 is two_balls' own Perlin(0)
            want = sample_perlin(p, seed)
            got = o.texture_sample(tid, p, [0.0, 0.0])
            assert np.float64(want).tobytes() == got[0].tobytes() and got[0] == got[1] == got[2], (p, seed, want, got)
    o.close()
