"""A second, independent restatement of the reference's geometry path — pure Python, written from the Rust sources and
LITERAL: `Bvh::new` with its recursive median split (bvh.rs:36-91), the recursive `hit_node` with a cloned ray whose t_max
shrinks after a left hit (bvh.rs:93-124), `AABB::union` / `AABB::collide` with Rust's NaN-ignoring f64::min/max
(utility.rs:130-154), `hit_triangle` (hittable.rs:65-108), `hit_sphere` (hittable.rs:39-63) and the bounding boxes
(hittable.rs:124-140). The oracle's closest hits (ids, t, position, normal, uv) must equal it bit for bit; the only convention
added is the one stated in rtp.h for centroid ties (the reference's sort_unstable_by leaves them to the Rust std version):
equal keys order by LeafId. This also exercises the equivalence DESIGN.md §2 rests on, since the oracle and the product do not
recurse. No GPU needed."""
import math

import numpy as np
import pytest

import oracle
from rtp_b200 import _abi as A
from rtp_b200 import api, scenes

INF = float("inf")
SMOL = 1e-7  # utility.rs:31
MISS = 0xFFFFFFFF


def fmin(a, b):  # Rust f64::min: a NaN operand is ignored
    if a != a:
        return b
    if b != b:
        return a
    return a if a < b else b


def fmax(a, b):
    if a != a:
        return b
    if b != b:
        return a
    return a if a > b else b


def mul(a, b):  # IEEE multiply where Python would raise nothing but 0 * inf must be NaN (it is)
    return a * b


def div(a, b):  # IEEE divide: x / 0 = +-inf, 0 / 0 = NaN
    if b == 0.0:
        if a == 0.0 or a != a:
            return float("nan")
        return math.copysign(INF, a) * math.copysign(1.0, b)
    return a / b


class Box:
    def __init__(self, lo, hi):
        self.min, self.max = tuple(lo), tuple(hi)

    def union(self, o):  # utility.rs:130-135
        return Box([fmin(self.min[k], o.min[k]) for k in range(3)], [fmax(self.max[k], o.max[k]) for k in range(3)])

    def collide(self, o, inv, t_min, t_max):  # utility.rs:137-154
        t0 = [mul(self.min[k] - o[k], inv[k]) for k in range(3)]
        t1 = [mul(self.max[k] - o[k], inv[k]) for k in range(3)]
        lo = fmax(fmax(fmax(t_min, fmin(t0[0], t1[0])), fmin(t0[1], t1[1])), fmin(t0[2], t1[2]))
        hi = fmin(fmin(fmin(t_max, fmax(t0[0], t1[0])), fmax(t0[1], t1[1])), fmax(t0[2], t1[2]))
        return hi >= lo


class LiteralBvh:
    def __init__(self, sc):
        self.sc = sc
        self.mesh = sc.scene_data.mesh_table
        self.nodes = []  # ("leaf", box, leaf_id) | ("branch", box, left, right)
        content = [(i, self.bounding_box(h)) for i, h in enumerate(sc.hittables)]
        self.root = self.make(content, 0)

    def tri(self, h):
        m = self.mesh[int(h["mesh"])]
        i0 = int(h["triangle"])
        return [m.vertices[int(m.indices[i0 + k])] for k in range(3)], m.material

    def bounding_box(self, h):
        if int(h["kind"]) == A.HITTABLE_SPHERE:  # hittable.rs:124-129
            c, r = [float(x) for x in h["center"]], float(h["radius"])
            return Box([c[k] - r for k in range(3)], [c[k] + r for k in range(3)])
        (va, vb, vc), _ = self.tri(h)  # hittable.rs:131-140
        a, b, c = ([float(x) for x in v["position"]] for v in (va, vb, vc))
        return Box([fmin(fmin(a[k], b[k]), c[k]) for k in range(3)], [fmax(fmax(a[k], b[k]), c[k]) for k in range(3)])

    def make(self, content, axis):  # bvh.rs:36-67
        if len(content) == 1:
            leaf, box = content[0]
            self.nodes.append(("leaf", box, leaf))
            return len(self.nodes) - 1
        content = sorted(content, key=lambda e: (0.5 * (e[1].min[axis] + e[1].max[axis]), e[0]))
        half = len(content) // 2
        left = self.make(content[:half], (axis + 1) % 3)
        right = self.make(content[half:], (axis + 1) % 3)
        self.nodes.append(("branch", self.nodes[left][1].union(self.nodes[right][1]), left, right))
        return len(self.nodes) - 1

    def hit_leaf(self, h, o, d, t_min, t_max):
        if int(h["kind"]) == A.HITTABLE_SPHERE:  # hittable.rs:39-63
            c, r = [float(x) for x in h["center"]], float(h["radius"])
            oc = [o[k] - c[k] for k in range(3)]
            a = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]
            half_b = (d[0] * oc[0] + d[1] * oc[1]) + d[2] * oc[2]
            cc = ((oc[0] * oc[0] + oc[1] * oc[1]) + oc[2] * oc[2]) - r * r
            delta = half_b * half_b - a * cc
            if delta <= 0.0:
                return None
            sq = math.sqrt(delta)
            t = div(-half_b - sq, a)
            if t < t_min or t > t_max:
                t = div(-half_b + sq, a)
                if t < t_min or t > t_max:
                    return None
            p = [o[k] + t * d[k] for k in range(3)]
            n = [p[k] - c[k] for k in range(3)]
            ln = math.sqrt((n[0] * n[0] + n[1] * n[1]) + n[2] * n[2])
            n = [div(n[k], ln) for k in range(3)]
            return t, p, n, None, int(h["material"])
        (va, vb, vc), material = self.tri(h)  # hittable.rs:65-108
        a, b, c = ([float(x) for x in v["position"]] for v in (va, vb, vc))
        ba = [a[k] - b[k] for k in range(3)]
        ca = [a[k] - c[k] for k in range(3)]
        pa = [a[k] - o[k] for k in range(3)]
        det = ba[0] * ca[1] * d[2] + ba[1] * ca[2] * d[0] + ba[2] * ca[0] * d[1] - ba[0] * ca[2] * d[1] - ba[1] * ca[0] * d[2] - ba[2] * ca[1] * d[0]
        if abs(det) < SMOL:
            return None
        inv_det = 1.0 / det
        t = (pa[0] * (ba[1] * ca[2] - ba[2] * ca[1]) + pa[1] * (ba[2] * ca[0] - ba[0] * ca[2]) + pa[2] * (ba[0] * ca[1] - ba[1] * ca[0])) * inv_det
        u = (pa[0] * (ca[1] * d[2] - ca[2] * d[1]) + pa[1] * (ca[2] * d[0] - ca[0] * d[2]) + pa[2] * (ca[0] * d[1] - ca[1] * d[0])) * inv_det
        v = (pa[0] * (ba[2] * d[1] - ba[1] * d[2]) + pa[1] * (ba[0] * d[2] - ba[2] * d[0]) + pa[2] * (ba[1] * d[0] - ba[0] * d[1])) * inv_det
        w = 1.0 - u - v
        if t < t_min or t > t_max or u < 0.0 or v < 0.0 or w < 0.0:
            return None
        p = [o[k] + t * d[k] for k in range(3)]
        na, nb, nc = ([float(x) for x in vv["normal"]] for vv in (va, vb, vc))
        ta, tb, tc = ([float(x) for x in vv["uv"]] for vv in (va, vb, vc))
        n = [w * na[k] + u * nb[k] + v * nc[k] for k in range(3)]
        uv = [w * ta[k] + u * tb[k] + v * tc[k] for k in range(2)]
        return t, p, n, uv, int(material)

    def hit_node(self, node, o, d, inv, t_min, t_max):  # bvh.rs:93-119
        nd = self.nodes[node]
        if nd[0] == "leaf":
            if nd[1].collide(o, inv, t_min, t_max):
                r = self.hit_leaf(self.sc.hittables[nd[2]], o, d, t_min, t_max)
                return None if r is None else (r, nd[2])
            return None
        if not nd[1].collide(o, inv, t_min, t_max):
            return None
        hit = None
        new = self.hit_node(nd[2], o, d, inv, t_min, t_max)
        if new is not None:
            t_max = new[0][0]
            hit = new
        new = self.hit_node(nd[3], o, d, inv, t_min, t_max)
        if new is not None:
            hit = new
        return hit

    def hit(self, ray):  # bvh.rs:121-124, utility.rs:71-77
        o = [float(x) for x in ray["origin"]]
        d = [float(x) for x in ray["direction"]]
        inv = [div(1.0, d[k]) for k in range(3)]
        return self.hit_node(self.root, o, d, inv, float(ray["t_min"]), float(ray["t_max"]))


def soup_scene(seed, n_tri, with_duplicates):
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-1.5, 1.5, (n_tri, 3))
    pos = (centres[:, None, :] + rng.normal(0.0, 0.25, (n_tri, 3, 3))).reshape(-1, 3)
    if with_duplicates:  # coincident triangles (exact t ties), axis-aligned ones (zero-thickness boxes), shared vertices on a lattice
        pos[: 3 * 10] = np.round(pos[: 3 * 10] * 4) / 4
        pos[3 * 10: 3 * 20] = pos[: 3 * 10]
        pos[3 * 20: 3 * 30, 2] = 0.5
    nrm = rng.normal(size=(len(pos), 3))
    uvs = rng.uniform(size=(len(pos), 2))
    mesh_a = api.Mesh.from_arrays(pos[: len(pos) // 2 // 3 * 3], nrm[: len(pos) // 2 // 3 * 3], uvs[: len(pos) // 2 // 3 * 3], material=0)
    rest = len(pos) // 2 // 3 * 3
    perm = rng.permutation(len(pos) - rest).astype(np.uint32)  # an indexed mesh with shuffled, non-trivial indices
    perm = perm[: len(perm) // 3 * 3]
    mesh_b = api.Mesh.from_arrays(pos[rest:], nrm[rest:], uvs[rest:], indices=perm, material=1)
    mats = [api.Material.new(api.Scatter.Lambert, api.Absorb.WhiteBody, api.Emit.NONE)] * 3
    spheres = [api.Hittable.Sphere([0.2, -0.3, 0.1], 0.6, 2), api.Hittable.Sphere([0.0, -101.5, 0.0], 100.0, 2), api.Hittable.Sphere([0.2, -0.3, 0.1], 0.6, 1)]
    hittables = api.Hittable.concat([api.Hittable.triangles_of(mesh_a, 0), spheres[0], api.Hittable.triangles_of(mesh_b, 1), spheres[1], spheres[2]])
    sc = api.ExampleScene(scenes._bunny_camera(), api.SceneData(mats, [], [mesh_a, mesh_b]), "bvh", hittables, api.Emit.SkyGradient)
    return sc


def soup_rays(seed, n):
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=A.RAY_DTYPE)
    o = rng.normal(size=(n, 3))
    o *= rng.uniform(2.5, 4.0, (n, 1)) / np.linalg.norm(o, axis=1, keepdims=True)
    tgt = rng.uniform(-1.2, 1.2, (n, 3))
    d = tgt - o
    rays["origin"], rays["direction"] = o, d / np.linalg.norm(d, axis=1, keepdims=True)
    rays["t_min"], rays["t_max"] = 1e-3, np.inf
    k = n // 8
    rays["direction"][:k] *= rng.uniform(0.1, 7.0, (k, 1))          # not normalised
    rays["t_max"][k:2 * k] = rng.uniform(1.0, 5.0, k)                # finite t_max
    rays["origin"][2 * k:3 * k] = np.round(rays["origin"][2 * k:3 * k] * 4) / 4   # axis-parallel rays through the lattice
    axis = rng.integers(0, 3, k)
    dd = np.zeros((k, 3))
    dd[np.arange(k), axis] = -np.sign(rays["origin"][2 * k:3 * k][np.arange(k), axis] + 1e-9)
    rays["direction"][2 * k:3 * k] = dd
    rays["origin"][3 * k:4 * k] = rng.uniform(-1.0, 1.0, (k, 3))     # origins inside the soup
    return rays


@pytest.mark.parametrize("seed,n_tri,dup", [(1, 120, True), (2, 257, False), (3, 40, True)])
def test_literal_recursive_bvh_equals_oracle(seed, n_tri, dup):
    sc = soup_scene(seed, n_tri, dup)
    lit = LiteralBvh(sc)
    o = oracle.Scene(sc)
    info = o.info()
    assert info.n_leaves == len(sc.hittables) and info.n_nodes == len(lit.nodes) == 2 * len(sc.hittables) - 1
    # the depth-first leaf order of the literal tree is the oracle's
    order = []

    def walk(n):
        nd = lit.nodes[n]
        if nd[0] == "leaf":
            order.append(nd[2])
        else:
            walk(nd[2]); walk(nd[3])

    walk(lit.root)
    assert order == list(o.leaf_order())
    rays = soup_rays(seed + 10, 1600)
    got = o.hit_full(rays)
    n_hit = 0
    for k, ray in enumerate(rays):
        want = lit.hit(ray)
        g = got[k]
        if want is None:
            assert g["leaf"] == MISS, (k, g)
            continue
        (t, p, n, uv, material), leaf = want
        n_hit += 1
        assert (int(g["leaf"]), int(g["material"])) == (leaf, material), (k, g, want)
        assert np.float64(t).tobytes() == g["t"].tobytes(), (k, t, g["t"])
        assert np.array(p).tobytes() == g["position"].tobytes() and np.array(n).tobytes() == g["normal"].tobytes(), (k, p, n, g)
        if uv is not None:
            assert np.array(uv).tobytes() == g["uv"].tobytes(), (k, uv, g["uv"])
    assert n_hit > len(rays) // 3
    o.close()
