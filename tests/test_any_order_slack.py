"""The numerical lemma the any-order walk rests on, stress-tested on the CPU: for a primitive that passes its
static tests, the computed slab entry of its own box never exceeds the computed hit distance by more than slack(t). The formulas
are the reference's (hittable.rs:39-57, 65-99; utility.rs:137-154) evaluated with numpy doubles (IEEE, unfused); slack is the
real-number expression that rtp_device.cu `any_slack` bounds from above in f32, here with each primitive's OWN extent and magnitude
as the scene constants, which makes the bound as tight as it can get. Adversarial inputs: rays through vertices and edges (which
lie on box faces), rays almost parallel to the triangle (|det| just above 1e-7), tiny direction components (huge 1/d),
unnormalised directions, far origins, axis-aligned triangles, grazing rays on spheres of any size."""
import numpy as np

U = 2.0 ** -53
K = 8.0 * U * 1e7


def slab_entry(bmin, bmax, o, inv, t_min):
    t0 = (bmin - o) * inv
    t1 = (bmax - o) * inv
    lo = np.fmax(np.fmax(np.fmax(t_min, np.fmin(t0[:, 0], t1[:, 0])), np.fmin(t0[:, 1], t1[:, 1])), np.fmin(t0[:, 2], t1[:, 2]))
    hi = np.fmin(np.fmin(np.fmin(np.inf, np.fmax(t0[:, 0], t1[:, 0])), np.fmax(t0[:, 1], t1[:, 1])), np.fmax(t0[:, 2], t1[:, 2]))
    return lo, hi


def triangle_cases(rng, n):
    scale = 10.0 ** rng.uniform(-3, 1, (n, 1))              # triangle size
    where = 10.0 ** rng.uniform(-1, 4, (n, 1)) * rng.choice([-1.0, 1.0], (n, 3))
    a = where + rng.normal(size=(n, 3)) * scale
    b = a + rng.normal(size=(n, 3)) * scale
    c = a + rng.normal(size=(n, 3)) * scale
    flat = rng.random(n) < 0.15                               # axis-aligned: zero-thickness boxes
    ax = rng.integers(0, 3, n)
    b[flat, ax[flat]] = a[flat, ax[flat]]
    c[flat, ax[flat]] = a[flat, ax[flat]]
    # target: inside, on an edge, on a vertex
    w = rng.dirichlet([1.0, 1.0, 1.0], n)
    mode = rng.integers(0, 4, n)
    w[mode == 1, 2] = 0.0
    w[mode == 2] = np.eye(3)[rng.integers(0, 3, (mode == 2).sum())]
    w /= w.sum(axis=1, keepdims=True)
    tgt = w[:, :1] * a + w[:, 1:2] * b + w[:, 2:] * c
    nrm = np.cross(b - a, c - a)
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-300)
    dist = 10.0 ** rng.uniform(-2, 3, (n, 1)) * scale
    tilt = rng.normal(size=(n, 3))
    # from head-on to almost in-plane
    g = 10.0 ** rng.uniform(-7, 0, (n, 1))
    tang = tilt - (tilt * nrm).sum(axis=1, keepdims=True) * nrm
    tang /= np.maximum(np.linalg.norm(tang, axis=1, keepdims=True), 1e-300)
    dirn = g * nrm * rng.choice([-1.0, 1.0], (n, 1)) + np.sqrt(np.maximum(0.0, 1.0 - g * g)) * tang
    o = tgt - dirn * dist
    d = (tgt - o)
    d *= 10.0 ** rng.uniform(-3, 3, (n, 1)) / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-300)
    tiny = rng.random(n) < 0.2                                # one tiny direction component
    k = rng.integers(0, 3, n)
    d[tiny, k[tiny]] = 10.0 ** rng.uniform(-14, -6, tiny.sum()) * rng.choice([-1.0, 1.0], tiny.sum())
    return a, b, c, o, d


def test_triangle_gap_never_exceeds_slack():
    rng = np.random.default_rng(2024)
    worst, passed, n_abnormal = 0.0, 0, 0
    for _ in range(8):
        n = 400000
        a, b, c, o, d = triangle_cases(rng, n)
        t_min = 1e-3
        with np.errstate(all="ignore"):
            ba, ca, pa = a - b, a - c, a - o
            det = (ba[:, 0] * ca[:, 1] * d[:, 2] + ba[:, 1] * ca[:, 2] * d[:, 0] + ba[:, 2] * ca[:, 0] * d[:, 1]
                   - ba[:, 0] * ca[:, 2] * d[:, 1] - ba[:, 1] * ca[:, 0] * d[:, 2] - ba[:, 2] * ca[:, 1] * d[:, 0])
            inv_det = 1.0 / det
            t = (pa[:, 0] * (ba[:, 1] * ca[:, 2] - ba[:, 2] * ca[:, 1]) + pa[:, 1] * (ba[:, 2] * ca[:, 0] - ba[:, 0] * ca[:, 2])
                 + pa[:, 2] * (ba[:, 0] * ca[:, 1] - ba[:, 1] * ca[:, 0])) * inv_det
            u = (pa[:, 0] * (ca[:, 1] * d[:, 2] - ca[:, 2] * d[:, 1]) + pa[:, 1] * (ca[:, 2] * d[:, 0] - ca[:, 0] * d[:, 2])
                 + pa[:, 2] * (ca[:, 0] * d[:, 1] - ca[:, 1] * d[:, 0])) * inv_det
            v = (pa[:, 0] * (ba[:, 2] * d[:, 1] - ba[:, 1] * d[:, 2]) + pa[:, 1] * (ba[:, 0] * d[:, 2] - ba[:, 2] * d[:, 0])
                 + pa[:, 2] * (ba[:, 1] * d[:, 0] - ba[:, 0] * d[:, 1])) * inv_det
            w = 1.0 - u - v
            inv = 1.0 / d
            bmin, bmax = np.minimum(np.minimum(a, b), c), np.maximum(np.maximum(a, b), c)
            lo, hi = slab_entry(bmin, bmax, o, inv, t_min)
            ok = (np.abs(det) >= 1e-7) & (t >= t_min) & (u >= 0.0) & (v >= 0.0) & (w >= 0.0) & np.isfinite(t) & (hi >= lo)
            # the walk's own eligibility: finite, 1e-15 <= |1/d| <= 1e15, |coordinates| <= 1e15, kappa <= 1/4, e_uv <= 0.01
            E = (bmax - bmin).max(axis=1)
            A = np.maximum(np.abs(bmin), np.abs(bmax)).max(axis=1)
            P = np.abs(o).max(axis=1) + A
            D = np.abs(d).max(axis=1)
            Imax = np.abs(inv).max(axis=1)
            kappa = K * 6.0 * D * E * E
            e_uv = K * (6.0 * P * E * D + 6.06 * D * E * E) + 3.0 * U
            elig = (kappa <= 0.25) & (e_uv <= 0.01) & (np.abs(inv).min(axis=1) >= 1e-15) & (Imax <= 1e15)
            eta = (4.0 * e_uv + 5.0 * U) * E
            s0 = 4.0 * ((eta + U * P) * Imax * (1.0 + U) + K * 6.0 * P * E * E * (4.0 / 3.0) + 3.0 * U * t_min)
            s1 = 4.0 * ((K * 6.0 * D * E * E + 5.0 * U) * (4.0 / 3.0))
            slack = s0 + s1 * np.abs(t)
            gap = lo - t
        m = ok & elig
        passed += int(m.sum())
        assert (gap[m] <= slack[m]).all(), float((gap[m] / slack[m]).max())
        abnormal = m & (gap > 0)
        n_abnormal += int(abnormal.sum())
        if abnormal.any():
            worst = max(worst, float((gap[abnormal] / slack[abnormal]).max()))
    assert passed > 500000 and n_abnormal > 10000, (passed, n_abnormal)  # the case the bound exists for is exercised
    assert worst < 0.25, worst  # the safety factor is not being eaten: the tightest case uses under a quarter of the bound


def test_sphere_gap_never_exceeds_slack():
    rng = np.random.default_rng(77)
    worst, passed, n_abnormal = 0.0, 0, 0
    for _ in range(6):
        n = 400000
        r = 10.0 ** rng.uniform(-3, 3, n)
        c = 10.0 ** rng.uniform(-1, 3, (n, 1)) * rng.normal(size=(n, 3))
        o = c + rng.normal(size=(n, 3)) * (r * 10.0 ** rng.uniform(-1, 3, n))[:, None]
        to_c = c - o
        side = rng.normal(size=(n, 3))
        side -= (side * to_c).sum(axis=1, keepdims=True) * to_c / np.maximum((to_c * to_c).sum(axis=1, keepdims=True), 1e-300)
        side /= np.maximum(np.linalg.norm(side, axis=1, keepdims=True), 1e-300)
        off = rng.choice([0.0, 0.3, 0.9, 1.0, 1.0 - 1e-15, 1.0 - 1e-12, 1.0 - 1e-8, 1.0 + 1e-15], n)  # from centred to grazing
        tgt = c + side * (r * off)[:, None]
        # a third of the rays go for the six points where the sphere touches its own box, from along that axis: t ~ slab entry
        pole = rng.random(n) < 0.33
        k = rng.integers(0, 3, n)
        sgn = rng.choice([-1.0, 1.0], n)
        e = np.zeros((n, 3))
        e[np.arange(n), k] = sgn
        jit = rng.normal(size=(n, 3)) * (r * 10.0 ** rng.uniform(-12, -2, n))[:, None]
        jit[np.arange(n), k] = 0.0
        tgt_p = c + e * r[:, None] + jit
        o_p = tgt_p + e * (r * 10.0 ** rng.uniform(-2, 3, n))[:, None] + rng.normal(size=(n, 3)) * (r * 10.0 ** rng.uniform(-9, 0, n))[:, None]
        tgt[pole], o[pole] = tgt_p[pole], o_p[pole]
        d = tgt - o
        d *= 10.0 ** rng.uniform(-3, 3, (n, 1)) / np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-300)
        t_min = 1e-3
        with np.errstate(all="ignore"):
            oc = o - c
            a = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            half_b = (d[:, 0] * oc[:, 0] + d[:, 1] * oc[:, 1]) + d[:, 2] * oc[:, 2]
            cc = ((oc[:, 0] * oc[:, 0] + oc[:, 1] * oc[:, 1]) + oc[:, 2] * oc[:, 2]) - r * r
            delta = half_b * half_b - a * cc
            sq = np.sqrt(delta)
            t1, t2 = (-half_b - sq) / a, (-half_b + sq) / a
            t = np.where(t1 >= t_min, t1, t2)
            inv = 1.0 / d
            bmin, bmax = c - r[:, None], c + r[:, None]
            lo, hi = slab_entry(bmin, bmax, o, inv, t_min)
            ok = (delta > 0.0) & (t >= t_min) & np.isfinite(t) & (hi >= lo)
            Imax = np.abs(inv).max(axis=1)
            elig = (np.abs(inv).min(axis=1) >= 1e-15) & (Imax <= 1e15)
            Po = np.abs(o).max(axis=1)
            C = np.maximum(np.abs(bmin), np.abs(bmax)).max(axis=1)
            S = 3.0 * (Po + C) ** 2 + r * r
            g = np.sqrt(40.0 * U * S)
            s0 = 4.0 * ((3.0 * g + U * (2.0 * (C + r) + Po + C)) * Imax + 3.0 * U * t_min)
            s1 = 4.0 * (8.0 * U)
            slack = s0 + s1 * np.abs(t)
            gap = lo - t
        m = ok & elig
        passed += int(m.sum())
        assert (gap[m] <= slack[m]).all(), float((gap[m] / slack[m]).max())
        abnormal = m & (gap > 0)
        n_abnormal += int(abnormal.sum())
        if abnormal.any():
            worst = max(worst, float((gap[abnormal] / slack[abnormal]).max()))
    assert passed > 500000 and n_abnormal > 10000, (passed, n_abnormal)
    assert worst < 0.25, worst
