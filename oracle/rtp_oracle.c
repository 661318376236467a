/*
 * rtp_oracle.c — CPU ORACLE. TEST INFRASTRUCTURE ONLY (see rtp_oracle.h; PARITY UNPINNED).
 *
 * Every function cites the lines of /root/reference/src it restates. Scalar expressions keep
 * the reference's association order; build with -ffp-contract=off (Rust never fuses a*b+c) and
 * without -ffast-math. f64::min/max are IEEE minNum/maxNum (inline f64_min/f64_max), `as u32`/`as u8`/`as
 * isize` are saturating with NaN -> 0.
 */
#define _GNU_SOURCE
#include "rtp_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_PI 3.14159265358979323846264338327950288   /* std::f64::consts::PI  */
#define ORC_TAU 6.28318530717958647692528676655900577  /* std::f64::consts::TAU */
#define RAY_EPSILON 1e-3 /* utility.rs:30 */
#define SMOL 1e-7        /* utility.rs:31 */

static __thread char g_err[256];
const char* orc_last_error(void) { return g_err; }
static int fail(int code, const char* msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}

/* ------------------------------------------------------------------ vectors ------------- */

typedef struct { double x, y, z; } v3;

static inline v3 v3_make(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 v3_from(const double* p) { v3 r = {p[0], p[1], p[2]}; return r; }
static inline v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_scale(double s, v3 a) { return v3_make(s * a.x, s * a.y, s * a.z); }
static inline v3 v3_mul(v3 a, v3 b) { return v3_make(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
/* nalgebra dot on a static 3-vector: a + b + c, left to right */
static inline double v3_dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline double v3_norm_squared(v3 a) { return 0.0 + v3_dot(a, a); }
static inline double v3_norm(v3 a) { return sqrt(v3_norm_squared(a)); }
static inline v3 v3_normalize(v3 a) {
    double n = v3_norm(a);
    return v3_make(a.x / n, a.y / n, a.z / n);
}
static inline v3 v3_cross(v3 a, v3 b) {
    return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

/* f64::min / f64::max = IEEE minNum / maxNum: a NaN operand is dropped. Written inline (not libm
 * fmin/fmax calls) so the CPU baseline compiles to minsd/maxsd + blend like rustc's lowering. */
static inline double f64_min(double a, double b) { double m = a < b ? a : b; return (b != b) ? a : m; }
static inline double f64_max(double a, double b) { double m = a > b ? a : b; return (b != b) ? a : m; }

/* `x as u32` */
static inline uint32_t sat_u32(double x) {
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 4294967295.0) return 4294967295u;
    return (uint32_t)x;
}
/* `x as u8` */
static inline uint8_t sat_u8(double x) {
    if (!(x == x)) return 0;
    if (x <= 0.0) return 0;
    if (x >= 255.0) return 255;
    return (uint8_t)x;
}
/* `x as isize` */
static inline int64_t sat_i64(double x) {
    if (!(x == x)) return 0;
    if (x <= -9223372036854775808.0) return INT64_MIN;
    if (x >= 9223372036854775808.0) return INT64_MAX;
    return (int64_t)x;
}
/* f64::clamp: NaN stays NaN */
static inline double clampd(double x, double lo, double hi) {
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}

/* ------------------------------------------------------------------ randomness ---------- */

/* Philox4x32-10 (Salmon et al., SC'11; Random123 philox.h). Replaces rand 0.8's StdRng
 * (randomness.rs:5), which the reference seeds from entropy (main.rs:52). */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct {
    uint32_t key[2];
    uint32_t ctr[4]; /* ctr[2] = block currently cached */
    uint32_t block[4];
    uint32_t k;      /* next draw index */
    uint32_t cached; /* block index held in `block`, or 0xFFFFFFFF */
} orc_rng;

static void rng_init(orc_rng* r, uint64_t seed, uint32_t lo, uint32_t hi, uint32_t stream) {
    r->key[0] = (uint32_t)seed;
    r->key[1] = (uint32_t)(seed >> 32);
    r->ctr[0] = lo; r->ctr[1] = hi; r->ctr[2] = 0; r->ctr[3] = stream;
    r->k = 0;
    r->cached = 0xFFFFFFFFu;
}

/* rng.gen::<f64>() — rand 0.8 `Standard`: 53 high bits of a u64, times 2^-53 */
static double rng_next(orc_rng* r) {
    uint32_t b = r->k >> 1, pair = r->k & 1u;
    if (b != r->cached) {
        r->ctr[2] = b;
        orc_philox4x32_10(r->ctr, r->key, r->block);
        r->cached = b;
    }
    r->k++;
    uint64_t u = ((uint64_t)r->block[2 * pair + 1] << 32) | r->block[2 * pair];
    return (double)(u >> 11) * 0x1.0p-53;
}

void orc_rng_draws(uint64_t seed, uint32_t lo, uint32_t hi, uint32_t stream, uint32_t first,
                   uint32_t n, double* out) {
    orc_rng r;
    rng_init(&r, seed, lo, hi, stream);
    r.k = first;
    for (uint32_t i = 0; i < n; ++i) out[i] = rng_next(&r);
}

/* randomness.rs:21-34 UnitDisk */
static void sample_unit_disk(orc_rng* r, double* x, double* y) {
    for (;;) {
        double vx = 2.0 * rng_next(r) - 1.0;
        double vy = 2.0 * rng_next(r) - 1.0;
        if (0.0 + (vx * vx + vy * vy) < 1.0) { *x = vx; *y = vy; return; }
    }
}
/* randomness.rs:39-53 UnitBall */
static v3 sample_unit_ball(orc_rng* r) {
    for (;;) {
        v3 v;
        v.x = 2.0 * rng_next(r) - 1.0;
        v.y = 2.0 * rng_next(r) - 1.0;
        v.z = 2.0 * rng_next(r) - 1.0;
        if (v3_norm_squared(v) < 1.0) return v;
    }
}
/* randomness.rs:58-73 UnitSphere (Marsaglia) */
static v3 sample_unit_sphere(orc_rng* r) {
    for (;;) {
        double vx = 2.0 * rng_next(r) - 1.0;
        double vy = 2.0 * rng_next(r) - 1.0;
        double s = 0.0 + (vx * vx + vy * vy);
        if (s < 1.0) {
            double n = 2.0 * sqrt(1.0 - s);
            return v3_make(vx * n, vy * n, 1.0 - 2.0 * s);
        }
    }
}

/* randomness.rs:91-105 noise::integer — wrapping isize arithmetic, arithmetic >> */
int64_t orc_noise_integer(int64_t x, int64_t y, int64_t z, int64_t seed) {
    const uint64_t A = 0x369E6D3B899E43CFull, B = 0x53F89E7FFDA3B07Dull, C = 0x3B13C1CA4937E629ull,
                   D = 0x577C2C6E4019D645ull, E = 60493ull, F = 19990303ull, G = 1376312589ull;
    uint64_t h = A * (uint64_t)x + B * (uint64_t)y + C * (uint64_t)z + D * (uint64_t)seed;
    h = (uint64_t)((int64_t)h >> 13) ^ h;
    h = h * (h * h * E + F) + G;
    return (int64_t)h;
}
/* randomness.rs:108-110 noise::real */
static double noise_real(int64_t x, int64_t y, int64_t z, int64_t seed) {
    return (double)orc_noise_integer(x, y, z, seed) / 9223372036854775808.0; /* isize::MAX as f64 */
}

/* ------------------------------------------------------------------ scene --------------- */

typedef struct {
    double bmin[3], bmax[3];
    uint32_t left, right; /* branch */
    uint32_t leaf;        /* RTP_MISS on branches */
    uint32_t _pad;
} orc_node;

typedef struct {
    uint32_t id;
    double bmin[3], bmax[3];
} orc_item;

typedef struct {
    uint8_t* rgba; /* owned copy for Image textures */
    rtp_texture t;
} orc_texture;

typedef struct {
    orc_node* nodes;
    uint32_t n_nodes, root, depth;
} orc_tree; /* bvh.rs:27-34 Bvh of a NESTED Hittable::Bvh (its leaves are a run of orc_scene.nested) */

struct orc_scene {
    uint32_t root_kind;
    uint32_t n_meshes, n_hittables, n_materials, n_textures;
    rtp_mesh* meshes; /* deep copies */
    rtp_hittable* hittables;
    rtp_hittable* nested;   /* items of nested List / Bvh hittables (rtp_scene_desc.nested) */
    uint32_t n_nested;
    orc_tree* nested_tree;  /* [n_nested]: the tree of the nested Bvh whose run starts at that index, nodes == NULL elsewhere */
    rtp_material* materials;
    orc_texture* textures;
    rtp_emit background;
    orc_node* nodes;
    uint32_t n_nodes, root, depth;
};

typedef struct {
    uint64_t rays, node_visits, triangle_tests, sphere_tests;
} orc_counters;

typedef struct {
    double t;
    v3 position, normal;
    double u, v;
    uint32_t material; /* the MaterialId half of Option<(Hit, MaterialId)> (hittable.rs:19) */
} orc_hit;

/* hittable.rs:27-34, 124-147 bounding boxes. Returns 0, or 1 where the reference panics: the box of a Bvh (hittable.rs:32) */
static int hittable_bbox(const orc_scene* s, const rtp_hittable* h, double bmin[3], double bmax[3]) {
    if (h->kind == RTP_HITTABLE_BVH) return 1;
    if (h->kind == RTP_HITTABLE_LIST) { /* hittable.rs:142-147 bounding_box_list: AABB::default() when empty, else a left fold of unions */
        const rtp_hittable* items = s->nested + h->mesh;
        for (int k = 0; k < 3; ++k) bmin[k] = bmax[k] = 0.0;
        for (uint32_t i = 0; i < h->triangle; ++i) {
            double lo[3], hi[3];
            if (hittable_bbox(s, &items[i], lo, hi)) return 1;
            for (int k = 0; k < 3; ++k) { /* utility.rs:130-135 AABB::union */
                bmin[k] = i ? f64_min(bmin[k], lo[k]) : lo[k];
                bmax[k] = i ? f64_max(bmax[k], hi[k]) : hi[k];
            }
        }
        return 0;
    }
    if (h->kind == RTP_HITTABLE_SPHERE) {
        for (int k = 0; k < 3; ++k) {
            bmin[k] = h->center[k] - h->radius;
            bmax[k] = h->center[k] + h->radius;
        }
    } else {
        const rtp_mesh* m = &s->meshes[h->mesh];
        const double* a = m->vertices[m->indices[h->triangle + 0]].position;
        const double* b = m->vertices[m->indices[h->triangle + 1]].position;
        const double* c = m->vertices[m->indices[h->triangle + 2]].position;
        for (int k = 0; k < 3; ++k) {
            bmin[k] = f64_min(f64_min(a[k], b[k]), c[k]);
            bmax[k] = f64_max(f64_max(a[k], b[k]), c[k]);
        }
    }
    return 0;
}

/* bvh.rs:60-64: order by 0.5*(min+max) on the axis. The reference uses sort_unstable_by, whose
 * order among equal keys is unspecified; the oracle fixes it as (key, LeafId). */
static int item_cmp(const void* pa, const void* pb, void* arg) {
    const orc_item* a = (const orc_item*)pa;
    const orc_item* b = (const orc_item*)pb;
    int axis = *(const int*)arg;
    double ka = 0.5 * (a->bmin[axis] + a->bmax[axis]);
    double kb = 0.5 * (b->bmin[axis] + b->bmax[axis]);
    if (ka < kb) return -1;
    if (ka > kb) return 1;
    return (a->id > b->id) - (a->id < b->id);
}

/* bvh.rs:36-56 make_bvh: children are pushed before their parent, so the subtree over n leaves fills the 2n-1 consecutive node
 * slots starting at `base` (left subtree, right subtree, then the node itself): the numbering of the reference's sequential
 * `nodes.push`. Because the slots are known up front, the two halves of a big range are built on two threads (test
 * infrastructure for the 10 M-leaf scene; the result does not depend on it). */
typedef struct {
    orc_scene* s; orc_item* items; size_t n; int axis; uint32_t depth, base, spawn;
} bvh_job;
static uint32_t make_bvh_at(orc_scene* s, orc_item* items, size_t n, int axis, uint32_t depth, uint32_t base, uint32_t spawn);
static void* bvh_job_run(void* arg) {
    bvh_job* j = (bvh_job*)arg;
    make_bvh_at(j->s, j->items, j->n, j->axis, j->depth, j->base, j->spawn);
    return NULL;
}
static uint32_t make_bvh_at(orc_scene* s, orc_item* items, size_t n, int axis, uint32_t depth, uint32_t base, uint32_t spawn) {
    uint32_t seen = __atomic_load_n(&s->depth, __ATOMIC_RELAXED);
    while (depth > seen && !__atomic_compare_exchange_n(&s->depth, &seen, depth, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    const uint32_t self = base + (uint32_t)(2 * n - 2);
    if (n == 1) {
        orc_node* nd = &s->nodes[self];
        memcpy(nd->bmin, items[0].bmin, sizeof nd->bmin);
        memcpy(nd->bmax, items[0].bmax, sizeof nd->bmax);
        nd->left = nd->right = RTP_MISS;
        nd->leaf = items[0].id;
        return self;
    }
    qsort_r(items, n, sizeof *items, item_cmp, &axis);
    size_t half = n / 2; /* bvh.rs:66 split_at_mut(len/2) */
    const uint32_t left = base + (uint32_t)(2 * half - 2), right = self - 1;
    if (spawn > 0 && n > 65536) {
        bvh_job job = {s, items, half, (axis + 1) % 3, depth + 1, base, spawn - 1};
        pthread_t th;
        int threaded = pthread_create(&th, NULL, bvh_job_run, &job) == 0;
        if (!threaded) bvh_job_run(&job);
        make_bvh_at(s, items + half, n - half, (axis + 1) % 3, depth + 1, left + 1, spawn - 1);
        if (threaded) pthread_join(th, NULL);
    } else {
        make_bvh_at(s, items, half, (axis + 1) % 3, depth + 1, base, 0);
        make_bvh_at(s, items + half, n - half, (axis + 1) % 3, depth + 1, left + 1, 0);
    }
    orc_node* nd = &s->nodes[self];
    const orc_node* l = &s->nodes[left];
    const orc_node* r = &s->nodes[right];
    for (int k = 0; k < 3; ++k) { /* utility.rs:130-135 AABB::union */
        nd->bmin[k] = f64_min(l->bmin[k], r->bmin[k]);
        nd->bmax[k] = f64_max(l->bmax[k], r->bmax[k]);
    }
    nd->left = left;
    nd->right = right;
    nd->leaf = RTP_MISS;
    return self;
}
static uint32_t make_bvh(orc_scene* s, orc_item* items, size_t n, int axis, uint32_t depth) {
    uint32_t root = make_bvh_at(s, items, n, axis, depth, 0, 5);
    s->n_nodes = (uint32_t)(2 * n - 1);
    return root;
}

/* bvh.rs:70-91 Bvh::new over `n` hittables: boxes, then make_bvh. strict: NaN centroids and Bvh items are errors (the reference panics) */
static int build_tree(orc_scene* s, const rtp_hittable* leaves, uint32_t n, int strict, orc_tree* out) {
    orc_item* items = (orc_item*)malloc(sizeof(orc_item) * (n ? n : 1));
    for (uint32_t i = 0; i < n; ++i) {
        items[i].id = i;
        if (hittable_bbox(s, &leaves[i], items[i].bmin, items[i].bmax)) {
            free(items);
            return fail(RTP_ERR_INVALID, "bounding box of a Bvh: \"Do not take the bounding box of a Bvh\" (hittable.rs:32)");
        }
        for (int k = 0; k < 3; ++k) {
            double key = 0.5 * (items[i].bmin[k] + items[i].bmax[k]);
            if (key != key && strict) { /* partial_cmp().unwrap() panics, bvh.rs:63 */
                free(items);
                return fail(RTP_ERR_INVALID, "NaN bounding-box centroid");
            }
        }
    }
    orc_node* keep_nodes = s->nodes; uint32_t keep_n = s->n_nodes, keep_depth = s->depth;
    s->nodes = (orc_node*)calloc((size_t)2 * (n ? n : 1), sizeof(orc_node));
    s->n_nodes = 0; s->depth = 0;
    out->root = make_bvh(s, items, n, 0, 1);
    out->nodes = s->nodes; out->n_nodes = s->n_nodes; out->depth = s->depth;
    s->nodes = keep_nodes; s->n_nodes = keep_n; s->depth = keep_depth;
    free(items);
    return RTP_OK;
}

void orc_scene_destroy(orc_scene* s) {
    if (!s) return;
    if (s->meshes) {
        for (uint32_t i = 0; i < s->n_meshes; ++i) {
            free((void*)s->meshes[i].vertices);
            free((void*)s->meshes[i].indices);
        }
    }
    if (s->textures) for (uint32_t i = 0; i < s->n_textures; ++i) free(s->textures[i].rgba);
    if (s->nested_tree) for (uint32_t i = 0; i < s->n_nested; ++i) free(s->nested_tree[i].nodes);
    free(s->nested_tree); free(s->nested);
    free(s->meshes); free(s->hittables); free(s->materials); free(s->textures); free(s->nodes);
    free(s);
}

static int check_emit(const rtp_emit* e, uint32_t n_textures) {
    if (e->kind > RTP_EMIT_SKY_SPHERE) return 0;
    if (e->kind == RTP_EMIT_SKY_SPHERE && e->texture >= n_textures) return 0;
    return 1;
}

int orc_scene_create(const rtp_scene_desc* d, orc_scene** out) {
    if (!d || !out) return fail(RTP_ERR_INVALID, "null argument");
    if (d->abi_version != RTP_ABI_VERSION) return fail(RTP_ERR_INVALID, "abi version mismatch");
    if (d->root_kind > RTP_ROOT_LIST) return fail(RTP_ERR_INVALID, "bad root kind");
    if (d->root_kind == RTP_ROOT_BVH && d->n_hittables == 0)
        return fail(RTP_ERR_INVALID, "Bvh::new on an empty list is unreachable!() in the reference (bvh.rs:40)");
    orc_scene* s = (orc_scene*)calloc(1, sizeof *s);
    if (!s) return fail(RTP_ERR_NOMEM, "oom");
    s->root_kind = d->root_kind;
    s->n_meshes = d->n_meshes; s->n_hittables = d->n_hittables;
    s->n_materials = d->n_materials; s->n_textures = d->n_textures;
    s->background = d->background;
    s->meshes = (rtp_mesh*)calloc(d->n_meshes ? d->n_meshes : 1, sizeof(rtp_mesh));
    s->hittables = (rtp_hittable*)calloc(d->n_hittables ? d->n_hittables : 1, sizeof(rtp_hittable));
    s->materials = (rtp_material*)calloc(d->n_materials ? d->n_materials : 1, sizeof(rtp_material));
    s->textures = (orc_texture*)calloc(d->n_textures ? d->n_textures : 1, sizeof(orc_texture));
    for (uint32_t i = 0; i < d->n_meshes; ++i) {
        const rtp_mesh* m = &d->meshes[i];
        rtp_mesh* c = &s->meshes[i];
        *c = *m;
        rtp_vertex* v = (rtp_vertex*)malloc(sizeof(rtp_vertex) * (m->n_vertices ? m->n_vertices : 1));
        uint32_t* ix = (uint32_t*)malloc(sizeof(uint32_t) * (m->n_indices ? m->n_indices : 1));
        memcpy(v, m->vertices, sizeof(rtp_vertex) * m->n_vertices);
        memcpy(ix, m->indices, sizeof(uint32_t) * m->n_indices);
        c->vertices = v; c->indices = ix;
        for (uint32_t k = 0; k < m->n_indices; ++k)
            if (ix[k] >= m->n_vertices) { orc_scene_destroy(s); return fail(RTP_ERR_INVALID, "vertex index out of range"); }
        if (m->material >= d->n_materials) { orc_scene_destroy(s); return fail(RTP_ERR_INVALID, "mesh material out of range"); }
    }
    memcpy(s->materials, d->materials, sizeof(rtp_material) * d->n_materials);
    for (uint32_t i = 0; i < d->n_textures; ++i) {
        s->textures[i].t = d->textures[i];
        const rtp_texture* t = &d->textures[i];
        if (t->kind > RTP_TEXTURE_PERLIN) { orc_scene_destroy(s); return fail(RTP_ERR_INVALID, "bad texture kind"); }
        if (t->kind == RTP_TEXTURE_IMAGE) {
            size_t n = (size_t)t->width * t->height * 4;
            if (!t->rgba || n == 0) { orc_scene_destroy(s); return fail(RTP_ERR_INVALID, "empty image texture"); }
            s->textures[i].rgba = (uint8_t*)malloc(n);
            memcpy(s->textures[i].rgba, t->rgba, n);
            s->textures[i].t.rgba = s->textures[i].rgba;
        }
        if (t->kind == RTP_TEXTURE_CHECKER && (t->odd >= d->n_textures || t->even >= d->n_textures)) {
            orc_scene_destroy(s); return fail(RTP_ERR_INVALID, "checker texture id out of range");
        }
    }
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        const rtp_material* m = &d->materials[i];
        if (m->scatter > RTP_SCATTER_DIELECTRIC || m->absorb > RTP_ABSORB_ALBEDO_MAP || !check_emit(&m->emit, d->n_textures) ||
            (m->absorb == RTP_ABSORB_ALBEDO_MAP && m->absorb_texture >= d->n_textures)) {
            orc_scene_destroy(s); return fail(RTP_ERR_INVALID, "bad material");
        }
    }
    if (!check_emit(&d->background, d->n_textures)) { orc_scene_destroy(s); return fail(RTP_ERR_INVALID, "bad background"); }
    memcpy(s->hittables, d->hittables, sizeof(rtp_hittable) * d->n_hittables);
    if (d->n_nested && !d->nested) { orc_scene_destroy(s); return fail(RTP_ERR_INVALID, "null nested table"); }
    s->n_nested = d->n_nested;
    s->nested = (rtp_hittable*)calloc(d->n_nested ? d->n_nested : 1, sizeof(rtp_hittable));
    s->nested_tree = (orc_tree*)calloc(d->n_nested ? d->n_nested : 1, sizeof(orc_tree));
    if (d->n_nested) memcpy(s->nested, d->nested, sizeof(rtp_hittable) * d->n_nested);
    for (uint32_t pass = 0; pass < 2; ++pass) {
        const rtp_hittable* tab = pass ? s->nested : s->hittables;
        uint32_t n = pass ? s->n_nested : s->n_hittables;
        for (uint32_t i = 0; i < n; ++i) {
            const rtp_hittable* h = &tab[i];
            int ok = 1;
            if (h->kind == RTP_HITTABLE_SPHERE) ok = h->material < d->n_materials;
            else if (h->kind == RTP_HITTABLE_TRIANGLE)
                ok = h->mesh < d->n_meshes && (uint64_t)h->triangle + 3 <= d->meshes[h->mesh].n_indices;
            else if (h->kind == RTP_HITTABLE_LIST || h->kind == RTP_HITTABLE_BVH)
                /* a run of `nested`; a nested container's run lies entirely before the container itself, so nesting cannot cycle */
                ok = (uint64_t)h->mesh + h->triangle <= (pass ? i : s->n_nested);
            else ok = 0;
            if (!ok) { orc_scene_destroy(s); return fail(RTP_ERR_INVALID, "bad hittable"); }
            if (h->kind == RTP_HITTABLE_BVH && h->triangle == 0) { orc_scene_destroy(s); return fail(RTP_ERR_INVALID, "Bvh::new on an empty list is unreachable!() in the reference (bvh.rs:40)"); }
        }
    }
    /* nested Bvh::new (bvh.rs:70-91), inner ones first (their runs come first) */
    for (uint32_t pass = 0; pass < 2; ++pass) {
        const rtp_hittable* tab = pass ? s->hittables : s->nested;
        uint32_t n = pass ? s->n_hittables : s->n_nested;
        for (uint32_t i = 0; i < n; ++i) {
            const rtp_hittable* h = &tab[i];
            if (h->kind != RTP_HITTABLE_BVH || s->nested_tree[h->mesh].nodes) continue;
            int rc = build_tree(s, s->nested + h->mesh, h->triangle, 1, &s->nested_tree[h->mesh]);
            if (rc) { orc_scene_destroy(s); return rc; }
        }
    }
    /* bvh.rs:70-91 Bvh::new (also built for List roots of primitives so the differential test has leaf boxes) */
    if (d->n_hittables) {
        int any_bvh = 0;
        for (uint32_t i = 0; i < d->n_hittables; ++i) { double lo[3], hi[3]; any_bvh |= hittable_bbox(s, &s->hittables[i], lo, hi); }
        if (d->root_kind == RTP_ROOT_BVH || !any_bvh) {
            orc_tree t;
            int rc = build_tree(s, s->hittables, d->n_hittables, d->root_kind == RTP_ROOT_BVH, &t);
            if (rc) { orc_scene_destroy(s); return rc; }
            s->nodes = t.nodes; s->n_nodes = t.n_nodes; s->root = t.root; s->depth = t.depth;
        }
    }
    *out = s;
    return RTP_OK;
}

int orc_scene_get_info(const orc_scene* s, rtp_scene_info* info) {
    if (!s || !info) return fail(RTP_ERR_INVALID, "null argument");
    info->n_leaves = s->n_hittables;
    info->n_nodes = s->n_nodes;
    info->depth = s->depth;
    info->root_kind = s->root_kind;
    info->device_bytes = 0;
    return RTP_OK;
}

static void leaf_order_rec(const orc_scene* s, uint32_t node, uint32_t* out, size_t* n) {
    const orc_node* nd = &s->nodes[node];
    if (nd->leaf != RTP_MISS) { out[(*n)++] = nd->leaf; return; }
    leaf_order_rec(s, nd->left, out, n);
    leaf_order_rec(s, nd->right, out, n);
}
int orc_scene_leaf_order(const orc_scene* s, uint32_t* out, size_t cap) {
    if (!s || !out || cap < s->n_hittables) return fail(RTP_ERR_INVALID, "bad argument");
    size_t n = 0;
    if (s->n_hittables && s->nodes) leaf_order_rec(s, s->root, out, &n);
    return RTP_OK;
}
int orc_scene_node(const orc_scene* s, uint32_t node, double aabb[6], uint32_t lrl[3]) {
    if (!s || node >= s->n_nodes) return fail(RTP_ERR_INVALID, "bad node");
    memcpy(aabb, s->nodes[node].bmin, 24);
    memcpy(aabb + 3, s->nodes[node].bmax, 24);
    lrl[0] = s->nodes[node].left; lrl[1] = s->nodes[node].right; lrl[2] = s->nodes[node].leaf;
    return RTP_OK;
}
uint32_t orc_scene_root(const orc_scene* s) { return s->root; }

/* ------------------------------------------------------------------ intersection -------- */

typedef struct {
    v3 o, d, inv;
    double t_min, t_max;
} orc_xray; /* utility.rs:61-64 RayExpanded */

/* utility.rs:137-154 AABB::collide */
static inline int aabb_collide(const double bmin[3], const double bmax[3], const orc_xray* r) {
    double t0x = (bmin[0] - r->o.x) * r->inv.x, t0y = (bmin[1] - r->o.y) * r->inv.y, t0z = (bmin[2] - r->o.z) * r->inv.z;
    double t1x = (bmax[0] - r->o.x) * r->inv.x, t1y = (bmax[1] - r->o.y) * r->inv.y, t1z = (bmax[2] - r->o.z) * r->inv.z;
    double t_min = f64_max(f64_max(f64_max(r->t_min, f64_min(t0x, t1x)), f64_min(t0y, t1y)), f64_min(t0z, t1z));
    double t_max = f64_min(f64_min(f64_min(r->t_max, f64_max(t0x, t1x)), f64_max(t0y, t1y)), f64_max(t0z, t1z));
    return t_max >= t_min;
}

static orc_xray expand(const rtp_ray* ray) { /* utility.rs:71-77 Ray::expand */
    orc_xray r;
    r.o = v3_from(ray->origin);
    r.d = v3_from(ray->direction);
    r.inv = v3_make(1.0 / r.d.x, 1.0 / r.d.y, 1.0 / r.d.z);
    r.t_min = ray->t_min;
    r.t_max = ray->t_max;
    return r;
}

int orc_aabb_collide(const double bmin[3], const double bmax[3], const rtp_ray* ray) {
    orc_xray r = expand(ray);
    return aabb_collide(bmin, bmax, &r);
}

/* hittable.rs:39-63 hit_sphere */
static int hit_sphere(const rtp_hittable* h, const orc_xray* r, orc_hit* hit) {
    v3 center = v3_from(h->center);
    v3 to_center = v3_sub(r->o, center);
    double a = v3_norm_squared(r->d);
    double half_b = v3_dot(r->d, to_center);
    double c = v3_norm_squared(to_center) - h->radius * h->radius;
    double delta = half_b * half_b - a * c;
    if (delta <= 0.0) return 0;
    double sqrt_delta = sqrt(delta);
    double t = (-half_b - sqrt_delta) / a;
    if (t < r->t_min || t > r->t_max) {
        t = (-half_b + sqrt_delta) / a;
        if (t < r->t_min || t > r->t_max) return 0;
    }
    v3 position = v3_add(r->o, v3_scale(t, r->d)); /* utility.rs:67-69 Ray::at */
    v3 normal = v3_normalize(v3_sub(position, center));
    hit->t = t;
    hit->position = position;
    hit->normal = normal;
    hit->u = 0.5 - atan2(normal.z, normal.x) / ORC_TAU;
    hit->v = asin(normal.y) / ORC_PI + 0.5;
    return 1;
}

/* hittable.rs:65-108 hit_triangle (+ mesh.rs:25-30 get_triangle) */
static int hit_triangle(const orc_scene* s, const rtp_hittable* h, const orc_xray* r, orc_hit* hit) {
    const rtp_mesh* m = &s->meshes[h->mesh];
    rtp_vertex va = m->vertices[m->indices[h->triangle + 0]];
    rtp_vertex vb = m->vertices[m->indices[h->triangle + 1]];
    rtp_vertex vc = m->vertices[m->indices[h->triangle + 2]];
    v3 a = v3_from(va.position), b = v3_from(vb.position), c = v3_from(vc.position);
    v3 ba = v3_sub(a, b);
    v3 ca = v3_sub(a, c);
    v3 pa = v3_sub(a, r->o);
    v3 d = r->d;

    double det = ba.x * ca.y * d.z + ba.y * ca.z * d.x + ba.z * ca.x * d.y
               - ba.x * ca.z * d.y - ba.y * ca.x * d.z - ba.z * ca.y * d.x;
    if (fabs(det) < SMOL) return 0;
    double inv_det = 1.0 / det;

    double t = (pa.x * (ba.y * ca.z - ba.z * ca.y)
              + pa.y * (ba.z * ca.x - ba.x * ca.z)
              + pa.z * (ba.x * ca.y - ba.y * ca.x)) * inv_det;
    double u = (pa.x * (ca.y * d.z - ca.z * d.y)
              + pa.y * (ca.z * d.x - ca.x * d.z)
              + pa.z * (ca.x * d.y - ca.y * d.x)) * inv_det;
    double v = (pa.x * (ba.z * d.y - ba.y * d.z)
              + pa.y * (ba.x * d.z - ba.z * d.x)
              + pa.z * (ba.y * d.x - ba.x * d.y)) * inv_det;
    double w = 1.0 - u - v;

    if (t < r->t_min || t > r->t_max || u < 0.0 || v < 0.0 || w < 0.0) return 0;

    hit->t = t;
    hit->position = v3_add(r->o, v3_scale(t, d));
    hit->normal = v3_add(v3_add(v3_scale(w, v3_from(va.normal)), v3_scale(u, v3_from(vb.normal))),
                         v3_scale(v, v3_from(vc.normal)));
    hit->u = (w * va.uv[0] + u * vb.uv[0]) + v * vc.uv[0];
    hit->v = (w * va.uv[1] + u * vb.uv[1]) + v * vc.uv[1];
    return 1;
}

static inline uint32_t hittable_material(const orc_scene* s, const rtp_hittable* h) {
    return h->kind == RTP_HITTABLE_SPHERE ? h->material : s->meshes[h->mesh].material;
}

static int hit_node(const orc_scene* s, const orc_node* nodes, const rtp_hittable* leaves, const orc_xray* ray, uint32_t node, orc_hit* hit, uint32_t* leaf, orc_counters* c);

/* hittable.rs:18-25 Hittable::hit */
static int hittable_hit(const orc_scene* s, const rtp_hittable* h, const orc_xray* r, orc_hit* hit, orc_counters* c) {
    switch (h->kind) {
    case RTP_HITTABLE_SPHERE:
        c->sphere_tests++;
        if (!hit_sphere(h, r, hit)) return 0;
        hit->material = h->material;
        return 1;
    case RTP_HITTABLE_TRIANGLE:
        c->triangle_tests++;
        if (!hit_triangle(s, h, r, hit)) return 0;
        hit->material = s->meshes[h->mesh].material; /* hittable.rs:107 */
        return 1;
    case RTP_HITTABLE_LIST: { /* hittable.rs:110-120 hit_list */
        const rtp_hittable* items = s->nested + h->mesh;
        int found = 0;
        orc_xray ray = *r;
        orc_hit nh;
        for (uint32_t i = 0; i < h->triangle; ++i) {
            if (hittable_hit(s, &items[i], &ray, &nh, c)) {
                ray.t_max = nh.t;
                *hit = nh; found = 1;
            }
        }
        return found;
    }
    default: { /* RTP_HITTABLE_BVH: bvh.rs:121-124 Bvh::hit expands the ray again */
        const orc_tree* t = &s->nested_tree[h->mesh];
        orc_xray ray = *r;
        ray.inv = v3_make(1.0 / ray.d.x, 1.0 / ray.d.y, 1.0 / ray.d.z);
        uint32_t leaf;
        return hit_node(s, t->nodes, s->nested + h->mesh, &ray, t->root, hit, &leaf, c);
    }
    }
}

/* bvh.rs:93-119 Bvh::hit_node */
static int hit_node(const orc_scene* s, const orc_node* nodes, const rtp_hittable* leaves, const orc_xray* ray, uint32_t node, orc_hit* hit, uint32_t* leaf, orc_counters* c) {
    const orc_node* nd = &nodes[node];
    c->node_visits++;
    if (nd->leaf != RTP_MISS) {
        if (aabb_collide(nd->bmin, nd->bmax, ray)) {
            if (hittable_hit(s, &leaves[nd->leaf], ray, hit, c)) { *leaf = nd->leaf; return 1; }
        }
        return 0;
    }
    if (!aabb_collide(nd->bmin, nd->bmax, ray)) return 0;
    int found = 0;
    orc_xray r = *ray; /* bvh.rs:105 clone */
    orc_hit h;
    uint32_t l;
    if (hit_node(s, nodes, leaves, &r, nd->left, &h, &l, c)) {
        r.t_max = h.t; /* bvh.rs:107 */
        *hit = h; *leaf = l; found = 1;
    }
    if (hit_node(s, nodes, leaves, &r, nd->right, &h, &l, c)) {
        *hit = h; *leaf = l; found = 1; /* bvh.rs:111 replace */
    }
    return found;
}

/* hittable.rs:110-120 hit_list over the root list */
static int hit_list(const orc_scene* s, const orc_xray* ray, orc_hit* hit, uint32_t* leaf, orc_counters* c) {
    int found = 0;
    orc_xray r = *ray;
    orc_hit h;
    for (uint32_t i = 0; i < s->n_hittables; ++i) {
        if (hittable_hit(s, &s->hittables[i], &r, &h, c)) {
            r.t_max = h.t;
            *hit = h; *leaf = i; found = 1;
        }
    }
    return found;
}

/* `scene.hit(ray, scene_data)` on the root (render.rs:105,133). *leaf = index of the winning item of the root container */
static int scene_hit(const orc_scene* s, const rtp_ray* ray, int force_list, orc_hit* hit, uint32_t* leaf, orc_counters* c) {
    c->rays++;
    if (s->root_kind == RTP_ROOT_LIST || force_list) {
        orc_xray r; /* List does not expand the ray; inv is unused */
        r.o = v3_from(ray->origin); r.d = v3_from(ray->direction);
        r.inv = v3_make(0, 0, 0);
        r.t_min = ray->t_min; r.t_max = ray->t_max;
        return hit_list(s, &r, hit, leaf, c);
    }
    orc_xray r = expand(ray); /* bvh.rs:121-124 */
    return hit_node(s, s->nodes, s->hittables, &r, s->root, hit, leaf, c);
}

/* ------------------------------------------------------------------ textures ------------ */

static v3 texture_sample(const orc_scene* s, uint32_t tid, const orc_hit* hit);

/* texture.rs:40-49 sample_image */
static v3 sample_image(const orc_texture* tx, const orc_hit* hit) {
    double w = (double)tx->t.width, h = (double)tx->t.height;
    uint32_t i = sat_u32(clampd(hit->u * w, 0.0, w - 1.0));
    uint32_t j = sat_u32(clampd(hit->v * h, 0.0, h - 1.0));
    const uint8_t* p = tx->rgba + 4 * ((size_t)i + (size_t)j * tx->t.width); /* image.rs:31-33 */
    return v3_make((double)p[0] / 255.0, (double)p[1] / 255.0, (double)p[2] / 255.0);
}

/* texture.rs:70-76 grad_dot */
static double grad_dot(v3 p, int64_t cx, int64_t cy, int64_t cz, int64_t seed) {
    v3 grad = v3_make(noise_real(cx, cy, cz, seed + 1), noise_real(cx, cy, cz, seed + 2), noise_real(cx, cy, cz, seed + 3));
    return v3_dot(v3_sub(p, v3_make((double)cx, (double)cy, (double)cz)), grad);
}
static double mix(double a, double b, double t) { return (b - a) * t + a; } /* texture.rs:78-80 */
static double smootherstep(double t) { return (t * (t * 6.0 - 15.0) + 10.0) * t * t * t; } /* texture.rs:103 */

/* texture.rs:20-36 Texture::sample */
static v3 texture_sample(const orc_scene* s, uint32_t tid, const orc_hit* hit) {
    const orc_texture* tx = &s->textures[tid];
    switch (tx->t.kind) {
    case RTP_TEXTURE_MISSING: return v3_make(0, 0, 0);
    case RTP_TEXTURE_DEBUG_UVS: return v3_make(hit->u, hit->v, 0.0);
    case RTP_TEXTURE_SOLID: return v3_from(tx->t.rgb);
    case RTP_TEXTURE_IMAGE: return sample_image(tx, hit);
    case RTP_TEXTURE_CHECKER: { /* texture.rs:51-60 */
        v3 p = hit->position;
        if (fmod(floor(p.x) + floor(p.y) + floor(p.z), 2.0) == 0.0) return texture_sample(s, tx->t.even, hit);
        return texture_sample(s, tx->t.odd, hit);
    }
    case RTP_TEXTURE_NOISE: { /* texture.rs:62-68 */
        v3 p = hit->position;
        double x = noise_real(sat_i64(floor(p.x)), sat_i64(floor(p.y)), sat_i64(floor(p.z)), tx->t.seed);
        x = 0.5 * x + 0.5;
        return v3_make(x, x, x);
    }
    case RTP_TEXTURE_PERLIN: { /* texture.rs:82-119 */
        v3 p = hit->position;
        v3 fp = v3_make(floor(p.x), floor(p.y), floor(p.z));
        int64_t flx = sat_i64(fp.x), fly = sat_i64(fp.y), flz = sat_i64(fp.z);
        int64_t clx = flx + 1, cly = fly + 1, clz = flz + 1;
        int64_t seed = tx->t.seed;
        double k1 = grad_dot(p, flx, fly, flz, seed), k2 = grad_dot(p, clx, fly, flz, seed);
        double k3 = grad_dot(p, flx, cly, flz, seed), k4 = grad_dot(p, clx, cly, flz, seed);
        double k5 = grad_dot(p, flx, fly, clz, seed), k6 = grad_dot(p, clx, fly, clz, seed);
        double k7 = grad_dot(p, flx, cly, clz, seed), k8 = grad_dot(p, clx, cly, clz, seed);
        v3 t = v3_sub(p, fp);
        t = v3_make(smootherstep(t.x), smootherstep(t.y), smootherstep(t.z));
        double k12 = mix(k1, k2, t.x), k34 = mix(k3, k4, t.x), k56 = mix(k5, k6, t.x), k78 = mix(k7, k8, t.x);
        double k1234 = mix(k12, k34, t.y), k5678 = mix(k56, k78, t.y);
        double k = mix(k1234, k5678, t.z);
        double x = 0.5 * k + 0.5;
        return v3_make(x, x, x);
    }
    }
    return v3_make(0, 0, 0);
}

int orc_texture_sample(const orc_scene* s, uint32_t texture, const double position[3], const double uv[2], double rgb[3]) {
    if (!s || texture >= s->n_textures) return fail(RTP_ERR_INVALID, "bad texture id");
    orc_hit h;
    h.t = 0; h.position = v3_from(position); h.normal = v3_make(0, 0, 0); h.u = uv[0]; h.v = uv[1];
    v3 c = texture_sample(s, texture, &h);
    rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
    return RTP_OK;
}

/* ------------------------------------------------------------------ materials ----------- */

/* utility.rs:106-108 reflect */
static v3 reflect(v3 incident, v3 normal) {
    double k = 2.0 * v3_dot(incident, normal);
    return v3_sub(incident, v3_scale(k, normal));
}
/* utility.rs:110-119 refract */
static int refract(v3 incident, v3 normal, double eta, v3* out) {
    double cos_theta = v3_dot(normal, incident);
    double k = 1.0 - eta * eta * (1.0 - cos_theta * cos_theta);
    if (k < 0.0) return 0;
    *out = v3_sub(v3_scale(eta, incident), v3_scale(eta * cos_theta + sqrt(k), normal));
    return 1;
}

/* material.rs:49-60 Emit::evaluate */
static v3 emit_evaluate(const orc_scene* s, const rtp_emit* e, v3 dir, const orc_hit* hit) {
    switch (e->kind) {
    case RTP_EMIT_NONE: return v3_make(0, 0, 0);
    case RTP_EMIT_COLOR: return v3_from(e->rgb);
    case RTP_EMIT_DEBUG_NORMALS: return hit->normal;
    case RTP_EMIT_SKY_GRADIENT: {
        double t = 0.5 * (dir.y / v3_norm(dir) + 1.0);
        double a = 1.0 - t;
        return v3_make(a * 1.0 + t * 0.5, a * 1.0 + t * 0.7, a * 1.0 + t * 1.0);
    }
    case RTP_EMIT_SKY_SPHERE: return texture_sample(s, e->texture, hit);
    }
    return v3_make(0, 0, 0);
}

/* material.rs:74-81 Absorb::evaluate */
static v3 absorb_evaluate(const orc_scene* s, const rtp_material* m, const orc_hit* hit) {
    switch (m->absorb) {
    case RTP_ABSORB_BLACKBODY: return v3_make(0, 0, 0);
    case RTP_ABSORB_WHITEBODY: return v3_make(1, 1, 1);
    case RTP_ABSORB_ALBEDO: return v3_from(m->absorb_rgb);
    case RTP_ABSORB_ALBEDO_MAP: return texture_sample(s, m->absorb_texture, hit);
    }
    return v3_make(0, 0, 0);
}

/* material.rs:27-34 Scatter::evaluate → material.rs:115-180. Returns 1 and fills `out` on scatter. */
static int scatter_evaluate(const rtp_material* m, v3 dir, const orc_hit* hit, orc_rng* rng, rtp_ray* out) {
    v3 sd;
    switch (m->scatter) {
    case RTP_SCATTER_NONE: return 0;
    case RTP_SCATTER_LAMBERT: /* material.rs:115-130 */
        if (v3_dot(hit->normal, dir) > 0.0) return 0;
        sd = v3_normalize(v3_add(hit->normal, sample_unit_sphere(rng)));
        break;
    case RTP_SCATTER_METAL: { /* material.rs:132-152 */
        if (v3_dot(hit->normal, dir) > 0.0) return 0;
        v3 refl = reflect(dir, hit->normal);
        v3 fz = v3_scale(m->scatter_param, sample_unit_ball(rng));
        sd = v3_normalize(v3_add(refl, fz));
        if (v3_dot(hit->normal, sd) < 0.0) return 0;
        break;
    }
    case RTP_SCATTER_DIELECTRIC: { /* material.rs:154-180 */
        double eta; v3 normal;
        if (v3_dot(hit->normal, dir) > 0.0) { eta = m->scatter_param; normal = v3_neg(hit->normal); }
        else { eta = 1.0 / m->scatter_param; normal = hit->normal; }
        double q = (1.0 - eta) / (1.0 + eta);
        double r0 = q * q;                       /* powi(2) */
        double x = 1.0 + v3_dot(normal, dir);
        double x2 = x * x;
        double x5 = x * (x2 * x2);               /* powi(5) */
        double reflectance = r0 + (1.0 - r0) * x5;
        if (rng_next(rng) < reflectance) sd = reflect(dir, normal);
        else if (!refract(dir, normal, eta, &sd)) sd = reflect(dir, normal);
        break;
    }
    default: return 0;
    }
    out->origin[0] = hit->position.x; out->origin[1] = hit->position.y; out->origin[2] = hit->position.z;
    out->direction[0] = sd.x; out->direction[1] = sd.y; out->direction[2] = sd.z;
    out->t_min = RAY_EPSILON;
    out->t_max = INFINITY;
    return 1;
}

/* utility.rs:93-100 Hit::at_infinity */
static orc_hit hit_at_infinity(v3 dir) {
    orc_hit h;
    h.t = INFINITY;
    h.position = dir;
    h.normal = dir;
    h.u = 0.5 - atan2(dir.z, dir.x) / ORC_TAU;
    h.v = asin(dir.y) / ORC_PI + 0.5;
    return h;
}

/* ------------------------------------------------------------------ integrator ---------- */

/* render.rs:125-146 trace_path_continue; render.rs:102-122 trace_path_first differs only in
 * also reporting `hit`. Recursive exactly like the reference so the radiance fold
 * emit + absorb * (…) is evaluated inside-out. */
static v3 trace_path_rec(const orc_scene* s, const rtp_ray* ray, uint32_t depth, orc_rng* rng, orc_counters* c, int* first_hit) {
    if (!first_hit && depth == 0) return v3_make(0, 0, 0);
    orc_hit hit; uint32_t leaf;
    v3 dir = v3_from(ray->direction);
    if (scene_hit(s, ray, 0, &hit, &leaf, c)) {
        if (first_hit) *first_hit = 1;
        const rtp_material* m = &s->materials[hit.material];
        rtp_ray scattered;
        int has = scatter_evaluate(m, dir, &hit, rng, &scattered); /* material.rs:106 */
        v3 absorb = absorb_evaluate(s, m, &hit);                   /* material.rs:107 */
        v3 emit = emit_evaluate(s, &m->emit, dir, &hit);           /* material.rs:108 */
        v3 rest = v3_make(0, 0, 0);
        if (has) rest = v3_mul(absorb, trace_path_rec(s, &scattered, depth - 1, rng, c, NULL));
        return v3_add(emit, rest);
    }
    if (first_hit) *first_hit = 0;
    orc_hit inf = hit_at_infinity(dir);
    return emit_evaluate(s, &s->background, dir, &inf);
}

/* render.rs:32-52 Camera::shoot */
static void camera_shoot(const rtp_camera* cam, double u, double v, orc_rng* rng, rtp_ray* out) {
    double tan_fov = tan(0.5 * cam->fov);
    double lx = 0.0, ly = 0.0;
    if (rng) { /* drawn even when lens_radius == 0 (render.rs:36) */
        sample_unit_disk(rng, &lx, &ly);
        lx = cam->lens_radius * lx;
        ly = cam->lens_radius * ly;
    }
    v3 origin = v3_make(lx, ly, 0.0);
    v3 target = v3_make((2.0 * u - 1.0) * tan_fov * cam->focal_dist * cam->aspect_ratio,
                        (2.0 * v - 1.0) * tan_fov * cam->focal_dist,
                        -cam->focal_dist);
    v3 direction = v3_normalize(v3_sub(target, origin));
    const double* m = cam->orientation; /* columns x,y,z */
    /* utility.rs:185-191 transform_vector / transform_point */
    v3 wd = v3_make((m[0] * direction.x + m[3] * direction.y) + m[6] * direction.z,
                    (m[1] * direction.x + m[4] * direction.y) + m[7] * direction.z,
                    (m[2] * direction.x + m[5] * direction.y) + m[8] * direction.z);
    v3 wo = v3_make(((m[0] * origin.x + m[3] * origin.y) + m[6] * origin.z) + cam->position[0],
                    ((m[1] * origin.x + m[4] * origin.y) + m[7] * origin.z) + cam->position[1],
                    ((m[2] * origin.x + m[5] * origin.y) + m[8] * origin.z) + cam->position[2]);
    out->origin[0] = wo.x; out->origin[1] = wo.y; out->origin[2] = wo.z;
    out->direction[0] = wd.x; out->direction[1] = wd.y; out->direction[2] = wd.z;
    out->t_min = RAY_EPSILON;
    out->t_max = INFINITY;
}

void orc_camera_rays(const rtp_camera* cam, uint32_t width, uint32_t height, rtp_ray* rays) {
    for (uint32_t j = 0; j < height; ++j)
        for (uint32_t i = 0; i < width; ++i) {
            double u = ((double)i + 0.5) / (double)width;
            double v = ((double)j + 0.5) / (double)height;
            camera_shoot(cam, u, v, NULL, &rays[(size_t)i + (size_t)j * width]);
        }
}

/* One sample of one pixel: main.rs:70-83 with the counter-based stream of rtp.h */
static v3 trace_sample(const orc_scene* s, const rtp_camera* cam, const rtp_render_params* p, uint32_t i, uint32_t j,
                       uint32_t smp, orc_counters* c, int* hit) {
    orc_rng rng;
    rng_init(&rng, p->seed, j * p->width + i, smp, RTP_RNG_STREAM_PATH);
    /* render.rs:76-81 make_uv_jitter */
    double u = ((double)i + rng_next(&rng)) / (double)p->width;
    double v = ((double)j + rng_next(&rng)) / (double)p->height;
    rtp_ray ray;
    camera_shoot(cam, u, v, &rng, &ray);
    int h = 0;
    v3 col = trace_path_rec(s, &ray, p->max_bounce, &rng, c, &h);
    *hit = h;
    return col;
}

int orc_trace_one(const orc_scene* s, const rtp_camera* cam, const rtp_render_params* p, uint32_t i, uint32_t j, uint32_t smp,
                  double rgb[3], int* hit, uint32_t* n_rays) {
    if (!s || !cam || !p || p->max_bounce < 1) return fail(RTP_ERR_INVALID, "bad argument");
    orc_counters c = {0, 0, 0, 0};
    v3 col = trace_sample(s, cam, p, i, j, smp, &c, hit);
    rgb[0] = col.x; rgb[1] = col.y; rgb[2] = col.z;
    if (n_rays) *n_rays = (uint32_t)c.rays;
    return RTP_OK;
}

/* ------------------------------------------------------------------ drivers ------------- */

size_t orc_split_in_tiles(uint32_t fw, uint32_t fh, uint32_t tw, uint32_t th, uint32_t* out, size_t cap) {
    /* image.rs:151-167 */
    uint32_t nti = (fw + tw - 1) / tw, ntj = (fh + th - 1) / th;
    size_t n = 0;
    for (uint32_t tj = 0; tj < ntj; ++tj)
        for (uint32_t ti = 0; ti < nti; ++ti) {
            if (out && n < cap) {
                uint32_t oi = ti * tw, oj = tj * th;
                out[4 * n + 0] = oi; out[4 * n + 1] = oj;
                out[4 * n + 2] = tw < fw - oi ? tw : fw - oi;
                out[4 * n + 3] = th < fh - oj ? th : fh - oj;
            }
            ++n;
        }
    return n;
}

typedef struct {
    const orc_scene* scene;
    const rtp_camera* cam;
    const rtp_render_params* p;
    double* rgb; double* fg;
    uint32_t* tiles; size_t n_tiles; /* LIFO job queue, main.rs:41,58 */
    pthread_mutex_t lock;
    orc_counters total; uint64_t paths;
} render_job;

static void* render_worker(void* arg) {
    render_job* job = (render_job*)arg;
    const rtp_render_params* p = job->p;
    orc_counters c = {0, 0, 0, 0};
    uint64_t paths = 0;
    uint32_t ns = p->sample_end - p->sample_begin;
    for (;;) {
        uint32_t tile[4];
        pthread_mutex_lock(&job->lock);
        if (job->n_tiles == 0) { pthread_mutex_unlock(&job->lock); break; }
        memcpy(tile, job->tiles + 4 * (--job->n_tiles), sizeof tile); /* pop() from the end */
        pthread_mutex_unlock(&job->lock);
        for (uint32_t tj = 0; tj < tile[3]; ++tj) {
            /* rtp_render_params.row_offset / row_stride: only rows tile_y + row_offset + k * row_stride belong to this call */
            const uint32_t rs = p->row_stride ? p->row_stride : 1u, rel = tj + tile[1] - p->tile_y;
            if (rel < p->row_offset || (rel - p->row_offset) % rs != 0) continue;
            for (uint32_t ti = 0; ti < tile[2]; ++ti) {
                uint32_t i = ti + tile[0], j = tj + tile[1];
                v3 final_color = v3_make(0, 0, 0);
                double foreground = 0.0;
                for (uint32_t smp = p->sample_begin; smp < p->sample_end; ++smp) {
                    int hit;
                    v3 col = trace_sample(job->scene, job->cam, p, i, j, smp, &c, &hit);
                    final_color = v3_add(final_color, col); /* main.rs:80 */
                    if (hit) foreground += 1.0;
                }
                paths += ns;
                size_t px = (size_t)i + (size_t)j * p->width;
                if (!(p->flags & RTP_RENDER_RAW_SUMS)) { /* main.rs:86-87 */
                    double n = (double)p->num_samples;
                    final_color = v3_make(final_color.x / n, final_color.y / n, final_color.z / n);
                    foreground = foreground / n;
                }
                job->rgb[3 * px + 0] = final_color.x;
                job->rgb[3 * px + 1] = final_color.y;
                job->rgb[3 * px + 2] = final_color.z;
                if (job->fg) job->fg[px] = foreground;
            }
        }
    }
    pthread_mutex_lock(&job->lock);
    job->total.rays += c.rays; job->total.node_visits += c.node_visits;
    job->total.triangle_tests += c.triangle_tests; job->total.sphere_tests += c.sphere_tests;
    job->paths += paths;
    pthread_mutex_unlock(&job->lock);
    return NULL;
}

int orc_render(const orc_scene* s, const rtp_camera* cam, const rtp_render_params* p, double* rgb, double* fg,
               int n_threads, rtp_stats* stats) {
    if (!s || !cam || !p || !rgb) return fail(RTP_ERR_INVALID, "null argument");
    if (p->max_bounce < 1) return fail(RTP_ERR_INVALID, "assert!(depth >= 1) (render.rs:97)");
    if (p->width == 0 || p->height == 0 || p->sample_end < p->sample_begin) return fail(RTP_ERR_INVALID, "bad frame");
    uint32_t tx = p->tile_x, ty = p->tile_y;
    uint32_t tw = p->tile_w ? p->tile_w : p->width - tx, th = p->tile_h ? p->tile_h : p->height - ty;
    if (tx + tw > p->width || ty + th > p->height) return fail(RTP_ERR_INVALID, "tile outside frame");
    if (n_threads < 1) n_threads = 1;
    render_job job;
    memset(&job, 0, sizeof job);
    job.scene = s; job.cam = cam; job.p = p; job.rgb = rgb; job.fg = fg;
    size_t nt = orc_split_in_tiles(tw, th, 32, 32, NULL, 0); /* main.rs:26,36 */
    job.tiles = (uint32_t*)malloc(sizeof(uint32_t) * 4 * (nt ? nt : 1));
    orc_split_in_tiles(tw, th, 32, 32, job.tiles, nt);
    for (size_t k = 0; k < nt; ++k) { job.tiles[4 * k] += tx; job.tiles[4 * k + 1] += ty; }
    job.n_tiles = nt;
    pthread_mutex_init(&job.lock, NULL);
    pthread_t* th_ids = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int k = 0; k < n_threads; ++k) pthread_create(&th_ids[k], NULL, render_worker, &job);
    for (int k = 0; k < n_threads; ++k) pthread_join(th_ids[k], NULL);
    pthread_mutex_destroy(&job.lock);
    free(th_ids); free(job.tiles);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->rays = job.total.rays; stats->paths = job.paths;
        stats->node_visits = job.total.node_visits;
        stats->triangle_tests = job.total.triangle_tests; stats->sphere_tests = job.total.sphere_tests;
    }
    return RTP_OK;
}

typedef struct {
    const orc_scene* scene; const rtp_ray* rays; rtp_hit_full* hits;
    size_t begin, end; int mode;
    orc_counters c;
} trace_job;

static void* trace_worker(void* arg) {
    trace_job* job = (trace_job*)arg;
    const orc_scene* s = job->scene;
    for (size_t k = job->begin; k < job->end; ++k) {
        orc_hit h; uint32_t leaf;
        rtp_hit_full* o = &job->hits[k];
        if (s->n_hittables && scene_hit(s, &job->rays[k], job->mode == 1, &h, &leaf, &job->c)) {
            o->leaf = leaf;
            o->material = h.material;
            o->t = h.t;
            o->position[0] = h.position.x; o->position[1] = h.position.y; o->position[2] = h.position.z;
            o->normal[0] = h.normal.x; o->normal[1] = h.normal.y; o->normal[2] = h.normal.z;
            o->uv[0] = h.u; o->uv[1] = h.v;
        } else {
            memset(o, 0, sizeof *o);
            o->leaf = RTP_MISS; o->material = RTP_MISS; o->t = INFINITY;
        }
    }
    return NULL;
}

int orc_trace_closest(const orc_scene* s, const rtp_ray* rays, size_t n, rtp_hit_full* hits, int mode, int n_threads,
                      rtp_stats* stats) {
    if (!s || (n && (!rays || !hits))) return fail(RTP_ERR_INVALID, "null argument");
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_threads > n) n_threads = n ? (int)n : 1;
    trace_job* jobs = (trace_job*)calloc((size_t)n_threads, sizeof *jobs);
    pthread_t* ids = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int k = 0; k < n_threads; ++k) {
        jobs[k].scene = s; jobs[k].rays = rays; jobs[k].hits = hits; jobs[k].mode = mode;
        jobs[k].begin = n * (size_t)k / (size_t)n_threads;
        jobs[k].end = n * (size_t)(k + 1) / (size_t)n_threads;
        pthread_create(&ids[k], NULL, trace_worker, &jobs[k]);
    }
    orc_counters c = {0, 0, 0, 0};
    for (int k = 0; k < n_threads; ++k) {
        pthread_join(ids[k], NULL);
        c.rays += jobs[k].c.rays; c.node_visits += jobs[k].c.node_visits;
        c.triangle_tests += jobs[k].c.triangle_tests; c.sphere_tests += jobs[k].c.sphere_tests;
    }
    free(jobs); free(ids);
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->rays = c.rays; stats->node_visits = c.node_visits;
        stats->triangle_tests = c.triangle_tests; stats->sphere_tests = c.sphere_tests;
    }
    return RTP_OK;
}

/* ------------------------------------------------------------------ assets -------------- */

/* utility.rs:172-177 Transformation::lookat */
int orc_camera_lookat(const double position[3], const double target[3], const double up[3], rtp_camera* cam) {
    if (!position || !target || !up || !cam) return fail(RTP_ERR_INVALID, "null argument");
    v3 z = v3_normalize(v3_sub(v3_from(position), v3_from(target)));
    v3 x = v3_cross(v3_from(up), z);
    v3 y = v3_cross(z, x);
    cam->orientation[0] = x.x; cam->orientation[1] = x.y; cam->orientation[2] = x.z;
    cam->orientation[3] = y.x; cam->orientation[4] = y.y; cam->orientation[5] = y.z;
    cam->orientation[6] = z.x; cam->orientation[7] = z.y; cam->orientation[8] = z.z;
    cam->position[0] = position[0]; cam->position[1] = position[1]; cam->position[2] = position[2];
    return RTP_OK;
}

/* utility.rs:212-220 to_srgb_u8 */
void orc_frame_to_srgb8(const double* rgb, uint32_t width, uint32_t height, uint8_t* out) {
    size_t n = (size_t)width * height;
    for (size_t p = 0; p < n; ++p) {
        for (int k = 0; k < 3; ++k) out[4 * p + k] = sat_u8(255.0 * pow(clampd(rgb[3 * p + k], 0.0, 1.0), 1.0 / 2.2));
        out[4 * p + 3] = 0xff;
    }
}

/* mesh.rs:39-136 obj_parser. A line is `v|vn|vt|f`, one or more blanks, then the payload; lines
 * that do not parse are skipped (mesh.rs:117-120); trailing text is ignored. */
typedef struct { uint32_t p, n, t; } obj_index; /* n,t = RTP_MISS for None */

static const char* skip_space1(const char* s) { /* nom space1: one or more ' ' or '\t' */
    if (*s != ' ' && *s != '\t') return NULL;
    while (*s == ' ' || *s == '\t') ++s;
    return s;
}
static const char* parse_double(const char* s, double* out) {
    /* nom `double` does not skip leading whitespace */
    if (*s == ' ' || *s == '\t' || *s == '\n' || *s == '\r' || *s == '\0') return NULL;
    char* end;
    double v = strtod(s, &end);
    if (end == s) return NULL;
    *out = v;
    return end;
}
static const char* parse_vec(const char* s, double* out, int n) {
    for (int k = 0; k < n; ++k) {
        if (k) { s = skip_space1(s); if (!s) return NULL; }
        s = parse_double(s, &out[k]);
        if (!s) return NULL;
    }
    return s;
}
/* mesh.rs:58-70 parse_index: separated_list1("/", opt(integer)); position = [0]-1, texcoord = [1]-1, normal = [2]-1 */
static const char* parse_index(const char* s, obj_index* out) {
    uint32_t vals[3] = {0, 0, 0}; int have[3] = {0, 0, 0};
    int k = 0;
    for (;;) {
        uint64_t v = 0; int digits = 0;
        while (*s >= '0' && *s <= '9') { v = v * 10 + (uint64_t)(*s - '0'); ++s; ++digits; if (v > 0xFFFFFFFFull) return NULL; }
        if (k < 3) { have[k] = digits > 0; vals[k] = (uint32_t)v; }
        ++k;
        if (*s == '/') { ++s; continue; }
        break;
    }
    if (!have[0]) return NULL; /* "Position index not provided" */
    out->p = vals[0] - 1;
    out->t = have[1] ? vals[1] - 1 : RTP_MISS;
    out->n = have[2] ? vals[2] - 1 : RTP_MISS;
    return s;
}

typedef struct { void* data; size_t n, cap, elem; } vec_t;
static void* vec_push(vec_t* v) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 1024;
        v->data = realloc(v->data, v->cap * v->elem);
    }
    return (char*)v->data + (v->n++) * v->elem;
}

int orc_obj_load(const char* path, rtp_mesh* out) {
    if (!path || !out) return fail(RTP_ERR_INVALID, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(RTP_ERR_IO, "cannot open obj file");
    vec_t pos = {0, 0, 0, 24}, nrm = {0, 0, 0, 24}, tex = {0, 0, 0, 16};
    vec_t corners = {0, 0, 0, sizeof(obj_index)}, faces = {0, 0, 0, 8}; /* faces: first_vertex, num_vertices */
    char* line = NULL; size_t cap = 0;
    while (getline(&line, &cap, f) >= 0) {
        const char* s = line;
        double v[3];
        if (s[0] == 'v' && (s[1] == ' ' || s[1] == '\t')) {
            const char* q = skip_space1(s + 1);
            if (q && parse_vec(q, v, 3)) memcpy(vec_push(&pos), v, 24);
        } else if (s[0] == 'v' && s[1] == 'n') {
            const char* q = skip_space1(s + 2);
            if (q && parse_vec(q, v, 3)) memcpy(vec_push(&nrm), v, 24);
        } else if (s[0] == 'v' && s[1] == 't') {
            const char* q = skip_space1(s + 2);
            if (q && parse_vec(q, v, 2)) memcpy(vec_push(&tex), v, 16);
        } else if (s[0] == 'f') {
            const char* q = skip_space1(s + 1);
            if (!q) continue;
            size_t first = corners.n; uint32_t cnt = 0;
            for (;;) {
                obj_index ix;
                const char* e = parse_index(q, &ix);
                if (!e) break;
                *(obj_index*)vec_push(&corners) = ix; ++cnt;
                q = skip_space1(e);
                if (!q) break;
            }
            if (cnt == 0) continue; /* separated_list1 needs one element, else the line is skipped */
            uint32_t* fc = (uint32_t*)vec_push(&faces);
            fc[0] = (uint32_t)first; fc[1] = cnt;
        }
    }
    free(line); fclose(f);

    /* mesh.rs:151-166: unique (p,n,t) triples in first-seen order */
    size_t hcap = 16; while (hcap < corners.n * 2 + 16) hcap *= 2;
    uint32_t* slots = (uint32_t*)malloc(sizeof(uint32_t) * hcap);
    memset(slots, 0xff, sizeof(uint32_t) * hcap);
    obj_index* keys = (obj_index*)malloc(sizeof(obj_index) * (corners.n ? corners.n : 1));
    rtp_vertex* verts = (rtp_vertex*)malloc(sizeof(rtp_vertex) * (corners.n ? corners.n : 1));
    uint32_t* corner_vertex = (uint32_t*)malloc(sizeof(uint32_t) * (corners.n ? corners.n : 1));
    uint32_t nv = 0; int rc = RTP_OK;
    for (size_t k = 0; k < corners.n && rc == RTP_OK; ++k) {
        obj_index ix = ((obj_index*)corners.data)[k];
        uint64_t h = ((uint64_t)ix.p * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)ix.n * 0xC2B2AE3D27D4EB4Full) ^ ((uint64_t)ix.t * 0x165667B19E3779F9ull);
        size_t slot = (size_t)(h >> 20) & (hcap - 1);
        for (;;) {
            uint32_t id = slots[slot];
            if (id == RTP_MISS) {
                if (ix.p >= pos.n || (ix.n != RTP_MISS && ix.n >= nrm.n) || (ix.t != RTP_MISS && ix.t >= tex.n)) {
                    rc = fail(RTP_ERR_FORMAT, "obj index out of range (the reference panics)"); break;
                }
                slots[slot] = nv; keys[nv] = ix;
                rtp_vertex* v = &verts[nv];
                memcpy(v->position, (char*)pos.data + 24 * (size_t)ix.p, 24);
                if (ix.n != RTP_MISS) memcpy(v->normal, (char*)nrm.data + 24 * (size_t)ix.n, 24);
                else v->normal[0] = v->normal[1] = v->normal[2] = 0.0;
                if (ix.t != RTP_MISS) memcpy(v->uv, (char*)tex.data + 16 * (size_t)ix.t, 16);
                else v->uv[0] = v->uv[1] = 0.0;
                corner_vertex[k] = nv++;
                break;
            }
            if (keys[id].p == ix.p && keys[id].n == ix.n && keys[id].t == ix.t) { corner_vertex[k] = id; break; }
            slot = (slot + 1) & (hcap - 1);
        }
    }
    uint32_t* indices = (uint32_t*)malloc(sizeof(uint32_t) * (faces.n ? faces.n * 3 : 1));
    for (size_t k = 0; k < faces.n && rc == RTP_OK; ++k) { /* mesh.rs:169-179 */
        uint32_t* fc = (uint32_t*)faces.data + 2 * k;
        if (fc[1] != 3) { rc = fail(RTP_ERR_FORMAT, "Non-triangular face are not supported"); break; }
        indices[3 * k + 0] = corner_vertex[fc[0] + 0];
        indices[3 * k + 1] = corner_vertex[fc[0] + 1];
        indices[3 * k + 2] = corner_vertex[fc[0] + 2];
    }
    free(slots); free(keys); free(corner_vertex);
    free(pos.data); free(nrm.data); free(tex.data); free(corners.data);
    if (rc != RTP_OK) { free(verts); free(indices); free(faces.data); return rc; }
    out->vertices = verts; out->indices = indices;
    out->n_vertices = nv; out->n_indices = (uint32_t)(faces.n * 3);
    out->material = 0; out->_pad = 0; /* mesh.rs:181 MaterialId(0) */
    free(faces.data);
    return RTP_OK;
}

void orc_mesh_free(rtp_mesh* m) {
    if (!m) return;
    free((void*)m->vertices); free((void*)m->indices);
    m->vertices = NULL; m->indices = NULL; m->n_vertices = m->n_indices = 0;
}

/* image.rs:73-114 tga::load */
int orc_tga_load(const char* path, rtp_image* out) {
    if (!path || !out) return fail(RTP_ERR_INVALID, "null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(RTP_ERR_IO, "cannot open tga file");
    uint8_t hd[18];
    if (fread(hd, 1, 18, f) != 18) { fclose(f); return fail(RTP_ERR_IO, "short tga header"); }
    uint32_t w = hd[12] | (hd[13] << 8), h = hd[14] | (hd[15] << 8);
    uint8_t bpp = hd[16], desc = hd[17];
    if (hd[0] != 0 || hd[1] != 0 || hd[2] != 2 || (bpp != 24 && bpp != 32)) {
        fclose(f); return fail(RTP_ERR_FORMAT, "This tga header is not supported");
    }
    size_t bytes = (size_t)w * h * (bpp / 8);
    uint8_t* raw = (uint8_t*)malloc(bytes ? bytes : 1);
    if (fread(raw, 1, bytes, f) != bytes) { free(raw); fclose(f); return fail(RTP_ERR_IO, "short tga data"); }
    fclose(f);
    size_t out_bytes = (size_t)w * h * 4;
    uint8_t* rgba = (uint8_t*)malloc(out_bytes ? out_bytes : 1);
    const uint8_t* p = raw;
    for (uint32_t y = 0; y < h; ++y) {
        uint32_t yy = (desc & (1 << 5)) ? h - 1 - y : y;
        for (uint32_t x = 0; x < w; ++x) {
            uint8_t* o = rgba + 4 * ((size_t)x + (size_t)yy * w);
            o[0] = p[2]; o[1] = p[1]; o[2] = p[0];
            if (bpp == 32) { o[3] = p[3]; p += 4; } else { o[3] = 0xff; p += 3; }
        }
    }
    free(raw);
    out->rgba = rgba; out->width = w; out->height = h;
    return RTP_OK;
}

/* image.rs:116-137 tga::save */
int orc_tga_save(const rtp_image* img, const char* path) {
    if (!img || !path || !img->rgba) return fail(RTP_ERR_INVALID, "null argument");
    if (img->width > 0xFFFF || img->height > 0xFFFF) return fail(RTP_ERR_INVALID, "image too large for tga (try_into fails)");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(RTP_ERR_IO, "cannot create tga file");
    uint8_t hd[18]; memset(hd, 0, sizeof hd);
    hd[2] = 2; hd[16] = 32;
    hd[12] = (uint8_t)(img->width & 0xff); hd[13] = (uint8_t)(img->width >> 8);
    hd[14] = (uint8_t)(img->height & 0xff); hd[15] = (uint8_t)(img->height >> 8);
    fwrite(hd, 1, 18, f);
    size_t n = (size_t)img->width * img->height;
    for (size_t k = 0; k < n; ++k) {
        const uint8_t* p = img->rgba + 4 * k;
        uint8_t bgra[4] = {p[2], p[1], p[0], p[3]};
        fwrite(bgra, 1, 4, f);
    }
    fclose(f);
    return RTP_OK;
}

void orc_image_free(rtp_image* img) {
    if (!img) return;
    free(img->rgba); img->rgba = NULL; img->width = img->height = 0;
}
