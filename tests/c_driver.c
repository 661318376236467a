/* Plain-C driver of the drop-in boundary: includes include/rtp.h, links -lrtp_b200, makes the calls a non-Python host makes.
 * tests/test_abi_host.py compiles it with gcc and runs it: without a GPU it must get RTP_ERR_CUDA (there is no CPU fallback)
 * AFTER the host-only entry points worked; with a GPU (pytest -m gpu) it traces a small scene and checks the answers that can
 * be known in closed form. Exit code 0 = every expectation held. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "rtp.h"

static int fail(const char* what) {
    fprintf(stderr, "c_driver: %s (rtp_last_error: %s)\n", what, rtp_last_error());
    return 1;
}

int main(int argc, char** argv) {
    int want_gpu = argc > 1 && strcmp(argv[1], "gpu") == 0;
    if (rtp_abi_version() != RTP_ABI_VERSION) return fail("abi version");

    /* host-only entry points: lookat (utility.rs:172-177), tiles (image.rs:151-167), the random stream */
    rtp_camera cam;
    memset(&cam, 0, sizeof cam);
    const double pos[3] = {0.0, 0.0, 5.0}, target[3] = {0.0, 0.0, 0.0}, up[3] = {0.0, 1.0, 0.0};
    if (rtp_camera_lookat(pos, target, up, &cam) != RTP_OK) return fail("lookat");
    cam.aspect_ratio = 1.0; cam.fov = 0.7853981633974483; cam.focal_dist = 1.0; cam.lens_radius = 0.0;
    if (cam.orientation[8] != 1.0 || cam.position[2] != 5.0) return fail("lookat values");
    uint32_t tiles[4 * 16];
    size_t n_tiles = 0;
    if (rtp_split_in_tiles(100, 70, 32, 32, tiles, 16, &n_tiles) != RTP_OK || n_tiles != 12) return fail("split_in_tiles");
    double draws[4];
    if (rtp_rng_draws(1, 2, 3, RTP_RNG_STREAM_PATH, 0, 4, draws) != RTP_OK || !(draws[0] >= 0.0 && draws[0] < 1.0)) return fail("rng draws");

    /* scene: one emissive triangle x + y + z = 1 (example_scenes.rs:222-262) and a unit sphere at the origin, List root */
    rtp_vertex v[3];
    memset(v, 0, sizeof v);
    v[0].position[0] = 1.0; v[1].position[1] = 1.0; v[2].position[2] = 1.0;
    for (int k = 0; k < 3; ++k) v[k].normal[0] = v[k].normal[1] = v[k].normal[2] = 0.5773502691896258;
    const uint32_t idx[3] = {0, 1, 2};
    rtp_mesh mesh = {v, idx, 3, 3, 0, 0};
    rtp_hittable hs[2];
    memset(hs, 0, sizeof hs);
    hs[0].kind = RTP_HITTABLE_TRIANGLE; hs[0].mesh = 0; hs[0].triangle = 0;
    hs[1].kind = RTP_HITTABLE_SPHERE; hs[1].material = 1; hs[1].center[0] = -3.0; hs[1].radius = 1.0;
    rtp_material mats[2];
    memset(mats, 0, sizeof mats);
    mats[0].scatter = RTP_SCATTER_NONE; mats[0].absorb = RTP_ABSORB_BLACKBODY; mats[0].emit.kind = RTP_EMIT_DEBUG_NORMALS;
    mats[1].scatter = RTP_SCATTER_LAMBERT; mats[1].absorb = RTP_ABSORB_WHITEBODY; mats[1].emit.kind = RTP_EMIT_NONE;
    rtp_scene_desc desc;
    memset(&desc, 0, sizeof desc);
    desc.abi_version = RTP_ABI_VERSION; desc.root_kind = RTP_ROOT_BVH;
    desc.meshes = &mesh; desc.n_meshes = 1; desc.hittables = hs; desc.n_hittables = 2;
    desc.materials = mats; desc.n_materials = 2; desc.background.kind = RTP_EMIT_SKY_GRADIENT;

    uint32_t order[2];
    rtp_scene_info info;
    if (rtp_bvh_build_order(&desc, order, 2, &info) != RTP_OK || info.n_leaves != 2 || info.n_nodes != 3) return fail("bvh_build_order");
    if (order[0] != 1 || order[1] != 0) return fail("leaf order: the sphere's centroid (-3) sorts first on x");

    /* validation happens before any device is touched: a bad material id is RTP_ERR_INVALID with or without a GPU */
    hs[1].material = 7;
    rtp_scene* scene = NULL;
    if (rtp_scene_create(&desc, &scene) != RTP_ERR_INVALID || scene != NULL) return fail("bad material must be RTP_ERR_INVALID");
    hs[1].material = 1;

    int rc = rtp_scene_create(&desc, &scene);
    if (!want_gpu) {
        if (rc == RTP_OK) { rtp_scene_destroy(scene); printf("c_driver: a GPU is present; host checks passed\n"); return 0; }
        if (rc != RTP_ERR_CUDA || !strstr(rtp_last_error(), "no CPU fallback")) return fail("without a GPU rtp_scene_create must fail with RTP_ERR_CUDA");
        printf("c_driver: host checks passed; no GPU: %s\n", rtp_last_error());
        return 0;
    }
    if (rc != RTP_OK) return fail("rtp_scene_create");

    /* three rays: one through the triangle's centroid along (-1,-1,-1)/sqrt3 from (1,1,1): t = 2/sqrt(3) exactly up to rounding;
     * one at the sphere: t = 5 - 1; one that misses */
    rtp_ray rays[3];
    memset(rays, 0, sizeof rays);
    for (int k = 0; k < 3; ++k) { rays[k].t_min = 1e-3; rays[k].t_max = INFINITY; }
    rays[0].origin[0] = rays[0].origin[1] = rays[0].origin[2] = 1.0;
    rays[0].direction[0] = rays[0].direction[1] = rays[0].direction[2] = -1.0; /* not normalised on purpose (utility.rs:54) */
    rays[1].origin[0] = -3.0; rays[1].origin[2] = 5.0; rays[1].direction[2] = -1.0;
    rays[2].origin[1] = 10.0; rays[2].direction[1] = 1.0;
    rtp_hit hits[3];
    rtp_stats st;
    if (rtp_trace_closest(scene, rays, 3, hits, &st) != RTP_OK) return fail("rtp_trace_closest");
    if (hits[0].leaf != 0 || hits[0].material != 0 || fabs(hits[0].t - 2.0 / 3.0) > 1e-15) return fail("triangle hit: t = 2/3 along (-1,-1,-1)");
    if (hits[1].leaf != 1 || hits[1].material != 1 || hits[1].t != 4.0) return fail("sphere hit: t = 4");
    if (hits[2].leaf != RTP_MISS || !isinf(hits[2].t)) return fail("miss");
    if (st.rays != 3) return fail("stats.rays");

    /* a 16 x 16 frame, 2 spp: every pixel finite, sky above */
    rtp_render_params p;
    memset(&p, 0, sizeof p);
    p.width = 16; p.height = 16; p.num_samples = 2; p.max_bounce = 4; p.seed = 1; p.sample_end = 2;
    double rgb[16 * 16 * 3], fg[16 * 16];
    if (rtp_render(scene, &cam, &p, rgb, fg, &st) != RTP_OK) return fail("rtp_render");
    for (int k = 0; k < 16 * 16 * 3; ++k) if (!(rgb[k] >= 0.0 && rgb[k] <= 1.0 + 1e-12)) return fail("pixel out of range");
    if (st.paths != 16 * 16 * 2 || st.rays < st.paths) return fail("render stats");
    rtp_scene_destroy(scene);
    printf("c_driver: GPU checks passed (%llu rays in the frame)\n", (unsigned long long)st.rays);
    return 0;
}
