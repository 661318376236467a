/*
 * rtp_oracle.h — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the hot path of /root/reference (crate `raytracing2`), in the
 * reference's own f64 arithmetic and evaluation order. Only tests/, __graft_entry__.smoke()
 * and bench.py's CPU-baseline legs may load it; the product (raytracing-potato_b200/) never
 * links, imports or calls anything under oracle/.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path
 *  and cannot be compiled here (no rustc/cargo; nalgebra, rand, nom are not
 * vendored). The oracle is therefore pinned only by (i) analytic cases, (ii) its own
 * BVH-vs-brute-force differential, (iii) Random123's published Philox4x32-10 known answers,
 * (iv) asset invariants (vertex/index counts, TGA orientation) and (v) an independent numpy
 * restatement in tests/. See DESIGN.md §3.
 *
 * Third-party arithmetic restated from the published nalgebra 0.29 algorithm (source absent):
 * dot(a,b) = (ax*bx + ay*by) + az*bz; norm_squared = 0 + dot(v,v); normalize = v / sqrt(ns)
 * component-wise; M*v accumulates column by column (same bits as the row dot in that order);
 * cross = (ay*bz-az*by, az*bx-ax*bz, ax*by-ay*bx). powi(x,2)=x*x, powi(x,5)=x*((x*x)*(x*x)).
 */
#ifndef RTP_ORACLE_H
#define RTP_ORACLE_H

#include "../include/rtp.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

const char* orc_last_error(void);

/* randomness: the shared counter-based stream (rtp.h rtp_rng_draws) */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_rng_draws(uint64_t seed, uint32_t index_lo, uint32_t index_hi, uint32_t stream,
                   uint32_t first_draw, uint32_t n_draws, double* out);

/* assets */
int orc_obj_load(const char* path, rtp_mesh* out);  /* mesh.rs:112-183 */
void orc_mesh_free(rtp_mesh* mesh);
int orc_tga_load(const char* path, rtp_image* out); /* image.rs:73-114 */
int orc_tga_save(const rtp_image* image, const char* path); /* image.rs:116-137 */
void orc_image_free(rtp_image* image);
int orc_camera_lookat(const double position[3], const double target[3], const double up[3],
                      rtp_camera* camera);           /* utility.rs:172-177 */
void orc_frame_to_srgb8(const double* rgb, uint32_t width, uint32_t height, uint8_t* rgba_out);
size_t orc_split_in_tiles(uint32_t fw, uint32_t fh, uint32_t tw, uint32_t th, uint32_t* out, size_t cap);

/* scene */
int orc_scene_create(const rtp_scene_desc* desc, orc_scene** out); /* bvh.rs:70-91 */
void orc_scene_destroy(orc_scene* scene);
int orc_scene_get_info(const orc_scene* scene, rtp_scene_info* info);
int orc_scene_leaf_order(const orc_scene* scene, uint32_t* out, size_t cap);
/* node dump for layout tests: per node 6 doubles (min,max) + left,right,leaf (u32; leaf = RTP_MISS on branches) */
int orc_scene_node(const orc_scene* scene, uint32_t node, double aabb[6], uint32_t lrl[3]);
uint32_t orc_scene_root(const orc_scene* scene);

/* primitives, exposed for unit tests */
int orc_aabb_collide(const double bmin[3], const double bmax[3], const rtp_ray* ray); /* utility.rs:137-154 */

/* closest hit through the scene root (bvh.rs:121-124 or hittable.rs:110-120).
 * mode 0 = as the scene's root kind says, 1 = force brute-force list scan over the same leaves. */
int orc_trace_closest(const orc_scene* scene, const rtp_ray* rays, size_t n, rtp_hit_full* hits_out,
                      int mode, int n_threads, rtp_stats* stats);

/* camera rays at pixel centres (render.rs:32-52 with lens_radius forced to 0 and no draw) */
void orc_camera_rays(const rtp_camera* camera, uint32_t width, uint32_t height, rtp_ray* rays_out);

/* the worker loop of main.rs:36-92 with n_threads workers popping 32x32 tiles LIFO */
int orc_render(const orc_scene* scene, const rtp_camera* camera, const rtp_render_params* params,
               double* rgb_out, double* foreground_out, int n_threads, rtp_stats* stats);

/* one path, for fine-grained tests: returns final colour, first-hit flag and number of rays */
int orc_trace_one(const orc_scene* scene, const rtp_camera* camera, const rtp_render_params* params,
                  uint32_t i, uint32_t j, uint32_t s, double rgb[3], int* hit, uint32_t* n_rays);

/* texture sample at an explicit hit (texture.rs:20-36) */
int orc_texture_sample(const orc_scene* scene, uint32_t texture, const double position[3],
                       const double uv[2], double rgb[3]);
int64_t orc_noise_integer(int64_t x, int64_t y, int64_t z, int64_t seed); /* randomness.rs:91-105 */

#ifdef __cplusplus
}
#endif
#endif
