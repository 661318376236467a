// rtp_internal.h — structures shared by the host layer (rtp_host.cpp) and the device layer
// (rtp_device.cu). Not part of the ABI.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rtp.h"

namespace rtp {

// ---------------------------------------------------------------------------------------------
// Device layout (HBM). All records are 16-byte aligned so they load as 128-bit vectors.
//
// Nodes are stored in the reference's depth-first, left-before-right order (pre-order), so the
// reference traversal bvh.rs:93-119 becomes a stack-free walk: a node that passes its slab test
// is followed by node+1 (its left child, or for a leaf simply the next subtree), a node that
// fails jumps to `skip`, the first pre-order index after its subtree.
// ---------------------------------------------------------------------------------------------

constexpr uint32_t kNoPrim = 0xFFFFFFFFu;

struct alignas(64) DNode {  // 64 B
    double bmin[3];
    double bmax[3];
    uint32_t skip;  // pre-order index of the next node once this subtree is done
    uint32_t prim;  // leaf: primitive slot (= DFS rank of the leaf); branch: kNoPrim
    uint32_t kind;  // leaf: rtp_hittable_kind
    uint32_t _pad;  // List roots: 1 = test the box in DPrim first (a leaf of a nested Bvh), 0 = no gate (hittable.rs:110-120)
};
static_assert(sizeof(DNode) == 64, "DNode must be 64 bytes");

// One primitive per leaf (bvh.rs:43-47), stored in DFS-rank order so a traversal touches them
// in increasing address order. Triangle: a, ba = a-b, ca = a-c precomputed on the host with the
// same IEEE subtraction the reference performs per test (hittable.rs:71-72) — bit-identical.
// Sphere: center in a[], radius in ba[0]. The leaf's own exact f64 box (the reference's per-leaf
// slab gate, bvh.rs:96) travels with the primitive so gate + test need one 128-byte record.
struct alignas(16) DPrim {  // 128 B
    double a[3];
    double ba[3];
    double ca[3];
    uint32_t leaf;      // LeafId handed back to the caller
    uint32_t material;  // MaterialId
    double bmin[3];
    double bmax[3];
};
static_assert(sizeof(DPrim) == 128, "DPrim must be 128 bytes");

// Culling tree in f32: a 4-wide tree over the rank-ordered leaf sequence. A node holds the boxes of its (up to) four
// children, in rank order, as six 4-lane planes (structure of arrays), so one 128-byte line answers four slab tests and a
// ray picks its near / far planes by ADDRESS (offset 0 or 16 inside each axis block) instead of by select. Every child box is
// the child's exact f64 box rounded OUTWARD and inflated by 2^-21 * max(|min|,|max|) per axis, so that an f32 slab
// evaluation with the matching ray-side slack is a rigorous lower / upper bound of the reference's f64 evaluation
// (rtp_device.cu collide32). child[k]: internal = index of the child node; leaf = kWideLeaf | kind << 30 |
// kWideBig if big | primitive slot; unused = kWideEmpty with an inverted (+inf, -inf) box that no ray can enter.
constexpr uint32_t kWideLeaf = 0x80000000u;
constexpr uint32_t kWideBig = 0x20000000u;       // leaf child word: the primitive is one of the scene's big primitives (trace_any_kernel tests those up front)
constexpr uint32_t kWideSlotMask = 0x1FFFFFFFu;  // leaf child word: primitive slot
constexpr uint32_t kWideEmpty = 0xFFFFFFFFu;
struct alignas(128) DWide {  // 128 B = one L1/L2 line
    float plane[3][2][4];    // [axis][0 = min, 1 = max][child]
    uint32_t child[4];
    uint32_t big_mask;       // bit k: child k is, or contains, one of the scene's "big" primitives (any-order walk: exempt from distance culling)
    uint32_t _pad[3];
};
static_assert(sizeof(DWide) == 128, "DWide must be 128 bytes");

// Shading attributes of a triangle (mesh.rs:7-11 normals/uvs of its three vertices), read once
// per accepted path vertex, never during traversal.
struct alignas(16) DAttr {  // 128 B; uv first so that both arrays start 16-byte aligned
    double uv[3][2];
    double n[3][3];
    double _pad;
};
static_assert(sizeof(DAttr) == 128, "DAttr must be 128 bytes");

struct DTexture {
    uint32_t kind;
    uint32_t width, height;
    uint32_t odd, even;
    uint32_t _pad;
    int64_t seed;
    double rgb[3];
    const uint8_t* rgba;  // device pointer
};

struct DMaterial {
    uint32_t scatter, absorb, absorb_texture, emit_kind;
    uint32_t emit_texture, _pad[3];
    double scatter_param;
    double absorb_rgb[3];
    double emit_rgb[3];
};

struct DSceneView {  // passed by value to kernels
    const DNode* nodes;
    const DWide* wide;
    const double* wide_boxes;  // [node][child][min xyz, max xyz] exact f64 child boxes (counting kernels: violation check)
    const DWide* any_wide;     // the tree the any-order lanes walk: `wide`, or the order-free tree of a big scene
    const double* any_boxes;
    const DPrim* prims;
    const DAttr* attrs;
    const DMaterial* materials;
    const DTexture* textures;
    uint32_t n_nodes;
    uint32_t n_prims;
    uint32_t root_kind;
    uint32_t bg_kind;
    uint32_t bg_texture;
    uint32_t f32_culling;  // 1: scene magnitudes allow the f32 conservative slab test
    double bg_rgb[3];
    // any-order walk: front-to-back traversal with distance culling, exact because every leaf that can matter is
    // still tested and the rare ray whose answer could depend on the reference's visiting order is re-walked in order
    uint32_t any_order;    // 1: eligible rays of this scene use it
    uint32_t any_cap;      // entries of the per-lane any-order stack
    uint32_t n_big;        // primitives exempt from distance culling (the outsized ones: extent above 16x the median), at most kMaxBig
    uint32_t _pad_any;
    uint32_t big[8];       // their slots | kind << 31
    double any_E;          // largest box extent of a triangle that is not big
    double any_A;          // largest |coordinate| of such a triangle
    float any_Ef, any_Af;  // the same, rounded up to f32
    float any_Cf, any_Rf;  // spheres that are not big: largest |centre coordinate| and radius, rounded up; any_Rf < 0: there are none
};
constexpr uint32_t kMaxBig = 8;

// ---------------------------------------------------------------------------------------------
// Host-side flattened scene produced by rtp_host.cpp and uploaded by rtp_device.cu
// ---------------------------------------------------------------------------------------------
// Big arrays that were BUILT on a device (rtp_build.cu device_build_scene) and never existed on the host: device_scene_upload
// adopts them on that device and copies them peer to peer to the other devices of a multi-device scene.
struct DevArrays {
    bool valid = false;
    mutable bool adopted = false;  // the replica on `device` owns the arrays now (device_scene_upload); the others copied from them
    int device = -1;
    DNode* nodes = nullptr; size_t n_nodes = 0;
    DWide* wide = nullptr; double* wide_boxes = nullptr; size_t n_wide = 0;
    DPrim* prims = nullptr; DAttr* attrs = nullptr; size_t n_prims = 0;
};

struct FlatScene {
    DevArrays dev;
    size_t prim_count() const { return dev.valid ? dev.n_prims : prims.size(); }
    size_t node_count() const { return dev.valid ? dev.n_nodes : nodes.size(); }
    size_t wide_count() const { return dev.valid ? dev.n_wide : wide.size(); }
    std::vector<DNode> nodes;
    std::vector<DWide> wide;
    std::vector<double> wide_boxes;
    uint32_t wide_depth = 0;         // nodes on the longest root-to-leaf path of the 4-wide tree
    std::vector<DPrim> prims;
    std::vector<DAttr> attrs;
    std::vector<DMaterial> materials;
    std::vector<DTexture> textures;            // rgba = nullptr here; filled at upload
    std::vector<std::vector<uint8_t>> images;  // per texture (empty unless Image)
    std::vector<uint32_t> leaf_order;          // LeafId by DFS rank
    uint32_t root_kind = 0;
    uint32_t depth = 0;              // of the reference tree (bvh.rs), what rtp_scene_info reports
    uint32_t n_reference_nodes = 0;  // 2n-1
    uint32_t device_depth = 0;       // of the culling tree the kernels walk
    bool list_gates = false;         // List root with nested Bvh items: some primitives are gated by a box (DNode::_pad)
    bool any_ok = false;             // the any-order walk may be used on this scene (prepare_any_order, rtp_host.cpp)
    std::vector<DWide> free_wide;    // second culling tree over the Morton order of the leaves (any-order walk of big scenes); may be empty
    std::vector<double> free_boxes;
    uint32_t free_depth = 0;
    uint32_t n_big = 0;
    uint32_t big[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double any_E = 0.0, any_A = 0.0, any_C = 0.0, any_R = 0.0;
    bool any_spheres = false;        // some sphere is not big: the slack bound needs the sphere term
    double scene_mag = 0.0;    // largest |coordinate| of any node box
    bool boxes_finite = true;  // every box is finite and ordered (min <= max): precondition of the sign-selected slab test, of the
                               // f32 culling walk and of re-shaping the tree; otherwise the reference topology and the literal test are used
    rtp_emit background{};
};

// error plumbing (rtp_host.cpp)
int set_error(int code, const std::string& msg);

// host layer
// device_build: compute the reference leaf order on the GPU (rtp_build.cu) when the scene is large; host-only callers pass false
int flatten_scene(const rtp_scene_desc* desc, FlatScene* out, bool device_build = false);

// device layer (rtp_device.cu, rtp_build.cu)
struct DeviceScene;
int ensure_device();  // binds device 0 if rtp_init has not been called; RTP_ERR_CUDA without a usable sm_100 GPU
int device_reference_order(const double* boxes /* n x {min xyz, max xyz} */, uint32_t n, uint32_t* order_out /* item index by DFS rank */);
int device_reference_order_dev(const double* d_boxes, uint32_t n, uint32_t* d_order_out);  // the same with device pointers
// The whole of flatten_scene's geometry part on the current device for a flat (un-nested) Bvh scene: leaf boxes, the reference's
// depth-first order, the SAH culling tree over that order, primitive / attribute records, the 4-wide collapse and the any-order
// tables. Fills out->dev and the host-side metadata. Returns RTP_OK, a negative status, or 1 = "not applicable, use the host
// path" (irregular boxes). `timing`: phase times on stderr.
int device_build_scene(const rtp_scene_desc* d, FlatScene* out, bool timing);
void device_free_arrays(DevArrays* a);
int device_scene_upload(const FlatScene& flat, int device /* < 0: the bound device */, DeviceScene** out);
void device_scene_free(DeviceScene* ds);
uint64_t device_scene_bytes(const DeviceScene* ds);

}  // namespace rtp

struct rtp_scene {
    rtp::FlatScene flat;  // nodes/prims/attrs are released after upload; leaf_order is kept
    rtp::DeviceScene* dev = nullptr;          // devs[0]: the device the *_device entry points use
    std::vector<rtp::DeviceScene*> devs;      // one replica per device of rtp_scene_create_multi's mask
    uint32_t n_leaves = 0, n_nodes = 0;
};
