// rtp_build.cu — device side of `Bvh::new` (bvh.rs:36-91): the reference's depth-first leaf order computed on the GPU.
//
// make_bvh (bvh.rs:36-56) sorts a range by the AABB-centroid key of one axis (bvh.rs:58-67), cuts it at len/2 and recurses
// with the next axis; the axis depends only on the depth. All ranges of one depth are therefore sorted by the same axis,
// and the whole recursion is a loop over depths, each depth one segmented sort of ALL leaves by (range, key, LeafId)
// (ties by LeafId: the deterministic order shared with the host build and the oracle). The segmented sort is
// three stable LSD passes — LeafId, then the 64-bit order-preserving image of the f64 key, then the range's start position —
// done with cub::DeviceRadixSort (CUDA toolkit; library code, used for this once-per-scene stage only). The kernels here
// compute keys and split the ranges. Output: item index by rank, bit-identical to the host Builder (tests/test_gpu_parity.py).
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>

#include <cstdint>
#include <string>

#include "rtp_internal.h"

namespace rtp {

#define RTP_CUDA_B(expr)                                                                             \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) { cleanup(); return set_error(RTP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } \
    } while (0)

// keys of one depth: the item's LeafId (== its index: items arrive in LeafId order) and the sortable image of its centroid key
__global__ void __launch_bounds__(256) build_keys_kernel(const double* __restrict__ boxes, const uint32_t* __restrict__ perm, uint32_t n, int axis,
                                                          unsigned long long* __restrict__ key64) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* b = boxes + static_cast<size_t>(perm[i]) * 6;
    double key = 0.5 * (b[axis] + b[3 + axis]);  // bvh.rs:61-62
    key = key + 0.0;                             // -0.0 and +0.0 compare equal in partial_cmp: one image for both
    unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(key));
    u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);  // monotone map f64 -> u64 (no NaN: rejected by the host validation)
    key64[i] = u;
}

__global__ void __launch_bounds__(256) build_segkeys_kernel(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ start_of_item, uint32_t n,
                                                             uint32_t* __restrict__ segkey) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) segkey[i] = start_of_item[perm[i]];
}

// bvh.rs:64: the left half gets len/2 items. Position i holds item perm[i] of the range [start, start+len).
__global__ void __launch_bounds__(256) build_split_kernel(const uint32_t* __restrict__ perm, uint32_t n, uint32_t* __restrict__ start_of_item,
                                                           uint32_t* __restrict__ len_of_item) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t item = perm[i];
    const uint32_t s = start_of_item[item], l = len_of_item[item];
    if (l <= 1) return;
    const uint32_t half = l / 2;
    if (i - s < half) {
        len_of_item[item] = half;
    } else {
        start_of_item[item] = s + half;
        len_of_item[item] = l - half;
    }
}

__global__ void __launch_bounds__(256) build_init_kernel(uint32_t n, uint32_t* __restrict__ perm, uint32_t* __restrict__ start_of_item,
                                                          uint32_t* __restrict__ len_of_item) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    perm[i] = i; start_of_item[i] = 0; len_of_item[i] = n;
}

int device_reference_order(const double* boxes, uint32_t n, uint32_t* order_out) {
    double* d_boxes = nullptr;
    uint32_t *d_perm[2] = {nullptr, nullptr}, *d_k32[2] = {nullptr, nullptr}, *d_start = nullptr, *d_len = nullptr;
    unsigned long long* d_k64[2] = {nullptr, nullptr};
    void* d_temp = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_boxes); cudaFree(d_perm[0]); cudaFree(d_perm[1]); cudaFree(d_k32[0]); cudaFree(d_k32[1]); cudaFree(d_start); cudaFree(d_len);
        cudaFree(d_k64[0]); cudaFree(d_k64[1]); cudaFree(d_temp);
    };
    if (n == 0) return RTP_OK;
    if (int rc = ensure_device()) return rc;
    RTP_CUDA_B(cudaMalloc(reinterpret_cast<void**>(&d_boxes), static_cast<size_t>(n) * 6 * sizeof(double)));
    RTP_CUDA_B(cudaMemcpy(d_boxes, boxes, static_cast<size_t>(n) * 6 * sizeof(double), cudaMemcpyHostToDevice));
    for (int k = 0; k < 2; ++k) {
        RTP_CUDA_B(cudaMalloc(reinterpret_cast<void**>(&d_perm[k]), static_cast<size_t>(n) * 4));
        RTP_CUDA_B(cudaMalloc(reinterpret_cast<void**>(&d_k32[k]), static_cast<size_t>(n) * 4));
        RTP_CUDA_B(cudaMalloc(reinterpret_cast<void**>(&d_k64[k]), static_cast<size_t>(n) * 8));
    }
    RTP_CUDA_B(cudaMalloc(reinterpret_cast<void**>(&d_start), static_cast<size_t>(n) * 4));
    RTP_CUDA_B(cudaMalloc(reinterpret_cast<void**>(&d_len), static_cast<size_t>(n) * 4));
    size_t temp_bytes = 0, need = 0;
    cub::DoubleBuffer<uint32_t> kb32(d_k32[0], d_k32[1]), vb(d_perm[0], d_perm[1]);
    cub::DoubleBuffer<unsigned long long> kb64(d_k64[0], d_k64[1]);
    RTP_CUDA_B(cub::DeviceRadixSort::SortPairs(nullptr, need, kb64, vb, static_cast<int>(n), 0, 64));
    temp_bytes = need;
    RTP_CUDA_B(cub::DeviceRadixSort::SortPairs(nullptr, need, kb32, vb, static_cast<int>(n), 0, 32));
    temp_bytes = need > temp_bytes ? need : temp_bytes;
    RTP_CUDA_B(cudaMalloc(&d_temp, temp_bytes));

    const unsigned grid = (n + 255) / 256;
    build_init_kernel<<<grid, 256>>>(n, vb.Current(), d_start, d_len);
    RTP_CUDA_B(cudaGetLastError());
    int seg_bits = 1;
    while ((1ull << seg_bits) < n) ++seg_bits;  // range start positions are < n
    int axis = 0;
    for (uint64_t longest = n; longest > 1; longest = (longest + 1) / 2, axis = (axis + 1) % 3) {
        // pass 1 (least significant): LeafId. The value array IS the id, so sort it as keys carrying itself.
        RTP_CUDA_B(cudaMemcpyAsync(kb32.Current(), vb.Current(), static_cast<size_t>(n) * 4, cudaMemcpyDeviceToDevice));
        RTP_CUDA_B(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, kb32, vb, static_cast<int>(n), 0, seg_bits));
        // pass 2: centroid key of this depth's axis
        build_keys_kernel<<<grid, 256>>>(d_boxes, vb.Current(), n, axis, kb64.Current());
        RTP_CUDA_B(cudaGetLastError());
        RTP_CUDA_B(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, kb64, vb, static_cast<int>(n), 0, 64));
        // pass 3 (most significant): the range each item belongs to
        build_segkeys_kernel<<<grid, 256>>>(vb.Current(), d_start, n, kb32.Current());
        RTP_CUDA_B(cudaGetLastError());
        RTP_CUDA_B(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, kb32, vb, static_cast<int>(n), 0, seg_bits));
        build_split_kernel<<<grid, 256>>>(vb.Current(), n, d_start, d_len);
        RTP_CUDA_B(cudaGetLastError());
    }
    RTP_CUDA_B(cudaMemcpy(order_out, vb.Current(), static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost));
    cleanup();
    return RTP_OK;
}

}  // namespace rtp
