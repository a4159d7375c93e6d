// Device-side helpers shared by every kernel: mbarrier + TMA (cp.async.bulk.tensor) wrappers,
// warp/block reductions, and the per-nucleus record the kernels exchange through HBM.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nfx {

// One record per nucleus, written by the geometry/raster kernel and read by every consumer.
// (left, top) = patch origin (utils.rs:159-160, f32 arithmetic, trunc toward zero) minus the tile
// origin. The reference copies image rows max(top,0)..min(bottom,H) to patch rows starting at
// -min(top,0) (utils.rs:161-192): patch row pr shows image row top+pr iff that row exists AND
// pr < bottom-top. TMA's zero fill covers the first condition, (nvc, nvr) the second (bottom-top is
// P-1 when cy-P/2 < 0 < cy+P/2 has a fractional part, because both casts truncate toward zero).
struct NucInfo {
    int32_t left, top;     // window origin in TILE coordinates (may be negative / beyond the tile)
    int32_t nvc, nvr;      // min(right-left, P), min(bottom-top, P): patch columns/rows >= these are 0
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    // make the init visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMA: 2D tiled bulk tensor copy global -> shared, completion on an mbarrier ---------------
// Coordinates are signed element indices; out-of-bounds elements are filled with zeros, which is
// exactly the reference's zero padding at image borders (utils.rs:174-192).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int32_t x,
                                            int32_t y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
// shared -> global 2D store (bulk group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int32_t x,
                                             int32_t y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::
                     "l"(reinterpret_cast<uint64_t>(map)),
                 "r"(x), "r"(y), "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---- reductions -------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Reduce NV values per thread over the whole CTA. `scratch` needs NV * (blockDim/32) elements.
// Result valid in every thread of warp 0 (and broadcast through scratch[0..NV) after the call).
template <int NV, typename T>
__device__ __forceinline__ void block_sum(T (&v)[NV], T* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();   // scratch may still be read from a previous call
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) scratch[warp * NV + k] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            T s = (lane < nw) ? scratch[lane * NV + k] : T(0);
            s = warp_sum(s);
            v[k] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) scratch[k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = scratch[k];
}

// Patch layout in shared memory.
// MEASURED on B200 (scripts/tma_probe.cu): a tiled u8 TMA load faults ("illegal instruction") unless
// the innermost box coordinate is a multiple of 16 BYTES (negative / out-of-bounds coordinates are
// fine and zero-filled). A nucleus window starts at byte 3*left of the tile row, so each box starts
// at the 16-byte boundary below it and is 16 bytes wider: rows are 208 B (= 13 x 16) for a 64-pixel
// panel and the pixel data begins `o = (3*left) & 15` bytes into the row.
//   addr(r, c) = (c >> 6) * 208 * P + r * 208 + o + (c & 63) * 3
constexpr int kPanelPx = 64;
constexpr int kPanelData = 192;    // payload bytes of a panel row
constexpr int kPanelBytes = 208;   // TMA box width (payload + alignment slack)
__host__ __device__ __forceinline__ int patch_panels(int P) { return (P + kPanelPx - 1) / kPanelPx; }
// bytes between consecutive panels of a `rows`-row window (TMA destinations must stay 128-byte aligned)
__host__ __device__ __forceinline__ int panel_stride(int rows) { return (kPanelBytes * rows + 127) & ~127; }
__host__ __device__ __forceinline__ int window_smem_bytes(int P, int rows) { return patch_panels(P) * panel_stride(rows); }
__host__ __device__ __forceinline__ int patch_smem_bytes(int P) { return window_smem_bytes(P, P); }
__host__ __device__ __forceinline__ int patch_byte_offset(int left) { return (3 * left) & 15; }
// `rows` = rows per panel of the window in shared memory (P for a whole patch, the slab height otherwise)
__device__ __forceinline__ int patch_addr(int rows, int o, int r, int c) {
    return (c >> 6) * panel_stride(rows) + r * kPanelBytes + o + (c & 63) * 3;
}
// 32-bit words per bitmask row
__host__ __device__ __forceinline__ int mask_wpr(int P) { return (P + 31) / 32; }

// Issue the TMA boxes ({208 B, rows} each) that bring `rows` rows of one nucleus window into shared
// memory, panel k at smem + k*208*rows. ONE thread, after mbar_expect_tx(bar, panels*208*rows) -- the
// transaction count is the box payload, not the padded stride.
__device__ __forceinline__ void tma_load_window(uint8_t* smem, const CUtensorMap* map, int left, int top,
                                                int P, int rows, uint64_t* bar) {
    const int np = patch_panels(P);
    const int xal = ((3 * left) >> 4) << 4;   // arithmetic shift: floor for negative origins
    for (int k = 0; k < np; ++k)
        tma_load_2d(smem + (size_t)k * panel_stride(rows), map, xal + k * kPanelData, top, bar);
}
__device__ __forceinline__ void tma_load_patch(uint8_t* smem_patch, const CUtensorMap* map, int left,
                                               int top, int P, uint64_t* bar) {
    tma_load_window(smem_patch, map, left, top, P, P, bar);
}

// Four consecutive pixels (12 bytes) starting at byte address `b` of shared memory (any alignment).
__device__ __forceinline__ void load_quad(const uint8_t* smem, int b, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(smem + (b & ~3));
    const uint32_t sh = (b & 3) * 8;
    const uint32_t a0 = wp[0], a1 = wp[1], a2 = wp[2], a3 = wp[3];
    w0 = __funnelshift_r(a0, a1, sh);
    w1 = __funnelshift_r(a1, a2, sh);
    w2 = __funnelshift_r(a2, a3, sh);
}

}  // namespace nfx
