// staged.cu -- north-star kernel (1) on its own (batched patch gather, tile -> u8 patch array) and
// the adapters of the trait-level entry point.
//
//   k_gather     src/utils.rs:159-192  TMA box load (zero fill = the reference's zero padding) of the
//                                      16-byte aligned superset of the window, re-aligned with
//                                      funnel shifts and written with coalesced 128-bit stores.
//   k_expand     bitmask -> u8 0/1 masks (parity tap for src/utils.rs:152-157)
//   k_pack_batch reference `Batch` layout (patchs [n,3,P,P] f32 = k/255, masks [n,1,P,P] f32;
//                src/utils.rs:17, 172, 198-199) -> u8 interleaved patch array + bitmask, so that
//                FeatureSet::compute_features_batched inputs run through the same kernels.
#include "nfx_kernels.h"

namespace nfx {

namespace {

constexpr int kGatherThreads = 128;

// One CTA per nucleus: TMA box load(s) of the 16-byte-aligned superset of the window (zero fill
// outside the tile), then every thread re-aligns 16 output bytes with funnel shifts and writes them
// with one 128-bit store (output rows are 16-byte aligned, 3P bytes of payload each).
__global__ void __launch_bounds__(kGatherThreads)
k_gather(const int64_t n, const int P, const NucInfo* __restrict__ info,
         const __grid_constant__ CUtensorMap map_tile, uint8_t* __restrict__ out, const int64_t pitch) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const int64_t i = blockIdx.x;
    const NucInfo inf = info[i];
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, (uint32_t)(patch_panels(P) * kPanelBytes * P));
        tma_load_patch(smem_raw, &map_tile, inf.left, inf.top, P, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    const int o = patch_byte_offset(inf.left);
    const int vecs_per_row = (int)(pitch >> 4);   // 16-byte vectors per output row
    const int row_bytes = 3 * P;
    for (int k = threadIdx.x; k < P * vecs_per_row; k += kGatherThreads) {
        const int r = k / vecs_per_row, v = k - r * vecs_per_row;
        const int ob = v * 16;                                    // first payload byte of this vector
        // source: panel (ob / 192), byte (ob % 192) + o within the 208-byte panel row
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int b = ob + 4 * q;
            uint32_t val = 0u;
            if (b < row_bytes) {
                const int pn = b / kPanelData, within = b - pn * kPanelData;
                const int a = pn * panel_stride(P) + r * kPanelBytes + o + within;
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(smem_raw + (a & ~3));
                val = __funnelshift_r(wp[0], wp[1], (a & 3) * 8);
                // utils.rs:161-192: columns/rows the reference never copies stay zero
                const int c0 = b / 3;   // first pixel touched by this word
                if (r >= inf.nvr) val = 0u;
                else if (inf.nvc < P) {
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        if ((b + t) / 3 >= inf.nvc) val &= ~(0xffu << (8 * t));
                }
                (void)c0;
                if (b + 4 > row_bytes) val &= (1u << (8 * (row_bytes - b))) - 1u;   // pitch padding
            }
            w[q] = val;
        }
        *reinterpret_cast<uint4*>(out + (i * P + r) * pitch + ob) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

__global__ void k_expand(const int64_t total_px, const int P, const uint32_t* __restrict__ bitmask,
                         uint8_t* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total_px) return;
    const int wpr = mask_wpr(P);
    const int64_t i = k / (P * P);
    const int rem = (int)(k - i * P * P), r = rem / P, c = rem - r * P;
    out[k] = (bitmask[(i * P + r) * wpr + (c >> 5)] >> (c & 31)) & 1u;
}

// one warp per (nucleus, row, 32-column word)
__global__ void __launch_bounds__(256)
k_pack_batch(const int64_t n, const int P, const float* __restrict__ patchs,
             const float* __restrict__ masks, uint8_t* __restrict__ out, const int64_t pitch,
             uint32_t* __restrict__ bitmask, NucInfo* __restrict__ info, int* __restrict__ bad) {
    const int wpr = mask_wpr(P), lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= n * P * wpr) return;
    const int64_t i = gw / (P * wpr);
    const int rem = (int)(gw - i * P * wpr), r = rem / wpr, w = rem - r * wpr, c = 32 * w + lane;
    bool bit = false;
    if (c < P) {
        const int64_t plane = (int64_t)P * P, px = (int64_t)r * P + c;
        uint8_t* o = out + (i * P + r) * pitch + 3 * c;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float v = patchs[(i * 3 + ch) * plane + px] * 255.0f;
            const float u = rintf(v);
            if (!(fabsf(v - u) <= 1e-3f) || u < 0.f || u > 255.f) atomicAdd(bad, 1);
            o[ch] = (uint8_t)fminf(fmaxf(u, 0.f), 255.f);
        }
        bit = masks[i * plane + px] != 0.f;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, bit);
    if (lane == 0) bitmask[(i * P + r) * wpr + w] = word;
    if (rem == 0 && lane == 0) {
        NucInfo inf;
        inf.left = 0;
        inf.top = (int32_t)(i * P);
        inf.nvc = P;
        inf.nvr = P;
        info[i] = inf;
    }
}

}  // namespace

cudaError_t launch_gather(int64_t n, int P, const NucInfo* info, const CUtensorMap* map_tile,
                          uint8_t* patches, int64_t pitch, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int smem = patch_smem_bytes(P) + 16;   // the re-alignment reads one word past the payload
    if (smem > 32 * 1024) {   // static shared memory counts towards the 48 KB default limit too
        cudaError_t e = cudaFuncSetAttribute(k_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    k_gather<<<(unsigned)n, kGatherThreads, smem, s>>>(n, P, info, *map_tile, patches, pitch);
    return cudaGetLastError();
}

cudaError_t launch_expand_mask(int64_t n, int P, const uint32_t* bitmask, uint8_t* out, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int64_t total = n * P * P;
    k_expand<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(total, P, bitmask, out);
    return cudaGetLastError();
}

cudaError_t launch_pack_batch(int64_t n, int P, const float* patchs, const float* masks,
                              uint8_t* patches_u8, int64_t pitch, uint32_t* bitmask, NucInfo* info,
                              int* bad_count, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int64_t warps = n * P * mask_wpr(P);
    k_pack_batch<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>(n, P, patchs, masks, patches_u8,
                                                                     pitch, bitmask, info, bad_count);
    return cudaGetLastError();
}

}  // namespace nfx
