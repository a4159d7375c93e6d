// staged.cu -- north-star kernel (1) on its own (batched patch gather, tile -> u8 patch array) and
// the adapters of the trait-level entry point.
//
//   k_gather     src/utils.rs:159-192  one TMA box load (zero fill = the reference's zero padding)
//                                      + one TMA box store per 64-pixel panel; the only SM work is
//                                      the rare fix-up of the trunc-toward-zero window quirk.
//   k_expand     bitmask -> u8 0/1 masks (parity tap for src/utils.rs:152-157)
//   k_pack_batch reference `Batch` layout (patchs [n,3,P,P] f32 = k/255, masks [n,1,P,P] f32;
//                src/utils.rs:17, 172, 198-199) -> u8 interleaved patch array + bitmask, so that
//                FeatureSet::compute_features_batched inputs run through the same kernels.
#include "nfx_kernels.h"

namespace nfx {

namespace {

__global__ void __launch_bounds__(32)
k_gather(const int64_t n, const int P, const NucInfo* __restrict__ info,
         const __grid_constant__ CUtensorMap map_tile, const __grid_constant__ CUtensorMap map_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const int64_t i = blockIdx.x;
    const NucInfo inf = info[i];
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, (uint32_t)patch_smem_bytes(P));
        tma_load_patch(smem_raw, &map_tile, inf.left, inf.top, P, &bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (inf.nvc < P || inf.nvr < P) {
        for (int k = threadIdx.x; k < P * P; k += 32) {
            const int r = k / P, c = k - r * P;
            if (r >= inf.nvr || c >= inf.nvc) {
                const int a = patch_addr(P, r, c);
                smem_raw[a] = 0; smem_raw[a + 1] = 0; smem_raw[a + 2] = 0;
            }
        }
        fence_proxy_async();   // generic-proxy writes -> visible to the TMA store
        __syncwarp();
    }
    if (threadIdx.x == 0) {
        const int np = patch_panels(P);
        for (int k = 0; k < np; ++k)
            tma_store_2d(&map_out, smem_raw + (size_t)k * kPanelBytes * P, k * kPanelBytes, (int32_t)(i * P));
        tma_store_commit();
        tma_store_wait_all();
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
                          const CUtensorMap* map_patches, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int smem = patch_smem_bytes(P);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    k_gather<<<(unsigned)n, 32, smem, s>>>(n, P, info, *map_tile, *map_patches);
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
