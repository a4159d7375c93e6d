// glcm.cu -- north-star kernel (5): GLCM texture. One CTA per nucleus builds the masked symmetric
// co-occurrence histograms in shared memory with atomics and derives the 14 Haralick features.
//
// Replaces GlcmFeatureSet::compute_features_batched (src/features/texture.rs:24-168):
//   grey = mean_dim(-3)              texture.rs:36        ((r+g)+b)/3 with IEEE f32 ops, bit-exact
//   tch_utils::glcm::glcm            texture.rs:40-46     oracle/SPEC.md B5
//   tch_utils::glcm::glcm_features   texture.rs:48-63     oracle/SPEC.md B6
// for levels {32,64,128,254} x offsets {(0,1),(1,1),(1,0),(1,-1)} (texture.rs:19-20).
//
// The L x L matrix is never scanned: a nucleus has only ~K co-occurring pixel pairs per offset
// (K = mask area), so everything is driven by the compacted pair list.
//   * G = C + C^T is kept as a TRIANGULAR u16 histogram (cell (min,max)), L(L+1)/2 entries, and is
//     only needed for the two features that are non-linear in p_ij (entropy, angular second moment):
//     sum_cells f(G) = sum_pairs 2 f(G_pair)/G_pair.
//   * every other feature is a moment of the three marginal histograms p_x, p_{x+y}, p_{x-y}
//     (HXY1 = HXY2 = 2 HX identically).
//   * after each (offset, level) the triangular histogram is cleared by replaying the pair list.
#include <math_constants.h>

#include "nfx_kernels.h"

namespace nfx {

namespace {

constexpr int kGlcmThreads = 256;
constexpr int kMaxLevels = 254;
constexpr int kTriEntries = kMaxLevels * (kMaxLevels + 1) / 2;   // 32385 u16
constexpr int kTriBytes = ((kTriEntries * 2 + 15) / 16) * 16;

__device__ __constant__ int c_levels[4] = {32, 64, 128, 254};
__device__ __constant__ int c_off[4][2] = {{0, 1}, {1, 1}, {1, 0}, {1, -1}};   // (dy, dx)

struct GlcmSmem {
    int region_a;   // patch (until quantised) aliased with the triangular histogram
    int rows, q128, q254, pairs, hist, total;
};
__host__ __device__ inline GlcmSmem glcm_layout(int P) {
    GlcmSmem L;
    int a = patch_smem_bytes(P) > kTriBytes ? patch_smem_bytes(P) : kTriBytes;
    a = (a + 127) & ~127;
    L.region_a = 0;
    L.rows = a;
    L.q128 = L.rows + ((P * mask_wpr(P) * 4 + 15) & ~15);
    L.q254 = L.q128 + P * P;
    L.pairs = L.q254 + P * P;
    L.hist = L.pairs + P * P * 2;
    L.total = L.hist + (256 + 512 + 256) * 4;
    return L;
}

__device__ __forceinline__ int quant_level(const uint8_t* q128, const uint8_t* q254, int pos, int lv) {
    // SPEC.md B5: q = min(floor(grey * L), L-1). x32/x64/x128 are exact scalings of the same f32 grey,
    // so floor(grey*64) == floor(grey*128) >> 1 (and >> 2 for 32); only 254 needs its own plane.
    switch (lv) {
        case 0: return min((int)q128[pos] >> 2, 31);
        case 1: return min((int)q128[pos] >> 1, 63);
        case 2: return min((int)q128[pos], 127);
        default: return (int)q254[pos];
    }
}

__global__ void __launch_bounds__(kGlcmThreads)
k_glcm(const GlcmParams p, const __grid_constant__ CUtensorMap map) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = p.P, wpr = mask_wpr(P), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t i = blockIdx.x;
    const GlcmSmem L = glcm_layout(P);
    uint8_t* patch = smem_raw + L.region_a;
    uint32_t* tri32 = reinterpret_cast<uint32_t*>(smem_raw + L.region_a);
    uint16_t* tri16 = reinterpret_cast<uint16_t*>(smem_raw + L.region_a);
    uint32_t* rows = reinterpret_cast<uint32_t*>(smem_raw + L.rows);
    uint8_t* q128 = smem_raw + L.q128;
    uint8_t* q254 = smem_raw + L.q254;
    uint16_t* pairs = reinterpret_cast<uint16_t*>(smem_raw + L.pairs);
    uint32_t* hx = reinterpret_cast<uint32_t*>(smem_raw + L.hist);   // [256]
    uint32_t* hs = hx + 256;                                           // [512]
    uint32_t* hd = hs + 512;                                           // [256]
    __shared__ __align__(8) uint64_t bar;
    __shared__ double s_red[12 * (kGlcmThreads / 32)];
    __shared__ float s_lut[256];
    __shared__ int s_scan[kGlcmThreads / 32 + 1];
    __shared__ int s_box[2];

    const NucInfo inf = p.info[i];
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        s_box[0] = P; s_box[1] = -1;
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, (uint32_t)patch_smem_bytes(P));
        tma_load_patch(patch, &map, inf.left, inf.top, P, &bar);
    }
    {
        const uint32_t* gm = p.bitmask + i * (int64_t)P * wpr;
        int rmin = P, rmax = -1;
        for (int k = tid; k < P * wpr; k += kGlcmThreads) {
            const uint32_t b = gm[k];
            rows[k] = b;
            if (b) { const int r = k / wpr; rmin = min(rmin, r); rmax = max(rmax, r); }
        }
        s_lut[tid] = __fdiv_rn((float)tid, 255.0f);   // utils.rs:172  u8 -> f32 / 255.0
        rmin = warp_min(rmin); rmax = warp_max(rmax);
        if (lane == 0) { atomicMin(&s_box[0], rmin); atomicMax(&s_box[1], rmax); }
    }
    __syncthreads();
    const int rmin = s_box[0], rmax = s_box[1];
    const int o = patch_byte_offset(inf.left);
    mbar_wait(&bar, 0);
    if (inf.nvc < P || inf.nvr < P) {
        for (int k = tid; k < P * P; k += kGlcmThreads) {
            const int r = k / P, c = k - r * P;
            if (r >= inf.nvr || c >= inf.nvc) {
                const int a = patch_addr(P, o, r, c);
                patch[a] = 0; patch[a + 1] = 0; patch[a + 2] = 0;
            }
        }
        __syncthreads();
    }
    // ---- grey quantisation planes for the mask's rows (texture.rs:36 + SPEC.md B5), bit-exact ----
    const bool dbg_all = (p.dbg_grey != nullptr);
    {
        const int r0 = dbg_all ? 0 : max(rmin, 0), r1 = dbg_all ? P - 1 : rmax;
        for (int k = r0 * P + tid; k < (r1 + 1) * P; k += kGlcmThreads) {
            const int r = k / P, c = k - r * P;
            const int a = patch_addr(P, o, r, c);
            const float g = __fdiv_rn(__fadd_rn(__fadd_rn(s_lut[patch[a]], s_lut[patch[a + 1]]), s_lut[patch[a + 2]]), 3.0f);
            q128[k] = (uint8_t)(int)floorf(__fmul_rn(g, 128.0f));
            q254[k] = (uint8_t)min((int)floorf(__fmul_rn(g, 254.0f)), 253);
        }
    }
    __syncthreads();   // patch is dead from here on: region A becomes the triangular histogram
    for (int k = tid; k < kTriBytes / 16; k += kGlcmThreads)
        reinterpret_cast<uint4*>(smem_raw + L.region_a)[k] = make_uint4(0, 0, 0, 0);
    if (dbg_all) {
        int lv = 3;
        for (int t = 0; t < 4; ++t) if (c_levels[t] == p.dbg_levels) lv = t;
        for (int k = tid; k < P * P; k += kGlcmThreads)
            p.dbg_grey[i * (int64_t)P * P + k] = (uint8_t)quant_level(q128, q254, k, lv);
    }

    float* out = p.out ? p.out + i * (int64_t)p.out_stride + p.col_glcm : nullptr;

    for (int oi = 0; oi < kGlcmOffsets; ++oi) {
        const int dy = c_off[oi][0], dx = c_off[oi][1];
        const int dpos = dy * P + dx;
        // ---- compact the pair list: source pixels p with mask[p] & mask[p + (dy,dx)] ----
        __syncthreads();
        int npairs = 0;
        {
            const int items = P * wpr;
            int running = 0;
            for (int base = 0; base < items; base += kGlcmThreads) {
                const int k = base + tid;
                uint32_t pb = 0;
                int r = 0, w = 0;
                if (k < items) {
                    r = k / wpr; w = k - r * wpr;
                    const int r2 = r + dy;
                    if (r >= rmin && r <= rmax && r2 < P) {
                        const uint32_t* nr = rows + r2 * wpr;
                        uint32_t nb = nr[w];
                        if (dx == 1) nb = (nb >> 1) | ((w + 1 < wpr) ? (nr[w + 1] << 31) : 0u);
                        else if (dx == -1) nb = (nb << 1) | ((w > 0) ? (nr[w - 1] >> 31) : 0u);
                        pb = rows[k] & nb;
                        // neighbour column must exist inside the patch when P is not a multiple of 32
                        if (dx == 1 && w == wpr - 1 && (P & 31)) pb &= (1u << ((P & 31) - 1)) - 1u;
                    }
                }
                const int cnt = __popc(pb);
                // block exclusive scan of cnt
                int incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (lane == 31) s_scan[warp] = incl;
                __syncthreads();
                if (tid == 0) {
                    int acc = 0;
                    for (int t = 0; t < kGlcmThreads / 32; ++t) { const int v = s_scan[t]; s_scan[t] = acc; acc += v; }
                    s_scan[kGlcmThreads / 32] = acc;
                }
                __syncthreads();
                int pos = running + s_scan[warp] + incl - cnt;
                while (pb) {
                    const int c = 32 * w + __ffs(pb) - 1;
                    pb &= pb - 1;
                    pairs[pos++] = (uint16_t)(r * P + c);
                }
                running += s_scan[kGlcmThreads / 32];
                __syncthreads();
            }
            npairs = running;
        }
        const double T = 2.0 * (double)npairs;

        for (int lv = 0; lv < kGlcmLevels; ++lv) {
            const int NL = c_levels[lv];
            for (int k = tid; k < 1024; k += kGlcmThreads) hx[k] = 0u;   // hx, hs, hd are contiguous
            __syncthreads();
            // ---- pass 1: shared-memory atomics ----
            for (int k = tid; k < npairs; k += kGlcmThreads) {
                const int pos = pairs[k];
                const int a = quant_level(q128, q254, pos, lv), b = quant_level(q128, q254, pos + dpos, lv);
                const int lo = min(a, b), hi = max(a, b);
                const int cell = ((hi * (hi + 1)) >> 1) + lo;
                atomicAdd(&tri32[cell >> 1], 1u << ((cell & 1) * 16));
                atomicAdd(&hx[a], 1u);
                atomicAdd(&hx[b], 1u);
                atomicAdd(&hs[a + b], 2u);
                atomicAdd(&hd[hi - lo], 2u);
            }
            __syncthreads();
            // ---- pass 2: entropy and ASM from the cell counts ----
            double acc[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) acc[k] = 0.0;
            {
                float sum_g = 0.f, sum_lng = 0.f;
                for (int k = tid; k < npairs; k += kGlcmThreads) {
                    const int pos = pairs[k];
                    const int a = quant_level(q128, q254, pos, lv), b = quant_level(q128, q254, pos + dpos, lv);
                    const int lo = min(a, b), hi = max(a, b);
                    const int cell = ((hi * (hi + 1)) >> 1) + lo;
                    const uint32_t g = (uint32_t)tri16[cell] << (a == b ? 1 : 0);   // G_ab
                    sum_g += (float)g;
                    sum_lng += logf((float)g);
                    if (p.dbg_counts && p.dbg_levels == NL && p.dbg_dy == dy && p.dbg_dx == dx) {
                        uint32_t* dc = p.dbg_counts + i * (int64_t)NL * NL;
                        dc[a * NL + b] = g;
                        dc[b * NL + a] = g;
                    }
                }
                acc[0] = sum_g;
                acc[1] = sum_lng;
            }
            // ---- marginal moments ----
            for (int k = tid; k < NL; k += kGlcmThreads) {
                const double c = (double)hx[k], kk = (double)k;
                acc[2] += kk * c;
                acc[3] += kk * kk * c;
                if (c > 0.0) acc[4] += c * log(c);
                const double d = (double)hd[k];
                acc[5] += kk * d;
                acc[6] += kk * kk * d;
                acc[7] += d / (1.0 + kk * kk);
            }
            for (int k = tid; k < 2 * NL - 1; k += kGlcmThreads) {
                const double c = (double)hs[k], kk = (double)k;
                acc[8] += kk * c;
                acc[9] += kk * kk * c;
                if (c > 0.0) acc[10] += c * log(c);
            }
            block_sum<12>(acc, s_red);
            // ---- clear the triangular histogram by replaying the pairs ----
            for (int k = tid; k < npairs; k += kGlcmThreads) {
                const int pos = pairs[k];
                const int a = quant_level(q128, q254, pos, lv), b = quant_level(q128, q254, pos + dpos, lv);
                const int lo = min(a, b), hi = max(a, b);
                tri16[((hi * (hi + 1)) >> 1) + lo] = 0;
            }
            if (tid == 0 && out) {
                float* o = out + (lv * kGlcmOffsets + oi) * kGlcmFeat;
                if (npairs == 0) {
                    for (int f = 0; f < kGlcmFeat; ++f) o[f] = CUDART_NAN_F;   // 0/0 (SPEC.md B5)
                } else {
                    const double lnT = log(T);
                    const double asm_ = 2.0 * acc[0] / (T * T);
                    const double hxy = lnT - 2.0 * acc[1] / T;
                    const double mu = acc[2] / T, ei2 = acc[3] / T;
                    const double var = ei2 - mu * mu;
                    const double hxm = lnT - acc[4] / T;
                    const double dav = acc[5] / T, contrast = acc[6] / T, idm = acc[7] / T;
                    const double sav = acc[8] / T, es2 = acc[9] / T;
                    const double sent = lnT - acc[10] / T;
                    const double eij = 0.5 * (es2 - 2.0 * ei2);
                    o[0] = (float)((eij - mu * mu) / var);            // correlation
                    o[1] = (float)contrast;
                    o[2] = (float)dav;                                // dissimilarity
                    o[3] = (float)hxy;                                // entropy
                    o[4] = (float)asm_;
                    o[5] = (float)sav;
                    o[6] = (float)(es2 - sav * sav);                  // sum variance
                    o[7] = (float)sent;
                    o[8] = (float)var;                                // sum of squares
                    o[9] = (float)idm;
                    o[10] = (float)dav;                               // difference average
                    o[11] = (float)(contrast - dav * dav);            // difference variance
                    o[12] = (float)((hxy - 2.0 * hxm) / hxm);         // IMC1 (HXY1 = 2 HX)
                    o[13] = (float)sqrt(fmax(1.0 - exp(-2.0 * (2.0 * hxm - hxy)), 0.0));   // IMC2
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace

cudaError_t launch_glcm(const GlcmParams& p, const CUtensorMap* map, cudaStream_t s) {
    if (p.n <= 0) return cudaSuccess;
    const GlcmSmem L = glcm_layout(p.P);
    cudaError_t e = cudaFuncSetAttribute(k_glcm, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return e;
    k_glcm<<<(unsigned)p.n, kGlcmThreads, L.total, s>>>(p, *map);
    return cudaGetLastError();
}

}  // namespace nfx
