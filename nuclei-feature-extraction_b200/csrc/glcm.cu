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
// (K = mask area), so everything is driven by the compacted pair list (one packed u32 per pair).
//   * G = C + C^T is kept as a TRIANGULAR u16 histogram (cell (min,max)), L(L+1)/2 entries, and is
//     only needed for the two features that are non-linear in p_ij (entropy, angular second moment):
//     sum_cells f(G) = sum_pairs 2 f(G_pair)/G_pair.
//   * every other feature is a moment of the three marginal histograms p_x, p_{x+y}, p_{x-y}
//     (HXY1 = HXY2 = 2 HX identically); the moments are exact integer sums (REDUX).
//   * after each (offset, level) the triangular histogram is cleared by replaying the pair list.
//   * per-warp partial sums of all 16 combinations are parked in shared memory; the 14 features of
//     each combination are computed once at the end by one thread per combination.
#include <math_constants.h>

#include <type_traits>

#include "nfx_kernels.h"

namespace nfx {

namespace {

constexpr int kGlcmThreads = 256;
constexpr int kMaxLevels = 254;
constexpr int kTriEntries = kMaxLevels * (kMaxLevels + 1) / 2;   // 32385 u16
constexpr int kTriBytes = ((kTriEntries * 2 + 15) / 16) * 16;

__device__ __constant__ int c_levels[4] = {32, 64, 128, 254};
__device__ __constant__ int c_off[4][2] = {{0, 1}, {1, 1}, {1, 0}, {1, -1}};   // (dy, dx)

constexpr int kCombos = kGlcmLevels * kGlcmOffsets;   // 16
constexpr int kNI = 7, kNF = 4;                        // integer / float partial sums per combination
constexpr int kNW = kGlcmThreads / 32;

struct GlcmSmem {
    int region_a;   // patch (until quantised) aliased with the triangular histogram
    int rows, q128, q254, pairs, hist, part_i, part_f, total;
};
__host__ __device__ inline GlcmSmem glcm_layout(int P) {
    GlcmSmem L;
    int a = patch_smem_bytes(P) > kTriBytes ? patch_smem_bytes(P) : kTriBytes;
    a = (a + 127) & ~127;
    L.region_a = 0;
    L.rows = a;
    L.q128 = L.rows + ((P * mask_wpr(P) * 4 + 15) & ~15);
    const int plane = (P * P + 15) & ~15;           // odd patch sizes: keep every array 16-byte aligned
    L.q254 = L.q128 + plane;
    L.pairs = L.q254 + plane;                       // u32 per pair: a254 | b254<<8 | a128<<16 | b128<<24
    L.hist = L.pairs + ((P * P * 4 + 15) & ~15);
    L.part_i = L.hist + (256 + 512 + 256) * 4;
    L.part_f = L.part_i + kCombos * kNW * kNI * 8;  // u64
    L.total = L.part_f + kCombos * kNW * kNF * 4;
    return L;
}

// Levels of one pair for level index lv from its packed record (SPEC.md B5: q = min(floor(g*L), L-1)).
// x32/x64/x128 are exact scalings of the same f32 grey, so floor(g*64) = floor(g*128) >> 1 and
// floor(g*32) = floor(g*128) >> 2; the q128 plane is stored pre-clamped to 127, which commutes with
// the shifts (128 only occurs for g == 1). Only 254 needs its own plane.
__device__ __forceinline__ void pair_levels(uint32_t rec, int lv, int& a, int& b) {
    if (lv == 3) {
        a = rec & 0xff;
        b = (rec >> 8) & 0xff;
    } else {
        const int sh = 2 - lv;
        a = ((rec >> 16) & 0xff) >> sh;
        b = (rec >> 24) >> sh;
    }
}

// The 14 Haralick features (SPEC.md B6) from the sufficient statistics of one symmetric GLCM.
//   T = sum G = 2*pairs, sg = sum_pairs G_pair, slg = sum_pairs ln G_pair,
//   x1,x2 = sum k hx, sum k^2 hx ; d1,d2 = same for the difference histogram ; s1,s2 = sum histogram,
//   cx, cs = sum c ln c over hx / hs ; idm_sum = sum hd[k] / (1 + k^2).
__device__ __forceinline__ void haralick_write(float* o_, double T, double sg, double slg, double x1, double x2,
                                               double d1, double d2, double s1, double s2, double cx, double cs,
                                               double idm_sum) {
    const double lnT = log(T);
    const double asm_ = 2.0 * sg / (T * T);
    // Entropies are >= 0; the logarithmic sums arrive in fixed point (2^-16 quantum), so a matrix with ONE occupied
    // cell (asm == 1 exactly: the moments are exact integers) is pinned to its exact entropy 0 (-> IMC1 = 0/0 = NaN).
    const bool one_cell = asm_ == 1.0;
    const double hxy = one_cell ? 0.0 : fmax(lnT - 2.0 * slg / T, 0.0);
    const double mu = x1 / T, ei2 = x2 / T;
    const double var = ei2 - mu * mu;
    const double hxm = one_cell ? 0.0 : fmax(lnT - cx / T, 0.0);
    const double dav = d1 / T, contrast = d2 / T, idm = idm_sum / T;
    const double sav = s1 / T, es2 = s2 / T;
    const double sent = one_cell ? 0.0 : fmax(lnT - cs / T, 0.0);
    const double eij = 0.5 * (es2 - 2.0 * ei2);
    o_[0] = (float)((eij - mu * mu) / var);            // correlation
    o_[1] = (float)contrast;
    o_[2] = (float)dav;                                // dissimilarity
    o_[3] = (float)hxy;                                // entropy
    o_[4] = (float)asm_;
    o_[5] = (float)sav;
    o_[6] = (float)(es2 - sav * sav);                  // sum variance
    o_[7] = (float)sent;
    o_[8] = (float)var;                                // sum of squares
    o_[9] = (float)idm;
    o_[10] = (float)dav;                               // difference average
    o_[11] = (float)(contrast - dav * dav);            // difference variance
    o_[12] = (float)((hxy - 2.0 * hxm) / hxm);         // IMC1 (HXY1 = 2 HX)
    o_[13] = (float)sqrt(fmax(1.0 - exp(-2.0 * (2.0 * hxm - hxy)), 0.0));   // IMC2
}

template <bool SMALL>   // generic path (64 < P <= 128 in practice): one (offset, level) at a time
__global__ void __launch_bounds__(kGlcmThreads)
k_glcm_generic(const GlcmParams p, const __grid_constant__ CUtensorMap map) {
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
    uint32_t* pairs = reinterpret_cast<uint32_t*>(smem_raw + L.pairs);
    uint32_t* hx = reinterpret_cast<uint32_t*>(smem_raw + L.hist);   // [256]
    uint32_t* hs = hx + 256;                                           // [512]
    uint32_t* hd = hs + 512;                                           // [256]
    unsigned long long* part_i = reinterpret_cast<unsigned long long*>(smem_raw + L.part_i);   // [combo][warp][kNI]
    float* part_f = reinterpret_cast<float*>(smem_raw + L.part_f);                               // [combo][warp][kNF]
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_lut[256];
    __shared__ int s_scan[kNW + 1];
    __shared__ int s_box[2];
    __shared__ int s_npairs[kGlcmOffsets];

    const NucInfo inf = p.info[i];
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        s_box[0] = P;
        s_box[1] = -1;
    }
    __syncthreads();
    const float* greyp = p.grey ? p.grey + i * (int64_t)P * P : nullptr;   // f32 grey plane instead of the u8 window
    if (tid == 0 && !greyp) {
        mbar_expect_tx(&bar, (uint32_t)(patch_panels(P) * kPanelBytes * P));
        tma_load_patch(patch, &map, inf.left, inf.top, P, &bar);
    }
    {
        const uint32_t* gm = p.bitmask + i * (int64_t)P * wpr;
        int rmin = P, rmax = -1;
        for (int k = tid; k < P * wpr; k += kGlcmThreads) {
            const uint32_t b = gm[k];
            rows[k] = b;
            if (b) {
                const int r = k / wpr;
                rmin = min(rmin, r);
                rmax = max(rmax, r);
            }
        }
        s_lut[tid] = __fdiv_rn((float)tid, 255.0f);   // utils.rs:172  u8 -> f32 / 255.0
        rmin = warp_min(rmin);
        rmax = warp_max(rmax);
        if (lane == 0) {
            atomicMin(&s_box[0], rmin);
            atomicMax(&s_box[1], rmax);
        }
    }
    __syncthreads();
    const int rmin = s_box[0], rmax = s_box[1];
    const int o = patch_byte_offset(inf.left);
    if (!greyp) mbar_wait(&bar, 0);
    if (!greyp && (inf.nvc < P || inf.nvr < P)) {
        for (int k = tid; k < P * P; k += kGlcmThreads) {
            const int r = k / P, c = k - r * P;
            if (r >= inf.nvr || c >= inf.nvc) {
                const int a = patch_addr(P, o, r, c);
                patch[a] = 0;
                patch[a + 1] = 0;
                patch[a + 2] = 0;
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
            const float g = greyp ? greyp[k] : __fdiv_rn(
                __fadd_rn(__fadd_rn(s_lut[patch[a]], s_lut[patch[a + 1]]), s_lut[patch[a + 2]]), 3.0f);
            q128[k] = (uint8_t)min((int)floorf(__fmul_rn(g, 128.0f)), 127);
            q254[k] = (uint8_t)min((int)floorf(__fmul_rn(g, p.scale254)), 253);
        }
    }
    __syncthreads();   // patch is dead from here on: region A becomes the triangular histogram
    for (int k = tid; k < kTriBytes / 16; k += kGlcmThreads)
        reinterpret_cast<uint4*>(smem_raw + L.region_a)[k] = make_uint4(0, 0, 0, 0);
    for (int k = tid; k < 1024; k += kGlcmThreads) hx[k] = 0u;   // hx, hs, hd are contiguous
    if (dbg_all) {
        int lv = 3;
        for (int t = 0; t < 4; ++t)
            if (c_levels[t] == p.dbg_levels) lv = t;
        for (int k = tid; k < P * P; k += kGlcmThreads) {
            const int v = (lv == 3) ? q254[k] : (q128[k] >> (2 - lv));
            p.dbg_grey[i * (int64_t)P * P + k] = (uint8_t)v;
        }
    }

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
                    r = k / wpr;
                    w = k - r * wpr;
                    const int r2 = r + dy;
                    if (r >= rmin && r <= rmax && r2 < P) {
                        const uint32_t* nr = rows + r2 * wpr;
                        uint32_t nb = nr[w];
                        if (dx == 1) nb = (nb >> 1) | ((w + 1 < wpr) ? (nr[w + 1] << 31) : 0u);
                        else if (dx == -1) nb = (nb << 1) | ((w > 0) ? (nr[w - 1] >> 31) : 0u);
                        pb = rows[k] & nb;   // bits >= P are never set, so column P-1 has no right neighbour
                    }
                }
                const int cnt = __popc(pb);
                int incl = cnt;
#pragma unroll
                for (int o2 = 1; o2 < 32; o2 <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o2);
                    if (lane >= o2) incl += t;
                }
                if (lane == 31) s_scan[warp] = incl;
                __syncthreads();
                int wbase = 0, total = 0;
#pragma unroll
                for (int t = 0; t < kNW; ++t) {
                    const int v = s_scan[t];
                    wbase += (t < warp) ? v : 0;
                    total += v;
                }
                int pos = running + wbase + incl - cnt;
                while (pb) {
                    const int c = 32 * w + __ffs(pb) - 1;
                    pb &= pb - 1;
                    const int src = r * P + c, dst = src + dpos;
                    pairs[pos++] = (uint32_t)q254[src] | ((uint32_t)q254[dst] << 8) |
                                   ((uint32_t)q128[src] << 16) | ((uint32_t)q128[dst] << 24);
                }
                running += total;
                __syncthreads();
            }
            npairs = running;
        }
        if (tid == 0) s_npairs[oi] = npairs;

        for (int lv = 0; lv < kGlcmLevels; ++lv) {
            const int NL = c_levels[lv];
            const int combo = lv * kGlcmOffsets + oi;
            // ---- pass 1: shared-memory atomics (tri + marginal histograms are all-zero here) ----
            for (int k = tid; k < npairs; k += kGlcmThreads) {
                int a, b;
                pair_levels(pairs[k], lv, a, b);
                const int lo = min(a, b), hi = max(a, b);
                const int cell = ((hi * (hi + 1)) >> 1) + lo;
                atomicAdd(&tri32[cell >> 1], 1u << ((cell & 1) * 16));
                atomicAdd(&hx[a], 1u);
                atomicAdd(&hx[b], 1u);
                atomicAdd(&hs[a + b], 2u);
                atomicAdd(&hd[hi - lo], 2u);
            }
            __syncthreads();
            // ---- pass 2: entropy / ASM from the cell counts; moments of the marginals ----
            uint32_t vi[kNI] = {0, 0, 0, 0, 0, 0, 0};
            unsigned long long vl[kNI] = {0, 0, 0, 0, 0, 0, 0};
            float vf[kNF] = {0.f, 0.f, 0.f, 0.f};
            const bool dbg = p.dbg_counts && p.dbg_levels == NL && p.dbg_dy == dy && p.dbg_dx == dx;
            for (int k = tid; k < npairs; k += kGlcmThreads) {
                int a, b;
                pair_levels(pairs[k], lv, a, b);
                const int lo = min(a, b), hi = max(a, b);
                const uint32_t g = (uint32_t)tri16[((hi * (hi + 1)) >> 1) + lo] << (a == b ? 1 : 0);   // G_ab
                if (SMALL) vi[0] += g;
                else vl[0] += g;
                vf[0] += __logf((float)g);
                if (dbg) {
                    uint32_t* dc = p.dbg_counts + i * (int64_t)NL * NL;
                    dc[a * NL + b] = g;
                    dc[b * NL + a] = g;
                }
            }
            for (int k = tid; k < 2 * NL - 1; k += kGlcmThreads) {
                const uint32_t c = hs[k], kk = (uint32_t)k;
                if (SMALL) {
                    vi[5] += kk * c;
                    vi[6] += kk * kk * c;
                } else {
                    vl[5] += (unsigned long long)kk * c;
                    vl[6] += (unsigned long long)kk * kk * c;
                }
                if (c) vf[2] += (float)c * __logf((float)c);
                if (k < NL) {
                    const uint32_t cx = hx[k], cd = hd[k];
                    if (SMALL) {
                        vi[1] += kk * cx;
                        vi[2] += kk * kk * cx;
                        vi[3] += kk * cd;
                        vi[4] += kk * kk * cd;
                    } else {
                        vl[1] += (unsigned long long)kk * cx;
                        vl[2] += (unsigned long long)kk * kk * cx;
                        vl[3] += (unsigned long long)kk * cd;
                        vl[4] += (unsigned long long)kk * kk * cd;
                    }
                    if (cx) vf[1] += (float)cx * __logf((float)cx);
                    vf[3] += __fdividef((float)cd, 1.0f + (float)(kk * kk));
                }
            }
#pragma unroll
            for (int q = 0; q < kNI; ++q) {
                unsigned long long t;
                if (SMALL) t = __reduce_add_sync(0xffffffffu, vi[q]);
                else t = warp_sum(vl[q]);
                if (lane == 0) part_i[(combo * kNW + warp) * kNI + q] = t;
            }
#pragma unroll
            for (int q = 0; q < kNF; ++q) {
                const float t = warp_sum(vf[q]);
                if (lane == 0) part_f[(combo * kNW + warp) * kNF + q] = t;
            }
            __syncthreads();
            // ---- clear: triangular histogram by replaying the pairs, marginals densely ----
            for (int k = tid; k < npairs; k += kGlcmThreads) {
                int a, b;
                pair_levels(pairs[k], lv, a, b);
                const int lo = min(a, b), hi = max(a, b);
                tri16[((hi * (hi + 1)) >> 1) + lo] = 0;
            }
            for (int k = tid; k < 1024; k += kGlcmThreads) hx[k] = 0u;
            __syncthreads();
        }
    }
    // ---- 14 Haralick features per (level, offset): one thread per combination ----
    if (tid < kCombos && p.out) {
        const int combo = tid, oi = combo % kGlcmOffsets;
        float* o_ = p.out + i * (int64_t)p.out_stride + p.col_glcm + combo * kGlcmFeat;
        const int npairs = s_npairs[oi];
        if (npairs == 0) {
            for (int f = 0; f < kGlcmFeat; ++f) o_[f] = CUDART_NAN_F;   // 0/0 (SPEC.md B5)
        } else {
            double acc[kNI + kNF];
            for (int q = 0; q < kNI; ++q) {
                unsigned long long t = 0;
                for (int w = 0; w < kNW; ++w) t += part_i[(combo * kNW + w) * kNI + q];
                acc[q] = (double)t;
            }
            for (int q = 0; q < kNF; ++q) {
                double t = 0.0;
                for (int w = 0; w < kNW; ++w) t += (double)part_f[(combo * kNW + w) * kNF + q];
                acc[kNI + q] = t;
            }
            haralick_write(o_, 2.0 * (double)npairs, acc[0], acc[kNI + 0], acc[1], acc[2], acc[3], acc[4], acc[5], acc[6],
                           acc[kNI + 1], acc[kNI + 2], acc[kNI + 3]);
        }
    }
}


// =================================================================================================
// k_glcm64: P <= 64. All four levels of one offset are processed in ONE sweep over the pairs:
//   * the masked pixels are compacted once into a coordinate list (shared by the 4 offsets);
//   * pass 1 walks that list, tests the neighbour bit and issues the shared-memory atomics of all 4
//     levels (13 per pair); pass 2 walks it again for the cell counts. No pair list is kept: with a
//     4096-slot hash that brings the CTA to 69 KB of shared memory = 3 CTAs per SM (the kernel is
//     bound by shared-memory latency, so residency matters more than the re-derived pair);
//   * G for 32/64/128 levels lives in dense triangular u16 histograms (22 KB together); G for 254
//     levels (32 385 cells, <= K occupied) lives in an open-addressing hash table of 4096 slots
//     (linear probing; a 64 x 64 mask has at most 4032 pairs, so it can never fill up);
//   * p_x of 32 and 64 levels is the 128-level histogram folded by 4 / 2 (exact: q32 = q128 >> 2);
//   * p_{x-y} is never built: its three moments are linear in the pairs and live in registers;
//   * every partial sum is an integer (moments exactly, logarithmic terms in fixed point with a
//     2^-16 quantum, far below the 1e-4 tolerance) so that warp reduction is one REDUX each;
//   * tables are cleared densely with 128-bit stores. 3 block barriers per offset.
// =================================================================================================
constexpr int kG64Threads = 256, kG64NW = kG64Threads / 32;   // 384 / 512 threads measured slower (17.4 / 22.6 vs 16.2 / 20.2 ms per 200k)
constexpr int kTri32 = 32 * 33 / 2, kTri64 = 64 * 65 / 2, kTri128 = 128 * 129 / 2;
constexpr int kOffTri128 = 0;                                  // bytes inside region A
constexpr int kOffTri64 = kOffTri128 + kTri128 * 2;            // 16512
constexpr int kOffTri32 = kOffTri64 + kTri64 * 2;              // 20672
constexpr int kOffHash = ((kOffTri32 + kTri32 * 2 + 127) / 128) * 128;   // 21760
constexpr int kHashMax = 4096, kHashLg = 12;
constexpr int kRegionA64 = kOffHash + kHashMax * 4;            // 38144
constexpr int kMargWords = 2048;                               // hx128 @0, hx254 @128, hs128 @384, hs254 @640, hs64 x3 @1152, hs32 x8 @1536
constexpr int kNP = 11;                                        // partial sums per (level, offset)
constexpr float kLnFix = 0.6931471805599453f * 65536.0f;      // log2 -> ln, 16 fractional bits

struct Glcm64Smem {
    int rows, q128, q254, list, marg, parts, total;
};
__host__ __device__ inline Glcm64Smem glcm64_layout(int P) {
    Glcm64Smem L;
    L.marg = kRegionA64;
    L.rows = L.marg + kMargWords * 4;
    L.q128 = L.rows + ((P * mask_wpr(P) * 4 + 15) & ~15);
    const int plane = (P * P + 15) & ~15;           // odd patch sizes: keep every array 16-byte aligned
    L.q254 = L.q128 + plane;
    L.list = L.q254 + plane;
    L.parts = L.list + ((P * P * 2 + 15) & ~15);
    L.total = L.parts + kCombos * kG64NW * kNP * 4;
    return L;
}

__device__ __forceinline__ int tri_cell(int a, int b) {
    const int lo = min(a, b), hi = max(a, b);
    return ((hi * (hi + 1)) >> 1) + lo;
}
__device__ __forceinline__ void tri_add(uint32_t* tri32w, int cell) {
    atomicAdd(&tri32w[cell >> 1], 1u << ((cell & 1) * 16));
}
__device__ __forceinline__ uint32_t hash_slot(uint32_t key, int lg) { return (key * 0x9E3779B1u) >> (32 - lg); }
// slot = (key+1) << 16 | count
__device__ __forceinline__ void hash_add(uint32_t* tab, int lg, uint32_t key) {
    const uint32_t mask = (1u << lg) - 1u, tag = (key + 1u) << 16;
    uint32_t h = hash_slot(key, lg);
    while (true) {
        uint32_t cur = tab[h];
        if (cur == 0u) cur = atomicCAS(&tab[h], 0u, tag | 1u);
        else if ((cur & 0xffff0000u) == tag) { atomicAdd(&tab[h], 1u); return; }
        else { h = (h + 1u) & mask; continue; }
        if (cur == 0u) return;                                            // we inserted it
        if ((cur & 0xffff0000u) == tag) { atomicAdd(&tab[h], 1u); return; }   // somebody else did
        h = (h + 1u) & mask;
    }
}
__device__ __forceinline__ uint32_t hash_get(const uint32_t* tab, int lg, uint32_t key) {
    const uint32_t mask = (1u << lg) - 1u, tag = (key + 1u) << 16;
    uint32_t h = hash_slot(key, lg);
    while (true) {
        const uint32_t cur = tab[h];
        if ((cur & 0xffff0000u) == tag) return cur & 0xffffu;
        if (cur == 0u) return 0u;
        h = (h + 1u) & mask;
    }
}
__device__ __forceinline__ uint32_t fix_ln(uint32_t g) {   // ln(g) in 16.16 fixed point, g >= 1
    return __float2uint_rn(__log2f((float)g) * kLnFix);
}
__device__ __forceinline__ uint32_t fix_clnc(uint32_t c) {   // c ln c, 15 fractional bits: sum <= T ln T * 2^15 < 2^32 for T <= 8192
    return c ? __float2uint_rn((float)c * __log2f((float)c) * (0.5f * kLnFix)) : 0u;
}

__global__ void __launch_bounds__(kG64Threads, 3)
k_glcm64(const GlcmParams p, const __grid_constant__ CUtensorMap map) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = p.P, wpr = mask_wpr(P), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t i = blockIdx.x;
    const Glcm64Smem L = glcm64_layout(P);
    uint8_t* patch = smem_raw;
    uint32_t* tri128 = reinterpret_cast<uint32_t*>(smem_raw + kOffTri128);
    uint32_t* tri64 = reinterpret_cast<uint32_t*>(smem_raw + kOffTri64);
    uint32_t* tri32 = reinterpret_cast<uint32_t*>(smem_raw + kOffTri32);
    uint32_t* hash = reinterpret_cast<uint32_t*>(smem_raw + kOffHash);
    uint32_t* marg = reinterpret_cast<uint32_t*>(smem_raw + L.marg);
    // p_x of 128 / 254 levels (64 and 32 are folded from 128), p_{x+y} of all four. The two small sum histograms are
    // replicated (3 and 8 copies in adjacent words, picked by the lane): most pairs of a warp fall into the same few bins,
    // and same-address atomics serialise (6.7 and 4.7 wavefronts per instruction before).
    uint32_t* hx128 = marg, *hx254 = marg + 128, *hs128 = marg + 384, *hs254 = marg + 640, *hs64r = marg + 1152, *hs32r = marg + 1536;
    const int rep3 = lane % 3, rep8 = lane & 7;
    uint32_t* rows = reinterpret_cast<uint32_t*>(smem_raw + L.rows);
    uint16_t* qq = reinterpret_cast<uint16_t*>(smem_raw + L.q128);   // per pixel: q254 | q128 << 8 (spans the q128 and q254 slots of the layout)
    uint16_t* list = reinterpret_cast<uint16_t*>(smem_raw + L.list);
    uint32_t* parts = reinterpret_cast<uint32_t*>(smem_raw + L.parts);   // [combo][warp][kNP]
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_lut[256];
    __shared__ int s_scan[kG64NW + 1];
    __shared__ int s_np[kGlcmOffsets];
    __shared__ float s_idm[256];   // 2 / (1 + k^2): a pair adds 2 to G at distance k from the diagonal

    const NucInfo inf = p.info[i];
    const float* greyp = p.grey ? p.grey + i * (int64_t)P * P : nullptr;   // f32 grey plane instead of the u8 window
    if (tid == 0 && !greyp) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, (uint32_t)(patch_panels(P) * kPanelBytes * P));
        tma_load_patch(patch, &map, inf.left, inf.top, P, &bar);
    }
    if (tid < kGlcmOffsets) s_np[tid] = 0;
    if (tid < 256) {
        s_lut[tid] = __fdiv_rn((float)tid, 255.0f);   // utils.rs:172  u8 -> f32 / 255.0
        s_idm[tid] = __fdiv_rn(2.0f, 1.0f + (float)(tid * tid));
    }
    // ---- mask rows -> shared memory + compacted pixel list ((row << 8) | col) ----
    const uint32_t* gm = p.bitmask + i * (int64_t)P * wpr;
    int K = 0;
    for (int base = 0; base < P * wpr; base += kG64Threads) {
        const int k = base + tid;
        uint32_t bits = (k < P * wpr) ? gm[k] : 0u;
        if (k < P * wpr) rows[k] = bits;
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o2);
            if (lane >= o2) incl += t;
        }
        __syncthreads();
        if (lane == 31) s_scan[warp] = incl;
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int t = 0; t < kG64NW; ++t) {
            const int v = s_scan[t];
            wbase += (t < warp) ? v : 0;
            total += v;
        }
        int pos = K + wbase + incl - cnt;
        const int r = k / wpr, cb = (k - r * wpr) * 32;
        while (bits) {
            const int c = cb + __ffs(bits) - 1;
            bits &= bits - 1;
            list[pos++] = (uint16_t)((r << 8) | c);
        }
        K += total;
    }
    constexpr int lg = kHashLg;
    const int o = patch_byte_offset(inf.left);
    __syncthreads();
    if (!greyp) mbar_wait(&bar, 0);
    // ---- grey quantisation (texture.rs:36 + SPEC.md B5), bit-exact, masked pixels only ----
    const bool dbg_all = (p.dbg_grey != nullptr);
    auto quantise = [&](int r, int c) {
        uint32_t pr = 0, pg = 0, pb = 0;
        if (r < inf.nvr && c < inf.nvc) {   // utils.rs:161-192: the rest of the window is zero
            const int a = patch_addr(P, o, r, c);
            pr = patch[a]; pg = patch[a + 1]; pb = patch[a + 2];
        }
        const float g = __fdiv_rn(__fadd_rn(__fadd_rn(s_lut[pr], s_lut[pg]), s_lut[pb]), 3.0f);
        const int a128 = min((int)floorf(__fmul_rn(g, 128.0f)), 127), a254 = min((int)floorf(__fmul_rn(g, p.scale254)), 253);
        qq[r * P + c] = (uint16_t)(a254 | (a128 << 8));
    };
    auto quantise_grey = [&](int r, int c) {   // f32 batch (f32batch.cu): the grey plane is given
        const float g = greyp[r * P + c];
        const int a128 = min((int)floorf(__fmul_rn(g, 128.0f)), 127), a254 = min((int)floorf(__fmul_rn(g, p.scale254)), 253);
        qq[r * P + c] = (uint16_t)(a254 | (a128 << 8));
    };
    if (dbg_all) {
        for (int k = tid; k < P * P; k += kG64Threads) { if (greyp) quantise_grey(k / P, k % P); else quantise(k / P, k % P); }
    }
    // The list entry becomes (position, neighbour flags): bit 12 + oi is set when the pixel's neighbour at offset oi is
    // inside the window and masked. The sweeps then cost three 16-bit loads per pixel (entry, own levels, neighbour's
    // levels) instead of a mask-word load with its bit arithmetic and four byte loads (ncu round 2: 11 % of the kernel's
    // shared-memory wavefronts and a fifth of the two sweeps' instructions went into finding the pair).
    auto list_loop = [&](auto grey_tag) {
        for (int j = tid; j < K; j += kG64Threads) {
            const uint32_t rc = list[j];
            const int r = rc >> 8, c = rc & 255;
            if (!dbg_all) { if constexpr (decltype(grey_tag)::value) quantise_grey(r, c); else quantise(r, c); }
            uint32_t fl = 0u;
#pragma unroll
            for (int oi = 0; oi < kGlcmOffsets; ++oi) {
                const int r2 = r + c_off[oi][0], c2 = c + c_off[oi][1];
                if ((r2 < P) && ((unsigned)c2 < (unsigned)P) && ((rows[r2 * wpr + (c2 >> 5)] >> (c2 & 31)) & 1u)) fl |= 1u << oi;
            }
            list[j] = (uint16_t)((r * P + c) | (fl << 12));
        }
    };
    if (greyp) list_loop(std::true_type{}); else list_loop(std::false_type{});
    __syncthreads();   // the window is dead from here on: region A becomes the histograms
    {
        uint4* z = reinterpret_cast<uint4*>(smem_raw);
        for (int k = tid; k < (kOffHash + (4 << lg)) / 16; k += kG64Threads) z[k] = make_uint4(0, 0, 0, 0);
        uint4* zm = reinterpret_cast<uint4*>(marg);
        for (int k = tid; k < kMargWords / 4; k += kG64Threads) zm[k] = make_uint4(0, 0, 0, 0);
    }
    if (dbg_all) {
        int lv = 3;
        for (int t = 0; t < 4; ++t)
            if (c_levels[t] == p.dbg_levels) lv = t;
        for (int k = tid; k < P * P; k += kG64Threads)
            p.dbg_grey[i * (int64_t)P * P + k] = (uint8_t)((lv == 3) ? (qq[k] & 0xff) : ((qq[k] >> 8) >> (2 - lv)));
    }
    __syncthreads();

    for (int oi = 0; oi < kGlcmOffsets; ++oi) {
        const int dy = c_off[oi][0], dx = c_off[oi][1];
        const int dpos = dy * P + dx;
        // ---- pass 1: neighbour test, atomics of all four levels ----
        // The difference histogram p_{x-y} only ever enters through its moments sum k c, sum k^2 c and sum c / (1 + k^2),
        // which are linear in the pairs: they are accumulated in registers (its atomics were the most conflicted of all:
        // |a - b| is 0, 1 or 2 for most pairs, so up to 15 lanes of a warp hit the same word).
        int np_local = 0;
        uint32_t d1[4] = {0, 0, 0, 0}, d2[4] = {0, 0, 0, 0};
        float fi[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = tid; j < K; j += kG64Threads) {
            const uint32_t e = list[j];
            if (!((e >> (12 + oi)) & 1u)) continue;
            ++np_local;
            const int src = e & 0xfff;
            const uint32_t qa = qq[src], qb = qq[src + dpos];
            const int a3 = qa & 0xff, b3 = qb & 0xff, a2 = qa >> 8, b2 = qb >> 8;
            // 254 levels
            hash_add(hash, lg, (uint32_t)(min(a3, b3) * 256 + max(a3, b3)));
            atomicAdd(&hx254[a3], 1u);
            atomicAdd(&hx254[b3], 1u);
            atomicAdd(&hs254[a3 + b3], 2u);
            { const int k = abs(a3 - b3); d1[3] += k; d2[3] += k * k; fi[3] += s_idm[k]; }
            // 128 levels
            tri_add(tri128, tri_cell(a2, b2));
            atomicAdd(&hx128[a2], 1u);
            atomicAdd(&hx128[b2], 1u);
            atomicAdd(&hs128[a2 + b2], 2u);
            { const int k = abs(a2 - b2); d1[2] += k; d2[2] += k * k; fi[2] += s_idm[k]; }
            // 64 levels (p_x is folded from the 128-level histogram later)
            const int a1 = a2 >> 1, b1 = b2 >> 1;
            tri_add(tri64, tri_cell(a1, b1));
            atomicAdd(&hs64r[(a1 + b1) * 3 + rep3], 2u);
            { const int k = abs(a1 - b1); d1[1] += k; d2[1] += k * k; fi[1] += s_idm[k]; }
            // 32 levels
            const int a0 = a2 >> 2, b0 = b2 >> 2;
            tri_add(tri32, tri_cell(a0, b0));
            atomicAdd(&hs32r[(a0 + b0) * 8 + rep8], 2u);
            { const int k = abs(a0 - b0); d1[0] += k; d2[0] += k * k; fi[0] += s_idm[k]; }
        }
        np_local = __reduce_add_sync(0xffffffffu, np_local);
        if (lane == 0 && np_local) atomicAdd(&s_np[oi], np_local);
#pragma unroll
        for (int lv = 0; lv < 4; ++lv) {   // parked now: keeps the twelve accumulators out of pass 2's register budget
            const uint32_t t1 = __reduce_add_sync(0xffffffffu, d1[lv]), t2 = __reduce_add_sync(0xffffffffu, d2[lv]);
            const float tf = warp_sum(fi[lv]);
            if (lane == 0) {
                uint32_t* pp = parts + ((lv * kGlcmOffsets + oi) * kG64NW + warp) * kNP;
                pp[3] = 2u * t1;   // every pair counts twice in the symmetric matrix
                pp[4] = 2u * t2;
                pp[10] = __float_as_uint(tf);
            }
        }
        __syncthreads();
        // ---- pass 2: per-pair cell counts (entropy, ASM) for the four levels ----
        uint32_t sg[4] = {0, 0, 0, 0}, sl[4] = {0, 0, 0, 0};
        for (int j = tid; j < K; j += kG64Threads) {
            const uint32_t e = list[j];
            if (!((e >> (12 + oi)) & 1u)) continue;
            const int src = e & 0xfff;
            const uint32_t qa = qq[src], qb = qq[src + dpos];
            const int a3 = qa & 0xff, b3 = qb & 0xff, a2 = qa >> 8, b2 = qb >> 8;
            const int a1 = a2 >> 1, b1 = b2 >> 1, a0 = a2 >> 2, b0 = b2 >> 2;
            uint32_t g[4];
            g[3] = hash_get(hash, lg, (uint32_t)(min(a3, b3) * 256 + max(a3, b3))) << (a3 == b3 ? 1 : 0);
            g[2] = (uint32_t)reinterpret_cast<const uint16_t*>(tri128)[tri_cell(a2, b2)] << (a2 == b2 ? 1 : 0);
            g[1] = (uint32_t)reinterpret_cast<const uint16_t*>(tri64)[tri_cell(a1, b1)] << (a1 == b1 ? 1 : 0);
            g[0] = (uint32_t)reinterpret_cast<const uint16_t*>(tri32)[tri_cell(a0, b0)] << (a0 == b0 ? 1 : 0);
#pragma unroll
            for (int lv = 0; lv < 4; ++lv) {
                sg[lv] += g[lv];
                sl[lv] += fix_ln(g[lv]);
            }
            if (p.dbg_counts && p.dbg_dy == dy && p.dbg_dx == dx) {
                const int NL = p.dbg_levels;
                const int lv = NL == 32 ? 0 : (NL == 64 ? 1 : (NL == 128 ? 2 : 3));
                const int a = lv == 3 ? a3 : (a2 >> (2 - lv)), b = lv == 3 ? b3 : (b2 >> (2 - lv));
                uint32_t* dc = p.dbg_counts + i * (int64_t)NL * NL;
                dc[a * NL + b] = g[lv];
                dc[b * NL + a] = g[lv];
            }
        }
        // ---- marginal moments, one level at a time; warp REDUX, per-warp partials to shared memory ----
#pragma unroll
        for (int lv = 0; lv < 4; ++lv) {
            const int NL = 32 << lv;   // table geometry (254 levels use the 256-wide layout)
            uint32_t v[kNP];
#pragma unroll
            for (int q = 0; q < kNP; ++q) v[q] = 0u;
            v[0] = sg[lv];
            v[7] = sl[lv];
            for (int k = tid; k < 2 * NL; k += kG64Threads) {
                const uint32_t kk = (uint32_t)k;
                uint32_t c;                                                 // p_{x+y}
                if (lv == 3) c = hs254[k];
                else if (lv == 2) c = hs128[k];
                else if (lv == 1) c = hs64r[3 * k] + hs64r[3 * k + 1] + hs64r[3 * k + 2];
                else {
                    const uint4 u0 = reinterpret_cast<const uint4*>(hs32r)[2 * k], u1 = reinterpret_cast<const uint4*>(hs32r)[2 * k + 1];
                    c = (u0.x + u0.y) + (u0.z + u0.w) + (u1.x + u1.y) + (u1.z + u1.w);
                }
                v[5] += kk * c;
                v[6] += kk * kk * c;
                v[9] += fix_clnc(c);
                if (k < NL) {
                    uint32_t cx;                                            // p_x
                    if (lv == 3) cx = hx254[k];
                    else if (lv == 2) cx = hx128[k];
                    else if (lv == 1) cx = hx128[2 * k] + hx128[2 * k + 1];
                    else cx = hx128[4 * k] + hx128[4 * k + 1] + hx128[4 * k + 2] + hx128[4 * k + 3];
                    v[1] += kk * cx;
                    v[2] += kk * kk * cx;
                    v[8] += fix_clnc(cx);
                }
            }
            const int combo = lv * kGlcmOffsets + oi;
#pragma unroll
            for (int q = 0; q < kNP; ++q) {
                if (q == 3 || q == 4 || q == 10) continue;   // difference moments: written after pass 1
                const uint32_t t = __reduce_add_sync(0xffffffffu, v[q]);
                if (lane == 0) parts[(combo * kG64NW + warp) * kNP + q] = t;
            }
        }
        __syncthreads();
        // ---- dense clear of every table ----
        if (oi + 1 < kGlcmOffsets) {
            uint4* z = reinterpret_cast<uint4*>(smem_raw);
            for (int k = tid; k < (kOffHash + (4 << lg)) / 16; k += kG64Threads) z[k] = make_uint4(0, 0, 0, 0);
            uint4* zm = reinterpret_cast<uint4*>(marg);
            for (int k = tid; k < kMargWords / 4; k += kG64Threads) zm[k] = make_uint4(0, 0, 0, 0);
            __syncthreads();
        }
    }
    // ---- 14 Haralick features per (level, offset): one thread per combination ----
    if (tid < kCombos && p.out) {
        const int combo = tid, oi = combo % kGlcmOffsets;
        float* o_ = p.out + i * (int64_t)p.out_stride + p.col_glcm + combo * kGlcmFeat;
        const int npairs = s_np[oi];
        if (npairs == 0) {
            for (int f = 0; f < kGlcmFeat; ++f) o_[f] = CUDART_NAN_F;   // 0/0 (SPEC.md B5)
        } else {
            double acc[kNP];
            for (int q = 0; q < kNP; ++q) {
                unsigned long long t = 0;
                double tf = 0.0;
                for (int w = 0; w < kG64NW; ++w) {
                    const uint32_t v = parts[(combo * kG64NW + w) * kNP + q];
                    t += v;
                    tf += (double)__uint_as_float(v);   // q == 10 only: f32 partial sums of 2 / (1 + k^2)
                }
                acc[q] = q == 10 ? tf : (double)t;
            }
            haralick_write(o_, 2.0 * (double)npairs, acc[0], acc[7] / 65536.0, acc[1], acc[2], acc[3], acc[4], acc[5], acc[6],
                           acc[8] / 32768.0, acc[9] / 32768.0, acc[10]);
        }
    }
}


// =================================================================================================
// k_glcm_large: 128 < P <= 256 (BASELINE config 5, 256x256 windows). The quantised plane of ONE level
// family lives in shared memory at a time (q254, then q128: 64 KB), the window is streamed through
// 64-row slabs (aliased with the histogram region), co-occurring pairs are enumerated straight from
// the mask words, G is a dense triangular u16 histogram (counts <= 65 280 fit), moments are u64.
// =================================================================================================
constexpr int kLargeThreads = 1024;
constexpr int kLNW = kLargeThreads / 32;
struct GlcmLargeSmem {
    int plane, region_t, rows, hist, part_i, part_f, total;
};
__host__ __device__ inline GlcmLargeSmem glcm_large_layout(int P) {
    GlcmLargeSmem L;
    L.plane = 0;
    L.region_t = (P * (P + 4) + 127) & ~127;   // plane rows are P + 4 bytes apart: a 256-byte pitch would put every row on the same banks
    int t = window_smem_bytes(P, 64) > kTriBytes ? window_smem_bytes(P, 64) : kTriBytes;
    L.rows = L.region_t + ((t + 127) & ~127);
    L.hist = L.rows + ((P * mask_wpr(P) * 4 + 15) & ~15);
    L.part_i = L.hist + 3 * 1024 * 4;   // hx[256] hs[512] hd[256] for each of the (up to) three levels of a round
    L.part_f = L.part_i + kCombos * kLNW * kNI * 8;
    L.total = L.part_f + kCombos * kLNW * kNF * 4;
    return L;
}

__global__ void __launch_bounds__(kLargeThreads, 1)
k_glcm_large(const GlcmParams p, const __grid_constant__ CUtensorMap map /* box {208, 64 rows} */) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = p.P, wpr = mask_wpr(P), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t i = blockIdx.x;
    const GlcmLargeSmem L = glcm_large_layout(P);
    uint8_t* plane = smem_raw + L.plane;
    const int PP = P + 4;   // plane pitch
    uint8_t* slab = smem_raw + L.region_t;
    uint32_t* tri32 = reinterpret_cast<uint32_t*>(smem_raw + L.region_t);
    uint16_t* tri16 = reinterpret_cast<uint16_t*>(smem_raw + L.region_t);
    uint32_t* rows = reinterpret_cast<uint32_t*>(smem_raw + L.rows);
    uint32_t* hx = reinterpret_cast<uint32_t*>(smem_raw + L.hist);
    unsigned long long* part_i = reinterpret_cast<unsigned long long*>(smem_raw + L.part_i);
    float* part_f = reinterpret_cast<float*>(smem_raw + L.part_f);
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_lut[256];
    __shared__ unsigned long long s_npairs[kGlcmOffsets];
    __shared__ float s_idm[256];   // 2 / (1 + k^2): a pair adds 2 to G at distance k from the diagonal

    const NucInfo inf = p.info[i];
    const int o = patch_byte_offset(inf.left);
    const float* greyp = p.grey ? p.grey + i * (int64_t)P * P : nullptr;   // f32 grey plane instead of the u8 window
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (tid < 256) {
        s_lut[tid] = __fdiv_rn((float)tid, 255.0f);
        s_idm[tid] = __fdiv_rn(2.0f, 1.0f + (float)(tid * tid));
    }
    const uint32_t* gm = p.bitmask + i * (int64_t)P * wpr;
    for (int k = tid; k < P * wpr; k += kLargeThreads) rows[k] = gm[k];
    __syncthreads();
    const int nslab = (P + 63) / 64;
    const uint32_t slab_tx = (uint32_t)(patch_panels(P) * kPanelBytes * 64);
    uint32_t phase = 0;

    for (int round = 0; round < 2; ++round) {   // round 0: 254 levels (q254 plane); round 1: 128/64/32 (q128 plane)
        // ---- stream the window and quantise it (texture.rs:36 + SPEC.md B5, bit-exact) ----
        for (int sidx = 0; sidx < nslab; ++sidx) {
            const int row0 = sidx * 64, nrows = min(64, P - row0);
            __syncthreads();   // slab buffer (aliased with the histograms) is free
            if (!greyp) {
                if (tid == 0) {
                    mbar_expect_tx(&bar, slab_tx);
                    tma_load_window(slab, &map, inf.left, inf.top + row0, P, 64, &bar);
                }
                mbar_wait(&bar, phase);
                phase ^= 1u;
            }
            for (int k = tid; k < nrows * P; k += kLargeThreads) {
                const int lr = k / P, c = k - lr * P, r = row0 + lr;
                // only pixels under the mask ever enter a pair (the debug dump wants the whole plane)
                if (!p.dbg_grey && !((rows[r * wpr + (c >> 5)] >> (c & 31)) & 1u)) continue;
                uint32_t pr = 0, pg = 0, pb = 0;
                if (!greyp && r < inf.nvr && c < inf.nvc) {
                    const int a = patch_addr(64, o, lr, c);
                    pr = slab[a]; pg = slab[a + 1]; pb = slab[a + 2];
                }
                const float g = greyp ? greyp[r * P + c] : __fdiv_rn(__fadd_rn(__fadd_rn(s_lut[pr], s_lut[pg]), s_lut[pb]), 3.0f);
                plane[r * PP + c] = round == 0 ? (uint8_t)min((int)floorf(__fmul_rn(g, p.scale254)), 253)
                                              : (uint8_t)min((int)floorf(__fmul_rn(g, 128.0f)), 127);
            }
        }
        __syncthreads();
        if (p.dbg_grey) {
            const int lvd = p.dbg_levels == 32 ? 0 : (p.dbg_levels == 64 ? 1 : (p.dbg_levels == 128 ? 2 : 3));
            if ((round == 0) == (lvd == 3))
                for (int k = tid; k < P * P; k += kLargeThreads)
                    p.dbg_grey[i * (int64_t)P * P + k] = (uint8_t)(lvd == 3 ? plane[(k / P) * PP + k % P] : (plane[(k / P) * PP + k % P] >> (2 - lvd)));
        }
        for (int k = tid; k < kTriBytes / 16; k += kLargeThreads)
            reinterpret_cast<uint4*>(smem_raw + L.region_t)[k] = make_uint4(0, 0, 0, 0);
        for (int k = tid; k < 3 * 1024; k += kLargeThreads) hx[k] = 0u;
        __syncthreads();

        // Levels of this round: 254 alone (round 0), or 128 / 64 / 32 together (round 1: q64 = q128 >> 1, q32 = q128 >> 2,
        // so ONE enumeration of the co-occurring pairs feeds the three matrices: 12 atomics per pair in the first
        // sweep, three cell look-ups in the second). Slot s of the round holds level lv_hi - s.
        const int lv_hi = round == 0 ? 3 : 2, lv_lo = round == 0 ? 3 : 0, nlev = lv_hi - lv_lo + 1;
        // triangular u16 histograms of the slots, packed two cells per u32 word, back to back in region T
        int tri_off[3], tri_words_tot = 0;
        for (int s = 0; s < 3; ++s) {
            tri_off[s] = tri_words_tot;
            if (s < nlev) { const int NL = c_levels[lv_hi - s]; tri_words_tot += ((NL * (NL + 1) / 2) + 1) / 2; }
        }
        for (int oi = 0; oi < kGlcmOffsets; ++oi) {
            const int dy = c_off[oi][0], dx = c_off[oi][1], dpos = dy * PP + dx;
            unsigned long long g2[3] = {0, 0, 0}, np_local = 0;   // sum of the cell counts met (= sum of squares of the cells)
            float glg[3] = {0.f, 0.f, 0.f};                       // sum of ln(cell count)
            // moments of the difference histogram p_{x-y}: linear in the pairs, kept in registers (its atomics were the
            // most conflicted ones: |a - b| is 0, 1 or 2 for most pairs)
            uint32_t d1[3] = {0, 0, 0}, d2[3] = {0, 0, 0};        // <= 128 pairs per thread: 128 * 253^2 < 2^32
            float fi[3] = {0.f, 0.f, 0.f};
            for (int pass = 0; pass < 2; ++pass) {
                // Every lane prepares the pair bits of ONE mask word; the warp then walks its 32 words together, lane l taking
                // bit l of the word in turn (empty words are skipped warp-uniformly). Lanes that walked the bits of their own
                // word ran 15 of 32 at a time (words inside the nucleus are full, words outside are empty).
                for (int base = 0; base < P * wpr; base += kLargeThreads) {
                    const int k = base + tid;
                    uint32_t pbw = 0u;
                    if (k < P * wpr) {
                        const int r = k / wpr, w = k - r * wpr, r2 = r + dy;
                        if (r2 < P) {
                            const uint32_t* nr = rows + r2 * wpr;
                            uint32_t nb = nr[w];
                            if (dx == 1) nb = (nb >> 1) | ((w + 1 < wpr) ? (nr[w + 1] << 31) : 0u);
                            else if (dx == -1) nb = (nb << 1) | ((w > 0) ? (nr[w - 1] >> 31) : 0u);
                            pbw = rows[k] & nb;
                        }
                    }
                    if (!__any_sync(0xffffffffu, pbw != 0u)) continue;
                    for (int t = 0; t < 32; ++t) {
                        const uint32_t pbt = __shfl_sync(0xffffffffu, pbw, t);
                        if (pbt == 0u) continue;                      // warp-uniform
                        if (!((pbt >> lane) & 1u)) continue;
                        const int kt = base + (warp << 5) + t, r = kt / wpr, w = kt - r * wpr;
                        const int c = 32 * w + lane;
                        const int src = r * PP + c;
                        const int a0 = plane[src], b0 = plane[src + dpos];
                        if (pass == 0) ++np_local;
#pragma unroll
                        for (int s = 0; s < 3; ++s) {
                            if (s >= nlev) break;
                            const int a = a0 >> s, b = b0 >> s;
                            const int lo = min(a, b), hi = max(a, b);
                            const int cell = ((hi * (hi + 1)) >> 1) + lo;
                            uint32_t* hxs = hx + s * 1024;
                            if (pass == 0) {
                                if (s == 0) {   // slots 1, 2 (64 / 32 levels) are POOLED from the 128-level tables below
                                    atomicAdd(&tri32[tri_off[s] + (cell >> 1)], 1u << ((cell & 1) * 16));
                                    atomicAdd(&hxs[a], 1u);
                                    atomicAdd(&hxs[b], 1u);
                                    atomicAdd(&hxs[256 + a + b], 2u);
                                }
                                const uint32_t kd = (uint32_t)(hi - lo);
                                d1[s] += kd;
                                d2[s] += kd * kd;
                                fi[s] += s_idm[kd];
                            } else if (s == 0) {
                                const uint32_t g = (uint32_t)tri16[2 * tri_off[s] + cell] << (a == b ? 1 : 0);
                                g2[s] += g;
                                glg[s] += __logf((float)g);
                                const int NL = c_levels[lv_hi - s];
                                if (p.dbg_counts && p.dbg_levels == NL && p.dbg_dy == dy && p.dbg_dx == dx) {
                                    uint32_t* dc = p.dbg_counts + i * (int64_t)NL * NL;
                                    dc[a * NL + b] = g;
                                    dc[b * NL + a] = g;
                                }
                            }
                        }
                    }
                }
                __syncthreads();
                if (pass == 0 && nlev == 3) {
                    // q64 = q128 >> 1 and q32 = q128 >> 2 exactly, so the 64- and 32-level matrices are 2 x 2 poolings of the
                    // 128-level one: at 256 x 256 a nucleus has ~25 000 pairs per offset against 8 256 + 2 080 + 528 cells, and
                    // pooling replaces 8 of the 12 atomics per pair and two of the three cell look-ups of the second sweep.
                    const uint16_t* T128 = tri16 + 2 * tri_off[0];
                    uint16_t* T64 = tri16 + 2 * tri_off[1];
                    uint16_t* T32 = tri16 + 2 * tri_off[2];
                    uint32_t* hx64 = hx + 1024, *hx32 = hx + 2048;
                    auto tc = [](int hi, int lo) { return ((hi * (hi + 1)) >> 1) + lo; };
                    for (int t = tid; t < 64 * 64; t += kLargeThreads) {
                        const int A = t >> 6, B = t & 63;
                        if (B > A) continue;
                        uint32_t c = (uint32_t)T128[tc(2 * A, 2 * B)] + T128[tc(2 * A + 1, 2 * B + 1)] + T128[tc(2 * A + 1, 2 * B)];
                        if (A > B) c += T128[tc(2 * A, 2 * B + 1)];
                        T64[tc(A, B)] = (uint16_t)c;   // <= 65 280 pairs per offset in total
                    }
                    if (tid < 64) hx64[tid] = hx[2 * tid] + hx[2 * tid + 1];
                    if (tid >= 64 && tid < 96) { const int k = tid - 64; hx32[k] = hx[4 * k] + hx[4 * k + 1] + hx[4 * k + 2] + hx[4 * k + 3]; }
                    __syncthreads();
                    {
                        const int A = tid >> 5, B = tid & 31;   // 1024 threads = 32 x 32
                        if (B <= A) {
                            uint32_t c = (uint32_t)T64[tc(2 * A, 2 * B)] + T64[tc(2 * A + 1, 2 * B + 1)] + T64[tc(2 * A + 1, 2 * B)];
                            if (A > B) c += T64[tc(2 * A, 2 * B + 1)];
                            T32[tc(A, B)] = (uint16_t)c;
                        }
                    }
                    __syncthreads();
                    // p_{x+y}: a pair adds 2 at a + b; sums over the anti-diagonals of the pooled matrices
                    if (tid < 127) {
                        uint32_t c = 0;
                        for (int B = max(0, tid - 63); 2 * B <= tid; ++B) c += T64[tc(tid - B, B)];
                        hx64[256 + tid] = 2u * c;
                    } else if (tid >= 128 && tid < 128 + 63) {
                        const int k = tid - 128;
                        uint32_t c = 0;
                        for (int B = max(0, k - 31); 2 * B <= k; ++B) c += T32[tc(k - B, B)];
                        hx32[256 + k] = 2u * c;
                    }
                    // entropy / ASM sums of the pooled matrices, cell by cell: sum_pairs G = sum_cells c G, sum_pairs ln G = sum_cells c ln G
                    for (int t = tid; t < 64 * 64 + 32 * 32; t += kLargeThreads) {
                        const bool l64 = t < 64 * 64;
                        const int u = l64 ? t : t - 64 * 64, NLq = l64 ? 64 : 32, A = l64 ? (u >> 6) : (u >> 5), B = u & (NLq - 1);
                        if (B > A) continue;
                        const uint32_t c = l64 ? T64[tc(A, B)] : T32[tc(A, B)];
                        const uint32_t g = c << (A == B ? 1 : 0);
                        if (c) {
                            const unsigned long long cg = (unsigned long long)c * g;
                            const float cl = (float)c * __logf((float)g);
                            if (l64) { g2[1] += cg; glg[1] += cl; } else { g2[2] += cg; glg[2] += cl; }
                        }
                        if (p.dbg_counts && p.dbg_levels == NLq && p.dbg_dy == dy && p.dbg_dx == dx) {
                            uint32_t* dc = p.dbg_counts + i * (int64_t)NLq * NLq;
                            dc[A * NLq + B] = g;
                            dc[B * NLq + A] = g;
                        }
                    }
                    __syncthreads();
                }
            }
            np_local = warp_sum(np_local);
            if (tid == 0) s_npairs[oi] = 0ull;
            __syncthreads();
            if (lane == 0 && np_local) atomicAdd(&s_npairs[oi], np_local);
            // ---- marginal sums of every slot, partials per warp ----
            for (int s = 0; s < nlev; ++s) {
                const int lv = lv_hi - s, NL = c_levels[lv], combo = lv * kGlcmOffsets + oi;
                const uint32_t* hxs = hx + s * 1024;
                const uint32_t* hss = hxs + 256;
                unsigned long long vl[kNI] = {s == 0 ? g2[0] : (s == 1 ? g2[1] : g2[2]), 0, 0, 0, 0, 0, 0};
                float vf[kNF] = {s == 0 ? glg[0] : (s == 1 ? glg[1] : glg[2]), 0.f, 0.f, 0.f};
                vl[3] = 2ull * (s == 0 ? d1[0] : (s == 1 ? d1[1] : d1[2]));   // every pair counts twice in the symmetric matrix
                vl[4] = 2ull * (s == 0 ? d2[0] : (s == 1 ? d2[1] : d2[2]));
                vf[3] = s == 0 ? fi[0] : (s == 1 ? fi[1] : fi[2]);
                for (int k = tid; k < 2 * NL - 1; k += kLargeThreads) {
                    const unsigned long long c = hss[k], kk = (unsigned long long)k;
                    vl[5] += kk * c;
                    vl[6] += kk * kk * c;
                    if (c) vf[2] += (float)c * __logf((float)c);
                    if (k < NL) {
                        const unsigned long long cx = hxs[k];
                        vl[1] += kk * cx;
                        vl[2] += kk * kk * cx;
                        if (cx) vf[1] += (float)cx * __logf((float)cx);
                    }
                }
#pragma unroll
                for (int q = 0; q < kNI; ++q) {
                    const unsigned long long t = warp_sum(vl[q]);
                    if (lane == 0) part_i[(combo * kLNW + warp) * kNI + q] = t;
                }
#pragma unroll
                for (int q = 0; q < kNF; ++q) {
                    const float t = warp_sum(vf[q]);
                    if (lane == 0) part_f[(combo * kLNW + warp) * kNF + q] = t;
                }
            }
            __syncthreads();
            for (int k = tid; k < tri_words_tot; k += kLargeThreads) tri32[k] = 0u;
            for (int k = tid; k < nlev * 1024; k += kLargeThreads) hx[k] = 0u;
            __syncthreads();
        }
        // ---- features of this round's levels ----
        if (tid < kCombos && p.out) {
            const int combo = tid, lv = combo / kGlcmOffsets, oi = combo - lv * kGlcmOffsets;
            if (lv >= lv_lo && lv <= lv_hi) {
                float* o_ = p.out + i * (int64_t)p.out_stride + p.col_glcm + combo * kGlcmFeat;
                const unsigned long long npairs = s_npairs[oi];
                if (npairs == 0) {
                    for (int f = 0; f < kGlcmFeat; ++f) o_[f] = CUDART_NAN_F;
                } else {
                    double acc[kNI + kNF];
                    for (int q = 0; q < kNI; ++q) {
                        unsigned long long t = 0;
                        for (int w = 0; w < kLNW; ++w) t += part_i[(combo * kLNW + w) * kNI + q];
                        acc[q] = (double)t;
                    }
                    for (int q = 0; q < kNF; ++q) {
                        double t = 0.0;
                        for (int w = 0; w < kLNW; ++w) t += (double)part_f[(combo * kLNW + w) * kNF + q];
                        acc[kNI + q] = t;
                    }
                    haralick_write(o_, 2.0 * (double)npairs, acc[0], acc[kNI + 0], acc[1], acc[2], acc[3], acc[4], acc[5],
                                   acc[6], acc[kNI + 1], acc[kNI + 2], acc[kNI + 3]);
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace

int glcm_uses_slab_map(int P) { return P > 128; }

cudaError_t launch_glcm(const GlcmParams& p, const CUtensorMap* map, cudaStream_t s) {
    if (p.n <= 0) return cudaSuccess;
    if (p.P > 128) {   // `map` must be the 64-row slab map
        const int smem = glcm_large_layout(p.P).total;
        cudaError_t e = cudaFuncSetAttribute(k_glcm_large, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k_glcm_large<<<(unsigned)p.n, kLargeThreads, smem, s>>>(p, *map);
        return cudaGetLastError();
    }
    auto go = [&](auto kern, int threads, int smem) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)p.n, threads, smem, s>>>(p, *map);
        return cudaGetLastError();
    };
    if (p.P <= 64) return go(k_glcm64, kG64Threads, glcm64_layout(p.P).total);
    return go(k_glcm_generic<false>, kGlcmThreads, glcm_layout(p.P).total);
}

}  // namespace nfx
