// texture2.cu -- the two remaining texture sets of `all` / `texture` (src/args.rs:38-45), SURVEY.md 8f-1.
//
//   k_glrlm  GLRLMFeatureSet::compute_features_batched        src/features/texture.rs:178-310
//            tch_utils::glrlm::{glrlm, features::glrlm_features}   oracle/SPEC.md B9 (UNPINNED)
//   k_gabor  GaborFilterFeatureSet::compute_features_batched   src/features/texture.rs:322-369
//            tch_utils::gabor::apply_gabor_filter               oracle/SPEC.md B10 (UNPINNED)
//
// k_glrlm: one CTA per nucleus. The window is streamed through 64-row TMA slabs into a 24-level plane
// (bit-exact grey quantisation, same rule as the GLCM), then every masked pixel that STARTS a run
// walks it and bumps a 24x16 shared-memory histogram; 17 features per direction from exact counts.
//
// k_gabor: one CTA per nucleus (P <= 64). The isotropic-envelope Gabor kernel factorises exactly:
//   g(u,v) = G(u)G(v)cos(a u + b v) = [G(u)cos(a u)][G(v)cos(b v)] - [G(u)sin(a u)][G(v)sin(b v)]
// so each of the 48 filters is two separable passes (30 taps each instead of 900): a row pass over the
// mask's bounding box (+29 halo rows) into two shared-memory planes, and a column pass at the masked
// pixels only. Taps live in constant memory (warp-uniform index -> FFMA with a constant operand).
#include <math.h>
#include <math_constants.h>

#include "nfx_kernels.h"

namespace nfx {

namespace {

constexpr int kRlLevels = 24, kRlMax = 16;   // texture.rs:174-175
__device__ __constant__ int c_dirs[4][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}};   // (dx, dy) texture.rs:176

constexpr int kTexThreads = 256;
constexpr int kGaborFilters = 48, kGaborK = 30, kGaborLo = 14;   // 'same' padding: 14 before, 15 after
__device__ __constant__ float c_gabor[kGaborFilters][4][kGaborK];   // [filter][cu, su, cv, sv][tap]

__device__ __forceinline__ float grey_of(const float* lut, uint32_t r, uint32_t g, uint32_t b) {
    // texture.rs:189/332: mean_dim(-3) of u8/255 values = ((r+g)+b)/3 with IEEE f32 operations
    return __fdiv_rn(__fadd_rn(__fadd_rn(lut[r], lut[g]), lut[b]), 3.0f);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTexThreads)
k_glrlm(const TexParams p, const __grid_constant__ CUtensorMap map /* box {208, CS rows} */) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = p.P, wpr = mask_wpr(P), tid = threadIdx.x;
    const int CS = p.slab_rows, nslab = (P + CS - 1) / CS;
    const int64_t i = blockIdx.x;
    uint8_t* slab = smem_raw;
    uint8_t* plane = smem_raw + window_smem_bytes(P, CS);
    uint32_t* rows = reinterpret_cast<uint32_t*>(plane + ((P * P + 15) & ~15));
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_lut[256];
    __shared__ uint32_t s_R[kRlLevels * kRlMax];
    __shared__ double s_red[16 * (kTexThreads / 32)];

    const NucInfo inf = p.info[i];
    const int o = patch_byte_offset(inf.left);
    const uint32_t slab_tx = (uint32_t)(patch_panels(P) * kPanelBytes * CS);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, slab_tx);
        tma_load_window(slab, &map, inf.left, inf.top, P, CS, &bar);
    }
    s_lut[tid] = __fdiv_rn((float)tid, 255.0f);
    const uint32_t* gm = p.bitmask + i * (int64_t)P * wpr;
    double K = 0.0;
    for (int k = tid; k < P * wpr; k += kTexThreads) {
        const uint32_t b = gm[k];
        rows[k] = b;
        K += (double)__popc(b);
    }
    __syncthreads();
    for (int sidx = 0; sidx < nslab; ++sidx) {
        const int row0 = sidx * CS, nrows = min(CS, P - row0);
        mbar_wait(&bar, sidx & 1);
        for (int k = tid; k < nrows * P; k += kTexThreads) {
            const int lr = k / P, c = k - lr * P, r = row0 + lr;
            uint32_t pr = 0, pg = 0, pb = 0;
            if (r < inf.nvr && c < inf.nvc) {   // utils.rs:161-192: the rest of the window is zero
                const int a = patch_addr(CS, o, lr, c);
                pr = slab[a]; pg = slab[a + 1]; pb = slab[a + 2];
            }
            plane[r * P + c] = (uint8_t)min((int)floorf(__fmul_rn(grey_of(s_lut, pr, pg, pb), (float)kRlLevels)), kRlLevels - 1);
        }
        __syncthreads();
        if (tid == 0 && sidx + 1 < nslab) {
            mbar_expect_tx(&bar, slab_tx);
            tma_load_window(slab, &map, inf.left, inf.top + row0 + CS, P, CS, &bar);
        }
    }
    {
        double v[1] = {K};
        block_sum<1>(v, s_red);
        K = v[0];
    }
    float* out = p.out + i * (int64_t)p.out_stride + p.col_glrlm;
    const double cmid = (kRlLevels + 1) * 0.5;
    for (int d = 0; d < 4; ++d) {
        const int dx = c_dirs[d][0], dy = c_dirs[d][1];
        for (int k = tid; k < kRlLevels * kRlMax; k += kTexThreads) s_R[k] = 0u;
        __syncthreads();
        auto masked = [&](int r, int c) -> bool {
            return (unsigned)r < (unsigned)P && (unsigned)c < (unsigned)P && ((rows[r * wpr + (c >> 5)] >> (c & 31)) & 1u);
        };
        for (int k = tid; k < P * wpr; k += kTexThreads) {
            uint32_t bits = rows[k];
            const int r = k / wpr, cb = (k - r * wpr) * 32;
            while (bits) {
                const int c = cb + __ffs(bits) - 1;
                bits &= bits - 1;
                const int lv = plane[r * P + c];
                if (masked(r - dy, c - dx) && plane[(r - dy) * P + (c - dx)] == lv) continue;   // not a run start
                int len = 1, r2 = r + dy, c2 = c + dx;
                while (masked(r2, c2) && plane[r2 * P + c2] == lv) { ++len; r2 += dy; c2 += dx; }
                atomicAdd(&s_R[lv * kRlMax + min(len, kRlMax) - 1], 1u);
            }
        }
        __syncthreads();
        // ---- 17 features (oracle glrlm_features), float64 sums of exact counts ----
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = 0.0;
        for (int k = tid; k < kRlLevels * kRlMax; k += kTexThreads) {
            const double R = (double)s_R[k];
            const double ii = (double)(k / kRlMax + 1), jj = (double)(k % kRlMax + 1);
            const double u = (ii - cmid) / (cmid - 1.0), u2 = u * u;
            v[0] += R;                       // Nr
            v[1] += R / (jj * jj);           // SRE
            v[2] += R * jj * jj;             // LRE
            v[3] += R / (ii * ii);           // LGRE
            v[4] += R * ii * ii;             // HGRE
            v[5] += R / (ii * ii * jj * jj); // SRLGE
            v[6] += R * ii * ii / (jj * jj); // SRHGE
            v[7] += R * jj * jj / (ii * ii); // LRLGE
            v[8] += R * ii * ii * jj * jj;   // LRHGE
            v[9] += R * (1.0 - u2) / (jj * jj);    // short run, mid grey
            v[10] += R * (1.0 - u2) * jj * jj;     // long run, mid grey
            v[11] += R * u2 / (jj * jj);           // short run, extreme grey
            v[12] += R * u2 * jj * jj;             // long run, extreme grey
            v[13] += R * jj;                       // sum of run lengths
        }
        if (tid < kRlLevels) {   // grey-level non-uniformity: squared row sums
            double rs = 0.0;
            for (int j = 0; j < kRlMax; ++j) rs += (double)s_R[tid * kRlMax + j];
            v[14] = rs * rs;
        }
        if (tid < kRlMax) {      // run-length non-uniformity: squared column sums
            double cs = 0.0;
            for (int l = 0; l < kRlLevels; ++l) cs += (double)s_R[l * kRlMax + tid];
            v[15] = cs * cs;
        }
        block_sum<16>(v, s_red);
        if (tid == 0) {
            float* o_ = out + d * 17;
            const double nr = v[0], mean = v[13] / nr;
            o_[0] = (float)(v[1] / nr);  o_[1] = (float)(v[2] / nr);
            o_[2] = (float)(v[14] / nr); o_[3] = (float)(v[15] / nr);
            o_[4] = (float)(v[3] / nr);  o_[5] = (float)(v[4] / nr);
            o_[6] = (float)(v[5] / nr);  o_[7] = (float)(v[6] / nr);
            o_[8] = (float)(v[7] / nr);  o_[9] = (float)(v[8] / nr);
            o_[10] = (float)(v[9] / nr); o_[11] = (float)(v[10] / nr);
            o_[12] = (float)(v[11] / nr); o_[13] = (float)(v[12] / nr);
            o_[14] = (float)(nr / K);                        // run percentage
            o_[15] = (float)mean;                            // run length mean
            o_[16] = (float)(v[2] / nr - mean * mean);       // run length variance
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Dynamic smem: slab(window) | G[(P+29)][GS] f32 | A[(P+29)][P] f32 | B[(P+29)][P] f32 | rows | list u16
__global__ void __launch_bounds__(kTexThreads)
k_gabor(const TexParams p, const __grid_constant__ CUtensorMap map /* box {208, P rows} */) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = p.P, wpr = mask_wpr(P), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kTexThreads / 32;
    const int PH = P + kGaborK - 1, GS = P + kGaborK;   // padded height / row stride of G
    const int64_t i = blockIdx.x;
    uint8_t* patch = smem_raw;
    float* G = reinterpret_cast<float*>(smem_raw + patch_smem_bytes(P));
    float* A = G + PH * GS;
    float* B = A + PH * P;
    uint32_t* rows = reinterpret_cast<uint32_t*>(B + PH * P);
    uint16_t* list = reinterpret_cast<uint16_t*>(rows + P * wpr);
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_lut[256];
    __shared__ int s_scan[NW + 1];
    __shared__ int s_box[4];
    __shared__ double s_red[2 * NW];

    const NucInfo inf = p.info[i];
    const int o = patch_byte_offset(inf.left);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, (uint32_t)(patch_panels(P) * kPanelBytes * P));
        tma_load_patch(patch, &map, inf.left, inf.top, P, &bar);
        s_box[0] = P; s_box[1] = -1; s_box[2] = P; s_box[3] = -1;
    }
    s_lut[tid] = __fdiv_rn((float)tid, 255.0f);
    for (int k = tid; k < PH * GS; k += kTexThreads) G[k] = 0.f;
    __syncthreads();
    // ---- mask rows, bounding box, pixel list ----
    const uint32_t* gm = p.bitmask + i * (int64_t)P * wpr;
    int K = 0;
    {
        int rmin = P, rmax = -1, cmin = P, cmax = -1;
        for (int base = 0; base < P * wpr; base += kTexThreads) {
            const int k = base + tid;
            uint32_t bits = (k < P * wpr) ? gm[k] : 0u;
            if (k < P * wpr) rows[k] = bits;
            const int r = k / wpr, cb = (k - r * wpr) * 32;
            if (bits) {
                rmin = min(rmin, r); rmax = max(rmax, r);
                cmin = min(cmin, cb + __ffs(bits) - 1); cmax = max(cmax, cb + 31 - __clz(bits));
            }
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
            for (int t = 0; t < NW; ++t) {
                const int v = s_scan[t];
                wbase += (t < warp) ? v : 0;
                total += v;
            }
            int pos = K + wbase + incl - cnt;
            while (bits) {
                const int c = cb + __ffs(bits) - 1;
                bits &= bits - 1;
                list[pos++] = (uint16_t)((r << 8) | c);
            }
            K += total;
        }
        rmin = warp_min(rmin); rmax = warp_max(rmax); cmin = warp_min(cmin); cmax = warp_max(cmax);
        if (lane == 0) {
            atomicMin(&s_box[0], rmin); atomicMax(&s_box[1], rmax);
            atomicMin(&s_box[2], cmin); atomicMax(&s_box[3], cmax);
        }
    }
    __syncthreads();
    const int rmin = s_box[0], rmax = s_box[1], cmin = s_box[2], cmax = s_box[3];
    float* out = p.out + i * (int64_t)p.out_stride + p.col_gabor;
    mbar_wait(&bar, 0);   // never leave the CTA with a TMA still writing its shared memory
    if (K == 0) {   // empty mask: 0/0 = NaN (texture.rs:340-344)
        for (int k = tid; k < 2 * kGaborFilters; k += kTexThreads) out[k] = CUDART_NAN_F;
        return;
    }
    // ---- grey plane with the 'same' zero halo: only the part the bounding box can reach ----
    const int gr0 = max(rmin - kGaborLo, 0), gr1 = min(rmax + (kGaborK - 1 - kGaborLo), P - 1);
    const int gc0 = max(cmin - kGaborLo, 0), gc1 = min(cmax + (kGaborK - 1 - kGaborLo), P - 1);
    const int gw = gc1 - gc0 + 1;
    for (int k = tid; k < (gr1 - gr0 + 1) * gw; k += kTexThreads) {
        const int r = gr0 + k / gw, c = gc0 + k % gw;
        float g = 0.f;
        if (r < inf.nvr && c < inf.nvc) {
            const int a = patch_addr(P, o, r, c);
            g = grey_of(s_lut, patch[a], patch[a + 1], patch[a + 2]);
        }
        G[(r + kGaborLo) * GS + c + kGaborLo] = g;
    }
    __syncthreads();
    const int bw = cmax - cmin + 1, nrow = rmax - rmin + kGaborK;   // row pass: padded rows rmin .. rmax+29
    for (int f = 0; f < kGaborFilters; ++f) {
        // ---- row pass: A = G (*) cu, B = G (*) su over the bounding box columns ----
        for (int k = tid; k < nrow * bw; k += kTexThreads) {
            const int pr = rmin + k / bw, c = cmin + k % bw;   // padded row, patch column
            const float* g = G + pr * GS + c;                  // taps cover padded columns c .. c+29
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int t = 0; t < kGaborK; ++t) {
                const float x = g[t];
                a = fmaf(x, c_gabor[f][0][t], a);
                b = fmaf(x, c_gabor[f][1][t], b);
            }
            A[pr * P + c] = a;
            B[pr * P + c] = b;
        }
        __syncthreads();
        // ---- column pass at the masked pixels: out = A (*) cv - B (*) sv ; masked mean / variance ----
        double s[2] = {0.0, 0.0};
        for (int j = tid; j < K; j += kTexThreads) {
            const uint32_t rc = list[j];
            const float* a = A + (rc >> 8) * P + (rc & 255);   // padded rows r .. r+29
            const float* b = B + (rc >> 8) * P + (rc & 255);
            float v = 0.f;
#pragma unroll
            for (int t = 0; t < kGaborK; ++t) {
                v = fmaf(a[t * P], c_gabor[f][2][t], v);
                v = fmaf(-b[t * P], c_gabor[f][3][t], v);
            }
            s[0] += (double)v;
            s[1] += (double)v * (double)v;
        }
        block_sum<2>(s, s_red);   // also separates this filter's planes from the next row pass
        if (tid == 0) {
            const double mean = s[0] / (double)K;
            out[2 * f] = (float)mean;
            out[2 * f + 1] = (float)fmax(s[1] / (double)K - mean * mean, 0.0);
        }
    }
}

}  // namespace

int gabor_max_patch() { return 64; }

static bool g_taps_ready[64] = {};
static cudaError_t ensure_gabor_taps() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && g_taps_ready[dev]) return cudaSuccess;
    // oracle gabor_bank: taps on linspace(-1,1,30), theta = angle_idx*2*pi/8, sigma = 0.45 (texture.rs:319-334)
    static float h[kGaborFilters][4][kGaborK];
    const double freqs[6] = {0.5, 1.0, 2.0, 4.0, 6.0, 8.0}, sigma = 0.45, pi = 3.14159265358979323846;
    for (int a = 0; a < 8; ++a)
        for (int q = 0; q < 6; ++q) {
            const double th = a * 2.0 * pi / 8.0, wa = 2.0 * pi * freqs[q] * cos(th), wb = 2.0 * pi * freqs[q] * sin(th);
            for (int t = 0; t < kGaborK; ++t) {
                const double u = -1.0 + 2.0 * t / (kGaborK - 1), ge = exp(-u * u / (2.0 * sigma * sigma));
                h[a * 6 + q][0][t] = (float)(ge * cos(wa * u));
                h[a * 6 + q][1][t] = (float)(ge * sin(wa * u));
                h[a * 6 + q][2][t] = (float)(ge * cos(wb * u));
                h[a * 6 + q][3][t] = (float)(ge * sin(wb * u));
            }
        }
    e = cudaMemcpyToSymbol(c_gabor, h, sizeof(h));
    if (e == cudaSuccess && dev < 64) g_taps_ready[dev] = true;
    return e;
}

cudaError_t launch_glrlm(const TexParams& p, const CUtensorMap* map_cslab, cudaStream_t s) {
    if (p.n <= 0) return cudaSuccess;
    const int smem = window_smem_bytes(p.P, p.slab_rows) + ((p.P * p.P + 15) & ~15) + p.P * mask_wpr(p.P) * 4;
    cudaError_t e = cudaFuncSetAttribute(k_glrlm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k_glrlm<<<(unsigned)p.n, kTexThreads, smem, s>>>(p, *map_cslab);
    return cudaGetLastError();
}

cudaError_t launch_gabor(const TexParams& p, const CUtensorMap* map_patch, cudaStream_t s) {
    if (p.n <= 0) return cudaSuccess;
    cudaError_t e = ensure_gabor_taps();
    if (e != cudaSuccess) return e;
    const int P = p.P, PH = P + kGaborK - 1;
    const int smem = patch_smem_bytes(P) + PH * (P + kGaborK) * 4 + 2 * PH * P * 4 + P * mask_wpr(P) * 4 + P * P * 2;
    e = cudaFuncSetAttribute(k_gabor, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k_gabor<<<(unsigned)p.n, kTexThreads, smem, s>>>(p, *map_patch);
    return cudaGetLastError();
}

}  // namespace nfx
