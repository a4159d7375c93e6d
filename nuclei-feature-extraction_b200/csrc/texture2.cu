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
// so each of the 48 filters is two separable passes (30 taps each instead of 900): a first pass over the
// mask's bounding box plus its 29-pixel halo into shared-memory planes, and a second, register-tiled pass
// over the bounding box whose outputs are summed under the mask. Taps live in constant memory
// (warp-uniform index -> FFMA with a uniform-register operand).
#include <mutex>
#include <math.h>
#include <math_constants.h>

#include <type_traits>

#include "nfx_kernels.h"

namespace nfx {

namespace {

constexpr int kRlLevels = 24, kRlMax = 16;   // texture.rs:174-175
__device__ __constant__ int c_dirs[4][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}};   // (dx, dy) texture.rs:176

constexpr int kTexThreads = 256;
constexpr int kGaborFilters = 48, kGaborK = 30, kGaborLo = 14;   // 'same' padding: 14 before, 15 after
// Separable factors of the bank (oracle gabor_bank): per frequency q the 1-D profiles
//   [0] G(t)cos(w t)  [1] G(t)sin(w t)  with w = 2 pi f           (theta = 0 / 90 degrees)
//   [2] G(t)cos(w't)  [3] G(t)sin(w't)  with w' = 2 pi f cos(pi/4) (theta = 45 / 135 degrees)
// and the bare Gaussian envelope G(t). Rows and columns share them (the tap grid is symmetric).
__device__ __constant__ float c_gtap[6][4][kGaborK];
__device__ __constant__ float c_genv[kGaborK];
// Oblique angle pairs (theta, pi - theta): cos(a u +- b v) = [G cos au][G cos bv] -+ [G sin au][G sin bv] with a = w cos(theta),
// b = w sin(theta). [pair][q] = {G cos(a t), G sin(a t) (row pass), G cos(b t), G sin(b t) (column pass)}.
// pair 0: theta = 45 (a = b); pairs 1, 2: theta = 22.5 and 67.5 degrees -- only under NFX_RULE_GABOR_HALF_TURN, where the
// eight angles are i * pi / 8 (48 distinct filters) instead of i * 2 pi / 8 (24 distinct ones, theta and theta + pi coincide).
__device__ __constant__ float c_gobl[3][6][4][kGaborK];

__device__ __forceinline__ float grey_of(const float* lut, uint32_t r, uint32_t g, uint32_t b) {
    // texture.rs:189/332: mean_dim(-3) of u8/255 values = ((r+g)+b)/3 with IEEE f32 operations
    return __fdiv_rn(__fadd_rn(__fadd_rn(lut[r], lut[g]), lut[b]), 3.0f);
}

// ------------------------------------------------------------------------------------------------
constexpr int kRlCells = kRlLevels * kRlMax;
static_assert(kRlLevels == 24, "u = (i - 12.5)/11.5 below assumes 24 levels");

// Dynamic smem: region X (slab; windows of several slabs reuse it for the pixel list u16 (row << 8) | col) | plane[(P+2)^2] u8 | rows[P*wpr] u32 | list (one-slab windows)
// THREADS = 256 for P <= 128 (several CTAs per SM); 1024 for larger windows, whose 204 KB of shared memory leave
// room for one CTA per SM only (8 warps per SM measured 25 % issue-active at P = 256).
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_glrlm(const TexParams p, const __grid_constant__ CUtensorMap map /* box {208, CS rows} */) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = p.P, wpr = mask_wpr(P), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = THREADS / 32;
    const int CS = p.slab_rows, nslab = (P + CS - 1) / CS;
    const int64_t i = blockIdx.x;
    // region X = slab buffer while the window streams in, pixel list afterwards
    const int region_x = max(window_smem_bytes(P, CS), (P * P * 2 + 127) & ~127);
    uint8_t* slab = smem_raw;
    uint16_t* list = reinterpret_cast<uint16_t*>(smem_raw);   // (one-slab windows: moved below)
    uint8_t* plane = smem_raw + region_x;
    // plane[(r + 1) * PP + c + 1] = level of a masked pixel, 0xFF everywhere else (one-pixel border included): the run
    // walk compares ONE byte per step instead of a bounds check, a mask-word test and a level load.
    const int PP = P + 2;
    uint32_t* rows = reinterpret_cast<uint32_t*>(plane + ((PP * PP + 15) & ~15));
    // One-slab windows (P <= 64) keep the pixel list in its own region: it is built while the window is in flight and only
    // the MASKED pixels are quantised (a run never leaves the mask). Quantising every pixel of the rows that hold mask bits
    // was a fifth of the kernel's instructions (ncu round 2) for 900 of 4096 pixels that matter.
    const bool early = nslab == 1;
    if (early) list = reinterpret_cast<uint16_t*>(rows + ((P * wpr + 3) & ~3));
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_lut[256];
    __shared__ uint32_t s_R[4 * kRlCells];   // one 24 x 16 histogram per direction
    __shared__ int s_scan[NW + 1];

    const NucInfo inf = p.info[i];
    const int o = patch_byte_offset(inf.left);
    const uint32_t slab_tx = (uint32_t)(patch_panels(P) * kPanelBytes * CS);
    const float* greyp = p.grey ? p.grey + i * (int64_t)P * P : nullptr;   // f32 grey plane instead of the u8 window
    if (tid == 0 && !greyp) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, slab_tx);
        tma_load_window(slab, &map, inf.left, inf.top, P, CS, &bar);
    }
    if (tid < 256) s_lut[tid] = __fdiv_rn((float)tid, 255.0f);
    const uint32_t* gm = p.bitmask + i * (int64_t)P * wpr;
    for (int k = tid; k < P * wpr; k += THREADS) rows[k] = gm[k];
    for (int k = tid; k < ((PP * PP + 15) >> 4); k += THREADS) reinterpret_cast<uint4*>(plane)[k] = make_uint4(~0u, ~0u, ~0u, ~0u);
    __syncthreads();
    // ---- compacted list of the masked pixels ----
    int K = 0;
    auto build_list = [&]() {
        for (int base = 0; base < P * wpr; base += THREADS) {
            const int k = base + tid;
            uint32_t bits = (k < P * wpr) ? rows[k] : 0u;
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
            const int r = k / wpr, cb = (k - r * wpr) * 32;
            while (bits) {
                const int c = cb + __ffs(bits) - 1;
                bits &= bits - 1;
                list[pos++] = (uint16_t)((r << 8) | c);
            }
            K += total;
        }
        __syncthreads();
    };
    if (early) {
        build_list();
        if (!greyp) mbar_wait(&bar, 0);
        if (!greyp) {
            for (int j = tid; j < K; j += THREADS) {
                const uint32_t rc = list[j];
                const int r = rc >> 8, c = rc & 255;
                uint32_t pr = 0, pg = 0, pb = 0;
                if (r < inf.nvr && c < inf.nvc) {   // utils.rs:161-192: the rest of the window is zero
                    const int a = patch_addr(CS, o, r, c);
                    pr = slab[a]; pg = slab[a + 1]; pb = slab[a + 2];
                }
                plane[(r + 1) * PP + c + 1] = (uint8_t)min((int)floorf(__fmul_rn(grey_of(s_lut, pr, pg, pb), (float)kRlLevels)), kRlLevels - 1);
            }
        } else {   // f32 batch (f32batch.cu): the grey plane is given
            for (int j = tid; j < K; j += THREADS) {
                const uint32_t rc = list[j];
                const int r = rc >> 8, c = rc & 255;
                plane[(r + 1) * PP + c + 1] = (uint8_t)min((int)floorf(__fmul_rn(greyp[r * P + c], (float)kRlLevels)), kRlLevels - 1);
            }
        }
        __syncthreads();
    } else {
        // ---- 24-level plane (texture.rs:189 + SPEC.md B9), rows that hold mask bits only ----
        for (int sidx = 0; sidx < nslab; ++sidx) {
            const int row0 = sidx * CS, nrows = min(CS, P - row0);
            if (!greyp) mbar_wait(&bar, sidx & 1);
            for (int k = tid; k < nrows * P; k += THREADS) {
                const int lr = k / P, c = k - lr * P, r = row0 + lr;
                if (!((rows[r * wpr + (c >> 5)] >> (c & 31)) & 1u)) continue;
                uint32_t pr = 0, pg = 0, pb = 0;
                if (!greyp && r < inf.nvr && c < inf.nvc) {   // utils.rs:161-192: the rest of the window is zero
                    const int a = patch_addr(CS, o, lr, c);
                    pr = slab[a]; pg = slab[a + 1]; pb = slab[a + 2];
                }
                const float g = greyp ? greyp[r * P + c] : grey_of(s_lut, pr, pg, pb);
                plane[(r + 1) * PP + c + 1] = (uint8_t)min((int)floorf(__fmul_rn(g, (float)kRlLevels)), kRlLevels - 1);
            }
            __syncthreads();
            if (tid == 0 && !greyp && sidx + 1 < nslab) {
                mbar_expect_tx(&bar, slab_tx);
                tma_load_window(slab, &map, inf.left, inf.top + row0 + CS, P, CS, &bar);
            }
        }
        build_list();
    }
    float* out = p.out + i * (int64_t)p.out_stride + p.col_glrlm;
    // ---- run detection for the four directions: one sweep per direction, four histograms ----
    for (int k = tid; k < 4 * kRlCells; k += THREADS) s_R[k] = 0u;
    __syncthreads();
    for (int j = tid; j < K; j += THREADS) {
        const uint32_t rc = list[j];
        const int at = ((rc >> 8) + 1) * PP + (rc & 255) + 1;
        const uint32_t lv = plane[at];
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            const int step = c_dirs[d][1] * PP + c_dirs[d][0];
            if (plane[at - step] == lv) continue;   // not a run start
            int len = 1, q = at + step;
            while (plane[q] == lv) { ++len; q += step; }
            atomicAdd(&s_R[d * kRlCells + lv * kRlMax + min(len, kRlMax) - 1], 1u);
        }
    }
    __syncthreads();
    // ---- 17 features per direction: warp d owns direction d (12 cells per lane, one shuffle tree) ----
    if (warp < 4) {
        const uint32_t* R4 = s_R + warp * kRlCells;
        float v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = 0.f;
        for (int k = lane; k < kRlCells; k += 32) {
            const float R = (float)R4[k];
            if (R == 0.f) continue;
            const float fi = (float)(k / kRlMax + 1), fj = (float)(k % kRlMax + 1);
            const float ii = fi * fi, jj = fj * fj, ri = __fdiv_rn(1.0f, ii), rj = __fdiv_rn(1.0f, jj);
            const float u = (fi - 12.5f) * (1.0f / 11.5f), u2 = u * u, m2 = 1.0f - u2;
            v[0] += R;
            v[1] = fmaf(R, rj, v[1]);
            v[2] = fmaf(R, jj, v[2]);
            v[3] = fmaf(R, ri, v[3]);
            v[4] = fmaf(R, ii, v[4]);
            v[5] = fmaf(R, ri * rj, v[5]);
            v[6] = fmaf(R, ii * rj, v[6]);
            v[7] = fmaf(R, jj * ri, v[7]);
            v[8] = fmaf(R, ii * jj, v[8]);
            v[9] = fmaf(R, m2 * rj, v[9]);
            v[10] = fmaf(R, m2 * jj, v[10]);
            v[11] = fmaf(R, u2 * rj, v[11]);
            v[12] = fmaf(R, u2 * jj, v[12]);
            v[13] = fmaf(R, fj, v[13]);
        }
        if (lane < kRlLevels) {   // grey-level non-uniformity: squared row sums (exact integers)
            uint32_t rs = 0;
            for (int j = 0; j < kRlMax; ++j) rs += R4[lane * kRlMax + j];
            v[14] = (float)rs * (float)rs;
        }
        if (lane < kRlMax) {      // run-length non-uniformity: squared column sums
            uint32_t cs = 0;
            for (int l = 0; l < kRlLevels; ++l) cs += R4[l * kRlMax + lane];
            v[15] = (float)cs * (float)cs;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = warp_sum(v[q]);
        if (lane == 0) {
            float* o_ = out + warp * 17;
            const double nr = v[0], mean = (double)v[13] / nr;
            o_[0] = (float)(v[1] / nr);  o_[1] = (float)(v[2] / nr);
            o_[2] = (float)(v[14] / nr); o_[3] = (float)(v[15] / nr);
            o_[4] = (float)(v[3] / nr);  o_[5] = (float)(v[4] / nr);
            o_[6] = (float)(v[5] / nr);  o_[7] = (float)(v[6] / nr);
            o_[8] = (float)(v[7] / nr);  o_[9] = (float)(v[8] / nr);
            o_[10] = (float)(v[9] / nr); o_[11] = (float)(v[10] / nr);
            o_[12] = (float)(v[11] / nr); o_[13] = (float)(v[12] / nr);
            o_[14] = (float)(nr / (double)K);                // run percentage
            o_[15] = (float)mean;                            // run length mean
            o_[16] = (float)((double)v[2] / nr - mean * mean);   // run length variance
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_gabor. cos is even, so the kernels of theta and theta + 180 degrees are identical: 24 distinct
// filters. Per frequency f (w = 2 pi f, w' = w cos 45):
//   theta =   0: cos(w u)          -> columns with G (ONE plane for all six f), rows with G cos(w u)
//   theta =  90: cos(w v)          -> rows with G (ONE plane for all six f), columns with G cos(w v)
//                  (G cos is even: the 15 pair sums x[t] + x[29-t] are shared by the six profiles of a pass)
//   theta =  45: cos(w'u + w'v)    -> p - q   where p = (G cos w'u rows)(G cos w'v cols),
//   theta = 135: cos(-w'u + w'v)   -> p + q         q = (G sin w'u rows)(G sin w'v cols)
// Horizontal passes are register tiled (4 outputs per thread from 9 float4 loads, row index fastest across
// the lanes; the plane strides are 4 (mod 32) floats so that float4 loads and stores are bank-conflict
// free); vertical passes produce 2 or 4 outputs per thread from 31 or 33 loads down one column.
// Dynamic smem: G[(P+29)][GS] | A[(P+29)][PS] (the TMA window lands here first) | B[(P+29)][PS] | rows | slack
// (A and B together also hold the theta = 0 plane, stride GS; the slack absorbs strip loads past the last row).
__host__ __device__ constexpr int gabor_gs(int P) { return ((P + kGaborK + 31) & ~31) + 4; }   // >= P+34, = 4 mod 32
__host__ __device__ constexpr int gabor_ps(int P) { return ((P + 27) & ~31) + 4; }             // >= P,    = 4 mod 32

// TILED = false: the whole P x P window (P <= 64) is one tile and everything outside it is the zero
// padding of the reference's 'same' convolution. TILED = true (P > 64): grid.y enumerates 64 x 64 output
// tiles of the window; each CTA fetches its tile plus the 14/15-pixel halo (real pixels inside the
// window, zeros outside), runs the same passes with the planes sized for a 64-pixel tile, and writes
// per-tile partial sums (sum v, sum v^2 per distinct filter, tile pixel count) that k_gabor_finalize folds.
constexpr int kGaborTile = 64, kGaborFetch = kGaborTile + kGaborK - 1;   // 93 rows/cols fetched per tile
constexpr int kGaborDistinct = 48;                                       // distinct filters: 24 (full turn) or 48 (half turn)
constexpr int kGaborPartial = 2 * kGaborDistinct + 1;                    // (sum, sum^2) per distinct filter + pixel count

template <int kP, bool TILED, bool HALF>   // kP = 64: compile-time plane strides (immediate LDS offsets); 0: runtime; HALF: NFX_RULE_GABOR_HALF_TURN
__global__ void __launch_bounds__(kTexThreads, 2)
k_gabor(const TexParams p, const __grid_constant__ CUtensorMap map /* box {208, P rows} or {208, 93 rows} */) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int WP = p.P;                                   // the nucleus window
    const int P = TILED ? kGaborTile : (kP ? kP : p.P);   // the tile this CTA filters (local coordinates)
    const int wpr = mask_wpr(P), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kTexThreads / 32;
    const int PH = P + kGaborK - 1, GS = (kP || TILED) ? gabor_gs(kGaborTile) : gabor_gs(P),
              PS = (kP || TILED) ? gabor_ps(kGaborTile) : gabor_ps(P);
    const int64_t i = blockIdx.x;
    const int tiles_x = (WP + kGaborTile - 1) / kGaborTile;
    const int oy = TILED ? (int)(blockIdx.y / tiles_x) * kGaborTile : 0, ox = TILED ? (int)(blockIdx.y % tiles_x) * kGaborTile : 0;
    float* G = reinterpret_cast<float*>(smem_raw);
    float* A = G + ((PH * GS + 31) & ~31);
    float* B = A + ((PH * PS + 31) & ~31);
    uint8_t* patch = reinterpret_cast<uint8_t*>(A);   // the fetched pixels are consumed before A is written
    uint32_t* rows = reinterpret_cast<uint32_t*>(B + ((PH * PS + 31) & ~31));
    // (the bytes after `rows` are slack: the strip loads of the last rows may run up to 4 plane rows past B)
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_lut[256];
    __shared__ int s_box[5];
    __shared__ double s_part[NW][2 * kGaborDistinct];   // per warp: (sum, sum of squares) of the distinct filters
    // The oblique taps of the pair being filtered, [q][profile][t] (32 floats per profile: float4 loads). A phase keeps its 60
    // taps in registers; fetching them from constant memory (30 LDC.64 + 32 MOV per warp and phase, right after a barrier
    // when no other warp of the CTA can issue) drew 10 % of the kernel's stall samples (profiles/r2_tex_ncu.txt).
    __shared__ __align__(16) float s_tap[6][4][32];
    // slot of angle index k (x 6 frequencies): full turn 0, 45, 90, 135 -> 0, 6, 12, 18; half turn i * 22.5 degrees -> 6 i
    constexpr bool half_turn = HALF;   // its own instantiation: the default bank keeps compile-time table and slot indices
    const int nf = half_turn ? 48 : 24, slot0 = 0, slot90 = half_turn ? 24 : 12;

    const NucInfo inf = p.info[i];
    // fetched region: rows/cols [f0, f0 + fetch) of the window, f0 = -14 relative to the tile when TILED
    const int fy = TILED ? oy - kGaborLo : 0, fx = TILED ? ox - kGaborLo : 0;
    const int frows = TILED ? kGaborFetch : P;            // rows per panel of the fetched region
    const int o = patch_byte_offset(inf.left + fx);
    const float* greyp = p.grey ? p.grey + i * (int64_t)WP * WP : nullptr;   // f32 grey plane instead of the u8 window
    if (tid == 0) {
        if (!greyp) {
            mbar_init(&bar, 1);
            mbar_fence_init();
            const int fw = TILED ? kGaborFetch : P;
            mbar_expect_tx(&bar, (uint32_t)(patch_panels(fw) * kPanelBytes * frows));
            tma_load_window(patch, &map, inf.left + fx, inf.top + fy, fw, frows, &bar);
        }
        s_box[0] = P; s_box[1] = -1; s_box[2] = P; s_box[3] = -1; s_box[4] = 0;
    }
    s_lut[tid] = __fdiv_rn((float)tid, 255.0f);
    auto stage_taps = [&](int pair) {
        for (int k = tid; k < 6 * 4 * 32; k += kTexThreads) {
            const int t = k & 31;
            (&s_tap[0][0][0])[k] = t < kGaborK ? c_gobl[pair][k >> 7][(k >> 5) & 3][t] : 0.f;
        }
    };
    stage_taps(half_turn ? 1 : 0);
    for (int k = tid; k < PH * GS; k += kTexThreads) G[k] = 0.f;
    __syncthreads();
    // ---- mask rows of the tile, bounding box, pixel count (tile-local coordinates) ----
    const int gwpr = mask_wpr(WP);
    const uint32_t* gm = p.bitmask + i * (int64_t)WP * gwpr;
    {
        int rmin = P, rmax = -1, cmin = P, cmax = -1, cnt = 0;
        for (int k = tid; k < P * wpr; k += kTexThreads) {
            const int r = k / wpr, w = k - r * wpr, cb = w * 32;
            uint32_t bits = 0u;
            if (oy + r < WP && (ox >> 5) + w < gwpr) bits = gm[(oy + r) * gwpr + (ox >> 5) + w];
            rows[k] = bits;
            if (bits) {
                rmin = min(rmin, r); rmax = max(rmax, r);
                cmin = min(cmin, cb + __ffs(bits) - 1); cmax = max(cmax, cb + 31 - __clz(bits));
            }
            cnt += __popc(bits);
        }
        rmin = warp_min(rmin); rmax = warp_max(rmax); cmin = warp_min(cmin); cmax = warp_max(cmax);
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) {
            atomicMin(&s_box[0], rmin); atomicMax(&s_box[1], rmax);
            atomicMin(&s_box[2], cmin); atomicMax(&s_box[3], cmax);
            atomicAdd(&s_box[4], cnt);
        }
    }
    __syncthreads();
    const int rmin = s_box[0], rmax = s_box[1], cmin = s_box[2], cmax = s_box[3], K = s_box[4];
    float* out = p.out + i * (int64_t)p.out_stride + p.col_gabor;
    double* part = TILED ? p.gabor_partial + (i * (int64_t)gridDim.y + blockIdx.y) * kGaborPartial : nullptr;
    if (!greyp) mbar_wait(&bar, 0);   // never leave the CTA with a TMA still writing its shared memory
    if (K == 0) {
        if (TILED) {        // empty tile: contributes nothing
            for (int k = tid; k < kGaborPartial; k += kTexThreads) part[k] = 0.0;
        } else {            // empty mask: 0/0 = NaN (texture.rs:340-344)
            for (int k = tid; k < 2 * kGaborFilters; k += kTexThreads) out[k] = CUDART_NAN_F;
        }
        return;
    }
    if (TILED && tid == 0) part[kGaborPartial - 1] = (double)K;
    // ---- grey plane with its halo: only the part the bounding box can reach. Outside the nucleus
    //      window (and beyond what the reference copied, NucInfo) it is the zero padding of 'same' ----
    const int gr0 = rmin - kGaborLo, gr1 = rmax + (kGaborK - 1 - kGaborLo);   // tile-local rows, may leave [0,P)
    const int gc0 = cmin - kGaborLo, gc1 = cmax + (kGaborK - 1 - kGaborLo);
    const int gw = gc1 - gc0 + 1;
    const float inv_gw = 1.0f / (float)gw;
    const int wr_lim = min(WP, inf.nvr), wc_lim = min(WP, inf.nvc);
    if (!greyp) {
        for (int k = tid; k < (gr1 - gr0 + 1) * gw; k += kTexThreads) {
            const int kr = __float2int_rz(((float)k + 0.5f) * inv_gw);   // = k / gw exactly (see fdiv below)
            const int r = gr0 + kr, c = gc0 + (k - kr * gw);       // tile-local
            const int wr = oy + r, wc = ox + c;                    // window coordinates
            if (wr >= 0 && wr < wr_lim && wc >= 0 && wc < wc_lim) {
                const int a = patch_addr(frows, o, wr - fy, wc - fx);
                G[(r + kGaborLo) * GS + c + kGaborLo] = grey_of(s_lut, patch[a], patch[a + 1], patch[a + 2]);
            }
        }
    } else {   // f32 batch (f32batch.cu): the grey plane is given
        for (int k = tid; k < (gr1 - gr0 + 1) * gw; k += kTexThreads) {
            const int kr = __float2int_rz(((float)k + 0.5f) * inv_gw);
            const int r = gr0 + kr, c = gc0 + (k - kr * gw), wr = oy + r, wc = ox + c;
            if (wr >= 0 && wr < wr_lim && wc >= 0 && wc < wc_lim) G[(r + kGaborLo) * GS + c + kGaborLo] = greyp[wr * WP + wc];
        }
    }
    __syncthreads();   // the fetched pixels (aliased with A) are dead from here on

    const int cq0 = cmin & ~3, nquad = ((cmax - cq0) >> 2) + 1;   // column quads of the bounding box
    const int nrow = rmax - rmin + kGaborK;                       // padded rows rmin .. rmax+29
    const int rh = rmax - rmin + 1, ncol = cmax - cmin + 1;
    auto mbit = [&](int r, int c) -> bool { return (rows[r * wpr + (c >> 5)] >> (c & 31)) & 1u; };
    // k / d for the item loops (k < 2^12, d < 2^7) without the integer-division sequence: exact, because
    // (k + 0.5) / d is at least 0.5/d away from an integer while the f32 product is off by < 2^-10.
    auto fdiv = [](int k, float inv) -> int { return __float2int_rz(((float)k + 0.5f) * inv); };
    const float inv_nrow = 1.0f / (float)nrow, inv_rh = 1.0f / (float)rh, inv_ncol = 1.0f / (float)ncol,
                inv_npc = 1.0f / (float)(ncol + kGaborK - 1);
    // warp partials of NV doubles -> s_part[warp][slot0 ..): every slot is written by exactly one phase
    auto stash2 = [&](double a, double b, int slot0) {
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) { s_part[warp][slot0] = a; s_part[warp][slot0 + 1] = b; }
    };

    // Horizontal FIR, 4 outputs per thread from 9 float4 loads (row index fastest across the lanes):
    //   dA[pr][c0+m] = sum_t src[pr][c0+m+t] * ta[t]   (and dB with tb when TWO), pr = row_lo .. row_lo+nr-1.
    auto h_store = [&](const float* src, int ss, float* dA, float* dB, int ds, bool env, int q, int row_lo, int nr, float inv_nr) {
        const bool two = !env;
        float wa[32], wb[32];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const float4 u = reinterpret_cast<const float4*>(s_tap[q][0])[m], v = reinterpret_cast<const float4*>(s_tap[q][1])[m];
            wa[4 * m] = u.x; wa[4 * m + 1] = u.y; wa[4 * m + 2] = u.z; wa[4 * m + 3] = u.w;
            wb[4 * m] = v.x; wb[4 * m + 1] = v.y; wb[4 * m + 2] = v.z; wb[4 * m + 3] = v.w;
        }
        if (env) {
#pragma unroll
            for (int t = 0; t < kGaborK; ++t) wa[t] = c_genv[t];
        }
        for (int k = tid; k < nr * nquad; k += kTexThreads) {
            const int kq = fdiv(k, inv_nr), pr = row_lo + (k - kq * nr), c0 = cq0 + 4 * kq;
            const float4* g4 = reinterpret_cast<const float4*>(src + pr * ss + c0);
            float x[36];
#pragma unroll
            for (int m = 0; m < 9; ++m) {
                const float4 v = g4[m];
                x[4 * m] = v.x; x[4 * m + 1] = v.y; x[4 * m + 2] = v.z; x[4 * m + 3] = v.w;
            }
            float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int t = 0; t < kGaborK; ++t) {
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    a[m] = fmaf(x[m + t], wa[t], a[m]);
                    if (two) b[m] = fmaf(x[m + t], wb[t], b[m]);
                }
            }
            *reinterpret_cast<float4*>(dA + pr * ds + c0) = make_float4(a[0], a[1], a[2], a[3]);
            if (two) *reinterpret_cast<float4*>(dB + pr * ds + c0) = make_float4(b[0], b[1], b[2], b[3]);
        }
    };

    // ======== theta = 90 (filters 12..17): rows with the envelope (ONE plane), then a vertical pass with the six
    // cos(w v) profiles. The profiles are even (tap[t] = tap[29-t]): the 15 pair sums x[t] + x[29-t] are shared by
    // the six filters (15 adds + 6 x 15 FMAs instead of 180 FMAs). SH output rows per thread share their loads. ====
    h_store(G, GS, B, B, PS, true, 0, rmin, nrow, inv_nrow);
    __syncthreads();
    auto v_six = [&](auto sh_tag) {
        constexpr int SH = decltype(sh_tag)::value;
        double s[12];
#pragma unroll
        for (int h = 0; h < 12; ++h) s[h] = 0.0;
        const int nstrip = (rh + SH - 1) / SH;
        for (int k = tid; k < nstrip * ncol; k += kTexThreads) {
            const int ks = fdiv(k, inv_ncol), c = cmin + (k - ks * ncol), r0 = rmin + SH * ks;
            const float* src = B + r0 * PS + c;
            float x[kGaborK - 1 + SH];
#pragma unroll
            for (int j = 0; j < kGaborK - 1 + SH; ++j) x[j] = src[j * PS];
#pragma unroll
            for (int m = 0; m < SH; ++m) {
                float e[15];
#pragma unroll
                for (int t = 0; t < 15; ++t) e[t] = x[m + t] + x[m + 29 - t];
                const bool in = (r0 + m <= rmax) && mbit(r0 + m, c);
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    float v0 = 0.f, v1 = 0.f;
#pragma unroll
                    for (int t = 0; t < 15; ++t) {
                        if (t & 1) v1 = fmaf(e[t], c_gtap[q][0][t], v1);
                        else v0 = fmaf(e[t], c_gtap[q][0][t], v0);
                    }
                    const double v = in ? (double)(v0 + v1) : 0.0;
                    s[2 * q] += v; s[2 * q + 1] += v * v;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) stash2(s[2 * q], s[2 * q + 1], 2 * (slot90 + q));
    };
    // Strips of 4 rows for every box size. Round 2 measured 2-row strips (better thread fill on small boxes, but v_diag's
    // 2-row form is bound by the shared-memory pipe: 63 loads per 120 FMAs) and 8-row strips (37 loads per 480 FMAs, half
    // as many items): choosing per nucleus by a cost model was slower than always 4 (12.70 / 12.54 against 12.39 ms per 100k).
    v_six(std::integral_constant<int, 4>{});
    __syncthreads();

    // ======== theta = 0 (filters 0..5): columns with the envelope first (one plane over the padded columns the box
    // can reach, stride GS, spanning the A and B regions), then a horizontal pass with the six cos(w u) profiles. ====
    float* V = A;
    {
        const int npc = ncol + kGaborK - 1;   // padded columns cmin .. cmax+29
        const int nstrip = (rh + 3) / 4;
        for (int k = tid; k < nstrip * npc; k += kTexThreads) {
            const int ks = fdiv(k, inv_npc), pc = cmin + (k - ks * npc), r0 = rmin + 4 * ks;
            const float* src = G + r0 * GS + pc;
            float x[kGaborK + 3];
#pragma unroll
            for (int j = 0; j < kGaborK + 3; ++j) x[j] = src[j * GS];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                float v0 = 0.f, v1 = 0.f;
#pragma unroll
                for (int t = 0; t < kGaborK; ++t) {
                    if (t & 1) v1 = fmaf(x[m + t], c_genv[t], v1);
                    else v0 = fmaf(x[m + t], c_genv[t], v0);
                }
                V[(r0 + m) * GS + pc] = v0 + v1;
            }
        }
    }
    __syncthreads();
    {
        double s[12];
#pragma unroll
        for (int h = 0; h < 12; ++h) s[h] = 0.0;
        for (int k = tid; k < rh * nquad; k += kTexThreads) {
            const int kq = fdiv(k, inv_rh), r = rmin + (k - kq * rh), c0 = cq0 + 4 * kq;
            const float4* g4 = reinterpret_cast<const float4*>(V + r * GS + c0);
            float x[36];
#pragma unroll
            for (int m = 0; m < 9; ++m) {
                const float4 v = g4[m];
                x[4 * m] = v.x; x[4 * m + 1] = v.y; x[4 * m + 2] = v.z; x[4 * m + 3] = v.w;
            }
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                float e[15];
#pragma unroll
                for (int t = 0; t < 15; ++t) e[t] = x[m + t] + x[m + 29 - t];
                const bool in = mbit(r, c0 + m);
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    float v0 = 0.f, v1 = 0.f;
#pragma unroll
                    for (int t = 0; t < 15; ++t) {
                        if (t & 1) v1 = fmaf(e[t], c_gtap[q][0][t], v1);
                        else v0 = fmaf(e[t], c_gtap[q][0][t], v0);
                    }
                    const double v = in ? (double)(v0 + v1) : 0.0;
                    s[2 * q] += v; s[2 * q + 1] += v * v;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) stash2(s[2 * q], s[2 * q + 1], 2 * (slot0 + q));
    }
    __syncthreads();

    // ======== theta = 45 (filter 6+q) and 135 (filter 18+q): rows with cos(w'u) -> A and sin(w'u) -> B in one pass,
    // then p = (A columns, cos w'v), q = (B columns, sin w'v): the two filters are p - q and p + q. ========
    auto v_diag = [&](auto sh_tag, int q, int slot_minus, int slot_plus) {
        constexpr int SH = decltype(sh_tag)::value;
        float wc[32], ws[32];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const float4 u = reinterpret_cast<const float4*>(s_tap[q][2])[m], v = reinterpret_cast<const float4*>(s_tap[q][3])[m];
            wc[4 * m] = u.x; wc[4 * m + 1] = u.y; wc[4 * m + 2] = u.z; wc[4 * m + 3] = u.w;
            ws[4 * m] = v.x; ws[4 * m + 1] = v.y; ws[4 * m + 2] = v.z; ws[4 * m + 3] = v.w;
        }
        double s[4] = {0.0, 0.0, 0.0, 0.0};
        const int nstrip = (rh + SH - 1) / SH;
        for (int k = tid; k < nstrip * ncol; k += kTexThreads) {
            const int ks = fdiv(k, inv_ncol), c = cmin + (k - ks * ncol), r0 = rmin + SH * ks;
            float pp[SH], qq[SH];
            {
                const float* src = A + r0 * PS + c;
                float x[kGaborK - 1 + SH];
#pragma unroll
                for (int j = 0; j < kGaborK - 1 + SH; ++j) x[j] = src[j * PS];
#pragma unroll
                for (int m = 0; m < SH; ++m) {
                    float v0 = 0.f, v1 = 0.f;
#pragma unroll
                    for (int t = 0; t < kGaborK; ++t) {
                        if (t & 1) v1 = fmaf(x[m + t], wc[t], v1);
                        else v0 = fmaf(x[m + t], wc[t], v0);
                    }
                    pp[m] = v0 + v1;
                }
            }
            {
                const float* src = B + r0 * PS + c;
                float x[kGaborK - 1 + SH];
#pragma unroll
                for (int j = 0; j < kGaborK - 1 + SH; ++j) x[j] = src[j * PS];
#pragma unroll
                for (int m = 0; m < SH; ++m) {
                    float v0 = 0.f, v1 = 0.f;
#pragma unroll
                    for (int t = 0; t < kGaborK; ++t) {
                        if (t & 1) v1 = fmaf(x[m + t], ws[t], v1);
                        else v0 = fmaf(x[m + t], ws[t], v0);
                    }
                    qq[m] = v0 + v1;
                }
            }
#pragma unroll
            for (int m = 0; m < SH; ++m) {
                const bool in = (r0 + m <= rmax) && mbit(r0 + m, c);
                const double v45 = in ? (double)(pp[m] - qq[m]) : 0.0, v135 = in ? (double)(pp[m] + qq[m]) : 0.0;
                s[0] += v45;  s[1] += v45 * v45;
                s[2] += v135; s[3] += v135 * v135;
            }
        }
        stash2(s[0], s[1], 2 * (slot_minus + q));
        stash2(s[2], s[3], 2 * (slot_plus + q));
    };
    // full turn: the one oblique pair (45, 135) -> slots 6, 18. Half turn: (22.5, 157.5), (45, 135), (67.5, 112.5) -> 6 i.
    constexpr int npair = half_turn ? 3 : 1;
    for (int pi = 0; pi < npair; ++pi) {
        const int pair = half_turn ? (pi == 0 ? 1 : (pi == 1 ? 0 : 2)) : 0;
        const int sm = half_turn ? (pair == 1 ? 6 : (pair == 0 ? 12 : 18)) : 6;     // theta      : p - q
        const int sp = half_turn ? (pair == 1 ? 42 : (pair == 0 ? 36 : 30)) : 18;   // pi - theta : p + q
        if (pi > 0) { stage_taps(pair); __syncthreads(); }   // (the previous pair's last phase ended with a barrier)
        for (int q = 0; q < 6; ++q) {
            h_store(G, GS, A, B, PS, false, q, rmin, nrow, inv_nrow);
            __syncthreads();
            v_diag(std::integral_constant<int, 4>{}, q, sm, sp);
            __syncthreads();
        }
    }

    // ---- fold the warp partials in a fixed order: one thread per distinct filter ----
    if (tid < nf) {
        const int f = tid;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) { s0 += s_part[w][2 * f]; s1 += s_part[w][2 * f + 1]; }
        if (TILED) {
            part[2 * f] = s0;
            part[2 * f + 1] = s1;
        } else {
            const double Kd = (double)K, mean = s0 / Kd;
            const float mf = (float)mean, vf = (float)fmax(s1 / Kd - mean * mean, 0.0);
            out[2 * f] = mf; out[2 * f + 1] = vf;
            if (!half_turn) { out[2 * (f + 24)] = mf; out[2 * (f + 24) + 1] = vf; }   // theta + 180 degrees: same kernel
        }
    }
}

// one thread per (nucleus, distinct filter): fold the tiles, write theta and theta + 180 degrees
__global__ void k_gabor_finalize(const TexParams p, const int ntiles) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nf = p.gabor_half_turn ? 48 : 24;
    if (g >= p.n * nf) return;
    const int64_t i = g / nf;
    const int f = (int)(g - i * nf);
    const double* part = p.gabor_partial + i * (int64_t)ntiles * kGaborPartial;
    double s0 = 0.0, s1 = 0.0, K = 0.0;
    for (int t = 0; t < ntiles; ++t) {
        s0 += part[t * kGaborPartial + 2 * f];
        s1 += part[t * kGaborPartial + 2 * f + 1];
        K += part[t * kGaborPartial + kGaborPartial - 1];
    }
    float* out = p.out + i * (int64_t)p.out_stride + p.col_gabor;
    float mf = CUDART_NAN_F, vf = CUDART_NAN_F;   // empty mask: 0/0 (texture.rs:340-344)
    if (K > 0.0) {
        const double mean = s0 / K;
        mf = (float)mean;
        vf = (float)fmax(s1 / K - mean * mean, 0.0);
    }
    out[2 * f] = mf; out[2 * f + 1] = vf;
    if (!p.gabor_half_turn) { out[2 * (f + 24)] = mf; out[2 * (f + 24) + 1] = vf; }
}

}  // namespace

int gabor_max_patch() { return 256; }
int gabor_tiles(int P) { return P <= kGaborTile ? 1 : ((P + kGaborTile - 1) / kGaborTile) * ((P + kGaborTile - 1) / kGaborTile); }
int gabor_fetch_rows() { return kGaborFetch; }

// One copy of the taps per device, shared by every context on it: written once under a lock and completed before the flag
// is set (the streams of other contexts are not ordered after the copy).
static std::mutex g_taps_mu;
static bool g_taps_ready[64] = {};
static cudaError_t ensure_gabor_taps() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(g_taps_mu);
    if (dev < 64 && g_taps_ready[dev]) return cudaSuccess;
    // oracle gabor_bank: taps on linspace(-1,1,30), theta = angle_idx*2*pi/8, sigma = 0.45 (texture.rs:319-334)
    static float h[6][4][kGaborK], env[kGaborK];
    const double freqs[6] = {0.5, 1.0, 2.0, 4.0, 6.0, 8.0}, sigma = 0.45, pi = 3.14159265358979323846;
    for (int t = 0; t < kGaborK; ++t) {
        const double u = -1.0 + 2.0 * t / (kGaborK - 1), ge = exp(-u * u / (2.0 * sigma * sigma));
        env[t] = (float)ge;
        for (int q = 0; q < 6; ++q) {
            const double w = 2.0 * pi * freqs[q], w2 = w * cos(pi / 4.0);
            h[q][0][t] = (float)(ge * cos(w * u));
            h[q][1][t] = (float)(ge * sin(w * u));
            h[q][2][t] = (float)(ge * cos(w2 * u));
            h[q][3][t] = (float)(ge * sin(w2 * u));
        }
    }
    static float ho[3][6][4][kGaborK];
    const double th[3] = {pi / 4.0, pi / 8.0, 3.0 * pi / 8.0};
    for (int a = 0; a < 3; ++a)
        for (int q = 0; q < 6; ++q)
            for (int t = 0; t < kGaborK; ++t) {
                const double u = -1.0 + 2.0 * t / (kGaborK - 1), ge = exp(-u * u / (2.0 * sigma * sigma));
                const double w = 2.0 * pi * freqs[q], wa = w * cos(th[a]), wb = w * sin(th[a]);
                ho[a][q][0][t] = (float)(ge * cos(wa * u));
                ho[a][q][1][t] = (float)(ge * sin(wa * u));
                ho[a][q][2][t] = (float)(ge * cos(wb * u));
                ho[a][q][3][t] = (float)(ge * sin(wb * u));
            }
    e = cudaMemcpyToSymbol(c_gtap, h, sizeof(h));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_gobl, ho, sizeof(ho));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_genv, env, sizeof(env));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess && dev < 64) g_taps_ready[dev] = true;
    return e;
}

cudaError_t launch_glrlm(const TexParams& p, const CUtensorMap* map_cslab, cudaStream_t s) {
    if (p.n <= 0) return cudaSuccess;
    cudaError_t e;
    const int rx = window_smem_bytes(p.P, p.slab_rows) > ((p.P * p.P * 2 + 127) & ~127) ? window_smem_bytes(p.P, p.slab_rows) : ((p.P * p.P * 2 + 127) & ~127);
    const int smem = rx + (((p.P + 2) * (p.P + 2) + 15) & ~15) + ((p.P * mask_wpr(p.P) + 3) & ~3) * 4 + (p.P <= p.slab_rows ? p.P * p.P * 2 : 0);   // one-slab windows: + the list
    if (p.P > 128) {
        e = cudaFuncSetAttribute(k_glrlm<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k_glrlm<1024><<<(unsigned)p.n, 1024, smem, s>>>(p, *map_cslab);
        return cudaGetLastError();
    }
    e = cudaFuncSetAttribute(k_glrlm<kTexThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k_glrlm<kTexThreads><<<(unsigned)p.n, kTexThreads, smem, s>>>(p, *map_cslab);
    return cudaGetLastError();
}

// P <= 64: `map` is the whole-window map (box height P). P > 64: `map` is the 93-row halo map and
// p.gabor_partial holds n * gabor_tiles(P) * 49 doubles.
cudaError_t launch_gabor(const TexParams& p, const CUtensorMap* map, cudaStream_t s) {
    if (p.n <= 0) return cudaSuccess;
    cudaError_t e = ensure_gabor_taps();
    if (e != cudaSuccess) return e;
    const bool tiled = p.P > kGaborTile;
    const int P = tiled ? kGaborTile : p.P, PH = P + kGaborK - 1;
    const int GS = (tiled || P == 64) ? gabor_gs(kGaborTile) : gabor_gs(P), PS = (tiled || P == 64) ? gabor_ps(kGaborTile) : gabor_ps(P);
    const int smem = (((PH * GS + 31) & ~31) + 2 * ((PH * PS + 31) & ~31)) * 4 + P * mask_wpr(P) * 4 + P * P * 2;
    const int ntiles = gabor_tiles(p.P);
    auto go = [&](auto kern) -> cudaError_t {
        cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e2 != cudaSuccess) return e2;
        kern<<<dim3((unsigned)p.n, (unsigned)ntiles), kTexThreads, smem, s>>>(p, *map);
        return cudaGetLastError();
    };
    const bool half = p.gabor_half_turn != 0;
    if (!tiled) {
        if (half) return p.P == 64 ? go(k_gabor<64, false, true>) : go(k_gabor<0, false, true>);
        return p.P == 64 ? go(k_gabor<64, false, false>) : go(k_gabor<0, false, false>);
    }
    e = half ? go(k_gabor<64, true, true>) : go(k_gabor<64, true, false>);
    if (e != cudaSuccess) return e;
    k_gabor_finalize<<<(unsigned)((p.n * (p.gabor_half_turn ? 48 : 24) + 255) / 256), 256, 0, s>>>(p, ntiles);
    return cudaGetLastError();
}

}  // namespace nfx
