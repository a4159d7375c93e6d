// color.cu -- north-star kernels (1) patch gather (fused: one TMA box per nucleus, zero fill at the
// borders) and (3) masked colour/intensity statistics.
//
// Replaces ColorFeatureSet::compute_features_batched (src/features/color.rs:10-102):
//   hsv_from_rgb / hed_from_rgb   color.rs:45-46   (oracle/SPEC.md B3, B4)
//   mean_std x3                   color.rs:47-53, 117-134
//   circular_mean (batch coupled) color.rs:50-51, 144-155  -> k_hue_batch + k_hue_finalize
// and the gather of src/utils.rs:159-192 (the u8 window goes tile -> shared memory by TMA and is
// never materialised in HBM, neither as u8 nor as the reference's f32 [N,3,P,P]).
//
// k_color   : one CTA per nucleus; masked single-pass sums (exact integers for RGB and V, pivoted
//             f32 for S, H and HED), warp-shuffle + shared-memory reduction, 17 columns.
// k_hue_batch: the reference's `h.cos() * mask` broadcasts [N,P,P]*[N,1,P,P] to [N,N,P,P]: mean_h[i]
//             sums the hue of EVERY patch of the batch under mask i. It factorises into one P x P
//             image per batch, C[p] = sum_j cos h_j[p], S[p] = sum_j sin h_j[p], followed by masked
//             sums. One CTA per (batch, row slab): a TMA producer warp streams the slab of each of
//             the batch's patches through a 4-stage mbarrier ring (two patches per stage), 8 consumer
//             warps keep C,S in registers, then each warp folds the slab under the masks of its share
//             of nuclei (row prefix sums in fixed point, one step per run of mask bits).
// k_color_warp: the P = 64 form of k_color, one warp per nucleus.
// k_hue_finalize: sums slab partials in fixed order (bit-reproducible for any GPU count) + atan2.
#include <math_constants.h>

#include <mutex>

#include "nfx_kernels.h"

namespace nfx {

namespace {

constexpr int kColorThreads = 128;   // 64 threads (18 warps per SM): 0.56 -> 0.66 ms per 100k nuclei
// k_hue_batch<NCW>: NCW consumer warps + 1 TMA producer warp; a slab holds <= 256 pixel quads. NCW = 8 (one quad =
// 4 px per thread, 40 registers) is what is launched: 45 warps per SM instead of the 30 of NCW = 4 (0.665 -> 0.637 ms per
// 100 000 nuclei in round 1), and a lone (chunk, slab) CTA on an SM -- a trait-level call is ONE chunk -- no longer leaves
// each scheduler with a single warp that issues every fourth cycle (86 -> 66 us per 100 patches). Every pixel adds its
// patches in the same order whatever NCW is, so the results do not depend on it.
constexpr int kHueMaxQuads = 256;
constexpr int kHueChunk = 128;                   // nuclei whose NucInfo is staged in smem at a time (even: pairs never straddle)

// OD(v) = ln(max(v/255, 1e-6)) / ln(1e-6), f32 (SPEC.md B4); filled once per process.
__device__ float g_od_lut[256];

__global__ void k_init_od_lut() {
    const int v = threadIdx.x;
    const float x = fmaxf(__fdiv_rn((float)v, 255.0f), 1e-6f);
    g_od_lut[v] = __fdiv_rn(logf(x), -13.815510749816895f);
}

// inv([[0.65,0.70,0.29],[0.07,0.99,0.11],[0.27,0.57,0.78]]) in f32, M[k][c] (oracle HED_FROM_RGB)
#define HED_M00 1.87798273563385f
#define HED_M01 -1.0076786279678345f
#define HED_M02 -0.5561158061027527f
#define HED_M10 -0.06590805947780609f
#define HED_M11 1.134730339050293f
#define HED_M12 -0.135521799325943f
#define HED_M20 -0.6019073724746704f
#define HED_M21 -0.48041418194770813f
#define HED_M22 1.5735880136489868f

__device__ __forceinline__ float u8f(uint32_t v) {   // exact; compiles to one I2FP (full-rate ALU op on sm_100)
    return (float)(int)v;
}

struct Px {
    uint32_t r, g, b;
};
// Shared-memory loads from a 32-bit shared address kept in a register (a generic pointer to a static __shared__ array is
// re-derived from SR_CgaCtaId inside the loop otherwise).
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));   // read-only table: may be scheduled freely
    return v;
}
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t a, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}

// one lane of the (converged) warp, without keeping the lane index in a register
__device__ __forceinline__ bool elect_one() {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok));
    return ok != 0;
}

__device__ __forceinline__ float rcp_approx(float x) {   // MUFU.RCP, ~1 ulp
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// hexcone hue in sextants (SPEC.md B3): H = 60 t. Branch-free; d = max - min may be 0 (then t = 0).
// WRAP=true folds t into [0,6] like the reference's `mod 6`; the trigonometric consumers skip it.
template <bool WRAP>
__device__ __forceinline__ float hue_sextant(const Px& p, uint32_t mx, uint32_t d) {
    const bool isr = (mx == p.r), isg = (mx == p.g);
    const int gb = (int)p.g - (int)p.b, br = (int)p.b - (int)p.r, rg = (int)p.r - (int)p.g;
    const int num = isr ? gb : (isg ? br : rg);
    const float offs = isr ? 0.0f : (isg ? 2.0f : 4.0f);
    float t = fmaf((float)num, rcp_approx(u8f(max(d, 1u))), offs);
    if (WRAP) t += (t < 0.f) ? 6.0f : 0.0f;
    return t;
}

struct HsvHed {
    float h, s;         // degrees, [0,1]
    float hed[3];
    uint32_t mx;        // V * 255
};
__device__ __forceinline__ HsvHed convert(const Px& p, const float* lut) {
    HsvHed o;
    const uint32_t mx = max(p.r, max(p.g, p.b)), mn = min(p.r, min(p.g, p.b)), d = mx - mn;
    o.mx = mx;
    o.s = u8f(d) * rcp_approx(u8f(max(mx, 1u)));
    o.h = 60.0f * hue_sextant<true>(p, mx, d);
    const float a = lut[p.r], b = lut[p.g], c = lut[p.b];
    o.hed[0] = fmaxf(0.f, a * HED_M00 + b * HED_M10 + c * HED_M20);
    o.hed[1] = fmaxf(0.f, a * HED_M01 + b * HED_M11 + c * HED_M21);
    o.hed[2] = fmaxf(0.f, a * HED_M02 + b * HED_M12 + c * HED_M22);
    return o;
}

// ------------------------------------------------------------------------------------------------
// The window is processed in row slabs of CS = min(P, 64) rows (one slab for the default P = 64):
// dynamic smem = slab[panels*208*CS] | rows[P*wpr] u32 | lut[256] f32 | list[CS*P] u16 (byte address of the pixel in the slab).
// The slab holding the patch centre goes first because its centre pixel provides the pivots.
__global__ void __launch_bounds__(kColorThreads, 8)
k_color(const ColorParams p, const __grid_constant__ CUtensorMap map) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = p.P, wpr = mask_wpr(P), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int CS = p.slab_rows, nslab = (P + CS - 1) / CS;
    const int slab_bytes = window_smem_bytes(P, CS);
    const uint32_t slab_tx = (uint32_t)(patch_panels(P) * kPanelBytes * CS);
    constexpr int NW = kColorThreads / 32;
    const int64_t i = blockIdx.x;
    uint8_t* patch = smem_raw;
    uint32_t* rows = reinterpret_cast<uint32_t*>(smem_raw + slab_bytes);
    float* lut = reinterpret_cast<float*>(rows + P * wpr);
    uint16_t* list = reinterpret_cast<uint16_t*>(lut + 256);
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_scan[NW + 1];
    __shared__ uint32_t s_ri[NW][9];
    __shared__ float s_rf[NW][10];
    __shared__ float s_pv[5];

    const NucInfo inf = p.info[i];
    const int o = patch_byte_offset(inf.left);
    const int centre_slab = (P / 2) / CS;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, slab_tx);   // first slab: in flight while the mask is fetched
        tma_load_window(patch, &map, inf.left, inf.top + centre_slab * CS, P, CS, &bar);
    }
    for (int k = tid; k < 256; k += kColorThreads) lut[k] = g_od_lut[k];
    const uint32_t* gm = p.bitmask + i * (int64_t)P * wpr;
    for (int k = tid; k < P * wpr; k += kColorThreads) rows[k] = gm[k];
    __syncthreads();

    HsvHed pv;
    int Ktot = 0;
    uint32_t sr = 0, sg = 0, sb = 0, srr = 0, sgg = 0, sbb = 0, sv = 0, svv = 0;
    float s1[5] = {0, 0, 0, 0, 0}, s2[5] = {0, 0, 0, 0, 0};   // hed0, hed1, hed2, s, h (pivoted)
    for (int it = 0; it < nslab; ++it) {
        const int sidx = (it == 0) ? centre_slab : (it <= centre_slab ? it - 1 : it);
        const int row0 = sidx * CS, nrows = min(CS, P - row0);
        if (tid == 0 && it > 0) {
            mbar_expect_tx(&bar, slab_tx);
            tma_load_window(patch, &map, inf.left, inf.top + row0, P, CS, &bar);
        }
        // ---- while the slab is in flight: compact its mask rows into a list of pixel coordinates ----
        int K = 0;
        for (int base = 0; base < nrows * wpr; base += kColorThreads) {
            const int k = base + tid;
            uint32_t bits = (k < nrows * wpr) ? rows[row0 * wpr + k] : 0u;
            const int cnt = __popc(bits);
            int incl = cnt;
#pragma unroll
            for (int o2 = 1; o2 < 32; o2 <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o2);
                if (lane >= o2) incl += t;
            }
            __syncthreads();   // s_scan reuse
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
            const int abase = patch_addr(CS, 0, r, cb);   // a mask word never straddles a 64-pixel panel
            while (bits) {
                const int c = __ffs(bits) - 1;
                bits &= bits - 1;
                list[pos++] = (uint16_t)(abase + 3 * c);   // byte address of the pixel in the slab (without o): < 53 KB
            }
            K += total;
        }
        Ktot += K;
        __syncthreads();   // list visible
        mbar_wait(&bar, it & 1);
        if (inf.nvc < P || inf.nvr < row0 + nrows) {   // rare: part of the window is never copied (NucInfo)
            for (int k = tid; k < nrows * P; k += kColorThreads) {
                const int r = k / P, c = k - r * P;
                if (row0 + r >= inf.nvr || c >= inf.nvc) {
                    const int a = patch_addr(CS, o, r, c);
                    patch[a] = 0; patch[a + 1] = 0; patch[a + 2] = 0;
                }
            }
            fence_proxy_async();   // these generic-proxy stores precede the next slab's TMA (async proxy) write of the same bytes
            __syncthreads();
        }
        if (it == 0) {   // pivots (any value of the right magnitude removes the one-pass cancellation): one thread converts
            if (tid == 0) {
                const int a = patch_addr(CS, o, P / 2 - row0, P / 2);
                Px c = {patch[a], patch[a + 1], patch[a + 2]};
                const HsvHed t = convert(c, lut);
                s_pv[0] = t.h; s_pv[1] = t.s; s_pv[2] = t.hed[0]; s_pv[3] = t.hed[1]; s_pv[4] = t.hed[2];
            }
            __syncthreads();
            pv.h = s_pv[0]; pv.s = s_pv[1]; pv.hed[0] = s_pv[2]; pv.hed[1] = s_pv[3]; pv.hed[2] = s_pv[4];
        }
        const uint8_t* pbase = patch + o;
        for (int j = tid; j < K; j += kColorThreads) {
            const uint8_t* pp = pbase + list[j];
            const Px px = {pp[0], pp[1], pp[2]};
            const HsvHed c = convert(px, lut);
            sr += px.r; sg += px.g; sb += px.b;
            srr += px.r * px.r; sgg += px.g * px.g; sbb += px.b * px.b;
            sv += c.mx; svv += c.mx * c.mx;
            float d;
#pragma unroll
            for (int q = 0; q < 3; ++q) { d = c.hed[q] - pv.hed[q]; s1[q] += d; s2[q] = fmaf(d, d, s2[q]); }
            d = c.s - pv.s; s1[3] += d; s2[3] = fmaf(d, d, s2[3]);
            d = c.h - pv.h; s1[4] += d; s2[4] = fmaf(d, d, s2[4]);
        }
        if (it + 1 < nslab) __syncthreads();   // slab buffer and list are reused
    }
    const int K = Ktot;
    // ---- exact integer sums: one REDUX per warp; float sums: folded across the warps through shared
    //      memory first, then ONE warp runs the shuffle tree (4x fewer shuffles than a tree per warp) ----
    {
        const uint32_t vi[8] = {sr, sg, sb, srr, sgg, sbb, sv, svv};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t t = __reduce_add_sync(0xffffffffu, vi[q]);
            if (lane == 0) s_ri[warp][q] = t;
        }
        __syncthreads();   // the pixel list is dead: reuse it as float scratch [10][kColorThreads]
        float* fs = reinterpret_cast<float*>(list);
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            fs[q * kColorThreads + tid] = s1[q];
            fs[(5 + q) * kColorThreads + tid] = s2[q];
        }
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int q = 0; q < 10; ++q) {
                float a = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) a += fs[q * kColorThreads + w * 32 + lane];
                a = warp_sum(a);
                if (lane == 0) s_rf[0][q] = a;
            }
        }
    }
    __syncthreads();
    // ---- 17 columns, one lane each (out[6] = mean_h belongs to k_hue_finalize). Same f64 expressions for
    //      every column: m = s/K ; mean = (pivot + m)/den ; std = sqrt(max(ss/K - m^2, 0))/den ----
    if (warp == 0 && lane < 18 && lane != 6) {
        // column -> index of its sum and of its sum of squares in v[] (0 = unused), v = [K, 8 integer sums, s1[5], s2[5]]
        //              mean_r g  b  std_r g  b  -  mean_s mean_v std_h std_s std_v mean_hed0 1   2  std_hed0 1   2
        static const int ia[18] = {1, 2, 3, 1, 2, 3, 0, 12, 7, 13, 12, 7, 9, 10, 11, 9, 10, 11};
        static const int ib[18] = {0, 0, 0, 4, 5, 6, 0, 0, 0, 18, 17, 8, 0, 0, 0, 14, 15, 16};
        auto fetch = [&](int q) -> double {   // q in 1..18
            if (q <= 8) {
                unsigned long long t = 0;
#pragma unroll
                for (int w = 0; w < NW; ++w) t += s_ri[w][q - 1];
                return (double)t;
            }
            return (double)s_rf[0][q - 9];
        };
        const int a = ia[lane], b = ib[lane];
        const bool is_std = b != 0;
        const bool is8 = a <= 8;                                   // u8-valued channel: scale by 1/255
        const float pivf = lane == 7 ? pv.s : (lane == 12 ? pv.hed[0] : (lane == 13 ? pv.hed[1] : (lane == 14 ? pv.hed[2] : 0.f)));
        const double Kd = (double)K;
        const double m = fetch(a) / Kd;
        double val = (double)pivf + m;
        if (is_std) val = sqrt(fmax(fetch(b) / Kd - m * m, 0.0));
        float* out = p.out + i * (int64_t)p.out_stride + p.col_color;
        out[lane] = (float)(is8 ? val / 255.0 : val);
    }
}

// ------------------------------------------------------------------------------------------------
// k_color_warp: P = 64 (the reference's default patch size and every BASELINE configuration but the stress one). ONE WARP
// per nucleus, four nuclei per CTA, no block barrier after the table load. k_color spends more than a third of its
// instructions in per-warp fixed work (block scan, cross-warp reductions, barriers) that four warps repeat for ~7 pixels per
// thread; here one warp pays it once for ~29 pixels per lane:
//   * the window is taken in four 16-row sub-slabs (one TMA box {208 B, 16 rows} each = 32 mask words = one per lane);
//     sub-slabs without a mask bit are never fetched (a third of the DRAM traffic for the bench's nuclei);
//   * while a sub-slab is in flight its mask words are compacted into a list of byte addresses (warp scan);
//   * pivots come from the first masked pixel; sums are reduced with REDUX / one shuffle tree, the 17 columns are finished by
//     17 lanes exactly like k_color does.
constexpr int kCwWarps = 4, kCwRows = 16;
constexpr int kCwSlabBytes = kCwRows * kPanelBytes;              // 3328, a multiple of 128
constexpr int kCwWarpBytes = kCwSlabBytes + kCwRows * 64 * 2;    // + list: 5376
static_assert(kCwSlabBytes % 128 == 0 && kCwWarpBytes % 128 == 0, "TMA destinations must stay 128-byte aligned");

// ------------------------------------------------------------------------------------------------
// The pixel loop lives on the FMA pipe. On B200 FADD / FMUL / FFMA issue every cycle while
// ALU-pipe instructions (VIMNMX, ISETP, SEL, IADD3, LEA, I2FP) and IMAD issue every second cycle
// (scripts/f32x2_probe.cu, profiles/r2_f32x2_probe.txt); the round-1 integer form of this loop spent 50 ALU-pipe cycles per pixel.
//   * the optical-density table holds {OD(v), (float)v}: one 64-bit load per channel gives both, no conversions;
//   * max / min are FMNMX3, the sextant select is three predicated FFMAs (as in k_hue_batch), 1/d and 1/max are MUFU.RCP
//     of the float value plus 1e-30 (d = 0 -> numerator 0 -> 0, the reference's rule for grey pixels);
//   * the "integer" sums run in f32: a lane sees at most 128 pixels of a 64 x 64 window, so every partial sum stays below
//     2^24 and is exact; they are converted once and reduced with REDUX as before;
//   * the pivot is folded into the HED chain: max(y, 0) - pv = max(y - pv, -pv).
// A packed f32x2 version (two pixels per lane and step, FFMA2 / FADD2) was 5 % SLOWER than the scalar loop: FFMA2 issues every
// second cycle, so it saves issue slots but no FMA-pipe time, and the ALU pipe stayed the limit (profiles/README.md).
// Dynamic smem: per warp { slab[16 * 208] | list[1024] u16 } | lut2[256] float2.
__global__ void __launch_bounds__(32 * kCwWarps, 9)
k_color_warp(const ColorParams p, const __grid_constant__ CUtensorMap map /* box {208, 16 rows} */) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int P = 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2* lut2 = reinterpret_cast<float2*>(smem_raw + kCwWarps * kCwWarpBytes);
    __shared__ __align__(8) uint64_t s_bar[kCwWarps];
    for (int k = tid; k < 256; k += 32 * kCwWarps) lut2[k] = make_float2(g_od_lut[k], (float)k);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kCwWarps + warp;
    if (i >= p.n) return;

    uint8_t* slab = smem_raw + warp * kCwWarpBytes;
    uint16_t* list = reinterpret_cast<uint16_t*>(slab + kCwSlabBytes);
    uint64_t* bar = &s_bar[warp];
    const NucInfo inf = p.info[i];
    const int o = patch_byte_offset(inf.left);
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    const uint32_t* gm = p.bitmask + i * (int64_t)(P * 2);
    uint32_t w[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) w[s] = gm[s * 32 + lane];
    __syncwarp();
    float pvh = 0.f, pvs = 0.f, pv0 = 0.f, pv1 = 0.f, pv2 = 0.f;
    bool first = true;
    uint32_t phase = 0;
    int Ktot = 0;
    float fr1 = 0.f, fg1 = 0.f, fb1 = 0.f, fr2 = 0.f, fg2 = 0.f, fb2 = 0.f, fv1 = 0.f, fv2 = 0.f;   // exact (< 2^24 per lane)
    float s1[5] = {0, 0, 0, 0, 0}, s2[5] = {0, 0, 0, 0, 0};   // hed0, hed1, hed2, s, h (pivoted)
    const int abase = (lane >> 1) * kPanelBytes + (lane & 1) * 96;   // patch_addr(16, 0, lane / 2, 32 * (lane & 1))
    const uint32_t lut_a = smem_u32(lut2);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        uint32_t bits = w[s];
        if (!__any_sync(0xffffffffu, bits != 0u)) continue;   // warp-uniform: no masked pixel in these 16 rows
        const int row0 = s * kCwRows;
        if (lane == 0) {
            mbar_expect_tx(bar, (uint32_t)kCwSlabBytes);
            tma_load_window(slab, &map, inf.left, inf.top + row0, P, kCwRows, bar);
        }
        // ---- while the sub-slab is in flight: compact its 32 mask words into a list of byte addresses ----
        const int cnt = __popc(bits);
        int incl = cnt;
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o2);
            if (lane >= o2) incl += t;
        }
        const int K = __shfl_sync(0xffffffffu, incl, 31);
        int pos = incl - cnt;
        while (bits) {
            const int c = __ffs(bits) - 1;
            bits &= bits - 1;
            list[pos++] = (uint16_t)(abase + 3 * c);
        }
        Ktot += K;
        __syncwarp();   // list visible
        while (!mbar_try_wait(bar, phase)) {
        }
        phase ^= 1u;
        if (inf.nvc < P || inf.nvr < row0 + kCwRows) {   // rare: part of the window is never copied (NucInfo)
            for (int k = lane; k < kCwRows * P; k += 32) {
                const int r = k >> 6, c = k & 63;
                if (row0 + r >= inf.nvr || c >= inf.nvc) {
                    const int a = patch_addr(kCwRows, o, r, c);
                    slab[a] = 0; slab[a + 1] = 0; slab[a + 2] = 0;
                }
            }
            fence_proxy_async();   // generic-proxy stores before the next sub-slab's TMA (async proxy) write of the same bytes
            __syncwarp();
        }
        const uint8_t* pbase = slab + o;
        if (first) {   // pivots (any value of the right magnitude removes the one-pass cancellation): the first masked pixel
            const uint8_t* pp = pbase + list[0];
            const Px c = {pp[0], pp[1], pp[2]};
            const HsvHed t = convert(c, g_od_lut);
            pvh = t.h; pvs = t.s; pv0 = t.hed[0]; pv1 = t.hed[1]; pv2 = t.hed[2];
            first = false;
        }
        const float n0 = -pv0, n1 = -pv1, n2 = -pv2;
        for (int j = lane; j < K; j += 32) {
            const uint8_t* pp = pbase + list[j];
            const uint32_t r = pp[0], g = pp[1], b = pp[2];
            float oa, fr, ob, fg, oc, fb;   // optical density and value of each channel
            asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(oa), "=f"(fr) : "r"(lut_a + r * 8u));
            asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(ob), "=f"(fg) : "r"(lut_a + g * 8u));
            asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(oc), "=f"(fb) : "r"(lut_a + b * 8u));
            fr1 += fr; fg1 += fg; fb1 += fb;
            fr2 = fmaf(fr, fr, fr2); fg2 = fmaf(fg, fg, fg2); fb2 = fmaf(fb, fb, fb2);
            // HED minus pivot: max(y, 0) - pv = max(y - pv, -pv)
            float d0 = fmaxf(fmaf(oc, HED_M20, fmaf(ob, HED_M10, fmaf(oa, HED_M00, n0))), n0);
            float d1 = fmaxf(fmaf(oc, HED_M21, fmaf(ob, HED_M11, fmaf(oa, HED_M01, n1))), n1);
            float d2 = fmaxf(fmaf(oc, HED_M22, fmaf(ob, HED_M12, fmaf(oa, HED_M02, n2))), n2);
            s1[0] += d0; s2[0] = fmaf(d0, d0, s2[0]);
            s1[1] += d1; s2[1] = fmaf(d1, d1, s2[1]);
            s1[2] += d2; s2[2] = fmaf(d2, d2, s2[2]);
            // HSV
            const float mx = fmaxf(fr, fmaxf(fg, fb)), mn = fminf(fr, fminf(fg, fb));
            fv1 += mx; fv2 = fmaf(mx, mx, fv2);
            const float d = mx - mn;
            const float rd = rcp_approx(d + 1e-30f), rm = rcp_approx(mx + 1e-30f);
            float t;   // hue in sextants, wrapped into [0, 6): max = r -> (g-b)/d (+6 if negative) ; g -> 2 + (b-r)/d ; b -> 4 + (r-g)/d
            asm("{\n"
                ".reg .pred pr, pg, pn;\n"
                ".reg .f32 gb, br, rg;\n"
                "sub.f32 gb, %2, %3;\n"
                "sub.f32 br, %3, %1;\n"
                "sub.f32 rg, %1, %2;\n"
                "setp.eq.f32 pr, %4, %1;\n"
                "setp.eq.f32 pg, %4, %2;\n"
                "setp.lt.and.f32 pn, gb, 0f00000000, pr;\n"
                "fma.rn.f32 %0, rg, %5, 0f40800000;\n"        // 4
                "@pg fma.rn.f32 %0, br, %5, 0f40000000;\n"    // 2
                "@pr mul.f32 %0, gb, %5;\n"
                "@pn add.f32 %0, %0, 0f40C00000;\n"           // + 6
                "}\n"
                : "=&f"(t)
                : "f"(fr), "f"(fg), "f"(fb), "f"(mx), "f"(rd));
            const float dh = fmaf(t, 60.0f, -pvh), ds = fmaf(d, rm, -pvs);
            s1[3] += ds; s2[3] = fmaf(ds, ds, s2[3]);
            s1[4] += dh; s2[4] = fmaf(dh, dh, s2[4]);
        }
        __syncwarp();   // slab and list are reused by the next sub-slab
    }
    // ---- sums: every lane ends up with every total; lane 0 parks them in the (dead) list for the column lanes ----
    uint32_t* fi = reinterpret_cast<uint32_t*>(list);
    float* ff = reinterpret_cast<float*>(list) + 8;
    {
        const float vf[8] = {fr1, fg1, fb1, fr2, fg2, fb2, fv1, fv2};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t t = __reduce_add_sync(0xffffffffu, (uint32_t)__float2int_rn(vf[q]));
            if (lane == 0) fi[q] = t;
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const float a = warp_sum(s1[q]), b = warp_sum(s2[q]);
            if (lane == 0) { ff[q] = a; ff[5 + q] = b; }
        }
    }
    __syncwarp();
    // ---- 17 columns, one lane each (out[6] = mean_h belongs to k_hue_finalize); same f64 expressions as k_color ----
    if (lane < 18 && lane != 6) {
        static const int ia[18] = {1, 2, 3, 1, 2, 3, 0, 12, 7, 13, 12, 7, 9, 10, 11, 9, 10, 11};
        static const int ib[18] = {0, 0, 0, 4, 5, 6, 0, 0, 0, 18, 17, 8, 0, 0, 0, 14, 15, 16};
        auto fetch = [&](int q) -> double { return q <= 8 ? (double)fi[q - 1] : (double)ff[q - 9]; };   // q in 1..18
        const int a = ia[lane], b = ib[lane];
        const bool is_std = b != 0;
        const bool is8 = a <= 8;                                   // u8-valued channel: scale by 1/255
        const float pivf = lane == 7 ? pvs : (lane == 12 ? pv0 : (lane == 13 ? pv1 : (lane == 14 ? pv2 : 0.f)));
        const double Kd = (double)Ktot;
        const double m = fetch(a) / Kd;
        double val = (double)pivf + m;
        if (is_std) val = sqrt(fmax(fetch(b) / Kd - m * m, 0.0));
        float* out = p.out + i * (int64_t)p.out_stride + p.col_color;
        out[lane] = (float)(is8 ? val / 255.0 : val);
    }
}

// ------------------------------------------------------------------------------------------------
// k_hue_batch: grid = (n_batches, slabs); one CTA per (chunk, row slab), a TMA producer warp, 8 consumer warps that own one
// pixel quad of the union of the chunk's masks each (mean_h only ever reads C, S under some mask of the batch, so quads outside
// the union are never evaluated: 44 % of a 64 x 64 window for the bench's nuclei). Round 2 cut the per-patch fixed work
// (ncu: 44 of the 139 instructions per patch and warp were ring / barrier / address work, the masked fold was 12.5 % of the
// kernel, producer spinning 4.4 %):
//   * a ring stage holds TWO patches: one mbarrier wait, one arrive and one loop step per pair of patches;
//   * the per-patch constants of the quad loads (word offset, funnel shift, "window partly uncopied") are packed into one
//     u32 when the chunk's NucInfo is staged; every lane of an active warp computes (lanes without a quad re-read quad 0
//     and are dropped at the end), so the loop body has no ownership branches;
//   * the fold under each nucleus' mask uses ROW PREFIX SUMS of the (S, C) image in fixed point (exact integer differences,
//     no cancellation, any order gives the same bits): a mask word is walked run by run -- two 64-bit loads per run of
//     set bits instead of two loads and two adds per set bit. `fix_shift` keeps batch * P * 2^shift below 2^31;
//   * the producer lane sleeps between polls of the `empty` barrier.
// Dynamic smem: ring[kHue2Stages][2][stage_bytes] | pre[R][P + 2] int2 {S, C} (entry k = sum of the first k pixels).
constexpr int kHue2Stages = 4;
template <int NCW>
__global__ void __launch_bounds__(32 * NCW + 32, 5)
k_hue_batch(const ColorParams p, const __grid_constant__ CUtensorMap map, const int R, const int fix_shift) {
    constexpr int kConsumers = 32 * NCW, kThreads = kConsumers + 32;
    static_assert(kConsumers == kHueMaxQuads, "one quad per consumer thread");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = p.P, wpr = mask_wpr(P), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int stage_bytes = window_smem_bytes(P, R);
    const uint32_t stage_tx = (uint32_t)(patch_panels(P) * kPanelBytes * R);
    const int64_t b0 = (int64_t)blockIdx.x * p.batch_size;
    const int nb = (int)min((int64_t)p.batch_size, p.n - b0);
    const int slab = blockIdx.y, row0 = slab * R;
    uint8_t* ring = smem_raw;
    const int ppitch = P + 2;
    int2* pre = reinterpret_cast<int2*>(smem_raw + (size_t)kHue2Stages * 2 * stage_bytes);
    __shared__ __align__(8) uint64_t full[kHue2Stages], empty[kHue2Stages];
    __shared__ NucInfo s_info[kHueChunk];
    __shared__ __align__(8) uint32_t s_pinfo[kHueChunk + 2];
    __shared__ uint32_t s_union[64];
    __shared__ uint16_t s_qlist[kHueMaxQuads];
    __shared__ int s_nact;
    __shared__ float s_rcp[256];   // (pi/3) / d
    for (int k = tid; k < 256; k += kThreads) s_rcp[k] = k ? __fdiv_rn(1.0471975511965976f, (float)k) : 0.f;
    const int qpr = (P + 3) >> 2, nquads = R * qpr;   // the last quad of a row may hang over (P % 4 != 0): never stored
    const int nrows_u = min(R, P - row0), words_u = nrows_u * wpr;
    if (tid < 64) s_union[tid] = 0u;
    for (int k = tid; k < R * ppitch; k += kThreads) pre[k] = make_int2(0, 0);
    __syncthreads();
    {   // union of the chunk's masks over this slab (thread t walks words t, t + T, ... of the [nb][words_u] block)
        // four independent loads per round: one load per round left every CTA waiting ~11 L2 latencies before its first
        // TMA (ncu: 10 % of the kernel's stall samples sat in this loop for 3 % of its instructions)
        int j = tid / words_u, w = tid - j * words_u;
        const int dj = kThreads / words_u, dw = kThreads - dj * words_u;
        while (j < nb) {
            uint32_t v[4];
            int ww[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ww[u] = w;
                v[u] = (j < nb) ? p.bitmask[((b0 + j) * (int64_t)P + row0) * wpr + w] : 0u;
                j += dj; w += dw;
                if (w >= words_u) { w -= words_u; ++j; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (v[u]) atomicOr(&s_union[ww[u]], v[u]);
        }
    }
    __syncthreads();
    if (warp == 0) {
        int cnt = 0;
        for (int base = 0; base < nquads; base += 32) {
            const int q = base + lane, r = q / qpr, c = (q - r * qpr) * 4;
            const bool act = q < nquads && r < nrows_u && ((s_union[r * wpr + (c >> 5)] >> (c & 31)) & 0xFu) != 0u;
            const uint32_t b = __ballot_sync(0xffffffffu, act);
            if (act) s_qlist[cnt + __popc(b & ((1u << lane) - 1u))] = (uint16_t)q;
            cnt += __popc(b);
        }
        if (lane == 0) s_nact = cnt;
    }
    __syncthreads();
    const int nact = s_nact;
    if (nact == 0) {   // no nucleus of the chunk has a masked pixel in this slab
        for (int i = tid; i < nb; i += kThreads) {
            float* hp = p.hue_partial + ((b0 + i) * (int64_t)p.slabs + slab) * 2;
            hp[0] = 0.f;
            hp[1] = 0.f;
        }
        return;
    }
    const int nwarps_act = min(NCW, (nact + 31) / 32);   // consumer warps that own at least one quad
    if (tid == 0) {
        for (int s = 0; s < kHue2Stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nwarps_act); }
        mbar_fence_init();
    }
    const bool owner = (tid < kConsumers) && (tid < nact);
    const int q = (int)s_qlist[owner ? tid : 0];
    const int rr = q / qpr, c0 = (q - rr * qpr) * 4;
    uint32_t soff = (uint32_t)((c0 >> 6) * panel_stride(R) + rr * kPanelBytes + (c0 & 63) * 3);   // + word offset per patch
    const bool active = warp < nwarps_act;   // warp-uniform
    float C[4] = {0.f, 0.f, 0.f, 0.f}, S[4] = {0.f, 0.f, 0.f, 0.f};

    uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty), pinfo_a = smem_u32(s_pinfo), ring_a = smem_u32(ring),
             rcp_a = smem_u32(s_rcp);
    asm volatile("" : "+r"(full_a), "+r"(empty_a), "+r"(pinfo_a), "+r"(ring_a), "+r"(rcp_a), "+r"(soff));

    // four pixels of one patch: words w0..w2 hold r0 g0 b0 r1 | g1 b1 r2 g2 | b2 r3 g3 b3.
    // B200 issues FADD / FMUL / FFMA every cycle but ALU-pipe instructions (PRMT, VIMNMX, ISETP, SEL, IADD3, LEA, I2FP) and
    // IMAD only every second cycle (scripts/f32x2_probe.cu), and the integer form of this loop was ALU-pipe bound (ncu: 68 %
    // ALU, 23 % FMA). So the pixel is unpacked straight into FLOATS (PRMT against 0x4B000000 gives 2^23 + byte, exact), the
    // channel differences are FADDs (the bias cancels), and the sextant select is three predicated FFMA/FMULs instead of
    // integer subtractions, two selects and a conversion. The table index d = max - min is the difference of the bit patterns.
    auto hue_angle = [&](float fr, float fg, float fb) -> float {
        const float mx = fmaxf(fr, fmaxf(fg, fb)), mn = fminf(fr, fminf(fg, fb));
        const uint32_t d = (uint32_t)(__float_as_int(mx) - __float_as_int(mn));
        const float rd = lds_f32(rcp_a + d * 4u);   // (pi/3) / d, 0 for d = 0 (then every difference is 0 too)
        // h = 60 t degrees = t * pi/3 radians (no wrap needed under sin/cos): max = r -> (g-b)/d ; g -> 2 + (b-r)/d ; b -> 4 + (r-g)/d
        float ang;
        asm("{\n"
            ".reg .pred pr, pg;\n"
            ".reg .f32 gb, br, rg;\n"
            "sub.f32 gb, %2, %3;\n"
            "sub.f32 br, %3, %1;\n"
            "sub.f32 rg, %1, %2;\n"
            "setp.eq.f32 pr, %4, %1;\n"
            "setp.eq.f32 pg, %4, %2;\n"
            "fma.rn.f32 %0, rg, %5, 0f40860A92;\n"        // 4 pi / 3
            "@pg fma.rn.f32 %0, br, %5, 0f40060A92;\n"    // 2 pi / 3
            "@pr mul.f32 %0, gb, %5;\n"
            "}\n"
            : "=&f"(ang)
            : "f"(fr), "f"(fg), "f"(fb), "f"(mx), "f"(rd));
        return ang;
    };
    auto accumulate = [&](uint32_t w0, uint32_t w1, uint32_t w2) {
        constexpr uint32_t kBias = 0x4B000000u;   // 2^23: PRMT picks one byte of w and the three upper bytes of kBias
#define NFX_BF(w, k) __uint_as_float(__byte_perm((w), kBias, 0x7540 + (k)))
        const float a0 = hue_angle(NFX_BF(w0, 0), NFX_BF(w0, 1), NFX_BF(w0, 2));
        const float a1 = hue_angle(NFX_BF(w0, 3), NFX_BF(w1, 0), NFX_BF(w1, 1));
        const float a2 = hue_angle(NFX_BF(w1, 2), NFX_BF(w1, 3), NFX_BF(w2, 0));
        const float a3 = hue_angle(NFX_BF(w2, 1), NFX_BF(w2, 2), NFX_BF(w2, 3));
#undef NFX_BF
        C[0] += __cosf(a0); S[0] += __sinf(a0);
        C[1] += __cosf(a1); S[1] += __sinf(a1);
        C[2] += __cosf(a2); S[2] += __sinf(a2);
        C[3] += __cosf(a3); S[3] += __sinf(a3);
    };
    // bytes of the quad that the reference never copies (NucInfo) read as zero
    auto mask_uncopied = [&](const NucInfo& inf, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
        const bool rowdead = (row0 + rr) >= inf.nvr;
        const int nlive = rowdead ? 0 : min(max(inf.nvc - c0, 0), 4);   // live pixels of the quad
        const int nbytes = 3 * nlive;                                    // pixel k occupies bytes 3k..3k+2 of the 12-byte quad
        w0 = nbytes >= 4 ? w0 : (nbytes > 0 ? (w0 & ((1u << (8 * nbytes)) - 1u)) : 0u);
        w1 = nbytes >= 8 ? w1 : (nbytes > 4 ? (w1 & ((1u << (8 * (nbytes - 4))) - 1u)) : 0u);
        w2 = nbytes >= 12 ? w2 : (nbytes > 8 ? (w2 & ((1u << (8 * (nbytes - 8))) - 1u)) : 0u);
    };

    uint32_t gq = 0;   // pair counter over the whole batch (ring position)
    for (int base = 0; base < nb; base += kHueChunk) {
        const int cnt = min(kHueChunk, nb - base);
        __syncthreads();   // previous chunk fully consumed before the staged records are overwritten
        for (int k = tid; k < cnt + 1; k += kThreads) {
            uint32_t pi = 0u;
            if (k < cnt) {
                const NucInfo inf = p.info[b0 + base + k];
                s_info[k] = inf;
                const uint32_t ob = (uint32_t)patch_byte_offset(inf.left);
                pi = (ob & ~3u) | ((ob & 3u) << 11) | ((inf.nvc < P || inf.nvr < P) ? 0x10000u : 0u);
            }
            s_pinfo[k] = pi;
        }
        __syncthreads();
        const int npairs = (cnt + 1) >> 1;
        if (warp == NCW) {
            // ---- TMA producer warp (one elected lane) ----
            if (lane == 0) {
                for (int jp = 0; jp < npairs; ++jp) {
                    const uint32_t g = gq + (uint32_t)jp, s = g % kHue2Stages, ph = (g / kHue2Stages) & 1u;
                    if (g >= (uint32_t)kHue2Stages)
                        while (!mbar_try_wait(&empty[s], ph ^ 1u)) __nanosleep(40);
                    const int two = (2 * jp + 1 < cnt);
                    mbar_expect_tx(&full[s], stage_tx * (two ? 2u : 1u));
                    uint8_t* dst = ring + (size_t)s * 2 * stage_bytes;
                    const NucInfo i0 = s_info[2 * jp];
                    tma_load_window(dst, &map, i0.left, i0.top + row0, P, R, &full[s]);
                    if (two) {
                        const NucInfo i1 = s_info[2 * jp + 1];
                        tma_load_window(dst + stage_bytes, &map, i1.left, i1.top + row0, P, R, &full[s]);
                    }
                }
            }
        } else if (active) {
            uint32_t g = gq, pa = pinfo_a;
            for (int jp = 0; jp < npairs; ++jp, ++g, pa += 8u) {
                const uint32_t s = g % kHue2Stages, ph = (g / kHue2Stages) & 1u;
                while (!mbar_try_wait_a(full_a + s * 8u, ph)) {
                }
                uint32_t pi0, pi1;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(pi0), "=r"(pi1) : "r"(pa) : "memory");
                // soff is a multiple of 4, so the word alignment and the funnel shift depend on the patch only
                const uint32_t sb = ring_a + s * (uint32_t)(2 * stage_bytes) + soff;
                const uint32_t wa = sb + (pi0 & 0xffu), wb = sb + (uint32_t)stage_bytes + (pi1 & 0xffu);
                const uint32_t a0 = lds_u32(wa), a1 = lds_u32(wa + 4), a2 = lds_u32(wa + 8), a3 = lds_u32(wa + 12);
                const uint32_t b0w = lds_u32(wb), b1 = lds_u32(wb + 4), b2 = lds_u32(wb + 8), b3 = lds_u32(wb + 12);
                const uint32_t sha = pi0 >> 8, shb = pi1 >> 8;   // the funnel shift uses the low five bits
                uint32_t u0 = __funnelshift_r(a0, a1, sha), u1 = __funnelshift_r(a1, a2, sha), u2 = __funnelshift_r(a2, a3, sha);
                uint32_t v0 = __funnelshift_r(b0w, b1, shb), v1 = __funnelshift_r(b1, b2, shb), v2 = __funnelshift_r(b2, b3, shb);
                __syncwarp();
                if (elect_one()) mbar_arrive_a(empty_a + s * 8u);
                const bool two = 2 * jp + 1 < cnt;   // warp-uniform: the last pair of an odd chunk holds one patch
                if ((pi0 | pi1) & 0x10000u) {   // rare: a window partly never copied (NucInfo); own copy of the arithmetic
                    if (pi0 & 0x10000u) mask_uncopied(s_info[2 * jp], u0, u1, u2);
                    if (pi1 & 0x10000u) mask_uncopied(s_info[2 * jp + 1], v0, v1, v2);
                    accumulate(u0, u1, u2);
                    if (two) accumulate(v0, v1, v2);
                } else {
                    accumulate(u0, u1, u2);
                    if (two) accumulate(v0, v1, v2);
                }
            }
        }
        gq += (uint32_t)npairs;
    }
    __syncthreads();
    // ---- the slab's (S, C) image in fixed point, then inclusive row prefix sums: pre[r][k] = sum of pixels 0..k-1 ----
    if (owner) {
        const float scale = (float)(1u << fix_shift);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (c0 + k < P) pre[rr * ppitch + 1 + c0 + k] = make_int2(__float2int_rn(S[k] * scale), __float2int_rn(C[k] * scale));
    }
    __syncthreads();
    constexpr int NWARP = kThreads / 32;
    const int nrows = min(R, P - row0), words = nrows * wpr;
    for (int r = warp; r < nrows; r += NWARP) {
        int2* row = pre + r * ppitch;
        int cx = 0, cy = 0;
        for (int c = 0; c < P; c += 64) {
            const int e0 = 1 + c + 2 * lane;
            int2 a = (e0 <= P) ? row[e0] : make_int2(0, 0), b = (e0 + 1 <= P) ? row[e0 + 1] : make_int2(0, 0);
            const int lx = a.x + b.x, ly = a.y + b.y;
            int ix = lx, iy = ly;
#pragma unroll
            for (int o2 = 1; o2 < 32; o2 <<= 1) {
                const int tx = __shfl_up_sync(0xffffffffu, ix, o2), ty = __shfl_up_sync(0xffffffffu, iy, o2);
                if (lane >= o2) { ix += tx; iy += ty; }
            }
            a.x += cx + ix - lx; a.y += cy + iy - ly;
            b.x += a.x; b.y += a.y;
            if (e0 <= P) row[e0] = a;
            if (e0 + 1 <= P) row[e0 + 1] = b;
            cx += __shfl_sync(0xffffffffu, ix, 31);
            cy += __shfl_sync(0xffffffffu, iy, 31);
        }
    }
    __syncthreads();
    // ---- masked sums under each nucleus' mask, four nuclei per warp pass (their mask words are fetched together) ----
    constexpr int UN = 4;
    const float inv_scale = 1.0f / (float)(1u << fix_shift);
    for (int i0 = warp * UN; i0 < nb; i0 += NWARP * UN) {
        float ss[UN], sc[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) { ss[u] = 0.f; sc[u] = 0.f; }
        for (int w = lane; w < words; w += 32) {
            uint32_t bits[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u)
                bits[u] = (i0 + u < nb) ? p.bitmask[((b0 + i0 + u) * (int64_t)P + row0) * wpr + w] : 0u;
            const int r = w / wpr;
            const int2* rowp = pre + r * ppitch + (w - r * wpr) * 32;
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                uint32_t b = bits[u];
                while (b) {   // one iteration per run of set bits
                    const int s = __ffs(b) - 1;
                    const uint32_t nt = ~(b >> s);
                    const int e = s + (nt ? __ffs(nt) - 1 : 32);   // run = bits [s, e)
                    const int2 hi = rowp[e], lo = rowp[s];
                    ss[u] += (float)(hi.x - lo.x);
                    sc[u] += (float)(hi.y - lo.y);
                    b = (e >= 32) ? 0u : (b & (0xffffffffu << e));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const float s1 = warp_sum(ss[u]), c1 = warp_sum(sc[u]);
            if (lane == 0 && i0 + u < nb) {
                float* hp = p.hue_partial + ((b0 + i0 + u) * (int64_t)p.slabs + slab) * 2;
                hp[0] = s1 * inv_scale;
                hp[1] = c1 * inv_scale;
            }
        }
    }
}

__global__ void k_hue_finalize(const ColorParams p) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    float* out = p.out + i * (int64_t)p.out_stride + p.col_color;
    const float* hp = p.hue_partial + i * (int64_t)p.slabs * 2;
    float ss = 0.f, sc = 0.f;
    for (int s = 0; s < p.slabs; ++s) { ss += hp[2 * s]; sc += hp[2 * s + 1]; }
    // color.rs:154  (atan2(sin, cos).rad2deg + 360) fmod 360 ; empty mask -> 0/0 -> NaN
    float deg = atan2f(ss, sc) * 57.29577951308232f;
    deg = fmodf(deg + 360.0f, 360.0f);
    if (out[0] != out[0]) deg = CUDART_NAN_F;   // mean_r is NaN iff the mask is empty
    out[6] = deg;
}

}  // namespace

// rows per (chunk, slab) CTA of k_hue_batch: at most kHueMaxQuads pixel quads (one per consumer thread)
int hue_slab_rows(int P) { return max(1, min(P, kHueMaxQuads / ((P + 3) / 4))); }
int color_slab_rows(int P) { return P < 64 ? P : 64; }
int color_smem_bytes(int P) {
    const int cs = color_slab_rows(P);
    const int list_bytes = cs * P * 2 > 10 * kColorThreads * 4 ? cs * P * 2 : 10 * kColorThreads * 4;   // list / float scratch
    return window_smem_bytes(P, cs) + P * mask_wpr(P) * 4 + 256 * 4 + list_bytes;
}

// One table per device, shared by every context on it (the reference gives each rayon thread its own context and
// several of them share a GPU, utils.rs:215-221): filled once under a lock and COMPLETED before the flag is set, because
// the streams of other contexts are not ordered after `s`.
static std::mutex g_lut_mu;
static bool g_lut_ready[64] = {};
static cudaError_t ensure_lut(cudaStream_t s) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(g_lut_mu);
    if (dev < 64 && g_lut_ready[dev]) return cudaSuccess;
    k_init_od_lut<<<1, 256, 0, s>>>();
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess && dev < 64) g_lut_ready[dev] = true;
    return e;
}

cudaError_t launch_color(const ColorParams& p, const CUtensorMap* map, cudaStream_t s) {
    if (p.n <= 0) return cudaSuccess;
    cudaError_t e = ensure_lut(s);
    if (e != cudaSuccess) return e;
    const int smem = color_smem_bytes(p.P);
    if (smem > 32 * 1024) {   // static shared memory counts towards the 48 KB default limit too
        e = cudaFuncSetAttribute(k_color, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    k_color<<<(unsigned)p.n, kColorThreads, smem, s>>>(p, *map);
    return cudaGetLastError();
}

// P = 64 only; `map_slab` is the {208 B, 16 rows} slab map (hue_slab_rows(64) == 16) that k_hue_batch uses too.
cudaError_t launch_color_warp(const ColorParams& p, const CUtensorMap* map_slab, cudaStream_t s) {
    if (p.n <= 0) return cudaSuccess;
    if (p.P != 64 || hue_slab_rows(64) != kCwRows) return cudaErrorInvalidValue;
    cudaError_t e = ensure_lut(s);
    if (e != cudaSuccess) return e;
    const int smem = kCwWarps * kCwWarpBytes + 256 * 8;
    k_color_warp<<<(unsigned)((p.n + kCwWarps - 1) / kCwWarps), 32 * kCwWarps, smem, s>>>(p, *map_slab);
    return cudaGetLastError();
}

// Largest shift with batch * P * 2^shift < 2^31 (the row prefix sums of k_hue_batch are int32), at most 18.
static int hue_fix_shift(int64_t nb, int P) {
    int sh = 18;
    while (sh > 0 && (double)nb * P * (double)(1u << sh) >= 2147483648.0) --sh;
    return sh;
}

cudaError_t launch_hue_batch(const ColorParams& p, const CUtensorMap* map_slab, int R, cudaStream_t s) {
    if (p.n <= 0) return cudaSuccess;
    const int64_t nbatch = (p.n + p.batch_size - 1) / p.batch_size;
    dim3 grid((unsigned)nbatch, (unsigned)p.slabs);
    const int64_t nbmax = p.n < p.batch_size ? p.n : p.batch_size;
    if ((double)nbmax * p.P >= 2147483648.0) return cudaErrorInvalidValue;   // a chunk of > 2^31 / P nuclei
    const int smem = kHue2Stages * 2 * window_smem_bytes(p.P, R) + R * (p.P + 2) * (int)sizeof(int2);
    if (smem > 32 * 1024) {   // static shared memory counts towards the 48 KB default limit too
        cudaError_t e = cudaFuncSetAttribute(k_hue_batch<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    k_hue_batch<8><<<grid, 32 * 8 + 32, smem, s>>>(p, *map_slab, R, hue_fix_shift(nbmax, p.P));
    return cudaGetLastError();
}

cudaError_t launch_hue_finalize(const ColorParams& p, cudaStream_t s) {
    if (p.n <= 0) return cudaSuccess;
    k_hue_finalize<<<(unsigned)((p.n + 255) / 256), 256, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace nfx
