// ext.cu -- EXTENSION outputs: quantities BASELINE.json's north_star lists but the reference does NOT compute
// ("build them as clearly-separated extension outputs, never mixed into the drop-in column schema").
//   NFX_EXT_COLOR_MOMENTS  masked skewness / excess kurtosis of r, g, b, grey, s, v, haematoxylin, eosin, dab   (SPEC.md C1)
//   NFX_EXT_MASK_MOMENTS   raw / central / Hu moments of the raster mask (cv2.moments + cv2.HuMoments)           (SPEC.md C2)
//   NFX_EXT_CONTOUR        boundary length of the mask: crack count and Pratt's bit-quad perimeter                (SPEC.md C3)
//   NFX_EXT_GLCM_D2        BASELINE config 3's literal GLCM: 32 levels, distances 1 and 2, 4 angles, 14 Haralick (SPEC.md C4)
// They have no reference counterpart: oracle/nfx_oracle.py (ext_*) is their specification, pinned against scipy / cv2.
// Separate entry points (nfx_compute_ext / nfx_download_ext), separate output matrix, separate column names.
//
// These are not on the reference's hot path; the kernels are written for clarity: one CTA per nucleus, mask bits from the
// bitmask k_geom wrote, window pixels straight from the resident slide (rows of a window are contiguous 3P-byte runs).
#include <math_constants.h>

#include "nfx_kernels.h"

namespace nfx {

namespace {

constexpr int kExtThreads = 128;
constexpr int kExtGlcmThreads = 256;

// inv([[0.65,0.70,0.29],[0.07,0.99,0.11],[0.27,0.57,0.78]]) in f32 (oracle HED_FROM_RGB, SPEC.md B4)
__device__ __constant__ float c_hed[3][3] = {{1.87798273563385f, -1.0076786279678345f, -0.5561158061027527f},
                                             {-0.06590805947780609f, 1.134730339050293f, -0.135521799325943f},
                                             {-0.6019073724746704f, -0.48041418194770813f, 1.5735880136489868f}};

struct Rgb {
    uint32_t r, g, b;
};
// pixel (r, c) of nucleus window `inf` (utils.rs:159-192: zero outside the slide and beyond what the reference copies)
__device__ __forceinline__ Rgb window_pixel(const ExtParams& p, const NucInfo& inf, int r, int c) {
    Rgb o = {0u, 0u, 0u};
    const int64_t y = (int64_t)inf.top + r, x = (int64_t)inf.left + c;
    if (r < inf.nvr && c < inf.nvc && y >= 0 && y < p.th && x >= 0 && x < p.tw) {
        const uint8_t* px = p.tile + y * p.tpitch + 3 * x;
        o.r = px[0]; o.g = px[1]; o.b = px[2];
    }
    return o;
}

// the nine channels of SPEC.md C1 for one pixel, f32 as the reference-side conversions produce them (SPEC.md B3, B4, A12)
__device__ __forceinline__ void channels(const Rgb& px, float* ch) {
    const float r = __fdiv_rn((float)px.r, 255.0f), g = __fdiv_rn((float)px.g, 255.0f), b = __fdiv_rn((float)px.b, 255.0f);
    ch[0] = r; ch[1] = g; ch[2] = b;
    ch[3] = __fdiv_rn(__fadd_rn(__fadd_rn(r, g), b), 3.0f);
    const float mx = fmaxf(r, fmaxf(g, b)), mn = fminf(r, fminf(g, b));
    ch[4] = mx > 0.f ? __fdiv_rn(mx - mn, mx) : 0.f;
    ch[5] = mx;
    float od[3];
    const float v[3] = {r, g, b};
#pragma unroll
    for (int k = 0; k < 3; ++k) od[k] = __fdiv_rn(logf(fmaxf(v[k], 1e-6f)), -13.815510749816895f);
#pragma unroll
    for (int k = 0; k < 3; ++k) ch[6 + k] = fmaxf(0.f, od[0] * c_hed[0][k] + od[1] * c_hed[1][k] + od[2] * c_hed[2][k]);
}

// ------------------------------------------------------------------------------------------------
// colour moments + mask moments + contour: one CTA per nucleus. Dynamic smem: rows[P * wpr] u32.
__global__ void __launch_bounds__(kExtThreads) k_ext_stats(const ExtParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* rows = reinterpret_cast<uint32_t*>(smem_raw);
    const int P = p.P, wpr = mask_wpr(P), tid = threadIdx.x;
    const int64_t i = blockIdx.x;
    __shared__ double s_red[40 * (kExtThreads / 32)];
    __shared__ unsigned long long s_int[12 * (kExtThreads / 32)];
    __shared__ float s_piv[9];
    __shared__ int s_first;
    const NucInfo inf = p.info[i];
    const uint32_t* gm = p.bitmask + i * (int64_t)P * wpr;
    for (int k = tid; k < P * wpr; k += kExtThreads) rows[k] = gm[k];
    if (tid == 0) s_first = P * P;
    __syncthreads();
    auto mbit = [&](int r, int c) -> int {
        return (r >= 0 && r < P && c >= 0 && c < P) ? (int)((rows[r * wpr + (c >> 5)] >> (c & 31)) & 1u) : 0;
    };
    float* out = p.out + i * (int64_t)p.out_stride;

    // ---- mask moments (exact integer sums) and the crack count, word by word ----
    unsigned long long m[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // m00 m10 m01 m20 m11 m02 m30 m21 m12 m03 | cracks
    int first = P * P;
    for (int k = tid; k < P * wpr; k += kExtThreads) {
        const int r = k / wpr, w = k - r * wpr;
        uint32_t bits = rows[k];
        if (!bits) continue;
        first = min(first, r * P + 32 * w + __ffs(bits) - 1);
        const uint32_t lo = w > 0 ? rows[k - 1] >> 31 : 0u, hi = w + 1 < wpr ? rows[k + 1] << 31 : 0u;
        const uint32_t left = (bits << 1) | lo, right = (bits >> 1) | hi;
        const uint32_t up = r > 0 ? rows[k - wpr] : 0u, down = r + 1 < P ? rows[k + wpr] : 0u;
        m[10] += __popc(bits & ~left) + __popc(bits & ~right) + __popc(bits & ~up) + __popc(bits & ~down);
        const unsigned long long y = (unsigned long long)r;
        while (bits) {
            const unsigned long long x = (unsigned long long)(32 * w + __ffs(bits) - 1);
            bits &= bits - 1;
            m[0] += 1; m[1] += x; m[2] += y; m[3] += x * x; m[4] += x * y; m[5] += y * y;
            m[6] += x * x * x; m[7] += x * x * y; m[8] += x * y * y; m[9] += y * y * y;
        }
    }
    // ---- bit quads over the zero-padded mask: windows with top-left corner (r, c), r, c in [-1, P-1] ----
    unsigned long long q1 = 0, q2 = 0, q3 = 0, qd = 0;
    for (int k = tid; k < (P + 1) * (P + 1); k += kExtThreads) {
        const int r = k / (P + 1) - 1, c = k - (r + 1) * (P + 1) - 1;
        const int a = mbit(r, c), b = mbit(r, c + 1), cc = mbit(r + 1, c), d = mbit(r + 1, c + 1), n = a + b + cc + d;
        q1 += (n == 1);
        q3 += (n == 3);
        const bool diag = (n == 2) && (a == d);
        qd += diag;
        q2 += (n == 2) && !diag;
    }
    {
        unsigned long long v[15];
#pragma unroll
        for (int k = 0; k < 11; ++k) v[k] = m[k];
        v[11] = q1; v[12] = q2; v[13] = q3; v[14] = qd;
        const int lane = tid & 31, warp = tid >> 5;
        first = warp_min(first);
        if (lane == 0) atomicMin(&s_first, first);
        // 15 integer sums: shuffle tree per warp, then a serial fold (exact in any order)
        __shared__ unsigned long long s_all[15 * (kExtThreads / 32)];
#pragma unroll
        for (int k = 0; k < 15; ++k) {
            unsigned long long t = v[k];
            for (int o2 = 16; o2 > 0; o2 >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o2);
            if (lane == 0) s_all[warp * 15 + k] = t;
        }
        __syncthreads();
        if (tid == 0) {
            double s[15];
            for (int k = 0; k < 15; ++k) {
                unsigned long long t = 0;
                for (int w = 0; w < kExtThreads / 32; ++w) t += s_all[w * 15 + k];
                s[k] = (double)t;
            }
            s_int[0] = (unsigned long long)s[0];
            if (p.col_mask >= 0) {
                float* o = out + p.col_mask;
                for (int k = 0; k < 10; ++k) o[k] = (float)s[k];
                const double K = s[0];
                if (K > 0.0) {
                    const double cx = s[1] / K, cy = s[2] / K;
                    // central moments from the raw ones (cv2.moments)
                    const double mu20 = s[3] - cx * s[1], mu11 = s[4] - cx * s[2], mu02 = s[5] - cy * s[2];
                    const double mu30 = s[6] - cx * (3.0 * mu20 + cx * s[1]);
                    const double mu21 = s[7] - cx * (2.0 * mu11 + cx * s[2]) - cy * mu20;
                    const double mu12 = s[8] - cy * (2.0 * mu11 + cy * s[1]) - cx * mu02;
                    const double mu03 = s[9] - cy * (3.0 * mu02 + cy * s[2]);
                    o[10] = (float)mu20; o[11] = (float)mu11; o[12] = (float)mu02;
                    o[13] = (float)mu30; o[14] = (float)mu21; o[15] = (float)mu12; o[16] = (float)mu03;
                    const double i2 = 1.0 / (K * K), i3 = i2 / sqrt(K);
                    const double n20 = mu20 * i2, n11 = mu11 * i2, n02 = mu02 * i2, n30 = mu30 * i3, n21 = mu21 * i3, n12 = mu12 * i3,
                                 n03 = mu03 * i3;
                    const double a = n30 + n12, b = n21 + n03, c = n30 - 3.0 * n12, d = 3.0 * n21 - n03;
                    o[17] = (float)(n20 + n02);
                    o[18] = (float)((n20 - n02) * (n20 - n02) + 4.0 * n11 * n11);
                    o[19] = (float)(c * c + d * d);
                    o[20] = (float)(a * a + b * b);
                    o[21] = (float)(c * a * (a * a - 3.0 * b * b) + d * b * (3.0 * a * a - b * b));
                    o[22] = (float)((n20 - n02) * (a * a - b * b) + 4.0 * n11 * a * b);
                    o[23] = (float)(d * a * (a * a - 3.0 * b * b) - c * b * (3.0 * a * a - b * b));
                } else {
                    for (int k = 10; k < 24; ++k) o[k] = CUDART_NAN_F;
                }
            }
            if (p.col_contour >= 0) {
                out[p.col_contour] = (float)s[10];
                out[p.col_contour + 1] = (float)(s[12] + (s[11] + s[13] + 2.0 * s[14]) * 0.70710678118654752440);
            }
        }
    }
    if (p.col_color < 0) return;
    __syncthreads();
    // ---- colour moments: power sums of (x - pivot) in float64, pivot = the first masked pixel ----
    const int K = (int)s_int[0];
    if (K == 0) {
        for (int k = tid; k < 18; k += kExtThreads) out[p.col_color + k] = CUDART_NAN_F;
        return;
    }
    if (tid == 0) {
        float ch[9];
        channels(window_pixel(p, inf, s_first / P, s_first % P), ch);
        for (int k = 0; k < 9; ++k) s_piv[k] = ch[k];
    }
    __syncthreads();
    double s[36];   // [channel][power 1..4]
#pragma unroll
    for (int k = 0; k < 36; ++k) s[k] = 0.0;
    for (int k = tid; k < P * wpr; k += kExtThreads) {
        const int r = k / wpr, w = k - r * wpr;
        uint32_t bits = rows[k];
        while (bits) {
            const int c = 32 * w + __ffs(bits) - 1;
            bits &= bits - 1;
            float ch[9];
            channels(window_pixel(p, inf, r, c), ch);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const double d = (double)ch[q] - (double)s_piv[q], d2 = d * d;
                s[4 * q] += d; s[4 * q + 1] += d2; s[4 * q + 2] += d2 * d; s[4 * q + 3] += d2 * d2;
            }
        }
    }
    block_sum<36>(s, s_red);
    if (tid < 9) {
        const double Kd = (double)K, a1 = s[4 * tid] / Kd, a2 = s[4 * tid + 1] / Kd, a3 = s[4 * tid + 2] / Kd, a4 = s[4 * tid + 3] / Kd;
        const double m2 = a2 - a1 * a1, m3 = a3 - 3.0 * a1 * a2 + 2.0 * a1 * a1 * a1;
        const double m4 = a4 - 4.0 * a1 * a3 + 6.0 * a1 * a1 * a2 - 3.0 * a1 * a1 * a1 * a1;
        float sk = CUDART_NAN_F, ku = CUDART_NAN_F;
        // a channel that is constant over the mask has m2 = 0 exactly in exact arithmetic; its power sums of (x - pivot) are
        // exactly 0 here (the pivot IS that constant), so the test below is exact for that case
        if (m2 > 0.0 && a2 > 0.0) {
            sk = (float)(m3 / (m2 * sqrt(m2)));
            ku = (float)(m4 / (m2 * m2) - 3.0);
        }
        out[p.col_color + 2 * tid] = sk;
        out[p.col_color + 2 * tid + 1] = ku;
    }
}

// ------------------------------------------------------------------------------------------------
// GLCM with 32 levels at distances 1 and 2 (SPEC.md C4 = rules B5 / B6 on eight offsets): one CTA per nucleus.
// Dynamic smem: rows[P * wpr] u32 | q[P * P] u8 (level of every pixel of the window).
__device__ __constant__ int c_ext_off[8][2] = {{0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, 2}, {2, 2}, {2, 0}, {2, -2}};   // (dy, dx)

__global__ void __launch_bounds__(kExtGlcmThreads) k_ext_glcm(const ExtParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P = p.P, wpr = mask_wpr(P), tid = threadIdx.x;
    constexpr int L = 32;
    uint32_t* rows = reinterpret_cast<uint32_t*>(smem_raw);
    uint8_t* q = smem_raw + (((size_t)P * wpr * 4 + 15) & ~(size_t)15);
    __shared__ uint32_t T[L * L];
    __shared__ float pm[L * L];
    __shared__ double px[L], psum[2 * L - 1], pdif[L];
    __shared__ double s_red[16 * (kExtGlcmThreads / 32)];
    __shared__ float s_lut[256];
    __shared__ double s_mu;
    const int64_t i = blockIdx.x;
    const NucInfo inf = p.info[i];
    const uint32_t* gm = p.bitmask + i * (int64_t)P * wpr;
    for (int k = tid; k < P * wpr; k += kExtGlcmThreads) rows[k] = gm[k];
    s_lut[tid & 255] = __fdiv_rn((float)(tid & 255), 255.0f);
    __syncthreads();
    auto masked = [&](int r, int c) -> bool {
        return r >= 0 && r < P && c >= 0 && c < P && ((rows[r * wpr + (c >> 5)] >> (c & 31)) & 1u);
    };
    // grey levels of the masked pixels (texture.rs:36 + SPEC.md B5, the float32 operations of the GLCM kernels)
    for (int k = tid; k < P * P; k += kExtGlcmThreads) {
        const int r = k / P, c = k - r * P;
        uint8_t lv = 0;
        if (masked(r, c)) {
            const Rgb v = window_pixel(p, inf, r, c);
            const float g = __fdiv_rn(__fadd_rn(__fadd_rn(s_lut[v.r], s_lut[v.g]), s_lut[v.b]), 3.0f);
            lv = (uint8_t)min((int)floorf(__fmul_rn(g, 32.0f)), L - 1);
        }
        q[k] = lv;
    }
    float* out = p.out + i * (int64_t)p.out_stride + p.col_glcm;
    for (int oi = 0; oi < 8; ++oi) {
        const int dy = c_ext_off[oi][0], dx = c_ext_off[oi][1];
        __syncthreads();
        for (int k = tid; k < L * L; k += kExtGlcmThreads) T[k] = 0u;
        __syncthreads();
        for (int k = tid; k < P * P; k += kExtGlcmThreads) {
            const int r = k / P, c = k - r * P;
            if (masked(r, c) && masked(r + dy, c + dx)) {
                const int a = q[k], b = q[(r + dy) * P + c + dx];
                atomicAdd(&T[a * L + b], 1u);
                atomicAdd(&T[b * L + a], 1u);
            }
        }
        __syncthreads();
        // total (exact), normalised matrix in f32 as the reference-side rule does (G.float() / G.sum()), then float64
        double tot[1] = {0.0};
        for (int k = tid; k < L * L; k += kExtGlcmThreads) tot[0] += (double)T[k];
        block_sum<1>(tot, s_red);
        float* o = out + 14 * oi;
        if (tot[0] == 0.0) {   // 0 / 0: every feature NaN (SPEC.md B5)
            if (tid < 14) o[tid] = CUDART_NAN_F;
            continue;
        }
        const float ftot = (float)tot[0];
        for (int k = tid; k < L * L; k += kExtGlcmThreads) pm[k] = __fdiv_rn((float)T[k], ftot);
        __syncthreads();
        if (tid < L) {
            double s = 0.0;
            for (int j = 0; j < L; ++j) s += (double)pm[tid * L + j];
            px[tid] = s;
        } else if (tid < L + 2 * L - 1) {
            const int k = tid - L;
            double s = 0.0;
            for (int a = max(0, k - (L - 1)); a <= min(k, L - 1); ++a) s += (double)pm[a * L + (k - a)];
            psum[k] = s;
        } else if (tid < 2 * L + 2 * L - 1) {
            const int k = tid - (L + 2 * L - 1);
            double s = 0.0;
            for (int a = 0; a < L; ++a) {
                if (a + k < L) s += (double)pm[a * L + a + k];
                if (k > 0 && a - k >= 0) s += (double)pm[a * L + a - k];
            }
            pdif[k] = s;
        }
        __syncthreads();
        if (tid == 0) {
            double mu = 0.0;
            for (int a = 0; a < L; ++a) mu += a * px[a];
            s_mu = mu;
        }
        __syncthreads();
        const double mu = s_mu;
        // cell sums: ijp, contrast, dissimilarity, -p ln p, p^2, idm, (i-mu)^2 p, hxy1, hxy2
        double c9[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int k = tid; k < L * L; k += kExtGlcmThreads) {
            const int a = k / L, b = k - a * L;
            const double pv = (double)pm[k], d = (double)(a - b), pp = px[a] * px[b];
            c9[0] += (double)(a * b) * pv;
            c9[1] += d * d * pv;
            c9[2] += fabs(d) * pv;
            if (pv > 0.0) c9[3] -= pv * log(pv);
            c9[4] += pv * pv;
            c9[5] += pv / (1.0 + d * d);
            c9[6] += ((double)a - mu) * ((double)a - mu) * pv;
            if (pp > 0.0) { c9[7] -= pv * log(pp); c9[8] -= pp * log(pp); }
        }
        block_sum<9>(c9, s_red);
        if (tid == 0) {
            double var = 0.0, hx = 0.0, sa = 0.0, se = 0.0, da = 0.0;
            for (int a = 0; a < L; ++a) {
                var += ((double)a - mu) * ((double)a - mu) * px[a];
                if (px[a] > 0.0) hx -= px[a] * log(px[a]);
                da += a * pdif[a];
            }
            for (int k = 0; k < 2 * L - 1; ++k) {
                sa += k * psum[k];
                if (psum[k] > 0.0) se -= psum[k] * log(psum[k]);
            }
            double sv = 0.0, dv = 0.0;
            for (int k = 0; k < 2 * L - 1; ++k) sv += ((double)k - sa) * ((double)k - sa) * psum[k];
            for (int a = 0; a < L; ++a) dv += ((double)a - da) * ((double)a - da) * pdif[a];
            const double ent = c9[3];
            o[0] = (float)((c9[0] - mu * mu) / sqrt(var * var));    // sqrt(varx * vary), px = py by symmetry
            o[1] = (float)c9[1];
            o[2] = (float)c9[2];
            o[3] = (float)ent;
            o[4] = (float)c9[4];
            o[5] = (float)sa;
            o[6] = (float)sv;
            o[7] = (float)se;
            o[8] = (float)c9[6];
            o[9] = (float)c9[5];
            o[10] = (float)da;
            o[11] = (float)dv;
            o[12] = (float)((ent - c9[7]) / hx);                     // max(hx, hy) = hx
            o[13] = (float)sqrt(fmax(1.0 - exp(-2.0 * (c9[8] - ent)), 0.0));
        }
    }
}

}  // namespace

cudaError_t launch_ext(const ExtParams& p, cudaStream_t s) {
    if (p.n <= 0) return cudaSuccess;
    const int wpr = mask_wpr(p.P);
    if (p.col_color >= 0 || p.col_mask >= 0 || p.col_contour >= 0) {
        const int smem = p.P * wpr * 4;
        k_ext_stats<<<(unsigned)p.n, kExtThreads, smem, s>>>(p);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (p.col_glcm >= 0) {
        const int smem = (int)((((size_t)p.P * wpr * 4 + 15) & ~(size_t)15) + (size_t)p.P * p.P);
        if (smem > 24 * 1024) {   // static shared memory counts towards the 48 KB default limit too
            cudaError_t e = cudaFuncSetAttribute(k_ext_glcm, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
        }
        k_ext_glcm<<<(unsigned)p.n, kExtGlcmThreads, smem, s>>>(p);
        return cudaGetLastError();
    }
    return cudaSuccess;
}

}  // namespace nfx
