// f32batch.cu -- the trait-level entry (FeatureSet::compute_features_batched, src/features/mod.rs) for batches whose patch
// values are NOT k/255. The reference's own loaders only ever produce k/255 (src/utils.rs:172), and those batches take the
// u8 kernels; a caller that hands the trait arbitrary f32 tensors in [0,1] lands here instead of being refused:
//   * the texture sets run their usual kernels on an f32 grey plane (k_grey_f32 -> GlcmParams::grey / TexParams::grey);
//   * the colour set is evaluated straight from the f32 patches with the reference's formulas
//       hsv_from_rgb / hed_from_rgb   color.rs:45-46   (oracle/SPEC.md B3, B4)
//       mean_std x3                   color.rs:47-53, 117-134
//       circular_mean (batch coupled) color.rs:50-51, 144-155
//     in float64 sums (this path is written for clarity, not speed: one CTA per nucleus, no TMA, no tables).
// The geometry set needs no pixel and is shared with the u8 path.
#include <math_constants.h>

#include "nfx_kernels.h"

namespace nfx {

namespace {

// inv([[0.65,0.70,0.29],[0.07,0.99,0.11],[0.27,0.57,0.78]]) in f32, M[c][k] (oracle HED_FROM_RGB): hed_k = sum_c od_c M[c][k]
__device__ __constant__ float c_hed[3][3] = {{1.87798273563385f, -1.0076786279678345f, -0.5561158061027527f},
                                             {-0.06590805947780609f, 1.134730339050293f, -0.135521799325943f},
                                             {-0.6019073724746704f, -0.48041418194770813f, 1.5735880136489868f}};

__global__ void k_grey_f32(const int64_t n, const int P, const float* __restrict__ patchs, float* __restrict__ grey) {
    const int64_t plane = (int64_t)P * P, k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n * plane) return;
    const int64_t i = k / plane, px = k - i * plane;
    const float* b = patchs + i * 3 * plane + px;
    grey[k] = __fdiv_rn(__fadd_rn(__fadd_rn(b[0], b[plane]), b[2 * plane]), 3.0f);   // mean_dim(-3)
}

// hue in degrees, [0, 360) (oracle hsv_from_rgb)
__device__ __forceinline__ float hue_deg(float r, float g, float b, float mx, float d) {
    if (d == 0.f) return 0.f;
    float h;
    if (mx == r) {
        h = __fdiv_rn(g - b, d);
        h = h - 6.0f * floorf(__fdiv_rn(h, 6.0f));   // torch.remainder(x, 6): sign of the divisor
    } else if (mx == g) {
        h = __fdiv_rn(b - r, d) + 2.0f;
    } else {
        h = __fdiv_rn(r - g, d) + 4.0f;
    }
    return h * 60.0f;
}

constexpr int kF32Threads = 128;

// one CTA per nucleus: float64 sums of the nine channels and of their squares under the mask
__global__ void __launch_bounds__(kF32Threads)
k_color_f32(const int P, const float* __restrict__ patchs, const uint32_t* __restrict__ bitmask, float* __restrict__ out,
            const int out_stride, const int col_color) {
    const int64_t i = blockIdx.x, plane = (int64_t)P * P;
    const int wpr = mask_wpr(P), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* pr = patchs + i * 3 * plane;
    const uint32_t* gm = bitmask + i * (int64_t)P * wpr;
    double s1[9], s2[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) { s1[q] = 0.0; s2[q] = 0.0; }
    int K = 0;
    for (int k = tid; k < P * P; k += kF32Threads) {
        const int r_ = k / P, c_ = k - r_ * P;
        if (!((gm[r_ * wpr + (c_ >> 5)] >> (c_ & 31)) & 1u)) continue;
        ++K;
        const float r = pr[k], g = pr[plane + k], b = pr[2 * plane + k];
        const float mx = fmaxf(r, fmaxf(g, b)), mn = fminf(r, fminf(g, b)), d = mx - mn;
        float v[9];
        v[0] = r; v[1] = g; v[2] = b;
        v[3] = hue_deg(r, g, b, mx, d);
        v[4] = mx > 0.f ? __fdiv_rn(d, mx) : 0.f;
        v[5] = mx;
        const float od0 = __fdiv_rn(logf(fmaxf(r, 1e-6f)), -13.815510749816895f), od1 = __fdiv_rn(logf(fmaxf(g, 1e-6f)), -13.815510749816895f),
                    od2 = __fdiv_rn(logf(fmaxf(b, 1e-6f)), -13.815510749816895f);
#pragma unroll
        for (int q = 0; q < 3; ++q) v[6 + q] = fmaxf(od0 * c_hed[0][q] + od1 * c_hed[1][q] + od2 * c_hed[2][q], 0.f);
#pragma unroll
        for (int q = 0; q < 9; ++q) { s1[q] += (double)v[q]; s2[q] += (double)v[q] * (double)v[q]; }
    }
    __shared__ double s_s[kF32Threads / 32][18];
    __shared__ int s_k[kF32Threads / 32];
    K = __reduce_add_sync(0xffffffffu, K);
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        const double a = warp_sum(s1[q]), b = warp_sum(s2[q]);
        if (lane == 0) { s_s[warp][q] = a; s_s[warp][9 + q] = b; }
    }
    if (lane == 0) s_k[warp] = K;
    __syncthreads();
    if (tid < 18 && tid != 6) {
        // column -> channel:  mean_r g b | std_r g b | (mean_h) | mean_s mean_v | std_h std_s std_v | mean_hed x3 | std_hed x3
        static const int ch[18] = {0, 1, 2, 0, 1, 2, 0, 4, 5, 3, 4, 5, 6, 7, 8, 6, 7, 8};
        const bool is_std = (tid >= 3 && tid <= 5) || (tid >= 9 && tid <= 11) || tid >= 15;
        double a = 0.0, b = 0.0, Kd = 0.0;
        for (int w = 0; w < kF32Threads / 32; ++w) { a += s_s[w][ch[tid]]; b += s_s[w][9 + ch[tid]]; Kd += (double)s_k[w]; }
        const double m = a / Kd;   // empty mask: 0/0 = NaN like the reference
        out[i * (int64_t)out_stride + col_color + tid] = (float)(is_std ? sqrt(fmax(b / Kd - m * m, 0.0)) : m);
    }
}

// circular_mean broadcasts [N,P,P] * [N,1,P,P] to [N,N,P,P]: per chunk of the batch the images C[p] = sum_j cos h_j[p],
// S[p] = sum_j sin h_j[p] (patches added in index order), then masked sums per nucleus.
__global__ void k_hue_f32_images(const int64_t n, const int P, const int B, const float* __restrict__ patchs, float* __restrict__ img) {
    const int64_t plane = (int64_t)P * P, chunk = blockIdx.y;
    const int64_t px = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= plane) return;
    const int64_t j0 = chunk * B, j1 = j0 + B < n ? j0 + B : n;
    float C = 0.f, S = 0.f;
    for (int64_t j = j0; j < j1; ++j) {
        const float* b = patchs + j * 3 * plane + px;
        const float r = b[0], g = b[plane], bl = b[2 * plane];
        const float mx = fmaxf(r, fmaxf(g, bl)), mn = fminf(r, fminf(g, bl));
        const float a = hue_deg(r, g, bl, mx, mx - mn) * 0.017453292519943295f;
        float sn, cs;
        sincosf(a, &sn, &cs);
        C += cs;
        S += sn;
    }
    img[(chunk * plane + px) * 2] = S;
    img[(chunk * plane + px) * 2 + 1] = C;
}

// one warp per nucleus
__global__ void k_hue_f32_fold(const int64_t n, const int P, const int B, const uint32_t* __restrict__ bitmask,
                               const float* __restrict__ img, float* __restrict__ out, const int out_stride, const int col_color) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, wpr = mask_wpr(P);
    if (i >= n) return;
    const float* im = img + (i / B) * (int64_t)P * P * 2;
    const uint32_t* gm = bitmask + i * (int64_t)P * wpr;
    double ss = 0.0, sc = 0.0;
    int K = 0;
    for (int k = lane; k < P * P; k += 32) {
        const int r = k / P, c = k - r * P;
        if ((gm[r * wpr + (c >> 5)] >> (c & 31)) & 1u) { ss += (double)im[2 * k]; sc += (double)im[2 * k + 1]; ++K; }
    }
    ss = warp_sum(ss);
    sc = warp_sum(sc);
    K = __reduce_add_sync(0xffffffffu, K);
    if (lane == 0) {
        // color.rs:154  (atan2(sin, cos).rad2deg + 360) fmod 360 ; empty mask -> 0/0 -> NaN
        float deg = atan2f((float)ss, (float)sc) * 57.29577951308232f;
        deg = fmodf(deg + 360.0f, 360.0f);
        out[i * (int64_t)out_stride + col_color + 6] = K ? deg : CUDART_NAN_F;
    }
}

}  // namespace

cudaError_t launch_grey_f32(int64_t n, int P, const float* patchs, float* grey, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    const int64_t total = n * P * P;
    k_grey_f32<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(n, P, patchs, grey);
    return cudaGetLastError();
}

cudaError_t launch_color_f32(int64_t n, int P, int batch_size, const float* patchs, const uint32_t* bitmask, float* hue_images,
                             float* out, int out_stride, int col_color, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    k_color_f32<<<(unsigned)n, kF32Threads, 0, s>>>(P, patchs, bitmask, out, out_stride, col_color);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const int64_t chunks = (n + batch_size - 1) / batch_size;
    if (chunks > 65535) return cudaErrorInvalidValue;   // a trait call is ONE batch; 65 535 chunks are far beyond it
    k_hue_f32_images<<<dim3((unsigned)((P * P + 255) / 256), (unsigned)chunks), 256, 0, s>>>(n, P, batch_size, patchs, hue_images);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_hue_f32_fold<<<(unsigned)((n * 32 + 255) / 256), 256, 0, s>>>(n, P, batch_size, bitmask, hue_images, out, out_stride, col_color);
    return cudaGetLastError();
}

}  // namespace nfx
