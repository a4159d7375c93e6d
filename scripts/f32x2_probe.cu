// Issue rate of FFMA (three register operands) against the packed FFMA2 / FADD2 (f32x2) on one SM of a B200:
//   nvcc -arch=sm_100a -O3 scripts/f32x2_probe.cu -o scripts/f32x2_probe && scripts/f32x2_probe
// Prints warp-instructions per cycle and SM for 4..32 resident warps. Used to decide whether packing two pixels per
// instruction pays in k_color_warp / k_gabor (profiles/r2_f32x2_probe.txt).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ int imad(int a, int b, int c) { int d; asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

template <int MODE> __global__ void k(float* out, long long* cyc, float x, float y, int iters) {
    float a[8]; u64 p[8]; int q[8];
    const float m = x + threadIdx.x * 1e-9f, n = y;
    u64 mm, nn; asm("mov.b64 %0, {%1, %2};" : "=l"(mm) : "f"(m), "f"(m)); asm("mov.b64 %0, {%1, %2};" : "=l"(nn) : "f"(n), "f"(n));
    for (int i = 0; i < 8; ++i) { a[i] = i; asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"((float)i), "f"((float)i)); q[i] = i; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = fma1(a[i], m, n);
            if (MODE == 1) p[i] = fma2(p[i], mm, nn);
            if (MODE == 2) p[i] = add2(p[i], mm);
            if (MODE == 3) { a[i] = fma1(a[i], m, n); q[i] = imad(q[i], threadIdx.x, q[i]); }   // FFMA + IMAD (both fma pipe)
            if (MODE == 4) { a[i] = fma1(a[i], m, n); q[i] = (q[i] ^ (int)threadIdx.x) + it; }    // FFMA + ALU work
            if (MODE == 5) { p[i] = fma2(p[i], mm, nn); q[i] = (q[i] ^ (int)threadIdx.x) + it; }  // FFMA2 + ALU work
        }
    }
    const long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += a[i] + lo + hi + q[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int per_iter) {
    float* out; long long* cyc; cudaMalloc(&out, 4 * 1024); cudaMalloc(&cyc, 8);
    for (int warps : {4, 8, 16, 32}) {
        const int iters = 4096;
        k<MODE><<<1, warps * 32>>>(out, cyc, 1.0001f, 0.5f, iters);
        k<MODE><<<1, warps * 32>>>(out, cyc, 1.0001f, 0.5f, iters);
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-28s warps %2d: %.3f warp-instr/cycle/SM (%.3f per SMSP)\n", name, warps, (double)iters * per_iter * warps / c, (double)iters * per_iter * warps / c / 4);
    }
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("FFMA r,r,r", 8); run<1>("FFMA2", 8); run<2>("FADD2", 8); run<3>("FFMA + IMAD", 16); run<4>("FFMA + LOP3/IADD", 24); run<5>("FFMA2 + LOP3/IADD", 24);
    cudaError_t e = cudaDeviceSynchronize(); if (e) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
