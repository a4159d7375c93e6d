// Throughput of the warp-level mma.sync (legacy HMMA path) on one SM of a B200, to size a tensor-core Gabor bank:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/mma_probe.cu -o /tmp/mma_probe && /tmp/mma_probe
// Prints MMA instructions per cycle and SM, and the MAC rate per SM and clock, for 4..32 resident warps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// MODE 0: m16n8k16 f16 x f16 -> f32 ; 1: m16n8k16 bf16 -> f32 ; 2: m16n8k8 tf32 -> f32 ; 3: m16n8k16 f16 -> f16
template <int MODE> __global__ void k(float* out, long long* cyc, uint32_t seed, int iters) {
    constexpr int NACC = 8;
    float c[NACC][4];
    uint32_t a[4] = {seed, seed ^ 0x3c003c00u, seed + threadIdx.x, seed}, b[2] = {seed ^ 0x38003800u, seed};
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = 0; c[i][2] = 1; c[i][3] = 2; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (MODE == 0)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            if (MODE == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            if (MODE == 2)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            if (MODE == 3) {
                uint32_t* h = reinterpret_cast<uint32_t*>(c[i]);
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3,%4,%5}, {%6,%7}, {%0,%1};"
                             : "+r"(h[0]), "+r"(h[1]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int macs) {
    float* out; long long* cyc; cudaMalloc(&out, 4 * 1024); cudaMalloc(&cyc, 8);
    for (int warps : {4, 8, 16, 32}) {
        const int iters = 4096;
        k<MODE><<<1, warps * 32>>>(out, cyc, 0x3c003c00u, iters);
        k<MODE><<<1, warps * 32>>>(out, cyc, 0x3c003c00u, iters);
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        const double per_cycle = (double)iters * 8 * warps / c;
        printf("%-32s warps %2d: %.4f mma/cycle/SM = %.0f MAC/cycle/SM (%.1f cycles per mma and SMSP)\n", name, warps, per_cycle, per_cycle * macs, 4.0 / per_cycle);
    }
    cudaFree(out); cudaFree(cyc);
}
// whole chip, all SMs busy: sustained rate under power
template <int MODE> void run_chip(const char* name, int macs) {
    float* out; long long* cyc; cudaMalloc(&out, 4 * 148 * 8 * 512); cudaMalloc(&cyc, 8 * 148 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1 << 16, warps = 16;
    k<MODE><<<148 * 2, warps * 32>>>(out, cyc, 0x3c003c00u, 1024);
    cudaEventRecord(e0);
    k<MODE><<<148 * 2, warps * 32>>>(out, cyc, 0x3c003c00u, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)iters * 8 * warps * 148 * 2;
    printf("%-32s chip: %.3f ms, %.1f TMAC/s = %.1f TFLOP/s\n", name, ms, mmas * macs / ms / 1e9, 2 * mmas * macs / ms / 1e9);
}
int main() {
    run<0>("m16n8k16 f16 -> f32", 2048); run<1>("m16n8k16 bf16 -> f32", 2048); run<2>("m16n8k8 tf32 -> f32", 1024); run<3>("m16n8k16 f16 -> f16", 2048);
    run_chip<0>("m16n8k16 f16 -> f32", 2048); run_chip<2>("m16n8k8 tf32 -> f32", 1024);
    cudaError_t e = cudaDeviceSynchronize(); if (e) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
