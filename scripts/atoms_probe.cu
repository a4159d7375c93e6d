// Probe: shared-memory atomic throughput on B200 (peak for the GLCM roofline).
// modes: 0 = conflict-free (lane -> own bank), 1 = random addresses in a 16K-word table,
//        2 = all lanes same address, 3 = random 16-bit halves packed in 32-bit words (tri histogram)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__global__ void k(int mode, int iters, unsigned* sink) {
    extern __shared__ unsigned tab[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            unsigned idx;
            if (mode == 0) idx = ((it * 8 + u) * 32 + (threadIdx.x & 31) + (threadIdx.x >> 5) * 1024) & 16383;
            else if (mode == 2) idx = (threadIdx.x >> 5);
            else { x = x * 1664525u + 1013904223u; idx = (x >> 9) & 16383; }
            if (mode == 3) atomicAdd(&tab[idx], 1u << (((x >> 5) & 1) * 16));
            else atomicAdd(&tab[idx], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) sink[blockIdx.x] = tab[blockIdx.x & 1023];
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    unsigned* sink; cudaMalloc(&sink, 4 * 4096);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int threads : {256, 512, 1024}) for (int mode = 0; mode < 4; ++mode) {
        int blocks = pr.multiProcessorCount * (threads == 1024 ? 2 : 3), iters = 2000;
        k<<<blocks, threads, 65536>>>(mode, 10, sink);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a); k<<<blocks, threads, 65536>>>(mode, iters, sink); cudaEventRecord(b);
        cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b);
        double ops = (double)blocks * threads * iters * 8;
        printf("threads=%4d mode=%d: %.1f G atomics/s  (%.2f lane-atomics/clk/SM at 1.9 GHz, %d SMs)\n", threads, mode,
               ops / ms / 1e6, ops / (ms * 1e-3) / pr.multiProcessorCount / 1.9e9, pr.multiProcessorCount);
    }
    return 0;
}
