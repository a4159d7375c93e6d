// csv.cu -- output assembly on the device: the rows polars' CsvWriter writes for the
// reference's DataFrame (src/main.rs:76-89 hstack behind the `centroid` key of src/utils.rs:226-232,
// src/main.rs:163-166 writer; oracle/SPEC.md B12):
//     "cx,cy",f0,f1,...,f{F-1}\n
// with every f32 in Rust `Display` form (f32_display.h). The feature matrix and the centroids are already in
// HBM; formatting 419 cells per nucleus on host cores is the slowest stage of a 5 M nuclei job, so it is
// done here: k_csv_measure (cell lengths -> row lengths), an exclusive scan, k_csv_write (one warp per
// row: cells formatted into a shared-memory line buffer, flushed with coalesced stores).
#include <cub/device/device_scan.cuh>

#include "f32_display.h"
#include "nfx_kernels.h"

namespace nfx {

__device__ const uint64_t d_pow5_inv[31] = NFX_POW5_INV_SPLIT;
__device__ const uint64_t d_pow5[47] = NFX_POW5_SPLIT;

namespace {

constexpr int kWarps = 8;                       // warps (rows in flight) per CTA
constexpr int kChunkBytes = 32 * (NFX_F32_MAX_CHARS + 2);   // 32 cells of one row, worst case

__device__ __forceinline__ uint32_t cell_bits(const CsvParams& p, int64_t row, int col) {
    const float v = col < 2 ? reinterpret_cast<const float*>(p.centroids)[2 * row + col]
                            : p.features[row * p.F + (col - 2)];
    return __float_as_uint(v);
}
// bytes the cell occupies with what follows it: x -> ',' ; y -> '"' and ',' or '\n' ; feature -> ',' or '\n'
__device__ __forceinline__ int cell_extra(int col) { return col == 1 ? 2 : 1; }

__global__ void __launch_bounds__(kWarps * 32) k_csv_measure(CsvParams p) {
    const int lane = threadIdx.x & 31;
    const int ncols = p.F + 2;
    for (int64_t r = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); r < p.rows; r += (int64_t)gridDim.x * kWarps) {
        const int64_t row = p.row_lo + r;
        int len = 0;
        for (int c = lane; c < ncols; c += 32) len += f32_display_len(cell_bits(p, row, c), d_pow5_inv, d_pow5) + cell_extra(c);
        len = __reduce_add_sync(0xffffffffu, len);
        if (lane == 0) p.row_len[r] = (int64_t)len + 1;   // the opening quote of the key
    }
}

__global__ void __launch_bounds__(kWarps * 32) k_csv_write(CsvParams p) {
    __shared__ __align__(16) char line[kWarps][kChunkBytes];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ncols = p.F + 2;
    char* buf = line[w];
    for (int64_t r = (int64_t)blockIdx.x * kWarps + w; r < p.rows; r += (int64_t)gridDim.x * kWarps) {
        const int64_t row = p.row_lo + r;
        char* dst = p.text + p.row_off[r];
        if (lane == 0) *dst = '"';
        ++dst;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
            const int c = c0 + lane;
            Cell32 cell;
            int len = 0;
            if (c < ncols) {
                cell = f32_cell(cell_bits(p, row, c), d_pow5_inv, d_pow5);
                len = cell_len(cell) + cell_extra(c);
            }
            // exclusive scan of the cell lengths of this chunk
            int incl = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (c < ncols) {
                char* q = buf + (incl - len);
                q += cell_write(cell, q);
                if (c == 1) *q++ = '"';
                *q = (c != 0 && c == ncols - 1) ? '\n' : ',';
            }
            __syncwarp();
            for (int k = lane; k < total; k += 32) dst[k] = buf[k];
            __syncwarp();
            dst += total;
        }
    }
}

}  // namespace

cudaError_t csv_scan_bytes(int64_t rows, size_t* bytes) {
    *bytes = 0;
    return cub::DeviceScan::ExclusiveSum(nullptr, *bytes, (const int64_t*)nullptr, (int64_t*)nullptr, (int)(rows + 1));
}

cudaError_t launch_csv_measure(const CsvParams& p, void* scan_tmp, size_t scan_bytes, cudaStream_t s) {
    if (p.rows <= 0) return cudaSuccess;
    const int grid = (int)std::min<int64_t>((p.rows + kWarps - 1) / kWarps, 148 * 8);
    // row_len has rows+1 entries (the last is a zero written by the caller) so that the scan also yields the total
    k_csv_measure<<<grid, kWarps * 32, 0, s>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, (const int64_t*)p.row_len, p.row_off, (int)(p.rows + 1), s);
}

cudaError_t launch_csv_write(const CsvParams& p, cudaStream_t s) {
    if (p.rows <= 0) return cudaSuccess;
    const int grid = (int)std::min<int64_t>((p.rows + kWarps - 1) / kWarps, 148 * 8);
    k_csv_write<<<grid, kWarps * 32, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace nfx
