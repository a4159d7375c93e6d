// slide_decode.cu -- SURVEY.md 8f row 4: the slide decode path. The reference opens `.svs` files with OpenSlide and
// reads one region per nucleus under a Mutex (src/utils.rs:79-139); here level 0 of the TIFF container (tiff.cpp) is
// decoded block by block into the slide that is resident in HBM (nfx_slide_alloc). Default: the host decoder of
// jpeg_exact.cpp, whose pixels are libjpeg's bit for bit (what OpenSlide hands the reference), one block per host thread
// at a time, uploaded from pinned staging. NFX_DECODE_FAST (and streams the exact decoder does not cover): nvJPEG, so
// that the compressed bytes are all that crosses PCIe. Host threads feed nvJPEG (its Huffman stage runs on the CPU),
// each with its own decoder state, CUDA stream and device scratch; k_block_store moves a decoded block into the
// slide, clipped at the right / bottom edge (TIFF blocks are always full size) and interleaved when nvJPEG returns
// planes (Photometric = RGB means the three JPEG components ARE R, G, B: no colour transform, like libtiff / OpenSlide).
#include <nvjpeg.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "nfx_host.h"
#include "nfx_kernels.h"

namespace nfx {

namespace {

// src: interleaved RGB (planes == 0, pitch0) or three planes (planes == 1). dst: slide row-major u8 RGB.
__global__ void k_block_store(const uint8_t* s0, const uint8_t* s1, const uint8_t* s2, int pitch0, int pitch1, int pitch2,
                              int planes, uint8_t* dst, int64_t dst_pitch, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    uint8_t r, g, b;
    if (planes) {
        r = s0[(size_t)y * pitch0 + x];
        g = s1[(size_t)y * pitch1 + x];
        b = s2[(size_t)y * pitch2 + x];
    } else {
        const uint8_t* q = s0 + (size_t)y * pitch0 + 3 * x;
        r = q[0]; g = q[1]; b = q[2];
    }
    uint8_t* o = dst + (size_t)y * dst_pitch + 3 * (size_t)x;
    o[0] = r; o[1] = g; o[2] = b;
}

struct Worker {
    nvjpegJpegState_t state = nullptr;
    cudaStream_t stream = nullptr;
    uint8_t* scratch = nullptr;
    size_t scratch_bytes = 0;
    std::vector<uint8_t> merged;
    std::vector<uint8_t> rgb;        // exact path: decoded block
    uint8_t* pinned = nullptr;       // exact path: staging of the block on its way to the slide
    size_t pinned_bytes = 0;
    std::string err;
};

const char* nvjpeg_str(nvjpegStatus_t s) {
    switch (s) {
        case NVJPEG_STATUS_SUCCESS: return "success";
        case NVJPEG_STATUS_NOT_INITIALIZED: return "not initialised";
        case NVJPEG_STATUS_INVALID_PARAMETER: return "invalid parameter";
        case NVJPEG_STATUS_BAD_JPEG: return "bad JPEG stream";
        case NVJPEG_STATUS_JPEG_NOT_SUPPORTED: return "JPEG variant not supported";
        case NVJPEG_STATUS_ALLOCATOR_FAILURE: return "allocator failure";
        case NVJPEG_STATUS_EXECUTION_FAILED: return "execution failed";
        case NVJPEG_STATUS_ARCH_MISMATCH: return "architecture mismatch";
        case NVJPEG_STATUS_INTERNAL_ERROR: return "internal error";
        default: return "unknown nvJPEG status";
    }
}

}  // namespace

// Decodes every block of L into the slide at `slide` (pitch bytes per row, L.width x L.height pixels).
// Returns an empty string on success.
std::string decode_tiff_level(const uint8_t* file, const TiffLevel& L, uint8_t* slide, int64_t pitch, int device, int threads, bool fast) {
    if (L.compression != 7) return "only JPEG-compressed TIFF blocks (compression 7) are supported, found compression " + std::to_string(L.compression);
    if (L.photometric != 2 && L.photometric != 6) return "unsupported photometric interpretation " + std::to_string(L.photometric);
    nvjpegHandle_t handle = nullptr;
    nvjpegStatus_t st = nvjpegCreateSimple(&handle);
    if (st != NVJPEG_STATUS_SUCCESS) return std::string("nvjpegCreateSimple: ") + nvjpeg_str(st);
    const int64_t nblocks = (int64_t)L.offsets.size();
    const int T = (int)std::max<int64_t>(1, std::min<int64_t>(threads <= 0 ? (int)std::thread::hardware_concurrency() : threads, std::min<int64_t>(nblocks, 64)));
    std::vector<Worker> workers(T);
    std::atomic<int64_t> next{0};
    std::atomic<bool> failed{false};
    const bool rgb_components = L.photometric == 2;
    // abbreviated streams: tables (without their EOI) + block (without its SOI)
    const bool have_tables = L.jpeg_tables.size() >= 4;
    auto work = [&](int t) {
        Worker& w = workers[t];
        if (cudaSetDevice(device) != cudaSuccess) { w.err = "cudaSetDevice failed"; failed = true; return; }
        if (nvjpegJpegStateCreate(handle, &w.state) != NVJPEG_STATUS_SUCCESS) { w.err = "nvjpegJpegStateCreate failed"; failed = true; return; }
        if (cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking) != cudaSuccess) { w.err = "cudaStreamCreate failed"; failed = true; return; }
        for (;;) {
            const int64_t b = next.fetch_add(1);
            if (b >= nblocks || failed.load()) break;
            const uint8_t* data = file + L.offsets[b];
            size_t n = (size_t)L.counts[b];
            if (n == 0) continue;   // sparse file: the block stays black (the slide buffer was cleared)
            if (have_tables && n >= 2 && data[0] == 0xFF && data[1] == 0xD8) {
                w.merged.assign(L.jpeg_tables.begin(), L.jpeg_tables.end() - 2);
                w.merged.insert(w.merged.end(), data + 2, data + n);
                data = w.merged.data();
                n = w.merged.size();
            }
            const int64_t bx = b % L.across, by = b / L.across;
            const int64_t x0 = bx * L.block_w, y0 = by * L.block_h;
            if (!fast) {   // libjpeg's pixels (jpeg_exact.cpp); nvJPEG below only when the stream is outside its scope
                int ew = 0, eh = 0;
                std::string why;
                if (jpeg_decode_exact(data, n, rgb_components ? 0 : 1, w.rgb, ew, eh, why)) {
                    const int vw = (int)std::min<int64_t>(std::min<int64_t>(ew, L.block_w), L.width - x0);
                    const int vh = (int)std::min<int64_t>(std::min<int64_t>(eh, L.block_h), L.height - y0);
                    if (vw > 0 && vh > 0) {
                        // the staging buffer is reused by the next block of this worker: wait for the previous upload
                        if (cudaStreamSynchronize(w.stream) != cudaSuccess) { w.err = "decode stream failed"; failed = true; break; }
                        if (w.rgb.size() > w.pinned_bytes) {
                            if (w.pinned) cudaFreeHost(w.pinned);
                            w.pinned = nullptr;
                            if (cudaMallocHost((void**)&w.pinned, w.rgb.size()) != cudaSuccess) { w.err = "out of pinned memory"; failed = true; break; }
                            w.pinned_bytes = w.rgb.size();
                        }
                        memcpy(w.pinned, w.rgb.data(), w.rgb.size());
                        if (cudaMemcpy2DAsync(slide + (size_t)y0 * pitch + 3 * (size_t)x0, (size_t)pitch, w.pinned, (size_t)3 * ew, (size_t)3 * vw,
                                              (size_t)vh, cudaMemcpyHostToDevice, w.stream) != cudaSuccess) { w.err = "block upload failed"; failed = true; break; }
                    }
                    continue;
                }
            }
            int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
            nvjpegChromaSubsampling_t sub;
            nvjpegStatus_t s = nvjpegGetImageInfo(handle, data, n, &nc, &sub, ws, hs);
            if (s != NVJPEG_STATUS_SUCCESS) { w.err = std::string("block ") + std::to_string(b) + ": " + nvjpeg_str(s); failed = true; break; }
            if (nc != 3) { w.err = "block " + std::to_string(b) + " is not a 3-component JPEG"; failed = true; break; }
            const int bw = ws[0], bh = hs[0];
            const bool planes = rgb_components;
            if (planes && (ws[1] != bw || ws[2] != bw || hs[1] != bh || hs[2] != bh)) {
                w.err = "block " + std::to_string(b) + ": subsampled components in an RGB-photometric JPEG";
                failed = true;
                break;
            }
            const size_t need = (size_t)bw * bh * 3 + 1024;
            if (need > w.scratch_bytes) {
                if (w.scratch) cudaFree(w.scratch);
                w.scratch = nullptr;
                if (cudaMalloc((void**)&w.scratch, need) != cudaSuccess) { w.err = "out of device memory"; failed = true; break; }
                w.scratch_bytes = need;
            }
            nvjpegImage_t img;
            memset(&img, 0, sizeof img);
            if (planes) {
                for (int c = 0; c < 3; ++c) { img.channel[c] = w.scratch + (size_t)c * bw * bh; img.pitch[c] = (size_t)bw; }
            } else {
                img.channel[0] = w.scratch;
                img.pitch[0] = (size_t)bw * 3;
            }
            s = nvjpegDecode(handle, w.state, data, n, planes ? NVJPEG_OUTPUT_UNCHANGED : NVJPEG_OUTPUT_RGBI, &img, w.stream);
            if (s != NVJPEG_STATUS_SUCCESS) { w.err = std::string("block ") + std::to_string(b) + ": " + nvjpeg_str(s); failed = true; break; }
            const int vw = (int)std::min<int64_t>(std::min<int64_t>(bw, L.block_w), L.width - x0);
            const int vh = (int)std::min<int64_t>(std::min<int64_t>(bh, L.block_h), L.height - y0);
            if (vw > 0 && vh > 0) {
                dim3 blk(32, 8), grd((vw + 31) / 32, (vh + 7) / 8);
                k_block_store<<<grd, blk, 0, w.stream>>>(img.channel[0], img.channel[1], img.channel[2], (int)img.pitch[0], (int)img.pitch[1],
                                                        (int)img.pitch[2], planes ? 1 : 0, slide + (size_t)y0 * pitch + 3 * (size_t)x0, pitch, vw, vh);
                if (cudaGetLastError() != cudaSuccess) { w.err = "k_block_store launch failed"; failed = true; break; }
            }
            // the scratch is reused by the next block of this worker: same stream, so the order is kept
        }
        if (w.stream && cudaStreamSynchronize(w.stream) != cudaSuccess && w.err.empty()) { w.err = "decode stream failed"; failed = true; }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < T; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    std::string err;
    for (auto& w : workers) {
        if (err.empty() && !w.err.empty()) err = w.err;
        if (w.scratch) cudaFree(w.scratch);
        if (w.pinned) cudaFreeHost(w.pinned);
        if (w.stream) cudaStreamDestroy(w.stream);
        if (w.state) nvjpegJpegStateDestroy(w.state);
    }
    nvjpegDestroy(handle);
    return err;
}

}  // namespace nfx
