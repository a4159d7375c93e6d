// Host-side helpers shared by schema.cpp and api.cu.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace nfx {
std::string rust_f32_display(float x);
int set_cols(int set_index);                         // columns of set 0..4 (flat() order)
int column_offset(uint32_t mask, uint32_t bit);      // first column of `bit` within `mask`, or -1
extern thread_local std::string g_thread_error;      // last error without a context

// Directory 0 of a TIFF / BigTIFF as a flat table of compressed blocks (tiles, or full-width strips). tiff.cpp
struct TiffLevel {
    int64_t width = 0, height = 0;
    int block_w = 0, block_h = 0;
    int64_t across = 0, down = 0;            // blocks per row / per column, row-major table
    int compression = 1, photometric = 2, samples = 1;
    std::vector<int64_t> offsets, counts;    // byte range of every block inside the file
    std::vector<uint8_t> jpeg_tables;        // tag 347: SOI DQT DHT EOI shared by abbreviated block streams (may be empty)
};
bool tiff_parse(const uint8_t* file, int64_t len, TiffLevel& out, std::string& err);
// slide_decode.cu: every block of L into the slide (device pointer); "" on success. fast = false: the host decoder of
// jpeg_exact.cpp (libjpeg's pixels), nvJPEG only for streams it does not cover; fast = true: nvJPEG for every block.
std::string decode_tiff_level(const uint8_t* file, const TiffLevel& L, uint8_t* slide, int64_t pitch, int device, int threads, bool fast);
// jpeg_exact.cpp: baseline JPEG -> interleaved u8 RGB with libjpeg / libjpeg-turbo's default arithmetic, bit for bit.
// colourspace: 0 = components are R,G,B; 1 = YCbCr; -1 = libjpeg's own rule (JFIF / Adobe marker / component ids)
bool jpeg_decode_exact(const uint8_t* data, size_t len, int colourspace, std::vector<uint8_t>& out, int& W, int& H, std::string& err);
}  // namespace nfx
