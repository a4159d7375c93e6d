// Host-visible launchers of the sm_100a kernels (one translation unit per kernel family).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "nfx_device.cuh"

namespace nfx {

// Column counts per feature set (schema.cpp holds the names).
constexpr int kShapeCols = 12;
constexpr int kColorCols = 18;
constexpr int kGlcmFeat = 14;
constexpr int kGlcmLevels = 4;
constexpr int kGlcmOffsets = 4;
constexpr int kGlcmCols = kGlcmFeat * kGlcmLevels * kGlcmOffsets;   // 224
constexpr int kGlrlmCols = 68;
constexpr int kGaborCols = 96;

// ---- geom.cu: centroid + window + polygon raster + shape set --------------------------------
struct GeomParams {
    const float2* poly_xy;     // CSR vertices (slide coordinates, or CENTRED when centred_input)
    const int64_t* poly_off;   // [n+1]
    int64_t n;
    int P;
    int tile_ox, tile_oy;      // slide coordinates of tile pixel (0,0)
    int vmax;                  // max ring length in this launch
    int vsmem;                 // ring capacity of the shared-memory layout = min(vmax, kGeomRingSmem); longer rings work in HBM:
    int64_t n_giant;           // rings longer than vsmem: computed by a second launch over giant_list
    const int* giant_list;     // [n_giant] nucleus index of every long ring (slot s works in ring_scratch slot s)
    unsigned char* ring_scratch;   // [n_giant][geom_ring_bytes(vmax)] pts | sorted | stk of the long rings
    float2* centroid;          // [n] out (RASTER) / unused
    NucInfo* info;             // [n] out (RASTER) / unused
    uint32_t* bitmask;         // [n][P][wpr]: out when RASTER, in otherwise
    float* out;                // [n][out_stride] feature matrix, or nullptr
    int out_stride;
    int col_shape;             // first column of the geometry set in `out`, or -1
    uint32_t* ellipse_bits;    // optional debug tap [n][P][wpr], or nullptr
    float sample_off;          // 0 (SPEC.md B1/B2) or 0.5 (NFX_RULE_RASTER_PIXEL_CENTRE): pixel k samples k + off - P/2
    int slide_window;          // NFX_RULE_WINDOW_SLIDE: window origin saturates at 0, always P x P (utils.rs:96-109)
};
// raster=true : polygons are raw rings; computes centroid/info, rasterises, writes bitmask.
// raster=false: polygons are already centred and bitmask is an input (trait-level path).
cudaError_t launch_geom(const GeomParams& p, bool raster, bool shape, cudaStream_t s);
constexpr int kGeomRingSmem = 1400;   // vertices of a ring kept in shared memory (32 B each with the hull's work arrays, four nuclei per CTA)
size_t geom_ring_bytes(int vmax);     // bytes of one ring_scratch slot

// ---- color.cu ---------------------------------------------------------------------------------
struct ColorParams {
    int64_t n;
    int P;
    int batch_size;
    const NucInfo* info;
    const uint32_t* bitmask;
    float* out;
    int out_stride;
    int col_color;             // first column of the colour set
    float* hue_partial;        // [n][slabs][2] (sum sin, sum cos) scratch
    int slabs;
    int slab_rows;             // k_color row-slab height = color_slab_rows(P)
};
cudaError_t launch_color(const ColorParams& p, const CUtensorMap* map, cudaStream_t s);
cudaError_t launch_color_warp(const ColorParams& p, const CUtensorMap* map_slab16, cudaStream_t s);   // P == 64
cudaError_t launch_hue_batch(const ColorParams& p, const CUtensorMap* map_slab, int slab_rows,
                             cudaStream_t s);
cudaError_t launch_hue_finalize(const ColorParams& p, cudaStream_t s);
int color_smem_bytes(int P);
int color_slab_rows(int P);
int hue_slab_rows(int P);

// ---- glcm.cu ----------------------------------------------------------------------------------
struct GlcmParams {
    int64_t n;
    int P;
    const NucInfo* info;
    const uint32_t* bitmask;
    float* out;
    int out_stride;
    int col_glcm;
    // debug taps (nullptr unless requested)
    uint32_t* dbg_counts;      // [n][L][L] symmetric counts for (dbg_levels, dbg_dy, dbg_dx)
    int dbg_levels, dbg_dy, dbg_dx;
    uint8_t* dbg_grey;         // [n][P][P] quantised grey for dbg_levels
    float scale254;            // 254.0f, or 255.0f under NFX_RULE_GLCM_254_U8: q254 = min(floor(g * scale254), 253)
    const float* grey;         // nullptr, or [n][P][P] f32 grey planes (f32batch.cu): the window is then neither fetched nor looked up
};
// map = whole-window map for P <= 128, the 64-row slab map (color_slab_rows) when glcm_uses_slab_map(P)
cudaError_t launch_glcm(const GlcmParams& p, const CUtensorMap* map, cudaStream_t s);
int glcm_uses_slab_map(int P);

// ---- texture2.cu: GLRLM and Gabor sets -------------------------------------------------------------
struct TexParams {
    int64_t n;
    int P;
    int slab_rows;             // color_slab_rows(P): TMA box height of the slab map
    const NucInfo* info;
    const uint32_t* bitmask;
    float* out;
    int out_stride;
    int col_glrlm, col_gabor;  // first column of each set (or -1)
    double* gabor_partial;     // [n][gabor_tiles(P)][97] scratch when P > 64, else nullptr
    int gabor_half_turn;       // NFX_RULE_GABOR_HALF_TURN: angles i * pi / 8 (48 distinct filters) instead of i * 2 pi / 8
    const float* grey;         // nullptr, or [n][P][P] f32 grey planes (f32batch.cu): the window is then neither fetched nor looked up
};
cudaError_t launch_glrlm(const TexParams& p, const CUtensorMap* map_cslab, cudaStream_t s);
cudaError_t launch_gabor(const TexParams& p, const CUtensorMap* map, cudaStream_t s);   // map: see texture2.cu
int gabor_max_patch();
int gabor_tiles(int P);
int gabor_fetch_rows();

// ---- staged.cu: kernels (1) gather and the f32 batch packer -----------------------------------
// tile window -> u8 patch array [n*P rows][pitch bytes] through TMA load + TMA store.
cudaError_t launch_gather(int64_t n, int P, const NucInfo* info, const CUtensorMap* map_tile,
                          uint8_t* patches, int64_t pitch, cudaStream_t s);
// expand bitmask -> u8 0/1 masks [n][P][P]
cudaError_t launch_expand_mask(int64_t n, int P, const uint32_t* bitmask, uint8_t* out,
                               cudaStream_t s);
// reference Batch layout (patchs [n,3,P,P] f32 k/255, masks [n,1,P,P] f32) -> u8 patch array +
// bitmask + NucInfo rows pointing into the patch array.
cudaError_t launch_pack_batch(int64_t n, int P, const float* patchs, const float* masks,
                              uint8_t* patches_u8, int64_t pitch, uint32_t* bitmask, NucInfo* info,
                              int* bad_count, cudaStream_t s);

// ---- f32batch.cu: trait-level batches whose patch values are NOT k/255 (arbitrary f32 in [0,1]) ----
// grey = ((r + g) + b) / 3 with IEEE f32 operations (texture.rs:36 / 189 / 332) for the texture kernels' `grey` input
cudaError_t launch_grey_f32(int64_t n, int P, const float* patchs, float* grey, cudaStream_t s);
// the colour set (color.rs:10-102) straight from f32 patches: 17 columns per nucleus + the batch-coupled mean_h;
// hue_images = [ceil(n / batch_size)][P*P][2] f32 scratch
cudaError_t launch_color_f32(int64_t n, int P, int batch_size, const float* patchs, const uint32_t* bitmask, float* hue_images,
                             float* out, int out_stride, int col_color, cudaStream_t s);

// ---- ext.cu: extension outputs (north_star items the reference does not compute; never part of the drop-in schema) ----
constexpr int kExtColorCols = 18, kExtMaskCols = 24, kExtContourCols = 2, kExtGlcmCols = 112;
struct ExtParams {
    int64_t n;
    int P;
    const NucInfo* info;
    const uint32_t* bitmask;
    const uint8_t* tile;       // the resident slide (u8 interleaved RGB)
    int64_t tpitch, tw, th;
    float* out;                // [n][out_stride]
    int out_stride;
    int col_color, col_mask, col_contour, col_glcm;   // first column of each extension set, or -1
};
cudaError_t launch_ext(const ExtParams& p, cudaStream_t s);

// ---- csv.cu: output assembly -----------------------------------------------
struct CsvParams {
    const float2* centroids;   // [n] all rows of the context
    const float* features;     // [n][F]
    int F;
    int64_t row_lo, rows;      // the rows to format
    int64_t* row_len;          // [rows+1] scratch, entry `rows` must be zero
    int64_t* row_off;          // [rows+1] exclusive scan: byte offset of every row, total at [rows]
    char* text;                // output (k_csv_write)
};
cudaError_t csv_scan_bytes(int64_t rows, size_t* bytes);
cudaError_t launch_csv_measure(const CsvParams& p, void* scan_tmp, size_t scan_bytes, cudaStream_t s);
cudaError_t launch_csv_write(const CsvParams& p, cudaStream_t s);

}  // namespace nfx
