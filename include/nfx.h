/*
 * nfx.h -- C ABI of libnfx.so: the B200 (sm_100a) per-nucleus feature pipeline that replaces the
 * tch/libtorch tensor work of oxabz/nuclei-feature-extraction.
 *
 * Boundary: the library replaces the stage triple
 *     patch_loader -> move_tensors_to_device -> extract_features      (src/main.rs:149-151)
 * i.e. src/utils.rs:141-224 (batch builder + H2D) and the FeatureSet::compute_features_batched
 * implementations of src/features/{shape,color,texture}.rs.  The Rust host keeps the CLI, GeoJSON
 * parsing and the polars writers and binds these symbols through an `extern "C"` block
 * (INTEGRATION.md shows the shim).  Plain pointers and sizes only; no torch types.
 *
 * Threading: one nfx_ctx per (host thread, GPU), like one rayon worker of the reference
 * (src/utils.rs:215-221).  A context is not thread-safe; different contexts are independent.
 * Errors: nothing throws or aborts across the boundary; every call returns NFX_OK (0) or a negative
 * code and nfx_last_error() gives the message (the reference panics instead, src/main.rs:76-89).
 * There is NO CPU fallback: without a usable sm_100 device every compute call fails.
 */
#ifndef NFX_H
#define NFX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NFX_OK               0
#define NFX_ERR_INVALID     -1   /* bad argument */
#define NFX_ERR_CUDA        -2   /* CUDA runtime/driver error (message has the CUDA string) */
#define NFX_ERR_STATE       -3   /* call order: no tile / no polygons / nothing computed yet */
#define NFX_ERR_UNSUPPORTED -4   /* shape outside what the kernels handle (documented per call) */
#define NFX_ERR_NOMEM       -5

/* Feature-set selector bits in args::FeatureSet::flat() order (src/args.rs:35-49). */
#define NFX_FS_GEOMETRY 0x01u    /* ShapeFeatureSet        "geometry"      12 columns (shape.rs:113-128)  */
#define NFX_FS_COLOR    0x02u    /* ColorFeatureSet        "color"         18 columns (color.rs:80-100)   */
#define NFX_FS_GLCM     0x04u    /* GlcmFeatureSet         "GLCM"         224 columns (texture.rs:81-157) */
#define NFX_FS_GLRLM    0x08u    /* GLRLMFeatureSet        "GLRLM"         68 columns (texture.rs:243-301)*/
#define NFX_FS_GABOR    0x10u    /* GaborFilterFeatureSet  "gabor filter"  96 columns (texture.rs:346-361)*/
#define NFX_FS_TEXTURE  (NFX_FS_GLCM | NFX_FS_GLRLM | NFX_FS_GABOR)
#define NFX_FS_ALL      0x1Fu

typedef struct nfx_ctx nfx_ctx;

/* Mirrors the reference's operating-point flags (src/args.rs:93-108). */
typedef struct nfx_config {
    int32_t patch_size;   /* -p/--patch-size, default 64. Any size in [16, 256] (odd sizes included). */
    int32_t batch_size;   /* -b/--batch-size, default 100. mean_h couples the nuclei of one chunk
                             [k*B,(k+1)*B) (src/features/color.rs:50-51, 144-155; src/main.rs:148). */
    int32_t rule_flags;   /* NFX_RULE_* bits below; 0 = the rules of oracle/SPEC.md */
    int32_t reserved[5];  /* must be 0 */
} nfx_config;

/* Switches of the rules whose upstream source (tch-utils@d1c10c0) is not available (oracle/SPEC.md section B, "parity
 * unpinned"): the adopted position is 0, the flag selects the most plausible alternative, in the kernels and -- under the
 * same name -- in the oracle (oracle.RULES). tools/emit_reference_golden.rs + tests/test_reference_golden.py tell which
 * position the reference takes once someone can run it. */
#define NFX_RULE_RASTER_PIXEL_CENTRE 0x1  /* polygon / ellipse rasters sample the pixel CENTRE (c + 0.5 - P/2, r + 0.5 - P/2)
                                             instead of (c - P/2, r - P/2)          (src/utils.rs:152-157, shape.rs:80-87) */
#define NFX_RULE_GABOR_HALF_TURN     0x2  /* Gabor angles i * pi / 8 (48 distinct filters, three oblique angle pairs per frequency) instead of
                                             i * 2 pi / 8 (24 distinct filters, theta and theta + pi coincide)   (texture.rs:333-334) */
#define NFX_RULE_GLCM_254_U8         0x4  /* the 254-level GLCM quantises like an 8-bit image, q = min(floor(g * 255), 253), instead of
                                             min(floor(g * 254), 253) (src/features/texture.rs:19, 40-46). The power-of-two level
                                             counts keep floor(g * L): the kernels use floor(g*32) = floor(g*128) >> 2 */
#define NFX_RULE_WINDOW_SLIDE        0x8  /* the slide path's window (src/utils.rs:96-126) instead of the image path's
                                             (utils.rs:159-192): origin = ((cx - P/2) as u32, (cy - P/2) as u32) -- negative origins
                                             saturate to 0, i.e. the window is SHIFTED, not zero padded, at the left / top edge --
                                             and always P x P (OpenSlide returns black beyond the right / bottom edge). Set
                                             automatically by nfx_slide_load_tiff (.svs / .tif input). */

/* ---- lifetime ------------------------------------------------------------------------------- */
/* Replaces Device::Cuda(gpus[idx]) selection (src/utils.rs:215-221). cfg may be NULL (defaults). */
int nfx_create(int device, const nfx_config* cfg, nfx_ctx** out);
int nfx_destroy(nfx_ctx* ctx);
/* ctx may be NULL: returns the calling thread's last error from nfx_create / schema calls. */
const char* nfx_last_error(const nfx_ctx* ctx);
/* Library/ABI version and the sm arch the kernels were built for ("sm_100a"). */
const char* nfx_version(void);

/* ---- inputs --------------------------------------------------------------------------------- */
/* Stage one slide tile in HBM. Replaces load_input_image + the image Mutex (src/main.rs:20-35,
 * src/utils.rs:146). rgb: HOST pointer, u8 interleaved R,G,B, `h` rows of `row_stride_bytes`
 * (>= 3*w). (origin_x, origin_y) = slide coordinates of the tile's pixel (0,0); polygon coordinates
 * are slide coordinates. Pixels outside the tile read as 0, exactly like the reference's zero
 * padding at image borders (src/utils.rs:174-192). The copy is asynchronous on the context stream
 * when `rgb` is pinned memory. */
int nfx_tile_upload(nfx_ctx* ctx, const uint8_t* rgb, int64_t w, int64_t h,
                    int64_t row_stride_bytes, int64_t origin_x, int64_t origin_y);

/* Slides larger than one upload: reserve a W x H u8 RGB slide in HBM (a 100k x 100k slide is 30 GB of
 * the 180 GB) and stream tiles/bands into it with any number of nfx_slide_write_tile calls, from
 * pinned buffers the caller may reuse after nfx_sync. nfx_tile_upload(rgb,w,h,..) is exactly
 * nfx_slide_alloc(w,h,..) + one nfx_slide_write_tile of the whole image. Regions never written
 * read as undefined data; pixels outside the slide read as 0. */
int nfx_slide_alloc(nfx_ctx* ctx, int64_t w, int64_t h, int64_t origin_x, int64_t origin_y);
/* (x0, y0) = position of the tile's pixel (0,0) inside the slide allocated above. */
int nfx_slide_write_tile(nfx_ctx* ctx, const uint8_t* rgb, int64_t x0, int64_t y0, int64_t w, int64_t h,
                         int64_t row_stride_bytes);

/* ---- one slide on several GPUs (BASELINE config 4) ------------------------------------
 * The reference gives every rayon worker the same slide and moves each batch's patches to
 * gpus[worker % len] (src/utils.rs:211-224); GeoJSON index order is not spatial order, so every GPU's index range
 * touches every tile. Here the slide is resident on EVERY GPU, but each of N GPUs receives only 1/N of it from the
 * host over its own PCIe link and the rest from its peers over NVLink / NVSwitch:
 *   rank r: nfx_slide_alloc(W, H) ; nfx_slide_write_tile(rows of r) ; nfx_sync ; nfx_slide_export(&handle[r])
 *   all   : exchange the handles (any host mechanism) and make sure every rank has synchronised its upload
 *   rank r: for every peer q != r: nfx_slide_import_rows(handle[q], first row of q, rows of q)
 * nfx_slide_import_rows opens the peer's allocation with CUDA IPC (the peer is another PROCESS: one process per GPU)
 * and copies the rows device to device on the context stream. nfx_slide_copy_rows is the same transfer between two
 * contexts of ONE process (one host thread per GPU, like nfx-cli). Both slides must have the same size. */
typedef struct nfx_slide_handle {
    uint8_t ipc[64];      /* cudaIpcMemHandle_t of the slide allocation */
    int64_t width, height, pitch;
    int32_t device;
    int32_t reserved;
} nfx_slide_handle;
int nfx_slide_export(nfx_ctx* ctx, nfx_slide_handle* out);
int nfx_slide_import_rows(nfx_ctx* ctx, const nfx_slide_handle* peer, int64_t y0, int64_t rows);
int nfx_slide_copy_rows(nfx_ctx* dst, nfx_ctx* src, int64_t y0, int64_t rows);

/* Stage n polygons (GeoJSON ring 0 of each feature, closing duplicate included, exactly as
 * `geometry.coordinates[0]` parses to f32: src/geojson.rs:8-24) in CSR form:
 * poly_xy = [poly_off[n]][2] f32 (x,y) slide coordinates, poly_off = [n+1] vertex offsets.
 * Replaces the `&[Feature]` chunk handed to patch_loader (src/input.rs:13-30). */
int nfx_polygons_upload(nfx_ctx* ctx, int64_t n, const float* poly_xy, const int64_t* poly_off);

/* ---- the hot path --------------------------------------------------------------------------- */
/* Launch every kernel for the staged tile + polygons (asynchronous on the context stream):
 * centroid/centring (utils.rs:54-74), mask rasterisation (utils.rs:152-157), patch gather
 * (utils.rs:159-192, fused into the consumers through TMA) and the selected feature sets
 * (shape.rs:16-130, color.rs:10-102, texture.rs:24-168). */
int nfx_compute(nfx_ctx* ctx, uint32_t feature_mask);

/* Wait for nfx_compute and copy results to HOST memory.
 * centroids: [n][2] f32 (the key of utils.rs:226-232 is nfx_centroid_key of each row), may be NULL.
 * features:  [n][nfx_feature_count(mask)] f32 row-major, rows in INPUT order, columns in flat()
 * order (src/main.rs:76-89). May be NULL. */
int nfx_download(nfx_ctx* ctx, float* centroids, float* features);

/* nfx_polygons_upload + nfx_compute + nfx_download in one call: the drop-in for
 * `chunk -> patch_loader -> move_tensors_to_device -> extract_features` (src/main.rs:148-151) over
 * any number of chunks at once (n need not equal batch_size). */
int nfx_extract(nfx_ctx* ctx, int64_t n, const float* poly_xy, const int64_t* poly_off,
                uint32_t feature_mask, float* centroids, float* features);

int nfx_sync(nfx_ctx* ctx);

/* ---- extension outputs ----------------------------------------------------------------------------
 * Quantities BASELINE.json's north_star lists that the REFERENCE does not compute: masked skewness /
 * kurtosis, raw / central / Hu moments of the mask, a contour perimeter, and the GLCM at distance 2 (BASELINE config 3's
 * "32 grey levels, distances 1/2, 4 angles"). They have no reference counterpart -- oracle/SPEC.md section C defines
 * them -- and are NEVER part of the 418-column drop-in schema: their own bit mask, matrix and column names.
 * Needs a staged tile and polygons like nfx_compute; rows in input order, columns in the order of the bits below. */
#define NFX_EXT_COLOR_MOMENTS 0x1u   /* skew_<c>, kurtosis_<c> for c in r g b grey s v haematoxylin eosin dab   (18 columns) */
#define NFX_EXT_MASK_MOMENTS  0x2u   /* m00 m10 m01 m20 m11 m02 m30 m21 m12 m03 mu20 mu11 mu02 mu30 mu21 mu12 mu03 hu1..hu7 (24) */
#define NFX_EXT_CONTOUR       0x4u   /* contour_crack_length, contour_perimeter                                     (2) */
#define NFX_EXT_GLCM_D2       0x8u   /* <haralick>_<dy>_<dx>_32 for (dy,dx) in (0,1) (1,1) (1,0) (1,-1) (0,2) (2,2) (2,0) (2,-2) (112) */
#define NFX_EXT_ALL           0xFu
int nfx_ext_feature_count(uint32_t ext_mask);
const char* nfx_ext_feature_name(uint32_t ext_mask, int idx);
int nfx_compute_ext(nfx_ctx* ctx, uint32_t ext_mask);
int nfx_download_ext(nfx_ctx* ctx, float* out /* [n][nfx_ext_feature_count(ext_mask)] */);

/* ---- trait-level drop-in -------------------------------------------------------------------- */
/* FeatureSet::compute_features_batched(centroids, polygons, patchs, masks) (src/features/mod.rs:
 * 12-28) for ONE batch: patchs [n,3,P,P] f32 in [0,1], masks [n,1,P,P] f32 (non-zero = inside), polygons = CENTRED
 * rings (utils.rs:65-72) in CSR. A batch built by the reference's own loader holds values k/255 (utils.rs:172) and runs
 * on the u8 kernels; any other f32 values are accepted too: the texture sets then read an f32 grey plane and the colour
 * set is evaluated from the f32 patches (csrc/f32batch.cu; written for correctness, not speed).
 * `feature_set` is one NFX_FS_* bit (the trait call), or a union of bits: the batch is then uploaded
 * once and the columns of the sets follow each other in flat() order (src/args.rs:38-45), which saves
 * the 65 KB per nucleus of PCIe traffic each further per-set call would repeat. All pointers are HOST
 * pointers. out: [n][nfx_feature_count(feature_set)] f32. The whole call is one chunk for mean_h. */
int nfx_compute_features_batched(nfx_ctx* ctx, uint32_t feature_set, int64_t n,
                                 const float* centroids, const float* poly_xy,
                                 const int64_t* poly_off, const float* patchs, const float* masks,
                                 float* out);

/* ---- staged kernels (north-star kernels 1 and 2 on their own) and parity taps ---------------- */
/* Kernel (1): batched patch gather of the staged polygons' windows from the staged tile into a
 * device array of u8 patches [n][P][3P] (interleaved RGB), utils.rs:159-192. If out != NULL the
 * patches are copied to the HOST array out[n][P][P][3]. */
int nfx_gather_patches(nfx_ctx* ctx, uint8_t* out);
/* Kernel (2): polygon -> mask rasterisation only (utils.rs:152-157). If out != NULL the masks are
 * copied to HOST as u8 0/1 out[n][P][P]. */
int nfx_rasterize(nfx_ctx* ctx, uint8_t* out);
/* Ellipse masks drawn by the last geometry run (shape.rs:80-87), HOST u8 out[n][P][P]. */
int nfx_debug_ellipses(nfx_ctx* ctx, uint8_t* out);
/* Symmetric GLCM COUNTS (C + C^T, before normalisation) of the staged inputs for one
 * (levels, offset) pair, HOST uint32 out[n][levels][levels] (texture.rs:40-46). */
int nfx_debug_glcm_counts(nfx_ctx* ctx, int levels, int dy, int dx, uint32_t* out);
/* Quantised grey levels min(floor(grey*levels), levels-1), HOST u8 out[n][P][P] (texture.rs:36). */
int nfx_debug_grey_levels(nfx_ctx* ctx, int levels, uint8_t* out);

/* ---- schema (column names of src/features/{shape,color,texture}.rs; set names of src/args.rs) -- */
int nfx_feature_count(uint32_t feature_mask);
/* Column `idx` (0-based, WITHOUT the leading "centroid" key column) for the given mask, or NULL. */
const char* nfx_feature_name(uint32_t feature_mask, int idx);
/* args::FeatureSet::from_str (src/args.rs:18-32): case-insensitive geometry|color|glcm|glrlm|
 * gabor|texture|all -> bits. Returns NFX_ERR_INVALID for anything else. */
int nfx_parse_feature_set(const char* name, uint32_t* bits);
/* FeatureSet::name() of one bit (shape.rs:132-134, color.rs:104-106, texture.rs:169-171,312-314,
 * 371-373): "geometry", "color", "GLCM", "GLRLM", "gabor filter". */
const char* nfx_feature_set_name(uint32_t bit);
/* centroid_to_key_string (src/utils.rs:226-228): Rust `Display` of both f32, comma separated.
 * Returns the string length (excluding NUL) or NFX_ERR_INVALID if buf is too small. */
int nfx_centroid_key(float x, float y, char* buf, int buflen);

/* Rust `Display` of one f32 (shortest round-trip decimal, never an exponent): the formatting polars uses
 * for the key and the CSV writer. Returns the length or NFX_ERR_INVALID if buf is too small. */
int nfx_format_f32(float v, char* buf, int buflen);

/* ---- slide decode -------------------------------------------------------- */
/* The reference opens `.svs` slides with OpenSlide and reads one region per nucleus (src/utils.rs:79-139,
 * src/main.rs:20-35). An Aperio .svs is a TIFF / BigTIFF whose directory 0 is the full-resolution image in
 * JPEG-compressed tiles sharing one JPEGTables blob. nfx_slide_load_tiff parses that container (baseline TIFF
 * 6.0 + BigTIFF, either byte order, tiles or strips, 8-bit 3-sample, compression 7), allocates the slide
 * (origin 0,0) and decodes every block with nvJPEG straight into HBM on `threads` host threads (<= 0: all
 * cores); only compressed bytes cross PCIe. Photometric = RGB blocks are taken as R,G,B components (no colour
 * transform, as libtiff / OpenSlide do), Photometric = YCbCr blocks are converted. Real .svs files could not
 * be tested in this environment; the container and decode paths are tested on synthetic
 * files against libjpeg-turbo. `file` is the whole file in host memory (e.g. an mmap). */
typedef struct nfx_tiff_level {
    int64_t width, height;
    int32_t block_width, block_height;   /* tile size, or (width, rows per strip) */
    int64_t blocks;
    int32_t compression, photometric;    /* TIFF tags 259, 262 */
    int32_t jpeg_tables_bytes;           /* tag 347 */
} nfx_tiff_level;
int nfx_tiff_info(const uint8_t* file, int64_t len, nfx_tiff_level* out);   /* host only: no GPU needed */
int nfx_slide_load_tiff(nfx_ctx* ctx, const uint8_t* file, int64_t len, int32_t threads);   /* = _ex with flags 0 */
/* Decoder choice. Default (flags = 0): every JPEG block is decoded on the host threads by csrc/jpeg_exact.cpp, a restatement
 * of the IJG / libjpeg-turbo default path (integer "islow" IDCT, triangle-filter chroma upsampling, fixed-point YCbCr -> RGB)
 * whose pixels equal libjpeg's bit for bit -- the pixels OpenSlide hands the reference -- and uploaded from pinned staging;
 * streams outside its scope (progressive, 1x2 sampling, several scans) go through nvJPEG. NFX_DECODE_FAST: nvJPEG for every
 * block (only compressed bytes cross PCIe, about twice the tile rate, pixels within a few grey levels of libjpeg's). */
#define NFX_DECODE_FAST 0x1u
int nfx_slide_load_tiff_ex(nfx_ctx* ctx, const uint8_t* file, int64_t len, int32_t threads, uint32_t flags);
/* The exact decoder on its own (host only: no GPU needed): one baseline JPEG stream -> interleaved u8 RGB.
 * colourspace: 0 = the components are R,G,B (TIFF Photometric = RGB), 1 = YCbCr, -1 = libjpeg's rule (JFIF / Adobe marker /
 * component ids). rgb may be NULL to query the size; capacity in bytes. NFX_ERR_UNSUPPORTED for streams outside its scope. */
int nfx_jpeg_decode(const uint8_t* data, int64_t len, int32_t colourspace, uint8_t* rgb, int64_t capacity, int32_t* width, int32_t* height);
/* parity tap: a region of the resident slide back to the host, interleaved u8 RGB [h][w][3] */
int nfx_debug_slide_read(nfx_ctx* ctx, int64_t x0, int64_t y0, int64_t w, int64_t h, uint8_t* rgb);

/* ---- GeoJSON -> CSR polygon packing ------------------------------------- */
/* Replaces `serde_json::from_reader::<FeatureCollection>` (src/main.rs:37-42) for the model of
 * src/geojson.rs:8-24 (`features[].bbox` required, `features[].geometry.{type,coordinates}`, unknown
 * keys ignored, duplicate or missing fields rejected) and keeps ring 0 of every feature, which is all
 * preprocess_polygon reads (src/utils.rs:54-60). The text is scanned and parsed by `threads` host
 * threads (<= 0: all cores); the result is the CSR the other entry points take. Numbers are
 * deserialised like serde_json 1.0.107 without `float_roundtrip` does for an f32 field (u64
 * significand, one power-of-ten multiply/divide in f64, then `as f32`), so the coordinates and hence
 * the centroid keys are the reference's bit for bit. On error: NFX_ERR_INVALID, message with line and
 * column through nfx_last_error(NULL). */
typedef struct nfx_geojson nfx_geojson;
int nfx_geojson_parse(const char* text, int64_t len, int32_t threads, nfx_geojson** out);
int64_t nfx_geojson_count(const nfx_geojson* g);          /* features */
int64_t nfx_geojson_vertices(const nfx_geojson* g);       /* stored vertices of all rings 0 */
const float* nfx_geojson_xy(const nfx_geojson* g);        /* [vertices][2] f32 */
const int64_t* nfx_geojson_offsets(const nfx_geojson* g); /* [features+1] */
const float* nfx_geojson_bbox(const nfx_geojson* g);      /* [features][4], NaN where the array is shorter */
const int32_t* nfx_geojson_rings(const nfx_geojson* g);   /* [features] ring count (only ring 0 is kept) */
void nfx_geojson_free(nfx_geojson* g);
/* One JSON number token deserialised as an f32 field (the rule above); NFX_ERR_INVALID if it is not a number. */
int nfx_parse_f32(const char* token, int32_t len, float* out);

/* ---- output assembly ----------------------------------------------------- */
/* The text polars' CsvWriter writes for the reference's output DataFrame (src/main.rs:76-89 hstack behind
 * the `centroid` key of src/utils.rs:226-232; writer at src/main.rs:163-166): a header line, then per
 * nucleus `"cx,cy",f0,...,f{F-1}\n`, every f32 in Rust `Display` form (shortest round-trip digits, no
 * exponent, NaN / inf / -inf). Cells are formatted ON THE GPU from the resident result of nfx_compute
 * (k_csv_measure, scan, k_csv_write) and only the text crosses PCIe; call it per row range to stream a
 * large table through a fixed host buffer. *len receives the byte count (also when NFX_ERR_INVALID
 * reports that cap is too small; nothing is written then). `out` is a host pointer (pinned is faster). */
int nfx_csv_header(uint32_t feature_mask, char* out, int64_t cap, int64_t* len);
int nfx_csv_rows(nfx_ctx* ctx, int64_t row_lo, int64_t row_hi, char* out, int64_t cap, int64_t* len);
/* Same formatter for a caller-held DataFrame (host pointers: centroids [n][2], features [n][cols] f32), e.g.
 * the frames returned by nfx_compute_features_batched after the host's hstack. */
int nfx_csv_format(nfx_ctx* ctx, int64_t n, int32_t cols, const float* centroids, const float* features,
                   char* out, int64_t cap, int64_t* len);

/* ---- multi-GPU partition ------------------------------------------------------ */
/* Contiguous index ranges, one per part, boundaries rounded to multiples of batch_size so that
 * every reference chunk [k*B,(k+1)*B) (src/main.rs:148) lives on one GPU. bounds: [parts+1]. */
int nfx_partition(int64_t n, int32_t batch_size, int32_t parts, int64_t* bounds);

/* ---- measurement ----------------------------------------------------------------------------- */
typedef struct nfx_kernel_time {
    char    name[48];
    int64_t launches;
    double  total_ms;      /* sum of CUDA-event durations of this kernel on the context stream */
} nfx_kernel_time;
/* When enabled every kernel launch is bracketed by CUDA events on the context stream. */
int nfx_profile_enable(nfx_ctx* ctx, int enable);
int nfx_profile_reset(nfx_ctx* ctx);
/* Synchronises, then fills up to `max` entries; returns the number of distinct kernels. */
int nfx_profile_get(nfx_ctx* ctx, nfx_kernel_time* out, int max);
/* Whole-region timer on the context stream (CUDA events). */
int nfx_timer_start(nfx_ctx* ctx);
int nfx_timer_stop(nfx_ctx* ctx, float* elapsed_ms);   /* synchronises on the stop event */
/* Number of kernel launches issued by this context since creation. */
int64_t nfx_launch_count(const nfx_ctx* ctx);
/* Write `bytes` of device memory (> L2) to evict the L2 between timed iterations. */
int nfx_flush_l2(nfx_ctx* ctx);
/* Pinned host allocations for H2D/D2H staging by the caller. */
int nfx_host_alloc(void** p, int64_t bytes);
int nfx_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* NFX_H */
