// api.cu -- the C ABI of include/nfx.h: context, HBM staging, TMA descriptors, launch sequencing,
// measurement. No torch, no CPU fallback: every compute entry point fails if CUDA fails.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges are no-ops unless a profiler is attached
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/nfx.h"
#include "nfx_host.h"
#include "nfx_kernels.h"

using namespace nfx;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    cudaError_t ensure(size_t need) {
        if (need <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = need + need / 8 + 64;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct ProfRec {
    cudaEvent_t a, b;
    int kid;
};

}  // namespace

struct nfx_ctx {
    int device = 0;
    int P = 64, B = 100;
    uint32_t rules = 0;   // NFX_RULE_* (nfx_config.rule_flags; NFX_RULE_WINDOW_SLIDE is also set by nfx_slide_load_tiff)
    cudaStream_t stream = nullptr;
    std::string err;
    EncodeTiledFn encode = nullptr;

    // tile
    DevBuf<uint8_t> tile;
    int64_t tw = 0, th = 0, tpitch = 0, tox = 0, toy = 0;
    bool have_tile = false;
    CUtensorMap map_tile_patch, map_tile_slab, map_tile_cslab, map_tile_gabor;

    std::vector<std::pair<std::string, void*>> peers;   // CUDA IPC mappings of peer slides (nfx_slide_import_rows)

    // polygons
    DevBuf<float2> xy;
    DevBuf<int64_t> off;
    int64_t n = 0, nverts = 0;
    int vmax = 0;
    DevBuf<int> giant_list;            // nuclei whose ring is longer than kGeomRingSmem vertices: they work in HBM (geom.cu)
    DevBuf<unsigned char> ring_scratch;
    int64_t n_giant = 0;
    bool have_poly = false;

    // per nucleus
    DevBuf<float2> centroid;
    DevBuf<NucInfo> info;
    DevBuf<uint32_t> bitmask;
    DevBuf<float> out;
    DevBuf<float> ext_out;
    uint32_t ext_mask = 0;
    DevBuf<float> hue;
    DevBuf<uint32_t> ellipse;
    DevBuf<double> gabor_part;
    uint32_t computed_mask = 0;
    int out_cols = 0;
    bool have_geom = false;   // centroid/info/bitmask valid for the staged polygons

    // staged patch array (kernel (1) output / trait-level input)
    DevBuf<uint8_t> patches;
    int64_t ppitch = 0;
    CUtensorMap map_pat_patch, map_pat_slab, map_pat_cslab, map_pat_gabor;

    // scratch
    DevBuf<uint8_t> scratch8;
    DevBuf<float> scratchf;
    const float* grey_f32 = nullptr;   // set while a trait-level batch of arbitrary f32 patches is computed (f32batch.cu)
    DevBuf<uint32_t> scratch32;
    DevBuf<uint8_t> flush;
    int* d_bad = nullptr;

    // output assembly (csv.cu)
    DevBuf<int64_t> csv_len, csv_off;
    DevBuf<uint8_t> csv_tmp;
    DevBuf<char> csv_text;
    DevBuf<float> csv_in;

    // measurement
    bool profile = false;
    std::vector<ProfRec> recs;
    std::vector<std::string> knames;
    int64_t launches = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
};

namespace {

int fail(nfx_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    g_thread_error = msg;
    return code;
}
int cuda_fail(nfx_ctx* c, cudaError_t e, const char* what) {
    return fail(c, NFX_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                                     \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call);   \
    } while (0)

int kernel_id(nfx_ctx* c, const char* name) {
    for (size_t k = 0; k < c->knames.size(); ++k)
        if (c->knames[k] == name) return (int)k;
    c->knames.push_back(name);
    return (int)c->knames.size() - 1;
}

// Launch wrapper: counts launches and, when profiling, brackets the launch with CUDA events on the
// context stream (the stream the kernel runs on).
template <typename F>
cudaError_t timed(nfx_ctx* c, const char* name, int nlaunch, F&& f) {
    ProfRec r;
    const bool prof = c->profile;
    if (prof) {
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        r.kid = kernel_id(c, name);
        cudaEventRecord(r.a, c->stream);
    }
    nvtxRangePushA(name);   // the reference logs wall-clock per feature set (src/main.rs:55-70); here every kernel is a range
    cudaError_t e = f();
    nvtxRangePop();
    c->launches += nlaunch;
    if (prof) {
        cudaEventRecord(r.b, c->stream);
        c->recs.push_back(r);
    }
    return e;
}

// 2D u8 tensor map: dim0 = bytes per row (width_bytes), dim1 = rows, box = {192 B, box_h rows}.
int make_map(nfx_ctx* ctx, CUtensorMap* m, void* base, int64_t width_bytes, int64_t rows, int64_t pitch,
             int box_h) {
    const cuuint64_t gdim[2] = {(cuuint64_t)width_bytes, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)pitch};
    const cuuint32_t box[2] = {(cuuint32_t)kPanelBytes, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = ctx->encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[160];
        snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed (CUresult %d; w=%lld rows=%lld pitch=%lld box_h=%d)",
                 (int)r, (long long)width_bytes, (long long)rows, (long long)pitch, box_h);
        return fail(ctx, NFX_ERR_CUDA, b);
    }
    return NFX_OK;
}

int set_device(nfx_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    return NFX_OK;
}

struct Cols {
    int shape, color, glcm, glrlm, gabor, total;
};
Cols columns(uint32_t mask) {
    Cols c;
    c.shape = column_offset(mask, NFX_FS_GEOMETRY);
    c.color = column_offset(mask, NFX_FS_COLOR);
    c.glcm = column_offset(mask, NFX_FS_GLCM);
    c.glrlm = column_offset(mask, NFX_FS_GLRLM);
    c.gabor = column_offset(mask, NFX_FS_GABOR);
    c.total = nfx_feature_count(mask);
    return c;
}

int check_patch_size(nfx_ctx* ctx, uint32_t mask) {
    if (mask == 0 || (mask & ~NFX_FS_ALL)) return fail(ctx, NFX_ERR_INVALID, "empty or unknown feature mask");
    return NFX_OK;
}

// Geometry pass over the staged polygons: centroid, window, bitmask (+ shape columns).
int run_geom(nfx_ctx* ctx, bool shape, float* out, int stride, int col, uint32_t* ellipse_bits) {
    const int64_t n = ctx->n;
    const int wpr = mask_wpr(ctx->P);
    CK(ctx->centroid.ensure(n));
    CK(ctx->info.ensure(n));
    CK(ctx->bitmask.ensure((size_t)n * ctx->P * wpr));
    GeomParams g;
    g.poly_xy = ctx->xy.p;
    g.poly_off = ctx->off.p;
    g.n = n;
    g.P = ctx->P;
    g.tile_ox = (int)ctx->tox;
    g.tile_oy = (int)ctx->toy;
    g.vmax = std::max(ctx->vmax, 1);
    g.vsmem = std::min(g.vmax, kGeomRingSmem);
    g.n_giant = ctx->n_giant;
    g.giant_list = ctx->n_giant ? ctx->giant_list.p : nullptr;
    g.ring_scratch = ctx->n_giant ? ctx->ring_scratch.p : nullptr;
    g.centroid = ctx->centroid.p;
    g.info = ctx->info.p;
    g.bitmask = ctx->bitmask.p;
    g.out = out;
    g.out_stride = stride;
    g.col_shape = col;
    g.ellipse_bits = ellipse_bits;
    g.sample_off = (ctx->rules & NFX_RULE_RASTER_PIXEL_CENTRE) ? 0.5f : 0.0f;
    g.slide_window = (ctx->rules & NFX_RULE_WINDOW_SLIDE) ? 1 : 0;
    CK(timed(ctx, shape ? "k_geom<raster,shape>" : "k_geom<raster>", 1,
             [&] { return launch_geom(g, true, shape, ctx->stream); }));
    ctx->have_geom = true;
    return NFX_OK;
}

int run_color(nfx_ctx* ctx, int64_t n, int batch, const CUtensorMap* mp, const CUtensorMap* ms, float* out,
              int stride, int col) {
    const int R = hue_slab_rows(ctx->P), slabs = (ctx->P + R - 1) / R;
    CK(ctx->hue.ensure((size_t)n * slabs * 2));
    ColorParams c;
    c.n = n;
    c.P = ctx->P;
    c.batch_size = batch;
    c.info = ctx->info.p;
    c.bitmask = ctx->bitmask.p;
    c.out = out;
    c.out_stride = stride;
    c.col_color = col;
    c.hue_partial = ctx->hue.p;
    c.slabs = slabs;
    c.slab_rows = color_slab_rows(ctx->P);
    // P == 64: one warp per nucleus over 16-row sub-slabs (the slab map); other sizes: one CTA per nucleus
    if (ctx->P == 64) CK(timed(ctx, "k_color_warp", 1, [&] { return launch_color_warp(c, ms, ctx->stream); }));
    else CK(timed(ctx, "k_color", 1, [&] { return launch_color(c, mp, ctx->stream); }));
    CK(timed(ctx, "k_hue_batch", 1, [&] { return launch_hue_batch(c, ms, R, ctx->stream); }));
    CK(timed(ctx, "k_hue_finalize", 1, [&] { return launch_hue_finalize(c, ctx->stream); }));
    return NFX_OK;
}

int run_glcm(nfx_ctx* ctx, int64_t n, const CUtensorMap* mp, float* out, int stride, int col,
             uint32_t* dbg_counts, int dl, int ddy, int ddx, uint8_t* dbg_grey) {
    GlcmParams g;
    g.n = n;
    g.P = ctx->P;
    g.info = ctx->info.p;
    g.bitmask = ctx->bitmask.p;
    g.out = out;
    g.out_stride = stride;
    g.col_glcm = col;
    g.dbg_counts = dbg_counts;
    g.dbg_levels = dl;
    g.dbg_dy = ddy;
    g.dbg_dx = ddx;
    g.dbg_grey = dbg_grey;
    g.scale254 = (ctx->rules & NFX_RULE_GLCM_254_U8) ? 255.0f : 254.0f;
    g.grey = ctx->grey_f32;
    CK(timed(ctx, "k_glcm", 1, [&] { return launch_glcm(g, mp, ctx->stream); }));
    return NFX_OK;
}

int run_tex2(nfx_ctx* ctx, int64_t n, uint32_t mask, const CUtensorMap* map_cslab, const CUtensorMap* map_patch,
             const CUtensorMap* map_gabor, float* out, int stride, int col_glrlm, int col_gabor) {
    TexParams t;
    t.n = n;
    t.P = ctx->P;
    t.slab_rows = color_slab_rows(ctx->P);
    t.info = ctx->info.p;
    t.bitmask = ctx->bitmask.p;
    t.out = out;
    t.out_stride = stride;
    t.col_glrlm = col_glrlm;
    t.col_gabor = col_gabor;
    t.gabor_partial = nullptr;
    t.gabor_half_turn = (ctx->rules & NFX_RULE_GABOR_HALF_TURN) ? 1 : 0;
    t.grey = ctx->grey_f32;
    const bool gabor_tiled = gabor_tiles(ctx->P) > 1;
    if ((mask & NFX_FS_GABOR) && gabor_tiled) {
        CK(ctx->gabor_part.ensure((size_t)n * gabor_tiles(ctx->P) * 97));
        t.gabor_partial = ctx->gabor_part.p;
    }
    if (mask & NFX_FS_GLRLM) CK(timed(ctx, "k_glrlm", 1, [&] { return launch_glrlm(t, map_cslab, ctx->stream); }));
    if (mask & NFX_FS_GABOR)
        CK(timed(ctx, "k_gabor", gabor_tiled ? 2 : 1, [&] { return launch_gabor(t, gabor_tiled ? map_gabor : map_patch, ctx->stream); }));
    return NFX_OK;
}

int need_inputs(nfx_ctx* ctx, bool tile) {
    if (!ctx) return NFX_ERR_INVALID;
    if (tile && !ctx->have_tile) return fail(ctx, NFX_ERR_STATE, "no tile staged: call nfx_tile_upload first");
    if (!ctx->have_poly) return fail(ctx, NFX_ERR_STATE, "no polygons staged: call nfx_polygons_upload first");
    return NFX_OK;
}

int make_patch_array(nfx_ctx* ctx, int64_t n) {
    const int P = ctx->P;
    // rows carry 16 bytes of padding so that the 208-byte TMA box of the last panel stays inside the tensor
    ctx->ppitch = ((3 * P + 15) / 16) * 16 + 16;
    CK(ctx->patches.ensure((size_t)n * P * ctx->ppitch));
    const int R = hue_slab_rows(P);
    int rc;
    if ((rc = make_map(ctx, &ctx->map_pat_patch, ctx->patches.p, ctx->ppitch, n * P, ctx->ppitch, P))) return rc;
    if ((rc = make_map(ctx, &ctx->map_pat_slab, ctx->patches.p, ctx->ppitch, n * P, ctx->ppitch, R))) return rc;
    if ((rc = make_map(ctx, &ctx->map_pat_cslab, ctx->patches.p, ctx->ppitch, n * P, ctx->ppitch, color_slab_rows(P)))) return rc;
    if ((rc = make_map(ctx, &ctx->map_pat_gabor, ctx->patches.p, ctx->ppitch, n * P, ctx->ppitch, gabor_fetch_rows()))) return rc;
    return NFX_OK;
}

}  // namespace

extern "C" {

const char* nfx_last_error(const nfx_ctx* ctx) { return ctx ? ctx->err.c_str() : g_thread_error.c_str(); }

int nfx_create(int device, const nfx_config* cfg, nfx_ctx** out) {
    if (!out) return fail(nullptr, NFX_ERR_INVALID, "nfx_create: out is NULL");
    *out = nullptr;
    int P = 64, B = 100;
    uint32_t rules = 0;
    if (cfg) {
        if (cfg->patch_size) P = cfg->patch_size;
        if (cfg->batch_size) B = cfg->batch_size;
        rules = (uint32_t)cfg->rule_flags;
        for (int k = 0; k < 5; ++k)
            if (cfg->reserved[k]) return fail(nullptr, NFX_ERR_INVALID, "nfx_config.reserved must be 0");
    }
    if (rules & ~(uint32_t)(NFX_RULE_RASTER_PIXEL_CENTRE | NFX_RULE_GABOR_HALF_TURN | NFX_RULE_GLCM_254_U8 | NFX_RULE_WINDOW_SLIDE))
        return fail(nullptr, NFX_ERR_INVALID, "unknown nfx_config.rule_flags bit");
    if (P < 16 || P > 256) return fail(nullptr, NFX_ERR_UNSUPPORTED, "patch_size must be in [16,256]");
    if (B < 1) return fail(nullptr, NFX_ERR_INVALID, "batch_size must be >= 1");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceCount (no CUDA device: there is no CPU fallback)");
    if (device < 0 || device >= count) return fail(nullptr, NFX_ERR_INVALID, "GPU " + std::to_string(device) + " does not exist");   // args.rs:176-180
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10) return fail(nullptr, NFX_ERR_UNSUPPORTED, std::string("kernels are built for sm_100a only; device is ") + prop.name);
    nfx_ctx* ctx = new nfx_ctx();
    ctx->device = device;
    ctx->P = P;
    ctx->B = B;
    ctx->rules = rules;
    auto bail = [&](int rc) { nfx_destroy(ctx); return rc; };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(cuda_fail(nullptr, e, "cudaSetDevice"));
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(cuda_fail(nullptr, e, "cudaStreamCreate"));
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess) return bail(fail(nullptr, NFX_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver"));
    ctx->encode = (EncodeTiledFn)fn;
    if ((e = cudaEventCreate(&ctx->t0)) != cudaSuccess || (e = cudaEventCreate(&ctx->t1)) != cudaSuccess) return bail(cuda_fail(nullptr, e, "cudaEventCreate"));
    if ((e = cudaMalloc((void**)&ctx->d_bad, sizeof(int))) != cudaSuccess) return bail(cuda_fail(nullptr, e, "cudaMalloc"));
    *out = ctx;
    return NFX_OK;
}

int nfx_destroy(nfx_ctx* ctx) {
    if (!ctx) return NFX_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& r : ctx->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto& pr : ctx->peers) cudaIpcCloseMemHandle(pr.second);
    ctx->peers.clear();
    ctx->tile.release(); ctx->xy.release(); ctx->off.release(); ctx->centroid.release(); ctx->info.release();
    ctx->giant_list.release(); ctx->ring_scratch.release(); ctx->bitmask.release(); ctx->out.release(); ctx->ext_out.release(); ctx->hue.release(); ctx->ellipse.release(); ctx->gabor_part.release(); ctx->patches.release();
    ctx->scratch8.release(); ctx->scratchf.release(); ctx->scratch32.release(); ctx->flush.release();
    ctx->csv_len.release(); ctx->csv_off.release(); ctx->csv_tmp.release(); ctx->csv_text.release(); ctx->csv_in.release();
    if (ctx->d_bad) cudaFree(ctx->d_bad);
    if (ctx->t0) cudaEventDestroy(ctx->t0);
    if (ctx->t1) cudaEventDestroy(ctx->t1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return NFX_OK;
}

int nfx_slide_alloc(nfx_ctx* ctx, int64_t w, int64_t h, int64_t origin_x, int64_t origin_y) {
    if (!ctx) return NFX_ERR_INVALID;
    if (w <= 0 || h <= 0) return fail(ctx, NFX_ERR_INVALID, "nfx_slide_alloc: bad size");
    if (3 * w >= (1ll << 31) || h >= (1ll << 31)) return fail(ctx, NFX_ERR_UNSUPPORTED, "slide too large for 32-bit TMA coordinates");
    int rc = set_device(ctx);
    if (rc) return rc;
    const int64_t pitch = ((3 * w + 15) / 16) * 16;
    ctx->have_tile = false;
    CK(ctx->tile.ensure((size_t)pitch * h + 256));
    ctx->tw = w; ctx->th = h; ctx->tpitch = pitch; ctx->tox = origin_x; ctx->toy = origin_y;
    if ((rc = make_map(ctx, &ctx->map_tile_patch, ctx->tile.p, 3 * w, h, pitch, ctx->P))) return rc;
    if ((rc = make_map(ctx, &ctx->map_tile_slab, ctx->tile.p, 3 * w, h, pitch, hue_slab_rows(ctx->P)))) return rc;
    if ((rc = make_map(ctx, &ctx->map_tile_cslab, ctx->tile.p, 3 * w, h, pitch, color_slab_rows(ctx->P)))) return rc;
    if ((rc = make_map(ctx, &ctx->map_tile_gabor, ctx->tile.p, 3 * w, h, pitch, gabor_fetch_rows()))) return rc;
    ctx->have_tile = true;
    ctx->have_geom = false;   // window origins depend on the slide origin
    return NFX_OK;
}

int nfx_slide_write_tile(nfx_ctx* ctx, const uint8_t* rgb, int64_t x0, int64_t y0, int64_t w, int64_t h,
                         int64_t row_stride_bytes) {
    if (!ctx) return NFX_ERR_INVALID;
    if (!ctx->have_tile) return fail(ctx, NFX_ERR_STATE, "no slide allocated: call nfx_slide_alloc first");
    if (!rgb || w <= 0 || h <= 0 || row_stride_bytes < 3 * w || x0 < 0 || y0 < 0 || x0 + w > ctx->tw || y0 + h > ctx->th)
        return fail(ctx, NFX_ERR_INVALID, "nfx_slide_write_tile: tile does not fit the slide");
    int rc = set_device(ctx);
    if (rc) return rc;
    CK(cudaMemcpy2DAsync(ctx->tile.p + (size_t)y0 * ctx->tpitch + 3 * x0, ctx->tpitch, rgb, row_stride_bytes, 3 * w, h,
                         cudaMemcpyHostToDevice, ctx->stream));
    return NFX_OK;
}

int nfx_tile_upload(nfx_ctx* ctx, const uint8_t* rgb, int64_t w, int64_t h, int64_t row_stride_bytes,
                    int64_t origin_x, int64_t origin_y) {
    if (!ctx) return NFX_ERR_INVALID;
    if (!rgb || w <= 0 || h <= 0 || row_stride_bytes < 3 * w) return fail(ctx, NFX_ERR_INVALID, "nfx_tile_upload: bad arguments");
    int rc = nfx_slide_alloc(ctx, w, h, origin_x, origin_y);
    if (rc) return rc;
    return nfx_slide_write_tile(ctx, rgb, 0, 0, w, h, row_stride_bytes);
}

int nfx_slide_export(nfx_ctx* ctx, nfx_slide_handle* out) {
    if (!ctx || !out) return NFX_ERR_INVALID;
    if (!ctx->have_tile) return fail(ctx, NFX_ERR_STATE, "no slide allocated: call nfx_slide_alloc first");
    int rc = set_device(ctx);
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->tile.p));
    static_assert(sizeof(h) == sizeof(out->ipc), "cudaIpcMemHandle_t is 64 bytes");
    memcpy(out->ipc, &h, sizeof h);
    out->width = ctx->tw; out->height = ctx->th; out->pitch = ctx->tpitch; out->device = ctx->device;
    return NFX_OK;
}

static int check_rows(nfx_ctx* ctx, int64_t w, int64_t h, int64_t pitch, int64_t y0, int64_t rows) {
    if (!ctx->have_tile) return fail(ctx, NFX_ERR_STATE, "no slide allocated: call nfx_slide_alloc first");
    if (w != ctx->tw || h != ctx->th || pitch != ctx->tpitch) return fail(ctx, NFX_ERR_INVALID, "peer slide has a different size");
    if (y0 < 0 || rows < 0 || y0 + rows > ctx->th) return fail(ctx, NFX_ERR_INVALID, "rows outside the slide");
    return NFX_OK;
}

int nfx_slide_import_rows(nfx_ctx* ctx, const nfx_slide_handle* peer, int64_t y0, int64_t rows) {
    if (!ctx || !peer) return NFX_ERR_INVALID;
    int rc = check_rows(ctx, peer->width, peer->height, peer->pitch, y0, rows);
    if (rc) return rc;
    if ((rc = set_device(ctx))) return rc;
    if (rows == 0) return NFX_OK;
    const std::string key(reinterpret_cast<const char*>(peer->ipc), sizeof peer->ipc);
    void* base = nullptr;
    for (auto& pr : ctx->peers)
        if (pr.first == key) base = pr.second;
    if (!base) {
        cudaIpcMemHandle_t h;
        memcpy(&h, peer->ipc, sizeof h);
        CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peers.emplace_back(key, base);
    }
    // same pitch on both sides: the rows are one contiguous range (NVLink / NVSwitch when the peer is another GPU)
    CK(cudaMemcpyAsync(ctx->tile.p + (size_t)y0 * ctx->tpitch, static_cast<const uint8_t*>(base) + (size_t)y0 * ctx->tpitch,
                       (size_t)rows * ctx->tpitch, cudaMemcpyDeviceToDevice, ctx->stream));
    return NFX_OK;
}

int nfx_slide_copy_rows(nfx_ctx* ctx, nfx_ctx* src, int64_t y0, int64_t rows) {
    if (!ctx || !src) return NFX_ERR_INVALID;
    if (!src->have_tile) return fail(ctx, NFX_ERR_STATE, "source context has no slide");
    int rc = check_rows(ctx, src->tw, src->th, src->tpitch, y0, rows);
    if (rc) return rc;
    if ((rc = set_device(ctx))) return rc;
    if (rows == 0) return NFX_OK;
    CK(cudaMemcpyPeerAsync(ctx->tile.p + (size_t)y0 * ctx->tpitch, ctx->device, src->tile.p + (size_t)y0 * src->tpitch, src->device,
                           (size_t)rows * ctx->tpitch, ctx->stream));
    return NFX_OK;
}

int nfx_slide_load_tiff(nfx_ctx* ctx, const uint8_t* file, int64_t len, int32_t threads) {
    return nfx_slide_load_tiff_ex(ctx, file, len, threads, 0);
}

int nfx_jpeg_decode(const uint8_t* data, int64_t len, int32_t colourspace, uint8_t* rgb, int64_t capacity, int32_t* width, int32_t* height) {
    if (!data || len <= 0 || !width || !height || colourspace < -1 || colourspace > 1) return fail(nullptr, NFX_ERR_INVALID, "nfx_jpeg_decode: bad arguments");
    std::vector<uint8_t> out;
    int w = 0, h = 0;
    std::string err;
    if (!jpeg_decode_exact(data, (size_t)len, colourspace, out, w, h, err)) return fail(nullptr, NFX_ERR_UNSUPPORTED, "jpeg: " + err);
    *width = w;
    *height = h;
    if (rgb) {
        if ((int64_t)out.size() > capacity) return fail(nullptr, NFX_ERR_INVALID, "nfx_jpeg_decode: output buffer too small");
        memcpy(rgb, out.data(), out.size());
    }
    return NFX_OK;
}

int nfx_slide_load_tiff_ex(nfx_ctx* ctx, const uint8_t* file, int64_t len, int32_t threads, uint32_t flags) {
    if (!ctx) return NFX_ERR_INVALID;
    if (flags & ~(uint32_t)NFX_DECODE_FAST) return fail(ctx, NFX_ERR_INVALID, "unknown nfx_slide_load_tiff_ex flag");
    TiffLevel L;
    std::string err;
    if (!tiff_parse(file, len, L, err)) return fail(ctx, NFX_ERR_INVALID, "tiff: " + err);
    int rc = nfx_slide_alloc(ctx, L.width, L.height, 0, 0);
    if (rc) return rc;
    CK(cudaMemsetAsync(ctx->tile.p, 0, (size_t)ctx->tpitch * L.height, ctx->stream));   // sparse blocks stay black
    CK(cudaStreamSynchronize(ctx->stream));
    err = decode_tiff_level(file, L, ctx->tile.p, ctx->tpitch, ctx->device, threads, (flags & NFX_DECODE_FAST) != 0);
    if (!err.empty()) { ctx->have_tile = false; return fail(ctx, NFX_ERR_UNSUPPORTED, "tiff: " + err); }
    ctx->rules |= NFX_RULE_WINDOW_SLIDE;   // .svs / .tif input takes the reference's slide path (src/utils.rs:96-126)
    ctx->have_geom = false;
    return NFX_OK;
}

int nfx_debug_slide_read(nfx_ctx* ctx, int64_t x0, int64_t y0, int64_t w, int64_t h, uint8_t* rgb) {
    if (!ctx) return NFX_ERR_INVALID;
    if (!ctx->have_tile) return fail(ctx, NFX_ERR_STATE, "no slide resident");
    if (!rgb || w <= 0 || h <= 0 || x0 < 0 || y0 < 0 || x0 + w > ctx->tw || y0 + h > ctx->th)
        return fail(ctx, NFX_ERR_INVALID, "nfx_debug_slide_read: region outside the slide");
    int rc = set_device(ctx);
    if (rc) return rc;
    CK(cudaMemcpy2DAsync(rgb, (size_t)3 * w, ctx->tile.p + (size_t)y0 * ctx->tpitch + 3 * x0, ctx->tpitch, (size_t)3 * w, h,
                         cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return NFX_OK;
}

int nfx_polygons_upload(nfx_ctx* ctx, int64_t n, const float* poly_xy, const int64_t* poly_off) {
    if (!ctx) return NFX_ERR_INVALID;
    if (n < 0 || (n > 0 && (!poly_xy || !poly_off))) return fail(ctx, NFX_ERR_INVALID, "nfx_polygons_upload: bad arguments");
    int rc = set_device(ctx);
    if (rc) return rc;
    int64_t total = 0;
    int vmax = 0;
    if (n > 0) {
        if (poly_off[0] != 0) return fail(ctx, NFX_ERR_INVALID, "poly_off[0] must be 0");
        for (int64_t i = 0; i < n; ++i) {
            const int64_t v = poly_off[i + 1] - poly_off[i];
            if (v < 0) return fail(ctx, NFX_ERR_INVALID, "poly_off must be non-decreasing");
            if (v > (1 << 20)) return fail(ctx, NFX_ERR_UNSUPPORTED, "ring longer than 2^20 vertices");
            vmax = std::max<int64_t>(vmax, v);
        }
        total = poly_off[n];
    }
    // rings that do not fit the shared-memory layout of k_geom get a work slot in HBM
    ctx->n_giant = 0;
    if (vmax > kGeomRingSmem) {
        std::vector<int> list;
        for (int64_t i = 0; i < n; ++i)
            if (poly_off[i + 1] - poly_off[i] > kGeomRingSmem) list.push_back((int)i);
        ctx->n_giant = (int64_t)list.size();
        CK(ctx->giant_list.ensure(list.size()));
        CK(ctx->ring_scratch.ensure((size_t)ctx->n_giant * geom_ring_bytes(vmax)));
        CK(cudaMemcpyAsync(ctx->giant_list.p, list.data(), list.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));   // `list` is a local
    }
    CK(ctx->xy.ensure((size_t)total + 1));
    CK(ctx->off.ensure((size_t)n + 1));
    if (n > 0) {
        CK(cudaMemcpyAsync(ctx->xy.p, poly_xy, (size_t)total * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->off.p, poly_off, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->n = n; ctx->nverts = total; ctx->vmax = vmax;
    ctx->have_poly = true;
    ctx->have_geom = false;
    ctx->computed_mask = 0;
    ctx->ext_mask = 0;
    return NFX_OK;
}

int nfx_compute(nfx_ctx* ctx, uint32_t mask) {
    int rc = need_inputs(ctx, true);
    if (rc) return rc;
    if ((rc = check_patch_size(ctx, mask))) return rc;
    if ((rc = set_device(ctx))) return rc;
    const Cols c = columns(mask);
    const int64_t n = ctx->n;
    CK(ctx->out.ensure((size_t)std::max<int64_t>(n, 1) * c.total));
    ctx->out_cols = c.total;
    ctx->computed_mask = 0;
    if (n == 0) { ctx->computed_mask = mask; return NFX_OK; }
    if ((rc = run_geom(ctx, (mask & NFX_FS_GEOMETRY) != 0, ctx->out.p, c.total, c.shape, nullptr))) return rc;
    if (mask & NFX_FS_COLOR)
        if ((rc = run_color(ctx, n, ctx->B, &ctx->map_tile_cslab, &ctx->map_tile_slab, ctx->out.p, c.total, c.color))) return rc;
    if (mask & NFX_FS_GLCM)
        if ((rc = run_glcm(ctx, n, glcm_uses_slab_map(ctx->P) ? &ctx->map_tile_cslab : &ctx->map_tile_patch, ctx->out.p, c.total, c.glcm, nullptr, 0, 0, 0, nullptr))) return rc;
    if (mask & (NFX_FS_GLRLM | NFX_FS_GABOR))
        if ((rc = run_tex2(ctx, n, mask, &ctx->map_tile_cslab, &ctx->map_tile_patch, &ctx->map_tile_gabor, ctx->out.p, c.total, c.glrlm, c.gabor))) return rc;
    ctx->computed_mask = mask;
    return NFX_OK;
}

// ---- extension outputs (ext.cu) ---------------------------------------------------------------------
namespace {
const char* const kExtChan[9] = {"r", "g", "b", "grey", "s", "v", "haematoxylin", "eosin", "dab"};
const char* const kExtMask[24] = {"m00", "m10", "m01", "m20", "m11", "m02", "m30", "m21", "m12", "m03", "mu20", "mu11", "mu02", "mu30",
                                  "mu21", "mu12", "mu03", "hu1", "hu2", "hu3", "hu4", "hu5", "hu6", "hu7"};
const char* const kExtHaralick[14] = {"correlation", "contrast", "dissimilarity", "entropy", "angular_second_moment", "sum_average",
                                      "sum_variance", "sum_entropy", "sum_of_squares", "inverse_difference_moment", "difference_average",
                                      "difference_variance", "information_measure_correlation1", "information_measure_correlation2"};
const int kExtOff[8][2] = {{0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, 2}, {2, 2}, {2, 0}, {2, -2}};
thread_local std::string g_ext_name;
}  // namespace

int nfx_ext_feature_count(uint32_t m) {
    if (m & ~NFX_EXT_ALL) return -1;
    return ((m & NFX_EXT_COLOR_MOMENTS) ? kExtColorCols : 0) + ((m & NFX_EXT_MASK_MOMENTS) ? kExtMaskCols : 0) +
           ((m & NFX_EXT_CONTOUR) ? kExtContourCols : 0) + ((m & NFX_EXT_GLCM_D2) ? kExtGlcmCols : 0);
}

const char* nfx_ext_feature_name(uint32_t m, int idx) {
    if (idx < 0 || idx >= nfx_ext_feature_count(m)) return nullptr;
    if (m & NFX_EXT_COLOR_MOMENTS) {
        if (idx < kExtColorCols) { g_ext_name = std::string((idx & 1) ? "kurtosis_" : "skew_") + kExtChan[idx / 2]; return g_ext_name.c_str(); }
        idx -= kExtColorCols;
    }
    if (m & NFX_EXT_MASK_MOMENTS) {
        if (idx < kExtMaskCols) return kExtMask[idx];
        idx -= kExtMaskCols;
    }
    if (m & NFX_EXT_CONTOUR) {
        if (idx < kExtContourCols) return idx == 0 ? "contour_crack_length" : "contour_perimeter";
        idx -= kExtContourCols;
    }
    const int o = idx / 14;
    g_ext_name = std::string(kExtHaralick[idx % 14]) + "_" + std::to_string(kExtOff[o][0]) + "_" + std::to_string(kExtOff[o][1]) + "_32";
    return g_ext_name.c_str();
}

int nfx_compute_ext(nfx_ctx* ctx, uint32_t m) {
    int rc = need_inputs(ctx, true);
    if (rc) return rc;
    if (m == 0 || (m & ~NFX_EXT_ALL)) return fail(ctx, NFX_ERR_INVALID, "empty or unknown extension mask");
    if ((rc = set_device(ctx))) return rc;
    const int cols = nfx_ext_feature_count(m);
    const int64_t n = ctx->n;
    CK(ctx->ext_out.ensure((size_t)std::max<int64_t>(n, 1) * cols));
    ctx->ext_mask = 0;
    if (n == 0) { ctx->ext_mask = m; return NFX_OK; }
    if (!ctx->have_geom)
        if ((rc = run_geom(ctx, false, nullptr, 0, -1, nullptr))) return rc;
    ExtParams e;
    e.n = n; e.P = ctx->P; e.info = ctx->info.p; e.bitmask = ctx->bitmask.p;
    e.tile = ctx->tile.p; e.tpitch = ctx->tpitch; e.tw = ctx->tw; e.th = ctx->th;
    e.out = ctx->ext_out.p; e.out_stride = cols;
    int c = 0;
    e.col_color = (m & NFX_EXT_COLOR_MOMENTS) ? c : -1;  c += (m & NFX_EXT_COLOR_MOMENTS) ? kExtColorCols : 0;
    e.col_mask = (m & NFX_EXT_MASK_MOMENTS) ? c : -1;    c += (m & NFX_EXT_MASK_MOMENTS) ? kExtMaskCols : 0;
    e.col_contour = (m & NFX_EXT_CONTOUR) ? c : -1;      c += (m & NFX_EXT_CONTOUR) ? kExtContourCols : 0;
    e.col_glcm = (m & NFX_EXT_GLCM_D2) ? c : -1;
    CK(timed(ctx, "k_ext", 2, [&] { return launch_ext(e, ctx->stream); }));
    ctx->ext_mask = m;
    return NFX_OK;
}

int nfx_download_ext(nfx_ctx* ctx, float* out) {
    if (!ctx) return NFX_ERR_INVALID;
    if (!ctx->ext_mask) return fail(ctx, NFX_ERR_STATE, "nothing computed: call nfx_compute_ext first");
    int rc = set_device(ctx);
    if (rc) return rc;
    if (ctx->n > 0 && out)
        CK(cudaMemcpyAsync(out, ctx->ext_out.p, (size_t)ctx->n * nfx_ext_feature_count(ctx->ext_mask) * sizeof(float),
                           cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return NFX_OK;
}

int nfx_sync(nfx_ctx* ctx) {
    if (!ctx) return NFX_ERR_INVALID;
    int rc = set_device(ctx);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return NFX_OK;
}

int nfx_download(nfx_ctx* ctx, float* centroids, float* features) {
    if (!ctx) return NFX_ERR_INVALID;
    if (!ctx->computed_mask) return fail(ctx, NFX_ERR_STATE, "nothing computed: call nfx_compute first");
    int rc = set_device(ctx);
    if (rc) return rc;
    if (ctx->n > 0) {
        if (centroids) CK(cudaMemcpyAsync(centroids, ctx->centroid.p, (size_t)ctx->n * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
        if (features) CK(cudaMemcpyAsync(features, ctx->out.p, (size_t)ctx->n * ctx->out_cols * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return NFX_OK;
}

// ---- output assembly ----------------------------------------------------------
namespace {
int csv_format_device(nfx_ctx* ctx, const float2* d_centroids, const float* d_features, int F, int64_t row_lo,
                      int64_t rows, char* out, int64_t cap, int64_t* len) {
    *len = 0;
    if (rows == 0) return NFX_OK;
    CK(ctx->csv_len.ensure((size_t)rows + 1));
    CK(ctx->csv_off.ensure((size_t)rows + 1));
    size_t tmp = 0;
    CK(csv_scan_bytes(rows, &tmp));
    CK(ctx->csv_tmp.ensure(tmp));
    CsvParams p{d_centroids, d_features, F, row_lo, rows, ctx->csv_len.p, ctx->csv_off.p, nullptr};
    CK(cudaMemsetAsync(ctx->csv_len.p + rows, 0, sizeof(int64_t), ctx->stream));
    CK(timed(ctx, "k_csv_measure", 1, [&] { return launch_csv_measure(p, ctx->csv_tmp.p, tmp, ctx->stream); }));
    ctx->launches += 2;   // cub's two scan kernels
    int64_t total = 0;
    CK(cudaMemcpyAsync(&total, ctx->csv_off.p + rows, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *len = total;
    if (total > cap || !out) return fail(ctx, NFX_ERR_INVALID, "csv: output buffer too small, need " + std::to_string(total) + " bytes");
    CK(ctx->csv_text.ensure((size_t)total));
    p.text = ctx->csv_text.p;
    CK(timed(ctx, "k_csv_write", 1, [&] { return launch_csv_write(p, ctx->stream); }));
    CK(cudaMemcpyAsync(out, ctx->csv_text.p, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return NFX_OK;
}
}  // namespace

int nfx_csv_rows(nfx_ctx* ctx, int64_t row_lo, int64_t row_hi, char* out, int64_t cap, int64_t* len) {
    if (!ctx || !len) return NFX_ERR_INVALID;
    if (!ctx->computed_mask) return fail(ctx, NFX_ERR_STATE, "nothing computed: call nfx_compute first");
    if (row_lo < 0 || row_hi < row_lo || row_hi > ctx->n) return fail(ctx, NFX_ERR_INVALID, "nfx_csv_rows: bad row range");
    int rc = set_device(ctx);
    if (rc) return rc;
    return csv_format_device(ctx, ctx->centroid.p, ctx->out.p, ctx->out_cols, row_lo, row_hi - row_lo, out, cap, len);
}

int nfx_csv_format(nfx_ctx* ctx, int64_t n, int32_t cols, const float* centroids, const float* features, char* out,
                   int64_t cap, int64_t* len) {
    if (!ctx || !len || n < 0 || cols < 0 || (n > 0 && (!centroids || (cols > 0 && !features))))
        return ctx ? fail(ctx, NFX_ERR_INVALID, "nfx_csv_format: bad arguments") : NFX_ERR_INVALID;
    int rc = set_device(ctx);
    if (rc) return rc;
    *len = 0;
    if (n == 0) return NFX_OK;
    const size_t nc = 2 * (size_t)n, nf = (size_t)n * (size_t)cols;
    CK(ctx->csv_in.ensure(nc + nf + 4));
    CK(cudaMemcpyAsync(ctx->csv_in.p, centroids, nc * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (nf) CK(cudaMemcpyAsync(ctx->csv_in.p + nc, features, nf * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    return csv_format_device(ctx, reinterpret_cast<const float2*>(ctx->csv_in.p), ctx->csv_in.p + nc, cols, 0, n, out, cap, len);
}

int nfx_extract(nfx_ctx* ctx, int64_t n, const float* poly_xy, const int64_t* poly_off, uint32_t mask,
                float* centroids, float* features) {
    int rc = nfx_polygons_upload(ctx, n, poly_xy, poly_off);
    if (rc) return rc;
    if ((rc = nfx_compute(ctx, mask))) return rc;
    return nfx_download(ctx, centroids, features);
}

int nfx_compute_features_batched(nfx_ctx* ctx, uint32_t fs, int64_t n, const float* centroids,
                                 const float* poly_xy, const int64_t* poly_off, const float* patchs,
                                 const float* masks, float* out) {
    if (!ctx) return NFX_ERR_INVALID;
    (void)centroids;   // only used for the key column, which the host formats (utils.rs:226-232)
    if (fs == 0 || (fs & ~(uint32_t)NFX_FS_ALL))
        return fail(ctx, NFX_ERR_INVALID, "feature_set must be one NFX_FS_* bit (the trait call) or a union of them (one upload, columns in flat() order)");
    int rc = check_patch_size(ctx, fs);
    if (rc) return rc;
    if (n < 0 || (n > 0 && (!patchs || !masks || !out))) return fail(ctx, NFX_ERR_INVALID, "nfx_compute_features_batched: bad arguments");
    if (n == 0) return NFX_OK;
    if ((rc = set_device(ctx))) return rc;
    // the asserts of shape.rs:23-47 / color.rs:18-42 become argument checks on the CSR
    if (fs & NFX_FS_GEOMETRY)
        if ((rc = nfx_polygons_upload(ctx, n, poly_xy, poly_off))) return rc;   // centred rings
    const int P = ctx->P, wpr = mask_wpr(P);
    const size_t plane = (size_t)P * P;
    CK(ctx->scratchf.ensure((size_t)n * 4 * plane));
    float* d_patchs = ctx->scratchf.p;
    float* d_masks = d_patchs + (size_t)n * 3 * plane;
    CK(cudaMemcpyAsync(d_patchs, patchs, (size_t)n * 3 * plane * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_masks, masks, (size_t)n * plane * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = make_patch_array(ctx, n))) return rc;
    CK(ctx->info.ensure(n));
    CK(ctx->centroid.ensure(n));
    CK(ctx->bitmask.ensure((size_t)n * P * wpr));
    CK(cudaMemsetAsync(ctx->d_bad, 0, sizeof(int), ctx->stream));
    CK(timed(ctx, "k_pack_batch", 1, [&] {
        return launch_pack_batch(n, P, d_patchs, d_masks, ctx->patches.p, ctx->ppitch, ctx->bitmask.p,
                                 ctx->info.p, ctx->d_bad, ctx->stream);
    }));
    // Values that are not k/255 (no loader of the reference produces them, utils.rs:172, but the trait accepts any tensor):
    // the texture sets then read an f32 grey plane and the colour set is evaluated from the f32 patches (f32batch.cu).
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, ctx->d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const bool f32path = bad != 0;
    const Cols c = columns(fs);
    const int cols = c.total;
    CK(ctx->out.ensure((size_t)n * cols));
    if (fs & NFX_FS_GEOMETRY) {
        GeomParams g;
        g.poly_xy = ctx->xy.p; g.poly_off = ctx->off.p; g.n = n; g.P = P; g.tile_ox = 0; g.tile_oy = 0;
        g.vmax = std::max(ctx->vmax, 1); g.vsmem = std::min(g.vmax, kGeomRingSmem);
        g.n_giant = ctx->n_giant; g.giant_list = ctx->n_giant ? ctx->giant_list.p : nullptr; g.ring_scratch = ctx->n_giant ? ctx->ring_scratch.p : nullptr;
        g.centroid = ctx->centroid.p; g.info = ctx->info.p;
        g.bitmask = ctx->bitmask.p; g.out = ctx->out.p; g.out_stride = cols; g.col_shape = c.shape; g.ellipse_bits = nullptr;
        g.sample_off = (ctx->rules & NFX_RULE_RASTER_PIXEL_CENTRE) ? 0.5f : 0.0f;
        g.slide_window = 0;
        CK(timed(ctx, "k_geom<shape>", 1, [&] { return launch_geom(g, false, true, ctx->stream); }));
    }
    const int batch = (int)std::min<int64_t>(n, 1 << 30);   // the trait call IS one batch
    if (f32path) {
        if (fs & (NFX_FS_GLCM | NFX_FS_GLRLM | NFX_FS_GABOR)) {
            // the f32 masks are packed into the bitmask by now: their region of the scratch buffer takes the grey planes
            CK(timed(ctx, "k_grey_f32", 1, [&] { return launch_grey_f32(n, P, d_patchs, d_masks, ctx->stream); }));
            ctx->grey_f32 = d_masks;
        }
        if (fs & NFX_FS_COLOR) {
            const int64_t chunks = (n + batch - 1) / batch;
            CK(ctx->hue.ensure((size_t)chunks * plane * 2));
            CK(timed(ctx, "k_color_f32", 3, [&] {
                return launch_color_f32(n, P, batch, d_patchs, ctx->bitmask.p, ctx->hue.p, ctx->out.p, cols, c.color, ctx->stream);
            }));
        }
    } else if (fs & NFX_FS_COLOR) {
        if ((rc = run_color(ctx, n, batch, &ctx->map_pat_cslab, &ctx->map_pat_slab, ctx->out.p, cols, c.color))) return rc;
    }
    if (fs & NFX_FS_GLCM)
        rc = run_glcm(ctx, n, glcm_uses_slab_map(ctx->P) ? &ctx->map_pat_cslab : &ctx->map_pat_patch, ctx->out.p, cols, c.glcm, nullptr, 0, 0, 0, nullptr);
    if (!rc && (fs & (NFX_FS_GLRLM | NFX_FS_GABOR)))
        rc = run_tex2(ctx, n, fs, &ctx->map_pat_cslab, &ctx->map_pat_patch, &ctx->map_pat_gabor, ctx->out.p, cols, c.glrlm, c.gabor);
    ctx->grey_f32 = nullptr;
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, ctx->out.p, (size_t)n * cols * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_geom = false;
    ctx->have_poly = false;   // the staged rings were centred ones: force a fresh upload for nfx_compute
    ctx->computed_mask = 0;
    return NFX_OK;
}

int nfx_rasterize(nfx_ctx* ctx, uint8_t* out) {
    int rc = need_inputs(ctx, false);
    if (rc) return rc;
    if ((rc = set_device(ctx))) return rc;
    if (ctx->n == 0) return NFX_OK;
    if ((rc = run_geom(ctx, false, nullptr, 0, -1, nullptr))) return rc;
    if (out) {
        const size_t total = (size_t)ctx->n * ctx->P * ctx->P;
        CK(ctx->scratch8.ensure(total));
        CK(timed(ctx, "k_expand", 1, [&] { return launch_expand_mask(ctx->n, ctx->P, ctx->bitmask.p, ctx->scratch8.p, ctx->stream); }));
        CK(cudaMemcpyAsync(out, ctx->scratch8.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return NFX_OK;
}

int nfx_gather_patches(nfx_ctx* ctx, uint8_t* out) {
    int rc = need_inputs(ctx, true);
    if (rc) return rc;
    if ((rc = set_device(ctx))) return rc;
    if (ctx->n == 0) return NFX_OK;
    if (!ctx->have_geom)
        if ((rc = run_geom(ctx, false, nullptr, 0, -1, nullptr))) return rc;
    if ((rc = make_patch_array(ctx, ctx->n))) return rc;
    CK(timed(ctx, "k_gather", 1, [&] {
        return launch_gather(ctx->n, ctx->P, ctx->info.p, &ctx->map_tile_patch, ctx->patches.p, ctx->ppitch, ctx->stream);
    }));
    if (out)
        CK(cudaMemcpy2DAsync(out, (size_t)3 * ctx->P, ctx->patches.p, ctx->ppitch, (size_t)3 * ctx->P,
                             (size_t)ctx->n * ctx->P, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return NFX_OK;
}

int nfx_debug_ellipses(nfx_ctx* ctx, uint8_t* out) {
    int rc = need_inputs(ctx, false);
    if (rc) return rc;
    if (!out) return fail(ctx, NFX_ERR_INVALID, "out is NULL");
    if ((rc = set_device(ctx))) return rc;
    if (ctx->n == 0) return NFX_OK;
    const int wpr = mask_wpr(ctx->P);
    CK(ctx->ellipse.ensure((size_t)ctx->n * ctx->P * wpr));
    CK(ctx->scratchf.ensure((size_t)ctx->n * kShapeCols));
    if ((rc = run_geom(ctx, true, ctx->scratchf.p, kShapeCols, 0, ctx->ellipse.p))) return rc;
    const size_t total = (size_t)ctx->n * ctx->P * ctx->P;
    CK(ctx->scratch8.ensure(total));
    CK(timed(ctx, "k_expand", 1, [&] { return launch_expand_mask(ctx->n, ctx->P, ctx->ellipse.p, ctx->scratch8.p, ctx->stream); }));
    CK(cudaMemcpyAsync(out, ctx->scratch8.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return NFX_OK;
}

static int glcm_debug(nfx_ctx* ctx, int levels, int dy, int dx, uint32_t* counts, uint8_t* grey) {
    int rc = need_inputs(ctx, true);
    if (rc) return rc;
    if (levels != 32 && levels != 64 && levels != 128 && levels != 254) return fail(ctx, NFX_ERR_INVALID, "levels must be one of 32, 64, 128, 254 (texture.rs:19)");
    if ((rc = check_patch_size(ctx, NFX_FS_GLCM))) return rc;
    if ((rc = set_device(ctx))) return rc;
    if (ctx->n == 0) return NFX_OK;
    if (!ctx->have_geom)
        if ((rc = run_geom(ctx, false, nullptr, 0, -1, nullptr))) return rc;
    uint32_t* d_counts = nullptr;
    uint8_t* d_grey = nullptr;
    if (counts) {
        const size_t words = (size_t)ctx->n * levels * levels;
        CK(ctx->scratch32.ensure(words));
        CK(cudaMemsetAsync(ctx->scratch32.p, 0, words * 4, ctx->stream));
        d_counts = ctx->scratch32.p;
    }
    if (grey) {
        CK(ctx->scratch8.ensure((size_t)ctx->n * ctx->P * ctx->P));
        d_grey = ctx->scratch8.p;
    }
    if ((rc = run_glcm(ctx, ctx->n, glcm_uses_slab_map(ctx->P) ? &ctx->map_tile_cslab : &ctx->map_tile_patch, nullptr, 0, 0, d_counts, levels, dy, dx, d_grey))) return rc;
    if (counts) CK(cudaMemcpyAsync(counts, d_counts, (size_t)ctx->n * levels * levels * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (grey) CK(cudaMemcpyAsync(grey, d_grey, (size_t)ctx->n * ctx->P * ctx->P, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return NFX_OK;
}

int nfx_debug_glcm_counts(nfx_ctx* ctx, int levels, int dy, int dx, uint32_t* out) {
    if (!out) return fail(ctx, NFX_ERR_INVALID, "out is NULL");
    return glcm_debug(ctx, levels, dy, dx, out, nullptr);
}
int nfx_debug_grey_levels(nfx_ctx* ctx, int levels, uint8_t* out) {
    if (!out) return fail(ctx, NFX_ERR_INVALID, "out is NULL");
    return glcm_debug(ctx, levels, 0, 1, nullptr, out);
}

// ---- measurement ------------------------------------------------------------------------------
int nfx_profile_enable(nfx_ctx* ctx, int enable) {
    if (!ctx) return NFX_ERR_INVALID;
    ctx->profile = enable != 0;
    return NFX_OK;
}
int nfx_profile_reset(nfx_ctx* ctx) {
    if (!ctx) return NFX_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& r : ctx->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    ctx->recs.clear();
    return NFX_OK;
}
int nfx_profile_get(nfx_ctx* ctx, nfx_kernel_time* out, int max) {
    if (!ctx) return NFX_ERR_INVALID;
    int rc = set_device(ctx);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<nfx_kernel_time> acc(ctx->knames.size());
    for (size_t k = 0; k < acc.size(); ++k) {
        memset(&acc[k], 0, sizeof acc[k]);
        snprintf(acc[k].name, sizeof acc[k].name, "%s", ctx->knames[k].c_str());
    }
    for (auto& r : ctx->recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            acc[r.kid].launches += 1;
            acc[r.kid].total_ms += ms;
        }
    }
    int cnt = 0;
    for (auto& a : acc) {
        if (!a.launches) continue;
        if (out && cnt < max) out[cnt] = a;
        ++cnt;
    }
    return cnt;
}
int nfx_timer_start(nfx_ctx* ctx) {
    if (!ctx) return NFX_ERR_INVALID;
    int rc = set_device(ctx);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->t0, ctx->stream));
    return NFX_OK;
}
int nfx_timer_stop(nfx_ctx* ctx, float* ms) {
    if (!ctx || !ms) return NFX_ERR_INVALID;
    int rc = set_device(ctx);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->t1, ctx->stream));
    CK(cudaEventSynchronize(ctx->t1));
    CK(cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return NFX_OK;
}
int64_t nfx_launch_count(const nfx_ctx* ctx) { return ctx ? ctx->launches : 0; }

int nfx_flush_l2(nfx_ctx* ctx) {
    if (!ctx) return NFX_ERR_INVALID;
    int rc = set_device(ctx);
    if (rc) return rc;
    const size_t bytes = 256ull << 20;   // > 126 MB L2
    CK(ctx->flush.ensure(bytes));
    CK(cudaMemsetAsync(ctx->flush.p, 0x5a, bytes, ctx->stream));
    return NFX_OK;
}

int nfx_host_alloc(void** p, int64_t bytes) {
    if (!p || bytes < 0) return NFX_ERR_INVALID;
    cudaError_t e = cudaHostAlloc(p, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaHostAlloc");
    return NFX_OK;
}
int nfx_host_free(void* p) {
    if (!p) return NFX_OK;
    cudaError_t e = cudaFreeHost(p);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaFreeHost");
    return NFX_OK;
}

}  // extern "C"
