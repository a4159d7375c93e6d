// nfx_host.hpp -- C++ host layer above the C ABI (include/nfx.h), mirroring the reference's Rust host
// for the hot path: the `FeatureSet` trait and its five implementors (src/features/mod.rs:12-28,
// shape.rs:132-134, color.rs:104-106, texture.rs:169-171, 312-314, 371-373), args::FeatureSet
// (src/args.rs:7-73), the GeoJSON model (src/geojson.rs:8-24), the image loader (src/main.rs:20-35),
// the batch pipeline (src/main.rs:146-158) and the writers (src/main.rs:160-189).
// The reference's toolchain (rustc) is not in the build image, so the compiled host mirror is C++.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "nfx.h"

namespace nfxhost {

using Point = std::array<float, 2>;              // utils.rs:10  CratePoint
using Points = std::vector<Point>;               // utils.rs:12

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// f32, contiguous, row-major: the stand-in for tch::Tensor at the trait boundary
struct Tensor {
    std::vector<int64_t> size;
    std::vector<float> data;
};

// polars DataFrame subset: one Utf8 key column + f32 feature columns
struct DataFrame {
    std::vector<std::string> centroid;           // key column (utils.rs:226-232)
    std::vector<float> centroid_xy;              // the same centroids as numbers, [row][2] (for the device CSV formatter)
    std::vector<std::string> names;
    std::vector<std::vector<float>> columns;     // columns[j][row]
    size_t height() const { return centroid.size(); }
    void hstack(const DataFrame& other);         // main.rs:76-89 (asserts equal centroid columns)
    void vstack(const DataFrame& other);         // main.rs:96-108
};

// ---- args.rs:7-49 ---------------------------------------------------------------------------
enum class FeatureSetKind { Geometry, Color, Glcm, Glrlm, Gabor, Texture, All };
FeatureSetKind feature_set_from_str(const std::string& s);          // throws Error "{} is not a valid feature set"
std::vector<FeatureSetKind> flat(const std::vector<FeatureSetKind>& s);
uint32_t feature_mask(const std::vector<FeatureSetKind>& s);        // flat() as NFX_FS_* bits; throws on duplicates

// ---- one context per (host thread, GPU), utils.rs:215-221 ------------------------------------
class Context {
public:
    Context(int device, int patch_size, int batch_size);
    ~Context();
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    nfx_ctx* raw() const { return ctx_; }
    int patch_size() const { return patch_size_; }
    void check(int rc) const;                    // non-zero -> Error(nfx_last_error)
private:
    nfx_ctx* ctx_ = nullptr;
    int patch_size_;
};

// ---- features/mod.rs:12-28 --------------------------------------------------------------------
class FeatureSet {
public:
    virtual ~FeatureSet() = default;
    virtual const char* name() const = 0;
    virtual DataFrame compute_features_batched(const std::vector<Point>& centroids, const std::vector<Points>& polygons,
                                               const Tensor& patchs, const Tensor& masks) const;
protected:
    FeatureSet(Context& c, uint32_t bit) : ctx_(c), bit_(bit) {}
    Context& ctx_;
    uint32_t bit_;
};
struct ShapeFeatureSet : FeatureSet { explicit ShapeFeatureSet(Context& c) : FeatureSet(c, NFX_FS_GEOMETRY) {} const char* name() const override { return "geometry"; } };
struct ColorFeatureSet : FeatureSet { explicit ColorFeatureSet(Context& c) : FeatureSet(c, NFX_FS_COLOR) {} const char* name() const override { return "color"; } };
struct GlcmFeatureSet : FeatureSet { explicit GlcmFeatureSet(Context& c) : FeatureSet(c, NFX_FS_GLCM) {} const char* name() const override { return "GLCM"; } };
struct GLRLMFeatureSet : FeatureSet { explicit GLRLMFeatureSet(Context& c) : FeatureSet(c, NFX_FS_GLRLM) {} const char* name() const override { return "GLRLM"; } };
struct GaborFilterFeatureSet : FeatureSet { explicit GaborFilterFeatureSet(Context& c) : FeatureSet(c, NFX_FS_GABOR) {} const char* name() const override { return "gabor filter"; } };
std::vector<std::unique_ptr<FeatureSet>> to_fs(const std::vector<FeatureSetKind>& s, Context& ctx);   // args.rs:51-73

// ---- geojson.rs:8-24, main.rs:37-42 ------------------------------------------------------------
// The collection is kept as the CSR nfx_geojson_parse produces (ring 0 of every feature; the reference keeps
// all rings but reads only coordinates[0], utils.rs:55).
struct FeatureCollection {
    std::vector<float> xy;        // [vertices][2], f32 as serde parses them
    std::vector<int64_t> off;     // [features+1]
    std::vector<float> bbox;      // [features][4] (required field of Feature, unused downstream)
    size_t size() const { return off.empty() ? 0 : off.size() - 1; }
    const float* ring(size_t i) const { return xy.data() + 2 * off[i]; }
    size_t ring_len(size_t i) const { return (size_t)(off[i + 1] - off[i]); }
};
FeatureCollection load_geometry(const std::string& path, int threads = 0);   // multi-threaded, nfx_geojson_parse

// ---- main.rs:20-35 ------------------------------------------------------------------------------
struct Image {
    int64_t w = 0, h = 0;
    std::vector<uint8_t> rgb;                    // [h][w][3] (png / ppm)
    std::vector<uint8_t> tiff;                   // or: the bytes of a JPEG-compressed TIFF / .svs, decoded on the GPU (nvJPEG)
};
Image load_input_image(const std::string& path);   // png (8-bit, non-interlaced), binary ppm, svs / tif / tiff (JPEG blocks); jpg -> Error
void upload_image(Context& ctx, const Image& image);   // tile upload, or nvJPEG decode straight into the resident slide

// ---- args.rs:76-183 -----------------------------------------------------------------------------
struct Args {
    std::string geometry, slide, output;
    std::vector<FeatureSetKind> feature_sets;
    bool overwrite = false, verbose = false, via_trait = false, host_csv = false;
    int patch_size = 64;
    int thread_count = 0;
    std::vector<int> gpus;
    int batch_size = 100;
};
Args parse_args(int argc, char** argv);            // throws Error with a usage message
std::string validate_paths(const Args& a);         // returns the output extension; throws Error (exit 1 messages of args.rs)

// ---- main.rs:146-158 ----------------------------------------------------------------------------
// Contiguous index ranges per GPU (nfx_partition), one host thread + context each, merged in input order.
DataFrame extract(const FeatureCollection& geometry, const Image& image, const Args& args);
// The csv arm of main.rs:163-166 without a DataFrame in between: every GPU formats the rows of its own range from the
// result that is still resident (nfx_csv_rows) and only text comes back; the ranges are written in input order.
void extract_to_csv(const FeatureCollection& geometry, const Image& image, const Args& args, const std::string& path);
// Same result through the trait objects: masks and patches come back from kernels (1) and (2) as the
// reference's Batch tensors (utils.rs:17) and go through FeatureSet::compute_features_batched chunk by chunk.
DataFrame extract_via_trait(const FeatureCollection& geometry, const Image& image, const Args& args);

// ---- main.rs:160-189 ----------------------------------------------------------------------------
// csv (cells formatted on GPU `device` when >= 0, else on the host: same bytes), json; others -> Error
void write_output(const std::string& path, const std::string& ext, const DataFrame& df, int device = -1);

}  // namespace nfxhost
