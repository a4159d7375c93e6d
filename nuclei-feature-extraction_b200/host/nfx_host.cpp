// nfx_host.cpp -- implementation of the C++ host mirror (see nfx_host.hpp for the reference citations).
#include "nfx_host.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>

namespace nfxhost {

// ------------------------------------------------------------------------------------------------
// DataFrame
// ------------------------------------------------------------------------------------------------
void DataFrame::hstack(const DataFrame& o) {
    if (names.empty() && columns.empty() && centroid.empty()) { *this = o; return; }
    if (o.centroid != centroid) throw Error("centroid columns differ between feature sets");   // main.rs:76-79 assert
    for (size_t j = 0; j < o.names.size(); ++j) {
        if (std::find(names.begin(), names.end(), o.names[j]) != names.end())
            throw Error("duplicate column '" + o.names[j] + "'");                                 // DataFrame::new fails, main.rs:89
        names.push_back(o.names[j]);
        columns.push_back(o.columns[j]);
    }
}
void DataFrame::vstack(const DataFrame& o) {
    if (names.empty() && centroid.empty()) { *this = o; return; }
    if (o.names != names) throw Error("vstack: schemas differ");
    centroid.insert(centroid.end(), o.centroid.begin(), o.centroid.end());
    centroid_xy.insert(centroid_xy.end(), o.centroid_xy.begin(), o.centroid_xy.end());
    for (size_t j = 0; j < columns.size(); ++j) columns[j].insert(columns[j].end(), o.columns[j].begin(), o.columns[j].end());
}

// ------------------------------------------------------------------------------------------------
// args.rs:7-73
// ------------------------------------------------------------------------------------------------
FeatureSetKind feature_set_from_str(const std::string& s) {
    std::string l = s;
    for (auto& c : l) c = (char)std::tolower((unsigned char)c);
    if (l == "geometry") return FeatureSetKind::Geometry;
    if (l == "color") return FeatureSetKind::Color;
    if (l == "glcm") return FeatureSetKind::Glcm;
    if (l == "glrlm") return FeatureSetKind::Glrlm;
    if (l == "gabor") return FeatureSetKind::Gabor;
    if (l == "texture") return FeatureSetKind::Texture;
    if (l == "all") return FeatureSetKind::All;
    throw Error(s + " is not a valid feature set");   // args.rs:29
}
std::vector<FeatureSetKind> flat(const std::vector<FeatureSetKind>& s) {
    std::vector<FeatureSetKind> out;
    for (auto fs : s) {
        if (fs == FeatureSetKind::All)
            out.insert(out.end(), {FeatureSetKind::Geometry, FeatureSetKind::Color, FeatureSetKind::Glcm, FeatureSetKind::Glrlm, FeatureSetKind::Gabor});
        else if (fs == FeatureSetKind::Texture)
            out.insert(out.end(), {FeatureSetKind::Glcm, FeatureSetKind::Glrlm, FeatureSetKind::Gabor});
        else
            out.push_back(fs);
    }
    return out;
}
static uint32_t bit_of(FeatureSetKind k) {
    switch (k) {
        case FeatureSetKind::Geometry: return NFX_FS_GEOMETRY;
        case FeatureSetKind::Color: return NFX_FS_COLOR;
        case FeatureSetKind::Glcm: return NFX_FS_GLCM;
        case FeatureSetKind::Glrlm: return NFX_FS_GLRLM;
        case FeatureSetKind::Gabor: return NFX_FS_GABOR;
        default: return 0;
    }
}
uint32_t feature_mask(const std::vector<FeatureSetKind>& s) {
    if (s.empty()) throw Error("no feature set given");   // main.rs:76 features[0] panics
    uint32_t m = 0;
    for (auto k : flat(s)) {
        if (m & bit_of(k)) throw Error("duplicate feature set");   // main.rs:89
        m |= bit_of(k);
    }
    return m;
}

// ------------------------------------------------------------------------------------------------
// Context
// ------------------------------------------------------------------------------------------------
Context::Context(int device, int patch_size, int batch_size) : patch_size_(patch_size) {
    nfx_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.patch_size = patch_size;
    cfg.batch_size = batch_size;
    if (nfx_create(device, &cfg, &ctx_) != NFX_OK) throw Error(nfx_last_error(nullptr));
}
Context::~Context() { nfx_destroy(ctx_); }
void Context::check(int rc) const {
    if (rc != NFX_OK) throw Error(nfx_last_error(ctx_));
}

static std::string key_of(float x, float y) {
    char b[128];
    nfx_centroid_key(x, y, b, sizeof b);   // utils.rs:226-228
    return b;
}
static DataFrame frame_from(uint32_t mask, const std::vector<std::string>& keys, const float* cent_xy, const std::vector<float>& feat) {
    DataFrame df;
    df.centroid = keys;
    df.centroid_xy.assign(cent_xy, cent_xy + 2 * keys.size());
    const int F = nfx_feature_count(mask);
    const size_t n = keys.size();
    for (int j = 0; j < F; ++j) {
        df.names.push_back(nfx_feature_name(mask, j));
        std::vector<float> col(n);
        for (size_t i = 0; i < n; ++i) col[i] = feat[i * F + j];
        df.columns.push_back(std::move(col));
    }
    return df;
}

// ------------------------------------------------------------------------------------------------
// trait FeatureSet (features/mod.rs:12-28)
// ------------------------------------------------------------------------------------------------
DataFrame FeatureSet::compute_features_batched(const std::vector<Point>& centroids, const std::vector<Points>& polygons,
                                               const Tensor& patchs, const Tensor& masks) const {
    // the asserts of shape.rs:23-47 / color.rs:18-42
    if (patchs.size.size() != 4) throw Error("The patchs tensor must be 4 dimensional");
    if (masks.size.size() != 4) throw Error("The masks tensor must be 4 dimensional");
    if (patchs.size[1] != 3) throw Error("The patchs tensor must have 3 channels");
    if (masks.size[1] != 1) throw Error("The masks tensor must have 1 channel");
    if (patchs.size[0] != masks.size[0]) throw Error("The number of patchs and masks must be the same");
    if ((size_t)patchs.size[0] != centroids.size()) throw Error("The number of patchs and centroids must be the same");
    if ((size_t)patchs.size[0] != polygons.size()) throw Error("The number of patchs and polygons must be the same");
    const int64_t n = patchs.size[0];
    std::vector<float> xy;
    std::vector<int64_t> off(1, 0);
    for (auto& ring : polygons) {
        for (auto& p : ring) { xy.push_back(p[0]); xy.push_back(p[1]); }
        off.push_back((int64_t)xy.size() / 2);
    }
    std::vector<float> out((size_t)n * nfx_feature_count(bit_));
    ctx_.check(nfx_compute_features_batched(ctx_.raw(), bit_, n, &centroids[0][0], xy.data(), off.data(), patchs.data.data(),
                                            masks.data.data(), out.data()));
    std::vector<std::string> keys;
    for (auto& c : centroids) keys.push_back(key_of(c[0], c[1]));
    return frame_from(bit_, keys, centroids.empty() ? nullptr : &centroids[0][0], out);
}

std::vector<std::unique_ptr<FeatureSet>> to_fs(const std::vector<FeatureSetKind>& s, Context& ctx) {
    std::vector<std::unique_ptr<FeatureSet>> out;
    for (auto k : flat(s)) {
        switch (k) {
            case FeatureSetKind::Geometry: out.emplace_back(new ShapeFeatureSet(ctx)); break;
            case FeatureSetKind::Color: out.emplace_back(new ColorFeatureSet(ctx)); break;
            case FeatureSetKind::Glcm: out.emplace_back(new GlcmFeatureSet(ctx)); break;
            case FeatureSetKind::Glrlm: out.emplace_back(new GLRLMFeatureSet(ctx)); break;
            case FeatureSetKind::Gabor: out.emplace_back(new GaborFilterFeatureSet(ctx)); break;
            default: throw Error("unreachable");   // args.rs:70
        }
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// GeoJSON (geojson.rs:8-24, main.rs:37-42): the library's multi-threaded packer
// ------------------------------------------------------------------------------------------------
FeatureCollection load_geometry(const std::string& path, int threads) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) throw Error("cannot open " + path);
    struct stat st;
    if (::fstat(fd, &st) != 0) { ::close(fd); throw Error("cannot stat " + path); }
    const size_t len = (size_t)st.st_size;
    void* map = len ? ::mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;   // parsed in place: no copy of the text
    ::close(fd);
    if (len && map == MAP_FAILED) throw Error("cannot map " + path);
    nfx_geojson* g = nullptr;
    const int rc = nfx_geojson_parse(len ? (const char*)map : "", (int64_t)len, threads, &g);
    if (len) ::munmap(map, len);
    if (rc != NFX_OK) throw Error(nfx_last_error(nullptr));
    FeatureCollection fc;
    const size_t n = (size_t)nfx_geojson_count(g), nv = (size_t)nfx_geojson_vertices(g);
    fc.xy.assign(nfx_geojson_xy(g), nfx_geojson_xy(g) + 2 * nv);
    fc.off.assign(nfx_geojson_offsets(g), nfx_geojson_offsets(g) + n + 1);
    fc.bbox.assign(nfx_geojson_bbox(g), nfx_geojson_bbox(g) + 4 * n);
    nfx_geojson_free(g);
    return fc;
}

// ------------------------------------------------------------------------------------------------
// Images: PNG (8-bit, non-interlaced; grey, grey+alpha, RGB, RGBA, palette) through zlib, and binary PPM
// ------------------------------------------------------------------------------------------------
static std::string ext_of(const std::string& path) {
    const size_t d = path.find_last_of('.'), s = path.find_last_of('/');
    if (d == std::string::npos || (s != std::string::npos && d < s)) return "";
    return path.substr(d + 1);
}

static Image load_ppm(const std::vector<uint8_t>& buf) {
    size_t pos = 2;
    auto token = [&]() -> long {
        while (pos < buf.size()) {
            if (buf[pos] == '#') { while (pos < buf.size() && buf[pos] != '\n') ++pos; }
            else if (std::isspace(buf[pos])) ++pos;
            else break;
        }
        long v = 0;
        while (pos < buf.size() && std::isdigit(buf[pos])) v = v * 10 + (buf[pos++] - '0');
        return v;
    };
    Image im;
    im.w = token();
    im.h = token();
    const long maxv = token();
    ++pos;
    if (maxv != 255 || im.w <= 0 || im.h <= 0 || buf.size() < pos + (size_t)im.w * im.h * 3) throw Error("unsupported PPM (need binary P6, maxval 255)");
    im.rgb.assign(buf.begin() + pos, buf.begin() + pos + (size_t)im.w * im.h * 3);
    return im;
}

static Image load_png(const std::vector<uint8_t>& buf) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8) != 0) throw Error("not a PNG file");
    auto be32 = [&](size_t o) { return ((uint32_t)buf[o] << 24) | ((uint32_t)buf[o + 1] << 16) | ((uint32_t)buf[o + 2] << 8) | buf[o + 3]; };
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte;
    for (size_t pos = 8; pos + 12 <= buf.size();) {
        const uint32_t len = be32(pos);
        const std::string type((const char*)&buf[pos + 4], 4);
        if (pos + 12 + len > buf.size()) throw Error("truncated PNG");
        const uint8_t* d = &buf[pos + 8];
        if (type == "IHDR") { w = be32(pos + 8); h = be32(pos + 12); depth = d[8]; ctype = d[9]; interlace = d[12]; }
        else if (type == "PLTE") plte.assign(d, d + len);
        else if (type == "IDAT") idat.insert(idat.end(), d, d + len);
        else if (type == "IEND") break;
        pos += 12 + len;
    }
    if (depth != 8 || interlace != 0) throw Error("unsupported PNG (need 8-bit, non-interlaced)");
    const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!ch) throw Error("unsupported PNG colour type");
    const size_t stride = (size_t)w * ch;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawlen = raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size()) throw Error("PNG inflate failed");
    std::vector<uint8_t> img(stride * h);
    for (uint32_t y = 0; y < h; ++y) {   // PNG filters 0-4
        const uint8_t ft = raw[y * (stride + 1)];
        const uint8_t* src = &raw[y * (stride + 1) + 1];
        uint8_t* cur = &img[y * stride];
        const uint8_t* up = y ? &img[(y - 1) * stride] : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= (size_t)ch ? cur[x - ch] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)ch) ? up[x - ch] : 0;
            int v = src[x];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: {
                    const int pp = a + b - c, pa = std::abs(pp - a), pb = std::abs(pp - b), pc = std::abs(pp - c);
                    v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: throw Error("bad PNG filter");
            }
            cur[x] = (uint8_t)v;
        }
    }
    Image im;
    im.w = w;
    im.h = h;
    im.rgb.resize((size_t)w * h * 3);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        uint8_t r, g, b;
        if (ctype == 0 || ctype == 4) r = g = b = img[i * ch];
        else if (ctype == 3) {
            const size_t k = (size_t)img[i] * 3;
            if (k + 2 >= plte.size()) throw Error("PNG palette index out of range");
            r = plte[k]; g = plte[k + 1]; b = plte[k + 2];
        } else { r = img[i * ch]; g = img[i * ch + 1]; b = img[i * ch + 2]; }
        im.rgb[i * 3] = r; im.rgb[i * 3 + 1] = g; im.rgb[i * 3 + 2] = b;
    }
    return im;
}

Image load_input_image(const std::string& path) {
    const std::string ext = ext_of(path);
    if (ext == "svs" || ext == "tif" || ext == "tiff") {   // main.rs:21-24 (InputImage::Slide): level 0 of the TIFF container, decoded on the GPU
        Image im;
        std::ifstream tf(path, std::ios::binary | std::ios::ate);
        if (!tf) throw Error("cannot open " + path);
        const std::streamsize tlen = tf.tellg();
        tf.seekg(0);
        im.tiff.resize((size_t)tlen);
        if (tlen && !tf.read((char*)im.tiff.data(), tlen)) throw Error("cannot read " + path);
        nfx_tiff_level lv;
        if (nfx_tiff_info(im.tiff.data(), (int64_t)im.tiff.size(), &lv) != NFX_OK) throw Error(nfx_last_error(nullptr));
        im.w = lv.width;
        im.h = lv.height;
        return im;
    }
    if (ext == "jpg" || ext == "jpeg") throw Error("JPEG decode is not built into nfx-cli (no libjpeg in the image): convert to png, or use `python -m nfx.cli`");
    if (ext != "png" && ext != "ppm")
        throw Error("Unsupported input format. Please use one of the following : svs, png, jpg, jpeg");   // main.rs:29-31
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error("cannot open " + path);
    std::vector<uint8_t> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (buf.size() >= 2 && buf[0] == 'P' && buf[1] == '6') return load_ppm(buf);
    return load_png(buf);
}

void upload_image(Context& ctx, const Image& image) {
    if (!image.tiff.empty()) ctx.check(nfx_slide_load_tiff(ctx.raw(), image.tiff.data(), (int64_t)image.tiff.size(), 0));
    else ctx.check(nfx_tile_upload(ctx.raw(), image.rgb.data(), image.w, image.h, 3 * image.w, 0, 0));
}

// ------------------------------------------------------------------------------------------------
// args.rs:76-183
// ------------------------------------------------------------------------------------------------
static const char* kUsage =
    "Usage: nfx-cli [OPTIONS] <GEOMETRY> <SLIDE> <OUTPUT> [FEATURE_SETS]...\n"
    "  -o, --overwrite            overwrite the output file if it already exists\n"
    "  -p, --patch-size <N>       patch size in pixels [default: 64]\n"
    "  -t, --thread-count <N>     accepted for compatibility (one host thread per GPU is used)\n"
    "  -g, --gpus <ID>...         GPUs to use [default: 0; there is no CPU path]\n"
    "  -b, --batch-size <N>       number of patches per chunk [default: 100]\n"
    "  -v, --verbose\n"
    "      --via-trait            run through FeatureSet::compute_features_batched chunk by chunk\n";

Args parse_args(int argc, char** argv) {
    Args a;
    std::vector<std::string> pos;
    auto need_int = [&](int& i) -> int {
        if (i + 1 >= argc) throw Error(std::string("missing value for ") + argv[i] + "\n" + kUsage);
        char* e = nullptr;
        const long v = std::strtol(argv[++i], &e, 10);
        if (*e) throw Error(std::string("invalid value '") + argv[i] + "'\n" + kUsage);
        return (int)v;
    };
    bool only_pos = false;
    for (int i = 1; i < argc; ++i) {
        const std::string s = argv[i];
        if (only_pos || s.empty() || s[0] != '-') { pos.push_back(s); continue; }
        if (s == "--") only_pos = true;
        else if (s == "-o" || s == "--overwrite") a.overwrite = true;
        else if (s == "-v" || s == "--verbose") a.verbose = true;
        else if (s == "--via-trait") a.via_trait = true;
        else if (s == "--host-csv") a.host_csv = true;
        else if (s == "-p" || s == "--patch-size") a.patch_size = need_int(i);
        else if (s == "-t" || s == "--thread-count") a.thread_count = need_int(i);
        else if (s == "-b" || s == "--batch-size") a.batch_size = need_int(i);
        else if (s == "-g" || s == "--gpus") {
            while (i + 1 < argc && std::isdigit((unsigned char)argv[i + 1][0])) a.gpus.push_back(std::atoi(argv[++i]));
            if (a.gpus.empty()) throw Error(std::string("missing value for --gpus\n") + kUsage);
        } else if (s == "-h" || s == "--help") throw Error(kUsage);
        else throw Error("unexpected argument '" + s + "'\n" + kUsage);
    }
    if (pos.size() < 3) throw Error(std::string("the following required arguments were not provided: <GEOMETRY> <SLIDE> <OUTPUT>\n") + kUsage);
    a.geometry = pos[0];
    a.slide = pos[1];
    a.output = pos[2];
    for (size_t k = 3; k < pos.size(); ++k) a.feature_sets.push_back(feature_set_from_str(pos[k]));
    if (a.gpus.empty()) a.gpus.push_back(0);
    return a;
}

static bool exists(const std::string& p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0;
}

std::string validate_paths(const Args& a) {
    if (!exists(a.geometry)) throw Error("Geometry file does not exist : \"" + a.geometry + "\"");          // args.rs:138-141
    if (!exists(a.slide)) throw Error("Slide file does not exist : \"" + a.slide + "\"");                    // args.rs:142-145
    if (exists(a.output) && !a.overwrite)
        throw Error("Output file already exists : \"" + a.output + "\"\nUse --overwrite to overwrite it");   // args.rs:147-153
    const std::string ext = ext_of(a.output);
    if (ext.empty()) throw Error("Output file must have an extension");                                       // args.rs:158-161
    static const char* ok[] = {"csv", "parquet", "pqt", "json", "ipc", "feather"};
    if (std::find_if(std::begin(ok), std::end(ok), [&](const char* e) { return ext == e; }) == std::end(ok))
        throw Error("Unsupported output format. Please use one of the following : csv, parquet, json, ipc, feather");   // args.rs:162-165
    return ext;
}

// ------------------------------------------------------------------------------------------------
// pipeline (main.rs:146-158)
// ------------------------------------------------------------------------------------------------
static void csr_of(const FeatureCollection& g, size_t lo, size_t hi, std::vector<float>& xy, std::vector<int64_t>& off) {
    xy.assign(g.xy.begin() + 2 * g.off[lo], g.xy.begin() + 2 * g.off[hi]);   // ring 0 as stored
    off.resize(hi - lo + 1);
    for (size_t i = lo; i <= hi; ++i) off[i - lo] = g.off[i] - g.off[lo];
}

DataFrame extract(const FeatureCollection& geometry, const Image& image, const Args& args) {
    const uint32_t mask = feature_mask(args.feature_sets);
    const size_t n = geometry.size();
    const int F = nfx_feature_count(mask);
    std::vector<float> cent(2 * n), feat((size_t)F * n);
    std::vector<int64_t> bounds(args.gpus.size() + 1);
    if (nfx_partition((int64_t)n, args.batch_size, (int)args.gpus.size(), bounds.data()) != NFX_OK) throw Error("bad partition");
    std::vector<std::string> errors(args.gpus.size());
    std::vector<std::thread> threads;
    for (size_t g = 0; g < args.gpus.size(); ++g) {
        threads.emplace_back([&, g] {
            try {
                const size_t lo = (size_t)bounds[g], hi = (size_t)bounds[g + 1];
                if (hi <= lo) return;
                Context ctx(args.gpus[g], args.patch_size, args.batch_size);
                upload_image(ctx, image);
                std::vector<float> xy;
                std::vector<int64_t> off;
                csr_of(geometry, lo, hi, xy, off);
                ctx.check(nfx_extract(ctx.raw(), (int64_t)(hi - lo), xy.data(), off.data(), mask, &cent[2 * lo], &feat[(size_t)F * lo]));
                if (args.verbose) std::fprintf(stderr, "INFO Extracted features for %zu/%zu patches\n", hi, n);   // main.rs:152-157
            } catch (const std::exception& e) {
                errors[g] = e.what();
            }
        });
    }
    for (auto& t : threads) t.join();
    for (auto& e : errors)
        if (!e.empty()) throw Error(e);
    std::vector<std::string> keys(n);
    for (size_t i = 0; i < n; ++i) keys[i] = key_of(cent[2 * i], cent[2 * i + 1]);
    return frame_from(mask, keys, cent.data(), feat);
}

void extract_to_csv(const FeatureCollection& geometry, const Image& image, const Args& args, const std::string& path) {
    const uint32_t mask = feature_mask(args.feature_sets);
    const size_t n = geometry.size();
    std::vector<int64_t> bounds(args.gpus.size() + 1);
    if (nfx_partition((int64_t)n, args.batch_size, (int)args.gpus.size(), bounds.data()) != NFX_OK) throw Error("bad partition");
    std::vector<std::string> errors(args.gpus.size());
    std::vector<std::vector<char>> text(args.gpus.size());   // rows of the ranges after the first, kept until their turn
    std::ofstream f(path, std::ios::binary);
    if (!f) throw Error("cannot create " + path);
    {
        int64_t hl = 0;
        nfx_csv_header(mask, nullptr, 0, &hl);
        std::vector<char> header((size_t)hl);
        if (nfx_csv_header(mask, header.data(), hl, &hl) != NFX_OK) throw Error("csv header");
        f.write(header.data(), hl);
    }
    std::vector<std::thread> threads;
    for (size_t g = 0; g < args.gpus.size(); ++g) {
        threads.emplace_back([&, g] {
            try {
                const size_t lo = (size_t)bounds[g], hi = (size_t)bounds[g + 1];
                if (hi <= lo) return;
                const bool timing = std::getenv("NFX_CLI_TIMING") != nullptr;
                auto t0 = std::chrono::steady_clock::now();
                auto lap = [&](const char* what) {
                    const auto t1 = std::chrono::steady_clock::now();
                    if (timing) std::fprintf(stderr, "[nfx-cli]   gpu %d %-10s %.1f ms\n", args.gpus[g], what, std::chrono::duration<double, std::milli>(t1 - t0).count());
                    t0 = t1;
                };
                Context ctx(args.gpus[g], args.patch_size, args.batch_size);
                lap("context");
                upload_image(ctx, image);
                lap("image->HBM");
                std::vector<float> xy;
                std::vector<int64_t> off;
                csr_of(geometry, lo, hi, xy, off);
                ctx.check(nfx_polygons_upload(ctx.raw(), (int64_t)(hi - lo), xy.data(), off.data()));
                ctx.check(nfx_compute(ctx.raw(), mask));
                ctx.check(nfx_sync(ctx.raw()));
                lap("kernels");
                const int64_t m = (int64_t)(hi - lo), F = nfx_feature_count(mask);
                const int64_t block = std::max<int64_t>(1, std::min<int64_t>(m, (64ll << 20) / (4 * (F + 2))));
                // pinned staging for the text (the D2H runs at PCIe speed); the first range streams straight into the file
                struct Pinned {
                    char* p = nullptr; int64_t cap = 0;
                    void ensure(int64_t need) {
                        if (need <= cap) return;
                        if (p) nfx_host_free(p);
                        void* q = nullptr;
                        if (nfx_host_alloc(&q, need) != NFX_OK) throw Error("cannot allocate pinned host memory");
                        p = (char*)q; cap = need;
                    }
                    ~Pinned() { if (p) nfx_host_free(p); }
                } buf;
                buf.ensure(block * (F + 2) * 12 + 64);
                for (int64_t r = 0; r < m; r += block) {
                    const int64_t r1 = std::min(m, r + block);
                    int64_t len = 0;
                    int rc = nfx_csv_rows(ctx.raw(), r, r1, buf.p, buf.cap, &len);
                    if (rc != NFX_OK && len > buf.cap) {
                        buf.ensure(len);
                        rc = nfx_csv_rows(ctx.raw(), r, r1, buf.p, buf.cap, &len);
                    }
                    ctx.check(rc);
                    if (g == 0) f.write(buf.p, len);
                    else text[g].insert(text[g].end(), buf.p, buf.p + len);
                }
                lap("csv rows");
                if (args.verbose) std::fprintf(stderr, "INFO Extracted features for %zu/%zu patches\n", hi, n);   // main.rs:152-157
            } catch (const std::exception& e) {
                errors[g] = e.what();
            }
        });
    }
    for (auto& t : threads) t.join();
    for (auto& e : errors)
        if (!e.empty()) throw Error(e);
    for (auto& t : text) f.write(t.data(), (std::streamsize)t.size());
}

DataFrame extract_via_trait(const FeatureCollection& geometry, const Image& image, const Args& args) {
    feature_mask(args.feature_sets);   // duplicate / empty checks
    Context ctx(args.gpus[0], args.patch_size, args.batch_size);
    auto sets = to_fs(args.feature_sets, ctx);
    const int P = args.patch_size;
    const size_t n = geometry.size(), plane = (size_t)P * P;
    DataFrame all;
    for (size_t lo = 0; lo < n; lo += (size_t)args.batch_size) {   // par_chunks(batch_size), main.rs:148
        const size_t hi = std::min(n, lo + (size_t)args.batch_size), m = hi - lo;
        // the Batch of utils.rs:17, built by kernels (2) and (1) instead of the host loader
        upload_image(ctx, image);
        std::vector<float> xy;
        std::vector<int64_t> off;
        csr_of(geometry, lo, hi, xy, off);
        ctx.check(nfx_polygons_upload(ctx.raw(), (int64_t)m, xy.data(), off.data()));
        std::vector<uint8_t> mask_u8(m * plane), patch_u8(m * plane * 3);
        ctx.check(nfx_rasterize(ctx.raw(), mask_u8.data()));
        ctx.check(nfx_gather_patches(ctx.raw(), patch_u8.data()));
        std::vector<float> cent(2 * m), dummy;
        ctx.check(nfx_compute(ctx.raw(), NFX_FS_GEOMETRY));
        ctx.check(nfx_download(ctx.raw(), cent.data(), nullptr));
        Tensor patchs{{(int64_t)m, 3, P, P}, std::vector<float>(m * 3 * plane)}, masks{{(int64_t)m, 1, P, P}, std::vector<float>(m * plane)};
        std::vector<Point> centroids(m);
        std::vector<Points> polygons(m);
        for (size_t i = 0; i < m; ++i) {
            centroids[i] = {cent[2 * i], cent[2 * i + 1]};
            const float* ring = geometry.ring(lo + i);
            for (size_t k = 0; k < geometry.ring_len(lo + i); ++k)   // utils.rs:65-72: centred ring
                polygons[i].push_back({ring[2 * k] - centroids[i][0], ring[2 * k + 1] - centroids[i][1]});
            for (size_t px = 0; px < plane; ++px) {
                masks.data[i * plane + px] = mask_u8[i * plane + px] ? 1.0f : 0.0f;
                for (int ch = 0; ch < 3; ++ch)   // utils.rs:172: u8 as f32 / 255
                    patchs.data[(i * 3 + ch) * plane + px] = (float)patch_u8[(i * plane + px) * 3 + ch] / 255.0f;
            }
        }
        DataFrame chunk;
        for (auto& fs : sets) chunk.hstack(fs->compute_features_batched(centroids, polygons, patchs, masks));   // main.rs:52-89
        all.vstack(chunk);
    }
    return all;
}

// ------------------------------------------------------------------------------------------------
// writers (main.rs:160-189)
// ------------------------------------------------------------------------------------------------
static std::string f32s(float v) {
    char b[64];
    nfx_format_f32(v, b, sizeof b);
    return b;
}

// CSV through the device formatter (nfx_csv_format): row-major blocks of the frame go up, text comes back.
static void write_csv_device(std::ofstream& f, const DataFrame& df, int device) {
    Context ctx(device, 64, 100);
    const size_t n = df.height(), F = df.columns.size();
    const size_t block = std::max<size_t>(1, std::min<size_t>(n, (64u << 20) / (4 * (F + 2))));   // ~64 MB of cells per call
    std::vector<float> feat(block * F);
    std::vector<char> text(block * (F + 2) * 12 + 64);
    for (size_t lo = 0; lo < n; lo += block) {
        const size_t m = std::min(block, n - lo);
        for (size_t j = 0; j < F; ++j)
            for (size_t i = 0; i < m; ++i) feat[i * F + j] = df.columns[j][lo + i];
        int64_t len = 0;
        int rc = nfx_csv_format(ctx.raw(), (int64_t)m, (int32_t)F, &df.centroid_xy[2 * lo], feat.data(), text.data(), (int64_t)text.size(), &len);
        if (rc != NFX_OK && len > (int64_t)text.size()) {
            text.resize((size_t)len);
            rc = nfx_csv_format(ctx.raw(), (int64_t)m, (int32_t)F, &df.centroid_xy[2 * lo], feat.data(), text.data(), (int64_t)text.size(), &len);
        }
        ctx.check(rc);
        f.write(text.data(), len);
    }
}

void write_output(const std::string& path, const std::string& ext, const DataFrame& df, int device) {
    if (ext != "csv" && ext != "json")
        throw Error("nfx-cli writes csv and json; parquet / ipc need the polars writers of the Rust host (or `python -m nfx.cli`)");
    std::ofstream f(path, std::ios::binary);
    if (!f) throw Error("cannot create " + path);
    if (ext == "csv") {   // polars 0.32 CsvWriter (oracle/SPEC.md B12)
        f << "centroid";
        for (auto& nme : df.names) f << ',' << nme;
        f << '\n';
        if (device >= 0 && df.centroid_xy.size() == 2 * df.height()) {
            write_csv_device(f, df, device);
        } else {
            for (size_t i = 0; i < df.height(); ++i) {
                f << '"' << df.centroid[i] << '"';   // the key holds a comma: quoted like polars' CsvWriter does
                for (auto& col : df.columns) f << ',' << f32s(col[i]);   // NaN is a value, not a null, in polars
                f << '\n';
            }
        }
    } else {   // JSON lines
        for (size_t i = 0; i < df.height(); ++i) {
            f << "{\"centroid\":\"" << df.centroid[i] << '"';
            for (size_t j = 0; j < df.names.size(); ++j) {
                f << ",\"" << df.names[j] << "\":";
                if (std::isnan(df.columns[j][i]) || std::isinf(df.columns[j][i])) f << "null";
                else f << f32s(df.columns[j][i]);
            }
            f << "}\n";
        }
    }
}

}  // namespace nfxhost
