// schema.cpp -- host-side contract of the reference that does not touch the GPU: feature-set names
// (src/args.rs:7-49), column names and order (src/features/*.rs), the centroid key string
// (src/utils.rs:226-232) and the multi-GPU partition rule.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/nfx.h"
#include "f32_display.h"
#include "nfx_host.h"

namespace nfx {

// Rust `impl Display for f32` (f32_display.h): shortest round-trip digits, never an exponent, no ".0".
static const uint64_t kPow5Inv[] = NFX_POW5_INV_SPLIT;
static const uint64_t kPow5[] = NFX_POW5_SPLIT;
std::string rust_f32_display(float x) {
    uint32_t bits;
    memcpy(&bits, &x, 4);
    char buf[NFX_F32_MAX_CHARS];
    const int n = f32_display(bits, buf, kPow5Inv, kPow5);
    return std::string(buf, (size_t)n);
}

namespace {

const char* kShape[] = {"area", "major_axis", "minor_axis", "eccentricity", "orientation", "perimeter",
                        "equivalent_perimeter", "compacity", "eliptic_deviation", "convex_hull_area",
                        "convex_deffect", "convex_perimeter"};   // shape.rs:114-128
const char* kColor[] = {"mean_r", "mean_g", "mean_b", "std_r", "std_g", "std_b", "mean_h", "mean_s",
                        "mean_v", "std_h", "std_s", "std_v", "mean_haematoxylin", "mean_eosin", "mean_dab",
                        "std_haematoxylin", "std_eosin", "std_dab"};   // color.rs:80-100
const char* kGlcm[] = {"correlation", "contrast", "dissimilarity", "entropy", "angular_second_moment",
                       "sum_average", "sum_variance", "sum_entropy", "sum_of_squares",
                       "inverse_difference_moment", "difference_average", "difference_variance",
                       "information_measure_correlation1", "information_measure_correlation2"};   // texture.rs:81-157
const int kGlcmLv[] = {32, 64, 128, 254};                                // texture.rs:19
const int kGlcmOff[][2] = {{0, 1}, {1, 1}, {1, 0}, {1, -1}};             // texture.rs:20
const char* kGlrlm[] = {"short_run_emphasis", "long_run_emphasis", "gray_level_nonuniformity",
                        "run_length_nonuniformity", "low_gray_level_run_emphasis",
                        "high_gray_level_run_emphasis", "short_run_low_gray_level_emphasis",
                        "short_run_high_gray_level_emphasis", "long_run_low_gray_level_emphasis",
                        "long_run_high_gray_level_emphasis", "short_run_mid_gray_level_emphasis",
                        "long_run_mid_gray_level_emphasis", "short_run_extreme_gray_level_emphasis",
                        "long_run_extreme_gray_level_emphasis", "run_percentage", "run_length_mean",
                        "run_length_variance"};                           // texture.rs:243-301
const int kGlrlmDir[][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}};            // texture.rs:176
const double kGaborFreq[] = {0.5, 1.0, 2.0, 4.0, 6.0, 8.0};              // texture.rs:320

struct Schema {
    std::vector<std::string> cols[5];
    Schema() {
        for (const char* s : kShape) cols[0].push_back(s);
        for (const char* s : kColor) cols[1].push_back(s);
        char b[160];
        for (int L : kGlcmLv)
            for (auto& o : kGlcmOff)
                for (const char* f : kGlcm) {
                    snprintf(b, sizeof b, "%s_%d_%d_%d", f, o[0], o[1], L);
                    cols[2].push_back(b);
                }
        for (auto& d : kGlrlmDir)
            for (const char* f : kGlrlm) {
                snprintf(b, sizeof b, "%s_%d_%d", f, d[0], d[1]);
                cols[3].push_back(b);
            }
        for (int j = 0; j < 48; ++j) {   // texture.rs:346-361
            const float angle = (float)(j / 6) * 45.0f;
            const float freq = (float)kGaborFreq[j % 6];
            for (const char* s : {"mean", "variance"}) {
                snprintf(b, sizeof b, "gabor_angle_%s_frequency_%s_%s", rust_f32_display(angle).c_str(),
                         rust_f32_display(freq).c_str(), s);
                cols[4].push_back(b);
            }
        }
    }
};
const Schema& schema() {
    static Schema s;
    return s;
}

}  // namespace

int set_cols(int set_index) { return (int)schema().cols[set_index].size(); }

int column_offset(uint32_t mask, uint32_t bit) {
    int off = 0;
    for (int k = 0; k < 5; ++k) {
        if ((1u << k) == bit) return (mask & bit) ? off : -1;
        if (mask & (1u << k)) off += set_cols(k);
    }
    return -1;
}

thread_local std::string g_thread_error;

}  // namespace nfx

extern "C" {

int nfx_feature_count(uint32_t mask) {
    int n = 0;
    for (int k = 0; k < 5; ++k)
        if (mask & (1u << k)) n += nfx::set_cols(k);
    return n;
}

const char* nfx_feature_name(uint32_t mask, int idx) {
    if (idx < 0) return nullptr;
    for (int k = 0; k < 5; ++k) {
        if (!(mask & (1u << k))) continue;
        const auto& c = nfx::schema().cols[k];
        if (idx < (int)c.size()) return c[idx].c_str();
        idx -= (int)c.size();
    }
    return nullptr;
}

int nfx_parse_feature_set(const char* name, uint32_t* bits) {
    if (!name || !bits) return NFX_ERR_INVALID;
    std::string s(name);
    for (auto& ch : s) ch = (char)tolower((unsigned char)ch);
    struct { const char* n; uint32_t b; } tab[] = {
        {"geometry", NFX_FS_GEOMETRY}, {"color", NFX_FS_COLOR}, {"glcm", NFX_FS_GLCM},
        {"glrlm", NFX_FS_GLRLM}, {"gabor", NFX_FS_GABOR}, {"texture", NFX_FS_TEXTURE}, {"all", NFX_FS_ALL}};
    for (auto& t : tab)
        if (s == t.n) { *bits = t.b; return NFX_OK; }
    nfx::g_thread_error = std::string(name) + " is not a valid feature set";   // args.rs:29
    return NFX_ERR_INVALID;
}

const char* nfx_feature_set_name(uint32_t bit) {
    switch (bit) {
        case NFX_FS_GEOMETRY: return "geometry";
        case NFX_FS_COLOR: return "color";
        case NFX_FS_GLCM: return "GLCM";
        case NFX_FS_GLRLM: return "GLRLM";
        case NFX_FS_GABOR: return "gabor filter";
        default: return nullptr;
    }
}

int nfx_centroid_key(float x, float y, char* buf, int buflen) {
    const std::string s = nfx::rust_f32_display(x) + "," + nfx::rust_f32_display(y);
    if (!buf || (int)s.size() + 1 > buflen) return NFX_ERR_INVALID;
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

int nfx_format_f32(float v, char* buf, int buflen) {
    const std::string s = nfx::rust_f32_display(v);
    if (!buf || (int)s.size() + 1 > buflen) return NFX_ERR_INVALID;
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

int nfx_csv_header(uint32_t mask, char* out, int64_t cap, int64_t* len) {
    if (!len) return NFX_ERR_INVALID;
    std::string h = "centroid";                       // main.rs:80-88: the key column, then the sets in flat() order
    const int f = nfx_feature_count(mask);
    for (int j = 0; j < f; ++j) { h += ','; h += nfx_feature_name(mask, j); }
    h += '\n';
    *len = (int64_t)h.size();
    if (!out || cap < (int64_t)h.size()) { nfx::g_thread_error = "csv: header buffer too small"; return NFX_ERR_INVALID; }
    memcpy(out, h.data(), h.size());
    return NFX_OK;
}

int nfx_partition(int64_t n, int32_t batch_size, int32_t parts, int64_t* bounds) {
    if (n < 0 || batch_size <= 0 || parts <= 0 || !bounds) return NFX_ERR_INVALID;
    const int64_t chunks = (n + batch_size - 1) / batch_size;
    for (int p = 0; p <= parts; ++p) {
        const int64_t c = (chunks * p) / parts;   // chunk index boundary, balanced to within one chunk
        const int64_t b = c * batch_size;
        bounds[p] = b < n ? b : n;
    }
    bounds[parts] = n;
    return NFX_OK;
}

const char* nfx_version(void) { return "nfx 0.1.0 (sm_100a)"; }

}  // extern "C"
