// jpeg_exact.cpp -- a baseline JPEG decoder whose pixels are bit-identical to libjpeg / libjpeg-turbo with their default
// settings, i.e. to what OpenSlide hands the reference for a JPEG-compressed Aperio tile (src/utils.rs:96-126 reads regions
// through openslide-rs; OpenSlide decodes with libjpeg: slow-but-accurate integer IDCT, "fancy" triangle-filter chroma
// upsampling, fixed-point YCbCr -> RGB). nvJPEG's arithmetic differs by a few grey levels, which would move GLCM counts
// and every other integer-derived column of a slide run; this decoder restates the PUBLISHED algorithms of the IJG
// library (jidctint.c, jdsample.c, jdcolor.c, jdhuff.c, jdmainct.c edge rules) so that the default .svs / .tif path gives
// the reference's pixels. Host code: the entropy decode of JPEG is serial per block whatever the device, and the
// reference decodes on the host too; the tile is uploaded once and everything per nucleus runs on the GPU.
//
// Supported: 8-bit sequential Huffman JPEG (SOF0 / SOF1), one interleaved scan, 1 or 3 components, luma sampling
// 1x1 / 2x1 / 2x2 with 1x1 chroma, restart intervals, tables in the stream (TIFF JPEGTables are merged in front by the
// caller). Anything else returns false with a message and the caller falls back to nvJPEG.
#include <stdint.h>
#include <string.h>

#include <exception>
#include <string>
#include <vector>

#include "nfx_host.h"

namespace nfx {

namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
    bool present = false;
    uint8_t bits[17] = {0}, vals[256] = {0};
    int32_t maxcode[18], valoff[17];
    uint8_t look_n[256], look_v[256];   // 8-bit look-ahead (jdhuff.c HUFF_LOOKAHEAD)
    bool build() {
        int code = 0, k = 0;
        memset(look_n, 0, sizeof look_n);
        for (int l = 1; l <= 16; ++l) {
            valoff[l] = k - code;
            for (int i = 0; i < bits[l]; ++i, ++k, ++code) {
                if (k >= 256) return false;
                if (l <= 8) {
                    const int first = code << (8 - l), cnt = 1 << (8 - l);
                    if (first + cnt > 256) return false;
                    for (int j = 0; j < cnt; ++j) { look_n[first + j] = (uint8_t)l; look_v[first + j] = vals[k]; }
                }
            }
            if (code > (1 << l)) return false;
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        return true;
    }
};

struct Comp {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int pw = 0, ph = 0;            // plane size in samples (whole MCUs)
    int dw = 0, dh = 0;            // libjpeg's downsampled_width / downsampled_height
    std::vector<uint8_t> plane;
    int pred = 0;
};

struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t buf = 0;
    int n = 0;
    bool marker = false;   // a marker was met: the rest of the segment reads as zero bits
    void fill() {
        while (n <= 56) {
            uint32_t b = 0;
            if (!marker && p < end) {
                b = *p;
                if (b == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;
                    else { marker = true; b = 0; }
                } else {
                    ++p;
                }
            } else {
                marker = true;
            }
            buf |= (uint64_t)b << (56 - n);
            n += 8;
        }
    }
    inline uint32_t peek(int k) { return (uint32_t)(buf >> (64 - k)); }
    inline void skip(int k) { buf <<= k; n -= k; }
    inline int32_t receive_extend(int s) {   // jdhuff.c HUFF_EXTEND
        if (s == 0) return 0;
        if (n < s) fill();
        const int32_t v = (int32_t)peek(s);
        skip(s);
        return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
    }
    inline int decode(const Huff& h) {
        if (n < 16) fill();
        const uint32_t look = peek(8);
        int l = h.look_n[look];
        if (l) { skip(l); return h.look_v[look]; }
        l = 9;
        int32_t code = (int32_t)peek(9);
        while (l <= 16 && code > h.maxcode[l]) { ++l; code = (int32_t)peek(l); }
        if (l > 16) { skip(16); return -1; }
        skip(l);
        return h.vals[(code + h.valoff[l]) & 255];
    }
};

// jidctint.c (jpeg_idct_islow): CONST_BITS = 13, PASS1_BITS = 2
constexpr int32_t F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270, F_0_899976223 = 7373,
                  F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137, F_1_961570560 = 16069, F_2_053119869 = 16819,
                  F_2_562915447 = 20995, F_3_072711026 = 25172;
inline int32_t descale(int32_t x, int n) { return (x + (1 << (n - 1))) >> n; }

// post-IDCT range limit: index (x & 1023) of libjpeg's table = clamp(x + 128) for |x| < 512
inline uint8_t idct_limit(int32_t x) {
    const int i = x & 1023;
    return (uint8_t)(i < 128 ? i + 128 : (i < 512 ? 255 : (i < 896 ? 0 : i - 896)));
}

void idct_islow(const int16_t* coef, const uint16_t* q, uint8_t* out, int stride) {
    int32_t ws[64];
    for (int c = 0; c < 8; ++c) {
        const int16_t* in = coef + c;
        const uint16_t* qq = q + c;
        int32_t* w = ws + c;
        if (in[8] == 0 && in[16] == 0 && in[24] == 0 && in[32] == 0 && in[40] == 0 && in[48] == 0 && in[56] == 0) {
            const int32_t dc = (int32_t)((uint32_t)((int32_t)in[0] * qq[0]) << 2);
            for (int k = 0; k < 8; ++k) w[8 * k] = dc;
            continue;
        }
        int32_t z2 = in[16] * qq[16], z3 = in[48] * qq[48];
        int32_t z1 = (z2 + z3) * F_0_541196100;
        int32_t tmp2 = z1 + z3 * (-F_1_847759065), tmp3 = z1 + z2 * F_0_765366865;
        z2 = in[0] * qq[0]; z3 = in[32] * qq[32];
        int32_t tmp0 = (int32_t)((uint32_t)(z2 + z3) << 13), tmp1 = (int32_t)((uint32_t)(z2 - z3) << 13);
        const int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = in[56] * qq[56]; tmp1 = in[40] * qq[40]; tmp2 = in[24] * qq[24]; tmp3 = in[8] * qq[8];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        int32_t z4 = tmp1 + tmp3;
        const int32_t z5 = (z3 + z4) * F_1_175875602;
        tmp0 *= F_0_298631336; tmp1 *= F_2_053119869; tmp2 *= F_3_072711026; tmp3 *= F_1_501321110;
        z1 *= -F_0_899976223; z2 *= -F_2_562915447; z3 *= -F_1_961570560; z4 *= -F_0_390180644;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        w[0] = descale(tmp10 + tmp3, 11);  w[56] = descale(tmp10 - tmp3, 11);
        w[8] = descale(tmp11 + tmp2, 11);  w[48] = descale(tmp11 - tmp2, 11);
        w[16] = descale(tmp12 + tmp1, 11); w[40] = descale(tmp12 - tmp1, 11);
        w[24] = descale(tmp13 + tmp0, 11); w[32] = descale(tmp13 - tmp0, 11);
    }
    for (int r = 0; r < 8; ++r) {
        const int32_t* w = ws + 8 * r;
        uint8_t* o = out + (size_t)r * stride;
        if (w[1] == 0 && w[2] == 0 && w[3] == 0 && w[4] == 0 && w[5] == 0 && w[6] == 0 && w[7] == 0) {
            const uint8_t dc = idct_limit(descale(w[0], 5));
            for (int k = 0; k < 8; ++k) o[k] = dc;
            continue;
        }
        int32_t z2 = w[2], z3 = w[6];
        int32_t z1 = (z2 + z3) * F_0_541196100;
        int32_t tmp2 = z1 + z3 * (-F_1_847759065), tmp3 = z1 + z2 * F_0_765366865;
        int32_t tmp0 = (int32_t)((uint32_t)(w[0] + w[4]) << 13), tmp1 = (int32_t)((uint32_t)(w[0] - w[4]) << 13);
        const int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
        tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
        z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
        int32_t z4 = tmp1 + tmp3;
        const int32_t z5 = (z3 + z4) * F_1_175875602;
        tmp0 *= F_0_298631336; tmp1 *= F_2_053119869; tmp2 *= F_3_072711026; tmp3 *= F_1_501321110;
        z1 *= -F_0_899976223; z2 *= -F_2_562915447; z3 *= -F_1_961570560; z4 *= -F_0_390180644;
        z3 += z5; z4 += z5;
        tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
        o[0] = idct_limit(descale(tmp10 + tmp3, 18)); o[7] = idct_limit(descale(tmp10 - tmp3, 18));
        o[1] = idct_limit(descale(tmp11 + tmp2, 18)); o[6] = idct_limit(descale(tmp11 - tmp2, 18));
        o[2] = idct_limit(descale(tmp12 + tmp1, 18)); o[5] = idct_limit(descale(tmp12 - tmp1, 18));
        o[3] = idct_limit(descale(tmp13 + tmp0, 18)); o[4] = idct_limit(descale(tmp13 - tmp0, 18));
    }
}

// jdsample.c h2v1_fancy_upsample: one row, `dw` input columns -> 2 dw output columns
void h2v1_fancy_row(const uint8_t* in, int dw, uint8_t* out) {
    if (dw == 1) { out[0] = in[0]; out[1] = in[0]; return; }
    const uint8_t* p = in;
    int v = *p++;
    *out++ = (uint8_t)v;
    *out++ = (uint8_t)((v * 3 + p[0] + 2) >> 2);
    for (int c = dw - 2; c > 0; --c) {
        v = (*p++) * 3;
        *out++ = (uint8_t)((v + p[-2] + 1) >> 2);
        *out++ = (uint8_t)((v + p[0] + 2) >> 2);
    }
    v = *p;
    *out++ = (uint8_t)((v * 3 + p[-1] + 1) >> 2);
    *out++ = (uint8_t)v;
}

// jdsample.c h2v2_fancy_upsample: output row from the nearer input row `in0` (weight 3) and the farther one `in1`
void h2v2_fancy_row(const uint8_t* in0, const uint8_t* in1, int dw, uint8_t* out) {
    if (dw == 1) {
        const int s = in0[0] * 3 + in1[0];
        out[0] = (uint8_t)((s * 4 + 8) >> 4);
        out[1] = (uint8_t)((s * 4 + 7) >> 4);
        return;
    }
    int thiscol = (*in0++) * 3 + (*in1++), nextcol = (*in0++) * 3 + (*in1++), lastcol;
    *out++ = (uint8_t)((thiscol * 4 + 8) >> 4);
    *out++ = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
    lastcol = thiscol; thiscol = nextcol;
    for (int c = dw - 2; c > 0; --c) {
        nextcol = (*in0++) * 3 + (*in1++);
        *out++ = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4);
        *out++ = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
        lastcol = thiscol; thiscol = nextcol;
    }
    *out++ = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4);
    *out++ = (uint8_t)((thiscol * 4 + 7) >> 4);
}

inline uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

struct YccTables {   // jdcolor.c build_ycc_rgb_table, SCALEBITS = 16
    int cr_r[256], cb_b[256];
    int32_t cr_g[256], cb_g[256];
    YccTables() {
        const int32_t f140200 = (int32_t)(1.40200 * 65536 + 0.5), f177200 = (int32_t)(1.77200 * 65536 + 0.5),
                      f071414 = (int32_t)(0.71414 * 65536 + 0.5), f034414 = (int32_t)(0.34414 * 65536 + 0.5);
        for (int i = 0; i < 256; ++i) {
            const int32_t x = i - 128;
            cr_r[i] = (int)((f140200 * x + 32768) >> 16);
            cb_b[i] = (int)((f177200 * x + 32768) >> 16);
            cr_g[i] = -f071414 * x;
            cb_g[i] = -f034414 * x + 32768;
        }
    }
};
const YccTables& ycc() {
    static const YccTables t;
    return t;
}

}  // namespace

// colourspace: 0 = the components ARE R,G,B (TIFF Photometric = RGB), 1 = YCbCr, -1 = libjpeg's own rule (JFIF / Adobe marker /
// component ids). out: interleaved u8 RGB, w * h * 3 bytes.
static bool decode_impl(const uint8_t* data, size_t len, int colourspace, std::vector<uint8_t>& out, int& W, int& H, std::string& err);
bool jpeg_decode_exact(const uint8_t* data, size_t len, int colourspace, std::vector<uint8_t>& out, int& W, int& H, std::string& err) {
    try {
        return decode_impl(data, len, colourspace, out, W, H, err);
    } catch (const std::exception& e) {   // allocation failure on a damaged header
        err = e.what();
        return false;
    }
}
static bool decode_impl(const uint8_t* data, size_t len, int colourspace, std::vector<uint8_t>& out, int& W, int& H, std::string& err) {
    auto fail = [&](const char* m) { err = m; return false; };
    if (len < 4 || data[0] != 0xFF || data[1] != 0xD8) return fail("not a JPEG stream");
    uint16_t qt[4][64];
    bool qt_ok[4] = {false, false, false, false};
    Huff dc[4], ac[4];
    std::vector<Comp> comps;
    int restart = 0;
    bool jfif = false, adobe = false;
    int adobe_transform = 0;
    W = H = 0;
    size_t pos = 2;
    const uint8_t* scan = nullptr;
    while (pos + 4 <= len) {
        if (data[pos] != 0xFF) return fail("marker expected");
        while (pos < len && data[pos] == 0xFF) ++pos;   // fill bytes
        if (pos >= len) break;
        const int m = data[pos++];
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) break;
        if (pos + 2 > len) return fail("truncated segment");
        const size_t seglen = ((size_t)data[pos] << 8) | data[pos + 1];
        if (seglen < 2 || pos + seglen > len) return fail("truncated segment");
        const uint8_t* s = data + pos + 2;
        const size_t n = seglen - 2;
        if (m == 0xDB) {   // DQT
            size_t k = 0;
            while (k < n) {
                const int pq = s[k] >> 4, tq = s[k] & 15;
                ++k;
                if (tq > 3 || pq > 1 || k + 64 * (pq + 1) > n) return fail("bad DQT");
                for (int i = 0; i < 64; ++i) {
                    const int v = pq ? ((s[k] << 8) | s[k + 1]) : s[k];
                    k += pq + 1;
                    qt[tq][kZigzag[i]] = (uint16_t)v;
                }
                qt_ok[tq] = true;
            }
        } else if (m == 0xC4) {   // DHT
            size_t k = 0;
            while (k < n) {
                if (k + 17 > n) return fail("bad DHT");
                const int tc = s[k] >> 4, th = s[k] & 15;
                if (tc > 1 || th > 3) return fail("bad DHT");
                Huff& h = tc ? ac[th] : dc[th];
                int total = 0;
                h.bits[0] = 0;
                for (int i = 1; i <= 16; ++i) { h.bits[i] = s[k + i]; total += h.bits[i]; }
                k += 17;
                if (total > 256 || k + total > n) return fail("bad DHT");
                memcpy(h.vals, s + k, total);
                k += total;
                if (!h.build()) return fail("bad Huffman table");
                h.present = true;
            }
        } else if (m == 0xC0 || m == 0xC1) {   // SOF0 / SOF1
            if (n < 6 || s[0] != 8) return fail("only 8-bit sequential JPEG is decoded exactly");
            H = (s[1] << 8) | s[2];
            W = (s[3] << 8) | s[4];
            const int nc = s[5];
            if ((nc != 1 && nc != 3) || n < (size_t)(6 + 3 * nc) || W <= 0 || H <= 0) return fail("unsupported component count");
            comps.resize(nc);
            for (int c = 0; c < nc; ++c) {
                comps[c].id = s[6 + 3 * c];
                comps[c].h = s[7 + 3 * c] >> 4;
                comps[c].v = s[7 + 3 * c] & 15;
                comps[c].tq = s[8 + 3 * c] & 3;
            }
        } else if (m >= 0xC2 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return fail("progressive / lossless / arithmetic JPEG is not decoded exactly");
        } else if (m == 0xDD) {
            if (n < 2) return fail("bad DRI");
            restart = (s[0] << 8) | s[1];
        } else if (m == 0xE0) {
            if (n >= 5 && !memcmp(s, "JFIF", 5)) jfif = true;
        } else if (m == 0xEE) {
            if (n >= 12 && !memcmp(s, "Adobe", 5)) { adobe = true; adobe_transform = s[11]; }
        } else if (m == 0xDA) {   // SOS
            if (comps.empty()) return fail("SOS before SOF");
            if (n < 1 || s[0] != (int)comps.size() || n < (size_t)(1 + 2 * s[0] + 3)) return fail("only one interleaved scan is decoded exactly");
            for (int k = 0; k < s[0]; ++k) {
                Comp* c = nullptr;
                for (auto& cc : comps)
                    if (cc.id == s[1 + 2 * k]) c = &cc;
                if (!c || c != &comps[k]) return fail("scan components out of frame order");
                c->td = s[2 + 2 * k] >> 4;
                c->ta = s[2 + 2 * k] & 15;
                if (c->td > 3 || c->ta > 3) return fail("bad table selector");
            }
            scan = data + pos + seglen;
            break;
        }
        pos += seglen;
    }
    if (!scan || comps.empty()) return fail("no scan found");
    const int nc = (int)comps.size();
    int hmax = 1, vmax = 1;
    for (auto& c : comps) {
        if (c.h < 1 || c.v < 1 || c.h > 2 || c.v > 2) return fail("sampling factors beyond 2 are not decoded exactly");
        hmax = c.h > hmax ? c.h : hmax;
        vmax = c.v > vmax ? c.v : vmax;
        if (!qt_ok[c.tq] || !dc[c.td].present || !ac[c.ta].present) return fail("missing quantisation or Huffman table");
    }
    if (nc == 1) { hmax = comps[0].h = 1; vmax = comps[0].v = 1; }   // a one-component scan is never interleaved: 8 x 8 MCUs
    if (nc == 3) {
        if (comps[1].h != 1 || comps[1].v != 1 || comps[2].h != 1 || comps[2].v != 1) return fail("chroma sampling other than 1x1");
        if (comps[0].h == 1 && comps[0].v == 2) return fail("1x2 luma sampling is not decoded exactly");
    }
    const int mcuw = 8 * hmax, mcuh = 8 * vmax, mx = (W + mcuw - 1) / mcuw, my = (H + mcuh - 1) / mcuh;
    if ((int64_t)mx * mcuw * (int64_t)my * mcuh > ((int64_t)1 << 28)) return fail("picture larger than 2^28 samples (a slide block never is)");
    for (auto& c : comps) {
        c.pw = mx * c.h * 8;
        c.ph = my * c.v * 8;
        c.dw = (W * c.h + hmax - 1) / hmax;
        c.dh = (H * c.v + vmax - 1) / vmax;
        c.plane.assign((size_t)c.pw * c.ph, 0);
        c.pred = 0;
    }
    // ---- entropy decode + dequantise + IDCT, MCU by MCU ----
    BitReader br{scan, data + len};
    int16_t coef[64];
    int togo = restart;
    for (int y = 0; y < my; ++y) {
        for (int x = 0; x < mx; ++x) {
            if (restart && togo == 0) {
                // byte-align, expect RSTn, reset the predictors (jdhuff.c process_restart)
                br.buf = 0; br.n = 0;
                const uint8_t* p = br.p;
                while (p + 1 < br.end && !(p[0] == 0xFF && p[1] >= 0xD0 && p[1] <= 0xD7)) ++p;
                if (p + 1 >= br.end) return fail("restart marker missing");
                br.p = p + 2;
                br.marker = false;
                for (auto& c : comps) c.pred = 0;
                togo = restart;
            }
            for (auto& c : comps) {
                const Huff &hd = dc[c.td], &ha = ac[c.ta];
                for (int by = 0; by < c.v; ++by)
                    for (int bx = 0; bx < c.h; ++bx) {
                        memset(coef, 0, sizeof coef);
                        int s = br.decode(hd);
                        if (s < 0 || s > 11) return fail("bad DC code");
                        c.pred += br.receive_extend(s);
                        coef[0] = (int16_t)c.pred;
                        for (int k = 1; k < 64;) {
                            const int rs = br.decode(ha);
                            if (rs < 0) return fail("bad AC code");
                            const int r = rs >> 4;
                            s = rs & 15;
                            if (s) {
                                k += r;
                                if (k > 63) return fail("AC run beyond the block");
                                coef[kZigzag[k]] = (int16_t)br.receive_extend(s);
                                ++k;
                            } else {
                                if (r != 15) break;   // EOB
                                k += 16;
                            }
                        }
                        idct_islow(coef, qt[c.tq], c.plane.data() + (size_t)((y * c.v + by) * 8) * c.pw + (size_t)(x * c.h + bx) * 8, c.pw);
                    }
            }
            if (restart) --togo;
        }
    }
    // ---- upsample + colour convert ----
    out.assign((size_t)W * H * 3, 0);
    if (nc == 1) {
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const uint8_t v = comps[0].plane[(size_t)y * comps[0].pw + x];
                uint8_t* o = &out[((size_t)y * W + x) * 3];
                o[0] = o[1] = o[2] = v;
            }
        return true;
    }
    bool is_ycc;
    if (colourspace == 0) is_ycc = false;
    else if (colourspace == 1) is_ycc = true;
    else if (jfif) is_ycc = true;
    else if (adobe) is_ycc = adobe_transform != 0;
    else is_ycc = !(comps[0].id == 'R' && comps[1].id == 'G' && comps[2].id == 'B');   // jdapimin.c default_decompress_parms
    const int hs = comps[0].h, vs = comps[0].v;
    std::vector<uint8_t> up[2];
    const uint8_t* rowp[3];
    for (auto& u : up) u.assign((size_t)2 * comps[1].pw + 8, 0);
    const YccTables& T = ycc();
    for (int y = 0; y < H; ++y) {
        rowp[0] = comps[0].plane.data() + (size_t)y * comps[0].pw;
        for (int k = 1; k < 3; ++k) {
            const Comp& c = comps[k];
            if (hs == 1 && vs == 1) {
                rowp[k] = c.plane.data() + (size_t)y * c.pw;
            } else if (vs == 1) {   // h2v1
                h2v1_fancy_row(c.plane.data() + (size_t)y * c.pw, c.dw, up[k - 1].data());
                rowp[k] = up[k - 1].data();
            } else {                // h2v2: jdmainct.c duplicates the first / last REAL sample row as context
                const int r = y >> 1;
                int r1 = (y & 1) ? r + 1 : r - 1;
                r1 = r1 < 0 ? 0 : (r1 > c.dh - 1 ? c.dh - 1 : r1);
                h2v2_fancy_row(c.plane.data() + (size_t)r * c.pw, c.plane.data() + (size_t)r1 * c.pw, c.dw, up[k - 1].data());
                rowp[k] = up[k - 1].data();
            }
        }
        uint8_t* o = &out[(size_t)y * W * 3];
        if (is_ycc) {
            for (int x = 0; x < W; ++x, o += 3) {
                const int yy = rowp[0][x], cb = rowp[1][x], cr = rowp[2][x];
                o[0] = clamp8(yy + T.cr_r[cr]);
                o[1] = clamp8(yy + (int)((T.cb_g[cb] + T.cr_g[cr]) >> 16));
                o[2] = clamp8(yy + T.cb_b[cb]);
            }
        } else {
            for (int x = 0; x < W; ++x, o += 3) { o[0] = rowp[0][x]; o[1] = rowp[1][x]; o[2] = rowp[2][x]; }
        }
    }
    return true;
}

}  // namespace nfx
