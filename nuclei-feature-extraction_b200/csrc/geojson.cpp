// geojson.cpp -- GeoJSON -> CSR polygon packing. Replaces
// `serde_json::from_reader::<FeatureCollection>` (src/main.rs:37-42) with the model of src/geojson.rs:8-24
// and the ring-0 selection of preprocess_polygon (src/utils.rs:54-60). Host code, one pass per thread:
//   A  unescaped-quote parity per chunk        -> is the chunk start inside a string?
//   B  bracket depth delta per chunk           -> absolute depth at the chunk start
//   C  '{' at depth 2->3 and ']' at depth 2->1 -> where every feature object starts, where the array ends
//   D  the features are split evenly over the threads and parsed (strictly) into thread-local CSR pieces
//   E  pieces are concatenated in input order
// Numbers follow serde_json 1.0.107 (Cargo.lock:2298; `float_roundtrip` off, Cargo.toml:25): u64 significand,
// decimal exponent, one u64->f64 conversion and one multiply/divide by a power of ten (src/de.rs
// parse_integer / parse_decimal / f64_from_parts), then `as f32` -- NOT a correctly rounded strtof.
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <chrono>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nfx.h"
#include "nfx_host.h"

// A buffer that is filled by the parser threads: no value-initialisation pass over it.
template <typename T>
struct RawBuf {
    std::unique_ptr<T[]> p;
    size_t n = 0;
    void resize(size_t m) { p.reset(new T[m]); n = m; }
    T* data() const { return p.get(); }
    size_t size() const { return n; }
    T& operator[](size_t i) const { return p[i]; }
};
struct nfx_geojson {
    RawBuf<float> xy;        // ring 0 of every feature, [nv][2]
    RawBuf<int64_t> off;     // [n+1]
    RawBuf<float> bbox;      // [n][4], NaN where the bbox array is shorter
    RawBuf<int32_t> rings;   // [n] number of rings of the feature (only ring 0 is kept, utils.rs:55)
};

namespace {

double g_pow10[309];
std::once_flag g_pow10_once;
void init_pow10() {
    std::call_once(g_pow10_once, [] {
        char b[16];
        for (int k = 0; k <= 308; ++k) {   // the f64 nearest to 10^k, like the literals of serde_json's POW10
            snprintf(b, sizeof b, "1e%d", k);
            g_pow10[k] = strtod(b, nullptr);
        }
    });
}

struct ParseError {
    std::string msg;
    size_t pos;
};

inline bool is_ws(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r'; }
inline bool is_digit(char c) { return (unsigned)(c - '0') < 10u; }

// serde_json f64_from_parts
inline bool f64_from_parts(bool positive, uint64_t significand, int32_t exponent, double* out) {
    double f = (double)significand;
    for (;;) {
        const uint32_t a = exponent < 0 ? (uint32_t)(-(int64_t)exponent) : (uint32_t)exponent;
        if (a <= 308) {
            if (exponent >= 0) {
                f *= g_pow10[a];
                if (isinf(f)) return false;   // NumberOutOfRange
            } else {
                f /= g_pow10[a];
            }
            break;
        }
        if (f == 0.0) break;
        if (exponent >= 0) return false;
        f /= 1e308;
        exponent += 308;
    }
    *out = positive ? f : -f;
    return true;
}

// One JSON number at p (no leading whitespace) deserialised as f32. Returns the end, or nullptr (msg set).
const char* parse_f32(const char* p, const char* end, float* out, const char** msg) {
    bool positive = true;
    if (p < end && *p == '-') { positive = false; ++p; }
    if (p >= end || !is_digit(*p)) { *msg = "invalid number"; return nullptr; }
    uint64_t sig = 0;
    int32_t exponent = 0;
    bool is_float = false;
    if (*p == '0') {
        ++p;
        if (p < end && is_digit(*p)) { *msg = "invalid number"; return nullptr; }   // no leading zeros
    } else {
        while (p < end && is_digit(*p)) {
            const uint64_t d = (uint64_t)(*p - '0');
            if (sig >= UINT64_MAX / 10 && (sig > UINT64_MAX / 10 || d > UINT64_MAX % 10)) {
                // parse_long_integer: the remaining integer digits only scale the value
                while (p < end && is_digit(*p)) { ++p; ++exponent; }
                is_float = true;
                break;
            }
            sig = sig * 10 + d;
            ++p;
        }
    }
    if (p < end && *p == '.') {
        is_float = true;
        ++p;
        if (p >= end || !is_digit(*p)) { *msg = "invalid number"; return nullptr; }
        while (p < end && is_digit(*p)) {
            const uint64_t d = (uint64_t)(*p - '0');
            if (sig >= UINT64_MAX / 10 && (sig > UINT64_MAX / 10 || d > UINT64_MAX % 10)) {
                while (p < end && is_digit(*p)) ++p;   // parse_decimal_overflow: further digits are dropped
                break;
            }
            sig = sig * 10 + d;
            --exponent;
            ++p;
        }
    }
    if (p < end && (*p == 'e' || *p == 'E')) {
        is_float = true;
        ++p;
        bool epos = true;
        if (p < end && (*p == '+' || *p == '-')) { epos = *p == '+'; ++p; }
        if (p >= end || !is_digit(*p)) { *msg = "invalid number"; return nullptr; }
        int64_t e = 0;
        bool eover = false;
        while (p < end && is_digit(*p)) {
            e = e * 10 + (*p - '0');
            if (e > INT32_MAX) eover = true, e = INT32_MAX;
            ++p;
        }
        if (eover) {   // parse_exponent_overflow
            if (sig != 0 && epos) { *msg = "number out of range"; return nullptr; }
            *out = positive ? 0.0f : -0.0f;
            return p;
        }
        int64_t fe = epos ? (int64_t)exponent + e : (int64_t)exponent - e;
        fe = std::max<int64_t>(INT32_MIN, std::min<int64_t>(INT32_MAX, fe));   // saturating_add / _sub
        exponent = (int32_t)fe;
    }
    if (!is_float) {   // ParserNumber::U64 / I64 -> visit_u64 / visit_i64 -> `as f32` (one rounding)
        if (positive) { *out = (float)sig; return p; }
        const int64_t neg = (int64_t)(0 - sig);
        if (neg >= 0 && sig != 0) { *out = (float)(-(double)sig); return p; }   // does not fit i64: F64
        if (sig == 0) { *out = -0.0f; return p; }      // "-0" is F64(-0.0) in serde_json
        *out = (float)neg;
        return p;
    }
    double f;
    if (!f64_from_parts(positive, sig, exponent, &f)) { *msg = "number out of range"; return nullptr; }
    *out = (float)f;
    return p;
}

// ---- fast path of parse_f32: at most 19 digits in total can never overflow the u64 significand -------
const uint64_t kPow10u[9] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull};

// Up to 8 leading ASCII digits of the little-endian word v: returns how many (0..8) and their value.
inline int digits8(uint64_t v, uint32_t* val) {
    const uint64_t w = v - 0x3030303030303030ull;
    const uint64_t nondigit = (w | (v + 0x4646464646464646ull)) & 0x8080808080808080ull;
    const int n = nondigit ? (__builtin_ctzll(nondigit) >> 3) : 8;
    if (n == 0) { *val = 0; return 0; }
    uint64_t x = w << (64 - 8 * n);   // the n digits become the low-order decimal places, zeros in front
    x = (x * 10) + (x >> 8);
    x = (((x & 0x000000FF000000FFull) * (100ull + (1000000ull << 32))) +
         (((x >> 16) & 0x000000FF000000FFull) * (1ull + (10000ull << 32)))) >> 32;
    *val = (uint32_t)x;
    return n;
}

// Digit run at p: folds it into *sig, returns the number of digits. Needs 8 readable bytes past the run's end
// only while `end - p >= 8`; the tail goes byte by byte.
inline int digit_run(const char*& p, const char* end, uint64_t* sig) {
    int total = 0;
    while (end - p >= 8) {
        uint64_t v;
        memcpy(&v, p, 8);
        uint32_t val;
        const int n = digits8(v, &val);
        *sig = *sig * kPow10u[n] + val;
        p += n;
        total += n;
        if (n < 8) return total;
        if (total > 24) return total;   // the caller falls back to the exact slow path anyway
    }
    while (p < end && is_digit(*p)) { *sig = *sig * 10 + (uint64_t)(*p - '0'); ++p; ++total; }
    return total;
}

// Same result as parse_f32 for plain decimals `-?int[.frac]` with <= 19 digits; everything else (exponents,
// longer numbers, malformed tokens) is handed to parse_f32.
inline const char* parse_f32_fast(const char* p0, const char* end, float* out, const char** msg) {
    const char* p = p0;
    bool positive = true;
    if (p < end && *p == '-') { positive = false; ++p; }
    uint64_t sig = 0;
    const char* ip = p;
    const int ni = digit_run(p, end, &sig);
    if (ni == 0 || ni > 19 || (ni > 1 && *ip == '0')) return parse_f32(p0, end, out, msg);
    if (p < end && *p == '.') {
        ++p;
        const int nf = digit_run(p, end, &sig);
        if (nf == 0 || ni + nf > 19 || (p < end && (*p == 'e' || *p == 'E'))) return parse_f32(p0, end, out, msg);
        const double f = (double)sig / g_pow10[nf];      // f64_from_parts with exponent -nf
        *out = (float)(positive ? f : -f);
        return p;
    }
    if (p < end && (*p == 'e' || *p == 'E')) return parse_f32(p0, end, out, msg);
    if (!positive) return parse_f32(p0, end, out, msg);
    *out = (float)sig;                                     // ParserNumber::U64 -> `as f32`
    return p;
}

// ---- strict recursive-descent reader over [p,end) ------------------------------------------------
struct Reader {
    const char* base;
    const char* p;
    const char* end;
    [[noreturn]] void fail(const std::string& m) const { throw ParseError{m, (size_t)(p - base)}; }
    void ws() { while (p < end && is_ws(*p)) ++p; }
    bool eat(char c) { ws(); if (p < end && *p == c) { ++p; return true; } return false; }
    void expect(char c) { if (!eat(c)) fail(std::string("expected `") + c + "`"); }
    // string without unescaping; returns [b,e) of the raw contents; sets esc if it holds a backslash
    void raw_string(const char** b, const char** e, bool* esc) {
        expect('"');
        *b = p;
        *esc = false;
        while (p < end && *p != '"') {
            if ((unsigned char)*p < 0x20) fail("control character in string");
            if (*p == '\\') {   // serde_json checks the escape even in strings it skips (ignore_str / ignore_escape)
                *esc = true;
                ++p;
                if (p >= end) break;
                const char e2 = *p;
                if (e2 == 'u') {
                    if (end - p < 5) fail("EOF while parsing a string");
                    for (int k = 1; k <= 4; ++k)
                        if (!isxdigit((unsigned char)p[k])) fail("invalid escape");
                    p += 4;
                } else if (e2 != '"' && e2 != '\\' && e2 != '/' && e2 != 'b' && e2 != 'f' && e2 != 'n' && e2 != 'r' && e2 != 't') {
                    fail("invalid escape");
                }
            }
            ++p;
        }
        if (p >= end) fail("EOF while parsing a string");
        *e = p;
        ++p;
    }
    bool key_is(const char* b, const char* e, const char* name) const {
        const size_t n = strlen(name);
        return (size_t)(e - b) == n && memcmp(b, name, n) == 0;
    }
    float number() {
        ws();
        float v;
        const char* msg = nullptr;
        const char* q = parse_f32_fast(p, end, &v, &msg);
        if (!q) {
            if (p < end && (*p == '"' || *p == '[' || *p == '{' || *p == 't' || *p == 'f' || *p == 'n'))
                fail("invalid type: expected f32");
            fail(msg);
        }
        p = q;
        return v;
    }
    void skip_value(int depth = 0) {
        if (depth > 128) fail("recursion limit exceeded");
        ws();
        if (p >= end) fail("EOF while parsing a value");
        const char c = *p;
        if (c == '"') { const char *b, *e; bool esc; raw_string(&b, &e, &esc); return; }
        if (c == '{') {
            ++p;
            if (eat('}')) return;
            do { const char *b, *e; bool esc; raw_string(&b, &e, &esc); expect(':'); skip_value(depth + 1); } while (eat(','));
            expect('}');
            return;
        }
        if (c == '[') {
            ++p;
            if (eat(']')) return;
            do { skip_value(depth + 1); } while (eat(','));
            expect(']');
            return;
        }
        if (c == 't' && end - p >= 4 && !memcmp(p, "true", 4)) { p += 4; return; }
        if (c == 'f' && end - p >= 5 && !memcmp(p, "false", 5)) { p += 5; return; }
        if (c == 'n' && end - p >= 4 && !memcmp(p, "null", 4)) { p += 4; return; }
        float v;
        const char* msg = nullptr;
        const char* q = parse_f32(p, end, &v, &msg);
        if (!q) fail("expected value");
        p = q;
    }
};

struct Piece {   // what one thread produced
    std::vector<float> xy, bbox;
    std::vector<int64_t> nverts;
    std::vector<int32_t> rings;
};

// Feature (geojson.rs:16-20) at r.p == '{'. Unknown keys are ignored, duplicates and missing fields are errors.
void parse_feature(Reader& r, Piece& out) {
    r.expect('{');
    bool have_bbox = false, have_geom = false;
    if (!r.eat('}')) {
        do {
            const char *kb, *ke;
            bool esc;
            r.raw_string(&kb, &ke, &esc);
            r.expect(':');
            if (!esc && r.key_is(kb, ke, "bbox")) {
                if (have_bbox) r.fail("duplicate field `bbox`");
                have_bbox = true;
                float bb[4] = {NAN, NAN, NAN, NAN};
                if (!r.eat('[')) r.fail("invalid type: expected a sequence");
                int k = 0;
                if (!r.eat(']')) {
                    do { const float v = r.number(); if (k < 4) bb[k] = v; ++k; } while (r.eat(','));
                    r.expect(']');
                }
                out.bbox.insert(out.bbox.end(), bb, bb + 4);
            } else if (!esc && r.key_is(kb, ke, "geometry")) {
                if (have_geom) r.fail("duplicate field `geometry`");
                have_geom = true;
                if (!r.eat('{')) r.fail("invalid type: expected struct Geometry");
                bool have_type = false, have_coords = false;
                if (!r.eat('}')) {
                    do {
                        const char *gb, *ge;
                        bool gesc;
                        r.raw_string(&gb, &ge, &gesc);
                        r.expect(':');
                        if (!gesc && r.key_is(gb, ge, "type")) {
                            if (have_type) r.fail("duplicate field `type`");
                            have_type = true;
                            r.ws();
                            if (r.p >= r.end || *r.p != '"') r.fail("invalid type: expected a string");
                            const char *b, *e;
                            bool s;
                            r.raw_string(&b, &e, &s);
                        } else if (!gesc && r.key_is(gb, ge, "coordinates")) {   // Vec<Vec<Vec<f32>>>
                            if (have_coords) r.fail("duplicate field `coordinates`");
                            have_coords = true;
                            if (!r.eat('[')) r.fail("invalid type: expected a sequence");
                            int32_t nring = 0;
                            int64_t nv = 0;
                            if (!r.eat(']')) {
                                do {
                                    if (!r.eat('[')) r.fail("invalid type: expected a sequence");
                                    if (!r.eat(']')) {
                                        do {
                                            if (!r.eat('[')) r.fail("invalid type: expected a sequence");
                                            int k = 0;
                                            float xy[2] = {0, 0};
                                            if (!r.eat(']')) {
                                                do { const float v = r.number(); if (k < 2) xy[k] = v; ++k; } while (r.eat(','));
                                                r.expect(']');
                                            }
                                            if (nring == 0) {
                                                // point[0], point[1] (utils.rs:24-31) panic on shorter positions
                                                if (k < 2) r.fail("a position of ring 0 has fewer than two numbers");
                                                out.xy.push_back(xy[0]);
                                                out.xy.push_back(xy[1]);
                                                ++nv;
                                            }
                                        } while (r.eat(','));
                                        r.expect(']');
                                    }
                                    ++nring;
                                } while (r.eat(','));
                                r.expect(']');
                            }
                            if (nring == 0) r.fail("feature without a ring (coordinates[0], src/utils.rs:55)");
                            out.nverts.push_back(nv);
                            out.rings.push_back(nring);
                        } else {
                            r.skip_value();
                        }
                    } while (r.eat(','));
                    r.expect('}');
                }
                if (!have_type) r.fail("missing field `type`");
                if (!have_coords) r.fail("missing field `coordinates`");
            } else {
                r.skip_value();
            }
        } while (r.eat(','));
        r.expect('}');
    }
    if (!have_bbox) r.fail("missing field `bbox`");          // geojson.rs:18 (not an Option)
    if (!have_geom) r.fail("missing field `geometry`");
}

struct Chunk {
    size_t b = 0, e = 0;
    uint64_t quotes = 0;
    bool in_str = false;
    int64_t delta[2] = {0, 0};   // bracket depth change if the chunk starts outside / inside a string
    int64_t depth = 0;
    std::vector<size_t> starts, closes;
};

template <typename F>
void parallel_for(int threads, F&& f) {
    if (threads <= 1) { f(0); return; }
    std::vector<std::thread> th;
    th.reserve(threads);
    for (int t = 0; t < threads; ++t) th.emplace_back([&f, t] { f(t); });
    for (auto& x : th) x.join();
}

// 64 bytes -> one bit per byte: quotes, backslashes, opening and closing brackets
struct Bits {
    uint64_t quote, bslash, open, close;
};
#if defined(__SSE2__)
inline Bits classify64(const char* p) {
    const __m128i q = _mm_set1_epi8('"'), bs = _mm_set1_epi8('\\'), lo = _mm_set1_epi8(0x20);
    const __m128i ob = _mm_set1_epi8('{'), cb = _mm_set1_epi8('}');   // '[' | 0x20 == '{', ']' | 0x20 == '}'
    Bits r{0, 0, 0, 0};
    for (int k = 0; k < 4; ++k) {
        const __m128i x = _mm_loadu_si128((const __m128i*)(p + 16 * k));
        const __m128i xl = _mm_or_si128(x, lo);
        r.quote |= (uint64_t)(uint16_t)_mm_movemask_epi8(_mm_cmpeq_epi8(x, q)) << (16 * k);
        r.bslash |= (uint64_t)(uint16_t)_mm_movemask_epi8(_mm_cmpeq_epi8(x, bs)) << (16 * k);
        r.open |= (uint64_t)(uint16_t)_mm_movemask_epi8(_mm_cmpeq_epi8(xl, ob)) << (16 * k);
        r.close |= (uint64_t)(uint16_t)_mm_movemask_epi8(_mm_cmpeq_epi8(xl, cb)) << (16 * k);
    }
    return r;
}
#else
inline Bits classify64(const char* p) {
    Bits r{0, 0, 0, 0};
    for (int k = 0; k < 64; ++k) {
        const char c = p[k];
        r.quote |= (uint64_t)(c == '"') << k;
        r.bslash |= (uint64_t)(c == '\\') << k;
        r.open |= (uint64_t)(c == '{' || c == '[') << k;
        r.close |= (uint64_t)(c == '}' || c == ']') << k;
    }
    return r;
}
#endif
inline uint64_t prefix_xor(uint64_t m) {
    m ^= m << 1; m ^= m << 2; m ^= m << 4; m ^= m << 8; m ^= m << 16; m ^= m << 32;
    return m;
}

// Pass AB: unescaped-quote count, and the depth change under both hypotheses about the chunk start.
// `s` = "inside a string" assuming the chunk started outside one; the other hypothesis is its complement.
void scan_quotes_depth(const char* t, Chunk& c) {
    uint64_t quotes = 0;
    int64_t d0 = 0, d1 = 0;
    bool s = false;
    size_t p = c.b;
    auto scalar_until = [&](size_t stop) {   // a backslash always skips one byte, like pass C
        while (p < stop) {
            const char ch = t[p];
            if (ch == '\\') { p += 2; continue; }
            if (ch == '"') { ++quotes; s = !s; }
            else if (ch == '{' || ch == '[') { if (s) ++d1; else ++d0; }
            else if (ch == '}' || ch == ']') { if (s) --d1; else --d0; }
            ++p;
        }
    };
    while (p + 64 <= c.e) {
        const Bits m = classify64(t + p);
        if (m.bslash) { scalar_until(p + 64); continue; }
        const uint64_t S = prefix_xor(m.quote) ^ (s ? ~0ULL : 0ULL);
        quotes += (uint64_t)__builtin_popcountll(m.quote);
        d0 += __builtin_popcountll(m.open & ~S) - __builtin_popcountll(m.close & ~S);
        d1 += __builtin_popcountll(m.open & S) - __builtin_popcountll(m.close & S);
        s = S >> 63;
        p += 64;
    }
    scalar_until(c.e);
    c.quotes = quotes;
    c.delta[0] = d0;
    c.delta[1] = d1;
}

// Pass C: with the string state and the depth at the chunk start known, record every '{' met at depth 2
// (a feature of the top-level array) and every ']' that returns to depth 1.
void scan_starts(const char* t, Chunk& c) {
    bool s = c.in_str;
    int64_t d = c.depth;
    size_t p = c.b;
    auto scalar_until = [&](size_t stop) {
        while (p < stop) {
            const char ch = t[p];
            if (ch == '\\') { p += 2; continue; }
            if (ch == '"') s = !s;
            else if (!s) {
                if (ch == '{' || ch == '[') {
                    if (ch == '{' && d == 2) c.starts.push_back(p);
                    ++d;
                } else if (ch == '}' || ch == ']') {
                    --d;
                    if (ch == ']' && d == 1) c.closes.push_back(p);
                }
            }
            ++p;
        }
    };
    while (p + 64 <= c.e) {
        const Bits m = classify64(t + p);
        if (m.bslash) { scalar_until(p + 64); continue; }
        const uint64_t S = prefix_xor(m.quote) ^ (s ? ~0ULL : 0ULL);
        const uint64_t op = m.open & ~S, cl = m.close & ~S;
        const int nop = __builtin_popcountll(op), ncl = __builtin_popcountll(cl);
        if (d - ncl > 2) {
            d += nop - ncl;            // the depth stays above 2 inside the block: nothing to record
        } else {
            uint64_t ev = op | cl;
            while (ev) {
                const int k = __builtin_ctzll(ev);
                ev &= ev - 1;
                if ((op >> k) & 1) {
                    if (d == 2 && t[p + k] == '{') c.starts.push_back(p + k);
                    ++d;
                } else {
                    --d;
                    if (d == 1 && t[p + k] == ']') c.closes.push_back(p + k);
                }
            }
        }
        s = S >> 63;
        p += 64;
    }
    scalar_until(c.e);
}

}  // namespace

extern "C" {

int nfx_parse_f32(const char* token, int32_t len, float* out) {
    if (!token || len <= 0 || !out) return NFX_ERR_INVALID;
    init_pow10();
    const char* msg = nullptr;
    const char* q = parse_f32(token, token + len, out, &msg);
    if (!q || q != token + len) {
        nfx::g_thread_error = std::string("geojson: ") + (q ? "trailing characters" : msg);
        return NFX_ERR_INVALID;
    }
    return NFX_OK;
}

int nfx_geojson_parse(const char* text, int64_t len, int32_t threads, nfx_geojson** out) {
    if (!text || len < 0 || !out) return NFX_ERR_INVALID;
    *out = nullptr;
    init_pow10();
    if (threads <= 0) threads = (int32_t)std::max(1u, std::thread::hardware_concurrency());
    threads = std::min<int32_t>(threads, 256);
    try {
        // ---- top level up to the features array (FeatureCollection, geojson.rs:22-24) ---------------
        Reader top{text, text, text + len};
        top.expect('{');
        size_t arr_begin = 0;
        bool have_features = false;
        if (!top.eat('}')) {
            for (;;) {
                const char *kb, *ke;
                bool esc;
                top.raw_string(&kb, &ke, &esc);
                top.expect(':');
                if (!esc && top.key_is(kb, ke, "features")) {
                    have_features = true;
                    if (!top.eat('[')) top.fail("invalid type: expected a sequence");
                    arr_begin = (size_t)(top.p - text) - 1;
                    break;
                }
                top.skip_value();
                if (!top.eat(',')) { top.expect('}'); break; }
            }
        }
        if (!have_features) throw ParseError{"missing field `features`", (size_t)(top.p - text)};

        // ---- chunking; a chunk never starts right after a backslash, so no escape straddles a boundary ----
        const size_t n = (size_t)len;
        const size_t min_chunk = 1 << 16;
        int nchunk = (int)std::min<size_t>((size_t)threads * 4, std::max<size_t>(1, n / min_chunk));
        std::vector<Chunk> ch(nchunk);
        {
            size_t prev = 0;
            for (int k = 0; k < nchunk; ++k) {
                size_t e = k + 1 == nchunk ? n : std::max(prev, n * (size_t)(k + 1) / nchunk);
                while (e < n && e > 0 && text[e - 1] == '\\') ++e;
                ch[k].b = prev;
                ch[k].e = e;
                prev = e;
            }
        }
        auto over_chunks = [&](auto&& body) {
            parallel_for(threads, [&](int t) {
                for (int k = t; k < nchunk; k += threads) body(ch[k]);
            });
        };
        const bool timing = getenv("NFX_GEOJSON_TIMING") != nullptr;
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto lap = [&](const char* what, std::chrono::steady_clock::time_point& t0) {
            if (timing) fprintf(stderr, "[geojson] %-8s %.1f ms\n", what, std::chrono::duration<double, std::milli>(now() - t0).count());
            t0 = now();
        };
        auto t0 = now();
        // A+B: quote parity and depth deltas
        over_chunks([&](Chunk& c) { scan_quotes_depth(text, c); });
        {
            uint64_t q = 0;
            int64_t d = 0;
            for (auto& c : ch) { c.in_str = q & 1; c.depth = d; q += c.quotes; d += c.delta[c.in_str ? 1 : 0]; }
        }
        lap("scan", t0);
        // C: feature starts and array ends
        over_chunks([&](Chunk& c) { scan_starts(text, c); });
        size_t arr_end = n;
        bool closed = false;
        for (auto& c : ch) {
            for (size_t p : c.closes)
                if (p > arr_begin) { arr_end = p; closed = true; break; }
            if (closed) break;
        }
        if (!closed) throw ParseError{"EOF while parsing a list", n};
        std::vector<size_t> starts;
        {
            size_t total = 0;
            for (auto& c : ch) total += c.starts.size();
            starts.reserve(total);
            for (auto& c : ch)
                for (size_t p : c.starts)
                    if (p > arr_begin && p < arr_end) starts.push_back(p);
        }
        const size_t nf = starts.size();

        // the array must begin with the first feature (or be empty)
        {
            Reader r{text, text + arr_begin + 1, text + arr_end + 1};
            r.ws();
            const size_t first = (size_t)(r.p - text);
            if (nf == 0 ? first != arr_end : first != starts[0])
                throw ParseError{"invalid type: expected struct Feature", first};
        }
        lap("starts", t0);
        // ---- D: parse the features -----------------------------------------------------------------
        const int T = (int)std::min<size_t>((size_t)threads, std::max<size_t>(1, nf));
        std::vector<Piece> pieces(T);
        std::vector<ParseError> errs(T);
        std::vector<char> failed(T, 0);
        parallel_for(T, [&](int t) {
            const size_t lo = nf * (size_t)t / T, hi = nf * (size_t)(t + 1) / T;
            Piece& pc = pieces[t];
            if (hi > lo) {
                const size_t bytes = (hi < nf ? starts[hi] : arr_end) - starts[lo];
                pc.xy.reserve(bytes / 6);
                pc.bbox.reserve(4 * (hi - lo));
                pc.nverts.reserve(hi - lo);
                pc.rings.reserve(hi - lo);
            }
            try {
                Reader r{text, text, text + arr_end + 1};
                for (size_t i = lo; i < hi; ++i) {
                    r.p = text + starts[i];
                    parse_feature(r, pc);
                    r.ws();
                    const size_t next = i + 1 < nf ? starts[i + 1] : arr_end;
                    if (i + 1 < nf) {
                        if (!(r.p < r.end && *r.p == ',')) r.fail("expected `,` or `]`");
                        ++r.p;
                        r.ws();
                    }
                    if ((size_t)(r.p - text) != next) r.fail(i + 1 < nf ? "invalid type: expected struct Feature" : "expected `,` or `]`");
                }
            } catch (const ParseError& e) {
                errs[t] = e;
                failed[t] = 1;
            }
        });
        for (int t = 0; t < T; ++t)
            if (failed[t]) throw errs[t];   // the first error in input order, like a sequential reader

        lap("parse", t0);
        // ---- the rest of the top-level object ------------------------------------------------------
        top.p = text + arr_end + 1;
        while (top.eat(',')) {
            const char *kb, *ke;
            bool esc;
            top.raw_string(&kb, &ke, &esc);
            top.expect(':');
            if (!esc && top.key_is(kb, ke, "features")) top.fail("duplicate field `features`");
            top.skip_value();
        }
        top.expect('}');
        top.ws();
        if (top.p != top.end) top.fail("trailing characters");

        // ---- E: concatenate ------------------------------------------------------------------------
        auto* g = new nfx_geojson;
        std::vector<size_t> f0(T + 1, 0), v0(T + 1, 0);
        for (int t = 0; t < T; ++t) {
            f0[t + 1] = f0[t] + pieces[t].nverts.size();
            v0[t + 1] = v0[t] + pieces[t].xy.size() / 2;
        }
        g->off.resize(nf + 1);
        g->xy.resize(2 * v0[T]);
        g->bbox.resize(4 * nf);
        g->rings.resize(nf);
        parallel_for(T, [&](int t) {
            const Piece& pc = pieces[t];
            if (!pc.xy.empty()) memcpy(g->xy.data() + 2 * v0[t], pc.xy.data(), pc.xy.size() * sizeof(float));
            if (!pc.bbox.empty()) memcpy(g->bbox.data() + 4 * f0[t], pc.bbox.data(), pc.bbox.size() * sizeof(float));
            if (!pc.rings.empty()) memcpy(g->rings.data() + f0[t], pc.rings.data(), pc.rings.size() * sizeof(int32_t));
            int64_t o = (int64_t)v0[t];
            for (size_t i = 0; i < pc.nverts.size(); ++i) { g->off[f0[t] + i] = o; o += pc.nverts[i]; }
        });
        g->off[nf] = (int64_t)v0[T];
        lap("concat", t0);
        *out = g;
        return NFX_OK;
    } catch (const ParseError& e) {
        // line / column like serde_json's messages
        size_t line = 1, col = 0;
        const size_t lim = std::min<size_t>(e.pos, (size_t)len);
        for (size_t k = 0; k < lim; ++k) {
            if (text[k] == '\n') { ++line; col = 0; } else ++col;
        }
        nfx::g_thread_error = "geojson: " + e.msg + " at line " + std::to_string(line) + " column " + std::to_string(col + 1);
        return NFX_ERR_INVALID;
    } catch (const std::bad_alloc&) {
        nfx::g_thread_error = "geojson: out of host memory";
        return NFX_ERR_INVALID;
    }
}

int64_t nfx_geojson_count(const nfx_geojson* g) { return g ? (int64_t)g->off.size() - 1 : -1; }
int64_t nfx_geojson_vertices(const nfx_geojson* g) { return g ? (int64_t)g->xy.size() / 2 : -1; }
const float* nfx_geojson_xy(const nfx_geojson* g) { return g ? g->xy.data() : nullptr; }
const int64_t* nfx_geojson_offsets(const nfx_geojson* g) { return g ? g->off.data() : nullptr; }
const float* nfx_geojson_bbox(const nfx_geojson* g) { return g ? g->bbox.data() : nullptr; }
const int32_t* nfx_geojson_rings(const nfx_geojson* g) { return g ? g->rings.data() : nullptr; }
void nfx_geojson_free(nfx_geojson* g) { delete g; }

}  // extern "C"
