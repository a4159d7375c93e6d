// f32_display.h -- Rust `impl Display for f32` (what polars 0.32 writes into the `centroid` key,
// src/utils.rs:226-232, and into CSV cells, src/main.rs:163-166): the shortest decimal digits that read back
// to the same f32 (closest to the value when several exist), printed positionally -- never an exponent, no
// trailing ".0". The digit generation is Ulf Adams' Ryu (PLDI 2018) for binary32, written once for the host
// (schema.cpp) and for the device (csv.cu); the two power-of-five tables are passed in so that the device
// copy can live in global memory (per-thread indices would serialise in the constant bank).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define NFX_HD __host__ __device__ __forceinline__
#else
#define NFX_HD inline
#endif

// floor(2^(pow5bits(q)-1+59) / 5^q) + 1, q = 0..30
#define NFX_POW5_INV_SPLIT { \
    576460752303423489ull, 461168601842738791ull, 368934881474191033ull, 295147905179352826ull, \
    472236648286964522ull, 377789318629571618ull, 302231454903657294ull, 483570327845851670ull, \
    386856262276681336ull, 309485009821345069ull, 495176015714152110ull, 396140812571321688ull, \
    316912650057057351ull, 507060240091291761ull, 405648192073033409ull, 324518553658426727ull, \
    519229685853482763ull, 415383748682786211ull, 332306998946228969ull, 531691198313966350ull, \
    425352958651173080ull, 340282366920938464ull, 544451787073501542ull, 435561429658801234ull, \
    348449143727040987ull, 557518629963265579ull, 446014903970612463ull, 356811923176489971ull, \
    570899077082383953ull, 456719261665907162ull, 365375409332725730ull }
// 5^i scaled to 61 bits, i = 0..46
#define NFX_POW5_SPLIT { \
    1152921504606846976ull, 1441151880758558720ull, 1801439850948198400ull, 2251799813685248000ull, \
    1407374883553280000ull, 1759218604441600000ull, 2199023255552000000ull, 1374389534720000000ull, \
    1717986918400000000ull, 2147483648000000000ull, 1342177280000000000ull, 1677721600000000000ull, \
    2097152000000000000ull, 1310720000000000000ull, 1638400000000000000ull, 2048000000000000000ull, \
    1280000000000000000ull, 1600000000000000000ull, 2000000000000000000ull, 1250000000000000000ull, \
    1562500000000000000ull, 1953125000000000000ull, 1220703125000000000ull, 1525878906250000000ull, \
    1907348632812500000ull, 1192092895507812500ull, 1490116119384765625ull, 1862645149230957031ull, \
    1164153218269348144ull, 1455191522836685180ull, 1818989403545856475ull, 2273736754432320594ull, \
    1421085471520200371ull, 1776356839400250464ull, 2220446049250313080ull, 1387778780781445675ull, \
    1734723475976807094ull, 2168404344971008868ull, 1355252715606880542ull, 1694065894508600678ull, \
    2117582368135750847ull, 1323488980084844279ull, 1654361225106055349ull, 2067951531382569187ull, \
    1292469707114105741ull, 1615587133892632177ull, 2019483917365790221ull }

namespace nfx {

#define NFX_F32_MAX_CHARS 56   // "-0." + 37 zeros + 9 digits and some slack

struct Dec32 {
    uint32_t digits;   // 1..9 significant digits, no trailing zero (0 only for the value zero)
    int32_t exp10;     // value = digits * 10^exp10
    int32_t ndigits;
};

NFX_HD uint32_t ryu_pow5bits(int32_t e) { return (uint32_t)(((uint32_t)e * 1217359u) >> 19) + 1u; }
NFX_HD uint32_t ryu_log10pow2(int32_t e) { return ((uint32_t)e * 78913u) >> 18; }
NFX_HD uint32_t ryu_log10pow5(int32_t e) { return ((uint32_t)e * 732923u) >> 20; }
NFX_HD uint32_t ryu_mulshift(uint32_t m, uint64_t factor, int32_t shift) {   // (m * factor) >> shift, shift > 32
    const uint64_t b0 = (uint64_t)m * (uint32_t)factor;
    const uint64_t b1 = (uint64_t)m * (uint32_t)(factor >> 32);
    return (uint32_t)(((b0 >> 32) + b1) >> (shift - 32));
}
NFX_HD bool ryu_mult_pow5(uint32_t v, uint32_t p) {
    uint32_t c = 0;
    while (v != 0 && v % 5u == 0u) { v /= 5u; ++c; }
    return c >= p;
}

// Finite, non-zero magnitude given by its raw exponent and mantissa fields -> shortest decimal.
NFX_HD Dec32 ryu_f32(uint32_t ieee_mant, uint32_t ieee_exp, const uint64_t* pow5_inv, const uint64_t* pow5) {
    int32_t e2;
    uint32_t m2;
    if (ieee_exp == 0) { e2 = 1 - 127 - 23 - 2; m2 = ieee_mant; }
    else { e2 = (int32_t)ieee_exp - 127 - 23 - 2; m2 = (1u << 23) | ieee_mant; }
    const bool accept = (m2 & 1u) == 0u;
    const uint32_t mv = 4u * m2, mp = 4u * m2 + 2u;
    const uint32_t mm_shift = (ieee_mant != 0u || ieee_exp <= 1u) ? 1u : 0u;   // the gap below a power of two is half as wide
    const uint32_t mm = 4u * m2 - 1u - mm_shift;
    uint32_t vr, vp, vm;
    int32_t e10;
    bool vm_tz = false, vr_tz = false;
    uint32_t last = 0;
    if (e2 >= 0) {
        const uint32_t q = ryu_log10pow2(e2);
        e10 = (int32_t)q;
        const int32_t k = 59 + (int32_t)ryu_pow5bits((int32_t)q) - 1;
        const int32_t i = -e2 + (int32_t)q + k;
        vr = ryu_mulshift(mv, pow5_inv[q], i);
        vp = ryu_mulshift(mp, pow5_inv[q], i);
        vm = ryu_mulshift(mm, pow5_inv[q], i);
        if (q != 0 && (vp - 1u) / 10u <= vm / 10u) {
            const int32_t l = 59 + (int32_t)ryu_pow5bits((int32_t)q - 1) - 1;
            last = ryu_mulshift(mv, pow5_inv[q - 1], -e2 + (int32_t)q - 1 + l) % 10u;
        }
        if (q <= 9) {
            if (mv % 5u == 0u) vr_tz = ryu_mult_pow5(mv, q);
            else if (accept) vm_tz = ryu_mult_pow5(mm, q);
            else vp -= ryu_mult_pow5(mp, q) ? 1u : 0u;
        }
    } else {
        const uint32_t q = ryu_log10pow5(-e2);
        e10 = (int32_t)q + e2;
        const int32_t i = -e2 - (int32_t)q;
        const int32_t k = (int32_t)ryu_pow5bits(i) - 61;
        int32_t j = (int32_t)q - k;
        vr = ryu_mulshift(mv, pow5[i], j);
        vp = ryu_mulshift(mp, pow5[i], j);
        vm = ryu_mulshift(mm, pow5[i], j);
        if (q != 0 && (vp - 1u) / 10u <= vm / 10u) {
            j = (int32_t)q - 1 - ((int32_t)ryu_pow5bits(i + 1) - 61);
            last = ryu_mulshift(mv, pow5[i + 1], j) % 10u;
        }
        if (q <= 1) {
            vr_tz = true;
            if (accept) vm_tz = mm_shift == 1u;
            else --vp;
        } else if (q < 31) {
            vr_tz = (mv & ((1u << (q - 1)) - 1u)) == 0u;
        }
    }
    int32_t removed = 0;
    uint32_t out;
    if (vm_tz || vr_tz) {
        while (vp / 10u > vm / 10u) {
            vm_tz &= vm % 10u == 0u;
            vr_tz &= last == 0u;
            last = vr % 10u;
            vr /= 10u; vp /= 10u; vm /= 10u;
            ++removed;
        }
        if (vm_tz) {
            while (vm % 10u == 0u) {
                vr_tz &= last == 0u;
                last = vr % 10u;
                vr /= 10u; vp /= 10u; vm /= 10u;
                ++removed;
            }
        }
        if (vr_tz && last == 5u && vr % 2u == 0u) last = 4u;   // exactly half: round to even
        out = vr + (((vr == vm && (!accept || !vm_tz)) || last >= 5u) ? 1u : 0u);
    } else {
        while (vp / 10u > vm / 10u) {
            last = vr % 10u;
            vr /= 10u; vp /= 10u; vm /= 10u;
            ++removed;
        }
        out = vr + ((vr == vm || last >= 5u) ? 1u : 0u);
    }
    int32_t e = e10 + removed;
    while (out != 0u && out % 10u == 0u) { out /= 10u; ++e; }
    int32_t nd = 1;
    for (uint32_t t = out; t >= 10u; t /= 10u) ++nd;
    return Dec32{out, e, nd};
}

// Length of the positional text of a finite non-zero value (sign excluded).
NFX_HD int32_t positional_len(const Dec32& d) {
    const int32_t kk = d.ndigits + d.exp10;             // digits in front of the decimal point
    if (kk <= 0) return 2 - kk + d.ndigits;             // "0." zeros digits
    if (kk >= d.ndigits) return kk;                     // digits zeros
    return d.ndigits + 1;                               // dig.its
}

// Writes the positional text (sign excluded) at dst, returns its length.
NFX_HD int32_t positional_write(const Dec32& d, char* dst) {
    const int32_t kk = d.ndigits + d.exp10;
    char dig[10];
    uint32_t t = d.digits;
    for (int32_t k = d.ndigits - 1; k >= 0; --k) { dig[k] = (char)('0' + t % 10u); t /= 10u; }
    int32_t n = 0;
    if (kk <= 0) {
        dst[n++] = '0';
        dst[n++] = '.';
        for (int32_t k = 0; k < -kk; ++k) dst[n++] = '0';
        for (int32_t k = 0; k < d.ndigits; ++k) dst[n++] = dig[k];
    } else if (kk >= d.ndigits) {
        for (int32_t k = 0; k < d.ndigits; ++k) dst[n++] = dig[k];
        for (int32_t k = d.ndigits; k < kk; ++k) dst[n++] = '0';
    } else {
        for (int32_t k = 0; k < kk; ++k) dst[n++] = dig[k];
        dst[n++] = '.';
        for (int32_t k = kk; k < d.ndigits; ++k) dst[n++] = dig[k];
    }
    return n;
}

// One f32 analysed once: enough to know the length of its text and to write it.
struct Cell32 {
    Dec32 d;
    int32_t kind;   // 0 finite non-zero, 1 zero, 2 infinity, 3 NaN
    bool neg;
};
NFX_HD Cell32 f32_cell(uint32_t bits, const uint64_t* pow5_inv, const uint64_t* pow5) {
    Cell32 c;
    c.neg = (bits >> 31) != 0u;
    const uint32_t mant = bits & 0x7FFFFFu, ex = (bits >> 23) & 0xFFu;
    c.d = Dec32{0u, 0, 1};
    if (ex == 0xFFu) { c.kind = mant ? 3 : 2; return c; }
    if (ex == 0u && mant == 0u) { c.kind = 1; return c; }
    c.kind = 0;
    c.d = ryu_f32(mant, ex, pow5_inv, pow5);
    return c;
}
NFX_HD int32_t cell_len(const Cell32& c) {
    if (c.kind == 3) return 3;                                  // NaN (Rust prints no sign)
    const int32_t s = c.neg ? 1 : 0;
    if (c.kind == 2) return s + 3;                              // inf / -inf
    if (c.kind == 1) return s + 1;                              // 0 / -0
    return s + positional_len(c.d);
}
NFX_HD int32_t cell_write(const Cell32& c, char* dst) {
    if (c.kind == 3) { dst[0] = 'N'; dst[1] = 'a'; dst[2] = 'N'; return 3; }
    int32_t n = 0;
    if (c.neg) dst[n++] = '-';
    if (c.kind == 2) { dst[n++] = 'i'; dst[n++] = 'n'; dst[n++] = 'f'; return n; }
    if (c.kind == 1) { dst[n++] = '0'; return n; }
    return n + positional_write(c.d, dst + n);
}

// Rust Display of the f32 with these bits into dst (>= NFX_F32_MAX_CHARS bytes, not NUL-terminated); returns the length.
NFX_HD int32_t f32_display(uint32_t bits, char* dst, const uint64_t* pow5_inv, const uint64_t* pow5) {
    return cell_write(f32_cell(bits, pow5_inv, pow5), dst);
}
NFX_HD int32_t f32_display_len(uint32_t bits, const uint64_t* pow5_inv, const uint64_t* pow5) {
    return cell_len(f32_cell(bits, pow5_inv, pow5));
}

}  // namespace nfx
