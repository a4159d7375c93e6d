// tiff.cpp -- the container side of SURVEY.md 8f row 4. The reference reads `.svs` slides through OpenSlide
// (src/utils.rs:79-139, src/main.rs:20-35): an Aperio .svs is a TIFF / BigTIFF whose first directory is the full
// resolution image, cut into JPEG-compressed tiles that share one JPEGTables blob. This file parses that container
// (baseline TIFF 6.0 + the BigTIFF extension, either byte order, tiles or strips) into a flat table of compressed
// blocks; slide_decode.cu hands the blocks to nvJPEG. Host code, no dependency.
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/nfx.h"
#include "nfx_host.h"

namespace nfx {

namespace {

struct Reader {
    const uint8_t* p;
    uint64_t n;
    bool be = false, big = false;
    bool ok(uint64_t off, uint64_t len) const { return off <= n && len <= n - off; }
    uint64_t u(uint64_t off, int bytes) const {   // caller checked the range
        uint64_t v = 0;
        if (be) for (int k = 0; k < bytes; ++k) v = (v << 8) | p[off + k];
        else for (int k = bytes - 1; k >= 0; --k) v = (v << 8) | p[off + k];
        return v;
    }
};

int type_size(int t) {
    switch (t) {
        case 1: case 2: case 6: case 7: return 1;   // BYTE ASCII SBYTE UNDEFINED
        case 3: case 8: return 2;                   // SHORT SSHORT
        case 4: case 9: case 11: case 13: return 4; // LONG SLONG FLOAT IFD
        case 5: case 10: case 12: case 16: case 17: case 18: return 8;   // RATIONAL SRATIONAL DOUBLE LONG8 SLONG8 IFD8
        default: return 0;
    }
}

struct Entry {
    int tag = 0, type = 0;
    uint64_t count = 0, data_off = 0;   // data_off: where the values live (inline or pointed to)
};

bool entry_values(const Reader& r, const Entry& e, std::vector<uint64_t>& out, std::string& err) {
    const int ts = type_size(e.type);
    if (ts == 0 || ts > 8 || e.type == 5 || e.type == 10 || e.type == 11 || e.type == 12) { err = "unsupported TIFF field type"; return false; }
    if (e.count > (1ull << 32) || !r.ok(e.data_off, e.count * ts)) { err = "TIFF field runs past the end of the file"; return false; }
    out.resize(e.count);
    for (uint64_t k = 0; k < e.count; ++k) out[k] = r.u(e.data_off + k * ts, ts);
    return true;
}

}  // namespace

// Parses directory 0. On failure returns false with err set.
bool tiff_parse(const uint8_t* file, int64_t len, TiffLevel& L, std::string& err) {
    Reader r{file, (uint64_t)(len < 0 ? 0 : len)};
    if (!file || len < 16) { err = "not a TIFF file (too short)"; return false; }
    if (file[0] == 'I' && file[1] == 'I') r.be = false;
    else if (file[0] == 'M' && file[1] == 'M') r.be = true;
    else { err = "not a TIFF file (byte-order mark)"; return false; }
    const uint64_t magic = r.u(2, 2);
    uint64_t ifd = 0;
    if (magic == 42) { r.big = false; ifd = r.u(4, 4); }
    else if (magic == 43) {
        r.big = true;
        if (r.u(4, 2) != 8 || r.u(6, 2) != 0) { err = "unsupported BigTIFF offset size"; return false; }
        ifd = r.u(8, 8);
    } else { err = "not a TIFF file (magic)"; return false; }
    const int cnt_b = r.big ? 8 : 2, ent_b = r.big ? 20 : 12, val_b = r.big ? 8 : 4;
    if (!r.ok(ifd, cnt_b)) { err = "TIFF directory offset out of range"; return false; }
    const uint64_t nent = r.u(ifd, cnt_b);
    if (nent > 4096 || !r.ok(ifd + cnt_b, nent * ent_b)) { err = "TIFF directory runs past the end of the file"; return false; }

    L = TiffLevel();
    std::vector<uint64_t> v, offs, cnts;
    bool tiled = false, have_offs = false, have_cnts = false;
    uint64_t rows_per_strip = 0;
    for (uint64_t k = 0; k < nent; ++k) {
        const uint64_t at = ifd + cnt_b + k * ent_b;
        Entry e;
        e.tag = (int)r.u(at, 2);
        e.type = (int)r.u(at + 2, 2);
        e.count = r.u(at + 4, val_b);
        const int ts = type_size(e.type);
        const uint64_t bytes = e.count * (uint64_t)(ts ? ts : 1);
        e.data_off = bytes <= (uint64_t)val_b ? at + 4 + val_b : r.u(at + 4 + val_b, val_b);
        auto scalar = [&](uint64_t& dst) -> bool {
            if (!entry_values(r, e, v, err) || v.empty()) { if (err.empty()) err = "empty TIFF field"; return false; }
            dst = v[0];
            return true;
        };
        uint64_t t = 0;
        switch (e.tag) {
            case 256: if (!scalar(t)) return false; L.width = (int64_t)t; break;
            case 257: if (!scalar(t)) return false; L.height = (int64_t)t; break;
            case 258:
                if (!entry_values(r, e, v, err)) return false;
                for (uint64_t b : v) if (b != 8) { err = "only 8 bits per sample are supported"; return false; }
                break;
            case 259: if (!scalar(t)) return false; L.compression = (int)t; break;
            case 262: if (!scalar(t)) return false; L.photometric = (int)t; break;
            case 277: if (!scalar(t)) return false; L.samples = (int)t; break;
            case 278: if (!scalar(t)) return false; rows_per_strip = t; break;
            case 284: if (!scalar(t)) return false; if (t != 1) { err = "planar TIFF layout is not supported"; return false; } break;
            case 322: if (!scalar(t)) return false; L.block_w = (int)t; tiled = true; break;
            case 323: if (!scalar(t)) return false; L.block_h = (int)t; tiled = true; break;
            case 273: case 324: if (!entry_values(r, e, offs, err)) return false; have_offs = true; if (e.tag == 324) tiled = true; break;
            case 279: case 325: if (!entry_values(r, e, cnts, err)) return false; have_cnts = true; break;
            case 347:
                if (!r.ok(e.data_off, e.count)) { err = "JPEGTables run past the end of the file"; return false; }
                L.jpeg_tables.assign(file + e.data_off, file + e.data_off + e.count);
                break;
            default: break;
        }
    }
    if (L.width <= 0 || L.height <= 0) { err = "TIFF directory has no image size"; return false; }
    if (L.samples != 3) { err = "only 3-sample (RGB / YCbCr) images are supported"; return false; }
    if (!have_offs || !have_cnts || offs.size() != cnts.size()) { err = "TIFF directory has no consistent block table"; return false; }
    if (!tiled) {   // strips are full-width blocks
        L.block_w = (int)L.width;
        L.block_h = (int)(rows_per_strip == 0 || rows_per_strip > (uint64_t)L.height ? L.height : rows_per_strip);
    }
    if (L.block_w <= 0 || L.block_h <= 0) { err = "bad TIFF block size"; return false; }
    L.across = (L.width + L.block_w - 1) / L.block_w;
    L.down = (L.height + L.block_h - 1) / L.block_h;
    if ((int64_t)offs.size() != L.across * L.down) { err = "TIFF block table does not match the image size"; return false; }
    L.offsets.resize(offs.size());
    L.counts.resize(offs.size());
    for (size_t k = 0; k < offs.size(); ++k) {
        if (!r.ok(offs[k], cnts[k])) { err = "TIFF block runs past the end of the file"; return false; }
        L.offsets[k] = (int64_t)offs[k];
        L.counts[k] = (int64_t)cnts[k];
    }
    return true;
}

}  // namespace nfx

extern "C" int nfx_tiff_info(const uint8_t* file, int64_t len, nfx_tiff_level* out) {
    if (!out) return NFX_ERR_INVALID;
    nfx::TiffLevel L;
    std::string err;
    if (!nfx::tiff_parse(file, len, L, err)) {
        nfx::g_thread_error = "tiff: " + err;
        return NFX_ERR_INVALID;
    }
    out->width = L.width;
    out->height = L.height;
    out->block_width = L.block_w;
    out->block_height = L.block_h;
    out->blocks = (int64_t)L.offsets.size();
    out->compression = L.compression;
    out->photometric = L.photometric;
    out->jpeg_tables_bytes = (int32_t)L.jpeg_tables.size();
    return NFX_OK;
}
