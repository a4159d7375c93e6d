// geom.cu -- north-star kernels (2) polygon-to-mask rasterisation and (4) shape descriptors.
//
// One CTA (64 threads) per nucleus. Replaces, per nucleus:
//   preprocess_polygon                       src/utils.rs:54-74     (sequential f32 centroid, centring)
//   tch_utils::shapes::polygon               src/utils.rs:152-157   (oracle/SPEC.md B1, bit-exact)
//   patch window origin                      src/utils.rs:159-162   (f32, trunc toward zero)
//   center_of_mass / covariance / linalg_eig src/features/shape.rs:141-226 (closed-form slanv2)
//   tch_utils::shapes::ellipse + deviation   src/features/shape.rs:80-87, 209-217 (SPEC.md B2)
//   geometric_features::* and convex hull    src/features/shape.rs:89-97  (SPEC.md B7, B8)
//
// Rasterisation is edge-parallel scanline: each thread takes one polygon edge, walks only the rows
// that edge spans, evaluates the crossing abscissa X with exactly the float64 operations of the
// per-pixel rule, and XORs the prefix mask {c : c - P/2 < X} into the row's bit words. The result is
// identical to evaluating the even-odd rule at every pixel, at O(perimeter) instead of O(P*P*V) cost.
#include <math_constants.h>

#include "nfx_kernels.h"

namespace nfx {

namespace {

// One WARP per nucleus, four nuclei per CTA, warp-level synchronisation only -- for the raster-only and the shape variant
// alike: a ring has ~30 edges (a second warp only doubles the issue slots of the serial parts), one-warp CTAs are capped at
// 32 resident warps per SM by the CTA limit, and the kernel is latency bound (serial centroid fold, f64 divides, the hull's
// monotone chains). The shape variant ran two warps per nucleus with block barriers in round 1: its two hull chains now run
// on lanes 0 and 1 of the one warp (one instruction stream instead of two), every block reduction is a shuffle tree.
template <bool SHAPE> struct GeomCfg {
    static constexpr int kThreads = 32;   // threads per nucleus
    static constexpr int kNpc = 4;        // nuclei per CTA
};

// first index k in [0,P] such that (k - half) >= v   (exact; v may be any double). half = P/2 - sample offset.
__device__ __forceinline__ int first_index_geq(double v, int P, double half) {
    if (!(v == v)) return P;
    const double t = v + half;
    int k = (t <= 0.0) ? 0 : (t >= (double)P ? P : (int)ceil(t));
    while (k > 0 && ((double)(k - 1) - half) >= v) --k;
    while (k < P && ((double)k - half) < v) ++k;
    return k;
}

__device__ __forceinline__ uint32_t prefix_bits(int nbits) {   // nbits clamped to [0,32]
    return nbits >= 32 ? 0xffffffffu : (nbits <= 0 ? 0u : ((1u << nbits) - 1u));
}

// SPEC.md B2, one IEEE operation at a time.
__device__ __forceinline__ bool ellipse_inside(double x, double y, double cx, double cy, double cs,
                                               double sn, double a, double b) {
    const double dx = __dsub_rn(x, cx), dy = __dsub_rn(y, cy);
    const double xr = __dadd_rn(__dmul_rn(dx, cs), __dmul_rn(dy, sn));
    const double yr = __dsub_rn(__dmul_rn(dy, cs), __dmul_rn(dx, sn));
    const double u = __ddiv_rn(xr, a), v = __ddiv_rn(yr, b);
    const double val = __dadd_rn(__dmul_rn(u, u), __dmul_rn(v, v));
    return val <= 1.0;
}

// ---- SPEC.md B1 raster, executed by ONE warp ---------------------------------------------------------------------------
// Round 1 gave each lane one edge and let it walk that edge's rows (0 to ~10 of them, a float64 division per row): 13 of 32
// lanes were active on average (ncu). Now the (edge, row) pairs of a block of 64 edges are flattened over the lanes: every
// lane sets up at most two edges (row range, and ONE float64 reciprocal of dy), a warp scan of the row counts numbers the
// pairs, and lane l takes pairs l, l + 32, ... (binary search of the pair number in the scanned counts).
// Bit-exactness: the rule is X = xi + ((y - yi) * dx) / dy with every operation rounded to float64, and pixel c is toggled
// iff c - P/2 < X. A pair first evaluates X' = xi + ((y - yi) * dx) * (1 / dy), which is within a few ulp of X; only when
// X' + P/2 lies within 1e-9 (relative) of an integer -- where the two could fall on different sides of a pixel abscissa --
// is the division carried out. Either way the toggled prefix is the one the per-pixel rule gives.
constexpr int kEdgeBlock = 64;
struct EdgeRecs {               // per nucleus, in shared memory
    double xi[kEdgeBlock], yi[kEdgeBlock], dx[kEdgeBlock], dy[kEdgeBlock], rdy[kEdgeBlock];
    int r0[kEdgeBlock];
    int off[kEdgeBlock + 1];    // exclusive scan of the row counts
};

__device__ __forceinline__ void raster_warp(const float2* pts, int V, int P, double half, uint32_t* rows, EdgeRecs* er, int lane) {
    const int wpr = mask_wpr(P);
    for (int e0 = 0; e0 < V; e0 += kEdgeBlock) {
        int cnt[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int j = lane + 32 * u, k = e0 + j;
            cnt[u] = 0;
            if (k < V) {   // warp-uniform for u = 1 when the block holds <= 32 edges, divergent only in the last block
                const float2 a = pts[k], b = pts[(k + 1 == V) ? 0 : k + 1];
                if (!(a.y == b.y || !(a.y == a.y) || !(b.y == b.y))) {
                    const double ylo = fmin((double)a.y, (double)b.y), yhi = fmax((double)a.y, (double)b.y);
                    const int r0 = first_index_geq(ylo, P, half), r1 = first_index_geq(yhi, P, half);
                    if (r1 > r0) {
                        const double dye = __dsub_rn((double)b.y, (double)a.y);
                        er->xi[j] = (double)a.x;
                        er->yi[j] = (double)a.y;
                        er->dx[j] = __dsub_rn((double)b.x, (double)a.x);
                        er->dy[j] = dye;
                        er->rdy[j] = __ddiv_rn(1.0, dye);
                        er->r0[j] = r0;
                        cnt[u] = r1 - r0;
                    }
                }
            }
        }
        // exclusive scan over the 64 counts (lane l owns entries l and l + 32)
        int i0 = cnt[0], i1 = cnt[1];
#pragma unroll
        for (int o2 = 1; o2 < 32; o2 <<= 1) {
            const int t0 = __shfl_up_sync(0xffffffffu, i0, o2), t1 = __shfl_up_sync(0xffffffffu, i1, o2);
            if (lane >= o2) { i0 += t0; i1 += t1; }
        }
        const int tot0 = __shfl_sync(0xffffffffu, i0, 31), T = tot0 + __shfl_sync(0xffffffffu, i1, 31);
        er->off[lane] = i0 - cnt[0];
        er->off[lane + 32] = tot0 + i1 - cnt[1];
        if (lane == 0) er->off[kEdgeBlock] = T;
        __syncwarp();
        for (int t = lane; t < T; t += 32) {
            int j = 0;   // largest j with off[j] <= t and a non-empty range: off[j] <= t < off[j + 1]
#pragma unroll
            for (int step = kEdgeBlock / 2; step > 0; step >>= 1)
                if (er->off[j + step] <= t) j += step;
            const int r = er->r0[j] + (t - er->off[j]);
            const double y = (double)r - half;
            const double xi = er->xi[j];
            const double num = __dmul_rn(__dsub_rn(y, er->yi[j]), er->dx[j]);
            double X = __dadd_rn(xi, __dmul_rn(num, er->rdy[j]));
            const double tt = X + half, fr = tt - rint(tt);
            int nb;                                           // pixels c < nb satisfy (c - half) < X
            if (fabs(fr) > 1e-9 * fmax(1.0, fabs(tt))) {
                // X' is further than its own error from every pixel abscissa, and so is the rounding of X' + half: the first
                // index at or above X is ceil(X' + half), clamped to the row
                nb = tt <= 0.0 ? 0 : (tt >= (double)P ? P : (int)ceil(tt));
            } else {   // next to a pixel abscissa (or not finite): the rule's own division and the exact search
                X = __dadd_rn(xi, __ddiv_rn(num, er->dy[j]));
                if (!(X == X)) continue;                      // `x < NaN` is false for every pixel: nothing toggled
                nb = first_index_geq(X, P, half);
            }
            for (int w = 0; w < wpr; ++w) {
                const uint32_t m = prefix_bits(nb - 32 * w);
                if (m) atomicXor(&rows[r * wpr + w], m);
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ double cross3(double2 o, double2 a, double2 b) {
    return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x);
}

// bytes of one nucleus' shared memory before the edge records (rows | pts | sorted | stk), 16-byte aligned
__host__ __device__ __forceinline__ size_t geom_edge_offset(int P, int vmax, bool shape) {
    size_t b = (size_t)((P * mask_wpr(P) + 3) & ~3) * 4 + (size_t)((vmax + 1) & ~1) * 8;
    if (shape) b += (size_t)vmax * 16 + (size_t)vmax * 2 * 4;
    return (b + 15) & ~(size_t)15;
}

__host__ __device__ __forceinline__ size_t ring_bytes(int vmax) {
    return (size_t)((vmax + 1) & ~1) * 8 + (size_t)vmax * 16 + (((size_t)vmax * 2 * 4 + 15) & ~(size_t)15);
}

// GIANT = false: every ring of the launch lives in shared memory (nuclei whose ring is longer are skipped: the second launch
// takes them). GIANT = true: grid = the long rings only (p.giant_list[slot] = nucleus), work arrays in the HBM slot. Two
// instantiations instead of one pointer that may be either: ring accesses stay LDS on the common path (generic loads on the
// hull's serial chain had cost the shape kernel 6 %).
template <bool RASTER, bool SHAPE, bool GIANT>
__global__ void __launch_bounds__(GeomCfg<SHAPE>::kThreads * GeomCfg<SHAPE>::kNpc) k_geom(const GeomParams p, const int nuc_smem) {
    constexpr int kGeomThreads = GeomCfg<SHAPE>::kThreads, kNpc = GeomCfg<SHAPE>::kNpc;
    extern __shared__ __align__(16) unsigned char smem_all[];
    const int sub = kNpc > 1 ? (int)(threadIdx.x / kGeomThreads) : 0;            // which nucleus of the CTA
    unsigned char* smem_raw = smem_all + (size_t)sub * nuc_smem;
    const int P = p.P, wpr = mask_wpr(P), tid = kNpc > 1 ? (int)(threadIdx.x % kGeomThreads) : (int)threadIdx.x;
    const int64_t slot = (int64_t)blockIdx.x * kNpc + sub;
    if (kNpc > 1 && slot >= (GIANT ? p.n_giant : p.n)) return;   // whole warp; the multi-nucleus variant only uses warp-level barriers
    const int64_t i = GIANT ? (int64_t)p.giant_list[slot] : slot;
    auto sync = [] { if (kNpc > 1) __syncwarp(); else __syncthreads(); };
    const int64_t o0 = p.poly_off[i];
    const int V = (int)(p.poly_off[i + 1] - o0);
    if (!GIANT && V > p.vsmem) return;   // whole nucleus (warp or CTA): the GIANT launch computes it

    // shared layout: rows[P*wpr] u32 | pts[cap] float2 | sorted[cap] double2 | stk[2*cap] int | edge records.
    // A ring longer than the shared-memory capacity (kGeomRingSmem vertices) keeps pts / sorted / stk in its HBM slot.
    uint32_t* rows = reinterpret_cast<uint32_t*>(smem_raw);
    const int cap = GIANT ? p.vmax : p.vsmem;
    float2* pts = GIANT ? reinterpret_cast<float2*>(p.ring_scratch + (size_t)slot * ring_bytes(p.vmax))
                        : reinterpret_cast<float2*>(rows + ((P * wpr + 3) & ~3));
    double2* sorted = reinterpret_cast<double2*>(pts + ((cap + 1) & ~1));   // 16-byte aligned
    int* stk = reinterpret_cast<int*>(sorted + (SHAPE ? cap : 0));
    EdgeRecs* erecs = reinterpret_cast<EdgeRecs*>(smem_raw + geom_edge_offset(P, p.vsmem, SHAPE));   // RASTER only
    __shared__ float s_c_all[kNpc][2];
    __shared__ int s_hull_all[kNpc][2];
    int* s_hull = s_hull_all[sub];
    float* s_c = s_c_all[sub];

    for (int k = tid; k < V; k += kGeomThreads) pts[k] = p.poly_xy[o0 + k];
    if (RASTER)
        for (int k = tid; k < P * wpr; k += kGeomThreads) rows[k] = 0u;
    else
        for (int k = tid; k < P * wpr; k += kGeomThreads) rows[k] = p.bitmask[i * P * wpr + k];
    sync();

    if (RASTER) {
        if (tid == 0) {
            // utils.rs:56-64 -- sequential f32 fold in stored order, then one division each
            float ax = 0.f, ay = 0.f;
            for (int k = 0; k < V; ++k) {
                ax = __fadd_rn(ax, pts[k].x);
                ay = __fadd_rn(ay, pts[k].y);
            }
            const float cx = __fdiv_rn(ax, (float)V), cy = __fdiv_rn(ay, (float)V);
            s_c[0] = cx;
            s_c[1] = cy;
            p.centroid[i] = make_float2(cx, cy);
            // utils.rs:159-162 -- `as i64` truncates toward zero and saturates (NaN -> 0)
            const float half = (float)P / 2.0f;
            auto cast = [](float v) -> long long { return (v == v) ? __float2ll_rz(v) : 0ll; };
            const long long top = cast(__fsub_rn(cy, half)), left = cast(__fsub_rn(cx, half));
            const long long bottom = cast(__fadd_rn(cy, half)), right = cast(__fadd_rn(cx, half));
            auto clamp32 = [](long long v) -> int {
                return (int)max(-(1ll << 30), min(1ll << 30, v));
            };
            NucInfo inf;
            if (p.slide_window) {
                // utils.rs:104-105 -- `as u32` saturates: negative and NaN origins read from 0 (the window is shifted, not
                // padded); OpenSlide always returns P x P pixels, black beyond the slide (= TMA zero fill)
                auto as_u32 = [](float v) -> long long { return (v == v && v > 0.f) ? min(__float2ll_rz(v), 4294967295ll) : 0ll; };
                inf.left = clamp32(as_u32(__fsub_rn(cx, half)) - p.tile_ox);
                inf.top = clamp32(as_u32(__fsub_rn(cy, half)) - p.tile_oy);
                inf.nvc = P;
                inf.nvr = P;
            } else {
                inf.left = clamp32(left - p.tile_ox);
                inf.top = clamp32(top - p.tile_oy);
                inf.nvc = (int)max(0ll, min((long long)P, right - left));
                inf.nvr = (int)max(0ll, min((long long)P, bottom - top));
            }
            p.info[i] = inf;
        }
        sync();
        const float cx = s_c[0], cy = s_c[1];
        for (int k = tid; k < V; k += kGeomThreads) {   // utils.rs:65-72
            float2 v = pts[k];
            v.x = __fsub_rn(v.x, cx);
            v.y = __fsub_rn(v.y, cy);
            pts[k] = v;
        }
        sync();

        // ---- SPEC.md B1: scanline raster over flattened (edge, row) pairs, one warp ----
        if (tid < 32) raster_warp(pts, V, P, 0.5 * (double)P - (double)p.sample_off, rows, erecs, tid);
        sync();
        for (int k = tid; k < P * wpr; k += kGeomThreads) p.bitmask[i * P * wpr + k] = rows[k];
    }

    if (!SHAPE) return;
    float* out = p.out + i * (int64_t)p.out_stride + p.col_shape;

    // ---- polygon scalars, SPEC.md B7 (float64 on the centred ring) ----
    double poly_area = 0.0;   // lane 0 keeps it for the hull's deviation
    {
        double v2[2] = {0.0, 0.0};   // shoelace sum, perimeter
        for (int k = tid; k < V; k += kGeomThreads) {
            const float2 a = pts[k], b = pts[(k + 1 == V) ? 0 : k + 1];
            v2[0] += (double)a.x * (double)b.y - (double)b.x * (double)a.y;
            const double ex = (double)b.x - (double)a.x, ey = (double)b.y - (double)a.y;
            v2[1] += sqrt(ex * ex + ey * ey);
        }
        v2[0] = warp_sum(v2[0]);
        v2[1] = warp_sum(v2[1]);
        if (tid == 0) {
            const double area = 0.5 * fabs(v2[0]), per = v2[1];
            out[0] = (float)area;
            out[5] = (float)per;
            out[6] = (float)(2.0 * sqrt(CUDART_PI * area));
            out[7] = (float)((4.0 * CUDART_PI * area) / (per * per));
            poly_area = area;
        }
    }
    // ---- convex hull, SPEC.md B8: rank sort + two monotone chains (threads 0 and 32) ----
    for (int k = tid; k < V; k += kGeomThreads) {
        const float2 a = pts[k];
        int rank = 0;
        for (int m = 0; m < V; ++m) {
            const float2 b = pts[m];
            rank += (b.x < a.x) || (b.x == a.x && (b.y < a.y || (b.y == a.y && m < k)));
        }
        sorted[rank] = make_double2((double)a.x, (double)a.y);   // converted once: the chain is latency bound
    }
    sync();
    if (tid < 2) {   // lower chain on lane 0, upper chain on lane 1
        int* S = stk + (tid == 0 ? 0 : cap);
        int sz = 0;
        for (int t = 0; t < V; ++t) {
            const int idx = (tid == 0) ? t : V - 1 - t;
            const double2 q = sorted[idx];
            while (sz >= 2 && cross3(sorted[S[sz - 2]], sorted[S[sz - 1]], q) <= 0.0) --sz;
            S[sz++] = idx;
        }
        s_hull[tid] = sz;
    }
    sync();
    if (tid == 0) {
        const int nl = max(s_hull[0] - 1, 0), nu = max(s_hull[1] - 1, 0), h = nl + nu;
        auto hp = [&](int k) -> double2 { return sorted[k < nl ? stk[k] : stk[cap + (k - nl)]]; };
        double sh = 0.0, per = 0.0;
        for (int k = 0; k < h; ++k) {
            const double2 a = hp(k), b = hp(k + 1 == h ? 0 : k + 1);
            sh += a.x * b.y - b.x * a.y;
            const double ex = b.x - a.x, ey = b.y - a.y;
            per += sqrt(ex * ex + ey * ey);
        }
        const double harea = (h >= 3) ? 0.5 * fabs(sh) : 0.0;
        if (h < 2) per = 0.0;
        out[9] = (float)harea;
        out[10] = (float)((harea - poly_area) / harea);
        out[11] = (float)per;
    }

    // ---- mask moments (shape.rs:149-157, 219-226): exact integer sums over set bits ----
    double mom[6] = {0, 0, 0, 0, 0, 0};   // K, Sr, Sc, Srr, Src, Scc
    for (int r = tid; r < P; r += kGeomThreads) {
        unsigned long long n = 0, sc = 0, scc = 0;
        for (int w = 0; w < wpr; ++w) {
            uint32_t bits = rows[r * wpr + w];
            while (bits) {   // one step per RUN of set bits [c0, c1): closed-form sums of c and c^2 (a blob row is one run)
                const int s0 = __ffs(bits) - 1;
                const uint32_t nt = ~(bits >> s0);
                const int e0 = s0 + (nt ? __ffs(nt) - 1 : 32);
                const unsigned c0 = 32u * w + s0, c1 = 32u * w + e0, len = c1 - c0;
                n += len;
                sc += (unsigned long long)(c0 + c1 - 1u) * len / 2u;
                // sum_{c < k} c^2 = (k - 1) k (2k - 1) / 6
                scc += ((unsigned long long)(c1 - 1u) * c1 * (2u * c1 - 1u) - (unsigned long long)(c0 ? (c0 - 1u) : 0u) * c0 * (2u * c0 - 1u)) / 6u;
                bits = (e0 >= 32) ? 0u : (bits & (0xffffffffu << e0));
            }
        }
        mom[0] += (double)n;
        mom[1] += (double)(n * r);
        mom[2] += (double)sc;
        mom[3] += (double)(n * r * r);
        mom[4] += (double)(sc * r);
        mom[5] += (double)scc;
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) mom[k] = warp_sum(mom[k]);   // sums of integers < 2^53: exact in any order

    // every thread derives the same scalars (cheap) so no extra broadcast is needed
    const double K = mom[0];
    const float nanf_ = CUDART_NAN_F;
    float major = nanf_, minor = nanf_, angle = nanf_;
    // center_of_mass: f32 sum of exact integers / K in f32 (ATen mean = sum then true division)
    const float mr = __fdiv_rn((float)mom[1], (float)K), mc = __fdiv_rn((float)mom[2], (float)K);
    if (K > 0.0) {
        const double er = mom[1] / K, ec = mom[2] / K;
        const float a = (float)(mom[3] / K - er * er);      // var rows
        const float b = (float)(mom[4] / K - er * ec);      // cov
        const float d = (float)(mom[5] / K - ec * ec);      // var cols
        // LAPACK sgeev on [[a,b],[b,d]] (slanv2 closed form, oracle eig2x2_lapack)
        float l0, l1, v00, v01, v10, v11;   // V columns are the eigenvectors
        if (b == 0.f) {
            l0 = a; l1 = d; v00 = 1.f; v01 = 0.f; v10 = 0.f; v11 = 1.f;
        } else {
            const float pp = 0.5f * (a - d);
            const float rr = hypotf(pp, b);
            const float z = pp + copysignf(rr, pp);
            l0 = d + z;
            l1 = d - (b / z) * b;
            const float tau = hypotf(b, z);
            const float cs = z / tau, sn = b / tau;
            v00 = cs; v01 = -sn; v10 = sn; v11 = cs;
        }
        float m0, m1;   // shape.rs:192-196: ROW 0 of the eigenvector matrix if l0 > l1 else row 1
        if (l0 > l1) { major = sqrtf(l0); minor = sqrtf(l1); m0 = v00; m1 = v01; }
        else         { major = sqrtf(l1); minor = sqrtf(l0); m0 = v10; m1 = v11; }
        angle = atan2f(m0, m1);
        major *= 2.0f;
        minor *= 2.0f;
    }
    if (tid == 0) {
        out[1] = major;
        out[2] = minor;
        const float M = major * 0.5f, m = minor * 0.5f;                 // shape.rs:205-207
        out[3] = __fdiv_rn(sqrtf(__fsub_rn(__fmul_rn(M, M), __fmul_rn(m, m))), M);
        out[4] = angle;
    }

    // ---- ellipse raster (SPEC.md B2) by exact row intervals + |mask - ellipse| ----
    const float halfP = (float)P / 2.0f;
    const double ecx = (double)__fsub_rn(mc, halfP), ecy = (double)__fsub_rn(mr, halfP);   // shape.rs:71-73,84
    const double ea = (double)major, eb = (double)minor, ang = (double)angle;
    const bool drawable = (ea > 0.0) && (eb > 0.0) && isfinite(ea) && isfinite(eb) && isfinite(ang) &&
                          isfinite(ecx) && isfinite(ecy);
    double cs = 0.0, sn = 0.0;
    if (drawable) sincos(ang, &sn, &cs);
    const double half = 0.5 * (double)P - (double)p.sample_off;   // pixel k samples k - half
    const double ia2 = 1.0 / (ea * ea), ib2 = 1.0 / (eb * eb);
    const double qa = cs * cs * ia2 + sn * sn * ib2;
    double diff[1] = {0.0};
    for (int r = tid; r < P; r += kGeomThreads) {
        int clo = 1, chi = 0;   // empty
        if (drawable) {
            const double y = (double)r - half, dy = y - ecy;
            const double qb = 2.0 * dy * cs * sn * (ia2 - ib2);
            const double qc = dy * dy * (sn * sn * ia2 + cs * cs * ib2) - 1.0;
            const double disc = qb * qb - 4.0 * qa * qc;
            if (disc >= -1e-6 * (qb * qb + 4.0 * fabs(qa * qc))) {
                const double sq = sqrt(fmax(disc, 0.0));
                const double x1 = (-qb - sq) / (2.0 * qa) + ecx, x2 = (-qb + sq) / (2.0 * qa) + ecx;
                clo = max(0, (int)fmin(fmax(ceil(x1 + half) - 1.0, -1.0), (double)P));
                chi = min(P - 1, (int)fmax(fmin(floor(x2 + half) + 1.0, (double)P), -1.0));
                while (clo <= chi && !ellipse_inside((double)clo - half, y, ecx, ecy, cs, sn, ea, eb)) ++clo;
                while (chi >= clo && !ellipse_inside((double)chi - half, y, ecx, ecy, cs, sn, ea, eb)) --chi;
                // grow if the estimate was one pixel short (cannot happen for a well-conditioned row)
                while (clo > 0 && clo <= chi && ellipse_inside((double)(clo - 1) - half, y, ecx, ecy, cs, sn, ea, eb)) --clo;
                while (chi < P - 1 && clo <= chi && ellipse_inside((double)(chi + 1) - half, y, ecx, ecy, cs, sn, ea, eb)) ++chi;
            }
        }
        for (int w = 0; w < wpr; ++w) {
            uint32_t e = 0u;
            if (clo <= chi) {
                const int lo = max(clo - 32 * w, 0), hi = min(chi - 32 * w, 31);
                if (lo <= hi) e = prefix_bits(hi + 1) & ~prefix_bits(lo);
            }
            diff[0] += (double)__popc(e ^ rows[r * wpr + w]);
            if (p.ellipse_bits) p.ellipse_bits[i * P * wpr + r * wpr + w] = e;
        }
    }
    diff[0] = warp_sum(diff[0]);
    if (tid == 0) out[8] = __fdiv_rn((float)diff[0], (float)K);   // shape.rs:209-217 (f32 tensor / scalar)
}

}  // namespace

size_t geom_ring_bytes(int vmax) {   // pts | sorted | stk, every part 16-byte aligned (= ring_bytes below)
    return (size_t)((vmax + 1) & ~1) * 8 + (size_t)vmax * 16 + (((size_t)vmax * 2 * 4 + 15) & ~(size_t)15);
}

cudaError_t launch_geom(const GeomParams& p, bool raster, bool shape, cudaStream_t s) {
    if (p.n <= 0) return cudaSuccess;
    size_t nuc = geom_edge_offset(p.P, p.vsmem, shape);
    if (raster) nuc += sizeof(EdgeRecs);
    nuc = (nuc + 15) & ~(size_t)15;
    auto go = [&](auto kern, int threads, int npc, int64_t count) -> cudaError_t {
        if (count <= 0) return cudaSuccess;
        const size_t smem = nuc * npc;
        if (smem > 32 * 1024) {   // static shared memory counts towards the 48 KB default limit too
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<(unsigned)((count + npc - 1) / npc), threads * npc, smem, s>>>(p, (int)nuc);
        return cudaGetLastError();
    };
    constexpr int T1 = GeomCfg<true>::kThreads, N1 = GeomCfg<true>::kNpc, T0 = GeomCfg<false>::kThreads, N0 = GeomCfg<false>::kNpc;
    cudaError_t e;
    if (raster) {
        e = shape ? go(k_geom<true, true, false>, T1, N1, p.n) : go(k_geom<true, false, false>, T0, N0, p.n);
        if (e == cudaSuccess && p.n_giant > 0)
            e = shape ? go(k_geom<true, true, true>, T1, N1, p.n_giant) : go(k_geom<true, false, true>, T0, N0, p.n_giant);
        return e;
    }
    if (!shape) return cudaSuccess;
    e = go(k_geom<false, true, false>, T1, N1, p.n);
    if (e == cudaSuccess && p.n_giant > 0) e = go(k_geom<false, true, true>, T1, N1, p.n_giant);
    return e;
}

}  // namespace nfx
