// exact.cuh -- the reference-order FP64 evaluation of every interpolation method, per query.
//
// This is the "exact path": one query evaluated with the reference CPU class's operation order
// (GridH.cpp, cited per function), every FP64 operation spelled with a round-to-nearest intrinsic so
// that nvcc cannot contract a*b+c into an FMA.  The reference's discrete decisions -- floor/round
// centre, candidate enumeration order, strict-'<' tie breaks -- are decided by last-bit FP64 noise
// (SURVEY.md section 0, facts 3-4), so they are only reproducible this way.  Bilinear, bicubic and
// all neighbour selections come out bit-identical to the CPU reference; kriging -- a tolerance method -- is
// evaluated with fused multiply-adds and an expm1 form of the variogram: <1e-9 m from the reference.
//
// The tiled fast paths (upsample.cu, fill.cu) handle clean stencils in bulk and call into this file
// for every output whose footprint holds a NaN.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#ifndef AUVI_VARIOGRAM_INLINE
#define AUVI_VARIOGRAM_INLINE __forceinline__
#endif

namespace auvi {

enum Method : int { BILINEAR = 0, CUBIC = 1, KRIGING = 2, NN = 3, IDW = 4, BILINEAR_SEARCH = 5, IDW_KNN = 6 };
// (AUVI_KRIGING_FITTED = 7 is KRIGING evaluated with the grid's fitted variogram parameters: no method of its own here)

constexpr int kMaxRadius = 10;   // GridH.cpp:275, :339
constexpr int kMaxCand   = 45;   // 3 + 2*(2*10+1): the most the early-terminating search can hold

// Read-only view of the depth grid (row-major, row 0 = min_lat, GridD.cu:65-75).  A view may cover
// only rows [row0, row0+rows) of the global grid (multi-GPU slabs); all index logic uses the global
// dimensions, only the final address subtracts row0.
template <typename T>
struct GridView {
    const T* __restrict__ z;
    int n_lat, n_lon;            // global dimensions
    int64_t ld;                  // elements between consecutive rows (>= n_lon)
    int row0;                    // first global row held in z
    double min_lon, max_lon, min_lat, max_lat;
    double lon_step, lat_step;   // (max-min)/(n-1), GridH.cpp:156-157 / GridD.cu:52-53
    // matrix entry of the kriging system as a function of the separation h: nugget + sill * (1 - exp(-h * inv_range)).
    // The reference's constants are 1, 100, 1/10 (its variogram, GridH.cpp:371-376).  The opt-in AUVI_KRIGING_FITTED puts
    // the fitted model here in covariance form: entry = c1 * exp(-h / a) = c1 + (-c1) * (1 - exp(-h / a)), diagonal c0 + c1
    double vg_nugget, vg_sill, vg_inv_range;
    double vg_diag;              // diagonal of the kriging matrix: gamma(0) = the nugget in the reference (GridH.cpp:386-392)

    __device__ __forceinline__ double at(int j, int i) const {
        return static_cast<double>(__ldg(z + static_cast<int64_t>(j - row0) * ld + i));
    }
};

// ---- contraction-proof FP64 arithmetic ---------------------------------------------------------
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dsqrt(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000LL); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <typename T>
__device__ __forceinline__ bool outside(const GridView<T>& g, double lon, double lat) {
    return lon < g.min_lon || lon > g.max_lon || lat < g.min_lat || lat > g.max_lat;   // GridH.cpp:162
}
// Grid-space coordinate, GridH.cpp:167-168.
__device__ __forceinline__ double to_index_space(double c, double lo, double step) {
    return ddiv(dsub(c, lo), step);
}

// s / n for n in 1..4, bit-identical to the IEEE division the reference performs (`sum / count`, GridH.cpp:17), without the
// ~40-instruction FP64 division sequence: n = 1, 2, 4 are exact scalings (RN(s/4) and RN(s * 0.25) round the same real
// number); n = 3 is one Markstein correction step on the correctly rounded reciprocal -- q = RN(s * y), r = s - 3q (exact in
// an FMA), q' = RN(q + r * y) is the correctly rounded quotient (checked against the division on 3e8 values, incl. the
// neighbours of multiples of three).  Outside the exponent range where q or r could go subnormal: the division itself.
__device__ __forceinline__ double div_count(double s, int n) {
    if (n == 3) {
        const double y = 0.33333333333333331;                      // RN(1/3)
        const double q = dmul(s, y);
        const double r = fma(-3.0, q, s);
        const double a = fabs(s);
        return (a > 1e-280 && a < 1e300) ? fma(r, y, q) : ddiv(s, 3.0);
    }
    return dmul(s, n == 1 ? 1.0 : (n == 2 ? 0.5 : 0.25));
}

// Mean of the non-NaN members of four values, summed a,b,c,d in order (GridH.cpp:10-18).
__device__ __forceinline__ double mean_valid4(double a, double b, double c, double d) {
    double s = 0.0; int n = 0;
    if (!isnan(a)) { s = dadd(s, a); ++n; }
    if (!isnan(b)) { s = dadd(s, b); ++n; }
    if (!isnan(c)) { s = dadd(s, c); ++n; }
    if (!isnan(d)) { s = dadd(s, d); ++n; }
    return n ? div_count(s, n) : qnan();
}

// ---- bilinear, GridH.cpp:160-210 ----------------------------------------------------------------
template <typename T>
__device__ double exact_bilinear(const GridView<T>& g, double x, double y) {
    int x0 = static_cast<int>(floor(x)), y0 = static_cast<int>(floor(y));
    int x1 = min(x0 + 1, g.n_lon - 1), y1 = min(y0 + 1, g.n_lat - 1);
    double wx = dsub(x, static_cast<double>(x0)), wy = dsub(y, static_cast<double>(y0));
    double a = g.at(y0, x0), b = g.at(y0, x1), c = g.at(y1, x0), d = g.at(y1, x1);
    if (isnan(a) || isnan(b) || isnan(c) || isnan(d)) return mean_valid4(a, b, c, d);
    double ux = dsub(1.0, wx);
    double lo = dadd(dmul(ux, a), dmul(wx, b));
    double hi = dadd(dmul(ux, c), dmul(wx, d));
    return dadd(dmul(dsub(1.0, wy), lo), dmul(wy, hi));
}

// Catmull-Rom in the reference's polynomial form and summation order, GridH.cpp:215-217:
// 0.5*(2*p1 + (-p0+p2)*t + (2*p0-5*p1+4*p2-p3)*t*t + (-p0+3*p1-3*p2+p3)*t*t*t)
// The reference rounds after each of its 23 operations.  This form reaches the same bits in 19 (the tiled FP64 upsample is
// bound by the FP64 pipe), using only identities that hold for every finite input short of
// overflow / subnormal intermediates (no depth is within 2^-900 of either):
//   * 2*p0 and 4*p2 are exact, so RN(2*p0 - RN(5*p1)) = fma(2, p0, -RN(5*p1)) and RN(u + 4*p2) = fma(4, p2, u);
//   * halving commutes with rounding, so the final 0.5 * (...) moves into the terms: the last factor of each product is
//     th = t/2 (exact) and 2*p1 becomes p1 -- RN(RN(RN(p1 + lin/2) + quad/2) + cub/2) is half the reference's sum.
// Checked bit for bit against the literal expression on 2e8 inputs (signed zeros, t = 0, t within 1e-12 of 0 and 1).
__device__ __forceinline__ double catmull_rom_exact_h(double p0, double p1, double p2, double p3, double t, double th) {
    const double lin = dmul(dsub(p2, p0), th);                    // p2 - p0 == -p0 + p2, 3*p1 - p0 == -p0 + 3*p1: same reals, same roundings
    const double qc = dsub(__fma_rn(4.0, p2, __fma_rn(2.0, p0, -dmul(5.0, p1))), p3);
    const double quad = dmul(dmul(qc, t), th);
    const double cc = dadd(dsub(dsub(dmul(3.0, p1), p0), dmul(3.0, p2)), p3);
    const double cub = dmul(dmul(dmul(cc, t), t), th);
    return dadd(dadd(dadd(p1, lin), quad), cub);
}
// The same evaluation in two steps: the three coefficients depend on the taps only, so a caller that evaluates several t on one
// set of taps (the tiled upsample: f_lat output rows under one window of four input rows) forms them once.
struct CatmullCoef { double a, qc, cc; };
__device__ __forceinline__ CatmullCoef catmull_rom_coef(double p0, double p1, double p2, double p3) {
    CatmullCoef k;
    // p2 - p0 and 3*p1 - p0 are the reference's -p0 + p2 and -p0 + 3*p1 (the same real numbers, so the same roundings, signed
    // zeros included); spelled with the negation the compiler materialised -p0 with an FP64 add of its own (one operation in
    // eleven on a pipe-bound kernel)
    k.a = dsub(p2, p0);
    k.qc = dsub(__fma_rn(4.0, p2, __fma_rn(2.0, p0, -dmul(5.0, p1))), p3);
    k.cc = dadd(dsub(dsub(dmul(3.0, p1), p0), dmul(3.0, p2)), p3);
    return k;
}
__device__ __forceinline__ double catmull_rom_eval(const CatmullCoef& k, double p1, double t, double th) {
    const double lin = dmul(k.a, th);
    const double quad = dmul(dmul(k.qc, t), th);
    const double cub = dmul(dmul(dmul(k.cc, t), t), th);
    return dadd(dadd(dadd(p1, lin), quad), cub);
}
__device__ __forceinline__ double catmull_rom_exact(double p0, double p1, double p2, double p3, double t) {
    return catmull_rom_exact_h(p0, p1, p2, p3, t, dmul(0.5, t));
}

// ---- ring search, GridH.cpp:24-118 ---------------------------------------------------------------
// Candidates are kept as (distance, packed offset); values are re-read for the few that survive.
struct Cands {
    double d[kMaxCand];
    int16_t code[kMaxCand];      // (dj+10)*32 + (di+10), offsets from the search centre
    int n;
};
__device__ __forceinline__ int16_t pack_off(int di, int dj) { return static_cast<int16_t>((dj + 10) * 32 + (di + 10)); }
__device__ __forceinline__ int off_i(int16_t c) { return (c & 31) - 10; }
__device__ __forceinline__ int off_j(int16_t c) { return (c >> 5) - 10; }

template <typename T>
__device__ __forceinline__ void consider(const GridView<T>& g, int ci, int cj, int di, int dj,
                                         double x, double y, Cands& c) {
    int i = ci + di, j = cj + dj;
    if (isnan(g.at(j, i))) return;
    double ddi = dsub(dadd(static_cast<double>(i), 0.5), x);     // (i + 0.5) - x, :42-44
    double ddj = dsub(dadd(static_cast<double>(j), 0.5), y);
    c.d[c.n] = dsqrt(dadd(dmul(ddi, ddi), dmul(ddj, ddj)));
    c.code[c.n] = pack_off(di, dj);
    ++c.n;
}

template <typename T>
__device__ void ring_search(const GridView<T>& g, double x, double y, int ci, int cj, Cands& c) {
    c.n = 0;
    consider(g, ci, cj, 0, 0, x, y, c);                           // centre first
    for (int r = 1; r <= kMaxRadius; ++r) {
        const bool top_ok = cj - r >= 0, bot_ok = cj + r < g.n_lat;
        for (int dx = -r; dx <= r; ++dx) {                        // top before bottom, per column
            int i = ci + dx;
            if (i < 0 || i >= g.n_lon) continue;
            if (top_ok) consider(g, ci, cj, dx, -r, x, y, c);
            if (bot_ok) consider(g, ci, cj, dx, +r, x, y, c);
        }
        if (c.n >= 4) break;
        const bool lef_ok = ci - r >= 0, rig_ok = ci + r < g.n_lon;
        for (int dy = -r + 1; dy <= r - 1; ++dy) {                // left before right, per row
            int j = cj + dy;
            if (j < 0 || j >= g.n_lat) continue;
            if (lef_ok) consider(g, ci, cj, -r, dy, x, y, c);
            if (rig_ok) consider(g, ci, cj, +r, dy, x, y, c);
        }
        if (c.n >= 4) break;
    }
}

// Partial selection sort WITH SWAPS and strict '<' (GridH.cpp:123-140): the swap, not a stable
// shift, decides which of two exactly tied candidates is taken, so it is reproduced literally.
__device__ __forceinline__ void pick_four(Cands& c) {
    for (int m = 0; m < 4; ++m) {
        int best = m;
        for (int k = m + 1; k < c.n; ++k)
            if (c.d[k] < c.d[best]) best = k;
        double td = c.d[m]; c.d[m] = c.d[best]; c.d[best] = td;
        int16_t tc = c.code[m]; c.code[m] = c.code[best]; c.code[best] = tc;
    }
}

// Outcome of a search: up to four cells (global indices), their values and distances.
struct Picked {
    int i[4], j[4];
    double v[4], d[4];
    int found;                   // total candidates seen by the search (GridH's `found`)
};

template <typename T>
__device__ __forceinline__ void gather_picked(const GridView<T>& g, int ci, int cj, const Cands& c, Picked& p) {
    p.found = c.n;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < c.n) {
            p.i[k] = ci + off_i(c.code[k]); p.j[k] = cj + off_j(c.code[k]);
            p.v[k] = g.at(p.j[k], p.i[k]);  p.d[k] = c.d[k];
        } else {
            p.i[k] = -1; p.j[k] = -1; p.v[k] = qnan(); p.d[k] = qnan();
        }
    }
}

// found < 4: mean of what was found, in enumeration order (GridH.cpp:291-298, :350-356).
__device__ __forceinline__ double mean_found(const Picked& p) {
    double s = 0.0;
    for (int k = 0; k < p.found && k < 4; ++k) s = dadd(s, p.v[k]);
    return p.found > 0 ? div_count(s, p.found < 4 ? p.found : 4) : qnan();
}

template <typename T>
__device__ __forceinline__ void search_and_pick(const GridView<T>& g, double x, double y, int ci, int cj, Picked& p) {
    Cands c;
    ring_search(g, x, y, ci, cj, c);
    if (c.n >= 4) pick_four(c);
    gather_picked(g, ci, cj, c, p);
}

// ---- bicubic with 4-nearest-mean fallback, GridH.cpp:223-319 ---------------------------------------
template <typename T>
__device__ double exact_cubic(const GridView<T>& g, double x, double y, Picked* out_sel) {
    int xi = static_cast<int>(floor(x)), yi = static_cast<int>(floor(y));
    double tx = dsub(x, static_cast<double>(xi)), ty = dsub(y, static_cast<double>(yi));
    double col[4];
    bool dirty = false;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        int jj = clampi(yi - 1 + m, 0, g.n_lat - 1);              // clamp-to-edge, :240-247
        double p[4];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            p[n] = g.at(jj, clampi(xi - 1 + n, 0, g.n_lon - 1));
            dirty |= isnan(p[n]);
        }
        col[m] = catmull_rom_exact(p[0], p[1], p[2], p[3], tx);
    }
    if (!dirty) {
        if (out_sel) out_sel->found = -2;                         // clean stencil: no search ran
        return catmull_rom_exact(col[0], col[1], col[2], col[3], ty);
    }
    Picked p;
    search_and_pick(g, x, y, xi, yi, p);                          // floor centre, :281-289
    if (out_sel) *out_sel = p;
    if (p.found < 4) return mean_found(p);
    return mean_valid4(p.v[0], p.v[1], p.v[2], p.v[3]);
}

// Centre of the round-centred methods, GridH.cpp:333-336.
__device__ __forceinline__ int round_centre(double c, int n) {
    return clampi(static_cast<int>(round(c)), 0, n - 1);
}

// Variogram of the squared separation: nugget 1 + sill 100 * (1 - exp(-h / range 10)), GridH.cpp:371-376.
// The reference forms 1 - exp(-t) with t = h/10 ~ 1e-4 for bathymetry grids, i.e. by cancellation; here it is
// -expm1(-t): the Taylor polynomial (degree 7 below t = 2^-7, degree 11 below 2^-4: truncation < 1e-19 relative), else the
// library expm1.  Against the reference's own rounding of exp(-t) near 1 this moves gamma by ~1e-14 and the kriging
// prediction by < 1e-9 m (measured on every Grid-B fixture; tests hold kriging to 1e-6 m).
__device__ AUVI_VARIOGRAM_INLINE double variogram_sq(double h2, double nugget, double sill, double inv_range) {
    // h / range; h to ~1 ulp, far inside the tolerance.  (An FP32 rsqrt seed + one FP64 Newton step was measured instead of the
    // library's rsqrt: fewer instructions, but three of them on the quarter-rate XU pipe: kriging fill 3.18 -> 3.43 ms.)
    const double t = (h2 > 0.0 ? h2 * rsqrt(h2) : 0.0) * inv_range;
    double em1;                                                    // expm1(-t)
    if (t < 0.0078125) {                                           // bathymetry grids: t ~ 1e-4 .. 1e-3; degree 7: < 1e-19 relative
        double q = -1.0 / 5040.0;
        q = fma(q, t, 1.0 / 720.0);
        q = fma(q, t, -1.0 / 120.0);
        q = fma(q, t, 1.0 / 24.0);
        q = fma(q, t, -1.0 / 6.0);
        q = fma(q, t, 0.5);
        q = fma(q, t, -1.0);
        em1 = q * t;
    } else if (t < 0.0625) {
        double q = -1.0 / 39916800.0;
        q = fma(q, t, 1.0 / 3628800.0);
        q = fma(q, t, -1.0 / 362880.0);
        q = fma(q, t, 1.0 / 40320.0);
        q = fma(q, t, -1.0 / 5040.0);
        q = fma(q, t, 1.0 / 720.0);
        q = fma(q, t, -1.0 / 120.0);
        q = fma(q, t, 1.0 / 24.0);
        q = fma(q, t, -1.0 / 6.0);
        q = fma(q, t, 0.5);
        q = fma(q, t, -1.0);
        em1 = q * t;
    } else {
        em1 = expm1(-t);
    }
    return fma(-sill, em1, nugget);
}

// ---- ordinary kriging on the four picked cells, GridH.cpp:361-419 ----------------------------------
// Same system as the reference (5 x 5, Lagrange row), solved by Gaussian elimination without pivoting -- the
// pivots are the reference's Gauss-Jordan pivots, so its singularity test (|pivot| < 1e-12 -> mean of the four)
// fires on the same systems -- with fused multiply-adds and one reciprocal per pivot, then back substitution:
// ~65 FP64 operations instead of ~135 + 30 divisions.  Kriging is a tolerance method (the reference's own CPU
// and GPU paths differ through exp), so operation order is free here; the discrete part -- which four cells --
// is decided before this function and is bit-exact.
template <typename T>
__device__ double kriging_from_picked(const GridView<T>& g, const Picked& p, double lon, double lat) {
    double px[4], py[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {                                  // cell centres in degrees, :366-367
        px[k] = dadd(g.min_lon, dmul(dadd(static_cast<double>(p.i[k]), 0.5), g.lon_step));
        py[k] = dadd(g.min_lat, dmul(dadd(static_cast<double>(p.j[k]), 0.5), g.lat_step));
    }
    // symmetric variogram matrix with gamma(0) = 1 on the diagonal: six pair entries + four query entries
    double M[5][6];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        M[a][a] = g.vg_diag;                                       // gamma(0) = the nugget (1) in the reference
#pragma unroll
        for (int b = a + 1; b < 4; ++b) {
            const double dx = px[a] - px[b], dy = py[a] - py[b];
            M[a][b] = variogram_sq(fma(dx, dx, dy * dy), g.vg_nugget, g.vg_sill, g.vg_inv_range);
            M[b][a] = M[a][b];
        }
        M[a][4] = 1.0; M[4][a] = 1.0;
        const double dx = px[a] - lon, dy = py[a] - lat;           // raw query lon/lat, :380
        M[a][5] = variogram_sq(fma(dx, dx, dy * dy), g.vg_nugget, g.vg_sill, g.vg_inv_range);
    }
    M[4][4] = 0.0; M[4][5] = 1.0;
    double inv[5];
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const double piv = M[r][r];
        if (fabs(piv) < 1e-12) return mean_valid4(p.v[0], p.v[1], p.v[2], p.v[3]);   // :403-406
        inv[r] = 1.0 / piv;
#pragma unroll
        for (int k = r + 1; k < 5; ++k) {
            const double f = M[k][r] * inv[r];
#pragma unroll
            for (int q = r + 1; q < 6; ++q) M[k][q] = fma(-f, M[r][q], M[k][q]);
        }
    }
    double w[5];
#pragma unroll
    for (int r = 4; r >= 0; --r) {
        double acc = M[r][5];
#pragma unroll
        for (int q = r + 1; q < 5; ++q) acc = fma(-M[r][q], w[q], acc);
        w[r] = acc * inv[r];
    }
    double out = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) out = fma(w[k], p.v[k], out);      // :417-418
    return out;
}

// ---- EXTENSIONS: NN and IDW over the same selection ---------------
__device__ __forceinline__ void nearest_first(Picked& p) {          // found < 4: first strict minimum
    int m = p.found < 4 ? p.found : 4, best = 0;
    for (int k = 1; k < m; ++k) if (p.d[k] < p.d[best]) best = k;
    if (best) {
        double t = p.v[0]; p.v[0] = p.v[best]; p.v[best] = t;
        t = p.d[0]; p.d[0] = p.d[best]; p.d[best] = t;
        int u = p.i[0]; p.i[0] = p.i[best]; p.i[best] = u;
        u = p.j[0]; p.j[0] = p.j[best]; p.j[best] = u;
    }
}

__device__ __forceinline__ float rcp_sfu(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// IDW power 2 over min(found,4) picks.  Weights in FP32 through the SFU reciprocal (north_star:
// "FP32 FMA/SFU distance-weight math"); the SELECTION above stays FP64-exact.  d == 0 -> that value.
// The weight of a pick is rcp(float(d^2)) with d^2 the squared index-space distance formed exactly as the search
// forms it before its sqrt (GridH.cpp:42-44) -- the same expression, on the same bits, as the tiled kernel
// (fill.cu finish_four), so the point-list and the lattice entry points return identical IDW values.
__device__ __forceinline__ double idw_from_picked(const Picked& p, double x, double y) {
    int m = p.found < 4 ? p.found : 4;
    float num = 0.f, den = 0.f;
    // centre the values so the FP32 weighted mean keeps ~1e-7 relative accuracy on 10 km depths
    const double ref = p.v[0];
    for (int k = 0; k < m; ++k) {
        if (p.d[k] == 0.0) return p.v[k];
        const double ddi = dsub(dadd(static_cast<double>(p.i[k]), 0.5), x);
        const double ddj = dsub(dadd(static_cast<double>(p.j[k]), 0.5), y);
        const float w = rcp_sfu(static_cast<float>(dadd(dmul(ddi, ddi), dmul(ddj, ddj))));
        num = fmaf(w, static_cast<float>(p.v[k] - ref), num);
        den += w;
    }
    return ref + static_cast<double>(__fdividef(num, den));
}

// ---- OPT-IN: IDW over the TRUE four nearest valid cells ---------------------------------
// The reference's search stops at the first pass that brings its count to four (GridH.cpp:82,115), so its "four nearest"
// are nearest only among an order-dependent, early-terminated candidate list.  This method scans
// the whole radius-10 window in the same enumeration order WITHOUT the two early breaks and keeps the four smallest
// distances, ties going to the candidate enumerated first (a stable selection; the reference's swap-based selection has
// no meaning without its list).  Rings are left as soon as no farther cell can matter: a cell of ring r lies at
// distance >= r - 1 from a query whose centre is round(x), round(y).  Weights as IDW.  oracle: orc_idw_knn.
template <typename T>
__device__ double exact_idw_knn(const GridView<T>& g, double x, double y, Picked* out_sel) {
    const int ci = round_centre(x, g.n_lon), cj = round_centre(y, g.n_lat);
    Picked p;
    p.found = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { p.i[k] = -1; p.j[k] = -1; p.v[k] = qnan(); p.d[k] = __longlong_as_double(0x7ff0000000000000LL); }
    auto visit = [&](int i, int j) {
        const double v = g.at(j, i);
        if (isnan(v)) return;
        const double ddi = dsub(dadd(static_cast<double>(i), 0.5), x), ddj = dsub(dadd(static_cast<double>(j), 0.5), y);
        const double d = dsqrt(dadd(dmul(ddi, ddi), dmul(ddj, ddj)));
        ++p.found;
        if (!(d < p.d[3])) return;                                 // not among the four; an equal distance came later: loses
        int at = 3;
        while (at > 0 && d < p.d[at - 1]) { p.d[at] = p.d[at - 1]; p.v[at] = p.v[at - 1]; p.i[at] = p.i[at - 1]; p.j[at] = p.j[at - 1]; --at; }
        p.d[at] = d; p.v[at] = v; p.i[at] = i; p.j[at] = j;
    };
    visit(ci, cj);
    for (int r = 1; r <= kMaxRadius; ++r) {
        if (p.found >= 4 && p.d[3] <= static_cast<double>(r - 1)) break;   // every cell of rings >= r is at least r - 1 away
        const bool top_ok = cj - r >= 0, bot_ok = cj + r < g.n_lat;
        for (int dx = -r; dx <= r; ++dx) {
            const int i = ci + dx;
            if (i < 0 || i >= g.n_lon) continue;
            if (top_ok) visit(i, cj - r);
            if (bot_ok) visit(i, cj + r);
        }
        const bool lef_ok = ci - r >= 0, rig_ok = ci + r < g.n_lon;
        for (int dy = -r + 1; dy <= r - 1; ++dy) {
            const int j = cj + dy;
            if (j < 0 || j >= g.n_lat) continue;
            if (lef_ok) visit(ci - r, j);
            if (rig_ok) visit(ci + r, j);
        }
    }
    if (out_sel) *out_sel = p;
    return p.found > 0 ? idw_from_picked(p, x, y) : qnan();
}

// ---- one query, any method --------------------------------------------------------------------------
// lon/lat are the raw query coordinates; x/y their index-space images (precomputed by the caller so
// that lattice kernels can take them from per-axis tables).  NaN x or y means out of bounds.
template <typename T>
__device__ double interp_exact(const GridView<T>& g, int method, double lon, double lat,
                               double x, double y, Picked* sel) {
    if (sel) { sel->found = -1; for (int k = 0; k < 4; ++k) { sel->i[k] = -1; sel->j[k] = -1; } }
    if (isnan(x) || isnan(y)) return qnan();
    if (method == BILINEAR) { if (sel) sel->found = -2; return exact_bilinear(g, x, y); }
    if (method == CUBIC) return exact_cubic(g, x, y, sel);
    if (method == IDW_KNN) return exact_idw_knn(g, x, y, sel);
    if (method == BILINEAR_SEARCH) {
        // OPT-IN (not a reference method): bilinear wherever the reference's bilinear returns a
        // number; where it returns NaN (all four corners missing, GridH.cpp:186-198) the 4-nearest mean the bicubic
        // method falls back to (floor-centred ring search, GridH.cpp:272-318).
        const double b = exact_bilinear(g, x, y);
        if (!isnan(b)) { if (sel) sel->found = -2; return b; }
        Picked p;
        search_and_pick(g, x, y, static_cast<int>(floor(x)), static_cast<int>(floor(y)), p);
        if (sel) *sel = p;
        return p.found < 4 ? mean_found(p) : mean_valid4(p.v[0], p.v[1], p.v[2], p.v[3]);
    }
    Picked p;
    int ci = round_centre(x, g.n_lon), cj = round_centre(y, g.n_lat);
    search_and_pick(g, x, y, ci, cj, p);
    double out;
    if (method == KRIGING) {
        out = p.found < 4 ? mean_found(p) : kriging_from_picked(g, p, lon, lat);
    } else if (method == NN) {
        if (p.found < 4) nearest_first(p);
        out = p.found > 0 ? p.v[0] : qnan();
    } else {
        out = p.found > 0 ? idw_from_picked(p, x, y) : qnan();
    }
    if (sel) *sel = p;
    return out;
}

}  // namespace auvi
