/* oracle/interp_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded, FP64 restatement of the reference's CPU interpolation path
 * (code/src/GridH.cpp of devsaxena974/AUV-Real-Time-Interpolation) on a flat row-major grid.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product (libauvi.so) never does and has no CPU fallback.
 *
 * PARITY PIN: oracle/binding.py + tests/test_oracle_cpu.py check this file point-by-point
 * (bit-for-bit for bilinear, cubic and all neighbour selections; <=1e-9 m for kriging) against
 * oracle/_ref/libgridh_ref.so -- the unmodified reference compiled in place -- and against the
 * golden MAE/RMSE/Max rows of results/TestingResults1.csv (tests/golden/golden_metrics.json).
 * The NN and IDW methods are EXTENSIONS with no reference code:
 * their neighbour selection is pinned by the reference's own search, their value is "parity
 * unpinned" by construction.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fPIC -shared (oracle/Makefile).  -ffp-contract=off
 * matters: a fused multiply-add in `min + idx*step` flips floor()/round() decisions
 * (SURVEY.md section 0, fact 4).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_RADIUS 10          /* GridH.cpp:275, :339 */
#define ORC_MAX_CAND   48          /* reachable maximum is 45: 3 + 2*(2*10+1) */

enum { ORC_BILINEAR = 0, ORC_CUBIC = 1, ORC_KRIGING = 2, ORC_NN = 3, ORC_IDW = 4 };

typedef struct {
    const double* z;               /* [n_lat][n_lon], row 0 = min_lat */
    int n_lat, n_lon;
    double min_lon, max_lon, min_lat, max_lat;
    double lon_step, lat_step;     /* GridH.cpp:156-157 */
} orc_grid;

typedef struct {
    int i[ORC_MAX_CAND], j[ORC_MAX_CAND];
    double v[ORC_MAX_CAND], d[ORC_MAX_CAND];
    int n;
} orc_cands;

void orc_grid_init(orc_grid* g, const double* z, int n_lat, int n_lon,
                   double min_lon, double max_lon, double min_lat, double max_lat) {
    g->z = z; g->n_lat = n_lat; g->n_lon = n_lon;
    g->min_lon = min_lon; g->max_lon = max_lon; g->min_lat = min_lat; g->max_lat = max_lat;
    g->lon_step = (max_lon - min_lon) / (n_lon - 1);
    g->lat_step = (max_lat - min_lat) / (n_lat - 1);
}

static inline double cell(const orc_grid* g, int j, int i) { return g->z[(size_t)j * g->n_lon + i]; }
static inline int outside(const orc_grid* g, double lon, double lat) {
    return lon < g->min_lon || lon > g->max_lon || lat < g->min_lat || lat > g->max_lat;
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Mean of the non-NaN members of four values (GridH.cpp:10-18). */
static double mean_valid4(double a, double b, double c, double d) {
    double s = 0.0; int n = 0;
    if (!isnan(a)) { s += a; ++n; }
    if (!isnan(b)) { s += b; ++n; }
    if (!isnan(c)) { s += c; ++n; }
    if (!isnan(d)) { s += d; ++n; }
    return n ? s / n : NAN;
}

/* ---- ring search (GridH.cpp:24-118) ------------------------------------------------------ */
static void consider(const orc_grid* g, int i, int j, double x, double y, orc_cands* c) {
    double v = cell(g, j, i);
    if (isnan(v)) return;
    double di = (i + 0.5) - x, dj = (j + 0.5) - y;       /* cell-centre offset, index space */
    c->i[c->n] = i; c->j[c->n] = j; c->v[c->n] = v;
    c->d[c->n] = sqrt(di * di + dj * dj);
    ++c->n;
}

static void ring_search(const orc_grid* g, double x, double y, int ci, int cj, orc_cands* c) {
    c->n = 0;
    consider(g, ci, cj, x, y, c);                         /* centre first, :36-46 */
    for (int r = 1; r <= ORC_MAX_RADIUS; ++r) {
        int top = cj - r, bot = cj + r;
        for (int dx = -r; dx <= r; ++dx) {                /* top before bottom per column, :53-81 */
            int i = ci + dx;
            if (i < 0 || i >= g->n_lon) continue;
            if (top >= 0)        consider(g, i, top, x, y, c);
            if (bot < g->n_lat)  consider(g, i, bot, x, y, c);
        }
        if (c->n >= 4) break;                             /* :82 */
        int lef = ci - r, rig = ci + r;
        for (int dy = -r + 1; dy <= r - 1; ++dy) {        /* left before right per row, :86-114 */
            int j = cj + dy;
            if (j < 0 || j >= g->n_lat) continue;
            if (lef >= 0)        consider(g, lef, j, x, y, c);
            if (rig < g->n_lon)  consider(g, rig, j, x, y, c);
        }
        if (c->n >= 4) break;                             /* :115 */
    }
}

/* Partial selection sort with swaps, strict '<' (GridH.cpp:123-140).  The swap (not a stable
 * shift) is part of the observable behaviour when distances tie exactly. */
static void pick_four(orc_cands* c) {
    for (int m = 0; m < 4; ++m) {
        int best = m;
        for (int k = m + 1; k < c->n; ++k)
            if (c->d[k] < c->d[best]) best = k;
        double td = c->d[m]; c->d[m] = c->d[best]; c->d[best] = td;
        double tv = c->v[m]; c->v[m] = c->v[best]; c->v[best] = tv;
        int ti = c->i[m]; c->i[m] = c->i[best]; c->i[best] = ti;
        int tj = c->j[m]; c->j[m] = c->j[best]; c->j[best] = tj;
    }
}

static void export_sel(const orc_cands* c, int32_t* sel) {
    if (!sel) return;
    int m = c->n < 4 ? c->n : 4;
    for (int k = 0; k < 4; ++k) {
        sel[2 * k]     = k < m ? c->i[k] : -1;
        sel[2 * k + 1] = k < m ? c->j[k] : -1;
    }
}

static double mean_first(const orc_cands* c) {            /* found < 4 branch, :291-298, :350-356 */
    double s = 0.0;
    for (int k = 0; k < c->n; ++k) s += c->v[k];
    return c->n ? s / c->n : NAN;
}

/* ---- bilinear (GridH.cpp:160-210) --------------------------------------------------------- */
double orc_bilinear(const orc_grid* g, double lon, double lat) {
    if (outside(g, lon, lat)) return NAN;
    double x = (lon - g->min_lon) / g->lon_step;
    double y = (lat - g->min_lat) / g->lat_step;
    int x0 = (int)floor(x), y0 = (int)floor(y);
    int x1 = x0 + 1 < g->n_lon - 1 ? x0 + 1 : g->n_lon - 1;
    int y1 = y0 + 1 < g->n_lat - 1 ? y0 + 1 : g->n_lat - 1;
    double wx = x - x0, wy = y - y0;
    double a = cell(g, y0, x0), b = cell(g, y0, x1), c = cell(g, y1, x0), d = cell(g, y1, x1);
    if (isnan(a) || isnan(b) || isnan(c) || isnan(d)) return mean_valid4(a, b, c, d);
    double lo = (1 - wx) * a + wx * b;
    double hi = (1 - wx) * c + wx * d;
    return (1 - wy) * lo + wy * hi;
}

/* ---- bicubic Catmull-Rom with 4-nearest-mean fallback (GridH.cpp:215-319) ----------------- */
static double catmull_rom(double p0, double p1, double p2, double p3, double t) {
    return 0.5 * (2 * p1 + (-p0 + p2) * t + (2 * p0 - 5 * p1 + 4 * p2 - p3) * t * t
                  + (-p0 + 3 * p1 - 3 * p2 + p3) * t * t * t);
}

double orc_cubic(const orc_grid* g, double lon, double lat, int32_t* sel, int32_t* found) {
    if (found) *found = -1;
    if (sel) for (int k = 0; k < 8; ++k) sel[k] = -1;
    if (outside(g, lon, lat)) return NAN;
    double xf = (lon - g->min_lon) / g->lon_step;
    double yf = (lat - g->min_lat) / g->lat_step;
    int xi = (int)floor(xf), yi = (int)floor(yf);
    double tx = xf - xi, ty = yf - yi;

    double p[4][4];
    int dirty = 0;
    for (int m = 0; m < 4; ++m) {                         /* clamp-to-edge 4x4, :239-253 */
        int jj = clampi(yi - 1 + m, 0, g->n_lat - 1);
        for (int n = 0; n < 4; ++n) {
            int ii = clampi(xi - 1 + n, 0, g->n_lon - 1);
            p[m][n] = cell(g, jj, ii);
            dirty |= isnan(p[m][n]);
        }
    }
    if (!dirty) {
        double col[4];
        for (int m = 0; m < 4; ++m) col[m] = catmull_rom(p[m][0], p[m][1], p[m][2], p[m][3], tx);
        if (found) *found = -2;                           /* -2: clean stencil, no search */
        return catmull_rom(col[0], col[1], col[2], col[3], ty);
    }
    orc_cands c;
    ring_search(g, xf, yf, xi, yi, &c);                   /* floor centre, :281-289 */
    if (found) *found = c.n;
    if (c.n < 4) { export_sel(&c, sel); return mean_first(&c); }
    pick_four(&c);
    export_sel(&c, sel);
    return mean_valid4(c.v[0], c.v[1], c.v[2], c.v[3]);
}

/* ---- shared prologue of the round-centred methods (GridH.cpp:331-359) --------------------- */
static int round_centre_search(const orc_grid* g, double lon, double lat, orc_cands* c,
                               double* px, double* py) {
    double x = (lon - g->min_lon) / g->lon_step;
    double y = (lat - g->min_lat) / g->lat_step;
    int ci = clampi((int)round(x), 0, g->n_lon - 1);
    int cj = clampi((int)round(y), 0, g->n_lat - 1);
    ring_search(g, x, y, ci, cj, c);
    if (px) *px = x;
    if (py) *py = y;
    return c->n;
}

static double variogram(double h) {                       /* GridH.cpp:371-376 */
    return 1.0 + 100.0 * (1.0 - exp(-h / 10.0));
}

/* ---- ordinary kriging on the 4 selected cells (GridH.cpp:326-420) ------------------------- */
double orc_kriging(const orc_grid* g, double lon, double lat, int32_t* sel, int32_t* found) {
    if (found) *found = -1;
    if (sel) for (int k = 0; k < 8; ++k) sel[k] = -1;
    if (outside(g, lon, lat)) return NAN;
    orc_cands c;
    int n = round_centre_search(g, lon, lat, &c, NULL, NULL);
    if (found) *found = n;
    if (n < 4) { export_sel(&c, sel); return mean_first(&c); }
    pick_four(&c);
    export_sel(&c, sel);

    double px[4], py[4];
    for (int k = 0; k < 4; ++k) {                          /* cell centres in degrees, :366-367 */
        px[k] = g->min_lon + (c.i[k] + 0.5) * g->lon_step;
        py[k] = g->min_lat + (c.j[k] + 0.5) * g->lat_step;
    }
    double M[5][6];
    memset(M, 0, sizeof M);
    for (int a = 0; a < 4; ++a) {
        for (int b = 0; b < 4; ++b) {
            double dx = px[a] - px[b], dy = py[a] - py[b];
            M[a][b] = variogram(sqrt(dx * dx + dy * dy)); /* gamma(0) = 1 on the diagonal */
        }
        M[a][4] = 1.0;
        M[4][a] = 1.0;
        double dx = px[a] - lon, dy = py[a] - lat;         /* raw query lon/lat, :380, :393-397 */
        M[a][5] = variogram(sqrt(dx * dx + dy * dy));
    }
    M[4][5] = 1.0;
    for (int r = 0; r < 5; ++r) {                          /* Gauss-Jordan, no pivoting, :401-414 */
        double piv = M[r][r];
        if (fabs(piv) < 1e-12) return mean_valid4(c.v[0], c.v[1], c.v[2], c.v[3]);
        for (int q = r; q < 6; ++q) M[r][q] /= piv;
        for (int k = 0; k < 5; ++k) {
            if (k == r) continue;
            double f = M[k][r];
            for (int q = r; q < 6; ++q) M[k][q] -= f * M[r][q];
        }
    }
    double out = 0.0;
    for (int k = 0; k < 4; ++k) out += M[k][5] * c.v[k];
    return out;
}

/* ---- EXTENSION A7: nearest neighbour = first selected candidate --------------------------- */
double orc_nn(const orc_grid* g, double lon, double lat, int32_t* sel, int32_t* found) {
    if (found) *found = -1;
    if (sel) for (int k = 0; k < 8; ++k) sel[k] = -1;
    if (outside(g, lon, lat)) return NAN;
    orc_cands c;
    int n = round_centre_search(g, lon, lat, &c, NULL, NULL);
    if (found) *found = n;
    if (n == 0) return NAN;
    if (n >= 4) pick_four(&c);
    else {                                                  /* fewer than 4: first strict minimum */
        int best = 0;
        for (int k = 1; k < n; ++k) if (c.d[k] < c.d[best]) best = k;
        double tv = c.v[0]; c.v[0] = c.v[best]; c.v[best] = tv;
        double td = c.d[0]; c.d[0] = c.d[best]; c.d[best] = td;
        int ti = c.i[0]; c.i[0] = c.i[best]; c.i[best] = ti;
        int tj = c.j[0]; c.j[0] = c.j[best]; c.j[best] = tj;
    }
    export_sel(&c, sel);
    return c.v[0];
}

/* ---- EXTENSION A8: inverse-distance weighting, power 2, over the selected <=4 ------------- */
double orc_idw(const orc_grid* g, double lon, double lat, int32_t* sel, int32_t* found) {
    if (found) *found = -1;
    if (sel) for (int k = 0; k < 8; ++k) sel[k] = -1;
    if (outside(g, lon, lat)) return NAN;
    orc_cands c;
    int n = round_centre_search(g, lon, lat, &c, NULL, NULL);
    if (found) *found = n;
    if (n == 0) return NAN;
    if (n >= 4) pick_four(&c);
    export_sel(&c, sel);
    int m = n < 4 ? n : 4;
    double num = 0.0, den = 0.0;
    for (int k = 0; k < m; ++k) {
        if (c.d[k] == 0.0) return c.v[k];
        double w = 1.0 / (c.d[k] * c.d[k]);
        num += w * c.v[k];
        den += w;
    }
    return num / den;
}

/* ---- batch entry (GridH.cpp:422-448 loop) -------------------------------------------------- */
/* pts: n x {lon,lat,elev} (Point.h:9-13).  sel (optional): n x 8 int32, found (optional): n. */
int orc_batch(const orc_grid* g, int method, const double* pts, int64_t n, double* out,
              int32_t* sel, int32_t* found) {
    for (int64_t k = 0; k < n; ++k) {
        double lon = pts[3 * k], lat = pts[3 * k + 1];
        int32_t* s = sel ? sel + 8 * k : NULL;
        int32_t* f = found ? found + k : NULL;
        switch (method) {
            case ORC_BILINEAR: out[k] = orc_bilinear(g, lon, lat); if (f) *f = -2; break;
            case ORC_CUBIC:    out[k] = orc_cubic(g, lon, lat, s, f); break;
            case ORC_KRIGING:  out[k] = orc_kriging(g, lon, lat, s, f); break;
            case ORC_NN:       out[k] = orc_nn(g, lon, lat, s, f); break;
            case ORC_IDW:      out[k] = orc_idw(g, lon, lat, s, f); break;
            default: return 1;
        }
    }
    return 0;
}

/* ---- query generators used by the reference drivers --------------------------------------- */
/* Grid-B node queries: coord = min + idx*step, step = (max-min)/(n-1) (test_gebco.cpp:72-81). */
void orc_node_axis(double lo, double hi, int n, double* out) {
    double step = (hi - lo) / (n - 1);
    for (int k = 0; k < n; ++k) out[k] = lo + k * step;
}
/* Grid-A expanded lattice: coord = min + k*(max-min)/(new_n-1) (test_interpolation.cpp:99-106). */
void orc_lattice_axis(double lo, double hi, int new_n, double* out) {
    for (int k = 0; k < new_n; ++k) out[k] = lo + k * (hi - lo) / (new_n - 1);
}

/* ---- error metrics (error_calculator.cpp:5-45) --------------------------------------------- */
/* NaN interpolants add nothing to the numerator but still count in the denominator. */
double orc_mae(const double* truth, const double* est, int64_t n) {
    double s = 0.0;
    for (int64_t k = 0; k < n; ++k) if (!isnan(est[k])) s += fabs(truth[k] - est[k]);
    return s / (double)n;
}
double orc_rmse(const double* truth, const double* est, int64_t n) {
    double s = 0.0;
    for (int64_t k = 0; k < n; ++k) if (!isnan(est[k])) { double d = truth[k] - est[k]; s += d * d; }
    return sqrt(s / (double)n);
}
double orc_maxerr(const double* truth, const double* est, int64_t n) {
    double m = 0.0;
    for (int64_t k = 0; k < n; ++k) { double d = fabs(truth[k] - est[k]); if (d > m) m = d; }
    return m;
}

/* ---- synthetic Grid-A field (generate_csv_grids.cpp:32-70) -------------------------------- */
/* z = -(10 + 2x) + 100*exp(-((x-75)^2/450 + (y-50)^2/450)), x = 100*i/(n_lon-1), y likewise. */
void orc_synth_grid(int n_lat, int n_lon, double* z) {
    for (int j = 0; j < n_lat; ++j) {
        double y = 100.0 * j / (n_lat - 1);
        for (int i = 0; i < n_lon; ++i) {
            double x = 100.0 * i / (n_lon - 1);
            double base = -(10.0 + 2.0 * x);
            double bump = 100.0 * exp(-((x - 75.0) * (x - 75.0) / (2 * 15.0 * 15.0)
                                        + (y - 50.0) * (y - 50.0) / (2 * 15.0 * 15.0)));
            z[(size_t)j * n_lon + i] = base + bump;
        }
    }
}

/* ==== OPT-IN METHODS -- no reference code; these are the checkers of our own definitions ==== */

/* ---- IDW over the TRUE four nearest valid cells: the reference's enumeration (GridH.cpp:24-118) WITHOUT its two early
 * `count >= 4` breaks (:82, :115), i.e. the whole radius-10 window, then the four smallest distances with ties going to
 * the candidate enumerated first (a stable selection), then power-2 weights as orc_idw. ------------------------------ */
#define ORC_FULL_CAND 441
typedef struct { int i[ORC_FULL_CAND], j[ORC_FULL_CAND]; double v[ORC_FULL_CAND], d[ORC_FULL_CAND]; int n; } orc_full;

static void consider_full(const orc_grid* g, int i, int j, double x, double y, orc_full* c) {
    double v = cell(g, j, i);
    if (isnan(v)) return;
    double di = (i + 0.5) - x, dj = (j + 0.5) - y;
    c->i[c->n] = i; c->j[c->n] = j; c->v[c->n] = v;
    c->d[c->n] = sqrt(di * di + dj * dj);
    ++c->n;
}

double orc_idw_knn(const orc_grid* g, double lon, double lat, int32_t* sel, int32_t* found) {
    if (found) *found = -1;
    if (sel) for (int k = 0; k < 8; ++k) sel[k] = -1;
    if (outside(g, lon, lat)) return NAN;
    double x = (lon - g->min_lon) / g->lon_step;
    double y = (lat - g->min_lat) / g->lat_step;
    int ci = clampi((int)round(x), 0, g->n_lon - 1);
    int cj = clampi((int)round(y), 0, g->n_lat - 1);
    orc_full c;
    c.n = 0;
    consider_full(g, ci, cj, x, y, &c);
    for (int r = 1; r <= ORC_MAX_RADIUS; ++r) {           /* every ring, both passes, no early exit */
        int top = cj - r, bot = cj + r;
        for (int dx = -r; dx <= r; ++dx) {
            int i = ci + dx;
            if (i < 0 || i >= g->n_lon) continue;
            if (top >= 0)        consider_full(g, i, top, x, y, &c);
            if (bot < g->n_lat)  consider_full(g, i, bot, x, y, &c);
        }
        int lef = ci - r, rig = ci + r;
        for (int dy = -r + 1; dy <= r - 1; ++dy) {
            int j = cj + dy;
            if (j < 0 || j >= g->n_lat) continue;
            if (lef >= 0)        consider_full(g, lef, j, x, y, &c);
            if (rig < g->n_lon)  consider_full(g, rig, j, x, y, &c);
        }
    }
    /* NOTE: `found` of this method counts the candidates of the rings the device scans before it can stop; the oracle
     * reports the picks only (found = min(n, 4) is what the tests compare). */
    int m = c.n < 4 ? c.n : 4;
    if (found) *found = m;
    if (c.n == 0) return NAN;
    int taken[4] = {-1, -1, -1, -1};
    for (int p = 0; p < m; ++p) {                          /* stable selection: first strict minimum among the rest */
        int best = -1;
        for (int k = 0; k < c.n; ++k) {
            if (k == taken[0] || k == taken[1] || k == taken[2] || k == taken[3]) continue;
            if (best < 0 || c.d[k] < c.d[best]) best = k;
        }
        taken[p] = best;
        if (sel) { sel[2 * p] = c.i[best]; sel[2 * p + 1] = c.j[best]; }
    }
    double num = 0.0, den = 0.0;
    for (int p = 0; p < m; ++p) {
        int k = taken[p];
        if (c.d[k] == 0.0) return c.v[k];
        double w = 1.0 / (c.d[k] * c.d[k]);
        num += w * c.v[k];
        den += w;
    }
    return num / den;
}

/* ---- fitted variogram: exponential model c0 + c1 * (1 - exp(-h / a)) from the grid's own empirical semivariances at lags
 * 1, 2, 4, 8 cells along both axes.  sums16 layout as the device reduction: [axis 0 lon, 1 lat][lag][sum of squares, pairs].
 * Fit = the library's auvi_variogram_fit_from_sums restated (weighted least squares over a ladder of candidate ranges). ---- */
void orc_variogram_sums(const orc_grid* g, double* sums16) {
    for (int k = 0; k < 16; ++k) sums16[k] = 0.0;
    for (int r = 0; r < g->n_lat; ++r)
        for (int i = 0; i < g->n_lon; ++i) {
            double v = cell(g, r, i);
            if (isnan(v)) continue;
            for (int l = 0; l < 4; ++l) {
                int k = 1 << l;
                if (i + k < g->n_lon) { double w = cell(g, r, i + k); if (!isnan(w)) { sums16[2 * l] += (v - w) * (v - w); sums16[2 * l + 1] += 1.0; } }
                if (r + k < g->n_lat) { double w = cell(g, r + k, i); if (!isnan(w)) { sums16[8 + 2 * l] += (v - w) * (v - w); sums16[8 + 2 * l + 1] += 1.0; } }
            }
        }
}

int orc_variogram_fit(const double* sums16, double lon_step, double lat_step, double* out3) {
    double h[8], gam[8], w[8], h_max = 0.0;
    int n = 0;
    for (int a = 0; a < 2; ++a)
        for (int l = 0; l < 4; ++l) {
            double ss = sums16[a * 8 + 2 * l], cnt = sums16[a * 8 + 2 * l + 1];
            if (!(cnt > 0.0)) continue;
            h[n] = (double)(1 << l) * fabs(a == 0 ? lon_step : lat_step);
            gam[n] = ss / (2.0 * cnt);
            w[n] = cnt;
            if (h[n] > h_max) h_max = h[n];
            ++n;
        }
    if (n < 2 || !(h_max > 0.0)) return 1;
    double best_r = 0.0;
    int have = 0;
    for (int m = 0; m < 10; ++m) {
        double a = h_max * ldexp(1.0, m - 2), f[8];
        double sw = 0, sf = 0, sff = 0, sg = 0, sfg = 0;
        for (int k = 0; k < n; ++k) {
            f[k] = 1.0 - exp(-h[k] / a);
            sw += w[k]; sf += w[k] * f[k]; sff += w[k] * f[k] * f[k]; sg += w[k] * gam[k]; sfg += w[k] * f[k] * gam[k];
        }
        double det = sw * sff - sf * sf;
        double c1 = det != 0.0 ? (sw * sfg - sf * sg) / det : 0.0;
        double c0 = (sg - c1 * sf) / sw;
        if (!(c0 >= 0.0) || det == 0.0) { c0 = 0.0; c1 = sff > 0.0 ? sfg / sff : 0.0; }
        if (!(c1 > 0.0) || !isfinite(c1) || !isfinite(c0)) continue;
        double r = 0.0;
        for (int k = 0; k < n; ++k) { double e = gam[k] - c0 - c1 * f[k]; r += w[k] * e * e; }
        if (!have || r < best_r) { have = 1; best_r = r; out3[0] = c0; out3[1] = c1; out3[2] = a; }
    }
    return have ? 0 : 1;
}

/* Ordinary kriging on the reference's four picks (same search, same selection as orc_kriging) with the fitted model in
 * covariance form: C(h) = c1 * exp(-h / a) between distinct points and towards the query, c0 + c1 on the diagonal; the
 * 5 x 5 system [C 1; 1' 0][w; mu] = [c; 1] solved by Gaussian elimination with partial pivoting (any exact solver will do). */
double orc_kriging_fitted(const orc_grid* g, double lon, double lat, double c0, double c1, double a, int32_t* sel, int32_t* found) {
    if (found) *found = -1;
    if (sel) for (int k = 0; k < 8; ++k) sel[k] = -1;
    if (outside(g, lon, lat)) return NAN;
    orc_cands c;
    int n = round_centre_search(g, lon, lat, &c, NULL, NULL);
    if (found) *found = n;
    if (n < 4) { export_sel(&c, sel); return mean_first(&c); }
    pick_four(&c);
    export_sel(&c, sel);
    double px[4], py[4], M[5][6];
    for (int k = 0; k < 4; ++k) {
        px[k] = g->min_lon + (c.i[k] + 0.5) * g->lon_step;
        py[k] = g->min_lat + (c.j[k] + 0.5) * g->lat_step;
    }
    memset(M, 0, sizeof M);
    for (int p = 0; p < 4; ++p) {
        for (int q = 0; q < 4; ++q) {
            double dx = px[p] - px[q], dy = py[p] - py[q];
            M[p][q] = p == q ? c0 + c1 : c1 * exp(-sqrt(dx * dx + dy * dy) / a);
        }
        M[p][4] = 1.0; M[4][p] = 1.0;
        double dx = px[p] - lon, dy = py[p] - lat;
        M[p][5] = c1 * exp(-sqrt(dx * dx + dy * dy) / a);
    }
    M[4][5] = 1.0;
    for (int r = 0; r < 5; ++r) {
        int piv = r;
        for (int k = r + 1; k < 5; ++k) if (fabs(M[k][r]) > fabs(M[piv][r])) piv = k;
        if (fabs(M[piv][r]) < 1e-300) return mean_valid4(c.v[0], c.v[1], c.v[2], c.v[3]);
        if (piv != r) for (int q = 0; q < 6; ++q) { double t = M[r][q]; M[r][q] = M[piv][q]; M[piv][q] = t; }
        for (int k = 0; k < 5; ++k) {
            if (k == r) continue;
            double f = M[k][r] / M[r][r];
            for (int q = r; q < 6; ++q) M[k][q] -= f * M[r][q];
        }
    }
    double out = 0.0;
    for (int k = 0; k < 4; ++k) out += (M[k][5] / M[k][k]) * c.v[k];
    return out;
}

/* batch forms of the opt-in methods: method 6 = IDW_KNN, 7 = KRIGING_FITTED (params = {c0, c1, range}) */
int orc_batch_optin(const orc_grid* g, int method, const double* params, const double* pts, int64_t n, double* out,
                    int32_t* sel, int32_t* found) {
    for (int64_t k = 0; k < n; ++k) {
        double lon = pts[3 * k], lat = pts[3 * k + 1];
        int32_t* s = sel ? sel + 8 * k : NULL;
        int32_t* f = found ? found + k : NULL;
        if (method == 6) out[k] = orc_idw_knn(g, lon, lat, s, f);
        else if (method == 7) out[k] = orc_kriging_fitted(g, lon, lat, params[0], params[1], params[2], s, f);
        else return 1;
    }
    return 0;
}
