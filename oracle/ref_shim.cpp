// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A C-ABI wrapper around the *unmodified* reference CPU class GridH, compiled in place from
// /root/reference/code/src/GridH.cpp + error_calculator.cpp (see oracle/Makefile).  The result,
// oracle/_ref/libgridh_ref.so, is the strongest checker we have: it IS the reference.  It is used
//   (1) to pin oracle/interp_oracle.c (our plain-C restatement) point by point,
//   (2) to generate tests/golden/*.json with oracle/make_fixtures.py,
//   (3) as bench.py's `--impl reference` arm and `cpu_baseline` (kind = "reference").
// No reference source is copied into this repository: REF_GRIDH_CPP / REF_ERRCALC_CPP are
// include paths into the read-only checkout, supplied by the Makefile.
//
// The shim reaches the reference's file-static helpers (findCandidateNeighbors
// GridH.cpp:24-118, selectFourNearest GridH.cpp:123-140) by textual inclusion, so that the
// four selected neighbour indices can be dumped without touching the reference.

#include REF_GRIDH_CPP
#include REF_ERRCALC_CPP

#include <cstdint>
#include <cstring>
#include <thread>

namespace {

struct RefGrid {
    std::vector<std::vector<double>> rows;   // [n_lat][n_lon], row 0 = min_lat (GridH.cpp:153)
    GridH* h;
    int n_lat, n_lon;
    double min_lon, max_lon, min_lat, max_lat;
    double lon_step, lat_step;               // same expression as GridH.cpp:156-157
};

typedef std::vector<Point> (GridH::*BatchFn)(const std::vector<Point>&) const;

BatchFn pick(int method) {
    switch (method) {
        case 0: return &GridH::batchBilinearInterpolate;
        case 1: return &GridH::batchCubicInterpolate;
        case 2: return &GridH::batchOrdinaryKrigingInterpolate;
    }
    return nullptr;
}

void run_slice(const RefGrid* g, int method, const double* pts, int64_t lo, int64_t hi, double* out) {
    std::vector<Point> q(static_cast<size_t>(hi - lo));
    for (int64_t k = lo; k < hi; ++k) q[k - lo] = Point{pts[3 * k], pts[3 * k + 1], pts[3 * k + 2]};
    std::vector<Point> r = (g->h->*pick(method))(q);
    for (int64_t k = lo; k < hi; ++k) out[k] = r[k - lo].elev;
}

}  // namespace

extern "C" {

// Build a GridH from a dense row-major array (row 0 = min_lat), argument order of GridH.h:20-27.
void* refh_create(const double* rowmajor, int n_lat, int n_lon,
                  double min_lon, double max_lon, double min_lat, double max_lat) {
    RefGrid* g = new RefGrid;
    g->rows.assign(n_lat, std::vector<double>(n_lon));
    for (int j = 0; j < n_lat; ++j)
        std::memcpy(g->rows[j].data(), rowmajor + static_cast<size_t>(j) * n_lon, sizeof(double) * n_lon);
    g->n_lat = n_lat; g->n_lon = n_lon;
    g->min_lon = min_lon; g->max_lon = max_lon; g->min_lat = min_lat; g->max_lat = max_lat;
    g->lon_step = (max_lon - min_lon) / (n_lon - 1);
    g->lat_step = (max_lat - min_lat) / (n_lat - 1);
    g->h = new GridH(max_lat, min_lat, n_lat, max_lon, min_lon, n_lon, g->rows);
    return g;
}

void refh_destroy(void* p) {
    RefGrid* g = static_cast<RefGrid*>(p);
    if (!g) return;
    delete g->h;
    delete g;
}

// method: 0 bilinear, 1 cubic, 2 kriging.  pts = n x {lon,lat,elev} doubles (Point.h:9-13).
// Calls GridH::batch* (GridH.cpp:422-448) on `threads` contiguous slices (1 = as shipped).
int refh_batch(void* p, int method, const double* pts, int64_t n, double* out_elev, int threads) {
    RefGrid* g = static_cast<RefGrid*>(p);
    if (!g || !pick(method)) return 1;
    if (n <= 0) return 0;
    if (threads <= 1) { run_slice(g, method, pts, 0, n, out_elev); return 0; }
    std::vector<std::thread> pool;
    int64_t per = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        int64_t lo = t * per, hi = std::min<int64_t>(n, lo + per);
        if (lo >= hi) break;
        pool.emplace_back(run_slice, g, method, pts, lo, hi, out_elev);
    }
    for (auto& th : pool) th.join();
    return 0;
}

// Dump the neighbour selection the reference makes for one query.
//   centre_rule 0: floor centre (cubic fallback, GridH.cpp:233-234,281-289)
//   centre_rule 1: round+clamp centre (kriging, GridH.cpp:333-336)
// Writes found (candidate count, -1 if out of bounds) and, when found >= 4, the four
// (i,j) pairs in post-selection order into sel[8] = {i0,j0,i1,j1,...}; when found < 4 the
// first `found` candidates in enumeration order.
int refh_select4(void* p, int centre_rule, double lon, double lat, int32_t* sel) {
    RefGrid* g = static_cast<RefGrid*>(p);
    for (int k = 0; k < 8; ++k) sel[k] = -1;
    if (lon < g->min_lon || lon > g->max_lon || lat < g->min_lat || lat > g->max_lat) return -1;
    double x = (lon - g->min_lon) / g->lon_step;
    double y = (lat - g->min_lat) / g->lat_step;
    int ci, cj;
    if (centre_rule == 0) {
        ci = static_cast<int>(std::floor(x));
        cj = static_cast<int>(std::floor(y));
    } else {
        ci = std::max(0, std::min(int(std::round(x)), g->n_lon - 1));
        cj = std::max(0, std::min(int(std::round(y)), g->n_lat - 1));
    }
    const int cap = 441;
    std::vector<int> vi(cap), vj(cap);
    std::vector<double> vv(cap), vd(cap);
    int found = findCandidateNeighbors(g->rows, g->n_lon, g->n_lat, x, y, ci, cj, 10, cap,
                                       vi.data(), vj.data(), vv.data(), vd.data());
    if (found >= 4) selectFourNearest(vi.data(), vj.data(), vv.data(), vd.data(), found);
    for (int k = 0; k < std::min(found, 4); ++k) { sel[2 * k] = vi[k]; sel[2 * k + 1] = vj[k]; }
    return found;
}

void refh_select4_batch(void* p, int centre_rule, const double* pts, int64_t n,
                        int32_t* found, int32_t* sel) {
    for (int64_t k = 0; k < n; ++k)
        found[k] = refh_select4(p, centre_rule, pts[3 * k], pts[3 * k + 1], sel + 8 * k);
}

// error_calculator.cpp:5-45 on plain arrays.  which: 0 MAE, 1 RMSE, 2 Max.
double refh_metric(int which, const double* truth, const double* interp, int64_t n) {
    std::vector<Point> a(n), b(n);
    for (int64_t k = 0; k < n; ++k) { a[k] = Point{0, 0, truth[k]}; b[k] = Point{0, 0, interp[k]}; }
    switch (which) {
        case 0: return meanAbsoluteError(a, b);
        case 1: return rootMeanSquareError(a, b);
        default: return maxAbsoluteError(a, b);
    }
}

}  // extern "C"
