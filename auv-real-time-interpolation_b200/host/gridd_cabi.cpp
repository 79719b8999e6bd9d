// gridd_cabi.cpp -- our GridD class (host/GridD.cpp: the reference's signatures over libauvi) behind a few C functions, so that
// scripts can drive the CLASS a user of the reference calls -- std::vector<Point> in, std::vector<Point> out, the copy of the
// query vector included (GridD.cu:101) -- and time it with the drivers' own chrono (test_gebco.cpp:183-196).  Built as
// lib/libgridd_c.so; used by tools/ref_gpu_compare.py beside the same shim around the reference's own GridD
// (oracle/ref_gpu_shim.cu).  Not part of the C ABI of libauvi (include/auvi.h).
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

#include "include/GridD.h"

extern "C" {

void* ourd_create(const double* rowmajor, int n_lat, int n_lon, double min_lon, double max_lon, double min_lat, double max_lat) {
    std::vector<std::vector<double>> rows(n_lat, std::vector<double>(n_lon));
    for (int j = 0; j < n_lat; ++j) std::memcpy(rows[j].data(), rowmajor + static_cast<size_t>(j) * n_lon, sizeof(double) * n_lon);
    return new GridD(min_lon, max_lon, n_lon, min_lat, max_lat, n_lat, rows);
}

void ourd_destroy(void* p) { delete static_cast<GridD*>(p); }

// method 0 bilinear, 1 cubic, 2 kriging; pts = n x {lon,lat,elev}.  Returns the milliseconds of the batch* call alone.
double ourd_batch(void* p, int method, const double* pts, int64_t n, double* out_elev) {
    GridD* g = static_cast<GridD*>(p);
    std::vector<Point> q(static_cast<size_t>(n));
    for (int64_t k = 0; k < n; ++k) q[k] = Point{pts[3 * k], pts[3 * k + 1], pts[3 * k + 2]};
    const auto t0 = std::chrono::high_resolution_clock::now();
    std::vector<Point> res = method == 0 ? g->batchBilinearInterpolate(q)
                           : method == 1 ? g->batchCubicInterpolate(q) : g->batchOrdinaryKrigingInterpolate(q);
    const auto t1 = std::chrono::high_resolution_clock::now();
    for (int64_t k = 0; k < n; ++k) out_elev[k] = res[k].elev;
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

}  // extern "C"
