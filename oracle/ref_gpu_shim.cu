// oracle/ref_gpu_shim.cu -- TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product).
//
// The reference's own GPU code -- code/src/kernels.cu (the three kernels, :173-546) and code/src/GridD.cu (the host
// class that mallocs, copies, launches and frees per call, :95-236) -- compiled UNMODIFIED for sm_100a where the files
// lie under /root/reference (oracle/Makefile, target ref_gpu), behind a C shim.  Both files are pulled into this one
// translation unit by textual inclusion (REF_KERNELS_CU / REF_GRIDD_CU are paths into the read-only checkout), which
// replaces the reference's CUDA_SEPARABLE_COMPILATION (CMakeLists.txt:58): GridD.cu only forward-declares the kernels.
// Result: oracle/_ref_gpu/libgridd_ref_sm100a.so, the secondary comparator SURVEY.md section 2.1 / BASELINE.md ask for:
// "the reference's kernels recompiled for sm_100a on the same box", timed beside points_kernel by bench.py.
//   refd_batch       end to end through GridD::batch* exactly as the drivers time it (test_gebco.cpp:183-196)
//   refd_kernel_ms   the same kernel with the reference's launch shape (256 threads, GridD.cu:116-124) on
//                    device-resident buffers, CUDA events -- kernel only
#include REF_KERNELS_CU
#include REF_GRIDD_CU

#include <chrono>
#include <cstdint>
#include <cstring>

namespace {
struct RefDev {
    GridD* g;
    double* d_grid;
    int n_lat, n_lon;
    double min_lon, max_lon, min_lat, max_lat, lon_step, lat_step;
};
}  // namespace

extern "C" {

void* refd_create(const double* rowmajor, int n_lat, int n_lon, double min_lon, double max_lon, double min_lat, double max_lat) {
    std::vector<std::vector<double>> rows(n_lat, std::vector<double>(n_lon));
    for (int j = 0; j < n_lat; ++j) std::memcpy(rows[j].data(), rowmajor + static_cast<size_t>(j) * n_lon, sizeof(double) * n_lon);
    RefDev* r = new RefDev;
    r->g = new GridD(min_lon, max_lon, n_lon, min_lat, max_lat, n_lat, rows);
    r->n_lat = n_lat; r->n_lon = n_lon;
    r->min_lon = min_lon; r->max_lon = max_lon; r->min_lat = min_lat; r->max_lat = max_lat;
    r->lon_step = (max_lon - min_lon) / (n_lon - 1);              // GridD.cu:52-53
    r->lat_step = (max_lat - min_lat) / (n_lat - 1);
    r->d_grid = nullptr;
    if (cudaMalloc(&r->d_grid, sizeof(double) * n_lat * n_lon) != cudaSuccess ||
        cudaMemcpy(r->d_grid, rowmajor, sizeof(double) * n_lat * n_lon, cudaMemcpyHostToDevice) != cudaSuccess) {
        delete r->g; delete r; return nullptr;
    }
    return r;
}

void refd_destroy(void* p) {
    RefDev* r = static_cast<RefDev*>(p);
    if (!r) return;
    cudaFree(r->d_grid);
    delete r->g;
    delete r;
}

// End to end, as the reference's drivers time it: std::vector<Point> in, std::vector<Point> out.  Returns milliseconds.
double refd_batch(void* p, int method, const double* pts, int64_t n, double* out_elev) {
    RefDev* r = static_cast<RefDev*>(p);
    std::vector<Point> q(static_cast<size_t>(n));
    for (int64_t k = 0; k < n; ++k) q[k] = Point{pts[3 * k], pts[3 * k + 1], pts[3 * k + 2]};
    const auto t0 = std::chrono::high_resolution_clock::now();
    std::vector<Point> res = method == 0 ? r->g->batchBilinearInterpolate(q)
                           : method == 1 ? r->g->batchCubicInterpolate(q) : r->g->batchOrdinaryKrigingInterpolate(q);
    const auto t1 = std::chrono::high_resolution_clock::now();
    for (int64_t k = 0; k < n; ++k) out_elev[k] = res[k].elev;
    return std::chrono::duration<double, std::milli>(t1 - t0).count();
}

// Kernel only: points and results resident, `reps` launches between two events.  Returns milliseconds per launch (< 0: error).
double refd_kernel_ms(void* p, int method, const double* pts, int64_t n, int reps, double* out_elev) {
    RefDev* r = static_cast<RefDev*>(p);
    Point* d_pts = nullptr;
    double* d_res = nullptr;
    if (cudaMalloc(&d_pts, sizeof(Point) * n) != cudaSuccess || cudaMalloc(&d_res, sizeof(double) * n) != cudaSuccess) return -1.0;
    cudaMemcpy(d_pts, pts, sizeof(Point) * n, cudaMemcpyHostToDevice);   // n x {lon,lat,elev} doubles == Point[n]
    const int block = 256, grid = static_cast<int>((n + block - 1) / block);   // GridD.cu:116-117
    auto launch = [&] {
        if (method == 0) bilinearInterpolationKernel<<<grid, block>>>(r->d_grid, d_pts, d_res, static_cast<int>(n), r->min_lon, r->max_lon, r->min_lat, r->max_lat, r->n_lon, r->n_lat, r->lon_step, r->lat_step);
        else if (method == 1) cubicInterpolationKernel<<<grid, block>>>(r->d_grid, d_pts, d_res, static_cast<int>(n), r->min_lon, r->max_lon, r->min_lat, r->max_lat, r->n_lon, r->n_lat, r->lon_step, r->lat_step);
        else krigingInterpolationKernel<<<grid, block>>>(r->d_grid, d_pts, d_res, static_cast<int>(n), r->min_lon, r->max_lon, r->min_lat, r->max_lat, r->n_lon, r->n_lat, r->lon_step, r->lat_step);
    };
    launch();
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int k = 0; k < reps; ++k) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const bool ok = cudaGetLastError() == cudaSuccess;
    if (out_elev) cudaMemcpy(out_elev, d_res, sizeof(double) * n, cudaMemcpyDeviceToHost);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_pts); cudaFree(d_res);
    return ok ? ms / reps : -1.0;
}

}  // extern "C"
