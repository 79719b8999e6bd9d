// points.cu -- Point-list mode: the kernels behind GridD::batch{Bilinear,Cubic,OrdinaryKriging}
// Interpolate (reference: code/src/GridD.cu:95-236 launching code/src/kernels.cu:173-546), plus the
// NN / IDW extension methods.  One thread per query, exact FP64 path (exact.cuh).
//
// Layout: queries arrive as the reference's wire format, an array of {lon,lat,elev} records (Point.h:9-13, 24 B, stride
// configurable; the host path packs them to 16-byte {lon,lat}).  A thread loads its own record straight into registers:
// one 16-byte load for packed records, two 8-byte loads otherwise (the second hits the sector the first brought into L1;
// round 1 staged the records through shared memory, which cost a barrier and measured 8 % slower than the reference's own
// bilinear kernel on the same box: profiles/r02_points_vs_reference_gpu.txt).  BILINEAR -- four gathers and a dozen
// flops, latency-bound on the L2 gathers -- evaluates TWO queries per thread so that eight gathers are in flight.
// Results are written as one coalesced double per query; the optional selection dump (4 x (i,j) + found) is for parity tests.
#include "exact.cuh"
#include "launch.h"

namespace auvi {

constexpr int kPointsBlock = 256;

template <int METHOD> struct PointsPerThread { static constexpr int value = METHOD == BILINEAR ? 2 : 1; };

template <typename T, int METHOD>
__global__ void __launch_bounds__(kPointsBlock)
points_kernel(GridView<T> g, const double* __restrict__ pts, int64_t stride_dbl, int64_t n,
              double* __restrict__ out, int32_t* __restrict__ sel, int32_t* __restrict__ found) {
    constexpr int PPT = PointsPerThread<METHOD>::value;
    const int64_t base = static_cast<int64_t>(blockIdx.x) * (kPointsBlock * PPT);
    double lon[PPT], lat[PPT];
    bool live[PPT];
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
        const int64_t q = base + u * kPointsBlock + threadIdx.x;
        live[u] = q < n;
        lon[u] = 0.0; lat[u] = 0.0;
        if (!live[u]) continue;
        if (stride_dbl == 2 && (reinterpret_cast<uintptr_t>(pts) & 15) == 0) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(pts) + q);
            lon[u] = v.x; lat[u] = v.y;
        } else {
            const double* p = pts + q * stride_dbl;
            lon[u] = __ldg(p); lat[u] = __ldg(p + 1);
        }
    }
    double res[PPT];
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
        if (!live[u]) continue;
        const int64_t q = base + u * kPointsBlock + threadIdx.x;
        double x = qnan(), y = qnan();
        if (!outside(g, lon[u], lat[u])) {
            x = to_index_space(lon[u], g.min_lon, g.lon_step);
            y = to_index_space(lat[u], g.min_lat, g.lat_step);
        }
        if (sel) {
            Picked p;
            res[u] = interp_exact<T>(g, METHOD, lon[u], lat[u], x, y, &p);
            found[q] = p.found;
            int4 a, b;
            const bool has = p.found >= 0;
            a.x = has ? p.i[0] : -1; a.y = has ? p.j[0] : -1; a.z = has ? p.i[1] : -1; a.w = has ? p.j[1] : -1;
            b.x = has ? p.i[2] : -1; b.y = has ? p.j[2] : -1; b.z = has ? p.i[3] : -1; b.w = has ? p.j[3] : -1;
            reinterpret_cast<int4*>(sel)[q * 2] = a;
            reinterpret_cast<int4*>(sel)[q * 2 + 1] = b;
        } else {
            res[u] = interp_exact<T>(g, METHOD, lon[u], lat[u], x, y, nullptr);
        }
    }
#pragma unroll
    for (int u = 0; u < PPT; ++u)
        if (live[u]) out[base + u * kPointsBlock + threadIdx.x] = res[u];
}

template <typename T>
cudaError_t launch_points_t(const GridView<T>& g, int method, const double* pts, int64_t stride_dbl,
                            int64_t n, double* out, int32_t* sel, int32_t* found, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    switch (method) {
#define AUVI_CASE(M) case M: { const int64_t per = kPointsBlock * PointsPerThread<M>::value; \
        points_kernel<T, M><<<static_cast<unsigned>((n + per - 1) / per), kPointsBlock, 0, st>>>(g, pts, stride_dbl, n, out, sel, found); } break;
        AUVI_CASE(BILINEAR) AUVI_CASE(CUBIC) AUVI_CASE(KRIGING) AUVI_CASE(NN) AUVI_CASE(IDW) AUVI_CASE(BILINEAR_SEARCH) AUVI_CASE(IDW_KNN)
#undef AUVI_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_points(const GridDesc& d, int method, const double* pts, int64_t stride_dbl, int64_t n,
                          double* out, int32_t* sel, int32_t* found, cudaStream_t st) {
    if (d.dtype == DT_F64) return launch_points_t<double>(make_view<double>(d), method, pts, stride_dbl, n, out, sel, found, st);
    return launch_points_t<float>(make_view<float>(d), method, pts, stride_dbl, n, out, sel, found, st);
}

}  // namespace auvi
