// points.cu -- Point-list mode: the kernels behind GridD::batch{Bilinear,Cubic,OrdinaryKriging}
// Interpolate (reference: code/src/GridD.cu:95-236 launching code/src/kernels.cu:173-546), plus the
// NN / IDW extension methods.  One thread per query, exact FP64 path (exact.cuh).
//
// Layout: queries arrive as the reference's wire format, an array of {lon,lat,elev} records
// (Point.h:9-13, 24 B, stride configurable).  A block first stages its 256 records through shared
// memory with fully coalesced 8-byte loads (the reference reads them with a 24-byte stride, wasting
// a third of every sector), then each thread evaluates its query.  Results are written as one
// coalesced double per query; the optional selection dump (4 x (i,j) + found) is for parity tests.
#include "exact.cuh"
#include "launch.h"

namespace auvi {

constexpr int kPointsBlock = 256;

template <typename T, int METHOD>
__global__ void __launch_bounds__(kPointsBlock)
points_kernel(GridView<T> g, const double* __restrict__ pts, int64_t stride_dbl, int64_t n,
              double* __restrict__ out, int32_t* __restrict__ sel, int32_t* __restrict__ found) {
    __shared__ double s_lonlat[kPointsBlock * 2];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kPointsBlock;
    const int live = static_cast<int>(min(static_cast<int64_t>(kPointsBlock), n - base));

    if (stride_dbl == 3) {
        // 256 records = 768 contiguous doubles; keep lon,lat (2 of every 3).
        for (int k = threadIdx.x; k < live * 3; k += kPointsBlock) {
            int rec = k / 3, fld = k - rec * 3;
            double v = __ldg(pts + base * 3 + k);
            if (fld < 2) s_lonlat[rec * 2 + fld] = v;
        }
    } else {
        if (threadIdx.x < live) {
            const double* p = pts + (base + threadIdx.x) * stride_dbl;
            s_lonlat[threadIdx.x * 2] = __ldg(p);
            s_lonlat[threadIdx.x * 2 + 1] = __ldg(p + 1);
        }
    }
    __syncthreads();
    if (threadIdx.x >= live) return;

    const double lon = s_lonlat[threadIdx.x * 2], lat = s_lonlat[threadIdx.x * 2 + 1];
    double x = qnan(), y = qnan();
    if (!outside(g, lon, lat)) {
        x = to_index_space(lon, g.min_lon, g.lon_step);
        y = to_index_space(lat, g.min_lat, g.lat_step);
    }
    const int64_t q = base + threadIdx.x;
    if (sel) {
        Picked p;
        out[q] = interp_exact<T>(g, METHOD, lon, lat, x, y, &p);
        found[q] = p.found;
        int4 a, b;
        const bool has = p.found >= 0;
        a.x = has ? p.i[0] : -1; a.y = has ? p.j[0] : -1; a.z = has ? p.i[1] : -1; a.w = has ? p.j[1] : -1;
        b.x = has ? p.i[2] : -1; b.y = has ? p.j[2] : -1; b.z = has ? p.i[3] : -1; b.w = has ? p.j[3] : -1;
        reinterpret_cast<int4*>(sel)[q * 2] = a;
        reinterpret_cast<int4*>(sel)[q * 2 + 1] = b;
    } else {
        out[q] = interp_exact<T>(g, METHOD, lon, lat, x, y, nullptr);
    }
}

template <typename T>
cudaError_t launch_points_t(const GridView<T>& g, int method, const double* pts, int64_t stride_dbl,
                            int64_t n, double* out, int32_t* sel, int32_t* found, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const unsigned blocks = static_cast<unsigned>((n + kPointsBlock - 1) / kPointsBlock);
    switch (method) {
#define AUVI_CASE(M) case M: points_kernel<T, M><<<blocks, kPointsBlock, 0, st>>>(g, pts, stride_dbl, n, out, sel, found); break;
        AUVI_CASE(BILINEAR) AUVI_CASE(CUBIC) AUVI_CASE(KRIGING) AUVI_CASE(NN) AUVI_CASE(IDW) AUVI_CASE(BILINEAR_SEARCH)
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
