// ingest.cu -- device side of the Grid-B data preparation: the work the
// reference does on the host before its drivers start -- netCDF4 read + `iloc[::-1]` row flip
// (code/subset_bathymetry.py:8-18), setting the removed cells to NaN (:78-85), and, at synthetic scale,
// drawing the mask itself -- done on the GPU so that a grid never exists as CSV text or as a
// vector<vector<double>> (test_gebco.cpp:19-40 parses 34 GB of text at BASELINE config 4's size).
//
//   decode_raw_kernel   file-order elements (NetCDF: big-endian int16 / int32 / float / double) -> the grid's
//                       storage type, optional row flip, scale/offset; GEBCO int16 uploads 2 B per cell instead of 8
//   mask_cells_kernel   gather the truth of the listed cells, then overwrite them with NaN
//   mask_hash_kernel    counter-based mask: a cell is removed iff hash(global flat index, seed) < fraction, so
//                       every rank regenerates its slab + halo of the same global mask without communication
// All three are one pass over their data: HBM-bound, element-wise.
#include "launch.h"

namespace auvi {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

template <typename T>
__global__ void __launch_bounds__(256)
decode_raw_kernel(const unsigned char* __restrict__ raw, int nc_type, int big_endian, int flip_rows, double scale,
                  double offset, int n_lat, int n_lon, T* __restrict__ out, int64_t ld) {
    const int64_t total = static_cast<int64_t>(n_lat) * n_lon;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; k < total; k += static_cast<int64_t>(gridDim.x) * 256) {
        const int r = static_cast<int>(k / n_lon), c = static_cast<int>(k - static_cast<int64_t>(r) * n_lon);
        double v;
        if (nc_type == 3) {
            uint16_t u = __ldg(reinterpret_cast<const uint16_t*>(raw) + k);
            if (big_endian) u = static_cast<uint16_t>((u << 8) | (u >> 8));
            v = static_cast<double>(static_cast<int16_t>(u));
        } else if (nc_type == 4 || nc_type == 5) {
            uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(raw) + k);
            if (big_endian) u = __byte_perm(u, 0, 0x0123);
            v = nc_type == 4 ? static_cast<double>(static_cast<int32_t>(u)) : static_cast<double>(__uint_as_float(u));
        } else {
            const uint32_t* q = reinterpret_cast<const uint32_t*>(raw) + 2 * k;   // 4-byte aligned is all a file offset gives
            uint32_t a = __ldg(q), b = __ldg(q + 1);
            const uint64_t u = big_endian ? (static_cast<uint64_t>(__byte_perm(a, 0, 0x0123)) << 32) | __byte_perm(b, 0, 0x0123)
                                          : (static_cast<uint64_t>(b) << 32) | a;
            v = __longlong_as_double(static_cast<long long>(u));
        }
        v = dadd(dmul(v, scale), offset);                           // identity for GEBCO (scale 1, offset 0): exact
        const int rr = flip_rows ? n_lat - 1 - r : r;
        out[static_cast<int64_t>(rr) * ld + c] = static_cast<T>(v);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
mask_cells_kernel(T* __restrict__ z, int64_t ld, int n_lon, int row0, int rows, const int64_t* __restrict__ idx, int64_t n,
                  T* __restrict__ truth) {
    for (int64_t t = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; t < n; t += static_cast<int64_t>(gridDim.x) * 256) {
        const int64_t k = __ldg(idx + t);
        const int64_t r = k / n_lon - row0;
        const int c = static_cast<int>(k % n_lon);
        if (r < 0 || r >= rows) continue;                          // another rank's slab
        T* cell = z + r * ld + c;
        if (truth) truth[t] = *cell;
        *cell = static_cast<T>(qnan());
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
mask_hash_kernel(T* __restrict__ z, int64_t ld, int n_lon, int row0, int rows, uint64_t threshold, uint64_t seed,
                 unsigned long long* __restrict__ n_masked) {
    const int64_t total = static_cast<int64_t>(rows) * n_lon;
    unsigned long long mine = 0;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; k < total; k += static_cast<int64_t>(gridDim.x) * 256) {
        const int64_t r = k / n_lon;
        const int c = static_cast<int>(k - r * n_lon);
        const uint64_t flat = static_cast<uint64_t>(r + row0) * static_cast<uint64_t>(n_lon) + c;   // GLOBAL index
        if ((splitmix64(flat ^ (seed * 0xD1342543DE82EF95ull)) >> 11) < threshold) {
            z[r * ld + c] = static_cast<T>(qnan());
            ++mine;
        }
    }
    if (n_masked) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_masked, mine);
    }
}

static int grid_blocks(int64_t n) {
    int64_t b = (n + 255) / 256;
    return static_cast<int>(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

cudaError_t launch_decode_raw(const void* raw, int nc_type, int big_endian, int flip_rows, double scale, double offset,
                              int n_lat, int n_lon, void* out, int64_t ld, int dtype, cudaStream_t st) {
    const int blocks = grid_blocks(static_cast<int64_t>(n_lat) * n_lon);
    const unsigned char* r = static_cast<const unsigned char*>(raw);
    if (dtype == DT_F64)
        decode_raw_kernel<double><<<blocks, 256, 0, st>>>(r, nc_type, big_endian, flip_rows, scale, offset, n_lat, n_lon, static_cast<double*>(out), ld);
    else
        decode_raw_kernel<float><<<blocks, 256, 0, st>>>(r, nc_type, big_endian, flip_rows, scale, offset, n_lat, n_lon, static_cast<float*>(out), ld);
    return cudaGetLastError();
}

cudaError_t launch_mask_cells(const GridDesc& d, const int64_t* idx, int64_t n, void* truth, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int blocks = grid_blocks(n);
    if (d.dtype == DT_F64)
        mask_cells_kernel<double><<<blocks, 256, 0, st>>>(static_cast<double*>(const_cast<void*>(d.z)), d.ld, d.n_lon, d.row0, d.rows, idx, n, static_cast<double*>(truth));
    else
        mask_cells_kernel<float><<<blocks, 256, 0, st>>>(static_cast<float*>(const_cast<void*>(d.z)), d.ld, d.n_lon, d.row0, d.rows, idx, n, static_cast<float*>(truth));
    return cudaGetLastError();
}

cudaError_t launch_mask_hash(const GridDesc& d, double fraction, uint64_t seed, unsigned long long* n_masked, cudaStream_t st) {
    // 53-bit uniform < fraction  <=>  (hash >> 11) < fraction * 2^53
    const double f = fraction < 0.0 ? 0.0 : (fraction > 1.0 ? 1.0 : fraction);
    const uint64_t threshold = static_cast<uint64_t>(f * 9007199254740992.0);
    const int blocks = grid_blocks(static_cast<int64_t>(d.rows) * d.n_lon);
    if (d.dtype == DT_F64)
        mask_hash_kernel<double><<<blocks, 256, 0, st>>>(static_cast<double*>(const_cast<void*>(d.z)), d.ld, d.n_lon, d.row0, d.rows, threshold, seed, n_masked);
    else
        mask_hash_kernel<float><<<blocks, 256, 0, st>>>(static_cast<float*>(const_cast<void*>(d.z)), d.ld, d.n_lon, d.row0, d.rows, threshold, seed, n_masked);
    return cudaGetLastError();
}

}  // namespace auvi
