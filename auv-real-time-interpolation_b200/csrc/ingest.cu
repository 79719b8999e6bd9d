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
    // a work item is 256 consecutive cells of one row: one division per item, none per cell
    const int chunks = (n_lon + 255) / 256;
    const int64_t items = static_cast<int64_t>(n_lat) * chunks;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const int r = static_cast<int>(it / chunks), c = static_cast<int>(it - static_cast<int64_t>(r) * chunks) * 256 + static_cast<int>(threadIdx.x);
        if (c >= n_lon) continue;
        const int64_t k = static_cast<int64_t>(r) * n_lon + c;
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

// A work item is 1024 consecutive cells of one row: one division per item, none per cell (a 65536^2 grid is 4.3 G cells).
template <typename T>
__global__ void __launch_bounds__(256)
mask_hash_kernel(T* __restrict__ z, int64_t ld, int n_lon, int row0, int rows, uint64_t threshold, uint64_t seed,
                 unsigned long long* __restrict__ n_masked) {
    constexpr int kChunk = 1024;
    const int chunks = (n_lon + kChunk - 1) / kChunk;
    const int64_t items = static_cast<int64_t>(rows) * chunks;
    const uint64_t salt = seed * 0xD1342543DE82EF95ull;
    unsigned long long mine = 0;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const int64_t r = it / chunks;
        const int c_lo = static_cast<int>(it - r * chunks) * kChunk, c_hi = min(n_lon, c_lo + kChunk);
        const uint64_t flat0 = static_cast<uint64_t>(r + row0) * static_cast<uint64_t>(n_lon);   // GLOBAL index of the row's first cell
        T* const zr = z + r * ld;
        for (int c = c_lo + static_cast<int>(threadIdx.x); c < c_hi; c += 256) {
            if ((splitmix64((flat0 + c) ^ salt) >> 11) < threshold) {
                zr[c] = static_cast<T>(qnan());
                ++mine;
            }
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


// ---- CSV matrix text -> grid, parsed on the device -------------------------------------------------------------
// The reference's driver reads its grids from CSV text with getline + stod into a vector<vector<double>>
// (test_gebco.cpp:19-40, test_interpolation.cpp readGridCSV).  Here the text is uploaded as it is and parsed in four
// passes: (1) delimiters (',' and '\n') counted per 4 KiB block, (2) exclusive scan of the counts, (3) the position of
// every delimiter written out, (4) one thread per field converts its characters.  Decimal -> double is Clinger's fast
// path: up to 15 significant digits and a power of ten up to 10^22 are exact in FP64, so one multiplication or division
// is correctly rounded -- the value stod returns.  Fields outside the fast path (more digits, larger exponents) are
// flagged and converted on the host with strtod by the caller (api.cu), so every value equals the reference's.
namespace csv {

constexpr int kBlockBytes = 4096, kThreads = 256, kPerThread = kBlockBytes / kThreads;

__device__ __forceinline__ bool is_delim(char ch) { return ch == ',' || ch == '\n'; }

__global__ void __launch_bounds__(kThreads)
count_kernel(const char* __restrict__ text, int64_t n, int* __restrict__ counts) {
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kBlockBytes + threadIdx.x * kPerThread;
    int c = 0;
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) c += (base + k < n) && is_delim(text[base + k]);
    __shared__ int warp_sum[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kThreads / 32; ++w) t += warp_sum[w];
        counts[blockIdx.x] = t;
    }
}

// one CTA walks the block counts in slices of 1024 with a running carry: offsets[b] = sum of counts before b
__global__ void __launch_bounds__(1024)
scan_kernel(const int* __restrict__ counts, int64_t n_blocks, int64_t* __restrict__ offsets, int64_t* __restrict__ total) {
    __shared__ int64_t warp_tot[32];
    __shared__ int64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t b0 = 0; b0 < n_blocks; b0 += 1024) {
        const int64_t b = b0 + threadIdx.x;
        const int64_t v = b < n_blocks ? counts[b] : 0;
        int64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int64_t w = warp_tot[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_tot[lane] = wi - w;                                // exclusive over warps
        }
        __syncthreads();
        const int64_t carry = carry_s;
        if (b < n_blocks) offsets[b] = carry + warp_tot[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

__global__ void __launch_bounds__(kThreads)
positions_kernel(const char* __restrict__ text, int64_t n, const int64_t* __restrict__ offsets, int64_t* __restrict__ pos,
                 int64_t max_fields) {
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kBlockBytes + threadIdx.x * kPerThread;
    int mine = 0;
    bool d[kPerThread];
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) { d[k] = (base + k < n) && is_delim(text[base + k]); mine += d[k]; }
    // exclusive prefix of `mine` over the CTA
    __shared__ int warp_sum[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += warp_sum[w];
    int64_t at = offsets[blockIdx.x] + before + incl - mine;
#pragma unroll
    for (int k = 0; k < kPerThread; ++k)
        if (d[k]) { if (at < max_fields) pos[at] = base + k; ++at; }
}

// status bits per run: 1 = a row does not have n_cols fields, 2 = a field that is not a number, 4 = slow fields exist
template <typename T>
__global__ void __launch_bounds__(256)
parse_kernel(const char* __restrict__ text, const int64_t* __restrict__ pos, int64_t n_fields, int n_cols, T* __restrict__ out,
             int64_t ld, unsigned* __restrict__ status, unsigned long long* __restrict__ n_slow, int64_t* __restrict__ slow_idx,
             int64_t slow_cap) {
    const double p10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16,
                            1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    for (int64_t f = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; f < n_fields; f += static_cast<int64_t>(gridDim.x) * 256) {
        int64_t a = f ? pos[f - 1] + 1 : 0, b = pos[f];            // [a, b) = the field's characters
        const bool row_end = text[b] == '\n';
        if (row_end != ((f + 1) % n_cols == 0)) { atomicOr(status, 1u); continue; }
        while (a < b && (text[a] == ' ' || text[a] == '\t')) ++a;
        while (b > a && (text[b - 1] == ' ' || text[b - 1] == '\t' || text[b - 1] == '\r')) --b;
        double v = 0.0;
        bool ok = a < b, slow = false;
        if (ok) {
            bool neg = false;
            int64_t q = a;
            if (text[q] == '-' || text[q] == '+') { neg = text[q] == '-'; ++q; }
            const int len = static_cast<int>(b - q);
            auto lower = [&](int64_t at) { const char ch = text[at]; return (ch >= 'A' && ch <= 'Z') ? static_cast<char>(ch + 32) : ch; };
            if (len == 3 && lower(q) == 'n' && lower(q + 1) == 'a' && lower(q + 2) == 'n') v = qnan();
            else if ((len == 3 || len == 8) && lower(q) == 'i' && lower(q + 1) == 'n' && lower(q + 2) == 'f') v = __longlong_as_double(0x7ff0000000000000LL);
            else {
                uint64_t m = 0;
                int digits = 0, seen = 0, e10 = 0;
                bool dot = false;
                for (; q < b; ++q) {
                    const char ch = text[q];
                    if (ch >= '0' && ch <= '9') {
                        ++seen;
                        if (m || ch != '0') {
                            if (digits < 19) { m = m * 10 + static_cast<uint64_t>(ch - '0'); ++digits; if (dot) --e10; }
                            else { slow = true; if (!dot) ++e10; }
                        } else if (dot) --e10;
                    } else if (ch == '.' && !dot) dot = true;
                    else break;
                }
                if (!seen) ok = false;
                if (ok && q < b && (text[q] == 'e' || text[q] == 'E')) {
                    ++q;
                    bool eneg = false;
                    if (q < b && (text[q] == '-' || text[q] == '+')) { eneg = text[q] == '-'; ++q; }
                    int e = 0, ed = 0;
                    for (; q < b && text[q] >= '0' && text[q] <= '9'; ++q) { if (e < 100000) e = e * 10 + (text[q] - '0'); ++ed; }
                    if (!ed) ok = false;
                    e10 += eneg ? -e : e;
                }
                if (ok && q != b) ok = false;                       // trailing characters: not a plain number
                if (ok) {
                    if (m == 0) v = 0.0;
                    else if (!slow && m < (1ull << 53) && e10 >= -22 && e10 <= 22)
                        v = e10 < 0 ? ddiv(static_cast<double>(m), p10[-e10]) : dmul(static_cast<double>(m), p10[e10]);
                    else if (!slow && m < (1ull << 53) && e10 > 22 && e10 <= 22 + 15 && static_cast<double>(m) * p10[e10 - 22] < 9007199254740992.0)
                        v = dmul(dmul(static_cast<double>(m), p10[e10 - 22]), p10[22]);   // still exact: m * 10^(e-22) < 2^53
                    else slow = true;
                }
            }
            if (neg) v = -v;
        }
        if (!ok) { atomicOr(status, 2u); continue; }
        const int64_t r = f / n_cols, c = f % n_cols;
        if (slow) {
            atomicOr(status, 4u);
            const unsigned long long slot = atomicAdd(n_slow, 1ull);
            if (static_cast<int64_t>(slot) < slow_cap) slow_idx[slot] = f;
            v = qnan();
        }
        out[r * ld + c] = static_cast<T>(v);
    }
}

template <typename T>
__global__ void patch_kernel(T* __restrict__ out, int64_t ld, int n_cols, const int64_t* __restrict__ idx,
                             const double* __restrict__ val, int64_t n) {
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; k < n; k += static_cast<int64_t>(gridDim.x) * 256) {
        const int64_t f = idx[k];
        out[(f / n_cols) * ld + f % n_cols] = static_cast<T>(val[k]);
    }
}

}  // namespace csv

size_t csv_block_count(int64_t n_bytes) { return static_cast<size_t>((n_bytes + csv::kBlockBytes - 1) / csv::kBlockBytes); }

cudaError_t launch_csv_index(const char* text, int64_t n_bytes, int* counts, int64_t* offsets, int64_t* total, int64_t* pos,
                             int64_t max_fields, cudaStream_t st) {
    const unsigned blocks = static_cast<unsigned>(csv_block_count(n_bytes));
    csv::count_kernel<<<blocks, csv::kThreads, 0, st>>>(text, n_bytes, counts);
    csv::scan_kernel<<<1, 1024, 0, st>>>(counts, blocks, offsets, total);
    csv::positions_kernel<<<blocks, csv::kThreads, 0, st>>>(text, n_bytes, offsets, pos, max_fields);
    return cudaGetLastError();
}

cudaError_t launch_csv_parse(const char* text, const int64_t* pos, int64_t n_fields, int n_cols, void* out, int64_t ld, int dtype,
                             unsigned* status, unsigned long long* n_slow, int64_t* slow_idx, int64_t slow_cap, cudaStream_t st) {
    const int blocks = grid_blocks(n_fields);
    if (dtype == DT_F64)
        csv::parse_kernel<double><<<blocks, 256, 0, st>>>(text, pos, n_fields, n_cols, static_cast<double*>(out), ld, status, n_slow, slow_idx, slow_cap);
    else
        csv::parse_kernel<float><<<blocks, 256, 0, st>>>(text, pos, n_fields, n_cols, static_cast<float*>(out), ld, status, n_slow, slow_idx, slow_cap);
    return cudaGetLastError();
}

cudaError_t launch_csv_patch(void* out, int64_t ld, int n_cols, int dtype, const int64_t* idx, const double* val, int64_t n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int blocks = grid_blocks(n);
    if (dtype == DT_F64) csv::patch_kernel<double><<<blocks, 256, 0, st>>>(static_cast<double*>(out), ld, n_cols, idx, val, n);
    else csv::patch_kernel<float><<<blocks, 256, 0, st>>>(static_cast<float*>(out), ld, n_cols, idx, val, n);
    return cudaGetLastError();
}

}  // namespace auvi
