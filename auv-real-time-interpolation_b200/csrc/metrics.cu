// metrics.cu -- device-side error metrics: the reduction behind meanAbsoluteError /
// rootMeanSquareError / maxAbsoluteError (reference: code/src/error_calculator.cpp:5-45).
//
// Convention kept from the reference: an estimate that is NaN adds nothing to the MAE / RMSE
// numerators but still counts in the denominator n (error_calculator.cpp:12-16, :26-31); the
// maximum ignores NaN because `NaN > m` is false (:40-43).  Pure HBM stream: 2 x sizeof(T) bytes per
// element, one pass of 16-byte loads.  Two-stage and deterministic (fixed partition, fixed tree), no atomics.
#include "launch.h"
#include <type_traits>

namespace auvi {

constexpr int kMetBlock = 256;
constexpr int kMetBlocks = 148 * 8;
constexpr int kMetChunk = 2 * kMetBlock * 4;   // elements per work item (FP32: two float4 per thread and array)

struct Partial { double sum_abs, sum_sq, max_abs, n_nan, cnt; };

__device__ __forceinline__ void combine(Partial& a, const Partial& b) {
    a.sum_abs += b.sum_abs; a.sum_sq += b.sum_sq; a.max_abs = fmax(a.max_abs, b.max_abs); a.n_nan += b.n_nan;
    a.cnt += b.cnt;
}

__device__ __forceinline__ Partial block_reduce(Partial v) {
    __shared__ Partial s[kMetBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Partial w;
        w.sum_abs = __shfl_down_sync(0xffffffffu, v.sum_abs, o);
        w.sum_sq = __shfl_down_sync(0xffffffffu, v.sum_sq, o);
        w.max_abs = __shfl_down_sync(0xffffffffu, v.max_abs, o);
        w.n_nan = __shfl_down_sync(0xffffffffu, v.n_nan, o);
        w.cnt = __shfl_down_sync(0xffffffffu, v.cnt, o);
        combine(v, w);
    }
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        Partial z{0.0, 0.0, 0.0, 0.0, 0.0};
        v = threadIdx.x < kMetBlock / 32 ? s[threadIdx.x] : z;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            Partial w;
            w.sum_abs = __shfl_down_sync(0xffffffffu, v.sum_abs, o);
            w.sum_sq = __shfl_down_sync(0xffffffffu, v.sum_sq, o);
            w.max_abs = __shfl_down_sync(0xffffffffu, v.max_abs, o);
            w.n_nan = __shfl_down_sync(0xffffffffu, v.n_nan, o);
        w.cnt = __shfl_down_sync(0xffffffffu, v.cnt, o);
            combine(v, w);
        }
    }
    return v;
}

// VEC: both arrays 16-byte aligned -- two 16-byte vectors of each per thread and work item (kMetChunk elements).
template <typename T, bool VEC>
__global__ void __launch_bounds__(kMetBlock)
metrics_partial_kernel(const T* __restrict__ truth, const T* __restrict__ est, int64_t n, Partial* __restrict__ part) {
    constexpr int V = 16 / static_cast<int>(sizeof(T));
    using Vec = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    Partial acc{0.0, 0.0, 0.0, 0.0, 0.0};
    auto cell = [&](T tv, T ev) {
        const double t = static_cast<double>(tv), e = static_cast<double>(ev);
        const double d = fabs(t - e);
        if (isnan(e)) acc.n_nan += 1.0;
        else { acc.sum_abs += d; acc.sum_sq += d * d; }
        if (d > acc.max_abs) acc.max_abs = d;                  // false for NaN, like the reference
        acc.cnt += 1.0;
    };
    const int64_t items = (n + kMetChunk - 1) / kMetChunk;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const int64_t lo = it * kMetChunk, hi = min(n, lo + kMetChunk);
        if (VEC) {
#pragma unroll
            for (int u = 0; u < kMetChunk / (kMetBlock * V); ++u) {
                const int64_t k = lo + (u * kMetBlock + static_cast<int>(threadIdx.x)) * V;
                if (k + V <= hi) {
                    const Vec t = __ldcs(reinterpret_cast<const Vec*>(truth + k));
                    const Vec e = __ldcs(reinterpret_cast<const Vec*>(est + k));
                    const T* const tp = reinterpret_cast<const T*>(&t);
                    const T* const ep = reinterpret_cast<const T*>(&e);
#pragma unroll
                    for (int j = 0; j < V; ++j) cell(tp[j], ep[j]);
                } else {
                    for (int64_t j = k; j < hi; ++j) cell(__ldcs(truth + j), __ldcs(est + j));
                }
            }
        } else {
            for (int64_t k = lo + threadIdx.x; k < hi; k += kMetBlock) cell(__ldcs(truth + k), __ldcs(est + k));
        }
    }
    acc = block_reduce(acc);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// The same sums over the cells a gap fill produced: cell (r,c) counts iff the MASKED grid holds NaN there, the
// estimate is the filled grid's cell and the truth the unmasked grid's -- the RMSE of test_gebco.cpp:150-230 without
// gathering the removed cells into point lists.  3 x sizeof(T) bytes per cell, a pure
// HBM stream: a work item is kMetChunk consecutive cells of one row (one division per item, none per cell); a thread
// takes two 16-byte vectors of each of the three arrays per item (six loads in flight), so the three streams run at the
// copy bandwidth (round 2: scalar loads + a 64-bit division per cell ran at 0.24 of it).  VEC = false: rows that are not
// 16-byte aligned (odd pitch, odd base) take the same walk with scalar loads.  Deterministic: fixed partition, fixed tree.

template <typename T>
__device__ __forceinline__ void metrics_cell(T m, T tv, T ev, double& sum_abs, double& sum_sq, double& max_abs, int& n_nan, int& cnt) {
    if (m == m) return;                                            // the masked grid kept this cell
    const double t = static_cast<double>(tv), e = static_cast<double>(ev);
    const double d = fabs(t - e);
    if (isnan(e)) n_nan += 1;
    else { sum_abs += d; sum_sq += d * d; }
    if (d > max_abs) max_abs = d;                                  // false for NaN, like the reference
    cnt += 1;
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(kMetBlock)
metrics_masked_partial_kernel(const T* __restrict__ masked, int64_t ld_m, const T* __restrict__ filled, int64_t ld_f,
                              const T* __restrict__ truth, int64_t ld_t, int64_t rows, int cols, Partial* __restrict__ part) {
    constexpr int V = 16 / static_cast<int>(sizeof(T));           // cells per 16-byte vector
    using Vec = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    const int chunks = (cols + kMetChunk - 1) / kMetChunk;
    const int64_t items = rows * chunks;
    Partial acc{0.0, 0.0, 0.0, 0.0, 0.0};
    double sum_abs = 0.0, sum_sq = 0.0, max_abs = 0.0;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const int64_t r = it / chunks;
        const int c_lo = static_cast<int>(it - r * chunks) * kMetChunk;
        const int c_hi = min(cols, c_lo + kMetChunk);
        const T* const mr = masked + r * ld_m;
        const T* const fr = filled + r * ld_f;
        const T* const tr = truth + r * ld_t;
        int n_nan = 0, cnt = 0;
        if (VEC) {
#pragma unroll
            for (int u = 0; u < kMetChunk / (kMetBlock * V); ++u) {
                const int c = c_lo + (u * kMetBlock + static_cast<int>(threadIdx.x)) * V;
                if (c + V <= c_hi) {
                    const Vec m = __ldcs(reinterpret_cast<const Vec*>(mr + c));
                    const Vec f = __ldcs(reinterpret_cast<const Vec*>(fr + c));
                    const Vec t = __ldcs(reinterpret_cast<const Vec*>(tr + c));
                    const T* const mp = reinterpret_cast<const T*>(&m);
                    const T* const fp = reinterpret_cast<const T*>(&f);
                    const T* const tp = reinterpret_cast<const T*>(&t);
#pragma unroll
                    for (int k = 0; k < V; ++k) metrics_cell(mp[k], tp[k], fp[k], sum_abs, sum_sq, max_abs, n_nan, cnt);
                } else {
                    for (int k = c; k < c_hi; ++k) metrics_cell(__ldcs(mr + k), __ldcs(tr + k), __ldcs(fr + k), sum_abs, sum_sq, max_abs, n_nan, cnt);
                }
            }
        } else {
            for (int c = c_lo + static_cast<int>(threadIdx.x); c < c_hi; c += kMetBlock)
                metrics_cell(__ldcs(mr + c), __ldcs(tr + c), __ldcs(fr + c), sum_abs, sum_sq, max_abs, n_nan, cnt);
        }
        acc.n_nan += static_cast<double>(n_nan);
        acc.cnt += static_cast<double>(cnt);
    }
    acc.sum_abs = sum_abs; acc.sum_sq = sum_sq; acc.max_abs = max_abs;
    acc = block_reduce(acc);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(kMetBlock)
metrics_final_kernel(const Partial* __restrict__ part, int n_part, double* __restrict__ result5) {
    Partial acc{0.0, 0.0, 0.0, 0.0, 0.0};
    for (int k = threadIdx.x; k < n_part; k += kMetBlock) combine(acc, part[k]);
    acc = block_reduce(acc);
    if (threadIdx.x == 0) {
        result5[0] = acc.sum_abs; result5[1] = acc.sum_sq; result5[2] = acc.max_abs; result5[3] = acc.n_nan; result5[4] = acc.cnt;
    }
}

size_t metrics_scratch_bytes() { return sizeof(Partial) * kMetBlocks; }

cudaError_t launch_metrics(const void* truth, const void* est, int dtype, int64_t n, void* scratch,
                           double* result4, cudaStream_t st, LaunchInfo* info) {
    const int64_t want = (n + kMetChunk - 1) / kMetChunk;
    const int blocks = static_cast<int>(want < 1 ? 1 : (want > kMetBlocks ? kMetBlocks : want));
    Partial* part = static_cast<Partial*>(scratch);
    const bool vec = reinterpret_cast<uintptr_t>(truth) % 16 == 0 && reinterpret_cast<uintptr_t>(est) % 16 == 0;
#define AUVI_M(T, V) metrics_partial_kernel<T, V><<<blocks, kMetBlock, 0, st>>>(static_cast<const T*>(truth), static_cast<const T*>(est), n, part)
    if (dtype == DT_F64) { if (vec) AUVI_M(double, true); else AUVI_M(double, false); }
    else { if (vec) AUVI_M(float, true); else AUVI_M(float, false); }
#undef AUVI_M
    metrics_final_kernel<<<1, kMetBlock, 0, st>>>(part, blocks, result4);
    if (info) info->launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_metrics_masked(const void* masked, int64_t ld_m, const void* filled, int64_t ld_f, const void* truth,
                                  int64_t ld_t, int dtype, int64_t rows, int cols, void* scratch, double* result5,
                                  cudaStream_t st, LaunchInfo* info) {
    const int64_t want = rows * ((cols + kMetChunk - 1) / kMetChunk);
    const int blocks = static_cast<int>(want < 1 ? 1 : (want > kMetBlocks ? kMetBlocks : want));
    Partial* part = static_cast<Partial*>(scratch);
    const size_t es = dtype == DT_F64 ? 8 : 4;
    auto aligned = [es](const void* p, int64_t ld) { return reinterpret_cast<uintptr_t>(p) % 16 == 0 && (ld * es) % 16 == 0; };
    const bool vec = aligned(masked, ld_m) && aligned(filled, ld_f) && aligned(truth, ld_t);
#define AUVI_MM(T, V) metrics_masked_partial_kernel<T, V><<<blocks, kMetBlock, 0, st>>>( \
        static_cast<const T*>(masked), ld_m, static_cast<const T*>(filled), ld_f, static_cast<const T*>(truth), ld_t, rows, cols, part)
    if (dtype == DT_F64) { if (vec) AUVI_MM(double, true); else AUVI_MM(double, false); }
    else { if (vec) AUVI_MM(float, true); else AUVI_MM(float, false); }
#undef AUVI_MM
    metrics_final_kernel<<<1, kMetBlock, 0, st>>>(part, blocks, result5);
    if (info) info->launches += 2;
    return cudaGetLastError();
}

// ---- empirical semivariances for the fitted variogram (opt-in AUVI_KRIGING_FITTED) -------------
// The reference hard-codes its variogram (nugget 1, sill 100, range 10 degrees: GridH.cpp:371-376).  The opt-in fits the
// same exponential model to the grid itself.  This kernel delivers the data of that fit: for the lags k = 1, 2, 4, 8 cells
// along each axis, the sum of squared differences and the number of pairs of VALID cells k apart.  One pass, each cell
// reads its eight partners (L1/L2 hits); deterministic: fixed partition over kVgBlocks blocks, fixed reduction trees.
// sums16 layout: [axis (0 lon, 1 lat)][lag index 0..3][0 sum of squares, 1 pair count].
constexpr int kVgBlocks = 148 * 4;

template <typename T>
__global__ void __launch_bounds__(kMetBlock)
variogram_partial_kernel(const T* __restrict__ z, int64_t ld, int rows, int n_lon, double* __restrict__ part) {
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    const int chunks = (n_lon + kMetBlock - 1) / kMetBlock;       // work item: kMetBlock consecutive cells of one row
    const int64_t items = static_cast<int64_t>(rows) * chunks;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const int r = static_cast<int>(it / chunks), i = static_cast<int>(it - static_cast<int64_t>(r) * chunks) * kMetBlock + static_cast<int>(threadIdx.x);
        if (i >= n_lon) continue;
        const double v = static_cast<double>(__ldg(z + static_cast<int64_t>(r) * ld + i));
        if (isnan(v)) continue;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const int k = 1 << l;
            if (i + k < n_lon) {
                const double w = static_cast<double>(__ldg(z + static_cast<int64_t>(r) * ld + i + k));
                if (!isnan(w)) { const double d = v - w; acc[2 * l] = fma(d, d, acc[2 * l]); acc[2 * l + 1] += 1.0; }
            }
            if (r + k < rows) {
                const double w = static_cast<double>(__ldg(z + static_cast<int64_t>(r + k) * ld + i));
                if (!isnan(w)) { const double d = v - w; acc[8 + 2 * l] = fma(d, d, acc[8 + 2 * l]); acc[8 + 2 * l + 1] += 1.0; }
            }
        }
    }
    __shared__ double s[kMetBlock / 32][16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double v = 0.0;
        for (int w = 0; w < kMetBlock / 32; ++w) v += s[w][threadIdx.x];
        part[static_cast<int64_t>(blockIdx.x) * 16 + threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(kMetBlock)
variogram_final_kernel(const double* __restrict__ part, int n_part, double* __restrict__ sums16) {
    __shared__ double s[kMetBlock / 16][16];
    const int k = threadIdx.x & 15, lane16 = threadIdx.x >> 4;   // 16 groups of 16 threads: thread (lane16, k) sums entry k
    double v = 0.0;
    for (int b = lane16; b < n_part; b += kMetBlock / 16) v += part[static_cast<int64_t>(b) * 16 + k];
    s[lane16][k] = v;
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0.0;
        for (int w = 0; w < kMetBlock / 16; ++w) t += s[w][threadIdx.x];
        sums16[threadIdx.x] = t;
    }
}

size_t variogram_scratch_bytes() { return sizeof(double) * 16 * kVgBlocks; }

cudaError_t launch_variogram_sums(const GridDesc& d, void* scratch, double* sums16, cudaStream_t st, LaunchInfo* info) {
    double* part = static_cast<double*>(scratch);
    if (d.dtype == DT_F64)
        variogram_partial_kernel<double><<<kVgBlocks, kMetBlock, 0, st>>>(static_cast<const double*>(d.z), d.ld, d.rows, d.n_lon, part);
    else
        variogram_partial_kernel<float><<<kVgBlocks, kMetBlock, 0, st>>>(static_cast<const float*>(d.z), d.ld, d.rows, d.n_lon, part);
    variogram_final_kernel<<<1, kMetBlock, 0, st>>>(part, kVgBlocks, sums16);
    if (info) info->launches += 2;
    return cudaGetLastError();
}

}  // namespace auvi
