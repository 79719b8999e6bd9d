// metrics.cu -- device-side error metrics: the reduction behind meanAbsoluteError /
// rootMeanSquareError / maxAbsoluteError (reference: code/src/error_calculator.cpp:5-45).
//
// Convention kept from the reference: an estimate that is NaN adds nothing to the MAE / RMSE
// numerators but still counts in the denominator n (error_calculator.cpp:12-16, :26-31); the
// maximum ignores NaN because `NaN > m` is false (:40-43).  Pure HBM stream: 2 x sizeof(T) bytes per
// element, one pass.  Two-stage and deterministic (fixed partition, fixed tree), no atomics.
#include "launch.h"

namespace auvi {

constexpr int kMetBlock = 256;
constexpr int kMetBlocks = 148 * 8;

struct Partial { double sum_abs, sum_sq, max_abs, n_nan; };

__device__ __forceinline__ void combine(Partial& a, const Partial& b) {
    a.sum_abs += b.sum_abs; a.sum_sq += b.sum_sq; a.max_abs = fmax(a.max_abs, b.max_abs); a.n_nan += b.n_nan;
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
        combine(v, w);
    }
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        Partial z{0.0, 0.0, 0.0, 0.0};
        v = threadIdx.x < kMetBlock / 32 ? s[threadIdx.x] : z;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            Partial w;
            w.sum_abs = __shfl_down_sync(0xffffffffu, v.sum_abs, o);
            w.sum_sq = __shfl_down_sync(0xffffffffu, v.sum_sq, o);
            w.max_abs = __shfl_down_sync(0xffffffffu, v.max_abs, o);
            w.n_nan = __shfl_down_sync(0xffffffffu, v.n_nan, o);
            combine(v, w);
        }
    }
    return v;
}

template <typename T>
__global__ void __launch_bounds__(kMetBlock)
metrics_partial_kernel(const T* __restrict__ truth, const T* __restrict__ est, int64_t n, Partial* __restrict__ part) {
    Partial acc{0.0, 0.0, 0.0, 0.0};
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * kMetBlock + threadIdx.x; k < n;
         k += static_cast<int64_t>(gridDim.x) * kMetBlock) {
        const double t = static_cast<double>(__ldcs(truth + k)), e = static_cast<double>(__ldcs(est + k));
        const double d = fabs(t - e);
        if (isnan(e)) acc.n_nan += 1.0;
        else { acc.sum_abs += d; acc.sum_sq += d * d; }
        if (d > acc.max_abs) acc.max_abs = d;                  // false for NaN, like the reference
    }
    acc = block_reduce(acc);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(kMetBlock)
metrics_final_kernel(const Partial* __restrict__ part, int n_part, double* __restrict__ result4) {
    Partial acc{0.0, 0.0, 0.0, 0.0};
    for (int k = threadIdx.x; k < n_part; k += kMetBlock) combine(acc, part[k]);
    acc = block_reduce(acc);
    if (threadIdx.x == 0) {
        result4[0] = acc.sum_abs; result4[1] = acc.sum_sq; result4[2] = acc.max_abs; result4[3] = acc.n_nan;
    }
}

size_t metrics_scratch_bytes() { return sizeof(Partial) * kMetBlocks; }

cudaError_t launch_metrics(const void* truth, const void* est, int dtype, int64_t n, void* scratch,
                           double* result4, cudaStream_t st, LaunchInfo* info) {
    int64_t want = (n + kMetBlock - 1) / kMetBlock;
    const int blocks = static_cast<int>(want < 1 ? 1 : (want > kMetBlocks ? kMetBlocks : want));
    Partial* part = static_cast<Partial*>(scratch);
    if (dtype == DT_F64)
        metrics_partial_kernel<double><<<blocks, kMetBlock, 0, st>>>(static_cast<const double*>(truth),
                                                                     static_cast<const double*>(est), n, part);
    else
        metrics_partial_kernel<float><<<blocks, kMetBlock, 0, st>>>(static_cast<const float*>(truth),
                                                                    static_cast<const float*>(est), n, part);
    metrics_final_kernel<<<1, kMetBlock, 0, st>>>(part, blocks, result4);
    if (info) info->launches += 2;
    return cudaGetLastError();
}

}  // namespace auvi
