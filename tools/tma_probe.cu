// probe: 2-D TMA tile loads with negative / odd start coordinates, f64 and f32
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../auv-real-time-interpolation_b200/csrc/tma.cuh"
using namespace auvi;
template <typename T>
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x, int y, int bw, int bh, T* out) {
    extern __shared__ __align__(128) unsigned char raw[];
    T* tile = reinterpret_cast<T*>(raw);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, bw * bh * sizeof(T)); tma_load_2d(tile, &tmap, x, y, &bar); }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <typename T> int run(int W, int H, int bw, int bh, int x, int y, CUtensorMapDataType dt) {
    void* sym; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    Enc enc = (Enc)sym;
    size_t pitch = (W * sizeof(T) + 15) / 16 * 16;
    std::vector<T> h(pitch / sizeof(T) * H);
    for (int j = 0; j < H; ++j) for (int i = 0; i < W; ++i) h[j * pitch / sizeof(T) + i] = 100 * j + i + 1;
    T* d; cudaMalloc(&d, pitch * H); cudaMemcpy(d, h.data(), pitch * H, cudaMemcpyHostToDevice);
    T* out; cudaMalloc(&out, bw * bh * sizeof(T));
    CUtensorMap m; cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t str[1] = {pitch};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, dt, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d ", (int)r);
    k<T><<<1, 128, bw * bh * sizeof(T)>>>(m, x, y, bw, bh, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("W=%d H=%d box=%dx%d at (%d,%d) es=%zu -> %s", W, H, bw, bh, x, y, sizeof(T), cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<T> o(bw * bh); cudaMemcpy(o.data(), out, bw * bh * sizeof(T), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int j = 0; j < bh; ++j) for (int i = 0; i < bw; ++i) {
            int gx = x + i, gy = y + j; T want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? (T)(100 * gy + gx + 1) : (T)0;
            if (o[j * bw + i] != want) ++bad;
        }
        printf(" mismatches=%d", bad);
    }
    printf("\n");
    return 0;
}
int main(int argc, char** argv) {
    int es = atoi(argv[1]), W = atoi(argv[2]), H = atoi(argv[3]), bw = atoi(argv[4]), bh = atoi(argv[5]), x = atoi(argv[6]), y = atoi(argv[7]);
    if (es == 8) return run<double>(W, H, bw, bh, x, y, CU_TENSOR_MAP_DATA_TYPE_FLOAT64);
    return run<float>(W, H, bw, bh, x, y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}
