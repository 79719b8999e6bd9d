// fill.cu -- full-grid gap fill: every NaN cell of the grid gets METHOD(query at its own node), valid
// cells pass through.  This is the structured form of the reference's Grid-B procedure
// (test_gebco.cpp:150-196: one query per removed cell, built by gridIndexToGeo :72-81, evaluated by
// GridH::batch{Cubic,OrdinaryKriging}Interpolate, GridH.cpp:223-420) and of BASELINE config 4
// (IDW / nearest-neighbour on a 70 % masked grid).  Methods: CUBIC (always the ring-search
// 4-nearest mean here: a masked cell is inside its own 4x4 stencil), KRIGING, NN, IDW.
//
// Design (one CTA = 256 threads = one 64 x 32 tile of cells):
//   1. The tile + a 12-cell halo (88 x 56 cells: search radius 10 + a centre that FP64 noise may move
//      by one cell) is staged into shared memory by ONE TMA 2-D box load.
//   2. Warps turn it into validity bitmasks with __ballot_sync (one 96-bit row per tile row); from here
//      on the ring search of GridH.cpp:24-118 is bit arithmetic: a 21-bit window per row, popcounts
//      for the early-termination rule, find-first-set for the enumeration order.
//   3. Valid cells are copied to the output tile; masked cells are COMPACTED into a CTA-wide queue
//      (warp-aggregated shared-memory atomics), so every warp of the search phase is full of real
//      queries regardless of the mask pattern.
//   4. One thread per query: termination ring from the bitmasks, candidates in the reference's
//      enumeration order with their FP64 squared distances (same operation order as GridH.cpp:42-44),
//      the reference's partial selection sort WITH swaps on a per-thread list in shared memory.
//      Distances are compared squared; sqrt is monotone, so the order can only differ when two squared
//      distances are within a few ulps -- those queries (and lists longer than 12) are redone by the
//      literal per-query path (exact.cuh).  Values of the four picks come from the shared tile.
//   5. Method epilogue (mean / NN / FP32 IDW weights / FP64 kriging solve), then the whole output
//      tile is written with coalesced 16-byte stores.
//
// Algorithmic HBM bytes: sizeof(T) read + sizeof(T) written per cell; the kernel is
// issue-bound (integer/bit work + FP64 distance math), not HBM-bound.
#include <cstdlib>
#include <cstring>

#include "exact.cuh"
#include "launch.h"
#include "tma.cuh"

namespace auvi {

constexpr int kFW = 64, kFH = 32;              // tile of output cells
constexpr int kFHalo = 12;
constexpr int kFBW = kFW + 2 * kFHalo;         // 88
constexpr int kFBH = kFH + 2 * kFHalo;         // 56
constexpr int kFThreads = 256;
constexpr int kFNMax = 12;                     // per-thread candidate list capacity (longer lists: literal path)
constexpr int kFRedoMax = 128;

struct FillAxis {
    const double* coord;
    const double* pos;
    const int* base;
};

template <typename T>
struct FillParams {
    GridView<T> g;
    FillAxis lat, lon;
    int64_t row_begin, row_end;
    int rows_resident_lo, rows_resident_hi;   // global rows [lo,hi) present in g.z
    T* out;
    int64_t out_ld;
    int use_tma;
    int vec_ok;
};

template <typename T>
struct FillSmem {
    alignas(128) T tile[kFBH * kFBW];
    alignas(16) T out[kFH * kFW];
    double d2[kFNMax * kFThreads];
    double x[kFW], y[kFH];
    uint32_t mask[kFBH * 4];
    int cx[kFW], cy[kFH];
    uint16_t code[kFNMax * kFThreads];
    uint16_t queue[kFW * kFH];
    uint16_t redo[kFRedoMax];
    int qn, rn;
    uint64_t bar;
};

template <typename T, int METHOD>
__device__ __noinline__ T fill_cell_literal(const FillParams<T>* p, int64_t J, int I) {
    return static_cast<T>(interp_exact<T>(p->g, METHOD, __ldg(p->lon.coord + I), __ldg(p->lat.coord + J),
                                          __ldg(p->lon.pos + I), __ldg(p->lat.pos + J), nullptr));
}

template <typename T, int METHOD>
__global__ void __launch_bounds__(kFThreads, 3)
fill_tiled_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ FillParams<T> p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FillSmem<T>& s = *reinterpret_cast<FillSmem<T>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = p.g.n_lon;
    const int I0 = blockIdx.x * kFW;
    const int64_t J0 = p.row_begin + static_cast<int64_t>(blockIdx.y) * kFH;
    const int c0 = I0 - kFHalo;                                   // 16-byte aligned for f32 and f64
    const int r0 = static_cast<int>(J0) - kFHalo;

    if (tid == 0) { s.qn = 0; s.rn = 0; }
    if (p.use_tma) {
        if (tid == 0) { prefetch_tmap(&tmap); mbar_init(&s.bar, 1); }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&s.bar, static_cast<uint32_t>(kFBW * kFBH * sizeof(T)));
            tma_load_2d(s.tile, &tmap, c0, r0 - p.g.row0, &s.bar);
        }
    } else {
        for (int k = tid; k < kFBW * kFBH; k += kFThreads) {
            const int lr = k / kFBW, lc = k - lr * kFBW;
            const int gr = r0 + lr, gc = c0 + lc;
            T v = static_cast<T>(qnan());
            if (gr >= p.rows_resident_lo && gr < p.rows_resident_hi && gc >= 0 && gc < W)
                v = __ldg(p.g.z + static_cast<int64_t>(gr - p.g.row0) * p.g.ld + gc);
            s.tile[k] = v;
        }
    }
    // per-axis query tables of this tile (overlaps the TMA flight): index-space position and search centre
    if (tid < kFW) {
        const int I = I0 + tid;
        double x = qnan();
        int c = 0;
        if (I < W) {
            x = __ldg(p.lon.pos + I);
            c = METHOD == CUBIC ? __ldg(p.lon.base + I) : (isnan(x) ? 0 : round_centre(x, p.g.n_lon));
        }
        s.x[tid] = x; s.cx[tid] = c;
    } else if (tid < kFW + kFH) {
        const int t = tid - kFW;
        const int64_t J = J0 + t;
        double y = qnan();
        int c = 0;
        if (J < p.row_end) {
            y = __ldg(p.lat.pos + J);
            c = METHOD == CUBIC ? __ldg(p.lat.base + J) : (isnan(y) ? 0 : round_centre(y, p.g.n_lat));
        }
        s.y[t] = y; s.cy[t] = c;
    }
    if (p.use_tma) mbar_wait(&s.bar, 0);
    else __syncthreads();

    // ---- validity bitmasks: bit c of row r = tile cell (r,c) holds a number and lies inside the grid ----
    for (int r = warp; r < kFBH; r += kFThreads / 32) {
        const int gr = r0 + r;
        const bool row_ok = gr >= p.rows_resident_lo && gr < p.rows_resident_hi;
#pragma unroll
        for (int seg = 0; seg < 3; ++seg) {
            const int c = seg * 32 + lane, gc = c0 + c;
            bool ok = false;
            if (row_ok && c < kFBW && gc >= 0 && gc < W) ok = !isnan(s.tile[r * kFBW + c]);
            const uint32_t word = __ballot_sync(0xffffffffu, ok);
            if (lane == 0) s.mask[r * 4 + seg] = word;
        }
        if (lane == 0) s.mask[r * 4 + 3] = 0u;
    }
    __syncthreads();

    // ---- pass valid cells through, compact masked cells into the queue ----------------------------------
#pragma unroll
    for (int it = 0; it < kFW * kFH / kFThreads; ++it) {
        const int k = it * kFThreads + tid;
        const int lj = k / kFW, li = k % kFW;
        const bool in_range = (J0 + lj < p.row_end) && (I0 + li < W);
        const int tr = lj + kFHalo, tc = li + kFHalo;
        const bool valid = (s.mask[tr * 4 + (tc >> 5)] >> (tc & 31)) & 1u;
        const bool todo = in_range && !valid;
        if (in_range && valid) s.out[k] = s.tile[tr * kFBW + tc];
        const uint32_t m = __ballot_sync(0xffffffffu, todo);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(&s.qn, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (todo) s.queue[base + __popc(m & ((1u << lane) - 1u))] = static_cast<uint16_t>(k);
    }
    __syncthreads();

    // ---- one thread per masked cell --------------------------------------------------------------------
    const int qn = s.qn;
    for (int q = tid; q < qn; q += kFThreads) {
        const int k = s.queue[q];
        const int lj = k / kFW, li = k % kFW;
        const double x = s.x[li], y = s.y[lj];
        if (isnan(x) || isnan(y)) { s.out[k] = static_cast<T>(qnan()); continue; }   // query out of bounds
        const int cig = s.cx[li], cjg = s.cy[lj];
        const int ci = cig - c0, cj = cjg - r0;                     // search centre in tile coordinates
        bool literal = (ci < 11) | (ci > kFBW - 12) | (cj < 11) | (cj > kFBH - 12);   // never for node queries
        double result = 0.0;
        if (!literal) {
            const int sh0 = ci - 10, wi = sh0 >> 5, sh = sh0 & 31;
            auto win = [&](int row) -> uint32_t {                  // validity of columns ci-10..ci+10 of a tile row
                return __funnelshift_r(s.mask[row * 4 + wi], s.mask[row * 4 + wi + 1], sh) & 0x1FFFFFu;
            };
            // -- where does the reference's search stop?  (GridH.cpp:48-117: break checks after each pass)
            uint32_t wt[kMaxRadius + 1], wb[kMaxRadius + 1];
            const uint32_t w0 = win(cj);
            wt[0] = w0; wb[0] = w0;
            int n = (w0 >> 10) & 1;
            int r_end = kMaxRadius, lr_end = 1;                     // last ring visited; did its left/right pass run?
            bool done = false;
#pragma unroll
            for (int r = 1; r <= kMaxRadius; ++r) {
                if (!done) {
                    wt[r] = win(cj - r); wb[r] = win(cj + r);
                    const uint32_t tbm = ((2u << (2 * r)) - 1u) << (10 - r);
                    n += __popc(wt[r] & tbm) + __popc(wb[r] & tbm);
                    if (n >= 4) { done = true; r_end = r; lr_end = 0; }
                    else {
                        const uint32_t lrm = (1u << (10 - r)) | (1u << (10 + r));
                        int c = __popc(w0 & lrm);
#pragma unroll
                        for (int d = 1; d < r; ++d) c += __popc(wt[d] & lrm) + __popc(wb[d] & lrm);
                        n += c;
                        if (n >= 4) { done = true; r_end = r; lr_end = 1; }
                    }
                }
            }
            if (n > kFNMax) literal = true;
            if (!literal) {
                // -- candidates in enumeration order, squared distances in the reference's operation order
                const double cxf = dadd(__int2double_rn(cig), 0.5), cyf = dadd(__int2double_rn(cjg), 0.5);
                double* const ld2 = s.d2 + tid;
                uint16_t* const lcode = s.code + tid;
                int cnt = 0;
                auto push = [&](int dx, int dy, double d2) {
                    ld2[cnt * kFThreads] = d2;
                    lcode[cnt * kFThreads] = static_cast<uint16_t>((dy + 10) * 32 + (dx + 10));
                    ++cnt;
                };
                auto sq_off = [&](double cf, int d, double q) -> double {   // ((c + d + 0.5) - q)^2, GridH.cpp:42-43
                    const double t = dsub(dadd(cf, __int2double_rn(d)), q);
                    return dmul(t, t);
                };
                if ((w0 >> 10) & 1) push(0, 0, dadd(sq_off(cxf, 0, x), sq_off(cyf, 0, y)));
#pragma unroll
                for (int r = 1; r <= kMaxRadius; ++r) {
                    if (r <= r_end) {
                        const uint32_t tbm = ((2u << (2 * r)) - 1u) << (10 - r);
                        const uint32_t top = wt[r] & tbm, bot = wb[r] & tbm;
                        uint32_t m = top | bot;
                        if (m) {
                            const double dyt = sq_off(cyf, -r, y), dyb = sq_off(cyf, r, y);
                            while (m) {                           // columns left to right; top before bottom
                                const int b = __ffs(m) - 1;
                                m &= m - 1;
                                const double dx2 = sq_off(cxf, b - 10, x);
                                if ((top >> b) & 1u) push(b - 10, -r, dadd(dx2, dyt));
                                if ((bot >> b) & 1u) push(b - 10, r, dadd(dx2, dyb));
                            }
                        }
                        if (r < r_end || lr_end) {
                            const uint32_t lb = 1u << (10 - r), rb = 1u << (10 + r);
                            uint32_t any = (w0 & (lb | rb));
#pragma unroll
                            for (int d = 1; d < r; ++d) any |= (wt[d] | wb[d]) & (lb | rb);
                            if (any) {
                                const double dxl = sq_off(cxf, -r, x), dxr = sq_off(cxf, r, x);
#pragma unroll
                                for (int dy = -r + 1; dy <= r - 1; ++dy) {   // rows top to bottom; left before right
                                    const uint32_t wr = dy < 0 ? wt[-dy] : (dy == 0 ? w0 : wb[dy]);
                                    if (wr & (lb | rb)) {
                                        const double dy2 = sq_off(cyf, dy, y);
                                        if (wr & lb) push(-r, dy, dadd(dxl, dy2));
                                        if (wr & rb) push(r, dy, dadd(dxr, dy2));
                                    }
                                }
                            }
                        }
                    }
                }
                // cnt == n by construction
                // -- the reference's partial selection sort with swaps (GridH.cpp:123-140) on squared distances
                constexpr double kSafe = 1.0 - 8.8817841970012523e-16;      // 1 - 2^-50
                bool unsure = false;
                // NN only needs the first pick: pass 0 (with fewer than four candidates the oracle's "first strict
                // minimum" is the same scan); the other methods run the four passes when four candidates exist.
                const int n_pass = METHOD == NN ? (cnt > 0 ? 1 : 0) : (cnt >= 4 ? 4 : 0);
                for (int m = 0; m < n_pass; ++m) {
                    int best = m;
                    const double dm = ld2[m * kFThreads];
                    double dbest = dm, thr = dmul(dm, kSafe);
                    for (int kk = m + 1; kk < cnt; ++kk) {
                        const double dk = ld2[kk * kFThreads];
                        if (dk < thr) { best = kk; dbest = dk; thr = dmul(dk, kSafe); }
                        else if (dk < dbest) unsure = true;        // within a few ulps: sqrt may tie
                    }
                    if (best != m) {
                        ld2[m * kFThreads] = dbest; ld2[best * kFThreads] = dm;
                        const uint16_t cm = lcode[m * kFThreads];
                        lcode[m * kFThreads] = lcode[best * kFThreads]; lcode[best * kFThreads] = cm;
                    }
                }
                if (unsure) literal = true;
                if (!literal) {
                    // -- gather the picks (first min(cnt,4) list entries) and finish the method
                    const int np = cnt < 4 ? cnt : 4;
                    Picked pk;
                    pk.found = cnt;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (e < np) {
                            const int code = lcode[e * kFThreads];
                            const int dx = (code & 31) - 10, dy = (code >> 5) - 10;
                            pk.i[e] = cig + dx; pk.j[e] = cjg + dy;
                            pk.v[e] = static_cast<double>(s.tile[(cj + dy) * kFBW + ci + dx]);
                            pk.d[e] = ld2[e * kFThreads];            // SQUARED distance
                        } else { pk.i[e] = -1; pk.j[e] = -1; pk.v[e] = qnan(); pk.d[e] = qnan(); }
                    }
                    if (METHOD == CUBIC) {
                        result = cnt < 4 ? mean_found(pk) : mean_valid4(pk.v[0], pk.v[1], pk.v[2], pk.v[3]);
                    } else if (METHOD == NN) {
                        result = cnt > 0 ? pk.v[0] : qnan();
                    } else if (METHOD == IDW) {
                        if (cnt == 0) result = qnan();
                        else {
                            // FP32 weights 1/d^2 through the SFU reciprocal; values centred on the first pick
                            float num = 0.f, den = 0.f;
                            const double ref = pk.v[0];
                            bool hit = false;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                if (e < np && !hit) {
                                    if (pk.d[e] == 0.0) { result = pk.v[e]; hit = true; }
                                    else {
                                        const float w = __frcp_rn(static_cast<float>(pk.d[e]));
                                        num = fmaf(w, static_cast<float>(pk.v[e] - ref), num);
                                        den += w;
                                    }
                                }
                            }
                            if (!hit) result = ref + static_cast<double>(__fdividef(num, den));
                        }
                    } else {   // KRIGING
                        if (cnt < 4) result = mean_found(pk);
                        else result = kriging_from_picked(p.g, pk, __ldg(p.lon.coord + I0 + li), __ldg(p.lat.coord + J0 + lj));
                    }
                }
            }
        }
        if (literal) {
            const int slot = atomicAdd(&s.rn, 1);
            if (slot < kFRedoMax) s.redo[slot] = static_cast<uint16_t>(k);
            else s.out[k] = fill_cell_literal<T, METHOD>(&p, J0 + lj, I0 + li);
        } else {
            s.out[k] = static_cast<T>(result);
        }
    }
    __syncthreads();
    // ---- queries the bitmask path handed back: literal per-query evaluation ---------------------------------
    const int rn = min(s.rn, kFRedoMax);
    for (int q = tid; q < rn; q += kFThreads) {
        const int k = s.redo[q];
        s.out[k] = fill_cell_literal<T, METHOD>(&p, J0 + k / kFW, I0 + k % kFW);
    }
    if (rn) __syncthreads();

    // ---- write the tile: coalesced 16-byte stores ----------------------------------------------------------
    constexpr int VEC = 16 / static_cast<int>(sizeof(T));
    const int rows_here = static_cast<int>(min(static_cast<int64_t>(kFH), p.row_end - J0));
    T* const out_tile = p.out + (J0 - p.row_begin) * p.out_ld + I0;
    if (p.vec_ok && I0 + kFW <= W) {
        for (int k = tid; k < rows_here * (kFW / VEC); k += kFThreads) {
            const int lj = k / (kFW / VEC), lv = k % (kFW / VEC);
            const int4 v = *reinterpret_cast<const int4*>(&s.out[lj * kFW + lv * VEC]);
            __stcs(reinterpret_cast<int4*>(out_tile + lj * p.out_ld + lv * VEC), v);
        }
    } else {
        for (int k = tid; k < rows_here * kFW; k += kFThreads) {
            const int lj = k / kFW, li = k % kFW;
            if (I0 + li < W) __stcs(out_tile + lj * p.out_ld + li, s.out[k]);
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
template <typename T, int METHOD>
static cudaError_t launch_fill_t(const GridDesc& d, const AxisTables& lat, const AxisTables& lon, int64_t row_begin,
                                 int64_t row_end, void* out, int64_t out_ld, cudaStream_t st, LaunchInfo* info) {
    // rows within the ring-search reach of the block must be resident (slab + halo)
    int need_lo = static_cast<int>(row_begin) - (kMaxRadius + 2), need_hi = static_cast<int>(row_end) - 1 + (kMaxRadius + 2);
    need_lo = need_lo < 0 ? 0 : need_lo;
    need_hi = need_hi > d.n_lat - 1 ? d.n_lat - 1 : need_hi;
    if (need_lo < d.row0 || need_hi >= d.row0 + d.rows) return cudaErrorInvalidValue;

    FillParams<T> p;
    p.g = make_view<T>(d);
    p.lat = FillAxis{lat.coord, lat.pos, lat.base};
    p.lon = FillAxis{lon.coord, lon.pos, lon.base};
    p.row_begin = row_begin; p.row_end = row_end;
    p.rows_resident_lo = d.row0; p.rows_resident_hi = d.row0 + d.rows;
    p.out = static_cast<T*>(out); p.out_ld = out_ld;
    const size_t es = sizeof(T);
    p.vec_ok = (reinterpret_cast<uintptr_t>(out) % 16 == 0 && (out_ld * es) % 16 == 0) ? 1 : 0;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    p.use_tma = make_grid_tensor_map(d, kFBW, kFBH, &tmap) ? 1 : 0;
    auto kern = fill_tiled_kernel<T, METHOD>;
    const size_t smem = sizeof(FillSmem<T>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    dim3 grid(static_cast<unsigned>((d.n_lon + kFW - 1) / kFW), static_cast<unsigned>((row_end - row_begin + kFH - 1) / kFH));
    kern<<<grid, kFThreads, smem, st>>>(tmap, p);
    if (info) { info->launches += 1; info->used_tma = p.use_tma; }
    return cudaGetLastError();
}

cudaError_t launch_fill(const GridDesc& d, int method, const AxisTables& lat, const AxisTables& lon, int64_t row_begin,
                        int64_t row_end, void* out, int64_t out_ld, cudaStream_t st, LaunchInfo* info) {
    if (row_end <= row_begin) return cudaSuccess;
#define AUVI_CASE(T, M) \
    case M: return launch_fill_t<T, M>(d, lat, lon, row_begin, row_end, out, out_ld, st, info);
    if (d.dtype == DT_F64) {
        switch (method) { AUVI_CASE(double, CUBIC) AUVI_CASE(double, KRIGING) AUVI_CASE(double, NN) AUVI_CASE(double, IDW) }
    } else {
        switch (method) { AUVI_CASE(float, CUBIC) AUVI_CASE(float, KRIGING) AUVI_CASE(float, NN) AUVI_CASE(float, IDW) }
    }
#undef AUVI_CASE
    return cudaErrorInvalidValue;
}

}  // namespace auvi
