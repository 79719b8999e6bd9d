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
constexpr int kFSq = 3;                        // squared-offset tables cover |offset| <= kFSq
constexpr int kFBins = 80;                     // (termination code 0..6) x (candidate count class 0..9) + spare

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
};

template <typename T>
struct FillSmem {
    alignas(128) T tile[kFBH * kFBW];
    double d2[kFNMax * kFThreads];            // per-thread candidate lists: squared distances
    double sqx[kFW * (2 * kFSq + 1)];         // ((cx + d + 0.5) - x)^2 per tile column, d = -kFSq..kFSq
    double sqy[kFH * (2 * kFSq + 1)];
    double x[kFW], y[kFH];
    uint32_t sorted[kFW * kFH];               // queries ordered by (termination, candidate count)
    uint32_t mask[kFBH * 4];
    int cx[kFW], cy[kFH];
    int hist[kFBins];
    uint16_t code[kFNMax * kFThreads];        // per-thread candidate lists: packed (dy,dx) offsets
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

// ((c + d + 0.5) - q)^2 in the reference's operation order (GridH.cpp:42-44); cf = c + 0.5 is exact.
__device__ __forceinline__ double sq_offset(double cf, int d, double q) {
    const double t = dsub(dadd(cf, __int2double_rn(d)), q);
    return dmul(t, t);
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
    T* const out_tile = p.out + (J0 - p.row_begin) * p.out_ld + I0;

    if (tid == 0) { s.qn = 0; s.rn = 0; }
    if (tid < kFBins) s.hist[tid] = 0;
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
    // per-axis query tables of this tile (overlaps the TMA flight): index-space position, search centre and
    // the squared offsets of the near rings
    if (tid < kFW) {
        const int I = I0 + tid;
        double x = qnan();
        int c = 0;
        if (I < W) {
            x = __ldg(p.lon.pos + I);
            c = METHOD == CUBIC ? __ldg(p.lon.base + I) : (isnan(x) ? 0 : round_centre(x, p.g.n_lon));
        }
        s.x[tid] = x; s.cx[tid] = c;
        const double cf = dadd(__int2double_rn(c), 0.5);
#pragma unroll
        for (int d = -kFSq; d <= kFSq; ++d) s.sqx[tid * (2 * kFSq + 1) + d + kFSq] = sq_offset(cf, d, x);
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
        const double cf = dadd(__int2double_rn(c), 0.5);
#pragma unroll
        for (int d = -kFSq; d <= kFSq; ++d) s.sqy[t * (2 * kFSq + 1) + d + kFSq] = sq_offset(cf, d, y);
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

    // ---- pass valid cells through (coalesced), compact masked cells into the queue -------------------------
#pragma unroll
    for (int it = 0; it < kFW * kFH / kFThreads; ++it) {
        const int k = it * kFThreads + tid;
        const int lj = k / kFW, li = k % kFW;
        const bool in_range = (J0 + lj < p.row_end) && (I0 + li < W);
        const int tr = lj + kFHalo, tc = li + kFHalo;
        const bool valid = (s.mask[tr * 4 + (tc >> 5)] >> (tc & 31)) & 1u;
        const bool todo = in_range && !valid;
        if (in_range && valid) __stcs(out_tile + lj * p.out_ld + li, s.tile[tr * kFBW + tc]);
        const uint32_t m = __ballot_sync(0xffffffffu, todo);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(&s.qn, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (todo) s.queue[base + __popc(m & ((1u << lane) - 1u))] = static_cast<uint16_t>(k);
    }
    __syncthreads();

    // ---- phase A: where does each query's search stop, and with how many candidates? ------------------------
    // (GridH.cpp:48-117: the count is checked after each top/bottom pass and after each left/right pass.)
    // Queries are then ORDERED by (termination ring/pass, candidate count) so that the warps of phase B run
    // the same ring code and the same list lengths: the data-dependent loops stop diverging.
    const int qn = s.qn;
    auto window = [&](int row, int wi, int sh) -> uint32_t {       // validity of columns ci-10..ci+10 of a tile row
        return __funnelshift_r(s.mask[row * 4 + wi], s.mask[row * 4 + wi + 1], sh) & 0x1FFFFFu;
    };
    for (int q = tid; q < qn; q += kFThreads) {
        const int k = s.queue[q];
        const int lj = k / kFW, li = k % kFW;
        uint32_t rec = 0xffffffffu;                                // "not for phase B"
        if (isnan(s.x[li]) || isnan(s.y[lj])) {
            __stcs(out_tile + lj * p.out_ld + li, static_cast<T>(qnan()));   // query out of bounds
        } else {
            const int ci = s.cx[li] - c0, cj = s.cy[lj] - r0;       // search centre in tile coordinates
            bool literal = (ci < 11) | (ci > kFBW - 12) | (cj < 11) | (cj > kFBH - 12);   // never for node queries
            if (!literal) {
                const int sh0 = ci - 10, wi = sh0 >> 5, sh = sh0 & 31;
                uint32_t wt[kMaxRadius + 1], wb[kMaxRadius + 1];
                const uint32_t w0 = window(cj, wi, sh);
                wt[0] = w0; wb[0] = w0;
                int n = (w0 >> 10) & 1;
                int r_end = kMaxRadius, lr_end = 1;                 // last ring visited; did its left/right pass run?
                bool done = false;
#pragma unroll
                for (int r = 1; r <= kMaxRadius; ++r) {
                    if (!done) {
                        wt[r] = window(cj - r, wi, sh); wb[r] = window(cj + r, wi, sh);
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
                else {
                    const int tcode = r_end <= 3 ? (r_end - 1) * 2 + lr_end : 6;
                    const int ncls = n < 4 ? 9 : n - 4;             // n in 4..12 -> 0..8
                    const int bin = tcode * 10 + ncls;
                    atomicAdd(&s.hist[bin], 1);
                    rec = static_cast<uint32_t>(k) | (static_cast<uint32_t>(r_end) << 12) |
                          (static_cast<uint32_t>(lr_end) << 16) | (static_cast<uint32_t>(n) << 17) |
                          (static_cast<uint32_t>(bin) << 22);
                }
            }
            if (literal) {
                const int slot = atomicAdd(&s.rn, 1);
                if (slot < kFRedoMax) s.redo[slot] = static_cast<uint16_t>(k);
                else __stcs(out_tile + lj * p.out_ld + li, fill_cell_literal<T, METHOD>(&p, J0 + lj, I0 + li));
            }
        }
        s.sorted[q] = rec;                                          // parked here until the scatter below
    }
    __syncthreads();
    if (warp == 0) {                                               // exclusive scan of the histogram
        int carry = 0;
        for (int b0 = 0; b0 < kFBins; b0 += 32) {
            const int b = b0 + lane;
            int v = b < kFBins ? s.hist[b] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (b < kFBins) s.hist[b] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) s.qn = carry;                               // queries that go to phase B
    }
    __syncthreads();
    // scatter: records move from the parking order to the bin order (read all, sync, then write)
    uint32_t mine[kFW * kFH / kFThreads];
#pragma unroll
    for (int it = 0; it < kFW * kFH / kFThreads; ++it) {
        const int q = it * kFThreads + tid;
        mine[it] = q < qn ? s.sorted[q] : 0xffffffffu;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kFW * kFH / kFThreads; ++it) {
        if (mine[it] != 0xffffffffu) s.sorted[atomicAdd(&s.hist[mine[it] >> 22], 1)] = mine[it];
    }
    __syncthreads();

    // ---- phase B: one thread per query, bin order ---------------------------------------------------------------
    const int qb = s.qn;
    for (int q = tid; q < qb; q += kFThreads) {
        const uint32_t rec = s.sorted[q];
        const int k = rec & 0xfff, r_end = (rec >> 12) & 15, lr_end = (rec >> 16) & 1, n = (rec >> 17) & 31;
        const int lj = k / kFW, li = k % kFW;
        const int cig = s.cx[li], cjg = s.cy[lj];
        const int ci = cig - c0, cj = cjg - r0;
        const int sh0 = ci - 10, wi = sh0 >> 5, sh = sh0 & 31;
        const double x = s.x[li], y = s.y[lj];
        const double cxf = dadd(__int2double_rn(cig), 0.5), cyf = dadd(__int2double_rn(cjg), 0.5);
        const double* const sqx = s.sqx + li * (2 * kFSq + 1) + kFSq;
        const double* const sqy = s.sqy + lj * (2 * kFSq + 1) + kFSq;
        auto sqdx = [&](int d) -> double { return (d >= -kFSq && d <= kFSq) ? sqx[d] : sq_offset(cxf, d, x); };
        auto sqdy = [&](int d) -> double { return (d >= -kFSq && d <= kFSq) ? sqy[d] : sq_offset(cyf, d, y); };

        // -- candidates in the reference's enumeration order with their squared distances
        double* const ld2 = s.d2 + tid;
        uint16_t* const lcode = s.code + tid;
        int cnt = 0;
        auto push = [&](int dx, int dy, double d2) {
            ld2[cnt * kFThreads] = d2;
            lcode[cnt * kFThreads] = static_cast<uint16_t>((dy + 10) * 32 + (dx + 10));
            ++cnt;
        };
        uint32_t wt[kMaxRadius + 1], wb[kMaxRadius + 1];
        const uint32_t w0 = window(cj, wi, sh);
        wt[0] = w0; wb[0] = w0;
        if ((w0 >> 10) & 1) push(0, 0, dadd(sqdx(0), sqdy(0)));
#pragma unroll
        for (int r = 1; r <= kMaxRadius; ++r) {
            if (r <= r_end) {
                wt[r] = window(cj - r, wi, sh); wb[r] = window(cj + r, wi, sh);
                const uint32_t tbm = ((2u << (2 * r)) - 1u) << (10 - r);
                const uint32_t top = wt[r] & tbm, bot = wb[r] & tbm;
                uint32_t m = top | bot;
                if (m) {
                    const double dyt = sqdy(-r), dyb = sqdy(r);
                    while (m) {                                   // columns left to right; top before bottom
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
                        const double dx2 = sqdx(b - 10);
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
                        const double dxl = sqdx(-r), dxr = sqdx(r);
#pragma unroll
                        for (int dy = -r + 1; dy <= r - 1; ++dy) {   // rows top to bottom; left before right
                            const uint32_t wr = dy < 0 ? wt[-dy] : (dy == 0 ? w0 : wb[dy]);
                            if (wr & (lb | rb)) {
                                const double dy2 = sqdy(dy);
                                if (wr & lb) push(-r, dy, dadd(dxl, dy2));
                                if (wr & rb) push(r, dy, dadd(dxr, dy2));
                            }
                        }
                    }
                }
            }
        }
        // cnt == n by construction (phase A counted the same bits)
        // -- the reference's partial selection sort with swaps (GridH.cpp:123-140) on squared distances.
        // NN only needs the first pick: pass 0 (with fewer than four candidates the oracle's "first strict
        // minimum" is the same scan); the other methods run the four passes when four candidates exist.
        constexpr double kSafe = 1.0 - 8.8817841970012523e-16;      // 1 - 2^-50
        bool unsure = (cnt != n);
        const int n_pass = METHOD == NN ? (cnt > 0 ? 1 : 0) : (cnt >= 4 ? 4 : 0);
        for (int m = 0; m < n_pass; ++m) {
            int best = m;
            const double dm = ld2[m * kFThreads];
            double dbest = dm, thr = dmul(dm, kSafe);
            for (int kk = m + 1; kk < cnt; ++kk) {
                const double dk = ld2[kk * kFThreads];
                if (dk < thr) { best = kk; dbest = dk; thr = dmul(dk, kSafe); }
                else if (dk < dbest) unsure = true;                // within a few ulps: sqrt may tie
            }
            if (best != m) {
                ld2[m * kFThreads] = dbest; ld2[best * kFThreads] = dm;
                const uint16_t cm = lcode[m * kFThreads];
                lcode[m * kFThreads] = lcode[best * kFThreads]; lcode[best * kFThreads] = cm;
            }
        }
        if (unsure) {
            const int slot = atomicAdd(&s.rn, 1);
            if (slot < kFRedoMax) s.redo[slot] = static_cast<uint16_t>(k);
            else __stcs(out_tile + lj * p.out_ld + li, fill_cell_literal<T, METHOD>(&p, J0 + lj, I0 + li));
            continue;
        }
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
                pk.d[e] = ld2[e * kFThreads];                        // SQUARED distance
            } else { pk.i[e] = -1; pk.j[e] = -1; pk.v[e] = qnan(); pk.d[e] = qnan(); }
        }
        double result;
        if (METHOD == CUBIC) {
            result = cnt < 4 ? mean_found(pk) : mean_valid4(pk.v[0], pk.v[1], pk.v[2], pk.v[3]);
        } else if (METHOD == NN) {
            result = cnt > 0 ? pk.v[0] : qnan();
        } else if (METHOD == IDW) {
            result = qnan();
            if (cnt > 0) {
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
        __stcs(out_tile + lj * p.out_ld + li, static_cast<T>(result));
    }
    __syncthreads();
    // ---- queries the bitmask path handed back: literal per-query evaluation ---------------------------------
    const int rn = min(s.rn, kFRedoMax);
    for (int q = tid; q < rn; q += kFThreads) {
        const int k = s.redo[q];
        __stcs(out_tile + (k / kFW) * p.out_ld + (k % kFW), fill_cell_literal<T, METHOD>(&p, J0 + k / kFW, I0 + k % kFW));
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
