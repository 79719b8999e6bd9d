// upsample.cu -- lattice mode: every output cell (J,I) of a separable query lattice.
//
// Replaces, for lattice-shaped query sets, the reference's per-point loop
// GridH::batch*Interpolate (GridH.cpp:422-448) / GridD::batch* (GridD.cu:95-236) as driven by
// generateExpandedGridQueryPoints (test_interpolation.cpp:91-109, Grid A upsampling) and by the
// grid-node queries of test_gebco.cpp:72-81,150-160 (Grid B gap fill).  The query coordinates are
// never materialised as 24-byte Points: they live in two per-axis FP64 tables (launch.h AxisTables).
//
// Two kernels:
//
//  upsample_tiled_kernel  (bilinear / bicubic) -- the HBM-bound fast path.
//     CTA = 256 threads = 256 consecutive output columns x TJ output rows.  The input footprint of
//     the tile (+ stencil halo) is staged into shared memory by ONE TMA 2-D box load
//     (cp.async.bulk.tensor, zero-filled outside the grid, then patched to the reference's
//     clamp-to-edge rule) or, when TMA cannot address the grid, by coalesced loads.
//     Each thread owns one output column: its column base/fraction come from the longitude table,
//     it keeps the horizontal pass of four consecutive input rows in registers and slides that
//     window down as the (warp-uniform) latitude table advances, so every output costs one
//     shared-memory read sweep / f_lat + one 4-tap vertical combine + one coalesced store.
//     FP64 grids are evaluated in the reference's exact operation order (bit-identical results);
//     FP32 grids use FP32 tap weights (|err| ~1e-7 relative).  Any output whose stencil touches a
//     NaN (or lies out of bounds) comes out NaN and is re-evaluated by the exact path (exact.cuh),
//     which reproduces the reference's ring-search fallback.
//
//  lattice_exact_kernel   (kriging / NN / IDW, and the gap-fill modes) -- one thread per output,
//     exact path; with FILL it passes valid cells through untouched.
//
// Algorithmic HBM bytes per output cell: s_out + s_in/(f_lat*f_lon).
#include <cstdlib>
#include <cstring>

#include "exact.cuh"
#include "launch.h"
#include "tma.cuh"

namespace auvi {

constexpr int kColThreads = 128;   // threads along the output columns; each owns 16 bytes of every row
constexpr int kRowGroups = 2;      // row halves of a tile, swept by separate thread groups
constexpr int kTileThreads = kColThreads * kRowGroups;
constexpr int kTileRowsMax = 128;  // output rows per CTA (upper bound; host picks TJ <= this)
#ifndef AUVI_F64_MINB
#define AUVI_F64_MINB 3        // resident CTAs per SM the FP64 kernels are compiled for (4 = 64 registers: measured, lost -- DESIGN.md section 10)
#endif

struct AxisDev {
    const double* coord;
    const double* pos;
    const int* base;
    int n;
};

template <typename T>
struct TileParams {
    GridView<T> g;
    AxisDev lat, lon;
    int64_t row_begin, row_end;    // output rows [row_begin,row_end)
    T* out;
    int64_t out_ld;
    int tj;                        // output rows per CTA
    int bw, bh;                    // shared-memory input box (elements), per column slab
    int box_stride;                // elements between the boxes of consecutive slabs (a multiple of 128 bytes: TMA destination)
    int slabs;                     // 1, 2 or 4: the tile's column threads are split into slabs, each with its own input box
                                   // (incl. its own stencil halo), when one box would exceed the 256-element TMA limit
    int use_tma;
    int vec_ok;                    // out and out_ld are 16-byte aligned: vector stores allowed
    int rows_lo, rows_hi;          // global rows [lo,hi) resident in g.z (a row slab of a sharded grid)
};

__device__ __forceinline__ void cr_weights(float t, float& w0, float& w1, float& w2, float& w3) {
    // Catmull-Rom tap weights (the expansion of GridH.cpp:215-217 by tap).
    float t2 = t * t, t3 = t2 * t;
    w0 = 0.5f * (-t + 2.f * t2 - t3);
    w1 = 0.5f * (2.f - 5.f * t2 + 3.f * t3);
    w2 = 0.5f * (t + 4.f * t2 - 3.f * t3);
    w3 = 0.5f * (-t2 + t3);
}

template <typename T> __device__ __forceinline__ void store_stream(T* p, T v);
template <> __device__ __forceinline__ void store_stream<float>(float* p, float v) { __stcs(p, v); }
template <> __device__ __forceinline__ void store_stream<double>(double* p, double v) { __stcs(p, v); }
// one 16-byte streaming store per thread and output row
__device__ __forceinline__ void store_stream_vec(float* p, const float (&v)[4]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
}
__device__ __forceinline__ void store_stream_vec(double* p, const double (&v)[2]) {
    __stcs(reinterpret_cast<double2*>(p), make_double2(v[0], v[1]));
}

// Out-of-line exact evaluation of one lattice cell (NaN in the footprint, or out of bounds): keeps
// the ring search and its local arrays out of the streaming loop's register budget.
template <typename T, int METHOD>
__device__ __noinline__ T lattice_cell_exact(const TileParams<T>* p, int64_t J, int I) {
    return static_cast<T>(interp_exact<T>(p->g, METHOD, __ldg(p->lon.coord + I), __ldg(p->lat.coord + J),
                                          __ldg(p->lon.pos + I), __ldg(p->lat.pos + J), nullptr));
}

// CTA = 128 column threads x 2 row groups.  A column thread owns COLS = 16/sizeof(T) adjacent output
// columns (one 16-byte vector store per output row); a row group sweeps its half of the tile's rows
// top to bottom, keeping the horizontal pass of the TAPS input rows under the current output row in
// registers and sliding that window as the (group-uniform) latitude table advances.
//
// WIN (FP32 bicubic at longitude factors 1 and 2 only; 0 = off): with a thread owning four adjacent output columns, adjacent
// lanes read their taps 4 / f_lon words apart -- 4-way (f_lon = 1) and 2-way (f_lon = 2) bank conflicts on sixteen 4-byte
// shared-memory loads per tile row, which made these launches LSU-bound (0.80-0.88 of the HBM peak,
// profiles/r02_ncu_upsample_small_factor.txt).  Here the thread loads the WINDOW that holds all its taps with three vector
// loads -- eight words from an 8-byte aligned start at f_lon = 1 (LDS.64, LDS.128, LDS.64), six at f_lon = 2 (3 x LDS.64) --
// and column c applies five weights to window words s_c .. s_c + 4 (s_c = c, resp. c / 2): its four Catmull-Rom weights
// shifted by k_c in {0, 1}, the fifth zero.  k_c absorbs the FP64 noise that moves floor() of a node-aligned query one cell
// down.  The host verifies on the longitude table that every column fits this pattern
// (window_mode) and launches the generic form otherwise.  A zero weight on a NaN tap yields NaN: such an output goes to the
// exact re-evaluation like any other dirty one.
template <typename T, int METHOD, int WIN = 0>
__global__ void __launch_bounds__(kTileThreads, sizeof(T) == 4 ? (WIN == 1 ? 3 : 4) : AUVI_F64_MINB)   // WIN = 1: wide boxes, <= 3 CTAs fit an SM anyway
upsample_tiled_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TileParams<T> p) {
    constexpr bool kCubic = (METHOD == CUBIC);
    static_assert(WIN == 0 || (METHOD == CUBIC && sizeof(T) == 4), "window loads: FP32 bicubic only");
    constexpr int WN = WIN == 1 ? 8 : 6;                           // window words
    constexpr int LO = kCubic ? 1 : 0;            // taps start at base-LO
    constexpr int TAPS = kCubic ? 4 : 2;
    constexpr bool kF64 = sizeof(T) == 8;
    constexpr int COLS = 16 / static_cast<int>(sizeof(T));
    constexpr int kTileCols = kColThreads * COLS;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* tile = reinterpret_cast<T*>(smem_raw);                      // [bh][bw]
    __shared__ uint64_t bar;
    __shared__ int s_top[kTileRowsMax + 1];                            // local row of the first tap
    __shared__ double2 s_ty[kF64 ? kTileRowsMax + 1 : 1];              // frac(pos_y) (NaN if out of bounds) and half of it
    __shared__ float4 s_wy[kF64 ? 1 : kTileRowsMax + 1];               // FP32 vertical tap weights

    const int tid = threadIdx.x;
    const int W = p.lon.n;
    const int I0 = blockIdx.x * kTileCols;
    const int64_t J0 = p.row_begin + static_cast<int64_t>(blockIdx.y) * p.tj;
    const int nJ = static_cast<int>(min(static_cast<int64_t>(p.tj), p.row_end - J0));
    // global column of tile column 0, rounded down to a 16-byte boundary: with no swizzle/interleave
    // the TMA unit faults ("illegal instruction") on a box whose first element is not 16-byte aligned
    // in global memory (measured on B200: profiles/r01_tma_alignment_probe.txt)
    const int r0 = __ldg(p.lat.base + J0) - LO;                    // global row of tile row 0
    const int bw = p.bw, bh = p.bh;
    // column slabs: slab s serves column threads [s*kColThreads/S, (s+1)*kColThreads/S) from its own box
    const int S = p.slabs, slab_cols = kTileCols / S, box_elems = bw * bh, box_stride = p.box_stride;
    auto slab_c0 = [&](int sl) -> int {                           // global column of box column 0 of slab sl
        const int Is = I0 + sl * slab_cols;
        return (__ldg(p.lon.base + (Is < W ? Is : W - 1)) - LO) & ~(COLS - 1);
    };

    // ---- stage the input box -------------------------------------------------------------------
    if (p.use_tma) {
        if (tid == 0) { prefetch_tmap(&tmap); mbar_init(&bar, 1); }
        __syncthreads();
        if (tid == 0) {
            int live = 0;
            for (int sl = 0; sl < S; ++sl) live += (I0 + sl * slab_cols < W);
            mbar_expect_tx(&bar, static_cast<uint32_t>(live * box_elems * sizeof(T)));
            for (int sl = 0; sl < live; ++sl) tma_load_2d(tile + sl * box_stride, &tmap, slab_c0(sl), r0 - p.g.row0, &bar);
        }
    } else {
        for (int sl = 0; sl < S && I0 + sl * slab_cols < W; ++sl) {
            const int c0 = slab_c0(sl);
            for (int k = tid; k < box_elems; k += kTileThreads) {  // coalesced, clamp-to-edge
                int lr = k / bw, lc = k - lr * bw;
                int gr = clampi(r0 + lr, 0, p.g.n_lat - 1), gc = clampi(c0 + lc, 0, p.g.n_lon - 1);
                // bh is the worst-case span over all tiles: a box may reach past the rows this tile uses -- and past
                // the resident slab.  Such rows are never read by the stencil; stay inside the allocation.
                gr = clampi(gr, p.rows_lo, p.rows_hi - 1);
                tile[sl * box_stride + k] = __ldg(p.g.z + static_cast<int64_t>(gr - p.g.row0) * p.g.ld + gc);
            }
        }
    }

    // ---- per-row and per-column setup (overlaps the TMA flight) ---------------------------------
    if (tid < nJ) {
        const double py = __ldg(p.lat.pos + J0 + tid);
        const int by = __ldg(p.lat.base + J0 + tid);
        const double ty = dsub(py, static_cast<double>(by));       // NaN stays NaN
        s_top[tid] = by - LO - r0;
        if constexpr (kF64) {
            s_ty[tid] = make_double2(ty, dmul(0.5, ty));
        } else {
            float4 w;
            if (kCubic) cr_weights(static_cast<float>(ty), w.x, w.y, w.z, w.w);
            else { w.x = 1.f - static_cast<float>(ty); w.y = static_cast<float>(ty); w.z = 0.f; w.w = 0.f; }
            s_wy[tid] = w;
        }
    }
    const int ct = tid % kColThreads, rg = tid / kColThreads;
    const int Ic = I0 + ct * COLS;                                 // first of this thread's columns
    const int my_slab = ct / (kColThreads / S);
    const int c0 = slab_c0(my_slab);                               // this thread's box
    const T* const my_tile = tile + my_slab * box_stride;
    int ox[COLS];
    double txd[COLS];
    double thx[kF64 && kCubic ? COLS : 1];                         // txd / 2: the last factor of each Catmull-Rom product (exact.cuh)
    float wx[kF64 ? 1 : COLS][4];
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
        const int I = Ic + c;
        double px = 0.0;
        int bx = c0 + LO;
        if (I < W) { px = __ldg(p.lon.pos + I); bx = __ldg(p.lon.base + I); }
        txd[c] = dsub(px, static_cast<double>(bx));
        ox[c] = bx - LO - c0;
        if constexpr (kF64 && kCubic) thx[c] = dmul(0.5, txd[c]);
        if constexpr (!kF64) {
            if (kCubic) cr_weights(static_cast<float>(txd[c]), wx[c][0], wx[c][1], wx[c][2], wx[c][3]);
            else { wx[c][0] = 1.f - static_cast<float>(txd[c]); wx[c][1] = static_cast<float>(txd[c]); wx[c][2] = 0.f; wx[c][3] = 0.f; }
        }
    }
    int w0 = 0;                                                    // WIN: first window word (box column, even)
    float w5[WIN ? COLS : 1][5];
    if constexpr (WIN > 0) {
        int lo = ox[0];
#pragma unroll
        for (int c = 1; c < COLS; ++c) {
            if (Ic + c >= W) ox[c] = ox[c - 1];                    // columns past the lattice: inside the window, never stored
            lo = min(lo, ox[c]);
        }
        w0 = lo & ~1;
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            const int k = ox[c] - w0 - (WIN == 2 ? c / 2 : c);     // 0 or 1 (host-checked: window_mode)
#pragma unroll
            for (int j = 0; j < 5; ++j)
                w5[c][j] = k == 0 ? (j < 4 ? wx[c][j < 4 ? j : 0] : 0.f) : (j >= 1 ? wx[c][j >= 1 ? j - 1 : 0] : 0.f);
        }
    }

    if (p.use_tma) {
        mbar_wait(&bar, 0);
        // The reference clamps stencil indices to the grid edge (GridH.cpp:242-247, :172-173); TMA
        // zero-fills instead, so border tiles copy the edge row/column into their out-of-grid halo.
        for (int sl = 0; sl < S && I0 + sl * slab_cols < W; ++sl) {
            const int b0 = slab_c0(sl);
            const bool border = b0 < 0 || b0 + bw > p.g.n_lon || r0 < 0 || r0 + bh > p.g.n_lat;
            if (!border) continue;
            T* const box = tile + sl * box_stride;
            for (int k = tid; k < box_elems; k += kTileThreads) {
                int lr = k / bw, lc = k - lr * bw;
                int gr = r0 + lr, gc = b0 + lc;
                int cr = clampi(gr, 0, p.g.n_lat - 1), cc = clampi(gc, 0, p.g.n_lon - 1);
                if (cr != gr || cc != gc) {
                    int sr = cr - r0, sc = cc - b0;                // in-grid source, never rewritten
                    if (sr >= 0 && sr < bh && sc >= 0 && sc < bw) box[k] = box[sr * bw + sc];
                }
            }
        }
    }
    __syncthreads();
    if (Ic >= W) return;

    // this row group's share of the tile rows
    const int per_group = (nJ + kRowGroups - 1) / kRowGroups;
    const int jr_begin = rg * per_group, jr_end = min(nJ, jr_begin + per_group);
    if (jr_begin >= jr_end) return;
    const bool full = p.vec_ok && (Ic + COLS <= W);
    T* out_row = p.out + (J0 - p.row_begin + jr_begin) * p.out_ld + Ic;
    const int64_t out_ld = p.out_ld;

    // ---- horizontal pass of one tile row for this thread's columns --------------------------------
    // A NaN anywhere in a footprint (or a NaN weight: out of bounds) surfaces in the horizontal pass
    // of some row the output uses, so it is enough to probe those (cheaper than probing every output).
    // FP32: the probe is a running sum (one FADD).  FP64: the kernel is bound by the FP64 pipe, so the probe is integer work on
    // the OUTPUTS instead -- every window row feeds some output of this thread, a NaN tap / weight therefore reaches one -- as the
    // running maximum of |v|'s high word (non-finite: >= 0x7ff00000).
    T probe = 0;
    int probe_hi = 0;
    auto hrow = [&](const T* r, T (&dst)[COLS]) {
        if constexpr (WIN > 0) {
            const float* const rp = reinterpret_cast<const float*>(r) + w0;
            float win[WN];
            const float2 a = *reinterpret_cast<const float2*>(rp);
            win[0] = a.x; win[1] = a.y;
            if constexpr (WIN == 1) {
                const float4 b = *reinterpret_cast<const float4*>(rp + 2);
                const float2 d = *reinterpret_cast<const float2*>(rp + 6);
                win[2] = b.x; win[3] = b.y; win[4] = b.z; win[5] = b.w; win[6] = d.x; win[7] = d.y;
            } else {
                const float2 b = *reinterpret_cast<const float2*>(rp + 2);
                const float2 d = *reinterpret_cast<const float2*>(rp + 4);
                win[2] = b.x; win[3] = b.y; win[4] = d.x; win[5] = d.y;
            }
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                constexpr int kHalf = WIN == 2 ? 2 : 1;
                const int sc = c / kHalf;
                dst[c] = fmaf(w5[c][4], win[sc + 4], fmaf(w5[c][3], win[sc + 3], fmaf(w5[c][2], win[sc + 2], fmaf(w5[c][1], win[sc + 1], w5[c][0] * win[sc]))));
                probe += dst[c];
            }
        } else {
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            const T* q = r + ox[c];
            if constexpr (kF64) {
                if constexpr (kCubic) dst[c] = catmull_rom_exact_h(q[0], q[1], q[2], q[3], txd[c], thx[c]);
                else dst[c] = dadd(dmul(dsub(1.0, txd[c]), q[0]), dmul(txd[c], q[1]));
            } else {
                if constexpr (kCubic) dst[c] = fmaf(wx[c][3], q[3], fmaf(wx[c][2], q[2], fmaf(wx[c][1], q[1], wx[c][0] * q[0])));
                else dst[c] = fmaf(wx[c][1], q[1], wx[c][0] * q[0]);
            }
            if constexpr (!kF64) probe += dst[c];
        }
        }
    };
    // one output row from the window; `ph` = slot holding the window's first row (a constant once the
    // phase loop below is unrolled, so the window never moves between registers)
    // FP64 bicubic: the Catmull-Rom coefficients of the vertical pass depend on the window only -- formed once per window
    // position, shared by the f_lat output rows under it (same operations, same bits: exact.cuh)
    CatmullCoef vk[kF64 && kCubic ? COLS : 1];
    // FP64 bicubic: s_ty[jr] / s_top[jr] are loaded one output row ahead, which takes the shared-memory latency out of the
    // dependent chain load -> multiply chain -> store of a row (2x1: 0.70 -> 0.78 of the HBM peak, 4x4: 0.93 -> 0.96, 2x2: 0.81 -> 0.82,
    // profiles/r02_upsample_small_factor_ab.txt)
    double2 ty_next = make_double2(0.0, 0.0);
    // Shared-space addresses of the row tables, pinned in registers: left to itself the compiler re-derives them (S2R
    // SR_CgaCtaId, LDC, MOV, LEA) at every window step of the sweep (FP64 bicubic 2x2: 0.505 -> 0.495 ms at 8192^2;
    // for the FP32 window kernel of factor 1 the same was mixed -- 1x1 0.77 -> 0.79, 4x1 -0.6 % -- and is not used).
    [[maybe_unused]] uint32_t sty_addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_ty));
    [[maybe_unused]] uint32_t stop_addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_top));
    if constexpr (kF64 && kCubic) asm volatile("" : "+r"(sty_addr), "+r"(stop_addr));
    int top_next = 0;
    // The same for the FP32 window-load kernel of longitude factor 1 (2x1 0.84 -> 0.91, 1x1 0.72 -> 0.77, 4x1 0.97 -> 0.99); the
    // factor-2 form LOSES with it (2x2 0.98 -> 0.94: eight more spill bytes inside its loop), so it keeps loading in place.
    constexpr bool kAheadF32 = WIN == 1;
    [[maybe_unused]] float4 wy_next = make_float4(0.f, 0.f, 0.f, 0.f);
    auto emit = [&](int jr, const T (&h)[TAPS][COLS], int ph) {
        T v[COLS];
        if constexpr (kF64 && kCubic) {
            const double2 ty2 = ty_next;
            // s_ty[jr + 1], s_top[jr + 1]: one entry of slack behind the last row
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(ty_next.x), "=d"(ty_next.y) : "r"(sty_addr + (jr + 1) * 16));
            asm volatile("ld.shared.s32 %0, [%1];" : "=r"(top_next) : "r"(stop_addr + (jr + 1) * 4));
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                v[c] = catmull_rom_eval(vk[c], h[(ph + 1) % TAPS][c], ty2.x, ty2.y);
                probe_hi = max(probe_hi, __double2hiint(v[c]) & 0x7fffffff);
            }
        } else if constexpr (kF64) {
            const double2 ty2 = s_ty[jr];
            const double ty = ty2.x;
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                if constexpr (kCubic)
                    v[c] = catmull_rom_exact_h(h[ph % TAPS][c], h[(ph + 1) % TAPS][c], h[(ph + 2) % TAPS][c], h[(ph + 3) % TAPS][c], ty, ty2.y);
                else v[c] = dadd(dmul(dsub(1.0, ty), h[ph % TAPS][c]), dmul(ty, h[(ph + 1) % TAPS][c]));
                probe_hi = max(probe_hi, __double2hiint(v[c]) & 0x7fffffff);
            }
        } else {
            float4 w;
            if constexpr (kAheadF32) { w = wy_next; wy_next = s_wy[jr + 1]; top_next = s_top[jr + 1]; }
            else w = s_wy[jr];
            probe += w.x;
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                if constexpr (kCubic)
                    v[c] = fmaf(w.w, h[(ph + 3) % TAPS][c], fmaf(w.z, h[(ph + 2) % TAPS][c], fmaf(w.y, h[(ph + 1) % TAPS][c], w.x * h[ph % TAPS][c])));
                else v[c] = fmaf(w.y, h[(ph + 1) % TAPS][c], w.x * h[ph % TAPS][c]);
            }
        }
        if (full) {
            store_stream_vec(out_row, v);
        } else {
#pragma unroll
            for (int c = 0; c < COLS; ++c)
                if (Ic + c < W) store_stream<T>(out_row + c, v[c]);
        }
        out_row += out_ld;
    };

    T h[TAPS][COLS];
    int jr = jr_begin;
    int top = s_top[jr];                                           // window = tile rows [top, top+TAPS)
    if constexpr (kF64 && kCubic) { ty_next = s_ty[jr]; top_next = top; }
    if constexpr (kAheadF32) { wy_next = s_wy[jr]; top_next = top; }
    const T* next_row = my_tile + top * bw;
#pragma unroll
    for (int k = 0; k < TAPS; ++k, next_row += bw) hrow(next_row, h[k]);
#pragma unroll 1
    for (;;) {
#pragma unroll
        for (int ph = 0; ph < TAPS; ++ph) {                        // slot (ph+k)%TAPS holds tile row top+k
            if constexpr (kF64 && kCubic) {                        // at every window position (formed lazily, inside the loop
#pragma unroll                                                     // below, it LOST 7 %: profiles/r02_upsample_small_factor_ab.txt)
                for (int c = 0; c < COLS; ++c)
                    vk[c] = catmull_rom_coef(h[ph % TAPS][c], h[(ph + 1) % TAPS][c], h[(ph + 2) % TAPS][c], h[(ph + 3) % TAPS][c]);
#pragma unroll 1
                while (jr < jr_end && top_next == top) {           // uniform across the row group
                    emit(jr, h, ph);
                    ++jr;
                }
            } else if constexpr (kAheadF32) {
#pragma unroll 1
                while (jr < jr_end && top_next == top) {
                    emit(jr, h, ph);
                    ++jr;
                }
            } else {
#pragma unroll 1
                while (jr < jr_end && s_top[jr] == top) {          // uniform across the row group
                    emit(jr, h, ph);
                    ++jr;
                }
            }
            if (jr >= jr_end) goto swept;
            hrow(next_row, h[ph]);                                 // the oldest slot takes tile row top+TAPS
            next_row += bw;
            ++top;
        }
    }
swept:
    const bool dirty = kF64 ? probe_hi >= 0x7ff00000 : isnan(probe);
    if (!dirty) return;
    // ---- second sweep, only for threads that produced a NaN: re-read the cells this thread wrote
    // and replace every NaN by the exact per-query evaluation (ring-search fallbacks, NaN-corner
    // means, out-of-bounds NaN).  Kept out of the streaming loop so that the call and its local
    // arrays do not cost the clean path registers.
    out_row = p.out + (J0 - p.row_begin + jr_begin) * p.out_ld + Ic;
    for (int jr = jr_begin; jr < jr_end; ++jr, out_row += p.out_ld) {
#pragma unroll 1
        for (int c = 0; c < COLS; ++c) {
            if (Ic + c >= W) break;
            if (isnan(out_row[c])) store_stream<T>(out_row + c, lattice_cell_exact<T, METHOD>(&p, J0 + jr, Ic + c));
        }
    }
}

// ---- one thread per output, exact path ----------------------------------------------------------
template <typename T, int METHOD, bool FILL>
__global__ void __launch_bounds__(256)
lattice_exact_kernel(const GridView<T> g, const AxisDev lat, const AxisDev lon, int64_t row_begin,
                     int64_t row_end, T* __restrict__ out, int64_t out_ld, int32_t* __restrict__ sel) {
    const int W = lon.n;
    const int64_t total = (row_end - row_begin) * W;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; k < total;
         k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t jr = k / W;
        const int I = static_cast<int>(k - jr * W);
        const int64_t J = row_begin + jr;
        Picked pk;
        pk.found = -3;                                             // -3: valid cell passed through
        double v;
        bool done = false;
        if (FILL) {
            v = g.at(static_cast<int>(J), I);
            done = !isnan(v);
        }
        if (!done) {
            v = interp_exact<T>(g, METHOD, __ldg(lon.coord + I), __ldg(lat.coord + J), __ldg(lon.pos + I),
                                __ldg(lat.pos + J), sel ? &pk : nullptr);
        }
        store_stream<T>(out + jr * out_ld + I, static_cast<T>(v));
        if (sel) {
            int32_t* s = sel + k * 9;
            const bool has = pk.found >= 0;
            s[0] = pk.found;
            for (int q = 0; q < 4; ++q) { s[1 + 2 * q] = has ? pk.i[q] : -1; s[2 + 2 * q] = has ? pk.j[q] : -1; }
        }
    }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

bool make_grid_tensor_map(const GridDesc& d, int box_w, int box_h, CUtensorMap* out, bool nan_fill) {
    const size_t es = d.dtype == DT_F64 ? 8 : 4;
    static const bool disabled = getenv("AUVI_NO_TMA") != nullptr;       // debugging / A-B measurements only
    if (disabled) return false;
    if (box_w > 256 || box_h > 256 || box_w < 1 || box_h < 1) return false;
    if ((d.ld * es) % 16 != 0 || (reinterpret_cast<uintptr_t>(d.z) % 16) != 0) return false;
    if ((box_w * es) % 16 != 0) return false;
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(d.n_lon), static_cast<cuuint64_t>(d.rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(d.ld * es)};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, d.dtype == DT_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                    const_cast<void*>(d.z), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Largest input span (in cells) any tile of `tile` consecutive outputs needs along one axis, when the
// first cell of a tile is (base - lo) rounded down to a multiple of `align` (a power of two).
static int max_span(const int* h_base, int64_t begin, int64_t end, int tile, int taps, int lo, int align) {
    int worst = 0;
    for (int64_t a = begin; a < end; a += tile) {
        int64_t b = (a + tile - 1 < end - 1) ? a + tile - 1 : end - 1;
        int first = (h_base[a] - lo) & ~(align - 1);
        int s = h_base[b] - lo + taps - first;
        if (s > worst) worst = s;
    }
    return worst;
}

// Does every group of four adjacent output columns (one thread of the tiled kernel) fit the window-load form of the FP32
// bicubic kernel?  Mirrors the kernel's own arithmetic: tap start of column c = base - 1 (columns past the lattice inherit
// their left neighbour's), window start = the smallest of the four rounded down to even, shift k_c = start_c - window - s_c
// with s_c = c / 2 (mode 2: longitude factor 2) or c (mode 1: factor 1) must be 0 or 1.  Box starts are multiples of four
// columns, so parities agree between global and box columns.  Returns 2, 1 or 0 (generic kernel).
static int window_mode(const AxisTables& lon) {
    static const bool off = getenv("AUVI_NO_WINDOW") != nullptr;  // A/B measurements only
    if (off) return 0;
    if (lon.window_mode_cache && *lon.window_mode_cache >= 0) return *lon.window_mode_cache;
    int verdict = 0;
    for (int mode = 2; mode >= 1 && !verdict; --mode) {
        bool ok = true;
        for (int g = 0; g < lon.n && ok; g += 4) {
            int start[4], lo = 0;
            for (int c = 0; c < 4; ++c) {
                start[c] = g + c < lon.n ? lon.h_base[g + c] - 1 : start[c - 1];
                lo = c == 0 || start[c] < lo ? start[c] : lo;
            }
            const int w0 = lo & ~1;
            for (int c = 0; c < 4; ++c) {
                const int k = start[c] - w0 - (mode == 2 ? c / 2 : c);
                ok = ok && (k == 0 || k == 1);
            }
        }
        if (ok) verdict = mode;
    }
    if (lon.window_mode_cache) *lon.window_mode_cache = verdict;
    return verdict;
}

template <typename T, int METHOD>
static cudaError_t launch_tiled(const GridDesc& d, const AxisTables& lat, const AxisTables& lon, int64_t row_begin,
                                int64_t row_end, void* out, int64_t out_ld, cudaStream_t st, LaunchInfo* info) {
    const int taps = METHOD == CUBIC ? 4 : 2;
    const int lo = METHOD == CUBIC ? 1 : 0;
    const size_t es = sizeof(T);
    const int align = static_cast<int>(16 / es);
    static const int tj_cap = getenv("AUVI_TILE_ROWS") ? atoi(getenv("AUVI_TILE_ROWS")) : kTileRowsMax;   // A/B measurements only
    int tj = tj_cap >= 16 && tj_cap <= kTileRowsMax ? tj_cap : kTileRowsMax;
    const int tile_cols = kColThreads * align;
    int slabs = 1;
    int bw = max_span(lon.h_base, 0, lon.n, tile_cols, taps, lo, align);
    bw = (bw + align - 1) / align * align;
    while (bw > 256 && slabs < 4) {                               // a TMA box holds at most 256 elements per dimension
        slabs *= 2;
        bw = max_span(lon.h_base, 0, lon.n, tile_cols / slabs, taps, lo, align);
        bw = (bw + align - 1) / align * align;
    }
    int bh = max_span(lat.h_base, row_begin, row_end, tj, taps, lo, 1);
    // Shared-memory budget per CTA.  Bilinear gains from four CTAs per SM (4 x 56 KB); bicubic re-runs its three-row
    // horizontal warm-up per row group, so it prefers tall tiles and settles for two or three CTAs (measured:
    // tools/run_upsample.py, 2x2 / 1x1 / 4x1 / 1x4 on 16384^2).
    const size_t budget = (METHOD == CUBIC ? 96 : 56) * 1024;
    while (static_cast<size_t>(slabs) * bw * bh * es > budget && tj > 16) {
        tj /= 2;
        bh = max_span(lat.h_base, row_begin, row_end, tj, taps, lo, 1);
    }
    // slab check: every input row a tile touches (after the reference's clamp) must be resident;
    // bicubic outputs with a NaN in their footprint fall back to the radius-10 ring search.
    const int reach = METHOD == CUBIC ? kMaxRadius + 1 : 0;
    int need_lo = lat.h_base[row_begin] - lo - reach, need_hi = lat.h_base[row_end - 1] - lo + taps - 1 + reach;
    need_lo = need_lo < 0 ? 0 : need_lo;
    need_hi = need_hi > d.n_lat - 1 ? d.n_lat - 1 : need_hi;
    if (need_lo < d.row0 || need_hi >= d.row0 + d.rows) return cudaErrorInvalidValue;

    TileParams<T> p;
    p.g = make_view<T>(d);
    p.lat = AxisDev{lat.coord, lat.pos, lat.base, lat.n};
    p.lon = AxisDev{lon.coord, lon.pos, lon.base, lon.n};
    p.row_begin = row_begin; p.row_end = row_end;
    p.out = static_cast<T*>(out); p.out_ld = out_ld;
    p.tj = tj; p.bw = bw; p.bh = bh; p.slabs = slabs;
    p.box_stride = static_cast<int>((static_cast<size_t>(bw) * bh * es + 127) / 128 * 128 / es);
    p.vec_ok = (reinterpret_cast<uintptr_t>(out) % 16 == 0 && (out_ld * es) % 16 == 0) ? 1 : 0;
    p.rows_lo = d.row0; p.rows_hi = d.row0 + d.rows;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    // The tensor map covers the resident slab; box rows outside it are zero-filled.  The slab check
    // above guarantees that every row a tile actually reads is resident, and rows outside the GLOBAL
    // grid are patched to the clamped edge row inside the kernel.
    p.use_tma = make_grid_tensor_map(d, bw, bh, &tmap) ? 1 : 0;

    size_t smem = static_cast<size_t>(slabs) * p.box_stride * es;
    auto kern = upsample_tiled_kernel<T, METHOD>;
    int win = 0;
    if constexpr (sizeof(T) == 4 && METHOD == CUBIC) {
        win = window_mode(lon);
        if (win == 2) kern = upsample_tiled_kernel<T, METHOD, 2>;
        if (win == 1) kern = upsample_tiled_kernel<T, METHOD, 1>;
        if (win) smem += 32;                                       // a window may end a few words past its thread's last tap
    }
    // the 48 KB a kernel gets without opting in count its STATIC shared memory (2.7 KB here) too: a dynamic size of 46-48 KB
    // used to fail the launch with "invalid argument" (found by test_lattice_f32_bicubic_window_loads: 64 x 300, 3 x 2)
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    dim3 grid(static_cast<unsigned>((lon.n + tile_cols - 1) / tile_cols),
              static_cast<unsigned>((row_end - row_begin + tj - 1) / tj));
    kern<<<grid, kTileThreads, smem, st>>>(tmap, p);
    if (info) { info->launches += 1; info->used_tma = p.use_tma; info->used_window = win; }
    return cudaGetLastError();
}

template <typename T, int METHOD, bool FILL>
static cudaError_t launch_exact(const GridDesc& d, const AxisTables& lat, const AxisTables& lon, int64_t row_begin,
                                int64_t row_end, void* out, int64_t out_ld, int32_t* sel, cudaStream_t st,
                                LaunchInfo* info) {
    // every row the ring search (radius 10 around a floor/round centre) may read must be resident
    int need_lo = lat.h_base[row_begin] - (kMaxRadius + 1), need_hi = lat.h_base[row_end - 1] + (kMaxRadius + 2);
    need_lo = need_lo < 0 ? 0 : need_lo;
    need_hi = need_hi > d.n_lat - 1 ? d.n_lat - 1 : need_hi;
    if (need_lo < d.row0 || need_hi >= d.row0 + d.rows) return cudaErrorInvalidValue;
    if (info) { info->launches += 1; info->used_tma = 0; }
    const int64_t total = (row_end - row_begin) * lon.n;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 64) blocks = 148 * 64;
    lattice_exact_kernel<T, METHOD, FILL><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        make_view<T>(d), AxisDev{lat.coord, lat.pos, lat.base, lat.n}, AxisDev{lon.coord, lon.pos, lon.base, lon.n},
        row_begin, row_end, static_cast<T*>(out), out_ld, sel);
    return cudaGetLastError();
}

template <typename T>
static cudaError_t launch_lattice_t(const GridDesc& d, int method, const AxisTables& lat, const AxisTables& lon,
                                    int64_t row_begin, int64_t row_end, void* out, int64_t out_ld, int fill,
                                    int32_t* sel, cudaStream_t st, LaunchInfo* info) {
    if (!fill && !sel) {
        if (method == BILINEAR) return launch_tiled<T, BILINEAR>(d, lat, lon, row_begin, row_end, out, out_ld, st, info);
        if (method == CUBIC) return launch_tiled<T, CUBIC>(d, lat, lon, row_begin, row_end, out, out_ld, st, info);
    }
    if (!sel && method != IDW_KNN && (fill || method == KRIGING || method == NN || method == IDW)) {   // IDW_KNN: per-query path
        static const bool v0 = getenv("AUVI_FILL_V0") != nullptr;        // A-B measurements only
        if (!v0) return launch_fill(d, method, lat, lon, row_begin, row_end, out, out_ld, fill, st, info);
    }
#define AUVI_CASE(M)                                                                                         \
    case M:                                                                                                  \
        return fill ? launch_exact<T, M, true>(d, lat, lon, row_begin, row_end, out, out_ld, sel, st, info)  \
                    : launch_exact<T, M, false>(d, lat, lon, row_begin, row_end, out, out_ld, sel, st, info);
    switch (method) {
        AUVI_CASE(BILINEAR) AUVI_CASE(CUBIC) AUVI_CASE(KRIGING) AUVI_CASE(NN) AUVI_CASE(IDW) AUVI_CASE(BILINEAR_SEARCH) AUVI_CASE(IDW_KNN)
        default: return cudaErrorInvalidValue;
    }
#undef AUVI_CASE
}

cudaError_t launch_lattice(const GridDesc& d, int method, const AxisTables& lat, const AxisTables& lon,
                           int64_t row_begin, int64_t row_end, void* out, int64_t out_ld, int fill, int32_t* sel,
                           cudaStream_t st, LaunchInfo* info) {
    if (row_end <= row_begin) return cudaSuccess;
    if (d.dtype == DT_F64)
        return launch_lattice_t<double>(d, method, lat, lon, row_begin, row_end, out, out_ld, fill, sel, st, info);
    return launch_lattice_t<float>(d, method, lat, lon, row_begin, row_end, out, out_ld, fill, sel, st, info);
}

}  // namespace auvi
