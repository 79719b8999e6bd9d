// fill.cu -- the tiled ring-search kernel.
//
// FILL = true: full-grid gap fill -- every NaN cell of the grid gets METHOD(query at its own node), valid cells pass
// through.  This is the structured form of the reference's Grid-B procedure (test_gebco.cpp:150-196: one query per
// removed cell, built by gridIndexToGeo :72-81, evaluated by GridH::batch*Interpolate, GridH.cpp:160-420) and of
// BASELINE config 4 (IDW / nearest-neighbour on a 70 % masked grid).  Methods: BILINEAR (four corners, NaN-corner mean:
// no search; inside this kernel only as the first stage of BILINEAR_SEARCH), CUBIC (always the ring-search 4-nearest mean here: a masked cell is inside its own 4x4 stencil), KRIGING,
// NN, IDW, and the opt-in BILINEAR_SEARCH.
// FILL = false: KRIGING / NN / IDW on an upsampling lattice (test_interpolation.cpp:283-297): every output cell a query.
//
// Design (one CTA = 256 threads = one 64 x 32 tile of output cells; DESIGN.md section 5.3 has the numbers):
//   1. The tile's grid cells + a 12-cell halo (88 x 56: search radius 10 + a centre that FP64 noise may move by one
//      cell) are staged into shared memory by ONE TMA 2-D box load; meanwhile per-axis tables (index-space position,
//      search centre, squared offsets of the near rings in the reference's operation order) and two small LUTs.
//   2. Warps turn the block into validity bitmasks with __ballot_sync (one 96-bit row per block row); from here on the
//      ring search of GridH.cpp:24-118 is bit arithmetic.
//   3. A warp takes whole tile rows: valid cells go straight to the output, masked cells are COMPACTED into a CTA-wide
//      queue (popcounts of the warp-uniform row words, one shared atomic per row).
//   4. Phase A1, one thread per query: the 5 x 5 block around the centre as ONE word whose bit order is the reference's
//      enumeration order; four popcounts give the pass at which the reference stops, a mask gives its candidates.
//      Searches that end inside the block with <= 8 candidates take the near path; the others get the general
//      termination scan (phase A2, rings to radius 10).
//   5. Records are counting-sorted by (path, candidate count) so that a warp's queries have the same list length.
//   6. Phase B: warps draw chunks of 32 records from a CTA-wide counter.  Near path: the candidate list lives in
//      registers (length 4 / 6 / 8 chosen per chunk) and the reference's partial selection sort WITH swaps runs on it
//      with selects; distances are compared squared and the sorted chain is checked afterwards for a gap that sqrt
//      rounding could close -- such queries are replayed with the square roots taken first.  General path: list in
//      shared memory, the selection literally with a sqrt guard per comparison; what it cannot decide goes to the
//      literal per-query path (exact.cuh).
//   7. Epilogue on the four picks: mean / first pick / FP32 IDW weights through the SFU reciprocal / kriging (its FP64
//      solve runs as a phase of its own over the recorded picks).
//   8. PATCH instantiations (kriging, IDW, 4-nearest mean by default; every method when the output is peer memory): results are
//      patched into the staged tile -- a masked cell is never read by a search, candidates are valid cells -- and the finished
//      tile leaves as 16-byte row stores; otherwise one streaming store per query and early pass-through stores.
//
// Algorithmic HBM bytes: sizeof(T) read + sizeof(T) written per cell; the kernel is issue-bound (integer /
// bit work + FP64 distance compares), not HBM-bound.
//
// BILINEAR gap fill has a kernel of its own at the end of this file (bilinear_fill_kernel: persistent CTAs, two-deep TMA
// ring, 16-byte stores) -- it needs no search and is HBM-bound.  The same restructuring of THIS kernel (round 2, "v5":
// persistent CTAs, two-deep TMA ring, vectorised bitmasks, results patched into the staged tile and stored as whole rows)
// was built, parity-checked and measured on the same box against this form: it lost 6-13 % at every mask fraction
// (profiles/r02_fill_v5_ab.txt), because the kernel is bound by issue slots and kernel text, not by
// the TMA latency three resident CTAs per SM already hide.  Out-of-bounds cells of the staged block arrive as NaN (TMA
// fill mode), so "outside the grid / outside the resident slab" and "masked" are one test.
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
constexpr int kFCells = kFW * kFH;
constexpr int kFPer = kFCells / kFThreads;     // queries per thread when every cell is masked
constexpr int kFNMax = 13;                     // per-thread candidate list capacity of the general path (3 + a full ring-2 row pass)
constexpr int kFRedoMax = 128;
constexpr int kFReplayMax = 256;                // near-tie queries replayed with square roots (more: literal path)
constexpr int kFSq = 3;                        // squared-offset tables cover |offset| <= kFSq
constexpr int kFNear = 8;                      // near path: at most this many candidates, all within the 5 x 5 block
constexpr int kFUnroll = 3;                    // rings unrolled with their row windows in registers; further rings: rolled loops
constexpr int kFBinNear = 110;                 // bins 0..109: general path (termination x count); 110..114: near path by count
constexpr int kFBins = 128;                    // (the bin travels in 7 bits of a record; 127 is reserved)

// 5 x 5 block around the search centre as one word whose bit order IS the reference's enumeration order
// (GridH.cpp:36-117): bit 0 the centre; 1..6 ring-1 top/bottom rows (per column left to right: top, bottom);
// 7..8 ring-1 left/right; 9..18 ring-2 top/bottom rows; 19..24 ring-2 left/right (per row top to bottom: left,
// right).  The cells the reference has visited at each of its `count >= 4` checks are then the low 7 / 9 / 19 / 25
// bits, and find-first-set walks the candidates in list order.
__host__ __device__ constexpr int near_pos(int dy, int dx) {
    return (dy == 0 && dx == 0) ? 0
         : (dy == -1 || dy == 1) && dx >= -1 && dx <= 1 ? 1 + 2 * (dx + 1) + (dy > 0)
         : (dy == 0 && (dx == -1 || dx == 1)) ? 7 + (dx > 0)
         : (dy == -2 || dy == 2) ? 9 + 2 * (dx + 2) + (dy > 0)
         : 19 + 2 * (dy + 1) + (dx > 0);
}
constexpr uint32_t kC1 = 0x7Fu, kC2 = 0x1FFu, kC3 = 0x7FFFFu, kC4 = 0x1FFFFFFu;

constexpr double kSafeRatio = 1.0 - 8.8817841970012523e-16;   // 1 - 2^-50: a gap that sqrt rounding cannot close

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
    int n_out_cols;                           // lattice columns (== g.n_lon on the node lattice)
    int f_lat, f_lon;                         // integer upsampling factors of the lattice axes (1 on the node lattice)
    int use_tma;
    int tiles_x, n_tiles;                     // bilinear_fill_kernel: tiles per tile row, tiles in all (row-major order)
    int vec_ok;                               // bilinear_fill_kernel: out / out_ld allow 16-byte stores
};

template <typename T>
struct FillSmem {
    alignas(128) T tile[kFBH * kFBW];
    double d2[kFNMax * kFThreads];            // general path: per-thread candidate lists, squared distances.  Before the
                                              // general path runs, its first 8 KB hold two queues (see the kernel)
    double sqx[kFW * (2 * kFSq + 1)];         // ((cx + d + 0.5) - x)^2 per tile column, d = -kFSq..kFSq
    double sqy[kFH * (2 * kFSq + 1)];
    double x[kFW], y[kFH];
    uint32_t rec[kFCells];                    // query records, ordered by bin after the scatter
    uint32_t mask[kFBH * 4];
    uint32_t lut[5 * 32];                     // 5-bit row window of block row dy+2 -> its bits in enumeration order
    uint32_t cell_off[32];                    // enumeration position -> byte offsets into the sqx / sqy rows: x | y << 16
    int cell_tile[32];                        // enumeration position -> tile offset dy*kFBW + dx
    int cell_dxy[32];                         // enumeration position -> (dx & 0xffff) | dy << 16
    int cx[kFW], cy[kFH];
    int hist[kFBins];
    uint16_t code[kFNMax * kFThreads];        // general path: per-thread candidate lists, packed (dy,dx) offsets
    uint16_t reck[kFCells];                   // cell (lj*kFW + li) of each record
    uint16_t redo[kFRedoMax];
    uint16_t replay[kFReplayMax];
    int qn, rn, dn, tn, q_near, next;
    uint64_t bar;
};

static_assert(sizeof(FillSmem<float>) <= 75 * 1024, "three CTAs per SM need <= 75 KB each (f32 grids)");
static_assert(sizeof(FillSmem<double>) <= 113 * 1024, "two CTAs per SM (f64 grids)");

// A result's way out.  PATCH = false: a streaming store to the output.  PATCH = true: into the staged tile in shared memory
// (`dst` then points there) -- a masked cell is never read by a search, candidates are valid cells -- and the finished tile
// leaves with 16-byte row stores at the end of the kernel: what the output wants when it is PEER memory (the gather fused
// into the kernel), where per-query 4-byte stores crawl over NVLink.
template <bool PATCH, typename T>
__device__ __forceinline__ void put(T* dst, T v) {
    if (PATCH) *dst = v;
    else __stcs(dst, v);
}

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

// ---- helper: selection in registers ------------------------------------------------------------------------------
// The reference's partial selection sort with swaps (GridH.cpp:123-140) on a list of N (distance, cell) pairs held
// in registers: pass m finds the FIRST strict minimum of entries m..N-1 and swaps it with entry m.  Cells are
// distinct, so the winner is recognised by its cell and the swap is spelled with selects: no register is indexed
// dynamically.  Unused tail entries hold +inf and never win.
template <int N, int PASSES>
__device__ __forceinline__ void select_in_registers(double (&d)[N], int (&c)[N]) {
#pragma unroll
    for (int m = 0; m < PASSES && m + 1 < N; ++m) {
        double dbest = d[m];
        int cbest = c[m];
#pragma unroll
        for (int k = m + 1; k < N; ++k) {
            const bool lt = d[k] < dbest;
            dbest = lt ? d[k] : dbest; cbest = lt ? c[k] : cbest;
        }
#pragma unroll
        for (int k = m + 1; k < N; ++k) {
            const bool here = c[k] == cbest;
            d[k] = here ? d[m] : d[k]; c[k] = here ? c[m] : c[k];
        }
        d[m] = dbest; c[m] = cbest;
    }
}

// ---- helper: ordinary kriging on four picks, out of line ---------------------------------------------------------------
// One copy of the FP64 system assembly + solve (exact.cuh, GridH.cpp:361-419) shared by every call site, with its own
// register allocation: inlined into each list-length variant of the near path it cost the whole kernel spills.
template <typename T>
__device__ __noinline__ double kriging_four(const GridView<T>* g, int i0, int i1, int i2, int i3, int j0, int j1, int j2,
                                            int j3, double v0, double v1, double v2, double v3, double lon, double lat) {
    Picked pk;
    pk.found = 4;
    pk.i[0] = i0; pk.i[1] = i1; pk.i[2] = i2; pk.i[3] = i3;
    pk.j[0] = j0; pk.j[1] = j1; pk.j[2] = j2; pk.j[3] = j3;
    pk.v[0] = v0; pk.v[1] = v1; pk.v[2] = v2; pk.v[3] = v3;
    pk.d[0] = pk.d[1] = pk.d[2] = pk.d[3] = 0.0;
    return kriging_from_picked(*g, pk, lon, lat);
}

// ---- helper: the method's value from four picks ---------------------------------------------------------------------
// v = the picks' values in selection order, d2 = their SQUARED index-space distances (> 0), (pi,pj) their global cells.
template <typename T, int METHOD>
__device__ __forceinline__ T finish_four(const FillParams<T>& p, const T (&v)[4], const double (&d2)[4], const int (&pi)[4],
                                         const int (&pj)[4], int I, int64_t J) {
    if (METHOD == NN) return v[0];
    if (METHOD == CUBIC || METHOD == BILINEAR_SEARCH) {
        // fallbackAverage (GridH.cpp:10-18) of four numbers: ((0 + a) + b + c + d) / 4
        const double sum = dadd(dadd(dadd(dadd(0.0, static_cast<double>(v[0])), static_cast<double>(v[1])),
                                     static_cast<double>(v[2])), static_cast<double>(v[3]));
        return static_cast<T>(dmul(sum, 0.25));                     // == sum / 4, bit for bit
    }
    if (METHOD == IDW) {
        // FP32 weights 1/d^2 through the SFU reciprocal; values centred on the first pick (== exact.cuh idw_from_picked)
        if (d2[0] == 0.0) return v[0];                               // on a cell centre (picks are in ascending order)
        float num = 0.f, den = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float w = rcp_sfu(static_cast<float>(d2[e]));
            if (e) num = fmaf(w, static_cast<float>(v[e] - v[0]), num);
            den += w;
        }
        return static_cast<T>(v[0] + static_cast<T>(__fdividef(num, den)));
    }
    return static_cast<T>(kriging_four<T>(&p.g, pi[0], pi[1], pi[2], pi[3], pj[0], pj[1], pj[2], pj[3], static_cast<double>(v[0]),
                                          static_cast<double>(v[1]), static_cast<double>(v[2]), static_cast<double>(v[3]),
                                          __ldg(p.lon.coord + I), __ldg(p.lat.coord + J)));
}

// ---- helper: fewer than four candidates (the search ran out of rings) -------------------------------------------------
template <typename T, int METHOD>
__device__ __noinline__ T finish_few(int cnt, double v0, double v1, double v2, double d0, double d1, double d2) {
    Picked pk;
    pk.found = cnt;
    pk.v[0] = v0; pk.v[1] = v1; pk.v[2] = v2; pk.v[3] = qnan();
    pk.d[0] = d0; pk.d[1] = d1; pk.d[2] = d2; pk.d[3] = qnan();
    if (cnt == 0) return static_cast<T>(qnan());
    if (METHOD == NN) return static_cast<T>(pk.v[0]);              // the caller ran the first-minimum pass
    if (METHOD == IDW) {
        float num = 0.f, den = 0.f;
        for (int e = 0; e < cnt; ++e) {
            if (pk.d[e] == 0.0) return static_cast<T>(pk.v[e]);
            const float w = rcp_sfu(static_cast<float>(pk.d[e]));
            num = fmaf(w, static_cast<float>(pk.v[e] - pk.v[0]), num);
            den += w;
        }
        return static_cast<T>(pk.v[0] + static_cast<double>(__fdividef(num, den)));
    }
    return static_cast<T>(mean_found(pk));                          // CUBIC and KRIGING: GridH.cpp:291-298, :350-356
}

// ---- helper: one near-path query ---------------------------------------------------------------------------------------
// The candidates are the set bits of `cand`, in list order; their squared distances come from the per-tile tables
// (same operations as GridH.cpp:42-44).  The selection runs on squared distances (REPLAY = false): sqrt is monotone,
// so a pass can only come out differently if a remaining value lies above the pass minimum by less than sqrt
// rounding can close.  That is checked on the sorted chain afterwards; such a query returns false and is replayed
// with REPLAY = true, which takes the square roots first -- the reference's own comparison, bit for bit.
template <typename T, int METHOD, int N, bool REPLAY, bool PATCH>
__device__ __forceinline__ bool near_query(FillSmem<T>& s, const FillParams<T>& p, uint32_t cand, int q, int k, int c0,
                                           int r0, int I0, int64_t J0, T* out_tile, int64_t tile_ld) {
    constexpr int kPasses = METHOD == NN ? 1 : 4;
    const int lj = k / kFW, li = k % kFW;
    const char* const sqx = reinterpret_cast<const char*>(s.sqx + li * (2 * kFSq + 1) + kFSq - 2);   // [dx + 2]
    const char* const sqy = reinterpret_cast<const char*>(s.sqy + lj * (2 * kFSq + 1) + kFSq - 2);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double d[N];
    int c[N];
#pragma unroll
    for (int e = 0; e < N; ++e) {
        const bool has = cand != 0u;
        const int b = (__ffs(cand) - 1) & 31;                      // exhausted list: cell 31, overwritten by +inf
        cand &= cand - 1u;
        const uint32_t off = s.cell_off[b];
        double v = dadd(*reinterpret_cast<const double*>(sqx + (off & 0xffffu)),
                        *reinterpret_cast<const double*>(sqy + (off >> 16)));
        if (REPLAY) v = dsqrt(v);                                   // GridH.cpp:44
        d[e] = has ? v : inf;
        c[e] = b;
    }
    select_in_registers<N, kPasses>(d, c);
    if (!REPLAY) {
        // smallest value left behind that is strictly above the last pick, then the chain of picks
        double rest = inf;
#pragma unroll
        for (int e = kPasses; e < N; ++e) rest = (d[e] > d[kPasses - 1] && d[e] < rest) ? d[e] : rest;
        bool tie = d[kPasses - 1] != rest && dmul(rest, kSafeRatio) < d[kPasses - 1];
#pragma unroll
        for (int e = 0; e + 1 < kPasses; ++e) tie |= d[e] != d[e + 1] && dmul(d[e + 1], kSafeRatio) < d[e];
        if (tie) return false;
    }
    if (METHOD == KRIGING) {
        // the FP64 system is solved in a phase of its own (kernel, "kriging of the near-path picks"): here only the
        // picks are recorded, so the selection does not carry the solver's registers
        s.rec[q] = static_cast<uint32_t>(c[0] | (c[1] << 5) | (c[2] << 10) | (c[3] << 15)) | (127u << 25);
        return true;
    }
    const int cig = s.cx[li], cjg = s.cy[lj];
    const T* const centre = s.tile + (cjg - r0) * kFBW + (cig - c0);
    T v[4];
    double d2[4];
    const int pi[4] = {0, 0, 0, 0}, pj[4] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        v[e] = centre[s.cell_tile[c[e]]];
        if (REPLAY) {                                               // the squared distance as first formed, not sqrt(.)^2
            const uint32_t off = s.cell_off[c[e]];
            d2[e] = dadd(*reinterpret_cast<const double*>(sqx + (off & 0xffffu)), *reinterpret_cast<const double*>(sqy + (off >> 16)));
        } else d2[e] = d[e];
    }
    put<PATCH>(out_tile + lj * tile_ld + li, finish_four<T, METHOD>(p, v, d2, pi, pj, I0 + li, J0 + lj));
    return true;
}

template <typename T, int METHOD, bool FILL, bool PATCH = false>
__global__ void __launch_bounds__(kFThreads, 3)
fill_tiled_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ FillParams<T> p) {
    constexpr bool kPatch = PATCH && FILL && METHOD != BILINEAR && METHOD != BILINEAR_SEARCH;   // bilinear READS masked corners
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FillSmem<T>& s = *reinterpret_cast<FillSmem<T>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = p.g.n_lon;                                       // grid columns; the lattice has p.n_out_cols
    const int I0 = blockIdx.x * kFW;
    const int64_t J0 = p.row_begin + static_cast<int64_t>(blockIdx.y) * kFH;
    // first grid column / row of the staged block: the tile's first node minus the halo, the column rounded down to a
    // 16-byte boundary for TMA (on the node lattice I0 - 12 already is one)
    const int c0 = (I0 / p.f_lon - kFHalo) & ~3;
    const int r0 = static_cast<int>(J0 / p.f_lat) - kFHalo;
    T* const g_out_tile = p.out + (J0 - p.row_begin) * p.out_ld + I0;
    // where results go: the output, or (PATCH) the tile's own cells inside the staged block
    T* const out_tile = kPatch ? s.tile + kFHalo * kFBW + kFHalo : g_out_tile;
    const int64_t tile_ld = kPatch ? kFBW : p.out_ld;
    // two queues live in the general path's list storage until that path starts:
    uint16_t* const queue = reinterpret_cast<uint16_t*>(s.d2);     // masked cells of the tile, compacted (read by phase A1)
    uint16_t* const defer = queue + kFCells;                       // A1 -> A2: queries the 5 x 5 block cannot decide

    if (tid == 0) { s.qn = 0; s.rn = 0; s.dn = 0; s.tn = 0; s.next = 0; }
    if (tid < kFBins) s.hist[tid] = 0;
    if (tid < 5 * 32) {
        const int dy = tid / 32 - 2, w = tid & 31;
        uint32_t v = 0;
#pragma unroll
        for (int dx = -2; dx <= 2; ++dx) if ((w >> (dx + 2)) & 1) v |= 1u << near_pos(dy, dx);
        s.lut[tid] = v;
    } else if (tid < 6 * 32) {
        const int b = tid - 5 * 32;                                 // cells 25..31 are never looked up; keep them harmless
        const int dy = b < 25 ? b / 5 - 2 : 0, dx = b < 25 ? b % 5 - 2 : 0;
        const int pos = b < 25 ? near_pos(dy, dx) : b;
        s.cell_off[pos] = static_cast<uint32_t>((dx + 2) * 8) | (static_cast<uint32_t>((dy + 2) * 8) << 16);
        s.cell_tile[pos] = dy * kFBW + dx;
        s.cell_dxy[pos] = (dx & 0xffff) | (dy << 16);
    }
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
        if (I < p.n_out_cols) {
            x = __ldg(p.lon.pos + I);
            c = (METHOD == CUBIC || METHOD == BILINEAR || METHOD == BILINEAR_SEARCH) ? __ldg(p.lon.base + I) : (isnan(x) ? 0 : round_centre(x, p.g.n_lon));
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
            c = (METHOD == CUBIC || METHOD == BILINEAR || METHOD == BILINEAR_SEARCH) ? __ldg(p.lat.base + J) : (isnan(y) ? 0 : round_centre(y, p.g.n_lat));
        }
        s.y[t] = y; s.cy[t] = c;
        const double cf = dadd(__int2double_rn(c), 0.5);
#pragma unroll
        for (int d = -kFSq; d <= kFSq; ++d) s.sqy[t * (2 * kFSq + 1) + d + kFSq] = sq_offset(cf, d, y);
    }
    if (p.use_tma) mbar_wait(&s.bar, 0);
    else __syncthreads();

    // ---- validity bitmasks: bit c of row r = tile cell (r,c) holds a number and lies inside the grid ----
    // (cells outside the grid or outside the resident slab arrive as NaN: the TMA map fills them so, the plain loader
    // writes them so)
    for (int r = warp; r < kFBH; r += kFThreads / 32) {
#pragma unroll
        for (int seg = 0; seg < 3; ++seg) {
            const int c = seg * 32 + lane;
            bool ok = false;
            if (c < kFBW) ok = !isnan(s.tile[r * kFBW + c]);
            const uint32_t word = __ballot_sync(0xffffffffu, ok);
            if (lane == 0) s.mask[r * 4 + seg] = word;
        }
        if (lane == 0) s.mask[r * 4 + 3] = 0u;
    }
    __syncthreads();

    // ---- pass valid cells through (coalesced), compact masked cells into the queue -------------------------
    // A warp takes whole tile rows: the row's validity is two words of the bitmask, the same for every lane, so the
    // queue slots come from popcounts instead of ballots, one shared-memory atomic per row.
    {
        const int ncols = min(kFW, p.n_out_cols - I0);
        const uint32_t range_lo = ncols >= 32 ? 0xffffffffu : (1u << ncols) - 1u;
        const uint32_t range_hi = ncols >= 64 ? 0xffffffffu : (ncols > 32 ? (1u << (ncols - 32)) - 1u : 0u);
        const uint32_t below = (1u << lane) - 1u;
        for (int lj = warp; lj < kFH; lj += kFThreads / 32) {
            if (J0 + lj >= p.row_end) break;
            const uint32_t* const mrow = s.mask + (lj + kFHalo) * 4;
            const uint32_t v_lo = __funnelshift_r(mrow[0], mrow[1], kFHalo), v_hi = __funnelshift_r(mrow[1], mrow[2], kFHalo);
            const uint32_t todo_lo = FILL ? ~v_lo & range_lo : range_lo, todo_hi = FILL ? ~v_hi & range_hi : range_hi;
            const int n_lo = __popc(todo_lo), n_row = n_lo + __popc(todo_hi);
            int base = 0;
            if (lane == 0 && n_row) base = atomicAdd(&s.qn, n_row);
            base = __shfl_sync(0xffffffffu, base, 0);
            const T* const trow = s.tile + (lj + kFHalo) * kFBW + kFHalo;
            T* const orow = g_out_tile + lj * p.out_ld;
            if (FILL && !kPatch) {
                if ((v_lo & range_lo) >> lane & 1u) __stcs(orow + lane, trow[lane]);
                if ((v_hi & range_hi) >> lane & 1u) __stcs(orow + 32 + lane, trow[32 + lane]);
            }
            if (todo_lo >> lane & 1u) queue[base + __popc(todo_lo & below)] = static_cast<uint16_t>(lj * kFW + lane);
            if (todo_hi >> lane & 1u) queue[base + n_lo + __popc(todo_hi & below)] = static_cast<uint16_t>(lj * kFW + 32 + lane);
        }
    }
    __syncthreads();

    // ---- BILINEAR: no search -- the four corners around the query sit in the tile (GridH.cpp:160-210) -----------------
    if (METHOD == BILINEAR || METHOD == BILINEAR_SEARCH) {
        const int qn_b = s.qn;
        for (int q = tid; q < qn_b; q += kFThreads) {
            const int k = queue[q];
            const int lj = k / kFW, li = k % kFW;
            const double x = s.x[li], y = s.y[lj];
            double result = qnan();
            if (!isnan(x) && !isnan(y)) {
                const int x0 = s.cx[li], y0 = s.cy[lj];             // floor(x), floor(y)
                const int x1 = min(x0 + 1, p.g.n_lon - 1), y1 = min(y0 + 1, p.g.n_lat - 1);
                const double wx = dsub(x, __int2double_rn(x0)), wy = dsub(y, __int2double_rn(y0));
                const T* const r0p = s.tile + (y0 - r0) * kFBW - c0;
                const T* const r1p = s.tile + (y1 - r0) * kFBW - c0;
                const double a = static_cast<double>(r0p[x0]), b = static_cast<double>(r0p[x1]);
                const double c = static_cast<double>(r1p[x0]), d = static_cast<double>(r1p[x1]);
                if (isnan(a) || isnan(b) || isnan(c) || isnan(d)) result = mean_valid4(a, b, c, d);
                else {
                    const double ux = dsub(1.0, wx);
                    const double lo = dadd(dmul(ux, a), dmul(wx, b));
                    const double hi = dadd(dmul(ux, c), dmul(wx, d));
                    result = dadd(dmul(dsub(1.0, wy), lo), dmul(wy, hi));
                }
            }
            // BILINEAR_SEARCH (opt-in): a query whose four corners are all missing goes on to the ring search
            if (METHOD == BILINEAR_SEARCH && isnan(result) && !isnan(x) && !isnan(y)) defer[atomicAdd(&s.dn, 1)] = static_cast<uint16_t>(k);
            else put<kPatch>(out_tile + lj * tile_ld + li, static_cast<T>(result));
        }
        if (METHOD == BILINEAR) return;
        __syncthreads();
        const int left = s.dn;
        for (int q = tid; q < left; q += kFThreads) queue[q] = defer[q];
        __syncthreads();
        if (tid == 0) { s.qn = left; s.dn = 0; }
        __syncthreads();
    }

    auto window = [&](int row, int wi, int sh) -> uint32_t {       // validity of columns ci-10..ci+10 of a tile row
        return __funnelshift_r(s.mask[row * 4 + wi], s.mask[row * 4 + wi + 1], sh) & 0x1FFFFFu;
    };
    auto to_literal = [&](int k) {                                  // hand a query to the literal per-query path
        const int slot = atomicAdd(&s.rn, 1);
        if (slot < kFRedoMax) s.redo[slot] = static_cast<uint16_t>(k);
        else put<kPatch>(out_tile + (k / kFW) * tile_ld + (k % kFW), fill_cell_literal<T, METHOD>(&p, J0 + k / kFW, I0 + k % kFW));
    };

    // ---- phase A1: the 5 x 5 block around each query's centre decides most searches ---------------------------
    // (GridH.cpp:48-117: the count is checked after each top/bottom pass and after each left/right pass.)  A query
    // whose search ends inside the block with at most kFNear candidates takes the near path: its record is the set
    // of candidate cells.  The others are deferred to the general termination scan (phase A2).
    const int qn = s.qn;
    for (int qbase = 0; qbase < qn; qbase += kFThreads) {
        const int q = qbase + tid;
        const bool live = q < qn;
        const int k = live ? queue[q] : 0;
        const int lj = k / kFW, li = k % kFW;
        uint32_t rec = 0xffffffffu;                                // "no record"
        int near_bin = -1;
        if (!live) {
        } else if (isnan(s.x[li]) || isnan(s.y[lj])) {
            put<kPatch>(out_tile + lj * tile_ld + li, static_cast<T>(qnan()));   // query out of bounds
        } else {
            const int ci = s.cx[li] - c0, cj = s.cy[lj] - r0;       // search centre in tile coordinates
            if ((ci < 11) | (ci > kFBW - 12) | (cj < 11) | (cj > kFBH - 12)) to_literal(k);   // never for node queries
            else {
                const int sh0 = ci - 2, wi = sh0 >> 5, sh = sh0 & 31;
                uint32_t blk = 0;
#pragma unroll
                for (int dy = -2; dy <= 2; ++dy) {
                    const uint32_t* mrow = s.mask + (cj + dy) * 4 + wi;
                    blk |= s.lut[(dy + 2) * 32 + (__funnelshift_r(mrow[0], mrow[1], sh) & 31u)];
                }
                const int n1 = __popc(blk & kC1), n2 = __popc(blk & kC2), n3 = __popc(blk & kC3), n4 = __popc(blk);
                const uint32_t cm = n1 >= 4 ? kC1 : (n2 >= 4 ? kC2 : (n3 >= 4 ? kC3 : kC4));
                const int n = n1 >= 4 ? n1 : (n2 >= 4 ? n2 : (n3 >= 4 ? n3 : n4));
                if (n >= 4 && n <= kFNear) {
                    near_bin = kFBinNear + n - 4;
                    rec = (blk & cm) | (static_cast<uint32_t>(near_bin) << 25);
                } else {
                    defer[atomicAdd(&s.dn, 1)] = static_cast<uint16_t>(q);
                }
            }
        }
        if (near_bin >= 0) atomicAdd(&s.hist[near_bin], 1);
        if (live) { s.rec[q] = rec; s.reck[q] = static_cast<uint16_t>(k); }
    }
    __syncthreads();
    // ---- phase A2: general termination scan (rings up to radius 10) for the deferred queries ---------------
    const int dn = s.dn;
    for (int t = tid; t < dn; t += kFThreads) {
        const int q = defer[t];
        const int k = s.reck[q];
        const int lj = k / kFW, li = k % kFW;
        const int ci = s.cx[li] - c0, cj = s.cy[lj] - r0;
        const int sh0 = ci - 10, wi = sh0 >> 5, sh = sh0 & 31;
        uint32_t wt[kFUnroll + 1], wb[kFUnroll + 1];
        const uint32_t w0 = window(cj, wi, sh);
        wt[0] = w0; wb[0] = w0;
        int n = (w0 >> 10) & 1;
        int r_end = kMaxRadius, lr_end = 1;                         // last ring visited; did its left/right pass run?
        bool done = false;
#pragma unroll
        for (int r = 1; r <= kFUnroll; ++r) {
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
        for (int r = kFUnroll + 1; r <= kMaxRadius && !done; ++r) {   // sparse neighbourhoods: rolled, windows re-read
            const uint32_t tbm = ((2u << (2 * r)) - 1u) << (10 - r);
            n += __popc(window(cj - r, wi, sh) & tbm) + __popc(window(cj + r, wi, sh) & tbm);
            if (n >= 4) { done = true; r_end = r; lr_end = 0; }
            else {
                const uint32_t lrm = (1u << (10 - r)) | (1u << (10 + r));
                for (int dy = -r + 1; dy <= r - 1; ++dy) n += __popc(window(cj + dy, wi, sh) & lrm);
                if (n >= 4) { done = true; r_end = r; lr_end = 1; }
            }
        }
        if (n > kFNMax) to_literal(k);
        else {
            // rings 1..5 each get bins of their own (round 2: with rings >= 4 lumped together, a warp of a 90 % masked tile mixed
            // searches of 4, 5 and 6 rings: -4.5 % at 90 %, -11 % at 97 %, +0.5 % at 70 %: profiles/r02_fill_bins_ab.txt)
            const int tcode = r_end <= 5 ? (r_end - 1) * 2 + lr_end : 10;
            const int ncls = n < 4 ? 9 : n - 4;                     // n in 4..12 -> 0..8
            const int bin = tcode * 10 + ncls;
            atomicAdd(&s.hist[bin], 1);
            s.rec[q] = static_cast<uint32_t>(r_end) | (static_cast<uint32_t>(lr_end) << 4) |
                       (static_cast<uint32_t>(n) << 5) | (static_cast<uint32_t>(bin) << 25);
        }
    }
    __syncthreads();
    // ---- order the records by bin: the warps of phase B then run the same code on lists of the same length ----
    if (warp == 0) {                                               // exclusive scan of the histogram
        int carry = 0;
#pragma unroll
        for (int b0 = 0; b0 < kFBins; b0 += 32) {
            const int b = b0 + lane;
            const int v = s.hist[b];
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            s.hist[b] = carry + incl - v;
            if (b == kFBinNear) s.q_near = carry + incl - v;        // first near-path record
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) s.qn = carry;                               // records that go to phase B
    }
    __syncthreads();
    {   // scatter: records move from queue order to bin order (read all, sync, then write)
        uint32_t mine[kFPer];
        uint16_t minek[kFPer];
#pragma unroll
        for (int it = 0; it < kFPer; ++it) {
            const int q = it * kFThreads + tid;
            mine[it] = q < qn ? s.rec[q] : 0xffffffffu;
            minek[it] = q < qn ? s.reck[q] : 0;
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < kFPer; ++it) {
            if (mine[it] != 0xffffffffu) {
                const int dst = atomicAdd(&s.hist[mine[it] >> 25], 1);
                s.rec[dst] = mine[it]; s.reck[dst] = minek[it];
            }
        }
    }
    __syncthreads();
    const int qb = s.qn, q_near = s.q_near;

    // ---- phase B, general path (as a function of the record): searches that leave the 5 x 5 block or hold more than
    // kFNear candidates.  Candidates in list order into a per-thread list in shared memory, the reference's selection
    // literally on it (squared distances, with the sqrt guard of each comparison), then the method.
    auto general_query = [&](int q) {
        const uint32_t rec = s.rec[q];
        const int k = s.reck[q];
        const int r_end = rec & 15, lr_end = (rec >> 4) & 1, n = (rec >> 5) & 31;
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
        uint32_t wt[kFUnroll + 1], wb[kFUnroll + 1];
        const uint32_t w0 = window(cj, wi, sh);
        wt[0] = w0; wb[0] = w0;
        if ((w0 >> 10) & 1) push(0, 0, dadd(sqdx(0), sqdy(0)));
#pragma unroll
        for (int r = 1; r <= kFUnroll; ++r) {
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
        for (int r = kFUnroll + 1; r <= r_end; ++r) {               // sparse neighbourhoods: the same walk, rolled
            const uint32_t tbm = ((2u << (2 * r)) - 1u) << (10 - r);
            const uint32_t top = window(cj - r, wi, sh) & tbm, bot = window(cj + r, wi, sh) & tbm;
            uint32_t m = top | bot;
            if (m) {
                const double dyt = sqdy(-r), dyb = sqdy(r);
                while (m) {
                    const int b = __ffs(m) - 1;
                    m &= m - 1;
                    const double dx2 = sqdx(b - 10);
                    if ((top >> b) & 1u) push(b - 10, -r, dadd(dx2, dyt));
                    if ((bot >> b) & 1u) push(b - 10, r, dadd(dx2, dyb));
                }
            }
            if (r < r_end || lr_end) {
                const uint32_t lb = 1u << (10 - r), rb = 1u << (10 + r);
                double dxl = 0.0, dxr = 0.0;
                bool have = false;
                for (int dy = -r + 1; dy <= r - 1; ++dy) {
                    const uint32_t wr = window(cj + dy, wi, sh);
                    if (wr & (lb | rb)) {
                        if (!have) { dxl = sqdx(-r); dxr = sqdx(r); have = true; }
                        const double dy2 = sqdy(dy);
                        if (wr & lb) push(-r, dy, dadd(dxl, dy2));
                        if (wr & rb) push(r, dy, dadd(dxr, dy2));
                    }
                }
            }
        }
        // cnt == n by construction (phase A counted the same bits)
        // -- the reference's partial selection sort with swaps (GridH.cpp:123-140) on squared distances.
        // NN only needs the first pick: pass 0 (with fewer than four candidates the oracle's "first strict
        // minimum" is the same scan); the other methods run the four passes when four candidates exist.
        bool unsure = (cnt != n);
        const int n_pass = METHOD == NN ? (cnt > 0 ? 1 : 0) : (cnt >= 4 ? 4 : 0);
        for (int m = 0; m < n_pass; ++m) {
            int best = m;
            const double dm = ld2[m * kFThreads];
            double dbest = dm, thr = dmul(dm, kSafeRatio);
            for (int kk = m + 1; kk < cnt; ++kk) {
                const double dk = ld2[kk * kFThreads];
                if (dk < thr) { best = kk; dbest = dk; thr = dmul(dk, kSafeRatio); }
                else if (dk < dbest) unsure = true;                // within a few ulps: sqrt may tie
            }
            if (best != m) {
                ld2[m * kFThreads] = dbest; ld2[best * kFThreads] = dm;
                const uint16_t cm = lcode[m * kFThreads];
                lcode[m * kFThreads] = lcode[best * kFThreads]; lcode[best * kFThreads] = cm;
            }
        }
        if (unsure) { to_literal(k); return; }
        if (cnt >= 4) {
            T v[4];
            double d2v[4];
            int pi[4], pj[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int code = lcode[e * kFThreads];
                const int dx = (code & 31) - 10, dy = (code >> 5) - 10;
                v[e] = s.tile[(cj + dy) * kFBW + ci + dx];
                d2v[e] = ld2[e * kFThreads];
                pi[e] = cig + dx; pj[e] = cjg + dy;
            }
            put<kPatch>(out_tile + lj * tile_ld + li, finish_four<T, METHOD>(p, v, d2v, pi, pj, I0 + li, J0 + lj));
        } else {                                                    // the search ran out of rings (GridH.cpp:291-298)
            double fv[3], fd[3];
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                const int code = e < cnt ? lcode[e * kFThreads] : (10 * 32 + 10);
                fv[e] = static_cast<double>(s.tile[(cj + (code >> 5) - 10) * kFBW + ci + (code & 31) - 10]);
                fd[e] = e < cnt ? ld2[e * kFThreads] : 0.0;
            }
            put<kPatch>(out_tile + lj * tile_ld + li, finish_few<T, METHOD>(cnt, fv[0], fv[1], fv[2], fd[0], fd[1], fd[2]));
        }
    };

    // ---- phase B: warps draw chunks of 32 records from a CTA-wide counter ------------------------------------------------
    // General-path chunks first (they are the long ones), then the near-path chunks; the bins are ordered by candidate
    // count, so a near chunk's queries have (nearly) the same count and the list length is a compile-time constant
    // chosen per chunk.  Drawing chunks keeps every warp busy until the records run out whatever the mix.
    {
        const int gen_chunks = (q_near + 31) >> 5, all_chunks = gen_chunks + ((qb - q_near + 31) >> 5);
        for (;;) {
            int ch = 0;
            if (lane == 0) ch = atomicAdd(&s.next, 1);
            ch = __shfl_sync(0xffffffffu, ch, 0);
            if (ch >= all_chunks) break;
            if (ch < gen_chunks) {
                const int q = ch * 32 + lane;
                if (q < q_near) general_query(q);
                __syncwarp();
                continue;
            }
            const int q = q_near + (ch - gen_chunks) * 32 + lane;
            const bool active = q < qb;
            const uint32_t cand = active ? (s.rec[q] & kC4) : 0u;
            const int nmax = __reduce_max_sync(0xffffffffu, __popc(cand));
            if (active) {
                const int k = s.reck[q];
                bool ok;
                if (nmax <= 4) ok = near_query<T, METHOD, 4, false, kPatch>(s, p, cand, q, k, c0, r0, I0, J0, out_tile, tile_ld);
                else if (nmax <= 6) ok = near_query<T, METHOD, 6, false, kPatch>(s, p, cand, q, k, c0, r0, I0, J0, out_tile, tile_ld);
                else ok = near_query<T, METHOD, kFNear, false, kPatch>(s, p, cand, q, k, c0, r0, I0, J0, out_tile, tile_ld);
                if (!ok) {
                    const int slot = atomicAdd(&s.tn, 1);
                    if (slot < kFReplayMax) s.replay[slot] = static_cast<uint16_t>(q);
                    else to_literal(k);
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- replay of the near-path queries that met a near tie: the same selection on square-rooted distances ----------
    const int tn = min(s.tn, kFReplayMax);
    for (int t = tid; t < tn; t += kFThreads) {
        const int q = s.replay[t];
        near_query<T, METHOD, kFNear, true, kPatch>(s, p, s.rec[q] & kC4, q, s.reck[q], c0, r0, I0, J0, out_tile, tile_ld);
    }
    __syncthreads();
    // ---- kriging of the near-path picks: full warps, nothing else live ------------------------------------------------
    if (METHOD == KRIGING) {
        for (int q = q_near + tid; q < qb; q += kFThreads) {
            const uint32_t rec = s.rec[q];
            if ((rec >> 25) != 127u) continue;                      // went to the literal path
            const int k = s.reck[q];
            const int lj = k / kFW, li = k % kFW;
            const int cig = s.cx[li], cjg = s.cy[lj];
            const T* const centre = s.tile + (cjg - r0) * kFBW + (cig - c0);
            Picked pk;
            pk.found = 4;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = (rec >> (5 * e)) & 31;
                const int dxy = s.cell_dxy[c];
                pk.i[e] = cig + static_cast<int16_t>(dxy & 0xffff); pk.j[e] = cjg + (dxy >> 16);
                pk.v[e] = static_cast<double>(centre[s.cell_tile[c]]);
                pk.d[e] = 0.0;
            }
            put<kPatch>(out_tile + lj * tile_ld + li,
                   static_cast<T>(kriging_from_picked(p.g, pk, __ldg(p.lon.coord + I0 + li), __ldg(p.lat.coord + J0 + lj))));
        }
    }
    // ---- queries the bitmask paths handed back: literal per-query evaluation ---------------------------------
    const int rn = min(s.rn, kFRedoMax);
    for (int q = tid; q < rn; q += kFThreads) {
        const int k = s.redo[q];
        put<kPatch>(out_tile + (k / kFW) * tile_ld + (k % kFW), fill_cell_literal<T, METHOD>(&p, J0 + k / kFW, I0 + k % kFW));
    }
    // ---- PATCH: the finished tile leaves -- pass-through cells and patched results in one sweep of 16-byte row stores ------
    if (kPatch) {
        __syncthreads();
        constexpr int VEC = 16 / static_cast<int>(sizeof(T)), VPR = kFW / VEC;
        const int nrows = static_cast<int>(min(static_cast<int64_t>(kFH), p.row_end - J0));
        const int ncols = min(kFW, p.n_out_cols - I0);
        if (p.vec_ok) {
            for (int k = tid; k < nrows * VPR; k += kFThreads) {
                const int lj = k / VPR, col = (k - lj * VPR) * VEC;
                const T* const src = out_tile + lj * kFBW + col;
                T* const dst = g_out_tile + lj * p.out_ld + col;
                if (col + VEC <= ncols) {
                    if constexpr (sizeof(T) == 4) __stcs(reinterpret_cast<float4*>(dst), *reinterpret_cast<const float4*>(src));
                    else __stcs(reinterpret_cast<double2*>(dst), *reinterpret_cast<const double2*>(src));
                } else {
                    for (int c = 0; col + c < ncols; ++c) __stcs(dst + c, src[c]);
                }
            }
        } else {
            for (int k = tid; k < nrows * kFW; k += kFThreads) {
                const int lj = k / kFW, li = k - lj * kFW;
                if (li < ncols) __stcs(g_out_tile + lj * p.out_ld + li, out_tile[lj * kFBW + li]);
            }
        }
    }
}

// ---- bilinear gap fill: no search, no queue -- a streaming kernel ------------------------------------------------------
// Every masked cell reads its four corners (GridH.cpp:160-210: NaN-corner mean, else the three lerps in the reference's
// order); valid cells pass through.  Persistent CTAs, 128 x 32 tiles + a 4 x 2 halo (floor() of a node query may come
// out one below the node: SURVEY.md section 0 fact 3) staged by TMA into a two-deep ring, the next tile in flight while
// this one is evaluated; a thread owns 16 bytes of a tile row and emits them with one streaming store.  HBM-bound:
// sizeof(T) read + sizeof(T) written per cell.
constexpr int kBW = 128, kBH = 32;             // tile of cells
constexpr int kBHaloC = 4, kBHaloR = 2;
constexpr int kBBW = kBW + 2 * kBHaloC;        // 136
constexpr int kBBH = kBH + 2 * kBHaloR;        // 36
constexpr int kBThreads = 256;

// mean_valid4 (exact.cuh, GridH.cpp:10-18) without branches, for the streaming kernel below where every lane takes it with
// its own count.  The sum adds +0.0 in place of a missing corner: s + (+0.0) == s for every s but -0.0, and s is never -0.0
// here (it starts at +0.0, and +0.0 + (-0.0) = +0.0), so the bits are those of the reference's "skip the NaNs" loop.  The
// division by the count is the product with 1 / n from a five-entry table -- exact for n = 1, 2, 4; for n = 3 the same
// Markstein step as div_count (r = s - 3q exact in an FMA, q' = RN(q + r * RN(1/3)) is the correctly rounded quotient),
// which leaves q alone when r = 0; n = 0 multiplies by NaN.  Same bits as mean_valid4 (A/B: sha1 of whole filled grids).
__constant__ double kInvCount[5] = {__builtin_nan(""), 1.0, 0.5, 0.33333333333333331, 0.25};
__constant__ double kCount[5] = {0.0, 1.0, 2.0, 3.0, 4.0};
__device__ __noinline__ double ddiv_cold(double a, double b) { return ddiv(a, b); }   // out of line: never speculated into the stream
template <typename T>
__device__ __forceinline__ double mean_valid4_flat(T a, T b, T c, T d, int& n_valid) {
    const bool va = a == a, vb = b == b, vc = c == c, vd = d == d;
    const int n = static_cast<int>(va) + static_cast<int>(vb) + static_cast<int>(vc) + static_cast<int>(vd);
    const T zero = static_cast<T>(0);
    double s = dadd(0.0, static_cast<double>(va ? a : zero));
    s = dadd(s, static_cast<double>(vb ? b : zero));
    s = dadd(s, static_cast<double>(vc ? c : zero));
    s = dadd(s, static_cast<double>(vd ? d : zero));
    n_valid = n;
    const double inv = kInvCount[n], cnt = kCount[n];
    double q = dmul(s, inv);
    const double r = __fma_rn(-cnt, q, s);
    q = __fma_rn(r, inv, q);
    const double mag = fabs(s);
    if (!(mag > 1e-280 && mag < 1e300) && s != 0.0) q = ddiv_cold(s, cnt);   // where q or r could go subnormal / overflow: the division itself
    return q;
}

constexpr int kNoCell = -(1 << 30);            // x0 table entry of a query whose position is NaN (out of bounds)

// The three lerps of GridH.cpp:200-209 in the reference's order.  Out of line: a gap fill never gets here (see the kernel).
__device__ __noinline__ double bilinear_lerp_cold(double a, double b, double c, double d, double x, double y, int x0, int y0) {
    const double wx = dsub(x, __int2double_rn(x0)), wy = dsub(y, __int2double_rn(y0));
    const double ux = dsub(1.0, wx);
    const double lo = dadd(dmul(ux, a), dmul(wx, b));
    const double hi = dadd(dmul(ux, c), dmul(wx, d));
    return dadd(dmul(dsub(1.0, wy), lo), dmul(wy, hi));
}

template <typename T>
struct BilinSmem {
    alignas(128) T tile[2][kBBH * kBBW];
    double x[kBW], y[kBH];
    int x0[kBW], y0[kBH];
    uint64_t bar[2];
};

template <typename T>
__global__ void __launch_bounds__(kBThreads, sizeof(T) == 4 ? 4 : 2)
bilinear_fill_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ FillParams<T> p) {
    constexpr int VEC = 16 / static_cast<int>(sizeof(T)), VPR = kBW / VEC;
    constexpr uint32_t kBoxBytes = kBBW * kBBH * sizeof(T);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BilinSmem<T>& s = *reinterpret_cast<BilinSmem<T>*>(smem_raw);
    const int tid = threadIdx.x;
    if (p.use_tma && tid == 0) { prefetch_tmap(&tmap); mbar_init(&s.bar[0], 1); mbar_init(&s.bar[1], 1); }
    __syncthreads();
    auto request_tile = [&](int t, int b) {
        const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
        mbar_expect_tx(&s.bar[b], kBoxBytes);
        tma_load_2d(s.tile[b], &tmap, tx * kBW - kBHaloC, static_cast<int>(p.row_begin) + ty * kBH - kBHaloR - p.g.row0, &s.bar[b]);
    };
    int tile = blockIdx.x;
    if (p.use_tma && tid == 0 && tile < p.n_tiles) request_tile(tile, 0);
    for (int it = 0; tile < p.n_tiles; ++it, tile += gridDim.x) {
        const int buf = it & 1;
        T* const tl = s.tile[buf];
        const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
        const int I0 = tx * kBW;
        const int64_t J0 = p.row_begin + static_cast<int64_t>(ty) * kBH;
        const int c0 = I0 - kBHaloC, r0 = static_cast<int>(J0) - kBHaloR;
        if (p.use_tma) {
            const int nxt = tile + gridDim.x;
            if (tid == 0 && nxt < p.n_tiles) request_tile(nxt, buf ^ 1);
        } else {
            for (int k = tid; k < kBBW * kBBH; k += kBThreads) {
                const int lr = k / kBBW, lc = k - lr * kBBW;
                const int gr = r0 + lr, gc = c0 + lc;
                T v = static_cast<T>(qnan());
                if (gr >= p.rows_resident_lo && gr < p.rows_resident_hi && gc >= 0 && gc < p.g.n_lon)
                    v = __ldg(p.g.z + static_cast<int64_t>(gr - p.g.row0) * p.g.ld + gc);
                tl[k] = v;
            }
        }
        if (tid < kBW) {
            const int I = I0 + tid;
            const bool in = I < p.n_out_cols;
            const double x = in ? __ldg(p.lon.pos + I) : qnan();
            s.x[tid] = x;
            s.x0[tid] = isnan(x) ? kNoCell : __ldg(p.lon.base + I);
        } else if (tid < kBW + kBH) {
            const int t = tid - kBW;
            const bool in = J0 + t < p.row_end;
            s.y[t] = in ? __ldg(p.lat.pos + J0 + t) : qnan();
            s.y0[t] = in ? __ldg(p.lat.base + J0 + t) : 0;
        }
        __syncthreads();                                            // tables (and the plain-load tile) visible
        if (p.use_tma) mbar_wait(&s.bar[buf], (it >> 1) & 1);

        const int nrows = static_cast<int>(min(static_cast<int64_t>(kBH), p.row_end - J0));
        const int ncols = min(kBW, p.n_out_cols - I0);
        T* const out_tile = p.out + (J0 - p.row_begin) * p.out_ld + I0;
        // A thread keeps its column group for the whole tile (kBThreads is a multiple of the vectors per row): the corner
        // columns of its cells -- offsets into a tile row, kNoCell where the query is out of bounds -- are formed once.
        static_assert(kBThreads % VPR == 0, "a thread's column group is fixed");
        const int col = (tid % VPR) * VEC;
        int xo0[VEC], xo1[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
            const int x0 = col + c < kBW ? s.x0[col + c] : kNoCell;
            xo0[c] = x0 == kNoCell ? kNoCell : x0 - c0;
            xo1[c] = min(x0 + 1, p.g.n_lon - 1) - c0;
        }
        for (int lj = tid / VPR; lj < nrows && col < ncols; lj += kBThreads / VPR) {
            const T* const src = tl + (lj + kBHaloR) * kBBW + kBHaloC + col;
            T v[VEC];
            if constexpr (sizeof(T) == 4) { const float4 q = *reinterpret_cast<const float4*>(src); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
            else { const double2 q = *reinterpret_cast<const double2*>(src); v[0] = q.x; v[1] = q.y; }
            const double y = s.y[lj];
            const int y0 = s.y0[lj], y1 = min(y0 + 1, p.g.n_lat - 1);
            const T* const r0p = tl + (y0 - r0) * kBBW;
            const T* const r1p = tl + (y1 - r0) * kBBW;
            const bool y_ok = !isnan(y);
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
                if (v[c] == v[c]) continue;                         // valid cell: passes through
                double result = qnan();
                if (y_ok && xo0[c] != kNoCell) {
                    const T a = r0p[xo0[c]], b = r0p[xo1[c]], cc = r1p[xo0[c]], d = r1p[xo1[c]];
                    // A masked cell queried at its own node is one of its four corners (floor() of the node position is the
                    // node or, by FP64 noise, the one below: SURVEY.md section 0 fact 3), so this is the NaN-corner mean
                    // (GridH.cpp:186-198) for every query of a gap fill; the lerp stays reachable, out of line.
                    int n_valid;
                    result = mean_valid4_flat<T>(a, b, cc, d, n_valid);
                    if (n_valid == 4)
                        result = bilinear_lerp_cold(static_cast<double>(a), static_cast<double>(b), static_cast<double>(cc), static_cast<double>(d),
                                                    s.x[col + c], y, xo0[c] + c0, y0);
                }
                v[c] = static_cast<T>(result);
            }
            T* const dst = out_tile + lj * p.out_ld + col;
            if (p.vec_ok && col + VEC <= ncols) {
                if constexpr (sizeof(T) == 4) __stcs(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
                else __stcs(reinterpret_cast<double2*>(dst), make_double2(v[0], v[1]));
            } else {
#pragma unroll
                for (int c = 0; c < VEC; ++c) if (col + c < ncols) __stcs(dst + c, v[c]);
            }
        }
        if (p.use_tma) fence_proxy_async_smem();
        __syncthreads();
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
// CTAs of `kern` that fit one SM times the SMs of the current device: the size of a persistent grid.
template <typename K>
static int resident_ctas(K kern, int threads, size_t smem) {
    int dev = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess) { cudaGetLastError(); return 148; }
    return sms * (per_sm < 1 ? 1 : per_sm);
}

template <typename T>
static cudaError_t fill_params(const GridDesc& d, const AxisTables& lat, const AxisTables& lon, int64_t row_begin, int64_t row_end,
                               void* out, int64_t out_ld, bool fill, int tile_w, int tile_h, FillParams<T>* pp) {
    FillParams<T>& p = *pp;
    p.g = make_view<T>(d);
    p.lat = FillAxis{lat.coord, lat.pos, lat.base};
    p.lon = FillAxis{lon.coord, lon.pos, lon.base};
    p.row_begin = row_begin; p.row_end = row_end;
    p.rows_resident_lo = d.row0; p.rows_resident_hi = d.row0 + d.rows;
    p.out = static_cast<T*>(out); p.out_ld = out_ld;
    p.n_out_cols = lon.n;
    p.f_lat = d.n_lat > 1 ? (lat.n - 1) / (d.n_lat - 1) : 1;
    p.f_lon = d.n_lon > 1 ? (lon.n - 1) / (d.n_lon - 1) : 1;
    if (p.f_lat < 1 || p.f_lon < 1 || (fill && (p.f_lat != 1 || p.f_lon != 1))) return cudaErrorInvalidValue;
    const int64_t tiles_x = (lon.n + tile_w - 1) / tile_w, tiles_y = (row_end - row_begin + tile_h - 1) / tile_h;
    if (tiles_x * tiles_y > (1ll << 30)) return cudaErrorInvalidValue;
    p.tiles_x = static_cast<int>(tiles_x); p.n_tiles = static_cast<int>(tiles_x * tiles_y);
    p.vec_ok = (reinterpret_cast<uintptr_t>(out) % 16 == 0 && (out_ld * sizeof(T)) % 16 == 0) ? 1 : 0;
    return cudaSuccess;
}

template <typename T>
static cudaError_t launch_bilinear_fill(const GridDesc& d, const AxisTables& lat, const AxisTables& lon, int64_t row_begin,
                                        int64_t row_end, void* out, int64_t out_ld, cudaStream_t st, LaunchInfo* info) {
    FillParams<T> p;
    cudaError_t e = fill_params<T>(d, lat, lon, row_begin, row_end, out, out_ld, true, kBW, kBH, &p);
    if (e != cudaSuccess) return e;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    p.use_tma = make_grid_tensor_map(d, kBBW, kBBH, &tmap, /*nan_fill=*/true) ? 1 : 0;
    auto kern = bilinear_fill_kernel<T>;
    const size_t smem = sizeof(BilinSmem<T>);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    static const int slots = resident_ctas(kern, kBThreads, smem);
    const int grid = p.n_tiles < slots ? p.n_tiles : slots;
    kern<<<grid, kBThreads, smem, st>>>(tmap, p);
    if (info) { info->launches += 1; info->used_tma = p.use_tma; }
    return cudaGetLastError();
}

template <typename T, int METHOD, bool FILL>
static cudaError_t launch_fill_t(const GridDesc& d, const AxisTables& lat, const AxisTables& lon, int64_t row_begin,
                                 int64_t row_end, void* out, int64_t out_ld, cudaStream_t st, LaunchInfo* info) {
    // rows within the ring-search reach of the block must be resident (slab + halo)
    int need_lo = lat.h_base[row_begin] - (kMaxRadius + 2), need_hi = lat.h_base[row_end - 1] + 1 + (kMaxRadius + 2);
    if (METHOD == BILINEAR) { need_lo = lat.h_base[row_begin]; need_hi = lat.h_base[row_end - 1] + 1; }
    need_lo = need_lo < 0 ? 0 : need_lo;
    need_hi = need_hi > d.n_lat - 1 ? d.n_lat - 1 : need_hi;
    if (need_lo < d.row0 || need_hi >= d.row0 + d.rows) return cudaErrorInvalidValue;
    if constexpr (METHOD == BILINEAR && FILL) {
        return launch_bilinear_fill<T>(d, lat, lon, row_begin, row_end, out, out_ld, st, info);
    } else {
        FillParams<T> p;
        p.g = make_view<T>(d);
        p.lat = FillAxis{lat.coord, lat.pos, lat.base};
        p.lon = FillAxis{lon.coord, lon.pos, lon.base};
        p.row_begin = row_begin; p.row_end = row_end;
        p.rows_resident_lo = d.row0; p.rows_resident_hi = d.row0 + d.rows;
        p.out = static_cast<T*>(out); p.out_ld = out_ld;
        p.n_out_cols = lon.n;
        p.f_lat = d.n_lat > 1 ? (lat.n - 1) / (d.n_lat - 1) : 1;
        p.f_lon = d.n_lon > 1 ? (lon.n - 1) / (d.n_lon - 1) : 1;
        p.tiles_x = 0; p.n_tiles = 0;
        p.vec_ok = (reinterpret_cast<uintptr_t>(out) % 16 == 0 && (out_ld * sizeof(T)) % 16 == 0) ? 1 : 0;
        if (p.f_lat < 1 || p.f_lon < 1 || (FILL && (p.f_lat != 1 || p.f_lon != 1))) return cudaErrorInvalidValue;
        CUtensorMap tmap;
        memset(&tmap, 0, sizeof tmap);
        p.use_tma = make_grid_tensor_map(d, kFBW, kFBH, &tmap, /*nan_fill=*/true) ? 1 : 0;
        auto kern = fill_tiled_kernel<T, METHOD, FILL>;
        if constexpr (FILL && METHOD != BILINEAR_SEARCH) {
            // Results patched into the staged tile + 16-byte row stores, or one streaming store per query?  Into local HBM the
            // two are within 1 % for IDW and the 4-nearest mean, row stores win 2-4 % for kriging and lose 3.5 % for NN
            // (profiles/r02_fill_patch_ab.txt).  Into PEER memory (the gather fused into the kernel)
            // per-query 4-byte stores crawl over NVLink: row stores for every method.  AUVI_FILL_PATCH=0/1 forces (A/B, tests).
            static const int force = [] { const char* e = getenv("AUVI_FILL_PATCH"); return e ? atoi(e) : -1; }();
            const bool patch = force >= 0 ? force != 0 : (METHOD != NN || output_is_peer_memory(out));
            if (patch) kern = fill_tiled_kernel<T, METHOD, FILL, true>;
        }
        const size_t smem = sizeof(FillSmem<T>);
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        dim3 grid(static_cast<unsigned>((lon.n + kFW - 1) / kFW), static_cast<unsigned>((row_end - row_begin + kFH - 1) / kFH));
        kern<<<grid, kFThreads, smem, st>>>(tmap, p);
        if (info) { info->launches += 1; info->used_tma = p.use_tma; }
        return cudaGetLastError();
    }
}

cudaError_t launch_fill(const GridDesc& d, int method, const AxisTables& lat, const AxisTables& lon, int64_t row_begin,
                        int64_t row_end, void* out, int64_t out_ld, int fill, cudaStream_t st, LaunchInfo* info) {
    if (row_end <= row_begin) return cudaSuccess;
#define AUVI_CASE(T, M, F) \
    case M: return launch_fill_t<T, M, F>(d, lat, lon, row_begin, row_end, out, out_ld, st, info);
    if (fill) {
        if (d.dtype == DT_F64) {
            switch (method) { AUVI_CASE(double, BILINEAR, true) AUVI_CASE(double, CUBIC, true) AUVI_CASE(double, KRIGING, true)
                              AUVI_CASE(double, NN, true) AUVI_CASE(double, IDW, true) AUVI_CASE(double, BILINEAR_SEARCH, true) }
        } else {
            switch (method) { AUVI_CASE(float, BILINEAR, true) AUVI_CASE(float, CUBIC, true) AUVI_CASE(float, KRIGING, true)
                              AUVI_CASE(float, NN, true) AUVI_CASE(float, IDW, true) AUVI_CASE(float, BILINEAR_SEARCH, true) }
        }
    } else {                                   // upsampling lattice, the search-based methods (every cell is a query)
        if (d.dtype == DT_F64) {
            switch (method) { AUVI_CASE(double, KRIGING, false) AUVI_CASE(double, NN, false) AUVI_CASE(double, IDW, false) }
        } else {
            switch (method) { AUVI_CASE(float, KRIGING, false) AUVI_CASE(float, NN, false) AUVI_CASE(float, IDW, false) }
        }
    }
#undef AUVI_CASE
    return cudaErrorInvalidValue;
}

}  // namespace auvi
