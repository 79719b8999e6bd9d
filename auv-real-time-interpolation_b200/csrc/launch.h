// launch.h -- host-side descriptors and the launch entry points each kernel file exports to api.cu.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "exact.cuh"

namespace auvi {

enum DType : int { DT_F64 = 0, DT_F32 = 1 };

// How a lattice axis turns an integer index into a coordinate (both are reference driver code):
//   AXIS_EXPANDED: c = lo + k*(hi-lo)/(new_n-1)          test_interpolation.cpp:99-106 (Grid A)
//   AXIS_NODES:    c = lo + k*step, step=(hi-lo)/(n-1)   test_gebco.cpp:72-81          (Grid B)
enum AxisKind : int { AXIS_EXPANDED = 0, AXIS_NODES = 1 };

// Device-resident grid (or a row slab of it) as the C-ABI holds it.
struct GridDesc {
    const void* z = nullptr;       // device pointer, row-major, row 0 = min_lat
    int dtype = DT_F64;
    int n_lat = 0, n_lon = 0;      // GLOBAL dimensions
    int64_t ld = 0;                // elements between consecutive rows (>= n_lon)
    int row0 = 0, rows = 0;        // slab held in z: global rows [row0, row0+rows)
    double min_lon = 0, max_lon = 0, min_lat = 0, max_lat = 0;
    double lon_step = 0, lat_step = 0;
    double vg_nugget = 1.0, vg_sill = 100.0, vg_inv_range = 0.1, vg_diag = 1.0;   // what the kriging kernels use (GridH.cpp:371-376)
};

template <typename T>
inline GridView<T> make_view(const GridDesc& d) {
    GridView<T> v;
    v.z = static_cast<const T*>(d.z);
    v.n_lat = d.n_lat; v.n_lon = d.n_lon; v.ld = d.ld; v.row0 = d.row0;
    v.min_lon = d.min_lon; v.max_lon = d.max_lon; v.min_lat = d.min_lat; v.max_lat = d.max_lat;
    v.lon_step = d.lon_step; v.lat_step = d.lat_step;
    v.vg_nugget = d.vg_nugget; v.vg_sill = d.vg_sill; v.vg_inv_range = d.vg_inv_range; v.vg_diag = d.vg_diag;
    return v;
}

// Per-axis query tables of a separable lattice, one entry per OUTPUT index.  Built once on the host
// with the reference drivers' exact expressions (api.cu), so every floor/round/tie decision
// downstream sees the same last-bit FP64 noise the reference CPU path sees.
struct AxisTables {
    const double* coord = nullptr; // device: raw lon (or lat) of output index k
    const double* pos = nullptr;   // device: index-space image (c-min)/step; NaN when out of bounds
    const int* base = nullptr;     // device: floor(pos); for out-of-bounds entries the nearest valid one
    const int* h_base = nullptr;   // host copy of base (tile sizing)
    int n = 0;                     // number of output indices
    int* window_mode_cache = nullptr;  // host, optional: -1 = not yet checked, else upsample.cu's window_mode() verdict for this axis
};

// points.cu
cudaError_t launch_points(const GridDesc& d, int method, const double* pts, int64_t stride_dbl, int64_t n,
                          double* out, int32_t* sel, int32_t* found, cudaStream_t st);

// What a lattice launch did (diagnostics surfaced through the C-ABI).
struct LaunchInfo {
    int launches = 0;              // kernels launched
    int used_tma = 0;              // 1 if the input tiles were staged by TMA
    int used_window = 0;           // window-load form of the FP32 bicubic kernel: 0 generic, 1 / 2 = longitude factor
};

// upsample.cu -- lattice mode: out[(J-row_begin)*out_ld + I] for J in [row_begin,row_end), all I.
// fill != 0: cells whose own grid value is valid are passed through (Grid-B gap fill on AXIS_NODES).
// sel (optional) is 9 int32 per output cell: {found, i0,j0,...,i3,j3}.
cudaError_t launch_lattice(const GridDesc& d, int method, const AxisTables& lat, const AxisTables& lon,
                           int64_t row_begin, int64_t row_end, void* out, int64_t out_ld,
                           int fill, int32_t* sel, cudaStream_t st, LaunchInfo* info);

// fill.cu -- the tiled ring-search kernel.  fill != 0: full-grid gap fill on the node lattice (AXIS_NODES, factor 1),
// BILINEAR, CUBIC (ring-search mean), KRIGING, NN, IDW.  fill == 0: KRIGING / NN / IDW on an upsampling lattice
// (integer factors), every output cell a query.
cudaError_t launch_fill(const GridDesc& d, int method, const AxisTables& lat, const AxisTables& lon, int64_t row_begin,
                        int64_t row_end, void* out, int64_t out_ld, int fill, cudaStream_t st, LaunchInfo* info);

// metrics.cu -- MAE / RMSE / Max / NaN count of est against truth (error_calculator.cpp:5-45).
// scratch: device buffer of metrics_scratch_bytes(); result5 (device): {sum|d|, sum d^2, max|d|, #NaN, #compared}.
size_t metrics_scratch_bytes();
cudaError_t launch_metrics(const void* truth, const void* est, int dtype, int64_t n, void* scratch,
                           double* result5, cudaStream_t st, LaunchInfo* info);
// The same over the cells that are NaN in `masked` (the cells a gap fill produced), 2-D with row pitches.
cudaError_t launch_metrics_masked(const void* masked, int64_t ld_m, const void* filled, int64_t ld_f, const void* truth,
                                  int64_t ld_t, int dtype, int64_t rows, int cols, void* scratch, double* result5,
                                  cudaStream_t st, LaunchInfo* info);

// metrics.cu -- empirical semivariances of the grid for the fitted variogram (opt-in AUVI_KRIGING_FITTED): for the lags
// 1, 2, 4, 8 cells along each axis, sums[16] = {sum of squared differences, pair count} x {lon lags, lat lags} over the
// pairs of valid cells; deterministic two-stage reduction.  scratch: variogram_scratch_bytes().
size_t variogram_scratch_bytes();
cudaError_t launch_variogram_sums(const GridDesc& d, void* scratch, double* sums16, cudaStream_t st, LaunchInfo* info);

// ingest.cu -- Grid-B data preparation on the device (decode a NetCDF variable, mask cells).
cudaError_t launch_decode_raw(const void* raw, int nc_type, int big_endian, int flip_rows, double scale, double offset,
                              int n_lat, int n_lon, void* out, int64_t ld, int dtype, cudaStream_t st);
cudaError_t launch_mask_cells(const GridDesc& d, const int64_t* idx, int64_t n, void* truth, cudaStream_t st);
cudaError_t launch_mask_hash(const GridDesc& d, double fraction, uint64_t seed, unsigned long long* n_masked, cudaStream_t st);
// CSV matrix text on the device: delimiter index (counts per 4 KiB block -> scan -> positions), then one thread per field.
size_t csv_block_count(int64_t n_bytes);
cudaError_t launch_csv_index(const char* text, int64_t n_bytes, int* counts, int64_t* offsets, int64_t* total, int64_t* pos,
                             int64_t max_fields, cudaStream_t st);
cudaError_t launch_csv_parse(const char* text, const int64_t* pos, int64_t n_fields, int n_cols, void* out, int64_t ld, int dtype,
                             unsigned* status, unsigned long long* n_slow, int64_t* slow_idx, int64_t slow_cap, cudaStream_t st);
cudaError_t launch_csv_patch(void* out, int64_t ld, int n_cols, int dtype, const int64_t* idx, const double* val, int64_t n,
                             cudaStream_t st);

// api.cu: does `p` lie in memory of ANOTHER device -- mapped by auvi_peer_open (CUDA IPC) or allocated on a peer device of this
// process?  The gap-fill kernel then leaves its tiles as 16-byte row stores instead of per-query stores.
bool output_is_peer_memory(const void* p);

// Encodes a 2-D tiled tensor map over the grid slab; returns false when TMA cannot address it.  nan_fill: box elements
// outside the slab arrive as NaN instead of zero (the gap-fill kernels treat "outside" and "masked" alike).
bool make_grid_tensor_map(const GridDesc& d, int box_w, int box_h, CUtensorMap* out, bool nan_fill = false);

}  // namespace auvi
