/* include/auvi.h -- the C ABI of libauvi.so, the B200-native replacement for the GPU half of
 * devsaxena974/AUV-Real-Time-Interpolation (code/src/GridD.cu + code/src/kernels.cu).
 *
 * Everything the reference's host code needs from the device goes through these entry points:
 * plain pointers and sizes, no C++/torch types.  The C++ class `GridD` with the reference's own
 * signatures (auv-real-time-interpolation_b200/host/GridD.{h,cpp}) is written on top of it;
 * INTEGRATION.md shows the binding a maintainer of the reference would add.  All functions return
 * 0 on success, non-zero on failure with a message available from auvi_last_error().  There is no
 * CPU fallback: without a CUDA device every compute call fails.
 *
 * Reference citations are file:line in /root/reference/code.
 */
#ifndef AUVI_H
#define AUVI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct auvi_grid auvi_grid;   /* opaque: a depth grid (or a row slab of one) resident on one GPU */

/* Interpolation methods.  0-2 are the reference's three (include/GridD.h:69,77,85);
 * 3-4 are extensions defined on the reference's own neighbour search. */
enum { AUVI_BILINEAR = 0, AUVI_CUBIC = 1, AUVI_KRIGING = 2, AUVI_NN = 3, AUVI_IDW = 4,
       /* opt-in, changes results: bilinear, and where the reference's bilinear returns NaN (all four
        * corners missing, GridH.cpp:186-198) the ring-search 4-nearest mean of the bicubic fallback (GridH.cpp:272-318) */
       AUVI_BILINEAR_SEARCH = 5,
       /* opt-in: IDW over the TRUE four nearest valid cells of the radius-10 window -- the reference's search without its two
        * early `count >= 4` breaks (GridH.cpp:82,115), ties to the cell enumerated first */
       AUVI_IDW_KNN = 6,
       /* opt-in: ordinary kriging on the reference's four picks with the exponential variogram FITTED to the grid instead of
        * the constants of GridH.cpp:371-376 (auvi_grid_fit_variogram; fitted on first use) */
       AUVI_KRIGING_FITTED = 7 };
/* Storage type of the depth grid and of lattice outputs. */
enum { AUVI_F64 = 0, AUVI_F32 = 1 };
/* How a lattice axis maps an output index to a coordinate:
 *   AUVI_AXIS_EXPANDED  c = lo + k*(hi-lo)/(new_n-1), new_n = f*(n-1)+1  (test_interpolation.cpp:91-109)
 *   AUVI_AXIS_NODES     c = lo + k*((hi-lo)/(n-1))                        (test_gebco.cpp:72-81)      */
enum { AUVI_AXIS_EXPANDED = 0, AUVI_AXIS_NODES = 1 };

/* ---- grid lifetime: replaces GridD::GridD + GridD::initialize (src/GridD.cu:41-83) and
 *      GridD::cleanup / ~GridD (src/GridD.cu:59-62,86-92) ------------------------------------- */

/* Upload a dense row-major host grid (row 0 = min_lat, as GridD.cu:67-72 flattens it).
 * dtype = type of host_rowmajor and of the device copy.  device = CUDA ordinal. */
int auvi_grid_create(const void* host_rowmajor, int dtype, int64_t n_lat, int64_t n_lon,
                     double min_lon, double max_lon, double min_lat, double max_lat,
                     int device, auvi_grid** out);

/* Upload a row slab of a larger grid: host_rows holds global rows [row0, row0+rows) densely (n_lon elements per row); n_lat
 * is the GLOBAL row count.  What one rank of a row-sharded job holds: its rows plus the halo. */
int auvi_grid_create_slab(const void* host_rows, int dtype, int64_t n_lat, int64_t n_lon, int64_t row0, int64_t rows,
                          double min_lon, double max_lon, double min_lat, double max_lat, int device, auvi_grid** out);

/* Adopt (borrow, never free) a grid -- or a row slab of one -- already resident in device memory.
 * dev_rows holds global rows [row0, row0+rows) with `ld` elements between rows; n_lat is the
 * GLOBAL row count (multi-GPU row sharding: each rank passes its slab + halo). */
int auvi_grid_adopt(const void* dev_rows, int dtype, int64_t n_lat, int64_t n_lon, int64_t ld,
                    int64_t row0, int64_t rows,
                    double min_lon, double max_lon, double min_lat, double max_lat,
                    int device, auvi_grid** out);

int auvi_grid_destroy(auvi_grid* g);      /* idempotent on NULL; large device blocks go to a process-wide cache */
int auvi_trim(void);                      /* release the cached device blocks (up to AUVI_CACHE_GB, default 40 GiB, are kept for reuse) */

/* ---- Point-list mode: replaces GridD::batch{Bilinear,Cubic,OrdinaryKriging}Interpolate
 *      (src/GridD.cu:95-150,156-193,199-236) and the three kernels of src/kernels.cu:173-546 ---- */

/* host_pts: n records of `stride_bytes` whose first two doubles are lon,lat (24 for struct
 * Point{lon,lat,elev}, include/Point.h:9-13).  The interpolated depth of record k is written to
 * (char*)host_out + k*out_stride_bytes (8 for a dense double array, 24 to fill Point::elev in
 * place); NaN outside the bounds (kernels.cu:193-196).  Synchronous.  Pinned staging and device
 * buffers persist across calls; copies and kernels of successive chunks overlap. */
int auvi_interp_points(auvi_grid* g, int method, const void* host_pts, int64_t n,
                       int64_t stride_bytes, void* host_out, int64_t out_stride_bytes);

/* Same on device-resident buffers, asynchronous on `stream` (a cudaStream_t, may be NULL).
 * dev_sel (optional, n x 8 int32) / dev_found (optional, n int32) receive the neighbour selection
 * {i0,j0,...,i3,j3} and candidate count (-1 outside, -2 no search needed) for parity checks. */
int auvi_interp_points_device(auvi_grid* g, int method, const void* dev_pts, int64_t n,
                              int64_t stride_bytes, double* dev_out_elev,
                              int32_t* dev_sel, int32_t* dev_found, void* stream);

/* ---- Lattice mode: the structured form of the same batches.  The query set is the separable
 *      lattice the reference drivers build -- generateExpandedGridQueryPoints
 *      (test_interpolation.cpp:91-109, axis kind EXPANDED, factor f => f*(n-1)+1 outputs per axis)
 *      or the grid nodes of test_gebco.cpp:72-81,150-160 (axis kind NODES, factor 1) -- without
 *      materialising 24-byte Points.  Output cell (J,I) equals what GridH::batch* returns for the
 *      query {lon(I), lat(J)}.  Rows [row_begin,row_end) only: the unit of multi-GPU sharding. ---- */

/* Output lattice dimensions for a grid and factors. */
int auvi_lattice_dims(const auvi_grid* g, int axis_kind, int f_lat, int f_lon,
                      int64_t* out_rows, int64_t* out_cols);

/* Upsample (fill = 0) or gap-fill (fill = 1: cells whose own value is valid are copied through,
 * NaN cells get method(query at that node); requires axis kind NODES and factors 1).
 * dev_out: (row_end-row_begin) rows of out_ld elements of the grid's dtype.  dev_sel9 (optional):
 * per output cell {found, i0,j0,...,i3,j3} int32.  Asynchronous on stream. */
int auvi_lattice_device(auvi_grid* g, int method, int axis_kind, int f_lat, int f_lon, int fill,
                        int64_t row_begin, int64_t row_end, void* dev_out, int64_t out_ld,
                        int32_t* dev_sel9, void* stream);

/* Host-buffer form: result rows are copied to host_out (dense, out_cols elements per row) inside
 * the call, row chunks double-buffered so that kernels overlap the device->host copies;
 * synchronous. */
int auvi_lattice(auvi_grid* g, int method, int axis_kind, int f_lat, int f_lon, int fill,
                 int64_t row_begin, int64_t row_end, void* host_out);

/* ---- Error metrics on device: replaces meanAbsoluteError / rootMeanSquareError /
 *      maxAbsoluteError (src/error_calculator.cpp:5-45), same NaN/denominator convention.
 *      out3 = {MAE, RMSE, Max}; out_nan = number of NaN estimates.  Synchronous on `stream`. */
int auvi_error_metrics_device(const void* dev_truth, const void* dev_est, int dtype, int64_t n,
                              double* out3, int64_t* out_nan, void* stream);

/* The same three numbers for a gap fill, without gathering the removed cells into point lists: cell (r,c) of rows
 * [row_begin,row_end) counts iff `masked` holds NaN there; estimate = dev_filled[(r-row_begin)*filled_ld + c], truth =
 * dev_truth[(r-row_begin)*truth_ld + c] (both of the grid's dtype).  This is the RMSE block of test_gebco.cpp:198-230
 * (SURVEY.md s8(f) N3).  out_count = number of cells compared (the reference's N). */
int auvi_fill_metrics_device(auvi_grid* masked, const void* dev_filled, int64_t filled_ld, const void* dev_truth,
                             int64_t truth_ld, int64_t row_begin, int64_t row_end, double* out3, int64_t* out_nan,
                             int64_t* out_count, void* stream);

/* ---- Grid-B data preparation: replaces the reference's host-side tool chain for this path -- netCDF4 read + row flip
 *      + seeded removal of cells + CSV round trip (code/subset_bathymetry.py:8-85) and the CSV reader of the driver
 *      (test_gebco.cpp:19-40) -- SURVEY.md s8(f) N1.  Host-only entries work without a GPU. ------------------------ */

/* A variable of a NetCDF-3 classic (CDF-1) or 64-bit-offset (CDF-2) file held in memory. */
typedef struct auvi_nc_var {
    int32_t nc_type;       /* 1 byte, 2 char, 3 short, 4 int, 5 float, 6 double */
    int32_t elem_bytes;
    int32_t ndims;
    int32_t has_fill;      /* _FillValue attribute present (reported only: the decoders do not map it to NaN --
                            * GEBCO tiles carry no missing cells; mask with auvi_grid_mask_cells / _hash) */
    int64_t shape[4];
    int64_t n_elems;
    int64_t data_offset;   /* bytes from the start of the file image; elements are big-endian, row-major */
    double scale_factor, add_offset, fill_value;
} auvi_nc_var;

/* Host only: locate `var_name` (GEBCO: "elevation", "lat", "lon") in the file image. */
int auvi_netcdf3_find(const void* file_image, int64_t n_bytes, const char* var_name, auvi_nc_var* out);
/* Host only: decode a whole variable to doubles (coordinate axes; applies scale_factor / add_offset). */
int auvi_netcdf3_read_f64(const void* file_image, int64_t n_bytes, const char* var_name, double* out, int64_t n_out);
/* Host only: out_idx[0..n) = numpy.random.seed(seed); numpy.random.choice(total, n, replace=False)
 * (subset_bathymetry.py:32-39: the legacy MT19937 stream, Fisher-Yates permutation prefix).  Like numpy it permutes all
 * `total` indices: 8 bytes of host memory per cell, sequential -- meant for GEBCO tiles; at BASELINE config 4's size use
 * auvi_grid_mask_hash. */
int auvi_legacy_choice(int64_t total, int64_t n, uint32_t seed, int64_t* out_idx);

/* Upload file-order elements (nc_type 3..6, big- or little-endian) and decode them on the device into a grid of
 * `dtype`: value*scale+offset, optionally flipping the row order (subset_bathymetry.py:16-17 `iloc[::-1]`).
 * GEBCO int16 tiles move 2 bytes per cell over PCIe instead of 8. */
int auvi_grid_create_raw(const void* host_raw, int nc_type, int big_endian, int flip_rows, double scale, double offset,
                         int dtype, int64_t n_lat, int64_t n_lon, double min_lon, double max_lon, double min_lat,
                         double max_lat, int device, auvi_grid** out);
/* A grid from CSV matrix text (one row per latitude, comma separated, "nan" for missing cells: reduced_data.csv of
 * subset_bathymetry.py:78-85, the files readGridCSV parses, test_gebco.cpp:19-40).  The text is uploaded as it is and
 * parsed on the device; every value equals what std::stod returns for the cell.  Rows must have equal field counts.
 * Device scratch while parsing: the text plus 16 bytes per cell.  auvi_csv_dims is host only. */
int auvi_csv_dims(const char* text, int64_t n_bytes, int64_t* n_rows, int64_t* n_cols);
int auvi_grid_create_csv(const char* text, int64_t n_bytes, int dtype, double min_lon, double max_lon, double min_lat,
                         double max_lat, int device, auvi_grid** out);
/* Remove cells: host_flat_idx[k] = row*n_lon+col (subset_bathymetry.py:39).  The cells become NaN in the grid's own
 * storage (also for adopted memory); host_truth (optional, n elements of the grid's dtype) receives their former values
 * in list order -- the third column of reference_missing.csv (:49-56).  Cells outside a slab are skipped (truth NaN). */
int auvi_grid_mask_cells(auvi_grid* g, const int64_t* host_flat_idx, int64_t n, void* host_truth);
/* Remove each cell with probability `fraction`, decided by a hash of (GLOBAL flat index, seed): ranks holding slabs of
 * one grid draw the same mask without communication (BASELINE config 4).  Asynchronous on stream unless out_masked. */
int auvi_grid_mask_hash(auvi_grid* g, double fraction, uint64_t seed, int64_t* out_masked, void* stream);
/* Copy grid rows [row_begin,row_end) back to a dense host array of the grid's dtype. */
int auvi_grid_read(auvi_grid* g, int64_t row_begin, int64_t row_end, void* host_out);

/* ---- Opt-in: fitted variogram (changes results, so never the default).  The model is
 *      gamma(h) = c0 + c1 * (1 - exp(-h / range)), h in degrees as in GridH.cpp:371-380.  auvi_grid_fit_variogram reduces the
 *      empirical semivariances of the grid at lags 1, 2, 4, 8 cells along both axes on the device and fits (c0, c1, range)
 *      by weighted least squares over a ladder of candidate ranges; needs the whole grid resident.  A row slab takes the
 *      parameters from auvi_grid_set_variogram.  AUVI_KRIGING_FITTED solves the kriging system in covariance form
 *      (c1 * exp(-h / range) off the diagonal, c0 + c1 on it) on the same four cells the reference's search picks. ---- */
int auvi_grid_fit_variogram(auvi_grid* g, double* out_c0_c1_range);      /* out: 3 doubles */
int auvi_grid_set_variogram(auvi_grid* g, double c0, double c1, double range);
/* Host only: the fit itself, from the 16 sums the device reduction delivers ([axis][lag 1,2,4,8][sum of squares, pairs]). */
int auvi_variogram_fit_from_sums(const double* sums16, double lon_step, double lat_step, double* out_c0_c1_range);

/* ---- Peer memory (one process per GPU): gather the row shards WITHOUT a collective.  The consumer exports its result
 *      buffer (any address inside a cudaMalloc allocation), producers map it and pass the mapped address as dev_out of
 *      auvi_lattice_device: the kernel's 16-byte stores go to the peer over NVLink, compute and transfer are one kernel and no
 *      staging copy exists (SURVEY.md s8(e) "fusion with the collective").  handle72: 72 bytes to move between processes. */
int auvi_peer_export(const void* dev_ptr, unsigned char* handle72);
int auvi_peer_open(const unsigned char* handle72, void** out_ptr);     /* in another process of the same node */
int auvi_peer_close(void* ptr);                                        /* a pointer returned by auvi_peer_open */

/* ---- Several GPUs of one box, ONE process (the reference has no multi-GPU path: GridD owns one device grid,
 *      src/GridD.cu:41-83, and every batch runs on it, :95-236).  Every output cell is independent, so the work shards
 *      with no data-path collective: lattice / gap-fill jobs by blocks of grid rows -- shard k holds
 *      its rows + a 14-row halo, replicated at upload, never exchanged -- and point lists by slices of the query list
 *      over a replicated grid.  One host thread per device drives the single-GPU entries above. ------------------- */
typedef struct auvi_multi auvi_multi;

/* Upload one dense host grid to n_gpus devices (devices: CUDA ordinals, NULL = 0..n_gpus-1; an ordinal may repeat).
 * replicate = 0: row slabs + halo (lattice / gap fill; what BASELINE config 4 needs at 65536^2);
 * replicate = 1: every device holds the whole grid (also serves point lists: GridD::batch*). */
int auvi_multi_create(const void* host_rowmajor, int dtype, int64_t n_lat, int64_t n_lon, double min_lon, double max_lon,
                      double min_lat, double max_lat, int n_gpus, const int* devices, int replicate, auvi_multi** out);
int auvi_multi_destroy(auvi_multi* m);
/* Host only (no GPU needed): the row plan auvi_multi_create uses for shard k of n_gpus.  out6 = {own_lo, own_hi, in_lo, in_hi,
 * row_lo, row_hi}: grid rows owned, grid rows held (own + 14-row halo; all rows when replicated), lattice rows produced at f_lat. */
int auvi_multi_plan(int64_t n_lat, int n_gpus, int k, int f_lat, int replicate, int64_t* out6);
int auvi_multi_count(const auvi_multi* m);
/* Shard k: its device, its single-GPU handle (borrowed) and the lattice rows [row_lo,row_hi) it produces at factor f_lat. */
int auvi_multi_shard(const auvi_multi* m, int k, int f_lat, int* device, auvi_grid** grid, int64_t* row_lo, int64_t* row_hi);
/* auvi_grid_mask_hash on every shard: one global mask, no communication. */
int auvi_multi_mask_hash(auvi_multi* m, double fraction, uint64_t seed);
/* auvi_lattice over all devices at once; host_out is the whole lattice (rows x cols, dense).  Synchronous. */
int auvi_multi_lattice(auvi_multi* m, int method, int axis_kind, int f_lat, int f_lon, int fill, void* host_out);
/* Device-resident form, asynchronous: dev_out[k] = where shard k's kernel writes the FIRST of its rows (pitch out_ld
 * elements).  Separate buffers leave the result sharded; addresses inside one buffer on one device (after
 * auvi_multi_enable_peer) gather it there -- the kernels' own stores cross NVLink, compute and gather are one kernel. */
int auvi_multi_lattice_device(auvi_multi* m, int method, int axis_kind, int f_lat, int f_lon, int fill, void* const* dev_out,
                              int64_t out_ld);
int auvi_multi_enable_peer(auvi_multi* m, int root_shard);   /* every shard's device may store to root_shard's device */
int auvi_multi_sync(auvi_multi* m);                          /* waits for every device */
float auvi_multi_last_kernel_ms(const auvi_multi* m);        /* device time of the last job, max over the devices */
/* auvi_interp_points with the query list cut into one slice per device (needs replicate = 1). */
int auvi_multi_interp_points(auvi_multi* m, int method, const void* host_pts, int64_t n, int64_t stride_bytes, void* host_out,
                             int64_t out_stride_bytes);

/* Host only: first-touch `bytes` of freshly allocated host memory at `p` from the library's worker threads (GridD::batch*
 * must return a new std::vector<Point> by value -- include/GridD.h:69-85 --: a 120 MB result vector costs ~40 ms of
 * single-threaded page faults otherwise).  Writes one zero byte per 4 KiB page: only for memory whose contents do not matter yet. */
int auvi_host_prefault(void* p, int64_t bytes);

/* ---- diagnostics ------------------------------------------------------------------------------ */
const char* auvi_last_error(void);          /* thread-local message of the last failure */
float auvi_last_kernel_ms(const auvi_grid* g); /* device time of the kernels of the last synchronous call */
int64_t auvi_launch_count(void);            /* kernels launched by this library so far (process-wide) */
int auvi_uses_tma(const auvi_grid* g);      /* 1 if the last lattice launch staged tiles by TMA */
int auvi_uses_window(const auvi_grid* g);   /* window-load form of the FP32 bicubic kernel in the last lattice launch: 0 = generic,
                                               1 / 2 = the form for longitude factor 1 / 2 (csrc/upsample.cu) */
int auvi_device_count(void);                /* CUDA devices visible (0 when none: compute calls fail) */
int auvi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AUVI_H */
