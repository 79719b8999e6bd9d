// api.cu -- the C ABI of libauvi.so (include/auvi.h): grid lifetime, the host-buffer and
// device-buffer forms of the three calling modes (point list, lattice, metrics), diagnostics.
//
// Replaces the host half of the reference's device class, code/src/GridD.cu:
//   GridD::initialize  (:65-83)   -> auvi_grid_create        (one upload, grid stays resident)
//   GridD::batch*      (:95-236)  -> auvi_interp_points      (persistent pinned staging + device
//                                    buffers, chunked so H2D / kernel / D2H of successive chunks
//                                    overlap on two streams; the reference mallocs, copies
//                                    synchronously and frees on every call)
//   GridD::cleanup     (:86-92)   -> auvi_grid_destroy
// and adds the structured lattice mode (upsample.cu) for the query sets the reference drivers build.
//
// Host arithmetic that decides discrete outcomes (grid steps, lattice coordinates, index-space
// images) is written with the reference's exact expressions and compiled with -ffp-contract=off
// (Makefile): SURVEY.md section 0 facts 3-4.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/auvi.h"
#include "hostpool.h"
#include "launch.h"

using namespace auvi;

namespace auvi { int set_error(const std::string& msg); }   // for the host-only translation units (prep.cpp)
extern "C" int auvi_grid_fit_variogram(auvi_grid* g, double* out_c0_c1_range);

namespace {

thread_local std::string t_error;
std::atomic<int64_t> g_launches{0};

int fail(const std::string& msg) { t_error = msg; return 1; }
int fail_cuda(const char* what, cudaError_t e) {
    t_error = std::string(what) + ": " + cudaGetErrorString(e);
    return 2;
}
#define AUVI_CUDA(call)                                                   \
    do {                                                                  \
        cudaError_t e_ = (call);                                          \
        if (e_ != cudaSuccess) return fail_cuda(#call, e_);               \
    } while (0)

struct AxisOwned {
    std::vector<double> coord, pos;
    std::vector<int> base;
    double* d_coord = nullptr;
    double* d_pos = nullptr;
    int* d_base = nullptr;
    mutable int window_mode = -1;                                  // upsample.cu: verdict of the window-load pattern check, cached
    AxisTables view() const {
        AxisTables t;
        t.coord = d_coord; t.pos = d_pos; t.base = d_base; t.h_base = base.data();
        t.window_mode_cache = &window_mode;
        t.n = static_cast<int>(coord.size());
        return t;
    }
};

// Process-wide cache of device blocks (grid storage, lattice staging, axis tables).  cudaMalloc/cudaFree cost up
// to ~150 ms per grid lifetime on a context that also maps a large pinned host buffer -- every cudaFree is a
// device-wide synchronisation plus page-table work (measured: profiles/r01_e2e_pieces.txt); a caller that creates
// and destroys grids per batch -- as the reference's drivers do -- should not pay that every time.  Blocks are
// only returned here after the device is idle (auvi_grid_destroy synchronises), so reuse on any stream is safe.
// At most kCacheBytes (40 GiB by default) are held; auvi_trim() releases them.
struct DevBlock { void* p; size_t bytes; int device; };
std::mutex g_cache_mu;
std::vector<DevBlock> g_cache;
std::map<void*, size_t> g_block_bytes;          // every live block handed out by cached_malloc -> the size it was ALLOCATED with
size_t g_cache_held = 0;
// AUVI_CACHE_GB overrides (0 = no caching).  The default keeps a 65536^2 f32 slab (17 GB) plus its staging: a cudaFree of
// 8.6 GB costs 230 ms, a third of an end-to-end step at BASELINE config 4 on two GPUs (profiles/r02_bench_n2_first.json).
const size_t kCacheBytes = [] { const char* e = getenv("AUVI_CACHE_GB"); return (e ? static_cast<size_t>(atol(e)) : 40ull) << 30; }();
constexpr size_t kCacheMinBlock = 4ull << 10;

cudaError_t cached_malloc(void** out, size_t bytes, int device) {
    if (bytes >= kCacheMinBlock) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        size_t best = g_cache.size();
        for (size_t k = 0; k < g_cache.size(); ++k)
            if (g_cache[k].device == device && g_cache[k].bytes >= bytes && g_cache[k].bytes <= bytes + bytes / 4 + 4096 &&
                (best == g_cache.size() || g_cache[k].bytes < g_cache[best].bytes)) best = k;
        if (best != g_cache.size()) {
            *out = g_cache[best].p;
            g_cache_held -= g_cache[best].bytes;
            g_block_bytes[*out] = g_cache[best].bytes;              // a reused block keeps its real size
            g_cache.erase(g_cache.begin() + best);
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {                         // give the cache back and retry once
        cudaGetLastError();
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (auto& b : g_cache) cudaFree(b.p);
        g_cache.clear(); g_cache_held = 0;
        e = cudaMalloc(out, bytes);
    }
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        g_block_bytes[*out] = bytes;
    }
    return e;
}

// The block goes back to the cache under the size it was allocated with (a reused block may be up to 25 % larger than
// what its last user asked for), so the 6 GiB cap counts real memory.  `asked` is only a fallback for foreign pointers.
void cached_free(void* p, size_t asked, int device) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_cache_mu);
    size_t bytes = asked;
    auto it = g_block_bytes.find(p);
    if (it != g_block_bytes.end()) { bytes = it->second; g_block_bytes.erase(it); }
    if (bytes >= kCacheMinBlock && bytes <= kCacheBytes) {
        while (!g_cache.empty() && g_cache_held + bytes > kCacheBytes) {   // evict oldest
            cudaFree(g_cache.front().p);
            g_cache_held -= g_cache.front().bytes;
            g_cache.erase(g_cache.begin());
        }
        g_cache.push_back(DevBlock{p, bytes, device});
        g_cache_held += bytes;
        return;
    }
    cudaFree(p);
}

// A block that must not be reused (an error left work in flight on it): plain cudaFree, bookkeeping dropped.
void uncached_free(void* p) {
    if (!p) return;
    { std::lock_guard<std::mutex> lk(g_cache_mu); g_block_bytes.erase(p); }
    cudaFree(p);
}

constexpr int64_t kPointChunk = 1 << 20;            // most queries per pipeline stage (16 MiB in, 8 MiB out)
constexpr int64_t kPointChunkMin = 1 << 16;         // smaller batches are cut into ~4 stages so that packing overlaps the copies
constexpr int64_t kLatticeChunkBytes = 256ll << 20; // device staging per pipeline stage, lattice host form
constexpr int64_t kBounceBytes = 64ll << 20;        // ... when the destination is pageable: stage size of the pinned bounce ring

}  // namespace

int auvi::set_error(const std::string& msg) { return fail(msg); }

struct PointStaging {
    int device = 0;
    double* h_in[2] = {nullptr, nullptr};
    double* h_out[2] = {nullptr, nullptr};
    double* d_in[2] = {nullptr, nullptr};
    double* d_out[2] = {nullptr, nullptr};
};

struct auvi_grid {
    GridDesc d;
    int device = 0;
    void* owned = nullptr;                            // device allocation we must free (create), else null
    size_t owned_bytes = 0;
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t ev_k0[2] = {nullptr, nullptr}, ev_k1[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    // point-list staging (double-buffered)
    double* h_in[2] = {nullptr, nullptr};             // pinned, packed {lon,lat}
    double* h_out[2] = {nullptr, nullptr};            // pinned
    double* d_in[2] = {nullptr, nullptr};
    double* d_out[2] = {nullptr, nullptr};
    PointStaging* staging = nullptr;                  // owner of the four above (process-wide pool)
    // lattice host-form staging
    void* d_rows[2] = {nullptr, nullptr};
    size_t d_rows_bytes = 0;
    std::map<long long, AxisOwned*> axes;             // key: which*2^40 + kind*2^32 + factor
    float last_ms = 0.f;
    int last_tma = 0, last_window = 0;
    // opt-in AUVI_KRIGING_FITTED: exponential model c0 + c1 * (1 - exp(-h / range)) fitted to this grid (or set by the caller)
    bool vg_valid = false;
    double vg_c0 = 0.0, vg_c1 = 0.0, vg_range = 0.0;
};

namespace {

int make_streams(auvi_grid* g) {
    for (int k = 0; k < 2; ++k) {
        AUVI_CUDA(cudaStreamCreateWithFlags(&g->st[k], cudaStreamNonBlocking));
        AUVI_CUDA(cudaEventCreate(&g->ev_k0[k]));
        AUVI_CUDA(cudaEventCreate(&g->ev_k1[k]));
        AUVI_CUDA(cudaEventCreateWithFlags(&g->ev_done[k], cudaEventDisableTiming));
    }
    return 0;
}

int check_common(const auvi_grid* g, int method) {
    if (!g) return fail("null grid handle");
    if (method < AUVI_BILINEAR || method > AUVI_KRIGING_FITTED) return fail("unknown interpolation method");
    return 0;
}

// The descriptor and kernel method a C-ABI method runs with.  AUVI_KRIGING_FITTED is the kriging kernel on the grid's own
// fitted model in covariance form (exact.cuh GridView::vg_*); the fit happens on first use.
int resolve_method(auvi_grid* g, int method, GridDesc* d, int* kernel_method) {
    *d = g->d;
    *kernel_method = method;
    if (method != AUVI_KRIGING_FITTED) return 0;
    if (!g->vg_valid) {
        double unused[3];
        if (auvi_grid_fit_variogram(g, unused)) return 1;
    }
    d->vg_nugget = g->vg_c1; d->vg_sill = -g->vg_c1; d->vg_inv_range = 1.0 / g->vg_range; d->vg_diag = g->vg_c0 + g->vg_c1;
    *kernel_method = AUVI_KRIGING;
    return 0;
}

void free_axis(AxisOwned* a, int device) {
    cached_free(a->d_coord, sizeof(double) * a->coord.size(), device);
    cached_free(a->d_pos, sizeof(double) * a->pos.size(), device);
    cached_free(a->d_base, sizeof(int) * a->base.size(), device);
    delete a;
}

// Per-axis lattice tables.  `which` 0 = longitude (columns), 1 = latitude (rows).
int get_axis(auvi_grid* g, int which, int kind, int factor, AxisOwned** out) {
    const long long key = (static_cast<long long>(which) << 40) | (static_cast<long long>(kind) << 32) | factor;
    auto it = g->axes.find(key);
    if (it != g->axes.end()) { *out = it->second; return 0; }
    const int n = which == 0 ? g->d.n_lon : g->d.n_lat;
    const double lo = which == 0 ? g->d.min_lon : g->d.min_lat;
    const double hi = which == 0 ? g->d.max_lon : g->d.max_lat;
    const double step = which == 0 ? g->d.lon_step : g->d.lat_step;
    if (factor < 1) return fail("lattice factor must be >= 1");
    if (kind == AUVI_AXIS_NODES && factor != 1) return fail("AUVI_AXIS_NODES requires factor 1");
    const int64_t n_out64 = static_cast<int64_t>(factor) * (n - 1) + 1;
    if (n_out64 > (1ll << 30)) return fail("lattice axis too long");
    const int n_out = static_cast<int>(n_out64);
    AxisOwned* a = new (std::nothrow) AxisOwned;
    if (!a) return fail("out of host memory");
    a->coord.resize(n_out); a->pos.resize(n_out); a->base.resize(n_out);
    for (int k = 0; k < n_out; ++k) {
        double c;
        if (kind == AUVI_AXIS_NODES) c = lo + k * step;                       // test_gebco.cpp:79-80
        else c = lo + k * (hi - lo) / (n_out - 1);                            // test_interpolation.cpp:99-104
        const double p = (c - lo) / step;                                      // GridH.cpp:167-168
        const bool out_of_bounds = c < lo || c > hi;                           // GridH.cpp:162 (per axis)
        a->coord[k] = c;
        a->pos[k] = out_of_bounds ? std::nan("") : p;
        int b = static_cast<int>(std::floor(p));
        a->base[k] = b < 0 ? 0 : (b > n - 1 ? n - 1 : b);
    }
    cudaError_t e;
    if ((e = cached_malloc(reinterpret_cast<void**>(&a->d_coord), sizeof(double) * n_out, g->device)) != cudaSuccess ||
        (e = cached_malloc(reinterpret_cast<void**>(&a->d_pos), sizeof(double) * n_out, g->device)) != cudaSuccess ||
        (e = cached_malloc(reinterpret_cast<void**>(&a->d_base), sizeof(int) * n_out, g->device)) != cudaSuccess ||
        (e = cudaMemcpy(a->d_coord, a->coord.data(), sizeof(double) * n_out, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(a->d_pos, a->pos.data(), sizeof(double) * n_out, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(a->d_base, a->base.data(), sizeof(int) * n_out, cudaMemcpyHostToDevice)) != cudaSuccess) {
        free_axis(a, g->device);
        return fail_cuda("axis table upload", e);
    }
    g->axes[key] = a;
    *out = a;
    return 0;
}

int fill_desc(GridDesc& d, int dtype, int64_t n_lat, int64_t n_lon, double min_lon, double max_lon,
              double min_lat, double max_lat) {
    if (dtype != AUVI_F64 && dtype != AUVI_F32) return fail("dtype must be AUVI_F64 or AUVI_F32");
    if (n_lat < 2 || n_lon < 2) return fail("grid needs at least 2 points per axis");
    if (n_lat > (1ll << 30) || n_lon > (1ll << 30)) return fail("grid axis too long");
    d.dtype = dtype;
    d.n_lat = static_cast<int>(n_lat); d.n_lon = static_cast<int>(n_lon);
    d.min_lon = min_lon; d.max_lon = max_lon; d.min_lat = min_lat; d.max_lat = max_lat;
    d.lon_step = (max_lon - min_lon) / (d.n_lon - 1);                          // GridD.cu:52-53
    d.lat_step = (max_lat - min_lat) / (d.n_lat - 1);
    return 0;
}

// Device scratch of one metrics call: the partial sums and the five results, one cached block.
struct MetricsScratch {
    void* scratch = nullptr;
    double* result5 = nullptr;
    size_t bytes = 0;
    int device = 0;
    int open() {
        AUVI_CUDA(cudaGetDevice(&device));
        bytes = (metrics_scratch_bytes() + 255) / 256 * 256 + 4096;
        AUVI_CUDA(cached_malloc(&scratch, bytes, device));
        result5 = reinterpret_cast<double*>(static_cast<char*>(scratch) + bytes - 4096);
        return 0;
    }
    int close(cudaError_t e, cudaStream_t st, double (&h)[5]) {    // copies the results out and returns the block
        if (e == cudaSuccess) e = cudaMemcpyAsync(h, result5, sizeof h, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) { cached_free(scratch, bytes, device); return 0; }
        uncached_free(scratch);
        return fail_cuda("metrics", e);
    }
};

// Pinned + device staging of the point-list pipeline, kept process-wide: a caller that builds a GridD per batch -- as
// the reference's drivers do -- should not pin 48 MB of host memory and free it again on every grid.
std::mutex g_staging_mu;
std::vector<PointStaging*> g_staging_free;

void release_staging(PointStaging* ps) {
    for (int k = 0; k < 2; ++k) {
        if (ps->h_in[k]) cudaFreeHost(ps->h_in[k]);
        if (ps->h_out[k]) cudaFreeHost(ps->h_out[k]);
        cudaFree(ps->d_in[k]); cudaFree(ps->d_out[k]);
    }
    delete ps;
}

int ensure_point_staging(auvi_grid* g) {
    if (g->h_in[0]) return 0;
    PointStaging* ps = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_staging_mu);
        for (size_t k = 0; k < g_staging_free.size(); ++k)
            if (g_staging_free[k]->device == g->device) { ps = g_staging_free[k]; g_staging_free.erase(g_staging_free.begin() + k); break; }
    }
    if (!ps) {
        ps = new (std::nothrow) PointStaging;
        if (!ps) return fail("out of host memory");
        ps->device = g->device;
        cudaError_t e = cudaSuccess;
        for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
            e = cudaHostAlloc(reinterpret_cast<void**>(&ps->h_in[k]), sizeof(double) * 2 * kPointChunk, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&ps->h_out[k]), sizeof(double) * kPointChunk, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ps->d_in[k]), sizeof(double) * 2 * kPointChunk);
            if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ps->d_out[k]), sizeof(double) * kPointChunk);
        }
        if (e != cudaSuccess) { release_staging(ps); return fail_cuda("point staging", e); }
    }
    g->staging = ps;
    for (int k = 0; k < 2; ++k) { g->h_in[k] = ps->h_in[k]; g->h_out[k] = ps->h_out[k]; g->d_in[k] = ps->d_in[k]; g->d_out[k] = ps->d_out[k]; }
    return 0;
}

// Pinned bounce rings of the lattice host form (pageable destinations) and of the pageable grid upload: a process-wide
// pool, so that neither a caller that makes one call after another nor several host threads driving several devices
// (multi.cu) pin and unpin 128 MiB per call.  auvi_trim() releases the idle ones.
struct BounceRing { char* buf[2]; size_t bytes; };
std::mutex g_bounce_mu;
std::vector<BounceRing> g_bounce_free;

int acquire_bounce(size_t& need, char* (&out)[2]) {   // `need` comes back as the ring's real stage size
    {
        std::lock_guard<std::mutex> lk(g_bounce_mu);
        for (size_t k = 0; k < g_bounce_free.size(); ++k)
            if (g_bounce_free[k].bytes >= need) {
                out[0] = g_bounce_free[k].buf[0]; out[1] = g_bounce_free[k].buf[1];
                need = g_bounce_free[k].bytes;
                g_bounce_free.erase(g_bounce_free.begin() + k);
                return 0;
            }
        if (!g_bounce_free.empty()) {                              // too small: replace it rather than keep both
            cudaFreeHost(g_bounce_free.back().buf[0]); cudaFreeHost(g_bounce_free.back().buf[1]);
            g_bounce_free.pop_back();
        }
    }
    out[0] = out[1] = nullptr;
    for (int k = 0; k < 2; ++k) {
        const cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&out[k]), need, cudaHostAllocPortable);
        if (e != cudaSuccess) { if (out[0]) cudaFreeHost(out[0]); out[0] = out[1] = nullptr; return fail_cuda("pinned bounce ring", e); }
    }
    return 0;
}

void release_bounce(char* (&ring)[2], size_t bytes) {
    if (!ring[0]) return;
    std::lock_guard<std::mutex> lk(g_bounce_mu);
    g_bounce_free.push_back(BounceRing{{ring[0], ring[1]}, bytes});
    ring[0] = ring[1] = nullptr;
}

// Dense host rows -> pitched device rows.  A pinned source goes in one 2-D copy; a pageable one (a std::vector) is
// copied by host threads into the pinned ring stage by stage while the previous stage is on the wire, instead of the
// driver's own staged copy.
cudaError_t upload_rows(void* dev, size_t dev_pitch, const void* host, size_t row_bytes, int64_t n_rows) {
    cudaPointerAttributes attr;
    bool pageable = true;
    if (cudaPointerGetAttributes(&attr, host) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
    else cudaGetLastError();
    const size_t total = row_bytes * static_cast<size_t>(n_rows);
    if (!pageable || total < (8u << 20))
        return cudaMemcpy2D(dev, dev_pitch, host, row_bytes, row_bytes, static_cast<size_t>(n_rows), cudaMemcpyHostToDevice);
    int64_t stage_rows = kBounceBytes / static_cast<int64_t>(row_bytes);
    if (stage_rows < 1) stage_rows = 1;
    size_t need = static_cast<size_t>(stage_rows) * row_bytes;
    char* ring[2] = {nullptr, nullptr};
    if (acquire_bounce(need, ring)) return cudaErrorMemoryAllocation;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
    HostPool& pool = HostPool::get();
    int64_t c = 0;
    for (int64_t r = 0; r < n_rows && e == cudaSuccess; r += stage_rows, ++c) {
        const int b = static_cast<int>(c & 1);
        const int64_t cnt = r + stage_rows < n_rows ? stage_rows : n_rows - r;
        if (c >= 2) e = cudaEventSynchronize(ev[b]);
        if (e != cudaSuccess) break;
        const char* src = static_cast<const char*>(host) + static_cast<size_t>(r) * row_bytes;
        char* const dst = ring[b];
        pool.for_range(cnt * static_cast<int64_t>(row_bytes), pool.size(), 4096, [&](int64_t lo, int64_t hi) {
            std::memcpy(dst + lo, src + lo, static_cast<size_t>(hi - lo));
        });
        e = cudaMemcpy2DAsync(static_cast<char*>(dev) + static_cast<size_t>(r) * dev_pitch, dev_pitch, ring[b], row_bytes, row_bytes,
                              static_cast<size_t>(cnt), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaEventRecord(ev[b], st);
    }
    if (st) { const cudaError_t e2 = cudaStreamSynchronize(st); if (e == cudaSuccess) e = e2; }
    for (int k = 0; k < 2; ++k) if (ev[k]) cudaEventDestroy(ev[k]);
    if (st) cudaStreamDestroy(st);
    release_bounce(ring, need);
    return e;
}

// Host threads for packing / unpacking point records (memory-bound loops): a few are enough to reach the copy rate.
int pack_threads(int64_t cnt) {
    if (cnt < 32768) return 1;
    const int hw = HostPool::get().size();
    return hw < 8 ? hw : 8;
}

}  // namespace

extern "C" {

// First-touch a fresh host allocation from the library's worker threads (one write per 4 KiB page): a caller that must
// return a newly allocated 100 MB std::vector -- GridD::batch*, whose signature returns the result vector by value -- pays
// ~40 ms of single-threaded page faults for it otherwise (profiles/r02_points_vs_reference_gpu.txt).
int auvi_host_prefault(void* p, int64_t bytes) {
    if (!p || bytes <= 0) return 0;
    HostPool& pool = HostPool::get();
    char* const base = static_cast<char*>(p);
    pool.for_range(bytes, bytes < (2 << 20) ? 1 : pool.size(), 4096, [&](int64_t lo, int64_t hi) {
        for (int64_t at = lo; at < hi; at += 4096) base[at] = 0;
    });
    return 0;
}

int auvi_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int auvi_version(void) { return 101; }

int auvi_trim(void) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (auto& b : g_cache) { cudaSetDevice(b.device); cudaFree(b.p); }
    g_cache.clear();
    g_cache_held = 0;
    std::lock_guard<std::mutex> lk2(g_staging_mu);
    for (PointStaging* ps : g_staging_free) { cudaSetDevice(ps->device); release_staging(ps); }
    g_staging_free.clear();
    {
        std::lock_guard<std::mutex> lk3(g_bounce_mu);
        for (BounceRing& r : g_bounce_free) { cudaFreeHost(r.buf[0]); cudaFreeHost(r.buf[1]); }
        g_bounce_free.clear();
    }
    return 0;
}
const char* auvi_last_error(void) { return t_error.c_str(); }
int64_t auvi_launch_count(void) { return g_launches.load(); }
float auvi_last_kernel_ms(const auvi_grid* g) { return g ? g->last_ms : 0.f; }
int auvi_uses_tma(const auvi_grid* g) { return g ? g->last_tma : 0; }
int auvi_uses_window(const auvi_grid* g) { return g ? g->last_window : 0; }

int auvi_grid_create_slab(const void* host_rows, int dtype, int64_t n_lat, int64_t n_lon, int64_t row0, int64_t rows,
                          double min_lon, double max_lon, double min_lat, double max_lat, int device, auvi_grid** out) {
    if (!out) return fail("null output handle");
    *out = nullptr;
    if (!host_rows) return fail("null host grid");
    if (row0 < 0 || rows < 1 || row0 + rows > n_lat) return fail("slab rows outside the grid");
    if (auvi_device_count() <= 0) return fail("no CUDA device: libauvi has no CPU fallback");
    auvi_grid* g = new (std::nothrow) auvi_grid;
    if (!g) return fail("out of host memory");
    if (fill_desc(g->d, dtype, n_lat, n_lon, min_lon, max_lon, min_lat, max_lat)) { delete g; return 1; }
    g->device = device;
    const size_t es = dtype == AUVI_F64 ? 8 : 4;
    // rows padded to a 16-byte pitch so that TMA can address any grid width
    const int64_t ld = (n_lon * es + 15) / 16 * 16 / es;
    cudaError_t e = cudaSetDevice(device);
    g->owned_bytes = static_cast<size_t>(ld) * rows * es;
    if (e == cudaSuccess) e = cached_malloc(&g->owned, g->owned_bytes, device);
    if (e == cudaSuccess) e = upload_rows(g->owned, ld * es, host_rows, n_lon * es, rows);
    if (e != cudaSuccess) { uncached_free(g->owned); delete g; return fail_cuda("grid upload", e); }
    g->d.z = g->owned; g->d.ld = ld; g->d.row0 = static_cast<int>(row0); g->d.rows = static_cast<int>(rows);
    if (make_streams(g)) { auvi_grid_destroy(g); return 2; }
    *out = g;
    return 0;
}

int auvi_grid_create(const void* host_rowmajor, int dtype, int64_t n_lat, int64_t n_lon,
                     double min_lon, double max_lon, double min_lat, double max_lat,
                     int device, auvi_grid** out) {
    if (!out) return fail("null output handle");
    *out = nullptr;
    if (!host_rowmajor) return fail("null host grid");
    return auvi_grid_create_slab(host_rowmajor, dtype, n_lat, n_lon, 0, n_lat, min_lon, max_lon, min_lat, max_lat, device, out);
}

int auvi_grid_adopt(const void* dev_rows, int dtype, int64_t n_lat, int64_t n_lon, int64_t ld,
                    int64_t row0, int64_t rows,
                    double min_lon, double max_lon, double min_lat, double max_lat,
                    int device, auvi_grid** out) {
    if (!out) return fail("null output handle");
    *out = nullptr;
    if (!dev_rows) return fail("null device grid");
    if (auvi_device_count() <= 0) return fail("no CUDA device: libauvi has no CPU fallback");
    if (ld < n_lon) return fail("ld must be >= n_lon");
    if (row0 < 0 || rows < 1 || row0 + rows > n_lat) return fail("slab rows outside the grid");
    auvi_grid* g = new (std::nothrow) auvi_grid;
    if (!g) return fail("out of host memory");
    if (fill_desc(g->d, dtype, n_lat, n_lon, min_lon, max_lon, min_lat, max_lat)) { delete g; return 1; }
    g->device = device;
    g->d.z = dev_rows; g->d.ld = ld; g->d.row0 = static_cast<int>(row0); g->d.rows = static_cast<int>(rows);
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete g; return fail_cuda("cudaSetDevice", e); }
    if (make_streams(g)) { auvi_grid_destroy(g); return 2; }
    *out = g;
    return 0;
}

int auvi_grid_destroy(auvi_grid* g) {
    if (!g) return 0;
    cudaSetDevice(g->device);
    cudaDeviceSynchronize();
    for (auto& kv : g->axes) free_axis(kv.second, g->device);
    for (int k = 0; k < 2; ++k) {
        cached_free(g->d_rows[k], g->d_rows_bytes, g->device);
        if (g->ev_k0[k]) cudaEventDestroy(g->ev_k0[k]);
        if (g->ev_k1[k]) cudaEventDestroy(g->ev_k1[k]);
        if (g->ev_done[k]) cudaEventDestroy(g->ev_done[k]);
        if (g->st[k]) cudaStreamDestroy(g->st[k]);
    }
    cached_free(g->owned, g->owned_bytes, g->device);
    if (g->staging) {
        std::lock_guard<std::mutex> lk(g_staging_mu);
        g_staging_free.push_back(g->staging);
    }
    delete g;
    return 0;
}

// ---- point list ----------------------------------------------------------------------------------

int auvi_interp_points_device(auvi_grid* g, int method, const void* dev_pts, int64_t n,
                              int64_t stride_bytes, double* dev_out_elev,
                              int32_t* dev_sel, int32_t* dev_found, void* stream) {
    if (check_common(g, method)) return 1;
    if (n < 0) return fail("negative point count");
    if (n == 0) return 0;
    if (!dev_pts || !dev_out_elev) return fail("null device buffer");
    if (stride_bytes < 16 || stride_bytes % 8) return fail("stride_bytes must be a multiple of 8 and >= 16");
    if ((dev_sel == nullptr) != (dev_found == nullptr)) return fail("dev_sel and dev_found go together");
    if (g->d.row0 != 0 || g->d.rows != g->d.n_lat) return fail("point-list mode needs the whole grid resident");
    AUVI_CUDA(cudaSetDevice(g->device));
    GridDesc desc;
    int km = method;
    if (resolve_method(g, method, &desc, &km)) return 1;
    cudaError_t e = launch_points(desc, km, static_cast<const double*>(dev_pts), stride_bytes / 8, n,
                                  dev_out_elev, dev_sel, dev_found, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail_cuda("point-list launch", e);
    g_launches.fetch_add(1);
    return 0;
}

// On an error exit copies and kernels of earlier chunks may still be in flight on the staging buffers the next call
// reuses: wait for both streams before handing the error back.
static void quiesce(auvi_grid* g) {
    const std::string keep = t_error;
    for (int k = 0; k < 2; ++k) if (g && g->st[k]) cudaStreamSynchronize(g->st[k]);
    cudaGetLastError();
    t_error = keep;
}

static int interp_points_host(auvi_grid* g, int method, const void* host_pts, int64_t n,
                              int64_t stride_bytes, void* host_out, int64_t out_stride_bytes) {
    if (check_common(g, method)) return 1;
    if (n < 0) return fail("negative point count");
    if (n == 0) return 0;                                        // GridD.cu:96-98: nothing to do
    if (!host_pts || !host_out) return fail("null host buffer");
    if (stride_bytes < 16 || stride_bytes % 8) return fail("stride_bytes must be a multiple of 8 and >= 16");
    if (out_stride_bytes < 8 || out_stride_bytes % 8) return fail("out_stride_bytes must be a multiple of 8");
    if (g->d.row0 != 0 || g->d.rows != g->d.n_lat) return fail("point-list mode needs the whole grid resident");
    AUVI_CUDA(cudaSetDevice(g->device));
    if (ensure_point_staging(g)) return 2;
    GridDesc desc;
    int km = method;
    if (resolve_method(g, method, &desc, &km)) return 1;

    const int64_t sd = stride_bytes / 8, od = out_stride_bytes / 8;
    const double* src = static_cast<const double*>(host_pts);
    double* dst = static_cast<double*>(host_out);
    int64_t chunk = (n + 3) / 4;                                   // ~4 stages for a small batch, 1 Mi points at most
    chunk = chunk < kPointChunkMin ? kPointChunkMin : (chunk > kPointChunk ? kPointChunk : chunk);
    const int64_t n_chunks = (n + chunk - 1) / chunk;
    float ms_total = 0.f;
    auto drain = [&](int64_t c) -> int {                         // chunk c: wait, unpack, account
        const int b = static_cast<int>(c & 1);
        const int64_t lo = c * chunk, cnt = (n - lo < chunk) ? n - lo : chunk;
        AUVI_CUDA(cudaEventSynchronize(g->ev_done[b]));
        const double* r = g->h_out[b];
        double* const d0 = dst + lo * od;
        HostPool::get().for_range(cnt, pack_threads(cnt), 1024, [&](int64_t k0, int64_t k1) {
            for (int64_t k = k0; k < k1; ++k) d0[k * od] = r[k];
        });
        float ms = 0.f;
        AUVI_CUDA(cudaEventElapsedTime(&ms, g->ev_k0[b], g->ev_k1[b]));
        ms_total += ms;
        return 0;
    };
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int b = static_cast<int>(c & 1);
        const int64_t lo = c * chunk, cnt = (n - lo < chunk) ? n - lo : chunk;
        if (c >= 2 && drain(c - 2)) return 2;                     // buffer b is free again after this
        double* pin = g->h_in[b];
        const double* s = src + lo * sd;
        HostPool::get().for_range(cnt, pack_threads(cnt), 1024, [&](int64_t k0, int64_t k1) {
            for (int64_t k = k0; k < k1; ++k) { pin[2 * k] = s[k * sd]; pin[2 * k + 1] = s[k * sd + 1]; }
        });
        cudaStream_t st = g->st[b];
        AUVI_CUDA(cudaMemcpyAsync(g->d_in[b], pin, sizeof(double) * 2 * cnt, cudaMemcpyHostToDevice, st));
        AUVI_CUDA(cudaEventRecord(g->ev_k0[b], st));
        cudaError_t e = launch_points(desc, km, g->d_in[b], 2, cnt, g->d_out[b], nullptr, nullptr, st);
        if (e != cudaSuccess) return fail_cuda("point-list launch", e);
        g_launches.fetch_add(1);
        AUVI_CUDA(cudaEventRecord(g->ev_k1[b], st));
        AUVI_CUDA(cudaMemcpyAsync(g->h_out[b], g->d_out[b], sizeof(double) * cnt, cudaMemcpyDeviceToHost, st));
        AUVI_CUDA(cudaEventRecord(g->ev_done[b], st));
    }
    for (int64_t c = (n_chunks >= 2 ? n_chunks - 2 : 0); c < n_chunks; ++c)
        if (drain(c)) return 2;
    g->last_ms = ms_total;
    return 0;
}

int auvi_interp_points(auvi_grid* g, int method, const void* host_pts, int64_t n,
                       int64_t stride_bytes, void* host_out, int64_t out_stride_bytes) {
    const int rc = interp_points_host(g, method, host_pts, n, stride_bytes, host_out, out_stride_bytes);
    if (rc) quiesce(g);
    return rc;
}

// ---- lattice ---------------------------------------------------------------------------------------

int auvi_lattice_dims(const auvi_grid* g, int axis_kind, int f_lat, int f_lon, int64_t* out_rows, int64_t* out_cols) {
    if (!g) return fail("null grid handle");
    if (axis_kind != AUVI_AXIS_EXPANDED && axis_kind != AUVI_AXIS_NODES) return fail("unknown axis kind");
    if (f_lat < 1 || f_lon < 1) return fail("lattice factor must be >= 1");
    if (axis_kind == AUVI_AXIS_NODES && (f_lat != 1 || f_lon != 1)) return fail("AUVI_AXIS_NODES requires factor 1");
    if (out_rows) *out_rows = static_cast<int64_t>(f_lat) * (g->d.n_lat - 1) + 1;
    if (out_cols) *out_cols = static_cast<int64_t>(f_lon) * (g->d.n_lon - 1) + 1;
    return 0;
}

int auvi_lattice_device(auvi_grid* g, int method, int axis_kind, int f_lat, int f_lon, int fill,
                        int64_t row_begin, int64_t row_end, void* dev_out, int64_t out_ld,
                        int32_t* dev_sel9, void* stream) {
    if (check_common(g, method)) return 1;
    int64_t rows = 0, cols = 0;
    if (auvi_lattice_dims(g, axis_kind, f_lat, f_lon, &rows, &cols)) return 1;
    if (fill && (axis_kind != AUVI_AXIS_NODES)) return fail("fill mode requires AUVI_AXIS_NODES");
    if (row_begin < 0 || row_end > rows || row_begin > row_end) return fail("row range outside the lattice");
    if (row_begin == row_end) return 0;
    if (!dev_out) return fail("null device output");
    if (out_ld < cols) return fail("out_ld must be >= lattice columns");
    AUVI_CUDA(cudaSetDevice(g->device));
    AxisOwned *lat = nullptr, *lon = nullptr;
    if (get_axis(g, 1, axis_kind, f_lat, &lat) || get_axis(g, 0, axis_kind, f_lon, &lon)) return 1;
    LaunchInfo info;
    GridDesc desc;
    int km = method;
    if (resolve_method(g, method, &desc, &km)) return 1;
    cudaError_t e = launch_lattice(desc, km, lat->view(), lon->view(), row_begin, row_end, dev_out, out_ld,
                                   fill, dev_sel9, static_cast<cudaStream_t>(stream), &info);
    if (e == cudaErrorInvalidValue && (g->d.row0 != 0 || g->d.rows != g->d.n_lat))
        return fail("row range needs grid rows outside the resident slab (halo too small)");
    if (e != cudaSuccess) return fail_cuda("lattice launch", e);
    g_launches.fetch_add(info.launches);
    g->last_tma = info.used_tma;
    g->last_window = info.used_window;
    return 0;
}

int auvi_lattice(auvi_grid* g, int method, int axis_kind, int f_lat, int f_lon, int fill,
                 int64_t row_begin, int64_t row_end, void* host_out) {
    if (check_common(g, method)) return 1;
    int64_t rows = 0, cols = 0;
    if (auvi_lattice_dims(g, axis_kind, f_lat, f_lon, &rows, &cols)) return 1;
    if (row_begin < 0 || row_end > rows || row_begin > row_end) return fail("row range outside the lattice");
    if (row_begin == row_end) return 0;
    if (!host_out) return fail("null host output");
    AUVI_CUDA(cudaSetDevice(g->device));
    const size_t es = g->d.dtype == AUVI_F64 ? 8 : 4;
    const int64_t row_bytes = cols * static_cast<int64_t>(es);
    // Pageable destination (a std::vector, a numpy array): a device->host copy into it is staged by the driver at
    // 16-20 GB/s.  Bounce through our own pinned ring instead and move the rows out with a few host threads while the
    // next chunk is on the wire.  A pinned / registered destination is written directly.
    cudaPointerAttributes attr;
    bool pageable = true;
    if (cudaPointerGetAttributes(&attr, host_out) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
    else cudaGetLastError();
    static const int64_t chunk_knob = [] {                            // tuning knob, MiB
        const char* e = getenv("AUVI_LATTICE_CHUNK_MB");
        const long v = e ? atol(e) : 0;
        return v > 0 ? static_cast<int64_t>(v) << 20 : 0ll;
    }();
    const int64_t chunk_bytes = chunk_knob ? chunk_knob : (pageable ? kBounceBytes : kLatticeChunkBytes);
    int64_t chunk_rows = chunk_bytes / row_bytes;
    if (chunk_rows < 1) chunk_rows = 1;
    if (chunk_rows > row_end - row_begin) chunk_rows = row_end - row_begin;
    const size_t need = static_cast<size_t>(chunk_rows) * row_bytes;
    if (need > g->d_rows_bytes) {
        for (int k = 0; k < 2; ++k) { cached_free(g->d_rows[k], g->d_rows_bytes, g->device); g->d_rows[k] = nullptr; }
        g->d_rows_bytes = 0;
        for (int k = 0; k < 2; ++k) AUVI_CUDA(cached_malloc(&g->d_rows[k], need, g->device));
        g->d_rows_bytes = need;
    }
    char* bounce[2] = {nullptr, nullptr};
    size_t bounce_bytes = need;
    if (pageable) {
        if (acquire_bounce(bounce_bytes, bounce)) return 2;
    }
    float ms_total = 0.f;
    int64_t c = 0;
    int64_t lo_of[2] = {0, 0}, hi_of[2] = {0, 0};
    auto drain = [&](int b) -> int {                               // chunk in staging slot b: wait, move out, account
        AUVI_CUDA(cudaEventSynchronize(g->ev_done[b]));
        if (pageable) {
            char* const dst = static_cast<char*>(host_out) + (lo_of[b] - row_begin) * row_bytes;
            const int64_t bytes = (hi_of[b] - lo_of[b]) * row_bytes;
            HostPool& pool = HostPool::get();                          // a plain copy: bandwidth-bound, more threads help
            const char* const from = bounce[b];
            pool.for_range(bytes, bytes < (1 << 20) ? 1 : pool.size(), 4096, [&](int64_t lo, int64_t hi) {
                std::memcpy(dst + lo, from + lo, static_cast<size_t>(hi - lo));
            });
        }
        float ms = 0.f;
        AUVI_CUDA(cudaEventElapsedTime(&ms, g->ev_k0[b], g->ev_k1[b]));
        ms_total += ms;
        return 0;
    };
    int rc = 0;
    for (int64_t r = row_begin; r < row_end && !rc; r += chunk_rows, ++c) {
        const int b = static_cast<int>(c & 1);
        const int64_t r_hi = (r + chunk_rows < row_end) ? r + chunk_rows : row_end;
        cudaStream_t st = g->st[b];
        if (c >= 2 && (rc = drain(b))) break;                     // staging slot b is free again after this
        lo_of[b] = r; hi_of[b] = r_hi;
        cudaError_t e = cudaEventRecord(g->ev_k0[b], st);
        if (e == cudaSuccess && auvi_lattice_device(g, method, axis_kind, f_lat, f_lon, fill, r, r_hi, g->d_rows[b], cols, nullptr, st)) { rc = 2; break; }
        if (e == cudaSuccess) e = cudaEventRecord(g->ev_k1[b], st);
        char* const target = pageable ? bounce[b] : static_cast<char*>(host_out) + (r - row_begin) * row_bytes;
        if (e == cudaSuccess) e = cudaMemcpyAsync(target, g->d_rows[b], static_cast<size_t>(r_hi - r) * row_bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaEventRecord(g->ev_done[b], st);
        if (e != cudaSuccess) { rc = fail_cuda("lattice copy", e); break; }
    }
    for (int64_t k = (c >= 2 ? c - 2 : 0); k < c && !rc; ++k) rc = drain(static_cast<int>(k & 1));
    if (rc) quiesce(g);
    if (pageable) { cudaDeviceSynchronize(); release_bounce(bounce, bounce_bytes); }
    g->last_ms = ms_total;
    return rc;
}

// ---- peer memory: gather without a collective -------------------------------------------------------------------
// One process per GPU.  The consumer exports its result buffer, the producers map it and pass the mapped address as
// `dev_out` of auvi_lattice_device: the kernel's stores then travel over NVLink themselves.

namespace {
std::mutex g_peer_mu;
std::map<void*, void*> g_peer_base;                   // pointer handed out by auvi_peer_open -> base of the IPC mapping
std::map<uintptr_t, size_t> g_peer_range;             // base of an IPC mapping -> its size
}

extern "C++" bool auvi::output_is_peer_memory(const void* p) {
    {
        std::lock_guard<std::mutex> lk(g_peer_mu);
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        auto it = g_peer_range.upper_bound(a);
        if (it != g_peer_range.begin()) { --it; if (a < it->first + it->second) return true; }
    }
    cudaPointerAttributes attr;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaPointerGetAttributes(&attr, p) == cudaSuccess)
        return attr.type == cudaMemoryTypeDevice && attr.device != dev;      // a peer device of this process
    cudaGetLastError();
    return false;
}

int auvi_peer_export(const void* dev_ptr, unsigned char* handle72) {
    if (!dev_ptr || !handle72) return fail("null argument");
    // the IPC handle names the whole allocation: keep the offset of dev_ptr inside it next to the handle
    typedef CUresult (*range_fn)(CUdeviceptr*, size_t*, CUdeviceptr);
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) {
        cudaGetLastError();
        return fail("cuMemGetAddressRange is not available");
    }
    CUdeviceptr base = 0;
    size_t size = 0;
    if (reinterpret_cast<range_fn>(sym)(&base, &size, reinterpret_cast<CUdeviceptr>(dev_ptr)) != CUDA_SUCCESS)
        return fail("not a device allocation");
    cudaIpcMemHandle_t h;
    AUVI_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
    static_assert(sizeof h == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(handle72, &h, 64);
    const int64_t off = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(dev_ptr) - base);
    std::memcpy(handle72 + 64, &off, 8);
    return 0;
}

int auvi_peer_open(const unsigned char* handle72, void** out_ptr) {
    if (!handle72 || !out_ptr) return fail("null argument");
    cudaIpcMemHandle_t h;
    int64_t off = 0;
    std::memcpy(&h, handle72, 64);
    std::memcpy(&off, handle72 + 64, 8);
    void* base = nullptr;
    AUVI_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *out_ptr = static_cast<char*>(base) + off;
    size_t size = 0;
    {   // the extent of the mapping, so that kernels can tell that an output pointer is peer memory
        typedef CUresult (*range_fn)(CUdeviceptr*, size_t*, CUdeviceptr);
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        CUdeviceptr b = 0;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q) == cudaSuccess && sym &&
            reinterpret_cast<range_fn>(sym)(&b, &size, reinterpret_cast<CUdeviceptr>(base)) != CUDA_SUCCESS) size = 0;
        else cudaGetLastError();
    }
    std::lock_guard<std::mutex> lk(g_peer_mu);
    g_peer_base[*out_ptr] = base;
    if (size) g_peer_range[reinterpret_cast<uintptr_t>(base)] = size;
    return 0;
}

int auvi_peer_close(void* ptr) {
    if (!ptr) return 0;
    void* base = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_peer_mu);
        auto it = g_peer_base.find(ptr);
        if (it == g_peer_base.end()) return fail("not a pointer from auvi_peer_open");
        base = it->second;
        g_peer_base.erase(it);
        g_peer_range.erase(reinterpret_cast<uintptr_t>(base));
    }
    AUVI_CUDA(cudaIpcCloseMemHandle(base));
    return 0;
}

// ---- opt-in: fitted variogram ---------------------------------------------------------------
// Exponential model gamma(h) = c0 + c1 * (1 - exp(-h / a)), h in degrees as in GridH.cpp:371-380, fitted to the empirical
// semivariances of the grid at lags 1, 2, 4, 8 cells along both axes (device reduction: metrics.cu): for each candidate
// range a = h_max * 2^(m-2), m = 0..9, weighted least squares in (c0, c1) with the pair counts as weights (c0 clamped at
// 0, then least squares through the origin); the candidate with the smallest residual wins.  oracle: orc_fit_variogram.
int auvi_variogram_fit_from_sums(const double* sums16, double lon_step, double lat_step, double* out3) {
    double h[8], gam[8], w[8];
    int n = 0;
    double h_max = 0.0;
    for (int a = 0; a < 2; ++a)
        for (int l = 0; l < 4; ++l) {
            const double ss = sums16[a * 8 + 2 * l], cnt = sums16[a * 8 + 2 * l + 1];
            if (!(cnt > 0.0)) continue;
            h[n] = static_cast<double>(1 << l) * std::fabs(a == 0 ? lon_step : lat_step);
            gam[n] = ss / (2.0 * cnt);
            w[n] = cnt;
            if (h[n] > h_max) h_max = h[n];
            ++n;
        }
    if (n < 2 || !(h_max > 0.0)) return fail("variogram fit: fewer than two lags have pairs of valid cells");
    double best_r = 0.0, best[3] = {0.0, 0.0, 0.0};
    bool have = false;
    for (int m = 0; m < 10; ++m) {
        const double a = h_max * std::ldexp(1.0, m - 2);
        double sw = 0, sf = 0, sff = 0, sg = 0, sfg = 0;
        double f[8];
        for (int k = 0; k < n; ++k) {
            f[k] = 1.0 - std::exp(-h[k] / a);
            sw += w[k]; sf += w[k] * f[k]; sff += w[k] * f[k] * f[k]; sg += w[k] * gam[k]; sfg += w[k] * f[k] * gam[k];
        }
        const double det = sw * sff - sf * sf;
        double c1 = det != 0.0 ? (sw * sfg - sf * sg) / det : 0.0;
        double c0 = (sg - c1 * sf) / sw;
        if (!(c0 >= 0.0) || det == 0.0) { c0 = 0.0; c1 = sff > 0.0 ? sfg / sff : 0.0; }
        if (!(c1 > 0.0) || !std::isfinite(c1) || !std::isfinite(c0)) continue;
        double r = 0.0;
        for (int k = 0; k < n; ++k) { const double e = gam[k] - c0 - c1 * f[k]; r += w[k] * e * e; }
        if (!have || r < best_r) { have = true; best_r = r; best[0] = c0; best[1] = c1; best[2] = a; }
    }
    if (!have) return fail("variogram fit: no candidate range gives a positive sill (constant grid?)");
    out3[0] = best[0]; out3[1] = best[1]; out3[2] = best[2];
    return 0;
}

int auvi_grid_fit_variogram(auvi_grid* g, double* out_c0_c1_range) {
    if (!g) return fail("null grid handle");
    if (!out_c0_c1_range) return fail("null output");
    if (g->d.row0 != 0 || g->d.rows != g->d.n_lat)
        return fail("the variogram fit needs the whole grid resident (on a row slab, pass the parameters with auvi_grid_set_variogram)");
    AUVI_CUDA(cudaSetDevice(g->device));
    void* scratch = nullptr;
    const size_t bytes = (variogram_scratch_bytes() + 255) / 256 * 256 + 256;
    AUVI_CUDA(cached_malloc(&scratch, bytes, g->device));
    double* d_sums = reinterpret_cast<double*>(static_cast<char*>(scratch) + bytes - 256);
    LaunchInfo info;
    double sums[16];
    cudaError_t e = launch_variogram_sums(g->d, scratch, d_sums, g->st[0], &info);
    if (e == cudaSuccess) e = cudaMemcpyAsync(sums, d_sums, sizeof sums, cudaMemcpyDeviceToHost, g->st[0]);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g->st[0]);
    if (e != cudaSuccess) { uncached_free(scratch); return fail_cuda("variogram sums", e); }
    cached_free(scratch, bytes, g->device);
    g_launches.fetch_add(info.launches);
    double p[3];
    if (auvi_variogram_fit_from_sums(sums, g->d.lon_step, g->d.lat_step, p)) return 1;
    g->vg_c0 = p[0]; g->vg_c1 = p[1]; g->vg_range = p[2]; g->vg_valid = true;
    out_c0_c1_range[0] = p[0]; out_c0_c1_range[1] = p[1]; out_c0_c1_range[2] = p[2];
    return 0;
}

int auvi_grid_set_variogram(auvi_grid* g, double c0, double c1, double range) {
    if (!g) return fail("null grid handle");
    if (!(c0 >= 0.0) || !(c1 > 0.0) || !(range > 0.0)) return fail("variogram needs c0 >= 0, c1 > 0, range > 0");
    g->vg_c0 = c0; g->vg_c1 = c1; g->vg_range = range; g->vg_valid = true;
    return 0;
}

// ---- metrics ---------------------------------------------------------------------------------------

int auvi_error_metrics_device(const void* dev_truth, const void* dev_est, int dtype, int64_t n,
                              double* out3, int64_t* out_nan, void* stream) {
    if (dtype != AUVI_F64 && dtype != AUVI_F32) return fail("dtype must be AUVI_F64 or AUVI_F32");
    if (n <= 0) return fail("Error: Reference or interpolated points vector is empty or sizes do not match.");
    if (!dev_truth || !dev_est || !out3) return fail("null buffer");
    if (auvi_device_count() <= 0) return fail("no CUDA device: libauvi has no CPU fallback");
    MetricsScratch ms;
    if (ms.open()) return 2;
    LaunchInfo info;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    cudaError_t e = launch_metrics(dev_truth, dev_est, dtype, n, ms.scratch, ms.result5, st, &info);
    double h[5];
    if (ms.close(e, st, h)) return 2;
    g_launches.fetch_add(info.launches);
    out3[0] = h[0] / static_cast<double>(n);                       // error_calculator.cpp:17
    out3[1] = std::sqrt(h[1] / static_cast<double>(n));            // :32
    out3[2] = h[2];
    if (out_nan) *out_nan = static_cast<int64_t>(h[3]);
    return 0;
}

int auvi_fill_metrics_device(auvi_grid* masked, const void* dev_filled, int64_t filled_ld, const void* dev_truth,
                             int64_t truth_ld, int64_t row_begin, int64_t row_end, double* out3, int64_t* out_nan,
                             int64_t* out_count, void* stream) {
    if (!masked) return fail("null grid handle");
    if (!dev_filled || !dev_truth || !out3) return fail("null buffer");
    const GridDesc& d = masked->d;
    if (row_begin < d.row0 || row_end > d.row0 + d.rows || row_begin >= row_end) return fail("row range outside the resident slab");
    if (filled_ld < d.n_lon || truth_ld < d.n_lon) return fail("row pitch must be >= n_lon");
    AUVI_CUDA(cudaSetDevice(masked->device));
    MetricsScratch ms;
    if (ms.open()) return 2;
    LaunchInfo info;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t es = d.dtype == AUVI_F64 ? 8 : 4;
    const char* m0 = static_cast<const char*>(d.z) + static_cast<size_t>(row_begin - d.row0) * d.ld * es;
    cudaError_t e = launch_metrics_masked(m0, d.ld, dev_filled, filled_ld, dev_truth, truth_ld, d.dtype, row_end - row_begin,
                                          d.n_lon, ms.scratch, ms.result5, st, &info);
    double h[5];
    if (ms.close(e, st, h)) return 2;
    g_launches.fetch_add(info.launches);
    const double n = h[4];
    if (out_count) *out_count = static_cast<int64_t>(n);
    if (out_nan) *out_nan = static_cast<int64_t>(h[3]);
    if (n <= 0.0) return fail("Error: Reference or interpolated points vector is empty or sizes do not match.");
    out3[0] = h[0] / n; out3[1] = std::sqrt(h[1] / n); out3[2] = h[2];
    return 0;
}

// ---- Grid-B data preparation -------------------------------------------------

int auvi_grid_create_raw(const void* host_raw, int nc_type, int big_endian, int flip_rows, double scale, double offset,
                         int dtype, int64_t n_lat, int64_t n_lon, double min_lon, double max_lon, double min_lat,
                         double max_lat, int device, auvi_grid** out) {
    if (!out) return fail("null output handle");
    *out = nullptr;
    if (!host_raw) return fail("null host buffer");
    if (nc_type < 3 || nc_type > 6) return fail("raw element type must be NC_SHORT(3), NC_INT(4), NC_FLOAT(5) or NC_DOUBLE(6)");
    if (auvi_device_count() <= 0) return fail("no CUDA device: libauvi has no CPU fallback");
    auvi_grid* g = new (std::nothrow) auvi_grid;
    if (!g) return fail("out of host memory");
    if (fill_desc(g->d, dtype, n_lat, n_lon, min_lon, max_lon, min_lat, max_lat)) { delete g; return 1; }
    g->device = device;
    const size_t es = dtype == AUVI_F64 ? 8 : 4, rs = nc_type == 3 ? 2 : (nc_type == 6 ? 8 : 4);
    const int64_t ld = (n_lon * es + 15) / 16 * 16 / es;
    const size_t raw_bytes = static_cast<size_t>(n_lat) * n_lon * rs;
    void* d_raw = nullptr;
    cudaError_t e = cudaSetDevice(device);
    g->owned_bytes = static_cast<size_t>(ld) * n_lat * es;
    if (e == cudaSuccess) e = cached_malloc(&g->owned, g->owned_bytes, device);
    if (e == cudaSuccess) e = cached_malloc(&d_raw, raw_bytes, device);
    if (e == cudaSuccess) e = cudaMemcpy(d_raw, host_raw, raw_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_decode_raw(d_raw, nc_type, big_endian, flip_rows, scale, offset, g->d.n_lat, g->d.n_lon,
                                                g->owned, ld, dtype, nullptr);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (d_raw) cached_free(d_raw, raw_bytes, device);
    if (e != cudaSuccess) { uncached_free(g->owned); delete g; return fail_cuda("raw grid upload", e); }
    g_launches.fetch_add(1);
    g->d.z = g->owned; g->d.ld = ld; g->d.row0 = 0; g->d.rows = g->d.n_lat;
    if (make_streams(g)) { auvi_grid_destroy(g); return 2; }
    *out = g;
    return 0;
}

int auvi_csv_dims(const char* text, int64_t n_bytes, int64_t* n_rows, int64_t* n_cols) {
    if (!text || n_bytes < 0) return fail("null CSV text");
    while (n_bytes > 0 && (text[n_bytes - 1] == '\n' || text[n_bytes - 1] == '\r' || text[n_bytes - 1] == ' ')) --n_bytes;
    if (n_bytes == 0) return fail("Grid data is empty.");                       // test_gebco.cpp:127-128
    int64_t cols = 1, rows = 1;
    const char* first_nl = static_cast<const char*>(memchr(text, '\n', static_cast<size_t>(n_bytes)));
    const int64_t first_len = first_nl ? first_nl - text : n_bytes;
    for (int64_t k = 0; k < first_len; ++k) cols += text[k] == ',';
    for (const char* p = first_nl; p; p = static_cast<const char*>(memchr(p + 1, '\n', static_cast<size_t>(text + n_bytes - p - 1)))) ++rows;
    if (n_rows) *n_rows = rows;
    if (n_cols) *n_cols = cols;
    return 0;
}

int auvi_grid_create_csv(const char* text, int64_t n_bytes, int dtype, double min_lon, double max_lon, double min_lat,
                         double max_lat, int device, auvi_grid** out) {
    if (!out) return fail("null output handle");
    *out = nullptr;
    int64_t n_rows = 0, n_cols = 0;
    if (auvi_csv_dims(text, n_bytes, &n_rows, &n_cols)) return 1;
    while (n_bytes > 0 && (text[n_bytes - 1] == '\n' || text[n_bytes - 1] == '\r' || text[n_bytes - 1] == ' ')) --n_bytes;
    if (auvi_device_count() <= 0) return fail("no CUDA device: libauvi has no CPU fallback");
    auvi_grid* g = new (std::nothrow) auvi_grid;
    if (!g) return fail("out of host memory");
    if (fill_desc(g->d, dtype, n_rows, n_cols, min_lon, max_lon, min_lat, max_lat)) { delete g; return 1; }
    g->device = device;
    const size_t es = dtype == AUVI_F64 ? 8 : 4;
    const int64_t ld = (n_cols * es + 15) / 16 * 16 / es, n_fields = n_rows * n_cols, n_text = n_bytes + 1;
    const size_t n_blocks = csv_block_count(n_text);
    // device scratch: text (+ the final newline), per-block counts and offsets, delimiter positions, slow-field list
    struct Blk { void* p = nullptr; size_t bytes = 0; };
    Blk b_text{nullptr, static_cast<size_t>(n_text)}, b_counts{nullptr, n_blocks * sizeof(int)}, b_off{nullptr, n_blocks * sizeof(int64_t)},
        b_pos{nullptr, static_cast<size_t>(n_fields) * sizeof(int64_t)}, b_slow{nullptr, static_cast<size_t>(n_fields) * sizeof(int64_t)},
        b_misc{nullptr, 4096};
    Blk* all[] = {&b_text, &b_counts, &b_off, &b_pos, &b_slow, &b_misc};
    auto release = [&](bool also_grid) {
        for (Blk* b : all) if (b->p) cached_free(b->p, b->bytes, device);
        if (also_grid) { if (g->owned) cached_free(g->owned, g->owned_bytes, device); delete g; }
    };
    cudaError_t e = cudaSetDevice(device);
    g->owned_bytes = static_cast<size_t>(ld) * n_rows * es;
    if (e == cudaSuccess) e = cached_malloc(&g->owned, g->owned_bytes, device);
    for (Blk* b : all) if (e == cudaSuccess) e = cached_malloc(&b->p, b->bytes, device);
    if (e != cudaSuccess) { release(true); return fail_cuda("CSV scratch allocation", e); }
    char* d_text = static_cast<char*>(b_text.p);
    int64_t* d_total = static_cast<int64_t*>(b_misc.p);
    unsigned* d_status = reinterpret_cast<unsigned*>(d_total + 1);
    unsigned long long* d_nslow = reinterpret_cast<unsigned long long*>(d_total + 2);
    const char nl = '\n';
    e = cudaMemcpy(d_text, text, static_cast<size_t>(n_bytes), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_text + n_bytes, &nl, 1, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(b_misc.p, 0, 64);
    if (e == cudaSuccess) e = launch_csv_index(d_text, n_text, static_cast<int*>(b_counts.p), static_cast<int64_t*>(b_off.p), d_total,
                                               static_cast<int64_t*>(b_pos.p), n_fields, nullptr);
    int64_t h_misc[3] = {0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(h_misc, b_misc.p, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { release(true); return fail_cuda("CSV delimiter index", e); }
    if (h_misc[0] != n_fields) {
        release(true);
        return fail("CSV rows do not all have the same number of fields (" + std::to_string(h_misc[0]) + " fields for " +
                    std::to_string(n_rows) + " x " + std::to_string(n_cols) + ")");
    }
    e = launch_csv_parse(d_text, static_cast<const int64_t*>(b_pos.p), n_fields, static_cast<int>(n_cols), g->owned, ld, dtype, d_status,
                         d_nslow, static_cast<int64_t*>(b_slow.p), n_fields, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(h_misc, b_misc.p, 24, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { release(true); return fail_cuda("CSV parse", e); }
    g_launches.fetch_add(4);
    const unsigned status = static_cast<unsigned>(h_misc[1] & 0xffffffff);
    if (status & 1u) { release(true); return fail("CSV rows do not all have the same number of fields"); }
    if (status & 2u) { release(true); return fail("CSV holds a field that is not a number (or an empty field)"); }
    const int64_t n_slow = h_misc[2];
    if (n_slow > 0) {
        // fields beyond the exact fast path (more than 19 digits, |exponent| > 22 ...): strtod on the host, like the reference
        std::vector<int64_t> idx(static_cast<size_t>(n_slow)), pos(static_cast<size_t>(n_fields));
        std::vector<double> val(static_cast<size_t>(n_slow));
        e = cudaMemcpy(idx.data(), b_slow.p, sizeof(int64_t) * n_slow, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(pos.data(), b_pos.p, sizeof(int64_t) * n_fields, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) {
            for (int64_t k = 0; k < n_slow; ++k) {
                const int64_t f = idx[k], a = f ? pos[f - 1] + 1 : 0, b = pos[f];
                const std::string cell(text + a, static_cast<size_t>((b < n_bytes ? b : n_bytes) - a));
                val[k] = std::strtod(cell.c_str(), nullptr);
            }
            void* d_val = nullptr;
            e = cached_malloc(&d_val, sizeof(double) * n_slow, device);
            if (e == cudaSuccess) e = cudaMemcpy(d_val, val.data(), sizeof(double) * n_slow, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) e = launch_csv_patch(g->owned, ld, static_cast<int>(n_cols), dtype, static_cast<const int64_t*>(b_slow.p),
                                                       static_cast<const double*>(d_val), n_slow, nullptr);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            if (d_val) cached_free(d_val, sizeof(double) * n_slow, device);
            g_launches.fetch_add(1);
        }
        if (e != cudaSuccess) { release(true); return fail_cuda("CSV slow-field patch", e); }
    }
    e = cudaDeviceSynchronize();
    release(false);
    if (e != cudaSuccess) { uncached_free(g->owned); delete g; return fail_cuda("CSV parse", e); }
    g->d.z = g->owned; g->d.ld = ld; g->d.row0 = 0; g->d.rows = g->d.n_lat;
    if (make_streams(g)) { auvi_grid_destroy(g); return 2; }
    *out = g;
    return 0;
}

int auvi_grid_mask_cells(auvi_grid* g, const int64_t* host_flat_idx, int64_t n, void* host_truth) {
    if (!g) return fail("null grid handle");
    if (n < 0) return fail("negative cell count");
    if (n == 0) return 0;
    if (!host_flat_idx) return fail("null index list");
    const int64_t total = static_cast<int64_t>(g->d.n_lat) * g->d.n_lon;
    for (int64_t k = 0; k < n; ++k)
        if (host_flat_idx[k] < 0 || host_flat_idx[k] >= total) return fail("cell index outside the grid");
    AUVI_CUDA(cudaSetDevice(g->device));
    const size_t es = g->d.dtype == AUVI_F64 ? 8 : 4;
    void *d_idx = nullptr, *d_truth = nullptr;
    cudaError_t e = cached_malloc(&d_idx, sizeof(int64_t) * n, g->device);
    if (e == cudaSuccess && host_truth) e = cached_malloc(&d_truth, es * n, g->device);
    if (e == cudaSuccess) e = cudaMemcpy(d_idx, host_flat_idx, sizeof(int64_t) * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && d_truth) e = cudaMemset(d_truth, 0xff, es * n);          // cells of other slabs stay NaN
    if (e == cudaSuccess) e = launch_mask_cells(g->d, static_cast<const int64_t*>(d_idx), n, d_truth, nullptr);
    if (e == cudaSuccess && d_truth) e = cudaMemcpy(host_truth, d_truth, es * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (d_idx) cached_free(d_idx, sizeof(int64_t) * n, g->device);
    if (d_truth) cached_free(d_truth, es * n, g->device);
    if (e != cudaSuccess) return fail_cuda("mask cells", e);
    g_launches.fetch_add(1);
    return 0;
}

int auvi_grid_mask_hash(auvi_grid* g, double fraction, uint64_t seed, int64_t* out_masked, void* stream) {
    if (!g) return fail("null grid handle");
    if (!(fraction >= 0.0 && fraction <= 1.0)) return fail("mask fraction must lie in [0,1]");
    AUVI_CUDA(cudaSetDevice(g->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void* d_cnt = nullptr;
    if (out_masked) {
        AUVI_CUDA(cached_malloc(&d_cnt, 4096, g->device));
        AUVI_CUDA(cudaMemsetAsync(d_cnt, 0, 8, st));
    }
    cudaError_t e = launch_mask_hash(g->d, fraction, seed, static_cast<unsigned long long*>(d_cnt), st);
    if (e == cudaSuccess && out_masked) {
        unsigned long long h = 0;
        e = cudaMemcpyAsync(&h, d_cnt, 8, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        *out_masked = static_cast<int64_t>(h);
    }
    if (d_cnt) cached_free(d_cnt, 4096, g->device);
    if (e != cudaSuccess) return fail_cuda("mask hash", e);
    g_launches.fetch_add(1);
    return 0;
}

int auvi_grid_read(auvi_grid* g, int64_t row_begin, int64_t row_end, void* host_out) {
    if (!g) return fail("null grid handle");
    if (row_begin < g->d.row0 || row_end > g->d.row0 + g->d.rows || row_begin > row_end) return fail("row range outside the resident slab");
    if (row_begin == row_end) return 0;
    if (!host_out) return fail("null host buffer");
    AUVI_CUDA(cudaSetDevice(g->device));
    const size_t es = g->d.dtype == AUVI_F64 ? 8 : 4;
    AUVI_CUDA(cudaMemcpy2D(host_out, g->d.n_lon * es,
                           static_cast<const char*>(g->d.z) + static_cast<size_t>(row_begin - g->d.row0) * g->d.ld * es,
                           g->d.ld * es, g->d.n_lon * es, row_end - row_begin, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
