// gebco_gapfill.cpp -- the reference's whole Grid-B procedure as one C++ program over libauvi's C ABI.
//
// What the reference spreads over a Python tool and a C++ driver --
//   code/subset_bathymetry.py   netCDF4 read, row flip, seeded removal of cells, three CSV files
//   code/test_gebco.cpp         CSV parse into vector<vector<double>>, 24-byte Point lists, three methods on CPU and
//                               GPU, MAE / RMSE / Max per method, result rows appended to TestingResults1.csv
// -- runs here without pandas, CSV text or Point lists: the GEBCO tile's int16 bytes go to the GPU as they lie in
// the NetCDF-3 file, the removed cells are the ones numpy.random.seed(42); choice(total, n, replace=False) picks, every
// removed cell is filled at its own node by the full-grid gap fill, and the error metrics are reduced on the device
// against the unmasked grid.  Output: one row per method in the column order of test_gebco.cpp:208-228
//   <impl>,<method>,<grid_points>,<batch_size>,<removal_fraction>,<time_ms>,<MAE>,<RMSE>,<Max>
// so the rows can be appended to the reference's results file.
//
// usage: gebco_gapfill <tile.nc> <removal_fraction> <min_lon> <max_lon> <min_lat> <max_lat> [seed=42] [device=0]
// (the bounds are the N/S/W/E numbers in the tile's file name: that is how test_gebco.cpp:132-133 was filled in)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include <cuda_runtime_api.h>

#include "auvi.h"

namespace {

[[noreturn]] void die(const char* where) {
    std::cerr << "Error: " << auvi_last_error() << " at " << where << std::endl;   // GridD.h:9-16 convention
    std::exit(1);
}
#define CHECK(call) do { if ((call) != 0) die(#call); } while (0)
#define CUDA_CHECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { std::cerr << "CUDA Error: " << cudaGetErrorString(e_) << " at " #call << std::endl; std::exit(1); } } while (0)

}  // namespace

int main(int argc, char** argv) {
    if (argc < 7) {
        std::cerr << "usage: " << argv[0] << " <tile.nc> <removal_fraction> <min_lon> <max_lon> <min_lat> <max_lat> [seed] [device]\n";
        return 2;
    }
    const double fraction = std::atof(argv[2]);
    const double min_lon = std::atof(argv[3]), max_lon = std::atof(argv[4]), min_lat = std::atof(argv[5]), max_lat = std::atof(argv[6]);
    const uint32_t seed = argc > 7 ? static_cast<uint32_t>(std::strtoul(argv[7], nullptr, 10)) : 42u;
    const int device = argc > 8 ? std::atoi(argv[8]) : 0;

    std::ifstream f(argv[1], std::ios::binary | std::ios::ate);
    if (!f) { std::cerr << "Error: unable to open " << argv[1] << std::endl; return 1; }
    std::vector<char> nc(static_cast<size_t>(f.tellg()));
    f.seekg(0);
    f.read(nc.data(), static_cast<std::streamsize>(nc.size()));

    auvi_nc_var v;
    CHECK(auvi_netcdf3_find(nc.data(), static_cast<int64_t>(nc.size()), "elevation", &v));
    if (v.ndims != 2) { std::cerr << "Error: 'elevation' is not a 2-D variable" << std::endl; return 1; }
    const int64_t n_lat = v.shape[0], n_lon = v.shape[1], total = n_lat * n_lon;
    std::cout << "Grid dimensions: " << n_lon << " x " << n_lat << std::endl;

    // the unmasked grid (truth) and the grid that gets masked: both decoded on the device from the file bytes
    auvi_grid *truth = nullptr, *masked = nullptr;
    CHECK(auvi_grid_create_raw(nc.data() + v.data_offset, v.nc_type, 1, /*flip_rows=*/1, v.scale_factor, v.add_offset,
                               AUVI_F64, n_lat, n_lon, min_lon, max_lon, min_lat, max_lat, device, &truth));
    CHECK(auvi_grid_create_raw(nc.data() + v.data_offset, v.nc_type, 1, 1, v.scale_factor, v.add_offset, AUVI_F64, n_lat,
                               n_lon, min_lon, max_lon, min_lat, max_lat, device, &masked));
    const int64_t n_remove = static_cast<int64_t>(static_cast<double>(total) * fraction);   // subset_bathymetry.py:35
    std::vector<int64_t> removed(static_cast<size_t>(n_remove));
    CHECK(auvi_legacy_choice(total, n_remove, seed, removed.data()));
    CHECK(auvi_grid_mask_cells(masked, removed.data(), n_remove, nullptr));
    std::cout << "Selected " << n_remove << " points for removal." << std::endl;

    CUDA_CHECK(cudaSetDevice(device));
    double *d_filled = nullptr, *d_truth = nullptr;
    CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_filled), sizeof(double) * total));
    CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_truth), sizeof(double) * total));
    {   // the truth grid as a dense device array (rows of n_lon)
        std::vector<double> h(static_cast<size_t>(total));
        CHECK(auvi_grid_read(truth, 0, n_lat, h.data()));
        CUDA_CHECK(cudaMemcpy(d_truth, h.data(), sizeof(double) * total, cudaMemcpyHostToDevice));
    }
    const struct { int id; const char* name; } methods[] = {{AUVI_BILINEAR, "Bilinear"}, {AUVI_CUBIC, "Cubic"},
                                                            {AUVI_KRIGING, "Kriging"}, {AUVI_NN, "Nearest"}, {AUVI_IDW, "IDW"}};
    for (const auto& m : methods) {
        CUDA_CHECK(cudaDeviceSynchronize());
        const auto t0 = std::chrono::high_resolution_clock::now();
        CHECK(auvi_lattice_device(masked, m.id, AUVI_AXIS_NODES, 1, 1, /*fill=*/1, 0, n_lat, d_filled, n_lon, nullptr, nullptr));
        CUDA_CHECK(cudaDeviceSynchronize());
        const auto t1 = std::chrono::high_resolution_clock::now();
        double err[3];
        int64_t n_nan = 0, count = 0;
        CHECK(auvi_fill_metrics_device(masked, d_filled, n_lon, d_truth, n_lon, 0, n_lat, err, &n_nan, &count, nullptr));
        const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        std::printf("GPU,%s,%lld,%lld,%g,%.3f,%.10g,%.10g,%.10g\n", m.name, static_cast<long long>(total),
                    static_cast<long long>(count), fraction, ms, err[0], err[1], err[2]);
        std::fprintf(stderr, "%s: %lld NaN outputs\n", m.name, static_cast<long long>(n_nan));
    }
    cudaFree(d_filled); cudaFree(d_truth);
    auvi_grid_destroy(masked);
    auvi_grid_destroy(truth);
    return 0;
}
