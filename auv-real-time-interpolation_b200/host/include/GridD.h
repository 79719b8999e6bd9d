// GridD.h -- device-resident bathymetry grid with the reference's public interface
// (code/include/GridD.h:21-95), implemented over the C ABI of libauvi.so (include/auvi.h).
//
// Drop-in contract: same class name, same constructor/method signatures and -- because the
// reference's drivers are compiled against the reference's own copy of this header -- the same
// data-member layout (GridD.h:23-28).  `d_grid` no longer points at raw grid memory: it carries
// the opaque auvi_grid* handle; nothing outside GridD.cpp ever dereferenced it.
// No <cuda_runtime.h> is needed to use this class.  The include guard equals the reference's.
#ifndef GRIDD_H
#define GRIDD_H

#include <vector>
#include "Point.h"

class GridD {
private:
    double* d_grid;                       // opaque libauvi handle (reference: device grid pointer)
    int num_lon, num_lat;
    double min_lon, max_lon;
    double min_lat, max_lat;
    double lon_step, lat_step;
    bool initialized;

    void initialize(const std::vector<std::vector<double>>& elevation_data);

public:
    // Argument order of the reference (GridD.h:49-51): longitude axis first, then latitude;
    // elevation_data[row][col], row 0 = min_latitude.
    GridD(double min_longitude, double max_longitude, int longitude_points,
          double min_latitude, double max_latitude, int latitude_points,
          const std::vector<std::vector<double>>& elevation_data);
    ~GridD();

    void cleanup();                       // idempotent (GridD.cu:86-92)

    // Each returns a copy of query_points with .elev replaced by the interpolated depth; NaN for
    // queries outside the bounds; the input itself when it is empty or the grid is not initialised.
    std::vector<Point> batchBilinearInterpolate(const std::vector<Point>& query_points);
    std::vector<Point> batchCubicInterpolate(const std::vector<Point>& query_points);
    std::vector<Point> batchOrdinaryKrigingInterpolate(const std::vector<Point>& query_points);

    double bilinearInterpolate(double lon, double lat);   // one-point batch (GridD.cu:239-245)
};

#endif
