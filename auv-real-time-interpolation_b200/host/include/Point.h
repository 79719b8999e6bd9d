// Point.h -- the query/result record of the interpolation API (drop-in for the reference's
// code/include/Point.h:9-13): three doubles, 24 bytes, array-of-structs.  This layout is the wire
// format of GridD::batch*: libauvi reads {lon,lat} at a 24-byte stride and writes `elev` in place.
// Same include guard as the reference header so that either one may be seen first.
#ifndef POINT_H
#define POINT_H

struct Point {
    double lon;   // degrees east
    double lat;   // degrees north
    double elev;  // depth / elevation in metres (output of the batch calls)
};

static_assert(sizeof(Point) == 24, "Point is the 24-byte wire format of the batch calls");

#endif
