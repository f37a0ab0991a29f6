// Internal: pixel -> ECEF math shared by geo.cu (dunk_world_coordinates) and pipeline.cu (the pose stage builds
// its object points on the device).  geotransform::get_world_coordinates, feature_database/src/elevationdb.rs:64-104.
#pragma once
#include "ctx.h"

struct dunk_elevation {
    dunk_ctx* ctx = nullptr;
    double gt_dataset[6];
    double gt_elev_inv[6];
    int has_elevation = 0;
    int x_size = 0, y_size = 0;
    double* heights = nullptr;   // device, y_size x x_size (the `elevation` table in row-id order)
};

namespace dunk {

constexpr double kWgs84A = 6378137.0;
constexpr double kWgs84F = 1.0 / 298.257223563;
constexpr double kWgs84Es = 2 * kWgs84F - kWgs84F * kWgs84F;
constexpr double kDegToRad = 0.017453292519943296;

struct GeoParams {
    double gt[6], inv[6];
    int has_elev, x_size, y_size;
};

inline GeoParams make_geo_params(const dunk_elevation* e) {
    GeoParams p;
    for (int i = 0; i < 6; ++i) { p.gt[i] = e->gt_dataset[i]; p.inv[i] = e->gt_elev_inv[i]; }
    p.has_elev = e->has_elevation; p.x_size = e->x_size; p.y_size = e->y_size;
    return p;
}

__device__ __forceinline__ double round_half_away(double v) { return v >= 0 ? floor(v + 0.5) : ceil(v - 0.5); }

// reference-image pixel (x, y) -> ECEF metres; returns false (and NaN coordinates) when the elevation sample
// does not exist (diesel NotFound in the reference)
__device__ __forceinline__ bool world_point(const GeoParams& p, const double* __restrict__ heights, double x, double y,
                                            double (&out)[3]) {
    // GeoTransform::apply (no FMA contraction: GDAL is built without it on x86-64)
    const double gx = __dadd_rn(__dadd_rn(p.gt[0], __dmul_rn(x, p.gt[1])), __dmul_rn(y, p.gt[2]));
    const double gy = __dadd_rn(__dadd_rn(p.gt[3], __dmul_rn(x, p.gt[4])), __dmul_rn(y, p.gt[5]));
    double h = 0.0;
    bool found = true;
    if (p.has_elev) {
        const double ex = __dadd_rn(__dadd_rn(p.inv[0], __dmul_rn(gx, p.inv[1])), __dmul_rn(gy, p.inv[2]));
        const double ey = __dadd_rn(__dadd_rn(p.inv[3], __dmul_rn(gx, p.inv[4])), __dmul_rn(gy, p.inv[5]));
        // elevation::get_elevation: row id = round(y) * x_size + round(x) + 1 (f64::round, i32 arithmetic)
        const long long idx = (long long)round_half_away(ey) * p.x_size + (long long)round_half_away(ex);
        if (idx >= 0 && idx < (long long)p.x_size * p.y_size) h = heights[idx];
        else {
            h = nan("");
            found = false;
        }
    }
    // convert_coordinates(coordinates.1, coordinates.0, height): EPSG:4326 (lat, lon) -> EPSG:4978
    const double phi = gy * kDegToRad, lam = gx * kDegToRad;
    const double s = sin(phi), c = cos(phi);
    const double N = kWgs84A / sqrt(1.0 - kWgs84Es * s * s);
    out[0] = (N + h) * c * cos(lam);
    out[1] = (N + h) * c * sin(lam);
    out[2] = (N * (1.0 - kWgs84Es) + h) * s;
    return found;
}

}  // namespace dunk
