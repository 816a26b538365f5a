// gtp_common.cuh - geometry of the MODIS scan-wise tie-point interpolators.
//
// Everything the kernels need to know about one (coarse, fine, width) configuration:
// how many tie-point rows/columns a scan has, which tie-point quad a fine pixel
// belongs to and at which sub-pixel coordinate it sits inside that quad.  The
// reference realises these maps by replicating arrays
// (/root/reference/geotiepoints/_modis_interpolator.pyx:196-249 coordinates,
// :410-593 expansion); here they are closed-form integer functions evaluated per
// pixel.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define GTP_HD __host__ __device__ __forceinline__
#else
#define GTP_HD inline
#endif

#define GTP_EARTH_RADIUS 6370997.0   // metres, _modis_utils.pyx:23
#define GTP_RK 6370.997              // km,     _modis_interpolator.pyx:10
#define GTP_H 709.0                  // km,     _modis_interpolator.pyx:12
#define GTP_PI 3.14159265358979323846
#define GTP_DEG2RAD (GTP_PI / 180.0)
#define GTP_RAD2DEG (180.0 / GTP_PI)

// per-axis scheme and out-of-range handling of the generic grid interpolators (include/gtp_b200.h has the same)
#ifndef GTP_GRID_NEAREST
#define GTP_GRID_NEAREST 0
#define GTP_GRID_LINEAR 1
#define GTP_GRID_CUBIC 3
#define GTP_GRID_OOB_EXTRAPOLATE 0
#define GTP_GRID_OOB_FILL 1
#define GTP_GRID_OOB_CLAMP 2
#endif

struct CviirsCfg {
    int L;    // tie-point rows per scan: 10 (1 km) or 2 (5 km)   :111-115
    int W;    // tie-point columns: 1354 | 271 | 270              :113-119
    int km;   // coarse pixel size in km (1 | 5) = "scan_width"   :120, :313
    int p;    // fine pixels per 1 km                             :122
    int f;    // fine pixels per coarse pixel                     :123
    int Wf;   // fine columns = 1354 * p                          :124
    int n;    // fine rows per scan = L * f                       :140
    int F5;   // 5 km only: 2 * p partial columns left and right  :128
};

// 0 ok; 1 unsupported resolution; 2 unsupported 5 km width (:33-36)
GTP_HD int gtp_make_cfg(int coarse_res, int fine_res, int coarse_width, CviirsCfg* c)
{
    if (coarse_res == 1000) { c->L = 10; c->W = 1354; }
    else if (coarse_res == 5000) {
        c->L = 2;
        c->W = coarse_width ? coarse_width : 271;
        if (c->W != 270 && c->W != 271) return 2;
    } else return 1;
    if (fine_res != 1000 && fine_res != 500 && fine_res != 250) return 1;
    if (fine_res >= coarse_res) return 1;
    c->km = coarse_res / 1000;
    c->p = 1000 / fine_res;
    c->f = c->p * c->km;
    c->Wf = 1354 * c->p;
    c->n = c->L * c->f;
    c->F5 = 2 * c->p;
    return 0;
}

// tie-point quad column of fine column i                          (a6, a7)
GTP_HD int gtp_quad_col(const CviirsCfg& c, int i)
{
    int q = (c.km == 1) ? (i / c.f) : ((i >= c.F5) ? (i - c.F5) / c.f : 0);
    return q > c.W - 2 ? c.W - 2 : q;
}

// first fine column whose quad column is q, and one past the last
GTP_HD int gtp_quad_col_begin(const CviirsCfg& c, int q)
{
    if (q <= 0) return 0;
    if (q > c.W - 2) return c.Wf;
    return (c.km == 1) ? q * c.f : c.F5 + q * c.f;
}

// tie-point quad row of fine row j within its scan                 (a6, a7)
GTP_HD int gtp_quad_row(const CviirsCfg& c, int j)
{
    if (c.km != 1) return 0;
    int h = c.f / 2;
    int r = (j >= h) ? (j - h) / c.f : 0;
    return r > 8 ? 8 : r;
}

// sub-pixel coordinate x[i] (a small integer; float32 in the reference)   (a4, a5)
GTP_HD int gtp_coord_x(const CviirsCfg& c, int i)
{
    if (c.km == 1) return (i < c.Wf - c.f) ? (i % c.f) : c.f + (i - (c.Wf - c.f));
    if (i < 2) return i - 2;                                   // x[0] = -2, x[1] = -1
    if (c.W == 271) { if (i >= c.Wf - 2) return 5 + (i - (c.Wf - 2)); }
    else            { if (i >= c.Wf - 7) return 5 + (i - (c.Wf - 7)); }
    return (i - 2) % c.f;
}

// sub-pixel coordinate y[j] TIMES TWO (so that the half-integers stay integral)
GTP_HD int gtp_coord_y2(const CviirsCfg& c, int j)
{
    if (c.km != 1) return 2 * (j - 2);
    int h = c.f / 2;
    if (j < h) return 2 * (j - h) + 1;                         // 0.5 + j - h
    if (j >= c.n - h) return 2 * (c.f + h - (c.n - j)) + 1;    // (f + 0.5) + (h - (n - j))
    return 2 * ((j + h) % c.f) + 1;                            // ((j + h) mod f) + 0.5
}
