// gtp_legacy.cu - the reference's LEGACY spline interpolators on sm_100a.
//
// modis5kmto1km / modis1kmto500m / modis1kmto250m (/root/reference/geotiepoints/__init__.py:49-152):
// GeoInterpolator (geointerpolator.py:10-75) over Interpolator (interpolator.py:54-229) with
// along-track order 1, cross-track order 3 and chunk_size = one scan:
//   lon / lat -> Cartesian xyz; fill_borders("y", "x") - per scan a first and a last row extrapolated
//   linearly to the scan's first / last fine row, then a first / last column extrapolated to the first /
//   last fine column; RectBivariateSpline(rows, cols, xyz, s = 0, kx = 1, ky = 3) - linear along the rows,
//   the interpolating cubic spline without knots at the second and second-to-last column (not-a-knot)
//   along the columns; xyz -> lon / lat.
//
// Everything but the column spline is linear in the data with weights that do not depend on the column, so
// the border rows are never formed: a fine row is the linear blend (inside the scan) or extrapolation
// (above the first / below the last tie row of the scan) of two of the scan's tie rows, applied AFTER the
// column spline.  Three launches per call:
//   1. legacy_coef_kernel (one thread): the forward-elimination coefficients of the not-a-knot tridiagonal
//      system - they depend on the column positions only;
//   2. legacy_rows_kernel: one thread per (tie row, Cartesian component): xyz of the row incl. the border
//      column(s), then the Thomas algorithm for the second derivatives M.  1355 dependent steps per thread;
//      the whole pass moves 48 B per tie point (3 B per 250 m output pixel) and is latency-, not
//      bandwidth-bound: ~60 us whatever the batch size up to ~40 granules;
//   3. legacy_eval_kernel: one thread per (scan, fine column): walks down the scan's tie rows, evaluates
//      the cubic of its column interval per tie row (y, M of the two interval ends: loads that a warp shares),
//      blends / extrapolates along the rows and back-transforms with the reference's formulas
//      (geointerpolator.py:63-75).  Coalesced 8-byte stores, 256 B per warp and output array.
// The back-transform is the rounding-faithful one (library acos / asin / sqrt, ~370 FP64-pipe slots per
// pixel): this first version is FP64-bound like gtp_generic.cu, not HBM-bound.
#include <cuda_runtime.h>
#include <math.h>
#include "gtp_kernels.h"

namespace {

constexpr double kR = GTP_EARTH_RADIUS;

struct LegacyGeom {
    int kind;                  // 0: 5 km -> 1 km, 1: 1 km -> 500 m, 2: 1 km -> 250 m
    int W;                     // tie columns (271 | 1354)
    int n;                     // spline points per row = W + border columns (273 | 1355)
    int first;                 // 1 when a first border column is added (5 km)
    int L;                     // tie rows per scan (2 | 10)
    int nf;                    // fine rows per scan (10 | 20 | 40)
    int Wf;                    // fine columns (1354 | 2708 | 5416)
};

__host__ __device__ inline LegacyGeom legacy_geom(int kind)
{
    LegacyGeom g;
    g.kind = kind;
    if (kind == 0) { g.W = 271; g.n = 273; g.first = 1; g.L = 2; g.nf = 10; g.Wf = 1354; }
    else if (kind == 1) { g.W = 1354; g.n = 1355; g.first = 0; g.L = 10; g.nf = 20; g.Wf = 2708; }
    else { g.W = 1354; g.n = 1355; g.first = 0; g.L = 10; g.nf = 40; g.Wf = 5416; }
    return g;
}

// position of spline point i of a row, in the tie-point coordinate of __init__.py (:54-58, :106-110, :133-141)
__host__ __device__ inline double aug_col(const LegacyGeom& g, int i)
{
    if (g.kind == 0) {
        if (i == 0) return 0.0;                                   // cols1km[0]
        if (i == g.n - 1) return 1353.0 / 5.0;                    // cols1km[-1]
        return (double)(2 + 5 * (i - 1)) / 5.0;                   // cols5km
    }
    if (i == g.n - 1) return (g.kind == 1) ? 2707.0 / 2.0 : 5415.0 / 4.0;
    return (double)i;
}

__host__ __device__ inline double fine_col(const LegacyGeom& g, int j)
{
    return (g.kind == 0) ? (double)j / 5.0 : (g.kind == 1 ? (double)j / 2.0 : (double)j / 4.0);
}

// the column interval fine column j lies in (index of its left spline point)
__host__ __device__ inline int col_interval(const LegacyGeom& g, int j)
{
    if (g.kind == 0) return (j < 2) ? 0 : 1 + (j - 2) / 5;
    return (g.kind == 1) ? j / 2 : j / 4;
}

// coef[0 .. m): cp (upper / den), coef[m .. 2m): 1 / den, coef[2m .. 3m): lower; m = n - 2 unknowns M_1 .. M_{n-2}
__global__ void legacy_coef_kernel(double* __restrict__ coef, int kind)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const LegacyGeom g = legacy_geom(kind);
    const int n = g.n, m = n - 2;
    double cp_prev = 0.0;
    for (int k = 0; k < m; ++k) {
        const double h0 = aug_col(g, k + 1) - aug_col(g, k), h1 = aug_col(g, k + 2) - aug_col(g, k + 1);
        double lower = h0, diag = 2.0 * (h0 + h1), upper = h1;
        if (k == 0) { diag += h0 * (h0 + h1) / h1; upper -= h0 * h0 / h1; }            // not-a-knot at point 1
        if (k == m - 1) { diag += h1 * (h1 + h0) / h0; lower -= h1 * h1 / h0; }        // ... and at point n - 2
        const double den = (k == 0) ? diag : diag - lower * cp_prev;
        const double cp = upper / den;
        coef[k] = cp; coef[m + k] = 1.0 / den; coef[2 * m + k] = lower;
        cp_prev = cp;
    }
}

// one thread per (tie row, component): Y[row][comp][n] and M[row][comp][n]
__global__ void __launch_bounds__(128)
legacy_rows_kernel(const double* __restrict__ lon, const double* __restrict__ lat, const double* __restrict__ coef,
                   double* __restrict__ Y, double* __restrict__ M, long long n_rows, int kind)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_rows * 3) return;
    const LegacyGeom g = legacy_geom(kind);
    const long long row = t / 3;
    const int comp = (int)(t - row * 3);
    const int n = g.n, m = n - 2, W = g.W;
    const double* plon = lon + row * W;
    const double* plat = lat + row * W;
    double* y = Y + t * n;
    double* mm = M + t * n;
    // pass 0: lonlat2xyz (geointerpolator.py:52-60), this thread's component only
    for (int c = 0; c < W; ++c) {
        const double lo = plon[c] * GTP_DEG2RAD, la = plat[c] * GTP_DEG2RAD;
        double v;
        if (comp == 2) v = kR * sin(la);
        else {
            const double rc = kR * cos(la);
            v = (comp == 0) ? rc * cos(lo) : rc * sin(lo);
        }
        y[c + g.first] = v;
    }
    // border columns (interpolator.py:100-148, _linear_extrapolate :33-51)
    if (g.first) {
        const double p0 = aug_col(g, 1), p1 = aug_col(g, 2), x = aug_col(g, 0);
        y[0] = y[2] + (x - p1) / (p0 - p1) * (y[1] - y[2]);
    }
    {
        const double p0 = aug_col(g, n - 3), p1 = aug_col(g, n - 2), x = aug_col(g, n - 1);
        y[n - 1] = y[n - 2] + (x - p1) / (p0 - p1) * (y[n - 3] - y[n - 2]);
    }
    // Thomas algorithm: forward elimination (dp kept in mm[1 .. m]), back substitution
    const double* cp = coef;
    const double* iden = coef + m;
    const double* low = coef + 2 * m;
    double y0 = y[0], y1 = y[1], dp_prev = 0.0;
    double x0 = aug_col(g, 0), x1 = aug_col(g, 1);
    for (int k = 0; k < m; ++k) {
        const double y2 = y[k + 2], x2 = aug_col(g, k + 2);
        const double rhs = 6.0 * ((y2 - y1) / (x2 - x1) - (y1 - y0) / (x1 - x0));
        const double dp = (rhs - low[k] * dp_prev) * iden[k];
        mm[k + 1] = dp;
        dp_prev = dp; y0 = y1; y1 = y2; x0 = x1; x1 = x2;
    }
    double u_next = mm[m];                      // M_{n-2}
    for (int k = m - 2; k >= 0; --k) {
        const double u = mm[k + 1] - cp[k] * u_next;
        mm[k + 1] = u;
        u_next = u;
    }
    {
        const double h0 = aug_col(g, 1) - aug_col(g, 0), h1 = aug_col(g, 2) - aug_col(g, 1);
        mm[0] = ((h0 + h1) * mm[1] - h0 * mm[2]) / h1;
        const double g0 = aug_col(g, n - 1) - aug_col(g, n - 2), g1 = aug_col(g, n - 2) - aug_col(g, n - 3);
        mm[n - 1] = ((g0 + g1) * mm[n - 2] - g0 * mm[n - 3]) / g1;
    }
}

// xyz -> lon / lat as geointerpolator.py:63-75 (thr = 0.8)
__device__ __forceinline__ void legacy_backtransform(double x, double y, double z, double& lon, double& lat)
{
    const double rho = sqrt(x * x + y * y);
    const double sy = (y > 0.0) ? 1.0 : ((y < 0.0) ? -1.0 : ((y == 0.0) ? 0.0 : y));
    lon = (acos(x / rho) * GTP_RAD2DEG) * sy;
    const double nz = z / kR;
    if (fabs(nz) < 0.8) lat = 90.0 - acos(nz) * GTP_RAD2DEG;
    else {
        const double sz = (z > 0.0) ? 1.0 : ((z < 0.0) ? -1.0 : ((z == 0.0) ? 0.0 : z));
        lat = sz * (90.0 - asin(rho / kR) * GTP_RAD2DEG);
    }
}

// one thread per (scan, fine column)
__global__ void __launch_bounds__(128)
legacy_eval_kernel(const double* __restrict__ Y, const double* __restrict__ M, double* __restrict__ out_lon,
                   double* __restrict__ out_lat, long long n_scans, int kind)
{
    const LegacyGeom g = legacy_geom(kind);
    const int blocks_per_scan = (g.Wf + blockDim.x - 1) / blockDim.x;
    const long long scan = blockIdx.x / blocks_per_scan;
    const int j = (int)(blockIdx.x - scan * blocks_per_scan) * blockDim.x + threadIdx.x;
    if (scan >= n_scans || j >= g.Wf) return;
    const int n = g.n;
    const int i = col_interval(g, j);
    const double xa = aug_col(g, i), h = aug_col(g, i + 1) - xa, t = fine_col(g, j) - xa;
    const double inv_h = 1.0 / h, h6 = h / 6.0, inv_6h = 1.0 / (6.0 * h);
    const double* y = Y + (scan * g.L * 3) * n + i;
    const double* mm = M + (scan * g.L * 3) * n + i;
    double* olon = out_lon + scan * (long long)g.nf * g.Wf + j;
    double* olat = out_lat + scan * (long long)g.nf * g.Wf + j;

    // the cubic of this column interval for one tie row (three components)
    auto spline_row = [&](int r, double (&v)[3]) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const long long o = ((long long)r * 3 + c) * n;
            const double y0 = y[o], y1 = y[o + 1], m0 = mm[o], m1 = mm[o + 1];
            const double b = (y1 - y0) * inv_h - h6 * (2.0 * m0 + m1);
            v[c] = y0 + t * (b + t * (0.5 * m0 + t * ((m1 - m0) * inv_6h)));
        }
    };
    // fine row jr of the scan sits at tie-row coordinate pos(jr); rows above the first / below the last tie
    // row extrapolate the first / last pair (fill_borders('y'), interpolator.py:150-197)
    const int f = g.nf / g.L;                                       // fine rows per tie row: 5 | 2 | 4
    double top[3], bot[3];
    spline_row(0, top);
    int jr = 0;
    for (int r = 0; r + 1 < g.L; ++r) {
        spline_row(r + 1, bot);
        // tie rows sit at fine-row index r f + off: off = 2 (5 km), 0.5 (500 m), 1.5 (250 m)
        const double off = (g.kind == 0) ? 2.0 : (g.kind == 1 ? 0.5 : 1.5);
        const int j_end = (r + 2 == g.L) ? g.nf : (int)((r + 1) * f + off) + 1;      // rows up to the lower tie row
        for (; jr < j_end; ++jr) {
            const double w = ((double)jr - off) / (double)f - (double)r;
            const double x = top[0] + w * (bot[0] - top[0]);
            const double yy = top[1] + w * (bot[1] - top[1]);
            const double z = top[2] + w * (bot[2] - top[2]);
            double lo, la;
            legacy_backtransform(x, yy, z, lo, la);
            __stcs(olon + (long long)jr * g.Wf, lo);
            __stcs(olat + (long long)jr * g.Wf, la);
        }
        top[0] = bot[0]; top[1] = bot[1]; top[2] = bot[2];
    }
}

}  // namespace

size_t gtp_legacy_workspace_bytes(int kind, long long n_scans)
{
    const LegacyGeom g = legacy_geom(kind);
    const size_t rows = (size_t)n_scans * g.L;
    return (2 * rows * 3 * g.n + 3 * (size_t)(g.n - 2)) * sizeof(double);
}

void gtp_legacy_shape(int kind, int* tie_rows_per_scan, int* tie_cols, int* fine_rows_per_scan, int* fine_cols)
{
    const LegacyGeom g = legacy_geom(kind);
    *tie_rows_per_scan = g.L; *tie_cols = g.W; *fine_rows_per_scan = g.nf; *fine_cols = g.Wf;
}

cudaError_t gtp_launch_legacy_f64(const double* lon, const double* lat, double* out_lon, double* out_lat,
                                  long long n_scans, int kind, void* workspace, cudaStream_t stream)
{
    if (n_scans <= 0) return cudaSuccess;
    const LegacyGeom g = legacy_geom(kind);
    const long long rows = n_scans * g.L;
    double* Y = (double*)workspace;
    double* M = Y + rows * 3 * g.n;
    double* coef = M + rows * 3 * g.n;
    legacy_coef_kernel<<<1, 1, 0, stream>>>(coef, kind);
    const long long threads = rows * 3;
    legacy_rows_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, stream>>>(lon, lat, coef, Y, M, rows, kind);
    const long long blocks = n_scans * ((g.Wf + 127) / 128);
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    legacy_eval_kernel<<<(unsigned)blocks, 128, 0, stream>>>(Y, M, out_lon, out_lat, n_scans, kind);
    return cudaGetLastError();
}
