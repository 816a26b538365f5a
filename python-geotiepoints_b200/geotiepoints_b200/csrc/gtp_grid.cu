// gtp_grid.cu - the reference's generic rectilinear-grid interpolators on sm_100a (third item).
//
// /root/reference/geotiepoints/interpolator.py:232-378 (Single / Multiple Grid | Spline Interpolator) and
// geointerpolator.py:76-113 (GeoGridInterpolator, GeoSplineInterpolator): tie points on a rectilinear grid
// (two strictly ascending 1-D coordinate vectors, any spacing), results on the product of two 1-D vectors of fine
// coordinates.  The reference hands the work to scipy: RegularGridInterpolator (method "linear" - the default -,
// "nearest", "slinear", "cubic", ...) or RectBivariateSpline(kx, ky, s = 0); the Geo classes run it on the
// Cartesian x / y / z of the lon / lat tie points (lonlat2xyz, geointerpolator.py:52-60) and back-transform
// (xyz2lonlat, :63-75).  All of these are tensor products of 1-D schemes, per axis one of
//   nearest | linear (RegularGridInterpolator "linear" / "slinear", spline order 1) |
//   cubic = the interpolating cubic spline with the not-a-knot end conditions ("cubic", spline order 3),
// and out of the tie-point range one of: extrapolate with the end polynomial (fill_value = None), a fill value
// (bounds_error = False), or the end value (FITPACK clamps the arguments of RectBivariateSpline).
//
// Launches (gtp_launch_grid): grid_locate_kernel per axis (binary search: interval, offset, weight, out-of-range
// flag of every fine coordinate - 1-D vectors, a few KB); GEO: grid_xyz_kernel (lon / lat -> x / y / z planes and,
// per tie point, the anchor of the back-transform: 8 planes); cubic along x: coefficients of the tridiagonal system
// (one thread; they depend on the knots only), Thomas algorithm per (component, tie row) for the second derivatives
// along x; cubic along y: grid_rows_kernel evaluates the x-scheme of every tie row at the fine columns (Z: tie rows
// x fine columns), Thomas per (component, fine column) along y - neighbouring threads on neighbouring addresses.
// Evaluation: GEO with linear / nearest on both axes (the reference's default): grid_geo_linear4_kernel, a thread
// per (fine row, four consecutive fine columns) - the cell's tie columns blended along y and its anchor loaded once,
// four independent back-transform chains, 256-bit stores; everything else: grid_eval_kernel, a thread per output
// pixel (x-scheme on the two tie rows around it, or Z and its second derivatives; y-scheme; GEO: anchor-relative
// back-transform - gtp_anchor_bt.cuh; the reference's formulas with library functions where the pixel is further
// than ~60 km from its tie point), coalesced 8-byte streaming stores.
#include <cstdint>
#include <cuda_runtime.h>
#include <math.h>
#include "gtp_kernels.h"
#include "gtp_anchor_bt.cuh"

namespace {

constexpr double kR = GTP_EARTH_RADIUS;

// per axis: what grid_locate_kernel leaves for every fine coordinate
struct AxisLoc {
    const int* idx;             // interval: tie[idx] <= x < tie[idx + 1], clipped to [0, n - 2]
    const double* d;            // x - tie[idx]
    const double* h;            // tie[idx + 1] - tie[idx]
    const double* u;            // d / h; nearest: 0 | 1
    const unsigned char* oob;   // x outside [tie[0], tie[n - 1]]
};

struct GridArgs {
    int ny, nx, n_comp;
    long long my, mx;
    const double* V;            // [component][tie row][tie column]
    const double* Mx;           // second derivatives along x, like V                       (cubic x, y not cubic)
    const double* Z;            // x-scheme at the fine columns: [component][tie row][fine column]  (cubic y)
    const double* My;           // second derivatives of Z along y                                   (cubic y)
    AxisLoc y, x;
    const double* lon;          // GEO: the tie points themselves (anchors)
    const double* lat;
    const double* anchors;      // GEO: kAnchorPacked planes, one element per tie point each (grid_xyz_kernel)
    double* out;                // [component][fine row][fine column]; GEO: lon plane
    double* out2;               // GEO: lat plane
    int fill_oob;               // out-of-range pixels get `fill`
    double fill;
    int nearest_y, nearest_x;   // u is 0 | 1: take that end (the other one stays out of the arithmetic)
};

__device__ __forceinline__ double cubic_piece(double y0, double y1, double m0, double m1, double h, double d)
{
    const double b = (y1 - y0) / h - h * (2.0 * m0 + m1) / 6.0;
    return y0 + d * (b + d * (0.5 * m0 + d * ((m1 - m0) / (6.0 * h))));
}

__global__ void __launch_bounds__(256)
grid_locate_kernel(const double* __restrict__ tie, int n, const double* __restrict__ fine, long long m, int method, int clamp,
                   int* __restrict__ idx, double* __restrict__ d, double* __restrict__ h, double* __restrict__ u,
                   unsigned char* __restrict__ oob)
{
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= m) return;
    double x = fine[j];
    const double lo_end = tie[0], hi_end = tie[n - 1];
    const bool outside = (x < lo_end) || (x > hi_end);
    if (clamp) x = (x < lo_end) ? lo_end : ((x > hi_end) ? hi_end : x);
    int lo = 0, hi = n - 1;                                          // tie[lo] <= x < tie[hi] (or clipped)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= tie[mid]) lo = mid; else hi = mid;
    }
    const double hh = tie[lo + 1] - tie[lo], dd = x - tie[lo];
    double uu = dd / hh;
    if (method == GTP_GRID_NEAREST) uu = (uu <= 0.5) ? 0.0 : ((uu > 0.5) ? 1.0 : uu);     // NaN stays NaN
    idx[j] = lo; d[j] = dd; h[j] = hh; u[j] = uu; oob[j] = outside ? 1 : 0;
}

__global__ void __launch_bounds__(256)
grid_xyz_kernel(const double* __restrict__ lon, const double* __restrict__ lat, double* __restrict__ xyz,
                double* __restrict__ anchors, long long n)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    double sl, cl, sp, cp;
    sincos(lon[t] * GTP_DEG2RAD, &sl, &cl);                          // lonlat2xyz, geointerpolator.py:52-60
    sincos(lat[t] * GTP_DEG2RAD, &sp, &cp);
    const double rc = kR * cp;
    const double x = rc * cl, y = rc * sl, z = kR * sp;
    xyz[t] = x; xyz[n + t] = y; xyz[2 * n + t] = z;
    // the tie point as an anchor of the back-transform, prepared once for all the pixels around it
    anchor_bt_pack(anchor_bt_setup(kR, lon[t], lat[t], x, y, z), kR, 0.8, anchors, n, t);
}

// coef[0 .. m): upper / den, [m .. 2m): 1 / den, [2m .. 3m): lower of the not-a-knot system, m = n - 2 unknowns;
// [3m .. 3m + n - 1): 6 / h of the n - 1 intervals
__global__ void grid_coef_kernel(const double* __restrict__ knots, int n, double* __restrict__ coef)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const int m = n - 2;
    double cp_prev = 0.0;
    for (int k = 0; k < m; ++k) {
        const double h0 = knots[k + 1] - knots[k], h1 = knots[k + 2] - knots[k + 1];
        double lower = h0, diag = 2.0 * (h0 + h1), upper = h1;
        if (k == 0) { diag += h0 * (h0 + h1) / h1; upper -= h0 * h0 / h1; }
        if (k == m - 1) { diag += h1 * (h1 + h0) / h0; lower -= h1 * h1 / h0; }
        const double den = (k == 0) ? diag : diag - lower * cp_prev;
        const double cp = upper / den;
        coef[k] = cp; coef[m + k] = 1.0 / den; coef[2 * m + k] = lower;
        cp_prev = cp;
    }
    for (int j = 0; j < n - 1; ++j) coef[3 * m + j] = 6.0 / (knots[j + 1] - knots[j]);
}

// second derivatives of the not-a-knot spline through n points y[0], y[st], ...: mm[0], mm[st], ...
__device__ __forceinline__ void thomas_solve(const double* __restrict__ y, double* __restrict__ mm, long long st, int n,
                                             const double* __restrict__ knots, const double* __restrict__ coef)
{
    const int m = n - 2;
    const double* cp = coef;
    const double* iden = coef + m;
    const double* low = coef + 2 * m;
    const double* s6 = coef + 3 * m;
    constexpr int kBatch = 8;                                        // loads of eight steps issued together
    double y1 = y[st], dp_prev = 0.0;
    double left = (y1 - y[0]) * s6[0];                               // 6 (y_{k+1} - y_k) / h_k
    for (int k0 = 0; k0 < m; k0 += kBatch) {
        double yb[kBatch], lb[kBatch], ib[kBatch], sb[kBatch];
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
            const int k = min(k0 + q, m - 1);
            yb[q] = y[(long long)(k + 2) * st]; lb[q] = low[k]; ib[q] = iden[k]; sb[q] = s6[k + 1];
        }
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
            const int k = k0 + q;
            if (k < m) {
                const double right = (yb[q] - y1) * sb[q];
                const double dp = ((right - left) - lb[q] * dp_prev) * ib[q];
                mm[(long long)(k + 1) * st] = dp;
                dp_prev = dp; y1 = yb[q]; left = right;
            }
        }
    }
    double next = dp_prev;
    for (int k = m - 2; k >= 0; --k) {
        const double v = mm[(long long)(k + 1) * st] - cp[k] * next;
        mm[(long long)(k + 1) * st] = v;
        next = v;
    }
    const double h0 = knots[1] - knots[0], h1 = knots[2] - knots[1];
    mm[0] = ((h0 + h1) * mm[st] - h0 * mm[2 * st]) / h1;
    const double g0 = knots[n - 1] - knots[n - 2], g1 = knots[n - 2] - knots[n - 3];
    mm[(long long)(n - 1) * st] = ((g0 + g1) * mm[(long long)(n - 2) * st] - g0 * mm[(long long)(n - 3) * st]) / g1;
}

// n_lines independent splines of n points: point stride st, line l starts at (l / inner) * outer_stride + (l % inner)
__global__ void __launch_bounds__(128)
grid_thomas_kernel(const double* __restrict__ y, double* __restrict__ mm, long long n_lines, long long inner,
                   long long outer_stride, long long st, int n, const double* __restrict__ knots,
                   const double* __restrict__ coef)
{
    const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (l >= n_lines) return;
    const long long base = (l / inner) * outer_stride + (l % inner);
    thomas_solve(y + base, mm + base, st, n, knots, coef);
}

// the x-scheme of component c on tie row r for the pixel's column (interval ix)
template <int MX>
__device__ __forceinline__ double row_value(const GridArgs& a, int c, int r, int ix, double dx, double hx, double ux)
{
    const long long o = ((long long)c * a.ny + r) * a.nx + ix;
    const double v0 = a.V[o], v1 = a.V[o + 1];
    if (MX == 3) return cubic_piece(v0, v1, a.Mx[o], a.Mx[o + 1], hx, dx);
    if (a.nearest_x) return (ux == 0.0) ? v0 : ((ux == 1.0) ? v1 : ux);     // _evaluate_nearest; NaN coordinate -> NaN
    return v0 * (1.0 - ux) + v1 * ux;                                // _evaluate_linear / evaluate_linear_2d
}

// Z[c][r][j] for every tie row r and fine column j (cubic along y needs whole columns of it)
template <int MX>
__global__ void __launch_bounds__(256)
grid_rows_kernel(GridArgs a, double* __restrict__ Z)
{
    // blockIdx.x: (component, tie row); blockIdx.y: 256 fine columns
    const int c = (int)(blockIdx.x / (unsigned)a.ny), r = (int)(blockIdx.x - (unsigned)c * (unsigned)a.ny);
    const long long j = blockIdx.y * (long long)blockDim.x + threadIdx.x;
    if (j >= a.mx) return;
    Z[((long long)c * a.ny + r) * a.mx + j] = row_value<MX>(a, c, r, a.x.idx[j], a.x.d[j], a.x.h[j], a.x.u[j]);
}

template <int MX, int MY>
__device__ __forceinline__ double pixel_value(const GridArgs& a, int c, long long j, int iy, double dy, double hy, double uy,
                                              int ix, double dx, double hx, double ux)
{
    if (MY == 3) {
        const long long o = ((long long)c * a.ny + iy) * a.mx + j;
        return cubic_piece(a.Z[o], a.Z[o + a.mx], a.My[o], a.My[o + a.mx], hy, dy);
    }
    if (a.nearest_y) return (uy == 0.0) ? row_value<MX>(a, c, iy, ix, dx, hx, ux)
                                        : ((uy == 1.0) ? row_value<MX>(a, c, iy + 1, ix, dx, hx, ux) : uy);
    return row_value<MX>(a, c, iy, ix, dx, hx, ux) * (1.0 - uy) + row_value<MX>(a, c, iy + 1, ix, dx, hx, ux) * uy;
}

template <int MX, int MY, bool GEO>
__global__ void __launch_bounds__(256)
grid_eval_kernel(GridArgs a)
{
    // a CTA row per fine row (blockIdx.x), 256 fine columns per CTA (blockIdx.y): no 64-bit division per pixel
    const long long i = blockIdx.x, j = blockIdx.y * (long long)blockDim.x + threadIdx.x;
    if (j >= a.mx) return;
    const long long n_px = a.my * a.mx, t = i * a.mx + j;
    const int iy = a.y.idx[i], ix = a.x.idx[j];
    const double dy = a.y.d[i], hy = a.y.h[i], uy = a.y.u[i];
    const double dx = a.x.d[j], hx = a.x.h[j], ux = a.x.u[j];
    const bool filled = a.fill_oob && (a.y.oob[i] | a.x.oob[j]);
    if (!GEO) {
        for (int c = 0; c < a.n_comp; ++c) {
            const double v = filled ? a.fill : pixel_value<MX, MY>(a, c, j, iy, dy, hy, uy, ix, dx, hx, ux);
            __stcs(a.out + (long long)c * n_px + t, v);
        }
        return;
    }
    double lon, lat;
    if (filled) xyz_to_lonlat_reference(kR, 0.8, a.fill, a.fill, a.fill, lon, lat);
    else {
        const double x = pixel_value<MX, MY>(a, 0, j, iy, dy, hy, uy, ix, dx, hx, ux);
        const double y = pixel_value<MX, MY>(a, 1, j, iy, dy, hy, uy, ix, dx, hx, ux);
        const double z = pixel_value<MX, MY>(a, 2, j, iy, dy, hy, uy, ix, dx, hx, ux);
        // xyz2lonlat (geointerpolator.py:63-75) around the nearer tie row / column of the pixel's cell
        const int ar = iy + ((uy > 0.5) ? 1 : 0), ac = ix + ((ux > 0.5) ? 1 : 0);
        const long long tie = (long long)ar * a.nx + ac, n = (long long)a.ny * a.nx;
        const AnchorBT anchor = anchor_bt_unpack(a.anchors, n, tie, a.lon[tie], a.lat[tie], a.V[tie], a.V[n + tie], a.V[2 * n + tie]);
        anchor_bt_eval(anchor, kR, 0.8, x, y, z, lon, lat);
    }
    __stcs(a.out + t, lon);
    __stcs(a.out2 + t, lat);
}

// GeoGridInterpolator's default ("linear", also "nearest"): one thread per (fine row, 4 consecutive fine columns).
// The two tie columns of a cell are blended along y once per cell and component, the anchor (the cell's tie point
// in the nearer tie row) is loaded once per cell; a pixel is then 9 FP64 operations + the back-transform.  With one
// thread per pixel the kernel spent ~300 instructions per pixel, three quarters of them on indices, the loads of
// twelve tie values and of the anchor.
__global__ void __launch_bounds__(128, 8)
grid_geo_linear4_kernel(GridArgs a, int vector_stores)
{
    const long long i = blockIdx.x, j0 = 4 * (blockIdx.y * (long long)blockDim.x + threadIdx.x);
    if (j0 >= a.mx) return;
    const long long n = (long long)a.ny * a.nx, t0 = i * a.mx + j0;
    const int iy = a.y.idx[i];
    const double uy = a.y.u[i];
    const bool oob_y = a.fill_oob && a.y.oob[i];
    const int ar = iy + ((uy > 0.5) ? 1 : 0);
    const double* const row0 = a.V + (long long)iy * a.nx;           // tie row iy of component 0; + nx: the next one
    auto along_y = [&](const double* p) -> double {
        if (a.nearest_y) return (uy == 0.0) ? p[0] : ((uy == 1.0) ? p[a.nx] : uy);
        return p[0] * (1.0 - uy) + p[a.nx] * uy;
    };
    double lon[4], lat[4];
    int ix[4];
    double ux[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long j = (j0 + k < a.mx) ? j0 + k : a.mx - 1;     // the ragged end repeats the last column (not stored)
        ix[k] = a.x.idx[j]; ux[k] = a.x.u[j];
    }
    // one cell's columns blended along y, and its anchor
    struct Cell { double left[3], right[3]; AnchorBT anchor; };
    auto load_cell = [&](int c0) -> Cell {
        Cell q;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            q.left[c] = along_y(row0 + c * n + c0);
            q.right[c] = along_y(row0 + c * n + c0 + 1);
        }
        const long long tie = (long long)ar * a.nx + c0;
        q.anchor = anchor_bt_unpack(a.anchors, n, tie, a.lon[tie], a.lat[tie], a.V[tie], a.V[n + tie], a.V[2 * n + tie]);
        return q;
    };
    auto pixel = [&](const Cell& q, double u, double& lo, double& la) {
        double v[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
            v[c] = a.nearest_x ? ((u == 0.0) ? q.left[c] : ((u == 1.0) ? q.right[c] : u)) : q.left[c] * (1.0 - u) + q.right[c] * u;
        anchor_bt_eval(q.anchor, kR, 0.8, v[0], v[1], v[2], lo, la);
    };
    if (ix[0] == ix[1] && ix[0] == ix[2] && ix[0] == ix[3]) {        // the usual case: one cell, four independent chains
        const Cell q = load_cell(ix[0]);
#pragma unroll
        for (int k = 0; k < 4; ++k) pixel(q, ux[k], lon[k], lat[k]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const Cell q = load_cell(ix[k]);
            pixel(q, ux[k], lon[k], lat[k]);
        }
    }
    if (a.fill_oob) {                                                // scipy: result[out_of_bounds] = fill_value, then xyz2lonlat
        double flon, flat;
        xyz_to_lonlat_reference(kR, 0.8, a.fill, a.fill, a.fill, flon, flat);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long j = (j0 + k < a.mx) ? j0 + k : a.mx - 1;
            if (oob_y || a.x.oob[j]) { lon[k] = flon; lat[k] = flat; }
        }
    }
    if (vector_stores && j0 + 4 <= a.mx) {
        asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(a.out + t0), "d"(lon[0]), "d"(lon[1]), "d"(lon[2]), "d"(lon[3]) : "memory");
        asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(a.out2 + t0), "d"(lat[0]), "d"(lat[1]), "d"(lat[2]), "d"(lat[3]) : "memory");
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (j0 + k < a.mx) { __stcs(a.out + t0 + k, lon[k]); __stcs(a.out2 + t0 + k, lat[k]); }
    }
}

inline size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }

struct AxisWs { int* idx; double *d, *h, *u; unsigned char* oob; };

inline size_t axis_bytes(long long m) { return align16((size_t)m * 4) + 3 * align16((size_t)m * 8) + align16((size_t)m); }

inline AxisWs carve_axis(char*& p, long long m)
{
    AxisWs w;
    w.idx = (int*)p; p += align16((size_t)m * 4);
    w.d = (double*)p; p += align16((size_t)m * 8);
    w.h = (double*)p; p += align16((size_t)m * 8);
    w.u = (double*)p; p += align16((size_t)m * 8);
    w.oob = (unsigned char*)p; p += align16((size_t)m);
    return w;
}

template <int MX, int MY>
cudaError_t launch_eval(const GridArgs& a, bool geo, cudaStream_t stream)
{
    const long long pieces = (a.mx + 255) / 256;
    if (a.my > 0x7fffffffLL || pieces > 65535) return cudaErrorInvalidConfiguration;
    const dim3 blocks((unsigned)a.my, (unsigned)pieces, 1);
    if (geo) grid_eval_kernel<MX, MY, true><<<blocks, 256, 0, stream>>>(a);
    else grid_eval_kernel<MX, MY, false><<<blocks, 256, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace

size_t gtp_grid_workspace_bytes(int ny, int nx, long long my, long long mx, int n_comp, int geo, int method_y, int method_x)
{
    size_t b = axis_bytes(my) + axis_bytes(mx);
    const size_t tie = (size_t)n_comp * ny * nx * 8, rows = (size_t)n_comp * ny * (size_t)mx * 8;
    if (geo) b += align16(tie) + align16((size_t)kAnchorPacked * ny * nx * 8);
    if (method_x == GTP_GRID_CUBIC) b += align16(tie) + align16((size_t)4 * nx * 8);
    if (method_y == GTP_GRID_CUBIC) b += 2 * align16(rows) + align16((size_t)4 * ny * 8);
    return b;
}

// values: n_comp planes of ny x nx (geo: lon plane; values2: lat plane, n_comp = 3 internally);
// out: n_comp planes of my x mx (geo: lon; out2: lat).  *launches gets the number of kernels launched.
cudaError_t gtp_launch_grid(const double* tie_y, int ny, const double* tie_x, int nx, const double* values,
                            const double* values2, int n_comp, int geo, const double* fine_y, long long my,
                            const double* fine_x, long long mx, int method_y, int method_x, int oob_mode, double fill,
                            double* out, double* out2, void* workspace, cudaStream_t stream, int* launches)
{
    int count = 0;
    if (launches) *launches = 0;
    if (my <= 0 || mx <= 0) return cudaSuccess;
    if (geo) n_comp = 3;
    char* p = (char*)workspace;
    const AxisWs wy = carve_axis(p, my), wx = carve_axis(p, mx);
    const int clamp = (oob_mode == GTP_GRID_OOB_CLAMP);
    grid_locate_kernel<<<(unsigned)((my + 255) / 256), 256, 0, stream>>>(tie_y, ny, fine_y, my, method_y, clamp, wy.idx, wy.d, wy.h, wy.u, wy.oob);
    grid_locate_kernel<<<(unsigned)((mx + 255) / 256), 256, 0, stream>>>(tie_x, nx, fine_x, mx, method_x, clamp, wx.idx, wx.d, wx.h, wx.u, wx.oob);
    count += 2;
    const long long n_tie = (long long)ny * nx;
    GridArgs a;
    a.ny = ny; a.nx = nx; a.n_comp = n_comp; a.my = my; a.mx = mx;
    a.V = values; a.Mx = nullptr; a.Z = nullptr; a.My = nullptr;
    a.y = AxisLoc{wy.idx, wy.d, wy.h, wy.u, wy.oob};
    a.x = AxisLoc{wx.idx, wx.d, wx.h, wx.u, wx.oob};
    a.lon = values; a.lat = values2; a.anchors = nullptr; a.out = out; a.out2 = out2;
    a.fill_oob = (oob_mode == GTP_GRID_OOB_FILL); a.fill = fill;
    a.nearest_y = (method_y == GTP_GRID_NEAREST); a.nearest_x = (method_x == GTP_GRID_NEAREST);
    if (geo) {
        double* xyz = (double*)p; p += align16((size_t)3 * n_tie * 8);
        double* anchors = (double*)p; p += align16((size_t)kAnchorPacked * n_tie * 8);
        grid_xyz_kernel<<<(unsigned)((n_tie + 255) / 256), 256, 0, stream>>>(values, values2, xyz, anchors, n_tie);
        ++count;
        a.V = xyz; a.anchors = anchors;
    }
    if (method_x == GTP_GRID_CUBIC) {
        double* Mx = (double*)p; p += align16((size_t)n_comp * n_tie * 8);
        double* coef = (double*)p; p += align16((size_t)4 * nx * 8);
        grid_coef_kernel<<<1, 32, 0, stream>>>(tie_x, nx, coef);
        const long long lines = (long long)n_comp * ny;                 // line l: row l of the stacked planes, points 1 apart
        grid_thomas_kernel<<<(unsigned)((lines + 127) / 128), 128, 0, stream>>>(a.V, Mx, lines, 1, nx, 1, nx, tie_x, coef);
        count += 2;
        a.Mx = Mx;
    }
    if (method_y == GTP_GRID_CUBIC) {
        const long long n_rows = (long long)n_comp * ny * mx;
        double* Z = (double*)p; p += align16((size_t)n_rows * 8);
        double* My = (double*)p; p += align16((size_t)n_rows * 8);
        double* coef = (double*)p; p += align16((size_t)4 * ny * 8);
        const long long pieces = (mx + 255) / 256;
        if (pieces > 65535) return cudaErrorInvalidConfiguration;
        const dim3 blocks((unsigned)(n_comp * ny), (unsigned)pieces, 1);
        if (method_x == GTP_GRID_CUBIC) grid_rows_kernel<3><<<blocks, 256, 0, stream>>>(a, Z);
        else grid_rows_kernel<1><<<blocks, 256, 0, stream>>>(a, Z);
        grid_coef_kernel<<<1, 32, 0, stream>>>(tie_y, ny, coef);
        const long long lines = (long long)n_comp * mx;                 // line l: component l / mx, fine column l % mx; points mx apart
        grid_thomas_kernel<<<(unsigned)((lines + 127) / 128), 128, 0, stream>>>(Z, My, lines, mx, (long long)ny * mx, mx, ny, tie_y, coef);
        count += 3;
        a.Z = Z; a.My = My;
    }
    cudaError_t e;
    const long long quad_pieces = ((mx + 3) / 4 + 127) / 128;
    if (geo && method_y != GTP_GRID_CUBIC && method_x != GTP_GRID_CUBIC && my <= 0x7fffffffLL && quad_pieces <= 65535) {
        // 256-bit stores: every row starts 32-byte aligned
        const int vec = ((((uintptr_t)out | (uintptr_t)out2) & 31) == 0) && (mx % 4 == 0);
        grid_geo_linear4_kernel<<<dim3((unsigned)my, (unsigned)quad_pieces, 1), 128, 0, stream>>>(a, vec);
        e = cudaGetLastError();
    }
    else if (method_y == GTP_GRID_CUBIC) e = launch_eval<1, 3>(a, geo != 0, stream);   // the x-scheme is in Z already
    else if (method_x == GTP_GRID_CUBIC) e = launch_eval<3, 1>(a, geo != 0, stream);
    else e = launch_eval<1, 1>(a, geo != 0, stream);
    ++count;
    if (launches) *launches = count;
    return e;
}
