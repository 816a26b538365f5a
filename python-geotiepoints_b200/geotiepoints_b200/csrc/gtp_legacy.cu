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
// column spline.  Four launches per call:
//   1. legacy_coef_kernel (one thread): the forward-elimination coefficients of the not-a-knot tridiagonal
//      system - they depend on the column positions only;
//   2. legacy_xyz_kernel: lon / lat -> xyz of every tie point into Y[component][point][tie row] (32 x 32 tiles
//      through shared memory: coalesced on both sides); legacy_thomas_kernel: one thread per (component, tie
//      row): the border column(s), then the Thomas algorithm for the second derivatives M - ~2700 dependent
//      steps per thread, neighbouring threads on neighbouring addresses.  48 B per tie point (3 B per 250 m
//      output pixel); latency-, not bandwidth-bound;
//   3. legacy_eval_1km_kernel<P> (1 km -> 250 / 500 m): one thread per (scan, tie-column interval) = P fine
//      columns: per tie row one cubic (y, M of the two interval ends) evaluated at t = k / P, per tie-row pair
//      one anchor, rows blended / extrapolated linearly, anchor-relative back-transform, one 256-bit (128-bit)
//      streaming store per fine row and output array.  legacy_eval_kernel (5 km -> 1 km, or outputs not aligned
//      for the vector stores): the same with one thread per (scan, fine column) and 8-byte stores.
// The back-transform is anchor-relative (legacy_backtransform_fast below; the reference's formulas with
// library acos / asin / sqrt only next to the poles and for out-of-range anchors).
#include <cstdint>
#include <cuda_runtime.h>
#include <math.h>
#include "gtp_kernels.h"
#include "gtp_anchor_bt.cuh"

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

// coef[0 .. m): cp (upper / den), coef[m .. 2m): 1 / den, coef[2m .. 3m): lower; m = n - 2 unknowns M_1 .. M_{n-2};
// coef[3m .. 3m + n - 1): 6 / h_j of the n - 1 intervals (the right-hand sides without divisions)
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
    for (int j = 0; j < n - 1; ++j) coef[3 * m + j] = 6.0 / (aug_col(g, j + 1) - aug_col(g, j));
}

// Layout of the scratch arrays Y (Cartesian tie points incl. the border columns) and M (second derivatives):
// [component][spline point][tie row], tie row fastest - the tridiagonal solves walk along the spline points
// with one thread per tie row, so neighbouring threads touch neighbouring addresses.
__device__ __forceinline__ long long ym_index(int comp, int point, long long row, int n, long long n_rows)
{
    return ((long long)comp * n + point) * n_rows + row;
}

// lonlat2xyz (geointerpolator.py:52-60) of every tie point; 32 x 32 tiles through shared memory so that both
// the row-major reads of lon / lat and the row-fastest writes of Y are coalesced
__global__ void __launch_bounds__(256)
legacy_xyz_kernel(const double* __restrict__ lon, const double* __restrict__ lat, double* __restrict__ Y,
                  long long n_rows, int kind)
{
    const LegacyGeom g = legacy_geom(kind);
    __shared__ double tile[3][32][33];
    const int tiles_x = (g.W + 31) / 32;
    const long long tile_y = blockIdx.x / tiles_x;
    const int col0 = (int)(blockIdx.x - tile_y * tiles_x) * 32;
    const long long row0 = tile_y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long long row = row0 + ty + 8 * k;
        const int col = col0 + tx;
        if (row < n_rows && col < g.W) {
            const double lo = lon[row * g.W + col] * GTP_DEG2RAD, la = lat[row * g.W + col] * GTP_DEG2RAD;
            double sl, cl, sp, cp;
            sincos(lo, &sl, &cl);
            sincos(la, &sp, &cp);
            const double rc = kR * cp;
            tile[0][tx][ty + 8 * k] = rc * cl;
            tile[1][tx][ty + 8 * k] = rc * sl;
            tile[2][tx][ty + 8 * k] = kR * sp;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int col = col0 + ty + 8 * k;
        const long long row = row0 + tx;
        if (row < n_rows && col < g.W) {
#pragma unroll
            for (int c = 0; c < 3; ++c) Y[ym_index(c, col + g.first, row, g.n, n_rows)] = tile[c][ty + 8 * k][tx];
        }
    }
}

// one thread per (component, tie row): the border columns, then the Thomas algorithm for M
__global__ void __launch_bounds__(128)
legacy_thomas_kernel(const double* __restrict__ coef, double* __restrict__ Y, double* __restrict__ M,
                     long long n_rows, int kind)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_rows * 3) return;
    const LegacyGeom g = legacy_geom(kind);
    const int comp = (int)(t / n_rows);
    const long long row = t - (long long)comp * n_rows;
    const int n = g.n, m = n - 2;
    double* y = Y + ym_index(comp, 0, row, n, n_rows);
    double* mm = M + ym_index(comp, 0, row, n, n_rows);
    const long long st = n_rows;                                   // stride between consecutive spline points
    // border columns (interpolator.py:100-148, _linear_extrapolate :33-51)
    if (g.first) {
        const double p0 = aug_col(g, 1), p1 = aug_col(g, 2), x = aug_col(g, 0);
        y[0] = y[2 * st] + (x - p1) / (p0 - p1) * (y[st] - y[2 * st]);
    }
    {
        const double p0 = aug_col(g, n - 3), p1 = aug_col(g, n - 2), x = aug_col(g, n - 1);
        y[(n - 1) * st] = y[(n - 2) * st] + (x - p1) / (p0 - p1) * (y[(n - 3) * st] - y[(n - 2) * st]);
    }
    // Thomas algorithm: forward elimination (dp kept in mm[1 .. m]), back substitution.  Eight steps at a time:
    // their loads are issued together (every step touches another cache line, tie row fastest), then the
    // dependent recurrence runs out of registers - one memory latency per eight steps instead of per step.
    const double* cp = coef;
    const double* iden = coef + m;
    const double* low = coef + 2 * m;
    const double* s6 = coef + 3 * m;             // 6 / h of interval j
    constexpr int kBatch = 8;
    double y0 = y[0], y1 = y[st], dp_prev = 0.0;
    double left = (y1 - y0) * s6[0];             // 6 (y_{k+1} - y_k) / h_k of the interval left of point k + 1
    for (int k0 = 0; k0 < m; k0 += kBatch) {
        double yb[kBatch], lb[kBatch], ib[kBatch], sb[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int k = min(k0 + u, m - 1);
            yb[u] = y[(long long)(k + 2) * st]; lb[u] = low[k]; ib[u] = iden[k]; sb[u] = s6[k + 1];
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int k = k0 + u;
            if (k < m) {
                const double right = (yb[u] - y1) * sb[u];
                const double dp = ((right - left) - lb[u] * dp_prev) * ib[u];
                mm[(long long)(k + 1) * st] = dp;
                dp_prev = dp; y1 = yb[u]; left = right;
            }
        }
    }
    double u_next = dp_prev;                    // M_{n-2} = the last dp
    for (int k0 = m - 2; k0 >= 0; k0 -= kBatch) {
        double db[kBatch], cb[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int k = max(k0 - u, 0);
            db[u] = mm[(long long)(k + 1) * st]; cb[u] = cp[k];
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int k = k0 - u;
            if (k >= 0) {
                const double v = db[u] - cb[u] * u_next;
                mm[(long long)(k + 1) * st] = v;
                u_next = v;
            }
        }
    }
    {
        const double h0 = aug_col(g, 1) - aug_col(g, 0), h1 = aug_col(g, 2) - aug_col(g, 1);
        mm[0] = ((h0 + h1) * mm[st] - h0 * mm[2 * st]) / h1;
        const double g0 = aug_col(g, n - 1) - aug_col(g, n - 2), g1 = aug_col(g, n - 2) - aug_col(g, n - 3);
        mm[(n - 1) * st] = ((g0 + g1) * mm[(n - 2) * st] - g0 * mm[(n - 3) * st]) / g1;
    }
}

// one thread per (scan, fine column)
__global__ void __launch_bounds__(128, 8)
legacy_eval_kernel(const double* __restrict__ lon_in, const double* __restrict__ lat_in,
                   const double* __restrict__ Y, const double* __restrict__ M, double* __restrict__ out_lon,
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
    const long long n_rows = n_scans * g.L, row0 = scan * g.L;
    double* olon = out_lon + scan * (long long)g.nf * g.Wf + j;
    double* olat = out_lat + scan * (long long)g.nf * g.Wf + j;

    // the cubic of this column interval for one tie row (three components)
    auto spline_row = [&](int r, double (&v)[3]) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const long long o = ym_index(c, i, row0 + r, n, n_rows);
            const double y0 = Y[o], y1 = Y[o + n_rows], m0 = M[o], m1 = M[o + n_rows];
            const double b = (y1 - y0) * inv_h - h6 * (2.0 * m0 + m1);
            v[c] = y0 + t * (b + t * (0.5 * m0 + t * ((m1 - m0) * inv_6h)));
        }
    };
    // fine row jr of the scan sits at tie-row coordinate pos(jr); rows above the first / below the last tie
    // row extrapolate the first / last pair (fill_borders('y'), interpolator.py:150-197)
    const int f = g.nf / g.L;                                       // fine rows per tie row: 5 | 2 | 4
    // the anchor of a tie-row pair: the tie point at the left end of this column interval in the upper row
    const int ia = (g.first && i == 0) ? 1 : i;                     // 5 km: the border column is not a tie point
    const int tie_col = ia - g.first;
    double top[3], bot[3];
    spline_row(0, top);
    int jr = 0;
    for (int r = 0; r + 1 < g.L; ++r) {
        const long long tie = (scan * g.L + r) * g.W + tie_col;
        const AnchorBT anchor = anchor_bt_setup(kR, lon_in[tie], lat_in[tie], Y[ym_index(0, ia, row0 + r, n, n_rows)],
                                                  Y[ym_index(1, ia, row0 + r, n, n_rows)], Y[ym_index(2, ia, row0 + r, n, n_rows)]);
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
            anchor_bt_eval(anchor, kR, 0.8, x, yy, z, lo, la);
            __stcs(olon + (long long)jr * g.Wf, lo);
            __stcs(olat + (long long)jr * g.Wf, la);
        }
        top[0] = bot[0]; top[1] = bot[1]; top[2] = bot[2];
    }
}

// 1 km -> 250 m / 500 m: one thread per (scan, tie-column interval) = P fine columns.  The cubic's coefficients
// (per tie row and component) and the anchor (per tie-row pair) are set up once per interval instead of once
// per fine column, and a fine row leaves as one 256-bit (128-bit) store per output array.
template <int P>
__device__ __forceinline__ void store_fine(double* p, const double (&v)[P]);
template <>
__device__ __forceinline__ void store_fine<4>(double* p, const double (&v)[4])
{
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}
template <>
__device__ __forceinline__ void store_fine<2>(double* p, const double (&v)[2])
{
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v[0]), "d"(v[1]) : "memory");
}

template <int P>
__global__ void __launch_bounds__(128, P == 4 ? 3 : 4)              // 166 | 128 registers without spills
legacy_eval_1km_kernel(const double* __restrict__ lon_in, const double* __restrict__ lat_in,
                       const double* __restrict__ Y, const double* __restrict__ M, double* __restrict__ out_lon,
                       double* __restrict__ out_lat, long long n_scans)
{
    const LegacyGeom g = legacy_geom(P == 4 ? 2 : 1);
    constexpr int kIntervals = 1354;                                 // tie columns = intervals (the last one ends at the border column)
    const int blocks_per_scan = (kIntervals + blockDim.x - 1) / blockDim.x;
    const long long scan = blockIdx.x / blocks_per_scan;
    const int i = (int)(blockIdx.x - scan * blocks_per_scan) * blockDim.x + threadIdx.x;
    if (scan >= n_scans || i >= kIntervals) return;
    const int n = g.n;
    const double h = aug_col(g, i + 1) - (double)i;                  // 1, or 0.75 | 0.5 in the last interval
    const double inv_h = 1.0 / h, h6 = h / 6.0, inv_6h = 1.0 / (6.0 * h);
    const long long n_rows = n_scans * g.L, row0 = scan * g.L;
    double* olon = out_lon + scan * (long long)g.nf * g.Wf + (long long)i * P;
    double* olat = out_lat + scan * (long long)g.nf * g.Wf + (long long)i * P;

    // (y, M) of the interval's two ends on one tie row, three components: 12 loads from 12 cache lines
    struct Raw { double y0[3], y1[3], m0[3], m1[3]; };
    auto load_row = [&](int r) -> Raw {
        Raw q;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const long long o = ym_index(c, i, row0 + r, n, n_rows);
            q.y0[c] = Y[o]; q.y1[c] = Y[o + n_rows]; q.m0[c] = M[o]; q.m1[c] = M[o + n_rows];
        }
        return q;
    };
    // the cubic of this interval for one tie row at the P fine columns t = k / P (three components)
    auto spline_row = [&](const Raw& q, double (&v)[3][P]) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double b = (q.y1[c] - q.y0[c]) * inv_h - h6 * (2.0 * q.m0[c] + q.m1[c]), c2 = 0.5 * q.m0[c], c3 = (q.m1[c] - q.m0[c]) * inv_6h;
#pragma unroll
            for (int k = 0; k < P; ++k) {
                const double t = (double)k / (double)P;
                v[c][k] = q.y0[c] + t * (b + t * (c2 + t * c3));
            }
        }
    };
    constexpr double off = (P == 4) ? 1.5 : 0.5;                     // tie row r sits at fine-row index r P + off
    double top[3][P], bot[3][P];
    Raw cur = load_row(0);
    spline_row(cur, top);
    Raw nxt = load_row(1);                                           // always one tie row ahead of the arithmetic
    int jr = 0;
    for (int r = 0; r + 1 < g.L; ++r) {
        const long long tie = (scan * g.L + r) * g.W + i;
        const AnchorBT anchor = anchor_bt_setup(kR, lon_in[tie], lat_in[tie], cur.y0[0], cur.y0[1], cur.y0[2]);
        spline_row(nxt, bot);
        cur = nxt;
        if (r + 2 < g.L) nxt = load_row(r + 2);
        const int j_end = (r + 2 == g.L) ? g.nf : (int)((r + 1) * P + off) + 1;
        for (; jr < j_end; ++jr) {
            const double w = ((double)jr - off) / (double)P - (double)r;
            double lo[P], la[P];
#pragma unroll
            for (int k = 0; k < P; ++k) {
                const double x = top[0][k] + w * (bot[0][k] - top[0][k]);
                const double yy = top[1][k] + w * (bot[1][k] - top[1][k]);
                const double z = top[2][k] + w * (bot[2][k] - top[2][k]);
                anchor_bt_eval(anchor, kR, 0.8, x, yy, z, lo[k], la[k]);
            }
            store_fine<P>(olon + (long long)jr * g.Wf, lo);
            store_fine<P>(olat + (long long)jr * g.Wf, la);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k = 0; k < P; ++k) top[c][k] = bot[c][k];
    }
}

}  // namespace

size_t gtp_legacy_workspace_bytes(int kind, long long n_scans)
{
    const LegacyGeom g = legacy_geom(kind);
    const size_t rows = (size_t)n_scans * g.L;
    return (2 * rows * 3 * g.n + 3 * (size_t)(g.n - 2) + (size_t)(g.n - 1)) * sizeof(double);
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
    const long long tiles = ((rows + 31) / 32) * ((g.W + 31) / 32);
    if (tiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    legacy_xyz_kernel<<<(unsigned)tiles, 256, 0, stream>>>(lon, lat, Y, rows, kind);
    const long long threads = rows * 3;
    legacy_thomas_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, stream>>>(coef, Y, M, rows, kind);
    // 1 km kinds with outputs aligned for the vector stores: a thread per tie-column interval; else per fine column
    const uintptr_t align = (kind == 2) ? 31 : 15;
    if (kind != 0 && ((((uintptr_t)out_lon | (uintptr_t)out_lat) & align) == 0)) {
        const long long blocks = n_scans * ((1354 + 127) / 128);
        if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        if (kind == 2) legacy_eval_1km_kernel<4><<<(unsigned)blocks, 128, 0, stream>>>(lon, lat, Y, M, out_lon, out_lat, n_scans);
        else legacy_eval_1km_kernel<2><<<(unsigned)blocks, 128, 0, stream>>>(lon, lat, Y, M, out_lon, out_lat, n_scans);
        return cudaGetLastError();
    }
    const long long blocks = n_scans * ((g.Wf + 127) / 128);
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    legacy_eval_kernel<<<(unsigned)blocks, 128, 0, stream>>>(lon, lat, Y, M, out_lon, out_lat, n_scans, kind);
    return cudaGetLastError();
}
