// gtp_generic.cu - rounding-faithful kernels for every configuration and dtype.
//
// One CTA owns one (scan, strip of tie-point columns).  It stages the Cartesian tie
// points of the strip (all tie rows of the scan, plus one halo column) and the
// expansion/alignment coefficients of its tie-point quads in shared memory, then
// streams the fine pixels of the strip row by row with fully coalesced stores.
// Every input element is read once from HBM (the halo column is an L2 hit) and every
// output element written once: 17.5 B (f64) / 8.75 B (f32) of HBM traffic per output
// pixel for 1 km -> 250 m.
//
// These kernels reproduce the reference's ROUNDING POINTS, not just its formulas
// (SURVEY.md 8a "precision model"): storage in T, every expression holding a double
// literal or a libm call evaluated in double and rounded to T at the assignment,
// float*float kept in float, no FMA contraction (this file is compiled with
// -fmad=false).  That is what keeps the float32 results within one float32 ulp of
// the reference.  It is FP64-pipe bound; the float64 1 km fast path lives in
// gtp_fast.cu and falls back to this one where its preconditions do not hold.
//
// Reference lines restated here (paths relative to /root/reference/geotiepoints):
//   tie_xyz            _modis_utils.pyx:26-40        lonlat2xyz
//   xyz_to_lonlat      _modis_utils.pyx:43-65        xyz2lonlat
//   quad_exp_ali       _modis_interpolator.pyx:41-82 _compute_expansion_alignment
//   cviirs pixel body  _modis_interpolator.pyx:328-344, :379-398
//   simple pixel body  _simple_modis_interpolator.pyx:71-277 + scipy map_coordinates(order=1,'nearest')
#include <type_traits>
#include <cuda_runtime.h>
#include <math.h>
#include "gtp_common.cuh"
#include "gtp_kernels.h"
#include "gtp_fast_f32_math.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kStripQuads = 64;   // tie-point quads (columns) per CTA

__device__ __forceinline__ int sign_of(double x) { return x > 0.0 ? 1 : (x < 0.0 ? -1 : 0); }

template <typename T>
__device__ __forceinline__ T deg2rad_t(T x) { return (T)((double)x * GTP_DEG2RAD); }

template <typename T>
__device__ __forceinline__ void tie_xyz(T lon, T lat, T& x, T& y, T& z)
{
    T lon_rad = deg2rad_t(lon), lat_rad = deg2rad_t(lat);
    double slon, clon, slat, clat;
    sincos((double)lon_rad, &slon, &clon);
    sincos((double)lat_rad, &slat, &clat);
    x = (T)((GTP_EARTH_RADIUS * clat) * clon);
    y = (T)((GTP_EARTH_RADIUS * clat) * slon);
    z = (T)(GTP_EARTH_RADIUS * slat);
}

template <typename T>
__device__ __forceinline__ void xyz_to_lonlat(T xt, T yt, T zt, T& lon, T& lat)
{
    double x = (double)xt, y = (double)yt, z = (double)zt;
    T thr = (T)0.8;
    double rho = sqrt(x * x + y * y);
    lon = (T)((acos(x / rho) * GTP_RAD2DEG) * (double)sign_of(y));
    if ((z < ((double)thr * GTP_EARTH_RADIUS)) && (z > ((-1.0 * (double)thr) * GTP_EARTH_RADIUS)))
        lat = (T)(90.0 - acos(z / GTP_EARTH_RADIUS) * GTP_RAD2DEG);
    else
        lat = (T)((double)sign_of(z) * (90.0 - asin(rho / GTP_EARTH_RADIUS) * GTP_RAD2DEG));
}

template <typename T>
__device__ __forceinline__ void quad_exp_ali(T satz_a, T satz_b, int km, T& c_exp, T& c_ali)
{
    T za = deg2rad_t(satz_a), zb = deg2rad_t(satz_b);
    T phi_a = (T)asin((GTP_RK * sin((double)za)) / (GTP_RK + GTP_H));
    T phi_b = (T)asin((GTP_RK * sin((double)zb)) / (GTP_RK + GTP_H));
    T theta_a = za - phi_a;
    T theta_b = zb - phi_b;
    T phi = (T)((double)(T)(phi_a + phi_b) / 2.0);
    double sphi, cphi;
    sincos((double)phi, &sphi, &cphi);
    T zeta = (T)asin(((GTP_RK + GTP_H) * sphi) / GTP_RK);
    T theta = zeta - phi;
    T den = (theta_a == theta_b) ? (T)((double)theta_a * 2.0) : (T)(theta_a - theta_b);
    c_exp = (T)(4.0 * ((((double)(T)(theta_a + theta_b) / 2.0) - (double)theta) / (double)den));
    T sin_beta_2 = (T)((double)km / (2.0 * GTP_H));
    double szeta, czeta;
    sincos((double)zeta, &szeta, &czeta);
    T d = (T)(((((GTP_RK + GTP_H) / GTP_RK) * cphi) - czeta) * (double)sin_beta_2);
    T dd = d * d;   // powf(d, 2) for float32, pow(d, 2) for float64 (:68)
    T e = (T)(czeta - sqrt((czeta * czeta) - (double)dd));
    c_ali = (T)(((4.0 * (double)e) * szeta) / (double)den);
}

// ---------------------------------------------------------------------------------
// CVIIRS (satellite-zenith driven) interpolator, any configuration.
// ---------------------------------------------------------------------------------
// SERIES (float32 5 km -> 1 km only): the back-transform of a pixel is the anchor-relative series of
// gtp_fast_f32_math.cuh around the float32-rounded corner A of its quad (one QuadF32 per quad in shared
// memory, set up once per strip) instead of acos / asin / sqrt / two divisions per pixel; quads the
// series cannot take (6 % of a day's, next to the poles) keep the library formulas.  The blend above it
// is unchanged, so the float32 fine xyz - and with them the result - stay the reference's.
// XYZ = true stores the fine Cartesian coordinates (_compute_fine_xyz, :346-398) in out_lon / out_lat /
// out_z = x / y / z instead of their back-transform.
template <typename T, bool XYZ, bool SERIES>
__global__ void __launch_bounds__(kThreads)
cviirs_generic_kernel(const T* __restrict__ lon, const T* __restrict__ lat, const T* __restrict__ satz,
                      T* __restrict__ out_lon, T* __restrict__ out_lat, T* __restrict__ out_z,
                      long long n_scans, int n_strips, CviirsCfg cfg)
{
    static_assert(!SERIES || (std::is_same<T, float>::value && !XYZ), "the series back-transform is float32 lon / lat only");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = cfg.L, W = cfg.W;
    const int TC = kStripQuads + 1;                 // tie columns held (with halo)
    T* sx = reinterpret_cast<T*>(smem_raw);         // [L][TC]
    T* sy = sx + L * TC;
    T* sz = sy + L * TC;
    T* sce = sz + L * TC;                           // [L-1][kStripQuads]
    T* sca = sce + (L - 1) * kStripQuads;
    // SERIES: one QuadF32 per quad behind the coefficient planes (8-byte aligned: an even number of floats precedes it)
    gtpfast32::QuadF32* sq = reinterpret_cast<gtpfast32::QuadF32*>(sca + (L - 1) * kStripQuads);

    const long long scan = blockIdx.x / n_strips;
    const int strip = (int)(blockIdx.x % n_strips);
    if (scan >= n_scans) return;
    const int q0 = strip * kStripQuads;             // first quad column of this CTA
    const int nq = min(kStripQuads, (W - 1) - q0);  // quads in this CTA
    const int nt = nq + 1;                          // tie columns needed

    const T* slon = lon + scan * (long long)L * W;
    const T* slat = lat + scan * (long long)L * W;
    const T* ssatz = satz + scan * (long long)L * W;

    for (int t = threadIdx.x; t < L * nt; t += kThreads) {
        int r = t / nt, c = t - r * nt;
        T x, y, z;
        tie_xyz(slon[r * W + q0 + c], slat[r * W + q0 + c], x, y, z);
        sx[r * TC + c] = x; sy[r * TC + c] = y; sz[r * TC + c] = z;
    }
    for (int t = threadIdx.x; t < (L - 1) * nq; t += kThreads) {
        int r = t / nq, c = t - r * nq;
        T ce, ca;
        quad_exp_ali(ssatz[r * W + q0 + c], ssatz[r * W + q0 + c + 1], cfg.km, ce, ca);
        sce[r * kStripQuads + c] = ce; sca[r * kStripQuads + c] = ca;
    }
    __syncthreads();
    if constexpr (SERIES) {
        for (int t = threadIdx.x; t < (L - 1) * nq; t += kThreads) {
            int r = t / nq, c = t - r * nq;
            const gtpfast32::TieGeoF A = gtpfast32::tie_geo_f32(slon[r * W + q0 + c], slat[r * W + q0 + c]);
            gtpfast32::Vec3F B, C, D;
            B.x = sx[r * TC + c + 1]; B.y = sy[r * TC + c + 1]; B.z = sz[r * TC + c + 1];
            C.x = sx[(r + 1) * TC + c + 1]; C.y = sy[(r + 1) * TC + c + 1]; C.z = sz[(r + 1) * TC + c + 1];
            D.x = sx[(r + 1) * TC + c]; D.y = sy[(r + 1) * TC + c]; D.z = sz[(r + 1) * TC + c];
            sq[r * kStripQuads + c] = gtpfast32::quad_setup_f32(A, B, C, D, sce[r * kStripQuads + c], sca[r * kStripQuads + c],
                                                                gtpfast32::kMaxRelHoriz5, gtpfast32::kMaxRelZ5);
        }
        __syncthreads();
    }

    const int i_lo = gtp_quad_col_begin(cfg, q0);
    const int i_hi = (q0 + nq >= W - 1) ? cfg.Wf : gtp_quad_col_begin(cfg, q0 + nq);
    const int ncols = i_hi - i_lo;
    const T fT = (T)cfg.f;
    T* olon = out_lon + scan * (long long)cfg.n * cfg.Wf;
    T* olat = out_lat + scan * (long long)cfg.n * cfg.Wf;

    if constexpr (SERIES) {
        // 5 km -> 1 km float32 (f = 5, one quad row per scan): the same rounding points written as in
        // gtp_fast_f32_math.cuh (blend_f32: 22 instead of 44 float <-> double conversions per pixel, which
        // is what bounds the loop below), constants instead of run-time divisions, series back-transform
        constexpr int kInner = 5 * kStripQuads;
        for (int t = threadIdx.x; t < 10 * ncols; t += kThreads) {
            int j, ii;
            if (ncols == kInner) { j = t / kInner; ii = t - j * kInner; }
            else { j = t / ncols; ii = t - j * ncols; }
            const int i = i_lo + ii;
            const int qc = gtpfast::quad_col_5km(i, W);                                 // a7
            const int c = qc - q0;
            const float s_s = (float)(i - 2 - 5 * qc) / 5.0f;                           // a5, :341
            const float s_t = (float)(j - 2) / 5.0f;                                    // :342
            const gtpfast32::QuadF32& q = sq[c];
            const gtpfast32::RowF32 rc = gtpfast32::row_setup_f32(q, s_t);
            const double k1d = gtpfast32::mul_rn((double)s_s, gtpfast32::add_rn(1.0, -(double)s_s));
            double v[3];
            gtpfast32::blend_f32(q, rc, s_s, k1d, v);
            float lo, la;
            if (q.mode != gtpfast::kSlow) gtpfast32::backtransform_f32<gtpfast::kBoth, true>(q, v, lo, la);
            else xyz_to_lonlat((float)v[0], (float)v[1], (float)v[2], lo, la);
            olon[(long long)j * cfg.Wf + i] = lo;
            olat[(long long)j * cfg.Wf + i] = la;
        }
        return;
    }

    for (int t = threadIdx.x; t < cfg.n * ncols; t += kThreads) {
        int j = t / ncols, i = i_lo + (t - j * ncols);
        int r = gtp_quad_row(cfg, j);
        int c = gtp_quad_col(cfg, i) - q0;
        T s_s = (T)gtp_coord_x(cfg, i) / fT;                       // :341
        T s_t = ((T)gtp_coord_y2(cfg, j) * (T)0.5) / fT;           // :342
        T ce = sce[r * kStripQuads + c], ca = sca[r * kStripQuads + c];
        T a_track = s_t;
        T a_scan = (T)(((double)s_s + (((double)s_s * (1.0 - (double)s_s)) * (double)ce))
                       + (((double)s_t * (1.0 - (double)s_t)) * (double)ca));   // :344
        const T* comp[3] = {sx, sy, sz};
        T fine[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            T A = comp[k][r * TC + c], B = comp[k][r * TC + c + 1];
            T D = comp[k][(r + 1) * TC + c], C = comp[k][(r + 1) * TC + c + 1];
            // :396-398  (1.0 - a) * A is a double product, a * B a T product
            T scan1 = (T)(((1.0 - (double)a_scan) * (double)A) + (double)(T)(a_scan * B));
            T scan2 = (T)(((1.0 - (double)a_scan) * (double)D) + (double)(T)(a_scan * C));
            fine[k] = (T)(((1.0 - (double)a_track) * (double)scan1) + (double)(T)(a_track * scan2));
        }
        if (XYZ) {
            olon[(long long)j * cfg.Wf + i] = fine[0];
            olat[(long long)j * cfg.Wf + i] = fine[1];
            out_z[(scan * cfg.n + j) * cfg.Wf + i] = fine[2];
        } else {
            T lo, la;
            xyz_to_lonlat(fine[0], fine[1], fine[2], lo, la);
            olon[(long long)j * cfg.Wf + i] = lo;
            olat[(long long)j * cfg.Wf + i] = la;
        }
    }
}

// ---------------------------------------------------------------------------------
// Simple (lon/lat only) interpolator: order-1 map_coordinates + edge extrapolation.
// ---------------------------------------------------------------------------------
template <typename T>
struct SimpleTile {
    const T* v[3];   // shared-memory component planes [10][TC]
    int TC;          // pitch
    int q0;          // first tie column held
    int W;           // tie columns of the granule
    int res;         // 2 | 4
    int Wf;          // fine columns
};

// scipy.ndimage.map_coordinates(order=1, mode='nearest') sample of component k at
// fine row j / fine column i of the scan: the coordinate is not clamped, the two taps
// per axis are (scipy >= 1.6); accumulation ((v*wy)*wx) in tap order, in double.
template <typename T>
__device__ __forceinline__ T map_coord(const SimpleTile<T>& t, int k, int j, int i)
{
    T yc = (T)(((double)j * (1.0 / (double)t.res)) - (((double)t.res * (1.0 / 16.0)) + (1.0 / 8.0)));
    T xc = (T)((double)i * (1.0 / (double)t.res));
    double cy = (double)yc, cx = (double)xc;
    double fly = floor(cy), flx = floor(cx);
    int sy = (int)fly, sx = (int)flx;
    double wy0 = 1.0 - (cy - fly), wx0 = 1.0 - (cx - flx);
    double wy1 = 1.0 - wy0, wx1 = 1.0 - wx0;
    int y0 = max(sy, 0), y1 = min(sy + 1, 9);
    y0 = min(y0, 9); y1 = max(y1, 0);
    int x0 = min(max(sx, 0), t.W - 1) - t.q0, x1 = min(max(sx + 1, 0), t.W - 1) - t.q0;
    const T* v = t.v[k];
    double acc = 0.0;
    acc = acc + (((double)v[y0 * t.TC + x0] * wy0) * wx0);
    acc = acc + (((double)v[y0 * t.TC + x1] * wy0) * wx1);
    acc = acc + (((double)v[y1 * t.TC + x0] * wy1) * wx0);
    acc = acc + (((double)v[y1 * t.TC + x1] * wy1) * wx1);
    return (T)acc;
}

// after _extrapolate_xyz_rightmost_columns (:139-150)
template <typename T>
__device__ __forceinline__ T col_extrap(const SimpleTile<T>& t, int k, int j, int i)
{
    const int nc = (t.res == 4) ? 3 : 1;
    T v = map_coord(t, k, j, i);
    if (i >= t.Wf - nc) {
        T diff = map_coord(t, k, j, t.Wf - nc - 1) - map_coord(t, k, j, t.Wf - nc - 2);
        v = v + diff * (T)(i - (t.Wf - nc) + 1);
    }
    return v;
}

template <typename T>
__device__ __forceinline__ T simple_ycoord(int res, int j)
{
    return (T)(((double)j * (1.0 / (double)res)) - (((double)res * (1.0 / 16.0)) + (1.0 / 8.0)));
}

// after _interpolate_xyz_250 / _interpolate_xyz_500 (:157-277)
template <typename T>
__device__ __forceinline__ T simple_value(const SimpleTile<T>& t, int k, int j, int i)
{
    int lo, hi;   // the two fine rows the extrapolation line goes through
    if (t.res == 4) {
        if (j < 2) { lo = 2; hi = 5; } else if (j >= 38) { lo = 34; hi = 37; } else return col_extrap(t, k, j, i);
    } else {
        if (j < 1) { lo = 1; hi = 2; } else if (j >= 19) { lo = 17; hi = 18; } else return col_extrap(t, k, j, i);
    }
    T vlo = col_extrap(t, k, lo, i), vhi = col_extrap(t, k, hi, i);
    T ylo = simple_ycoord<T>(t.res, lo), yhi = simple_ycoord<T>(t.res, hi), yj = simple_ycoord<T>(t.res, j);
    T m = (vhi - vlo) / (yhi - ylo);
    T b = vhi - m * yhi;
    return m * yj + b;
}

template <typename T, bool XYZ>
__global__ void __launch_bounds__(kThreads)
simple_generic_kernel(const T* __restrict__ lon, const T* __restrict__ lat,
                      T* __restrict__ out_lon, T* __restrict__ out_lat, T* __restrict__ out_z,
                      long long n_scans, int n_strips, int W, int res)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // tie columns held: the strip, one halo column to the right, and for the strip that
    // owns the right edge the two columns left of it that the column extrapolation reads
    const int TC = kStripQuads + 4;
    T* sx = reinterpret_cast<T*>(smem_raw);
    T* sy = sx + 10 * TC;
    T* sz = sy + 10 * TC;

    const long long scan = blockIdx.x / n_strips;
    const int strip = (int)(blockIdx.x % n_strips);
    if (scan >= n_scans) return;
    const int Wf = W * res;
    const int c_first = strip * kStripQuads;                     // first tie column owned
    const int c_last = min(c_first + kStripQuads, W);            // one past the last owned
    const int q0 = max(c_first - 2, 0);                          // first tie column held
    const int q1 = min(c_last + 1, W);                           // one past the last held
    const int nt = q1 - q0;

    const T* slon = lon + scan * 10LL * W;
    const T* slat = lat + scan * 10LL * W;
    for (int t = threadIdx.x; t < 10 * nt; t += kThreads) {
        int r = t / nt, c = t - r * nt;
        T x, y, z;
        tie_xyz(slon[r * W + q0 + c], slat[r * W + q0 + c], x, y, z);
        sx[r * TC + c] = x; sy[r * TC + c] = y; sz[r * TC + c] = z;
    }
    __syncthreads();

    SimpleTile<T> tile;
    tile.v[0] = sx; tile.v[1] = sy; tile.v[2] = sz;
    tile.TC = TC; tile.q0 = q0; tile.W = W; tile.res = res; tile.Wf = Wf;

    const int n = 10 * res;
    const int i_lo = c_first * res, i_hi = c_last * res;
    const int ncols = i_hi - i_lo;
    T* olon = out_lon + scan * (long long)n * Wf;
    T* olat = out_lat + scan * (long long)n * Wf;
    for (int t = threadIdx.x; t < n * ncols; t += kThreads) {
        int j = t / ncols, i = i_lo + (t - j * ncols);
        T fx = simple_value(tile, 0, j, i);
        T fy = simple_value(tile, 1, j, i);
        T fz = simple_value(tile, 2, j, i);
        if (XYZ) {
            olon[(long long)j * Wf + i] = fx;
            olat[(long long)j * Wf + i] = fy;
            out_z[(scan * n + j) * Wf + i] = fz;
        } else {
            T lo, la;
            xyz_to_lonlat(fx, fy, fz, lo, la);
            olon[(long long)j * Wf + i] = lo;
            olat[(long long)j * Wf + i] = la;
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------
// launchers (declared in gtp_kernels.h)
// ---------------------------------------------------------------------------------
namespace {

template <typename T, bool XYZ>
cudaError_t launch_cviirs(const T* lon, const T* lat, const T* satz, T* out_a, T* out_b, T* out_c,
                          long long n_scans, const CviirsCfg& cfg, cudaStream_t stream, bool allow_series)
{
    if (n_scans <= 0) return cudaSuccess;
    const int n_strips = (cfg.W - 1 + kStripQuads - 1) / kStripQuads;
    const size_t smem = sizeof(T) * (3 * cfg.L * (kStripQuads + 1) + 2 * (cfg.L - 1) * kStripQuads);
    const long long blocks = n_scans * n_strips;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    if constexpr (std::is_same<T, float>::value && !XYZ) {
        // 5 km -> 1 km float32: series back-transform.  Its acceptance limits assume the coordinate range of
        // f = 5 (|a_track| <= 1.4, |a_scan| <= 2.3); 5 km -> 500 | 250 m reach further (a5: not scaled with f)
        if (allow_series && cfg.km == 5 && cfg.p == 1) {
            static_assert(sizeof(gtpfast32::QuadF32) % 8 == 0 && (3 * (kStripQuads + 1) * 2 + 2 * kStripQuads) % 2 == 0, "QuadF32 alignment");
            const size_t smem_series = smem + sizeof(gtpfast32::QuadF32) * (cfg.L - 1) * kStripQuads;
            cviirs_generic_kernel<T, false, true><<<(unsigned)blocks, kThreads, smem_series, stream>>>(
                lon, lat, satz, out_a, out_b, out_c, n_scans, n_strips, cfg);
            return cudaGetLastError();
        }
    }
    cviirs_generic_kernel<T, XYZ, false><<<(unsigned)blocks, kThreads, smem, stream>>>(
        lon, lat, satz, out_a, out_b, out_c, n_scans, n_strips, cfg);
    return cudaGetLastError();
}

template <typename T, bool XYZ>
cudaError_t launch_simple(const T* lon, const T* lat, T* out_a, T* out_b, T* out_c,
                          long long n_scans, int width, int res, cudaStream_t stream)
{
    if (n_scans <= 0) return cudaSuccess;
    const int n_strips = (width + kStripQuads - 1) / kStripQuads;
    const size_t smem = sizeof(T) * 3 * 10 * (kStripQuads + 4);
    const long long blocks = n_scans * n_strips;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    simple_generic_kernel<T, XYZ><<<(unsigned)blocks, kThreads, smem, stream>>>(
        lon, lat, out_a, out_b, out_c, n_scans, n_strips, width, res);
    return cudaGetLastError();
}

}  // namespace

template <typename T>
cudaError_t gtp_launch_cviirs_generic(const T* lon, const T* lat, const T* satz, T* out_lon, T* out_lat,
                                      long long n_scans, const CviirsCfg& cfg, cudaStream_t stream, bool allow_series)
{
    return launch_cviirs<T, false>(lon, lat, satz, out_lon, out_lat, (T*)nullptr, n_scans, cfg, stream, allow_series);
}

template <typename T>
cudaError_t gtp_launch_cviirs_generic_xyz(const T* lon, const T* lat, const T* satz, T* out_x, T* out_y, T* out_z,
                                          long long n_scans, const CviirsCfg& cfg, cudaStream_t stream)
{
    return launch_cviirs<T, true>(lon, lat, satz, out_x, out_y, out_z, n_scans, cfg, stream, false);
}

template <typename T>
cudaError_t gtp_launch_simple_generic(const T* lon, const T* lat, T* out_lon, T* out_lat,
                                      long long n_scans, int width, int res, cudaStream_t stream)
{
    return launch_simple<T, false>(lon, lat, out_lon, out_lat, (T*)nullptr, n_scans, width, res, stream);
}

template <typename T>
cudaError_t gtp_launch_simple_generic_xyz(const T* lon, const T* lat, T* out_x, T* out_y, T* out_z,
                                          long long n_scans, int width, int res, cudaStream_t stream)
{
    return launch_simple<T, true>(lon, lat, out_x, out_y, out_z, n_scans, width, res, stream);
}

template cudaError_t gtp_launch_cviirs_generic<float>(const float*, const float*, const float*, float*, float*,
                                                      long long, const CviirsCfg&, cudaStream_t, bool);
template cudaError_t gtp_launch_cviirs_generic<double>(const double*, const double*, const double*, double*, double*,
                                                       long long, const CviirsCfg&, cudaStream_t, bool);
template cudaError_t gtp_launch_cviirs_generic_xyz<float>(const float*, const float*, const float*, float*, float*, float*,
                                                          long long, const CviirsCfg&, cudaStream_t);
template cudaError_t gtp_launch_cviirs_generic_xyz<double>(const double*, const double*, const double*, double*, double*, double*,
                                                           long long, const CviirsCfg&, cudaStream_t);
template cudaError_t gtp_launch_simple_generic_xyz<float>(const float*, const float*, float*, float*, float*,
                                                          long long, int, int, cudaStream_t);
template cudaError_t gtp_launch_simple_generic_xyz<double>(const double*, const double*, double*, double*, double*,
                                                           long long, int, int, cudaStream_t);
template cudaError_t gtp_launch_simple_generic<float>(const float*, const float*, float*, float*,
                                                      long long, int, int, cudaStream_t);
template cudaError_t gtp_launch_simple_generic<double>(const double*, const double*, double*, double*,
                                                       long long, int, int, cudaStream_t);
