// gtp_vii.cu - VII (EPS-SG METimage) tie-point interpolation on sm_100a (second item).
//
// /root/reference/geotiepoints/viiinterpolator.py: tie points every `factor` pixels along and across the
// track, `S` tie rows per scan; the pixel grid of a scan is (S - 1) factor rows x (n_act - 1) factor columns
// and never falls between two scans (:55-62).  tie_points_interpolation (:23-79) = two linear
// xarray.interp calls, first along then across the track; tie_points_geo_interpolation (:83-134) does that
// on lon / lat directly or - any |lat| > 60 deg, longitudes spanning > 180 deg - on Cartesian coordinates of
// the sphere with the IUGG mean radius (:20) followed by _xyz2lonlat (:157-178).
//
// One thread per output pixel (coalesced 8-byte streaming stores); the four tie points of a pixel are loads
// that `factor` x `factor` neighbouring pixels share.  Cartesian mode: vii_xyz_kernel turns the tie points
// into x / y / z once (scratch: 24 B per tie point = 24 / factor^2 B per pixel), the pixel kernel blends the
// three components and back-transforms with the reference's formulas (library acos / asin / sqrt: FP64-bound,
// like gtp_generic.cu).  Algorithmic bytes per pixel: 8 n_out written + 8 n_in / factor^2 read.
#include <cuda_runtime.h>
#include <math.h>
#include "gtp_kernels.h"
#include "gtp_anchor_bt.cuh"

namespace {

constexpr double kMeanR = 6371008.7714;          // viiinterpolator.py:20

struct ViiGeom {
    long long n_alt, n_act;     // tie points along / across the track
    int S, F;                   // tie rows per scan, sub-sampling factor
    long long rows_per_scan;    // (S - 1) F
    long long out_cols;         // (n_act - 1) F
};

__device__ __forceinline__ double lerp_ref(double lo, double hi, double w)
{
    return lo + (hi - lo) * w;                    // scipy interp1d: y_lo + slope (x - x_lo), x_hi - x_lo = factor
}

// pixel (row p, column q) of the output: its upper-left tie point and the two weights
__device__ __forceinline__ void vii_locate(const ViiGeom& g, long long p, long long q, long long& tie, double& wr, double& wc)
{
    const long long scan = p / g.rows_per_scan, pl = p - scan * g.rows_per_scan;
    const long long r0 = scan * g.S + pl / g.F, c0 = q / g.F;
    wr = (double)(pl % g.F) / (double)g.F;
    wc = (double)(q % g.F) / (double)g.F;
    tie = r0 * g.n_act + c0;
}

__device__ __forceinline__ double vii_blend(const double* __restrict__ d, const ViiGeom& g, long long tie, double wr, double wc)
{
    const double left = lerp_ref(d[tie], d[tie + g.n_act], wr);                  // first along the track (:74) ...
    const double right = lerp_ref(d[tie + 1], d[tie + g.n_act + 1], wr);
    return lerp_ref(left, right, wc);                                            // ... then across (:75)
}

template <int N>
__global__ void __launch_bounds__(256)
vii_interp_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out_a,
                  double* __restrict__ out_b, ViiGeom g, long long n_px)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_px) return;
    const long long p = t / g.out_cols, q = t - p * g.out_cols;
    long long tie; double wr, wc;
    vii_locate(g, p, q, tie, wr, wc);
    __stcs(out_a + t, vii_blend(a, g, tie, wr, wc));
    if (N == 2) __stcs(out_b + t, vii_blend(b, g, tie, wr, wc));
}

__global__ void __launch_bounds__(256)
vii_xyz_kernel(const double* __restrict__ lon, const double* __restrict__ lat, double* __restrict__ xyz, long long n)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    double sl, cl, sp, cp;
    sincos(lon[t] * GTP_DEG2RAD, &sl, &cl);                                      // _lonlat2xyz, :137-154
    sincos(lat[t] * GTP_DEG2RAD, &sp, &cp);
    const double rc = kMeanR * cp;
    xyz[t] = rc * cl; xyz[n + t] = rc * sl; xyz[2 * n + t] = kMeanR * sp;
}

__global__ void __launch_bounds__(256)
vii_geo_cartesian_kernel(const double* __restrict__ lon_in, const double* __restrict__ lat_in, const double* __restrict__ xyz,
                         double* __restrict__ out_lon, double* __restrict__ out_lat, ViiGeom g, long long n_px, double z_threshold)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n_px) return;
    const long long p = t / g.out_cols, q = t - p * g.out_cols;
    long long tie; double wr, wc;
    vii_locate(g, p, q, tie, wr, wc);
    const long long n = g.n_alt * g.n_act;
    const double x = vii_blend(xyz, g, tie, wr, wc), y = vii_blend(xyz + n, g, tie, wr, wc), z = vii_blend(xyz + 2 * n, g, tie, wr, wc);
    // _xyz2lonlat (:157-178), anchor-relative around the pixel's upper-left tie point (gtp_anchor_bt.cuh)
    const AnchorBT anchor = anchor_bt_setup(kMeanR, lon_in[tie], lat_in[tie], xyz[tie], xyz[n + tie], xyz[2 * n + tie], z_threshold);
    double lon, lat;
    anchor_bt_eval(anchor, kMeanR, z_threshold, x, y, z, lon, lat);
    __stcs(out_lon + t, lon);
    __stcs(out_lat + t, lat);
}

ViiGeom make_geom(long long n_alt, long long n_act, int S, int F)
{
    ViiGeom g;
    g.n_alt = n_alt; g.n_act = n_act; g.S = S; g.F = F;
    g.rows_per_scan = (long long)(S - 1) * F;
    g.out_cols = (n_act - 1) * F;
    return g;
}

}  // namespace

size_t gtp_vii_workspace_bytes(long long n_alt, long long n_act) { return (size_t)(3 * n_alt * n_act) * sizeof(double); }

// b == nullptr: one array
cudaError_t gtp_launch_vii_interp(const double* a, const double* b, double* out_a, double* out_b,
                                  long long n_alt, long long n_act, int S, int F, cudaStream_t stream)
{
    const ViiGeom g = make_geom(n_alt, n_act, S, F);
    const long long n_px = (n_alt / S) * g.rows_per_scan * g.out_cols;
    if (n_px <= 0) return cudaSuccess;
    const long long blocks = (n_px + 255) / 256;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    if (b) vii_interp_kernel<2><<<(unsigned)blocks, 256, 0, stream>>>(a, b, out_a, out_b, g, n_px);
    else vii_interp_kernel<1><<<(unsigned)blocks, 256, 0, stream>>>(a, nullptr, out_a, nullptr, g, n_px);
    return cudaGetLastError();
}

cudaError_t gtp_launch_vii_geo_cartesian(const double* lon, const double* lat, double* out_lon, double* out_lat,
                                         long long n_alt, long long n_act, int S, int F, double z_threshold,
                                         void* workspace, cudaStream_t stream)
{
    const ViiGeom g = make_geom(n_alt, n_act, S, F);
    const long long n_px = (n_alt / S) * g.rows_per_scan * g.out_cols, n = n_alt * n_act;
    if (n_px <= 0) return cudaSuccess;
    const long long blocks = (n_px + 255) / 256;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    vii_xyz_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(lon, lat, (double*)workspace, n);
    vii_geo_cartesian_kernel<<<(unsigned)blocks, 256, 0, stream>>>(lon, lat, (const double*)workspace, out_lon, out_lat, g, n_px, z_threshold);
    return cudaGetLastError();
}
