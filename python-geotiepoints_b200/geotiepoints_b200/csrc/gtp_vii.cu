// gtp_vii.cu - VII (EPS-SG METimage) tie-point interpolation on sm_100a (second item).
//
// /root/reference/geotiepoints/viiinterpolator.py: tie points every `factor` pixels along and across the
// track, `S` tie rows per scan; the pixel grid of a scan is (S - 1) factor rows x (n_act - 1) factor columns
// and never falls between two scans (:55-62).  tie_points_interpolation (:23-79) = two linear
// xarray.interp calls, first along then across the track; tie_points_geo_interpolation (:83-134) does that
// on lon / lat directly or - any |lat| > 60 deg, longitudes spanning > 180 deg - on Cartesian coordinates of
// the sphere with the IUGG mean radius (:20) followed by _xyz2lonlat (:157-178).
//
// Plain arrays (vii_interp_kernel) and other factors: one thread per output pixel (coalesced 8-byte streaming
// stores); the four tie points of a pixel are loads that `factor` x `factor` neighbouring pixels share.  Cartesian
// mode: vii_xyz_kernel turns the tie points into x / y / z and into anchors of the back-transform once (scratch:
// 88 B per tie point = 88 / factor^2 B per pixel); factors 4 and 8: vii_geo_cells_kernel, a thread per (output row,
// tie cell) - row weight, along-track blends and anchor once per cell, 256-bit stores -, else
// vii_geo_cartesian_kernel per pixel; both back-transform anchor-relative (gtp_anchor_bt.cuh).
// Algorithmic bytes per pixel: 8 n_out written + 8 n_in / factor^2 read.
#include <cstdint>
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

// pixel (row p, column q) of the output: its upper-left tie point and the two weights.  32-bit arithmetic: a
// 64-bit division is ~100 instructions, and six of them per pixel were most of the first version of this kernel.
__device__ __forceinline__ void vii_locate(const ViiGeom& g, unsigned p, unsigned q, long long& tie, double& wr, double& wc)
{
    const unsigned rps = (unsigned)g.rows_per_scan, F = (unsigned)g.F;
    const unsigned scan = p / rps, pl = p - scan * rps;
    const unsigned rl = pl / F, c0 = q / F;
    wr = (double)(pl - rl * F) / (double)g.F;
    wc = (double)(q - c0 * F) / (double)g.F;
    tie = ((long long)scan * g.S + rl) * g.n_act + c0;
}

// one CTA row per output row (blockIdx.x), 256 columns per CTA (blockIdx.y)
__device__ __forceinline__ bool vii_pixel(const ViiGeom& g, unsigned& p, unsigned& q, long long& t)
{
    p = blockIdx.x; q = blockIdx.y * blockDim.x + threadIdx.x;
    t = (long long)p * g.out_cols + q;
    return q < (unsigned)g.out_cols;
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
    unsigned p, q; long long t;
    if (!vii_pixel(g, p, q, t)) return;
    long long tie; double wr, wc;
    vii_locate(g, p, q, tie, wr, wc);
    __stcs(out_a + t, vii_blend(a, g, tie, wr, wc));
    if (N == 2) __stcs(out_b + t, vii_blend(b, g, tie, wr, wc));
}

__global__ void __launch_bounds__(256)
vii_xyz_kernel(const double* __restrict__ lon, const double* __restrict__ lat, double* __restrict__ xyz, long long n,
               double z_threshold)
{
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= n) return;
    double sl, cl, sp, cp;
    sincos(lon[t] * GTP_DEG2RAD, &sl, &cl);                                      // _lonlat2xyz, :137-154
    sincos(lat[t] * GTP_DEG2RAD, &sp, &cp);
    const double rc = kMeanR * cp;
    const double x = rc * cl, y = rc * sl, z = kMeanR * sp;
    xyz[t] = x; xyz[n + t] = y; xyz[2 * n + t] = z;
    // the tie point as an anchor of the back-transform (gtp_anchor_bt.cuh), shared by the factor^2 pixels of its cell
    anchor_bt_pack(anchor_bt_setup(kMeanR, lon[t], lat[t], x, y, z, z_threshold), kMeanR, z_threshold, xyz + 3 * n, n, t);
}

__global__ void __launch_bounds__(256)
vii_geo_cartesian_kernel(const double* __restrict__ lon_in, const double* __restrict__ lat_in, const double* __restrict__ xyz,
                         double* __restrict__ out_lon, double* __restrict__ out_lat, ViiGeom g, long long n_px, double z_threshold)
{
    unsigned p, q; long long t;
    if (!vii_pixel(g, p, q, t)) return;
    long long tie; double wr, wc;
    vii_locate(g, p, q, tie, wr, wc);
    const long long n = g.n_alt * g.n_act;
    const double x = vii_blend(xyz, g, tie, wr, wc), y = vii_blend(xyz + n, g, tie, wr, wc), z = vii_blend(xyz + 2 * n, g, tie, wr, wc);
    // _xyz2lonlat (:157-178), anchor-relative around the pixel's upper-left tie point (gtp_anchor_bt.cuh)
    const AnchorBT anchor = anchor_bt_unpack(xyz + 3 * n, n, tie, lon_in[tie], lat_in[tie], xyz[tie], xyz[n + tie], xyz[2 * n + tie]);
    double lon, lat;
    anchor_bt_eval(anchor, kMeanR, z_threshold, x, y, z, lon, lat);
    __stcs(out_lon + t, lon);
    __stcs(out_lat + t, lat);
}

// Factor 4 | 8 (METimage's is 8), outputs aligned for 256-bit stores: one thread per (output row, tie cell) = F
// pixels.  The cell's anchor, its row weight and the two along-track blends per component are set up once, a pixel
// is 6 FP64 operations + the back-transform, and a thread stores 32-byte pieces.  The per-pixel kernel above spends
// 320 instructions per pixel, 240 of them on indices, weights and the loads of the four tie points and the anchor.
template <int F>
__global__ void __launch_bounds__(128, 8)
vii_geo_cells_kernel(const double* __restrict__ lon_in, const double* __restrict__ lat_in, const double* __restrict__ xyz,
                     double* __restrict__ out_lon, double* __restrict__ out_lat, ViiGeom g, double z_threshold)
{
    const unsigned p = blockIdx.x, c0 = blockIdx.y * blockDim.x + threadIdx.x;
    if (c0 >= (unsigned)(g.n_act - 1)) return;
    const unsigned rps = (unsigned)g.rows_per_scan;
    const unsigned scan = p / rps, pl = p - scan * rps, rl = pl / F;
    const double wr = (double)(pl - rl * F) / (double)F;
    const long long tie = ((long long)scan * g.S + rl) * g.n_act + c0, n = g.n_alt * g.n_act;
    double left[3], slope[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double* d = xyz + c * n + tie;
        left[c] = lerp_ref(d[0], d[g.n_act], wr);                                // vii_blend, :74
        slope[c] = lerp_ref(d[1], d[g.n_act + 1], wr) - left[c];
    }
    const AnchorBT anchor = anchor_bt_unpack(xyz + 3 * n, n, tie, lon_in[tie], lat_in[tie], xyz[tie], xyz[n + tie], xyz[2 * n + tie]);
    double lo[F], la[F];
#pragma unroll
    for (int k = 0; k < F; ++k) {
        const double wc = (double)k / (double)F;                                 // :75
        anchor_bt_eval(anchor, kMeanR, z_threshold, left[0] + slope[0] * wc, left[1] + slope[1] * wc, left[2] + slope[2] * wc,
                       lo[k], la[k]);
    }
    const long long t = (long long)p * g.out_cols + (long long)c0 * F;
#pragma unroll
    for (int k = 0; k < F; k += 4) {
        asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(out_lon + t + k), "d"(lo[k]), "d"(lo[k + 1]), "d"(lo[k + 2]), "d"(lo[k + 3]) : "memory");
        asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(out_lat + t + k), "d"(la[k]), "d"(la[k + 1]), "d"(la[k + 2]), "d"(la[k + 3]) : "memory");
    }
}

ViiGeom make_geom(long long n_alt, long long n_act, int S, int F)
{
    ViiGeom g;
    g.n_alt = n_alt; g.n_act = n_act; g.S = S; g.F = F;
    g.rows_per_scan = (long long)(S - 1) * F;
    g.out_cols = (n_act - 1) * F;
    return g;
}

// rows in grid.x (< 2^31), 256-column pieces in grid.y (< 65536)
bool vii_grid(const ViiGeom& g, long long n_px, dim3& blocks)
{
    const long long rows = n_px / g.out_cols, pieces = (g.out_cols + 255) / 256;
    if (rows > 0x7fffffffLL || pieces > 65535 || g.out_cols > 0x7fffffffLL) return false;
    blocks = dim3((unsigned)rows, (unsigned)pieces, 1);
    return true;
}

}  // namespace

size_t gtp_vii_workspace_bytes(long long n_alt, long long n_act) { return (size_t)((3 + kAnchorPacked) * n_alt * n_act) * sizeof(double); }

// b == nullptr: one array
cudaError_t gtp_launch_vii_interp(const double* a, const double* b, double* out_a, double* out_b,
                                  long long n_alt, long long n_act, int S, int F, cudaStream_t stream)
{
    const ViiGeom g = make_geom(n_alt, n_act, S, F);
    const long long n_px = (n_alt / S) * g.rows_per_scan * g.out_cols;
    if (n_px <= 0) return cudaSuccess;
    dim3 blocks;
    if (!vii_grid(g, n_px, blocks)) return cudaErrorInvalidConfiguration;
    if (b) vii_interp_kernel<2><<<blocks, 256, 0, stream>>>(a, b, out_a, out_b, g, n_px);
    else vii_interp_kernel<1><<<blocks, 256, 0, stream>>>(a, nullptr, out_a, nullptr, g, n_px);
    return cudaGetLastError();
}

cudaError_t gtp_launch_vii_geo_cartesian(const double* lon, const double* lat, double* out_lon, double* out_lat,
                                         long long n_alt, long long n_act, int S, int F, double z_threshold,
                                         void* workspace, cudaStream_t stream)
{
    const ViiGeom g = make_geom(n_alt, n_act, S, F);
    const long long n_px = (n_alt / S) * g.rows_per_scan * g.out_cols, n = n_alt * n_act;
    if (n_px <= 0) return cudaSuccess;
    dim3 blocks;
    if (!vii_grid(g, n_px, blocks)) return cudaErrorInvalidConfiguration;
    vii_xyz_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(lon, lat, (double*)workspace, n, z_threshold);
    const long long rows = n_px / g.out_cols, cell_pieces = (n_act - 1 + 127) / 128;
    if ((F == 4 || F == 8) && ((((uintptr_t)out_lon | (uintptr_t)out_lat) & 31) == 0) && cell_pieces <= 65535) {
        const dim3 cells((unsigned)rows, (unsigned)cell_pieces, 1);
        if (F == 4) vii_geo_cells_kernel<4><<<cells, 128, 0, stream>>>(lon, lat, (const double*)workspace, out_lon, out_lat, g, z_threshold);
        else vii_geo_cells_kernel<8><<<cells, 128, 0, stream>>>(lon, lat, (const double*)workspace, out_lon, out_lat, g, z_threshold);
        return cudaGetLastError();
    }
    vii_geo_cartesian_kernel<<<blocks, 256, 0, stream>>>(lon, lat, (const double*)workspace, out_lon, out_lat, g, n_px, z_threshold);
    return cudaGetLastError();
}
