// gtp_anchor_bt.cuh - anchor-relative xyz -> lon / lat from a displacement (gtp_legacy.cu, gtp_vii.cu, gtp_grid.cu).
//
// The idea of gtp_fast_math.cuh for interpolators that produce plain Cartesian coordinates: every fine pixel
// lies within ~1.5 tie intervals of a tie point A whose lon / lat are INPUTS.  With (e, h) the displacement in
// A's horizontal frame scaled by 1 / rho_A and dz the one along the polar axis:
//   lon = lon_A + atan(e / (1 + h)),   lat = lat_A + series in dz (|z| < thr R) or in rho / rho_A - 1.
// ~30 FP64 instructions per pixel instead of ~370 for acos / asin / sqrt / two divisions, accurate to ~1e-13 deg
// while |e|, |h| <= 0.01 and |dz| <= 2e-3 R (checked per pixel; beyond that - next to the poles - and for
// anchors outside [-180, 180] x [-90, 90] the reference's formulas are evaluated as they are:
// geointerpolator.py:63-75 / viiinterpolator.py:157-178, which differ in the radius only).
#pragma once
#include "gtp_fast_math.cuh"

struct AnchorBT {
    double x, y, z, ue0, ue1, lon0, lat0;
    double a1, a2, a3, a4, a5;          // latitude series in dz (metres), low latitudes
    double b1, b2, b3, b4, b5;          // ... in v = rho / rho_A - 1, high latitudes
    bool ok, low, high;
};

// the reference's formulas (thr: z threshold as a fraction of R)
__device__ __forceinline__ void xyz_to_lonlat_reference(double R, double thr, double x, double y, double z, double& lon, double& lat)
{
    const double rho = sqrt(x * x + y * y);
    const double sy = (y > 0.0) ? 1.0 : ((y < 0.0) ? -1.0 : ((y == 0.0) ? 0.0 : y));
    lon = (acos(x / rho) * GTP_RAD2DEG) * sy;
    const double thr_z = thr * R;
    if (z < thr_z && z > -thr_z) lat = 90.0 - acos(z / R) * GTP_RAD2DEG;
    else {
        const double sz = (z > 0.0) ? 1.0 : ((z < 0.0) ? -1.0 : ((z == 0.0) ? 0.0 : z));
        lat = sz * (90.0 - asin(rho / R) * GTP_RAD2DEG);
    }
}

// thr_margin: the latitude series are prepared for |sin lat_A| < thr + 0.01 (low) / > thr - 0.01 (high)
__device__ __forceinline__ AnchorBT anchor_bt_setup(double R, double lon, double lat, double x, double y, double z, double thr = 0.8)
{
    using namespace gtpfast;
    AnchorBT a;
    a.x = x; a.y = y; a.z = z; a.lon0 = lon; a.lat0 = lat;
    const double inv_R = 1.0 / R;
    double ir;
    const double rho = fast_sqrt_rsqrt(fma(x, x, y * y), ir);
    const double cl = x * ir, sl = y * ir;
    a.ue0 = -sl * ir; a.ue1 = cl * ir;
    const double s = z * inv_R, c = rho * inv_R;                      // sin / cos of the latitude of A
    a.ok = (fabs(lon) <= 180.0) && (fabs(lat) <= 90.0) && (c > 1e-6);
    a.low = fabs(s) < thr + 0.01; a.high = fabs(s) > thr - 0.01;
    a.a1 = a.a2 = a.a3 = a.a4 = a.a5 = 0.0;
    a.b1 = a.b2 = a.b3 = a.b4 = a.b5 = 0.0;
    const double s2 = s * s, c2 = c * c;
    if (a.low) {                        // d^k/dz^k of asin(z / R) at z_A, 1 / R^k folded in
        const double inv_c = R * ir, ur = inv_c * inv_c * inv_R;
        double m = inv_c * (kRad2Deg * inv_R);
        a.a1 = m;
        m *= ur; a.a2 = 0.5 * s * m;
        m *= ur; a.a3 = fma(2.0, s2, 1.0) * kSixth * m;
        m *= ur; a.a4 = s * fma(2.0, s2, 3.0) * 0.125 * m;
        m *= ur; a.a5 = fma(s2, fma(8.0, s2, 24.0), 3.0) * k1_40 * m;
    }
    if (a.high) {                       // sign(z) acos(c (1 + v)) around v = 0; sin|lat_A| = |s|
        const double sa = fabs(s), inv_s = fast_rcp(sa), u = inv_s * inv_s;
        const double sign = (z < 0.0) ? -1.0 : 1.0;
        double m = inv_s * c * (-sign * kRad2Deg);
        a.b1 = m;
        m *= u * c; a.b2 = 0.5 * c * m;
        m *= u * c; a.b3 = fma(2.0, c2, 1.0) * kSixth * m;
        m *= u * c; a.b4 = c * fma(2.0, c2, 3.0) * 0.125 * m;
        m *= u * c; a.b5 = fma(c2, fma(8.0, c2, 24.0), 3.0) * k1_40 * m;
    }
    return a;
}

// An anchor per tie point, prepared once (by the kernel that converts the tie points) instead of once per pixel:
// 8 doubles - the horizontal frame, ONE latitude series (the low-latitude one if |z_A| < thr R, else the
// high-latitude one) and which.  A pixel on the other side of the reference's |z| = thr R switch than its anchor
// takes the reference's formulas (a line of pixels along |lat| = 53.13 deg).
constexpr int kAnchorPacked = 8;

// element m of tie point t lives at base[m * n + t] (n tie points): lanes on neighbouring tie points read
// neighbouring addresses - as 64-byte records the loads of a warp touched 32 sectors per instruction
__device__ __forceinline__ void anchor_bt_pack(const AnchorBT& a, double R, double thr, double* __restrict__ base, long long n, long long t)
{
    const bool low = fabs(a.z) < thr * R;
    double* const d = base + t;
    d[0] = a.ue0; d[n] = a.ue1;
    d[2 * n] = low ? a.a1 : a.b1; d[3 * n] = low ? a.a2 : a.b2; d[4 * n] = low ? a.a3 : a.b3;
    d[5 * n] = low ? a.a4 : a.b4; d[6 * n] = low ? a.a5 : a.b5;
    d[7 * n] = !a.ok ? 2.0 : (low ? 0.0 : 1.0);
}

__device__ __forceinline__ AnchorBT anchor_bt_unpack(const double* __restrict__ base, long long n, long long t, double lon, double lat,
                                                     double x, double y, double z)
{
    const double* const p = base + t;
    AnchorBT a;
    a.x = x; a.y = y; a.z = z; a.lon0 = lon; a.lat0 = lat;
    a.ue0 = __ldg(p); a.ue1 = __ldg(p + n);
    a.a1 = a.b1 = __ldg(p + 2 * n); a.a2 = a.b2 = __ldg(p + 3 * n); a.a3 = a.b3 = __ldg(p + 4 * n);
    a.a4 = a.b4 = __ldg(p + 5 * n); a.a5 = a.b5 = __ldg(p + 6 * n);
    const double kind = __ldg(p + 7 * n);
    a.ok = kind < 1.5; a.low = kind == 0.0; a.high = kind == 1.0;
    return a;
}

__device__ __forceinline__ void anchor_bt_eval(const AnchorBT& a, double R, double thr, double x, double y, double z,
                                               double& lon, double& lat)
{
    using namespace gtpfast;
    const double dx = x - a.x, dy = y - a.y, dz = z - a.z;
    const double e = fma(a.ue0, dx, a.ue1 * dy);
    const double h = fma(a.ue1, dx, -a.ue0 * dy);
    const double thr_z = thr * R;
    const bool use_low = (z < thr_z) && (z > -thr_z);                 // the reference's own test
    const bool fits = a.ok && (fabs(e) <= 0.01) && (fabs(h) <= 0.01) && (fabs(dz) <= 2e-3 * R) && (use_low ? a.low : a.high);
    if (!fits) { xyz_to_lonlat_reference(R, thr, x, y, z, lon, lat); return; }
    const double g = 1.0 + h, h2 = h * h;
    double rcp = 1.0 - h;                                             // 1 / (1 + h) = (1 - h)(1 + h^2)(1 + h^4) + O(h^8)
    rcp = fma(h2, rcp, rcp);
    rcp = fma(h2 * h2, rcp, rcp);
    const double t = e * rcp, t2 = t * t;
    const double at = fma(t * t2, fma(t2, fma(t2, -1.0 / 7.0, 0.2), -kThird), t);
    double lo = fma(at, kRad2Deg, a.lon0);
    if (lo > 180.0) lo -= 360.0;
    else if (lo < -180.0) lo += 360.0;
    lon = lo;
    if (use_low) lat = fma(dz, fma(dz, fma(dz, fma(dz, fma(dz, a.a5, a.a4), a.a3), a.a2), a.a1), a.lat0);
    else {
        const double v = fma(g, t2 * fma(t2, fma(t2, 0.0625, -0.125), 0.5), h);       // rho / rho_A - 1
        lat = fma(v, fma(v, fma(v, fma(v, fma(v, a.b5, a.b4), a.b3), a.b2), a.b1), a.lat0);
    }
}
