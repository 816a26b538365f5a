// gtp_fast_math.cuh - anchor-relative float64 math of the CVIIRS 1 km fast path.
//
// Same algorithm as the reference (/root/reference/geotiepoints/_modis_interpolator.pyx
// :41-82 expansion/alignment, :328-344 a_track/a_scan, :379-398 bilinear blend of the
// Cartesian tie points; _modis_utils.pyx:26-65 transforms), re-formulated so that the
// per-pixel work is ~25 FP64 FMAs instead of two inverse-trig calls, a sqrt and two
// divisions (SURVEY.md 7 "hard part 1"):
//
//   * every fine pixel of a tie-point quad lies within a few km of the quad's corner A,
//     whose lon/lat are INPUTS.  With the quad written as  v = A + a P + t Q + a t S
//     (P = B-A, Q = D-A, S = C-D-B+A;  a = a_scan, t = a_track), the displacement
//     d = v - A is projected once per quad on the local frame at A
//     (east e, horizontal-radial h, polar axis z) and pre-scaled by 1/rho_A and 1/R;
//   * lon = lon_A + atan(e' / (1 + h'))        - one Newton-refined reciprocal and a
//                                                3-term series (|e'|,|h'| <= 0.01);
//   * lat = lat_A + Taylor series of asin((z_A + d_z)/R) around z_A/R  (|z| < 0.8 R,
//     the reference's low-latitude branch, _modis_utils.pyx:62-63), or of
//     sign(z) acos(rho/R) around rho_A/R (high-latitude branch, :65), 5 terms.
//
// The reformulation is exact up to the truncation of the series; a quad is accepted
// (`Quad::mode`) only when its size guarantees a truncation error below 2e-13 deg, else
// the caller falls back to the rounding-faithful per-pixel evaluation.  The functions
// are __host__ __device__ so that tests/host_emulation.cpp can run exactly this math
// on the CPU against the oracle.
#pragma once
#include <math.h>
#include <string.h>
#include "gtp_common.cuh"

// Under nvcc these functions are device-only so that they may read __constant__ data: a
// double literal costs two UMOV issue slots per use (measured: 126 UMOV per tie row), a
// constant-bank operand costs none.  g++ (tests/host_emulation.cpp) sees plain inline
// functions and static const doubles.
#ifdef __CUDACC__
#define GTP_FD __device__ __forceinline__
#define GTP_FD_COLD static __device__ __noinline__      // rarely taken paths: keep them out of the hot I-cache lines
#define GTP_DCONST(name, value) static __constant__ double name = value
#else
#define GTP_FD inline
#define GTP_FD_COLD inline
#define GTP_DCONST(name, value) static const double name = value
#endif

namespace gtpfast {

GTP_DCONST(kDeg2Rad, GTP_DEG2RAD);
GTP_DCONST(kRad2Deg, GTP_RAD2DEG);
GTP_DCONST(kEarthR, GTP_EARTH_RADIUS);
GTP_DCONST(kInvEarthR, 1.0 / GTP_EARTH_RADIUS);
GTP_DCONST(kSinPhiScale, GTP_RK / (GTP_RK + GTP_H));           // R / (R + H)
GTP_DCONST(kK, (GTP_RK + GTP_H) / GTP_RK);                     // (R + H) / R
GTP_DCONST(kInv2H, 1.0 / (2.0 * GTP_H));
GTP_DCONST(kTwoOverPi, 0.63661977236758138243);
GTP_DCONST(kPio2Hi, 1.57079632679489655800e+00);
GTP_DCONST(kPio2Lo, 6.12323399573676603587e-17);
GTP_DCONST(kS1, -1.66666666666666324348e-01);
GTP_DCONST(kS2, 8.33333333332248946124e-03);
GTP_DCONST(kS3, -1.98412698298579493134e-04);
GTP_DCONST(kS4, 2.75573137070700676789e-06);
GTP_DCONST(kS5, -2.50507602534068634195e-08);
GTP_DCONST(kS6, 1.58969099521155010221e-10);
GTP_DCONST(kC1, 4.16666666666666019037e-02);
GTP_DCONST(kC2, -1.38888888888741095749e-03);
GTP_DCONST(kC3, 2.48015872894767294178e-05);
GTP_DCONST(kC4, -2.75573143513906633035e-07);
GTP_DCONST(kC5, 2.08757232129817482790e-09);
GTP_DCONST(kC6, -1.13596475577881948265e-11);
GTP_DCONST(kThird, 1.0 / 3.0);
GTP_DCONST(kSixth, 1.0 / 6.0);
GTP_DCONST(k1_24, 1.0 / 24.0);
GTP_DCONST(k1_40, 1.0 / 40.0);
GTP_DCONST(k1_120, 1.0 / 120.0);
GTP_DCONST(k1_720, 1.0 / 720.0);
GTP_DCONST(k1_5040, 1.0 / 5040.0);
GTP_DCONST(k3_40, 0.075);
GTP_DCONST(kABoundExtra, 2.3);     // |a_scan| bound of the pixels right of the last quad
GTP_DCONST(kABound, 2.3);          // float32 path (gtp_fast_f32_math.cuh): one bound for every lane
GTP_DCONST(kATBound, 3.3);
GTP_DCONST(kTBound, 1.4);          // |a_track| bound (first / last rows of a scan extrapolate)
GTP_DCONST(kIncMax, 2.0e-3);       // largest angle step (rad) the row-to-row rotation takes
GTP_DCONST(kDzMin, 1.0e-9);
GTP_DCONST(kLonC3, GTP_DEG2RAD * GTP_DEG2RAD / 3.0);                           // atan series in degrees
GTP_DCONST(kLonC5, GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD / 5.0);
GTP_DCONST(kV1, 0.5 * GTP_DEG2RAD * GTP_DEG2RAD);                              // sqrt(1 + t^2) - 1 series, t in degrees
GTP_DCONST(kV2, -0.125 * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD);
GTP_DCONST(kV3, 0.0625 * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD);
GTP_DCONST(kMaxRelHoriz, 0.01);
GTP_DCONST(kMaxRelHorizWide, 0.1);   // polar quads (cos lat <= kMaxCosWide) up to here take the long series
GTP_DCONST(kMaxCosWide, 0.3);
// atan(t) = t (1 - t^2/3 + t^4/5 - ... + t^12/13) with t in degrees (|t| <= 0.112 rad: next term 3e-16 rad)
GTP_DCONST(kLonC7, -GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD / 7.0);
GTP_DCONST(kLonC9, GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD / 9.0);
GTP_DCONST(kLonC11, -GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD / 11.0);
GTP_DCONST(kLonC13, GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD / 13.0);
// sqrt(1 + x) - 1 = x/2 - x^2/8 + x^3/16 - 5x^4/128 + 7x^5/256 - 21x^6/1024, x = (t deg2rad)^2
GTP_DCONST(kV4, -5.0 / 128.0 * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD);
GTP_DCONST(kV5, 7.0 / 256.0 * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD);
GTP_DCONST(kV6, -21.0 / 1024.0 * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD * GTP_DEG2RAD);
GTP_DCONST(kMaxRelZ, 2.0e-3);
GTP_DCONST(kSmallHoriz, 2.5e-3);
GTP_DCONST(kLatSwitch, 0.8);

enum QuadMode { kSlow = 0, kLow = 1, kHigh = 2, kBoth = 3, kLowSmall = 5, kWide = 8 };   // kLowSmall: kLow with Quad::small
// kWide: polar quad whose displacement is up to 10 % of its distance from the axis (8-term high-latitude series)

// The reference's latitude switch: z < thr*R and z > -thr*R with thr = 0.8
#define GTP_LAT_SWITCH 0.8
// Acceptance limits of the series (see header), in the constant table above:
//   kMaxRelHoriz  bound on |d_horizontal| / rho_A;  kMaxRelZ  bound on |d_z| / R;
//   a_max (1 | kABoundExtra), kTBound  bounds on |a_scan|, |a_track| used to bound |d|

struct TieGeo {            // geometry of one tie point (lonlat2xyz and the local frame)
    double sl, cl;         // sin / cos of the longitude
    double sp, cp;         // sin / cos of the latitude
    double lon, lat;       // the inputs, degrees
};

struct Vec3 { double x, y, z; };

struct TieZen {            // satellite-zenith chain of one tie point (:52-57, :73-78),
    double za;             // kept as sines/cosines instead of angles: zeta = deg2rad(satz)
    double sz, cz;         // sin / cos zeta
    double sphi, cphi;     // sin / cos of phi = asin(R sin(zeta) / (R + H))
};

struct Quad {
    int mode;
    double ce, ca;                       // c_expansion, c_alignment
    // frame components of P, Q, S scaled by 1/rho_A (h), 1/R (w) and 180/(pi rho_A) (e: degrees)
    double Pe, Ph, Pw, Qe, Qh, Qw, Se, Sh, Sw;
    double lon0, lat0;                   // anchor, degrees
    double s0;                           // z_A / R
    double a1, a2, a3, a4, a5;           // low-latitude series (degrees per w^k)
    double a6, a7, a8;                   // 5 km quads only (quad_classify_long)
    double b1, b2, b3, b4, b5;           // high-latitude series (degrees per v^k), v = (rho-rho_A)/rho_A
    double b6, b7, b8;                   // kWide only
    bool wrap;                           // lon may cross +-180
    bool small;                          // low-latitude quad small enough for the shortened series
};

// ---- lean elementary functions ------------------------------------------------------------
// The tie-point / quad set-up runs once per 16 output pixels but CUDA's general-purpose
// sincos / asin / sqrt / division carry argument-reduction slow paths, special-case fix-ups
// and dozens of 64-bit immediates; measured (profiles/r01_fast_v2_ncu.txt) they made the
// set-up two thirds of all issued instructions.  The inputs here are bounded angles and the
// contract is 1e-10 deg (1.7e-12 relative on the Cartesian tie points), so: Cody-Waite
// reduction by pi/2 with a two-part constant, the fdlibm minimax kernels on [-pi/4, pi/4]
// (error < 2^-57), and MUFU-seeded Newton iterations for 1/x and sqrt(x).

// 1/x to ~1 ulp (no special cases: 0 and inf give NaN, which routes the quad to the slow path)
GTP_FD double fast_rcp(double x)
{
#ifdef __CUDA_ARCH__
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
#else
    return 1.0 / x;
#endif
}

// sqrt(x) and 1/sqrt(x) to ~1 ulp for x > 0 (0 gives NaN -> slow path)
GTP_FD double fast_sqrt_rsqrt(double x, double& rsq)
{
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    rsq = y;
    const double sq = x * y;
    return fma(fma(-sq, sq, x), 0.5 * y, sq);
#else
    const double sq = sqrt(x);
    rsq = 1.0 / sq;
    return sq;
#endif
}

GTP_FD double fast_sqrt(double x) { double y; return fast_sqrt_rsqrt(x, y); }

// x with its sign bit flipped where `mask` has bit 63 set (an integer op instead of an FP64 one)
GTP_FD double xor_sign(double x, unsigned long long mask)
{
#ifdef __CUDA_ARCH__
    return __longlong_as_double(__double_as_longlong(x) ^ (long long)mask);
#else
    unsigned long long u;
    memcpy(&u, &x, 8); u ^= mask; memcpy(&x, &u, 8);
    return x;
#endif
}

struct SinCos { double s, c; };
GTP_FD_COLD SinCos library_sincos(double x) { SinCos r; sincos(x, &r.s, &r.c); return r; }

// sin and cos of |x| <= 8 (longitudes, latitudes, zenith angles in radians); the caller checks the range
GTP_FD void sincos_bounded(double x, double& sn, double& cs)
{
    const double kf = rint(x * kTwoOverPi);
    double r = fma(-kf, kPio2Hi, x);
    r = fma(-kf, kPio2Lo, r);
    const int quad = (int)kf;
    const double z = r * r;
    double ps = fma(z, kS6, kS5);
    ps = fma(z, ps, kS4);
    ps = fma(z, ps, kS3);
    ps = fma(z, ps, kS2);
    ps = fma(z, ps, kS1);
    const double sr = fma(r * z, ps, r);
    double pc = fma(z, kC6, kC5);
    pc = fma(z, pc, kC4);
    pc = fma(z, pc, kC3);
    pc = fma(z, pc, kC2);
    pc = fma(z, pc, kC1);
    const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
    const double s0 = (quad & 1) ? cr : sr;
    const double c0 = (quad & 1) ? sr : cr;
    sn = xor_sign(s0, (unsigned long long)(quad & 2) << 62);
    cs = xor_sign(c0, (unsigned long long)((quad + 1) & 2) << 62);
}

GTP_FD void fast_sincos(double x, double& sn, double& cs)
{
    if (!(fabs(x) <= 8.0)) { const SinCos r = library_sincos(x); sn = r.s; cs = r.c; return; }   // out-of-range / NaN: library path
    sincos_bounded(x, sn, cs);
}

// (sin, cos)(x + d) from (sin, cos)(x) for |d| <= kIncMax: sin d = d - d^3/6 (next term 2.7e-16),
// cos d - 1 = -d^2/2 + d^4/24.  Nine such steps (one scan) add < 2e-15 of rounding.
GTP_FD void rotate_small(double& sn, double& cs, double d)
{
    const double d2 = d * d;
    const double sd = fma(d * d2, -kSixth, d);
    const double cm = d2 * fma(d2, k1_24, -0.5);
    const double s1 = fma(sn, cm, fma(cs, sd, sn));
    const double c1 = fma(cs, cm, fma(-sn, sd, cs));
    sn = s1; cs = c1;
}

GTP_FD TieGeo tie_geo(double lon, double lat)
{
    TieGeo t;
    t.lon = lon; t.lat = lat;
    fast_sincos(lon * kDeg2Rad, t.sl, t.cl);
    fast_sincos(lat * kDeg2Rad, t.sp, t.cp);
    return t;
}

GTP_FD Vec3 tie_xyz(const TieGeo& t)             // _modis_utils.pyx:38-40
{
    Vec3 v;
    const double rc = kEarthR * t.cp;
    v.x = rc * t.cl; v.y = rc * t.sl; v.z = kEarthR * t.sp;
    return v;
}

GTP_FD void tie_zen_phi(TieZen& t)
{
    t.sphi = t.sz * kSinPhiScale;                                  // :74
    t.cphi = fast_sqrt(fma(-t.sphi, t.sphi, 1.0));
}

GTP_FD TieZen tie_zen(double satz)
{
    TieZen t;
    t.za = satz * kDeg2Rad;
    fast_sincos(t.za, t.sz, t.cz);
    tie_zen_phi(t);
    return t;
}

struct TieFull { TieGeo g; TieZen z; };
GTP_FD_COLD TieFull tie_full_library(double lon, double lat, double satz)
{
    TieFull f;
    f.g = tie_geo(lon, lat);
    f.z = tie_zen(satz);
    return f;
}

// field by field: an aggregate copy out of the cold call's return slot would pin g and z in local
// memory on the hot path too (measured: 4x the local loads in the SIMPLE instantiation)
template <bool WITH_ZEN>
GTP_FD void take_tie(TieGeo& g, TieZen& z, const TieFull& f)
{
    g.sl = f.g.sl; g.cl = f.g.cl; g.sp = f.g.sp; g.cp = f.g.cp; g.lon = f.g.lon; g.lat = f.g.lat;
    if (WITH_ZEN) { z.za = f.z.za; z.sz = f.z.sz; z.cz = f.z.cz; z.sphi = f.z.sphi; z.cphi = f.z.cphi; }
}

// First tie row of a scan: full evaluation (library functions for out-of-range / NaN inputs).
template <bool WITH_ZEN>
GTP_FD void tie_first(TieGeo& g, TieZen& z, double lon, double lat, double satz)
{
    const double xl = lon * kDeg2Rad, xp = lat * kDeg2Rad, za = satz * kDeg2Rad;
    if ((fabs(xl) <= 8.0) && (fabs(xp) <= 8.0) && (fabs(za) <= 8.0)) {
        g.lon = lon; g.lat = lat;
        sincos_bounded(xl, g.sl, g.cl);
        sincos_bounded(xp, g.sp, g.cp);
        if (WITH_ZEN) { z.za = za; sincos_bounded(za, z.sz, z.cz); tie_zen_phi(z); }
    } else {
        const TieFull f = tie_full_library(lon, lat, WITH_ZEN ? satz : 0.0);
        take_tie<WITH_ZEN>(g, z, f);
    }
}

template <bool WITH_ZEN>
GTP_FD_COLD TieFull tie_first_cold(double lon, double lat, double satz)
{
    TieFull f;
    tie_first<WITH_ZEN>(f.g, f.z, lon, lat, satz);
    return f;
}

// The tie point below `g` / `z` in the same tie column (the kernel walks down the ten tie rows of
// a scan): consecutive rows are ~1 km apart, so the sines and cosines follow from the previous
// row's by a small rotation (9 FMAs per angle instead of a 23-instruction sincos + quadrant
// logic).  Steps beyond kIncMax (poles, the antimeridian), NaN and out-of-range inputs take the
// full evaluation, out of line.
template <bool WITH_ZEN>
GTP_FD void tie_next(TieGeo& g, TieZen& z, double lon, double lat, double satz)
{
    const double dl = (lon - g.lon) * kDeg2Rad, dp = (lat - g.lat) * kDeg2Rad;
    const double za = satz * kDeg2Rad, dzr = WITH_ZEN ? za - z.za : 0.0;
    if ((fabs(dl) <= kIncMax) && (fabs(dp) <= kIncMax) && (fabs(dzr) <= kIncMax)) {
        g.lon = lon; g.lat = lat;
        rotate_small(g.sl, g.cl, dl);
        rotate_small(g.sp, g.cp, dp);
        if (WITH_ZEN) { z.za = za; rotate_small(z.sz, z.cz, dzr); tie_zen_phi(z); }
    } else {
        const TieFull f = tie_first_cold<WITH_ZEN>(lon, lat, satz);
        take_tie<WITH_ZEN>(g, z, f);
    }
}

template <bool WITH_ZEN>
GTP_FD void tie_advance(TieGeo& g, TieZen& z, double lon, double lat, double satz, bool first)
{
    if (first) tie_first<WITH_ZEN>(g, z, lon, lat, satz);
    else tie_next<WITH_ZEN>(g, z, lon, lat, satz);
}

// Expansion / alignment of the quad whose upper corners are a (left) and b (right),
// _modis_interpolator.pyx:41-70, without inverse trigonometry.  With zm = (za + zb)/2:
//   phi_a, phi_b are known through their sines/cosines; dphi = phi_b - phi_a and the mean
//   phi follow from angle-difference identities with small-argument series;
//   (theta_a + theta_b)/2 - theta  =  zm - zeta   (phi is the exact mean of phi_a, phi_b), and
//   zeta - zm = asin(sin(zeta) cos(zm) - cos(zeta) sin(zm)) is a tiny angle;
//   theta_a - theta_b = dphi - dz;
//   e = cos(zeta) - sqrt(cos(zeta)^2 - d^2) = (d^2 / 2 cos(zeta)) (1 + d^2 / 4 cos(zeta)^2)
//   (d <= 7e-4, next term 1e-14 relative; the reference's own form cancels 7 digits there).
// Returns false (ce, ca meaningless) when the two zenith angles are too far apart for the
// small-argument series or so close that the reference's theta_a == theta_b test (:62) may
// fire; the caller then uses quad_exp_ali_exact.  Straight-line on purpose: the check is the
// last thing evaluated so that the compiler can interleave this dependent chain with the
// quad projection that follows it in the kernel.
GTP_FD bool quad_exp_ali(double za, double sza, double cza, double sphi_a, double cphi_a,
                         double zb, double sphi_b, double cphi_b, int km, double& ce, double& ca)
{
    const double K = kK;
    // dphi = asin(sin(phi_b - phi_a))
    const double sd = fma(sphi_b, cphi_a, -(sphi_a * cphi_b));
    const double dz = zb - za;
    const double sd2 = sd * sd;
    const double dphi = fma(sd * sd2, fma(sd2, k3_40, kSixth), sd);
    // mean phi = phi_a + dphi / 2
    const double hp = 0.5 * dphi, hp2 = hp * hp;
    const double sh = fma(hp * hp2, fma(hp2, k1_120, -kSixth), hp);
    const double ch = fma(hp2, fma(hp2, k1_24, -0.5), 1.0);
    const double sphi = fma(sphi_a, ch, cphi_a * sh);
    const double cphi = fma(cphi_a, ch, -(sphi_a * sh));
    const double szeta = K * sphi;                                   // sin(zeta) of :82
    double inv_czeta;
    const double czeta = fast_sqrt_rsqrt(fma(-szeta, szeta, 1.0), inv_czeta);
    // zm = za + dz / 2
    const double hz = 0.5 * dz, hz2 = hz * hz;
    const double shz = fma(hz * hz2, fma(hz2, fma(hz2, -k1_5040, k1_120), -kSixth), hz);
    const double chz = fma(hz2, fma(hz2, fma(hz2, -k1_720, k1_24), -0.5), 1.0);
    const double szm = fma(sza, chz, cza * shz);
    const double czm = fma(cza, chz, -(sza * shz));
    // delta = zeta - zm
    const double sdel = fma(szeta, czm, -(czeta * szm));
    const double delta = fma(sdel * (sdel * sdel), kSixth, sdel);
    const double den = dphi - dz;                                    // theta_a - theta_b
    const double inv_den = 4.0 * fast_rcp(den);
    ce = -delta * inv_den;
    const double d = (K * cphi - czeta) * ((double)km * kInv2H);
    const double u = (d * d) * inv_czeta;
    const double e = (0.5 * u) * fma(0.25 * u, inv_czeta, 1.0);
    ca = (e * szeta) * inv_den;
    return fabs(sd) <= 0.01 && fabs(dz) <= 0.02 && fabs(dz) >= kDzMin;      // false for NaN
}

// the same coefficients, written exactly like the reference (:52-70), library functions
struct ExpAli { double ce, ca; };

GTP_FD_COLD ExpAli quad_exp_ali_exact(double za, double zb, int km)
{
    double ce, ca;
    const double phi_a = asin((GTP_RK * sin(za)) / (GTP_RK + GTP_H));
    const double phi_b = asin((GTP_RK * sin(zb)) / (GTP_RK + GTP_H));
    const double theta_a = za - phi_a, theta_b = zb - phi_b;
    const double phi = (phi_a + phi_b) / 2.0;
    const double zeta = asin(((GTP_RK + GTP_H) * sin(phi)) / GTP_RK);
    const double theta = zeta - phi;
    const double den = (theta_a == theta_b) ? theta_a * 2.0 : theta_a - theta_b;
    ce = 4.0 * ((((theta_a + theta_b) / 2.0) - theta) / den);
    const double d = ((((GTP_RK + GTP_H) / GTP_RK) * cos(phi)) - cos(zeta)) * ((double)km / (2.0 * GTP_H));
    const double e = cos(zeta) - sqrt((cos(zeta) * cos(zeta)) - (d * d));
    ca = ((4.0 * e) * sin(zeta)) / den;
    ExpAli r; r.ce = ce; r.ca = ca;
    return r;
}

GTP_FD void quad_coefficients(const TieZen& a, double zb, double sphi_b, double cphi_b, int km,
                              double& ce, double& ca)
{
    if (!quad_exp_ali(a.za, a.sz, a.cz, a.sphi, a.cphi, zb, sphi_b, cphi_b, km, ce, ca)) {
        const ExpAli r = quad_exp_ali_exact(a.za, zb, km);
        ce = r.ce; ca = r.ca;
    }
}

GTP_FD double absd(double x) { return fabs(x); }

// A, B = upper left / right, D, C = lower left / right corner of the quad; ce, ca from
// quad_exp_ali (they only depend on the zenith angles of A and B).  a_max bounds |a_scan| over
// the pixels this quad is evaluated at: 1 for the P pixels inside the quad (s_s <= 3/4, and
// |ce|, |ca| <= 1/4 is part of the acceptance), kABoundExtra for the P pixels right of the last
// quad (s_s <= 7/4, s_s (1 - s_s) >= -21/16).
struct QuadSize { double horiz, wmax, inv_c; };

// frame projection and size of the quad: straight-line
GTP_FD QuadSize quad_project(Quad& q, const TieGeo& A, const Vec3& Axyz, const Vec3& B, const Vec3& C, const Vec3& D,
                             double a_max)
{
    const double inv_c = fast_rcp(A.cp);
    const double inv_rho = inv_c * kInvEarthR;
    const double inv_rho_deg = inv_rho * kRad2Deg;
    const double inv_R = kInvEarthR;
    const double Px = B.x - Axyz.x, Py = B.y - Axyz.y, Pz = B.z - Axyz.z;
    const double Qx = D.x - Axyz.x, Qy = D.y - Axyz.y, Qz = D.z - Axyz.z;
    const double Sx = (C.x - D.x) - Px, Sy = (C.y - D.y) - Py, Sz = (C.z - D.z) - Pz;
    q.Pe = fma(-A.sl, Px, A.cl * Py) * inv_rho_deg; q.Ph = fma(A.cl, Px, A.sl * Py) * inv_rho; q.Pw = Pz * inv_R;
    q.Qe = fma(-A.sl, Qx, A.cl * Qy) * inv_rho_deg; q.Qh = fma(A.cl, Qx, A.sl * Qy) * inv_rho; q.Qw = Qz * inv_R;
    q.Se = fma(-A.sl, Sx, A.cl * Sy) * inv_rho_deg; q.Sh = fma(A.cl, Sx, A.sl * Sy) * inv_rho; q.Sw = Sz * inv_R;
    q.lon0 = A.lon; q.lat0 = A.lat; q.s0 = A.sp;

    // size of the displacement over the whole quad (incl. the extrapolated edge pixels), radians
    const double at_max = a_max * kTBound;
    const double horiz = a_max * fma(absd(q.Pe), kDeg2Rad, absd(q.Ph)) + kTBound * fma(absd(q.Qe), kDeg2Rad, absd(q.Qh))
                         + at_max * fma(absd(q.Se), kDeg2Rad, absd(q.Sh));
    const double wmax = a_max * absd(q.Pw) + kTBound * absd(q.Qw) + at_max * absd(q.Sw);
    QuadSize sz; sz.horiz = horiz; sz.wmax = wmax; sz.inv_c = inv_c;
    return sz;
}

// acceptance test, evaluation mode and series coefficients
GTP_FD void quad_classify(Quad& q, const TieGeo& A, const QuadSize& sz, double ce, double ca)
{
    const double horiz = sz.horiz, wmax = sz.wmax, inv_c = sz.inv_c;
    q.wrap = false;
    q.small = false;
    q.ce = ce; q.ca = ca;
    // written so that any NaN (lon/lat/satz) fails the test and takes the slow path
    // ... and so that an anchor outside [-180, 180] x [-90, 90] (e.g. the MOD03 fill value -999) does too: the
    // reference folds it through lonlat2xyz -> xyz2lonlat, the anchor-relative series would keep it
    const bool sane = (wmax <= kMaxRelZ) && (absd(q.ce) <= 0.25) && (absd(q.ca) <= 0.25) && (A.cp > 1e-6)
                      && (absd(A.lon) <= 180.0) && (absd(A.lat) <= 90.0);
    const bool narrow = sane && (horiz <= kMaxRelHoriz);
    // next to a pole the quad may span up to 10 % of its distance from the axis: longer series (kWide)
    const bool wide = !narrow && sane && (horiz <= kMaxRelHorizWide) && (A.cp <= kMaxCosWide);
    if (!(narrow || wide)) { q.mode = kSlow; return; }

    const double as0 = absd(q.s0);
    const double margin = 1e-9;
    if (wide) q.mode = kWide;
    else if (as0 + wmax < kLatSwitch - margin) q.mode = kLow;          // z_hi < 0.8 R and z_lo > -0.8 R
    else if (as0 - wmax > kLatSwitch + margin) q.mode = kHigh;         // z_lo > 0.8 R or z_hi < -0.8 R
    else q.mode = kBoth;
    q.wrap = (absd(q.lon0) + 1.05 * horiz * kRad2Deg) >= 180.0;

    if (q.mode & kLow) {
        // d/dw^k of asin(s + w) at w = 0, over k!  (s = sin lat_A, c = cos lat_A)
        const double s = A.sp, s2 = s * s, u = inv_c * inv_c;
        double m = inv_c * kRad2Deg;
        q.a1 = m;
        m *= u; q.a2 = 0.5 * s * m;
        m *= u; q.a3 = fma(2.0, s2, 1.0) * kSixth * m;
        m *= u; q.a4 = s * fma(2.0, s2, 3.0) * 0.125 * m;
        m *= u; q.a5 = fma(s2, fma(8.0, s2, 24.0), 3.0) * k1_40 * m;
        // shortened series (pixel_eval<kLowSmall>): 1/(1+h) to h^4, atan to t^3, asin to w^4.
        // Dropped terms: h^5 t <= 2.4e-16 rad, t^5/5 <= 2e-14 rad = 1.1e-12 deg, a5 w^5 <= 2e-14 deg.
        const double w2 = wmax * wmax;
        q.small = (q.mode == kLow) && (horiz <= kSmallHoriz) && (absd(q.a5) * (w2 * w2 * wmax) <= 2e-14);
    }
    if (q.mode & (kHigh | kWide)) {
        // lat = sign(z) * acos(rho / R);  rho / R = c (1 + v),  v = (rho - rho_A) / rho_A, c = cos(lat_A):
        // lat = lat_A + sum_k b_k v^k with b_k = -sign(z) rad2deg c^k a_k, a_k the Taylor coefficients of
        // asin around c.  They obey  a_{m+2} = (c (m+1)(2m+1) a_{m+1} + m^2 a_m) / ((1 - c^2)(m+2)(m+1)), so
        //   b_{m+2} = (c^2/s^2) ((2m+1)/(m+2) b_{m+1} + m^2/((m+2)(m+1)) b_m),   b_1 = -sign(z) rad2deg c/s.
        // kWide: 8 terms, (c v / s)^9 <= 5e-14 with c <= 0.3, |v| <= 0.1.
        const double c = A.cp, sa = absd(A.sp);
        const double inv_s = fast_rcp(sa);
        const double cs = c * inv_s, kap = cs * cs;
        const double sgn = (A.sp < 0.0) ? kRad2Deg : -kRad2Deg;   // -sign(z) * rad2deg
        q.b1 = cs * sgn;
        q.b2 = kap * (0.5 * q.b1);
        q.b3 = kap * fma(1.0 / 6.0, q.b1, q.b2);
        q.b4 = kap * fma(1.25, q.b3, (1.0 / 3.0) * q.b2);
        q.b5 = kap * fma(1.4, q.b4, 0.45 * q.b3);
        if (wide) {
            q.b6 = kap * fma(1.5, q.b5, (16.0 / 30.0) * q.b4);
            q.b7 = kap * fma(11.0 / 7.0, q.b6, (25.0 / 42.0) * q.b5);
            q.b8 = kap * fma(1.625, q.b7, (36.0 / 56.0) * q.b6);
        }
    }
}

GTP_FD void quad_setup(Quad& q, const TieGeo& A, const Vec3& Axyz, const Vec3& B, const Vec3& C, const Vec3& D,
                       double ce, double ca, double a_max)
{
    const QuadSize sz = quad_project(q, A, Axyz, B, C, D, a_max);
    quad_classify(q, A, sz, ce, ca);
}

struct RowCoef {           // per (quad, fine row): the a_track-dependent half of the bilinear form
    double Ue, Uh, Uw, Ve, Vh, Vw;     // (e, h, w) = a_col U + V, with the row's alignment term folded into V
};

// s_t = a_track of the fine row, c_t = s_t (1 - s_t)
GTP_FD RowCoef row_setup(const Quad& q, double s_t, double c_t)
{
    RowCoef r;
    const double a_base = c_t * q.ca;            // the per-(quad, row) part of a_scan (:344)
    r.Ue = fma(s_t, q.Se, q.Pe); r.Uh = fma(s_t, q.Sh, q.Ph); r.Uw = fma(s_t, q.Sw, q.Pw);
    r.Ve = fma(a_base, r.Ue, s_t * q.Qe); r.Vh = fma(a_base, r.Uh, s_t * q.Qh); r.Vw = fma(a_base, r.Uw, s_t * q.Qw);
    return r;
}

// One fine pixel.  `a_col` = s_s + s_s (1 - s_s) c_exp is the per-(quad, column) part of
// a_scan (:344).  MODE and WRAP are compile-time so that the row loops of the kernel are
// branch-free; they are quad properties (Quad::mode, ::wrap).  e is in degrees (Quad).
template <int MODE, bool WRAP>
GTP_FD void pixel_eval(const Quad& q, const RowCoef& r, double a_col, double& lon, double& lat)
{
    const double e = fma(a_col, r.Ue, r.Ve);
    const double h = fma(a_col, r.Uh, r.Vh);
    const double w = fma(a_col, r.Uw, r.Vw);
    if (MODE == kLowSmall) {
        // t = e / (1 + h) with 1/(1+h) = 1 - h + h^2 - h^3 + h^4 (|h| <= 2.5e-3)
        const double rcp = fma(h, fma(h, fma(h, h - 1.0, 1.0), -1.0), 1.0);
        const double t = e * rcp;
        lon = fma(t, fma(t * t, -kLonC3, 1.0), q.lon0);              // atan(t) = t - t^3/3, in degrees
        lat = fma(w, fma(w, fma(w, fma(w, q.a4, q.a3), q.a2), q.a1), q.lat0);
        return;
    }
    // t = e / (1 + h): 1/(1+h) ~ 1 - h + h^2, one Newton step (|h| <= 0.01 -> rel. error 1e-12)
    const double g = 1.0 + h;
    double rcp = fma(h, h - 1.0, 1.0);
    rcp = fma(rcp, fma(-g, rcp, 1.0), rcp);
    const double t = e * rcp;
    const double t2 = t * t;
    // atan(t) = t - t^3/3 + t^5/5   (|t| <= 0.011 rad -> next term 3e-15 rad), t in degrees
    double lo = fma(t, fma(t2, fma(t2, kLonC5, -kLonC3), 1.0), q.lon0);
    if (WRAP) {
        if (lo > 180.0) lo -= 360.0;
        else if (lo < -180.0) lo += 360.0;
    }
    lon = lo;

    double la_low = 0.0, la_high = 0.0;
    if (MODE & kLow)
        la_low = fma(w, fma(w, fma(w, fma(w, fma(w, q.a5, q.a4), q.a3), q.a2), q.a1), q.lat0);
    if (MODE & kHigh) {
        // rho / rho_A = (1 + h) sqrt(1 + t^2);  v = h + g (t^2/2 - t^4/8 + t^6/16)
        const double v = fma(g, t2 * fma(t2, fma(t2, kV3, kV2), kV1), h);
        la_high = fma(v, fma(v, fma(v, fma(v, fma(v, q.b5, q.b4), q.b3), q.b2), q.b1), q.lat0);
    }
    if (MODE == kLow) lat = la_low;
    else if (MODE == kHigh) lat = la_high;
    else {
        const double zr = q.s0 + w;
        lat = (zr < kLatSwitch && zr > -kLatSwitch) ? la_low : la_high;
    }
}

// ---- 5 km tie points: the same anchor-relative evaluation with 8-term series ---------------------
// A 5 km quad spans 5 x 5 ... 25 x 5 km and is extrapolated to a_track = -0.4 ... 1.4 and, at the
// swath edges, a_scan = -0.4 ... 2.2 (_get_coords_5km, :231-249): displacements of up to 60 km from
// corner A.  The series are therefore taken to the 8th order with coefficients from the Taylor
// recurrence of asin around x0 (1 - x0^2 = q):
//     a_{m+2} = (x0 (2m+1)/(m+2) a_{m+1} + m^2/((m+2)(m+1)) a_m) / q,
// low latitudes: x0 = sin(lat_A), variable w = dz/R; high latitudes: x0 = cos(lat_A), variable c v
// (folded into b_k, see quad_classify).  A quad is accepted when the 9th term is below 1e-13 deg
// over the quad's extent, |t| <= 0.11 (atan to t^13) - otherwise the rounding-faithful formulas.
GTP_DCONST(kLongTerm, 1.0e-13);

GTP_FD void quad_classify_long(Quad& q, const TieGeo& A, const QuadSize& sz, double ce, double ca)
{
    const double horiz = sz.horiz, wmax = sz.wmax, inv_c = sz.inv_c;
    q.wrap = false;
    q.small = false;
    q.ce = ce; q.ca = ca;
    const bool sane = (horiz <= kMaxRelHorizWide) && (wmax <= 0.05)
                      && (absd(q.ce) <= 0.25) && (absd(q.ca) <= 0.25) && (A.cp > 1e-6)
                      && (absd(A.lon) <= 180.0) && (absd(A.lat) <= 90.0);      // fill values: see quad_classify
    if (!sane) { q.mode = kSlow; return; }
    const double as0 = absd(q.s0);
    const double margin = 1e-9;
    if (as0 + wmax < kLatSwitch - margin) q.mode = kLow;
    else if (as0 - wmax > kLatSwitch + margin) q.mode = kHigh;
    else q.mode = kBoth;
    q.wrap = (absd(q.lon0) + 1.05 * horiz * kRad2Deg) >= 180.0;

    bool ok = true;
    if (q.mode & kLow) {
        const double s = A.sp, u = inv_c * inv_c, su = s * u;
        q.a1 = inv_c * kRad2Deg;
        q.a2 = (0.5 * su) * q.a1;
        q.a3 = fma(su, q.a2, ((1.0 / 6.0) * u) * q.a1);
        q.a4 = fma(1.25 * su, q.a3, ((1.0 / 3.0) * u) * q.a2);
        q.a5 = fma(1.4 * su, q.a4, (0.45 * u) * q.a3);
        q.a6 = fma(1.5 * su, q.a5, ((16.0 / 30.0) * u) * q.a4);
        q.a7 = fma((11.0 / 7.0) * su, q.a6, ((25.0 / 42.0) * u) * q.a5);
        q.a8 = fma(1.625 * su, q.a7, ((36.0 / 56.0) * u) * q.a6);
        const double a9 = fma((15.0 / 9.0) * su, q.a8, ((49.0 / 72.0) * u) * q.a7);
        const double w2 = wmax * wmax, w4 = w2 * w2;
        ok = ok && (absd(a9) * (w4 * w4 * wmax) <= kLongTerm);
    }
    if (q.mode & kHigh) {
        const double c = A.cp, sa = absd(A.sp);
        const double inv_s = fast_rcp(sa);
        const double cs = c * inv_s, kap = cs * cs;
        const double sgn = (A.sp < 0.0) ? kRad2Deg : -kRad2Deg;   // -sign(z) * rad2deg
        q.b1 = cs * sgn;
        q.b2 = kap * (0.5 * q.b1);
        q.b3 = kap * fma(1.0 / 6.0, q.b1, q.b2);
        q.b4 = kap * fma(1.25, q.b3, (1.0 / 3.0) * q.b2);
        q.b5 = kap * fma(1.4, q.b4, 0.45 * q.b3);
        q.b6 = kap * fma(1.5, q.b5, (16.0 / 30.0) * q.b4);
        q.b7 = kap * fma(11.0 / 7.0, q.b6, (25.0 / 42.0) * q.b5);
        q.b8 = kap * fma(1.625, q.b7, (36.0 / 56.0) * q.b6);
        const double b9 = kap * fma(15.0 / 9.0, q.b8, (49.0 / 72.0) * q.b7);
        const double vmax = 1.1 * horiz;                      // v = h + g (sqrt(1 + t^2) - 1)
        const double v2 = vmax * vmax, v4 = v2 * v2;
        ok = ok && (absd(b9) * (v4 * v4 * vmax) <= kLongTerm);
    }
    if (!ok) q.mode = kSlow;
    // most 5 km quads are still within the limits the 5-term forms of the 1 km path are good for
    // (pixel_eval<kLow | kHigh>: atan to t^5, a Newton reciprocal, series to the 5th order)
    q.small = (q.mode == kLow || q.mode == kHigh) && !q.wrap && (horiz <= kMaxRelHoriz) && (wmax <= kMaxRelZ);
}

// One fine pixel of a 5 km quad.  MODE = kLow | kHigh | kBoth (kBoth also does the +-180 wrap).
template <int MODE>
GTP_FD void pixel_eval_long(const Quad& q, const RowCoef& r, double a_col, double& lon, double& lat)
{
    const double e = fma(a_col, r.Ue, r.Ve);
    const double h = fma(a_col, r.Uh, r.Vh);
    const double g = 1.0 + h;
    const double t = e * fast_rcp(g);
    const double t2 = t * t;
    double p = fma(t2, kLonC13, kLonC11);
    p = fma(t2, p, kLonC9);
    p = fma(t2, p, kLonC7);
    p = fma(t2, p, kLonC5);
    p = fma(t2, p, -kLonC3);
    double lo = fma(t, fma(t2, p, 1.0), q.lon0);
    if (MODE == kBoth) {
        if (lo > 180.0) lo -= 360.0;
        else if (lo < -180.0) lo += 360.0;
    }
    lon = lo;
    double la_low = 0.0, la_high = 0.0, w = 0.0;
    if (MODE & kLow) {
        w = fma(a_col, r.Uw, r.Vw);
        double a = fma(w, q.a8, q.a7);
        a = fma(w, a, q.a6);
        a = fma(w, a, q.a5);
        a = fma(w, a, q.a4);
        a = fma(w, a, q.a3);
        a = fma(w, a, q.a2);
        a = fma(w, a, q.a1);
        la_low = fma(w, a, q.lat0);
    }
    if (MODE & kHigh) {
        double sq = fma(t2, kV6, kV5);
        sq = fma(t2, sq, kV4);
        sq = fma(t2, sq, kV3);
        sq = fma(t2, sq, kV2);
        sq = fma(t2, sq, kV1);
        const double v = fma(g, t2 * sq, h);
        double b = fma(v, q.b8, q.b7);
        b = fma(v, b, q.b6);
        b = fma(v, b, q.b5);
        b = fma(v, b, q.b4);
        b = fma(v, b, q.b3);
        b = fma(v, b, q.b2);
        b = fma(v, b, q.b1);
        la_high = fma(v, b, q.lat0);
    }
    if (MODE == kLow) lat = la_low;
    else if (MODE == kHigh) lat = la_high;
    else {
        const double zr = q.s0 + w;
        lat = (zr < kLatSwitch && zr > -kLatSwitch) ? la_low : la_high;
    }
}

// sub-pixel coordinate x[i] of _get_coords_5km (a5) for fine column i of a 5 km -> 1 km scan and
// the quad column it belongs to (a7): F5 = 2 partial columns left, 2 (W = 271) or 7 (W = 270) right
GTP_FD int quad_col_5km(int i, int W) { const int c = (i >= 2) ? (i - 2) / 5 : 0; return c > W - 2 ? W - 2 : c; }
GTP_FD int coord_x_5km(int i, int W) { return i - 2 - 5 * quad_col_5km(i, W); }

// One fine pixel of a kWide quad (always wrap-checked: these quads sit next to a pole).
GTP_FD void pixel_eval_wide(const Quad& q, const RowCoef& r, double a_col, double& lon, double& lat)
{
    const double e = fma(a_col, r.Ue, r.Ve);
    const double h = fma(a_col, r.Uh, r.Vh);
    const double g = 1.0 + h;
    const double t = e * fast_rcp(g);
    const double t2 = t * t;
    double p = fma(t2, kLonC13, kLonC11);
    p = fma(t2, p, kLonC9);
    p = fma(t2, p, kLonC7);
    p = fma(t2, p, kLonC5);
    p = fma(t2, p, -kLonC3);
    double lo = fma(t, fma(t2, p, 1.0), q.lon0);
    if (lo > 180.0) lo -= 360.0;
    else if (lo < -180.0) lo += 360.0;
    lon = lo;
    double sq = fma(t2, kV6, kV5);
    sq = fma(t2, sq, kV4);
    sq = fma(t2, sq, kV3);
    sq = fma(t2, sq, kV2);
    sq = fma(t2, sq, kV1);
    const double v = fma(g, t2 * sq, h);
    double b = fma(v, q.b8, q.b7);
    b = fma(v, b, q.b6);
    b = fma(v, b, q.b5);
    b = fma(v, b, q.b4);
    b = fma(v, b, q.b3);
    b = fma(v, b, q.b2);
    b = fma(v, b, q.b1);
    lat = fma(v, b, q.lat0);
}

// ---- rounding-faithful per-pixel evaluation (fallback for quads the series cannot take) ----
GTP_FD int sign_of(double x) { return x > 0.0 ? 1 : (x < 0.0 ? -1 : 0); }

GTP_FD void xyz_to_lonlat(double x, double y, double z, double& lon, double& lat)
{
    const double rho = sqrt(x * x + y * y);
    lon = (acos(x / rho) * GTP_RAD2DEG) * (double)sign_of(y);
    if ((z < (GTP_LAT_SWITCH * GTP_EARTH_RADIUS)) && (z > ((-1.0 * GTP_LAT_SWITCH) * GTP_EARTH_RADIUS)))
        lat = 90.0 - acos(z / GTP_EARTH_RADIUS) * GTP_RAD2DEG;
    else
        lat = (double)sign_of(z) * (90.0 - asin(rho / GTP_EARTH_RADIUS) * GTP_RAD2DEG);
}

GTP_FD double blend(double a_scan, double a_track, double A, double B, double C, double D)
{
    const double scan1 = (1.0 - a_scan) * A + a_scan * B;
    const double scan2 = (1.0 - a_scan) * D + a_scan * C;
    return (1.0 - a_track) * scan1 + a_track * scan2;
}

struct LonLat { double lon, lat; };

// by-value arguments on purpose: a noinline call with reference arguments would force the
// caller's tie points into local memory
GTP_FD_COLD LonLat pixel_eval_slow(Vec3 A, Vec3 B, Vec3 C, Vec3 D, double ce, double ca, double s_s, double s_t)
{
    const double a_scan = (s_s + (s_s * (1.0 - s_s)) * ce) + (s_t * (1.0 - s_t)) * ca;
    const double x = blend(a_scan, s_t, A.x, B.x, C.x, D.x);
    const double y = blend(a_scan, s_t, A.y, B.y, C.y, D.y);
    const double z = blend(a_scan, s_t, A.z, B.z, C.z, D.z);
    LonLat r;
    xyz_to_lonlat(x, y, z, r.lon, r.lat);
    return r;
}

// y[j] of _get_coords_1km (a4): 0.5 + j - h on the first h rows of the scan, continuing
// past f on the last h rows, ((j + h) mod f) + 0.5 in between (f = P, h = P / 2, n = 10 P)
template <int P>
GTP_FD double fine_row_y(int j)
{
    constexpr int kHalf = P / 2, kN = 10 * P;
    return (j < kHalf) ? (0.5 + j - kHalf)
         : (j >= kN - kHalf) ? ((P + 0.5) + (kHalf - (kN - j)))
         : (((j + kHalf) % P) + 0.5);
}

// ---- ECEF output (gtp_fast_xyz.cu): the bilinear form itself, v(a, t) = A + a P + t Q + a t S ----
struct QuadXyz { Vec3 A, P, Q, S; };      // P = B - A, Q = D - A, S = (C - D) - P  (:360-363, :396-398)

GTP_FD QuadXyz quadxyz_setup(const Vec3& A, const Vec3& B, const Vec3& C, const Vec3& D)
{
    QuadXyz q;
    q.A = A;
    q.P.x = B.x - A.x; q.P.y = B.y - A.y; q.P.z = B.z - A.z;
    q.Q.x = D.x - A.x; q.Q.y = D.y - A.y; q.Q.z = D.z - A.z;
    q.S.x = (C.x - D.x) - q.P.x; q.S.y = (C.y - D.y) - q.P.y; q.S.z = (C.z - D.z) - q.P.z;
    return q;
}

// one fine row: v = V + a_col U with a_scan = a_col + a_row, a_row = s_t (1 - s_t) c_ali  (:344)
struct RowXyz { Vec3 U, V; };

GTP_FD RowXyz rowxyz_setup(const QuadXyz& q, double s_t, double a_row)
{
    RowXyz r;
    r.U.x = fma(s_t, q.S.x, q.P.x); r.U.y = fma(s_t, q.S.y, q.P.y); r.U.z = fma(s_t, q.S.z, q.P.z);
    r.V.x = fma(a_row, r.U.x, fma(s_t, q.Q.x, q.A.x));
    r.V.y = fma(a_row, r.U.y, fma(s_t, q.Q.y, q.A.y));
    r.V.z = fma(a_row, r.U.z, fma(s_t, q.Q.z, q.A.z));
    return r;
}

}  // namespace gtpfast
