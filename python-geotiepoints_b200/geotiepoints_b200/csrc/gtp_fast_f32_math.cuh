// gtp_fast_f32_math.cuh - float32 CVIIRS 1 km fast path: rounding-faithful blend, anchor-relative
// back-transform.
//
// The float32 specialisation of the reference stores everything in float but evaluates every
// expression that holds a double literal or a libm call in double and rounds to float at the
// assignment (SURVEY.md 8a "precision model", read off the C Cython generates).  Its fine
// Cartesian coordinates are therefore float32 (0.25-0.5 m granularity) BEFORE the back-transform,
// and a result within 1e-5 deg of the reference has to reproduce those rounding points, not just
// the formulas.  So this path
//   * computes the tie points, the expansion/alignment coefficients, a_scan and the three
//     bilinear blends with exactly the reference's float/double mix (explicit conversions,
//     plain * and + only; the translation unit is compiled with -fmad=false), and
//   * back-transforms the float32 xyz (promoted to double, as the reference does,
//     _modis_utils.pyx:51-57) with the anchor-relative series of gtp_fast_math.cuh around the
//     FLOAT32-ROUNDED corner A of the quad: two inverse-trig calls, a sqrt and two divisions per
//     pixel become ~25 FMAs, accurate to ~1e-13 deg before the final rounding to float32.
// Per pixel that is ~22 float<->double conversions (16/clk/SM on B200; the last rounding of the fine
// xyz is done on the FP64 pipe, round_to_f32) and ~55 FP64 instructions,
// against ~370 FP64-pipe slots of the rounding-faithful kernel in gtp_generic.cu.
#pragma once
#include "gtp_fast_math.cuh"

namespace gtpfast32 {

using namespace gtpfast;

GTP_DCONST(kLatSwitchF32, 0.800000011920928955078125);     // (double)0.8f, _modis_utils.pyx:48

// No-FMA double arithmetic: the reference is built for baseline x86-64 (no contraction).
#ifdef __CUDA_ARCH__
GTP_FD double mul_rn(double a, double b) { return __dmul_rn(a, b); }
GTP_FD double add_rn(double a, double b) { return __dadd_rn(a, b); }
GTP_FD float fmul_rn(float a, float b) { return __fmul_rn(a, b); }
GTP_FD float fsub_rn(float a, float b) { return __fsub_rn(a, b); }
GTP_FD float fadd_rn(float a, float b) { return __fadd_rn(a, b); }
#else
GTP_FD double mul_rn(double a, double b) { volatile double r = a * b; return r; }
GTP_FD double add_rn(double a, double b) { volatile double r = a + b; return r; }
GTP_FD float fmul_rn(float a, float b) { volatile float r = a * b; return r; }
GTP_FD float fsub_rn(float a, float b) { volatile float r = a - b; return r; }
GTP_FD float fadd_rn(float a, float b) { volatile float r = a + b; return r; }
#endif

// x rounded to float32 precision, as a double: what `(double)(float)x` gives, but on the FP64 pipe
// (Veltkamp's splitting with 2^29 + 1: 3 instructions) instead of two F2F conversions on the XU pipe,
// which the kernel is bound by (ncu: XU 80 % busy, F2F = 21 % of the instructions).  Round to nearest;
// a tie (the low 29 bits exactly 1000...0, one value in 5e8) may break either way.
GTP_FD double round_to_f32(double x)
{
#ifdef __CUDA_ARCH__
    const double c = mul_rn(x, 536870913.0);
    return add_rn(c, -add_rn(c, -x));
#else
    return (double)(float)x;
#endif
}

struct TieGeoF {           // geometry of one float32 tie point
    float x, y, z;         // lonlat2xyz, rounded to float (_modis_utils.pyx:38-40)
    double lam, phi;       // (double) of the float32 radians the reference feeds to sin/cos
    double sl, cl, sp, cp; // their sines / cosines (double)
};

struct Vec3F { float x, y, z; };

struct TieZenF { float za, phi, theta; };   // :52-57 rounded to float

GTP_FD TieGeoF tie_geo_f32(float lon, float lat)
{
    TieGeoF t;
    const float lon_rad = (float)mul_rn((double)lon, kDeg2Rad);      // deg2rad, rounded to float (:76-77)
    const float lat_rad = (float)mul_rn((double)lat, kDeg2Rad);
    t.lam = (double)lon_rad; t.phi = (double)lat_rad;
    fast_sincos(t.lam, t.sl, t.cl);
    fast_sincos(t.phi, t.sp, t.cp);
    const double rc = mul_rn(kEarthR, t.cp);
    t.x = (float)mul_rn(rc, t.cl);
    t.y = (float)mul_rn(rc, t.sl);
    t.z = (float)mul_rn(kEarthR, t.sp);
    return t;
}

GTP_FD TieZenF tie_zen_f32(float satz)
{
    TieZenF t;
    t.za = (float)mul_rn((double)satz, kDeg2Rad);
    double s, c;
    fast_sincos((double)t.za, s, c);
    t.phi = (float)asin(mul_rn(GTP_RK, s) / (GTP_RK + GTP_H));       // :74
    t.theta = fsub_rn(t.za, t.phi);                                  // :78
    return t;
}

// _compute_expansion_alignment for one quad, float32 rounding points (:41-70)
GTP_FD void quad_exp_ali_f32(const TieZenF& a, float phi_b, float theta_b, int km, float& c_exp, float& c_ali)
{
    const float phi = (float)((double)fadd_rn(a.phi, phi_b) / 2.0);
    double sphi, cphi;
    fast_sincos((double)phi, sphi, cphi);
    const float zeta = (float)asin(mul_rn(GTP_RK + GTP_H, sphi) / GTP_RK);
    const float theta = fsub_rn(zeta, phi);
    const float den = (a.theta == theta_b) ? (float)mul_rn((double)a.theta, 2.0) : fsub_rn(a.theta, theta_b);
    c_exp = (float)mul_rn(4.0, add_rn((double)fadd_rn(a.theta, theta_b) / 2.0, -(double)theta) / (double)den);
    const float sin_beta_2 = (float)((double)km / (2.0 * GTP_H));
    double szeta, czeta;
    fast_sincos((double)zeta, szeta, czeta);
    const float d = (float)mul_rn(add_rn(mul_rn((GTP_RK + GTP_H) / GTP_RK, cphi), -czeta), (double)sin_beta_2);
    const float dd = fmul_rn(d, d);                                   // powf(d, 2)
    const float e = (float)add_rn(czeta, -sqrt(add_rn(mul_rn(czeta, czeta), -(double)dd)));
    c_ali = (float)(mul_rn(mul_rn(4.0, (double)e), szeta) / (double)den);
}

// ---- the same set-up, walked down a tie column ------------------------------------------------------
// tie_geo_f32 / tie_zen_f32 / quad_exp_ali_f32 cost five sincos, two asin, a sqrt and six divisions per
// tie row and thread (~500 FP64 instructions: as much as the 16 pixels of the quad).  The kernel walks a
// tie column down the ten tie rows of a scan, and consecutive rows (and the neighbouring column) are
// ~1e-4 rad apart, so every one of those functions is evaluated as a small step from a value the thread
// already holds:
//   * sin / cos of the float32 longitude, latitude and zenith angle: rotation of the previous row's by
//     the (exact) difference of the float32 radians;
//   * phi = asin(R sin(za) / (R + H)) (:74): Newton step on sin(phi) = x from the previous row's phi
//     (asin_step: 4th-order start, one rotation, first-order finish; error < 1e-17 rad);
//   * sin / cos of the quad's mean phi (:59): rotation of corner a's; zeta = asin((R + H) sin(phi) / R)
//     (:82): asin_step from corner a's zenith angle, whose sine and cosine the thread has;
//   * divisions by Markstein's two-FMA correction of x * (1 / y).
// The float32 values the reference keeps (x, y, z, phi, theta, zeta, c_exp, c_ali) are roundings of
// doubles that agree with glibc's to a few ulps, i.e. they differ from the reference's in ~1e-8 of the
// tie points (the row-0 evaluation with sincos / asin has the same property).  Steps beyond the limits
// below, NaN and out-of-range inputs restart from the full evaluation.
GTP_DCONST(kStepMax, 4.0e-3);            // largest rotation / asin_step the series below are good for (rad)
GTP_DCONST(kRcpRkH, 1.0 / (GTP_RK + GTP_H));
GTP_DCONST(kRcpRk, 1.0 / GTP_RK);

// (sin, cos)(x + d) from (sin, cos)(x), |d| <= 5e-3: sin d to d^5 (next 1.5e-20), cos d to d^6 (next 1e-23)
GTP_FD void rotate_step(double& sn, double& cs, double d)
{
    const double d2 = d * d;
    const double sd = fma(d * d2, fma(d2, k1_120, -kSixth), d);
    const double cm = d2 * fma(d2, fma(d2, -k1_720, k1_24), -0.5);
    const double s1 = fma(sn, cm, fma(cs, sd, sn));
    const double c1 = fma(cs, cm, fma(-sn, sd, cs));
    sn = s1; cs = c1;
}

// n / c for a constant c with rc = RN(1 / c): q = n rc corrected by its own residual (correctly rounded
// except for a handful of n per binade)
GTP_FD double div_const(double n, double c, double rc)
{
    const double q = n * rc;
    return fma(fma(-q, c, n), rc, q);
}

// The angle with sine x from a nearby angle of sine s0 and cosine c0, |x - s0| <= kStepMax c0: returns the
// step and replaces (s0, c0) by (x, its cosine).  With u = (x - s0) / c0, w = tan:
//   asin(x) - asin(s0) = u + w u^2 / 2 + (1 + 3 w^2) u^3 / 6 + w (9 + 15 w^2) u^4 / 24 + O(u^5)
// then one Newton step on the rotated sine with d(asin) / dx = (1 + w d1) / c0 (relative error d1^2).
GTP_FD double asin_step(double x, double& s0, double& c0)
{
    const double ic = fast_rcp(c0);
    const double u = (x - s0) * ic, w = s0 * ic, w2 = w * w;
    const double t3 = fma(w2, 0.5, kSixth);
    const double t4 = w * fma(w2, 0.625, 0.375);
    const double d1 = fma(u * u, fma(u, fma(u, t4, t3), 0.5 * w), u);
    double s1 = s0, c1 = c0;
    rotate_step(s1, c1, d1);
    const double d2 = (x - s1) * fma(w * d1, ic, ic);
    s0 = x;
    c0 = fma(-s1, d2, c1);
    return d1 + d2;
}

GTP_FD double deg2rad_f32(float v) { return (double)(float)mul_rn((double)v, kDeg2Rad); }     // :76-77

struct TieColF {
    double lam, sl, cl;                 // float32 longitude in radians (as a double), sine, cosine
    double phi, sp, cp;                 // latitude
    double za, sz, cz;                  // zenith angle
    double pt, spt, cpt;                // phi of :74 before its rounding to float32, sine, cosine
};

template <bool WITH_ZEN>
GTP_FD void tie_col_first(TieColF& t, float lon, float lat, float satz)
{
    t.lam = deg2rad_f32(lon); t.phi = deg2rad_f32(lat);
    fast_sincos(t.lam, t.sl, t.cl);
    fast_sincos(t.phi, t.sp, t.cp);
    if (WITH_ZEN) {
        t.za = deg2rad_f32(satz);
        fast_sincos(t.za, t.sz, t.cz);
        t.spt = mul_rn(GTP_RK, t.sz) / (GTP_RK + GTP_H);
        t.pt = asin(t.spt);
        t.cpt = sqrt(fma(-t.spt, t.spt, 1.0));
    }
}

template <bool WITH_ZEN>
GTP_FD_COLD TieColF tie_col_first_cold(float lon, float lat, float satz)
{
    TieColF t;
    t.za = t.sz = t.cz = t.pt = t.spt = t.cpt = 0.0;
    tie_col_first<WITH_ZEN>(t, lon, lat, satz);
    return t;
}

template <bool WITH_ZEN>
GTP_FD void tie_col_take(TieColF& t, const TieColF& f)
{
    t.lam = f.lam; t.sl = f.sl; t.cl = f.cl; t.phi = f.phi; t.sp = f.sp; t.cp = f.cp;
    if (WITH_ZEN) { t.za = f.za; t.sz = f.sz; t.cz = f.cz; t.pt = f.pt; t.spt = f.spt; t.cpt = f.cpt; }
}

// the tie point below t in the same tie column
template <bool WITH_ZEN>
GTP_FD void tie_col_next(TieColF& t, float lon, float lat, float satz)
{
    const double lam = deg2rad_f32(lon), phi = deg2rad_f32(lat), za = WITH_ZEN ? deg2rad_f32(satz) : 0.0;
    const double dl = lam - t.lam, dp = phi - t.phi, dz = WITH_ZEN ? za - t.za : 0.0;
    bool ok = (fabs(dl) <= kStepMax) && (fabs(dp) <= kStepMax) && (fabs(dz) <= kStepMax) && (fabs(lam) <= 8.0) && (fabs(phi) <= 8.0);
    double sz = t.sz, cz = t.cz, x = 0.0;
    if (WITH_ZEN) {
        rotate_step(sz, cz, dz);
        x = div_const(mul_rn(GTP_RK, sz), GTP_RK + GTP_H, kRcpRkH);
        ok = ok && (fabs(x - t.spt) <= kStepMax * t.cpt) && (za >= 0.0) && (za <= 1.3);
    }
    if (ok) {
        t.lam = lam; t.phi = phi;
        rotate_step(t.sl, t.cl, dl);
        rotate_step(t.sp, t.cp, dp);
        if (WITH_ZEN) {
            t.za = za; t.sz = sz; t.cz = cz;
            const double step = asin_step(x, t.spt, t.cpt);
            t.pt += step;
        }
    } else {
        const TieColF f = tie_col_first_cold<WITH_ZEN>(lon, lat, satz);
        tie_col_take<WITH_ZEN>(t, f);
    }
}

// lonlat2xyz rounded to float (_modis_utils.pyx:38-40)
GTP_FD Vec3F tie_col_xyz(const TieColF& t)
{
    Vec3F v;
    const double rc = mul_rn(kEarthR, t.cp);
    v.x = (float)mul_rn(rc, t.cl);
    v.y = (float)mul_rn(rc, t.sl);
    v.z = (float)mul_rn(kEarthR, t.sp);
    return v;
}

GTP_FD TieGeoF tie_col_geo(const TieColF& t)
{
    TieGeoF g;
    const Vec3F v = tie_col_xyz(t);
    g.x = v.x; g.y = v.y; g.z = v.z;
    g.lam = t.lam; g.phi = t.phi; g.sl = t.sl; g.cl = t.cl; g.sp = t.sp; g.cp = t.cp;
    return g;
}

GTP_FD TieZenF tie_col_zen(const TieColF& t)
{
    TieZenF z;
    z.za = (float)t.za;                       // exact: t.za is a float32 value
    z.phi = (float)t.pt;                      // :74
    z.theta = fsub_rn(z.za, z.phi);           // :78
    return z;
}

struct ExpAliF { float ce, ca; };
GTP_FD_COLD ExpAliF quad_exp_ali_f32_cold(float za, float phi_a, float theta_a, float phi_b, float theta_b, int km)
{
    TieZenF a; a.za = za; a.phi = phi_a; a.theta = theta_a;
    ExpAliF r;
    quad_exp_ali_f32(a, phi_b, theta_b, km, r.ce, r.ca);
    return r;
}

// _compute_expansion_alignment (:41-70) for the quad whose upper left corner is the tie point of state t
// (za = (float)t.za etc. in a) and whose upper right corner has phi_b, theta_b
GTP_FD void quad_exp_ali_col(const TieColF& t, const TieZenF& a, float phi_b, float theta_b, int km, float& c_exp, float& c_ali)
{
    const float phi = fmul_rn(0.5f, fadd_rn(a.phi, phi_b));            // (float)((double)(phi_a + phi_b) / 2.0)
    const double h = (double)phi - t.pt;
    double sphi = t.spt, cphi = t.cpt;
    rotate_step(sphi, cphi, h);
    const double x = div_const(mul_rn(GTP_RK + GTP_H, sphi), GTP_RK, kRcpRk);
    double sz = t.sz, cz = t.cz;
    const bool ok = (fabs(h) <= kStepMax) && (fabs(x - sz) <= kStepMax * cz) && (fabs(phi) >= 1e-30f);
    const double dzeta = asin_step(x, sz, cz);
    const float zeta = (float)(t.za + dzeta);
    const double eps = ((double)zeta - t.za) - dzeta;                   // the rounding of zeta
    const double szeta = fma(eps, fma(-0.5 * eps, sz, cz), sz);
    const double czeta = fma(eps, fma(-0.5 * eps, cz, -sz), cz);
    const float theta = fsub_rn(zeta, phi);
    const float den = (a.theta == theta_b) ? fmul_rn(a.theta, 2.0f) : fsub_rn(a.theta, theta_b);
    const double dend = (double)den, rd = fast_rcp(dend);
    const double num = add_rn((double)fadd_rn(a.theta, theta_b) * 0.5, -(double)theta);
    double q = num * rd;
    q = fma(fma(-q, dend, num), rd, q);
    const float ce = (float)mul_rn(4.0, q);
    const float sin_beta_2 = (float)((double)km / (2.0 * GTP_H));
    const float d = (float)mul_rn(add_rn(mul_rn((GTP_RK + GTP_H) / GTP_RK, cphi), -czeta), (double)sin_beta_2);
    const float dd = fmul_rn(d, d);
    const float e = (float)add_rn(czeta, -fast_sqrt(add_rn(mul_rn(czeta, czeta), -(double)dd)));
    const double n2 = mul_rn(mul_rn(4.0, (double)e), szeta);
    double q2 = n2 * rd;
    q2 = fma(fma(-q2, dend, n2), rd, q2);
    const float ca = (float)q2;
    // the step limits, a zero / tiny / non-finite denominator and NaN: the reference's own operations
    if (ok && (fabsf(ce) <= 1e30f) && (fabsf(ca) <= 1e30f) && (fabsf(den) >= 1e-30f)) { c_exp = ce; c_ali = ca; }
    else { const ExpAliF r = quad_exp_ali_f32_cold(a.za, a.phi, a.theta, phi_b, theta_b, km); c_exp = r.ce; c_ali = r.ca; }
}

// Everything the pixels of one quad need.
struct QuadF32 {
    int mode;
    bool wrap;
    double A[3], D[3];                  // corners A and D widened to double (exact)
    float B[3], C[3];                   // corners B and C stay float: a * B is a float product (:396-397)
    float ce, ca;
    double ue0, ue1, uh0, uh1;          // rows of the local frame at A', pre-scaled by 1 / rho_A'
    double lon0;                        // longitude of A', degrees
    double lat0_low, lat0_high;         // latitude of A' by the low / high latitude formula, degrees
    double s0;                          // z_A' / R
    double a1, a2, a3, a4, a5;          // series of asin((z_A' + dz)/R) around z_A'/R in dz (metres), degrees
    double b1, b2, b3, b4, b5;          // series of sign(z) acos(rho/R) around rho_A'/R, degrees
};

// max_horiz / max_w: acceptance limits of the series, |d_horizontal| / rho_A' and |d_z| / R over the quad.
// The defaults keep the truncation error below ~1e-13 deg (1 km quads).  5 km quads (gtp_generic.cu) pass
// kMaxRelHoriz5 / kMaxRelZ5: the float32 result needs ~1e-8 deg, and the first neglected terms are
// t^7 / 7 <= 3e-12 rad of longitude and a6 w^6 <= 1500 deg x 1e-12 of latitude there.
GTP_DCONST(kMaxRelHoriz5, 0.03);
GTP_DCONST(kMaxRelZ5, 1.0e-2);

GTP_FD QuadF32 quad_setup_f32(const TieGeoF& A, const Vec3F& B, const Vec3F& C, const Vec3F& D, float ce, float ca,
                              double max_horiz, double max_w)
{
    QuadF32 q;
    q.ce = ce; q.ca = ca;
    q.wrap = false;
    q.a1 = q.a2 = q.a3 = q.a4 = q.a5 = 0.0;
    q.b1 = q.b2 = q.b3 = q.b4 = q.b5 = 0.0;
    q.lat0_low = q.lat0_high = 0.0;
    q.A[0] = (double)A.x; q.A[1] = (double)A.y; q.A[2] = (double)A.z;
    q.D[0] = (double)D.x; q.D[1] = (double)D.y; q.D[2] = (double)D.z;
    q.B[0] = B.x; q.B[1] = B.y; q.B[2] = B.z;
    q.C[0] = C.x; q.C[1] = C.y; q.C[2] = C.z;

    // local frame at A' = the float32-rounded corner A (what the reference back-transforms)
    double inv_rho;
    const double rho = fast_sqrt_rsqrt(fma(q.A[0], q.A[0], q.A[1] * q.A[1]), inv_rho);
    const double cl = q.A[0] * inv_rho, sl = q.A[1] * inv_rho;
    q.ue0 = -sl * inv_rho; q.ue1 = cl * inv_rho;
    q.uh0 = q.ue1; q.uh1 = -q.ue0;
    // lon(A') = lambda + asin(sin(lon(A') - lambda)): A' is within half a float32 ulp of the exact point
    q.lon0 = (A.lam + fma(A.cl, sl, -(A.sl * cl))) * kRad2Deg;
    q.s0 = q.A[2] * kInvEarthR;

    // displacement bounds over the quad (same role as in gtp_fast_math.cuh), on the FP32 pipe: they only
    // feed comparisons with a few per cent of slack (kBoundSlack)
    const float Px = fsub_rn(B.x, A.x), Py = fsub_rn(B.y, A.y), Pz = fsub_rn(B.z, A.z);
    const float Qx = fsub_rn(D.x, A.x), Qy = fsub_rn(D.y, A.y), Qz = fsub_rn(D.z, A.z);
    const float Sx = fsub_rn(fsub_rn(C.x, D.x), Px), Sy = fsub_rn(fsub_rn(C.y, D.y), Py), Sz = fsub_rn(fsub_rn(C.z, D.z), Pz);
    const float e0 = (float)q.ue0, e1 = (float)q.ue1;
    const float Pe = fmaf(e0, Px, e1 * Py), Ph = fmaf(e1, Px, -e0 * Py);
    const float Qe = fmaf(e0, Qx, e1 * Qy), Qh = fmaf(e1, Qx, -e0 * Qy);
    const float Se = fmaf(e0, Sx, e1 * Sy), Sh = fmaf(e1, Sx, -e0 * Sy);
    // 2 m of slack for the float32 roundings of the fine xyz themselves
    const float kA = (float)kABound, kT = (float)kTBound, kAT = (float)kATBound;
    const double horiz = (double)(1.0001f * (kA * (fabsf(Pe) + fabsf(Ph)) + kT * (fabsf(Qe) + fabsf(Qh))
                                             + kAT * (fabsf(Se) + fabsf(Sh)) + 2.0f * fabsf(e0) + 2.0f * fabsf(e1)));
    const double wmax = (double)(1.0001f * (kA * fabsf(Pz) + kT * fabsf(Qz) + kAT * fabsf(Sz) + 2.0f)) * kInvEarthR;
    const bool ok = (horiz <= max_horiz) && (wmax <= max_w)
                    && (fabsf(ce) <= 1.0f) && (fabsf(ca) <= 0.5f) && (A.cp > 1e-6)
                    // an anchor outside [-180, 180] x [-90, 90] (the MOD03 fill value -999): the reference folds it
                    // through lonlat2xyz -> xyz2lonlat, the anchor-relative series would keep it
                    && (fabs(A.lam) <= 3.1415928) && (fabs(A.phi) <= 1.5707964);
    if (!ok) { q.mode = kSlow; return q; }

    const double zr_lo = q.s0 - wmax, zr_hi = q.s0 + wmax;
    const double margin = 1e-9;
    if (zr_hi < kLatSwitchF32 - margin && zr_lo > -kLatSwitchF32 + margin) q.mode = kLow;
    else if (zr_lo > kLatSwitchF32 + margin || zr_hi < -kLatSwitchF32 - margin) q.mode = kHigh;
    else q.mode = kBoth;
    q.wrap = (absd(q.lon0) + 1.05 * horiz * kRad2Deg) >= 180.0;

    if (q.mode & kLow) {
        // asin around s0 = z_A'/R with c0 = sqrt(1 - s0^2); asin(s0) = phi + asin(s0 cos phi - c0 sin phi)
        const double s = q.s0, s2 = s * s;
        double inv_c;
        const double c0 = fast_sqrt_rsqrt(fma(-s, s, 1.0), inv_c);
        const double u = inv_c * inv_c;
        q.lat0_low = (A.phi + fma(s, A.cp, -(c0 * A.sp))) * kRad2Deg;
        // coefficients per power of dz (metres): 1 / R^k folded in
        const double ur = u * kInvEarthR;
        double m = inv_c * (kRad2Deg * kInvEarthR);
        q.a1 = m;
        m *= ur; q.a2 = 0.5 * s * m;
        m *= ur; q.a3 = fma(2.0, s2, 1.0) * kSixth * m;
        m *= ur; q.a4 = s * fma(2.0, s2, 3.0) * 0.125 * m;
        m *= ur; q.a5 = fma(s2, fma(8.0, s2, 24.0), 3.0) * k1_40 * m;
    }
    if (q.mode & kHigh) {
        // sign(z) acos(rho/R) around c = rho_A'/R with s' = sqrt(1 - c^2):
        // acos(c) = |phi| - asin(|sin phi| c - cos phi s')
        const double c = rho * kInvEarthR, c2 = c * c;
        double inv_s;
        const double sa = fast_sqrt_rsqrt(fma(-c, c, 1.0), inv_s);
        const double u = inv_s * inv_s;
        const double sign = (q.A[2] < 0.0) ? -1.0 : 1.0;
        q.lat0_high = sign * (absd(A.phi) - fma(absd(A.sp), c, -(A.cp * sa))) * kRad2Deg;
        double m = inv_s * c * (-sign * kRad2Deg);
        q.b1 = m;
        m *= u * c; q.b2 = 0.5 * c * m;
        m *= u * c; q.b3 = fma(2.0, c2, 1.0) * kSixth * m;
        m *= u * c; q.b4 = c * fma(2.0, c2, 3.0) * 0.125 * m;
        m *= u * c; q.b5 = fma(c2, fma(8.0, c2, 24.0), 3.0) * k1_40 * m;
    }
    return q;
}

GTP_FD QuadF32 quad_setup_f32(const TieGeoF& A, const Vec3F& B, const Vec3F& C, const Vec3F& D, float ce, float ca)
{
    return quad_setup_f32(A, B, C, D, ce, ca, kMaxRelHoriz, kMaxRelZ);
}

// per (quad, fine row): a_track = s_t (:343) and the s_t-dependent part of a_scan (:344)
struct RowF32 {
    float s_t;
    double omt;          // 1.0 - a_track
    double row_ca;       // (s_t (1 - s_t)) c_ali
};

GTP_FD RowF32 row_setup_f32(const QuadF32& q, float s_t)
{
    RowF32 r;
    r.s_t = s_t;
    r.omt = add_rn(1.0, -(double)s_t);
    r.row_ca = mul_rn(mul_rn((double)s_t, r.omt), (double)q.ca);
    return r;
}

// The float32 fine xyz of one pixel, with the reference's rounding points (:344, :396-398).
// k1d = (double)s_s * (1.0 - (double)s_s)
// v[k] are float32 values held in doubles (the reference promotes them to double next, :51-57)
GTP_FD void blend_f32(const QuadF32& q, const RowF32& r, float s_s, double k1d, double (&v)[3])
{
    const double a_d = add_rn(add_rn((double)s_s, mul_rn(k1d, (double)q.ce)), r.row_ca);
    const float a = (float)a_d;
    const double oma = add_rn(1.0, -(double)a);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float s1 = (float)add_rn(mul_rn(oma, q.A[k]), (double)fmul_rn(a, q.B[k]));
        const float s2 = (float)add_rn(mul_rn(oma, q.D[k]), (double)fmul_rn(a, q.C[k]));
        v[k] = round_to_f32(add_rn(mul_rn(r.omt, (double)s1), (double)fmul_rn(r.s_t, s2)));
    }
}

// ---- the same blend on the FP32 pipe: error-free transforms instead of double arithmetic -----------
// blend_f32 spends 7 float<->double conversions (16 / clk / SM) and 11 FP64 instructions (2 issue slots
// each) per component and pixel.  The reference's value of one edge,
//     s = (float)((1.0 - a) * A + (double)(float)(a * B))          (_modis_interpolator.pyx:396-397)
// is the float32 rounding of E = A + a (B - A) - ep up to the rounding of its double operations
// (<= 2^-29 float32 ulps), with ep = a B - RN32(a B), the error of the float product, itself a float
// (one FMA).  With beta = B - A exact in float32 and |a beta| <= |A| / 2:
//     sigma = RN32(A + a beta)                   1 FMA
//     eps   = RN32((A - sigma) + a beta)         1 ADD (exact, Sterbenz) + 1 FMA: the residual of sigma to 2^-24
//     tau   = eps - ep                           1 MUL + 1 FMA (ep) + 1 ADD; |tau| <~ 1 ulp(sigma)
//     s     = RN32(sigma + tau)
// tau carries a relative error <= 2^-22, so s can be wrong only where sigma + tau lies within 2^-22 |tau| of
// a rounding boundary: s is evaluated with tau (1 +- 2^-20) and the pixel is FLAGGED when the two differ
// (one pixel in ~1e5); the caller re-evaluates flagged rows with blend_f32.  Everything is written on
// pairs of float lanes (fma.rn.f32x2 & co. on sm_100: one issue slot for two lanes).  The second blend,
//     v = (float)((1.0 - s_t) * s1 + (double)(float)(s_t * s2))    (:398)
// has a 4-bit weight, so (1.0 - s_t) s1 and the sum are exact in double and one float FMA rounds identically.
struct F2 { float x, y; };

#ifdef __CUDA_ARCH__
typedef unsigned long long P2;          // two float lanes in one 64-bit register
GTP_FD P2 p2_make(float lo, float hi) { P2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
GTP_FD F2 p2_get(P2 v) { F2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
GTP_FD P2 p2_fma(P2 a, P2 b, P2 c) { P2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
GTP_FD P2 p2_mul(P2 a, P2 b) { P2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
GTP_FD P2 p2_add(P2 a, P2 b) { P2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
#else
typedef F2 P2;
GTP_FD P2 p2_make(float lo, float hi) { P2 r; r.x = lo; r.y = hi; return r; }
GTP_FD F2 p2_get(P2 v) { return v; }
GTP_FD P2 p2_fma(P2 a, P2 b, P2 c) { P2 d; d.x = fmaf(a.x, b.x, c.x); d.y = fmaf(a.y, b.y, c.y); return d; }
GTP_FD P2 p2_mul(P2 a, P2 b) { P2 d; d.x = fmul_rn(a.x, b.x); d.y = fmul_rn(a.y, b.y); return d; }
GTP_FD P2 p2_add(P2 a, P2 b) { P2 d; d.x = fadd_rn(a.x, b.x); d.y = fadd_rn(a.y, b.y); return d; }
#endif
GTP_FD P2 p2_dup(float v) { return p2_make(v, v); }

// One edge of a quad for a pair of lanes: the anchor corner, the other corner and their difference.
struct EdgeP2 { P2 A, B, d; };

#define GTP_EFT_UP 1.00000095367431640625f      // 1 + 2^-20
#define GTP_EFT_DN 0.999999046325683593750f     // 1 - 2^-20

// s of the header comment for two lanes; acc collects (s+ - s-)^2: non-zero (or NaN) = flagged
GTP_FD P2 eft_edge(P2 a, P2 na, const EdgeP2& e, P2& acc)
{
    const P2 m1 = p2_dup(-1.0f);
    const P2 sigma = p2_fma(a, e.d, e.A);
    const P2 d0 = p2_fma(sigma, m1, e.A);
    const P2 eps = p2_fma(a, e.d, d0);
    const P2 p = p2_mul(a, e.B);
    const P2 nep = p2_fma(na, e.B, p);
    const P2 tau = p2_add(eps, nep);
    const P2 sp = p2_fma(tau, p2_dup(GTP_EFT_UP), sigma);
    const P2 sm = p2_fma(tau, p2_dup(GTP_EFT_DN), sigma);
    const P2 df = p2_fma(sm, m1, sp);
    acc = p2_fma(df, df, acc);
    return sp;
}

// v - A_top for two lanes: the second blend (:398) and the displacement from the anchor (exact)
GTP_FD P2 eft_track(P2 s1, P2 s2, P2 st, P2 omt, P2 A)
{
    const P2 v = p2_fma(omt, s1, p2_mul(st, s2));
    return p2_fma(A, p2_dup(-1.0f), v);
}

// What the FP32-pipe blend of one quad needs.  ok: in every component the corners A and D are at least
// `margin` times larger than the quad's extent ext = max(|B - A|, |C - D|, |D - A|) (or the component is 0
// at all four corners).  Then B - A and C - D are exact (Sterbenz), |a beta| <= |A| / 2 for |a| <= margin / 2
// - 0.7, and |v - A| <= (|1 - t| a + |t| (1 + a)) ext < |A| / 2, so A - sigma and v - A are exact too:
// margin 8 for |a| <= 1.3, |t| <= 1.375 (the quad's own pixels), 16 for the lane that extrapolates right of
// the last quad (|a| <= 3.3).  Quads that fail (a component changes sign within a few quad widths: ~1 % of
// a day) take blend_f32.
struct QuadEft {
    float A[3], B[3], C[3], D[3];
    float dAB[3], dDC[3];               // B - A, C - D
    bool ok;
};

GTP_FD QuadEft quad_eft_setup(const Vec3F& A, const Vec3F& B, const Vec3F& C, const Vec3F& D, float margin)
{
    QuadEft q;
    q.A[0] = A.x; q.A[1] = A.y; q.A[2] = A.z; q.B[0] = B.x; q.B[1] = B.y; q.B[2] = B.z;
    q.C[0] = C.x; q.C[1] = C.y; q.C[2] = C.z; q.D[0] = D.x; q.D[1] = D.y; q.D[2] = D.z;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        q.dAB[k] = fsub_rn(q.B[k], q.A[k]);
        q.dDC[k] = fsub_rn(q.C[k], q.D[k]);
        const float ext = fmaxf(fmaxf(fabsf(q.dAB[k]), fabsf(q.dDC[k])), fabsf(fsub_rn(q.D[k], q.A[k])));
        const float amin = fminf(fabsf(q.A[k]), fabsf(q.D[k]));
        // NaN corners fail the first comparison; 1e-20: no underflow in the error terms of the products
        ok = ok && (amin >= fmul_rn(margin, ext)) && (amin >= 1e-20f || (amin == 0.0f && ext == 0.0f));
    }
    q.ok = ok;
    return q;
}

// The FP32-pipe blend of ONE pixel (gtp_fast5_f32.cu: a thread per quad, 5 km -> 1 km).  The two lanes hold
// (x, y) of the upper edge, (x, y) of the lower edge, and - for z - (upper edge, lower edge).  At 5 km the
// along-track weight s_t = (j - 2) / 5 is not a dyadic fraction, so the second blend (:398) is an edge blend of
// its own: (1.0 - s_t) s1 + (float)(s_t s2) = s1 + s_t (s2 - s1) - ep with s2 - s1 exact (Sterbenz).
// Returns the displacement of the fine xyz from corner A (exact); acc collects the rounding-boundary flags.
struct QuadEftP2 { EdgeP2 top, bot, zz; float A[3]; };

GTP_FD QuadEftP2 quad_eft_pack(const QuadEft& q)
{
    QuadEftP2 e;
    e.top.A = p2_make(q.A[0], q.A[1]); e.top.B = p2_make(q.B[0], q.B[1]); e.top.d = p2_make(q.dAB[0], q.dAB[1]);
    e.bot.A = p2_make(q.D[0], q.D[1]); e.bot.B = p2_make(q.C[0], q.C[1]); e.bot.d = p2_make(q.dDC[0], q.dDC[1]);
    e.zz.A = p2_make(q.A[2], q.D[2]); e.zz.B = p2_make(q.B[2], q.C[2]); e.zz.d = p2_make(q.dAB[2], q.dDC[2]);
    e.A[0] = q.A[0]; e.A[1] = q.A[1]; e.A[2] = q.A[2];
    return e;
}

GTP_FD void eft_pixel(const QuadEftP2& q, float a, P2 t2, P2 nt2, P2& acc, float (&d)[3])
{
    const P2 a2 = p2_dup(a), na2 = p2_dup(-a), m1 = p2_dup(-1.0f);
    const P2 s1 = eft_edge(a2, na2, q.top, acc), s2 = eft_edge(a2, na2, q.bot, acc);
    const F2 sz = p2_get(eft_edge(a2, na2, q.zz, acc));
    EdgeP2 track;
    track.A = s1; track.B = s2; track.d = p2_fma(s1, m1, s2);
    const F2 vxy = p2_get(eft_edge(t2, nt2, track, acc));
    track.A = p2_dup(sz.x); track.B = p2_dup(sz.y); track.d = p2_dup(fsub_rn(sz.y, sz.x));
    const F2 vz = p2_get(eft_edge(t2, nt2, track, acc));
    d[0] = fsub_rn(vxy.x, q.A[0]); d[1] = fsub_rn(vxy.y, q.A[1]); d[2] = fsub_rn(vz.x, q.A[2]);
}

// xyz2lonlat of the float32 xyz, anchor-relative in double, result rounded to float (:51-65)
// (dx, dy, dz) = the float32 xyz minus corner A (exact either way), zd = the float32 z
template <int MODE, bool WRAP>
GTP_FD void backtransform_f32_d(const QuadF32& q, double dx, double dy, double dz, double zd, float& lon, float& lat)
{
    const double e = fma(q.ue0, dx, q.ue1 * dy);
    const double h = fma(q.uh0, dx, q.uh1 * dy);
    const double w = dz;
    const double g = 1.0 + h;
    double rcp = fma(h, h - 1.0, 1.0);
    rcp = fma(rcp, fma(-g, rcp, 1.0), rcp);
    const double t = e * rcp;
    const double t2 = t * t;
    const double at = fma(t * t2, fma(t2, 0.2, -kThird), t);
    double lo = fma(at, kRad2Deg, q.lon0);
    if (WRAP) {
        if (lo > 180.0) lo -= 360.0;
        else if (lo < -180.0) lo += 360.0;
    }
    lon = (float)lo;
    double la_low = 0.0, la_high = 0.0;
    if (MODE & kLow)
        la_low = fma(w, fma(w, fma(w, fma(w, fma(w, q.a5, q.a4), q.a3), q.a2), q.a1), q.lat0_low);
    if (MODE & kHigh) {
        const double vv = fma(g, t2 * fma(t2, fma(t2, 0.0625, -0.125), 0.5), h);
        la_high = fma(vv, fma(vv, fma(vv, fma(vv, fma(vv, q.b5, q.b4), q.b3), q.b2), q.b1), q.lat0_high);
    }
    double la;
    if (MODE == kLow) la = la_low;
    else if (MODE == kHigh) la = la_high;
    else {
        // the reference's own test, on the float32 z promoted to double, thr = 0.8f (:62)
        const double thr_r = kLatSwitchF32 * GTP_EARTH_RADIUS;
        la = (zd < thr_r && zd > -thr_r) ? la_low : la_high;
    }
    lat = (float)la;
}

template <int MODE, bool WRAP>
GTP_FD void backtransform_f32(const QuadF32& q, const double (&v)[3], float& lon, float& lat)
{
    backtransform_f32_d<MODE, WRAP>(q, v[0] - q.A[0], v[1] - q.A[1], v[2] - q.A[2], v[2], lon, lat);
}

// The same back-transform for the P pixels of a fine row side by side (their dependent FP64 chains
// interleave), from the displacements (dx, dy, dz) = float32 xyz - corner A.  One latitude series per
// quad: HIGH in v = rho / rho_A - 1 (|z| >= 0.8 R), else in dz.
struct WideBT { double ue0, ue1, lon0, c0, c1, c2, c3, c4, c5; };

template <bool HIGH>
GTP_FD WideBT wide_bt_setup(const QuadF32& q)
{
    WideBT b;
    b.ue0 = q.ue0; b.ue1 = q.ue1; b.lon0 = q.lon0;
    b.c0 = HIGH ? q.lat0_high : q.lat0_low;
    b.c1 = HIGH ? q.b1 : q.a1; b.c2 = HIGH ? q.b2 : q.a2; b.c3 = HIGH ? q.b3 : q.a3;
    b.c4 = HIGH ? q.b4 : q.a4; b.c5 = HIGH ? q.b5 : q.a5;
    return b;
}

template <int P, bool HIGH>
GTP_FD void backtransform_f32_wide(const WideBT& b, const double (&dx)[P], const double (&dy)[P], const double (&dz)[P],
                                   float (&vlon)[P], float (&vlat)[P])
{
    // xyz2lonlat, anchor-relative (backtransform_f32_d), the P pixels side by side
    double e[P], h[P], g[P], rcp[P], t[P], t2[P];
#pragma unroll
    for (int k = 0; k < P; ++k) e[k] = b.ue1 * dy[k];
#pragma unroll
    for (int k = 0; k < P; ++k) h[k] = -b.ue0 * dy[k];
#pragma unroll
    for (int k = 0; k < P; ++k) e[k] = fma(b.ue0, dx[k], e[k]);
#pragma unroll
    for (int k = 0; k < P; ++k) h[k] = fma(b.ue1, dx[k], h[k]);
    // 1 / (1 + h) = (1 - h) (1 + h^2) (1 + h^4) (1 - h^8)^-1, |h| <= 0.01: three levels deep
#pragma unroll
    for (int k = 0; k < P; ++k) rcp[k] = 1.0 - h[k];
#pragma unroll
    for (int k = 0; k < P; ++k) t2[k] = h[k] * h[k];
#pragma unroll
    for (int k = 0; k < P; ++k) g[k] = 1.0 + h[k];
#pragma unroll
    for (int k = 0; k < P; ++k) rcp[k] = fma(t2[k], rcp[k], rcp[k]);
#pragma unroll
    for (int k = 0; k < P; ++k) t2[k] = t2[k] * t2[k];
#pragma unroll
    for (int k = 0; k < P; ++k) rcp[k] = fma(t2[k], rcp[k], rcp[k]);
#pragma unroll
    for (int k = 0; k < P; ++k) t[k] = e[k] * rcp[k];
#pragma unroll
    for (int k = 0; k < P; ++k) t2[k] = t[k] * t[k];
#pragma unroll
    for (int k = 0; k < P; ++k) e[k] = fma(t2[k], 0.2, -kThird);
#pragma unroll
    for (int k = 0; k < P; ++k) rcp[k] = t[k] * t2[k];
#pragma unroll
    for (int k = 0; k < P; ++k) t[k] = fma(rcp[k], e[k], t[k]);
#pragma unroll
    for (int k = 0; k < P; ++k) vlon[k] = (float)fma(t[k], kRad2Deg, b.lon0);
    double x[P], pl[P];
    if (HIGH) {
#pragma unroll
        for (int k = 0; k < P; ++k) x[k] = fma(t2[k], 0.0625, -0.125);
#pragma unroll
        for (int k = 0; k < P; ++k) x[k] = fma(t2[k], x[k], 0.5);
#pragma unroll
        for (int k = 0; k < P; ++k) x[k] = t2[k] * x[k];
#pragma unroll
        for (int k = 0; k < P; ++k) x[k] = fma(g[k], x[k], h[k]);
    } else {
#pragma unroll
        for (int k = 0; k < P; ++k) x[k] = dz[k];
    }
#pragma unroll
    for (int k = 0; k < P; ++k) pl[k] = fma(x[k], b.c5, b.c4);
#pragma unroll
    for (int k = 0; k < P; ++k) pl[k] = fma(x[k], pl[k], b.c3);
#pragma unroll
    for (int k = 0; k < P; ++k) pl[k] = fma(x[k], pl[k], b.c2);
#pragma unroll
    for (int k = 0; k < P; ++k) pl[k] = fma(x[k], pl[k], b.c1);
#pragma unroll
    for (int k = 0; k < P; ++k) vlat[k] = (float)fma(x[k], pl[k], b.c0);
}

struct LonLatF { float lon, lat; };

// the reference's formulas with library functions, for quads the series cannot take
GTP_FD_COLD LonLatF backtransform_f32_slow(float xf, float yf, float zf)
{
    const double x = (double)xf, y = (double)yf, z = (double)zf;
    const double rho = sqrt(x * x + y * y);
    LonLatF r;
    r.lon = (float)((acos(x / rho) * GTP_RAD2DEG) * (double)sign_of(y));
    const double thr_r = kLatSwitchF32 * GTP_EARTH_RADIUS;
    if ((z < thr_r) && (z > -thr_r)) r.lat = (float)(90.0 - acos(z / GTP_EARTH_RADIUS) * GTP_RAD2DEG);
    else r.lat = (float)((double)sign_of(z) * (90.0 - asin(rho / GTP_EARTH_RADIUS) * GTP_RAD2DEG));
    return r;
}

// ---- simple_modis_interpolator, float32 (_simple_modis_interpolator.pyx:13-277) --------------------
// The reference samples the float32 tie-point xyz with scipy's order-1 map_coordinates (double
// accumulation t = 0; t += (v wy) wx per tap, result cast to float32, :121-132), then extrapolates the
// last nc = 3 | 1 fine columns (:139-150) and the first / last 2 | 1 fine rows of the scan (:157-277)
// in float32 arithmetic.  The corners are float32 and the weights multiples of 1/8, so every tap product
// and partial sum is exact in double: the cast to float32 is the only rounding, and any way of writing
// the bilinear form gives scipy's value bit for bit (tests/test_fast_math_host.py checks the float32 xyz
// against the oracle with array_equal).
// One fine row of a quad: its left edge L = (1 - ty) A + ty D and right-minus-left dR, per component.
// All of it is exact in double (24-bit corners, 3-bit weights), which is why the evaluation order is free.
struct SimpleRowF32 { double L[3], dR[3]; };

GTP_FD SimpleRowF32 simple_row_f32(const QuadF32& q, double ty)
{
    SimpleRowF32 r;
    const double wy0 = 1.0 - ty;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        r.L[c] = fma(q.D[c], ty, q.A[c] * wy0);
        r.dR[c] = fma((double)q.C[c], ty, (double)q.B[c] * wy0) - r.L[c];
    }
    return r;
}

// the sample at column fraction tx (k / P inside the quad), a float32 value held in a double
GTP_FD void simple_sample_f32(const SimpleRowF32& r, double tx, double (&v)[3])
{
#pragma unroll
    // the conversion pair rounds exactly like the reference's cast (ties included) and is cheaper here than the
    // three-instruction split on the FP64 pipe, which is this kernel's bottleneck (+5 % at 250 m, round 2 A/B)
    for (int c = 0; c < 3; ++c) v[c] = (double)(float)fma(tx, r.dR[c], r.L[c]);
}

// The pixels right of the last tie column (fine columns Wf - P .. Wf - 1, evaluated on the quad of tie
// columns 1352 | 1353): column Wf - P sits on tie column 1353 (both taps clamp to it, 'nearest'), the
// nc = P - 1 columns after it are that value + diff (o + 1), diff = v[Wf - nc - 1] - v[Wf - nc - 2] (:146-150)
GTP_FD void simple_sample_right_f32(const SimpleRowF32& r, int P, int k, double (&v)[3])
{
    double prev[3];
    if (k > 0) simple_sample_f32(r, (double)(P - 1) / (double)P, prev);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double e = round_to_f32(r.L[c] + r.dR[c]);
        if (k == 0) { v[c] = e; continue; }
        const float ef = (float)e;
        const float diff = fsub_rn(ef, (float)prev[c]);
        v[c] = (double)fadd_rn(ef, fmul_rn(diff, (float)k));
    }
}

// first / last rows of a scan: the line through the samples of two interior rows, float32 arithmetic
// (_compute_slope_offset / _extrapolate, :227-277): m = (v_hi - v_lo) / (y_hi - y_lo), b = v_hi - m y_hi
GTP_FD float simple_line_f32(float v_lo, float v_hi, float y_lo, float y_hi, float y)
{
#ifdef __CUDA_ARCH__
    const float m = __fdiv_rn(fsub_rn(v_hi, v_lo), fsub_rn(y_hi, y_lo));
#else
    volatile float m = fsub_rn(v_hi, v_lo) / fsub_rn(y_hi, y_lo);
#endif
    const float b = fsub_rn(v_hi, fmul_rn(m, y_hi));
    return fadd_rn(fmul_rn(m, y), b);
}

// yc[j] of _compute_yx_coordinate_arrays (:71-81): j / res - (res / 16 + 1 / 8), exact in float32
GTP_FD float simple_ycoord_f32(int res, int j)
{
    return (float)((double)j * (1.0 / (double)res) - ((double)res * (1.0 / 16.0) + 0.125));
}

// What one fine row m of quad row qr (0..8) needs: the row itself, or - for the first / last rows of a
// scan - the two interior rows the extrapolation line goes through (:157-225): 250 m rows (2, 5) /
// (34, 37), 500 m rows (1, 2) / (17, 18), and their yc.
template <int P>
struct SimpleRowsF32 {
    bool edge;
    SimpleRowF32 a, b;          // the row | the line's two rows
    float y_lo, y_hi, y;
};

template <int P>
GTP_FD SimpleRowsF32<P> simple_rows_f32(const QuadF32& q, int qr, int m)
{
    constexpr int kHalf = P / 2;
    SimpleRowsF32<P> r;
    const double ty0 = ((qr == 0) ? (0.5 - kHalf) : 0.5) / P;
    const bool top = (qr == 0) && (m < kHalf), bottom = (qr == 8) && (m >= P);
    r.edge = top || bottom;
    r.y_lo = r.y_hi = r.y = 0.0f;
    if (!r.edge) {
        r.a = simple_row_f32(q, ty0 + (double)m / (double)P);
        r.b = r.a;
        return r;
    }
    const int m_lo = top ? kHalf : 0, m_hi = top ? kHalf + P - 1 : P - 1;
    const int j0 = (qr == 0) ? 0 : 8 * P + kHalf;                       // scan row of m = 0
    r.a = simple_row_f32(q, ty0 + (double)m_lo / (double)P);
    r.b = simple_row_f32(q, ty0 + (double)m_hi / (double)P);
    r.y_lo = simple_ycoord_f32(P, j0 + m_lo); r.y_hi = simple_ycoord_f32(P, j0 + m_hi); r.y = simple_ycoord_f32(P, j0 + m);
    return r;
}

// pixel k of the quad column; `right`: the lane that evaluates the pixels right of the last tie column
template <int P>
GTP_FD void simple_pixel_f32(const SimpleRowsF32<P>& r, bool right, int k, double (&v)[3])
{
    if (!r.edge) {
        if (right) simple_sample_right_f32(r.a, P, k, v);
        else simple_sample_f32(r.a, (double)k / (double)P, v);
        return;
    }
    double lo[3], hi[3];
    if (right) { simple_sample_right_f32(r.a, P, k, lo); simple_sample_right_f32(r.b, P, k, hi); }
    else { simple_sample_f32(r.a, (double)k / (double)P, lo); simple_sample_f32(r.b, (double)k / (double)P, hi); }
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (double)simple_line_f32((float)lo[c], (float)hi[c], r.y_lo, r.y_hi, r.y);
}

}  // namespace gtpfast32
