// host_emulation.cpp - runs the float64 fast-path MATH of the CUDA kernel on the CPU.
//
// TEST INFRASTRUCTURE (tests/test_fast_math_host.py).  It includes the very header the
// kernel is built from (csrc/gtp_fast_math.cuh, __host__ __device__ functions) and replaces
// only the warp orchestration of gtp_fast.cu by plain loops, so the series, the acceptance
// test and the fallback can be checked against the oracle without a GPU.  It is not part of
// the product and is never loaded by geotiepoints_b200.
#include <cstdint>
#include <vector>
#include "../python-geotiepoints_b200/geotiepoints_b200/csrc/gtp_fast_math.cuh"
#include "../python-geotiepoints_b200/geotiepoints_b200/csrc/gtp_fast_f32_math.cuh"

using namespace gtpfast;

// simple == true: simple_modis_interpolator = the same bilinear form with c_exp = c_ali = 0.
// Like the kernel, every tie column is walked down the ten tie rows of a scan with tie_advance
// (row 0: full sincos, rows 1..9: small rotations of the row above).
template <int P>
static void run(const double* lon, const double* lat, const double* satz, long n_scans,
                double* out_lon, double* out_lat, long* mode_counts, bool simple = false)
{
    const int W = 1354, L = 10, Wf = W * P, N = L * P, H = P / 2;
    std::vector<TieGeo> geo((size_t)L * W);
    std::vector<TieZen> zen((size_t)L * W);
    for (long s = 0; s < n_scans; ++s) {
        for (int c = 0; c < W; ++c) {
            TieGeo g = {0.0, 1.0, 0.0, 1.0, 0.0, 0.0};
            TieZen z = {0.0, 0.0, 1.0, 0.0, 1.0};
            for (int r = 0; r < L; ++r) {
                const long o = (s * L + r) * W + c;
                if (simple) tie_advance<false>(g, z, lon[o], lat[o], 0.0, r == 0);
                else tie_advance<true>(g, z, lon[o], lat[o], satz[o], r == 0);
                geo[(size_t)r * W + c] = g; zen[(size_t)r * W + c] = z;
            }
        }
        for (int qc = 0; qc <= W - 1; ++qc) {            // qc == W-1: the "extra" pixels of the last quad
            const bool extra = (qc == W - 1);
            const int ca = extra ? W - 2 : qc, cb = ca + 1, xoff = extra ? P : 0;
            for (int qr = 0; qr < L - 1; ++qr) {
                const TieGeo A = geo[(size_t)qr * W + ca];
                const Vec3 Ax = tie_xyz(A), B = tie_xyz(geo[(size_t)qr * W + cb]),
                           D = tie_xyz(geo[(size_t)(qr + 1) * W + ca]), C = tie_xyz(geo[(size_t)(qr + 1) * W + cb]);
                double ce = 0.0, cal = 0.0;
                if (!simple) {
                    const TieZen za = zen[(size_t)qr * W + ca], zb = zen[(size_t)qr * W + cb];
                    quad_coefficients(za, zb.za, zb.sphi, zb.cphi, 1, ce, cal);
                }
                Quad q;
                quad_setup(q, A, Ax, B, C, D, ce, cal, extra ? (double)kABoundExtra : 1.0);
                mode_counts[q.mode == kWide ? 5 : q.mode]++;
                if (q.mode != kSlow && q.mode != kWide && q.small) mode_counts[4]++;
                const int j0 = (qr == 0) ? 0 : qr * P + H;
                const int j1 = (qr == L - 2) ? N : (qr + 1) * P + H;
                for (int j = j0; j < j1; ++j) {
                    const double s_t = fine_row_y<P>(j) / (double)P;
                    const RowCoef rc = (q.mode != kSlow) ? row_setup(q, s_t, s_t * (1.0 - s_t)) : RowCoef();
                    for (int k = 0; k < P; ++k) {
                        const double s_s = (double)(xoff + k) / (double)P;
                        double lo, la;
                        const double a_col = fma(s_s * (1.0 - s_s), q.ce, s_s);
                        if (q.mode == kSlow) { const LonLat ll = pixel_eval_slow(Ax, B, C, D, q.ce, q.ca, s_s, s_t); lo = ll.lon; la = ll.lat; }
                        else if (q.mode == kWide) pixel_eval_wide(q, rc, a_col, lo, la);
                        else if (q.wrap) pixel_eval<kBoth, true>(q, rc, a_col, lo, la);
                        else if (q.small) pixel_eval<kLowSmall, false>(q, rc, a_col, lo, la);
                        else if (q.mode == kLow) pixel_eval<kLow, false>(q, rc, a_col, lo, la);
                        else if (q.mode == kHigh) pixel_eval<kHigh, false>(q, rc, a_col, lo, la);
                        else pixel_eval<kBoth, false>(q, rc, a_col, lo, la);
                        const long o = (s * N + j) * (long)Wf + (long)ca * P + xoff + k;
                        out_lon[o] = lo; out_lat[o] = la;
                    }
                }
            }
        }
    }
}

extern "C" int gtp_emul_simple_fast_f64(const double* lon, const double* lat, long n_scans, int fine_res,
                                        double* out_lon, double* out_lat, long* mode_counts)
{
    for (int k = 0; k < 6; ++k) mode_counts[k] = 0;
    if (fine_res == 250) run<4>(lon, lat, nullptr, n_scans, out_lon, out_lat, mode_counts, true);
    else if (fine_res == 500) run<2>(lon, lat, nullptr, n_scans, out_lon, out_lat, mode_counts, true);
    else return 1;
    return 0;
}

extern "C" int gtp_emul_cviirs_fast_f64(const double* lon, const double* lat, const double* satz, long n_scans,
                                        int fine_res, double* out_lon, double* out_lat, long* mode_counts)
{
    for (int k = 0; k < 6; ++k) mode_counts[k] = 0;
    if (fine_res == 250) run<4>(lon, lat, satz, n_scans, out_lon, out_lat, mode_counts);
    else if (fine_res == 500) run<2>(lon, lat, satz, n_scans, out_lon, out_lat, mode_counts);
    else return 1;
    return 0;
}


// ---- ECEF output (csrc/gtp_fast_xyz.cu): the same tie-point walk, then the bilinear form itself ----
template <int P>
static void run_xyz(const double* lon, const double* lat, const double* satz, long n_scans,
                    double* out_x, double* out_y, double* out_z)
{
    const int W = 1354, L = 10, Wf = W * P, N = L * P, H = P / 2;
    const bool simple = (satz == nullptr);
    std::vector<TieGeo> geo((size_t)L * W);
    std::vector<TieZen> zen((size_t)L * W);
    for (long s = 0; s < n_scans; ++s) {
        for (int c = 0; c < W; ++c) {
            TieGeo g = {0.0, 1.0, 0.0, 1.0, 0.0, 0.0};
            TieZen z = {0.0, 0.0, 1.0, 0.0, 1.0};
            for (int r = 0; r < L; ++r) {
                const long o = (s * L + r) * W + c;
                if (simple) tie_advance<false>(g, z, lon[o], lat[o], 0.0, r == 0);
                else tie_advance<true>(g, z, lon[o], lat[o], satz[o], r == 0);
                geo[(size_t)r * W + c] = g; zen[(size_t)r * W + c] = z;
            }
        }
        for (int qc = 0; qc <= W - 1; ++qc) {
            const bool extra = (qc == W - 1);
            const int ca = extra ? W - 2 : qc, cb = ca + 1, xoff = extra ? P : 0;
            for (int qr = 0; qr < L - 1; ++qr) {
                double ce = 0.0, cal = 0.0;
                if (!simple) {
                    const TieZen za = zen[(size_t)qr * W + ca], zb = zen[(size_t)qr * W + cb];
                    quad_coefficients(za, zb.za, zb.sphi, zb.cphi, 1, ce, cal);
                }
                const QuadXyz q = quadxyz_setup(tie_xyz(geo[(size_t)qr * W + ca]), tie_xyz(geo[(size_t)qr * W + cb]),
                                                tie_xyz(geo[(size_t)(qr + 1) * W + cb]), tie_xyz(geo[(size_t)(qr + 1) * W + ca]));
                const int j0 = (qr == 0) ? 0 : qr * P + H;
                const int j1 = (qr == L - 2) ? N : (qr + 1) * P + H;
                for (int j = j0; j < j1; ++j) {
                    const double s_t = fine_row_y<P>(j) / (double)P;
                    const RowXyz rw = rowxyz_setup(q, s_t, simple ? 0.0 : (s_t * (1.0 - s_t)) * cal);
                    for (int k = 0; k < P; ++k) {
                        const double s_s = (double)(xoff + k) / (double)P;
                        const double a_col = simple ? s_s : fma(s_s * (1.0 - s_s), ce, s_s);
                        const long o = (s * N + j) * (long)Wf + (long)ca * P + xoff + k;
                        out_x[o] = fma(a_col, rw.U.x, rw.V.x);
                        out_y[o] = fma(a_col, rw.U.y, rw.V.y);
                        out_z[o] = fma(a_col, rw.U.z, rw.V.z);
                    }
                }
            }
        }
    }
}

// satz == NULL: the simple interpolator
extern "C" int gtp_emul_xyz_fast_f64(const double* lon, const double* lat, const double* satz, long n_scans,
                                     int fine_res, double* out_x, double* out_y, double* out_z)
{
    if (fine_res == 250) run_xyz<4>(lon, lat, satz, n_scans, out_x, out_y, out_z);
    else if (fine_res == 500) run_xyz<2>(lon, lat, satz, n_scans, out_x, out_y, out_z);
    else return 1;
    return 0;
}

// ---- float64 5 km -> 1 km fast path (csrc/gtp_fast5.cu) ------------------------------------------
extern "C" int gtp_emul_cviirs_fast5_f64(const double* lon, const double* lat, const double* satz, long n_scans,
                                         int W, double* out_lon, double* out_lat, long* mode_counts)
{
    for (int k = 0; k < 6; ++k) mode_counts[k] = 0;
    if (W != 270 && W != 271) return 1;
    const int Wf = 1354, nq = W - 1;
    for (long s = 0; s < n_scans; ++s) {
        for (int qc = 0; qc < nq; ++qc) {
            TieGeo top[2], bot[2];
            TieZen zen[2], unused = {0.0, 0.0, 1.0, 0.0, 1.0};
            for (int k = 0; k < 2; ++k) {
                const long o = (s * 2) * W + qc + k;
                top[k] = TieGeo{0.0, 1.0, 0.0, 1.0, 0.0, 0.0};
                zen[k] = TieZen{0.0, 0.0, 1.0, 0.0, 1.0};
                tie_first<true>(top[k], zen[k], lon[o], lat[o], satz[o]);
                bot[k] = top[k];
                tie_next<false>(bot[k], unused, lon[o + W], lat[o + W], 0.0);
            }
            const Vec3 A = tie_xyz(top[0]), B = tie_xyz(top[1]), D = tie_xyz(bot[0]), C = tie_xyz(bot[1]);
            double ce, cal;
            quad_coefficients(zen[0], zen[1].za, zen[1].sphi, zen[1].cphi, 5, ce, cal);
            const double a_max = (qc != nq - 1) ? 1.0 : (W == 271 ? 1.4 : 3.0);
            Quad q;
            const QuadSize sz = quad_project(q, top[0], A, B, C, D, a_max);
            quad_classify_long(q, top[0], sz, ce, cal);
            mode_counts[q.mode]++;
            if (q.mode != kSlow && q.small) mode_counts[4]++;
            const int i0 = (qc == 0) ? 0 : 2 + 5 * qc, i1 = (qc == nq - 1) ? Wf : 2 + 5 * qc + 5;
            for (int j = 0; j < 10; ++j) {
                const double s_t = (double)(j - 2) / 5.0;
                const RowCoef rc = (q.mode != kSlow) ? row_setup(q, s_t, s_t * (1.0 - s_t)) : RowCoef();
                for (int i = i0; i < i1; ++i) {
                    const double s_s = (double)coord_x_5km(i, W) / 5.0;
                    const double a_col = fma(s_s * (1.0 - s_s), q.ce, s_s);
                    double lo, la;
                    if (q.mode == kSlow) { const LonLat ll = pixel_eval_slow(A, B, C, D, q.ce, q.ca, s_s, s_t); lo = ll.lon; la = ll.lat; }
                    else if (q.small && q.mode == kLow) pixel_eval<kLow, false>(q, rc, a_col, lo, la);      // the kernel: whole warps only
                    else if (q.small && q.mode == kHigh) pixel_eval<kHigh, false>(q, rc, a_col, lo, la);
                    else if (!q.wrap && q.mode == kLow) pixel_eval_long<kLow>(q, rc, a_col, lo, la);
                    else if (!q.wrap && q.mode == kHigh) pixel_eval_long<kHigh>(q, rc, a_col, lo, la);
                    else pixel_eval_long<kBoth>(q, rc, a_col, lo, la);
                    const long o = (s * 10 + j) * (long)Wf + i;
                    out_lon[o] = lo; out_lat[o] = la;
                }
            }
        }
    }
    return 0;
}

// ---- float32 fast path (gtp_fast_f32_math.cuh) -------------------------------------------------
// one quad with the blend on the FP32 pipe, exactly as quad_rows_eft of gtp_fast_f32.cu does it
template <int P, bool HIGH>
static bool eft_quad(const gtpfast32::QuadF32& q, const gtpfast32::QuadEft& qe, bool extra, float s_t, int j0, int j1,
                     float* out_lon, float* out_lat, long base, long Wf, long* mode_counts)
{
    using namespace gtpfast32;
    EdgeP2 top[3], bot[3];
    for (int c = 0; c < 3; ++c) {
        top[c].A = p2_dup(qe.A[c]); top[c].B = p2_dup(qe.B[c]); top[c].d = p2_dup(qe.dAB[c]);
        bot[c].A = p2_dup(qe.D[c]); bot[c].B = p2_dup(qe.C[c]); bot[c].d = p2_dup(qe.dDC[c]);
    }
    const WideBT bt = wide_bt_setup<HIGH>(q);
    const double ced = (double)q.ce, xo = extra ? 1.0 : 0.0;
    double a_col[P];
    for (int k = 0; k < P; ++k) { const double ss = xo + (double)k / P; a_col[k] = fma(ss * (1.0 - ss), ced, ss); }
    const P2 m1 = p2_dup(-1.0f);
    P2 acc = p2_dup(0.0f);
    for (int j = j0; j < j1; ++j, s_t += 1.0f / P) {
        const RowF32 rc = row_setup_f32(q, s_t);
        const P2 st2 = p2_dup(s_t), omt2 = p2_dup(fsub_rn(1.0f, s_t));
        double dx[P], dy[P], dz[P];
        for (int k = 0; k < P; k += 2) {
            const P2 a2 = p2_make((float)add_rn(a_col[k], rc.row_ca), (float)add_rn(a_col[k + 1], rc.row_ca));
            const P2 na2 = p2_mul(a2, m1);
            F2 d[3];
            const P2 acc_before = acc;
            for (int c = 0; c < 3; ++c) {
                const P2 s1 = eft_edge(a2, na2, top[c], acc), s2 = eft_edge(a2, na2, bot[c], acc);
                d[c] = p2_get(eft_track(s1, s2, st2, omt2, top[c].A));
            }
            dx[k] = (double)d[0].x; dx[k + 1] = (double)d[0].y;
            dy[k] = (double)d[1].x; dy[k + 1] = (double)d[1].y;
            dz[k] = (double)d[2].x; dz[k + 1] = (double)d[2].y;
            // cross-check against the reference's arithmetic wherever nothing was flagged in this pair
            const F2 f0 = p2_get(acc_before), f1 = p2_get(acc);
            if (f0.x == f1.x && f0.y == f1.y) {
                for (int u = 0; u < 2; ++u) {
                    const float s_s = (float)((extra ? P : 0) + k + u) / (float)P;
                    double v[3];
                    blend_f32(q, rc, s_s, mul_rn((double)s_s, add_rn(1.0, -(double)s_s)), v);
                    if (v[0] - q.A[0] != dx[k + u] || v[1] - q.A[1] != dy[k + u] || v[2] - q.A[2] != dz[k + u]) mode_counts[7]++;
                }
            }
        }
        float vlon[P], vlat[P];
        backtransform_f32_wide<P, HIGH>(bt, dx, dy, dz, vlon, vlat);
        for (int k = 0; k < P; ++k) { out_lon[base + j * Wf + k] = vlon[k]; out_lat[base + j * Wf + k] = vlat[k]; }
    }
    const F2 fl = p2_get(acc);
    return !(fl.x == 0.0f) || !(fl.y == 0.0f);
}

template <int P>
static void run_f32(const float* lon, const float* lat, const float* satz, long n_scans,
                    float* out_lon, float* out_lat, long* mode_counts)
{
    using namespace gtpfast32;
    const int W = 1354, L = 10, Wf = W * P, N = L * P, H = P / 2;
    std::vector<TieColF> col((size_t)L * W);
    for (long s = 0; s < n_scans; ++s) {
        // like the kernel: every tie column is walked down the ten tie rows (row 0: full evaluation)
        for (int c = 0; c < W; ++c) {
            TieColF t;
            for (int r = 0; r < L; ++r) {
                const long o = (s * L + r) * W + c;
                if (r == 0) tie_col_first<true>(t, lon[o], lat[o], satz[o]);
                else tie_col_next<true>(t, lon[o], lat[o], satz[o]);
                col[(size_t)r * W + c] = t;
            }
        }
        for (int qc = 0; qc <= W - 1; ++qc) {
            const bool extra = (qc == W - 1);
            const int ca = extra ? W - 2 : qc, cb = ca + 1, xoff = extra ? P : 0;
            for (int qr = 0; qr < L - 1; ++qr) {
                auto geo = [&](int r, int c) { return tie_col_geo(col[(size_t)r * W + c]); };
                auto v3 = [](const TieGeoF& t) { Vec3F v; v.x = t.x; v.y = t.y; v.z = t.z; return v; };
                const TieGeoF A = geo(qr, ca);
                const Vec3F B = v3(geo(qr, cb)), D = v3(geo(qr + 1, ca)), C = v3(geo(qr + 1, cb));
                const TieZenF za = tie_col_zen(col[(size_t)qr * W + ca]), zb = tie_col_zen(col[(size_t)qr * W + cb]);
                float ce, cal;
                quad_exp_ali_col(col[(size_t)qr * W + ca], za, zb.phi, zb.theta, 1, ce, cal);
                const QuadF32 q = quad_setup_f32(A, B, C, D, ce, cal);
                mode_counts[q.mode]++;
                const int j0 = (qr == 0) ? 0 : qr * P + H;
                const int j1 = (qr == L - 2) ? N : (qr + 1) * P + H;
                float s_t = ((qr == 0) ? (0.5f - H) : 0.5f) / P;
                Vec3F Av; Av.x = A.x; Av.y = A.y; Av.z = A.z;
                const QuadEft qe = quad_eft_setup(Av, B, C, D, extra ? 16.0f : 8.0f);
                // like the kernel: the FP32-pipe blend for plain low / high latitude quads, and the whole quad
                // again with blend_f32 when any of its values sat on a rounding boundary
                bool faithful = !qe.ok || q.wrap || q.mode == kBoth || q.mode == kSlow;
                if (!faithful) {
                    mode_counts[5]++;
                    faithful = (q.mode == kHigh) ? eft_quad<P, true>(q, qe, extra, s_t, j0, j1, out_lon, out_lat, (s * N) * (long)Wf + (long)ca * P + xoff, Wf, mode_counts)
                                                 : eft_quad<P, false>(q, qe, extra, s_t, j0, j1, out_lon, out_lat, (s * N) * (long)Wf + (long)ca * P + xoff, Wf, mode_counts);
                    if (faithful) mode_counts[6]++;
                }
                if (!faithful) continue;
                for (int j = j0; j < j1; ++j, s_t += 1.0f / P) {
                    const RowF32 rc = row_setup_f32(q, s_t);
                    for (int k = 0; k < P; ++k) {
                        const float s_s = (float)(xoff + k) / (float)P;
                        const double k1d = mul_rn((double)s_s, add_rn(1.0, -(double)s_s));
                        double v[3];
                        float lo, la;
                        blend_f32(q, rc, s_s, k1d, v);
                        if (q.mode == kSlow) { const LonLatF ll = backtransform_f32_slow((float)v[0], (float)v[1], (float)v[2]); lo = ll.lon; la = ll.lat; }
                        else backtransform_f32<kBoth, true>(q, v, lo, la);
                        const long o = (s * N + j) * (long)Wf + (long)ca * P + xoff + k;
                        out_lon[o] = lo; out_lat[o] = la;
                    }
                }
            }
        }
    }
}

// mode_counts[5..7]: quads on the FP32-pipe blend, of those evaluated again with blend_f32, pixels where the
// FP32-pipe blend differs from blend_f32 although no value was flagged (must stay 0)
extern "C" int gtp_emul_cviirs_fast_f32(const float* lon, const float* lat, const float* satz, long n_scans,
                                        int fine_res, float* out_lon, float* out_lat, long* mode_counts)
{
    for (int k = 0; k < 8; ++k) mode_counts[k] = 0;
    if (fine_res == 250) run_f32<4>(lon, lat, satz, n_scans, out_lon, out_lat, mode_counts);
    else if (fine_res == 500) run_f32<2>(lon, lat, satz, n_scans, out_lon, out_lat, mode_counts);
    else return 1;
    return 0;
}

// ---- float32 simple interpolator (simple_pixel_f32 of gtp_fast_f32_math.cuh) ----------------------
// xyz != NULL additionally returns the float32 fine Cartesian coordinates (for the bit-exactness check
// against the oracle's new_xyz)
template <int P>
static void run_simple_f32(const float* lon, const float* lat, long n_scans, float* out_lon, float* out_lat,
                           float* const* xyz, long* mode_counts)
{
    using namespace gtpfast32;
    const int W = 1354, L = 10, Wf = W * P, N = L * P, H = P / 2;
    std::vector<TieColF> col((size_t)L * W);
    for (long s = 0; s < n_scans; ++s) {
        for (int c = 0; c < W; ++c) {
            TieColF t;
            for (int r = 0; r < L; ++r) {
                const long o = (s * L + r) * W + c;
                if (r == 0) tie_col_first<false>(t, lon[o], lat[o], 0.0f);
                else tie_col_next<false>(t, lon[o], lat[o], 0.0f);
                col[(size_t)r * W + c] = t;
            }
        }
        for (int qc = 0; qc <= W - 1; ++qc) {
            const bool extra = (qc == W - 1);
            const int ca = extra ? W - 2 : qc, cb = ca + 1, xoff = extra ? P : 0;
            for (int qr = 0; qr < L - 1; ++qr) {
                auto geo = [&](int r, int c) { return tie_col_geo(col[(size_t)r * W + c]); };
                auto v3 = [](const TieGeoF& t) { Vec3F v; v.x = t.x; v.y = t.y; v.z = t.z; return v; };
                const TieGeoF A = geo(qr, ca);
                const Vec3F B = v3(geo(qr, cb)), D = v3(geo(qr + 1, ca)), C = v3(geo(qr + 1, cb));
                const QuadF32 q = quad_setup_f32(A, B, C, D, 0.0f, 0.0f);
                mode_counts[q.mode]++;
                const int nrows = P + ((qr == 0 || qr == L - 2) ? H : 0);
                const int j0 = (qr == 0) ? 0 : qr * P + H;
                for (int m = 0; m < nrows; ++m) {
                    const SimpleRowsF32<P> rows = simple_rows_f32<P>(q, qr, m);
                    for (int k = 0; k < P; ++k) {
                        double v[3];
                        float lo, la;
                        simple_pixel_f32<P>(rows, extra, k, v);
                        if (q.mode == kSlow) { const LonLatF ll = backtransform_f32_slow((float)v[0], (float)v[1], (float)v[2]); lo = ll.lon; la = ll.lat; }
                        else if (q.wrap) backtransform_f32<kBoth, true>(q, v, lo, la);
                        else if (q.mode == kLow) backtransform_f32<kLow, false>(q, v, lo, la);
                        else if (q.mode == kHigh) backtransform_f32<kHigh, false>(q, v, lo, la);
                        else backtransform_f32<kBoth, false>(q, v, lo, la);
                        const long o = (s * N + j0 + m) * (long)Wf + (long)ca * P + xoff + k;
                        out_lon[o] = lo; out_lat[o] = la;
                        if (xyz) for (int c = 0; c < 3; ++c) xyz[c][o] = (float)v[c];
                    }
                }
            }
        }
    }
}

extern "C" int gtp_emul_simple_fast_f32(const float* lon, const float* lat, long n_scans, int fine_res,
                                        float* out_lon, float* out_lat, float* out_x, float* out_y, float* out_z,
                                        long* mode_counts)
{
    for (int k = 0; k < 5; ++k) mode_counts[k] = 0;
    float* const xyz[3] = {out_x, out_y, out_z};
    if (fine_res == 250) run_simple_f32<4>(lon, lat, n_scans, out_lon, out_lat, out_x ? xyz : nullptr, mode_counts);
    else if (fine_res == 500) run_simple_f32<2>(lon, lat, n_scans, out_lon, out_lat, out_x ? xyz : nullptr, mode_counts);
    else return 1;
    return 0;
}

// ---- float32 5 km -> 1 km fast path (csrc/gtp_fast5_f32.cu) ----------------------------------------------
// like the kernel: tie column walked from the upper to the lower tie row, quad_exp_ali_col, the blend on the
// FP32 pipe (eft_pixel) for plain low / high latitude quads - the whole quad again with blend_f32 when a value
// sat on a rounding boundary -, blend_f32 for the rest.  mode_counts[4]: pixels whose FP32-pipe blend differs from
// blend_f32 without having been flagged (must stay 0); [5]: quads on the FP32 pipe.
extern "C" int gtp_emul_cviirs5_series_f32(const float* lon, const float* lat, const float* satz, long n_scans,
                                           int W, float* out_lon, float* out_lat, long* mode_counts)
{
    using namespace gtpfast32;
    for (int k = 0; k < 6; ++k) mode_counts[k] = 0;
    if (W != 270 && W != 271) return 1;
    const int Wf = 1354, nq = W - 1;
    std::vector<TieColF> top(W), bot(W);
    for (long s = 0; s < n_scans; ++s) {
        for (int c = 0; c < W; ++c) {
            TieColF t;
            const long o = (s * 2) * W + c;
            tie_col_first<true>(t, lon[o], lat[o], satz[o]);
            top[c] = t;
            tie_col_next<false>(t, lon[o + W], lat[o + W], 0.0f);
            bot[c] = t;
        }
        for (int qc = 0; qc < nq; ++qc) {
            const TieGeoF A = tie_col_geo(top[qc]);
            const Vec3F B = tie_col_xyz(top[qc + 1]), D = tie_col_xyz(bot[qc]), C = tie_col_xyz(bot[qc + 1]);
            const TieZenF za = tie_col_zen(top[qc]), zb = tie_col_zen(top[qc + 1]);
            float ce, cal;
            quad_exp_ali_col(top[qc], za, zb.phi, zb.theta, 5, ce, cal);
            const QuadF32 q = quad_setup_f32(A, B, C, D, ce, cal, kMaxRelHoriz5, kMaxRelZ5);
            Vec3F Av; Av.x = A.x; Av.y = A.y; Av.z = A.z;
            const QuadEft qe0 = quad_eft_setup(Av, B, C, D, (qc == nq - 1) ? 32.0f : 8.0f);
            const QuadEftP2 qe = quad_eft_pack(qe0);
            mode_counts[q.mode]++;
            const int i0 = (qc == 0) ? 0 : 2 + 5 * qc, i1 = (qc == nq - 1) ? Wf : 2 + 5 * qc + 5;
            bool eft = q.mode != kSlow && qe0.ok && !q.wrap && q.mode != kBoth;
            if (eft) mode_counts[5]++;
            for (int pass = 0; pass < 2; ++pass) {
                P2 acc = p2_dup(0.0f);
                for (int j = 0; j < 10; ++j) {
                    const float s_t = (float)(j - 2) / 5.0f;
                    const RowF32 rc = row_setup_f32(q, s_t);
                    const P2 t2 = p2_dup(s_t), nt2 = p2_dup(-s_t);
                    for (int i = i0; i < i1; ++i) {
                        const float s_s = (float)coord_x_5km(i, W) / 5.0f;
                        const double k1d = mul_rn((double)s_s, add_rn(1.0, -(double)s_s));
                        double v[3];
                        float lo, la;
                        blend_f32(q, rc, s_s, k1d, v);
                        if (eft) {
                            const float a = (float)add_rn(add_rn((double)s_s, mul_rn(k1d, (double)q.ce)), rc.row_ca);
                            float d[3];
                            const P2 before = acc;
                            eft_pixel(qe, a, t2, nt2, acc, d);
                            const F2 f0 = p2_get(before), f1 = p2_get(acc);
                            if (f0.x == f1.x && f0.y == f1.y
                                && ((double)d[0] != v[0] - q.A[0] || (double)d[1] != v[1] - q.A[1] || (double)d[2] != v[2] - q.A[2])) mode_counts[4]++;
                            if (q.mode == kHigh) backtransform_f32_d<kHigh, false>(q, (double)d[0], (double)d[1], (double)d[2], 0.0, lo, la);
                            else backtransform_f32_d<kLow, false>(q, (double)d[0], (double)d[1], (double)d[2], 0.0, lo, la);
                        } else if (q.mode == kSlow) { const LonLatF ll = backtransform_f32_slow((float)v[0], (float)v[1], (float)v[2]); lo = ll.lon; la = ll.lat; }
                        else backtransform_f32<kBoth, true>(q, v, lo, la);
                        const long o = (s * 10 + j) * (long)Wf + i;
                        out_lon[o] = lo; out_lat[o] = la;
                    }
                }
                const F2 fl = p2_get(acc);
                if (!eft || ((fl.x == 0.0f) && (fl.y == 0.0f))) break;
                eft = false;            // flagged: the literal arithmetic for the whole quad
            }
        }
    }
    return 0;
}
