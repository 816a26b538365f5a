// gtp_fast_f32.cu - float32 1 km -> 250 m / 500 m fast paths (CVIIRS and simple interpolator) for sm_100a.
//
// Same work decomposition as gtp_fast.cu (one thread per tie-point quad column, ten tie rows
// top to bottom, right neighbour by warp shuffle, cp.async prefetch of the next tie row, one
// 128-bit / 64-bit streaming store per lane, fine row and output array); the math is
// gtp_fast_f32_math.cuh: the reference's float32 rounding points for everything up to the fine
// Cartesian coordinates, then the anchor-relative double-precision back-transform.
// Compiled with -fmad=false: the rounding-faithful part must not be contracted; the series use
// explicit fma().
#include <cuda_runtime.h>
#include "gtp_kernels.h"
#include "gtp_fast_f32_math.cuh"

namespace {

using namespace gtpfast32;

constexpr int kW = 1354;
constexpr int kL = 10;
constexpr int kQuadColsPerWarp = 31;
constexpr int kWarpsPerScan = (kW + kQuadColsPerWarp - 1) / kQuadColsPerWarp;   // 44
constexpr int kThreads = 128;
constexpr unsigned kFull = 0xffffffffu;

template <int P>
__device__ __forceinline__ void store_px(float* p, const float (&v)[P]);

template <>
__device__ __forceinline__ void store_px<4>(float* p, const float (&v)[4])
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}

template <>
__device__ __forceinline__ void store_px<2>(float* p, const float (&v)[2])
{
    asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
}

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int P, int MODE, bool WRAP>
__device__ __forceinline__ void quad_rows(const QuadF32& q, const float (&s_s)[P], const double (&k1d)[P],
                                          float s_t0, int nrows, float* olon, float* olat)
{
    constexpr long long kPitch = (long long)kW * P;
    float s_t = s_t0;
#pragma unroll 1
    for (int m = 0; m < nrows; ++m) {
        const RowF32 rc = row_setup_f32(q, s_t);
        float vlon[P], vlat[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
            double v[3];
            blend_f32(q, rc, s_s[k], k1d[k], v);
            if (MODE == kSlow) {
                const LonLatF ll = backtransform_f32_slow((float)v[0], (float)v[1], (float)v[2]);
                vlon[k] = ll.lon; vlat[k] = ll.lat;
            } else {
                backtransform_f32<MODE, WRAP>(q, v, vlon[k], vlat[k]);
            }
        }
        store_px<P>(olon, vlon);
        store_px<P>(olat, vlat);
        olon += kPitch; olat += kPitch;
        s_t += 1.0f / P;             // exact: multiples of 1/8
    }
}

// The same rows with the blend on the FP32 pipe (gtp_fast_f32_math.cuh, "error-free transforms"): pixel
// pairs in the two lanes of the packed float instructions, one component after the other; the displacement
// from corner A goes to double for the back-transform (3 conversions per pixel instead of 23), which is
// written step by step over the P pixels of the row so that their dependent FP64 chains interleave (one
// pixel after the other they ran at the FP64 latency, 300 stall samples per instruction).  HIGH: the
// latitude series in v = rho / rho_A - 1 (|z| >= 0.8 R) instead of w = dz / R; quads that need the
// per-pixel latitude switch or the +-180 wrap are rare and stay with quad_rows.  Returns true when any
// value of the quad sat within 2^-20 of a rounding boundary (one quad in ~1500): the caller then evaluates
// the whole quad again with blend_f32 (the stores are idempotent).
template <int P, bool HIGH>
__device__ __forceinline__ bool quad_rows_eft(const QuadF32& q, const QuadEft& qe, bool extra, float s_t0, int nrows,
                                              float* olon, float* olat)
{
    constexpr long long kPitch = (long long)kW * P;
    EdgeP2 top[3], bot[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        top[c].A = p2_dup(qe.A[c]); top[c].B = p2_dup(qe.B[c]); top[c].d = p2_dup(qe.dAB[c]);
        bot[c].A = p2_dup(qe.D[c]); bot[c].B = p2_dup(qe.C[c]); bot[c].d = p2_dup(qe.dDC[c]);
    }
    const WideBT bt = wide_bt_setup<HIGH>(q);
    // s_s + s_s (1 - s_s) c_exp (:344): s_s = (xoff + k) / P, exact products, one rounding
    const double ced = (double)q.ce, xo = extra ? 1.0 : 0.0;
    double a_col[P];
#pragma unroll
    for (int k = 0; k < P; ++k) {
        const double ss = xo + (double)k / P;
        a_col[k] = fma(ss * (1.0 - ss), ced, ss);
    }
    const P2 m1 = p2_dup(-1.0f);
    P2 acc = p2_dup(0.0f);
    float s_t = s_t0;
#pragma unroll 1
    for (int m = 0; m < nrows; ++m) {
        const RowF32 rc = row_setup_f32(q, s_t);
        const P2 st2 = p2_dup(s_t), omt2 = p2_dup(fsub_rn(1.0f, s_t));      // exact: multiples of 1/8
        float af[P];
#pragma unroll
        for (int k = 0; k < P; ++k) af[k] = (float)add_rn(a_col[k], rc.row_ca);
        double dx[P], dy[P], dz[P];
#pragma unroll
        for (int k = 0; k < P; k += 2) {
            const P2 a2 = p2_make(af[k], af[k + 1]);
            const P2 na2 = p2_mul(a2, m1);
            F2 d[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const P2 s1 = eft_edge(a2, na2, top[c], acc);
                const P2 s2 = eft_edge(a2, na2, bot[c], acc);
                d[c] = p2_get(eft_track(s1, s2, st2, omt2, top[c].A));
            }
            dx[k] = (double)d[0].x; dx[k + 1] = (double)d[0].y;
            dy[k] = (double)d[1].x; dy[k + 1] = (double)d[1].y;
            dz[k] = (double)d[2].x; dz[k + 1] = (double)d[2].y;
        }
        float vlon[P], vlat[P];
        backtransform_f32_wide<P, HIGH>(bt, dx, dy, dz, vlon, vlat);
        store_px<P>(olon, vlon);
        store_px<P>(olat, vlat);
        olon += kPitch; olat += kPitch;
        s_t += 1.0f / P;
    }
    const F2 fl = p2_get(acc);
    return !(fl.x == 0.0f) || !(fl.y == 0.0f);
}

template <int P>
__global__ void __launch_bounds__(kThreads, 4)
cviirs_fast_1km_f32_kernel(const float* __restrict__ lon, const float* __restrict__ lat,
                           const float* __restrict__ satz, float* __restrict__ out_lon,
                           float* __restrict__ out_lat, long long n_scans)
{
    constexpr int kWf = kW * P;
    constexpr int kN = kL * P;
    constexpr int kHalf = P / 2;

    const long long warp_global = (blockIdx.x * (long long)kThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long scan = warp_global / kWarpsPerScan;
    if (scan >= n_scans) return;
    const int w = (int)(warp_global - scan * kWarpsPerScan);
    const int c = w * kQuadColsPerWarp + lane;
    const int tc = min(c, kW - 1);
    const bool extra = (c == kW - 1);      // the lane of tie column 1353: pixels right of the last quad
    const bool quad_lane = (lane < kQuadColsPerWarp) && (c <= kW - 2);
    const bool active = quad_lane || extra;
    const int quad_col = extra ? kW - 2 : c;
    const int xoff = extra ? P : 0;

    const long long scan_base = scan * (long long)(kL * kW);
    const float* plon = lon + scan_base + tc;
    const float* plat = lat + scan_base + tc;
    const float* psatz = satz + scan_base + tc;
    float* olon = out_lon + scan * (long long)kN * kWf + (long long)quad_col * P + xoff;
    float* olat = out_lat + scan * (long long)kN * kWf + (long long)quad_col * P + xoff;

    float s_s[P];
    double k1d[P];
#pragma unroll
    for (int k = 0; k < P; ++k) {
        s_s[k] = (float)(xoff + k) / (float)P;                      // x[i] / f, a float division (:341)
        k1d[k] = mul_rn((double)s_s[k], add_rn(1.0, -(double)s_s[k]));
    }

    const bool last_warp = (w == kWarpsPerScan - 1);      // holds tie column 1353 and its left neighbour
    TieColF st;                                            // own tie column, walked down the scan
    TieGeoF topA;
    Vec3F topB;
    float top_ce = 0.0f, top_ca = 0.0f;
    __shared__ float stage[2][3][kThreads];
    cp_async4(&stage[0][0][threadIdx.x], plon);
    cp_async4(&stage[0][1][threadIdx.x], plat);
    cp_async4(&stage[0][2][threadIdx.x], psatz);
    cp_async_commit();

#pragma unroll 1
    for (int r = 0; r < kL; ++r) {
        cp_async_wait_all();
        const float clon = stage[r & 1][0][threadIdx.x], clat = stage[r & 1][1][threadIdx.x],
                    csatz = stage[r & 1][2][threadIdx.x];
        if (r + 1 < kL) {
            plon += kW; plat += kW; psatz += kW;
            cp_async4(&stage[(r + 1) & 1][0][threadIdx.x], plon);
            cp_async4(&stage[(r + 1) & 1][1][threadIdx.x], plat);
            cp_async4(&stage[(r + 1) & 1][2][threadIdx.x], psatz);
            cp_async_commit();
        }
        if (r == 0) tie_col_first<true>(st, clon, clat, csatz);
        else tie_col_next<true>(st, clon, clat, csatz);
        const TieGeoF own = tie_col_geo(st);
        const TieZenF zen = tie_col_zen(st);
        TieGeoF curA = own;
        Vec3F curB;
        curB.x = __shfl_down_sync(kFull, own.x, 1);
        curB.y = __shfl_down_sync(kFull, own.y, 1);
        curB.z = __shfl_down_sync(kFull, own.z, 1);
        TieColF zst = st;                // the state the quad's zenith chain starts from: its upper LEFT corner's
        TieZenF za = zen;
        float phi_b = __shfl_down_sync(kFull, zen.phi, 1);
        float theta_b = __shfl_down_sync(kFull, zen.theta, 1);
        if (last_warp) {
            // the lane of tie column 1353 evaluates the quad of columns 1352 | 1353 right of its last tie
            // column: A = the left neighbour's tie point (by shuffle), B = its own
            TieGeoF lg; TieColF ls; TieZenF lz;
            lg.x = __shfl_up_sync(kFull, own.x, 1); lg.y = __shfl_up_sync(kFull, own.y, 1); lg.z = __shfl_up_sync(kFull, own.z, 1);
            ls.lam = __shfl_up_sync(kFull, st.lam, 1); ls.sl = __shfl_up_sync(kFull, st.sl, 1); ls.cl = __shfl_up_sync(kFull, st.cl, 1);
            ls.phi = __shfl_up_sync(kFull, st.phi, 1); ls.sp = __shfl_up_sync(kFull, st.sp, 1); ls.cp = __shfl_up_sync(kFull, st.cp, 1);
            ls.za = __shfl_up_sync(kFull, st.za, 1); ls.sz = __shfl_up_sync(kFull, st.sz, 1); ls.cz = __shfl_up_sync(kFull, st.cz, 1);
            ls.pt = __shfl_up_sync(kFull, st.pt, 1); ls.spt = __shfl_up_sync(kFull, st.spt, 1); ls.cpt = __shfl_up_sync(kFull, st.cpt, 1);
            lz.za = __shfl_up_sync(kFull, zen.za, 1); lz.phi = __shfl_up_sync(kFull, zen.phi, 1); lz.theta = __shfl_up_sync(kFull, zen.theta, 1);
            if (extra) {
                lg.lam = ls.lam; lg.phi = ls.phi; lg.sl = ls.sl; lg.cl = ls.cl; lg.sp = ls.sp; lg.cp = ls.cp;
                curA = lg;
                curB.x = own.x; curB.y = own.y; curB.z = own.z;
                zst = ls; za = lz;
                phi_b = zen.phi; theta_b = zen.theta;
            }
        }

        if (r > 0 && active) {
            const int qr = r - 1;
            const int nrows = P + ((qr == 0 || qr == kL - 2) ? kHalf : 0);
            const float s_t0 = ((qr == 0) ? (0.5f - kHalf) : 0.5f) / P;     // y[j] / f (:342)
            Vec3F curAv; curAv.x = curA.x; curAv.y = curA.y; curAv.z = curA.z;
            const QuadF32 q = quad_setup_f32(topA, topB, curB, curAv, top_ce, top_ca);
            Vec3F topAv; topAv.x = topA.x; topAv.y = topA.y; topAv.z = topA.z;
            const QuadEft qe = quad_eft_setup(topAv, topB, curB, curAv, extra ? 16.0f : 8.0f);
            // FP32-pipe blend for plain low- / high-latitude quads; the rest (a component near 0, the latitude
            // switch or the +-180 wrap inside the quad, a value on a rounding boundary) the reference's way
            bool faithful = !qe.ok || q.wrap || q.mode == kBoth;
            if (q.mode == kSlow) quad_rows<P, kSlow, false>(q, s_s, k1d, s_t0, nrows, olon, olat);
            else {
                if (!faithful) faithful = (q.mode == kHigh) ? quad_rows_eft<P, true>(q, qe, extra, s_t0, nrows, olon, olat)
                                                            : quad_rows_eft<P, false>(q, qe, extra, s_t0, nrows, olon, olat);
                if (faithful) quad_rows<P, kBoth, true>(q, s_s, k1d, s_t0, nrows, olon, olat);
            }
            olon += (long long)nrows * kWf;
            olat += (long long)nrows * kWf;
        }
        topA = curA;
        topB = curB;

        // expansion / alignment of the quad row below this tie row (upper corners only, :307-308)
        if (r + 1 < kL && active) quad_exp_ali_col(zst, za, phi_b, theta_b, 1, top_ce, top_ca);
    }
}

template <int P>
cudaError_t launch(const float* lon, const float* lat, const float* satz, float* out_lon, float* out_lat,
                   long long n_scans, cudaStream_t stream)
{
    const long long warps = n_scans * kWarpsPerScan;
    const long long blocks = (warps * 32 + kThreads - 1) / kThreads;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    cviirs_fast_1km_f32_kernel<P><<<(unsigned)blocks, kThreads, 0, stream>>>(lon, lat, satz, out_lon, out_lat, n_scans);
    return cudaGetLastError();
}

// ---- simple_modis_interpolator, float32: same decomposition, no zenith angles ----
template <int MODE, bool WRAP>
__device__ __forceinline__ void simple_backtransform(const QuadF32& q, const double (&v)[3], float& lon, float& lat)
{
    if (MODE == kSlow) {
        const LonLatF ll = backtransform_f32_slow((float)v[0], (float)v[1], (float)v[2]);
        lon = ll.lon; lat = ll.lat;
    } else {
        backtransform_f32<MODE, WRAP>(q, v, lon, lat);
    }
}

// The fine rows of one quad.  Interior rows (all but the first / last 2 | 1 rows of a scan) are one
// bilinear sample per pixel; the scan-edge rows, which only quad rows 0 and 8 have, go through the
// reference's float32 line fit (simple_rows_f32 / simple_pixel_f32) in a loop of their own.
template <int P, int MODE, bool WRAP>
__device__ __forceinline__ void quad_rows_simple(const QuadF32& q, bool right, int qr, int nrows, float* olon, float* olat)
{
    constexpr long long kPitch = (long long)kW * P;
    constexpr int kHalf = P / 2;
    const int m0 = (qr == 0) ? kHalf : 0;                       // interior rows [m0, m0 + P)
    double ty = 0.5 / P;                                        // row fraction of row m0: the first fine row below the tie row
#pragma unroll 1
    for (int m = m0; m < m0 + P; ++m) {
        const SimpleRowF32 row = simple_row_f32(q, ty);
        float vlon[P], vlat[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
            double v[3];
            if (right) simple_sample_right_f32(row, P, k, v);
            else simple_sample_f32(row, (double)k / (double)P, v);
            simple_backtransform<MODE, WRAP>(q, v, vlon[k], vlat[k]);
        }
        store_px<P>(olon + m * kPitch, vlon);
        store_px<P>(olat + m * kPitch, vlat);
        ty += 1.0 / P;
    }
    if (nrows == P) return;
    const int e0 = (qr == 0) ? 0 : P;                           // scan-edge rows [e0, e0 + kHalf)
#pragma unroll 1
    for (int m = e0; m < e0 + kHalf; ++m) {
        const SimpleRowsF32<P> rows = simple_rows_f32<P>(q, qr, m);
        float vlon[P], vlat[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
            double v[3];
            simple_pixel_f32<P>(rows, right, k, v);
            simple_backtransform<MODE, WRAP>(q, v, vlon[k], vlat[k]);
        }
        store_px<P>(olon + m * kPitch, vlon);
        store_px<P>(olat + m * kPitch, vlat);
    }
}

template <int P>
__global__ void __launch_bounds__(kThreads, 4)
simple_fast_1km_f32_kernel(const float* __restrict__ lon, const float* __restrict__ lat,
                           float* __restrict__ out_lon, float* __restrict__ out_lat, long long n_scans)
{
    constexpr int kWf = kW * P;
    constexpr int kN = kL * P;
    constexpr int kHalf = P / 2;

    const long long warp_global = (blockIdx.x * (long long)kThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long scan = warp_global / kWarpsPerScan;
    if (scan >= n_scans) return;
    const int w = (int)(warp_global - scan * kWarpsPerScan);
    const int c = w * kQuadColsPerWarp + lane;
    const int tc = min(c, kW - 1);
    const bool extra = (c == kW - 1);      // the lane of tie column 1353: pixels right of the last tie column
    const bool quad_lane = (lane < kQuadColsPerWarp) && (c <= kW - 2);
    const bool active = quad_lane || extra;
    const int quad_col = extra ? kW - 2 : c;
    const int xoff = extra ? P : 0;

    const long long scan_base = scan * (long long)(kL * kW);
    const float* plon = lon + scan_base + tc;
    const float* plat = lat + scan_base + tc;
    float* olon = out_lon + scan * (long long)kN * kWf + (long long)quad_col * P + xoff;
    float* olat = out_lat + scan * (long long)kN * kWf + (long long)quad_col * P + xoff;

    TieColF st;                    // own tie column, walked down the scan
    TieGeoF topA;
    Vec3F topB;
    __shared__ float stage[2][2][kThreads];
    cp_async4(&stage[0][0][threadIdx.x], plon);
    cp_async4(&stage[0][1][threadIdx.x], plat);
    cp_async_commit();

#pragma unroll 1
    for (int r = 0; r < kL; ++r) {
        cp_async_wait_all();
        const float clon = stage[r & 1][0][threadIdx.x], clat = stage[r & 1][1][threadIdx.x];
        if (r + 1 < kL) {
            plon += kW; plat += kW;
            cp_async4(&stage[(r + 1) & 1][0][threadIdx.x], plon);
            cp_async4(&stage[(r + 1) & 1][1][threadIdx.x], plat);
            cp_async_commit();
        }
        if (r == 0) tie_col_first<false>(st, clon, clat, 0.0f);
        else tie_col_next<false>(st, clon, clat, 0.0f);
        const TieGeoF own = tie_col_geo(st);
        TieGeoF curA = own;
        Vec3F curB;
        curB.x = __shfl_down_sync(kFull, own.x, 1);
        curB.y = __shfl_down_sync(kFull, own.y, 1);
        curB.z = __shfl_down_sync(kFull, own.z, 1);
        if (w == kWarpsPerScan - 1) {
            // the lane of tie column 1353: A = the left neighbour's tie point (by shuffle), B = its own
            TieGeoF lg;
            lg.x = __shfl_up_sync(kFull, own.x, 1); lg.y = __shfl_up_sync(kFull, own.y, 1); lg.z = __shfl_up_sync(kFull, own.z, 1);
            lg.lam = __shfl_up_sync(kFull, own.lam, 1); lg.sl = __shfl_up_sync(kFull, own.sl, 1); lg.cl = __shfl_up_sync(kFull, own.cl, 1);
            lg.phi = __shfl_up_sync(kFull, own.phi, 1); lg.sp = __shfl_up_sync(kFull, own.sp, 1); lg.cp = __shfl_up_sync(kFull, own.cp, 1);
            if (extra) { curA = lg; curB.x = own.x; curB.y = own.y; curB.z = own.z; }
        }

        if (r > 0 && active) {
            const int qr = r - 1;
            const int nrows = P + ((qr == 0 || qr == kL - 2) ? kHalf : 0);
            Vec3F curAv; curAv.x = curA.x; curAv.y = curA.y; curAv.z = curA.z;
            const QuadF32 q = quad_setup_f32(topA, topB, curB, curAv, 0.0f, 0.0f);
            if (q.mode == kSlow) quad_rows_simple<P, kSlow, false>(q, extra, qr, nrows, olon, olat);
            else if (q.wrap) quad_rows_simple<P, kBoth, true>(q, extra, qr, nrows, olon, olat);
            else if (q.mode == kLow) quad_rows_simple<P, kLow, false>(q, extra, qr, nrows, olon, olat);
            else if (q.mode == kHigh) quad_rows_simple<P, kHigh, false>(q, extra, qr, nrows, olon, olat);
            else quad_rows_simple<P, kBoth, false>(q, extra, qr, nrows, olon, olat);
            olon += (long long)nrows * kWf;
            olat += (long long)nrows * kWf;
        }
        topA = curA;
        topB = curB;
    }
}

template <int P>
cudaError_t launch_simple(const float* lon, const float* lat, float* out_lon, float* out_lat, long long n_scans, cudaStream_t stream)
{
    const long long warps = n_scans * kWarpsPerScan;
    const long long blocks = (warps * 32 + kThreads - 1) / kThreads;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    simple_fast_1km_f32_kernel<P><<<(unsigned)blocks, kThreads, 0, stream>>>(lon, lat, out_lon, out_lat, n_scans);
    return cudaGetLastError();
}

}  // namespace

bool gtp_cviirs_fast_f32_supported(const CviirsCfg& cfg)
{
    return cfg.km == 1 && (cfg.p == 4 || cfg.p == 2);
}

cudaError_t gtp_launch_cviirs_fast_f32(const float* lon, const float* lat, const float* satz,
                                       float* out_lon, float* out_lat,
                                       long long n_scans, const CviirsCfg& cfg, cudaStream_t stream)
{
    if (n_scans <= 0) return cudaSuccess;
    const uintptr_t align = (cfg.p == 4) ? 16 : 8;       // 128-bit / 64-bit stores; the row pitch already is
    if (((uintptr_t)out_lon | (uintptr_t)out_lat) & (align - 1))
        return gtp_launch_cviirs_generic<float>(lon, lat, satz, out_lon, out_lat, n_scans, cfg, stream);
    return cfg.p == 4 ? launch<4>(lon, lat, satz, out_lon, out_lat, n_scans, stream)
                      : launch<2>(lon, lat, satz, out_lon, out_lat, n_scans, stream);
}

bool gtp_simple_fast_f32_supported(int width, int res)
{
    return width == kW && (res == 4 || res == 2);
}

cudaError_t gtp_launch_simple_fast_f32(const float* lon, const float* lat, float* out_lon, float* out_lat,
                                       long long n_scans, int res, cudaStream_t stream)
{
    if (n_scans <= 0) return cudaSuccess;
    const uintptr_t align = (res == 4) ? 16 : 8;
    if (((uintptr_t)out_lon | (uintptr_t)out_lat) & (align - 1))
        return gtp_launch_simple_generic<float>(lon, lat, out_lon, out_lat, n_scans, kW, res, stream);
    return res == 4 ? launch_simple<4>(lon, lat, out_lon, out_lat, n_scans, stream)
                    : launch_simple<2>(lon, lat, out_lon, out_lat, n_scans, stream);
}
