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
// from corner A goes to double for the back-transform (3 conversions per pixel instead of 23).  One
// instance serves low- and high-latitude quads (the latitude series is chosen per quad: `high`); quads that
// need the per-pixel latitude switch or the +-180 wrap are rare and stay with quad_rows.  Returns true when
// any value of the quad sat within 2^-20 of a rounding boundary (one quad in ~1500): the caller then
// evaluates the whole quad again with blend_f32 (the stores are idempotent).
template <int P>
__device__ __forceinline__ bool quad_rows_eft(const QuadF32& q, const QuadEft& qe, const float (&s_s)[P],
                                              const double (&k1d)[P], float s_t0, int nrows, float* olon, float* olat)
{
    constexpr long long kPitch = (long long)kW * P;
    const bool high = (q.mode == kHigh);
    EdgeP2 top[3], bot[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        top[c].A = p2_dup(qe.A[c]); top[c].B = p2_dup(qe.B[c]); top[c].d = p2_dup(qe.dAB[c]);
        bot[c].A = p2_dup(qe.D[c]); bot[c].B = p2_dup(qe.C[c]); bot[c].d = p2_dup(qe.dDC[c]);
    }
    // the latitude series of this quad: in w = dz / R (low) or in v = rho / rho_A - 1 (high)
    const double c0 = high ? q.lat0_high : q.lat0_low;
    const double c1 = high ? q.b1 : q.a1, c2 = high ? q.b2 : q.a2, c3 = high ? q.b3 : q.a3,
                 c4 = high ? q.b4 : q.a4, c5 = high ? q.b5 : q.a5;
    const double ced = (double)q.ce;
    double a_col[P];                                    // s_s + s_s (1 - s_s) c_exp, the row-independent part of a_scan
#pragma unroll
    for (int k = 0; k < P; ++k) a_col[k] = add_rn((double)s_s[k], mul_rn(k1d[k], ced));
    const P2 m1 = p2_dup(-1.0f);
    P2 acc = p2_dup(0.0f);
    float s_t = s_t0;
#pragma unroll 1
    for (int m = 0; m < nrows; ++m) {
        const RowF32 rc = row_setup_f32(q, s_t);
        const P2 st2 = p2_dup(s_t), omt2 = p2_dup(fsub_rn(1.0f, s_t));      // exact: multiples of 1/8
        float vlon[P], vlat[P];
#pragma unroll
        for (int k = 0; k < P; k += 2) {
            const P2 a2 = p2_make((float)add_rn(a_col[k], rc.row_ca), (float)add_rn(a_col[k + 1], rc.row_ca));
            const P2 na2 = p2_mul(a2, m1);
            F2 d[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const P2 s1 = eft_edge(a2, na2, top[c], acc);
                const P2 s2 = eft_edge(a2, na2, bot[c], acc);
                d[c] = p2_get(eft_track(s1, s2, st2, omt2, top[c].A));
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const double dx = (double)(u ? d[0].y : d[0].x), dy = (double)(u ? d[1].y : d[1].x),
                             dz = (double)(u ? d[2].y : d[2].x);
                const double e = fma(q.ue0, dx, q.ue1 * dy);
                const double h = fma(q.uh0, dx, q.uh1 * dy);
                const double g = 1.0 + h;
                double rcp = fma(h, h - 1.0, 1.0);
                rcp = fma(rcp, fma(-g, rcp, 1.0), rcp);
                const double t = e * rcp;
                const double t2 = t * t;
                const double at = fma(t * t2, fma(t2, 0.2, -kThird), t);
                vlon[k + u] = (float)fma(at, kRad2Deg, q.lon0);
                double x = dz * kInvEarthR;
                if (high) x = fma(g, t2 * fma(t2, fma(t2, 0.0625, -0.125), 0.5), h);
                vlat[k + u] = (float)fma(x, fma(x, fma(x, fma(x, fma(x, c5, c4), c3), c2), c1), c0);
            }
        }
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
        const TieGeoF own = tie_geo_f32(clon, clat);
        TieGeoF curA = own;
        Vec3F curB;
        curB.x = __shfl_down_sync(kFull, own.x, 1);
        curB.y = __shfl_down_sync(kFull, own.y, 1);
        curB.z = __shfl_down_sync(kFull, own.z, 1);
        if (extra) {                 // A = tie column 1352 (recomputed by this one lane), B = own column 1353
            const long long o = scan_base + (long long)r * kW + (kW - 2);
            curA = tie_geo_f32(__ldg(lon + o), __ldg(lat + o));
            curB.x = own.x; curB.y = own.y; curB.z = own.z;
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
                if (!faithful) faithful = quad_rows_eft<P>(q, qe, s_s, k1d, s_t0, nrows, olon, olat);
                if (faithful) quad_rows<P, kBoth, true>(q, s_s, k1d, s_t0, nrows, olon, olat);
            }
            olon += (long long)nrows * kWf;
            olat += (long long)nrows * kWf;
        }
        topA = curA;
        topB = curB;

        if (r + 1 < kL) {
            // expansion / alignment of the quad row below this tie row (upper corners only, :307-308)
            const TieZenF zen = tie_zen_f32(csatz);
            float phi_b = __shfl_down_sync(kFull, zen.phi, 1);
            float theta_b = __shfl_down_sync(kFull, zen.theta, 1);
            TieZenF za = zen;
            if (extra) {
                za = tie_zen_f32(__ldg(satz + scan_base + (long long)r * kW + (kW - 2)));
                phi_b = zen.phi; theta_b = zen.theta;
            }
            quad_exp_ali_f32(za, phi_b, theta_b, 1, top_ce, top_ca);
        }
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
        const TieGeoF own = tie_geo_f32(clon, clat);
        TieGeoF curA = own;
        Vec3F curB;
        curB.x = __shfl_down_sync(kFull, own.x, 1);
        curB.y = __shfl_down_sync(kFull, own.y, 1);
        curB.z = __shfl_down_sync(kFull, own.z, 1);
        if (extra) {                 // A = tie column 1352 (recomputed by this one lane), B = own column 1353
            const long long o = scan_base + (long long)r * kW + (kW - 2);
            curA = tie_geo_f32(__ldg(lon + o), __ldg(lat + o));
            curB.x = own.x; curB.y = own.y; curB.z = own.z;
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
