// gtp_fast5_f32.cu - float32 CVIIRS 5 km -> 1 km fast path for sm_100a (BASELINE configs[2], float32).
//
// modis_5km_to_1km on float32 arrays - the only dtype the stock reference accepts for it.
// Same decomposition as gtp_fast5.cu (one THREAD owns one tie-point quad = 5 fine columns x 10 fine rows;
// right-hand corners from lane + 1 by shuffle; the partial columns at the swath edges go to otherwise idle
// "virtual" lanes; a warp stages every fine row in shared memory and stores it as one contiguous span), with
// the float32 arithmetic of gtp_fast_f32_math.cuh: the reference's float32 rounding points for the tie points
// and the expansion / alignment coefficients (tie column walked from the upper to the lower tie row), the
// blend on the FP32 pipe with error-free transforms (eft_pixel: both blends are edge blends here, s_t = (j - 2)
// / 5 is not dyadic), the anchor-relative double-precision back-transform around the float32-rounded corner A.
// A quad in which any value sat on a rounding boundary, quads with a Cartesian component near 0, quads that
// straddle the latitude switch or +-180 deg are evaluated with the literal float / double mix (blend_f32),
// polar quads with the library formulas.  Replaces the strip kernel of gtp_generic.cu for this configuration
// (thread per pixel, quad data re-read from shared memory per pixel: 6.9e10 px/s).
#include <cstdint>
#include <cuda_runtime.h>
#include "gtp_kernels.h"
#include "gtp_fast_f32_math.cuh"

namespace {

using namespace gtpfast32;

constexpr int kFineCols = 1354;
constexpr int kF = 5;
constexpr int kRows = 10;
constexpr int kQuadsPerWarp = 30;
constexpr int kThreads5 = 128;        // 4 independent warps (strips) per CTA: they start together and share instruction fetches
constexpr unsigned kFull = 0xffffffffu;
constexpr int kSpanMax = 152;

enum Path { kNone = 0, kEftLow = 1, kEftHigh = 2, kFaithful = 3, kLibrary = 4 };

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(kFull, v, src); }
__device__ __forceinline__ float shfl_f(float v, int src) { return __shfl_sync(kFull, v, src); }

// the inputs of a quad's set-up, so that a virtual lane can take them from the lane that owns the quad
struct QuadIn { TieGeoF A; Vec3F B, C, D; float ce, ca; };

__device__ __forceinline__ void take_quad_in(QuadIn& q, int src, bool take)
{
#define GTP_TAKE_F(f) { const float v_ = shfl_f(f, src); if (take) f = v_; }
#define GTP_TAKE_D(f) { const double v_ = shfl_d(f, src); if (take) f = v_; }
    GTP_TAKE_F(q.A.x) GTP_TAKE_F(q.A.y) GTP_TAKE_F(q.A.z)
    GTP_TAKE_D(q.A.lam) GTP_TAKE_D(q.A.phi) GTP_TAKE_D(q.A.sl) GTP_TAKE_D(q.A.cl) GTP_TAKE_D(q.A.sp) GTP_TAKE_D(q.A.cp)
    GTP_TAKE_F(q.B.x) GTP_TAKE_F(q.B.y) GTP_TAKE_F(q.B.z)
    GTP_TAKE_F(q.C.x) GTP_TAKE_F(q.C.y) GTP_TAKE_F(q.C.z)
    GTP_TAKE_F(q.D.x) GTP_TAKE_F(q.D.y) GTP_TAKE_F(q.D.z)
    GTP_TAKE_F(q.ce) GTP_TAKE_F(q.ca)
#undef GTP_TAKE_F
#undef GTP_TAKE_D
}

// One fine row of a lane's quad (kF pixels, masked) into the warp's staging buffer.
template <int PATH>
__device__ __forceinline__ void stage_row(const QuadF32& q, const QuadEftP2& qe, const float (&s_s)[kF], const double (&a_col)[kF],
                                          const double (&k1d)[kF], float s_t, unsigned kmask, P2& acc, float* my_lon, float* my_lat)
{
    const RowF32 rc = row_setup_f32(q, s_t);
    const P2 t2 = p2_dup(s_t), nt2 = p2_dup(-s_t);
#pragma unroll
    for (int k = 0; k < kF; ++k) {
        float lo, la;
        if (PATH == kEftLow || PATH == kEftHigh) {
            const float a = (float)add_rn(a_col[k], rc.row_ca);                       // a_scan as the reference rounds it (:344)
            float d[3];
            eft_pixel(qe, a, t2, nt2, acc, d);
            if (PATH == kEftLow) backtransform_f32_d<kLow, false>(q, (double)d[0], (double)d[1], (double)d[2], 0.0, lo, la);
            else backtransform_f32_d<kHigh, false>(q, (double)d[0], (double)d[1], (double)d[2], 0.0, lo, la);
        } else {
            double v[3];
            blend_f32(q, rc, s_s[k], k1d[k], v);
            if (PATH == kFaithful) backtransform_f32<kBoth, true>(q, v, lo, la);
            else { const LonLatF ll = backtransform_f32_slow((float)v[0], (float)v[1], (float)v[2]); lo = ll.lon; la = ll.lat; }
        }
        if ((kmask >> k) & 1u) { my_lon[k] = lo; my_lat[k] = la; }
    }
}

__device__ __forceinline__ void copy_row_out(const float2* span_lon, const float2* span_lat, int span2, int lane,
                                             float* dst_lon, float* dst_lat)
{
    __syncwarp();
    __stcs(reinterpret_cast<float2*>(dst_lon) + lane, span_lon[lane]);
    __stcs(reinterpret_cast<float2*>(dst_lat) + lane, span_lat[lane]);
    __stcs(reinterpret_cast<float2*>(dst_lon) + lane + 32, span_lon[lane + 32]);
    __stcs(reinterpret_cast<float2*>(dst_lat) + lane + 32, span_lat[lane + 32]);
    if (lane + 64 < span2) {
        __stcs(reinterpret_cast<float2*>(dst_lon) + lane + 64, span_lon[lane + 64]);
        __stcs(reinterpret_cast<float2*>(dst_lat) + lane + 64, span_lat[lane + 64]);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kThreads5, 16 * 32 / kThreads5)
cviirs_fast_5km_f32_kernel(const float* __restrict__ lon, const float* __restrict__ lat, const float* __restrict__ satz,
                           float* __restrict__ out_lon, float* __restrict__ out_lat, long long n_scans, int W, int warps_per_scan)
{
    const long long warp_global = (blockIdx.x * (long long)kThreads5 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long scan = warp_global / warps_per_scan;
    if (scan >= n_scans) return;
    const int w = (int)(warp_global - scan * warps_per_scan);
    const int n_quads = W - 1;
    const int qc = w * kQuadsPerWarp + lane;
    const int tc = min(qc, W - 1);
    const bool real = (lane < kQuadsPerWarp) && (qc < n_quads);
    const bool first_warp = (w == 0), last_warp = (w == warps_per_scan - 1);

    // ---- the two tie points of this lane's column: the upper one in full, the lower one by a step from it ----
    const long long base = scan * 2 * (long long)W + tc;
    TieColF col;
    tie_col_first<true>(col, __ldg(lon + base), __ldg(lat + base), __ldg(satz + base));
    QuadIn in;
    in.A = tie_col_geo(col);
    const TieZenF zen = tie_col_zen(col);
    const TieColF top = col;
    tie_col_next<false>(col, __ldg(lon + base + W), __ldg(lat + base + W), 0.0f);
    in.D = tie_col_xyz(col);
    in.B.x = __shfl_down_sync(kFull, in.A.x, 1); in.B.y = __shfl_down_sync(kFull, in.A.y, 1); in.B.z = __shfl_down_sync(kFull, in.A.z, 1);
    in.C.x = __shfl_down_sync(kFull, in.D.x, 1); in.C.y = __shfl_down_sync(kFull, in.D.y, 1); in.C.z = __shfl_down_sync(kFull, in.D.z, 1);
    const float phi_b = __shfl_down_sync(kFull, zen.phi, 1), theta_b = __shfl_down_sync(kFull, zen.theta, 1);
    in.ce = in.ca = 0.0f;
    if (real) quad_exp_ali_col(top, zen, phi_b, theta_b, 5, in.ce, in.ca);

    // ---- lane roles (as in gtp_fast5.cu): virtual lanes take over the partial columns at the swath edges ----
    int src_quad = qc, xoff = 0;
    unsigned kmask = real ? 0x1fu : 0u;
    if (first_warp || last_warp) {        // warp-uniform
        const int last_lane = (n_quads - 1) - w * kQuadsPerWarp;
        const int n_extra = (W == 271) ? 2 : 7;
        const bool v_first = first_warp && (lane == 31);
        const bool v_last1 = last_warp && (lane == last_lane + 1);
        const bool v_last2 = last_warp && (lane == last_lane + 2) && (n_extra > kF);
        if (first_warp) take_quad_in(in, 0, v_first);
        if (last_warp) take_quad_in(in, last_lane, v_last1 || v_last2);
        if (v_first) { src_quad = 0; xoff = -kF; kmask = 0x18u; }                       // k = 3, 4 -> x = -2, -1
        if (v_last1) { src_quad = n_quads - 1; xoff = kF; kmask = (n_extra >= kF) ? 0x1fu : ((1u << n_extra) - 1u); }
        if (v_last2) { src_quad = n_quads - 1; xoff = 2 * kF; kmask = (1u << (n_extra - kF)) - 1u; }
    }

    // ---- quad set-up ----
    const QuadF32 q = quad_setup_f32(in.A, in.B, in.C, in.D, in.ce, in.ca, kMaxRelHoriz5, kMaxRelZ5);
    Vec3F Av; Av.x = in.A.x; Av.y = in.A.y; Av.z = in.A.z;
    // the last quad extrapolates up to s_s = 2.2 (270 columns: |a_scan| <= 4.9 for |c_exp| <= 1): wider margin
    const QuadEft qe0 = quad_eft_setup(Av, in.B, in.C, in.D, (src_quad == n_quads - 1) ? 32.0f : 8.0f);
    const QuadEftP2 qe = quad_eft_pack(qe0);
    int path = kNone;
    if (kmask != 0u) {
        if (q.mode == kSlow) path = kLibrary;
        else if (!qe0.ok || q.wrap || q.mode == kBoth) path = kFaithful;
        else path = (q.mode == kHigh) ? kEftHigh : kEftLow;
    }

    float s_s[kF];
    double a_col[kF], k1d[kF];
    const double ced = (double)q.ce;
#pragma unroll
    for (int k = 0; k < kF; ++k) {
        s_s[k] = __fdiv_rn((float)(xoff + k), 5.0f);                                   // x[i] / f, a float division (:341)
        k1d[k] = mul_rn((double)s_s[k], add_rn(1.0, -(double)s_s[k]));
        a_col[k] = add_rn((double)s_s[k], mul_rn(k1d[k], ced));
    }

    // ---- pixels: row by row, staged in shared memory so that the warp stores contiguous spans ----
    __shared__ __align__(16) float s_lon_all[kThreads5 / 32][kSpanMax];
    __shared__ __align__(16) float s_lat_all[kThreads5 / 32][kSpanMax];
    float* const s_lon = s_lon_all[threadIdx.x >> 5];
    float* const s_lat = s_lat_all[threadIdx.x >> 5];
    const int i_start = first_warp ? 0 : 2 + kF * kQuadsPerWarp * w;
    const int i_end = last_warp ? kFineCols : 2 + kF * kQuadsPerWarp * (w + 1);
    const int span2 = (i_end - i_start) >> 1;
    float* const my_lon = s_lon + ((2 + kF * src_quad + xoff) - i_start);
    float* const my_lat = s_lat + ((2 + kF * src_quad + xoff) - i_start);
    const float2* const span_lon = reinterpret_cast<const float2*>(s_lon);
    const float2* const span_lat = reinterpret_cast<const float2*>(s_lat);
    const long long row0 = scan * (long long)(kRows * kFineCols) + i_start;

    // pass 0: every lane on its own path; pass 1 (rare): the whole warp again with the literal arithmetic when a
    // value of any of its quads sat on a rounding boundary (the stores are idempotent)
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        P2 acc = p2_dup(0.0f);
        float* dst_lon = out_lon + row0;
        float* dst_lat = out_lat + row0;
#pragma unroll 1
        for (int j = 0; j < kRows; ++j) {
            const float s_t = __fdiv_rn((float)(j - 2), 5.0f);                         // y[j] / f (:342)
            if (path == kEftLow) stage_row<kEftLow>(q, qe, s_s, a_col, k1d, s_t, kmask, acc, my_lon, my_lat);
            else if (path == kEftHigh) stage_row<kEftHigh>(q, qe, s_s, a_col, k1d, s_t, kmask, acc, my_lon, my_lat);
            else if (path == kFaithful) stage_row<kFaithful>(q, qe, s_s, a_col, k1d, s_t, kmask, acc, my_lon, my_lat);
            else if (path == kLibrary) stage_row<kLibrary>(q, qe, s_s, a_col, k1d, s_t, kmask, acc, my_lon, my_lat);
            copy_row_out(span_lon, span_lat, span2, lane, dst_lon, dst_lat);
            dst_lon += kFineCols; dst_lat += kFineCols;
        }
        const F2 fl = p2_get(acc);
        const bool flagged = !(fl.x == 0.0f) || !(fl.y == 0.0f);
        if (!__any_sync(kFull, flagged)) break;
        if (path == kEftLow || path == kEftHigh) path = kFaithful;
    }
}

}  // namespace

bool gtp_cviirs_fast5_f32_supported(const CviirsCfg& cfg)
{
    return cfg.km == 5 && cfg.p == 1 && (cfg.W == 271 || cfg.W == 270);
}

cudaError_t gtp_launch_cviirs_fast5_f32(const float* lon, const float* lat, const float* satz, float* out_lon, float* out_lat,
                                        long long n_scans, const CviirsCfg& cfg, cudaStream_t stream)
{
    if (n_scans <= 0) return cudaSuccess;
    // 64-bit stores need 8-byte aligned outputs (the row pitch, 5416 B, and every span start already are)
    if ((((uintptr_t)out_lon | (uintptr_t)out_lat) & 7) != 0)
        return gtp_launch_cviirs_generic<float>(lon, lat, satz, out_lon, out_lat, n_scans, cfg, stream, true);
    const int warps_per_scan = (cfg.W - 1 + kQuadsPerWarp - 1) / kQuadsPerWarp;
    const long long blocks = (n_scans * warps_per_scan * 32 + kThreads5 - 1) / kThreads5;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    cviirs_fast_5km_f32_kernel<<<(unsigned)blocks, kThreads5, 0, stream>>>(lon, lat, satz, out_lon, out_lat, n_scans, cfg.W, warps_per_scan);
    return cudaGetLastError();
}
