// gtp_fast5.cu - float64 CVIIRS 5 km -> 1 km fast path for sm_100a.
//
// modis_5km_to_1km (modisinterpolator.py:37-41 -> _modis_interpolator.pyx:17-38 with
// coarse_resolution = 5000, fine_resolution = 1000): a scan has 2 tie rows of 271 (or 270) tie
// points and becomes 10 x 1354 fine pixels; every tie-point quad covers 5 fine columns x all 10
// fine rows (a_track = (j - 2) / 5 = -0.4 ... 1.4), the first quad also the 2 columns left of it
// (x = -2, -1), the last one the 2 (W = 271) or 7 (W = 270) columns right of it (x = 5 ... 11)
// (_get_coords_5km :231-249, _expand_tiepoint_array_5km :502-593).
//
// One THREAD owns one quad: it evaluates its two tie points (the lower one by rotating the upper
// one's sines / cosines), takes the right-hand corners from lane + 1 by shuffle, sets the quad up
// once (gtp_fast_math.cuh: frame projection, expansion / alignment, 8-term series) and produces its
// 50 pixels with ~30 FP64 instructions each; a warp stages every fine row in shared memory and
// stores it as one contiguous span.  The edge columns are evaluated by otherwise idle
// lanes ("virtual" lanes) that receive the finished quad of lane 0 / of the last quad by shuffle, so
// no warp waits for one thread that has 12 columns instead of 5.
//
// HBM traffic: 16 B written per pixel + 3 x 8 B per tie point = 16.96 B/px.
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>
#include "gtp_kernels.h"
#include "gtp_fast_math.cuh"

namespace {

using namespace gtpfast;

constexpr int kFineCols = 1354;
constexpr int kF = 5;                // fine pixels per coarse pixel (5 km -> 1 km)
constexpr int kRows = 10;            // fine rows per scan
constexpr int kQuadsPerWarp = 30;    // 150 fine columns: every warp's span starts on an even column (16 B);
                                     // lane 30 provides the right-hand corners of lane 29
constexpr int kThreads5 = 32;     // one warp = one strip: a CTA slot is free again as soon as its strip is done
constexpr unsigned kFull = 0xffffffffu;

// s_t = y / f, y = j - 2 (:231-249, :341-343) and s_s = x / f for x = -5 ... 11 (index x + 5)
__constant__ double kTrack5[kRows] = {-2.0 / 5.0, -1.0 / 5.0, 0.0 / 5.0, 1.0 / 5.0, 2.0 / 5.0,
                                      3.0 / 5.0, 4.0 / 5.0, 5.0 / 5.0, 6.0 / 5.0, 7.0 / 5.0};
__constant__ double kScan5[17] = {-5.0 / 5.0, -4.0 / 5.0, -3.0 / 5.0, -2.0 / 5.0, -1.0 / 5.0, 0.0 / 5.0,
                                  1.0 / 5.0, 2.0 / 5.0, 3.0 / 5.0, 4.0 / 5.0, 5.0 / 5.0, 6.0 / 5.0,
                                  7.0 / 5.0, 8.0 / 5.0, 9.0 / 5.0, 10.0 / 5.0, 11.0 / 5.0};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(kFull, v, src); }

// the four Cartesian corners of a quad: only the rounding-faithful fallback needs them after the set-up,
// so they wait in shared memory (12 doubles per lane) instead of in 24 registers through the row loops
struct QuadCorners { Vec3 A, B, C, D; };

__device__ __forceinline__ void take_quad_from(Quad& q, int src, bool take)
{
#define GTP_TAKE(f) { const double v_ = shfl_d(f, src); if (take) f = v_; }
    { const int m_ = __shfl_sync(kFull, q.mode, src); if (take) q.mode = m_; }
    { const int w_ = __shfl_sync(kFull, (int)q.wrap, src); if (take) q.wrap = (w_ != 0); }
    { const int s_ = __shfl_sync(kFull, (int)q.small, src); if (take) q.small = (s_ != 0); }
    GTP_TAKE(q.ce) GTP_TAKE(q.ca)
    GTP_TAKE(q.Pe) GTP_TAKE(q.Ph) GTP_TAKE(q.Pw) GTP_TAKE(q.Qe) GTP_TAKE(q.Qh) GTP_TAKE(q.Qw)
    GTP_TAKE(q.Se) GTP_TAKE(q.Sh) GTP_TAKE(q.Sw)
    GTP_TAKE(q.lon0) GTP_TAKE(q.lat0) GTP_TAKE(q.s0)
    GTP_TAKE(q.a1) GTP_TAKE(q.a2) GTP_TAKE(q.a3) GTP_TAKE(q.a4) GTP_TAKE(q.a5) GTP_TAKE(q.a6) GTP_TAKE(q.a7) GTP_TAKE(q.a8)
    GTP_TAKE(q.b1) GTP_TAKE(q.b2) GTP_TAKE(q.b3) GTP_TAKE(q.b4) GTP_TAKE(q.b5) GTP_TAKE(q.b6) GTP_TAKE(q.b7) GTP_TAKE(q.b8)
#undef GTP_TAKE
}

constexpr int kSpanMax = 152;        // fine columns one warp produces per row: 30 x 5 (+ 2 at either swath edge)

// store two adjacent pixels of the float64 result: as they are, or rounded once to float32
__device__ __forceinline__ void store_pair(double* dst, int pair, double2 v) { __stcs(reinterpret_cast<double2*>(dst) + pair, v); }
__device__ __forceinline__ void store_pair(float* dst, int pair, double2 v)
{
    __stcs(reinterpret_cast<float2*>(dst) + pair, make_float2((float)v.x, (float)v.y));
}

// One fine row of a lane's quad (kF pixels) into the warp's staging buffer.  Every evaluation mode writes
// its pixels itself (arrays merged across the modes would live in local memory); lanes of the inner warps
// own all 5 columns, so only the swath-edge warps mask.  SHORT: the 5-term forms of the 1 km path (Quad::small).
template <int MODE, bool MASKED, bool SHORT>
__device__ __forceinline__ void stage_row(const Quad& q, const RowCoef& rc, const double (&a_col)[kF], unsigned kmask,
                                          double* my_lon, double* my_lat)
{
#pragma unroll
    for (int k = 0; k < kF; ++k) {
        double lo, la;
        if (SHORT) pixel_eval<MODE, false>(q, rc, a_col[k], lo, la);
        else pixel_eval_long<MODE>(q, rc, a_col[k], lo, la);
        if (!MASKED || ((kmask >> k) & 1u)) { my_lon[k] = lo; my_lat[k] = la; }
    }
}

// The warp's staged row -> global memory: 150 | 152 columns = 75 | 76 pairs, two full rounds of 32 lanes
// and one of 11 | 12.
template <typename TOut>
__device__ __forceinline__ void copy_row_out(const double2* span_lon, const double2* span_lat, int span2, int lane,
                                             TOut*& dst_lon, TOut*& dst_lat, int pitch)
{
    __syncwarp();
    store_pair(dst_lon, lane, span_lon[lane]);
    store_pair(dst_lat, lane, span_lat[lane]);
    store_pair(dst_lon, lane + 32, span_lon[lane + 32]);
    store_pair(dst_lat, lane + 32, span_lat[lane + 32]);
    if (lane + 64 < span2) {
        store_pair(dst_lon, lane + 64, span_lon[lane + 64]);
        store_pair(dst_lat, lane + 64, span_lat[lane + 64]);
    }
    __syncwarp();
    dst_lon += pitch; dst_lat += pitch;
}

// TIn / TZen / TOut as in gtp_fast.cu: all-double is the reference's interface; float lon/lat, float
// or int16 * zen_scale + zen_offset zenith angles and float | double results serve gtp_cviirs_interp_mixed.
template <typename TIn, typename TZen, typename TOut>
__global__ void __launch_bounds__(kThreads5, 16)
modis_fast_5km_f64_kernel(const TIn* __restrict__ lon, const TIn* __restrict__ lat,
                          const TZen* __restrict__ satz, TOut* __restrict__ out_lon,
                          TOut* __restrict__ out_lat, long long n_scans, int W, int warps_per_scan,
                          double zen_scale, double zen_offset)
{
    const long long warp_global = (blockIdx.x * (long long)kThreads5 + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long scan = warp_global / warps_per_scan;
    if (scan >= n_scans) return;
    const int w = (int)(warp_global - scan * warps_per_scan);
    const int n_quads = W - 1;
    const int qc = w * kQuadsPerWarp + lane;            // tie column / quad column of this lane
    const int tc = min(qc, W - 1);
    const bool real = (lane < kQuadsPerWarp) && (qc < n_quads);
    const bool first_warp = (w == 0), last_warp = (w == warps_per_scan - 1);

    // ---- the two tie points of this lane's column -------------------------------------------
    const long long base = scan * 2 * (long long)W + tc;
    const double lon0 = (double)__ldg(lon + base), lat0 = (double)__ldg(lat + base);
    const double lon1 = (double)__ldg(lon + base + W), lat1 = (double)__ldg(lat + base + W);
    double satz0 = (double)__ldg(satz + base);
    if (!std::is_same<TZen, double>::value) satz0 = fma(satz0, zen_scale, zen_offset);
    TieGeo top = {0.0, 1.0, 0.0, 1.0, 0.0, 0.0};
    TieZen zen = {0.0, 0.0, 1.0, 0.0, 1.0}, unused = zen;
    tie_first<true>(top, zen, lon0, lat0, satz0);
    TieGeo bot = top;
    tie_next<false>(bot, unused, lon1, lat1, 0.0);
    QuadCorners x;
    x.A = tie_xyz(top);
    x.D = tie_xyz(bot);
    x.B.x = __shfl_down_sync(kFull, x.A.x, 1); x.B.y = __shfl_down_sync(kFull, x.A.y, 1); x.B.z = __shfl_down_sync(kFull, x.A.z, 1);
    x.C.x = __shfl_down_sync(kFull, x.D.x, 1); x.C.y = __shfl_down_sync(kFull, x.D.y, 1); x.C.z = __shfl_down_sync(kFull, x.D.z, 1);
    double zb = __shfl_down_sync(kFull, zen.za, 1);
    const double sphi_b = __shfl_down_sync(kFull, zen.sphi, 1);
    const double cphi_b = __shfl_down_sync(kFull, zen.cphi, 1);

    // ---- quad set-up (real lanes) ------------------------------------------------------------
    // a_scan reaches 0.8 + |c_exp|/4 + |c_ali| 0.56 inside the swath, 1.2 / 2.2 (+ 0.24 / 2.64 |c_exp|)
    // in the columns right of the last quad (W = 271 / 270); the last quad's bound covers those too
    const bool last_quad = real && (qc == n_quads - 1);
    const double a_max = !last_quad ? 1.0 : (W == 271 ? 1.4 : 3.0);
    Quad q;
    q.mode = kSlow; q.wrap = false; q.ce = q.ca = 0.0;
    if (!real) zb = zen.za + 1e-3;       // lanes without a quad would see two equal zenith angles
    double ce, ca;
    quad_coefficients(zen, zb, sphi_b, cphi_b, 5, ce, ca);
    {
        const QuadSize sz = quad_project(q, top, x.A, x.B, x.C, x.D, a_max);
        quad_classify_long(q, top, sz, ce, ca);
    }
    __shared__ double corners[12][kThreads5];
    {
        double* const cs = &corners[0][threadIdx.x];
        cs[0 * kThreads5] = x.A.x; cs[1 * kThreads5] = x.A.y; cs[2 * kThreads5] = x.A.z;
        cs[3 * kThreads5] = x.B.x; cs[4 * kThreads5] = x.B.y; cs[5 * kThreads5] = x.B.z;
        cs[6 * kThreads5] = x.C.x; cs[7 * kThreads5] = x.C.y; cs[8 * kThreads5] = x.C.z;
        cs[9 * kThreads5] = x.D.x; cs[10 * kThreads5] = x.D.y; cs[11 * kThreads5] = x.D.z;
    }
    __syncwarp();                             // virtual lanes read another lane's slot
    int corner_lane = threadIdx.x;            // whose corners this lane evaluates (virtual lanes: the source quad's)

    // ---- lane roles --------------------------------------------------------------------------
    // real lanes: columns x = 0..4 of their quad.  Virtual lanes (idle otherwise) take over the
    // partial columns: warp 0, lane 31: x = -2, -1 of quad 0; last warp, the lanes after the last
    // quad: x = 5..9 and x = 10, 11 of the last quad.
    int src_quad = qc, xoff = 0;
    unsigned kmask = real ? 0x1fu : 0u;
    if (first_warp || last_warp) {        // warp-uniform
        const int last_lane = (n_quads - 1) - w * kQuadsPerWarp;      // lane of the last quad (last warp)
        const int n_extra = (W == 271) ? 2 : 7;
        const bool v_first = first_warp && (lane == 31);
        const bool v_last1 = last_warp && (lane == last_lane + 1);
        const bool v_last2 = last_warp && (lane == last_lane + 2) && (n_extra > kF);
        if (first_warp) take_quad_from(q, 0, v_first);
        if (last_warp) take_quad_from(q, last_lane, v_last1 || v_last2);
        if (v_first) corner_lane = (threadIdx.x & ~31);
        if (v_last1 || v_last2) corner_lane = (threadIdx.x & ~31) + last_lane;
        if (v_first) { src_quad = 0; xoff = -kF; kmask = 0x18u; }                       // k = 3, 4 -> x = -2, -1
        if (v_last1) { src_quad = n_quads - 1; xoff = kF; kmask = (n_extra >= kF) ? 0x1fu : ((1u << n_extra) - 1u); }
        if (v_last2) { src_quad = n_quads - 1; xoff = 2 * kF; kmask = (1u << (n_extra - kF)) - 1u; }
    }
    // ---- pixels: row by row, staged in shared memory so that the warp stores contiguous spans ----
    // A lane's 5 columns are 40 B apart from its neighbour's: stored directly, every store instruction
    // would touch 10 cache lines with 8 B each (measured: 29 % of the roofline, FP64 pipe 21 % busy).
    // Instead the warp's pixels of one fine row (one contiguous span of 150 or 152 columns that starts
    // on an even column) go through a per-warp buffer and leave as 128-bit stores, 512 contiguous bytes
    // per instruction.
    __shared__ __align__(16) double s_lon[kThreads5 / 32][kSpanMax];
    __shared__ __align__(16) double s_lat[kThreads5 / 32][kSpanMax];
    const int i_start = first_warp ? 0 : 2 + kF * kQuadsPerWarp * w;
    const int i_end = last_warp ? kFineCols : 2 + kF * kQuadsPerWarp * (w + 1);
    const int span2 = (i_end - i_start) >> 1;                        // in 16-byte pairs
    double* const my_lon = s_lon[threadIdx.x >> 5] + ((2 + kF * src_quad + xoff) - i_start);   // slot of column k = 0
    double* const my_lat = s_lat[threadIdx.x >> 5] + ((2 + kF * src_quad + xoff) - i_start);
    const double2* const span_lon = reinterpret_cast<const double2*>(s_lon[threadIdx.x >> 5]);
    const double2* const span_lat = reinterpret_cast<const double2*>(s_lat[threadIdx.x >> 5]);
    const long long row0 = scan * (long long)(kRows * kFineCols) + i_start;

    double a_col[kF];
#pragma unroll
    for (int k = 0; k < kF; ++k) {
        const double s_s = kScan5[min(xoff + k + 5, 16)];           // masked columns: any valid entry
        a_col[k] = fma(s_s * (1.0 - s_s), q.ce, s_s);                // :341, :344
    }
    const int mode = (kmask == 0u) ? -1 : ((q.mode == kSlow) ? kSlow : ((q.wrap || q.mode == kBoth) ? kBoth : q.mode));

    const bool edge = first_warp || last_warp;                       // warp-uniform
    TOut* dst_lon = out_lon + row0;
    TOut* dst_lat = out_lat + row0;
    // Inner warps whose quads all evaluate the same way - nearly all of them - run a row loop without
    // the per-row dispatch (lanes 30 and 31 own no quad there and only help with the copy).
    const bool all_low = !edge && __all_sync(kFull, mode == kLow || mode < 0);
    const bool all_high = !edge && __all_sync(kFull, mode == kHigh || mode < 0);
    const bool all_short = __all_sync(kFull, q.small || mode < 0);
    if (all_low && all_short) {
#pragma unroll 1
        for (int j = 0; j < kRows; ++j) {
            if (mode >= 0) {
                const double s_t = kTrack5[j];
                const RowCoef rc = row_setup(q, s_t, s_t * (1.0 - s_t));
                stage_row<kLow, false, true>(q, rc, a_col, kmask, my_lon, my_lat);
            }
            copy_row_out(span_lon, span_lat, span2, lane, dst_lon, dst_lat, kFineCols);
        }
        return;
    }
    if (all_high && all_short) {
#pragma unroll 1
        for (int j = 0; j < kRows; ++j) {
            if (mode >= 0) {
                const double s_t = kTrack5[j];
                const RowCoef rc = row_setup(q, s_t, s_t * (1.0 - s_t));
                stage_row<kHigh, false, true>(q, rc, a_col, kmask, my_lon, my_lat);
            }
            copy_row_out(span_lon, span_lat, span2, lane, dst_lon, dst_lat, kFineCols);
        }
        return;
    }
    if (all_low) {
#pragma unroll 1
        for (int j = 0; j < kRows; ++j) {
            if (mode >= 0) {
                const double s_t = kTrack5[j];
                const RowCoef rc = row_setup(q, s_t, s_t * (1.0 - s_t));
                stage_row<kLow, false, false>(q, rc, a_col, kmask, my_lon, my_lat);
            }
            copy_row_out(span_lon, span_lat, span2, lane, dst_lon, dst_lat, kFineCols);
        }
        return;
    }
    if (all_high) {
#pragma unroll 1
        for (int j = 0; j < kRows; ++j) {
            if (mode >= 0) {
                const double s_t = kTrack5[j];
                const RowCoef rc = row_setup(q, s_t, s_t * (1.0 - s_t));
                stage_row<kHigh, false, false>(q, rc, a_col, kmask, my_lon, my_lat);
            }
            copy_row_out(span_lon, span_lat, span2, lane, dst_lon, dst_lat, kFineCols);
        }
        return;
    }
#pragma unroll 1
    for (int j = 0; j < kRows; ++j) {
        const double s_t = kTrack5[j];
        if (mode == kLow || mode == kHigh || mode == kBoth) {
            const RowCoef rc = row_setup(q, s_t, s_t * (1.0 - s_t));
            if (mode == kLow) stage_row<kLow, true, false>(q, rc, a_col, kmask, my_lon, my_lat);
            else if (mode == kHigh) stage_row<kHigh, true, false>(q, rc, a_col, kmask, my_lon, my_lat);
            else stage_row<kBoth, true, false>(q, rc, a_col, kmask, my_lon, my_lat);
        } else if (mode == kSlow) {
#pragma unroll 1
            for (int k = 0; k < kF; ++k) {
                if (!((kmask >> k) & 1u)) continue;
                const double* const cs = &corners[0][corner_lane];
                Vec3 cA, cB, cC, cD;
                cA.x = cs[0 * kThreads5]; cA.y = cs[1 * kThreads5]; cA.z = cs[2 * kThreads5];
                cB.x = cs[3 * kThreads5]; cB.y = cs[4 * kThreads5]; cB.z = cs[5 * kThreads5];
                cC.x = cs[6 * kThreads5]; cC.y = cs[7 * kThreads5]; cC.z = cs[8 * kThreads5];
                cD.x = cs[9 * kThreads5]; cD.y = cs[10 * kThreads5]; cD.z = cs[11 * kThreads5];
                const LonLat ll = pixel_eval_slow(cA, cB, cC, cD, q.ce, q.ca, kScan5[xoff + k + 5], s_t);
                my_lon[k] = ll.lon; my_lat[k] = ll.lat;
            }
        }
        copy_row_out(span_lon, span_lat, span2, lane, dst_lon, dst_lat, kFineCols);
    }
}

}  // namespace

bool gtp_cviirs_fast5_f64_supported(const CviirsCfg& cfg)
{
    return cfg.km == 5 && cfg.p == 1 && (cfg.W == 271 || cfg.W == 270);
}

namespace {

template <typename TIn, typename TZen, typename TOut>
cudaError_t launch_fast5(const TIn* lon, const TIn* lat, const TZen* satz, TOut* out_lon, TOut* out_lat,
                         long long n_scans, int W, double zen_scale, double zen_offset, cudaStream_t stream)
{
    // 9 warps of 30 quads: 270 (W = 271) or 269 (W = 270) quads; the last warp has lanes to spare for
    // the columns right of the last quad (lane 30 | lanes 29, 30), warp 0 uses lane 31 for the left ones
    const int warps_per_scan = (W - 1 + kQuadsPerWarp - 1) / kQuadsPerWarp;
    const long long warps = n_scans * warps_per_scan;
    const long long blocks = (warps * 32 + kThreads5 - 1) / kThreads5;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    modis_fast_5km_f64_kernel<TIn, TZen, TOut><<<(unsigned)blocks, kThreads5, 0, stream>>>(
        lon, lat, satz, out_lon, out_lat, n_scans, W, warps_per_scan, zen_scale, zen_offset);
    return cudaGetLastError();
}

}  // namespace

cudaError_t gtp_launch_cviirs_fast5_f64(const double* lon, const double* lat, const double* satz,
                                        double* out_lon, double* out_lat,
                                        long long n_scans, const CviirsCfg& cfg, cudaStream_t stream)
{
    if (n_scans <= 0) return cudaSuccess;
    // 128-bit stores need 16-byte aligned outputs (the row pitch and every span start already are)
    if ((((uintptr_t)out_lon | (uintptr_t)out_lat) & 15) != 0)
        return gtp_launch_cviirs_generic<double>(lon, lat, satz, out_lon, out_lat, n_scans, cfg, stream);
    return launch_fast5(lon, lat, satz, out_lon, out_lat, n_scans, cfg.W, 1.0, 0.0, stream);
}

// float32 lon / lat, float32 | int16 zenith angles, float64 math, float32 | float64 results (5 km -> 1 km)
cudaError_t gtp_launch_cviirs5_mixed(const float* lon, const float* lat, const void* satz, int zen_is_i16,
                                     double zen_scale, double zen_offset, void* out_lon, void* out_lat, int out_is_f64,
                                     long long n_scans, int W, cudaStream_t stream)
{
    if (n_scans <= 0) return cudaSuccess;
    if (W != 270 && W != 271) return cudaErrorInvalidValue;
    const uintptr_t out_align = out_is_f64 ? 15 : 7;
    if ((((uintptr_t)out_lon | (uintptr_t)out_lat) & out_align) || (((uintptr_t)lon | (uintptr_t)lat) & 3)
        || ((uintptr_t)satz & (zen_is_i16 ? 1 : 3)))
        return cudaErrorMisalignedAddress;
    if (zen_is_i16) {
        const short* z = (const short*)satz;
        return out_is_f64 ? launch_fast5(lon, lat, z, (double*)out_lon, (double*)out_lat, n_scans, W, zen_scale, zen_offset, stream)
                          : launch_fast5(lon, lat, z, (float*)out_lon, (float*)out_lat, n_scans, W, zen_scale, zen_offset, stream);
    }
    const float* z = (const float*)satz;
    return out_is_f64 ? launch_fast5(lon, lat, z, (double*)out_lon, (double*)out_lat, n_scans, W, zen_scale, zen_offset, stream)
                      : launch_fast5(lon, lat, z, (float*)out_lon, (float*)out_lat, n_scans, W, zen_scale, zen_offset, stream);
}
