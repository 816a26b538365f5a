// gtp_fast.cu - float64 CVIIRS 1 km -> 250 m / 500 m fast path for sm_100a.
//
// One THREAD owns one tie-point quad column of one scan and walks down the ten tie
// rows of the scan; the neighbouring column's tie-point data comes from the next lane
// by warp shuffle; shared memory only holds one private slot per thread for the next tie
// row's inputs (cp.async double buffering), so there is no barrier.
// For every quad the thread projects the quad on the local frame of its corner A once
// (gtp_fast_math.cuh) and then produces the quad's P x (P | P + P/2) fine pixels with
// ~16 FP64 FMAs each; a warp therefore writes 31 x 32 B = 992 contiguous bytes per
// fine row and output array with one 256-bit store per lane (st.global.v4.f64).
//
// HBM traffic per output pixel (1 km -> 250 m): 16 B written + 3 x 8 B / 16 read =
// 17.5 B, the algorithmic minimum; the halo column a warp shares with its neighbour
// is an L2 hit.  Warps cover 31 quad columns (lane 31 only provides the right
// neighbour's tie points), 44 warps per scan; the thread of tie column 1353 produces
// the P extra pixels right of the last quad (x = P .. 2P-1, extrapolation).
#include <type_traits>
#include <cuda_runtime.h>
#include "gtp_kernels.h"
#include "gtp_fast_math.cuh"
#include "gtp_fast_io.cuh"

namespace {

using namespace gtpfast;

constexpr int kW = 1354;            // tie columns at 1 km
constexpr int kL = 10;              // tie rows per scan
constexpr int kQuadColsPerWarp = 31;
constexpr int kWarpsPerScan = (kW + kQuadColsPerWarp - 1) / kQuadColsPerWarp;   // 44
// CTA size, 16 warps per SM either way (measured): one-warp CTAs free their slot as soon as their strip is
// done, +2 ... 5 % for the short strips of 500 m and +3 % for the MOD03 types; the simple interpolator at 250 m
// is 2.5 % faster with four warps per CTA
template <int P, bool SIMPLE>
struct FastThreads { static constexpr int value = (SIMPLE && P == 4) ? 128 : 32; };
constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarpsPerSm = 16;      // resident warps per SM the register allocation is tuned for (128 registers)

// the fine rows [0, nrows) of one quad, first row at a_track = s_t0, branch-free per pixel
template <int P, int MODE, bool WRAP, typename TOut>
__device__ __forceinline__ void quad_rows(const Quad& q, const double (&a_col)[P], double s_t0, int nrows,
                                          TOut* out_lon, TOut* out_lat, long long out_idx)
{
    // one element index for both outputs (2 registers) instead of two running pointers (4)
    constexpr long long kPitch = (long long)kW * P;
    double s_t = s_t0;
#pragma unroll 1
    for (int m = 0; m < nrows; ++m) {
        const RowCoef rc = row_setup(q, s_t, s_t * (1.0 - s_t));
        double vlon[P], vlat[P];
#pragma unroll
        for (int k = 0; k < P; ++k) {
            if (MODE == kWide) pixel_eval_wide(q, rc, a_col[k], vlon[k], vlat[k]);
            else pixel_eval<MODE, WRAP>(q, rc, a_col[k], vlon[k], vlat[k]);
        }
        store_px<P>(out_lon + out_idx, vlon);
        store_px<P>(out_lat + out_idx, vlat);
        out_idx += kPitch;
        s_t += 1.0 / P;              // exact: multiples of 1/8
    }
}

// The rarely taken evaluation modes (quads straddling |lat| = 53.13 deg or the antimeridian, polar
// quads).  Inlined on purpose: as a real (noinline) function this cost 25 % of the whole kernel even
// on granules that never call it (the caller's register allocation around the call), and ptxas 12.9
// turned the 256-bit vector store inside the non-entry function into a single 64-bit store - only
// pixel 0 of 4 was written, caught by the antimeridian parity test (gpurun_out/per_granule.txt, ab10.txt).
template <int P, typename TOut>
__device__ __forceinline__ void quad_rows_cold(const Quad* qp, const double* a_colp, double s_t0, int nrows,
                                            TOut* out_lon, TOut* out_lat, long long out_idx)
{
    const Quad q = *qp;
    double a_col[P];
#pragma unroll
    for (int k = 0; k < P; ++k) a_col[k] = a_colp[k];
    if (q.mode == kWide) quad_rows<P, kWide, true>(q, a_col, s_t0, nrows, out_lon, out_lat, out_idx);
    else if (q.wrap) quad_rows<P, kBoth, true>(q, a_col, s_t0, nrows, out_lon, out_lat, out_idx);
    else quad_rows<P, kBoth, false>(q, a_col, s_t0, nrows, out_lon, out_lat, out_idx);
}

// SIMPLE = true is simple_modis_interpolator: the same bilinear form with a_scan = s_s and
// a_track = s_t, i.e. c_expansion = c_alignment = 0 and no zenith-angle input.  In exact
// arithmetic the reference's order-1 map_coordinates + linear extrapolation of the last 3 (1)
// columns and first/last 2 (1) rows (_simple_modis_interpolator.pyx:85-225) IS that bilinear
// form evaluated at t_y = j/res - (res/16 + 1/8) - r, t_x = i/res - c: the
// fine-row / column fractions coincide with the CVIIRS ones of the same resolution.
//
// TIn / TZen / TOut: the element types of lon & lat, of the zenith angle and of the outputs.  The
// reference's interface is all-float64 (double, double, double); the other instantiations serve
// gtp_cviirs_interp_mixed: float32 lon/lat, float32 or int16 * zen_scale + zen_offset
// zenith angles, float64 MATH throughout, results rounded once to float32 or kept in float64.
template <int P, bool SIMPLE, typename TIn, typename TZen, typename TOut>
__global__ void __launch_bounds__(FastThreads<P, SIMPLE>::value, kWarpsPerSm * 32 / FastThreads<P, SIMPLE>::value)
modis_fast_1km_f64_kernel(const TIn* __restrict__ lon, const TIn* __restrict__ lat,
                          const TZen* __restrict__ satz, TOut* __restrict__ out_lon,
                          TOut* __restrict__ out_lat, long long n_scans, double zen_scale, double zen_offset)
{
    constexpr bool kZenDecode = !std::is_same<TZen, double>::value;      // float64 zenith angles are taken as they are
    constexpr int kWf = kW * P;          // fine columns
    constexpr int kN = kL * P;           // fine rows per scan
    constexpr int kHalf = P / 2;

    constexpr int kThreads = FastThreads<P, SIMPLE>::value;
    const long long warp_global = (blockIdx.x * (long long)kThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long scan = warp_global / kWarpsPerScan;
    if (scan >= n_scans) return;
    const int w = (int)(warp_global - scan * kWarpsPerScan);
    const int c = w * kQuadColsPerWarp + lane;          // tie column this lane computes
    const int tc = min(c, kW - 1);
    const bool last_warp = (w == kWarpsPerScan - 1);
    const bool extra = (c == kW - 1);                   // also owns the pixels right of the last quad
    const bool quad_lane = (lane < kQuadColsPerWarp) && (c <= kW - 2);
    const bool active = quad_lane || extra;
    const int quad_col = extra ? kW - 2 : c;
    const int xoff = extra ? P : 0;

    // element indices instead of five running pointers: the bases are kernel parameters (uniform)
    long long tie_idx = scan * (long long)(kL * kW) + tc;
    long long out_idx = scan * (long long)kN * kWf + (long long)quad_col * P + xoff;

    // carried from the tie row above: this lane's tie point (frame + anchor; the next row's sines
    // and cosines are rotated out of it), corner B, and the quad row's expansion / alignment
    // coefficients (they depend on the upper corners only, :307-308)
    TieGeo top = {0.0, 1.0, 0.0, 1.0, 0.0, 0.0};
    TieZen zen = {0.0, 0.0, 1.0, 0.0, 1.0};
    Vec3 topB = {0.0, 0.0, 0.0};
    double top_ce = 0.0, top_ca = 0.0;
    __shared__ double stage[2][3][kThreads];     // double-buffered next-row inputs, one slot per thread
    constexpr bool kPark = (P == 4);
    __shared__ double park[kPark ? 14 : 1][kThreads];      // per-thread parking slots (see the row loops)
    stage_fetch(&stage[0][0][threadIdx.x], lon, tie_idx);
    stage_fetch(&stage[0][1][threadIdx.x], lat, tie_idx);
    if (!SIMPLE) stage_fetch(&stage[0][2][threadIdx.x], satz, tie_idx);
    cp_async_commit();

#pragma unroll 1
    for (int r = 0; r < kL; ++r) {
        cp_async_wait_all();
        const double clon = stage_value(&stage[r & 1][0][threadIdx.x], lon, tie_idx);
        const double clat = stage_value(&stage[r & 1][1][threadIdx.x], lat, tie_idx);
        double csatz = SIMPLE ? 0.0 : stage_value(&stage[r & 1][2][threadIdx.x], satz, tie_idx);
        if (!SIMPLE && kZenDecode) csatz = fma(csatz, zen_scale, zen_offset);
        if (r + 1 < kL) {            // fetch the next tie row while this one is worked on
            tie_idx += kW;           // even: an int16's half of its staged pair is the same in every row
            stage_fetch(&stage[(r + 1) & 1][0][threadIdx.x], lon, tie_idx);
            stage_fetch(&stage[(r + 1) & 1][1][threadIdx.x], lat, tie_idx);
            if (!SIMPLE) stage_fetch(&stage[(r + 1) & 1][2][threadIdx.x], satz, tie_idx);
            cp_async_commit();
        }
        // this lane's tie point of row r; corner B of the lane's quad column comes from lane + 1
        TieGeo own = top;
        tie_advance<!SIMPLE>(own, zen, clon, clat, csatz, r == 0);
        const Vec3 own_xyz = tie_xyz(own);
        Vec3 curB;
        curB.x = __shfl_down_sync(kFull, own_xyz.x, 1);
        curB.y = __shfl_down_sync(kFull, own_xyz.y, 1);
        curB.z = __shfl_down_sync(kFull, own_xyz.z, 1);
        TieGeo A = top;                  // corner A of the quad row above this tie row
        Vec3 D = own_xyz;
        if (last_warp) {                 // warp-uniform: the lane of tie column 1353 evaluates the quad of its left neighbour
            TieGeo up;
            up.sl = __shfl_up_sync(kFull, top.sl, 1); up.cl = __shfl_up_sync(kFull, top.cl, 1);
            up.sp = __shfl_up_sync(kFull, top.sp, 1); up.cp = __shfl_up_sync(kFull, top.cp, 1);
            up.lon = __shfl_up_sync(kFull, top.lon, 1); up.lat = __shfl_up_sync(kFull, top.lat, 1);
            Vec3 upD;
            upD.x = __shfl_up_sync(kFull, own_xyz.x, 1);
            upD.y = __shfl_up_sync(kFull, own_xyz.y, 1);
            upD.z = __shfl_up_sync(kFull, own_xyz.z, 1);
            if (extra) { A = up; D = upD; curB = own_xyz; }
        }

        if (r > 0 && active) {
            // quad row qr = r - 1 covers the fine rows [j0, j0 + nrows) of the scan (a6); its first
            // row sits at y = 0.5 (a4), except for the top quad row whose first h rows extrapolate
            const int qr = r - 1;
            const int nrows = P + ((qr == 0 || qr == kL - 2) ? kHalf : 0);
            const double s_t0 = ((qr == 0) ? (0.5 - kHalf) : 0.5) / P;
            const Vec3 A_xyz = tie_xyz(A);
            Quad q;
            quad_setup(q, A, A_xyz, topB, curB, D, top_ce, top_ca, extra ? (double)kABoundExtra : 1.0);
            if (q.mode != kSlow) {
                // State that is only needed again after the row loops (this tie row's frame, corner B, the
                // zenith chain: 14 doubles) waits in shared memory meanwhile: with those 28 registers the
                // compiler interleaves the 4 pixel chains of a fine row instead of evaluating them one
                // after the other (ncu: "wait" was 47 % of the stall samples in the row loops; +2..6 % at
                // 250 m).  Not at 500 m: 28 shared-memory accesses per quad cost more than they buy there.
                volatile double* const pk = &park[0][threadIdx.x];
                if (kPark) {
                    pk[0 * kThreads] = own.sl; pk[1 * kThreads] = own.cl; pk[2 * kThreads] = own.sp; pk[3 * kThreads] = own.cp;
                    pk[4 * kThreads] = own.lon; pk[5 * kThreads] = own.lat;
                    pk[6 * kThreads] = curB.x; pk[7 * kThreads] = curB.y; pk[8 * kThreads] = curB.z;
                    if (!SIMPLE) {
                        pk[9 * kThreads] = zen.za; pk[10 * kThreads] = zen.sz; pk[11 * kThreads] = zen.cz;
                        pk[12 * kThreads] = zen.sphi; pk[13 * kThreads] = zen.cphi;
                    }
                }
                // s_s + s_s (1 - s_s) c_exp  (:341, :344); s_s = k / P inside the quad, 1 + k / P right of it
                double a_col[P];
                double s_base = extra ? 1.0 : 0.0;
                // SIMPLE: a_col = s_s does not depend on the quad; keep the compiler from hoisting the P
                // values out of the tie-row loop, where they would cost 2 P registers through every loop
                if (SIMPLE) asm volatile("" : "+d"(s_base));
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    const double s_s = s_base + (double)k / (double)P;
                    a_col[k] = SIMPLE ? s_s : fma(s_s * (1.0 - s_s), q.ce, s_s);
                }
                if (!q.wrap && q.small) quad_rows<P, kLowSmall, false>(q, a_col, s_t0, nrows, out_lon, out_lat, out_idx);
                else if (!q.wrap && q.mode == kLow) quad_rows<P, kLow, false>(q, a_col, s_t0, nrows, out_lon, out_lat, out_idx);
                else if (!q.wrap && q.mode == kHigh) quad_rows<P, kHigh, false>(q, a_col, s_t0, nrows, out_lon, out_lat, out_idx);
                else {
                    const Quad qc = q;
                    double a_colc[P];
#pragma unroll
                    for (int k = 0; k < P; ++k) a_colc[k] = a_col[k];
                    quad_rows_cold<P>(&qc, a_colc, s_t0, nrows, out_lon, out_lat, out_idx);
                }
                if (kPark) {
                    own.sl = pk[0 * kThreads]; own.cl = pk[1 * kThreads]; own.sp = pk[2 * kThreads]; own.cp = pk[3 * kThreads];
                    own.lon = pk[4 * kThreads]; own.lat = pk[5 * kThreads];
                    curB.x = pk[6 * kThreads]; curB.y = pk[7 * kThreads]; curB.z = pk[8 * kThreads];
                    if (!SIMPLE) {
                        zen.za = pk[9 * kThreads]; zen.sz = pk[10 * kThreads]; zen.cz = pk[11 * kThreads];
                        zen.sphi = pk[12 * kThreads]; zen.cphi = pk[13 * kThreads];
                    }
                }
            } else {
#pragma unroll 1
                for (int m = 0; m < nrows; ++m) {
                    const double s_t = s_t0 + m * (1.0 / P);
                    double vlon[P], vlat[P];
#pragma unroll 1
                    for (int k = 0; k < P; ++k) {
                        const LonLat ll = pixel_eval_slow(A_xyz, topB, curB, D, q.ce, q.ca,
                                                          (double)(xoff + k) / (double)P, s_t);
                        vlon[k] = ll.lon; vlat[k] = ll.lat;
                    }
                    store_px<P>(out_lon + out_idx + (long long)m * kWf, vlon);
                    store_px<P>(out_lat + out_idx + (long long)m * kWf, vlat);
                }
            }
            out_idx += (long long)nrows * kWf;
        }
        top = own;
        topB = curB;

        if (!SIMPLE && r + 1 < kL) {
            // expansion / alignment of the quad row BELOW this tie row, from this row's zenith angles
            double zb = __shfl_down_sync(kFull, zen.za, 1);
            double sphi_b = __shfl_down_sync(kFull, zen.sphi, 1);
            double cphi_b = __shfl_down_sync(kFull, zen.cphi, 1);
            TieZen za = zen;
            if (last_warp) {
                TieZen up;
                up.za = __shfl_up_sync(kFull, zen.za, 1);
                up.sz = __shfl_up_sync(kFull, zen.sz, 1); up.cz = __shfl_up_sync(kFull, zen.cz, 1);
                up.sphi = __shfl_up_sync(kFull, zen.sphi, 1); up.cphi = __shfl_up_sync(kFull, zen.cphi, 1);
                if (extra) { za = up; zb = zen.za; sphi_b = zen.sphi; cphi_b = zen.cphi; }
            }
            // lanes without a quad (lane 31, clamped columns) would see two equal zenith angles
            if (active) quad_coefficients(za, zb, sphi_b, cphi_b, 1, top_ce, top_ca);
        }
    }
}

}  // namespace

bool gtp_cviirs_fast_f64_supported(const CviirsCfg& cfg)
{
    return cfg.km == 1 && (cfg.p == 4 || cfg.p == 2);
}

bool gtp_simple_fast_f64_supported(int width, int res)
{
    return width == kW && (res == 4 || res == 2);
}

namespace {

// 256-bit (P = 4, float64), 128-bit (P = 2 float64 | P = 4 float32) or 64-bit (P = 2 float32) stores
// need aligned outputs; the row pitch already is
template <typename TOut>
bool outputs_aligned(const TOut* out_lon, const TOut* out_lat, int p)
{
    const uintptr_t align = (uintptr_t)p * sizeof(TOut);
    return (((uintptr_t)out_lon | (uintptr_t)out_lat) & (align - 1)) == 0;
}

template <int P, bool SIMPLE, typename TIn, typename TZen, typename TOut>
cudaError_t launch_fast(const TIn* lon, const TIn* lat, const TZen* satz, TOut* out_lon, TOut* out_lat,
                        long long n_scans, double zen_scale, double zen_offset, cudaStream_t stream)
{
    const long long warps = n_scans * kWarpsPerScan;
    constexpr int kThreads = FastThreads<P, SIMPLE>::value;
    const long long blocks = (warps * 32 + kThreads - 1) / kThreads;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    modis_fast_1km_f64_kernel<P, SIMPLE, TIn, TZen, TOut><<<(unsigned)blocks, kThreads, 0, stream>>>(
        lon, lat, satz, out_lon, out_lat, n_scans, zen_scale, zen_offset);
    return cudaGetLastError();
}

template <typename TZen, typename TOut>
cudaError_t launch_mixed(const float* lon, const float* lat, const TZen* satz, double zen_scale, double zen_offset,
                         TOut* out_lon, TOut* out_lat, long long n_scans, int p, cudaStream_t stream)
{
    if (!outputs_aligned(out_lon, out_lat, p)) return cudaErrorMisalignedAddress;
    // an int16 zenith angle is fetched as the aligned 4-byte pair that holds it
    if (((uintptr_t)lon | (uintptr_t)lat | (uintptr_t)satz) & 3) return cudaErrorMisalignedAddress;
    return p == 4 ? launch_fast<4, false>(lon, lat, satz, out_lon, out_lat, n_scans, zen_scale, zen_offset, stream)
                  : launch_fast<2, false>(lon, lat, satz, out_lon, out_lat, n_scans, zen_scale, zen_offset, stream);
}

}  // namespace

cudaError_t gtp_launch_cviirs_fast_f64(const double* lon, const double* lat, const double* satz,
                                       double* out_lon, double* out_lat,
                                       long long n_scans, const CviirsCfg& cfg, cudaStream_t stream)
{
    if (n_scans <= 0) return cudaSuccess;
    if (!outputs_aligned(out_lon, out_lat, cfg.p))
        return gtp_launch_cviirs_generic<double>(lon, lat, satz, out_lon, out_lat, n_scans, cfg, stream);
    return cfg.p == 4 ? launch_fast<4, false>(lon, lat, satz, out_lon, out_lat, n_scans, 1.0, 0.0, stream)
                      : launch_fast<2, false>(lon, lat, satz, out_lon, out_lat, n_scans, 1.0, 0.0, stream);
}

cudaError_t gtp_launch_simple_fast_f64(const double* lon, const double* lat, double* out_lon, double* out_lat,
                                       long long n_scans, int res, cudaStream_t stream)
{
    if (n_scans <= 0) return cudaSuccess;
    if (!outputs_aligned(out_lon, out_lat, res))
        return gtp_launch_simple_generic<double>(lon, lat, out_lon, out_lat, n_scans, kW, res, stream);
    return res == 4 ? launch_fast<4, true>(lon, lat, (const double*)nullptr, out_lon, out_lat, n_scans, 1.0, 0.0, stream)
                    : launch_fast<2, true>(lon, lat, (const double*)nullptr, out_lon, out_lat, n_scans, 1.0, 0.0, stream);
}

// simple_modis_interpolator on float32 lon / lat: float64 math, float32 | float64 results (width 1354)
cudaError_t gtp_launch_simple_mixed(const float* lon, const float* lat, void* out_lon, void* out_lat, int out_is_f64,
                                    long long n_scans, int res, cudaStream_t stream)
{
    if (n_scans <= 0) return cudaSuccess;
    if (res != 4 && res != 2) return cudaErrorInvalidValue;
    if (((uintptr_t)lon | (uintptr_t)lat) & 3) return cudaErrorMisalignedAddress;
    const double* none = nullptr;
    if (out_is_f64) {
        double* olon = (double*)out_lon; double* olat = (double*)out_lat;
        if (!outputs_aligned(olon, olat, res)) return cudaErrorMisalignedAddress;
        return res == 4 ? launch_fast<4, true>(lon, lat, none, olon, olat, n_scans, 1.0, 0.0, stream)
                        : launch_fast<2, true>(lon, lat, none, olon, olat, n_scans, 1.0, 0.0, stream);
    }
    float* olon = (float*)out_lon; float* olat = (float*)out_lat;
    if (!outputs_aligned(olon, olat, res)) return cudaErrorMisalignedAddress;
    return res == 4 ? launch_fast<4, true>(lon, lat, none, olon, olat, n_scans, 1.0, 0.0, stream)
                    : launch_fast<2, true>(lon, lat, none, olon, olat, n_scans, 1.0, 0.0, stream);
}

// float32 lon / lat, float32 (zen_is_i16 == 0) or int16 zenith angles decoded as v * zen_scale +
// zen_offset, float64 math, float32 (out_is_f64 == 0) or float64 results; 1 km tie points only
cudaError_t gtp_launch_cviirs_mixed(const float* lon, const float* lat, const void* satz, int zen_is_i16,
                                    double zen_scale, double zen_offset, void* out_lon, void* out_lat, int out_is_f64,
                                    long long n_scans, int p, cudaStream_t stream)
{
    if (n_scans <= 0) return cudaSuccess;
    if (p != 4 && p != 2) return cudaErrorInvalidValue;
    if (zen_is_i16) {
        const short* z = (const short*)satz;
        return out_is_f64 ? launch_mixed(lon, lat, z, zen_scale, zen_offset, (double*)out_lon, (double*)out_lat, n_scans, p, stream)
                          : launch_mixed(lon, lat, z, zen_scale, zen_offset, (float*)out_lon, (float*)out_lat, n_scans, p, stream);
    }
    const float* z = (const float*)satz;
    return out_is_f64 ? launch_mixed(lon, lat, z, zen_scale, zen_offset, (double*)out_lon, (double*)out_lat, n_scans, p, stream)
                      : launch_mixed(lon, lat, z, zen_scale, zen_offset, (float*)out_lon, (float*)out_lat, n_scans, p, stream);
}
