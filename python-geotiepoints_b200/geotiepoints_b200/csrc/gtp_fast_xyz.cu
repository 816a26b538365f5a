// gtp_fast_xyz.cu - ECEF output of the 1 km interpolators for sm_100a.
//
// The fine Cartesian coordinates x, y, z (metres, sphere of radius 6370997 m) that the reference
// computes in _compute_fine_xyz (_modis_interpolator.pyx:346-398) - or new_xyz of the simple
// interpolator (_simple_modis_interpolator.pyx:52-63) - and then hands to xyz2lonlat (:293).  The
// step after geolocation (resampling: kd-tree or gradient search on ECEF points) converts lon / lat
// straight back to these, so the kernel stops here: no back-transform, three outputs.
//
// Same decomposition as gtp_fast.cu: one thread per tie-point quad column of one scan, ten tie rows
// walked from top to bottom, the neighbouring column's tie point by warp shuffle, the next tie row
// staged with cp.async into one private shared-memory slot per thread.  Per quad the bilinear form
// (:396-398) with a_track = s_t and a_scan = s_s + s_s (1 - s_s) c_exp + s_t (1 - s_t) c_ali (:344) is
//     v(a, t) = A + a P + t Q + a t S,  P = B - A, Q = D - A, S = (C - D) - P
// so a fine row costs U = P + t S, V = A + t Q + a_row U (a_row: the alignment term) and each pixel
// three FMAs, v = V + a_col U.  ~15 FP64 instructions per pixel: HBM-bound by a wide margin.
//
// HBM traffic per output pixel (1 km -> 250 m): float64 24 B written + 24 B / 16 read = 25.5 B;
// float32 results from MOD03's on-disk types (float32 lon / lat, int16 zenith): 12 + 10 / 16 = 12.6 B.
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
#ifndef GTP_XYZ_THREADS
#define GTP_XYZ_THREADS 128
#endif
#ifndef GTP_XYZ_MIN_BLOCKS
#define GTP_XYZ_MIN_BLOCKS 3     // measured (gpurun_out/ab_xyz.txt): 2-3 CTAs 92.7 %, 4: 90 %, 5: 87 %, 6: 79 %, 8: 67 % of the
                                 // HBM roofline - a pure store stream wants few resident warps
#endif
constexpr int kThreads = GTP_XYZ_THREADS;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ Vec3 shfl_down1(const Vec3& v)
{
    Vec3 r;
    r.x = __shfl_down_sync(kFull, v.x, 1); r.y = __shfl_down_sync(kFull, v.y, 1); r.z = __shfl_down_sync(kFull, v.z, 1);
    return r;
}

__device__ __forceinline__ Vec3 shfl_up1(const Vec3& v)
{
    Vec3 r;
    r.x = __shfl_up_sync(kFull, v.x, 1); r.y = __shfl_up_sync(kFull, v.y, 1); r.z = __shfl_up_sync(kFull, v.z, 1);
    return r;
}

// SIMPLE: c_expansion = c_alignment = 0, no zenith angles (see gtp_fast.cu).  TIn / TZen / TOut as in
// gtp_fast.cu: (double, double, double) is the reference's float64 layout, the others MOD03's types.
template <int P, bool SIMPLE, typename TIn, typename TZen, typename TOut>
__global__ void __launch_bounds__(kThreads, GTP_XYZ_MIN_BLOCKS)
modis_fast_1km_xyz_kernel(const TIn* __restrict__ lon, const TIn* __restrict__ lat, const TZen* __restrict__ satz,
                          TOut* __restrict__ out_x, TOut* __restrict__ out_y, TOut* __restrict__ out_z,
                          long long n_scans, double zen_scale, double zen_offset)
{
    constexpr bool kZenDecode = !std::is_same<TZen, double>::value;
    constexpr int kWf = kW * P;          // fine columns
    constexpr int kN = kL * P;           // fine rows per scan
    constexpr int kHalf = P / 2;

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

    long long tie_idx = scan * (long long)(kL * kW) + tc;
    long long out_idx = scan * (long long)kN * kWf + (long long)quad_col * P + xoff;

    // carried from the tie row above: this lane's tie point (its sines and cosines are rotated into the
    // next row's), corners A and B of the quad row, and the quad row's expansion / alignment coefficients
    TieGeo top = {0.0, 1.0, 0.0, 1.0, 0.0, 0.0};
    TieZen zen = {0.0, 0.0, 1.0, 0.0, 1.0};
    Vec3 topA = {0.0, 0.0, 0.0}, topB = {0.0, 0.0, 0.0};
    double top_ce = 0.0, top_ca = 0.0;
    __shared__ double stage[2][3][kThreads];
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
        if (r + 1 < kL) {
            tie_idx += kW;
            stage_fetch(&stage[(r + 1) & 1][0][threadIdx.x], lon, tie_idx);
            stage_fetch(&stage[(r + 1) & 1][1][threadIdx.x], lat, tie_idx);
            if (!SIMPLE) stage_fetch(&stage[(r + 1) & 1][2][threadIdx.x], satz, tie_idx);
            cp_async_commit();
        }
        tie_advance<!SIMPLE>(top, zen, clon, clat, csatz, r == 0);        // top is now this row's tie point
        const Vec3 own_xyz = tie_xyz(top);
        Vec3 curB = shfl_down1(own_xyz);
        Vec3 A = topA, D = own_xyz;
        if (last_warp) {                 // the lane of tie column 1353 evaluates the quad of its left neighbour
            const Vec3 upA = shfl_up1(topA), upD = shfl_up1(own_xyz);
            if (extra) { A = upA; D = upD; curB = own_xyz; }
        }

        if (r > 0 && active) {
            const int qr = r - 1;
            const int nrows = P + ((qr == 0 || qr == kL - 2) ? kHalf : 0);
            double s_t = ((qr == 0) ? (0.5 - kHalf) : 0.5) / P;
            const QuadXyz q = quadxyz_setup(A, topB, curB, D);
            double a_col[P];
            const double s_base = extra ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < P; ++k) {
                const double s_s = s_base + (double)k / (double)P;
                a_col[k] = SIMPLE ? s_s : fma(s_s * (1.0 - s_s), top_ce, s_s);
            }
            constexpr long long kPitch = (long long)kW * P;
#pragma unroll 1
            for (int m = 0; m < nrows; ++m) {
                const RowXyz rw = rowxyz_setup(q, s_t, SIMPLE ? 0.0 : (s_t * (1.0 - s_t)) * top_ca);
                double vx[P], vy[P], vz[P];
#pragma unroll
                for (int k = 0; k < P; ++k) {
                    vx[k] = fma(a_col[k], rw.U.x, rw.V.x); vy[k] = fma(a_col[k], rw.U.y, rw.V.y); vz[k] = fma(a_col[k], rw.U.z, rw.V.z);
                }
                store_px<P>(out_x + out_idx, vx);
                store_px<P>(out_y + out_idx, vy);
                store_px<P>(out_z + out_idx, vz);
                out_idx += kPitch;
                s_t += 1.0 / P;              // exact: multiples of 1/8
            }
        }
        topA = own_xyz;
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
            if (active) quad_coefficients(za, zb, sphi_b, cphi_b, 1, top_ce, top_ca);
        }
    }
}

template <typename TOut>
bool outputs_aligned(const void* a, const void* b, const void* c, int p)
{
    const uintptr_t align = (uintptr_t)p * sizeof(TOut);
    return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & (align - 1)) == 0;
}

template <int P, bool SIMPLE, typename TIn, typename TZen, typename TOut>
cudaError_t launch_xyz(const void* lon, const void* lat, const void* satz, void* out_x, void* out_y, void* out_z,
                       long long n_scans, double zen_scale, double zen_offset, cudaStream_t stream)
{
    if (!outputs_aligned<TOut>(out_x, out_y, out_z, P)) return cudaErrorMisalignedAddress;
    if (((uintptr_t)lon | (uintptr_t)lat) & (sizeof(TIn) - 1)) return cudaErrorMisalignedAddress;
    // an int16 zenith angle is fetched as the aligned 4-byte pair that holds it
    if ((uintptr_t)satz & (sizeof(TZen) == 8 ? 7 : 3)) return cudaErrorMisalignedAddress;
    const long long warps = n_scans * kWarpsPerScan;
    const long long blocks = (warps * 32 + kThreads - 1) / kThreads;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    modis_fast_1km_xyz_kernel<P, SIMPLE, TIn, TZen, TOut><<<(unsigned)blocks, kThreads, 0, stream>>>(
        (const TIn*)lon, (const TIn*)lat, (const TZen*)satz, (TOut*)out_x, (TOut*)out_y, (TOut*)out_z,
        n_scans, zen_scale, zen_offset);
    return cudaGetLastError();
}

template <bool SIMPLE, typename TIn, typename TZen, typename TOut>
cudaError_t launch_xyz_p(int p, const void* lon, const void* lat, const void* satz, void* out_x, void* out_y, void* out_z,
                         long long n_scans, double zen_scale, double zen_offset, cudaStream_t stream)
{
    return p == 4 ? launch_xyz<4, SIMPLE, TIn, TZen, TOut>(lon, lat, satz, out_x, out_y, out_z, n_scans, zen_scale, zen_offset, stream)
                  : launch_xyz<2, SIMPLE, TIn, TZen, TOut>(lon, lat, satz, out_x, out_y, out_z, n_scans, zen_scale, zen_offset, stream);
}

template <bool SIMPLE, typename TIn, typename TZen>
cudaError_t launch_xyz_out(int out_is_f64, int p, const void* lon, const void* lat, const void* satz,
                           void* out_x, void* out_y, void* out_z, long long n_scans, double zen_scale, double zen_offset,
                           cudaStream_t stream)
{
    return out_is_f64
        ? launch_xyz_p<SIMPLE, TIn, TZen, double>(p, lon, lat, satz, out_x, out_y, out_z, n_scans, zen_scale, zen_offset, stream)
        : launch_xyz_p<SIMPLE, TIn, TZen, float>(p, lon, lat, satz, out_x, out_y, out_z, n_scans, zen_scale, zen_offset, stream);
}

}  // namespace

// in_is_f64: lon / lat (and, for CVIIRS, the zenith angles) are float64 - the reference's layout; otherwise
// lon / lat are float32 and zen_kind picks float32 (0) or int16 (1) zenith angles, decoded as
// v * zen_scale + zen_offset.  satz == nullptr selects the simple interpolator.  p = 4 | 2.
cudaError_t gtp_launch_xyz_fast(const void* lon, const void* lat, const void* satz, int in_is_f64, int zen_kind,
                                double zen_scale, double zen_offset, void* out_x, void* out_y, void* out_z,
                                int out_is_f64, long long n_scans, int p, cudaStream_t stream)
{
    if (n_scans <= 0) return cudaSuccess;
    if (p != 4 && p != 2) return cudaErrorInvalidValue;
    if (!satz) {
        return in_is_f64
            ? launch_xyz_out<true, double, double>(out_is_f64, p, lon, lat, nullptr, out_x, out_y, out_z, n_scans, 1.0, 0.0, stream)
            : launch_xyz_out<true, float, double>(out_is_f64, p, lon, lat, nullptr, out_x, out_y, out_z, n_scans, 1.0, 0.0, stream);
    }
    if (in_is_f64)
        return launch_xyz_out<false, double, double>(out_is_f64, p, lon, lat, satz, out_x, out_y, out_z, n_scans, 1.0, 0.0, stream);
    return zen_kind == 1
        ? launch_xyz_out<false, float, short>(out_is_f64, p, lon, lat, satz, out_x, out_y, out_z, n_scans, zen_scale, zen_offset, stream)
        : launch_xyz_out<false, float, float>(out_is_f64, p, lon, lat, satz, out_x, out_y, out_z, n_scans, zen_scale, zen_offset, stream);
}
