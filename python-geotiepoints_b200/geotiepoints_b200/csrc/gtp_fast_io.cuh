// gtp_fast_io.cuh - global-memory access helpers shared by the 1 km fast kernels (gtp_fast.cu,
// gtp_fast_xyz.cu): vector streaming stores of P fine pixels and the per-thread cp.async staging of
// the next tie row.
#pragma once
#include <cuda_runtime.h>

namespace gtpfast {

template <int P>
__device__ __forceinline__ void store_px(double* p, const double (&v)[P]);

template <>
__device__ __forceinline__ void store_px<4>(double* p, const double (&v)[4])
{
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
}

template <>
__device__ __forceinline__ void store_px<2>(double* p, const double (&v)[2])
{
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v[0]), "d"(v[1]) : "memory");
}

// float32 outputs of the float64 math (gtp_cviirs_interp_mixed): rounded once, at the store
template <int P>
__device__ __forceinline__ void store_px(float* p, const double (&v)[P]);

template <>
__device__ __forceinline__ void store_px<4>(float* p, const double (&v)[4])
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"((float)v[0]), "f"((float)v[1]),
                 "f"((float)v[2]), "f"((float)v[3]) : "memory");
}

template <>
__device__ __forceinline__ void store_px<2>(float* p, const double (&v)[2])
{
    asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"((float)v[0]), "f"((float)v[1]) : "memory");
}

// 8-byte asynchronous global -> shared copy (LDGSTS): the next tie row is fetched while the
// current quad row is computed without holding the values in registers (with ~165 live
// registers ptxas otherwise parks a register prefetch in local memory and stalls on it).
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gmem_src) : "memory");
}

// One tie-point input element -> its 8-byte staging slot, and back as a double.  float64 is the
// reference's layout; float32 and int16 (MOD03 stores SensorZenith as int16 * 0.01,
// testdata/create_modis_test_data.py:95-97) are the on-disk types, widened here instead of on the
// host.  An int16 travels as the aligned 4-byte pair that holds it (cp.async moves >= 4 bytes).
__device__ __forceinline__ void stage_fetch(double* slot, const double* src, long long idx) { cp_async8(slot, src + idx); }
__device__ __forceinline__ void stage_fetch(double* slot, const float* src, long long idx) { cp_async4(slot, src + idx); }
__device__ __forceinline__ void stage_fetch(double* slot, const short* src, long long idx) { cp_async4(slot, src + (idx & ~1LL)); }
__device__ __forceinline__ double stage_value(const double* slot, const double*, long long) { return *slot; }
__device__ __forceinline__ double stage_value(const double* slot, const float*, long long)
{
    return (double)*reinterpret_cast<const float*>(slot);
}
__device__ __forceinline__ double stage_value(const double* slot, const short*, long long idx)
{
    const int pair = *reinterpret_cast<const int*>(slot);
    return (double)(short)((idx & 1) ? (pair >> 16) : (pair & 0xffff));
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

}  // namespace gtpfast
