// gtp_kernels.h - internal launcher declarations shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include "gtp_common.cuh"

// gtp_generic.cu: rounding-faithful kernels (all configurations, float32/float64).  allow_series: float32
// 5 km -> 1 km may back-transform with the anchor-relative series (same float32 fine xyz, ~3x faster)
template <typename T>
cudaError_t gtp_launch_cviirs_generic(const T* lon, const T* lat, const T* satz, T* out_lon, T* out_lat,
                                      long long n_scans, const CviirsCfg& cfg, cudaStream_t stream, bool allow_series = false);
template <typename T>
cudaError_t gtp_launch_simple_generic(const T* lon, const T* lat, T* out_lon, T* out_lat,
                                      long long n_scans, int width, int res, cudaStream_t stream);

// ECEF output: x / y / z of _compute_fine_xyz instead of their back-transform
template <typename T>
cudaError_t gtp_launch_cviirs_generic_xyz(const T* lon, const T* lat, const T* satz, T* out_x, T* out_y, T* out_z,
                                          long long n_scans, const CviirsCfg& cfg, cudaStream_t stream);
template <typename T>
cudaError_t gtp_launch_simple_generic_xyz(const T* lon, const T* lat, T* out_x, T* out_y, T* out_z,
                                          long long n_scans, int width, int res, cudaStream_t stream);
// gtp_fast_xyz.cu: float64 math at 1 km (width 1354), p = 4 | 2; satz == nullptr: simple interpolator;
// returns cudaErrorMisalignedAddress when the vector stores / staged loads cannot be used
cudaError_t gtp_launch_xyz_fast(const void* lon, const void* lat, const void* satz, int in_is_f64, int zen_kind,
                                double zen_scale, double zen_offset, void* out_x, void* out_y, void* out_z,
                                int out_is_f64, long long n_scans, int p, cudaStream_t stream);

// gtp_fast.cu: float64 anchor-relative fast paths
cudaError_t gtp_launch_cviirs_fast_f64(const double* lon, const double* lat, const double* satz,
                                       double* out_lon, double* out_lat,
                                       long long n_scans, const CviirsCfg& cfg, cudaStream_t stream);
bool gtp_cviirs_fast_f64_supported(const CviirsCfg& cfg);
cudaError_t gtp_launch_simple_fast_f64(const double* lon, const double* lat, double* out_lon, double* out_lat,
                                       long long n_scans, int res, cudaStream_t stream);
bool gtp_simple_fast_f64_supported(int width, int res);
// the same kernel on the on-disk types: float32 lon/lat, float32 | int16 zenith, float32 | float64 out
cudaError_t gtp_launch_cviirs_mixed(const float* lon, const float* lat, const void* satz, int zen_is_i16,
                                    double zen_scale, double zen_offset, void* out_lon, void* out_lat, int out_is_f64,
                                    long long n_scans, int p, cudaStream_t stream);

cudaError_t gtp_launch_simple_mixed(const float* lon, const float* lat, void* out_lon, void* out_lat, int out_is_f64,
                                    long long n_scans, int res, cudaStream_t stream);

// gtp_fast5.cu: float64 5 km -> 1 km fast path (271 or 270 tie columns)
cudaError_t gtp_launch_cviirs_fast5_f64(const double* lon, const double* lat, const double* satz,
                                        double* out_lon, double* out_lat,
                                        long long n_scans, const CviirsCfg& cfg, cudaStream_t stream);
bool gtp_cviirs_fast5_f64_supported(const CviirsCfg& cfg);
cudaError_t gtp_launch_cviirs5_mixed(const float* lon, const float* lat, const void* satz, int zen_is_i16,
                                     double zen_scale, double zen_offset, void* out_lon, void* out_lat, int out_is_f64,
                                     long long n_scans, int W, cudaStream_t stream);

// gtp_fast_f32.cu: float32 CVIIRS 1 km fast path (rounding-faithful blend, anchor-relative back-transform)
cudaError_t gtp_launch_cviirs_fast_f32(const float* lon, const float* lat, const float* satz,
                                       float* out_lon, float* out_lat,
                                       long long n_scans, const CviirsCfg& cfg, cudaStream_t stream);
bool gtp_cviirs_fast_f32_supported(const CviirsCfg& cfg);
// float32 simple interpolator at width 1354: scipy's tap sum and the float32 edge extrapolations
// reproduced exactly, anchor-relative back-transform
cudaError_t gtp_launch_simple_fast_f32(const float* lon, const float* lat, float* out_lon, float* out_lat,
                                       long long n_scans, int res, cudaStream_t stream);
bool gtp_simple_fast_f32_supported(int width, int res);

// gtp_legacy.cu: the reference's legacy spline interpolators (kind = GTP_LEGACY_*), float64
size_t gtp_legacy_workspace_bytes(int kind, long long n_scans);
void gtp_legacy_shape(int kind, int* tie_rows_per_scan, int* tie_cols, int* fine_rows_per_scan, int* fine_cols);
cudaError_t gtp_launch_legacy_f64(const double* lon, const double* lat, double* out_lon, double* out_lat,
                                  long long n_scans, int kind, void* workspace, cudaStream_t stream);

// gtp_vii.cu: VII tie-point interpolation (geotiepoints/viiinterpolator.py), float64
size_t gtp_vii_workspace_bytes(long long n_alt, long long n_act);
cudaError_t gtp_launch_vii_interp(const double* a, const double* b, double* out_a, double* out_b,
                                  long long n_alt, long long n_act, int S, int F, cudaStream_t stream);
cudaError_t gtp_launch_vii_geo_cartesian(const double* lon, const double* lat, double* out_lon, double* out_lat,
                                         long long n_alt, long long n_act, int S, int F, double z_threshold,
                                         void* workspace, cudaStream_t stream);

// gtp_grid.cu: generic rectilinear-grid interpolators (geotiepoints/interpolator.py:232-378, geointerpolator.py:76-113)
size_t gtp_grid_workspace_bytes(int ny, int nx, long long my, long long mx, int n_comp, int geo, int method_y, int method_x);
cudaError_t gtp_launch_grid(const double* tie_y, int ny, const double* tie_x, int nx, const double* values,
                            const double* values2, int n_comp, int geo, const double* fine_y, long long my,
                            const double* fine_x, long long mx, int method_y, int method_x, int oob_mode, double fill,
                            double* out, double* out2, void* workspace, cudaStream_t stream, int* launches);

// gtp_fast5_f32.cu: float32 5 km (270 | 271 columns) -> 1 km fast path
bool gtp_cviirs_fast5_f32_supported(const CviirsCfg& cfg);
cudaError_t gtp_launch_cviirs_fast5_f32(const float* lon, const float* lat, const float* satz, float* out_lon, float* out_lat,
                                        long long n_scans, const CviirsCfg& cfg, cudaStream_t stream);
