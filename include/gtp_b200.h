/*
 * gtp_b200.h - C ABI of libgtp_b200.so: MODIS geolocation tie-point interpolation on
 * NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the hot path of pytroll/python-geotiepoints 1.9.0.
 * Each entry point replaces one compiled operator of the reference (paths relative to
 * the reference tree, geotiepoints/...):
 *
 *   gtp_cviirs_interp_{f32,f64}        _modis_interpolator.pyx:17-38  `interpolate(lon1, lat1,
 *                                      satz1, coarse_resolution, fine_resolution,
 *                                      coarse_scan_width)` -> MODISInterpolator.interpolate
 *                                      (:130-194), i.e. what modisinterpolator.py:21-62
 *                                      (modis_1km_to_250m / _500m, modis_5km_to_1km /
 *                                      _500m / _250m) call.
 *   gtp_simple_interp_{f32,f64}        _simple_modis_interpolator.pyx:13-64
 *                                      `interpolate_geolocation_cartesian(lon_array,
 *                                      lat_array, coarse_resolution, fine_resolution)`,
 *                                      i.e. what simple_modis_interpolator.py:45-61 call.
 *   gtp_*_interp_host_{f32,f64}        the same operators for HOST buffers (numpy
 *                                      callers): pipelined H2D / kernel / D2H inside.
 *
 * Conventions: plain pointers and sizes, no exceptions, re-entrant and thread-safe; the only
 * process state is a thread-local error string, a launch counter and the cache of idle
 * streams + device scratch of the host entry points (gtp_trim() empties it).  Every function returns 0 on success, a GTP_E_*
 * code (>0) for invalid arguments, or -(cudaError_t) for CUDA failures;
 * gtp_last_error() describes the last failure on the calling thread.  The *_interp_*
 * device entry points are asynchronous on `stream` (a cudaStream_t, NULL = default
 * stream); the caller owns and pre-allocates every buffer.  Arrays are C-contiguous
 * row-major, rows = scans * rows_per_scan.  There is no CPU fallback: without a CUDA
 * device every compute entry point fails.
 *
 * Shapes: CVIIRS  in  3 x (n_scans*L, W)      L=10,W=1354 (1 km) | L=2,W=270|271 (5 km)
 *                                 out 2 x (n_scans*L*f, 1354*p) p=1000/fine_res, f=p*coarse_res/1000
 *                        simple  in  2 x (n_scans*10, width)
 *                                 out 2 x (n_scans*10*res, width*res)  res = 2 | 4
 */
#ifndef GTP_B200_H
#define GTP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTP_OK 0
#define GTP_E_RESOLUTION 1   /* unsupported coarse/fine resolution pair                         */
#define GTP_E_WIDTH_5KM 2    /* 5 km tie points with a width other than 270/271 (reference:     */
                             /* NotImplementedError, _modis_interpolator.pyx:33-36)             */
#define GTP_E_NULL 3         /* null pointer with n_scans > 0                                   */
#define GTP_E_SHAPE 4        /* negative n_scans / width too small                              */
#define GTP_E_NO_DEVICE 5    /* no usable CUDA device                                           */

const char* gtp_version(void);
const char* gtp_last_error(void);
/* number of CUDA devices visible (0 when none / no driver); never fails */
int gtp_device_count(void);

/* Output geometry for a configuration: rows per scan in / out, columns in / out.
 * coarse_width is only read for coarse_res == 5000 (0 means 271). */
int gtp_cviirs_geometry(int coarse_res, int fine_res, int coarse_width,
                        int* rows_per_scan_in, int* cols_in, int* rows_per_scan_out, int* cols_out);

/* ---- device-pointer entry points (asynchronous on `stream`) -------------------- */
int gtp_cviirs_interp_f32(const float* lon, const float* lat, const float* satz, int64_t n_scans,
                          int coarse_res, int fine_res, int coarse_width,
                          float* out_lon, float* out_lat, int device, void* stream);
int gtp_cviirs_interp_f64(const double* lon, const double* lat, const double* satz, int64_t n_scans,
                          int coarse_res, int fine_res, int coarse_width,
                          double* out_lon, double* out_lat, int device, void* stream);
int gtp_simple_interp_f32(const float* lon, const float* lat, int64_t n_scans, int64_t width, int res_factor,
                          float* out_lon, float* out_lat, int device, void* stream);
int gtp_simple_interp_f64(const double* lon, const double* lat, int64_t n_scans, int64_t width, int res_factor,
                          double* out_lon, double* out_lat, int device, void* stream);

/* Same operators, but always through the rounding-faithful generic kernels (never the
 * anchor-relative fast paths); used by tests and by the bench to show the two side by side. */
int gtp_cviirs_interp_generic_f64(const double* lon, const double* lat, const double* satz, int64_t n_scans,
                                  int coarse_res, int fine_res, int coarse_width,
                                  double* out_lon, double* out_lat, int device, void* stream);

int gtp_cviirs_interp_generic_f32(const float* lon, const float* lat, const float* satz, int64_t n_scans,
                                  int coarse_res, int fine_res, int coarse_width,
                                  float* out_lon, float* out_lat, int device, void* stream);
int gtp_simple_interp_generic_f64(const double* lon, const double* lat, int64_t n_scans, int64_t width, int res_factor,
                                  double* out_lon, double* out_lat, int device, void* stream);

/* ---- host-pointer entry points (synchronous; copies inside) -------------------- */
/* Inputs and outputs are host buffers (pageable or pinned).  The call splits the scans
 * into chunks and overlaps H2D, kernel and D2H on three cached streams; it returns when the
 * outputs are complete.  Pinned buffers (gtp_host_alloc) reach PCIe line rate. */
int gtp_cviirs_interp_host_f32(const float* lon, const float* lat, const float* satz, int64_t n_scans,
                               int coarse_res, int fine_res, int coarse_width,
                               float* out_lon, float* out_lat, int device);
int gtp_cviirs_interp_host_f64(const double* lon, const double* lat, const double* satz, int64_t n_scans,
                               int coarse_res, int fine_res, int coarse_width,
                               double* out_lon, double* out_lat, int device);
int gtp_simple_interp_host_f32(const float* lon, const float* lat, int64_t n_scans, int64_t width, int res_factor,
                               float* out_lon, float* out_lat, int device);
int gtp_simple_interp_host_f64(const double* lon, const double* lat, int64_t n_scans, int64_t width, int res_factor,
                               double* out_lon, double* out_lat, int device);

/* ---- on-disk types in, float64 math (extension) --------- */
/* MOD03 stores Longitude / Latitude as float32 and SensorZenith as int16 * 0.01
 * (reference: testdata/create_modis_test_data.py:85-99).  These entry points take
 * those arrays as they are - float32 lon / lat, zenith angles as float32 (GTP_ZEN_F32) or int16
 * (GTP_ZEN_I16), decoded in the kernel as value * zen_scale + zen_offset degrees - run the float64
 * algorithm of gtp_cviirs_interp_f64 on the widened values and write float32 (GTP_OUT_F32, rounded
 * once at the store) or float64 (GTP_OUT_F64) results.  The result equals the reference's float64
 * path on `lon.astype(float64), lat.astype(float64), zenith * scale + offset` (to the float64
 * parity of this library), NOT its float32 path, whose fine Cartesian coordinates are float32
 * (0.25-0.5 m of rounding noise).  coarse_res 1000 with fine_res 250 | 500, or coarse_res 5000
 * (coarse_width 270 | 271, 0 = 271) with fine_res 1000.  Inputs must be 4-byte aligned (int16: each
 * chunk of scans starts on an even element, which whole scans do), outputs 16-byte (float64 at
 * 250 m: 32-byte) aligned. */
#define GTP_ZEN_F32 0
#define GTP_ZEN_I16 1
#define GTP_OUT_F32 0
#define GTP_OUT_F64 1
int gtp_cviirs_interp_mixed(const float* lon, const float* lat, const void* satz, int zen_kind,
                            double zen_scale, double zen_offset, int64_t n_scans, int coarse_res, int fine_res,
                            int coarse_width, void* out_lon, void* out_lat, int out_kind, int device, void* stream);
int gtp_cviirs_interp_mixed_host(const float* lon, const float* lat, const void* satz, int zen_kind,
                                 double zen_scale, double zen_offset, int64_t n_scans, int coarse_res, int fine_res,
                                 int coarse_width, void* out_lon, void* out_lat, int out_kind, int device);

/* simple_modis_interpolator (gtp_simple_interp_f64) on float32 lon / lat of width 1354: float64 math,
 * float32 | float64 results */
int gtp_simple_interp_mixed(const float* lon, const float* lat, int64_t n_scans, int res_factor,
                            void* out_lon, void* out_lat, int out_kind, int device, void* stream);
int gtp_simple_interp_mixed_host(const float* lon, const float* lat, int64_t n_scans, int res_factor,
                                 void* out_lon, void* out_lat, int out_kind, int device);

/* ---- ECEF output (extension) ---------------------------------- */
/* The fine Cartesian coordinates x, y, z (metres; sphere of radius 6370997 m, _modis_utils.pyx:23) that
 * the reference computes in MODISInterpolator._compute_fine_xyz (_modis_interpolator.pyx:346-398) -
 * new_xyz in _simple_modis_interpolator.pyx:52-63 - and hands to xyz2lonlat (:293): the same operators
 * stopped one step earlier, for consumers (resampling) that convert lon / lat straight back to ECEF.
 * Three outputs of the shape of out_lon / out_lat above.  The _f32 / _f64 forms follow the reference's
 * arithmetic in that type for every configuration (float64 1 km tie points of width 1354 with
 * 4- | 2-element aligned outputs take gtp_fast_xyz.cu, everything else the rounding-faithful kernels;
 * _generic_ forces the latter); the _mixed forms take MOD03's on-disk types like gtp_cviirs_interp_mixed
 * (float64 math, float32 | float64 out; coarse_res 1000 with fine_res 250 | 500 only). */
int gtp_cviirs_interp_xyz_f32(const float* lon, const float* lat, const float* satz, int64_t n_scans,
                              int coarse_res, int fine_res, int coarse_width,
                              float* out_x, float* out_y, float* out_z, int device, void* stream);
int gtp_cviirs_interp_xyz_f64(const double* lon, const double* lat, const double* satz, int64_t n_scans,
                              int coarse_res, int fine_res, int coarse_width,
                              double* out_x, double* out_y, double* out_z, int device, void* stream);
int gtp_cviirs_interp_xyz_generic_f64(const double* lon, const double* lat, const double* satz, int64_t n_scans,
                                      int coarse_res, int fine_res, int coarse_width,
                                      double* out_x, double* out_y, double* out_z, int device, void* stream);
int gtp_simple_interp_xyz_f32(const float* lon, const float* lat, int64_t n_scans, int64_t width, int res_factor,
                              float* out_x, float* out_y, float* out_z, int device, void* stream);
int gtp_simple_interp_xyz_f64(const double* lon, const double* lat, int64_t n_scans, int64_t width, int res_factor,
                              double* out_x, double* out_y, double* out_z, int device, void* stream);
int gtp_simple_interp_xyz_generic_f64(const double* lon, const double* lat, int64_t n_scans, int64_t width, int res_factor,
                                      double* out_x, double* out_y, double* out_z, int device, void* stream);
int gtp_cviirs_interp_xyz_host_f32(const float* lon, const float* lat, const float* satz, int64_t n_scans,
                                   int coarse_res, int fine_res, int coarse_width,
                                   float* out_x, float* out_y, float* out_z, int device);
int gtp_cviirs_interp_xyz_host_f64(const double* lon, const double* lat, const double* satz, int64_t n_scans,
                                   int coarse_res, int fine_res, int coarse_width,
                                   double* out_x, double* out_y, double* out_z, int device);
int gtp_simple_interp_xyz_host_f32(const float* lon, const float* lat, int64_t n_scans, int64_t width, int res_factor,
                                   float* out_x, float* out_y, float* out_z, int device);
int gtp_simple_interp_xyz_host_f64(const double* lon, const double* lat, int64_t n_scans, int64_t width, int res_factor,
                                   double* out_x, double* out_y, double* out_z, int device);
int gtp_cviirs_interp_xyz_mixed(const float* lon, const float* lat, const void* satz, int zen_kind,
                                double zen_scale, double zen_offset, int64_t n_scans, int coarse_res, int fine_res,
                                void* out_x, void* out_y, void* out_z, int out_kind, int device, void* stream);
int gtp_cviirs_interp_xyz_mixed_host(const float* lon, const float* lat, const void* satz, int zen_kind,
                                     double zen_scale, double zen_offset, int64_t n_scans, int coarse_res, int fine_res,
                                     void* out_x, void* out_y, void* out_z, int out_kind, int device);
int gtp_simple_interp_xyz_mixed(const float* lon, const float* lat, int64_t n_scans, int res_factor,
                                void* out_x, void* out_y, void* out_z, int out_kind, int device, void* stream);
int gtp_simple_interp_xyz_mixed_host(const float* lon, const float* lat, int64_t n_scans, int res_factor,
                                     void* out_x, void* out_y, void* out_z, int out_kind, int device);

/* ---- legacy spline interpolators ------------------------------------------
 * Replace geotiepoints.modis5kmto1km / modis1kmto500m / modis1kmto250m
 * (/root/reference/geotiepoints/__init__.py:49-152): GeoInterpolator (geointerpolator.py:10-75) over
 * Interpolator (interpolator.py:54-229) with along-track order 1, cross-track order 3, chunk_size = one
 * scan, fill_borders("y", "x"), i.e. scipy's RectBivariateSpline(kx = 1, ky = 3, s = 0) on the Cartesian
 * tie points.  float64 lon / lat in degrees, whole scans:
 *   GTP_LEGACY_5KM_1KM    (2 n_scans, 271)   -> (10 n_scans, 1354)
 *   GTP_LEGACY_1KM_500M   (10 n_scans, 1354) -> (20 n_scans, 2708)
 *   GTP_LEGACY_1KM_250M   (10 n_scans, 1354) -> (40 n_scans, 5416)
 * The device form takes device pointers and allocates its scratch (48 bytes per tie point) stream-ordered
 * (cudaMallocAsync) on `stream`; the host form copies in chunks like the other host entry points. */
#define GTP_LEGACY_5KM_1KM 0
#define GTP_LEGACY_1KM_500M 1
#define GTP_LEGACY_1KM_250M 2
int gtp_legacy_interp_f64(const double* lon, const double* lat, int64_t n_scans, int kind,
                          double* out_lon, double* out_lat, int device, void* stream);
int gtp_legacy_interp_host_f64(const double* lon, const double* lat, int64_t n_scans, int kind,
                               double* out_lon, double* out_lat, int device);

/* ---- VII (EPS-SG METimage) tie-point interpolation -----------------------------
 * Replace geotiepoints.viiinterpolator.tie_points_interpolation (viiinterpolator.py:23-79, one array per
 * call) and tie_points_geo_interpolation (:83-134).  `data` is (n_tie_alt, n_tie_act) float64, C-contiguous,
 * n_tie_alt a multiple of scan_alt_tie_points (else GTP_E_SHAPE - the reference's ValueError); the result is
 * (n_tie_alt / S * (S - 1) * factor, (n_tie_act - 1) * factor).  `cartesian` is the reference's decision
 * (:108: any |lat| > lat_threshold_use_cartesian or a longitude span > 180 deg), taken by the caller, who has
 * the arrays; z_threshold = z_threshold_use_xy (:86). */
int gtp_vii_interp_f64(const double* data, int64_t n_tie_alt, int64_t n_tie_act, int scan_alt_tie_points,
                       int tie_points_factor, double* out, int device, void* stream);
int gtp_vii_interp_host_f64(const double* data, int64_t n_tie_alt, int64_t n_tie_act, int scan_alt_tie_points,
                            int tie_points_factor, double* out, int device);
int gtp_vii_geo_interp_f64(const double* lon, const double* lat, int64_t n_tie_alt, int64_t n_tie_act,
                           int scan_alt_tie_points, int tie_points_factor, int cartesian, double z_threshold,
                           double* out_lon, double* out_lat, int device, void* stream);
int gtp_vii_geo_interp_host_f64(const double* lon, const double* lat, int64_t n_tie_alt, int64_t n_tie_act,
                                int scan_alt_tie_points, int tie_points_factor, int cartesian, double z_threshold,
                                double* out_lon, double* out_lat, int device);

/* ---- generic rectilinear-grid interpolators (third item) ------------------------
 * Replace what geotiepoints.interpolator.SingleGridInterpolator / SingleSplineInterpolator
 * (/root/reference/geotiepoints/interpolator.py:232-340: scipy's RegularGridInterpolator /
 * RectBivariateSpline(s = 0)) and geointerpolator.GeoGridInterpolator / GeoSplineInterpolator
 * (geointerpolator.py:76-113: the same on the Cartesian x / y / z of lon / lat, then xyz2lonlat :63-75) compute.
 * Tie points: `values` = n_arrays planes of (n_tie_y, n_tie_x) float64 on the rectilinear grid tie_y x tie_x
 * (strictly ascending; descending grids are flipped by the caller, as scipy does).  Result: n_arrays planes of
 * (n_fine_y, n_fine_x) at the product of the two fine coordinate vectors.  Per axis one scheme:
 *   GTP_GRID_NEAREST, GTP_GRID_LINEAR ("linear" / "slinear" / spline order 1),
 *   GTP_GRID_CUBIC (interpolating cubic spline, not-a-knot ends: "cubic" / spline order 3; needs >= 4 tie points).
 * Fine coordinates outside the tie-point range: GTP_GRID_OOB_EXTRAPOLATE (scipy's fill_value = None: the end
 * polynomial continues), GTP_GRID_OOB_FILL (`fill_value`, scipy's bounds_error = False; the geo form
 * back-transforms x = y = z = fill_value like the reference does), GTP_GRID_OOB_CLAMP (the end value: FITPACK
 * clamps the arguments of RectBivariateSpline).  bounds_error = True is the caller's check.
 * Device forms take device pointers for every array (coordinate vectors included) and allocate their scratch
 * stream-ordered; host forms stage everything through device memory and return when the results are there. */
#define GTP_GRID_NEAREST 0
#define GTP_GRID_LINEAR 1
#define GTP_GRID_CUBIC 3
#define GTP_GRID_OOB_EXTRAPOLATE 0
#define GTP_GRID_OOB_FILL 1
#define GTP_GRID_OOB_CLAMP 2
int gtp_grid_interp_f64(const double* tie_y, int64_t n_tie_y, const double* tie_x, int64_t n_tie_x,
                        const double* values, int n_arrays, const double* fine_y, int64_t n_fine_y,
                        const double* fine_x, int64_t n_fine_x, int method_y, int method_x, int oob_mode,
                        double fill_value, double* out, int device, void* stream);
int gtp_grid_interp_host_f64(const double* tie_y, int64_t n_tie_y, const double* tie_x, int64_t n_tie_x,
                             const double* values, int n_arrays, const double* fine_y, int64_t n_fine_y,
                             const double* fine_x, int64_t n_fine_x, int method_y, int method_x, int oob_mode,
                             double fill_value, double* out, int device);
int gtp_geogrid_interp_f64(const double* tie_y, int64_t n_tie_y, const double* tie_x, int64_t n_tie_x,
                           const double* lon, const double* lat, const double* fine_y, int64_t n_fine_y,
                           const double* fine_x, int64_t n_fine_x, int method_y, int method_x, int oob_mode,
                           double fill_value, double* out_lon, double* out_lat, int device, void* stream);
int gtp_geogrid_interp_host_f64(const double* tie_y, int64_t n_tie_y, const double* tie_x, int64_t n_tie_x,
                                const double* lon, const double* lat, const double* fine_y, int64_t n_fine_y,
                                const double* fine_x, int64_t n_fine_x, int method_y, int method_x, int oob_mode,
                                double fill_value, double* out_lon, double* out_lat, int device);

/* The host entry points keep up to 4 idle sets of (3 streams + chunk scratch, ~300 MB of
 * device memory each) per process for the next call; gtp_trim() frees them. */
int gtp_trim(void);

/* Device-to-host relay of the host entry points (process-wide; -1 = off, the default).  On a multi-GPU host
 * whose PCIe root ports are not equally fast, a process may route its results over NVLink through a
 * staging buffer on `device` (peer access required) and copy them to the host from there:
 *   kernel on GPU a -> cudaMemcpyPeerAsync a -> `device` -> cudaMemcpyAsync `device` -> host.
 * Results are bit-identical; only the path of the bytes changes.  Measured on the 8 x B200 hosts of this
 * pool (profiles/r02_pcie_probe2_n8.txt): GPUs 0-3 get 11 GB/s each and GPUs 4-7 18 GB/s each when all
 * eight copy at once (118 GB/s), while GPUs 4-7 alone sustain 177 GB/s. */
int gtp_set_d2h_relay(int device);
int gtp_get_d2h_relay(void);

/* page-locked host memory for the host entry points */
int gtp_host_alloc(void** ptr, uint64_t bytes);
int gtp_host_free(void* ptr);

/* Number of kernels this library has launched in the calling process (all threads);
 * bench.py reports the delta over its timed region as `gpu_launches`. */
uint64_t gtp_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GTP_B200_H */
