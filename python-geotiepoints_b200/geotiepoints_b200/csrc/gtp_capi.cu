// gtp_capi.cu - the extern "C" boundary of libgtp_b200.so (declared in include/gtp_b200.h).
//
// Thin by design: argument validation with the reference's error conditions
// (/root/reference/geotiepoints/_modis_interpolator.pyx:33-36), kernel selection, and
// the chunked H2D / kernel / D2H pipeline of the host-buffer entry points.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>
#include <cuda_runtime.h>
#include "../../../include/gtp_b200.h"
#include "gtp_kernels.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, const char* detail = "")
{
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}

int cuda_fail(cudaError_t e, const char* where)
{
    snprintf(g_err, sizeof(g_err), "CUDA error in %s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
    return -(int)e;
}

int cfg_error(int rc)
{
    if (rc == 2) return fail(GTP_E_WIDTH_5KM, "Can't interpolate if 5km tiepoints have less than 270 columns.%s");
    return fail(GTP_E_RESOLUTION, "unsupported coarse/fine resolution pair%s");
}

// RAII "run on this device, then put the caller's device back"
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device)
    {
        err = cudaGetDevice(&prev);
        if (err != cudaSuccess) return;
        if (device >= 0 && device != prev) {
            err = cudaSetDevice(device);
            switched = (err == cudaSuccess);
        }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

template <typename T>
int cviirs_device(const T* lon, const T* lat, const T* satz, int64_t n_scans, int coarse_res, int fine_res,
                  int coarse_width, T* out_lon, T* out_lat, int device, void* stream, bool allow_fast)
{
    CviirsCfg cfg;
    int rc = gtp_make_cfg(coarse_res, fine_res, coarse_width, &cfg);
    if (rc) return cfg_error(rc);
    if (n_scans < 0) return fail(GTP_E_SHAPE, "n_scans must be >= 0%s");
    if (n_scans == 0) return GTP_OK;
    if (!lon || !lat || !satz || !out_lon || !out_lat) return fail(GTP_E_NULL, "null buffer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) { cudaGetLastError(); return fail(GTP_E_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(guard.err)); }
    cudaError_t e;
    if constexpr (sizeof(T) == 8) {
        if (allow_fast && gtp_cviirs_fast_f64_supported(cfg))
            e = gtp_launch_cviirs_fast_f64(lon, lat, satz, out_lon, out_lat, n_scans, cfg, (cudaStream_t)stream);
        else if (allow_fast && gtp_cviirs_fast5_f64_supported(cfg))
            e = gtp_launch_cviirs_fast5_f64(lon, lat, satz, out_lon, out_lat, n_scans, cfg, (cudaStream_t)stream);
        else
            e = gtp_launch_cviirs_generic<T>(lon, lat, satz, out_lon, out_lat, n_scans, cfg, (cudaStream_t)stream);
    } else {
        if (allow_fast && gtp_cviirs_fast_f32_supported(cfg))
            e = gtp_launch_cviirs_fast_f32(lon, lat, satz, out_lon, out_lat, n_scans, cfg, (cudaStream_t)stream);
        else if (allow_fast && gtp_cviirs_fast5_f32_supported(cfg))
            e = gtp_launch_cviirs_fast5_f32(lon, lat, satz, out_lon, out_lat, n_scans, cfg, (cudaStream_t)stream);
        else
            e = gtp_launch_cviirs_generic<T>(lon, lat, satz, out_lon, out_lat, n_scans, cfg, (cudaStream_t)stream, allow_fast);
    }
    if (e != cudaSuccess) return cuda_fail(e, "cviirs launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GTP_OK;
}

template <typename T>
int simple_device(const T* lon, const T* lat, int64_t n_scans, int64_t width, int res, T* out_lon, T* out_lat,
                  int device, void* stream, bool allow_fast)
{
    if (res != 2 && res != 4) return fail(GTP_E_RESOLUTION, "simple interpolator: res_factor must be 2 or 4%s");
    if (n_scans < 0) return fail(GTP_E_SHAPE, "n_scans must be >= 0%s");
    if (width < 4 || width > (1 << 24)) return fail(GTP_E_SHAPE, "simple interpolator: width must be in [4, 2^24]%s");
    if (n_scans == 0) return GTP_OK;
    if (!lon || !lat || !out_lon || !out_lat) return fail(GTP_E_NULL, "null buffer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) { cudaGetLastError(); return fail(GTP_E_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(guard.err)); }
    cudaError_t e;
    if constexpr (sizeof(T) == 8) {
        if (allow_fast && gtp_simple_fast_f64_supported((int)width, res))
            e = gtp_launch_simple_fast_f64(lon, lat, out_lon, out_lat, n_scans, res, (cudaStream_t)stream);
        else
            e = gtp_launch_simple_generic<T>(lon, lat, out_lon, out_lat, n_scans, (int)width, res, (cudaStream_t)stream);
    } else {
        if (allow_fast && gtp_simple_fast_f32_supported((int)width, res))
            e = gtp_launch_simple_fast_f32(lon, lat, out_lon, out_lat, n_scans, res, (cudaStream_t)stream);
        else
            e = gtp_launch_simple_generic<T>(lon, lat, out_lon, out_lat, n_scans, (int)width, res, (cudaStream_t)stream);
    }
    if (e != cudaSuccess) return cuda_fail(e, "simple launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GTP_OK;
}

// ---------------------------------------------------------------------------------
// Host-buffer pipeline.  The scans are cut into chunks; chunk k runs H2D -> kernel ->
// D2H in order on stream k % kStreams, so the two copy engines and the SMs overlap
// across chunks.  Streams and device scratch live in a HostPipe that is checked out of
// a small per-process cache for the duration of one call and handed back afterwards
// (creating streams and mapping ~300 MB of device memory costs as much as moving one
// granule over PCIe; measured, profiles/r01_e2e_probe.txt).  A pipe is owned by exactly
// one call at a time, so concurrent callers (dask worker threads) never share state.
// ---------------------------------------------------------------------------------
constexpr int kStreams = 3;
constexpr int kMaxIdlePipes = 4;

int64_t chunk_out_bytes()
{
    // per output array and chunk; GTP_CHUNK_MB is a tuning knob for the e2e probe
    static const int64_t bytes = [] {
        const char* env = getenv("GTP_CHUNK_MB");
        long mb = env ? atol(env) : 0;
        if (mb < 1 || mb > 4096) mb = 64;
        return (int64_t)mb << 20;
    }();
    return bytes;
}

// D2H relay (gtp_set_d2h_relay): on multi-GPU hosts whose PCIe root ports are not equally fast (measured on the
// 8 x B200 boxes of this pool: GPUs 0-3 share ~74 GB/s of device-to-host bandwidth, GPUs 4-7 get 44 GB/s
// each, and with all eight copying the host takes 118 GB/s in total; profiles/r02_pcie_probe2_n8.txt) a rank
// may send its results over NVLink to a staging buffer on a GPU with a faster port and copy them to the host
// from there.  -1 = off.
std::atomic<int> g_relay_device{-1};

struct HostPipe {
    int device = -1;
    cudaStream_t streams[kStreams] = {nullptr, nullptr, nullptr};
    char* scratch[kStreams] = {nullptr, nullptr, nullptr};
    size_t capacity[kStreams] = {0, 0, 0};
    // relay side: streams, staging buffers and events on the relay device
    int relay_device = -1;
    cudaStream_t relay_streams[kStreams] = {nullptr, nullptr, nullptr};
    char* relay_buf[kStreams] = {nullptr, nullptr, nullptr};
    size_t relay_cap[kStreams] = {0, 0, 0};
    cudaEvent_t copied[kStreams] = {nullptr, nullptr, nullptr};      // on the source stream: peer copy done
    cudaEvent_t drained[kStreams] = {nullptr, nullptr, nullptr};     // on the relay stream: staging buffer read
    bool drained_valid[kStreams] = {false, false, false};

    // called with the SOURCE device current; leaves it current
    cudaError_t reserve_relay(int s, size_t bytes, int relay)
    {
        cudaError_t e = cudaSuccess;
        if (relay_device != relay) { destroy_relay(); relay_device = relay; }
        if (!copied[s]) { e = cudaEventCreateWithFlags(&copied[s], cudaEventDisableTiming); if (e != cudaSuccess) return e; }
        int can = 0;
        e = cudaDeviceCanAccessPeer(&can, device, relay);
        if (e != cudaSuccess) return e;
        if (!can) return cudaErrorPeerAccessUnsupported;
        e = cudaDeviceEnablePeerAccess(relay, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
        if (e != cudaSuccess) return e;
        e = cudaSetDevice(relay);
        if (e != cudaSuccess) return e;
        if (!relay_streams[s]) e = cudaStreamCreateWithFlags(&relay_streams[s], cudaStreamNonBlocking);
        if (e == cudaSuccess && !drained[s]) e = cudaEventCreateWithFlags(&drained[s], cudaEventDisableTiming);
        if (e == cudaSuccess && relay_cap[s] < bytes) {
            if (relay_buf[s]) { cudaFree(relay_buf[s]); relay_buf[s] = nullptr; relay_cap[s] = 0; }
            e = cudaMalloc((void**)&relay_buf[s], bytes);
            if (e == cudaSuccess) relay_cap[s] = bytes;
        }
        const cudaError_t back = cudaSetDevice(device);
        return e != cudaSuccess ? e : back;
    }
    void destroy_relay()
    {
        if (relay_device >= 0) {
            int cur = -1;
            cudaGetDevice(&cur);
            if (cudaSetDevice(relay_device) == cudaSuccess) {
                for (int s = 0; s < kStreams; ++s) {
                    if (relay_buf[s]) cudaFree(relay_buf[s]);
                    if (relay_streams[s]) cudaStreamDestroy(relay_streams[s]);
                    if (drained[s]) cudaEventDestroy(drained[s]);
                }
            }
            cudaGetLastError();
            if (cur >= 0) cudaSetDevice(cur);
        }
        for (int s = 0; s < kStreams; ++s) {
            relay_buf[s] = nullptr; relay_streams[s] = nullptr; drained[s] = nullptr; relay_cap[s] = 0; drained_valid[s] = false;
            if (copied[s]) { cudaEventDestroy(copied[s]); copied[s] = nullptr; }
        }
        relay_device = -1;
    }

    cudaError_t reserve(int s, size_t bytes)
    {
        if (!streams[s]) {
            cudaError_t e = cudaStreamCreateWithFlags(&streams[s], cudaStreamNonBlocking);
            if (e != cudaSuccess) return e;
        }
        if (capacity[s] >= bytes) return cudaSuccess;
        if (scratch[s]) { cudaFree(scratch[s]); scratch[s] = nullptr; capacity[s] = 0; }
        cudaError_t e = cudaMalloc((void**)&scratch[s], bytes);
        if (e == cudaSuccess) capacity[s] = bytes;
        return e;
    }
    void destroy()
    {
        destroy_relay();
        for (int s = 0; s < kStreams; ++s) {
            if (scratch[s]) cudaFree(scratch[s]);
            if (streams[s]) cudaStreamDestroy(streams[s]);
            scratch[s] = nullptr; streams[s] = nullptr; capacity[s] = 0;
        }
    }
};

std::mutex g_pipe_mutex;
std::vector<HostPipe*> g_idle_pipes;

// called with `device` current
HostPipe* acquire_pipe(int device)
{
    for (;;) {
        HostPipe* p = nullptr;
        {
            std::lock_guard<std::mutex> lock(g_pipe_mutex);
            for (size_t i = 0; i < g_idle_pipes.size(); ++i)
                if (g_idle_pipes[i]->device == device) {
                    p = g_idle_pipes[i];
                    g_idle_pipes.erase(g_idle_pipes.begin() + i);
                    break;
                }
        }
        if (!p) break;
        // a set kept from before a cudaDeviceReset / context change holds stale handles: drop it as it is
        // (its streams and scratch went with the context) and look on
        const cudaError_t q = p->streams[0] ? cudaStreamQuery(p->streams[0]) : cudaSuccess;
        if (q == cudaSuccess || q == cudaErrorNotReady) return p;
        cudaGetLastError();
        delete p;
    }
    HostPipe* p = new HostPipe;
    p->device = device;
    return p;
}

void release_pipe(HostPipe* p, bool healthy)
{
    if (healthy) {
        std::lock_guard<std::mutex> lock(g_pipe_mutex);
        if ((int)g_idle_pipes.size() < kMaxIdlePipes) { g_idle_pipes.push_back(p); return; }
    }
    p->destroy();      // the caller still has p->device current
    delete p;
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

#define GTP_CUDA_TRY(expr, where)                                  \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) { status = cuda_fail(_e, where); goto done; } \
    } while (0)

// Byte-level core: `in` are n_in host arrays of in_bytes[a] bytes per scan each; the n_out outputs have
// out_bytes bytes per scan each.  launch(d_in, d_out, n_scans_of_chunk, stream) enqueues the kernel.
constexpr int kMaxInputs = 3;
constexpr int kMaxOutputs = 3;

template <typename Launch>
int host_pipeline_n(const void* const* in, const size_t* in_bytes, int n_in, void* const* out, int n_out,
                    size_t out_bytes, int64_t n_scans, int device, Launch launch)
{
    if (n_scans == 0) return GTP_OK;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) { cudaGetLastError(); return fail(GTP_E_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(guard.err)); }
    int current = -1;
    if (cudaGetDevice(&current) != cudaSuccess) { cudaGetLastError(); return fail(GTP_E_NO_DEVICE, "no usable CUDA device%s"); }
    int status = GTP_OK;
    int64_t chunk = std::max<int64_t>(1, chunk_out_bytes() / (int64_t)out_bytes);
    chunk = std::min(chunk, n_scans);
    const int64_t n_chunks = (n_scans + chunk - 1) / chunk;
    const int n_streams = (int)std::min<int64_t>(kStreams, n_chunks);
    size_t in_off[kMaxInputs + 1] = {0};
    for (int a = 0; a < n_in; ++a) in_off[a + 1] = in_off[a] + align_up((size_t)chunk * in_bytes[a]);
    const size_t out_chunk = align_up((size_t)chunk * out_bytes);
    HostPipe* pipe = acquire_pipe(current);
    void* d_in[kStreams][kMaxInputs] = {};
    void* d_out[kStreams][kMaxOutputs] = {};
    int relay = g_relay_device.load(std::memory_order_relaxed);
    if (relay == current) relay = -1;
    if (relay < 0 && pipe->relay_device >= 0) pipe->destroy_relay();
    for (int s = 0; s < n_streams; ++s) {
        GTP_CUDA_TRY(pipe->reserve(s, in_off[n_in] + n_out * out_chunk), "scratch alloc");
        for (int a = 0; a < n_in; ++a) d_in[s][a] = pipe->scratch[s] + in_off[a];
        for (int a = 0; a < n_out; ++a) d_out[s][a] = pipe->scratch[s] + in_off[n_in] + a * out_chunk;
        if (relay >= 0) GTP_CUDA_TRY(pipe->reserve_relay(s, n_out * out_chunk, relay), "D2H relay set-up");
    }
    for (int64_t k = 0; k < n_chunks; ++k) {
        const int s = (int)(k % n_streams);
        const cudaStream_t st = pipe->streams[s];
        const int64_t s0 = k * chunk, ns = std::min(chunk, n_scans - s0);
        for (int a = 0; a < n_in; ++a)
            GTP_CUDA_TRY(cudaMemcpyAsync(d_in[s][a], (const char*)in[a] + (size_t)s0 * in_bytes[a], (size_t)ns * in_bytes[a],
                                         cudaMemcpyHostToDevice, st), "H2D");
        status = launch(d_in[s], d_out[s], ns, st);
        if (status != GTP_OK) goto done;
        if (relay < 0) {
            for (int a = 0; a < n_out; ++a)
                GTP_CUDA_TRY(cudaMemcpyAsync((char*)out[a] + (size_t)s0 * out_bytes, d_out[s][a], (size_t)ns * out_bytes,
                                             cudaMemcpyDeviceToHost, st), "D2H");
        } else {
            // results -> staging buffer on the relay GPU over NVLink (ordered on this stream), then, on the relay
            // GPU's stream, staging buffer -> host; the staging buffer is free again when `drained` fires
            if (pipe->drained_valid[s]) GTP_CUDA_TRY(cudaStreamWaitEvent(st, pipe->drained[s], 0), "relay wait");
            for (int a = 0; a < n_out; ++a)
                GTP_CUDA_TRY(cudaMemcpyPeerAsync(pipe->relay_buf[s] + a * out_chunk, relay, d_out[s][a], current,
                                                 (size_t)ns * out_bytes, st), "peer copy to the relay GPU");
            GTP_CUDA_TRY(cudaEventRecord(pipe->copied[s], st), "relay event");
            GTP_CUDA_TRY(cudaSetDevice(relay), "relay device");
            cudaError_t re = cudaStreamWaitEvent(pipe->relay_streams[s], pipe->copied[s], 0);
            for (int a = 0; a < n_out && re == cudaSuccess; ++a)
                re = cudaMemcpyAsync((char*)out[a] + (size_t)s0 * out_bytes, pipe->relay_buf[s] + a * out_chunk,
                                     (size_t)ns * out_bytes, cudaMemcpyDeviceToHost, pipe->relay_streams[s]);
            if (re == cudaSuccess) re = cudaEventRecord(pipe->drained[s], pipe->relay_streams[s]);
            const cudaError_t back = cudaSetDevice(current);
            GTP_CUDA_TRY(re, "D2H from the relay GPU");
            GTP_CUDA_TRY(back, "relay device");
            pipe->drained_valid[s] = true;
        }
    }
done:
    bool healthy = true;
    for (int s = 0; s < kStreams; ++s) {
        if (!pipe->streams[s]) continue;
        cudaError_t e = cudaStreamSynchronize(pipe->streams[s]);
        if (e == cudaSuccess && pipe->relay_streams[s]) {
            e = cudaStreamSynchronize(pipe->relay_streams[s]);      // a stream handle carries its device
            pipe->drained_valid[s] = false;
        }
        if (e != cudaSuccess) {
            healthy = false;
            if (status == GTP_OK) status = cuda_fail(e, "pipeline sync");
        }
    }
    release_pipe(pipe, healthy && status == GTP_OK);
    return status;
}

// the two-output (lon, lat) form
template <typename Launch>
int host_pipeline_bytes(const void* const* in, const size_t* in_bytes, int n_in, void* out_lon, void* out_lat,
                        size_t out_bytes, int64_t n_scans, int device, Launch launch)
{
    void* const out[2] = {out_lon, out_lat};
    return host_pipeline_n(in, in_bytes, n_in, out, 2, out_bytes, n_scans, device,
        [&](void* const* d_in, void* const* d_out, int64_t ns, cudaStream_t st) { return launch(d_in, d_out[0], d_out[1], ns, st); });
}

template <typename T, int N_IN, typename Launch>
int host_pipeline(const T* const (&in)[N_IN], T* out_lon, T* out_lat, int64_t n_scans,
                  int64_t in_elems_per_scan, int64_t out_elems_per_scan, int device, Launch launch)
{
    static_assert(N_IN <= kMaxInputs, "at most three input arrays");
    const void* in_v[N_IN];
    size_t in_bytes[N_IN];
    for (int a = 0; a < N_IN; ++a) { in_v[a] = in[a]; in_bytes[a] = (size_t)in_elems_per_scan * sizeof(T); }
    return host_pipeline_bytes(in_v, in_bytes, N_IN, out_lon, out_lat, (size_t)out_elems_per_scan * sizeof(T), n_scans, device,
        [&](void* const* d_in, void* d_lon, void* d_lat, int64_t ns, cudaStream_t st) {
            T* typed[N_IN];
            for (int a = 0; a < N_IN; ++a) typed[a] = (T*)d_in[a];
            return launch(typed, (T*)d_lon, (T*)d_lat, ns, st);
        });
}

// ---- mixed-type operator: on-disk types in, float64 math, float32 | float64 out ----
// geometry of the mixed operator: 1 km -> 250 | 500 m or 5 km (270 | 271 columns) -> 1 km
struct MixedCfg { int rows_per_scan, width, out_rows_per_scan, out_cols; };

int mixed_check(const void* lon, const void* lat, const void* satz, int zen_kind, int64_t n_scans, int coarse_res,
                int fine_res, int coarse_width, const void* out_lon, const void* out_lat, int out_kind, MixedCfg* mc)
{
    if (coarse_res == 1000 && (fine_res == 250 || fine_res == 500)) {
        const int p = 1000 / fine_res;
        *mc = MixedCfg{10, 1354, 10 * p, 1354 * p};
    } else if (coarse_res == 5000 && fine_res == 1000) {
        const int w = coarse_width ? coarse_width : 271;
        if (w != 270 && w != 271) return fail(GTP_E_WIDTH_5KM, "Can't interpolate if 5km tiepoints have less than 270 columns.%s");
        *mc = MixedCfg{2, w, 10, 1354};
    } else {
        return fail(GTP_E_RESOLUTION, "mixed operator: 1000 -> 250 | 500 or 5000 -> 1000 only%s");
    }
    if (zen_kind != GTP_ZEN_F32 && zen_kind != GTP_ZEN_I16) return fail(GTP_E_RESOLUTION, "mixed operator: unknown zenith type%s");
    if (out_kind != GTP_OUT_F32 && out_kind != GTP_OUT_F64) return fail(GTP_E_RESOLUTION, "mixed operator: unknown output type%s");
    if (n_scans < 0) return fail(GTP_E_SHAPE, "n_scans must be >= 0%s");
    if (n_scans > 0 && (!lon || !lat || !satz || !out_lon || !out_lat)) return fail(GTP_E_NULL, "null buffer%s");
    return GTP_OK;
}

int simple_mixed_check(const void* lon, const void* lat, int64_t n_scans, int res, const void* out_lon, const void* out_lat,
                       int out_kind)
{
    if (res != 2 && res != 4) return fail(GTP_E_RESOLUTION, "simple interpolator: res_factor must be 2 or 4%s");
    if (out_kind != GTP_OUT_F32 && out_kind != GTP_OUT_F64) return fail(GTP_E_RESOLUTION, "mixed operator: unknown output type%s");
    if (n_scans < 0) return fail(GTP_E_SHAPE, "n_scans must be >= 0%s");
    if (n_scans > 0 && (!lon || !lat || !out_lon || !out_lat)) return fail(GTP_E_NULL, "null buffer%s");
    return GTP_OK;
}

int simple_mixed_device(const float* lon, const float* lat, int64_t n_scans, int res, void* out_lon, void* out_lat,
                        int out_kind, int device, void* stream)
{
    int rc = simple_mixed_check(lon, lat, n_scans, res, out_lon, out_lat, out_kind);
    if (rc != GTP_OK || n_scans == 0) return rc;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) { cudaGetLastError(); return fail(GTP_E_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(guard.err)); }
    cudaError_t e = gtp_launch_simple_mixed(lon, lat, out_lon, out_lat, out_kind == GTP_OUT_F64, n_scans, res, (cudaStream_t)stream);
    if (e == cudaErrorMisalignedAddress) { cudaGetLastError(); return fail(GTP_E_SHAPE, "mixed operator: inputs must be 4-byte and outputs 16-byte (float64: 32-byte) aligned%s"); }
    if (e != cudaSuccess) return cuda_fail(e, "simple mixed launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GTP_OK;
}

int mixed_device(const float* lon, const float* lat, const void* satz, int zen_kind, double zen_scale, double zen_offset,
                 int64_t n_scans, int coarse_res, int fine_res, int coarse_width, void* out_lon, void* out_lat, int out_kind,
                 int device, void* stream)
{
    MixedCfg mc;
    int rc = mixed_check(lon, lat, satz, zen_kind, n_scans, coarse_res, fine_res, coarse_width, out_lon, out_lat, out_kind, &mc);
    if (rc != GTP_OK || n_scans == 0) return rc;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) { cudaGetLastError(); return fail(GTP_E_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(guard.err)); }
    cudaError_t e = (coarse_res == 1000)
        ? gtp_launch_cviirs_mixed(lon, lat, satz, zen_kind == GTP_ZEN_I16, zen_scale, zen_offset, out_lon, out_lat,
                                  out_kind == GTP_OUT_F64, n_scans, 1000 / fine_res, (cudaStream_t)stream)
        : gtp_launch_cviirs5_mixed(lon, lat, satz, zen_kind == GTP_ZEN_I16, zen_scale, zen_offset, out_lon, out_lat,
                                   out_kind == GTP_OUT_F64, n_scans, mc.width, (cudaStream_t)stream);
    if (e == cudaErrorMisalignedAddress) { cudaGetLastError(); return fail(GTP_E_SHAPE, "mixed operator: inputs must be 4-byte and outputs 16-byte (float64: 32-byte) aligned%s"); }
    if (e != cudaSuccess) return cuda_fail(e, "mixed launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GTP_OK;
}

template <typename T>
int cviirs_host(const T* lon, const T* lat, const T* satz, int64_t n_scans, int coarse_res, int fine_res,
                int coarse_width, T* out_lon, T* out_lat, int device)
{
    CviirsCfg cfg;
    int rc = gtp_make_cfg(coarse_res, fine_res, coarse_width, &cfg);
    if (rc) return cfg_error(rc);
    if (n_scans < 0) return fail(GTP_E_SHAPE, "n_scans must be >= 0%s");
    if (n_scans == 0) return GTP_OK;
    if (!lon || !lat || !satz || !out_lon || !out_lat) return fail(GTP_E_NULL, "null buffer%s");
    const T* const in[3] = {lon, lat, satz};
    return host_pipeline<T, 3>(in, out_lon, out_lat, n_scans, (int64_t)cfg.L * cfg.W, (int64_t)cfg.n * cfg.Wf, device,
        [&](T* const* d_in, T* d_lon, T* d_lat, int64_t ns, cudaStream_t st) {
            return cviirs_device<T>(d_in[0], d_in[1], d_in[2], ns, coarse_res, fine_res, coarse_width,
                                    d_lon, d_lat, -1, (void*)st, true);
        });
}

template <typename T>
int simple_host(const T* lon, const T* lat, int64_t n_scans, int64_t width, int res, T* out_lon, T* out_lat, int device)
{
    if (res != 2 && res != 4) return fail(GTP_E_RESOLUTION, "simple interpolator: res_factor must be 2 or 4%s");
    if (n_scans < 0) return fail(GTP_E_SHAPE, "n_scans must be >= 0%s");
    if (width < 4 || width > (1 << 24)) return fail(GTP_E_SHAPE, "simple interpolator: width must be in [4, 2^24]%s");
    if (n_scans == 0) return GTP_OK;
    if (!lon || !lat || !out_lon || !out_lat) return fail(GTP_E_NULL, "null buffer%s");
    const T* const in[2] = {lon, lat};
    return host_pipeline<T, 2>(in, out_lon, out_lat, n_scans, 10 * width, 10 * res * width * res, device,
        [&](T* const* d_in, T* d_lon, T* d_lat, int64_t ns, cudaStream_t st) {
            return simple_device<T>(d_in[0], d_in[1], ns, width, res, d_lon, d_lat, -1, (void*)st, true);
        });
}

// ---- ECEF output: x / y / z of _compute_fine_xyz instead of their back-transform ----
const char* const kMisaligned = "ECEF operator: inputs must be aligned to their element size (int16: 4 bytes) and outputs to 4 (250 m) | 2 (500 m) elements%s";

template <typename T>
int cviirs_xyz_device(const T* lon, const T* lat, const T* satz, int64_t n_scans, int coarse_res, int fine_res,
                      int coarse_width, T* out_x, T* out_y, T* out_z, int device, void* stream, bool allow_fast)
{
    CviirsCfg cfg;
    int rc = gtp_make_cfg(coarse_res, fine_res, coarse_width, &cfg);
    if (rc) return cfg_error(rc);
    if (n_scans < 0) return fail(GTP_E_SHAPE, "n_scans must be >= 0%s");
    if (n_scans == 0) return GTP_OK;
    if (!lon || !lat || !satz || !out_x || !out_y || !out_z) return fail(GTP_E_NULL, "null buffer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) { cudaGetLastError(); return fail(GTP_E_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(guard.err)); }
    cudaError_t e = cudaErrorMisalignedAddress;
    if constexpr (sizeof(T) == 8) {
        if (allow_fast && gtp_cviirs_fast_f64_supported(cfg))
            e = gtp_launch_xyz_fast(lon, lat, satz, 1, 0, 1.0, 0.0, out_x, out_y, out_z, 1, n_scans, cfg.p, (cudaStream_t)stream);
    }
    if (e == cudaErrorMisalignedAddress)       // not taken, or the vector stores do not fit: rounding-faithful kernel
        e = gtp_launch_cviirs_generic_xyz<T>(lon, lat, satz, out_x, out_y, out_z, n_scans, cfg, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "cviirs xyz launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GTP_OK;
}

template <typename T>
int simple_xyz_device(const T* lon, const T* lat, int64_t n_scans, int64_t width, int res, T* out_x, T* out_y, T* out_z,
                      int device, void* stream, bool allow_fast)
{
    if (res != 2 && res != 4) return fail(GTP_E_RESOLUTION, "simple interpolator: res_factor must be 2 or 4%s");
    if (n_scans < 0) return fail(GTP_E_SHAPE, "n_scans must be >= 0%s");
    if (width < 4 || width > (1 << 24)) return fail(GTP_E_SHAPE, "simple interpolator: width must be in [4, 2^24]%s");
    if (n_scans == 0) return GTP_OK;
    if (!lon || !lat || !out_x || !out_y || !out_z) return fail(GTP_E_NULL, "null buffer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) { cudaGetLastError(); return fail(GTP_E_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(guard.err)); }
    cudaError_t e = cudaErrorMisalignedAddress;
    if constexpr (sizeof(T) == 8) {
        if (allow_fast && gtp_simple_fast_f64_supported((int)width, res))
            e = gtp_launch_xyz_fast(lon, lat, nullptr, 1, 0, 1.0, 0.0, out_x, out_y, out_z, 1, n_scans, res, (cudaStream_t)stream);
    }
    if (e == cudaErrorMisalignedAddress)
        e = gtp_launch_simple_generic_xyz<T>(lon, lat, out_x, out_y, out_z, n_scans, (int)width, res, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "simple xyz launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GTP_OK;
}

template <typename T>
int cviirs_xyz_host(const T* lon, const T* lat, const T* satz, int64_t n_scans, int coarse_res, int fine_res,
                    int coarse_width, T* out_x, T* out_y, T* out_z, int device)
{
    CviirsCfg cfg;
    int rc = gtp_make_cfg(coarse_res, fine_res, coarse_width, &cfg);
    if (rc) return cfg_error(rc);
    if (n_scans < 0) return fail(GTP_E_SHAPE, "n_scans must be >= 0%s");
    if (n_scans == 0) return GTP_OK;
    if (!lon || !lat || !satz || !out_x || !out_y || !out_z) return fail(GTP_E_NULL, "null buffer%s");
    const void* in[3] = {lon, lat, satz};
    const size_t tie = (size_t)cfg.L * cfg.W * sizeof(T);
    const size_t in_bytes[3] = {tie, tie, tie};
    void* const out[3] = {out_x, out_y, out_z};
    return host_pipeline_n(in, in_bytes, 3, out, 3, (size_t)cfg.n * cfg.Wf * sizeof(T), n_scans, device,
        [&](void* const* d_in, void* const* d_out, int64_t ns, cudaStream_t st) {
            return cviirs_xyz_device<T>((const T*)d_in[0], (const T*)d_in[1], (const T*)d_in[2], ns, coarse_res, fine_res,
                                        coarse_width, (T*)d_out[0], (T*)d_out[1], (T*)d_out[2], -1, (void*)st, true);
        });
}

template <typename T>
int simple_xyz_host(const T* lon, const T* lat, int64_t n_scans, int64_t width, int res, T* out_x, T* out_y, T* out_z, int device)
{
    if (res != 2 && res != 4) return fail(GTP_E_RESOLUTION, "simple interpolator: res_factor must be 2 or 4%s");
    if (n_scans < 0) return fail(GTP_E_SHAPE, "n_scans must be >= 0%s");
    if (width < 4 || width > (1 << 24)) return fail(GTP_E_SHAPE, "simple interpolator: width must be in [4, 2^24]%s");
    if (n_scans == 0) return GTP_OK;
    if (!lon || !lat || !out_x || !out_y || !out_z) return fail(GTP_E_NULL, "null buffer%s");
    const void* in[2] = {lon, lat};
    const size_t tie = (size_t)10 * width * sizeof(T);
    const size_t in_bytes[2] = {tie, tie};
    void* const out[3] = {out_x, out_y, out_z};
    return host_pipeline_n(in, in_bytes, 2, out, 3, tie * res * res, n_scans, device,
        [&](void* const* d_in, void* const* d_out, int64_t ns, cudaStream_t st) {
            return simple_xyz_device<T>((const T*)d_in[0], (const T*)d_in[1], ns, width, res,
                                        (T*)d_out[0], (T*)d_out[1], (T*)d_out[2], -1, (void*)st, true);
        });
}

// on-disk types in (float32 lon / lat, float32 | int16 zenith; satz == NULL: simple interpolator), float64
// math, float32 | float64 ECEF out; 1 km tie points only
int xyz_mixed_check(const void* lon, const void* lat, const void* satz, bool simple, int zen_kind, int64_t n_scans,
                    int coarse_res, int fine_res, const void* out_x, const void* out_y, const void* out_z, int out_kind)
{
    if (coarse_res != 1000 || (fine_res != 250 && fine_res != 500))
        return fail(GTP_E_RESOLUTION, "mixed ECEF operator: 1000 -> 250 | 500 only%s");
    if (!simple && zen_kind != GTP_ZEN_F32 && zen_kind != GTP_ZEN_I16) return fail(GTP_E_RESOLUTION, "mixed operator: unknown zenith type%s");
    if (out_kind != GTP_OUT_F32 && out_kind != GTP_OUT_F64) return fail(GTP_E_RESOLUTION, "mixed operator: unknown output type%s");
    if (n_scans < 0) return fail(GTP_E_SHAPE, "n_scans must be >= 0%s");
    if (n_scans > 0 && (!lon || !lat || (!simple && !satz) || !out_x || !out_y || !out_z)) return fail(GTP_E_NULL, "null buffer%s");
    return GTP_OK;
}

int xyz_mixed_device(const float* lon, const float* lat, const void* satz, bool simple, int zen_kind, double zen_scale,
                     double zen_offset, int64_t n_scans, int coarse_res, int fine_res, void* out_x, void* out_y, void* out_z,
                     int out_kind, int device, void* stream)
{
    int rc = xyz_mixed_check(lon, lat, satz, simple, zen_kind, n_scans, coarse_res, fine_res, out_x, out_y, out_z, out_kind);
    if (rc != GTP_OK || n_scans == 0) return rc;
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) { cudaGetLastError(); return fail(GTP_E_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(guard.err)); }
    cudaError_t e = gtp_launch_xyz_fast(lon, lat, simple ? nullptr : satz, 0, zen_kind == GTP_ZEN_I16, zen_scale, zen_offset,
                                        out_x, out_y, out_z, out_kind == GTP_OUT_F64, n_scans, 1000 / fine_res, (cudaStream_t)stream);
    if (e == cudaErrorMisalignedAddress) { cudaGetLastError(); return fail(GTP_E_SHAPE, kMisaligned); }
    if (e != cudaSuccess) return cuda_fail(e, "mixed xyz launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return GTP_OK;
}

int xyz_mixed_host(const float* lon, const float* lat, const void* satz, bool simple, int zen_kind, double zen_scale,
                   double zen_offset, int64_t n_scans, int coarse_res, int fine_res, void* out_x, void* out_y, void* out_z,
                   int out_kind, int device)
{
    int rc = xyz_mixed_check(lon, lat, satz, simple, zen_kind, n_scans, coarse_res, fine_res, out_x, out_y, out_z, out_kind);
    if (rc != GTP_OK || n_scans == 0) return rc;
    const void* in[3] = {lon, lat, satz};
    const size_t tie = 10 * 1354;
    const size_t in_bytes[3] = {tie * 4, tie * 4, tie * (zen_kind == GTP_ZEN_I16 ? 2 : 4)};
    const int p = 1000 / fine_res;
    void* const out[3] = {out_x, out_y, out_z};
    return host_pipeline_n(in, in_bytes, simple ? 2 : 3, out, 3, tie * p * p * (out_kind == GTP_OUT_F64 ? 8 : 4), n_scans, device,
        [&](void* const* d_in, void* const* d_out, int64_t ns, cudaStream_t st) {
            return xyz_mixed_device((const float*)d_in[0], (const float*)d_in[1], simple ? nullptr : d_in[2], simple, zen_kind,
                                    zen_scale, zen_offset, ns, coarse_res, fine_res, d_out[0], d_out[1], d_out[2], out_kind,
                                    -1, (void*)st);
        });
}

}  // namespace

extern "C" {

const char* gtp_version(void) { return "gtp_b200 0.1.0 (sm_100a; restates python-geotiepoints 1.9.0 MODIS interpolators)"; }
const char* gtp_last_error(void) { return g_err; }
uint64_t gtp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int gtp_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int gtp_cviirs_geometry(int coarse_res, int fine_res, int coarse_width,
                        int* rows_per_scan_in, int* cols_in, int* rows_per_scan_out, int* cols_out)
{
    CviirsCfg cfg;
    int rc = gtp_make_cfg(coarse_res, fine_res, coarse_width, &cfg);
    if (rc) return cfg_error(rc);
    if (rows_per_scan_in) *rows_per_scan_in = cfg.L;
    if (cols_in) *cols_in = cfg.W;
    if (rows_per_scan_out) *rows_per_scan_out = cfg.n;
    if (cols_out) *cols_out = cfg.Wf;
    return GTP_OK;
}

int gtp_cviirs_interp_f32(const float* lon, const float* lat, const float* satz, int64_t n_scans,
                          int coarse_res, int fine_res, int coarse_width,
                          float* out_lon, float* out_lat, int device, void* stream)
{ return cviirs_device<float>(lon, lat, satz, n_scans, coarse_res, fine_res, coarse_width, out_lon, out_lat, device, stream, true); }

int gtp_cviirs_interp_f64(const double* lon, const double* lat, const double* satz, int64_t n_scans,
                          int coarse_res, int fine_res, int coarse_width,
                          double* out_lon, double* out_lat, int device, void* stream)
{ return cviirs_device<double>(lon, lat, satz, n_scans, coarse_res, fine_res, coarse_width, out_lon, out_lat, device, stream, true); }

int gtp_cviirs_interp_generic_f64(const double* lon, const double* lat, const double* satz, int64_t n_scans,
                                  int coarse_res, int fine_res, int coarse_width,
                                  double* out_lon, double* out_lat, int device, void* stream)
{ return cviirs_device<double>(lon, lat, satz, n_scans, coarse_res, fine_res, coarse_width, out_lon, out_lat, device, stream, false); }

int gtp_cviirs_interp_generic_f32(const float* lon, const float* lat, const float* satz, int64_t n_scans,
                                  int coarse_res, int fine_res, int coarse_width,
                                  float* out_lon, float* out_lat, int device, void* stream)
{ return cviirs_device<float>(lon, lat, satz, n_scans, coarse_res, fine_res, coarse_width, out_lon, out_lat, device, stream, false); }

int gtp_simple_interp_f32(const float* lon, const float* lat, int64_t n_scans, int64_t width, int res_factor,
                          float* out_lon, float* out_lat, int device, void* stream)
{ return simple_device<float>(lon, lat, n_scans, width, res_factor, out_lon, out_lat, device, stream, true); }

int gtp_simple_interp_f64(const double* lon, const double* lat, int64_t n_scans, int64_t width, int res_factor,
                          double* out_lon, double* out_lat, int device, void* stream)
{ return simple_device<double>(lon, lat, n_scans, width, res_factor, out_lon, out_lat, device, stream, true); }

int gtp_simple_interp_generic_f64(const double* lon, const double* lat, int64_t n_scans, int64_t width, int res_factor,
                                  double* out_lon, double* out_lat, int device, void* stream)
{ return simple_device<double>(lon, lat, n_scans, width, res_factor, out_lon, out_lat, device, stream, false); }

int gtp_cviirs_interp_host_f32(const float* lon, const float* lat, const float* satz, int64_t n_scans,
                               int coarse_res, int fine_res, int coarse_width, float* out_lon, float* out_lat, int device)
{ return cviirs_host<float>(lon, lat, satz, n_scans, coarse_res, fine_res, coarse_width, out_lon, out_lat, device); }

int gtp_cviirs_interp_host_f64(const double* lon, const double* lat, const double* satz, int64_t n_scans,
                               int coarse_res, int fine_res, int coarse_width, double* out_lon, double* out_lat, int device)
{ return cviirs_host<double>(lon, lat, satz, n_scans, coarse_res, fine_res, coarse_width, out_lon, out_lat, device); }

int gtp_simple_interp_host_f32(const float* lon, const float* lat, int64_t n_scans, int64_t width, int res_factor,
                               float* out_lon, float* out_lat, int device)
{ return simple_host<float>(lon, lat, n_scans, width, res_factor, out_lon, out_lat, device); }

int gtp_simple_interp_host_f64(const double* lon, const double* lat, int64_t n_scans, int64_t width, int res_factor,
                               double* out_lon, double* out_lat, int device)
{ return simple_host<double>(lon, lat, n_scans, width, res_factor, out_lon, out_lat, device); }

int gtp_simple_interp_mixed(const float* lon, const float* lat, int64_t n_scans, int res_factor,
                            void* out_lon, void* out_lat, int out_kind, int device, void* stream)
{ return simple_mixed_device(lon, lat, n_scans, res_factor, out_lon, out_lat, out_kind, device, stream); }

int gtp_simple_interp_mixed_host(const float* lon, const float* lat, int64_t n_scans, int res_factor,
                                 void* out_lon, void* out_lat, int out_kind, int device)
{
    int rc = simple_mixed_check(lon, lat, n_scans, res_factor, out_lon, out_lat, out_kind);
    if (rc != GTP_OK || n_scans == 0) return rc;
    const void* in[2] = {lon, lat};
    const size_t tie = 10 * 1354;
    const size_t in_bytes[2] = {tie * 4, tie * 4};
    const size_t out_bytes = tie * res_factor * res_factor * (out_kind == GTP_OUT_F64 ? 8 : 4);
    return host_pipeline_bytes(in, in_bytes, 2, out_lon, out_lat, out_bytes, n_scans, device,
        [&](void* const* d_in, void* d_lon, void* d_lat, int64_t ns, cudaStream_t st) {
            return simple_mixed_device((const float*)d_in[0], (const float*)d_in[1], ns, res_factor, d_lon, d_lat,
                                       out_kind, -1, (void*)st);
        });
}

int gtp_cviirs_interp_mixed(const float* lon, const float* lat, const void* satz, int zen_kind,
                            double zen_scale, double zen_offset, int64_t n_scans, int coarse_res, int fine_res,
                            int coarse_width, void* out_lon, void* out_lat, int out_kind, int device, void* stream)
{
    return mixed_device(lon, lat, satz, zen_kind, zen_scale, zen_offset, n_scans, coarse_res, fine_res, coarse_width,
                        out_lon, out_lat, out_kind, device, stream);
}

int gtp_cviirs_interp_mixed_host(const float* lon, const float* lat, const void* satz, int zen_kind,
                                 double zen_scale, double zen_offset, int64_t n_scans, int coarse_res, int fine_res,
                                 int coarse_width, void* out_lon, void* out_lat, int out_kind, int device)
{
    MixedCfg mc;
    int rc = mixed_check(lon, lat, satz, zen_kind, n_scans, coarse_res, fine_res, coarse_width, out_lon, out_lat, out_kind, &mc);
    if (rc != GTP_OK || n_scans == 0) return rc;
    const void* in[3] = {lon, lat, satz};
    const size_t tie = (size_t)mc.rows_per_scan * mc.width;
    const size_t in_bytes[3] = {tie * 4, tie * 4, tie * (zen_kind == GTP_ZEN_I16 ? 2 : 4)};
    const size_t out_bytes = (size_t)mc.out_rows_per_scan * mc.out_cols * (out_kind == GTP_OUT_F64 ? 8 : 4);
    return host_pipeline_bytes(in, in_bytes, 3, out_lon, out_lat, out_bytes, n_scans, device,
        [&](void* const* d_in, void* d_lon, void* d_lat, int64_t ns, cudaStream_t st) {
            return mixed_device((const float*)d_in[0], (const float*)d_in[1], d_in[2], zen_kind, zen_scale, zen_offset,
                                ns, coarse_res, fine_res, coarse_width, d_lon, d_lat, out_kind, -1, (void*)st);
        });
}

// ---- ECEF output ----
int gtp_cviirs_interp_xyz_f32(const float* lon, const float* lat, const float* satz, int64_t n_scans,
                              int coarse_res, int fine_res, int coarse_width,
                              float* out_x, float* out_y, float* out_z, int device, void* stream)
{ return cviirs_xyz_device<float>(lon, lat, satz, n_scans, coarse_res, fine_res, coarse_width, out_x, out_y, out_z, device, stream, true); }

int gtp_cviirs_interp_xyz_f64(const double* lon, const double* lat, const double* satz, int64_t n_scans,
                              int coarse_res, int fine_res, int coarse_width,
                              double* out_x, double* out_y, double* out_z, int device, void* stream)
{ return cviirs_xyz_device<double>(lon, lat, satz, n_scans, coarse_res, fine_res, coarse_width, out_x, out_y, out_z, device, stream, true); }

int gtp_cviirs_interp_xyz_generic_f64(const double* lon, const double* lat, const double* satz, int64_t n_scans,
                                      int coarse_res, int fine_res, int coarse_width,
                                      double* out_x, double* out_y, double* out_z, int device, void* stream)
{ return cviirs_xyz_device<double>(lon, lat, satz, n_scans, coarse_res, fine_res, coarse_width, out_x, out_y, out_z, device, stream, false); }

int gtp_simple_interp_xyz_f32(const float* lon, const float* lat, int64_t n_scans, int64_t width, int res_factor,
                              float* out_x, float* out_y, float* out_z, int device, void* stream)
{ return simple_xyz_device<float>(lon, lat, n_scans, width, res_factor, out_x, out_y, out_z, device, stream, true); }

int gtp_simple_interp_xyz_f64(const double* lon, const double* lat, int64_t n_scans, int64_t width, int res_factor,
                              double* out_x, double* out_y, double* out_z, int device, void* stream)
{ return simple_xyz_device<double>(lon, lat, n_scans, width, res_factor, out_x, out_y, out_z, device, stream, true); }

int gtp_simple_interp_xyz_generic_f64(const double* lon, const double* lat, int64_t n_scans, int64_t width, int res_factor,
                                      double* out_x, double* out_y, double* out_z, int device, void* stream)
{ return simple_xyz_device<double>(lon, lat, n_scans, width, res_factor, out_x, out_y, out_z, device, stream, false); }

int gtp_cviirs_interp_xyz_host_f32(const float* lon, const float* lat, const float* satz, int64_t n_scans,
                                   int coarse_res, int fine_res, int coarse_width,
                                   float* out_x, float* out_y, float* out_z, int device)
{ return cviirs_xyz_host<float>(lon, lat, satz, n_scans, coarse_res, fine_res, coarse_width, out_x, out_y, out_z, device); }

int gtp_cviirs_interp_xyz_host_f64(const double* lon, const double* lat, const double* satz, int64_t n_scans,
                                   int coarse_res, int fine_res, int coarse_width,
                                   double* out_x, double* out_y, double* out_z, int device)
{ return cviirs_xyz_host<double>(lon, lat, satz, n_scans, coarse_res, fine_res, coarse_width, out_x, out_y, out_z, device); }

int gtp_simple_interp_xyz_host_f32(const float* lon, const float* lat, int64_t n_scans, int64_t width, int res_factor,
                                   float* out_x, float* out_y, float* out_z, int device)
{ return simple_xyz_host<float>(lon, lat, n_scans, width, res_factor, out_x, out_y, out_z, device); }

int gtp_simple_interp_xyz_host_f64(const double* lon, const double* lat, int64_t n_scans, int64_t width, int res_factor,
                                   double* out_x, double* out_y, double* out_z, int device)
{ return simple_xyz_host<double>(lon, lat, n_scans, width, res_factor, out_x, out_y, out_z, device); }

int gtp_cviirs_interp_xyz_mixed(const float* lon, const float* lat, const void* satz, int zen_kind,
                                double zen_scale, double zen_offset, int64_t n_scans, int coarse_res, int fine_res,
                                void* out_x, void* out_y, void* out_z, int out_kind, int device, void* stream)
{
    return xyz_mixed_device(lon, lat, satz, false, zen_kind, zen_scale, zen_offset, n_scans, coarse_res, fine_res,
                            out_x, out_y, out_z, out_kind, device, stream);
}

int gtp_cviirs_interp_xyz_mixed_host(const float* lon, const float* lat, const void* satz, int zen_kind,
                                     double zen_scale, double zen_offset, int64_t n_scans, int coarse_res, int fine_res,
                                     void* out_x, void* out_y, void* out_z, int out_kind, int device)
{
    return xyz_mixed_host(lon, lat, satz, false, zen_kind, zen_scale, zen_offset, n_scans, coarse_res, fine_res,
                          out_x, out_y, out_z, out_kind, device);
}

int gtp_simple_interp_xyz_mixed(const float* lon, const float* lat, int64_t n_scans, int res_factor,
                                void* out_x, void* out_y, void* out_z, int out_kind, int device, void* stream)
{
    if (res_factor != 2 && res_factor != 4) return fail(GTP_E_RESOLUTION, "simple interpolator: res_factor must be 2 or 4%s");
    return xyz_mixed_device(lon, lat, nullptr, true, 0, 1.0, 0.0, n_scans, 1000, 1000 / res_factor,
                            out_x, out_y, out_z, out_kind, device, stream);
}

int gtp_simple_interp_xyz_mixed_host(const float* lon, const float* lat, int64_t n_scans, int res_factor,
                                     void* out_x, void* out_y, void* out_z, int out_kind, int device)
{
    if (res_factor != 2 && res_factor != 4) return fail(GTP_E_RESOLUTION, "simple interpolator: res_factor must be 2 or 4%s");
    return xyz_mixed_host(lon, lat, nullptr, true, 0, 1.0, 0.0, n_scans, 1000, 1000 / res_factor,
                          out_x, out_y, out_z, out_kind, device);
}

// ---- legacy spline interpolators ----
namespace {
// Stream-ordered scratch (cudaMallocAsync) comes from the current device's default pool; keep freed blocks in
// the pool: the default gives them back at every synchronisation, and a cudaMalloc of a few GB costs more than
// the kernels.  gtp_trim() empties the pool again.
void keep_pooled_scratch()
{
    int cur = -1;
    cudaMemPool_t pool = nullptr;
    if (cudaGetDevice(&cur) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, cur) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
}

int legacy_device(const double* lon, const double* lat, int64_t n_scans, int kind, double* out_lon, double* out_lat,
                  int device, void* stream)
{
    if (kind < GTP_LEGACY_5KM_1KM || kind > GTP_LEGACY_1KM_250M) return fail(GTP_E_RESOLUTION, "legacy interpolator: unknown kind%s");
    if (n_scans < 0) return fail(GTP_E_SHAPE, "n_scans must be >= 0%s");
    if (n_scans == 0) return GTP_OK;
    if (!lon || !lat || !out_lon || !out_lat) return fail(GTP_E_NULL, "null buffer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) { cudaGetLastError(); return fail(GTP_E_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(guard.err)); }
    keep_pooled_scratch();
    void* ws = nullptr;
    cudaError_t e = cudaMallocAsync(&ws, gtp_legacy_workspace_bytes(kind, n_scans), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "legacy workspace");
    e = gtp_launch_legacy_f64(lon, lat, out_lon, out_lat, n_scans, kind, ws, (cudaStream_t)stream);
    const cudaError_t ef = cudaFreeAsync(ws, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "legacy launch");
    if (ef != cudaSuccess) return cuda_fail(ef, "legacy workspace free");
    g_launches.fetch_add(4, std::memory_order_relaxed);
    return GTP_OK;
}
}  // namespace

int gtp_legacy_interp_f64(const double* lon, const double* lat, int64_t n_scans, int kind,
                          double* out_lon, double* out_lat, int device, void* stream)
{ return legacy_device(lon, lat, n_scans, kind, out_lon, out_lat, device, stream); }

int gtp_legacy_interp_host_f64(const double* lon, const double* lat, int64_t n_scans, int kind,
                               double* out_lon, double* out_lat, int device)
{
    if (kind < GTP_LEGACY_5KM_1KM || kind > GTP_LEGACY_1KM_250M) return fail(GTP_E_RESOLUTION, "legacy interpolator: unknown kind%s");
    if (n_scans < 0) return fail(GTP_E_SHAPE, "n_scans must be >= 0%s");
    if (n_scans == 0) return GTP_OK;
    if (!lon || !lat || !out_lon || !out_lat) return fail(GTP_E_NULL, "null buffer%s");
    int L, W, nf, Wf;
    gtp_legacy_shape(kind, &L, &W, &nf, &Wf);
    const double* const in[2] = {lon, lat};
    return host_pipeline<double, 2>(in, out_lon, out_lat, n_scans, (int64_t)L * W, (int64_t)nf * Wf, device,
        [&](double* const* d_in, double* d_lon, double* d_lat, int64_t ns, cudaStream_t st) {
            return legacy_device(d_in[0], d_in[1], ns, kind, d_lon, d_lat, -1, (void*)st);
        });
}

// ---- VII tie-point interpolation ----
namespace {
int vii_check(const void* a, const void* out, int64_t n_alt, int64_t n_act, int S, int F)
{
    if (S < 2 || F < 1 || n_act < 2 || n_alt < 0) return fail(GTP_E_SHAPE, "VII: need scan_alt_tie_points >= 2, tie_points_factor >= 1, >= 2 tie columns%s");
    if (n_alt % S != 0) return fail(GTP_E_SHAPE, "The number of tie points in the along-route dimension must be a multiple of scan_alt_tie_points%s");
    if (n_alt > 0 && (!a || !out)) return fail(GTP_E_NULL, "null buffer%s");
    return GTP_OK;
}

// cartesian < 0: one array (b, out_b unused); 0: two arrays interpolated as they are; 1: lon / lat through xyz
int vii_device(const double* a, const double* b, int64_t n_alt, int64_t n_act, int S, int F, int cartesian, double z_threshold,
               double* out_a, double* out_b, int device, void* stream)
{
    int rc = vii_check(a, out_a, n_alt, n_act, S, F);
    if (rc != GTP_OK || n_alt == 0) return rc;
    if (cartesian >= 0 && (!b || !out_b)) return fail(GTP_E_NULL, "null buffer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) { cudaGetLastError(); return fail(GTP_E_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(guard.err)); }
    cudaError_t e;
    if (cartesian == 1) {
        keep_pooled_scratch();
        void* ws = nullptr;
        e = cudaMallocAsync(&ws, gtp_vii_workspace_bytes(n_alt, n_act), (cudaStream_t)stream);
        if (e != cudaSuccess) return cuda_fail(e, "VII workspace");
        e = gtp_launch_vii_geo_cartesian(a, b, out_a, out_b, n_alt, n_act, S, F, z_threshold, ws, (cudaStream_t)stream);
        cudaFreeAsync(ws, (cudaStream_t)stream);
        g_launches.fetch_add(2, std::memory_order_relaxed);
    } else {
        e = gtp_launch_vii_interp(a, cartesian < 0 ? nullptr : b, out_a, cartesian < 0 ? nullptr : out_b, n_alt, n_act, S, F, (cudaStream_t)stream);
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    if (e != cudaSuccess) return cuda_fail(e, "VII launch");
    return GTP_OK;
}

int vii_host(const double* a, const double* b, int64_t n_alt, int64_t n_act, int S, int F, int cartesian, double z_threshold,
             double* out_a, double* out_b, int device)
{
    int rc = vii_check(a, out_a, n_alt, n_act, S, F);
    if (rc != GTP_OK || n_alt == 0) return rc;
    if (cartesian >= 0 && (!b || !out_b)) return fail(GTP_E_NULL, "null buffer%s");
    const int64_t n_scans = n_alt / S, in_per_scan = (int64_t)S * n_act, out_per_scan = (int64_t)(S - 1) * F * (n_act - 1) * F;
    if (cartesian < 0) {
        const void* in_v[1] = {a};
        const size_t in_b[1] = {(size_t)in_per_scan * 8};
        void* const out_v[1] = {out_a};
        return host_pipeline_n(in_v, in_b, 1, out_v, 1, (size_t)out_per_scan * 8, n_scans, device,
            [&](void* const* d_in, void* const* d_out, int64_t ns, cudaStream_t st) {
                return vii_device((const double*)d_in[0], nullptr, ns * S, n_act, S, F, -1, z_threshold, (double*)d_out[0], nullptr, -1, (void*)st);
            });
    }
    const double* const in[2] = {a, b};
    return host_pipeline<double, 2>(in, out_a, out_b, n_scans, in_per_scan, out_per_scan, device,
        [&](double* const* d_in, double* d_a, double* d_b, int64_t ns, cudaStream_t st) {
            return vii_device(d_in[0], d_in[1], ns * S, n_act, S, F, cartesian, z_threshold, d_a, d_b, -1, (void*)st);
        });
}
}  // namespace

int gtp_vii_interp_f64(const double* data, int64_t n_tie_alt, int64_t n_tie_act, int scan_alt_tie_points, int tie_points_factor,
                       double* out, int device, void* stream)
{ return vii_device(data, nullptr, n_tie_alt, n_tie_act, scan_alt_tie_points, tie_points_factor, -1, 0.8, out, nullptr, device, stream); }

int gtp_vii_interp_host_f64(const double* data, int64_t n_tie_alt, int64_t n_tie_act, int scan_alt_tie_points, int tie_points_factor,
                            double* out, int device)
{ return vii_host(data, nullptr, n_tie_alt, n_tie_act, scan_alt_tie_points, tie_points_factor, -1, 0.8, out, nullptr, device); }

int gtp_vii_geo_interp_f64(const double* lon, const double* lat, int64_t n_tie_alt, int64_t n_tie_act, int scan_alt_tie_points,
                           int tie_points_factor, int cartesian, double z_threshold, double* out_lon, double* out_lat,
                           int device, void* stream)
{ return vii_device(lon, lat, n_tie_alt, n_tie_act, scan_alt_tie_points, tie_points_factor, cartesian ? 1 : 0, z_threshold, out_lon, out_lat, device, stream); }

int gtp_vii_geo_interp_host_f64(const double* lon, const double* lat, int64_t n_tie_alt, int64_t n_tie_act, int scan_alt_tie_points,
                                int tie_points_factor, int cartesian, double z_threshold, double* out_lon, double* out_lat, int device)
{ return vii_host(lon, lat, n_tie_alt, n_tie_act, scan_alt_tie_points, tie_points_factor, cartesian ? 1 : 0, z_threshold, out_lon, out_lat, device); }

// ---- generic rectilinear-grid interpolators (third item) ----
namespace {
int grid_check(int64_t ny, int64_t nx, int n_arrays, int64_t my, int64_t mx, int method_y, int method_x, int oob_mode)
{
    auto known = [](int m) { return m == GTP_GRID_NEAREST || m == GTP_GRID_LINEAR || m == GTP_GRID_CUBIC; };
    if (!known(method_y) || !known(method_x)) return fail(GTP_E_RESOLUTION, "grid interpolator: unknown scheme%s");
    if (oob_mode < GTP_GRID_OOB_EXTRAPOLATE || oob_mode > GTP_GRID_OOB_CLAMP) return fail(GTP_E_RESOLUTION, "grid interpolator: unknown out-of-range mode%s");
    if (n_arrays < 1 || my < 0 || mx < 0) return fail(GTP_E_SHAPE, "grid interpolator: negative size%s");
    if (ny < 2 || nx < 2 || ny > 0x7fffffff || nx > 0x7fffffff) return fail(GTP_E_SHAPE, "grid interpolator: need at least 2 tie points per dimension%s");
    if ((method_y == GTP_GRID_CUBIC && ny < 4) || (method_x == GTP_GRID_CUBIC && nx < 4))
        return fail(GTP_E_SHAPE, "grid interpolator: the cubic scheme needs at least 4 tie points%s");
    return GTP_OK;
}

// geo: values = lon, values2 = lat, out / out2 = lon / lat
int grid_device(const double* tie_y, int64_t ny, const double* tie_x, int64_t nx, const double* values, const double* values2,
                int n_arrays, bool geo, const double* fine_y, int64_t my, const double* fine_x, int64_t mx, int method_y,
                int method_x, int oob_mode, double fill, double* out, double* out2, int device, void* stream)
{
    int rc = grid_check(ny, nx, n_arrays, my, mx, method_y, method_x, oob_mode);
    if (rc != GTP_OK) return rc;
    if (my == 0 || mx == 0) return GTP_OK;
    if (!tie_y || !tie_x || !values || !fine_y || !fine_x || !out || (geo && (!values2 || !out2))) return fail(GTP_E_NULL, "null buffer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) { cudaGetLastError(); return fail(GTP_E_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(guard.err)); }
    keep_pooled_scratch();
    void* ws = nullptr;
    cudaError_t e = cudaMallocAsync(&ws, gtp_grid_workspace_bytes((int)ny, (int)nx, my, mx, geo ? 3 : n_arrays, geo, method_y, method_x), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "grid workspace");
    int launches = 0;
    e = gtp_launch_grid(tie_y, (int)ny, tie_x, (int)nx, values, values2, n_arrays, geo, fine_y, my, fine_x, mx, method_y, method_x,
                        oob_mode, fill, out, out2, ws, (cudaStream_t)stream, &launches);
    const cudaError_t ef = cudaFreeAsync(ws, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "grid launch");
    if (ef != cudaSuccess) return cuda_fail(ef, "grid workspace free");
    g_launches.fetch_add((uint64_t)launches, std::memory_order_relaxed);
    return GTP_OK;
}

// everything staged through stream-ordered device memory on a stream of its own; returns when the results are in `out`
int grid_host(const double* tie_y, int64_t ny, const double* tie_x, int64_t nx, const double* values, const double* values2,
              int n_arrays, bool geo, const double* fine_y, int64_t my, const double* fine_x, int64_t mx, int method_y,
              int method_x, int oob_mode, double fill, double* out, double* out2, int device)
{
    int rc = grid_check(ny, nx, n_arrays, my, mx, method_y, method_x, oob_mode);
    if (rc != GTP_OK) return rc;
    if (my == 0 || mx == 0) return GTP_OK;
    if (!tie_y || !tie_x || !values || !fine_y || !fine_x || !out || (geo && (!values2 || !out2))) return fail(GTP_E_NULL, "null buffer%s");
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) { cudaGetLastError(); return fail(GTP_E_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(guard.err)); }
    keep_pooled_scratch();
    cudaStream_t st = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e != cudaSuccess) return cuda_fail(e, "grid stream");
    const size_t tie_b = (size_t)ny * nx * 8, out_b = (size_t)my * mx * 8;
    const int n_in = geo ? 2 : n_arrays, n_out = geo ? 2 : n_arrays;
    const size_t total = (size_t)(ny + nx + my + mx) * 8 + 64 + (size_t)n_in * tie_b + (size_t)n_out * out_b;
    char* base = nullptr;
    e = cudaMallocAsync((void**)&base, total, st);
    if (e != cudaSuccess) { cudaStreamDestroy(st); return cuda_fail(e, "grid staging"); }
    char* p = base;
    auto put = [&](const void* src, size_t bytes) -> double* {
        double* d = (double*)p;
        if (e == cudaSuccess) e = cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st);
        p += (bytes + 15) & ~(size_t)15;
        return d;
    };
    const double* d_ty = put(tie_y, (size_t)ny * 8);
    const double* d_tx = put(tie_x, (size_t)nx * 8);
    const double* d_fy = put(fine_y, (size_t)my * 8);
    const double* d_fx = put(fine_x, (size_t)mx * 8);
    const double *d_v = nullptr, *d_v2 = nullptr;
    if (geo) { d_v = put(values, tie_b); d_v2 = put(values2, tie_b); }
    else d_v = put(values, (size_t)n_arrays * tie_b);
    double* d_out = (double*)p;
    double* d_out2 = geo ? d_out + (size_t)my * mx : nullptr;
    if (e == cudaSuccess) {
        rc = grid_device(d_ty, ny, d_tx, nx, d_v, d_v2, n_arrays, geo, d_fy, my, d_fx, mx, method_y, method_x, oob_mode, fill,
                         d_out, d_out2, -1, (void*)st);
        if (rc == GTP_OK) {
            if (geo) {
                e = cudaMemcpyAsync(out, d_out, out_b, cudaMemcpyDeviceToHost, st);
                if (e == cudaSuccess) e = cudaMemcpyAsync(out2, d_out2, out_b, cudaMemcpyDeviceToHost, st);
            } else e = cudaMemcpyAsync(out, d_out, (size_t)n_arrays * out_b, cudaMemcpyDeviceToHost, st);
        }
    }
    cudaFreeAsync(base, st);
    const cudaError_t es = cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    if (rc != GTP_OK) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "grid copy");
    if (es != cudaSuccess) return cuda_fail(es, "grid synchronize");
    return GTP_OK;
}
}  // namespace

int gtp_grid_interp_f64(const double* tie_y, int64_t n_tie_y, const double* tie_x, int64_t n_tie_x, const double* values, int n_arrays,
                        const double* fine_y, int64_t n_fine_y, const double* fine_x, int64_t n_fine_x, int method_y, int method_x,
                        int oob_mode, double fill_value, double* out, int device, void* stream)
{ return grid_device(tie_y, n_tie_y, tie_x, n_tie_x, values, nullptr, n_arrays, false, fine_y, n_fine_y, fine_x, n_fine_x, method_y, method_x, oob_mode, fill_value, out, nullptr, device, stream); }

int gtp_grid_interp_host_f64(const double* tie_y, int64_t n_tie_y, const double* tie_x, int64_t n_tie_x, const double* values, int n_arrays,
                             const double* fine_y, int64_t n_fine_y, const double* fine_x, int64_t n_fine_x, int method_y, int method_x,
                             int oob_mode, double fill_value, double* out, int device)
{ return grid_host(tie_y, n_tie_y, tie_x, n_tie_x, values, nullptr, n_arrays, false, fine_y, n_fine_y, fine_x, n_fine_x, method_y, method_x, oob_mode, fill_value, out, nullptr, device); }

int gtp_geogrid_interp_f64(const double* tie_y, int64_t n_tie_y, const double* tie_x, int64_t n_tie_x, const double* lon, const double* lat,
                           const double* fine_y, int64_t n_fine_y, const double* fine_x, int64_t n_fine_x, int method_y, int method_x,
                           int oob_mode, double fill_value, double* out_lon, double* out_lat, int device, void* stream)
{ return grid_device(tie_y, n_tie_y, tie_x, n_tie_x, lon, lat, 1, true, fine_y, n_fine_y, fine_x, n_fine_x, method_y, method_x, oob_mode, fill_value, out_lon, out_lat, device, stream); }

int gtp_geogrid_interp_host_f64(const double* tie_y, int64_t n_tie_y, const double* tie_x, int64_t n_tie_x, const double* lon, const double* lat,
                                const double* fine_y, int64_t n_fine_y, const double* fine_x, int64_t n_fine_x, int method_y, int method_x,
                                int oob_mode, double fill_value, double* out_lon, double* out_lat, int device)
{ return grid_host(tie_y, n_tie_y, tie_x, n_tie_x, lon, lat, 1, true, fine_y, n_fine_y, fine_x, n_fine_x, method_y, method_x, oob_mode, fill_value, out_lon, out_lat, device); }

int gtp_set_d2h_relay(int device)
{
    if (device >= 0) {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return fail(GTP_E_NO_DEVICE, "no usable CUDA device%s"); }
        if (device >= n) return fail(GTP_E_SHAPE, "gtp_set_d2h_relay: no such device%s");
    }
    g_relay_device.store(device < 0 ? -1 : device, std::memory_order_relaxed);
    return GTP_OK;
}

int gtp_get_d2h_relay(void) { return g_relay_device.load(std::memory_order_relaxed); }

int gtp_trim(void)
{
    std::vector<HostPipe*> pipes;
    {
        std::lock_guard<std::mutex> lock(g_pipe_mutex);
        pipes.swap(g_idle_pipes);
    }
    int status = GTP_OK;
    {   // the stream-ordered scratch of the legacy interpolators (default pool of every device)
        int n = 0, cur = -1;
        if (cudaGetDeviceCount(&n) == cudaSuccess && cudaGetDevice(&cur) == cudaSuccess) {
            for (int d = 0; d < n; ++d) {
                cudaMemPool_t pool = nullptr;
                if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
            }
        }
        cudaGetLastError();
    }
    for (HostPipe* p : pipes) {
        DeviceGuard guard(p->device);
        if (guard.err != cudaSuccess) { cudaGetLastError(); status = fail(GTP_E_NO_DEVICE, "no usable CUDA device: %s", cudaGetErrorString(guard.err)); }
        else p->destroy();
        delete p;
    }
    return status;
}

int gtp_host_alloc(void** ptr, uint64_t bytes)
{
    if (!ptr) return fail(GTP_E_NULL, "null pointer%s");
    cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) return cuda_fail(e, "cudaHostAlloc");
    return GTP_OK;
}

int gtp_host_free(void* ptr)
{
    if (!ptr) return GTP_OK;
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFreeHost");
    return GTP_OK;
}

}  // extern "C"
