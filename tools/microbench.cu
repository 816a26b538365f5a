// microbench.cu - per-SM instruction throughput of the pipes the interpolation kernels lean on
// (FP64 FMA, FP32 FMA, MUFU, F2F conversions, 64-bit shuffles) and a plain HBM write/copy rate.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void pipe_kernel(double* out, double seed)
{
    double a[ILP], g[ILP]; float f[ILP];
    for (int k = 0; k < ILP; ++k) { a[k] = seed + threadIdx.x * 1e-9 + k; f[k] = (float)a[k]; g[k] = a[k] * 0.5; }
    double b = 1.0000001, c = 1e-9; float fb = 1.0000001f, fc = 1e-9f;
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (OP == 0) a[k] = fma(a[k], b, c);                       // DFMA
            if (OP == 1) f[k] = fmaf(f[k], fb, fc);                    // FFMA
            if (OP == 2) f[k] = __frcp_rn(f[k]) + fc;                  // MUFU.RCP (+FADD)
            if (OP == 3) a[k] = (double)((float)a[k]) + c;             // F2F.F32.F64 + F2F.F64.F32 + DADD
            if (OP == 4) a[k] = __shfl_down_sync(0xffffffffu, a[k], 1) + c;  // 2x SHFL + DADD
            if (OP == 5) a[k] = a[k] * b;                              // DMUL
            if (OP == 6) a[k] = sqrt(a[k]) + b;                        // DSQRT sequence
            if (OP == 7) a[k] = b / a[k] + c;                          // DDIV sequence
            if (OP == 8) { a[k] = fma(a[k], b, c); a[k] = fma(a[k], b, c); a[k] = fma(a[k], b, c); a[k] = fma(a[k], b, c); }  // 4 DFMA
            if (OP == 9) { f[k] = (float)a[k]; a[k] = (double)(f[k] * fb); }   // F2F pair + FMUL (dependent)
            if (OP == 10) {                                            // 4 DFMA on a[] + independent F2F pair on g
                a[k] = fma(a[k], b, c); a[k] = fma(a[k], b, c); a[k] = fma(a[k], b, c); a[k] = fma(a[k], b, c);
                f[k] = (float)g[k]; g[k] = (double)(f[k] * fb);
            }
            if (OP == 11) { g[k] = (double)(float)g[k] * b; }          // F2F pair + DMUL
        }
    }
    double s = 0; for (int k = 0; k < ILP; ++k) s += a[k] + f[k] + g[k];
    if (s == 12345.678) out[0] = s;
}

__global__ void write_kernel(double4* out, size_t n)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        double4 v = make_double4(1.0, 2.0, 3.0, (double)i);
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(out + i), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
    }
}
__global__ void write16_kernel(double2* out, size_t n)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = make_double2(1.0, (double)i);
}
__global__ void copy_kernel(const double2* in, double2* out, size_t n)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}

template <int OP>
void run_pipe(const char* name, int sms, double clock_ghz, double ops_per_iter)
{
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * 4, threads = 512;
    pipe_kernel<OP><<<blocks, threads>>>(out, 1.0);
    cudaEventRecord(e0);
    pipe_kernel<OP><<<blocks, threads>>>(out, 1.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * ITERS * ILP * ops_per_iter;
    printf("%-34s %8.3f ms  %8.2f Gop/s  %6.1f op/clk/SM (at %.3f GHz)\n", name, ms, ops / ms * 1e-6,
           ops / (ms * 1e-3) / sms / (clock_ghz * 1e9), clock_ghz);
    cudaFree(out);
}

// `microbench loop <op> <seconds>`: keep one pipe (or the HBM write path) busy for a while so that
// tools/power_probe.py can read the sustained clock and power through NVML.
template <int OP>
void loop_pipe(int sms, double seconds)
{
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    float ms = 0; long launches = 0;
    while (ms < seconds * 1e3) {
        for (int i = 0; i < 16; ++i) pipe_kernel<OP><<<sms * 4, 512>>>(out, 1.0);
        launches += 16;
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    double ops = (double)launches * sms * 4 * 512 * ITERS * ILP * (OP == 8 ? 4 : 1);
    printf("loop op %d: %.1f Gop/s over %.2f s\n", OP, ops / ms * 1e-6, ms * 1e-3);
}

void loop_write(int sms, double seconds)
{
    size_t bytes = 4ull << 30; void* a; cudaMalloc(&a, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    float ms = 0; long launches = 0;
    while (ms < seconds * 1e3) {
        for (int i = 0; i < 4; ++i) write_kernel<<<sms * 8, 512>>>((double4*)a, bytes / 32);
        launches += 4;
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("loop write: %.1f GB/s over %.2f s\n", (double)launches * bytes / ms * 1e-6, ms * 1e-3);
}

int main(int argc, char** argv)
{
    if (argc == 4 && !strcmp(argv[1], "loop")) {
        cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
        const double seconds = atof(argv[3]);
        if (!strcmp(argv[2], "dfma")) loop_pipe<8>(p.multiProcessorCount, seconds);
        else if (!strcmp(argv[2], "ffma")) loop_pipe<1>(p.multiProcessorCount, seconds);
        else if (!strcmp(argv[2], "write")) loop_write(p.multiProcessorCount, seconds);
        return 0;
    }
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double ghz = clk_khz * 1e-6;
    printf("%s  SMs=%d  max clock %.3f GHz (op/clk figures assume the max clock; real clocks are lower under load)\n",
           p.name, p.multiProcessorCount, ghz);
    run_pipe<0>("DFMA", p.multiProcessorCount, ghz, 1);
    run_pipe<5>("DMUL", p.multiProcessorCount, ghz, 1);
    run_pipe<1>("FFMA", p.multiProcessorCount, ghz, 1);
    run_pipe<2>("MUFU.RCP+FADD", p.multiProcessorCount, ghz, 1);
    run_pipe<3>("F2F d->f->d + DADD", p.multiProcessorCount, ghz, 1);
    run_pipe<4>("SHFL.64 + DADD", p.multiProcessorCount, ghz, 1);
    run_pipe<6>("sqrt(double)+DADD", p.multiProcessorCount, ghz, 1);
    run_pipe<7>("div(double)+DADD", p.multiProcessorCount, ghz, 1);
    run_pipe<8>("4x DFMA", p.multiProcessorCount, ghz, 1);
    run_pipe<9>("F2F d->f, FMUL, F2F f->d", p.multiProcessorCount, ghz, 1);
    run_pipe<10>("4x DFMA + independent F2F pair", p.multiProcessorCount, ghz, 1);
    run_pipe<11>("F2F pair + DMUL", p.multiProcessorCount, ghz, 1);

    size_t bytes = 4ull << 30;
    void *a, *b; cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); write_kernel<<<p.multiProcessorCount * 8, 512>>>((double4*)a, bytes / 32); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("write 256-bit stores   %8.3f ms  %8.1f GB/s\n", ms, bytes / ms * 1e-6);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); write16_kernel<<<p.multiProcessorCount * 8, 512>>>((double2*)a, bytes / 16); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("write 128-bit stores   %8.3f ms  %8.1f GB/s\n", ms, bytes / ms * 1e-6);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); copy_kernel<<<p.multiProcessorCount * 8, 512>>>((const double2*)a, (double2*)b, bytes / 16); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("copy  128-bit          %8.3f ms  %8.1f GB/s (read+write)\n", ms, 2.0 * bytes / ms * 1e-6);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); cudaMemsetAsync(a, 0, bytes); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    printf("cudaMemset             %8.3f ms  %8.1f GB/s\n", ms, bytes / ms * 1e-6);
    return 0;
}
