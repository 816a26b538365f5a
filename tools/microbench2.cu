// microbench2.cu - round 2: can the FP32 pipe (scalar and packed f32x2), the FP64 pipe and the conversion
// unit work side by side?  The float32-faithful kernels want the blend on the FP32 pipe (error-free
// transforms) and the back-transform on the FP64 pipe, so what matters is the rate of MIXED streams.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench2 tools/microbench2.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITERS 2048
#define ILP 8

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// OP: 0 FFMA  1 FFMA2  2 FADD2  3 DFMA  4 DFMA + 2 FFMA  5 DFMA + 2 FFMA2  6 DFMA + 4 FFMA  7 DFMA + 4 FFMA2
//     8 LOP3/IADD pair  9 DFMA + 4 INT  10 F2F(f->d) + DFMA + 4 FFMA  11 F2F f->d alone  12 F2F d->f alone
//     13 2 DFMA + 6 FFMA2 + 1 F2F (the mix the float32 kernel would issue per ~1/12 pixel)  14 I2F.F64 alone
template <int OP>
__global__ void pipe_kernel(double* out, double seed, float fseed)
{
    double a[ILP]; float f[ILP], g[ILP]; unsigned long long p[ILP]; unsigned u[ILP];
    for (int k = 0; k < ILP; ++k) {
        a[k] = seed + threadIdx.x * 1e-9 + k; f[k] = fseed + k; g[k] = fseed * 0.5f + k;
        float2 t = make_float2(f[k], g[k]); p[k] = *(unsigned long long*)&t; u[k] = threadIdx.x + k;
    }
    const double b = 1.0000001, c = 1e-9; const float fb = 1.0000001f, fc = 1e-9f;
    float2 tb = make_float2(fb, fb), tc = make_float2(fc, fc);
    const unsigned long long pb = *(unsigned long long*)&tb, pc = *(unsigned long long*)&tc;
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (OP == 0) f[k] = fmaf(f[k], fb, fc);
            if (OP == 1) p[k] = ffma2(p[k], pb, pc);
            if (OP == 2) p[k] = fadd2(p[k], pc);
            if (OP == 3) a[k] = fma(a[k], b, c);
            if (OP == 4) { a[k] = fma(a[k], b, c); f[k] = fmaf(f[k], fb, fc); g[k] = fmaf(g[k], fb, fc); }
            if (OP == 5) { a[k] = fma(a[k], b, c); p[k] = ffma2(p[k], pb, pc); p[k] = ffma2(p[k], pb, pc); }
            if (OP == 6) { a[k] = fma(a[k], b, c); f[k] = fmaf(f[k], fb, fc); g[k] = fmaf(g[k], fb, fc);
                           f[k] = fmaf(f[k], fb, fc); g[k] = fmaf(g[k], fb, fc); }
            if (OP == 7) { a[k] = fma(a[k], b, c); p[k] = ffma2(p[k], pb, pc); p[k] = ffma2(p[k], pb, pc);
                           p[k] = ffma2(p[k], pb, pc); p[k] = ffma2(p[k], pb, pc); }
            if (OP == 8) { u[k] = (u[k] ^ 0x5bd1e995u) + 0x9e3779b9u; }
            if (OP == 9) { a[k] = fma(a[k], b, c); u[k] = (u[k] ^ 0x5bd1e995u) + 0x9e3779b9u; u[k] = (u[k] ^ 0x1bd1e995u) + 0x7e3779b9u; }
            if (OP == 10) { a[k] = fma(a[k], b, (double)f[k]); f[k] = fmaf(f[k], fb, fc); g[k] = fmaf(g[k], fb, fc);
                            f[k] = fmaf(f[k], fb, fc); g[k] = fmaf(g[k], fb, fc); }
            if (OP == 11) { a[k] = (double)f[k]; f[k] = __double_as_longlong(a[k]) & 0xff ? f[k] : fc; }
            if (OP == 12) { f[k] = (float)a[k]; a[k] = __hiloint2double(__float_as_int(f[k]), __double2loint(a[k])); }
            if (OP == 13) { a[k] = fma(a[k], b, (double)f[k]); a[k] = fma(a[k], b, c);
                            p[k] = ffma2(p[k], pb, pc); p[k] = ffma2(p[k], pb, pc); p[k] = ffma2(p[k], pb, pc);
                            p[k] = ffma2(p[k], pb, pc); p[k] = ffma2(p[k], pb, pc); p[k] = ffma2(p[k], pb, pc);
                            f[k] = __uint_as_float((unsigned)p[k]); }
            if (OP == 14) { a[k] = (double)(int)u[k]; u[k] = (unsigned)__double2loint(a[k]) + 3u; }
        }
    }
    double s = 0;
    for (int k = 0; k < ILP; ++k) s += a[k] + f[k] + g[k] + (double)p[k] + u[k];
    if (s == 12345.678) out[0] = s;
}

template <int OP>
void run_pipe(const char* name, int sms, double ghz, double lane_ops)
{
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 4, threads = 512;
    pipe_kernel<OP><<<blocks, threads>>>(out, 1.0, 1.5f);
    cudaEventRecord(e0);
    pipe_kernel<OP><<<blocks, threads>>>(out, 1.0, 1.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double groups = (double)blocks * threads * ITERS * ILP;
    printf("%-44s %8.3f ms  %7.2f groups/clk/SM  (%5.1f lane-ops/clk/SM at %.3f GHz)\n", name, ms,
           groups / (ms * 1e-3) / sms / (ghz * 1e9), groups * lane_ops / (ms * 1e-3) / sms / (ghz * 1e9), ghz);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz * 1e-6; const int sms = p.multiProcessorCount;
    printf("%s SMs=%d clock %.3f GHz; 'groups' = one iteration of the listed mix per thread\n", p.name, sms, ghz);
    run_pipe<0>("FFMA", sms, ghz, 1);
    run_pipe<1>("FFMA2 (2 lanes each)", sms, ghz, 2);
    run_pipe<2>("FADD2 (2 lanes each)", sms, ghz, 2);
    run_pipe<3>("DFMA", sms, ghz, 1);
    run_pipe<4>("DFMA + 2 FFMA", sms, ghz, 3);
    run_pipe<5>("DFMA + 2 FFMA2", sms, ghz, 5);
    run_pipe<6>("DFMA + 4 FFMA", sms, ghz, 5);
    run_pipe<7>("DFMA + 4 FFMA2", sms, ghz, 9);
    run_pipe<8>("LOP3 + IADD", sms, ghz, 2);
    run_pipe<9>("DFMA + 2 (LOP3 + IADD)", sms, ghz, 5);
    run_pipe<10>("F2F f->d + DFMA + 4 FFMA", sms, ghz, 6);
    run_pipe<11>("F2F f->d (+ LOP/select)", sms, ghz, 1);
    run_pipe<12>("F2F d->f (+ repack)", sms, ghz, 1);
    run_pipe<13>("F2F f->d + 2 DFMA + 6 FFMA2", sms, ghz, 15);
    run_pipe<14>("I2F.F64 + IADD", sms, ghz, 1);
    return 0;
}
