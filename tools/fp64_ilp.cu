// fp64_ilp.cu - FP64 FMA throughput per SM as a function of resident warps and independent
// chains per thread (how much ILP x TLP the FP64 pipe of sm_100a needs to saturate).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, double b, double c, int iters)
{
    double a[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) a[k] = 1.0 + threadIdx.x * 1e-9 + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) a[k] = fma(a[k], b, c);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += a[k];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
void run(int sms, int warps_per_sm, double ghz)
{
    double* out; cudaMalloc(&out, 8);
    const int iters = 1 << 14;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain<ILP><<<sms, warps_per_sm * 32>>>(out, 1.0000001, 1e-9, iters);
    cudaEventRecord(e0);
    chain<ILP><<<sms, warps_per_sm * 32>>>(out, 1.0000001, 1e-9, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)sms * warps_per_sm * 32 * iters * ILP;
    double per_clk = ops / (ms * 1e-3) / sms / (ghz * 1e9);
    printf("warps/SM %2d  ILP %d : %6.2f DFMA/clk/SM   (%.1f cycles per dependent DFMA per warp-chain)\n",
           warps_per_sm, ILP, per_clk, (double)warps_per_sm * 32 * ILP / per_clk / (warps_per_sm * ILP));
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6;
    for (int w : {4, 8, 12, 16, 24, 32}) {
        run<1>(p.multiProcessorCount, w, ghz);
        run<2>(p.multiProcessorCount, w, ghz);
        run<4>(p.multiProcessorCount, w, ghz);
        run<8>(p.multiProcessorCount, w, ghz);
    }
    return 0;
}
