// fp64_rate.cu -- developer microbenchmark: per-SM issue rates of the instructions the
// autocorrelation kernel is made of (DFMA, DADD+DMUL, I2F.F64), to know what its roofline is.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_rate fp64_rate.cu && ./fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* out, int iters, double seed, int iseed)
{
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + i + threadIdx.x;
    double x = seed * 0.5;
    int iv = iseed + threadIdx.x;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = __fma_rn(a[i], x, 1.0);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = __dadd_rn(a[i], __dmul_rn(a[i], x));
        } else if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) { a[i] = __dadd_rn(a[i], (double)(iv + i)); }
            iv = iv * 3 + 1;
        } else if (MODE == 3) { // one dependent chain: latency
            a[0] = __fma_rn(a[0], x, 1.0);
            a[0] = __fma_rn(a[0], x, 1.0);
            a[0] = __fma_rn(a[0], x, 1.0);
            a[0] = __fma_rn(a[0], x, 1.0);
            a[0] = __fma_rn(a[0], x, 1.0);
            a[0] = __fma_rn(a[0], x, 1.0);
            a[0] = __fma_rn(a[0], x, 1.0);
            a[0] = __fma_rn(a[0], x, 1.0);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int warps_per_sm, double ops_per_iter)
{
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * 2048);
    const int iters = 20000;
    k<MODE><<<sms, warps_per_sm * 32>>>(out, 10, 1.000001, 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms, warps_per_sm * 32>>>(out, iters, 1.000001, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * khz * 1e3;
    const double lane_ops = (double)iters * ops_per_iter * warps_per_sm * 32;
    printf("%-28s warps/SM %2d: %.1f lane-ops/clk/SM  (%.1f cycles per warp-iteration)\n", name, warps_per_sm,
           lane_ops / cycles, cycles / iters);
    cudaFree(out);
}

int main()
{
    for (int w : {4, 8, 16, 32}) {
        run<0>("DFMA x8 chains", w, 8);
        run<1>("DMUL+DADD x8 chains", w, 16);
        run<2>("I2F.F64+DADD x8", w, 16);
    }
    run<3>("DFMA dependent chain of 8", 4, 8);
    return 0;
}
