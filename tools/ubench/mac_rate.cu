// mac_rate.cu -- developer microbenchmark: multiply-accumulate rates per SM for the LPC residual of wide samples:
// 32-bit IMAD, IMAD.WIDE (32 x 32 + 64), the same sum as two 32-bit IMAD chains over 12-bit halves, and DFMA.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mac_rate mac_rate.cu && ./mac_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(long long* out, int iters, int seed)
{
    int q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = seed * (i + 3) + threadIdx.x;
    long long a64[8];
    int a32[8], b32[8];
    double ad[8], qd[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a64[i] = i; a32[i] = i; b32[i] = i; ad[i] = i; qd[i] = (double)q[i]; }
    int w = seed;
    double wd = (double)seed;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a32[i] += q[i] * w;                                          // IMAD
            else if (MODE == 1) a64[i] += (long long)q[i] * (long long)w;               // IMAD.WIDE
            else if (MODE == 2) { a32[i] += q[i] * (w >> 12); b32[i] += q[i] * (w & 0xFFF); }   // split on the fly: 2 ALU + 2 IMAD
            else if (MODE == 4) ad[i] = fma(qd[i], wd, ad[i]);                          // DFMA
            else if (MODE == 5) { a32[i] += q[i] * w; b32[i] += q[(i + 1) & 7] * w; }   // two IMAD chains, halves already split
            else if (MODE == 6) { a64[i] += (long long)q[i] * (long long)w; ad[i] = fma(qd[i], wd, ad[i]); }  // IMAD.WIDE + DFMA together
            else if (MODE == 7) { a32[i] += q[i] * w; ad[i] = fma(qd[i], wd, ad[i]); }  // IMAD + DFMA together
        }
        w = (w << 1) ^ (w >> 3) ^ it;          // (not linear in `it`: no strength reduction of the products)
        wd = fma(wd, 1.0000001, 3.0);
    }
    long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a64[i] + a32[i] + b32[i] + (long long)ad[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double macs_per_iter)
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    long long* out;
    cudaMalloc(&out, sizeof(long long) * sms * 1024);
    const int iters = 20000, threads = 1024;
    k<MODE><<<sms, threads>>>(out, 10, 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<sms, threads>>>(out, iters, 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * khz * 1e3;
    printf("%-52s %.1f lane-MACs/clk/SM\n", name, (double)iters * macs_per_iter * threads / cycles);
    cudaFree(out);
}

int main()
{
    run<0>("IMAD (32-bit)", 8);
    run<1>("IMAD.WIDE (32 x 32 + 64)", 8);
    run<2>("wide MAC as 2 IMAD over halves split on the fly", 8);
    run<5>("wide MAC as 2 IMAD over halves already split", 8);
    run<4>("DFMA", 8);
    run<6>("IMAD.WIDE + DFMA side by side (2 MACs)", 16);
    run<7>("IMAD + DFMA side by side (2 MACs)", 16);
    return 0;
}
