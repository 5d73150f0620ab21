// int_rate.cu -- developer microbenchmark: per-SM issue rates of the integer instructions the model-search
// kernel is made of (IADD3, IABS, VABSDIFF = __sad, IMAD, SHF, LOP3), alone and mixed.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o int_rate int_rate.cu && ./int_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(unsigned* out, int iters, int seed)
{
    unsigned a[8];
    int x = seed + threadIdx.x, y = seed * 3 + 1;
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = a[i] + (unsigned)x + (unsigned)i;                 // IADD3
            else if (MODE == 1) a[i] = __sad((int)a[i], x, (unsigned)y);            // VABSDIFF
            else if (MODE == 2) a[i] = (unsigned)abs((int)a[i] - x) + (unsigned)y;  // SUB + IABS + ADD
            else if (MODE == 3) a[i] = a[i] * (unsigned)x + (unsigned)y;            // IMAD
            else if (MODE == 4) { a[i] = __sad((int)a[i], x, (unsigned)y); a[i] = a[i] * 3u + (unsigned)x; }  // VABSDIFF + IMAD
            else if (MODE == 5) { a[i] = (a[i] >> 3) ^ (unsigned)x; }               // SHF + LOP3
            else if (MODE == 6) { a[i] = a[i] + (a[(i + 1) & 7] ^ (unsigned)x); }   // LOP3 + IADD3 (not foldable)
            else if (MODE == 7) { a[i] = (unsigned)abs((int)(a[i] ^ (unsigned)x)); } // LOP3 + IABS
            else if (MODE == 8) { a[i] = __sad((int)a[i], x, a[(i + 1) & 7] ^ (unsigned)y); }   // VABSDIFF + LOP3
            else if (MODE == 9) { a[i] = (a[i] >> (x & 7)) + a[(i + 3) & 7]; }       // SHF + IADD3
            else if (MODE == 10) { a[i] = a[i] ^ (a[(i + 1) & 7] & (unsigned)x); }   // LOP3 only (one 3-input op)
            else if (MODE == 11) { a[i] = (a[i] >> (x & 7)); a[i] = a[i] << (y & 3); } // SHF x2
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double ops_per_iter)
{
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    unsigned* out;
    cudaMalloc(&out, sizeof(unsigned) * sms * 1024);
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
    printf("%-34s %.1f lane-ops/clk/SM\n", name, (double)iters * ops_per_iter * threads / cycles);
    cudaFree(out);
}

int main()
{
    run<0>("IADD3", 8);
    run<1>("VABSDIFF (__sad)", 8);
    run<2>("SUB + IABS + ADD (as 3 ops)", 24);
    run<3>("IMAD", 8);
    run<4>("VABSDIFF + IMAD (2 ops)", 16);
    run<5>("SHF + LOP3 (2 ops)", 16);
    run<6>("LOP3 + IADD3 (2 ops)", 16);
    run<7>("LOP3 + IABS (2 ops)", 16);
    run<8>("VABSDIFF + LOP3 (2 ops)", 16);
    run<9>("SHF + IADD3 (2 ops)", 16);
    run<10>("LOP3 (1 op)", 8);
    run<11>("SHF + SHF (2 ops)", 16);
    return 0;
}
