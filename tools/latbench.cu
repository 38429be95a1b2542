// Dependent-issue latency of the instructions on K-TRK's serial loop-closure chain (one warp).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/latbench tools/latbench.cu && gpurun_out/latbench
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, long long* cyc, double a, double b, float fa, float fb) {
    double x = a + threadIdx.x;
    float y = fa + threadIdx.x;
    const int N = 2048;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (MODE == 0) x = fma(x, b, a);                         // DFMA
        if (MODE == 1) x = x + b;                                // DADD
        if (MODE == 2) y = fmaf(y, fb, fa);                      // FFMA
        if (MODE == 3) x = __shfl_xor_sync(0xffffffffu, x, 1) + b;   // 2 SHFL + DADD
        if (MODE == 4) y = __shfl_xor_sync(0xffffffffu, y, 1) + fb;  // SHFL + FADD
        if (MODE == 5) x = sqrt(x) + b;                          // DSQRT + DADD
        if (MODE == 6) x = a / x + b;                            // DDIV + DADD
        if (MODE == 7) x = atan(x) + b;                          // atan + DADD
        if (MODE == 8) y = atanf(y) + fb;
        if (MODE == 9) y = sqrtf(y) + fb;
        if (MODE == 10) y = fa / y + fb;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / N;
    out[threadIdx.x] = x + y;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 256 * 8); cudaMalloc(&c, 8);
    const char* names[] = {"DFMA", "DADD", "FFMA", "2xSHFL+DADD", "SHFL+FADD", "sqrt(double)+DADD", "div(double)+DADD", "atan(double)+DADD",
                           "atanf+FADD", "sqrtf+FADD", "div(float)+FADD"};
    for (int m = 0; m < 11; ++m) {
        switch (m) {
#define C(M) case M: k<M><<<1, 32>>>(d, c, 1.0000001, 0.9999999, 1.0001f, 0.9999f); break;
            C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10)
        }
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("%-22s %5lld clk per dependent op\n", names[m], h);
    }
    return 0;
}
