// Pipe-rate microbenchmarks for the correlator inner loop design (run on the B200 box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench tools/ubench.cu && gpurun_out/ubench
// Reports warp-instructions per clock per SM for FFMA, FFMA2 (packed fp32x2), LOP3/SHF, I2F,
// PRMT and mixes of them, so that the inner loop can be balanced between the FMA and ALU pipes.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(256) k(float* sink, int iters, float fa, float fb, uint32_t ua) {
    float x[8];
    float2 y[8];
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; y[i] = make_float2(x[i], x[i] + 1.f); u[i] = threadIdx.x * 7 + i; }
    const float2 a2 = make_float2(fa, fa), b2 = make_float2(fb, fb);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) x[i] = fmaf(x[i], fa, fb);                                  // FFMA
                if (MODE == 1) y[i] = __ffma2_rn(y[i], a2, b2);                            // FFMA2
                if (MODE == 2) u[i] = (u[i] & ua) ^ (u[i] >> 3);                           // SHF + LOP3
                if (MODE == 3) { x[i] = fmaf(x[i], fa, fb); u[i] = (u[i] ^ ua) | (u[i] << 1); }   // FFMA + (SHF,LOP3)
                if (MODE == 4) { y[i] = __ffma2_rn(y[i], a2, b2); u[i] = (u[i] ^ ua) | (u[i] << 1); }  // FFMA2 + (SHF,LOP3)
                if (MODE == 5) { x[i] += (float)(int)(u[i] & 0xffff); u[i] += ua; }        // I2F + FADD + LOP3 + IADD
                if (MODE == 6) { x[i] += __uint_as_float(__byte_perm(u[i], 0x4b400000u, 0x7610)) ; u[i] += ua; }  // PRMT + FADD + IADD
                if (MODE == 7) { y[i] = __ffma2_rn(y[i], a2, b2); u[i] = (u[i] ^ ua) + it; }    // FFMA2 + LOP3 + IADD
                if (MODE == 8) { y[i] = __ffma2_rn(y[i], a2, b2); x[i] = fmaf(x[i], fa, fb); }  // FFMA2 + FFMA
                if (MODE == 9) { y[i] = __ffma2_rn(y[i], a2, b2); u[i] = (u[i] ^ ua); }    // FFMA2 + LOP3 (1:1)
                if (MODE == 10) { y[i] = __ffma2_rn(y[i], a2, b2); u[i] = (u[i] ^ ua); u[(i + 1) & 7] = __funnelshift_l(u[i], u[(i+1)&7], 3); }    // FFMA2 + 2 ALU
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + y[i].x + y[i].y + __uint_as_float(u[i]);
    if (s == 123.456f) sink[0] = s;
}

template <int MODE>
void run(const char* name, double instr_per_inner, double flop_per_inner) {
    float* sink;
    cudaMalloc(&sink, 4);
    int dev, sms, clk;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * 8;
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<grid, 256>>>(sink, ITERS, 1.0000001f, 1e-9f, 0x5a5a5a5au);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double inner = (double)ITERS * 4 * 8 * 256.0 * grid;          // per-thread inner statements
    const double cycles = best * 1e-3 * clk * 1e3;                      // at max clock
    const double warp_inner_per_clk_sm = inner / 32.0 / cycles / sms;
    printf("%-34s %8.3f ms  %6.3f inner/clk/SM(warp)  %6.3f warp-instr/clk/SM  %7.2f TFLOP/s\n", name, best,
           warp_inner_per_clk_sm, warp_inner_per_clk_sm * instr_per_inner, inner * flop_per_inner / (best * 1e-3) / 1e12);
    cudaFree(sink);
}

int main() {
    run<0>("FFMA", 1, 2);
    run<1>("FFMA2", 1, 4);
    run<2>("SHF+LOP3", 2, 0);
    run<3>("FFMA + SHF+LOP3", 3, 2);
    run<4>("FFMA2 + SHF+LOP3", 3, 4);
    run<5>("I2F+FADD+LOP3+IADD", 4, 1);
    run<6>("PRMT+FADD+IADD", 3, 1);
    run<7>("FFMA2 + LOP3+IADD", 3, 4);
    run<8>("FFMA2 + FFMA", 2, 6);
    run<9>("FFMA2 + LOP3", 2, 4);
    run<10>("FFMA2 + LOP3 + SHF", 3, 4);
    return 0;
}
