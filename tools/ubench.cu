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

// Dependent-chain latencies (one warp): cycles per op.
__global__ void lat_kernel(long long* out, int iters, double a, double b, float fa, float fb) {
    double x = threadIdx.x + 1.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    double y = x;
    for (int i = 0; i < iters; ++i) y = y + a;
    long long t2 = clock64();
    float z = (float)y;
    for (int i = 0; i < iters; ++i) z = fmaf(z, fa, fb);
    long long t3 = clock64();
    double r = y;
    for (int i = 0; i < iters; ++i) { double q; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(r)); r = q + a; }
    long long t4 = clock64();
    double sq = r;
    for (int i = 0; i < iters; ++i) sq = sqrt(sq) + a;
    long long t5 = clock64();
    double dv = sq;
    for (int i = 0; i < iters; ++i) dv = a / dv + b;
    long long t6 = clock64();
    double at = dv;
    for (int i = 0; i < iters; ++i) at = atan(at) + a;
    long long t7 = clock64();
    float sh = z;
    for (int i = 0; i < iters; ++i) sh = __shfl_xor_sync(0xffffffffu, sh, 1) + fa;
    long long t8 = clock64();
    if (threadIdx.x == 0) {
        out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; out[5] = t6 - t5;
        out[6] = t7 - t6; out[7] = t8 - t7; out[8] = (long long)(x + y + z + r + sq + dv + at + sh);
    }
}
template <int MODE>
__global__ void __launch_bounds__(256) k64(double* sink, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = (MODE == 0) ? fma(x[i], a, b) : (x[i] + a);
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456) sink[0] = s;
}

int main() {
    {
        long long* d; cudaMalloc(&d, 128);
        const int iters = 2048;
        lat_kernel<<<1, 32>>>(d, iters, 1.0000001, 1e-9, 1.0000001f, 1e-9f);
        long long h[9]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        const char* nm[8] = {"DFMA", "DADD", "FFMA", "RCP64H+DADD", "sqrt(double)+DADD", "a/x+b (IEEE div)", "atan(double)+DADD", "SHFL+FADD"};
        for (int i = 0; i < 8; ++i) printf("latency %-22s %7.1f cycles/iter\n", nm[i], (double)h[i] / iters);
        double* sink; cudaMalloc(&sink, 8);
        int sms, clk; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
        for (int mode = 0; mode < 2; ++mode) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k64<0><<<sms * 8, 256>>>(sink, 4096, 1.0000001, 1e-9); else k64<1><<<sms * 8, 256>>>(sink, 4096, 1.0000001, 1e-9);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
            }
            const double ops = 4096.0 * 8 * 256 * sms * 8;
            printf("throughput %-6s %6.3f warp-instr/clk/SM  (%.2f Tops/s)\n", mode ? "DADD" : "DFMA", ops / 32 / (best * 1e-3 * clk * 1e3) / sms, ops / (best * 1e-3) / 1e12);
        }
    }

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
