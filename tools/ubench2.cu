// Pipe-rate microbenchmarks for the prefix-moment tracking kernel (K-TRKM) design (run on the B200 box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench2 tools/ubench2.cu && gpurun_out/ubench2
// warp-instructions per clock per SM of IDP.2A, IMAD, IADD3, I2F, MUFU, DADD, DFMA, SHFL, REDUX, LDS.128, STS.128.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(int* sink, int iters, int ia, int ib, float fa, double da) {
    __shared__ uint4 sm[256 * 2];
    int u[8]; float x[8]; double d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { u[i] = threadIdx.x * 7 + i; x[i] = threadIdx.x + i; d[i] = threadIdx.x + i; }
    sm[threadIdx.x] = make_uint4(1, 2, 3, 4); sm[threadIdx.x + 256] = make_uint4(1, 2, 3, 4);
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) u[i] = __dp2a_lo(ia + i, 0x0001, u[i]);              // IDP.2A
                if (MODE == 1) u[i] = u[i] * ia + ib;                               // IMAD
                if (MODE == 2) u[i] = u[i] + ia + ib;                               // IADD3
                if (MODE == 3) { x[i] += (float)u[i]; u[i] += ia; }                 // I2F + FADD + IADD
                if (MODE == 4) x[i] = __sinf(x[i]);                                 // FMUL + MUFU.SIN
                if (MODE == 5) d[i] = d[i] + da;                                    // DADD
                if (MODE == 6) d[i] = fma(d[i], da, da);                            // DFMA
                if (MODE == 7) u[i] = __shfl_xor_sync(0xffffffffu, u[i], 1);        // SHFL
                if (MODE == 8) u[i] = __reduce_add_sync(0xffffffffu, u[i]);         // REDUX
                if (MODE == 9) { uint4 v = sm[(threadIdx.x + u[i]) & 511]; u[i] = (v.x ^ v.y) + (v.z ^ v.w); }   // LDS.128 + 5 int
                if (MODE == 10) { sm[(threadIdx.x + i * 32) & 511] = make_uint4(u[i], it, r, i); }   // STS.128
                if (MODE == 11) { u[i] = __dp2a_lo(ia + i, 0x0001, u[i]); x[i] = fmaf(x[i], fa, fa); }   // IDP + FFMA
                if (MODE == 12) { u[i] = u[i] * ia + ib; x[i] = fmaf(x[i], fa, fa); }                    // IMAD + FFMA
                if (MODE == 13) { u[i] = __dp2a_lo(ia + i, 0x0001, u[i]); u[(i + 1) & 7] ^= ib; }          // IDP + LOP3
                if (MODE == 14) { x[i] = __int_as_float(u[i] + 0x4B400000) - 12582912.f; u[i] += ia; }     // magic int->float: IADD + FADD (+IADD)
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += u[i] + (int)x[i] + (int)d[i];
    if (s == 123456) sink[0] = s + sm[5].x;
}

template <int MODE>
void run(const char* name, double instr_per_inner) {
    int* sink; cudaMalloc(&sink, 4);
    int sms, clk;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * 8, iters = 1024;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<grid, 256>>>(sink, iters, 3, 5, 1.0000001f, 1.0000001);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    const double inner = (double)iters * 4 * 8 * 256.0 * grid;
    const double cycles = best * 1e-3 * clk * 1e3;
    const double w = inner / 32.0 / cycles / sms;
    printf("%-28s %8.3f ms  %6.3f inner/clk/SM(warp)  %6.3f warp-instr/clk/SM\n", name, best, w, w * instr_per_inner);
    cudaFree(sink);
}

int main() {
    run<0>("IDP.2A", 1); run<1>("IMAD", 1); run<2>("IADD3", 1); run<3>("I2F+FADD+IADD", 3); run<4>("FMUL+MUFU.SIN", 2);
    run<5>("DADD", 1); run<6>("DFMA", 1); run<7>("SHFL", 1); run<8>("REDUX", 1); run<9>("LDS.128(+5 int)", 6);
    run<10>("STS.128", 1); run<11>("IDP.2A + FFMA", 2); run<12>("IMAD + FFMA", 2); run<13>("IDP.2A + LOP3", 2);
    run<14>("magic i2f: IADD+FADD+IADD", 3);
    return 0;
}
