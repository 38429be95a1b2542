// Latency of a warp-wide gather of 32 scattered 16-byte entries from an L2-resident buffer, by load flavour.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_gather tools/ubench_gather.cu && /tmp/ubench_gather
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE> __device__ __forceinline__ uint4 ld(const uint4* p) {
    uint4 v;
    if (MODE == 0) asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 1) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 2) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 3) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 4) asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    if (MODE == 5) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
template <int MODE> __global__ void k(const uint4* buf, unsigned mask, int iters, int stride, long long* out, unsigned* sink) {
    unsigned idx = (blockIdx.x * 977u + threadIdx.x * stride) & mask;
    unsigned acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint4 v = ld<MODE>(buf + idx);
        acc += v.x;
        idx = (idx + 12289u * 13u + (v.y & 1u)) & mask;          // dependent chain, new lines every time
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
    if (acc == 0x12345u) *sink = acc;
}
int main() {
    const size_t n = 1 << 21;                 // 32 MB of entries
    uint4* buf; long long* out; unsigned* sink;
    cudaMalloc(&buf, n * 16); cudaMemset(buf, 0, n * 16); cudaMalloc(&out, 4096 * 8); cudaMalloc(&sink, 4);
    long long h[4096];
    const char* names[] = {"ld.global", "ld.global.cg", "ld.global.nc", "ld.volatile", "ld.L1::no_allocate", "ld.relaxed.gpu"};
    for (int stride : {1, 12, 97}) {
        for (int ctas : {1, 148, 1184}) {
            printf("lane stride %3d entries, %4d warps (1 per CTA):", stride, ctas);
            for (int m = 0; m < 6; ++m) {
                for (int rep = 0; rep < 2; ++rep) {
                    if (m == 0) k<0><<<ctas, 32>>>(buf, n - 1, 2000, stride, out, sink);
                    if (m == 1) k<1><<<ctas, 32>>>(buf, n - 1, 2000, stride, out, sink);
                    if (m == 2) k<2><<<ctas, 32>>>(buf, n - 1, 2000, stride, out, sink);
                    if (m == 3) k<3><<<ctas, 32>>>(buf, n - 1, 2000, stride, out, sink);
                    if (m == 4) k<4><<<ctas, 32>>>(buf, n - 1, 2000, stride, out, sink);
                    if (m == 5) k<5><<<ctas, 32>>>(buf, n - 1, 2000, stride, out, sink);
                    cudaDeviceSynchronize();
                }
                cudaMemcpy(h, out, ctas * 8, cudaMemcpyDeviceToHost);
                long long s = 0; for (int i = 0; i < ctas; ++i) s += h[i];
                printf("  %s %lld", names[m], s / ctas);
            }
            printf("\n");
        }
    }
    return 0;
}
