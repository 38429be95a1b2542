// Gather latency (32 scattered 16-byte entries, dependent chain) while other warps stream 256-bit stores through the same
// L2-resident buffer (the prefix-moment ring pattern).   blockIdx < n_read: readers; the others: writers.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(uint4* buf, unsigned mask, int iters, int n_read, int wsleep, long long* out, unsigned* sink) {
    if ((int)blockIdx.x < n_read) {
        unsigned idx = (blockIdx.x * 977u + threadIdx.x * 12u) & mask;
        unsigned acc = 0;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + idx));
            acc += v.x + v.z + v.w;
            idx = (idx + 12289u * 13u + (v.y & 1u)) & mask;
        }
        long long t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
        if (acc == 0x12345u) *sink = acc;
    } else {
        const int w = blockIdx.x - n_read;
        long long t0 = clock64();
        int n = 0;
        // every writer warp: blocks of 512 entries (8 KB), lane writes 16 consecutive entries with 8 x 256-bit stores
        for (unsigned b = w; clock64() - t0 < 3000000; b += gridDim.x - n_read, ++n) {
            uint4* dst = buf + (((size_t)b * 512) & mask) + threadIdx.x * 16;
            for (int v = 0; v < 16; v += 2)
                asm volatile("st.global.v8.u32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(dst + v), "r"(b) : "memory");
            if (wsleep) __nanosleep(wsleep);
        }
        if (threadIdx.x == 0) out[blockIdx.x] = n;
    }
}
int main() {
    const size_t n = 1 << 21;                 // 32 MB of entries
    uint4* buf; long long* out; unsigned* sink;
    cudaMalloc(&buf, n * 16); cudaMemset(buf, 0, n * 16); cudaMalloc(&out, 8192 * 8); cudaMalloc(&sink, 4);
    long long h[8192];
    for (int n_read : {148, 1184}) for (int n_write : {0, 148, 768}) for (int wsleep : {0, 2000}) {
        if (n_write == 0 && wsleep) continue;
        k<<<n_read + n_write, 32>>>(buf, n - 1, 4000, n_read, wsleep, out, sink);
        cudaDeviceSynchronize();
        cudaMemcpy(h, out, (n_read + n_write) * 8, cudaMemcpyDeviceToHost);
        long long s = 0, wb = 0; for (int i = 0; i < n_read; ++i) s += h[i];
        for (int i = n_read; i < n_read + n_write; ++i) wb += h[i];
        printf("readers %4d writers %4d sleep %4d ns: gather latency %lld cycles; writer blocks total %lld (%.2f TB/s written)\n", n_read, n_write, wsleep,
               s / n_read, wb, wb * 8192.0 / (3000000 / 1.965e9) / 1e12);
    }
    return 0;
}
