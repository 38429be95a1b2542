// Library core: error reporting, device selection, C/A code tables (K-CODE part 1),
// sample-format conversion (K-CVT) and the FP32 peak probe used as roofline denominator.
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace sydr {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += n; }

// ------------------------------------------------------------------------------------------
// C/A code generation on the device.  One thread per PRN runs the two 10-stage LFSRs
// (G1: x^10+x^3+1, G2: x^10+x^9+x^8+x^6+x^3+x^2+1, all-ones start) and combines G1 with G2
// delayed by the IS-GPS-200 Table 3-Ia delay.  Same construction as
// sydr/signal/ca.py:70-112 (chip 1 -> +1.0).
// ------------------------------------------------------------------------------------------
__constant__ int16_t c_g2_delay[kMaxPrn] = {
    5,   6,   7,   8,   17,  18,  139, 140, 141, 251, 252, 254, 255, 256, 257, 258, 469, 470, 471,
    472, 473, 474, 509, 512, 513, 514, 515, 516, 859, 860, 861, 862, 863, 950, 947, 948, 950};

__global__ void ca_code_kernel(uint32_t* padded_bits, int8_t* chips) {
    const int p = threadIdx.x;
    if (p >= kMaxPrn) return;
    uint32_t g1 = 0x3ff, g2 = 0x3ff;             // bit i = stage i+1
    uint32_t s1[32], s2[32];
    for (int w = 0; w < 32; ++w) s1[w] = s2[w] = 0;
    for (int i = 0; i < kCodeChips; ++i) {
        s1[i >> 5] |= ((g1 >> 9) & 1u) << (i & 31);
        s2[i >> 5] |= ((g2 >> 9) & 1u) << (i & 31);
        const uint32_t f1 = ((g1 >> 2) ^ (g1 >> 9)) & 1u;
        const uint32_t f2 = ((g2 >> 1) ^ (g2 >> 2) ^ (g2 >> 5) ^ (g2 >> 7) ^ (g2 >> 8) ^ (g2 >> 9)) & 1u;
        g1 = ((g1 << 1) | f1) & 0x3ff;
        g2 = ((g2 << 1) | f2) & 0x3ff;
    }
    const int d = c_g2_delay[p];
    uint32_t* pb = padded_bits + p * kCodeWords;
    for (int w = 0; w < kCodeWords; ++w) pb[w] = 0;
    auto chip = [&](int i) -> uint32_t {
        int j = i - d;
        if (j < 0) j += kCodeChips;
        return ((s1[i >> 5] >> (i & 31)) ^ (s2[j >> 5] >> (j & 31))) & 1u;
    };
    for (int i = 0; i < kCodeChips; ++i) {
        const uint32_t c = chip(i);
        chips[p * kCodeChips + i] = c ? 1 : -1;
        const int k = i + 1;                       // padded position
        pb[k >> 5] |= c << (k & 31);
    }
    pb[0] |= chip(kCodeChips - 1);                 // padded[0]    = c1022
    pb[1024 >> 5] |= chip(0) << (1024 & 31);       // padded[1024] = c0
}

static std::mutex g_tab_mutex;
static CodeTables g_tables[64];
static bool g_tables_ready[64];

int ensure_code_tables(CodeTables* out) {
    int dev = 0;
    SYDR_CUDA_CHECK(cudaGetDevice(&dev));
    SYDR_REQUIRE(dev >= 0 && dev < 64, SYDR_ERR_ARG, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    if (!g_tables_ready[dev]) {
        uint32_t* bits = nullptr;
        int8_t* chips = nullptr;
        SYDR_CUDA_CHECK(cudaMalloc(&bits, sizeof(uint32_t) * kMaxPrn * kCodeWords));
        SYDR_CUDA_CHECK(cudaMalloc(&chips, kMaxPrn * kCodeChips));
        ca_code_kernel<<<1, 64>>>(bits, chips);
        count_launch();
        SYDR_CUDA_CHECK(cudaGetLastError());
        SYDR_CUDA_CHECK(cudaDeviceSynchronize());
        g_tables[dev].padded_bits = bits;
        g_tables[dev].chips = chips;
        g_tables_ready[dev] = true;
    }
    *out = g_tables[dev];
    return SYDR_OK;
}

// ------------------------------------------------------------------------------------------
// K-CVT: any IQ format -> complex64.
// ------------------------------------------------------------------------------------------
__global__ void convert_kernel(const void* in, int dt, long long n, float2* out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = load_sample(in, dt, i);
}

// ------------------------------------------------------------------------------------------
// FP32 peak probe: 8 independent FMA chains per thread, 148*k CTAs of 256 threads.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* sink, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
          x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456f) sink[0] = s;
}

__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) sink[0] = s;
}
// dependent-chain latency probes (one warp)
__global__ void latency_kernel(long long* out, int iters, double a, double b, float fa, float fb) {
    double x = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    float y = threadIdx.x;
    for (int i = 0; i < iters; ++i) y = fmaf(y, fa, fb);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = (long long)(x + y); }
}

}  // namespace sydr

using namespace sydr;

extern "C" {

int sydr_abi_version(void) { return SYDR_ABI_VERSION; }
const char* sydr_last_error(void) { return g_err; }
void sydr_clear_error(void) { g_err[0] = '\0'; }
long long sydr_launch_count(void) { return g_launches.load(); }
void sydr_reset_launch_count(void) { g_launches = 0; }

int sydr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int sydr_set_device(int device) {
    SYDR_CUDA_CHECK(cudaSetDevice(device));
    return SYDR_OK;
}

int sydr_measure_fp32_peak(double* h_tflops, double* h_sm_clock_mhz) {
    SYDR_REQUIRE(h_tflops != nullptr, SYDR_ERR_ARG, "h_tflops is NULL");
    int dev = 0, sms = 0, clk = 0;
    SYDR_CUDA_CHECK(cudaGetDevice(&dev));
    SYDR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SYDR_CUDA_CHECK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    float* sink = nullptr;
    SYDR_CUDA_CHECK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    SYDR_CUDA_CHECK(cudaEventCreate(&e0));
    SYDR_CUDA_CHECK(cudaEventCreate(&e1));
    const int iters = 4096, grid = sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        SYDR_CUDA_CHECK(cudaEventRecord(e0));
        fp32_peak_kernel<<<grid, 256>>>(sink, iters, 1.0000001f, 1e-9f);
        count_launch();
        SYDR_CUDA_CHECK(cudaEventRecord(e1));
        SYDR_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        SYDR_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * 64.0 * iters * 256.0 * grid;
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *h_tflops = best;
    if (h_sm_clock_mhz) *h_sm_clock_mhz = clk / 1000.0;
    return SYDR_OK;
}

int sydr_measure_fp64_peak(double* h_tflops, double* h_dfma_latency_cycles, double* h_ffma_latency_cycles) {
    SYDR_REQUIRE(h_tflops != nullptr, SYDR_ERR_ARG, "h_tflops is NULL");
    int dev = 0, sms = 0;
    SYDR_CUDA_CHECK(cudaGetDevice(&dev));
    SYDR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* sink = nullptr;
    SYDR_CUDA_CHECK(cudaMalloc(&sink, 64));
    cudaEvent_t e0, e1;
    SYDR_CUDA_CHECK(cudaEventCreate(&e0));
    SYDR_CUDA_CHECK(cudaEventCreate(&e1));
    const int iters = 512, grid = sms * 8;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        SYDR_CUDA_CHECK(cudaEventRecord(e0));
        fp64_peak_kernel<<<grid, 256>>>(sink, iters, 1.0000001, 1e-9);
        count_launch();
        SYDR_CUDA_CHECK(cudaEventRecord(e1));
        SYDR_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        SYDR_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 64.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    *h_tflops = best;
    long long h[3] = {0, 0, 0};
    latency_kernel<<<1, 32>>>(reinterpret_cast<long long*>(sink), 4096, 1.0000001, 1e-9, 1.0000001f, 1e-9f);
    count_launch();
    SYDR_CUDA_CHECK(cudaMemcpy(h, sink, sizeof(h), cudaMemcpyDeviceToHost));
    if (h_dfma_latency_cycles) *h_dfma_latency_cycles = h[0] / 4096.0;
    if (h_ffma_latency_cycles) *h_ffma_latency_cycles = h[1] / 4096.0;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return SYDR_OK;
}

int sydr_ca_code(int prn, double* h_code1023) {
    SYDR_REQUIRE(prn >= 1 && prn <= kMaxPrn, SYDR_ERR_ARG, "PRN %d out of range 1..%d", prn, kMaxPrn);
    SYDR_REQUIRE(h_code1023 != nullptr, SYDR_ERR_ARG, "output pointer is NULL");
    CodeTables t;
    int rc = ensure_code_tables(&t);
    if (rc != SYDR_OK) return rc;
    int8_t tmp[kCodeChips];
    SYDR_CUDA_CHECK(cudaMemcpy(tmp, t.chips + (prn - 1) * kCodeChips, kCodeChips, cudaMemcpyDeviceToHost));
    for (int i = 0; i < kCodeChips; ++i) h_code1023[i] = (double)tmp[i];
    return SYDR_OK;
}

int sydr_convert_to_f32(const void* d_in, int iq_dtype, long long n_samples, float* d_out_c64, void* stream) {
    SYDR_REQUIRE(d_in && d_out_c64, SYDR_ERR_ARG, "NULL pointer");
    SYDR_REQUIRE(iq_dtype >= SYDR_IQ_I8 && iq_dtype <= SYDR_IQ_F64, SYDR_ERR_ARG, "bad iq_dtype %d", iq_dtype);
    if (n_samples <= 0) return SYDR_OK;
    const int threads = 256;
    long long blocks = (n_samples + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    convert_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(d_in, iq_dtype, n_samples,
                                                                        reinterpret_cast<float2*>(d_out_c64));
    count_launch();
    SYDR_CUDA_CHECK(cudaGetLastError());
    return SYDR_OK;
}

}  // extern "C"
