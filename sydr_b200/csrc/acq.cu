// K-CODE + K-ACQ: Parallel Code Phase Search acquisition for sm_100a.
//
// Reference semantics (file:line in /root/reference):
//   PCPS                          sydr/dsp/acquisition.py:9-74
//   TwoCorrelationPeakComparison  sydr/dsp/acquisition.py:78-115
//   code spectrum                 sydr/channel/channel_l1ca_borre.py:281-284,
//                                 sydr/signal/gnsssignal.py:35-70
//
// What the reference computes per (PRN, Doppler bin, non-coherent block) is
//     | sum_coh ifft( fft(x . carrier) . conj(fft(code)) ) |
// summed over the non-coherent blocks.  Restructured for the GPU without changing the result
// beyond FP32 rounding:
//   * fft(x . carrier) does not depend on the PRN: it is computed once per (bin, block) by
//     acq_fwd_kernel and shared by all PRNs (1/n_prn of the inverse-FFT work);
//   * the coherent sum commutes with the linear transforms, so the coh periods are summed in
//     the time domain before the single forward FFT;
//   * Doppler bins that differ by a multiple of fs/N (1 kHz for a 1 ms code period) have the same
//     forward spectrum, circularly shifted by that many bins (the extra carrier factor is
//     exp(2 pi i q n / N)): with 250 Hz steps only 4 of the 41 rows are transformed, the others
//     read a base row at a rotated index.  The forward spectra shrink from 82 MB to 8 MB at
//     25 MS/s and stay in L2;
//   * acq_ifft_kernel owns one (PRN, bin) row per CTA: spectrum multiply -> inverse FFT in
//     shared memory -> magnitude -> non-coherent sum in registers -> arg-max / second peak
//     with warp shuffles.  The correlation map is only written when the caller asks for it.
// The FFT is an in-place mixed-radix transform held in one shared-memory buffer: the forward
// transform is decimation-in-time (digit-reversed gather of the time samples, natural-order
// spectrum out) and the inverse is decimation-in-frequency (natural-order spectrum in,
// digit-reversed time samples out), so neither needs a permutation pass or a second buffer and
// the first inverse stage reads the spectra with fully coalesced loads.  N = 50 000 does not
// fit one SM: a radix-2 DIF split hands the even and odd output samples to the two CTAs of a
// cluster, which exchange their row maxima through distributed shared memory.
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "fft_radix.cuh"

namespace sydr {

// ------------------------------------------------------------------------------------------
// K-CODE: code spectra in FP64 with a global-memory transform (init-time only).
// ------------------------------------------------------------------------------------------
// UpsampleCode (gnsssignal.py:46-56): idx = trunc(ts * k / tc), ts = 1/fs, tc = 1/1.023e6.
__global__ void upsample_code_kernel(const int8_t* __restrict__ chips, const int* __restrict__ prns, int n_code,
                                     double fs, double2* __restrict__ out) {
    const int p = blockIdx.y;
    const int8_t* c = chips + (prns[p] - 1) * kCodeChips;
    const double ts = 1.0 / fs, tc = 1.0 / kCodeFreq;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n_code; k += gridDim.x * blockDim.x) {
        const int idx = (int)trunc(__ddiv_rn(__dmul_rn(ts, (double)k), tc));
        out[(size_t)p * n_code + k] = make_double2((double)c[idx], 0.0);
    }
}

// One out-of-place Stockham pass of radix R (any R, O(R^2) butterfly) in FP64.
__global__ void fft64_pass_kernel(const double2* __restrict__ in, double2* __restrict__ out, int n, int R, int p) {
    const int batch = blockIdx.y;
    in += (size_t)batch * n;
    out += (size_t)batch * n;
    const int t = n / R;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < t; i += gridDim.x * blockDim.x) {
        const int k = i % p;
        const int j = (i - k) * R + k;
        for (int q = 0; q < R; ++q) {
            double sr = 0.0, si = 0.0;
            for (int r = 0; r < R; ++r) {
                const double2 x = in[i + r * t];
                // twiddle exp(-2 pi i r k /(pR)) * exp(-2 pi i r q / R) = exp(-2 pi i r (k + q p)/(pR))
                const long long num = ((long long)r * (k + (long long)q * p)) % ((long long)p * R);
                double s, c;
                sincospi(-2.0 * (double)num / (double)((long long)p * R), &s, &c);
                sr += x.x * c - x.y * s;
                si += x.x * s + x.y * c;
            }
            out[j + q * p] = make_double2(sr, si);
        }
    }
}

// Direct O(N^2) DFT for lengths with large prime factors (init-time only).
__global__ void dft64_direct_kernel(const double2* __restrict__ in, double2* __restrict__ out, int n) {
    const int batch = blockIdx.y;
    in += (size_t)batch * n;
    out += (size_t)batch * n;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        double sr = 0.0, si = 0.0;
        for (int m = 0; m < n; ++m) {
            const long long num = ((long long)k * m) % n;
            double s, c;
            sincospi(-2.0 * (double)num / (double)n, &s, &c);
            const double2 x = in[m];
            sr += x.x * c - x.y * s;
            si += x.x * s + x.y * c;
        }
        out[k] = make_double2(sr, si);
    }
}

// conj(.) and optional scale -> float2 table used by the hot kernels.
__global__ void spectrum_finish_kernel(const double2* __restrict__ in, int n_total, double scale,
                                       float2* __restrict__ out32, double2* __restrict__ out64) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_total; i += gridDim.x * blockDim.x) {
        const double2 v = in[i];
        if (out32) out32[i] = make_float2((float)(v.x * scale), (float)(-v.y * scale));
        if (out64) out64[i] = make_double2(v.x, -v.y);
    }
}

static std::vector<int> factorize(int n, int maxr) {
    std::vector<int> f;
    for (int r : {4, 2, 3, 5, 7, 11, 13}) {
        if (r > maxr) continue;
        while (n % r == 0) { f.push_back(r); n /= r; }
    }
    if (n != 1) f.clear();
    return f;
}

// FP64 forward FFT of `batch` length-n rows: d_a (in/out), d_b scratch.  Result pointer returned.
static int fft64_batched(double2* d_a, double2* d_b, int n, int batch, cudaStream_t s, double2** result) {
    std::vector<int> f = factorize(n, 13);
    const int threads = 128;
    if (f.empty()) {
        dim3 grid((n + threads - 1) / threads, batch);
        dft64_direct_kernel<<<grid, threads, 0, s>>>(d_a, d_b, n);
        count_launch();
        SYDR_CUDA_CHECK(cudaGetLastError());
        *result = d_b;
        return SYDR_OK;
    }
    int p = 1;
    double2 *src = d_a, *dst = d_b;
    for (int R : f) {
        dim3 grid((n / R + threads - 1) / threads, batch);
        fft64_pass_kernel<<<grid, threads, 0, s>>>(src, dst, n, R, p);
        count_launch();
        SYDR_CUDA_CHECK(cudaGetLastError());
        p *= R;
        double2* t = src; src = dst; dst = t;
    }
    *result = src;
    return SYDR_OK;
}

// ------------------------------------------------------------------------------------------
// Shared-memory FFT stages
// ------------------------------------------------------------------------------------------
// In-place stage on block size M, radix R (L = M/R): elements base + k + r*L.
// DIF: butterfly then post-twiddle w_M^{k r};  DIT: pre-twiddle then butterfly.
// tw[k] = exp(-2 pi i k / M), k < L.
template <int R, int M, int N, int T, bool DIF>
__device__ __forceinline__ void stage_smem(float2* __restrict__ buf, const float2* __restrict__ tw) {
    constexpr int L = M / R;
    constexpr int NB = N / R;
    for (int b = threadIdx.x; b < NB; b += T) {
        const int blk = b / L, k = b - blk * L;
        float2* p = buf + blk * M + k;
        float2 u[R];
#pragma unroll
        for (int r = 0; r < R; ++r) u[r] = p[r * L];
        if (DIF) {
            Dft<R>::run(u);
            if (L > 1) twiddle_powers<R>(u, __ldg(tw + k));
        } else {
            if (L > 1) twiddle_powers<R>(u, __ldg(tw + k));
            Dft<R>::run(u);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) p[r * L] = u[r];
    }
}

// Compile-time plan: up to five radices (unused = 1).  Stage s works on block size M_s,
// M_1 = N, M_{s+1} = M_s / R_s.
template <int N_, int T_, int R1_, int R2_, int R3_, int R4_, int R5_>
struct Plan {
    static constexpr int N = N_, T = T_, R1 = R1_, R2 = R2_, R3 = R3_, R4 = R4_, R5 = R5_;
    static constexpr int M1 = N, M2 = M1 / R1, M3 = M2 / R2, M4 = M3 / R3, M5 = M4 / R4;
    static constexpr int NSTAGE = (R5 > 1) ? 5 : (R4 > 1) ? 4 : (R3 > 1) ? 3 : 2;
    static constexpr int RL = (R5 > 1) ? R5 : (R4 > 1) ? R4 : (R3 > 1) ? R3 : R2;   // last radix
    static constexpr int NBL = N / RL;                      // butterflies of the last stage
    static constexpr int ROUNDS = (NBL + T - 1) / T;
    static_assert(R1 * R2 * R3 * R4 * R5 == N, "radices must multiply to N");
    // twiddle-table offsets (entries) for stages 1..4 (L_s = M_s/R_s entries each)
    static constexpr int TW1 = 0, TW2 = TW1 + M1 / R1, TW3 = TW2 + M2 / R2, TW4 = TW3 + M3 / R3,
                         TWN = TW4 + M4 / R4;
};

// Natural index of last-stage butterfly b, output r:  n = n0(b) + r * (N / RL), where n0 is the
// digit reversal of b over the radices before the last one.
template <class P>
__device__ __forceinline__ int natural_base(int b) {
    // b = d1*(NBL/R1) + d2*(NBL/(R1 R2)) + ... ; n0 = d1 + R1*d2 + R1*R2*d3 + ...
    int n0 = 0, mult = 1, m = P::NBL;
    if (P::NSTAGE >= 2) { m /= P::R1; const int d = b / m; b -= d * m; n0 += d * mult; mult *= P::R1; }
    if (P::NSTAGE >= 3) { m /= P::R2; const int d = b / m; b -= d * m; n0 += d * mult; mult *= P::R2; }
    if (P::NSTAGE >= 4) { m /= P::R3; const int d = b / m; b -= d * m; n0 += d * mult; mult *= P::R3; }
    if (P::NSTAGE >= 5) { m /= P::R4; const int d = b / m; b -= d * m; n0 += d * mult; mult *= P::R4; }
    return n0;
}

struct AcqDev {                 // device-side view of a plan
    const float2* code_spec;    // [n_prn][n_code], conj(fft(code))/n_code, natural order
    float2* Y;                  // [n_bins_local][noncoh][n_code] forward spectra
    const float2* tw;           // per-stage twiddle tables (Plan::TW* offsets), half-plan length
    const float2* tw_split;     // w_N^e, e < N/2 (HALVES == 2 only)
    int n_code, n_prn, n_rows, bin_lo, coh, noncoh, chip;
    int n_base;                 // > 0: only the global Doppler rows 0 .. n_base-1 have forward spectra; global row R uses
                                // row R % n_base circularly shifted by R / n_base bins (they differ by multiples of fs/N)
    double fs, inter_freq, doppler_range, doppler_step;
};

// ---- forward: carrier wipe-off + coherent pre-sum + DIT FFT -> Y (natural order) ------------
template <class P, int HALVES>
__global__ void __launch_bounds__(P::T) acq_fwd_kernel(const AcqDev A, const void* __restrict__ iq, int dt) {
    extern __shared__ __align__(16) float2 fbuf[];
    constexpr int NH = P::N;                     // length handled by this CTA
    constexpr int NF = NH * HALVES;              // full code length
    const int h = (HALVES == 2) ? (blockIdx.x & 1) : 0;
    const int cell = blockIdx.x / HALVES;        // (row, block)
    const int row = cell / A.noncoh, blk = cell - row * A.noncoh;
    // acquisition.py:34,42: freq = IF - bins[b], bins = arange(-range, range+1, step)
    // with shared spectra the transformed rows are the *global* rows 0 .. n_base-1, whatever Doppler range this
    // plan owns: every shard of a multi-GPU search then works from identical base spectra
    const int grow = (A.n_base > 0) ? row : A.bin_lo + row;
    const double fbin = -A.doppler_range + (double)grow * A.doppler_step;
    const double freq = A.inter_freq - fbin;
    const long long blk0 = (long long)blk * A.coh * NF;

    // carrier-wiped, coherently summed sample n (acquisition.py:33,45,53)
    auto wiped = [&](int n) -> float2 {
        float2 acc = make_float2(0.f, 0.f);
        for (int m = 0; m < A.coh; ++m) {
            const long long na = (long long)m * NF + n;
            const double pp = __ddiv_rn(__dmul_rn((double)(na * 2), kPi), A.fs);     // phasePoints
            double turns = -(__dmul_rn(freq, pp)) * 0.15915494309189535;
            turns -= rint(turns);
            float s, c;
            sincospif((float)(2.0 * turns), &s, &c);
            const float2 x = load_sample(iq, dt, blk0 + na);
            acc.x += x.x * c - x.y * s;
            acc.y += x.x * s + x.y * c;
        }
        return acc;
    };
    auto input = [&](int n) -> float2 {          // time sample n of this CTA's (half) transform
        if (HALVES == 1) return wiped(n);
        const float2 a = wiped(n), b = wiped(n + NH);
        if (h == 0) return cadd(a, b);
        return cmul(csub(a, b), __ldg(A.tw_split + n));
    };

    // first DIT stage = last plan stage: contiguous RL elements, no twiddle (k = 0)
    constexpr int RL = P::RL;
    for (int b = threadIdx.x; b < P::NBL; b += P::T) {
        const int n0 = natural_base<P>(b);
        float2 u[RL];
#pragma unroll
        for (int r = 0; r < RL; ++r) u[r] = input(n0 + r * P::NBL);
        Dft<RL>::run(u);
#pragma unroll
        for (int r = 0; r < RL; ++r) fbuf[b * RL + r] = u[r];
    }
    __syncthreads();
    if (P::NSTAGE >= 5) { stage_smem<P::R4, P::M4, NH, P::T, false>(fbuf, A.tw + P::TW4); __syncthreads(); }
    if (P::NSTAGE >= 4) { stage_smem<P::R3, P::M3, NH, P::T, false>(fbuf, A.tw + P::TW3); __syncthreads(); }
    if (P::NSTAGE >= 3) { stage_smem<P::R2, P::M2, NH, P::T, false>(fbuf, A.tw + P::TW2); __syncthreads(); }
    // last DIT stage (plan stage 1, L = NH/R1): write natural-order spectrum to global
    {
        constexpr int R = P::R1, L = NH / R;
        float2* out = A.Y + ((size_t)cell * HALVES + h) * NH;
        for (int k = threadIdx.x; k < L; k += P::T) {
            float2 u[R];
#pragma unroll
            for (int r = 0; r < R; ++r) u[r] = fbuf[k + r * L];
            twiddle_powers<R>(u, __ldg(A.tw + P::TW1 + k));
            Dft<R>::run(u);
#pragma unroll
            for (int r = 0; r < R; ++r) out[k + r * L] = u[r];
        }
    }
}

// ---- row summary helpers --------------------------------------------------------------------
struct PeakRed { float v; int i; };
__device__ __forceinline__ PeakRed peak_better(PeakRed a, PeakRed b) {
    // larger value wins; equal values -> lower index (np.argmax returns the first occurrence)
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ PeakRed warp_peak(PeakRed p) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        PeakRed q;
        q.v = __shfl_xor_sync(0xffffffffu, p.v, o);
        q.i = __shfl_xor_sync(0xffffffffu, p.i, o);
        p = peak_better(p, q);
    }
    return p;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Is code index n searched for the second peak?  (acquisition.py:103-110, quirks kept)
__device__ __forceinline__ bool in_second_range(int n, int i1, int chip, int n_code) {
    const int e0 = i1 - chip, e1 = i1 + chip;
    if (e0 < 1) return n >= e1 && n < n_code - 1;
    if (e1 >= n_code) return n < e0;
    return n < e0 || (n >= e1 && n < n_code - 1);
}

// |.| with the hardware square root (one MUFU.SQRT, ~1 ulp) instead of the correctly rounded sqrtf, whose
// RSQ + Newton + range-check sequence is 10 instructions per sample: 8 % of this kernel.  The FP32
// transform in front of it already carries ~1e-6 of rounding noise.
__device__ __forceinline__ float sqrt_mufu(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ---- inverse: spectrum multiply + DIF FFT + |.| + non-coherent sum + peak search ------------
template <class P, int HALVES, bool CODE_IN_SMEM>
__global__ void __launch_bounds__(P::T) acq_ifft_kernel(const AcqDev A, sydr_acq_row* __restrict__ rows,
                                                        float* __restrict__ maps) {
    extern __shared__ __align__(16) float2 fbuf[];
    constexpr int NH = P::N, NF = NH * HALVES, T = P::T, RL = P::RL, ROUNDS = P::ROUNDS;
    __shared__ PeakRed s_peak[32];
    __shared__ float s_max[32];
    __shared__ PeakRed s_best;
    __shared__ float s_m2;
    __shared__ PeakRed x_peak[2];     // DSMEM exchange slots (HALVES == 2)
    __shared__ float x_m2[2];

    const int h = (HALVES == 2) ? (int)cluster_ctarank() : 0;
    // CTAs are ordered row-major over (Doppler row, PRN): the ~148 resident CTAs read a handful of
    // forward spectra Y[row] (all PRNs share them) and the 32 code spectra, which stay in L2
    const int cta = blockIdx.x / HALVES;
    const int row = cta / A.n_prn, slot = cta - row * A.n_prn;
    const int rowid = slot * A.n_rows + row;              // index into rows[] / maps[]
    const float2* __restrict__ C = A.code_spec + (size_t)slot * NF;
    float2* cbuf = fbuf + NH;                             // code spectrum copy (CODE_IN_SMEM)
    if (CODE_IN_SMEM) {
        for (int i = threadIdx.x; i < NF; i += T) cbuf[i] = __ldg(C + i);
        __syncthreads();
    }
    auto code_at = [&](int f) -> float2 { return CODE_IN_SMEM ? cbuf[f] : __ldg(C + f); };

    float acc[ROUNDS][RL];
#pragma unroll
    for (int q = 0; q < ROUNDS; ++q)
#pragma unroll
        for (int r = 0; r < RL; ++r) acc[q][r] = 0.f;

    for (int blk = 0; blk < A.noncoh; ++blk) {
        // forward spectrum of this row: its own, or base row (row % n_base) read q = row / n_base bins lower
        const int brow = (A.n_base > 0) ? (A.bin_lo + row) % A.n_base : row;
        const int qsh = (A.n_base > 0) ? (A.bin_lo + row) / A.n_base : 0;
        const float2* __restrict__ Y = A.Y + ((size_t)brow * A.noncoh + blk) * NF;
        // swapped product z = swap(Y[f] * C[f]); Y of a split transform is stored [parity][m]
        auto zin = [&](int f) -> float2 {
            int fy = f - qsh;
            fy += (fy < 0) ? NF : 0;
            const float2 y = (HALVES == 2) ? __ldg(Y + (size_t)(fy & 1) * NH + (fy >> 1)) : __ldg(Y + fy);
            const float2 z = cmul(y, code_at(f));
            return make_float2(z.y, z.x);
        };
        auto input = [&](int e) -> float2 {
            if (HALVES == 1) return zin(e);
            const float2 a = zin(e), b = zin(e + NH);
            if (h == 0) return cadd(a, b);
            return cmul(csub(a, b), __ldg(A.tw_split + e));
        };
        // stage 1 (M = NH): coalesced global loads straight into the butterfly
        {
            constexpr int R = P::R1, L = NH / R;
            for (int k = threadIdx.x; k < L; k += T) {
                float2 u[R];
#pragma unroll
                for (int r = 0; r < R; ++r) u[r] = input(k + r * L);
                Dft<R>::run(u);
                twiddle_powers<R>(u, __ldg(A.tw + P::TW1 + k));
#pragma unroll
                for (int r = 0; r < R; ++r) fbuf[k + r * L] = u[r];
            }
        }
        __syncthreads();
        if (P::NSTAGE >= 3) { stage_smem<P::R2, P::M2, NH, T, true>(fbuf, A.tw + P::TW2); __syncthreads(); }
        if (P::NSTAGE >= 4) { stage_smem<P::R3, P::M3, NH, T, true>(fbuf, A.tw + P::TW3); __syncthreads(); }
        if (P::NSTAGE >= 5) { stage_smem<P::R4, P::M4, NH, T, true>(fbuf, A.tw + P::TW4); __syncthreads(); }
        // last stage (L = 1): butterfly, magnitude, non-coherent accumulation in registers
#pragma unroll
        for (int q = 0; q < ROUNDS; ++q) {
            const int b = threadIdx.x + q * T;
            if (b < P::NBL) {
                float2 u[RL];
#pragma unroll
                for (int r = 0; r < RL; ++r) u[r] = fbuf[b * RL + r];
                Dft<RL>::run(u);
#pragma unroll
                for (int r = 0; r < RL; ++r) acc[q][r] += sqrt_mufu(u[r].x * u[r].x + u[r].y * u[r].y);   // L68
            }
        }
        __syncthreads();                                   // fbuf is rewritten by the next block
    }

    // ---- row peak: natural index of acc[q][r] is  HALVES * (n0(b) + r*NBL) + h
    PeakRed best; best.v = -1.f; best.i = 0x7fffffff;
#pragma unroll
    for (int q = 0; q < ROUNDS; ++q) {
        const int b = threadIdx.x + q * T;
        if (b < P::NBL) {
            const int n0 = natural_base<P>(b);
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                PeakRed c; c.v = acc[q][r]; c.i = HALVES * (n0 + r * P::NBL) + h;
                best = peak_better(best, c);
            }
        }
    }
    best = warp_peak(best);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_peak[warp] = best;
    __syncthreads();
    if (warp == 0) {
        PeakRed p = (lane < T / 32) ? s_peak[lane] : PeakRed{-1.f, 0x7fffffff};
        p = warp_peak(p);
        if (lane == 0) s_best = p;
    }
    __syncthreads();
    if (HALVES == 2) {
        // exchange the half-row maxima through DSMEM
        if (threadIdx.x == 0) {
            const PeakRed mine = s_best;
            for (uint32_t r = 0; r < 2; ++r) {
                const uint32_t dst = mapa_u32(smem_u32(&x_peak[h]), r);
                asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(dst), "r"(__float_as_uint(mine.v)) : "memory");
                asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(dst + 4), "r"(mine.i) : "memory");
            }
        }
        cluster_sync_all();
        if (threadIdx.x == 0) s_best = peak_better(x_peak[0], x_peak[1]);
        __syncthreads();
    }
    const PeakRed win = s_best;
    float m2 = -1.f;
#pragma unroll
    for (int q = 0; q < ROUNDS; ++q) {
        const int b = threadIdx.x + q * T;
        if (b < P::NBL) {
            const int n0 = natural_base<P>(b);
#pragma unroll
            for (int r = 0; r < RL; ++r) {
                const int n = HALVES * (n0 + r * P::NBL) + h;
                if (in_second_range(n, win.i, A.chip, NF)) m2 = fmaxf(m2, acc[q][r]);
            }
        }
    }
    m2 = warp_max(m2);
    if (lane == 0) s_max[warp] = m2;
    __syncthreads();
    if (warp == 0) {
        float v = (lane < T / 32) ? s_max[lane] : -1.f;
        v = warp_max(v);
        if (lane == 0) s_m2 = v;
    }
    __syncthreads();
    if (HALVES == 2) {
        if (threadIdx.x == 0) {
            for (uint32_t r = 0; r < 2; ++r) {
                const uint32_t dst = mapa_u32(smem_u32(&x_m2[h]), r);
                asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(dst), "r"(__float_as_uint(s_m2)) : "memory");
            }
        }
        cluster_sync_all();
        if (threadIdx.x == 0) s_m2 = fmaxf(x_m2[0], x_m2[1]);
        __syncthreads();
    }
    if (threadIdx.x == 0 && h == 0) {
        sydr_acq_row rr;
        rr.peak1 = win.v; rr.code_idx = win.i; rr.peak2 = s_m2; rr.reserved = 0;
        rows[rowid] = rr;
    }
    // ---- optional correlation-map row (drop-in mode): coalesced through shared memory
    if (maps != nullptr) {
        float* stage = reinterpret_cast<float*>(fbuf);      // NH floats, local index n' = n0 + r*NBL
#pragma unroll
        for (int q = 0; q < ROUNDS; ++q) {
            const int b = threadIdx.x + q * T;
            if (b < P::NBL) {
                const int n0 = natural_base<P>(b);
#pragma unroll
                for (int r = 0; r < RL; ++r) stage[n0 + r * P::NBL] = acc[q][r];
            }
        }
        __syncthreads();
        float* mrow = maps + (size_t)rowid * NF;
        for (int i = threadIdx.x; i < NH; i += T) mrow[HALVES * i + h] = stage[i];
    }
}

// ---- per-PRN reduction over Doppler rows ------------------------------------------------------
__global__ void acq_reduce_kernel(const sydr_acq_row* __restrict__ rows, const int* __restrict__ prns, int n_rows,
                                  int bin_lo, sydr_acq_peak* __restrict__ peaks) {
    const int slot = blockIdx.x;
    __shared__ PeakRed s_p[32];
    PeakRed best; best.v = -1.f; best.i = 0x7fffffff;
    for (int r = threadIdx.x; r < n_rows; r += blockDim.x) {
        PeakRed c; c.v = rows[slot * n_rows + r].peak1; c.i = r;
        best = peak_better(best, c);            // first maximum in C order = lowest row on ties
    }
    best = warp_peak(best);
    if ((threadIdx.x & 31) == 0) s_p[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        PeakRed p = (threadIdx.x < blockDim.x / 32) ? s_p[threadIdx.x] : PeakRed{-1.f, 0x7fffffff};
        p = warp_peak(p);
        if (threadIdx.x == 0) {
            const sydr_acq_row rr = rows[slot * n_rows + p.i];
            sydr_acq_peak o;
            o.prn = prns[slot]; o.freq_idx = bin_lo + p.i; o.code_idx = rr.code_idx;
            o.peak1 = rr.peak1; o.peak2 = rr.peak2; o.ratio = rr.peak1 / rr.peak2;
            peaks[slot] = o;
        }
    }
}

// ---- TwoCorrelationPeakComparison on a float64 map (drop-in function path) --------------------
struct PeakRed64 { double v; int i; };
__global__ void peak_rows_f64_kernel(const double* __restrict__ map, int n_code, int chip, double* __restrict__ row_p1,
                                     int* __restrict__ row_i1, double* __restrict__ row_p2) {
    const double* m = map + (size_t)blockIdx.x * n_code;
    __shared__ double sv[256];
    __shared__ int si[256];
    double bv = -1.0; int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < n_code; i += blockDim.x) {
        const double v = m[i];
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
    sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v = sv[threadIdx.x + o]; const int i = si[threadIdx.x + o];
            if (v > sv[threadIdx.x] || (v == sv[threadIdx.x] && i < si[threadIdx.x])) { sv[threadIdx.x] = v; si[threadIdx.x] = i; }
        }
        __syncthreads();
    }
    const double p1 = sv[0]; const int i1 = si[0];
    __syncthreads();
    double m2 = -1.0;
    for (int i = threadIdx.x; i < n_code; i += blockDim.x)
        if (in_second_range(i, i1, chip, n_code)) m2 = fmax(m2, m[i]);
    sv[threadIdx.x] = m2;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sv[threadIdx.x] = fmax(sv[threadIdx.x], sv[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) { row_p1[blockIdx.x] = p1; row_i1[blockIdx.x] = i1; row_p2[blockIdx.x] = sv[0]; }
}

}  // namespace sydr

using namespace sydr;

// ------------------------------------------------------------------------------------------
// Plan object
// ------------------------------------------------------------------------------------------
struct sydr_acq_plan {
    AcqDev dev;
    int n_bins_total, bin_hi;
    int halves;                  // 1 or 2
    int nh;                      // transform length per CTA
    std::vector<int> prns;
    int* d_prns;
    float2* d_code_spec;
    float2* d_Y;
    float2* d_tw;
    float2* d_tw_split;
    sydr_acq_row* d_rows;        // internal row summaries
    int device;
};

namespace {

// Plans for the supported code lengths (samples per 1 ms C/A period).
using Plan2000  = Plan<2000, 256, 10, 10, 5, 4, 1>;
using Plan4000  = Plan<4000, 256, 10, 10, 10, 4, 1>;
using Plan5000  = Plan<5000, 512, 10, 10, 10, 5, 1>;
using Plan8000  = Plan<8000, 512, 10, 10, 10, 8, 1>;
using Plan10000 = Plan<10000, 512, 10, 10, 10, 10, 1>;
using Plan12500 = Plan<12500, 512, 10, 10, 5, 5, 5>;
using Plan16000 = Plan<16000, 512, 10, 10, 10, 16, 1>;
using Plan20000 = Plan<20000, 512, 10, 10, 10, 20, 1>;
using Plan25000 = Plan<25000, 512, 10, 10, 10, 25, 1>;

struct PlanShape { int nh, halves; };

bool plan_shape(int n_code, PlanShape* s) {
    switch (n_code) {
        case 2000: case 4000: case 5000: case 8000: case 10000: case 12500: case 16000: case 20000: case 25000:
            s->nh = n_code; s->halves = 1; return true;
        case 32000: case 40000: case 50000:
            s->nh = n_code / 2; s->halves = 2; return true;
        default: return false;
    }
}

template <class P>
void plan_stage_sizes(int* L) {   // twiddle entries per stage
    L[0] = P::M1 / P::R1; L[1] = P::M2 / P::R2; L[2] = P::M3 / P::R3; L[3] = P::M4 / P::R4;
}

template <class P>
int build_twiddles(sydr_acq_plan* pl) {
    std::vector<float2> tw(P::TWN);
    const int Ms[4] = {P::M1, P::M2, P::M3, P::M4};
    const int off[4] = {P::TW1, P::TW2, P::TW3, P::TW4};
    int L[4];
    plan_stage_sizes<P>(L);
    for (int s = 0; s < 4; ++s)
        for (int k = 0; k < L[s]; ++k) {
            const double a = -2.0 * M_PI * (double)k / (double)Ms[s];
            tw[off[s] + k] = make_float2((float)cos(a), (float)sin(a));
        }
    SYDR_CUDA_CHECK(cudaMalloc(&pl->d_tw, sizeof(float2) * P::TWN));
    SYDR_CUDA_CHECK(cudaMemcpy(pl->d_tw, tw.data(), sizeof(float2) * P::TWN, cudaMemcpyHostToDevice));
    return SYDR_OK;
}

template <class P, int HALVES>
int launch_acq(sydr_acq_plan* pl, const void* d_iq, int dt, sydr_acq_row* d_rows, float* d_maps, cudaStream_t s) {
    const AcqDev& A = pl->dev;
    const size_t fft_bytes = sizeof(float2) * P::N;
    // forward: one CTA per (row, block[, half])
    {
        auto k = acq_fwd_kernel<P, HALVES>;
        SYDR_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fft_bytes));
        const int fwd_rows = (A.n_base > 0) ? A.n_base : A.n_rows;
        k<<<fwd_rows * A.noncoh * HALVES, P::T, fft_bytes, s>>>(A, d_iq, dt);
        count_launch();
        SYDR_CUDA_CHECK(cudaGetLastError());
    }
    // inverse: one CTA (or CTA pair) per (PRN, row)
    {
        constexpr bool kCodeSmem = (HALVES == 1) && (sizeof(float2) * P::N * 2 <= 200 * 1024);
        const size_t bytes = kCodeSmem ? 2 * fft_bytes : fft_bytes;
        auto k = acq_ifft_kernel<P, HALVES, kCodeSmem>;
        SYDR_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)(A.n_prn * A.n_rows * HALVES));
        lc.blockDim = dim3(P::T);
        lc.dynamicSmemBytes = bytes;
        lc.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = HALVES;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        lc.attrs = at;
        lc.numAttrs = 1;
        SYDR_CUDA_CHECK(cudaLaunchKernelEx(&lc, k, A, d_rows, d_maps));
        count_launch();
    }
    return SYDR_OK;
}

int dispatch_acq(sydr_acq_plan* pl, const void* d_iq, int dt, sydr_acq_row* d_rows, float* d_maps, cudaStream_t s) {
    switch (pl->dev.n_code) {
        case 2000: return launch_acq<Plan2000, 1>(pl, d_iq, dt, d_rows, d_maps, s);
        case 4000: return launch_acq<Plan4000, 1>(pl, d_iq, dt, d_rows, d_maps, s);
        case 5000: return launch_acq<Plan5000, 1>(pl, d_iq, dt, d_rows, d_maps, s);
        case 8000: return launch_acq<Plan8000, 1>(pl, d_iq, dt, d_rows, d_maps, s);
        case 10000: return launch_acq<Plan10000, 1>(pl, d_iq, dt, d_rows, d_maps, s);
        case 12500: return launch_acq<Plan12500, 1>(pl, d_iq, dt, d_rows, d_maps, s);
        case 16000: return launch_acq<Plan16000, 1>(pl, d_iq, dt, d_rows, d_maps, s);
        case 20000: return launch_acq<Plan20000, 1>(pl, d_iq, dt, d_rows, d_maps, s);
        case 25000: return launch_acq<Plan25000, 1>(pl, d_iq, dt, d_rows, d_maps, s);
        case 32000: return launch_acq<Plan16000, 2>(pl, d_iq, dt, d_rows, d_maps, s);
        case 40000: return launch_acq<Plan20000, 2>(pl, d_iq, dt, d_rows, d_maps, s);
        case 50000: return launch_acq<Plan25000, 2>(pl, d_iq, dt, d_rows, d_maps, s);
        default:
            set_error("no FFT plan for %d samples per code", pl->dev.n_code);
            return SYDR_ERR_UNSUPPORTED;
    }
}

int dispatch_twiddles(sydr_acq_plan* pl) {
    switch (pl->nh) {
        case 2000: return build_twiddles<Plan2000>(pl);
        case 4000: return build_twiddles<Plan4000>(pl);
        case 5000: return build_twiddles<Plan5000>(pl);
        case 8000: return build_twiddles<Plan8000>(pl);
        case 10000: return build_twiddles<Plan10000>(pl);
        case 12500: return build_twiddles<Plan12500>(pl);
        case 16000: return build_twiddles<Plan16000>(pl);
        case 20000: return build_twiddles<Plan20000>(pl);
        case 25000: return build_twiddles<Plan25000>(pl);
        default: set_error("no FFT plan for half length %d", pl->nh); return SYDR_ERR_UNSUPPORTED;
    }
}

// Code spectra of `prns` at n_code samples: conj(fft(upsampled code)).  Either output may be NULL.
int compute_code_spectra(const int* h_prns, int n_prn, double fs, int n_code, double scale, float2* d_out32,
                         double2* d_out64, cudaStream_t s) {
    CodeTables t;
    int rc = ensure_code_tables(&t);
    if (rc != SYDR_OK) return rc;
    int* d_prns = nullptr;
    double2 *a = nullptr, *b = nullptr, *res = nullptr;
    const size_t total = (size_t)n_prn * n_code;
    SYDR_CUDA_CHECK(cudaMalloc(&d_prns, sizeof(int) * n_prn));
    SYDR_CUDA_CHECK(cudaMalloc(&a, sizeof(double2) * total));
    SYDR_CUDA_CHECK(cudaMalloc(&b, sizeof(double2) * total));
    SYDR_CUDA_CHECK(cudaMemcpyAsync(d_prns, h_prns, sizeof(int) * n_prn, cudaMemcpyHostToDevice, s));
    dim3 grid((n_code + 255) / 256, n_prn);
    upsample_code_kernel<<<grid, 256, 0, s>>>(t.chips, d_prns, n_code, fs, a);
    count_launch();
    SYDR_CUDA_CHECK(cudaGetLastError());
    rc = fft64_batched(a, b, n_code, n_prn, s, &res);
    if (rc == SYDR_OK) {
        spectrum_finish_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(res, (int)total, scale, d_out32, d_out64);
        count_launch();
        if (cudaGetLastError() != cudaSuccess) rc = SYDR_ERR_CUDA;
    }
    cudaStreamSynchronize(s);
    cudaFree(d_prns);
    cudaFree(a);
    cudaFree(b);
    return rc;
}

}  // namespace

namespace sydr {
// FP64 forward FFT of host rows (interleaved complex128 in and out); used by the legacy setSatellite.
int fft64_host_rows(const double* h_in, int n, int batch, double* h_out) {
    SYDR_REQUIRE(h_in && h_out && n > 0 && batch > 0, SYDR_ERR_ARG, "fft64_host_rows: bad arguments");
    double2 *a = nullptr, *b = nullptr, *res = nullptr;
    const size_t bytes = sizeof(double2) * (size_t)n * batch;
    SYDR_CUDA_CHECK(cudaMalloc(&a, bytes));
    SYDR_CUDA_CHECK(cudaMalloc(&b, bytes));
    int rc = SYDR_OK;
    if (cudaMemcpy(a, h_in, bytes, cudaMemcpyHostToDevice) != cudaSuccess) rc = SYDR_ERR_CUDA;
    if (rc == SYDR_OK) rc = fft64_batched(a, b, n, batch, 0, &res);
    if (rc == SYDR_OK && cudaMemcpy(h_out, res, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) rc = SYDR_ERR_CUDA;
    if (rc == SYDR_ERR_CUDA) set_error("fft64_host_rows: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(a);
    cudaFree(b);
    return rc;
}
}  // namespace sydr

extern "C" {

int sydr_code_spectrum(int prn, double fs, double* h_spectrum_c128, long long n_code) {
    SYDR_REQUIRE(prn >= 1 && prn <= kMaxPrn, SYDR_ERR_ARG, "PRN %d out of range", prn);
    SYDR_REQUIRE(h_spectrum_c128 != nullptr && n_code > 0 && n_code <= (1 << 22), SYDR_ERR_ARG, "bad output/n_code");
    const long long expect = llrint(fs / (kCodeFreq / kCodeChips));
    SYDR_REQUIRE(expect == n_code, SYDR_ERR_ARG, "n_code %lld does not match round(fs/1000) = %lld", n_code, expect);
    double2* d_out = nullptr;
    SYDR_CUDA_CHECK(cudaMalloc(&d_out, sizeof(double2) * n_code));
    int rc = compute_code_spectra(&prn, 1, fs, (int)n_code, 1.0, nullptr, d_out, 0);
    if (rc == SYDR_OK && cudaMemcpy(h_spectrum_c128, d_out, sizeof(double2) * n_code, cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_error("copy of code spectrum failed");
        rc = SYDR_ERR_CUDA;
    }
    cudaFree(d_out);
    return rc;
}

int sydr_acq_plan_create(double fs, double inter_freq, double doppler_range, double doppler_step, int coh, int noncoh,
                         const int* h_prns, int n_prn, int bin_lo, int bin_hi, sydr_acq_plan** out_plan) {
    SYDR_REQUIRE(out_plan && h_prns, SYDR_ERR_ARG, "NULL pointer");
    SYDR_REQUIRE(fs > 0 && doppler_step > 0 && doppler_range >= 0, SYDR_ERR_ARG, "bad fs/doppler arguments");
    SYDR_REQUIRE(coh >= 1 && noncoh >= 1 && n_prn >= 1, SYDR_ERR_ARG, "coh, noncoh and n_prn must be >= 1");
    for (int i = 0; i < n_prn; ++i)
        SYDR_REQUIRE(h_prns[i] >= 1 && h_prns[i] <= kMaxPrn, SYDR_ERR_ARG, "PRN %d out of range", h_prns[i]);
    // samplesPerCode / samplesPerCodeChip: channel_l1ca_borre.py:283-284 (Python round = nearbyint)
    const int n_code = (int)nearbyint(fs * kCodeChips / kCodeFreq);
    const int chip = (int)nearbyint(fs / kCodeFreq);
    // len(np.arange(-R, R+1, step)) = ceil((2R+1)/step)
    const int n_bins = (int)ceil((2.0 * doppler_range + 1.0) / doppler_step);
    if (bin_hi < 0) bin_hi = n_bins;
    SYDR_REQUIRE(bin_lo >= 0 && bin_lo < bin_hi && bin_hi <= n_bins, SYDR_ERR_ARG, "bad bin range [%d,%d) of %d", bin_lo,
                 bin_hi, n_bins);
    PlanShape shape;
    SYDR_REQUIRE(plan_shape(n_code, &shape), SYDR_ERR_UNSUPPORTED,
                 "no FFT plan for %d samples per code (fs = %.0f Hz)", n_code, fs);

    sydr_acq_plan* pl = new sydr_acq_plan();
    memset(&pl->dev, 0, sizeof(pl->dev));
    pl->prns.assign(h_prns, h_prns + n_prn);
    pl->n_bins_total = n_bins;
    pl->bin_hi = bin_hi;
    pl->halves = shape.halves;
    pl->nh = shape.nh;
    pl->d_prns = nullptr; pl->d_code_spec = nullptr; pl->d_Y = nullptr; pl->d_tw = nullptr; pl->d_tw_split = nullptr;
    pl->d_rows = nullptr;
    cudaGetDevice(&pl->device);
    const int n_rows = bin_hi - bin_lo;
    int rc = SYDR_OK;
    auto fail = [&](int code) { sydr_acq_plan_destroy(pl); return code; };
    // rows R and R + G differ by G * step = fs / N when that ratio is a whole number: they share forward spectra
    int n_base = 0;
    {
        const double g = (fs / (double)n_code) / doppler_step;
        const int G = (int)llround(g);
        n_base = (G >= 1 && fabs(g - (double)G) < 1e-9) ? G : 0;    // global base rows, whatever the shard
    }
    if (cudaMalloc(&pl->d_prns, sizeof(int) * n_prn) != cudaSuccess ||
        cudaMalloc(&pl->d_code_spec, sizeof(float2) * (size_t)n_prn * n_code) != cudaSuccess ||
        cudaMalloc(&pl->d_Y, sizeof(float2) * (size_t)(n_base > 0 ? n_base : n_rows) * noncoh * n_code) != cudaSuccess ||
        cudaMalloc(&pl->d_rows, sizeof(sydr_acq_row) * (size_t)n_prn * n_rows) != cudaSuccess) {
        set_error("acquisition plan: device allocation failed (%s)", cudaGetErrorString(cudaGetLastError()));
        return fail(SYDR_ERR_CUDA);
    }
    cudaMemcpy(pl->d_prns, h_prns, sizeof(int) * n_prn, cudaMemcpyHostToDevice);
    rc = compute_code_spectra(h_prns, n_prn, fs, n_code, 1.0 / (double)n_code, pl->d_code_spec, nullptr, 0);
    if (rc != SYDR_OK) return fail(rc);
    rc = dispatch_twiddles(pl);
    if (rc != SYDR_OK) return fail(rc);
    if (shape.halves == 2) {
        std::vector<float2> tw(shape.nh);
        for (int e = 0; e < shape.nh; ++e) {
            const double a = -2.0 * M_PI * (double)e / (double)n_code;
            tw[e] = make_float2((float)cos(a), (float)sin(a));
        }
        if (cudaMalloc(&pl->d_tw_split, sizeof(float2) * shape.nh) != cudaSuccess) return fail(SYDR_ERR_CUDA);
        cudaMemcpy(pl->d_tw_split, tw.data(), sizeof(float2) * shape.nh, cudaMemcpyHostToDevice);
    }
    AcqDev& A = pl->dev;
    A.code_spec = pl->d_code_spec; A.Y = pl->d_Y; A.tw = pl->d_tw; A.tw_split = pl->d_tw_split;
    A.n_code = n_code; A.n_prn = n_prn; A.n_rows = n_rows; A.bin_lo = bin_lo; A.coh = coh; A.noncoh = noncoh;
    A.chip = chip; A.fs = fs; A.inter_freq = inter_freq; A.doppler_range = doppler_range; A.doppler_step = doppler_step;
    A.n_base = n_base;
    *out_plan = pl;
    return SYDR_OK;
}

int sydr_acq_plan_destroy(sydr_acq_plan* pl) {
    if (!pl) return SYDR_OK;
    cudaFree(pl->d_prns); cudaFree(pl->d_code_spec); cudaFree(pl->d_Y); cudaFree(pl->d_tw); cudaFree(pl->d_tw_split);
    cudaFree(pl->d_rows);
    delete pl;
    return SYDR_OK;
}

int sydr_acq_plan_info(const sydr_acq_plan* pl, int* n_code, int* n_bins_total, int* n_rows_local, int* samples_per_chip,
                       long long* required_samples) {
    SYDR_REQUIRE(pl != nullptr, SYDR_ERR_ARG, "plan is NULL");
    if (n_code) *n_code = pl->dev.n_code;
    if (n_bins_total) *n_bins_total = pl->n_bins_total;
    if (n_rows_local) *n_rows_local = pl->dev.n_rows;
    if (samples_per_chip) *samples_per_chip = pl->dev.chip;
    if (required_samples) *required_samples = (long long)pl->dev.n_code * pl->dev.coh * pl->dev.noncoh;
    return SYDR_OK;
}

int sydr_acq_plan_set_spectrum(sydr_acq_plan* pl, int prn_slot, const double* h_spec) {
    SYDR_REQUIRE(pl && h_spec, SYDR_ERR_ARG, "NULL pointer");
    SYDR_REQUIRE(prn_slot >= 0 && prn_slot < pl->dev.n_prn, SYDR_ERR_ARG, "prn_slot %d out of range", prn_slot);
    const int n = pl->dev.n_code;
    std::vector<float2> v(n);
    const double sc = 1.0 / (double)n;
    for (int i = 0; i < n; ++i) v[i] = make_float2((float)(h_spec[2 * i] * sc), (float)(h_spec[2 * i + 1] * sc));
    SYDR_CUDA_CHECK(cudaMemcpy(pl->d_code_spec + (size_t)prn_slot * n, v.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
    return SYDR_OK;
}

int sydr_acq_run(sydr_acq_plan* pl, const void* d_iq, int iq_dtype, long long n_samples, sydr_acq_peak* d_peaks,
                 sydr_acq_row* d_rows, float* d_maps, void* stream) {
    SYDR_REQUIRE(pl && d_iq, SYDR_ERR_ARG, "NULL pointer");
    SYDR_REQUIRE(iq_dtype >= SYDR_IQ_I8 && iq_dtype <= SYDR_IQ_F64, SYDR_ERR_ARG, "bad iq_dtype %d", iq_dtype);
    const long long need = (long long)pl->dev.n_code * pl->dev.coh * pl->dev.noncoh;
    SYDR_REQUIRE(n_samples >= need, SYDR_ERR_ARG, "acquisition needs %lld samples, got %lld", need, n_samples);
    cudaStream_t s = (cudaStream_t)stream;
    sydr_acq_row* rows = d_rows ? d_rows : pl->d_rows;
    int rc = dispatch_acq(pl, d_iq, iq_dtype, rows, d_maps, s);
    if (rc != SYDR_OK) return rc;
    if (d_peaks) {
        acq_reduce_kernel<<<pl->dev.n_prn, 128, 0, s>>>(rows, pl->d_prns, pl->dev.n_rows, pl->dev.bin_lo, d_peaks);
        count_launch();
        SYDR_CUDA_CHECK(cudaGetLastError());
    }
    return SYDR_OK;
}

int sydr_acq_reduce_rows(const sydr_acq_row* h_rows, const int* h_prns, int n_prn, int n_bins, sydr_acq_peak* h_peaks) {
    SYDR_REQUIRE(h_rows && h_prns && h_peaks, SYDR_ERR_ARG, "NULL pointer");
    for (int p = 0; p < n_prn; ++p) {
        int best = 0;
        for (int r = 1; r < n_bins; ++r)
            if (h_rows[p * n_bins + r].peak1 > h_rows[p * n_bins + best].peak1) best = r;
        const sydr_acq_row& rr = h_rows[p * n_bins + best];
        h_peaks[p].prn = h_prns[p]; h_peaks[p].freq_idx = best; h_peaks[p].code_idx = rr.code_idx;
        h_peaks[p].peak1 = rr.peak1; h_peaks[p].peak2 = rr.peak2; h_peaks[p].ratio = rr.peak1 / rr.peak2;
    }
    return SYDR_OK;
}

int sydr_peak_compare(const double* h_map, int n_bins, int n_code, int chip, int* h_freq_idx, int* h_code_idx,
                      double* h_ratio) {
    SYDR_REQUIRE(h_map && h_freq_idx && h_code_idx && h_ratio, SYDR_ERR_ARG, "NULL pointer");
    SYDR_REQUIRE(n_bins >= 1 && n_code >= 2, SYDR_ERR_ARG, "bad map shape %d x %d", n_bins, n_code);
    double *d_map = nullptr, *d_p1 = nullptr, *d_p2 = nullptr;
    int* d_i1 = nullptr;
    const size_t total = (size_t)n_bins * n_code;
    SYDR_CUDA_CHECK(cudaMalloc(&d_map, sizeof(double) * total));
    SYDR_CUDA_CHECK(cudaMalloc(&d_p1, sizeof(double) * n_bins));
    SYDR_CUDA_CHECK(cudaMalloc(&d_p2, sizeof(double) * n_bins));
    SYDR_CUDA_CHECK(cudaMalloc(&d_i1, sizeof(int) * n_bins));
    SYDR_CUDA_CHECK(cudaMemcpy(d_map, h_map, sizeof(double) * total, cudaMemcpyHostToDevice));
    peak_rows_f64_kernel<<<n_bins, 256>>>(d_map, n_code, chip, d_p1, d_i1, d_p2);
    count_launch();
    SYDR_CUDA_CHECK(cudaGetLastError());
    std::vector<double> p1(n_bins), p2(n_bins);
    std::vector<int> i1(n_bins);
    SYDR_CUDA_CHECK(cudaMemcpy(p1.data(), d_p1, sizeof(double) * n_bins, cudaMemcpyDeviceToHost));
    SYDR_CUDA_CHECK(cudaMemcpy(p2.data(), d_p2, sizeof(double) * n_bins, cudaMemcpyDeviceToHost));
    SYDR_CUDA_CHECK(cudaMemcpy(i1.data(), d_i1, sizeof(int) * n_bins, cudaMemcpyDeviceToHost));
    cudaFree(d_map); cudaFree(d_p1); cudaFree(d_p2); cudaFree(d_i1);
    int best = 0;
    for (int r = 1; r < n_bins; ++r)
        if (p1[r] > p1[best]) best = r;                // first maximum in C order
    *h_freq_idx = best;
    *h_code_idx = i1[best];
    *h_ratio = p1[best] / p2[best];
    return SYDR_OK;
}

}  // extern "C"
