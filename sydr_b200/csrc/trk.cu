// K-TRK: early/prompt/late correlators with carrier + code NCO replica generation and the
// Borre DLL/PLL loop closure, for sm_100a.
//
// Reference semantics (file:line in /root/reference):
//   EPL                     sydr/dsp/tracking.py:92-116
//   DLL_NNEML / PLL_costa   sydr/dsp/tracking.py:120-142
//   BorreLoopFilter         sydr/dsp/tracking.py:180-186
//   NCO update              sydr/channel/channel_l1ca_borre.py:363-429
//
// Structure.  One channel = one CTA, or one thread-block cluster of S CTAs when the channel
// count is small and per-epoch latency is what matters (epoch k+1 needs epoch k's loop-filter
// output, so epochs of a channel are a serial chain).  Each thread owns contiguous chunks of
// C samples: the carrier phase is seeded in FP64 per chunk and advanced by an FP32 phasor
// recurrence; the code chip under every sample is taken from a per-chunk sign mask whose chip
// transitions are located with the reference's exact FP64 expression
// ceil(fl(fl(i*step') + start)), so no sample is ever assigned to the wrong chip.
// IQ samples stay int8/int16 until they are in registers.  In the closed-loop kernel each
// CTA's slice of the *next* epoch is staged into shared memory by a TMA bulk copy
// (cp.async.bulk + mbarrier) while the current epoch is being correlated; cluster partial
// sums are all-gathered with st.async (DSMEM store + remote mbarrier complete_tx), after which
// every CTA closes the loops redundantly in FP64 (bit-identical), so one exchange per epoch
// is the only inter-CTA synchronisation.
#include "common.cuh"

namespace sydr {

// Samples per thread chunk: a multiple of the 16-byte vector and of 4, chosen so that the
// per-thread stride in shared memory (48 / 160 bytes) keeps LDS.128 (nearly) conflict free and
// that a cluster of 8 CTAs still has ~2 warps per scheduler at 25 MS/s.
// VPC (vectors per chunk) is a kernel template parameter: int8 3 (24 samples, 48 B); int16 3
// (12 samples, 48 B; latency mode, clusters of >= 4 CTAs) or 5 (20 samples, 80 B; throughput
// mode); complex64 10 (20 samples, 160 B).

struct EpochConst {
    double ca, cb;         // carrier phase in turns at sample j: ca*j + cb  (ca = -fc/fs, cb = rem/(2 pi))
    double start[3];       // linspace start  = remCode + spacing     tracking.py:110
    double step[3];        // linspace step'  = (stop-start)/n        numpy linspace
    double inv_step[3];    // ~1/step' (locates chip boundaries; every boundary is then pinned exactly)
    float w[4][2];         // carrier rotation by 1, 2, 3, 4 samples: exp(-j 2 pi k fc/fs)
    int n;                 // samples in the epoch
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

// a / b by Newton-Raphson on the hardware reciprocal seed with an FMA residual correction
// (<= 1 ulp, branch free; ~10 instructions instead of the ~40 of the IEEE division routine).
// Operands here are always finite, normal and non-zero.
__device__ __forceinline__ double drcp(double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(fma(-b, r, 1.0), r, r);
    r = fma(fma(-b, r, 1.0), r, r);
    return r;
}
__device__ __forceinline__ double ddiv(double a, double b) {
    const double r = drcp(b);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}

// int32 -> double and round-to-nearest-integer without touching the conversion pipe
// (magic-number tricks; exact for |j| < 2^31 and |x| < 2^51).
__device__ __forceinline__ double i2d(int j) {
    return __hiloint2double(0x43300000, (int)((unsigned)j ^ 0x80000000u)) - 4503601774854144.0;
}
__device__ __forceinline__ double drint(double x) {
    return (x + 6755399441055744.0) - 6755399441055744.0;
}
// fl(fl(j*step') + start): the value numpy's linspace produces for sample j.
__device__ __forceinline__ double code_phase(int j, double start, double step) {
    return dadd(dmul(i2d(j), step), start);
}
// ceil() of a double in (-2^31, 2^31): adding 1.5*2^52 with round-up leaves ceil(x) in the low word.
__device__ __forceinline__ int ceil_to_int(double x) {
    return __double2loint(__dadd_ru(x, 6755399441055744.0));
}
// Padded-code lookup with Python index semantics (negative wraps once; beyond 1024 is the
// reference's IndexError -> flagged).
__device__ __forceinline__ uint32_t chip_bit(const uint32_t* cb, int k, int& err) {
    if (k < 0) k += kPaddedChips;
    if (k < 0 || k >= kPaddedChips) { err = 1; k = min(max(k, 0), kPaddedChips - 1); }
    return (cb[k >> 5] >> (k & 31)) & 1u;
}

// First sample t in (jlo, jhi] whose code phase exceeds chip k, given an estimate: walks to the
// exact answer with the reference's own expression.
__device__ __noinline__ int boundary_exact(int t, int jlo, int jhi, int k, double start, double step) {
    t = max(jlo + 1, min(t, jhi));
    while (t > jlo + 1 && code_phase(t - 1, start, step) > i2d(k)) --t;
    while (t < jhi && !(code_phase(t, start, step) > i2d(k))) ++t;
    return t;
}

// Generic exact walk (Python wrap-around / out-of-range indices); rare.
__device__ __noinline__ uint32_t tap_mask_walk(int jlo, int jhi, int k0, double start, double step,
                                               const uint32_t* cb, int* err) {
    const int k1 = ceil_to_int(code_phase(jhi, start, step));
    int k = k0, e = 0;
    uint32_t cur = chip_bit(cb, k, e);
    uint32_t m = cur ? 0xffffffffu : 0u;
    for (int guard = 0; k < k1 && guard < 64; ++guard) {
        const int t = boundary_exact(jlo + 1, jlo, jhi, k, start, step);
        const int kn = ceil_to_int(code_phase(t, start, step));
        const uint32_t nxt = chip_bit(cb, kn, e);
        if (nxt != cur) m ^= 0xffffffffu << (t - jlo);
        cur = nxt;
        k = kn;
    }
    if (e) *err = 1;
    return m;
}

// Sign mask for samples [jlo, jlo+cnt) of correlator tap `s`; bit i set = chip +1.
// The chip under the first sample comes from the exact FP64 expression
// ceil(fl(fl(j*step')+start)); each following chip boundary is located from the real-valued
// crossing (k-start)/step' and re-derived with the exact expression whenever the crossing is
// within 1e-6 sample of an integer (rounding of the reference expression moves a boundary by
// < 1e-10 sample), so the mask equals code[ceil(linspace(...))] sample for sample.
// jd = (double)jlo.  One copy of the code serves the three taps (instruction-cache footprint).
__device__ __noinline__ uint32_t tap_mask(int jlo, double jd, int cnt, const EpochConst* ec, int s,
                                          const uint32_t* cb, int* err) {
    const double start = ec->start[s], step = ec->step[s], inv_step = ec->inv_step[s];
    const int jhi = jlo + cnt - 1;
    const int k0 = ceil_to_int(dadd(dmul(jd, step), start));
    if (k0 < 0 || k0 > kPaddedChips - 1) return tap_mask_walk(jlo, jhi, k0, start, step, cb, err);
    // window of padded-code bits k0 .. k0+31
    const int wi = k0 >> 5;
    const uint32_t w = __funnelshift_r(cb[wi], cb[wi + 1], k0 & 31);
    uint32_t m = (w & 1u) ? 0xffffffffu : 0u;
    for (int b = 0; b < 31; ++b) {
        const int k = k0 + b;
        const double js = dmul(dsub(i2d(k), start), inv_step);     // crossing of chip k, in samples
        int t = ceil_to_int(js);                                     // first integer >= js
        const double d = i2d(t) - js;                                // [0, 1)
        if (!(d > 1e-6 && d < 0.999999)) {
            if (t > jhi + 1) break;                                  // clearly beyond this chunk
            if (!(code_phase(jhi, start, step) > i2d(k))) break;     // exact: chip k lasts beyond the chunk
            t = boundary_exact(t, jlo, jhi, k, start, step);
        } else if (t > jhi) {
            break;
        }
        t = max(t, jlo + 1);
        if (k + 1 > kPaddedChips - 1) *err = 1;                      // the reference's IndexError
        if (((w >> b) ^ (w >> (b + 1))) & 1u) m ^= 0xffffffffu << (t - jlo);
    }
    return m;
}

__device__ __forceinline__ float chip_value(uint32_t mask, int i) {
    // +1.0f when mask bit i is set, -1.0f otherwise
    return __uint_as_float(0x3f800000u | ((~mask << (31 - i)) & 0x80000000u));
}

// Zero the samples of one vector that lie outside [vlo, vhi) (vector-relative; epoch start / end).
template <int DT>
__device__ __forceinline__ void mask_vector(uint4& raw, int vlo, int vhi) {
    uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (DT == SYDR_IQ_I16) {
            w[k] = (k >= vlo && k < vhi) ? w[k] : 0u;
        } else if (DT == SYDR_IQ_I8) {
            const uint32_t keep = ((2 * k >= vlo && 2 * k < vhi) ? 0x0000ffffu : 0u) |
                                  ((2 * k + 1 >= vlo && 2 * k + 1 < vhi) ? 0xffff0000u : 0u);
            w[k] &= keep;
        } else {
            w[k] = ((k >> 1) >= vlo && (k >> 1) < vhi) ? w[k] : 0u;
        }
    }
    raw = make_uint4(w[0], w[1], w[2], w[3]);
}

// Correlate one chunk of C = VPC*SPV samples starting at epoch-relative index j0 against the
// three taps.  `src` points at the chunk's first vector (shared or global memory, 16-byte
// aligned); `ec` lives in shared memory.  acc = {IE, QE, IP, QP, IL, QL}.
template <int DT, int VPC>
__device__ __forceinline__ void correlate_chunk(const uint4* src, int j0, const EpochConst& ec,
                                                const uint32_t* cb, float* acc, int& err) {
    constexpr int SPV = IqTraits<DT>::SPV;
    constexpr int C = SPV * VPC;
    const int lo = max(j0, 0);
    const int hi = min(j0 + C, ec.n);
    if (hi <= lo) return;
    const bool edge = (lo != j0) || (hi != j0 + C);
    uint4 cur = src[0];

    const double jd = i2d(lo);
    uint32_t m0 = tap_mask(lo, jd, hi - lo, &ec, 0, cb, &err) << (lo - j0);
    uint32_t m1 = tap_mask(lo, jd, hi - lo, &ec, 1, cb, &err) << (lo - j0);
    uint32_t m2 = tap_mask(lo, jd, hi - lo, &ec, 2, cb, &err) << (lo - j0);

// Carrier seed (tracking.py:102): phase of sample j0 in turns, FP64, reduced to [-0.5, 0.5].
    double turns = fma(ec.ca, i2d(j0), ec.cb);
    turns -= drint(turns);
    float pre, pim;
    __sincosf((float)turns * 6.283185307179586f, &pim, &pre);

    // rotations by 1..4 samples
    const float w1r = ec.w[0][0], w1i = ec.w[0][1], w2r = ec.w[1][0], w2i = ec.w[1][1];
    const float w3r = ec.w[2][0], w3i = ec.w[2][1], w4r = ec.w[3][0], w4i = ec.w[3][1];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f;
    constexpr int G = (SPV >= 4) ? 4 : SPV;                  // samples sharing one phasor base
#pragma unroll 1
    for (int v = 0; v < VPC; ++v) {
        const uint4 nxt = src[min(v + 1, VPC - 1)];          // next vector in flight while this one is used
        if (edge) mask_vector<DT>(cur, lo - (j0 + v * SPV), hi - (j0 + v * SPV));
        float re[SPV], im[SPV];
        decode_vec<DT>(cur, re, im);
        cur = nxt;
#pragma unroll
        for (int u = 0; u < SPV; ++u) {
            // carrier phasor of this sample: group base (pre, pim) times w^(u % G)
            float cr = pre, ci = pim;
            if ((u % G) == 1) { cr = pre * w1r - pim * w1i; ci = pre * w1i + pim * w1r; }
            if ((u % G) == 2) { cr = pre * w2r - pim * w2i; ci = pre * w2i + pim * w2r; }
            if ((u % G) == 3) { cr = pre * w3r - pim * w3i; ci = pre * w3i + pim * w3r; }
            // signal = replica * rfData                               tracking.py:105
            const float sr = re[u] * cr - im[u] * ci;
            const float si = re[u] * ci + im[u] * cr;
            const float c0 = chip_value(m0, u), c1 = chip_value(m1, u), c2 = chip_value(m2, u);
            a0 = fmaf(c0, sr, a0); a1 = fmaf(c0, si, a1);
            a2 = fmaf(c1, sr, a2); a3 = fmaf(c1, si, a3);
            a4 = fmaf(c2, sr, a4); a5 = fmaf(c2, si, a5);
            if ((u % G) == G - 1) {                          // advance the base by G samples
                const float gr = (G == 4) ? w4r : w2r, gi = (G == 4) ? w4i : w2i;
                const float t = pre * gr - pim * gi;
                pim = pre * gi + pim * gr;
                pre = t;
            }
        }
        m0 >>= SPV; m1 >>= SPV; m2 >>= SPV;
    }
    acc[0] += a0; acc[1] += a1; acc[2] += a2; acc[3] += a3; acc[4] += a4; acc[5] += a5;
}

// Tap constants of one correlator (numpy linspace arithmetic, tracking.py:110-112).
__device__ __forceinline__ void tap_const(double rem_code, double spacing, double code_step, int n,
                                          double& start, double& step, double& inv_step) {
    const double dn = i2d(n);
    start = dadd(rem_code, spacing);                           // shift
    const double stop = dadd(dmul(code_step, dn), start);      // codeStep*n + shift
    step = ddiv(dsub(stop, start), dn);                        // linspace step
    inv_step = drcp(step);
}
// Carrier constants: turns(j) = ca*j + cb; rotations by 1..4 samples.
__device__ __forceinline__ void carrier_const(double fc, double rem_carrier, double inv_fs, double& ca,
                                              double& cb, float (*w)[2]) {
    ca = -(fc * inv_fs);
    cb = rem_carrier * 0.15915494309189535;                    // 1/(2 pi)
    const double wt = ca - drint(ca);
    float s1, c1;
    sincospif((float)(2.0 * wt), &s1, &c1);
    const float c2 = c1 * c1 - s1 * s1, s2 = 2.f * c1 * s1;
    w[0][0] = c1; w[0][1] = s1;
    w[1][0] = c2; w[1][1] = s2;
    w[2][0] = c2 * c1 - s2 * s1; w[2][1] = c2 * s1 + s2 * c1;
    w[3][0] = c2 * c2 - s2 * s2; w[3][1] = 2.f * c2 * s2;
}

// Per-epoch constants from the NCO state (one thread; open-loop kernel).
__device__ __forceinline__ void make_epoch_const(EpochConst& ec, int n, double fs, double fc,
                                                 double rem_carrier, double rem_code,
                                                 double code_step, const double* spacing) {
    ec.n = n;
#pragma unroll
    for (int s = 0; s < 3; ++s) tap_const(rem_code, spacing[s], code_step, n, ec.start[s], ec.step[s], ec.inv_step[s]);
    carrier_const(fc, rem_carrier, 1.0 / fs, ec.ca, ec.cb, ec.w);
}

// Sum eight per-thread values over a warp with 9 shuffles (halving the value set at each of
// the first three butterfly levels).  On return lane L holds the warp total of value index
// ((L>>4)&1)*4 + ((L>>3)&1)*2 + ((L>>2)&1).
__device__ __forceinline__ float warp_sum8(const float* v, int lane) {
    float a[4], b[2], c;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h16 ? v[i] : v[i + 4];
        const float keep = h16 ? v[i + 4] : v[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h8 ? a[i] : a[i + 2];
        const float keep = h8 ? a[i + 2] : a[i];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        const float send = h4 ? b[0] : b[1];
        const float keep = h4 ? b[1] : b[0];
        c = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    return c;
}
__device__ __forceinline__ int sum8_index(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }

// Block-wide sum of eight accumulators.  On return (warps 0 and 1) lane L holds the block total
// of value index (L & 7).  `red` is [32][8] floats of shared memory.
__device__ __forceinline__ float block_sum8(const float* acc, float (*red)[8]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    const float w = warp_sum8(acc, lane);
    if ((lane & 3) == 0) red[warp][sum8_index(lane)] = w;
    __syncthreads();
    float t = 0.f;
    if (warp < 2) {
        for (int ww = lane >> 3; ww < nw; ww += 4) t += red[ww][lane & 7];
        t += __shfl_xor_sync(0xffffffffu, t, 8);
        t += __shfl_xor_sync(0xffffffffu, t, 16);
    }
    return t;
}

// ------------------------------------------------------------------------------------------
// Open-loop batch: one CTA per EPL call (the drop-in EPL() and the teacher-forced parity test).
// ------------------------------------------------------------------------------------------
template <int DT, int VPC>
__global__ void __launch_bounds__(256) epl_batch_kernel(const uint8_t* __restrict__ iq, long long iq_len,
                                                        double fs, const sydr_epl_args* __restrict__ args,
                                                        const uint32_t* __restrict__ code_bits,
                                                        double* __restrict__ out) {
    constexpr int SPV = IqTraits<DT>::SPV, BPS = IqTraits<DT>::BPS;
    constexpr int C = SPV * VPC;
    __shared__ uint32_t cb[kCodeWords];
    __shared__ EpochConst ec_sh;
    __shared__ float red[32][8];
    const sydr_epl_args a = args[blockIdx.x];
    if (threadIdx.x < kCodeWords) cb[threadIdx.x] = code_bits[(a.prn - 1) * kCodeWords + threadIdx.x];
    if (threadIdx.x == 0)
        make_epoch_const(ec_sh, a.n, fs, a.carrier_freq, a.rem_carrier, a.rem_code, a.code_step, a.spacing);
    __syncthreads();
    const EpochConst ec = ec_sh;
    const long long a0 = a.start & ~(long long)(SPV - 1);     // 16-byte aligned window start
    const int lead = (int)(a.start - a0);
    const int nchunks = (lead + a.n + C - 1) / C;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int err = 0;
    for (int q = threadIdx.x; q < nchunks; q += blockDim.x) {
        const uint4* src = reinterpret_cast<const uint4*>(iq + (a0 + (long long)q * C) * BPS);
        correlate_chunk<DT, VPC>(src, q * C - lead, ec, cb, acc, err);
    }
    const float tot = block_sum8(acc, red);
    if (threadIdx.x < 6) out[(long long)blockIdx.x * 6 + threadIdx.x] = (double)tot;
    (void)iq_len;
}

// ------------------------------------------------------------------------------------------
// Closed loop.
// ------------------------------------------------------------------------------------------
// The NCO / loop-filter state of one channel (mirrors sydr_trk_state), split by owner: the code
// loop lives in warp 0, the carrier loop in warp 1; the two chains never read each other's
// variables (channel_l1ca_borre.py:363-429), so they close concurrently.
struct CodeState {
    long long cur;
    int n_req;
    double code_freq, code_step, rem_code, nco_code_err, nco_code;
};
struct CarrierState {
    double carrier_freq, rem_carrier, nco_carrier_err, nco_carrier;
};
struct LoopConst {        // per-channel constants hoisted out of the epoch loop
    double inv_fs, fs;
    double dll_c1, dll_c2;   // tau2/tau1, pdi/tau1            tracking.py:183-184
    double pll_c1, pll_c2;
};

struct TrkParams {
    const uint8_t* iq;
    long long iq_alloc;      // samples in the whole d_iq allocation
    double fs;
    sydr_trk_state* states;
    sydr_trk_epoch* out;
    int* nepochs;
    const uint32_t* code_bits;
    int max_epochs;
    int Q;                   // chunks per CTA per epoch (window = Q*C samples)
    int use_tma;
    int append;              // records are indexed by the cumulative epoch count
    long long iq_len;        // > 0: overrides the states' iq_len
    long long* prof;         // optional [n_channels][8] phase cycle counters of thread 0 (NULL = off)
};

struct EpochCtl {            // published by warps 0 / 1 for every epoch
    EpochConst ec;
    long long a;             // epoch start sample (rec-relative)
    int stop;
};

constexpr int kMaxCluster = 8;
constexpr int kTrkMaxThreads = 640;

struct TrkShared {           // static shared memory of the closed-loop kernel
    uint32_t cb[kCodeWords];
    EpochCtl ctl;
    float red[32][8];
    alignas(16) float gather[2][kMaxCluster][8];
    alignas(8) uint64_t bar_data[2];
    uint64_t bar_gather[2];
    sydr_trk_state cfgs;     // the channel's state as loaded (constants live here)
    CodeState sc;            // owned by warp 0
    CarrierState sk;         // owned by warp 1
    LoopConst K;
    int n_hist[2];           // samples of epoch e (index e & 1), for the carrier warp
    int rec_base;            // index of this call's first record in the channel's output row
    int status;
    long long pc[8];         // diagnostics
    long long tprev;
};

// TMA bulk copy of one CTA's window of the epoch starting at sample `a` (executed by one lane).
template <int DT, int VPC>
__device__ __forceinline__ void trk_prefetch(TrkShared& sh, uint8_t* dst, const uint8_t* rec_base, long long rec_alloc,
                                             long long a, uint32_t rank, int Q, int buf) {
    constexpr int SPV = IqTraits<DT>::SPV, BPS = IqTraits<DT>::BPS;
    constexpr int C = SPV * VPC;
    const long long a0 = a & ~(long long)(SPV - 1);
    long long w0 = a0 + (long long)rank * Q * C;
    long long w1 = w0 + (long long)Q * C;
    if (w1 > rec_alloc) w1 = rec_alloc & ~(long long)(SPV - 1);
    const long long bytes = (w1 > w0) ? (w1 - w0) * BPS : 0;
    if (bytes > 0) {
        mbar_arrive_expect_tx(&sh.bar_data[buf], (uint32_t)bytes);
        const uint8_t* src = rec_base + w0 * BPS;
        long long off = 0;
        while (off < bytes) {
            const uint32_t piece = (uint32_t)min((long long)32768, bytes - off);
            tma_bulk_g2s(dst + off, src + off, piece, &sh.bar_data[buf]);
            off += piece;
        }
    } else {
        mbar_arrive(&sh.bar_data[buf]);
    }
}

// Sum over the cluster's ranks of partial-sum component k (fixed order -> identical in every CTA).
__device__ __forceinline__ double rank_sum(const TrkShared& sh, int slot, uint32_t S, int k) {
    double s = 0.0;
    for (uint32_t r = 0; r < S; ++r) s += (double)sh.gather[slot][r][k];
    return s;
}

// Warp 0, once per epoch.  CLOSE: all-gather the cluster's partial sums, close the CODE loop
// (DLL_NNEML + Borre filter + code NCO, channel_l1ca_borre.py:383-388, 422-429) and store its
// share of the epoch record.  Then publish stop flag, epoch bounds and the three tap constants
// of the next epoch (lanes 0-2: one correlator each) and request its TMA window (lane 4).
template <int DT, int VPC, bool CLOSE>
__device__ __noinline__ void trk_code_warp(TrkShared& sh, const TrkParams& P, uint8_t* win0, uint8_t* win1,
                                           const uint8_t* rec_base, long long rec_alloc, float part, uint32_t S,
                                           uint32_t rank, int ch, int epoch, int lane, bool prof) {
    constexpr int SPV = IqTraits<DT>::SPV;
    constexpr int C = SPV * VPC;
    const unsigned full = 0xffffffffu;
    CodeState st = sh.sc;
    int status = sh.status;
    if (CLOSE) {
        const int e = epoch - 1;                       // the epoch just correlated
        double ck;                                     // lane k (< 8): total of component k
        if (S > 1) {
            const int slot = e & 1;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __shfl_sync(full, part, k);
            if ((uint32_t)lane < S) {
                const uint32_t dst = mapa_u32(smem_u32(&sh.gather[slot][rank][0]), (uint32_t)lane);
                const uint32_t rb = mapa_u32(smem_u32(&sh.bar_gather[slot]), (uint32_t)lane);
                st_async_v4(dst, rb, v[0], v[1], v[2], v[3]);
                st_async_v4(dst + 16, rb, v[4], v[5], v[6], v[7]);
            }
            // st.async data is visible once the phase completes (complete_tx): no cluster fence needed
            mbar_wait(&sh.bar_gather[slot], (e >> 1) & 1);
            ck = rank_sum(sh, slot, S, lane & 7);
        } else {
            ck = (double)part;
        }
        if (prof) { const long long now = clock64(); sh.pc[5] += now - sh.tprev; sh.tprev = now; }   // all-gather
        // |E| (even lanes) and |L| (odd lanes)                                    tracking.py:126
        const double mx = __shfl_sync(full, ck, (lane & 1) ? 4 : 0), my = __shfl_sync(full, ck, (lane & 1) ? 5 : 1);
        const double mag = sqrt(dadd(dmul(mx, mx), dmul(my, my)));
        const double me = __shfl_sync(full, mag, 0), ml = __shfl_sync(full, mag, 1);
        const double errflag = __shfl_sync(full, ck, 6);
        const double code_err = ddiv(dsub(me, ml), dadd(me, ml));
        double nco_code = dmul(sh.K.dll_c1, dsub(code_err, st.nco_code_err));       // BorreLoopFilter
        nco_code = dadd(nco_code, dmul(sh.K.dll_c2, code_err));
        const long long e_start = st.cur;
        const int e_n = st.n_req;
        const double n = i2d(e_n);
        st.code_freq = dsub(st.code_freq, nco_code);                                 // L422
        st.rem_code = dadd(st.rem_code, dsub(dmul(n, st.code_step), (double)kCodeChips));   // L424
        st.code_step = ddiv(st.code_freq, sh.K.fs);                                  // L425
        st.cur += e_n;                                                               // L428
        st.n_req = ceil_to_int(ddiv(dsub((double)kCodeChips, st.rem_code), st.code_step));  // L429
        st.nco_code_err = code_err;
        st.nco_code = nco_code;
        if (errflag != 0.0) status = SYDR_ERR_STATE;
        if (rank == 0) {
            double* rec = reinterpret_cast<double*>(P.out + (long long)ch * P.max_epochs + sh.rec_base + e);
            if (lane < 6) rec[lane] = ck;              // i_early .. q_late
            else if (lane == 6) rec[6] = nco_code;
            else if (lane == 9) rec[9] = st.code_freq;
            else if (lane == 10) rec[10] = code_err;
            else if (lane == 12) rec[12] = (double)e_start;
            else if (lane == 13) rec[13] = n;
            else if (lane == 14) rec[14] = st.rem_code;
        }
        if (lane == 0) sh.sc = st;
        if (prof) { const long long now = clock64(); sh.pc[6] += now - sh.tprev; sh.tprev = now; }   // code loop
    }
    // ---- publish epoch `epoch`
    if (st.n_req <= 0 || (long long)st.n_req + SPV > (long long)S * P.Q * C) status = SYDR_ERR_STATE;
    const bool stop = (status != 0) || (sh.rec_base + epoch >= P.max_epochs) || (st.cur + st.n_req > sh.cfgs.iq_len);
    if (!stop) {
        double t_start, t_step, t_inv;
        tap_const(st.rem_code, sh.cfgs.spacing[min(lane, 2)], st.code_step, st.n_req, t_start, t_step, t_inv);
        if (lane < 3) {
            sh.ctl.ec.start[lane] = t_start;
            sh.ctl.ec.step[lane] = t_step;
            sh.ctl.ec.inv_step[lane] = t_inv;
        } else if (lane == 3) {
            sh.ctl.ec.n = st.n_req;
            sh.ctl.a = st.cur;
            sh.n_hist[epoch & 1] = st.n_req;
            if (S > 1) mbar_arrive_expect_tx(&sh.bar_gather[epoch & 1], 32u * S);   // arm this epoch's gather
        } else if (lane == 4 && P.use_tma) {
            const int buf = epoch & 1;
            if (epoch == 0) trk_prefetch<DT, VPC>(sh, win0, rec_base, rec_alloc, st.cur, rank, P.Q, 0);
            // next epoch's window, fetched while this one is correlated
            trk_prefetch<DT, VPC>(sh, buf ? win0 : win1, rec_base, rec_alloc, st.cur + st.n_req, rank, P.Q, buf ^ 1);
        }
    }
    if (lane == 5) { sh.ctl.stop = stop; sh.status = status; }
    __syncwarp();
}

// Warp 1, once per epoch.  CLOSE: close the CARRIER loop (remaining carrier phase, PLL_costa +
// Borre filter + carrier NCO, channel_l1ca_borre.py:364-365, 391-396, 423) from the prompt sums
// and store its share of the record; then publish the carrier constants of the next epoch.
template <bool CLOSE>
__device__ __noinline__ void trk_carrier_warp(TrkShared& sh, const TrkParams& P, float part, uint32_t S, uint32_t rank,
                                              int ch, int epoch, int lane) {
    const unsigned full = 0xffffffffu;
    CarrierState st = sh.sk;
    if (CLOSE) {
        const int e = epoch - 1;
        double ip, qp;
        if (S > 1) {
            const int slot = e & 1;
            mbar_wait(&sh.bar_gather[slot], (e >> 1) & 1);
            ip = rank_sum(sh, slot, S, 2);
            qp = rank_sum(sh, slot, S, 3);
        } else {
            ip = (double)__shfl_sync(full, part, 2);
            qp = (double)__shfl_sync(full, part, 3);
        }
        const double n = i2d(sh.n_hist[e & 1]);
        // L364-365: rem' = (rem - ((fc*2)*pi*n)/fs) mod 2 pi   (Python float %: result in [0, 2 pi))
        const double twopi = 2.0 * kPi;
        double rc = dsub(st.rem_carrier, ddiv(dmul(dmul(dmul(st.carrier_freq, 2.0), kPi), n), sh.K.fs));
        {
            const double q = floor(rc * 0.15915494309189535);
            rc = fma(-q, twopi, rc);
            if (rc < 0.0) rc += twopi;
            if (rc >= twopi) rc -= twopi;
        }
        st.rem_carrier = rc;
        const double ph_err = ddiv(atan(ddiv(qp, ip)), kGpsPi * 2.0);               // PLL_costa
        double nco_car = dmul(sh.K.pll_c1, dsub(ph_err, st.nco_carrier_err));        // BorreLoopFilter
        nco_car = dadd(nco_car, dmul(sh.K.pll_c2, ph_err));
        st.carrier_freq = dadd(st.carrier_freq, nco_car);                            // L423
        st.nco_carrier_err = ph_err;
        st.nco_carrier = nco_car;
        if (rank == 0) {
            double* rec = reinterpret_cast<double*>(P.out + (long long)ch * P.max_epochs + sh.rec_base + e);
            if (lane == 7) rec[7] = nco_car;
            else if (lane == 8) rec[8] = st.carrier_freq;
            else if (lane == 11) rec[11] = ph_err;
            else if (lane == 15) rec[15] = rc;
        }
        if (lane == 0) sh.sk = st;
    }
    double ca, cbb;
    float w[4][2];
    carrier_const(st.carrier_freq, st.rem_carrier, sh.K.inv_fs, ca, cbb, w);
    if (lane == 0) {
        sh.ctl.ec.ca = ca;
        sh.ctl.ec.cb = cbb;
    } else if (lane <= 4) {
        sh.ctl.ec.w[lane - 1][0] = w[lane - 1][0];
        sh.ctl.ec.w[lane - 1][1] = w[lane - 1][1];
    }
    __syncwarp();
}

template <int DT, int VPC>
__global__ void __launch_bounds__(kTrkMaxThreads) trk_borre_kernel(const TrkParams P) {
    constexpr int SPV = IqTraits<DT>::SPV, BPS = IqTraits<DT>::BPS;
    constexpr int C = SPV * VPC;
    extern __shared__ __align__(128) uint8_t dyn_smem[];
    __shared__ __align__(16) TrkShared sh;

    const uint32_t S = cluster_nctarank();
    const uint32_t rank = cluster_ctarank();
    const int ch = blockIdx.x / S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Q = P.Q;
    const uint32_t win_bytes = (uint32_t)Q * C * BPS;
    uint8_t* win[2] = {dyn_smem, dyn_smem + win_bytes};

    sydr_trk_state* gst = P.states + ch;
    if (tid == 0) {
        sh.cfgs = *gst;
        if (P.iq_len > 0) sh.cfgs.iq_len = P.iq_len;
        sh.rec_base = P.append ? (int)sh.cfgs.epochs_done : 0;
        const sydr_trk_state& g = sh.cfgs;
        sh.sc.cur = g.cur; sh.sc.n_req = (int)g.n_req;
        sh.sc.code_freq = g.code_freq; sh.sc.code_step = g.code_step; sh.sc.rem_code = g.rem_code;
        sh.sc.nco_code_err = g.nco_code_err; sh.sc.nco_code = g.nco_code;
        sh.sk.carrier_freq = g.carrier_freq; sh.sk.rem_carrier = g.rem_carrier;
        sh.sk.nco_carrier_err = g.nco_carrier_err; sh.sk.nco_carrier = g.nco_carrier;
        sh.K.fs = P.fs; sh.K.inv_fs = 1.0 / P.fs;
        sh.K.dll_c1 = g.dll_tau2 / g.dll_tau1; sh.K.dll_c2 = g.dll_pdi / g.dll_tau1;
        sh.K.pll_c1 = g.pll_tau2 / g.pll_tau1; sh.K.pll_c2 = g.pll_pdi / g.pll_tau1;
        sh.status = 0;
        for (int k = 0; k < 8; ++k) sh.pc[k] = 0;
        mbar_init(&sh.bar_data[0], 1);
        mbar_init(&sh.bar_data[1], 1);
        mbar_init(&sh.bar_gather[0], 1);
        mbar_init(&sh.bar_gather[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid < kCodeWords) sh.cb[tid] = P.code_bits[(sh.cfgs.prn - 1) * kCodeWords + tid];
    if (S > 1) cluster_sync_all();                  // remote mbarriers are initialised
    const uint8_t* rec_base = P.iq + sh.cfgs.iq_base * BPS;
    const long long rec_alloc = P.iq_alloc - sh.cfgs.iq_base;   // samples readable from rec_base

    const bool prof = (P.prof != nullptr) && tid == 0;
#define SYDR_TICK(k)                                   \
    if (prof) {                                        \
        const long long now__ = clock64();             \
        sh.pc[k] += now__ - sh.tprev;                  \
        sh.tprev = now__;                              \
    }
    if (prof) sh.tprev = clock64();
    int epoch = 0;
    if (warp == 0) trk_code_warp<DT, VPC, false>(sh, P, win[0], win[1], rec_base, rec_alloc, 0.f, S, rank, ch, 0, lane, prof);
    else if (warp == 1) trk_carrier_warp<false>(sh, P, 0.f, S, rank, ch, 0, lane);
    while (true) {
        const int buf = epoch & 1;
        SYDR_TICK(0)                                   // epoch constants + TMA issue
        __syncthreads();
        if (sh.ctl.stop) break;
        SYDR_TICK(1)                                   // barrier

        // ---- correlate this CTA's window
        const long long a = sh.ctl.a;
        const long long a0 = a & ~(long long)(SPV - 1);
        const int lead = (int)(a - a0);
        const long long wstart = (long long)rank * Q * C;          // relative to a0
        const int n_epoch = sh.ctl.ec.n;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int err = 0;
        if (P.use_tma) mbar_wait(&sh.bar_data[buf], (epoch >> 1) & 1);
        SYDR_TICK(2)                                   // wait for the staged window
        for (int q = tid; q < Q; q += blockDim.x) {
            const int j0 = (int)(wstart + (long long)q * C) - lead;
            if (j0 >= n_epoch) break;
            const uint4* src;
            if (P.use_tma) {
                src = reinterpret_cast<const uint4*>(win[buf] + (size_t)q * C * BPS);
            } else {
                src = reinterpret_cast<const uint4*>(rec_base + (a0 + wstart + (long long)q * C) * BPS);
            }
            // ctl.ec is only rewritten after the block-wide barrier inside block_sum8
            correlate_chunk<DT, VPC>(src, j0, sh.ctl.ec, sh.cb, acc, err);
        }
        SYDR_TICK(3)                                   // correlate (thread 0's chunks)
        acc[6] = err ? 1.f : 0.f;                      // code-index overflow anywhere aborts the channel
        const float part = block_sum8(acc, sh.red);    // warps 0 and 1: lane L holds component L & 7
        SYDR_TICK(4)                                   // block reduction (waits for the slowest warp)
        ++epoch;
        if (warp == 0)
            trk_code_warp<DT, VPC, true>(sh, P, win[0], win[1], rec_base, rec_alloc, part, S, rank, ch, epoch, lane, prof);
        else if (warp == 1)
            trk_carrier_warp<true>(sh, P, part, S, rank, ch, epoch, lane);
        // (the __syncthreads at the top of the next iteration orders ctl / window reuse)
    }

    // the window of the epoch that will not run was already requested: drain it before exit
    if (P.use_tma && epoch > 0) mbar_wait(&sh.bar_data[epoch & 1], (epoch >> 1) & 1);

    if (prof && rank == 0) {
        for (int k = 0; k < 7; ++k) P.prof[ch * 8 + k] = sh.pc[k];
        P.prof[ch * 8 + 7] = epoch;
    }
    if (tid == 0 && rank == 0) {
        gst->cur = sh.sc.cur; gst->n_req = sh.sc.n_req; gst->epochs_done = sh.cfgs.epochs_done + epoch;
        gst->carrier_freq = sh.sk.carrier_freq; gst->code_freq = sh.sc.code_freq; gst->code_step = sh.sc.code_step;
        gst->rem_carrier = sh.sk.rem_carrier; gst->rem_code = sh.sc.rem_code;
        gst->nco_code = sh.sc.nco_code; gst->nco_code_err = sh.sc.nco_code_err;
        gst->nco_carrier = sh.sk.nco_carrier; gst->nco_carrier_err = sh.sk.nco_carrier_err;
        gst->status = sh.status;
        P.nepochs[ch] = sh.rec_base + epoch;
    }
    if (S > 1) cluster_sync_all();                  // nobody leaves while peers may still write here
}

}  // namespace sydr

using namespace sydr;

namespace {

template <int DT, int VPC>
int launch_epl(const void* d_iq, long long iq_len, double fs, const sydr_epl_args* d_args, int n_calls,
               const uint32_t* bits, double* d_out, cudaStream_t s) {
    epl_batch_kernel<DT, VPC><<<n_calls, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(d_iq), iq_len, fs, d_args,
                                                 bits, d_out);
    count_launch();
    SYDR_CUDA_CHECK(cudaGetLastError());
    return SYDR_OK;
}

template <int DT, int VPC>
int launch_trk(const TrkParams& P, int n_channels, int cluster, int threads, cudaStream_t s) {
    constexpr int C = IqTraits<DT>::SPV * VPC;
    const size_t smem = P.use_tma ? (size_t)2 * P.Q * C * IqTraits<DT>::BPS : 0;
    SYDR_REQUIRE(smem <= 200 * 1024, SYDR_ERR_UNSUPPORTED,
                 "tracking window needs %zu B of shared memory; raise cfg.cluster", smem);
    auto kern = trk_borre_kernel<DT, VPC>;
    SYDR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(n_channels * cluster));
    lc.blockDim = dim3((unsigned)threads);
    lc.dynamicSmemBytes = smem;
    lc.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    lc.attrs = at;
    lc.numAttrs = 1;
    SYDR_CUDA_CHECK(cudaLaunchKernelEx(&lc, kern, P));
    count_launch();
    return SYDR_OK;
}

}  // namespace

static long long* g_trk_prof = nullptr;    // set by sydr_trk_profile_buffer (diagnostics)

extern "C" {

// Diagnostics: give the tracking kernel a device buffer of n_channels*8 int64 to fill with
// per-phase cycle counts of the loop-closing thread (NULL switches it off).
int sydr_trk_profile_buffer(long long* d_buf) {
    g_trk_prof = d_buf;
    return SYDR_OK;
}

int sydr_epl_batch(const void* d_iq, int iq_dtype, long long iq_len, double fs, const sydr_epl_args* d_args,
                   int n_calls, double* d_out, void* stream) {
    SYDR_REQUIRE(d_iq && d_args && d_out, SYDR_ERR_ARG, "NULL pointer");
    SYDR_REQUIRE(((uintptr_t)d_iq & 15) == 0, SYDR_ERR_ARG, "d_iq must be 16-byte aligned");
    SYDR_REQUIRE(fs > 0, SYDR_ERR_ARG, "fs must be positive");
    if (n_calls <= 0) return SYDR_OK;
    CodeTables t;
    int rc = ensure_code_tables(&t);
    if (rc != SYDR_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    switch (iq_dtype) {
        case SYDR_IQ_I8: return launch_epl<SYDR_IQ_I8, 3>(d_iq, iq_len, fs, d_args, n_calls, t.padded_bits, d_out, s);
        case SYDR_IQ_I16: return launch_epl<SYDR_IQ_I16, 5>(d_iq, iq_len, fs, d_args, n_calls, t.padded_bits, d_out, s);
        case SYDR_IQ_F32: return launch_epl<SYDR_IQ_F32, 10>(d_iq, iq_len, fs, d_args, n_calls, t.padded_bits, d_out, s);
        default:
            set_error("sydr_epl_batch: iq_dtype %d not supported (convert complex128 with sydr_convert_to_f32)", iq_dtype);
            return SYDR_ERR_UNSUPPORTED;
    }
}

int sydr_trk_run(const void* d_iq, int iq_dtype, long long iq_alloc_samples, double fs, sydr_trk_state* d_states,
                 int n_channels, sydr_trk_epoch* d_out, int max_epochs, int* d_nepochs, const sydr_trk_config* cfg,
                 void* stream) {
    SYDR_REQUIRE(d_iq && d_states && d_out && d_nepochs, SYDR_ERR_ARG, "NULL pointer");
    SYDR_REQUIRE(((uintptr_t)d_iq & 15) == 0, SYDR_ERR_ARG, "d_iq must be 16-byte aligned");
    SYDR_REQUIRE(fs >= 2.0e6, SYDR_ERR_UNSUPPORTED, "fs %.0f Hz below the supported 2 MHz", fs);
    SYDR_REQUIRE(max_epochs > 0, SYDR_ERR_ARG, "max_epochs must be positive");
    if (n_channels <= 0) return SYDR_OK;
    CodeTables t;
    int rc = ensure_code_tables(&t);
    if (rc != SYDR_OK) return rc;

    // Longest epoch we stage for: nominal code period + 0.2 % (code Doppler is < 1e-5).
    const long long n_max = (long long)(fs * 1.002e-3) + 64;
    int cluster = cfg ? cfg->cluster : 0;
    int threads = cfg ? cfg->threads : 0;
    const int use_tma = cfg ? (cfg->use_tma != 0) : 1;
    if (cluster <= 0) {
        // latency mode while clusters still fit one wave of the 148 SMs, else throughput mode
        cluster = 1;
        for (int c = 8; c >= 2; c >>= 1)
            if ((long long)n_channels * c <= 148) { cluster = c; break; }
    }
    SYDR_REQUIRE(cluster == 1 || cluster == 2 || cluster == 4 || cluster == 8, SYDR_ERR_ARG,
                 "cluster must be 1, 2, 4 or 8 (got %d)", cluster);
    int spv, vpc, bps;
    switch (iq_dtype) {
        case SYDR_IQ_I8: spv = 8; vpc = 3; bps = 2; break;
        case SYDR_IQ_I16: spv = 4; vpc = (cluster >= 4) ? 3 : 5; bps = 4; break;   // latency / throughput chunks
        case SYDR_IQ_F32: spv = 2; vpc = 10; bps = 8; break;
        default:
            set_error("sydr_trk_run: iq_dtype %d not supported", iq_dtype);
            return SYDR_ERR_UNSUPPORTED;
    }
    const int C = spv * vpc;
    // shared-memory budget: two windows of Q*C samples
    while (use_tma && cluster < 8 && 2 * ((n_max + spv + (long long)C * cluster - 1) / ((long long)C * cluster)) * C * bps > 200 * 1024)
        cluster <<= 1;                                     // (the chunk size chosen above is kept)
    const int Q = (int)((n_max + spv + (long long)C * cluster - 1) / ((long long)C * cluster));
    if (threads <= 0) {
        const int rounds = (Q + kTrkMaxThreads - 1) / kTrkMaxThreads;
        threads = (((Q + rounds - 1) / rounds) + 31) / 32 * 32;
        if (threads < 64) threads = 64;                    // warps 0 and 1 close the two loops
    }
    SYDR_REQUIRE(threads % 32 == 0 && threads >= 64 && threads <= kTrkMaxThreads, SYDR_ERR_ARG, "threads must be a multiple of 32 in [64, %d] (got %d)", kTrkMaxThreads, threads);

    TrkParams P;
    P.iq = reinterpret_cast<const uint8_t*>(d_iq);
    P.iq_alloc = iq_alloc_samples;
    P.fs = fs;
    P.states = d_states;
    P.out = d_out;
    P.nepochs = d_nepochs;
    P.code_bits = t.padded_bits;
    P.max_epochs = max_epochs;
    P.Q = Q;
    P.use_tma = use_tma;
    P.append = cfg ? (cfg->append != 0) : 0;
    P.iq_len = cfg ? cfg->iq_len : 0;
    P.prof = g_trk_prof;
    cudaStream_t s = (cudaStream_t)stream;
    switch (iq_dtype) {
        case SYDR_IQ_I8: return launch_trk<SYDR_IQ_I8, 3>(P, n_channels, cluster, threads, s);
        case SYDR_IQ_I16:
            return (vpc == 3) ? launch_trk<SYDR_IQ_I16, 3>(P, n_channels, cluster, threads, s)
                              : launch_trk<SYDR_IQ_I16, 5>(P, n_channels, cluster, threads, s);
        default: return launch_trk<SYDR_IQ_F32, 10>(P, n_channels, cluster, threads, s);
    }
}

int sydr_trk_state_init(sydr_trk_state* h, int prn, double fs, double carrier_freq, long long start_sample,
                        double dll_bw, double dll_damp, double dll_gain, double dll_pdi, double pll_bw,
                        double pll_damp, double pll_gain, double pll_pdi, double sp_early, double sp_prompt,
                        double sp_late) {
    SYDR_REQUIRE(h != nullptr, SYDR_ERR_ARG, "state pointer is NULL");
    SYDR_REQUIRE(prn >= 1 && prn <= kMaxPrn, SYDR_ERR_ARG, "PRN %d out of range", prn);
    memset(h, 0, sizeof(*h));
    h->prn = prn;
    h->cur = start_sample;
    h->carrier_freq = carrier_freq;
    h->code_freq = kCodeFreq;                                  // channel_l1ca_borre.py:111
    h->code_step = kCodeFreq / fs;                             // L250
    h->n_req = (long long)ceil((kCodeChips - 0.0) / h->code_step);   // L251
    // LoopFiltersCoefficients, tracking.py:56-61
    auto coeff = [](double bw, double z, double g, double* t1, double* t2) {
        const double wn = bw * 8.0 * z / (4.0 * z * z + 1);
        *t1 = g / (wn * wn);
        *t2 = 2.0 * z / wn;
    };
    coeff(dll_bw, dll_damp, dll_gain, &h->dll_tau1, &h->dll_tau2);
    coeff(pll_bw, pll_damp, pll_gain, &h->pll_tau1, &h->pll_tau2);
    h->dll_pdi = dll_pdi;
    h->pll_pdi = pll_pdi;
    h->spacing[0] = sp_early;
    h->spacing[1] = sp_prompt;
    h->spacing[2] = sp_late;
    return SYDR_OK;
}

}  // extern "C"
