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

// Samples per thread chunk: a multiple of the 16-byte vector, chosen so that the per-thread
// stride in shared memory (80/48/144 bytes) is bank-conflict free for LDS.128.
template <int DT> struct ChunkTraits;
template <> struct ChunkTraits<SYDR_IQ_I8>  { static constexpr int VPC = 3; };   // 24 samples, 48 B
template <> struct ChunkTraits<SYDR_IQ_I16> { static constexpr int VPC = 5; };   // 20 samples, 80 B
template <> struct ChunkTraits<SYDR_IQ_F32> { static constexpr int VPC = 9; };   // 18 samples, 144 B

struct EpochConst {
    double K;              // fl(fl(fc*2.0)*pi)                       tracking.py:102
    double rem_carrier;    // remainingCarrier
    double inv_fs;
    double start[3];       // linspace start  = remCode + spacing     tracking.py:110
    double step[3];        // linspace step'  = (stop-start)/n        numpy linspace
    double inv_step[3];
    float wre, wim;        // per-sample carrier rotation exp(-j*K/fs)
    int n;                 // samples in the epoch
    int err;
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// fl(fl(j*step') + start): the value numpy's linspace produces for sample j.
__device__ __forceinline__ double code_phase(int j, double start, double step) {
    return dadd(dmul((double)j, step), start);
}
// ceil() of a double in (-2^31, 2^31) without a conversion instruction: adding 1.5*2^52 with
// round-up leaves ceil(x) in the low mantissa word.
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

// Sign mask for samples [jlo, jlo+cnt) of one correlator tap; bit i set = chip +1.
__device__ __forceinline__ uint32_t tap_mask(int jlo, int cnt, double start, double step,
                                             double inv_step, const uint32_t* cb, int& err) {
    const int jhi = jlo + cnt - 1;
    int k = ceil_to_int(code_phase(jlo, start, step));
    const int k1 = ceil_to_int(code_phase(jhi, start, step));
    uint32_t cur = chip_bit(cb, k, err);
    uint32_t m = cur ? 0xffffffffu : 0u;
    int guard = 0;
    while (k < k1 && guard++ < 40) {
        // first sample t in (jlo, jhi] whose phase exceeds k, i.e. whose ceil() is >= k+1
        int t = (int)floor(dmul(dsub((double)k, start), inv_step)) + 1;
        t = max(jlo + 1, min(t, jhi));
        while (t > jlo + 1 && code_phase(t - 1, start, step) > (double)k) --t;
        while (t < jhi && !(code_phase(t, start, step) > (double)k)) ++t;
        // several chips may start at the same sample only if step' >= 1 (fs < 1.023 MHz)
        const int kn = ceil_to_int(code_phase(t, start, step));
        const uint32_t nxt = chip_bit(cb, kn, err);
        if (nxt != cur) m ^= 0xffffffffu << (t - jlo);
        cur = nxt;
        k = kn;
    }
    return m;
}

__device__ __forceinline__ float flip(float v, uint32_t notmask, int i) {
    // multiply by the chip (+1 when mask bit i is set, -1 otherwise)
    return __uint_as_float(__float_as_uint(v) ^ ((notmask << (31 - i)) & 0x80000000u));
}

// Correlate one chunk of C = VPC*SPV samples starting at epoch-relative index j0 against the
// three taps.  `src` points at the chunk's first vector (shared or global memory, 16-byte
// aligned).  acc = {IE, QE, IP, QP, IL, QL}.
template <int DT>
__device__ __forceinline__ void correlate_chunk(const uint4* src, int j0, const EpochConst& ec,
                                                const uint32_t* cb, float acc[6], int& err) {
    constexpr int SPV = IqTraits<DT>::SPV;
    constexpr int VPC = ChunkTraits<DT>::VPC;
    constexpr int C = SPV * VPC;
    const int lo = max(j0, 0);
    const int hi = min(j0 + C, ec.n);
    if (hi <= lo) return;
    const bool interior = (lo == j0) && (hi == j0 + C);

    uint32_t nm[3];
#pragma unroll
    for (int s = 0; s < 3; ++s)
        nm[s] = ~(tap_mask(lo, hi - lo, ec.start[s], ec.step[s], ec.inv_step[s], cb, err) << (lo - j0));

    // Carrier seed: theta = -(K * t_j0) + rem  (FP64), reduced to a fraction of a turn.
    const double theta = dadd(-dmul(ec.K, dmul((double)j0, ec.inv_fs)), ec.rem_carrier);
    double turns = theta * 0.15915494309189535;          // 1/(2*pi)
    turns -= rint(turns);
    float pre, pim;
    sincospif((float)(2.0 * turns), &pim, &pre);

    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f;
#pragma unroll
    for (int v = 0; v < VPC; ++v) {
        const uint4 raw = src[v];
        float re[SPV], im[SPV];
        decode_vec<DT>(raw, re, im);
#pragma unroll
        for (int u = 0; u < SPV; ++u) {
            const int i = v * SPV + u;
            float xr = re[u], xi = im[u];
            if (!interior) {
                const bool ok = (j0 + i >= lo) && (j0 + i < hi);
                xr = ok ? xr : 0.f;
                xi = ok ? xi : 0.f;
            }
            // signal = replica * rfData                               tracking.py:105
            const float sr = xr * pre - xi * pim;
            const float si = xr * pim + xi * pre;
            a0 += flip(sr, nm[0], i); a1 += flip(si, nm[0], i);
            a2 += flip(sr, nm[1], i); a3 += flip(si, nm[1], i);
            a4 += flip(sr, nm[2], i); a5 += flip(si, nm[2], i);
            const float npre = pre * ec.wre - pim * ec.wim;
            pim = pre * ec.wim + pim * ec.wre;
            pre = npre;
        }
    }
    acc[0] += a0; acc[1] += a1; acc[2] += a2; acc[3] += a3; acc[4] += a4; acc[5] += a5;
}

// Per-epoch constants from the NCO state (one thread).
__device__ __forceinline__ void make_epoch_const(EpochConst& ec, int n, double fs, double fc,
                                                 double rem_carrier, double rem_code,
                                                 double code_step, const double* spacing) {
    ec.n = n;
    ec.err = 0;
    ec.K = dmul(dmul(fc, 2.0), kPi);
    ec.rem_carrier = rem_carrier;
    ec.inv_fs = 1.0 / fs;
    const double dn = (double)n;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const double start = dadd(rem_code, spacing[s]);                 // shift
        const double stop = dadd(dmul(code_step, dn), start);            // codeStep*n + shift
        const double step = ddiv(dsub(stop, start), dn);                 // linspace step
        ec.start[s] = start;
        ec.step[s] = step;
        ec.inv_step[s] = 1.0 / step;
    }
    double wt = fc * ec.inv_fs;                                          // turns per sample
    wt -= rint(wt);
    float s_, c_;
    sincospif((float)(-2.0 * wt), &s_, &c_);
    ec.wre = c_;
    ec.wim = s_;
}

// Block-wide sum of NV accumulators; result valid in every lane of warp 0.
// `red` is [32][8] floats of shared memory.
template <int NV>
__device__ __forceinline__ void block_sum(float* acc, float (*red)[8], float* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = warp_sum(acc[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) red[warp][k] = acc[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            float v = (lane < nw) ? red[lane][k] : 0.f;
            out[k] = warp_sum(v);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Open-loop batch: one CTA per EPL call (the drop-in EPL() and the teacher-forced parity test).
// ------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256) epl_batch_kernel(const uint8_t* __restrict__ iq, long long iq_len,
                                                        double fs, const sydr_epl_args* __restrict__ args,
                                                        const uint32_t* __restrict__ code_bits,
                                                        double* __restrict__ out) {
    constexpr int SPV = IqTraits<DT>::SPV, BPS = IqTraits<DT>::BPS;
    constexpr int C = SPV * ChunkTraits<DT>::VPC;
    __shared__ uint32_t cb[kCodeWords];
    __shared__ EpochConst ec;
    __shared__ float red[32][8];
    const sydr_epl_args a = args[blockIdx.x];
    if (threadIdx.x < kCodeWords) cb[threadIdx.x] = code_bits[(a.prn - 1) * kCodeWords + threadIdx.x];
    if (threadIdx.x == 0)
        make_epoch_const(ec, a.n, fs, a.carrier_freq, a.rem_carrier, a.rem_code, a.code_step, a.spacing);
    __syncthreads();
    const long long a0 = a.start & ~(long long)(SPV - 1);     // 16-byte aligned window start
    const int lead = (int)(a.start - a0);
    const int nchunks = (lead + a.n + C - 1) / C;
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int err = 0;
    for (int q = threadIdx.x; q < nchunks; q += blockDim.x) {
        const uint4* src = reinterpret_cast<const uint4*>(iq + (a0 + (long long)q * C) * BPS);
        correlate_chunk<DT>(src, q * C - lead, ec, cb, acc, err);
    }
    float tot[6];
    block_sum<6>(acc, red, tot);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) out[(long long)blockIdx.x * 6 + k] = (double)tot[k];
    }
    (void)iq_len;
}

// ------------------------------------------------------------------------------------------
// Closed loop.
// ------------------------------------------------------------------------------------------
struct LoopState {        // registers of the loop-closing thread; mirrors sydr_trk_state
    long long cur, n_req, epochs_done;
    double carrier_freq, code_freq, code_step, rem_carrier, rem_code;
    double nco_code, nco_code_err, nco_carrier, nco_carrier_err;
};

// channel_l1ca_borre.py:363-429 for one epoch, given the six correlator sums.
__device__ __forceinline__ void close_loops(LoopState& st, const sydr_trk_state& cfgs, double fs,
                                            const double c[6], sydr_trk_epoch& rec) {
    const double n = (double)st.n_req;
    // L364-365: remaining carrier phase
    double rc = dsub(st.rem_carrier, ddiv(dmul(dmul(dmul(st.carrier_freq, 2.0), kPi), n), fs));
    const double twopi = 2.0 * kPi;
    rc = fmod(rc, twopi);
    if (rc != 0.0 && rc < 0.0) rc = dadd(rc, twopi);            // Python float % semantics
    st.rem_carrier = rc;
    // L383-388: DLL_NNEML + Borre filter
    const double me = sqrt(dadd(dmul(c[0], c[0]), dmul(c[1], c[1])));
    const double ml = sqrt(dadd(dmul(c[4], c[4]), dmul(c[5], c[5])));
    const double code_err = ddiv(dsub(me, ml), dadd(me, ml));
    double nco_code = dmul(ddiv(cfgs.dll_tau2, cfgs.dll_tau1), dsub(code_err, st.nco_code_err));
    nco_code = dadd(nco_code, dmul(ddiv(cfgs.dll_pdi, cfgs.dll_tau1), code_err));
    st.nco_code = nco_code;
    st.nco_code_err = code_err;
    // L391-396: PLL_costa (GPS pi) + Borre filter
    const double ph_err = ddiv(atan(ddiv(c[3], c[2])), kGpsPi * 2.0);
    double nco_car = dmul(ddiv(cfgs.pll_tau2, cfgs.pll_tau1), dsub(ph_err, st.nco_carrier_err));
    nco_car = dadd(nco_car, dmul(ddiv(cfgs.pll_pdi, cfgs.pll_tau1), ph_err));
    st.nco_carrier = nco_car;
    st.nco_carrier_err = ph_err;
    // L422-425: NCO update
    st.code_freq = dsub(st.code_freq, nco_code);
    st.carrier_freq = dadd(st.carrier_freq, nco_car);
    st.rem_code = dadd(st.rem_code, dsub(dmul(n, st.code_step), (double)kCodeChips));
    st.code_step = ddiv(st.code_freq, fs);
    // L428-429
    rec.start = (double)st.cur;
    rec.n = n;
    st.cur += st.n_req;
    st.n_req = (long long)ceil(ddiv(dsub((double)kCodeChips, st.rem_code), st.code_step));
    st.epochs_done += 1;
#pragma unroll
    for (int k = 0; k < 6; ++k) rec.corr[k] = c[k];
    rec.dll = nco_code;
    rec.pll = nco_car;
    rec.carrier_freq = st.carrier_freq;
    rec.code_freq = st.code_freq;
    rec.code_err = code_err;
    rec.carrier_err = ph_err;
    rec.rem_code = st.rem_code;
    rec.rem_carrier = st.rem_carrier;
}

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
    long long* prof;         // optional [n_channels][8] phase cycle counters of thread 0 (NULL = off)
};

struct EpochCtl {            // published by the loop thread each epoch
    EpochConst ec;
    long long a;             // epoch start sample (rec-relative)
    long long a_next;        // next epoch start (= a + n)
    int stop;
};

constexpr int kMaxCluster = 8;
constexpr int kTrkMaxThreads = 640;

template <int DT>
__global__ void __launch_bounds__(kTrkMaxThreads) trk_borre_kernel(const TrkParams P) {
    constexpr int SPV = IqTraits<DT>::SPV, BPS = IqTraits<DT>::BPS;
    constexpr int C = SPV * ChunkTraits<DT>::VPC;
    extern __shared__ __align__(128) uint8_t dyn_smem[];
    __shared__ uint32_t cb[kCodeWords];
    __shared__ EpochCtl ctl;
    __shared__ float red[32][8];
    __shared__ __align__(16) float gather[2][kMaxCluster][8];
    __shared__ __align__(8) uint64_t bar_data[2];
    __shared__ __align__(8) uint64_t bar_gather[2];

    const uint32_t S = cluster_nctarank();
    const uint32_t rank = cluster_ctarank();
    const int ch = blockIdx.x / S;
    const int tid = threadIdx.x;
    const int Q = P.Q;
    const uint32_t win_bytes = (uint32_t)Q * C * BPS;
    uint8_t* win[2] = {dyn_smem, dyn_smem + win_bytes};

    sydr_trk_state* gst = P.states + ch;
    __shared__ sydr_trk_state cfgs;                 // constant part (taus, spacing, base, len)
    if (tid == 0) {
        cfgs = *gst;
        mbar_init(&bar_data[0], 1);
        mbar_init(&bar_data[1], 1);
        mbar_init(&bar_gather[0], 1);
        mbar_init(&bar_gather[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid < kCodeWords) cb[tid] = P.code_bits[(cfgs.prn - 1) * kCodeWords + tid];
    if (S > 1) cluster_sync_all();                  // remote mbarriers are initialised
    const uint8_t* rec_base = P.iq + cfgs.iq_base * BPS;
    const long long rec_alloc = P.iq_alloc - cfgs.iq_base;   // samples readable from rec_base

    LoopState st;
    if (tid == 0) {
        st.cur = cfgs.cur; st.n_req = cfgs.n_req; st.epochs_done = cfgs.epochs_done;
        st.carrier_freq = cfgs.carrier_freq; st.code_freq = cfgs.code_freq; st.code_step = cfgs.code_step;
        st.rem_carrier = cfgs.rem_carrier; st.rem_code = cfgs.rem_code;
        st.nco_code = cfgs.nco_code; st.nco_code_err = cfgs.nco_code_err;
        st.nco_carrier = cfgs.nco_carrier; st.nco_carrier_err = cfgs.nco_carrier_err;
    }

    // Issue the TMA bulk copy of this CTA's window of the epoch starting at sample `a`.
    auto prefetch = [&](long long a, int buf) {
        const long long a0 = a & ~(long long)(SPV - 1);
        long long w0 = a0 + (long long)rank * Q * C;
        long long w1 = w0 + (long long)Q * C;
        if (w1 > rec_alloc) w1 = rec_alloc & ~(long long)(SPV - 1);
        const long long bytes = (w1 > w0) ? (w1 - w0) * BPS : 0;
        if (bytes > 0) {
            mbar_arrive_expect_tx(&bar_data[buf], (uint32_t)bytes);
            const uint8_t* src = rec_base + w0 * BPS;
            long long off = 0;
            while (off < bytes) {
                const uint32_t piece = (uint32_t)min((long long)32768, bytes - off);
                tma_bulk_g2s(win[buf] + off, src + off, piece, &bar_data[buf]);
                off += piece;
            }
        } else {
            mbar_arrive(&bar_data[buf]);
        }
    };

    int epoch = 0;
    int status = 0;
    long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = 0;
    const bool prof = (P.prof != nullptr) && tid == 0;
#define SYDR_TICK(k)                                   \
    if (prof) {                                        \
        const long long now__ = clock64();             \
        pc[k] += now__ - tprev;                        \
        tprev = now__;                                 \
    }
    if (prof) tprev = clock64();
    while (true) {
        const int buf = epoch & 1;
        if (tid == 0) {
            if (st.n_req <= 0 || st.n_req + SPV > (long long)S * Q * C) status = SYDR_ERR_STATE;
            const bool stop = (status != 0) || (epoch >= P.max_epochs) || (st.cur + st.n_req > cfgs.iq_len);
            ctl.stop = stop;
            if (!stop) {
                make_epoch_const(ctl.ec, (int)st.n_req, P.fs, st.carrier_freq, st.rem_carrier, st.rem_code,
                                 st.code_step, cfgs.spacing);
                ctl.a = st.cur;
                ctl.a_next = st.cur + st.n_req;
                if (P.use_tma) {
                    if (epoch == 0) prefetch(st.cur, 0);
                    prefetch(ctl.a_next, buf ^ 1);       // next epoch's window, while we compute
                }
            }
        }
        SYDR_TICK(0)                                   // epoch constants + TMA issue
        __syncthreads();
        if (ctl.stop) break;
        SYDR_TICK(1)                                   // barrier

        // ---- correlate this CTA's window
        const long long a = ctl.a;
        const long long a0 = a & ~(long long)(SPV - 1);
        const int lead = (int)(a - a0);
        const long long wstart = (long long)rank * Q * C;          // relative to a0
        float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int err = 0;
        if (P.use_tma) mbar_wait(&bar_data[buf], (epoch >> 1) & 1);
        SYDR_TICK(2)                                   // wait for the staged window
        for (int q = tid; q < Q; q += blockDim.x) {
            const int j0 = (int)(wstart + (long long)q * C) - lead;
            if (j0 >= ctl.ec.n) break;
            const uint4* src;
            if (P.use_tma) {
                src = reinterpret_cast<const uint4*>(win[buf] + (size_t)q * C * BPS);
            } else {
                src = reinterpret_cast<const uint4*>(rec_base + (a0 + wstart + (long long)q * C) * BPS);
            }
            correlate_chunk<DT>(src, j0, ctl.ec, cb, acc, err);
        }
        SYDR_TICK(3)                                   // correlate (thread 0's chunks)
        acc[6] = err ? 1.f : 0.f;                     // code-index overflow anywhere aborts the channel
        float part[7];
        block_sum<7>(acc, red, part);
        SYDR_TICK(4)                                   // block reduction (waits for the slowest warp)

        // ---- gather the cluster's partial sums and close the loops (warp 0)
        if (tid < 32) {
            double c[6];
            if (S > 1) {
                const int slot = epoch & 1;
                if (tid == 0) mbar_arrive_expect_tx(&bar_gather[slot], 32u * S);
                __syncwarp();
                if ((uint32_t)tid < S) {
                    const uint32_t dst = mapa_u32(smem_u32(&gather[slot][rank][0]), (uint32_t)tid);
                    const uint32_t rb = mapa_u32(smem_u32(&bar_gather[slot]), (uint32_t)tid);
                    st_async_v4(dst, rb, part[0], part[1], part[2], part[3]);
                    st_async_v4(dst + 16, rb, part[4], part[5], part[6], 0.f);
                }
                mbar_wait_cluster(&bar_gather[slot], (epoch >> 1) & 1);
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    double s = 0.0;
                    for (uint32_t r = 0; r < S; ++r) s += (double)gather[slot][r][k];
                    c[k] = s;
                }
                float e = 0.f;
                for (uint32_t r = 0; r < S; ++r) e += gather[slot][r][6];
                part[6] = e;
            } else {
#pragma unroll
                for (int k = 0; k < 6; ++k) c[k] = (double)part[k];
            }
            SYDR_TICK(5)                               // cluster all-gather
            if (tid == 0) {
                sydr_trk_epoch rec;
                close_loops(st, cfgs, P.fs, c, rec);
                if (part[6] != 0.f) status = SYDR_ERR_STATE;
                if (rank == 0) P.out[(long long)ch * P.max_epochs + epoch] = rec;
            }
            SYDR_TICK(6)                               // loop closure + record store
        }
        ++epoch;
        // (the __syncthreads at the top of the next iteration orders ctl/window reuse)
    }

    // the window of the epoch that will not run was already requested: drain it before exit
    if (P.use_tma && epoch > 0) mbar_wait(&bar_data[epoch & 1], (epoch >> 1) & 1);

    if (prof && rank == 0) {
        for (int k = 0; k < 7; ++k) P.prof[ch * 8 + k] = pc[k];
        P.prof[ch * 8 + 7] = epoch;
    }
    if (tid == 0 && rank == 0) {
        gst->cur = st.cur; gst->n_req = st.n_req; gst->epochs_done = st.epochs_done;
        gst->carrier_freq = st.carrier_freq; gst->code_freq = st.code_freq; gst->code_step = st.code_step;
        gst->rem_carrier = st.rem_carrier; gst->rem_code = st.rem_code;
        gst->nco_code = st.nco_code; gst->nco_code_err = st.nco_code_err;
        gst->nco_carrier = st.nco_carrier; gst->nco_carrier_err = st.nco_carrier_err;
        gst->status = status;
        P.nepochs[ch] = epoch;
    }
    if (S > 1) cluster_sync_all();                  // nobody leaves while peers may still write here
}

}  // namespace sydr

using namespace sydr;

namespace {

template <int DT>
int launch_epl(const void* d_iq, long long iq_len, double fs, const sydr_epl_args* d_args, int n_calls,
               const uint32_t* bits, double* d_out, cudaStream_t s) {
    epl_batch_kernel<DT><<<n_calls, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(d_iq), iq_len, fs, d_args,
                                                 bits, d_out);
    count_launch();
    SYDR_CUDA_CHECK(cudaGetLastError());
    return SYDR_OK;
}

template <int DT>
int launch_trk(const TrkParams& P, int n_channels, int cluster, int threads, cudaStream_t s) {
    constexpr int C = IqTraits<DT>::SPV * ChunkTraits<DT>::VPC;
    const size_t smem = P.use_tma ? (size_t)2 * P.Q * C * IqTraits<DT>::BPS : 0;
    SYDR_REQUIRE(smem <= 200 * 1024, SYDR_ERR_UNSUPPORTED,
                 "tracking window needs %zu B of shared memory; raise cfg.cluster", smem);
    auto kern = trk_borre_kernel<DT>;
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
        case SYDR_IQ_I8: return launch_epl<SYDR_IQ_I8>(d_iq, iq_len, fs, d_args, n_calls, t.padded_bits, d_out, s);
        case SYDR_IQ_I16: return launch_epl<SYDR_IQ_I16>(d_iq, iq_len, fs, d_args, n_calls, t.padded_bits, d_out, s);
        case SYDR_IQ_F32: return launch_epl<SYDR_IQ_F32>(d_iq, iq_len, fs, d_args, n_calls, t.padded_bits, d_out, s);
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

    int spv, vpc, bps;
    switch (iq_dtype) {
        case SYDR_IQ_I8: spv = 8; vpc = 3; bps = 2; break;
        case SYDR_IQ_I16: spv = 4; vpc = 5; bps = 4; break;
        case SYDR_IQ_F32: spv = 2; vpc = 9; bps = 8; break;
        default:
            set_error("sydr_trk_run: iq_dtype %d not supported", iq_dtype);
            return SYDR_ERR_UNSUPPORTED;
    }
    const int C = spv * vpc;
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
    // shared-memory budget: two windows of Q*C samples
    while (use_tma && cluster < 8 && 2 * ((n_max + spv + (long long)C * cluster - 1) / ((long long)C * cluster)) * C * bps > 200 * 1024)
        cluster <<= 1;
    const int Q = (int)((n_max + spv + (long long)C * cluster - 1) / ((long long)C * cluster));
    if (threads <= 0) {
        const int rounds = (Q + kTrkMaxThreads - 1) / kTrkMaxThreads;
        threads = (((Q + rounds - 1) / rounds) + 31) / 32 * 32;
        if (threads < 64) threads = 64;
    }
    SYDR_REQUIRE(threads % 32 == 0 && threads >= 32 && threads <= kTrkMaxThreads, SYDR_ERR_ARG, "threads must be a multiple of 32 in [32, %d] (got %d)", kTrkMaxThreads, threads);

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
    P.prof = g_trk_prof;
    cudaStream_t s = (cudaStream_t)stream;
    switch (iq_dtype) {
        case SYDR_IQ_I8: return launch_trk<SYDR_IQ_I8>(P, n_channels, cluster, threads, s);
        case SYDR_IQ_I16: return launch_trk<SYDR_IQ_I16>(P, n_channels, cluster, threads, s);
        default: return launch_trk<SYDR_IQ_F32>(P, n_channels, cluster, threads, s);
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
