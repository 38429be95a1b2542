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
#include <cstdlib>
#include <cstdio>
#include "trk_common.cuh"

namespace sydr {

// ------------------------------------------------------------------------------------------
// Open-loop batch: one CTA per EPL call (the drop-in EPL() and the teacher-forced parity test).
// ------------------------------------------------------------------------------------------
// 8-byte loads per segment of the half-chip path for a kernel instantiation (0 = path not built):
// int16 IQ with 12-sample chunks (25 MS/s) -> 7, 20-sample chunks (50 MS/s) -> 13, 4-sample
// chunks (10 MS/s) -> 3.
template <int DT, int VPC>
struct SegTraits { static constexpr int NV = 0; };
template <> struct SegTraits<SYDR_IQ_I16, 1> { static constexpr int NV = 3; };
template <> struct SegTraits<SYDR_IQ_I16, 3> { static constexpr int NV = 7; };
template <> struct SegTraits<SYDR_IQ_I16, 5> { static constexpr int NV = 13; };

template <int DT, int VPC>
__global__ void __launch_bounds__(256) epl_batch_kernel(const uint8_t* __restrict__ iq, long long iq_len,
                                                        double fs, const sydr_epl_args* __restrict__ args,
                                                        const uint32_t* __restrict__ code_bits,
                                                        double* __restrict__ out, int allow_seg) {
    constexpr int SPV = IqTraits<DT>::SPV, BPS = IqTraits<DT>::BPS;
    constexpr int C = SPV * VPC;
    constexpr int NV = SegTraits<DT, VPC>::NV;
    __shared__ uint32_t cb[kCodeWords];
    __shared__ EpochConst ec_sh;
    __shared__ float red[32][8];
    __shared__ uint32_t segtab[NV > 0 ? kSegTab : 4];
    __shared__ int seg_q[3];
    const sydr_epl_args a = args[blockIdx.x];
    // a call outside the recording or with no code for its PRN launches no loads: its six sums read NaN
    if ((unsigned)(a.prn - 1) >= (unsigned)kMaxPrn || a.n <= 0 || a.start < 0 || a.start + a.n > iq_len) {
        if (threadIdx.x < 6) out[(long long)blockIdx.x * 6 + threadIdx.x] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    if (threadIdx.x < kCodeWords) cb[threadIdx.x] = code_bits[(a.prn - 1) * kCodeWords + threadIdx.x];
    const long long a0 = a.start & ~(long long)(SPV - 1);     // 16-byte aligned window start
    const int lead = (int)(a.start - a0);
    if (threadIdx.x == 0) {
        make_epoch_const(ec_sh, a.n, fs, a.carrier_freq, a.rem_carrier, a.rem_code, a.code_step, a.spacing, C);
        ec_sh.seg = 0;
        if (NV > 0 && allow_seg) {
            const bool ok = seg_tap_offsets(a.spacing, seg_q);
            const double smin = fmin(ec_sh.start[0], fmin(ec_sh.start[1], ec_sh.start[2]));
            const double smax = fmax(ec_sh.start[0], fmax(ec_sh.start[1], ec_sh.start[2]));
            ec_sh.seg = ok && a.n > 0 && seg_epoch_ok<(NV > 0 ? NV : 1)>(smin, smax + a.code_step * (double)a.n * 1.000001, ec_sh.inv_step[1]);
            ec_sh.hb = ceil_to_int(2.0 * ec_sh.start[1]) - 1;
        }
    }
    __syncthreads();
    if (threadIdx.x < kMaxChunk)
        carrier_table_entry(ec_sh.ca, threadIdx.x, ec_sh.wtab[threadIdx.x][0], ec_sh.wtab[threadIdx.x][1]);
    if (NV > 0 && ec_sh.seg) build_seg_table(segtab, cb, seg_q);
    __syncthreads();
    const EpochConst& ec = ec_sh;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int err = 0;
    if (NV > 0 && ec.seg) {
        const uint8_t* base = iq + a0 * BPS;
        for (int h = ec.hb + (int)threadIdx.x;; h += blockDim.x)
            if (!correlate_segment<(NV > 0 ? NV : 1)>(base, -lead, 0x7fffffff, h, ec, segtab, seg_q, cb, acc, err)) break;
    } else {
        const int nchunks = (lead + a.n + C - 1) / C;
        for (int q = threadIdx.x; q < nchunks; q += blockDim.x) {
            const uint4* src = reinterpret_cast<const uint4*>(iq + (a0 + (long long)q * C) * BPS);
            correlate_chunk<DT, VPC>(src, q * C - lead, ec, cb, acc, err);
        }
    }
    const float tot = block_sum8(acc, red);
    if (threadIdx.x < 6) out[(long long)blockIdx.x * 6 + threadIdx.x] = (double)tot;
}

// One channel = one CTA or one cluster of S CTAs; W warps per CTA.  Per epoch:
//   (A) block barrier: the epoch constants published by warps 0 / 1 are visible;
//   (B) every warp correlates its chunks of the staged window, reduces its six sums with
//       shuffles and sends them to *every* CTA of the cluster (st.async + complete_tx on the
//       destination's mbarrier; plain stores + mbarrier.arrive when S = 1);
//   (C) warps 0 and 1 wait on that mbarrier, total the partial sums, close the code / carrier
//       loop in FP64, store the epoch record and publish the constants of the next epoch.
// The TMA window of the next epoch is requested right after (A) by the last warp, off the
// loop-closing path.
// LEAN = throughput instantiation: segment path only, <= 256 threads and <= 80 registers so that
// three channels share an SM and one channel's loop closure hides behind the others' correlation.
// An epoch the segment path cannot serve stops the channel with status kNeedGeneral; the host
// then repeats the launch with the general instantiation (sydr_trk_run).
// PROF = phase cycle counters compiled in (diagnostics instantiation; the counters sit on the serial
// chain, so the production instantiation does not carry them).
// KAP = the carrier warp closes the Kaplan loops (FLL-assisted PLL, lock indicators, C/N0, lock-state
// machine) instead of the Borre PLL; the code loop and everything else are shared.
// DENSE = same code under a tighter register budget (two CTAs of 288 threads per SM guaranteed, three of 192
// in practice): 4 % slower alone, but launches of several steps in flight pack 3 per SM instead of 2.
// PACK = throughput shape of the staged kernel (cfg.dense = 2): one CTA per channel with ONE staged window instead of
// two, so that two CTAs (two channels) share an SM and one channel's serial section -- gather, totals, FP64 loop closure,
// next epoch's constants: 40 % of an epoch for a CTA alone on its SM -- runs under the other's correlation.  The window
// of epoch k+1 is requested by the last warp as soon as every warp has released the window of epoch k (mbarrier), and
// lands while warps 0 / 1 close the loops.
template <int DT, int VPC, bool TMA, bool LEAN, bool PROF = false, bool KAP = false, bool DENSE = false, bool PACK = false>
__global__ void __launch_bounds__(LEAN ? kLeanThreads : (KAP ? kKaplanMaxThreads : (PACK ? kPackThreads : (DENSE ? kDenseMaxThreads : kTrkMaxThreads))),
                                  LEAN ? 3 : ((DENSE || PACK) ? 2 : 1))
trk_borre_kernel(const TrkParams P) {
    static_assert(!PACK || (TMA && !LEAN && !PROF && !KAP), "PACK is an instantiation of the staged Borre kernel");
    constexpr int SPV = IqTraits<DT>::SPV, BPS = IqTraits<DT>::BPS;
    constexpr int C = SPV * VPC;
    extern __shared__ __align__(128) uint8_t dyn_smem[];
    constexpr int NV = SegTraits<DT, VPC>::NV;
    // Integer IQ: the correlator sums of a warp are bounded, so they are exchanged in fixed point
    // (one REDUX per component instead of a shuffle tree, 32 B instead of 64 B per warp and CTA,
    // order-independent integer totals).  complex64 input has no bound: FP32 tree + FP64 totals.
    constexpr bool FIX = (DT != SYDR_IQ_F32);
    __shared__ __align__(16) TrkSharedT<(LEAN ? kLeanThreads / 32 : (PACK ? kPackThreads / 32 : kMaxCluster * kTrkMaxWarps)), (NV > 0 ? kSegTab : 1)> sh;
    const unsigned full = 0xffffffffu;

    const uint32_t S = cluster_nctarank();
    const uint32_t rank = cluster_ctarank();
    const int ch = blockIdx.x / S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, W = blockDim.x >> 5;
    const int Q = P.Q;
    const uint32_t win_bytes = TMA ? (uint32_t)(Q * C + kWinTail) * BPS : 0u;
    const int n_ent = (int)S * W;

    sydr_trk_state* gst = P.states + ch;
    if (tid == 0) {
        sh.cfgs = *gst;
        if (P.iq_len > 0) sh.cfgs.iq_len = P.iq_len;
        if (P.has_iq_base) sh.cfgs.iq_base = P.iq_base;
        sh.rec_base = P.append ? (int)sh.cfgs.epochs_done : (P.resume ? P.nepochs[ch] : 0);
        if (P.resume && sh.cfgs.status == kNeedGeneral) sh.cfgs.status = 0;
        if ((unsigned)(sh.cfgs.prn - 1) >= (unsigned)kMaxPrn && sh.cfgs.status == 0) sh.cfgs.status = SYDR_ERR_STATE;   // no code for this PRN
        const sydr_trk_state& g = sh.cfgs;
        sh.sc.cur = g.cur; sh.sc.n_req = (int)g.n_req;
        sh.sc.code_freq = g.code_freq; sh.sc.code_step = g.code_step; sh.sc.rem_code = g.rem_code;
        sh.sc.nco_code_err = g.nco_code_err; sh.sc.nco_code = g.nco_code;
        sh.sc.inv_step = drcp(g.code_step); sh.sc.inv_n = drcp((double)g.n_req);
        sh.sk.carrier_freq = g.carrier_freq; sh.sk.rem_carrier = g.rem_carrier;
        sh.sk.nco_carrier_err = g.nco_carrier_err; sh.sk.nco_carrier = g.nco_carrier;
        sh.K.fs = P.fs; sh.K.inv_fs = 1.0 / P.fs;
        sh.K.dll_c1 = g.dll_tau2 / g.dll_tau1; sh.K.dll_c2 = g.dll_pdi / g.dll_tau1;
        sh.K.pll_c1 = g.pll_tau2 / g.pll_tau1; sh.K.pll_c2 = g.pll_pdi / g.pll_tau1;
        sh.status = sh.cfgs.status;
        if constexpr (KAP) sh.kcfg = P.kstates[ch];
        sh.seg_ok = (NV > 0 && P.seg) ? (seg_tap_offsets(sh.cfgs.spacing, sh.seg_q) ? 1 : 0) : 0;
        for (int k = 0; k < 16; ++k) sh.pc[k] = 0;
        mbar_init(&sh.bar_data[0], 1);
        mbar_init(&sh.bar_data[1], 1);
        // cluster: one arrival (expect_tx) + S*W*64 bytes of st.async; single CTA: one arrival per warp
        mbar_init(&sh.bar_gather[0], S > 1 ? 1u : (uint32_t)W);
        mbar_init(&sh.bar_gather[1], S > 1 ? 1u : (uint32_t)W);
        mbar_init(&sh.bar_free, (uint32_t)W);       // PACK: one arrival per warp that has finished reading the window
        fence_mbar_init();
    }
    __syncthreads();
    if (tid < kCodeWords) sh.cb[tid] = P.code_bits[min(max(sh.cfgs.prn, 1), kMaxPrn) * kCodeWords - kCodeWords + tid];
    if (NV > 0 && sh.seg_ok) {
        __syncthreads();
        build_seg_table(sh.segtab, sh.cb, sh.seg_q);
    }
    if (S > 1) cluster_sync_all();                  // remote mbarriers are initialised
    const uint8_t* rec_base = P.iq + sh.cfgs.iq_base * BPS;
    const long long rec_alloc = P.iq_alloc - sh.cfgs.iq_base;   // samples readable from rec_base
    sydr_trk_epoch* out_row = (rank == 0) ? P.out + (long long)ch * P.max_epochs + sh.rec_base : nullptr;

    const bool prof = PROF && (P.prof != nullptr) && tid == 0;
#define SYDR_TICK(k)                                   \
    if (prof) {                                        \
        const long long now__ = clock64();             \
        sh.pc[k] += now__ - sh.tprev;                  \
        sh.tprev = now__;                              \
    }
    if (prof) sh.tprev = clock64();
    const bool prof1 = PROF && (P.prof != nullptr) && tid == 32;   // carrier warp: slots 10..13
#define SYDR_TICK1(k)                                  \
    if (prof1) {                                       \
        const long long now__ = clock64();             \
        sh.pc[k] += now__ - sh.tprev1;                 \
        sh.tprev1 = now__;                             \
    }
    if (prof1) sh.tprev1 = clock64();

    // loop state lives in registers of its owning warp (warp 0: code, warp 1: carrier)
    CodeState sc = sh.sc;
    CarrierState sk = sh.sk;
    struct NoKaplan {};
    typename std::conditional<KAP, KaplanRegs, NoKaplan>::type kr = {};
    if constexpr (KAP) {
        const sydr_kaplan_state& g = sh.kcfg;
        kr.ip_prev = g.ip_prev; kr.qp_prev = g.qp_prev; kr.fll_lock = g.fll_lock; kr.pll_lock = g.pll_lock;
        kr.cn0 = g.cn0; kr.pdpn = g.pdpn; kr.vel_memory = g.vel_memory; kr.fll_bw = g.fll_bw; kr.pll_bw = g.pll_bw;
        kr.accum_counter = g.accum_counter; kr.lock_state = g.lock_state; kr.flags = g.flags;
        kr.code_counter = g.code_counter;
        kr.atan_prev = atan(__ddiv_rn(g.qp_prev, g.ip_prev));
    }
    sydr_kaplan_epoch* kout_row = nullptr;
    if constexpr (KAP) kout_row = (rank == 0) ? P.kout + (long long)ch * P.max_epochs + sh.rec_base : nullptr;
    int status = sh.cfgs.status;                   // != 0: aborted earlier (< 0) or idle slot (> 0): no epochs
    int epoch = 0;
    // loop-invariant limits of the stop test
    const long long cap64 = (long long)S * Q * C - SPV;
    const unsigned n_cap = (unsigned)(cap64 < 0 ? 0 : (cap64 > 0x7fffffffLL ? 0x7fffffffLL : cap64));
    const int epoch_cap = P.max_epochs - sh.rec_base;
    const long long iq_len_reg = sh.cfgs.iq_len < rec_alloc ? sh.cfgs.iq_len : rec_alloc;   // never past the allocation
    while (true) {
        // ---- (C, second half) publish the constants of epoch `epoch`
        if (warp == 0) {
            if ((unsigned)(sc.n_req - 1) >= n_cap) status = SYDR_ERR_STATE;       // n_req <= 0 or beyond the staged window
            bool stop = (status != 0) || (epoch >= epoch_cap) || (sc.cur + sc.n_req > iq_len_reg);
            double t_start = 0.0, t_step = 0.0, t_stop = 0.0;
            int fast = 0, seg = 0, hb = 0, rounds = 0;
            if (!stop) {
                // tap constants (numpy linspace arithmetic, tracking.py:110-112), lane s < 3 = correlator s
                const double dn = i2d(sc.n_req);
                sc.inv_n = newton_rcp(dn, newton_rcp(dn, sc.inv_n));                  // n moves by +-1 at most
                t_start = dadd(sc.rem_code, sh.cfgs.spacing[min(lane, 2)]);
                t_stop = dadd(dmul(sc.code_step, dn), t_start);
                t_step = ddiv_by(dsub(t_stop, t_start), dn, sc.inv_n);
                fast = __all_sync(full, sc.inv_step >= (double)(C + 1)) && !(NV > 0 && sh.seg_ok);   // a chip outlasts a chunk
                if (NV > 0) {                                                        // every code index inside the padded code
                    const bool in = (t_start > -0.999) && (t_stop < (double)(kPaddedChips - 1) - 0.001);
                    seg = sh.seg_ok && __all_sync(full, in) && seg_epoch_ok<(NV > 0 ? NV : 1)>(0.0, 0.0, sc.inv_step);
                }
                if (LEAN && !seg) {                                                  // leave this epoch to the general kernel
                    stop = true;
                    status = kNeedGeneral;
                }
                if (NV > 0) {
                    // lattice index in front of this CTA's first sample, from the prompt tap; every lane
                    // evaluates it next to its own tap so that no single-lane FP64 tail extends the chain
                    const double p_start = dadd(sc.rem_code, sh.cfgs.spacing[1]);
                    const double p_stop = dadd(dmul(sc.code_step, dn), p_start);
                    const double p_step = ddiv_by(dsub(p_stop, p_start), dn, sc.inv_n);
                    const int lead_s = (int)(sc.cur & (long long)(SPV - 1));
                    const int wlo = max((int)rank * Q * C - lead_s, 0);
                    hb = ceil_to_int(2.0 * code_phase(wlo, p_start, p_step)) - 1;
                    // segments hb .. ceil(2 phase(n)) cover the epoch; W (a power of two) warps take 32 per round
                    if (LEAN) rounds = (ceil_to_int(2.0 * p_stop) - hb + 32 * W) >> (5 + 31 - __clz(W));
                }
            }
            if (!stop) {
                SYDR_TICK(8)
                if (lane < 3) {
                    sh.ctl.ec.start[lane] = t_start;
                    sh.ctl.ec.step[lane] = t_step;
                    sh.ctl.ec.inv_step[lane] = sc.inv_step;                        // ~1/step': estimates only
                    if (NV > 0 && lane == 1) {
                        sh.ctl.ec.seg = seg;
                        sh.ctl.ec.hb = hb;
                        sh.ctl.ec.rounds = rounds;
                    }
                } else if (lane == 3) {
                    sh.ctl.ec.n = sc.n_req;
                    sh.ctl.ec.fast = fast;
                    sh.ctl.a = sc.cur;
                    sh.n_hist[epoch & 1] = sc.n_req;
                    if (S > 1) mbar_arrive_expect_tx(&sh.bar_gather[epoch & 1], (FIX ? 32u : 64u) * (uint32_t)n_ent);   // arm this epoch's gather
                }
            }
            if (lane == 5) { sh.ctl.stop = stop; sh.status = status; }
            SYDR_TICK(9)
        } else if (warp == 1) {
            double ca, cbb;
            float w[4][2];
            carrier_const(sk.carrier_freq, sk.rem_carrier, sh.K.inv_fs, ca, cbb, w);
            if (lane == 0) {
                sh.ctl.ec.ca = ca;
                sh.ctl.ec.cb = cbb;
            } else if (lane <= 4) {
                sh.ctl.ec.w[lane - 1][0] = w[lane - 1][0];
                sh.ctl.ec.w[lane - 1][1] = w[lane - 1][1];
            }
            if (LEAN) {
                constexpr int RB = RoundTraits<(NV > 0 ? NV : 1)>::kRotBase;
                carrier_table_entry(ca, RB + lane, sh.rot[lane].x, sh.rot[lane].y);
                if (lane < kRotMax - 32) carrier_table_entry(ca, RB + 32 + lane, sh.rot[32 + lane].x, sh.rot[32 + lane].y);
            } else if (lane < kMaxChunk && !(NV > 0 && sh.seg_ok)) {
                // only the split-sum chunk path reads the table; a channel on the half-chip lattice
                // never takes it (an epoch the segment path cannot serve uses the per-sample masks)
                carrier_table_entry(ca, lane, sh.ctl.ec.wtab[lane][0], sh.ctl.ec.wtab[lane][1]);
            }
            SYDR_TICK1(13)                             // carrier warp: constants of the next epoch
        }
        SYDR_TICK(0)                                   // epoch constants
        __syncthreads();                               // (A)
        if (sh.ctl.stop) break;
        SYDR_TICK(1)                                   // barrier (waits for the slower closing warp)

        const int buf = PACK ? 0 : (epoch & 1);
        const long long a = sh.ctl.a;
        const long long a0 = a & ~(long long)(SPV - 1);
        const int lead = (int)(a - a0);
        const long long wstart = (long long)rank * Q * C;          // relative to a0
        const int n_epoch = sh.ctl.ec.n;
        if (TMA && warp == W - 1 && lane == 0) {
            if (epoch == 0) trk_prefetch<DT, VPC>(sh, dyn_smem, rec_base, rec_alloc, a, rank, Q, 0);
            // next epoch's window, fetched while this one is correlated (PACK: behind the correlation, below)
            if (!PACK) trk_prefetch<DT, VPC>(sh, dyn_smem + (size_t)(buf ^ 1) * win_bytes, rec_base, rec_alloc, a + n_epoch, rank, Q, buf ^ 1);
        }

        // ---- (B) correlate this CTA's window
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int err = 0;
        if (TMA) mbar_wait(&sh.bar_data[buf], PACK ? (epoch & 1) : ((epoch >> 1) & 1));
        SYDR_TICK(2)                                   // wait for the staged window
        if (NV > 0 && LEAN) {
            correlate_rounds<(NV > 0 ? NV : 1)>(rec_base + a * BPS, sh.ctl.ec.rounds, sh.ctl.ec, sh.rot, sh.segtab, sh.seg_q,
                                                sh.cb, acc, err);
        } else if (NV > 0 && sh.ctl.ec.seg) {
            // half-chip segments whose first sample lies in this CTA's window
            const uint8_t* base = TMA ? dyn_smem + (size_t)buf * win_bytes : rec_base + (a0 + wstart) * BPS;
            const int wlo = (int)wstart - lead;
            const int whi = (rank == S - 1) ? 0x7fffffff : wlo + Q * C;
            for (int h = sh.ctl.ec.hb + tid;; h += blockDim.x)
                if (!correlate_segment<(NV > 0 ? NV : 1)>(base, wlo, whi, h, sh.ctl.ec, sh.segtab, sh.seg_q, sh.cb, acc, err)) break;
        } else if (!LEAN) {
            for (int q = tid; q < Q; q += blockDim.x) {
                const int j0 = (int)(wstart + (long long)q * C) - lead;
                if (j0 >= n_epoch) break;
                // ctl.ec is only rewritten after every warp of the cluster has delivered its sums
                if (TMA) {          // shared-memory window (address space known to the compiler: LDS.128)
                    const uint4* src = reinterpret_cast<const uint4*>(dyn_smem + (size_t)buf * win_bytes + (size_t)q * C * BPS);
                    correlate_chunk<DT, VPC>(src, j0, sh.ctl.ec, sh.cb, acc, err);
                } else {
                    const uint4* src = reinterpret_cast<const uint4*>(rec_base + (a0 + wstart + (long long)q * C) * BPS);
                    correlate_chunk<DT, VPC>(src, j0, sh.ctl.ec, sh.cb, acc, err);
                }
            }
        }
        SYDR_TICK(3)                                   // correlate (thread 0's chunks)
        const int slot = epoch & 1;
        if (FIX) {
            // every lane ends up with the six warp totals; lanes 2r and 2r+1 send the two 16-byte
            // halves of the warp's entry to CTA r
            int q[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) q[k] = __reduce_add_sync(full, __float2int_rn(acc[k] * P.acc_scale));
            const int eflag = __any_sync(full, err) ? 1 : 0;       // code-index overflow anywhere aborts the channel
            const int hi = lane & 1;
            const uint32_t v0 = (uint32_t)(hi ? q[4] : q[0]), v1 = (uint32_t)(hi ? q[5] : q[1]);
            const uint32_t v2 = (uint32_t)(hi ? eflag : q[2]), v3 = (uint32_t)(hi ? 0 : q[3]);
            int* dst = reinterpret_cast<int*>(&sh.gather[slot][0][0]) + (rank * W + warp) * 8 + hi * 4;
            if (S > 1) {
                if (lane < 2 * (int)S) {
                    const uint32_t r = (uint32_t)lane >> 1;
                    st_async_v4u(mapa_u32(smem_u32(dst), r), mapa_u32(smem_u32(&sh.bar_gather[slot]), r), v0, v1, v2, v3);
                }
            } else {
                if (lane < 2) *reinterpret_cast<uint4*>(dst) = make_uint4(v0, v1, v2, v3);
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.bar_gather[slot]);
            }
        } else {
            acc[6] = err ? 1.f : 0.f;
            const float wsum = warp_sum8(acc, lane);   // lane L: warp total of component sum8_index(L)
            const int c = sum8_index(lane);
            double* dst = &sh.gather[slot][rank * W + warp][c];
            const double wd = (double)wsum;            // the only float -> double conversion of the epoch
            if (S > 1) {
                const uint32_t d = smem_u32(dst), bar = smem_u32(&sh.bar_gather[slot]);
                for (uint32_t r = lane & 3; r < S; r += 4)
                    st_async_b64(mapa_u32(d, r), mapa_u32(bar, r), wd);
            } else {
                if ((lane & 3) == 0) *dst = wd;
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.bar_gather[slot]);
            }
        }
        SYDR_TICK(4)                                   // warp reduction + send
        if (PACK) {
            // the single window: every warp releases it, the last warp refills it with the next epoch's samples
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.bar_free);
            if (warp == W - 1 && lane == 0) {
                mbar_wait(&sh.bar_free, epoch & 1);
                trk_prefetch<DT, VPC>(sh, dyn_smem, rec_base, rec_alloc, a + n_epoch, rank, Q, 0);
            }
        }
        const int e = epoch;
        ++epoch;
        // ---- (C) close the loops of epoch e
        if (warp < 2) {
            // what the loop updates need besides the sums, while the all-gather is in flight
            CodePre cpre = {0.0, 0.0};
            double rc_next = 0.0;
            if (warp == 0) cpre = code_pre(sc);
            else if constexpr (KAP) rc_next = carrier_pre_kaplan(sh, sk, sh.n_hist[e & 1]);
            else rc_next = carrier_pre(sh, sk, sh.n_hist[e & 1]);
            mbar_wait(&sh.bar_gather[slot], (e >> 1) & 1);       // st.async data is visible once the phase completes
            SYDR_TICK(5)                               // all-gather: wait for the slowest warp of the cluster
            SYDR_TICK1(10)                             // carrier warp: everything up to the gather
            const double ck = FIX ? gather_total_fixed<(LEAN ? 2 : 18)>(sh, slot, n_ent, lane, P.acc_inv) : gather_total(sh, slot, n_ent, lane);
            SYDR_TICK(7)                               // totals
            SYDR_TICK1(11)
            sydr_trk_epoch* rec = out_row ? out_row + e : nullptr;
            if (warp == 0) {
                code_close(sh, sc, status, ck, rec, lane, cpre);
            } else {
                if constexpr (KAP) carrier_close_kaplan(sh, sk, kr, ck, rc_next, rec, kout_row ? kout_row + e : nullptr, lane);
                else carrier_close(sh, sk, ck, rc_next, rec, lane);
            }
            SYDR_TICK(6)                               // loop closure
            SYDR_TICK1(12)
        }
    }

    // the window of the epoch that will not run was already requested: drain it before exit
    if (TMA && epoch > 0) mbar_wait(&sh.bar_data[PACK ? 0 : (epoch & 1)], PACK ? (epoch & 1) : ((epoch >> 1) & 1));

    if (warp == 0 && lane == 0) sh.sc = sc;
    if (warp == 1 && lane == 0) sh.sk = sk;
    if constexpr (KAP) if (warp == 1 && lane == 0) {
        sydr_kaplan_state& g = sh.kcfg;
        g.ip_prev = kr.ip_prev; g.qp_prev = kr.qp_prev; g.fll_lock = kr.fll_lock; g.pll_lock = kr.pll_lock;
        g.cn0 = kr.cn0; g.pdpn = kr.pdpn; g.vel_memory = kr.vel_memory; g.fll_bw = kr.fll_bw; g.pll_bw = kr.pll_bw;
        g.accum_counter = kr.accum_counter; g.lock_state = kr.lock_state; g.flags = kr.flags;
        g.code_counter = kr.code_counter;
    }
    __syncthreads();
    if (prof && rank == 0) {
        for (int k = 0; k < 15; ++k) P.prof[ch * 16 + k] = sh.pc[k];
        P.prof[ch * 16 + 15] = epoch;
    }
    if (tid == 0 && rank == 0) {
        gst->cur = sh.sc.cur; gst->n_req = sh.sc.n_req; gst->epochs_done = sh.cfgs.epochs_done + epoch;
        gst->carrier_freq = sh.sk.carrier_freq; gst->code_freq = sh.sc.code_freq; gst->code_step = sh.sc.code_step;
        gst->rem_carrier = sh.sk.rem_carrier; gst->rem_code = sh.sc.rem_code;
        gst->nco_code = sh.sc.nco_code; gst->nco_code_err = sh.sc.nco_code_err;
        gst->nco_carrier = sh.sk.nco_carrier; gst->nco_carrier_err = sh.sk.nco_carrier_err;
        gst->status = sh.status;
        if constexpr (KAP) P.kstates[ch] = sh.kcfg;
        P.nepochs[ch] = sh.rec_base + epoch;
    }
    if (S > 1) cluster_sync_all();                  // nobody leaves while peers may still write here
}

}  // namespace sydr

using namespace sydr;

namespace {

int g_trk_mode = 0;                        // 0 = auto, 1 = chunk paths only (sydr_trk_set_mode)

template <int DT, int VPC>
int launch_epl(const void* d_iq, long long iq_len, double fs, const sydr_epl_args* d_args, int n_calls,
               const uint32_t* bits, double* d_out, cudaStream_t s) {
    epl_batch_kernel<DT, VPC><<<n_calls, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(d_iq), iq_len, fs, d_args,
                                                 bits, d_out, g_trk_mode == 0 ? 1 : 0);
    count_launch();
    SYDR_CUDA_CHECK(cudaGetLastError());
    return SYDR_OK;
}

template <int DT, int VPC>
int launch_trk(const TrkParams& P, int n_channels, int cluster, int threads, int lean, cudaStream_t s) {
    constexpr int C = IqTraits<DT>::SPV * VPC;
    const size_t smem = P.use_tma ? (size_t)(P.dense == 2 ? 1 : 2) * (P.Q * C + kWinTail) * IqTraits<DT>::BPS : 0;
    SYDR_REQUIRE(smem <= 200 * 1024, SYDR_ERR_UNSUPPORTED,
                 "tracking window needs %zu B of shared memory; raise cfg.cluster", smem);
    constexpr bool HAS_LEAN = SegTraits<DT, VPC>::NV > 0;
    auto kern = P.use_tma ? trk_borre_kernel<DT, VPC, true, false> : trk_borre_kernel<DT, VPC, false, false>;
    if (P.prof != nullptr) kern = P.use_tma ? trk_borre_kernel<DT, VPC, true, false, true> : trk_borre_kernel<DT, VPC, false, false, true>;
    if (P.dense && threads <= kDenseMaxThreads && P.prof == nullptr && P.kstates == nullptr)
        kern = P.use_tma ? trk_borre_kernel<DT, VPC, true, false, false, false, true>
                         : trk_borre_kernel<DT, VPC, false, false, false, false, true>;
    if (P.kstates != nullptr)
        kern = P.use_tma ? trk_borre_kernel<DT, VPC, true, false, false, true> : trk_borre_kernel<DT, VPC, false, false, false, true>;
    if constexpr (DT != SYDR_IQ_F32) {
        if (P.dense == 2) kern = trk_borre_kernel<DT, VPC, true, false, false, false, false, true>;     // (trk_run_impl checked the shape)
    }
    size_t smem_launch = smem;
    if constexpr (HAS_LEAN) {
        if (lean) {
            kern = trk_borre_kernel<DT, VPC, false, true>;
            smem_launch = 0;                                    // the throughput loop reads global memory directly
        }
    }
    SYDR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_launch));
    static const bool trace = getenv("SYDR_TRK_TRACE") != nullptr;      // diagnostics: the shape of every tracking launch
    if (trace)
        fprintf(stderr, "[sydr] trk launch: %d channels, cluster %d, %d threads, %s, %zu B dynamic smem\n", n_channels, cluster, threads,
                lean ? "LEAN" : (P.dense == 2 ? "PACK" : (P.kstates ? "KAP" : (P.dense ? "DENSE" : "latency"))), smem_launch);
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)(n_channels * cluster));
    lc.blockDim = dim3((unsigned)threads);
    lc.dynamicSmemBytes = smem_launch;
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

// Vectors per chunk.  The split-sum path needs a chunk no longer than the distance between two
// consecutive chip-boundary positions of any tap (`gap` chips: 0.5 for the usual -0.5/0/+0.5
// spacing), so the largest instantiated chunk that fits is chosen; a longer chunk is still
// correct (per-sample masks) but slower.
static int pick_vpc(int iq_dtype, double fs, double gap) {
    const double room = gap * fs / kCodeFreq;              // samples between boundary positions
    switch (iq_dtype) {
        case SYDR_IQ_I8:  return (room >= 24.0) ? 3 : 1;                         // 24 / 8 samples
        case SYDR_IQ_I16: return (room >= 20.0) ? 5 : (room >= 12.0) ? 3 : 1;    // 20 / 12 / 4 samples
        default:          return (room >= 20.0) ? 10 : (room >= 12.0) ? 6 : 2;   // 20 / 12 / 4 samples
    }
}

#define SYDR_DISPATCH_VPC(FN, DTYPE, VPCV, ...)                                        \
    switch ((DTYPE) * 16 + (VPCV)) {                                                   \
        case SYDR_IQ_I8 * 16 + 1:   return FN<SYDR_IQ_I8, 1>(__VA_ARGS__);             \
        case SYDR_IQ_I8 * 16 + 3:   return FN<SYDR_IQ_I8, 3>(__VA_ARGS__);             \
        case SYDR_IQ_I16 * 16 + 1:  return FN<SYDR_IQ_I16, 1>(__VA_ARGS__);            \
        case SYDR_IQ_I16 * 16 + 3:  return FN<SYDR_IQ_I16, 3>(__VA_ARGS__);            \
        case SYDR_IQ_I16 * 16 + 5:  return FN<SYDR_IQ_I16, 5>(__VA_ARGS__);            \
        case SYDR_IQ_F32 * 16 + 2:  return FN<SYDR_IQ_F32, 2>(__VA_ARGS__);            \
        case SYDR_IQ_F32 * 16 + 6:  return FN<SYDR_IQ_F32, 6>(__VA_ARGS__);            \
        case SYDR_IQ_F32 * 16 + 10: return FN<SYDR_IQ_F32, 10>(__VA_ARGS__);           \
        default: break;                                                                \
    }

static int dispatch_trk(int iq_dtype, int vpc, const TrkParams& P, int n_channels, int cluster, int threads, int lean,
                        cudaStream_t s) {
    SYDR_DISPATCH_VPC(launch_trk, iq_dtype, vpc, P, n_channels, cluster, threads, lean, s)
    set_error("sydr_trk_run: no kernel for this chunk size");
    return SYDR_ERR_UNSUPPORTED;
}

static long long* g_trk_prof = nullptr;    // set by sydr_trk_profile_buffer (diagnostics)

extern "C" {

// Diagnostics: give the tracking kernel a device buffer of n_channels*8 int64 to fill with
// per-phase cycle counts of the loop-closing thread (NULL switches it off).
int sydr_trk_profile_buffer(long long* d_buf) {
    g_trk_prof = d_buf;
    return SYDR_OK;
}

// Diagnostics / tests: 0 = automatic path selection (half-chip segments where they apply),
// 1 = chunk paths only.
int sydr_trk_set_mode(int mode) {
    SYDR_REQUIRE(mode == 0 || mode == 1, SYDR_ERR_ARG, "mode must be 0 or 1");
    g_trk_mode = mode;
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
    if (iq_dtype != SYDR_IQ_I8 && iq_dtype != SYDR_IQ_I16 && iq_dtype != SYDR_IQ_F32) {
        set_error("sydr_epl_batch: iq_dtype %d not supported (convert complex128 with sydr_convert_to_f32)", iq_dtype);
        return SYDR_ERR_UNSUPPORTED;
    }
    SYDR_DISPATCH_VPC(launch_epl, iq_dtype, pick_vpc(iq_dtype, fs, 0.5), d_iq, iq_len, fs, d_args, n_calls,
                      t.padded_bits, d_out, s)
    set_error("sydr_epl_batch: no kernel for this chunk size");
    return SYDR_ERR_UNSUPPORTED;
}

static int trk_run_impl(const void* d_iq, int iq_dtype, long long iq_alloc_samples, double fs, sydr_trk_state* d_states,
                        int n_channels, sydr_trk_epoch* d_out, int max_epochs, int* d_nepochs, const sydr_trk_config* cfg,
                        void* stream, sydr_kaplan_state* d_kstates, sydr_kaplan_epoch* d_kout);

int sydr_trk_run(const void* d_iq, int iq_dtype, long long iq_alloc_samples, double fs, sydr_trk_state* d_states,
                 int n_channels, sydr_trk_epoch* d_out, int max_epochs, int* d_nepochs, const sydr_trk_config* cfg,
                 void* stream) {
    return trk_run_impl(d_iq, iq_dtype, iq_alloc_samples, fs, d_states, n_channels, d_out, max_epochs, d_nepochs, cfg,
                        stream, nullptr, nullptr);
}

int sydr_trk_run_kaplan(const void* d_iq, int iq_dtype, long long iq_alloc_samples, double fs, sydr_trk_state* d_states,
                        sydr_kaplan_state* d_kstates, int n_channels, sydr_trk_epoch* d_out, sydr_kaplan_epoch* d_kout,
                        int max_epochs, int* d_nepochs, const sydr_trk_config* cfg, void* stream) {
    SYDR_REQUIRE(d_kstates && d_kout, SYDR_ERR_ARG, "NULL pointer");
    return trk_run_impl(d_iq, iq_dtype, iq_alloc_samples, fs, d_states, n_channels, d_out, max_epochs, d_nepochs, cfg,
                        stream, d_kstates, d_kout);
}

}  // extern "C"

static int trk_run_impl(const void* d_iq, int iq_dtype, long long iq_alloc_samples, double fs, sydr_trk_state* d_states,
                        int n_channels, sydr_trk_epoch* d_out, int max_epochs, int* d_nepochs, const sydr_trk_config* cfg,
                        void* stream, sydr_kaplan_state* d_kstates, sydr_kaplan_epoch* d_kout) {
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
    int spv, bps;
    switch (iq_dtype) {
        case SYDR_IQ_I8: spv = 8; bps = 2; break;
        case SYDR_IQ_I16: spv = 4; bps = 4; break;
        case SYDR_IQ_F32: spv = 2; bps = 8; break;
        default:
            set_error("sydr_trk_run: iq_dtype %d not supported", iq_dtype);
            return SYDR_ERR_UNSUPPORTED;
    }
    const double gap = (cfg && cfg->min_tap_gap > 0.0) ? cfg->min_tap_gap : 0.5;
    const int vpc = pick_vpc(iq_dtype, fs, gap);
    const int C = spv * vpc;
    // PACK (cfg.dense = 2): one CTA per channel, one window, two CTAs per SM -- when the window leaves room for that
    const long long win1 = ((n_max + spv + (long long)C - 1) / C * C + kWinTail) * bps;
    const bool auto_shape = !cfg || cfg->cluster <= 0;
    const bool staged_ok = cluster == 1 && use_tma && iq_dtype != SYDR_IQ_F32 && d_kstates == nullptr && g_trk_prof == nullptr;
    bool pack = cfg && cfg->dense == 2 && staged_ok && win1 <= kPackWindowBytes && (threads <= 0 || threads <= kPackThreads);
    // Automatic throughput shapes (more channels than clusters of two fit: cluster = 1), measured at 25 MS/s int16 on one
    // B200 (profiles/r2/pack_shapes.txt), SM time per channel-epoch: staged kernel, one CTA of 384 threads per SM 3.75 us
    // (<= 148 channels); PACK, two CTAs per SM 3.1 us (<= 296 channels); LEAN, three CTAs per SM 4.0 us (beyond).
    bool staged_one = false;
    const bool auto_throughput = auto_shape && staged_ok && !(cfg && cfg->kernel == 1) && threads <= 0;
    if (auto_throughput && n_channels > kPackWave && win1 <= kPackWindowBytes) {
        // more channels than one wave of the PACK shape holds: waves of 296 channels, one launch each on the same stream (a wave
        // of the LEAN shape holds 444 channels but takes twice as long), the remainder in the shape that fits it
        for (int c0 = 0; c0 < n_channels; c0 += kPackWave) {
            const int nc = (n_channels - c0 < kPackWave) ? n_channels - c0 : kPackWave;
            rc = trk_run_impl(d_iq, iq_dtype, iq_alloc_samples, fs, d_states + c0, nc, d_out + (size_t)c0 * max_epochs, max_epochs,
                              d_nepochs + c0, cfg, stream, nullptr, nullptr);
            if (rc != SYDR_OK) return rc;
        }
        return SYDR_OK;
    }
    if (auto_throughput) {
        if (n_channels <= 148 && 2 * win1 <= 200 * 1024) staged_one = true;
        else if (n_channels <= kPackWave && win1 <= kPackWindowBytes) pack = true;
    }
    if (pack && threads <= 0) threads = kPackThreads;
    if (staged_one) threads = kTrkMaxThreads;
    // shared-memory budget: two windows of Q*C samples
    while (!pack && use_tma && cluster < 8 && 2 * (((n_max + spv + (long long)C * cluster - 1) / ((long long)C * cluster)) * C + kWinTail) * bps > 200 * 1024)
        cluster <<= 1;                                     // (the chunk size chosen above is kept)
    const int Q = (int)((n_max + spv + (long long)C * cluster - 1) / ((long long)C * cluster));
    if (threads <= 0) {
        // Two thirds of a thread per chunk / segment: the epoch is a latency chain, not a throughput
        // problem (2.00 us with 192 threads, 2.01 with 288 at 25 MS/s), and the smaller CTA leaves
        // registers and issue slots to whatever shares the SM (other steps in flight, ColdStartPool).
        // (half a thread per segment in the DENSE shape: 160 threads at 25 MS/s, four CTAs per SM)
        const int want = (cfg && cfg->dense) ? (Q + 1) / 2 : (2 * Q + 2) / 3;
        const int rounds = (want + kTrkMaxThreads - 1) / kTrkMaxThreads;
        threads = (((want + rounds - 1) / rounds) + 31) / 32 * 32;
        if (threads < 64) threads = 64;                    // warps 0 and 1 close the two loops
    }
    if (d_kstates != nullptr && threads > kKaplanMaxThreads) threads = kKaplanMaxThreads;
    SYDR_REQUIRE(threads % 32 == 0 && threads >= 64 && threads <= kTrkMaxThreads, SYDR_ERR_ARG, "threads must be a multiple of 32 in [64, %d] (got %d)", kTrkMaxThreads, threads);

    // Throughput instantiation (LEAN): int16 IQ whose half chip fits the segment path, one CTA of
    // <= 256 threads per channel reading global memory directly, four CTAs per SM.  Chosen
    // automatically when the channels outnumber the SMs' latency-mode capacity, or explicitly
    // with cluster = 1, use_tma = 0 and threads <= 256.
    int nv = 0;
    if (iq_dtype == SYDR_IQ_I16) nv = (vpc == 1) ? 3 : (vpc == 3) ? 7 : (vpc == 5) ? 13 : 0;
    const double half_chip = 0.5 * fs / kCodeFreq;
    const bool seg_fits = nv > 0 && g_trk_mode == 0 && half_chip >= 2 * nv - 2 + 0.05 && half_chip <= 2 * nv - 1 - 0.05;
    const bool lean = !pack && !staged_one && d_kstates == nullptr && seg_fits && ((auto_shape && cluster == 1) ||
                                   (!auto_shape && cluster == 1 && !use_tma && threads > 0 && threads <= kLeanThreads &&
                                    (threads & (threads - 1)) == 0));   // power of two: rounds by shift

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
    P.has_iq_base = cfg ? (cfg->use_iq_base != 0) : 0;
    P.iq_base = P.has_iq_base ? cfg->iq_base : 0;
    SYDR_REQUIRE(!P.has_iq_base || (P.iq_base & 7) == 0, SYDR_ERR_ARG, "cfg.iq_base must be a multiple of 8 samples");
    P.seg = (g_trk_mode == 0) ? 1 : 0;
    // Fixed-point exchange: |warp total| <= samples the warp handles x the largest sample magnitude.
    // A warp never handles more than twice its even share of the longest epoch (+ edge segments);
    // the scale is the largest power of two that keeps that bound inside int32.
    auto set_scale = [&](TrkParams& Q_, int warps) {
        const double per_warp = 2.0 * (double)n_max / (double)(warps < 1 ? 1 : warps) + 128.0;
        const double maxmag = (iq_dtype == SYDR_IQ_I8) ? 182.0 : 46342.0;      // |I + jQ| of a full-scale sample
        int k = (int)floor(log2(2147483647.0 / (per_warp * maxmag * 1.01)));
        if (k > 20) k = 20;
        Q_.acc_scale = (float)ldexp(1.0, k);
        Q_.acc_inv = ldexp(1.0, -k);
    };
    P.resume = 0;
    P.prof = g_trk_prof;
    P.kstates = d_kstates;
    P.kout = d_kout;
    P.dense = pack ? 2 : (cfg ? (cfg->dense != 0) : 0);
    cudaStream_t s = (cudaStream_t)stream;
    // Prefix-moment kernel (trkm.cu): the per-sample work done once per recording (int16 IQ) -- only on request
    // (cfg.kernel = 1): measured slower than the LEAN instantiation on this GPU, so the automatic
    // choice stays with the per-channel kernels.  Channels it cannot serve stop with kNeedGeneral and continue in the
    // general launch below.
    const int kernel_sel = cfg ? cfg->kernel : 0;
    SYDR_REQUIRE(kernel_sel >= 0 && kernel_sel <= 2, SYDR_ERR_ARG, "cfg.kernel must be 0, 1 or 2 (got %d)", kernel_sel);
    const bool moments_ok = iq_dtype == SYDR_IQ_I16 && g_trk_mode == 0 && d_kstates == nullptr && (g_trk_prof == nullptr || kernel_sel == 1);
    SYDR_REQUIRE(kernel_sel != 1 || moments_ok, SYDR_ERR_UNSUPPORTED,
                 "cfg.kernel = 1 (prefix-moment kernel) needs int16 IQ, the Borre loops and trk mode 0");
    bool moments = moments_ok && kernel_sel == 1;
    if (moments) {
        TrkParams PL = P;
        PL.use_tma = 1;
        const int rc2 = launch_trkm(PL, n_channels, cfg ? cfg->rec_channels : 0, cfg ? cfg->group : 0, s);
        if (rc2 == SYDR_ERR_UNSUPPORTED && kernel_sel == 0) moments = false;       // more channels on one recording than the device holds: per-channel kernels
        else if (rc2 != SYDR_OK) return rc2;
    }
    if (moments) {
        P.resume = 1;
        P.prof = nullptr;                    // (the counters of a diagnostics run belong to the prefix-moment kernel)
        if (auto_shape) threads = kTrkMaxThreads;
    } else if (lean) {
        TrkParams PL = P;
        PL.use_tma = 0;
        const int lt = (!auto_shape && threads > 0) ? threads : kLeanThreads;
        set_scale(PL, lt / 32);
        const int rc2 = dispatch_trk(iq_dtype, vpc, PL, n_channels, 1, lt, 1, s);
        if (rc2 != SYDR_OK) return rc2;
        // The general instantiation follows on the same stream: channels the LEAN kernel finished
        // exit at once; a channel it stopped in front of an epoch it cannot serve continues here.
        P.resume = 1;
        if (auto_shape) threads = kTrkMaxThreads;
    }
    set_scale(P, cluster * (threads / 32));
    return dispatch_trk(iq_dtype, vpc, P, n_channels, cluster, threads, 0, s);
}

extern "C" {

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
