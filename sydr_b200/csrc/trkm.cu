// K-TRKM: closed-loop E/P/L tracking in the prefix-moment formulation, throughput shape, for sm_100a.
//
// Reference semantics (file:line in /root/reference): EPL sydr/dsp/tracking.py:92-116, DLL_NNEML / PLL_costa
// tracking.py:120-142, BorreLoopFilter tracking.py:180-186, NCO update sydr/channel/channel_l1ca_borre.py:363-429.
//
// What the north star asks for -- "loads each int16 IQ tile once into shared memory and correlates every channel
// against it" -- with the per-sample work made independent of the channels:
//
//   One CTA = `group` channels (<= 4) of one recording, 16 equal warps, no roles.  A warp takes the next block of the
//   recording (736 new samples) by ticket, loads it once (24 samples per lane, 128-bit loads), and turns it into exact
//   integer prefix moments in a private 12 KB slot of shared memory, 16 bytes per sample:
//       E0[j] = sum_{i<j} x_i            E1[j] = sum_{i<j} E0[i+1] = sum_{m<j} (j-m) x_m      (complex, int32, mod 2^32)
//   counted from the block's first sample, 768 entries: the 32 extra ones let every segment that STARTS in the block end
//   in it (a segment is at most 31 samples long), so blocks do not depend on each other and nothing is exchanged between
//   warps but the epoch sums.  (IDP.2A unpacks an int16 I or Q and accumulates it in one instruction; the second moment is
//   one IADD per sample; a shuffle scan joins the lanes.  Wrap-around of the sums is harmless: only differences over <= 31
//   samples are used.)  The same warp then correlates every channel of the CTA against its block:
//       Between two consecutive chip-boundary samples a <= j < b of a channel (half a chip, ~12.2 samples at 25 MS/s) the
//       three code replicas are constant and the carrier advances by a few milliradians, so
//           sum_j x_j exp(i phi_j) = exp(i phi_c) [ S0 (1 - alpha^2 (L^2-1)/24) + i alpha/2 S1 ] + O(1e-6 S0),
//           S0 = E0[b]-E0[a],  W = E1[b]-E1[a] - L E0[a],  S1 = (L+1) S0 - 2 W,  L = b-a,  c = (a+b-1)/2,  alpha = -2 pi fc/fs.
//       One lane handles one boundary: one 16-byte read of the slot (the far end comes from the neighbouring lane), ~75
//       instructions per segment, no per-sample work.  The boundaries are located exactly as in trk.cu (the reference's own
//       FP64 expression ceil(fl(fl(j step') + start)) decides every sample within 1e-9 of a lattice crossing).
//   Epochs: the sums a warp collects for (channel, epoch) inside its block go, in fixed point, into the channel's epoch
//   accumulators (64-bit shared-memory atomics: integer addition, order independent); the warp that delivers the last
//   of the blocks an epoch overlaps closes the loops with the same FP64 code as K-TRK (trk_common.cuh: the NCO trajectory
//   arithmetic is the reference's, operation for operation), writes the epoch record and publishes the next epoch's
//   constants.  A warp whose block reaches into an epoch that is not published yet serves its other channels first.
//   A dependency always points to a block with a smaller number, so nothing can wait in a circle.
//
// Conditions (else the channel stops with status kNeedGeneral and the general kernel queued behind serves it, exactly
// like the LEAN instantiation of trk.cu): int16 IQ, spacings -0.5 / 0 / +0.5 chip around the prompt tap, half a chip
// between 1 and 30 samples, |alpha| x half chip <= 0.06 rad (|carrier| <= 19.5 kHz: error bound 4e-6 of the prompt
// magnitude at 45 dB-Hz), every code index inside the padded code, the channels of a CTA on the
// same recording (iq_base, a multiple of 4 samples).
#include "trk_common.cuh"

namespace sydr {

constexpr int kMWarps = 16;                     // warps per CTA
constexpr int kMThreads = 32 * kMWarps;
constexpr int kMMaxGroup = 4;                   // channels per CTA
constexpr int kMKS = 24;                        // samples per lane and block
constexpr int kMEnt = 32 * kMKS;                // entries per block: 768
constexpr int kMNew = kMEnt - 32;               // samples a block advances by: 736
constexpr int kMSlotBytes = kMEnt * 16;         // 12 KB per warp
constexpr int kMSeg = 31;                       // segments per round (32 boundaries)
constexpr int kMMaxPieces = 40;                 // blocks an epoch may overlap (an epoch is at most 39 * 736 - 735 samples long)
constexpr long long kMMaxSpan = 0x7fffffffLL - (1LL << 20);      // block-relative arithmetic is 32-bit
static_assert(kMKS % 4 == 0, "a lane loads whole 16-byte vectors");

struct MCtl {                 // per-epoch constants of one channel, published by the warp that closed the epoch before
    double start[3], step[3]; // numpy linspace constants of the three taps (tracking.py:110-112)
    double inv_step;          // ~1/step' of the prompt tap
    double ca, cb;            // carrier phase in turns at epoch-relative sample j: ca*j + cb
    long long a;              // epoch start sample (recording-relative)
    float ah, g2;             // alpha/2 = pi*ca; alpha^2/24
    int n, p0;
    int n_blocks;             // blocks the epoch overlaps: that many deliveries complete it
    int first_block;          // the first of them
    volatile int epoch;       // which epoch these constants belong to (-1 while they are being rewritten)
    int pad;
};

struct MChan {                // shared-memory state of one channel
    uint32_t cb[kCodeWords];                    // padded code, one bit per chip
    uint16_t stab[kPaddedChips + 1];            // byte 0 / 1 of entry k: top byte of +-1.0f for chips k / k+1
    MCtl ctl[2];                                // [epoch & 1]
    alignas(16) int part[2][kMMaxPieces][8];    // fixed-point sums (6) + error flag of the pieces of an epoch, [epoch & 1][block]
    int cnt[2];                                 // deliveries so far
    volatile int pub;                           // newest epoch whose constants are published (-1: none)
    volatile int stop_epoch;                    // the channel does not track this epoch or any later one
    sydr_trk_state cfgs;
    CodeState sc;
    CarrierState sk;
    LoopConst K;
    int rec_base, status, ch, active;
};

struct MShared {
    MChan chan[kMMaxGroup];
    float fix[kMWarps][8];                      // exact-evaluation corrections of ambiguous samples, per warp
    int ep[kMWarps][kMMaxGroup];                // per warp and channel: the epoch the warp's next block starts in (never decreases)
    long long origin;                           // recording-relative sample of block 0's first sample (multiple of 4)
    long long valid_lo, valid_hi;               // samples readable from the recording's base pointer
    long long iq_base, limit;
    int n_blocks, ok;
    int next_block;                             // ticket counter
    volatile int running;                       // channels still tracking
};

struct TrkmParams {
    TrkParams t;
    int n_channels;
    int group;                // channels per CTA
    int rec_channels;         // channel slots per recording (0: the channels form one sequence)
    int groups_per_rec;       // CTAs per recording
    double alpha_hc_max;      // |alpha| * half chip limit of the expansion
    int debug;
};

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, int a, int b, int c, int d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Entry of in-block sample s: lane l = s / 24 made it as its v-th (s = 24 l + v); stored lane-interleaved and skewed,
// word v * 32 + (l + v) % 32: the 32 lanes of a store write 32 consecutive 16-byte words, and the boundaries a round
// gathers (12 samples apart: two per producer lane) fall into different bank groups.
__device__ __forceinline__ uint32_t m_entry_offset(int s) {
    const int l = (s * 2731) >> 16;             // s / 24 for 0 <= s < 800
    const int v = s - kMKS * l;
    return (uint32_t)(v * 32 + ((l + v) & 31)) * 16u;
}

// ---- a block's prefix moments ---------------------------------------------------------------------------------------
__device__ __forceinline__ void m_produce(const TrkmParams& PM, const MShared& sh, uint32_t slot, int t, int lane) {
    const unsigned full = 0xffffffffu;
    const uint32_t* rec_ptr = reinterpret_cast<const uint32_t*>(PM.t.iq + sh.iq_base * 4);
    const long long b0 = sh.origin + (long long)t * kMNew;             // the block's first sample
    const long long s0 = b0 + lane * kMKS;                              // this lane's first sample
    uint32_t w[kMKS];
    if (b0 >= sh.valid_lo && b0 + kMEnt <= sh.valid_hi) {
        const uint4* src = reinterpret_cast<const uint4*>(rec_ptr + s0);
#pragma unroll
        for (int v = 0; v < kMKS / 4; ++v) {
            const uint4 q = __ldg(src + v);
            w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
        }
    } else {                                                           // edge of the allocation: guarded loads, zeros outside
#pragma unroll
        for (int v = 0; v < kMKS; ++v) {
            const long long s = s0 + v;
            w[v] = (s >= sh.valid_lo && s < sh.valid_hi) ? rec_ptr[s] : 0u;
        }
    }
    // pass 1: this lane's totals.  dp2a: I = low int16 of the word, Q = high int16.  t1 = sum_v (KS - v) x_v.
    int t0r = 0, t0i = 0, t1r = 0, t1i = 0;
#pragma unroll
    for (int v = 0; v < kMKS; ++v) {
        t0r = __dp2a_lo((int)w[v], 0x0001, t0r);
        t0i = __dp2a_lo((int)w[v], 0x0100, t0i);
        t1r += t0r;
        t1i += t0i;
    }
    // exclusive scan over the warp.  With u_m = t1_m - KS (m + 1) t0_m the second moment in front of lane l is
    // sum_{m<l} u_m + KS l sum_{m<l} t0_m: two plain sum scans.
    int s0r = t0r, s0i = t0i, s1r = t1r - kMKS * (lane + 1) * t0r, s1i = t1i - kMKS * (lane + 1) * t0i;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a0 = __shfl_up_sync(full, s0r, o), a1 = __shfl_up_sync(full, s0i, o);
        const int a2 = __shfl_up_sync(full, s1r, o), a3 = __shfl_up_sync(full, s1i, o);
        if (lane >= o) { s0r += a0; s0i += a1; s1r += a2; s1i += a3; }
    }
    int a0r = __shfl_up_sync(full, s0r, 1), a0i = __shfl_up_sync(full, s0i, 1);
    int a1r = __shfl_up_sync(full, s1r, 1), a1i = __shfl_up_sync(full, s1i, 1);
    if (lane == 0) { a0r = a0i = a1r = a1i = 0; }
    a1r += kMKS * lane * a0r;
    a1i += kMKS * lane * a0i;
    // pass 2: the entries of this lane's samples (exclusive prefixes), lane-interleaved
#pragma unroll
    for (int v = 0; v < kMKS; ++v) {
        sts128(slot + (uint32_t)(v * 32 + ((lane + v) & 31)) * 16u, a0r, a0i, a1r, a1i);
        a0r = __dp2a_lo((int)w[v], 0x0001, a0r);
        a0i = __dp2a_lo((int)w[v], 0x0100, a0i);
        a1r += a0r;
        a1i += a0i;
    }
    __syncwarp();
}

// Exact treatment of an ambiguous boundary (the crossing is within 1e-9 sample of an integer J): sample J was given to
// the segment behind the boundary; its three code indices are re-evaluated with the reference expression and the
// difference, times the wiped-off sample, goes to the warp's correction sums.
static __device__ __noinline__ void m_correct(MChan& ch, const MCtl& c, float* fix, uint32_t slot, int s_in_block, int J, int p) {
    const uint4 e0 = lds128(slot + m_entry_offset(s_in_block)), e1 = lds128(slot + m_entry_offset(s_in_block + 1));
    const float xr = (float)(int)(e1.x - e0.x), xi = (float)(int)(e1.y - e0.y);
    double turns = fma(c.ca, i2d(J), c.cb);
    turns -= drint(turns);
    float pr, pi;
    __sincosf((float)turns * 6.283185307179586f, &pi, &pr);
    const float zr = pr * xr - pi * xi, zi = pr * xi + pi * xr;
    int err = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int k_exact = ceil_to_int(code_phase(J, c.start[s], c.step[s]));
        const int k_seg = (p + 1 + s) >> 1;                       // ceil((H + q_s) / 2), H = p + 1, q = (-1, 0, +1)
        if (k_exact != k_seg) {
            const float d = sign_of_bit(chip_bit(ch.cb, k_exact, err)) - sign_of_bit(chip_bit(ch.cb, k_seg, err));
            atomicAdd(&fix[2 * s], d * zr);
            atomicAdd(&fix[2 * s + 1], d * zi);
        }
    }
    if (err) atomicAdd(&fix[6], 1.0f);
}

// The constants of the channel's next epoch from its NCO state (the warp that closed the epoch before, or set-up):
// ctl[epoch & 1], then `pub`.  An epoch the kernel cannot serve, the end of the data or of the record array stop the channel.
__device__ __forceinline__ void m_publish(const TrkmParams& PM, MShared& sh, MChan& ch, int epoch, int lane) {
    const unsigned full = 0xffffffffu;
    const TrkParams& P = PM.t;
    CodeState& sc = ch.sc;
    int status = ch.status;
    if (sc.n_req <= 0 || sc.n_req > 0x3fffffff) status = SYDR_ERR_STATE;
    const long long rec_alloc = P.iq_alloc - ch.cfgs.iq_base;
    const long long iq_len_reg = min(min((long long)ch.cfgs.iq_len, rec_alloc), sh.limit);
    bool stop = (status != 0) || (epoch >= P.max_epochs - ch.rec_base) || (sc.cur + sc.n_req > iq_len_reg);
    double t_start = 0.0, t_step = 0.0, t_stop = 0.0, inv_n = sc.inv_n;
    int p0 = 0;
    if (!stop) {
        const double dn = i2d(sc.n_req);
        inv_n = newton_rcp(dn, newton_rcp(dn, inv_n));
        t_start = dadd(sc.rem_code, ch.cfgs.spacing[min(lane, 2)]);             // tracking.py:110
        t_stop = dadd(dmul(sc.code_step, dn), t_start);
        t_step = ddiv_by(dsub(t_stop, t_start), dn, inv_n);                     // numpy linspace step
        const bool in = (t_start > -0.999) && (t_stop < (double)(kPaddedChips - 1) - 0.001);
        const double hc = 0.5 * sc.inv_step;                                     // half a chip, in samples
        const bool ok = __all_sync(full, in) && hc >= 1.0 && hc <= 30.0 && sc.n_req <= (kMMaxPieces - 1) * kMNew - kMNew + 1;
        const double p_start = dadd(sc.rem_code, ch.cfgs.spacing[1]);
        p0 = ceil_to_int(2.0 * p_start) - 1;                                     // lattice point in front of sample 0
        double ca, cbb;
        float w[4][2];
        carrier_const(ch.sk.carrier_freq, ch.sk.rem_carrier, ch.K.inv_fs, ca, cbb, w);
        const bool car_ok = fabs(2.0 * kPi * ca) * (0.5 * P.fs / kCodeFreq) <= PM.alpha_hc_max;   // the expansion needs |alpha| x half chip small
        if (!ok || !car_ok) { stop = true; status = kNeedGeneral; }
        if (!stop) {
            MCtl& c = ch.ctl[epoch & 1];
            if (lane == 0) c.epoch = -1;                                         // (a reader two epochs late must not take a half-written set)
            __syncwarp();
            __threadfence_block();
            if (lane < 3) {
                c.start[lane] = t_start;
                c.step[lane] = t_step;
            }
            if (lane == 0) {
                c.inv_step = sc.inv_step;
                c.p0 = p0;
                c.n = sc.n_req;
                c.a = sc.cur;
                c.ca = ca;
                c.cb = cbb;
                c.ah = (float)(kPi * ca);
                c.g2 = (float)((2.0 * kPi * ca) * (2.0 * kPi * ca) * (1.0 / 24.0));
                const long long lo = sc.cur - sh.origin, hi = lo + sc.n_req - 1;
                c.first_block = (int)(lo / kMNew);
                c.n_blocks = (int)(hi / kMNew) - c.first_block + 1;
                sc.inv_n = inv_n;
            }
        }
    }
    __syncwarp();
    __threadfence_block();
    if (lane == 0) {
        ch.status = status;
        if (stop) {
            ch.stop_epoch = epoch;
            atomicSub((int*)&sh.running, 1);
        } else {
            ch.ctl[epoch & 1].epoch = epoch;
            __threadfence_block();
            ch.pub = epoch;
        }
    }
    __syncwarp();
}

// Close epoch e of the channel (every delivery is in): totals -> DLL / PLL loop closure -> record -> next epoch's constants.
__device__ __forceinline__ void m_close(const TrkmParams& PM, MShared& sh, MChan& ch, int e, int lane) {
    const TrkParams& P = PM.t;
    const int slot = e & 1;
    // lane L adds the pieces (L >> 3) mod 4 of component L & 7: integer addition, the order does not matter
    const int n_pieces = ch.ctl[slot].n_blocks;
    long long tot = 0;
    for (int i = lane >> 3; i < n_pieces; i += 4) tot += ch.part[slot][i][lane & 7];
    tot += __shfl_xor_sync(0xffffffffu, tot, 8);
    tot += __shfl_xor_sync(0xffffffffu, tot, 16);
    if (lane == 8) ch.cnt[slot] = 0;                             // (the slot serves epoch e + 2, after the closure of e + 1)
    const double ck = (double)tot * P.acc_inv;
    sydr_trk_epoch* rec = P.out + (long long)ch.ch * P.max_epochs + ch.rec_base + e;
    const int n_epoch = ch.ctl[slot].n;
    const CodePre cpre = code_pre(ch.sc);
    const double rc_next = carrier_pre(ch, ch.sk, n_epoch);
    CodeState sc = ch.sc;
    CarrierState sk = ch.sk;
    int status = ch.status;
    code_close(ch, sc, status, ck, rec, lane, cpre);
    carrier_close(ch, sk, ck, rc_next, rec, lane);
    __syncwarp();
    if (lane == 0) {
        ch.sc = sc;
        ch.sk = sk;
        ch.status = status;
    }
    __syncwarp();
    m_publish(PM, sh, ch, e + 1, lane);
}

template <bool PROF>
__global__ void __launch_bounds__(kMThreads, 1) trkm_kernel(const TrkmParams PM) {
    extern __shared__ __align__(1024) uint8_t dyn_smem[];
    __shared__ __align__(16) MShared sh;
    const TrkParams& P = PM.t;
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = PM.group;
    // CTA -> channels: `group` consecutive channels, never across two recordings when the caller named the slots per recording
    int ch_lo, ch_hi;
    if (PM.rec_channels > 0) {
        const int r = blockIdx.x / PM.groups_per_rec, c_in_rec = blockIdx.x % PM.groups_per_rec;
        ch_lo = r * PM.rec_channels + c_in_rec * G;
        ch_hi = min(min((r + 1) * PM.rec_channels, PM.n_channels), ch_lo + G);
    } else {
        ch_lo = blockIdx.x * G;
        ch_hi = min(PM.n_channels, ch_lo + G);
    }
    const uint32_t slot = smem_u32(dyn_smem) + (uint32_t)warp * kMSlotBytes;

    // ---- set-up: channel states, the CTA's share of the recording, code tables, the first epoch's constants
    if (tid < G) {
        MChan& ch = sh.chan[tid];
        const int idx = ch_lo + tid;
        ch.ch = idx;
        ch.active = 0;
        ch.pub = -1;
        ch.stop_epoch = 0x7fffffff;
        ch.rec_base = 0;
        ch.cnt[0] = ch.cnt[1] = 0;
        if (idx < ch_hi) {
            ch.cfgs = P.states[idx];
            if (P.iq_len > 0) ch.cfgs.iq_len = P.iq_len;
            if (P.has_iq_base) ch.cfgs.iq_base = P.iq_base;
            ch.rec_base = P.append ? (int)ch.cfgs.epochs_done : 0;
            if ((unsigned)(ch.cfgs.prn - 1) >= (unsigned)kMaxPrn && ch.cfgs.status == 0) ch.cfgs.status = SYDR_ERR_STATE;
            const sydr_trk_state& g = ch.cfgs;
            ch.sc.cur = g.cur; ch.sc.n_req = (int)g.n_req;
            ch.sc.code_freq = g.code_freq; ch.sc.code_step = g.code_step; ch.sc.rem_code = g.rem_code;
            ch.sc.nco_code_err = g.nco_code_err; ch.sc.nco_code = g.nco_code;
            ch.sc.inv_step = drcp(g.code_step); ch.sc.inv_n = drcp((double)g.n_req);
            ch.sk.carrier_freq = g.carrier_freq; ch.sk.rem_carrier = g.rem_carrier;
            ch.sk.nco_carrier_err = g.nco_carrier_err; ch.sk.nco_carrier = g.nco_carrier;
            ch.K.fs = P.fs; ch.K.inv_fs = 1.0 / P.fs;
            ch.K.dll_c1 = g.dll_tau2 / g.dll_tau1; ch.K.dll_c2 = g.dll_pdi / g.dll_tau1;
            ch.K.pll_c1 = g.pll_tau2 / g.pll_tau1; ch.K.pll_c2 = g.pll_pdi / g.pll_tau1;
            ch.status = g.status;
            ch.active = (g.status == 0) ? 1 : 0;
            // spacings on the half-chip lattice at -1 / 0 / +1 half chips around the prompt tap
            int q[3];
            if (ch.active && !(seg_tap_offsets(g.spacing, q) && q[0] == -1 && q[2] == 1)) {
                ch.status = kNeedGeneral;
                ch.active = 0;
            }
        } else {
            ch.ch = -1;
            ch.status = 1;
            ch.cfgs.status = 1;
        }
    }
    __syncthreads();
    if (tid == 0) {
        // one recording (the first active channel's; a channel of another one is left to the general kernel), the samples from
        // the earliest epoch start to the latest end
        long long base = 0, lo = 0x7fffffffffffffffLL, hi = 0;
        int have = 0, running = 0;
        for (int c = 0; c < G; ++c) {
            MChan& ch = sh.chan[c];
            if (!ch.active) continue;
            const sydr_trk_state& g = ch.cfgs;
            if (!have) { base = g.iq_base; have = 1; }
            if (g.iq_base != base || (base & 3) != 0 || g.cur < 0) {
                ch.status = kNeedGeneral;
                ch.active = 0;
                continue;
            }
            lo = min(lo, (long long)g.cur);
            hi = max(hi, min((long long)g.iq_len, P.iq_alloc - base));
            ++running;
        }
        sh.iq_base = base;
        sh.ok = (running > 0 && hi > lo) ? 1 : 0;
        sh.origin = sh.ok ? (lo & ~3LL) : 0;
        sh.valid_lo = max(0LL, -base);
        sh.valid_hi = P.iq_alloc - base;
        long long span = hi - sh.origin;
        if (span > kMMaxSpan) span = kMMaxSpan;                        // the general kernel continues behind
        sh.limit = sh.origin + span;
        sh.n_blocks = sh.ok ? (int)(span / kMNew) + 1 : 0;
        if (!sh.ok)
            for (int c = 0; c < G; ++c)
                if (sh.chan[c].active) { sh.chan[c].status = kNeedGeneral; sh.chan[c].active = 0; --running; }
        sh.running = max(running, 0);
        sh.next_block = 0;
    }
    if (tid < kMWarps * 8) sh.fix[tid >> 3][tid & 7] = 0.f;
    __syncthreads();
    for (int c = 0; c < G; ++c) {
        MChan& ch = sh.chan[c];
        if (!ch.active) continue;
        if (tid < kCodeWords) ch.cb[tid] = P.code_bits[(ch.cfgs.prn - 1) * kCodeWords + tid];
    }
    __syncthreads();
    for (int c = 0; c < G; ++c) {
        MChan& ch = sh.chan[c];
        if (!ch.active) continue;
        for (int k = tid; k <= kPaddedChips; k += kMThreads) {
            const int k0 = min(k, kPaddedChips - 1), k1 = min(k + 1, kPaddedChips - 1);
            const uint32_t b0 = (ch.cb[k0 >> 5] >> (k0 & 31)) & 1u, b1 = (ch.cb[k1 >> 5] >> (k1 & 31)) & 1u;
            ch.stab[k] = (uint16_t)((b0 ? 0x3Fu : 0xBFu) | ((b1 ? 0x3Fu : 0xBFu) << 8));
        }
    }
    if (warp < G && sh.chan[warp].active) m_publish(PM, sh, sh.chan[warp], 0, lane);      // epoch 0
    __syncthreads();

    // ---- the blocks
    int* ep = sh.ep[warp];
    if (lane < kMMaxGroup) ep[lane] = 0;
    __syncwarp();
    float* fix = sh.fix[warp];
    long long pt0 = 0, pt = 0, c_prod = 0, c_piece = 0, c_wait = 0, c_close = 0, n_blk = 0, n_piece = 0, n_close = 0;
    if (PROF) pt0 = clock64();
    while (sh.ok) {
        if (PROF) pt = clock64();
        int t = 0;
        if (lane == 0) t = (sh.running > 0) ? atomicAdd(&sh.next_block, 1) : 0x7fffffff;
        t = __shfl_sync(full, t, 0);                                  // (one lane looks: the whole warp leaves or stays)
        if (t >= sh.n_blocks) break;
        m_produce(PM, sh, slot, t, lane);
        if (PROF) { const long long now = clock64(); c_prod += now - pt; pt = now; ++n_blk; }
        const int s_lo = t * kMNew, s_hi = s_lo + kMNew;              // the samples whose segments this block owns (block-0 relative)
        unsigned todo = 0;
        for (int c = 0; c < G; ++c)
            if (sh.chan[c].active) todo |= 1u << c;
        // The channel whose epoch ends soonest behind this block's first sample goes first: its pieces are what the next loop
        // closure waits for, and the warps with later blocks wait for that closure.
        int order = 0;
        {
            int key[kMMaxGroup];
#pragma unroll
            for (int c = 0; c < kMMaxGroup; ++c) {
                key[c] = 0x7fffffff;
                if (c < G && (todo >> c & 1u)) {
                    const MChan& ch = sh.chan[c];
                    const int e = ep[c];
                    if (ch.pub >= e && ch.stop_epoch > e) {
                        const MCtl& ctl = ch.ctl[e & 1];
                        long long k = ctl.a - sh.origin + ctl.n - s_lo;    // samples from the block's start to the epoch's end
                        if (k <= 0) k += ctl.n;                            // (that epoch is over: its successor's end, roughly)
                        key[c] = (int)min(k, 0x7ffffff0LL);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < kMMaxGroup; ++c) {
                int rank = 0;
#pragma unroll
                for (int d = 0; d < kMMaxGroup; ++d) rank += (key[d] < key[c] || (key[d] == key[c] && d < c)) ? 1 : 0;
                order |= c << (2 * rank);
            }
            order = __shfl_sync(full, order, 0);                          // one lane's view for all
        }
        int spins = 0;
        while (todo) {
            bool progressed = false;
            for (int k = 0; k < kMMaxGroup; ++k) {
                const int c = (order >> (2 * k)) & 3;
                if (c >= G || !(todo >> c & 1u)) continue;
                MChan& ch = sh.chan[c];
                // one (block, epoch) piece per pass: the part of epoch ep[c] that starts inside this block
                const int e = ep[c];
                const int stop_at = __shfl_sync(full, ch.stop_epoch, 0), pub = __shfl_sync(full, ch.pub, 0);   // one lane's view for all
                if (stop_at <= e) { todo &= ~(1u << c); progressed = true; continue; }
                if (pub < e) continue;                                // not published yet: the other channels first
                __threadfence_block();
                const MCtl& ctl = ch.ctl[e & 1];
                {
                    // The slot holds epoch e unless this warp comes two or more epochs late (every epoch in between is then
                    // closed without this block, i.e. lies in front of it): take up the epoch the slot holds now.
                    const int held = __shfl_sync(full, ctl.epoch, 0);
                    if (held != e) {
                        if (held > e) { ep[c] = held; progressed = true; }
                        continue;
                    }
                }
                const int n = ctl.n;
                const long long a64 = ctl.a - sh.origin;              // epoch start, block-0 relative
                if (a64 >= s_hi) { todo &= ~(1u << c); progressed = true; continue; }      // the channel starts behind this block
                const int a_rel = (int)a64;
                if (a_rel + n <= s_lo) { ep[c] = e + 1; progressed = true; continue; }     // this epoch ended in front of the block
                const int j_lo = max(s_lo - a_rel, 0), j_hi = min(s_hi - a_rel, n);        // epoch-relative samples of the piece
                // ---- correlate: segments that start in [j_lo, j_hi), 31 per round
                const double inv_step = ctl.inv_step;
                const float ah = ctl.ah, g2 = ctl.g2;
                const double ca_half = 0.5 * ctl.ca, cbt = ctl.cb;
                // lattice point in front of sample j_lo (one early: the rounding of this estimate must not skip a segment)
                int p = ceil_to_int(2.0 * fma(ctl.step[1], i2d(j_lo), ctl.start[1])) - 2;
                p = max(p, ctl.p0 - 1) + lane;
                double x = seg_crossing(dmul(0.5, i2d(p)), ctl.start[1], inv_step);
                const double dx = (double)kMSeg * 0.5 * inv_step;
                float aEr = 0.f, aEi = 0.f, aPr = 0.f, aPi = 0.f, aLr = 0.f, aLi = 0.f;
                while (true) {
                    bool amb;
                    const int B = seg_first_sample(x, amb);
                    const int Bc = min(max(B, 0), n);
                    if (__shfl_sync(full, Bc, 0) >= j_hi) break;       // the round starts behind the piece
                    const int Bn = __shfl_down_sync(full, Bc, 1);
                    const int L = Bn - Bc;                              // 0 for clipped segments
                    const bool mine = lane < kMSeg && Bc >= j_lo && Bc < j_hi && L > 0;
                    const int sb = min(max(a_rel + Bc - s_lo, 0), kMEnt - 1);             // in-block sample of the boundary
                    const uint4 e0 = lds128(slot + m_entry_offset(sb));
                    uint4 e1;
                    e1.x = __shfl_down_sync(full, e0.x, 1); e1.y = __shfl_down_sync(full, e0.y, 1);
                    e1.z = __shfl_down_sync(full, e0.z, 1); e1.w = __shfl_down_sync(full, e0.w, 1);
                    if (mine) {
                        const int d0r = (int)(e1.x - e0.x), d0i = (int)(e1.y - e0.y);   // S0
                        // W = sum (b - m) x_m = E1[b] - E1[a] - L E0[a];  S1 = 2 sum (m - c) x_m = (L + 1) S0 - 2 W   (mod 2^32, exact)
                        const int wr = (int)(e1.z - e0.z) - L * (int)e0.x, wi = (int)(e1.w - e0.w) - L * (int)e0.y;
                        const int d1r = (L + 1) * d0r - 2 * wr, d1i = (L + 1) * d0i - 2 * wi;
                        const float s0r = (float)d0r, s0i = (float)d0i, s1r = (float)d1r, s1i = (float)d1i;
                        const float g = 1.0f - g2 * (float)(L * L - 1);
                        const float yr = fmaf(g, s0r, -(ah * s1i)), yi = fmaf(g, s0i, ah * s1r);
                        // carrier phasor at the segment centre c = (2 Bc + L - 1) / 2 (tracking.py:102), FP64 turns
                        double turns = fma(ca_half, i2d(2 * Bc + L - 1), cbt);
                        turns -= drint(turns);
                        float pre, pim;
                        __sincosf((float)turns * 6.283185307179586f, &pim, &pre);
                        const float zr = pre * yr - pim * yi, zi = pre * yi + pim * yr;
                        const uint32_t se = ch.stab[min(max((p + 1) >> 1, 0), kPaddedChips)];
                        const float sa = __uint_as_float(__byte_perm(se, 0x00800000u, 0x0644));   // chip k: the early tap
                        const float sl = __uint_as_float(__byte_perm(se, 0x00800000u, 0x1644));   // chip k + 1: the late tap
                        const float sp = (p & 1) ? sa : sl;                                     // p even: the prompt tap reads chip k + 1
                        aEr = fmaf(sa, zr, aEr); aEi = fmaf(sa, zi, aEi);
                        aPr = fmaf(sp, zr, aPr); aPi = fmaf(sp, zi, aPi);
                        aLr = fmaf(sl, zr, aLr); aLi = fmaf(sl, zi, aLi);
                        if (amb && B >= 0 && B < n && !(PM.debug & 2)) m_correct(ch, ctl, fix, slot, sb, B, p);
                    }
                    if (__shfl_sync(full, Bc, 31) >= j_hi) break;      // the next round would start behind the piece
                    x += dx;
                    p += kMSeg;
                }
                // ---- deliver the piece's six sums in fixed point (integer addition: order independent)
                __syncwarp();
                float v[6] = {aEr, aEi, aPr, aPi, aLr, aLi};
                if (lane == kMSeg) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) v[k] = fix[k];         // lane 31 owns no segment: it carries the corrections
                }
                int q[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) q[k] = __reduce_add_sync(full, __float2int_rn(v[k] * P.acc_scale));
                const bool bad = fix[6] != 0.f;
                __syncwarp();
                if (lane < 8) fix[lane] = 0.f;
                int qq = q[0];
#pragma unroll
                for (int k = 1; k < 6; ++k) qq = (lane == k) ? q[k] : qq;
                if (lane == 6) qq = bad ? 1 : 0;
                if (lane < 8) ch.part[e & 1][t - ctl.first_block][lane] = (lane < 7) ? qq : 0;
                __threadfence_block();
                __syncwarp();
                int last = 0;
                if (lane == 0) last = (atomicAdd(&ch.cnt[e & 1], 1) + 1 == ctl.n_blocks) ? 1 : 0;
                last = __shfl_sync(full, last, 0);
                const bool ends_here = a_rel + n <= s_hi;              // the epoch ends inside this block
                if (PROF) { const long long now = clock64(); c_piece += now - pt; pt = now; ++n_piece; }
                if (last) {
                    __threadfence_block();
                    m_close(PM, sh, ch, e, lane);
                    if (PROF) { const long long now = clock64(); c_close += now - pt; pt = now; ++n_close; }
                }
                if (ends_here) ep[c] = e + 1;
                if (a_rel + n >= s_hi) todo &= ~(1u << c);             // (else the successor's first piece is in this block too)
                progressed = true;
            }
            if (!progressed) {
                __nanosleep(spins++ < 4 ? 100 : 400);                  // every open channel waits for a loop closure
                if (PROF) { const long long now = clock64(); c_wait += now - pt; pt = now; }
            } else {
                spins = 0;
            }
        }
    }
    if (PROF && P.prof != nullptr && lane == 0 && warp < 4 && ch_lo + (warp % max(G, 1)) < PM.n_channels) {
        // diagnostics: warps 0 .. G-1 of the CTA leave their cycle counters in the rows of the CTA's channels
        if (warp < G) {
            long long* pc = P.prof + (long long)(ch_lo + warp) * 16;
            pc[0] = clock64() - pt0; pc[1] = c_prod; pc[2] = c_piece; pc[3] = c_wait; pc[4] = c_close;
            pc[5] = n_blk; pc[6] = n_piece; pc[7] = n_close;
        }
    }
    __syncthreads();
    if (tid < G) {
        MChan& ch = sh.chan[tid];
        if (ch.ch >= 0) {
            if (!ch.active) {
                // idle slot, finished earlier, or left to the general kernel: no epochs from this launch
                if (ch.status == kNeedGeneral) P.states[ch.ch].status = kNeedGeneral;
                P.nepochs[ch.ch] = ch.rec_base;
            } else {
                const int epochs = (ch.stop_epoch != 0x7fffffff) ? ch.stop_epoch : max(ch.pub, 0);    // epochs closed by this launch
                sydr_trk_state* gst = P.states + ch.ch;
                gst->cur = ch.sc.cur; gst->n_req = ch.sc.n_req; gst->epochs_done = ch.cfgs.epochs_done + epochs;
                gst->carrier_freq = ch.sk.carrier_freq; gst->code_freq = ch.sc.code_freq; gst->code_step = ch.sc.code_step;
                gst->rem_carrier = ch.sk.rem_carrier; gst->rem_code = ch.sc.rem_code;
                gst->nco_code = ch.sc.nco_code; gst->nco_code_err = ch.sc.nco_code_err;
                gst->nco_carrier = ch.sk.nco_carrier; gst->nco_carrier_err = ch.sk.nco_carrier_err;
                gst->status = ch.status;
                P.nepochs[ch.ch] = ch.rec_base + epochs;
            }
        }
    }
}

}  // namespace sydr

using namespace sydr;

namespace sydr {

int g_trkm_debug = 0;

// Launch the prefix-moment kernel: `group` consecutive channels per CTA (<= 0: as many CTAs as fit one wave of the SMs);
// with rec_channels > 0 channels [r*rec_channels, (r+1)*rec_channels) belong to recording r and no CTA spans two recordings
// (a channel that is not on its CTA's recording is left to the general kernel).  CTAs are independent: any grid size.
int launch_trkm(const TrkParams& P0, int n_channels, int rec_channels, int group, cudaStream_t s) {
    SYDR_REQUIRE(rec_channels >= 0, SYDR_ERR_ARG, "rec_channels must not be negative");
    int dev = 0, sms = 0;
    SYDR_CUDA_CHECK(cudaGetDevice(&dev));
    SYDR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int rc = (rec_channels > 0 && rec_channels < n_channels) ? rec_channels : n_channels;
    const int n_rec = (n_channels + rc - 1) / rc;
    if (group <= 0) {
        group = 1;
        while (group < kMMaxGroup && n_rec * ((rc + group - 1) / group) > sms) ++group;
    }
    SYDR_REQUIRE(group >= 1 && group <= kMMaxGroup, SYDR_ERR_ARG, "group must be 1..%d (got %d)", kMMaxGroup, group);
    TrkmParams PM;
    PM.t = P0;
    // fixed-point scale of a delivery: one warp, one block, <= 64 segments of <= 31 full-scale samples
    {
        const double bound = 64.0 * 31.0 * 46342.0 * 1.01;
        int k = (int)floor(log2(2147483647.0 / bound));
        if (k > 20) k = 20;
        PM.t.acc_scale = (float)ldexp(1.0, k);
        PM.t.acc_inv = ldexp(1.0, -k);
    }
    PM.n_channels = n_channels;
    PM.group = group;
    PM.rec_channels = (rec_channels > 0 && rec_channels < n_channels) ? rec_channels : 0;
    PM.groups_per_rec = (rc + group - 1) / group;
    PM.alpha_hc_max = 0.06;
    PM.debug = g_trkm_debug;
    const size_t smem = (size_t)kMWarps * kMSlotBytes;
    auto kern = (P0.prof != nullptr) ? trkm_kernel<true> : trkm_kernel<false>;
    SYDR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_rec * PM.groups_per_rec, kMThreads, smem, s>>>(PM);
    count_launch();
    SYDR_CUDA_CHECK(cudaGetLastError());
    return SYDR_OK;
}

}  // namespace sydr

extern "C" int sydr_trkm_debug(int flags) { sydr::g_trkm_debug = flags; return 0; }
// (kept for the ABI: the launch shape is fixed in this formulation)
extern "C" int sydr_trkm_shape(int, int) { return 0; }
