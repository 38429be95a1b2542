// K-TRKM: closed-loop E/P/L tracking in the prefix-moment formulation, throughput shape, for sm_100a.
//
// Reference semantics (file:line in /root/reference): EPL sydr/dsp/tracking.py:92-116, DLL_NNEML / PLL_costa
// tracking.py:120-142, BorreLoopFilter tracking.py:180-186, NCO update sydr/channel/channel_l1ca_borre.py:363-429.
//
// What the north star asks for -- "loads each int16 IQ tile once ... and correlates every channel against it" -- taken
// one step further: the per-sample work is done ONCE PER RECORDING and does not depend on any channel; the channels
// of a recording then pick, wherever they run, the few values they need out of its result.
//
//   producer warps (two in every CTA; together they serve the CTA's recording, block by block, by ticket)
//       turn the recording into exact integer prefix moments, 16 bytes per sample,
//           E0[j] = sum_{i<j} x_i            E1[j] = sum_{i<j} E0[i+1] = sum_{m<j} (j-m) x_m      (complex, int32, mod 2^32)
//       in blocks of 512 entries that start afresh every 480 samples (entry k of block t belongs to sample 480 t + k,
//       sums counted from sample 480 t): the blocks do not depend on each other, one warp makes one block on its own
//       (IDP.2A unpacks an int16 I or Q and accumulates it in one instruction, the second moment is one IADD per
//       sample, a shuffle scan joins the lanes) and stores it with 256-bit stores into a ring of blocks per recording
//       in global memory -- 1 MB per recording, resident in the 126 MB L2.  Wrap-around of the sums is harmless: only
//       differences over <= 31 samples are ever used, those fit int32, so they are exact.
//   correlating warps (four per channel, one channel per CTA)
//       Between two consecutive chip-boundary samples a <= j < b of a channel (half a chip, ~12.2 samples at
//       25 MS/s) the three code replicas are constant and the carrier advances by a few milliradians, so
//           sum_j x_j exp(i phi_j) = exp(i phi_c) [ S0 (1 - alpha^2 (L^2-1)/24) + i alpha/2 S1 ] + O(1e-6 S0),
//           S0 = E0[b]-E0[a],  W = E1[b]-E1[a] - L E0[a],  S1 = (L+1) S0 - 2 W,  L = b-a,  c = (a+b-1)/2,  alpha = -2 pi fc/fs.
//       One lane handles one boundary: one 16-byte gather from the L2 ring (both ends of a segment from the block of its
//       first sample: the 32 extra entries of a block are there for that), ~70 instructions, no per-sample work.
//       The boundaries are located exactly as in trk.cu (the reference's own FP64 expression
//       ceil(fl(fl(j step') + start)) decides every sample within 1e-9 of a lattice crossing), the six sums of an
//       epoch are reduced in fixed point (order independent), and the loops are closed by the same FP64 code as
//       K-TRK (trk_common.cuh), so the NCO trajectory arithmetic is the reference's, operation for operation.
//   Flow control: a block is announced by a release store of its number into the ring slot's flag; every correlating
//   warp publishes the first block it may still read; a producer overwrites slot t % R once all of them passed t - R.
//   The ring (61 440 samples) is longer than an epoch, so the channels of a recording run at their own epoch phase.
//   The CTAs of one recording depend on each other (shared producers), those of different recordings do not: the
//   launch is cooperative (all CTAs resident), larger jobs are cut into several launches by recording.
//
// Conditions (else the channel stops with status kNeedGeneral and the general kernel queued behind serves it, exactly
// like the LEAN instantiation of trk.cu): int16 IQ, spacings -0.5 / 0 / +0.5 chip around the prompt tap, half a chip
// between 1 and 30 samples, |alpha| x half chip <= 0.06 rad (|carrier| <= 19.5 kHz: error bound 4e-6 of the prompt
// magnitude at 45 dB-Hz), every code index inside the padded code, the channel on the recording
// its slot belongs to (iq_base, a multiple of 4 samples).
#include "trk_common.cuh"

namespace sydr {

constexpr int kMCWMax = 8;                      // correlating warps per channel, at most (even: a lane keeps its lattice parity)
constexpr int kMProdMax = 4;                    // producer warps per CTA, at most
constexpr int kMKS = 16;                        // samples per producer lane
constexpr int kMBlkEnt = 32 * kMKS;             // entries per block: 512
constexpr int kMBlkNew = kMBlkEnt - 32;         // samples a block advances by: 480 (a segment is at most 31 samples long)
constexpr int kMRing = 128;                     // blocks of a recording's ring (power of two): 1 MB, 61 440 samples
constexpr int kMLead = 16;                      // blocks made without looking at the consumers
constexpr int kMSeg = 31;                       // segments per correlating round (32 boundaries)
constexpr int kMMaxThreads = 32 * (kMCWMax + kMProdMax);
constexpr long long kMMaxSpan = 0x7fffffffLL - (1LL << 20);      // ring-relative sample indices are 32-bit
static_assert((kMRing & (kMRing - 1)) == 0, "ring slot by mask");

struct MRec {                 // one per recording, written by trkm_plan_kernel
    long long origin;         // recording-relative sample of ring index 0 (multiple of 16)
    long long valid_lo, valid_hi;   // samples readable from the recording's base pointer
    long long iq_base;
    long long limit;          // no epoch of this launch reaches beyond this recording-relative sample
    int n_blocks;             // blocks the producers may have to make
    int ok;                   // the recording has channels to track
    unsigned next_ticket;     // next block to make
    int pad;
};

struct MCtl {                 // per-epoch constants of one channel, published by its two leader warps
    double start[3], step[3]; // numpy linspace constants of the three taps (tracking.py:110-112)
    double inv_step;          // ~1/step' of the prompt tap
    double ca, cb;            // carrier phase in turns at epoch-relative sample j: ca*j + cb
    long long a;              // epoch start sample (recording-relative)
    float ah, g2;             // alpha/2 = pi*ca; alpha^2/24
    int n, p0, stop, car_stop;
};

struct MChan {                // shared-memory state of the CTA's channel
    uint32_t cb[kCodeWords];                    // padded code, one bit per chip
    uint16_t stab[kPaddedChips + 1];            // byte 0 / 1 of entry k: top byte of +-1.0f for chips k / k+1
    MCtl ctl;
    alignas(16) int part[2][kMCWMax][8];        // fixed-point warp totals of an epoch, [epoch & 1][warp][component]
    float fix[kMCWMax][8];                      // exact-evaluation corrections of ambiguous samples, per warp
    alignas(8) uint64_t bar_part[2];
    sydr_trk_state cfgs;
    CodeState sc;
    CarrierState sk;
    LoopConst K;
    int n_hist[2];
    int rec_base, status, ch, active;
    MRec rec;
};

struct TrkmParams {
    TrkParams t;
    int n_channels;
    int ch_first;             // first channel of this launch (a multiple of rec_channels)
    int rec_channels;         // channel slots per recording
    int cw_n;                 // correlating warps per channel
    int prod_n;               // producer warps per CTA
    int prog_stride;          // progress slots per recording (rec_channels * cw_n, rounded up)
    double alpha_hc_max;      // |alpha| * half chip limit of the expansion
    MRec* recs;               // [n_rec]
    int* progress;            // [n_rec][prog_stride]: first block a correlating warp may still read (INT_MAX = finished)
    uint4* ring;              // [n_rec][kMRing][kMBlkEnt]
    int debug;
};

__device__ __forceinline__ void m_named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint4 ldg_cg128(const uint4* p) {                 // L2 only: the ring is rewritten by other SMs
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// First look at an entry: an ordinary load that does not allocate in L1 (strong loads of 32 scattered addresses are served
// one after the other: 2 000 cycles per gather).  What it returns is checked against the lap tag like everything else; a stale
// line is indistinguishable from a block that is not there yet and takes the strong load of the waiting path.
__device__ __forceinline__ uint4 ldg_na128(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg256(uint4* p, int a0, int a1, int a2, int a3, int b0, int b1, int b2, int b3) {
    asm volatile("st.global.v8.u32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0),
                 "r"(b1), "r"(b2), "r"(b3)
                 : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void st_relaxed_gpu(int* p, int v) {
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- plan: one CTA per recording ------------------------------------------------------------------------------------
__global__ void trkm_plan_kernel(const TrkmParams PM) {
    const TrkParams& P = PM.t;
    const int r = blockIdx.x;
    const int c_lo = PM.ch_first + r * PM.rec_channels, c_hi = min(c_lo + PM.rec_channels, PM.ch_first + PM.n_channels);
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        long long base = 0, lo = 0x7fffffffffffffffLL, hi = 0;
        int have = 0;
        for (int c = c_lo; c < c_hi; ++c) {
            const sydr_trk_state& g = P.states[c];
            if (g.status != 0) continue;
            const long long b = P.has_iq_base ? P.iq_base : (long long)g.iq_base;
            if (!have) { base = b; have = 1; }
            if (b != base) continue;                               // left to the general kernel by its own CTA
            lo = min(lo, (long long)g.cur);
            const long long len = P.iq_len > 0 ? P.iq_len : (long long)g.iq_len;
            hi = max(hi, min(len, P.iq_alloc - base));
        }
        MRec rec;
        rec.iq_base = base;
        rec.ok = (have && (base & 3) == 0 && hi > lo && lo >= 0) ? 1 : 0;
        rec.origin = rec.ok ? (lo & ~15LL) : 0;
        rec.valid_lo = max(0LL, -base);
        rec.valid_hi = P.iq_alloc - base;
        long long span = hi - rec.origin;
        if (span > kMMaxSpan) span = kMMaxSpan;                    // the general kernel continues behind
        rec.limit = rec.origin + span;
        rec.n_blocks = rec.ok ? (int)(span / kMBlkNew) + 1 : 0;    // entry indices 0 .. span
        rec.next_ticket = 0;
        rec.pad = 0;
        PM.recs[r] = rec;
        s_ok = rec.ok;
    }
    __syncthreads();
    // a slot of a channel that exists starts at block 0 (its CTA withdraws it if it does not take part), the others never hold the ring
    for (int i = threadIdx.x; i < PM.prog_stride; i += blockDim.x)
        PM.progress[r * PM.prog_stride + i] = (s_ok && c_lo + i / PM.cw_n < c_hi) ? 0 : 0x7fffffff;
}

// ---- producer warp --------------------------------------------------------------------------------------------------
// Block t: entries k = 0 .. 511 for the samples origin + 480 t + k, sums counted from the block's first sample.  Lane l owns
// the kMKS samples from position 16 l and writes the entries (exclusive prefixes) of exactly those positions.
template <bool PROF>
__device__ __forceinline__ void m_produce(const TrkmParams& PM, const MRec& rec, MRec* grec, int r, int lane, long long* pc) {
    const unsigned full = 0xffffffffu;
    const int n_blocks = rec.n_blocks;
    const uint32_t* rec_ptr = reinterpret_cast<const uint32_t*>(PM.t.iq + rec.iq_base * 4);
    uint4* gring = PM.ring + (size_t)r * kMRing * kMBlkEnt;
    const int* prog = PM.progress + r * PM.prog_stride;
    long long tp0 = 0, tp1 = 0, c_wait = 0, c_ld = 0, c_st = 0, n_blk = 0;
    if (PROF) tp0 = clock64();
    while (true) {
        if (PROF) tp1 = clock64();
        unsigned tk = 0;
        if (lane == 0) tk = atomicAdd(&grec->next_ticket, 1u);
        const int t = (int)__shfl_sync(full, tk, 0);
        if (t >= n_blocks) break;
        if (t >= kMLead) {
            // room in the ring?  block t overwrites block t - kMRing: every correlating warp must have moved past it
            bool done = false;
            while (true) {
                int mn = 0x7fffffff;
                for (int i = lane; i < PM.prog_stride; i += 32) mn = min(mn, ld_relaxed_gpu(prog + i));
                mn = __reduce_min_sync(full, mn);
                if (mn == 0x7fffffff) { done = true; break; }
                if (mn > t - kMRing) break;
                __nanosleep(1500);                                     // the ring is full: the producers are far ahead
            }
            if (done) break;                                            // every channel of the recording has finished
        }
        if (PROF) { const long long now = clock64(); c_wait += now - tp1; tp1 = now; }
        const long long s0 = rec.origin + (long long)t * kMBlkNew + lane * kMKS;       // this lane's first sample
        const long long b0 = s0 - lane * kMKS;
        uint32_t w[kMKS];
        if (b0 >= rec.valid_lo && b0 + kMBlkEnt <= rec.valid_hi) {
            const uint4* src = reinterpret_cast<const uint4*>(rec_ptr + s0);
#pragma unroll
            for (int v = 0; v < kMKS / 4; ++v) {
                const uint4 q = __ldg(src + v);
                w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
            }
        } else {                                                       // edge of the allocation: guarded loads, zeros outside
#pragma unroll
            for (int v = 0; v < kMKS; ++v) {
                const long long s = s0 + v;
                w[v] = (s >= rec.valid_lo && s < rec.valid_hi) ? rec_ptr[s] : 0u;
            }
        }
        // pass 1: this lane's totals.  dp2a: I = low int16 of the word, Q = high int16.  t1 = sum_v (16 - v) x_v.
        int t0r = 0, t0i = 0, t1r = 0, t1i = 0;
#pragma unroll
        for (int v = 0; v < kMKS; ++v) {
            t0r = __dp2a_lo((int)w[v], 0x0001, t0r);
            t0i = __dp2a_lo((int)w[v], 0x0100, t0i);
            t1r += t0r;
            t1i += t0i;
        }
        // exclusive scan over the warp.  With u_m = t1_m - 16 (m + 1) t0_m the second moment in front of lane l is
        // sum_{m<l} u_m + 16 l sum_{m<l} t0_m: two plain sum scans.
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
        if (PROF) { const long long now = clock64(); c_ld += now - tp1; tp1 = now; }
        // pass 2: the entries of this lane's samples (exclusive prefixes), two per 256-bit store.  Every 32-bit word carries
        // 26 bits of its sum (all that the differences need) and, on top, the lap of the ring the block belongs to: a reader
        // that finds the lap it expects in all four words of an entry holds that entry of that block -- no flag, no fence.
        const unsigned tag = ((unsigned)(t / kMRing) & 63u) << 26;
        auto word = [tag](int v) { return (int)(((unsigned)v & 0x03ffffffu) | tag); };
        uint4* dst = gring + (size_t)(t & (kMRing - 1)) * kMBlkEnt + lane * kMKS;
#pragma unroll
        for (int v = 0; v < kMKS; v += 2) {
            const int e0r = a0r, e0i = a0i, e1r = a1r, e1i = a1i;
            a0r = __dp2a_lo((int)w[v], 0x0001, a0r);
            a0i = __dp2a_lo((int)w[v], 0x0100, a0i);
            a1r += a0r;
            a1i += a0i;
            stg256(dst + v, word(e0r), word(e0i), word(e1r), word(e1i), word(a0r), word(a0i), word(a1r), word(a1i));
            a0r = __dp2a_lo((int)w[v + 1], 0x0001, a0r);
            a0i = __dp2a_lo((int)w[v + 1], 0x0100, a0i);
            a1r += a0r;
            a1i += a0i;
        }
        if (PROF) { c_st += clock64() - tp1; ++n_blk; }
    }
    if (PROF && pc != nullptr && lane == 0) {
        pc[8] = clock64() - tp0; pc[9] = c_wait; pc[10] = n_blk; pc[11] = c_ld; pc[12] = c_st;
    }
}

// ---- correlating warps ----------------------------------------------------------------------------------------------
__device__ __forceinline__ int sext26(unsigned v) { return (int)(v << 6) >> 6; }
__device__ __forceinline__ bool m_valid(const uint4& e, unsigned tag) {          // all four words are of the expected lap
    return ((((e.x ^ tag) | (e.y ^ tag)) | ((e.z ^ tag) | (e.w ^ tag))) >> 26) == 0u;
}
__device__ __forceinline__ uint4 m_entry(const uint4* ent, unsigned tag) {        // wait for an entry (rare: only the leading channel ever waits)
    uint4 e = ldg_cg128(ent);
    while (!m_valid(e, tag)) {
        __nanosleep(200);
        e = ldg_cg128(ent);
    }
    return e;
}
// Exact treatment of an ambiguous boundary (the crossing is within 1e-9 sample of an integer J): sample J was given to
// the segment behind the boundary; its three code indices are re-evaluated with the reference expression and the
// difference, times the wiped-off sample, goes to the warp's correction sums.
static __device__ __noinline__ void m_correct(MChan& ch, int cw, const uint4* ent, unsigned tag, int J, int p) {
    const MCtl& c = ch.ctl;
    const uint4 e0 = m_entry(ent, tag), e1 = m_entry(ent + 1, tag);
    const float xr = (float)sext26(e1.x - e0.x), xi = (float)sext26(e1.y - e0.y);
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
            atomicAdd(&ch.fix[cw][2 * s], d * zr);
            atomicAdd(&ch.fix[cw][2 * s + 1], d * zi);
        }
    }
    if (err) atomicAdd(&ch.fix[cw][6], 1.0f);
}

template <int kThreads, int kMinCtas, bool PROF>
__global__ void __launch_bounds__(kThreads, kMinCtas) trkm_kernel(const TrkmParams PM) {
    __shared__ __align__(16) MChan ch;
    const TrkParams& P = PM.t;
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int CW = PM.cw_n;
    const int idx = PM.ch_first + blockIdx.x;                          // one channel per CTA
    const int r = blockIdx.x / PM.rec_channels, c_in_rec = blockIdx.x % PM.rec_channels;
    int* my_prog = PM.progress + r * PM.prog_stride + c_in_rec * CW;

    // ---- set-up: the recording's plan, the channel's state and code tables
    if (tid == 0) {
        ch.rec = PM.recs[r];
        ch.ch = idx;
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
        // the recording the plan names, spacings on the half-chip lattice at -1 / 0 / +1 half chips around the prompt tap
        int q[3];
        if (ch.active && !(ch.rec.ok && g.iq_base == ch.rec.iq_base && g.cur >= ch.rec.origin && seg_tap_offsets(g.spacing, q) &&
                           q[0] == -1 && q[2] == 1)) {
            ch.status = kNeedGeneral;
            ch.active = 0;
        }
        mbar_init(&ch.bar_part[0], CW);
        mbar_init(&ch.bar_part[1], CW);
        fence_mbar_init();
    }
    __syncthreads();
    if (!ch.active) {
        // idle slot, finished earlier, or left to the general kernel: no epochs from this launch; the ring does not wait for it
        if (tid < CW) st_relaxed_gpu(my_prog + tid, 0x7fffffff);
        if (tid == 0) {
            if (ch.status == kNeedGeneral) P.states[idx].status = kNeedGeneral;
            P.nepochs[idx] = ch.rec_base;
        }
    } else {
        if (tid < kCodeWords) ch.cb[tid] = P.code_bits[(ch.cfgs.prn - 1) * kCodeWords + tid];
    }
    __syncthreads();
    if (ch.active) {
        for (int k = tid; k <= kPaddedChips; k += blockDim.x) {
            const int k0 = min(k, kPaddedChips - 1), k1 = min(k + 1, kPaddedChips - 1);
            const uint32_t b0 = (ch.cb[k0 >> 5] >> (k0 & 31)) & 1u, b1 = (ch.cb[k1 >> 5] >> (k1 & 31)) & 1u;
            ch.stab[k] = (uint16_t)((b0 ? 0x3Fu : 0xBFu) | ((b1 ? 0x3Fu : 0xBFu) << 8));
        }
    }
    __syncthreads();

    if (warp >= CW) {
        // ================================================================ producer warps (also of a CTA whose channel idles)
        if (ch.rec.ok) m_produce<PROF>(PM, ch.rec, PM.recs + r, r, lane, (PROF && P.prof != nullptr && warp == CW) ? P.prof + (long long)idx * 16 : nullptr);
        return;
    }
    if (!ch.active) return;
    // ==================================================================== correlating warps
    const int cw = warp;
    const uint4* gring = PM.ring + (size_t)r * kMRing * kMBlkEnt;
    sydr_trk_epoch* out_row = P.out + (long long)ch.ch * P.max_epochs + ch.rec_base;
    CodeState sc = ch.sc;
    CarrierState sk = ch.sk;
    int status = ch.cfgs.status;
    int epoch = 0;
    const int epoch_cap = P.max_epochs - ch.rec_base;
    const long long rec_alloc = P.iq_alloc - ch.cfgs.iq_base;
    const long long iq_len_reg = min(min((long long)ch.cfgs.iq_len, rec_alloc), ch.rec.limit);
    const long long origin = ch.rec.origin;
    int released = 0;
    long long tc0 = 0, c_flag = 0, c_e0 = 0, c_bar = 0, n_sleep = 0, n_round = 0, c_corr = 0, c_math = 0, c_amb = 0, n_amb = 0, tr = 0;
    const int pw = (PROF && P.prof != nullptr) ? 2 : -1;      // the warp whose time is recorded (no loop closure of its own)
    if (PROF) tc0 = clock64();
    while (true) {
        long long te0 = 0;
        if (PROF) te0 = clock64();
        // ---- publish the constants of epoch `epoch`
        if (cw == 0) {
            if (sc.n_req <= 0 || sc.n_req > 0x3fffffff) status = SYDR_ERR_STATE;
            bool stop = (status != 0) || (epoch >= epoch_cap) || (sc.cur + sc.n_req > iq_len_reg);
            double t_start = 0.0, t_step = 0.0, t_stop = 0.0;
            int p0 = 0;
            if (!stop) {
                const double dn = i2d(sc.n_req);
                sc.inv_n = newton_rcp(dn, newton_rcp(dn, sc.inv_n));
                t_start = dadd(sc.rem_code, ch.cfgs.spacing[min(lane, 2)]);             // tracking.py:110
                t_stop = dadd(dmul(sc.code_step, dn), t_start);
                t_step = ddiv_by(dsub(t_stop, t_start), dn, sc.inv_n);                  // numpy linspace step
                const bool in = (t_start > -0.999) && (t_stop < (double)(kPaddedChips - 1) - 0.001);
                const double hc = 0.5 * sc.inv_step;                                     // half a chip, in samples
                const bool ok = __all_sync(full, in) && hc >= 1.0 && hc <= 30.0;
                const double p_start = dadd(sc.rem_code, ch.cfgs.spacing[1]);
                p0 = ceil_to_int(2.0 * p_start) - 1;                                     // lattice point in front of sample 0
                if (!ok) { stop = true; status = kNeedGeneral; }
            }
            if (!stop) {
                if (lane < 3) {
                    ch.ctl.start[lane] = t_start;
                    ch.ctl.step[lane] = t_step;
                    if (lane == 1) { ch.ctl.inv_step = sc.inv_step; ch.ctl.p0 = p0; }
                } else if (lane == 3) {
                    ch.ctl.n = sc.n_req;
                    ch.ctl.a = sc.cur;
                    ch.n_hist[epoch & 1] = sc.n_req;
                }
            }
            if (lane == 5) { ch.ctl.stop = stop ? 1 : 0; ch.status = status; }
        } else if (cw == 1) {
            double ca, cbb;
            float w[4][2];
            carrier_const(sk.carrier_freq, sk.rem_carrier, ch.K.inv_fs, ca, cbb, w);
            if (lane == 0) {
                ch.ctl.ca = ca;
                ch.ctl.cb = cbb;
                ch.ctl.ah = (float)(kPi * ca);
                ch.ctl.g2 = (float)((2.0 * kPi * ca) * (2.0 * kPi * ca) * (1.0 / 24.0));
                // the expansion needs |alpha| x half chip small (nominal half chip: the code rate moves by < 1e-5)
                ch.ctl.car_stop = (fabs(2.0 * kPi * ca) * (0.5 * P.fs / kCodeFreq) > PM.alpha_hc_max) ? 1 : 0;
            }
        }
        m_named_barrier(1, CW * 32);                   // (A) the constants are visible
        if (PROF) { const long long now = clock64(); c_bar += now - te0; te0 = now; }
        if (ch.ctl.car_stop && !ch.ctl.stop) {         // leave the channel to the general kernel
            if (cw == 0 && lane == 0) ch.status = kNeedGeneral;
            break;
        }
        if (ch.ctl.stop) break;

        // ---- correlate: rounds of 31 segments, round r of the epoch belongs to warp r mod CW
        const MCtl& ctl = ch.ctl;
        const int n = ctl.n;
        const int a_rel = (int)(ctl.a - origin);                // ring-relative index of the epoch's first sample
        const double inv_step = ctl.inv_step;
        const float ah = ctl.ah;
        const float gtab = 1.0f - ctl.g2 * (float)(lane * lane - 1);          // lane L: 1 - alpha^2 (L^2-1)/24
        const double ca_half = 0.5 * ctl.ca, cbt = ctl.cb;
        int p = ctl.p0 + kMSeg * cw + lane;                    // this lane's front boundary (lattice point)
        double x = seg_crossing(dmul(0.5, i2d(p)), ctl.start[1], inv_step);
        const double dx = (double)(kMSeg * CW) * 0.5 * inv_step;
        float aAr = 0.f, aAi = 0.f, aBr = 0.f, aBi = 0.f;
        if (lane < 8) ch.fix[cw][lane] = 0.f;
        __syncwarp();
        // A round = 32 boundaries (31 segments).  The gathers of round k + 1 are in flight while round k is worked on.
        bool amb_n;
        int B_n = seg_first_sample(x, amb_n);
        int Bc_n = min(max(B_n, 0), n);
        bool more = __shfl_sync(full, Bc_n, 0) < n;            // else the round starts behind the epoch
        const uint4* ent_n = gring;
        int blk_n = 0, L_n = 0;
        bool far_n = false;
        uint4 e0_n = make_uint4(0u, 0u, 0u, 0u), e1_n = e0_n;
        auto gather = [&]() {                                  // the entries of the coming round's boundaries
            const int jr = a_rel + Bc_n;                       // ring-relative sample index of this lane's boundary
            blk_n = (int)((unsigned)jr / (unsigned)kMBlkNew);
            ent_n = gring + (size_t)(blk_n & (kMRing - 1)) * kMBlkEnt + (jr - blk_n * kMBlkNew);
            e0_n = ldg_na128(ent_n);
            L_n = __shfl_down_sync(full, Bc_n, 1) - Bc_n;      // 0 for lane 31 and for clipped segments
            // the far end of a segment that reaches into the next block: from the extra entries of its own block
            far_n = (__shfl_down_sync(full, blk_n, 1) != blk_n) && lane < kMSeg;
            if (far_n) e1_n = ldg_na128(ent_n + L_n);
        };
        if (more) gather();
        while (more) {
            if (PROF) tr = clock64();
            const int B = B_n, Bc = Bc_n, blk = blk_n, L = L_n;
            const bool amb = amb_n, far = far_n;
            const uint4* ent = ent_n;
            uint4 e0 = e0_n, e1 = e1_n;
            // what this warp may still read starts with this round: the blocks in front of it go back to the producers
            const int blk_first = __shfl_sync(full, blk, 0);
            if (blk_first > released) {
                released = blk_first;
                if (lane == 0) st_relaxed_gpu(my_prog + cw, released);
            }
            // the coming round
            x += dx;
            B_n = seg_first_sample(x, amb_n);
            Bc_n = min(max(B_n, 0), n);
            more = __shfl_sync(full, Bc_n, 0) < n;
            if (more) gather();
            // this round's entries must be of the lap this warp expects (else the block is not there yet: the leading channel waits)
            const unsigned tag = ((unsigned)(blk / kMRing) & 63u) << 26;
            long long tw = 0;
            if (PROF) { tw = clock64(); ++n_round; c_flag += tw - tr; }
            {
                unsigned bad = __ballot_sync(full, !(m_valid(e0, tag) && (!far || m_valid(e1, tag))));
                while (bad) {
                    // the block is not there yet (only a leading channel ever waits): watch ONE entry, the furthest one, at leisure,
                    // then look at all of them again
                    if (PROF) ++n_sleep;
                    const int h = 31 - __clz(bad);
                    if (lane == h) {
                        const uint4* pe = m_valid(e0, tag) ? ent + L : ent;
                        do {
                            __nanosleep(1000);
                        } while (!m_valid(ldg_cg128(pe), tag));
                    }
                    __syncwarp();
                    if (bad >> lane & 1u) {
                        e0 = ldg_cg128(ent);
                        if (far) e1 = ldg_cg128(ent + L);
                    }
                    bad = __ballot_sync(full, !(m_valid(e0, tag) && (!far || m_valid(e1, tag))));
                }
            }
            if (PROF) { tr = clock64(); c_e0 += tr - tw; }
            {
                const unsigned nx = __shfl_down_sync(full, e0.x, 1), ny = __shfl_down_sync(full, e0.y, 1);
                const unsigned nz = __shfl_down_sync(full, e0.z, 1), nw = __shfl_down_sync(full, e0.w, 1);
                if (!far) e1 = make_uint4(nx, ny, nz, nw);
            }
            // S0; W = sum (b - m) x_m = E1[b] - E1[a] - L E0[a];  S1 = 2 sum (m - c) x_m = (L + 1) S0 - 2 W   (modulo 2^26, exact)
            const unsigned u0r = e1.x - e0.x, u0i = e1.y - e0.y;
            const unsigned uwr = (e1.z - e0.z) - (unsigned)L * e0.x, uwi = (e1.w - e0.w) - (unsigned)L * e0.y;
            const int d0r = sext26(u0r), d0i = sext26(u0i);
            const int d1r = sext26((unsigned)(L + 1) * u0r - 2u * uwr), d1i = sext26((unsigned)(L + 1) * u0i - 2u * uwi);
            const float s0r = (float)d0r, s0i = (float)d0i, s1r = (float)d1r, s1i = (float)d1i;
            const float g = __shfl_sync(full, gtab, L & 31);
            const float yr = fmaf(g, s0r, -(ah * s1i)), yi = fmaf(g, s0i, ah * s1r);
            // carrier phasor at the segment centre c = (2 Bc + L - 1) / 2 (tracking.py:102), FP64 turns
            double turns = fma(ca_half, i2d(2 * Bc + L - 1), cbt);
            turns -= drint(turns);
            float pre, pim;
            __sincosf((float)turns * 6.283185307179586f, &pim, &pre);
            const float zr = pre * yr - pim * yi, zi = pre * yi + pim * yr;
            const uint32_t se = ch.stab[min(max((p + 1) >> 1, 0), kPaddedChips)];
            const float sa = __uint_as_float(__byte_perm(se, 0x00800000u, 0x0644));
            const float sb = __uint_as_float(__byte_perm(se, 0x00800000u, 0x1644));
            aAr = fmaf(sa, zr, aAr); aAi = fmaf(sa, zi, aAi);
            aBr = fmaf(sb, zr, aBr); aBi = fmaf(sb, zi, aBi);
            if (PROF) { const long long now = clock64(); c_math += now - tr + (long long)(aBr == 12345.f); tr = now; }
            if (__any_sync(full, amb && lane < kMSeg && B >= 0 && B < n)) {
                if (!(PM.debug & 2) && amb && lane < kMSeg && B >= 0 && B < n) m_correct(ch, cw, ent, tag, B, p);
                __syncwarp();
                if (PROF) ++n_amb;
            }
            if (PROF) c_amb += clock64() - tr;
            p += kMSeg * CW;
        }
        if (PROF) { const long long now = clock64(); c_corr += now - te0; te0 = now; }
        // ---- this warp's six sums in fixed point (integer addition: order independent)
        const bool even = ((ctl.p0 + kMSeg * cw + lane) & 1) == 0;        // p even: the prompt tap reads chip k + 1
        __syncwarp();
        float v[6] = {aAr, aAi, even ? aBr : aAr, even ? aBi : aAi, aBr, aBi};
        if (lane == kMSeg) {
#pragma unroll
            for (int k = 0; k < 6; ++k) v[k] = ch.fix[cw][k];           // lane 31 owns no segment: it carries the corrections
        }
        int q[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) q[k] = __reduce_add_sync(full, __float2int_rn(v[k] * P.acc_scale));
        const int slot_e = epoch & 1;
        if (lane == 0) {
            *reinterpret_cast<int4*>(&ch.part[slot_e][cw][0]) = make_int4(q[0], q[1], q[2], q[3]);
            *reinterpret_cast<int4*>(&ch.part[slot_e][cw][4]) = make_int4(q[4], q[5], ch.fix[cw][6] != 0.f ? 1 : 0, 0);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ch.bar_part[slot_e]);
        const int e = epoch;
        ++epoch;
        // ---- close the loops of epoch e (warp 0: code, warp 1: carrier)
        if (cw < 2) {
            CodePre cpre = {0.0, 0.0};
            double rc_next = 0.0;
            if (cw == 0) cpre = code_pre(sc);
            else rc_next = carrier_pre(ch, sk, ch.n_hist[e & 1]);
            mbar_wait(&ch.bar_part[slot_e], (e >> 1) & 1);
            long long tot = 0;
            for (int w2 = 0; w2 < CW; ++w2) tot += ch.part[slot_e][w2][lane & 7];
            const double ck = (double)tot * P.acc_inv;
            sydr_trk_epoch* rec = out_row + e;
            if (cw == 0) code_close(ch, sc, status, ck, rec, lane, cpre);
            else carrier_close(ch, sk, ck, rc_next, rec, lane);
        }
    }
    if (PROF && cw == pw && lane == 0) {
        long long* pc = P.prof + (long long)idx * 16;
        pc[0] = clock64() - tc0; pc[1] = c_flag; pc[2] = c_e0; pc[3] = c_bar; pc[4] = c_corr; pc[5] = n_sleep; pc[6] = n_round; pc[7] = epoch; pc[13] = c_math; pc[14] = c_amb; pc[15] = n_amb;
    }
    if (cw == 0 && lane == 0) ch.sc = sc;
    if (cw == 1 && lane == 0) ch.sk = sk;
    if (lane == 0) st_relaxed_gpu(my_prog + cw, 0x7fffffff);
    m_named_barrier(1, CW * 32);
    if (cw == 0 && lane == 0) {
        sydr_trk_state* gst = P.states + ch.ch;
        gst->cur = ch.sc.cur; gst->n_req = ch.sc.n_req; gst->epochs_done = ch.cfgs.epochs_done + epoch;
        gst->carrier_freq = ch.sk.carrier_freq; gst->code_freq = ch.sc.code_freq; gst->code_step = ch.sc.code_step;
        gst->rem_carrier = ch.sk.rem_carrier; gst->rem_code = ch.sc.rem_code;
        gst->nco_code = ch.sc.nco_code; gst->nco_code_err = ch.sc.nco_code_err;
        gst->nco_carrier = ch.sk.nco_carrier; gst->nco_carrier_err = ch.sk.nco_carrier_err;
        gst->status = ch.status;
        P.nepochs[ch.ch] = ch.rec_base + epoch;
    }
}

}  // namespace sydr

using namespace sydr;

namespace sydr {

int g_trkm_debug = 0;
int g_trkm_cw = 4, g_trkm_prod = 2;

namespace {
struct TrkmWorkspace { void* p = nullptr; size_t bytes = 0; };
TrkmWorkspace g_ws[16];
}

// Launch the prefix-moment kernel: one CTA per channel; channels [r*rec_channels, (r+1)*rec_channels) belong to recording r
// (rec_channels <= 0: all the channels are on one recording; a channel that is not on its slot's recording is left to the
// general kernel).  The CTAs of a recording share its producers, so a launch must be resident as a whole (cooperative
// launch); a job larger than the device holds at once is cut into several launches by recording.
// Returns SYDR_ERR_UNSUPPORTED (nothing launched) when not even one recording fits.
int launch_trkm(const TrkParams& P, int n_channels, int rec_channels, int cw, cudaStream_t s) {
    int dev = 0, sms = 0;
    SYDR_CUDA_CHECK(cudaGetDevice(&dev));
    SYDR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SYDR_REQUIRE(dev >= 0 && dev < 16, SYDR_ERR_UNSUPPORTED, "device index %d", dev);
    const int rc = (rec_channels > 0 && rec_channels < n_channels) ? rec_channels : n_channels;
    const int n_rec = (n_channels + rc - 1) / rc;
    SYDR_REQUIRE(cw == 0 || (cw >= 2 && cw <= kMCWMax && (cw & 1) == 0), SYDR_ERR_ARG, "correlating warps per channel must be 2, 4, 6 or 8 (got %d)", cw);
    const int cw_n = cw > 0 ? cw : g_trkm_cw, prod_n = g_trkm_prod;
    const int threads = 32 * (cw_n + prod_n);
    // <= 6 warps: three CTAs per SM (384 channels in one launch on 148 SMs); more: whatever the registers allow
    auto kern = threads <= 192 ? trkm_kernel<192, 3, false> : trkm_kernel<kMMaxThreads, 1, false>;
    if (P.prof != nullptr) kern = threads <= 192 ? trkm_kernel<192, 3, true> : trkm_kernel<kMMaxThreads, 1, true>;
    int per_sm = 0;
    SYDR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, 0));
    const int cap = per_sm * sms;
    if (rc > cap) {
        set_error("prefix-moment kernel: the %d channels of one recording exceed the %d CTAs the device holds at once", rc, cap);
        return SYDR_ERR_UNSUPPORTED;
    }
    const int rec_per_launch = cap / rc;
    const int prog_stride = (rc * cw_n + 31) & ~31;
    // workspace: plans, flags, progress slots, rings (L2 resident)
    const size_t rec_bytes = ((size_t)n_rec * sizeof(MRec) + 255) & ~(size_t)255;
    const size_t prog_bytes = ((size_t)n_rec * prog_stride * sizeof(int) + 255) & ~(size_t)255;
    const size_t ring_bytes = (size_t)n_rec * kMRing * kMBlkEnt * sizeof(uint4);
    const size_t need = rec_bytes + prog_bytes + ring_bytes;
    TrkmWorkspace& ws = g_ws[dev];
    if (ws.bytes < need) {
        // (a launch that still uses the old workspace is ordered in front of the free by cudaFree's implicit synchronisation)
        if (ws.p) SYDR_CUDA_CHECK(cudaFree(ws.p));
        ws.p = nullptr;
        ws.bytes = 0;
        SYDR_CUDA_CHECK(cudaMalloc(&ws.p, need));
        ws.bytes = need;
    }
    uint8_t* base = reinterpret_cast<uint8_t*>(ws.p);
    for (int r0 = 0; r0 < n_rec; r0 += rec_per_launch) {
        const int nr = (n_rec - r0 < rec_per_launch) ? n_rec - r0 : rec_per_launch;
        TrkmParams PM;
        PM.t = P;
        PM.ch_first = r0 * rc;
        PM.n_channels = (n_channels - PM.ch_first < nr * rc) ? n_channels - PM.ch_first : nr * rc;
        PM.rec_channels = rc;
        PM.cw_n = cw_n;
        PM.prod_n = prod_n;
        PM.prog_stride = prog_stride;
        PM.alpha_hc_max = 0.06;
        PM.recs = reinterpret_cast<MRec*>(base) + r0;
        PM.progress = reinterpret_cast<int*>(base + rec_bytes) + (size_t)r0 * prog_stride;
        PM.ring = reinterpret_cast<uint4*>(base + rec_bytes + prog_bytes) + (size_t)r0 * kMRing * kMBlkEnt;
        PM.debug = g_trkm_debug;
        // lap 63 in every word: nothing a reader of lap 0 takes for an entry
        SYDR_CUDA_CHECK(cudaMemsetAsync(PM.ring, 0xff, (size_t)nr * kMRing * kMBlkEnt * sizeof(uint4), s));
        trkm_plan_kernel<<<nr, 128, 0, s>>>(PM);
        count_launch();
        cudaLaunchConfig_t lc = {};
        lc.gridDim = dim3((unsigned)PM.n_channels);
        lc.blockDim = dim3((unsigned)threads);
        lc.dynamicSmemBytes = 0;
        lc.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeCooperative;
        at[0].val.cooperative = 1;
        lc.attrs = at;
        lc.numAttrs = 1;
        SYDR_CUDA_CHECK(cudaLaunchKernelEx(&lc, kern, PM));
        count_launch();
    }
    SYDR_CUDA_CHECK(cudaGetLastError());
    return SYDR_OK;
}

}  // namespace sydr

extern "C" int sydr_trkm_debug(int flags) { sydr::g_trkm_debug = flags; return 0; }
// Diagnostics / tuning: correlating warps per channel (2, 4, 6, 8) and producer warps per CTA (1 .. 4) of the next launches.
extern "C" int sydr_trkm_shape(int cw, int prod) {
    if (cw < 2 || cw > sydr::kMCWMax || (cw & 1) || prod < 1 || prod > sydr::kMProdMax) return SYDR_ERR_ARG;
    sydr::g_trkm_cw = cw;
    sydr::g_trkm_prod = prod;
    return 0;
}
