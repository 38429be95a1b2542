// K-TRKM: closed-loop E/P/L tracking in the prefix-moment formulation, throughput shape, for sm_100a.
//
// Reference semantics (file:line in /root/reference): EPL sydr/dsp/tracking.py:92-116, DLL_NNEML / PLL_costa
// tracking.py:120-142, BorreLoopFilter tracking.py:180-186, NCO update sydr/channel/channel_l1ca_borre.py:363-429.
//
// What the north star asks for -- "loads each int16 IQ tile once into shared memory and correlates every channel
// against it" -- taken one step further: the per-sample work is done ONCE PER RECORDING, not once per channel.
//
//   producer warps (per CTA, shared by its channels)
//       stage the recording tile by tile with TMA bulk copies (cp.async.bulk + mbarrier) and turn it into a ring of
//       exact integer prefix moments in shared memory, 16 bytes per sample:
//           P0[j] = sum_{i<j} x_i              P1[j] = sum_{i<j} 2 i x_i        (complex, int32, modulo 2^32)
//       (IDP.2A: one instruction unpacks an int16 I or Q and accumulates it; warp scan by shuffles; the carry between
//       tiles lives in registers.)  Wrap-around is harmless: only differences over <= 31 samples are ever used, and
//       those fit int32, so they are exact.
//   consumer warps (four per channel)
//       between two consecutive chip-boundary samples a <= j < b of a channel (half a chip, ~12.2 samples at 25 MS/s)
//       the three code replicas are constant and the carrier advances by a few milliradians, so
//           sum_j x_j exp(i phi_j) = exp(i phi_c) [ S0 (1 - alpha^2 (L^2-1)/24) + i alpha/2 S1 ] + O(1e-6 S0),
//           S0 = P0[b]-P0[a],   S1 = (P1[b]-P1[a]) - (a+b-1) S0,   L = b-a,   c = (a+b-1)/2,   alpha = -2 pi fc/fs.
//       One lane handles one boundary: two 16-byte reads of the ring, ~60 instructions, no per-sample work at all.
//       The boundaries are located exactly as in trk.cu (the reference's own FP64 expression
//       ceil(fl(fl(j step') + start)) decides every sample within 1e-9 of a lattice crossing), the six sums of an epoch
//       are reduced in fixed point (order independent), and the loops are closed by the same FP64 code as K-TRK
//       (trk_common.cuh), so the NCO trajectory arithmetic is the reference's, operation for operation.
//
// Conditions (else the channel stops with status kNeedGeneral and the general kernel queued behind serves it, exactly
// like the LEAN instantiation of trk.cu): int16 IQ, spacings -0.5 / 0 / +0.5 chip around the prompt tap, half a chip
// between 1 and 30 samples, |alpha| x half chip <= 0.06 rad (|carrier| <= 19.5 kHz: error bound 4e-6 of the prompt
// magnitude at 45 dB-Hz), every code index inside the padded code.
#include "trk_common.cuh"

namespace sydr {

constexpr int kMCW = 4;                         // consumer warps per channel (even: a lane keeps its lattice parity for an epoch)
constexpr int kMPW = 4;                         // producer warps per CTA
constexpr int kMMaxGroup = 4;                   // channels per CTA
constexpr int kMKS = 16;                        // samples per producer lane and super-tile
constexpr int kMTile = 32 * kMKS;               // samples per producer warp and super-tile
constexpr int kMSuper = kMPW * kMTile;          // samples per super-tile (one pass of the producer group): 2048
constexpr int kMRawBufs = 4;                    // TMA staging buffers (prefetch distance 3 super-tiles)
constexpr int kMSeg = 31;                       // segments per consumer round (32 boundaries)
constexpr int kMMaxThreads = 32 * (kMPW + kMMaxGroup * kMCW);

struct MCtl {                 // per-epoch constants of one channel, published by its two leader warps
    double start[3], step[3]; // numpy linspace constants of the three taps (tracking.py:110-112)
    double inv_step;          // ~1/step' of the prompt tap
    double ca, cb;            // carrier phase in turns at epoch-relative sample j: ca*j + cb
    long long a;              // epoch start sample (recording-relative)
    float ah, g2;             // alpha/2 = pi*ca; alpha^2/24
    int n, p0, stop, car_stop;
};

struct MChan {                // shared-memory state of one channel
    uint32_t cb[kCodeWords];                    // padded code, one bit per chip
    uint16_t stab[kPaddedChips + 1];            // byte 0 / 1 of entry k: top byte of +-1.0f for chips k / k+1
    MCtl ctl;
    alignas(16) int part[2][kMCW][8];           // fixed-point warp totals of an epoch, [epoch & 1][warp][component]
    float fix[kMCW][8];                         // exact-evaluation corrections of ambiguous samples, per warp
    alignas(8) uint64_t bar_part[2];
    sydr_trk_state cfgs;
    CodeState sc;
    CarrierState sk;
    LoopConst K;
    int n_hist[2];
    int rec_base, status, ch, active;
};

struct MShared {
    MChan chan[kMMaxGroup];
    alignas(16) int totals[2][kMPW][4];         // warp totals of a super-tile (producer scan)
    alignas(8) uint64_t bar_full[8];            // one per ring slot (super-tile), kMPW arrivals
    uint64_t bar_raw[kMRawBufs];
    volatile int progress[kMMaxGroup * kMCW];   // first super-tile a consumer warp may still read (INT_MAX = finished)
    volatile int go[2];                         // producer warp 0's verdict for a super-tile: 1 = produce, 0 = all consumers finished
    long long origin;                           // recording-relative sample of ring index 0 (multiple of kMSuper)
    long long valid_lo, valid_hi;               // samples readable from the recording's base pointer
    const uint8_t* rec_ptr;                     // sample 0 of the recording
    int n_super;                                // super-tiles the producers may have to make
    int rec_ok;
};

struct TrkmParams {
    TrkParams t;
    int n_channels;
    int group;                // channels per CTA
    int ring_super;           // super-tiles in the ring (power of two <= 8)
    double alpha_hc_max;      // |alpha| * half chip limit of the expansion
    int debug;
};

__device__ __forceinline__ uint32_t m_swz(uint32_t e) { return e ^ ((e >> 4) & 7u); }     // 16-entry lane stride -> 8 bank groups
__device__ __forceinline__ void m_named_barrier(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, int a, int b, int c, int d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- producer -----------------------------------------------------------------------------------------------------
// Super-tile k = ring-relative samples [k*kMSuper, (k+1)*kMSuper); producer warp pw owns [k*kMSuper + pw*kMTile, +kMTile),
// lane l the kMKS samples from j0 = ... + l*kMKS, and writes the entries of indices j0+1 .. j0+kMKS.
__device__ __forceinline__ void m_produce(MShared& sh, uint32_t ring_addr, uint32_t ring_mask, uint8_t* raw, const uint8_t* rec_base,
                                          int pw, int lane, int n_cons) {
    const unsigned full = 0xffffffffu;
    const int n_super = sh.n_super;
    const int ring_super = (int)((ring_mask + 1) / kMSuper);
    int c0r = 0, c0i = 0, c1r = 0, c1i = 0;                    // running prefix at the start of the super-tile (all producer threads)
    auto fast = [&](int k) {                                   // the whole super-tile is readable: TMA
        const long long s0 = sh.origin + (long long)k * kMSuper;
        return s0 >= sh.valid_lo && s0 + kMSuper <= sh.valid_hi;
    };
    auto issue = [&](int k) {                                  // one lane: stage super-tile k
        if (k < n_super && fast(k)) {
            const int b = k % kMRawBufs;
            mbar_arrive_expect_tx(&sh.bar_raw[b], kMSuper * 4);
            tma_bulk_g2s(raw + (size_t)b * kMSuper * 4, rec_base + (sh.origin + (long long)k * kMSuper) * 4, kMSuper * 4, &sh.bar_raw[b]);
        }
    };
    if (pw == 0 && lane == 0)
        for (int k = 0; k < kMRawBufs - 1; ++k) issue(k);
    int k = 0;
    for (; k < n_super; ++k) {
        const int j0 = k * kMSuper + pw * kMTile + lane * kMKS;          // ring-relative index of this lane's first sample
        uint32_t w[kMKS];
        if (fast(k)) {
            const int b = k % kMRawBufs;
            mbar_wait(&sh.bar_raw[b], (k / kMRawBufs) & 1);
            const uint4* src = reinterpret_cast<const uint4*>(raw + (size_t)b * kMSuper * 4) + (pw * kMTile + lane * kMKS) / 4;
#pragma unroll
            for (int v = 0; v < kMKS / 4; ++v) {
                const uint4 q = src[v];
                w[4 * v] = q.x; w[4 * v + 1] = q.y; w[4 * v + 2] = q.z; w[4 * v + 3] = q.w;
            }
        } else {                                                          // edge of the allocation: guarded loads, zeros outside
#pragma unroll
            for (int v = 0; v < kMKS; ++v) {
                const long long s = sh.origin + j0 + v;
                w[v] = (s >= sh.valid_lo && s < sh.valid_hi) ? reinterpret_cast<const uint32_t*>(rec_base)[s] : 0u;
            }
        }
        // pass 1: this lane's totals.  dp2a: I = low int16 of the word, Q = high int16.
        int t0r = 0, t0i = 0, t1r = 0, t1i = 0;
#pragma unroll
        for (int v = 0; v < kMKS; ++v) {
            t0r = __dp2a_lo((int)w[v], 0x0001, t0r);
            t0i = __dp2a_lo((int)w[v], 0x0100, t0i);
            t1r = __dp2a_lo((int)w[v], 2 * v, t1r);
            t1i = __dp2a_lo((int)w[v], (2 * v) << 8, t1i);
        }
        const int j2 = 2 * j0;                                            // P1 weights are 2 x (ring-relative index), modulo 2^32
        t1r += j2 * t0r;
        t1i += j2 * t0i;
        int s0r = t0r, s0i = t0i, s1r = t1r, s1i = t1i;                   // inclusive scan over the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a0 = __shfl_up_sync(full, s0r, o), a1 = __shfl_up_sync(full, s0i, o);
            const int a2 = __shfl_up_sync(full, s1r, o), a3 = __shfl_up_sync(full, s1i, o);
            if (lane >= o) { s0r += a0; s0i += a1; s1r += a2; s1i += a3; }
        }
        const int par = k & 1;
        if (lane == 31) *reinterpret_cast<int4*>(&sh.totals[par][pw][0]) = make_int4(s0r, s0i, s1r, s1i);
        if (pw == 0) {
            // room in the ring?  super-tile k overwrites k - ring_super: every consumer warp must have left it
            int go = 1;
            while (true) {
                const int p = (lane < n_cons) ? sh.progress[lane] : 0x7fffffff;
                const int mn = __reduce_min_sync(full, p);
                if (mn == 0x7fffffff) { go = 0; break; }
                if (mn > k - ring_super) break;
                __nanosleep(100);
            }
            if (lane == 0) sh.go[par] = go;
        }
        m_named_barrier(8, kMPW * 32);
        if (!sh.go[par]) break;
        if (pw == 0 && lane == 0) issue(k + kMRawBufs - 1);               // its buffer was last read in pass 1 of super-tile k-1
        // exclusive prefix at this lane's first sample
        int a0r = c0r + s0r - t0r, a0i = c0i + s0i - t0i, a1r = c1r + s1r - t1r, a1i = c1i + s1i - t1i;
#pragma unroll
        for (int q = 0; q < kMPW; ++q) {
            const int4 t = *reinterpret_cast<const int4*>(&sh.totals[par][q][0]);
            if (q < pw) { a0r += t.x; a0i += t.y; a1r += t.z; a1i += t.w; }
            c0r += t.x; c0i += t.y; c1r += t.z; c1i += t.w;
        }
        // pass 2: the entries.  P1[j0+v+1] = l1 + 2 j0 c0, l1 = (A1 - 2 j0 A0) + sum_{u<=v} 2u x_u
        int l1r = a1r - j2 * a0r, l1i = a1i - j2 * a0i;
#pragma unroll
        for (int v = 0; v < kMKS; ++v) {
            a0r = __dp2a_lo((int)w[v], 0x0001, a0r);
            a0i = __dp2a_lo((int)w[v], 0x0100, a0i);
            l1r = __dp2a_lo((int)w[v], 2 * v, l1r);
            l1i = __dp2a_lo((int)w[v], (2 * v) << 8, l1i);
            const uint32_t e = (uint32_t)(j0 + v + 1) & ring_mask;
            sts128(ring_addr + m_swz(e) * 16u, a0r, a0i, l1r + j2 * a0r, l1i + j2 * a0i);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.bar_full[k % ring_super]);          // release: the stores above are visible to the waiters
    }
    // left early (every consumer finished): the copies already requested must land before the shared memory is released
    if (pw == 0 && lane == 0)
        for (int kk = k + 1; kk < k + kMRawBufs - 1; ++kk)
            if (kk < n_super && fast(kk)) mbar_wait(&sh.bar_raw[kk % kMRawBufs], (kk / kMRawBufs) & 1);
}

// ---- consumer -----------------------------------------------------------------------------------------------------
// Exact treatment of an ambiguous boundary (the crossing is within 1e-9 sample of an integer J): sample J was given to
// the segment behind the boundary; its three code indices are re-evaluated with the reference expression and the
// difference, times the wiped-off sample, goes to the warp's correction sums.
static __device__ __noinline__ void m_correct(MChan& ch, int cw, uint32_t ring_addr, uint32_t ring_mask, int jr, int J, int p) {
    const MCtl& c = ch.ctl;
    const uint4 e0 = lds128(ring_addr + m_swz((uint32_t)jr & ring_mask) * 16u);
    const uint4 e1 = lds128(ring_addr + m_swz((uint32_t)(jr + 1) & ring_mask) * 16u);
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
            atomicAdd(&ch.fix[cw][2 * s], d * zr);
            atomicAdd(&ch.fix[cw][2 * s + 1], d * zi);
        }
    }
    if (err) atomicAdd(&ch.fix[cw][6], 1.0f);
}

__global__ void __launch_bounds__(kMMaxThreads, 1) trkm_kernel(const TrkmParams PM) {
    extern __shared__ __align__(128) uint8_t dyn_smem[];
    __shared__ __align__(16) MShared sh;
    const TrkParams& P = PM.t;
    const unsigned full = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = PM.group;
    const int ring_super = PM.ring_super;
    const uint32_t ring_entries = (uint32_t)ring_super * kMSuper, ring_mask = ring_entries - 1;
    const uint32_t ring_addr = smem_u32(dyn_smem);
    uint8_t* raw = dyn_smem + (size_t)ring_entries * 16;
    const int n_cons = G * kMCW;

    // ---- set-up: channel states, code tables, ring origin
    if (tid < G) {
        MChan& ch = sh.chan[tid];
        const int idx = blockIdx.x * G + tid;
        ch.ch = idx;
        ch.active = 0;
        if (idx < PM.n_channels) {
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
            // spacings must sit on the half-chip lattice at -1 / 0 / +1 half chips around the prompt tap
            int q[3];
            if (ch.active && !(seg_tap_offsets(g.spacing, q) && q[0] == -1 && q[2] == 1)) { ch.status = kNeedGeneral; ch.active = 0; }
        } else {
            ch.status = 1;
            ch.cfgs.status = 1;
        }
        mbar_init(&ch.bar_part[0], kMCW);
        mbar_init(&ch.bar_part[1], kMCW);
    }
    if (tid == 32) {
        for (int k = 0; k < 8; ++k) mbar_init(&sh.bar_full[k], kMPW);
        for (int k = 0; k < kMRawBufs; ++k) mbar_init(&sh.bar_raw[k], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        // one recording per CTA: the first active channel names it; a channel of another recording is left to the general kernel
        long long base = 0, lo = 0x7fffffffffffffffLL, hi = 0;
        int have = 0;
        for (int c = 0; c < G; ++c) {
            MChan& ch = sh.chan[c];
            if (!ch.active) continue;
            if (!have) { base = ch.cfgs.iq_base; have = 1; }
            if (ch.cfgs.iq_base != base || (base & 3) != 0) { ch.status = kNeedGeneral; ch.active = 0; continue; }
            lo = min(lo, (long long)ch.sc.cur);
            const long long alloc = P.iq_alloc - base;
            hi = max(hi, min((long long)ch.cfgs.iq_len, alloc));
        }
        sh.rec_ok = have;
        sh.origin = have ? (lo / kMSuper) * kMSuper : 0;
        sh.valid_lo = max(0LL, -base);
        sh.valid_hi = P.iq_alloc - base;
        sh.rec_ptr = P.iq + base * 4;
        const long long span = hi - sh.origin;
        sh.n_super = (have && span > 0) ? (int)min((span + kMSuper - 1) / kMSuper + 1, 0x3fffffffLL) : 0;
        *reinterpret_cast<uint4*>(dyn_smem) = make_uint4(0, 0, 0, 0);          // entry of ring index 0: the empty prefix
    }
    for (int i = tid; i < kMMaxGroup * kMCW; i += blockDim.x) sh.progress[i] = 0x7fffffff;
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
        for (int k = tid; k <= kPaddedChips; k += blockDim.x) {
            const int k0 = min(k, kPaddedChips - 1), k1 = min(k + 1, kPaddedChips - 1);
            const uint32_t b0 = (ch.cb[k0 >> 5] >> (k0 & 31)) & 1u, b1 = (ch.cb[k1 >> 5] >> (k1 & 31)) & 1u;
            ch.stab[k] = (uint16_t)((b0 ? 0x3Fu : 0xBFu) | ((b1 ? 0x3Fu : 0xBFu) << 8));
        }
        if (tid >= 32 * kMPW + c * 32 * kMCW && tid < 32 * kMPW + (c + 1) * 32 * kMCW && lane == 0)
            sh.progress[c * kMCW + ((warp - kMPW) % kMCW)] = -1;                // this consumer warp takes part (ring index 0 = "super-tile -1")
    }
    __syncthreads();

    if (warp < kMPW) {
        // ================================================================ producers
        if (sh.rec_ok) m_produce(sh, ring_addr, ring_mask, raw, sh.rec_ptr, warp, lane, n_cons);
    } else if (warp < kMPW + G * kMCW) {
        // ================================================================ consumers
        const int c = (warp - kMPW) / kMCW, cw = (warp - kMPW) % kMCW;
        MChan& ch = sh.chan[c];
        const int bar_id = 1 + c;
        if (ch.active) {
            sydr_trk_epoch* out_row = P.out + (long long)ch.ch * P.max_epochs + ch.rec_base;
            CodeState sc = ch.sc;
            CarrierState sk = ch.sk;
            int status = ch.cfgs.status;
            int epoch = 0;
            const int epoch_cap = P.max_epochs - ch.rec_base;
            const long long rec_alloc = P.iq_alloc - ch.cfgs.iq_base;
            const long long iq_len_reg = ch.cfgs.iq_len < rec_alloc ? ch.cfgs.iq_len : rec_alloc;
            const long long origin = sh.origin;
            int ready_upto = -1;                             // super-tiles [0, ready_upto] are known to be complete
            int released = -1;
            while (true) {
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
                m_named_barrier(bar_id, kMCW * 32);            // (A) the constants are visible
                if (ch.ctl.car_stop && !ch.ctl.stop) {         // leave the channel to the general kernel
                    if (cw == 0 && lane == 0) ch.status = kNeedGeneral;
                    break;
                }
                if (ch.ctl.stop) break;

                // ---- correlate: rounds of 31 segments, round r of the epoch belongs to warp r mod kMCW
                const MCtl& ctl = ch.ctl;
                const int n = ctl.n;
                const int a_rel = (int)(ctl.a - origin);                // ring-relative index of the epoch's first sample
                const double inv_step = (PM.debug & 4) ? 1.0 / ctl.step[1] : ctl.inv_step;
                const float ah = ctl.ah;
                const float gtab = 1.0f - ctl.g2 * (float)(lane * lane - 1);          // lane L: 1 - alpha^2 (L^2-1)/24
                const double ca_half = 0.5 * ctl.ca, cbt = ctl.cb;
                int p = ctl.p0 + kMSeg * cw + lane;                    // this lane's front boundary (lattice point)
                double x = seg_crossing(dmul(0.5, i2d(p)), ctl.start[1], inv_step);
                const double dx = (double)(kMSeg * kMCW) * 0.5 * inv_step;
                float aAr = 0.f, aAi = 0.f, aBr = 0.f, aBi = 0.f;
                if (lane < 8) ch.fix[cw][lane] = 0.f;
                __syncwarp();
                while (true) {
                    bool amb;
                    const int B = seg_first_sample(x, amb);
                    const int Bc = min(max(B, 0), n);
                    if (__shfl_sync(full, Bc, 0) >= n) break;          // the round starts behind the epoch
                    const int jr = a_rel + Bc;
                    const uint32_t slot = ring_addr + m_swz((uint32_t)jr & ring_mask) * 16u;
                    // every index of the round lies in or before the super-tile of lane 31's boundary
                    const int st_need = (__shfl_sync(full, jr, 31) - 1) >> 11;      // / kMSuper
                    static_assert(kMSuper == 2048, "shift above");
                    // what this warp may still read starts with this round: leave the super-tiles in front of it to the producers
                    const int st_first = (__shfl_sync(full, jr, 0) - 1) >> 11;
                    if (st_first > released) {
                        released = st_first;
                        if (lane == 0) sh.progress[c * kMCW + cw] = released;
                    }
                    while (ready_upto < st_need) {             // in order: a slot's phases are observed one by one
                        ++ready_upto;
                        mbar_wait(&sh.bar_full[ready_upto % ring_super], (ready_upto / ring_super) & 1);
                    }
                    const uint4 e0 = lds128(slot);
                    const uint4 e1 = lds128(__shfl_down_sync(full, slot, 1));
                    const int L = __shfl_down_sync(full, Bc, 1) - Bc;               // 0 for lane 31 and for clipped segments
                    const int d0r = (int)(e1.x - e0.x), d0i = (int)(e1.y - e0.y);
                    const int m = 2 * jr + L - 1;                                    // a + b - 1, ring-relative, modulo 2^32
                    const int d1r = (int)(e1.z - e0.z) - m * d0r, d1i = (int)(e1.w - e0.w) - m * d0i;
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
                    if (__any_sync(full, amb && lane < kMSeg && B >= 0 && B < n)) {
                        if (!(PM.debug & 2) && amb && lane < kMSeg && B >= 0 && B < n) m_correct(ch, cw, ring_addr, ring_mask, jr, B, p);
                        __syncwarp();
                    }
                    x += dx;
                    p += kMSeg * kMCW;
                    if (PM.debug & 1) x = seg_crossing(dmul(0.5, i2d(p)), ctl.start[1], inv_step);
                }
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
#pragma unroll
                    for (int w2 = 0; w2 < kMCW; ++w2) tot += ch.part[slot_e][w2][lane & 7];
                    const double ck = (double)tot * P.acc_inv;
                    sydr_trk_epoch* rec = out_row + e;
                    if (cw == 0) code_close(ch, sc, status, ck, rec, lane, cpre);
                    else carrier_close(ch, sk, ck, rc_next, rec, lane);
                }
            }
            if (cw == 0 && lane == 0) ch.sc = sc;
            if (cw == 1 && lane == 0) ch.sk = sk;
            if (lane == 0) sh.progress[c * kMCW + cw] = 0x7fffffff;
            m_named_barrier(bar_id, kMCW * 32);
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
        } else {
            // idle slot, finished earlier, or left to the general kernel: no epochs from this launch
            if (cw == 0 && lane == 0 && ch.ch < PM.n_channels) {
                if (ch.status == kNeedGeneral) P.states[ch.ch].status = kNeedGeneral;
                P.nepochs[ch.ch] = ch.rec_base;
            }
        }
    }
    __syncthreads();
}

}  // namespace sydr

using namespace sydr;

namespace sydr {

int g_trkm_debug = 0;
size_t trkm_smem_bytes(int ring_super) { return (size_t)ring_super * kMSuper * 16 + (size_t)kMRawBufs * kMSuper * 4; }

// Launch the prefix-moment kernel for n_channels channels, `group` consecutive channels per CTA (they must belong to
// one recording: same iq_base; a channel that does not is left to the general kernel).
int launch_trkm(const TrkParams& P, int n_channels, int group, cudaStream_t s) {
    SYDR_REQUIRE(group >= 1 && group <= kMMaxGroup, SYDR_ERR_ARG, "group must be 1..%d (got %d)", kMMaxGroup, group);
    TrkmParams PM;
    PM.t = P;
    PM.n_channels = n_channels;
    PM.group = group;
    PM.ring_super = 4;
    PM.alpha_hc_max = 0.06;
    PM.debug = g_trkm_debug;
    const size_t smem = trkm_smem_bytes(PM.ring_super);
    SYDR_CUDA_CHECK(cudaFuncSetAttribute(trkm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = 32 * (kMPW + group * kMCW);
    const int grid = (n_channels + group - 1) / group;
    trkm_kernel<<<grid, threads, smem, s>>>(PM);
    count_launch();
    SYDR_CUDA_CHECK(cudaGetLastError());
    return SYDR_OK;
}

}  // namespace sydr

extern "C" int sydr_trkm_debug(int flags) { sydr::g_trkm_debug = flags; return 0; }
