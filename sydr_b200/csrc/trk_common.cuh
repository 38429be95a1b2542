// Device helpers shared by the tracking kernels (trk.cu: chunk / half-chip segment formulations, latency and
// throughput shapes; trkm.cu: prefix-moment formulation): reference arithmetic of the code / carrier NCOs, chip
// boundaries, epoch constants and the Borre / Kaplan loop closures.  See trk.cu for the reference citations.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace sydr {

// Samples per thread chunk: a multiple of the 16-byte vector and of 4, chosen so that the
// per-thread stride in shared memory (48 / 160 bytes) keeps LDS.128 (nearly) conflict free and
// that a cluster of 8 CTAs still has ~2 warps per scheduler at 25 MS/s.
// VPC (vectors per chunk) is a kernel template parameter: int8 3 (24 samples, 48 B); int16 3
// (12 samples, 48 B; latency mode, clusters of >= 4 CTAs) or 5 (20 samples, 80 B; throughput
// mode); complex64 10 (20 samples, 160 B).

constexpr int kMaxChunk = 24;   // longest chunk (int8: 3 vectors of 8 samples)

struct EpochConst {
    double ca, cb;         // carrier phase in turns at sample j: ca*j + cb  (ca = -fc/fs, cb = rem/(2 pi))
    double start[3];       // linspace start  = remCode + spacing     tracking.py:110
    double step[3];        // linspace step'  = (stop-start)/n        numpy linspace
    double inv_step[3];    // ~1/step' (locates chip boundaries; every boundary is then pinned exactly)
    float w[4][2];         // carrier rotation by 1, 2, 3, 4 samples: exp(-j 2 pi k fc/fs)
    float wtab[kMaxChunk][2];   // carrier rotation by u samples, u < chunk length (split-sum path)
    int n;                 // samples in the epoch
    int fast;              // every tap keeps a chip for more than kMaxChunk samples: <= 1 flip per chunk and tap
    int seg;               // half-chip segment path applies to this epoch (see correlate_segment)
    int hb;                // first half-chip lattice index this CTA looks at
    int rounds;            // rounds of 32 segments per warp (throughput loop)
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
static __device__ __noinline__ int boundary_exact(int t, int jlo, int jhi, int k, double start, double step) {
    t = max(jlo + 1, min(t, jhi));
    while (t > jlo + 1 && code_phase(t - 1, start, step) > i2d(k)) --t;
    while (t < jhi && !(code_phase(t, start, step) > i2d(k))) ++t;
    return t;
}

// Generic exact walk (Python wrap-around / out-of-range indices); rare.
static __device__ __noinline__ uint32_t tap_mask_walk(int jlo, int jhi, int k0, double start, double step,
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
static __device__ __noinline__ uint32_t tap_mask(int jlo, double jd, int cnt, const EpochConst* ec, int s,
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

// ---- split-sum path -------------------------------------------------------------------------
// When a chip lasts longer than a chunk, each tap changes sign at most once inside a chunk, and
// with the usual half-chip spacing at most one *position* in the chunk carries a change (the
// prompt boundary and the early/late boundary alternate every half chip).  The chunk is then two
// runs of constant (E, P, L) signs split at sample T: accumulate Y = sum_u x_u w^u over the whole
// chunk and A = the same sum over u < T (the gate [u < T] is one saturating FADD on the FMA
// pipe), and apply carrier phasor and the six signs once per chunk.  No per-sample chip work is
// left, which matters because LOP3/SHF/PRMT issue at half the FFMA rate on sm_100a.
//
// Position of the first sample (offset from `lo`) whose chip index exceeds that of sample lo:
// the real-valued crossing x = (k - phase(lo)) / step' decides it unless x is within 1e-9 of an
// integer (the reference's two roundings move a crossing by < 1e-11 sample), in which case the
// caller falls back to the exact walk.  b0 / b1 = code bits of the chip under `lo` and the next.
__device__ __forceinline__ bool tap_split(double jd, const EpochConst& ec, int s, const uint32_t* cb, int& t,
                                          uint32_t& b0, uint32_t& b1) {
    const double ph = dadd(dmul(jd, ec.step[s]), ec.start[s]);        // numpy's linspace value at lo
    const double mg = __dadd_ru(ph, 6755399441055744.0);              // ceil(ph) in the low word
    const int k = __double2loint(mg);
    const double x = dmul(dsub(dsub(mg, 6755399441055744.0), ph), ec.inv_step[s]);   // >= 0, in samples
    const double xm = __dadd_rd(x, 6755399441055744.0);               // floor(x) in the low word
    const double fr = x - (xm - 6755399441055744.0);                  // [0, 1)
    t = __double2loint(xm) + 1;
    const int wi = k >> 5;
    const uint32_t w2 = __funnelshift_r(cb[wi & 31], cb[(wi & 31) + 1], k & 31);
    b0 = w2 & 1u;
    b1 = (w2 >> 1) & 1u;
    return (fr > 1e-9) && (fr < 1.0 - 1e-9) && (k >= 0) && (k < kPaddedChips - 1);
}

__device__ __forceinline__ float sign_of_bit(uint32_t bit) {      // bit 1 -> +1.0f, bit 0 -> -1.0f
    return __uint_as_float(0x3f800000u | ((bit ^ 1u) << 31));
}

// Correlate one chunk of C = VPC*SPV samples starting at epoch-relative index j0 against the
// three taps.  `src` points at the chunk's first vector (shared or global memory, 16-byte
// aligned); `ec` lives in shared memory.  acc = {IE, QE, IP, QP, IL, QL}.
template <int DT, int VPC>
__device__ __forceinline__ void correlate_chunk(const uint4* src, int j0, const EpochConst& ec,
                                                const uint32_t* cb, float* acc, int& err) {
    constexpr int SPV = IqTraits<DT>::SPV;
    constexpr int C = SPV * VPC;
    static_assert(C <= kMaxChunk, "chunk longer than the carrier table");
    const int lo = max(j0, 0);
    const int hi = min(j0 + C, ec.n);
    if (hi <= lo) return;
    const bool edge = (lo != j0) || (hi != j0 + C);
    const double jd = i2d(lo);

    // Carrier seed (tracking.py:102): phase of sample j0 in turns, FP64, reduced to [-0.5, 0.5].
    double turns = fma(ec.ca, i2d(j0), ec.cb);
    turns -= drint(turns);
    float pre, pim;
    __sincosf((float)turns * 6.283185307179586f, &pim, &pre);

    if (ec.fast) {
        int t0, t1, t2;
        uint32_t p0, p1, p2, q0, q1, q2;                    // code bits before / after each tap's flip
        bool ok = tap_split(jd, ec, 0, cb, t0, p0, q0);
        ok &= tap_split(jd, ec, 1, cb, t1, p1, q1);
        ok &= tap_split(jd, ec, 2, cb, t2, p2, q2);
        const int cnt = hi - lo;
        // flips beyond the valid samples do not exist for this chunk
        t0 = (t0 < cnt) ? t0 : 0x10000; t1 = (t1 < cnt) ? t1 : 0x10000; t2 = (t2 < cnt) ? t2 : 0x10000;
        const int tmin = min(t0, min(t1, t2));
        ok &= (t0 == tmin || t0 == 0x10000) && (t1 == tmin || t1 == 0x10000) && (t2 == tmin || t2 == 0x10000);
        if (ok) {
            const float tf = (float)min(tmin + (lo - j0), C);        // split position relative to j0
            float Yr = 0.f, Yi = 0.f, Ar = 0.f, Ai = 0.f;
#pragma unroll
            for (int v = 0; v < VPC; ++v) {
                uint4 cur = src[v];
                if (edge) mask_vector<DT>(cur, lo - (j0 + v * SPV), hi - (j0 + v * SPV));
                float re[SPV], im[SPV];
                decode_vec<DT>(cur, re, im);
#pragma unroll
                for (int u = 0; u < SPV; ++u) {
                    const int uu = v * SPV + u;
                    const float wr = ec.wtab[uu][0], wi = ec.wtab[uu][1];
                    const float yr = fmaf(re[u], wr, -(im[u] * wi));        // x_u * w^u
                    const float yi = fmaf(re[u], wi, im[u] * wr);
                    const float g = __saturatef(tf - (float)uu);            // 1 before the split, 0 after
                    Yr += yr; Yi += yi;
                    Ar = fmaf(g, yr, Ar); Ai = fmaf(g, yi, Ai);
                }
            }
            const float Br = Yr - Ar, Bi = Yi - Ai;
            // signal = replica * rfData (tracking.py:105): rotate both runs by the chunk phasor
            const float zar = pre * Ar - pim * Ai, zai = pre * Ai + pim * Ar;
            const float zbr = pre * Br - pim * Bi, zbi = pre * Bi + pim * Br;
            const float a0 = sign_of_bit(p0), a1 = sign_of_bit(p1), a2 = sign_of_bit(p2);
            const float b0 = sign_of_bit(t0 == tmin ? q0 : p0), b1 = sign_of_bit(t1 == tmin ? q1 : p1),
                        b2 = sign_of_bit(t2 == tmin ? q2 : p2);
            acc[0] += fmaf(a0, zar, b0 * zbr); acc[1] += fmaf(a0, zai, b0 * zbi);
            acc[2] += fmaf(a1, zar, b1 * zbr); acc[3] += fmaf(a1, zai, b1 * zbi);
            acc[4] += fmaf(a2, zar, b2 * zbr); acc[5] += fmaf(a2, zai, b2 * zbi);
            return;
        }
    }

    // ---- general path: per-sample sign masks (any chip rate, ties, index wrap-around)
    uint4 cur = src[0];
    uint32_t m0 = tap_mask(lo, jd, hi - lo, &ec, 0, cb, &err) << (lo - j0);
    uint32_t m1 = tap_mask(lo, jd, hi - lo, &ec, 1, cb, &err) << (lo - j0);
    uint32_t m2 = tap_mask(lo, jd, hi - lo, &ec, 2, cb, &err) << (lo - j0);

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

// ---- half-chip segment path ------------------------------------------------------------------
// With correlator spacings that are multiples of half a chip (the reference's -0.5 / 0 / +0.5,
// channel_GPS_L1CA_borre.ini:14-16 and channel_GPS_L1CA_kaplan.ini:13-14) every tap changes
// chip only where the prompt code phase crosses a multiple of 0.5: between two consecutive
// lattice points h-1 and h the three code values are constant.  One thread therefore owns one
// *segment* H (the samples with (H-1)/2 < phase <= H/2, ~12.2 at 25 MS/s): it sums
// x_u w^u over the segment with a two-chain Horner recurrence in packed FP32 (FFMA2), rotates
// the sum by the carrier phasor of its first sample and adds it to E, P and L with the three
// signs of the segment, which come from a per-channel table indexed by H (code index of tap s =
// ceil((H + q_s) / 2), q_s = 2 (spacing_s - spacing_prompt)).  There is no per-sample chip,
// gate or table work at all.
//
// Exactness.  The first sample of segment h+1 is B(h) = min{ j : phase(j) > h/2 }.  It is
// located from the real-valued crossing X(h) = (h/2 - start) / step', whose distance to the
// decision boundary of the reference expression ceil(fl(fl(j step') + start)) is < 1e-10
// sample for every tap; when X(h) is within 1e-9 of an integer J the sample J is given to
// segment h+1 and its three code indices are evaluated with the reference expression itself
// (seg_correct), so every sample carries exactly the chip the reference gives it.
constexpr int kSegTab = 2080;                 // lattice indices 0 .. 2079 (an epoch touches <= 2048)
constexpr double kMagic52 = 6755399441055744.0;   // 1.5 * 2^52

// Lattice crossing X(h) in samples: every thread that needs it evaluates this very expression.
__device__ __forceinline__ double seg_crossing(double hd_half, double start, double inv_step) {
    return dmul(dsub(hd_half, start), inv_step);
}
// B(h) = floor(X) + 1, or floor(X) when X is within 1e-9 above an integer; amb = X within 1e-9 of
// an integer (the sample B(h) then needs the exact evaluation).
__device__ __forceinline__ int seg_first_sample(double x, bool& amb) {
    const double xm = __dadd_rd(x, kMagic52);
    const double fr = x - (xm - kMagic52);                  // [0, 1)
    const bool low = fr < 1e-9;
    amb = low || (fr > 1.0 - 1e-9);
    return __double2loint(xm) + (low ? 0 : 1);
}

// Signs of the three taps for every lattice index: byte s of tab[H] is the top byte of +1.0f
// (0x3F) or -1.0f (0xBF) for tap s, so one PRMT per tap turns the entry into the sign factor.
__device__ __forceinline__ void build_seg_table(uint32_t* tab, const uint32_t* cb, const int* q) {
    for (int H = threadIdx.x; H < kSegTab; H += blockDim.x) {
        uint32_t e = 0;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int k = min(max((H + q[s] + 1) >> 1, 0), kPaddedChips - 1);
            const uint32_t bit = (cb[k >> 5] >> (k & 31)) & 1u;
            e |= (bit ? 0x3Fu : 0xBFu) << (8 * s);
        }
        tab[H] = e;
    }
}
__device__ __forceinline__ float seg_sign(uint32_t entry, int s) {
    return __uint_as_float(__byte_perm(entry, 0x00800000u, s == 0 ? 0x0644 : s == 1 ? 0x1644 : 0x2644));
}
// q_s = 2 (spacing_s - spacing_prompt) when every spacing is a multiple of half a chip around the
// prompt tap; returns false otherwise (the chunk paths then serve the channel).
__device__ __forceinline__ bool seg_tap_offsets(const double* spacing, int* q) {
    bool ok = true;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const double d = 2.0 * (spacing[s] - spacing[1]);
        const double r = drint(d);
        ok = ok && (d == r) && (fabs(r) <= 8.0);
        q[s] = (int)r;
    }
    return ok;
}

// Exact treatment of an ambiguous first sample J of segment h (rare): the difference between the
// code values the reference expression gives and the segment's, tap by tap, times the wiped-off
// sample.  Returned in registers so that the accumulators never live in local memory.
struct SegDelta { float d[6]; int err; };
static __device__ __noinline__ SegDelta seg_correct(uint32_t raw, int J, int h, float pr, float pi, const EpochConst* ec,
                                             const int* q, const uint32_t* cb) {
    const uint32_t t = raw ^ 0x80008000u;
    const float magic = 12582912.f + 32768.f;
    const float xr = __uint_as_float(__byte_perm(t, 0x4B400000u, 0x7610)) - magic;
    const float xi = __uint_as_float(__byte_perm(t, 0x4B400000u, 0x7632)) - magic;
    const float zr = pr * xr - pi * xi, zi = pr * xi + pi * xr;
    SegDelta o;
    o.err = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        const int k_exact = ceil_to_int(code_phase(J, ec->start[s], ec->step[s]));
        const int k_seg = (h + q[s] + 1) >> 1;
        float d = 0.f;
        if (k_exact != k_seg) d = sign_of_bit(chip_bit(cb, k_exact, o.err)) - sign_of_bit(chip_bit(cb, k_seg, o.err));
        o.d[2 * s] = d * zr;
        o.d[2 * s + 1] = d * zi;
    }
    return o;
}

// First sample of segment h+1 (see above); amb = it needs the exact evaluation.
__device__ __forceinline__ int seg_bound(int h, const EpochConst& ec, bool& amb) {
    return seg_first_sample(seg_crossing(dmul(0.5, i2d(h)), ec.start[1], ec.inv_step[1]), amb);
}

// Carrier phasor of epoch-relative sample j (tracking.py:102), FP64 seed.
__device__ __forceinline__ void seg_phasor(const EpochConst& ec, int j, float& pre, float& pim) {
    double turns = fma(ec.ca, i2d(j), ec.cb);
    turns -= drint(turns);
    __sincosf((float)turns * 6.283185307179586f, &pim, &pre);
}

// The arithmetic of one segment of int16 IQ.  w[0 .. 2NV) are the raw samples from the even
// address at or below the segment's first sample; samples [lo, hi) of them belong to the segment;
// (pre, pim) = carrier phasor of w[0], whose epoch-relative index is j0.  LMIN = 2 NV - 2 is the
// shortest unclipped segment.  CLAMP: h may lie outside the sign table (general kernel only).
template <int NV, bool CLAMP>
__device__ __forceinline__ void seg_body(uint32_t* w, int lo, int hi, int j0, float pre, float pim, int h, bool amb0,
                                         const EpochConst& ec, const uint32_t* tab, const int* q, const uint32_t* cb,
                                         float* acc, int& err) {
    constexpr int LMIN = 2 * NV - 2;
    const uint32_t raw_first = lo ? w[1] : w[0];       // for the exact evaluation of an ambiguous first sample

    // samples outside [lo, hi) do not belong to this segment
    if (hi - lo >= LMIN) {
        w[0] = lo ? 0u : w[0];
#pragma unroll
        for (int u = LMIN; u < 2 * NV; ++u) w[u] = (u < hi) ? w[u] : 0u;
    } else {                                           // clipped by the epoch's ends
#pragma unroll
        for (int u = 0; u < 2 * NV; ++u) w[u] = (u >= lo && u < hi) ? w[u] : 0u;
    }

    // Y = sum_u x_u w^u.  Samples u = 4m, 4m+1 (pairs k = 2m) and u = 4m+2, 4m+3 (pairs k = 2m+1)
    // form two sets of Horner chains in w^4; each set packs its even / odd sample chain in FFMA2.
    // Four independent chains halve the dependent depth; the sets are joined with one rotation by w^2.
    const float magic = 12582912.f + 32768.f;
    const float2 nm = make_float2(-magic, -magic);
    const float2 w4r = make_float2(ec.w[3][0], ec.w[3][0]);
    const float2 w4i = make_float2(ec.w[3][1], ec.w[3][1]);
    const float2 w4n = make_float2(-ec.w[3][1], -ec.w[3][1]);
    float2 RA = make_float2(0.f, 0.f), IA = RA, RB = RA, IB = RA;   // (even sample, odd sample) of set A / B
#pragma unroll
    for (int k = NV - 1; k >= 0; --k) {
        const uint32_t t0 = w[2 * k] ^ 0x80008000u, t1 = w[2 * k + 1] ^ 0x80008000u;
        float2 xr, xi;
        xr.x = __uint_as_float(__byte_perm(t0, 0x4B400000u, 0x7610));
        xr.y = __uint_as_float(__byte_perm(t1, 0x4B400000u, 0x7610));
        xi.x = __uint_as_float(__byte_perm(t0, 0x4B400000u, 0x7632));
        xi.y = __uint_as_float(__byte_perm(t1, 0x4B400000u, 0x7632));
        xr = __fadd2_rn(xr, nm);
        xi = __fadd2_rn(xi, nm);
        float2& R_ = (k & 1) ? RB : RA;
        float2& I_ = (k & 1) ? IB : IA;
        if (k >= NV - 2) {                               // first pair of its set
            R_ = xr;
            I_ = xi;
        } else {
            const float2 r2 = __ffma2_rn(w4n, I_, __ffma2_rn(w4r, R_, xr));
            I_ = __ffma2_rn(w4i, R_, __ffma2_rn(w4r, I_, xi));
            R_ = r2;
        }
    }
    float2 R = RA, I = IA;
    if (NV > 1) {                                        // set B starts two samples later
        const float2 w2r = make_float2(ec.w[1][0], ec.w[1][0]);
        const float2 w2i = make_float2(ec.w[1][1], ec.w[1][1]);
        const float2 w2n = make_float2(-ec.w[1][1], -ec.w[1][1]);
        R = __ffma2_rn(w2n, IB, __ffma2_rn(w2r, RB, RA));
        I = __ffma2_rn(w2i, RB, __ffma2_rn(w2r, IB, IA));
    }
    const float w1r = ec.w[0][0], w1i = ec.w[0][1];
    const float yr = fmaf(-w1i, I.y, fmaf(w1r, R.y, R.x));
    const float yi = fmaf(w1i, R.y, fmaf(w1r, I.y, I.x));
    // signal = replica * rfData (tracking.py:105)
    const float zr = pre * yr - pim * yi, zi = pre * yi + pim * yr;
    const uint32_t e = tab[CLAMP ? min(max(h, 0), kSegTab - 1) : h];
    const float s0 = seg_sign(e, 0), s1 = seg_sign(e, 1), s2 = seg_sign(e, 2);
    acc[0] = fmaf(s0, zr, acc[0]); acc[1] = fmaf(s0, zi, acc[1]);
    acc[2] = fmaf(s1, zr, acc[2]); acc[3] = fmaf(s1, zi, acc[3]);
    acc[4] = fmaf(s2, zr, acc[4]); acc[5] = fmaf(s2, zi, acc[5]);
    if (amb0) {                                        // first sample needs the exact code indices
        float pr = pre, pi = pim;
        if (lo) { pr = pre * w1r - pim * w1i; pi = pre * w1i + pim * w1r; }
        const SegDelta dlt = seg_correct(raw_first, j0 + lo, h, pr, pi, &ec, q, cb);
#pragma unroll
        for (int k = 0; k < 6; ++k) acc[k] += dlt.d[k];
        err |= dlt.err;
    }
}

// One segment from a window in shared or global memory.  `base` points at the sample whose
// epoch-relative index is `wlo` (8-byte aligned); the thread owns segment h if its first sample
// lies in [wlo, whi).  Returns false once the segment starts at or beyond min(whi, n): the
// caller's loop over h ends.
template <int NV>
__device__ __forceinline__ bool correlate_segment(const uint8_t* base, int wlo, int whi, int h, const EpochConst& ec,
                                                  const uint32_t* tab, const int* q, const uint32_t* cb, float* acc,
                                                  int& err) {
    bool amb0, amb1;
    const int B0 = seg_bound(h - 1, ec, amb0);
    const int B1 = seg_bound(h, ec, amb1);
    const int start = max(B0, 0), end = min(B1, ec.n);
    if (start >= min(whi, ec.n)) return false;
    if (end <= start || start < wlo) return true;
    const int off = start - wlo;                       // samples from `base`
    const int lo = off & 1;
    const uint2* src = reinterpret_cast<const uint2*>(base + (size_t)(off - lo) * 4);
    uint32_t w[2 * NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const uint2 v = src[k];
        w[2 * k] = v.x;
        w[2 * k + 1] = v.y;
    }
    float pre, pim;
    seg_phasor(ec, start - lo, pre, pim);
    seg_body<NV, true>(w, lo, lo + (end - start), start - lo, pre, pim, h, amb0 && B0 >= 0, ec, tab, q, cb, acc, err);
    return true;
}

// ---- throughput loop: every warp owns R*32 consecutive segments and walks them in R rounds of
// 32.  The samples of the *next* round (one contiguous piece of the recording, ~1.6 KB) travel
// into a private shared-memory slot of the warp with one TMA bulk copy (cp.async.bulk + mbarrier,
// no LSU work) while the current round is computed, and the first sample of segment h+1 is
// taken from the neighbouring lane instead of being located twice.
template <int NV>
struct RoundTraits {
    // a lane's windows of consecutive rounds start 32 segments apart: 32 (2 NV - 2) - 1 ..
    // 32 (2 NV - 1) + 1 samples; the carrier rotations by these distances are tabulated per epoch
    static constexpr int kRotBase = 32 * (2 * NV - 2) - 1;
    static constexpr int kRotN = 35;
};
constexpr int kRotMax = 36;

__device__ __forceinline__ void ldg64_nc(const void* p, uint32_t& a, uint32_t& b) {
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
}

// Throughput loop: every warp owns R*32 consecutive segments and walks them in R rounds of 32.
// The 2 NV words of a lane's *next* segment are requested from global memory (L1 / L2: the 12
// channels of a recording read the same lines) before the current segment is computed, so their
// latency hides behind ~200 instructions; the boundary behind a lane's segment comes from the
// neighbouring lane, the boundary in front of its next segment from one FP64 addition, and the
// carrier phasor from a tabulated rotation of the previous round's.
template <int NV>
__device__ __forceinline__ void correlate_rounds(const uint8_t* rec_epoch /* sample 0 of the epoch */, int rounds,
                                                 const EpochConst& ec, const float2* rot, const uint32_t* tab,
                                                 const int* q, const uint32_t* cb, float* acc, int& err) {
    using RT = RoundTraits<NV>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned full = 0xffffffffu;
    const int n = ec.n;
    // parity of the epoch's first sample inside its 8-byte pair
    const int al = (int)((reinterpret_cast<uintptr_t>(rec_epoch) >> 2) & 1);
    const int h0 = ec.hb + warp * rounds * 32;              // first segment of the warp
    int h = h0 + lane;
    // crossing of the lattice point in front of this lane's segment; advanced by 32 half chips per
    // round (x is only an estimate: seg_first_sample flags the cases that need the exact expression)
    const double start_b = ec.start[1], inv_step = ec.inv_step[1];
    double x = seg_crossing(dmul(0.5, i2d(h - 1)), start_b, inv_step);
    const double d32 = 16.0 * inv_step;
    bool ambc;
    int Bc = seg_first_sample(x, ambc);
    // the boundary behind the warp's last segment, evaluated like the next warp's first
    bool amb_unused;
    const int Bend = seg_first_sample(seg_crossing(dmul(0.5, i2d(h0 + rounds * 32 - 1)), start_b, inv_step), amb_unused);
    float pre = 1.f, pim = 0.f;
    int jprev = -0x40000000;                                // forces the FP64 seed in the first round
    uint32_t wa[2 * NV], wb[2 * NV];
    auto request = [&](int B, uint32_t* w) {                // the 2 NV words from the even address at the segment's start
        const int s0 = min(max(B, 0), n);
        const uint8_t* g = rec_epoch + (long long)(s0 - ((s0 + al) & 1)) * 4;
#pragma unroll
        for (int k = 0; k < NV; ++k) ldg64_nc(g + 8 * k, w[2 * k], w[2 * k + 1]);
    };
    request(Bc, wa);

#define SYDR_ROUND(WCUR, WNXT)                                                                             \
    {                                                                                                      \
        const double xn = x + d32;                                                                         \
        bool ambn;                                                                                         \
        const int Bn = seg_first_sample(xn, ambn);                                                         \
        int B1 = __shfl_down_sync(full, Bc, 1);                                                            \
        const int Bw = __shfl_sync(full, Bn, 0);                                                           \
        if (lane == 31) B1 = (r + 1 == rounds) ? Bend : Bw;                                                \
        const int start = max(Bc, 0), end = min(B1, n);                                                    \
        if (__all_sync(full, start >= n)) break;                                                           \
        if (r + 1 < rounds) request(Bn, WNXT);                                                             \
        if (end > start) {                                                                                 \
            const int lo = (start + al) & 1;                                                               \
            const int j0 = start - lo;                                                                     \
            const uint32_t d = (uint32_t)(j0 - jprev - RT::kRotBase);                                      \
            if (d < (uint32_t)RT::kRotN) {                                                                 \
                const float2 c = rot[d];                                                                   \
                const float t = pre * c.x - pim * c.y;                                                     \
                pim = pre * c.y + pim * c.x;                                                               \
                pre = t;                                                                                   \
            } else {                                                                                       \
                seg_phasor(ec, j0, pre, pim);                                                              \
            }                                                                                              \
            jprev = j0;                                                                                    \
            seg_body<NV, false>(WCUR, lo, lo + (end - start), j0, pre, pim, h, ambc && Bc >= 0, ec, tab,   \
                                q, cb, acc, err);                                                          \
        }                                                                                                  \
        x = xn;                                                                                            \
        Bc = Bn;                                                                                           \
        ambc = ambn;                                                                                       \
        h += 32;                                                                                           \
        ++r;                                                                                               \
    }

    for (int r = 0; r < rounds;) {
        SYDR_ROUND(wa, wb)
        if (r >= rounds) break;
        SYDR_ROUND(wb, wa)
    }
#undef SYDR_ROUND
}

// Does the segment path apply to an epoch?  Every code index must stay inside the padded code
// (no Python wrap-around, no IndexError) and a half chip must last between LMIN and LMIN + 1
// samples (clear of the integers, so that a segment never exceeds the 2 NV - 1 samples loaded).
template <int NV>
__device__ __forceinline__ bool seg_epoch_ok(double start_min, double stop_max, double inv_step) {
    const double hc = 0.5 * inv_step;
    return (hc >= (double)(2 * NV - 2) + 0.02) && (hc <= (double)(2 * NV - 1) - 0.02) &&
           (start_min > -0.999) && (stop_max < (double)(kPaddedChips - 1) - 0.001);
}

// a / b with a reciprocal rb of b that is correctly rounded or within an ulp: one multiply and
// the FMA residual correction (Markstein).  Returns the correctly rounded quotient for the
// operands of this file (checked exhaustively-at-random on the host against IEEE division:
// code_freq/fs, phase/(2 pi), (1023 - rem)/code_step, (stop - start)/n).
__device__ __forceinline__ double ddiv_by(double a, double b, double rb) {
    const double q = a * rb;
    return fma(fma(-b, q, a), rb, q);
}
__device__ __forceinline__ double newton_rcp(double b, double r) { return fma(fma(-b, r, 1.0), r, r); }

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

// Entry u of the carrier table: rotation by u samples, exp(2 pi i u ca), from the FP64 turn count.
__device__ __forceinline__ void carrier_table_entry(double ca, int u, float& c, float& s) {
    double tu = dmul(i2d(u), ca - drint(ca));
    tu -= drint(tu);
    sincospif((float)(2.0 * tu), &s, &c);
}

// Per-epoch constants from the NCO state (one thread; open-loop kernel).  `chunk` = samples per
// thread chunk of the calling kernel (decides whether the split-sum path applies).
__device__ __forceinline__ void make_epoch_const(EpochConst& ec, int n, double fs, double fc,
                                                 double rem_carrier, double rem_code,
                                                 double code_step, const double* spacing, int chunk) {
    ec.n = n;
    bool fast = true;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        tap_const(rem_code, spacing[s], code_step, n, ec.start[s], ec.step[s], ec.inv_step[s]);
        fast = fast && (ec.inv_step[s] >= (double)(chunk + 1));
    }
    ec.fast = fast ? 1 : 0;
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
// Closed loop.
// ------------------------------------------------------------------------------------------
// The NCO / loop-filter state of one channel (mirrors sydr_trk_state), split by owner: the code
// loop lives in warp 0, the carrier loop in warp 1; the two chains never read each other's
// variables (channel_l1ca_borre.py:363-429), so they close concurrently.
struct CodeState {
    long long cur;
    int n_req;
    double code_freq, code_step, rem_code, nco_code_err, nco_code;
    double inv_step;         // ~1/code_step, carried from epoch to epoch (one Newton step per epoch)
    double inv_n;            // ~1/n_req, refined with two Newton steps when n changes
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
    long long iq_base;       // with has_iq_base: overrides the states' iq_base (sliding window, may be negative)
    int has_iq_base;
    int seg;                 // half-chip segment path allowed (sampling rate fits the instantiation)
    float acc_scale;         // integer IQ: power of two that takes a warp's correlator sums into int32 (fixed-point all-reduce)
    double acc_inv;          // 1 / acc_scale
    int resume;              // follow-up of a LEAN launch: continue the record rows, serve kNeedGeneral channels
    long long* prof;         // optional [n_channels][16] phase cycle counters of thread 0 (NULL = off)
    int dense;               // 1: DENSE instantiation requested (steps in flight share the GPU); 2: PACK
    sydr_kaplan_state* kstates;   // Kaplan loop closure (KAP instantiation): per-channel state and per-epoch extras
    sydr_kaplan_epoch* kout;
};

struct EpochCtl {            // published by warps 0 / 1 for every epoch
    EpochConst ec;
    long long a;             // epoch start sample (rec-relative)
    int stop;
};

constexpr int kMaxCluster = 8;
constexpr int kWinTail = 32;     // samples staged beyond a CTA's window (segments that start inside may end outside)
constexpr int kTrkMaxThreads = 384;
constexpr int kDenseMaxThreads = 288;    // DENSE instantiation: <= 113 registers, three CTAs per SM beside other launches
constexpr int kLeanThreads = 256;
constexpr int kPackThreads = 320;        // PACK instantiation: two CTAs per SM under 102 registers
constexpr int kPackWave = 296;           // channels one launch of the PACK shape holds (two per SM)
constexpr int kPackWindowBytes = 108 * 1024;   // its single window: what two CTAs per SM leave of 227 KB beside 5 KB of static shared memory
constexpr int kKaplanMaxThreads = 384;   // the Kaplan carrier warp holds more state: 170 registers per thread instead of 102
constexpr int kNeedGeneral = 2;  // channel status: stopped in front of an epoch only the general kernel serves

constexpr int kTrkMaxWarps = kTrkMaxThreads / 32;

template <int NE, int NTAB>
struct TrkSharedT {          // static shared memory of the closed-loop kernel
    uint32_t cb[kCodeWords];
    EpochCtl ctl;
    // per-warp partial sums of every CTA of the cluster, [slot][rank * W + warp][component]
    alignas(16) double gather[2][NE][8];
    alignas(8) uint64_t bar_data[2];
    uint64_t bar_gather[2];
    uint64_t bar_free;       // PACK: the staged window has been read by every warp
    float2 rot[kRotMax];     // throughput loop: carrier rotation between a lane's consecutive windows
    sydr_trk_state cfgs;     // the channel's state as loaded (constants live here)
    CodeState sc;            // owned by warp 0
    CarrierState sk;         // owned by warp 1
    LoopConst K;
    int n_hist[2];           // samples of epoch e (index e & 1), for the carrier warp
    int rec_base;            // index of this call's first record in the channel's output row
    int status;
    int seg_ok;              // segment path usable for this channel (spacings on the half-chip lattice)
    int seg_q[3];            // tap offsets in half chips
    uint32_t segtab[NTAB];   // sign bytes of the three taps per lattice index
    long long pc[16];        // diagnostics
    long long tprev, tprev1;
    sydr_kaplan_state kcfg;  // Kaplan loop closure: the channel's configuration and state as loaded / to store
};

// TMA bulk copy of one CTA's window of the epoch starting at sample `a` (executed by one lane).
template <int DT, int VPC, class SH>
__device__ __forceinline__ void trk_prefetch(SH& sh, uint8_t* dst, const uint8_t* rec_base, long long rec_alloc,
                                             long long a, uint32_t rank, int Q, int buf) {
    constexpr int SPV = IqTraits<DT>::SPV, BPS = IqTraits<DT>::BPS;
    constexpr int C = SPV * VPC;
    const long long a0 = a & ~(long long)(SPV - 1);
    long long w0 = a0 + (long long)rank * Q * C;
    long long w1 = w0 + (long long)Q * C + kWinTail;
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

// Totals of the eight partial-sum components over every warp of every CTA of the cluster.
// Lane L adds the entries i = (L >> 3) mod 4 of component L & 7 in a fixed order (identical in
// every CTA, so the redundant loop closures agree bit for bit); on return every lane holds the
// total of component L & 7.
template <class SH>
__device__ __forceinline__ double gather_total(const SH& sh, int slot, int n_ent, int lane) {
    const int c = lane & 7;
    const double* g = &sh.gather[slot][0][c];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int i = lane >> 3;
    for (; i + 12 < n_ent; i += 16) {                 // four independent loads / adds in flight
        s0 += g[i * 8]; s1 += g[(i + 4) * 8]; s2 += g[(i + 8) * 8]; s3 += g[(i + 12) * 8];
    }
    for (; i < n_ent; i += 4) s0 += g[i * 8];
    double s = (s0 + s1) + (s2 + s3);
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    return s;
}

// acc + (int64)v in one instruction (sign extension and the 64-bit carry chain folded into IMAD.WIDE).
__device__ __forceinline__ long long mad_wide(int v, long long acc) {
    long long r;
    int one;
    asm("mov.s32 %0, 1;" : "=r"(one));
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(v), "r"(one), "l"(acc));
    return r;
}

// The same for the fixed-point exchange (integer IQ): entries are int32[8] warp totals scaled by a
// power of two; integer addition is associative, so the total does not depend on the order, the
// cluster shape or the warp count.  All loads are issued before the first addition.
template <int K, class SH>      // K entries per lane and pass (a lane walks entries lane/8, lane/8 + 4, ...)
__device__ __forceinline__ double gather_total_fixed(const SH& sh, int slot, int n_ent, int lane, double inv_scale) {
    const int c = lane & 7;
    const int* g = reinterpret_cast<const int*>(&sh.gather[slot][0][0]) + c;
    long long s = 0;
    for (int base = lane >> 3; base < n_ent; base += 4 * K) {
        int v[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i = base + 4 * k;
            v[k] = (i < n_ent) ? g[i * 8] : 0;
        }
        long long a0 = 0, a1 = 0;
#pragma unroll
        for (int k = 0; k < K; k += 2) { a0 = mad_wide(v[k], a0); a1 = mad_wide(v[k + 1], a1); }   // one IMAD.WIDE per entry
        s += a0 + a1;
    }
    s += __shfl_xor_sync(0xffffffffu, s, 8);
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    return (double)s * inv_scale;
}

// Warp 0: close the CODE loop of epoch e (DLL_NNEML + Borre filter + code NCO,
// channel_l1ca_borre.py:383-388, 422-429), store its share of the epoch record.
// The part of the code-loop update that does not need the correlator sums (evaluated while the
// all-gather is still in flight): epoch length as a double and the advanced code phase, L424.
struct CodePre { double n, rem_code; };
__device__ __forceinline__ CodePre code_pre(const CodeState& st) {
    CodePre p;
    p.n = i2d(st.n_req);
    p.rem_code = dadd(st.rem_code, dsub(dmul(p.n, st.code_step), (double)kCodeChips));   // L424
    return p;
}

template <class SH>
__device__ __forceinline__ void code_close(SH& sh, CodeState& st, int& status, double ck, sydr_trk_epoch* rec,
                                           int lane, const CodePre& pre) {
    const unsigned full = 0xffffffffu;
    // |E| (even lanes) and |L| (odd lanes)                                    tracking.py:126
    const double mx = __shfl_sync(full, ck, (lane & 1) ? 4 : 0), my = __shfl_sync(full, ck, (lane & 1) ? 5 : 1);
    const double mag = sqrt(dadd(dmul(mx, mx), dmul(my, my)));
    const double me = __shfl_sync(full, mag, 0), ml = __shfl_sync(full, mag, 1);
    const double errflag = __shfl_sync(full, ck, 6);
    const double code_err = ddiv(dsub(me, ml), dadd(me, ml));
    double nco_code = dmul(sh.K.dll_c1, dsub(code_err, st.nco_code_err));       // BorreLoopFilter
    nco_code = dadd(nco_code, dmul(sh.K.dll_c2, code_err));
    const long long e_start = st.cur;
    const double n = pre.n;
    st.code_freq = dsub(st.code_freq, nco_code);                                 // L422
    st.rem_code = pre.rem_code;                                                  // L424
    st.code_step = ddiv_by(st.code_freq, sh.K.fs, sh.K.inv_fs);                  // L425
    st.cur += st.n_req;                                                          // L428
    // The step moves by up to ~1e-6 relative from epoch to epoch (DLL noise): one Newton step would leave 1e-12, i.e.
    // 2e-8 sample at the end of an epoch, more than the 1e-9 sample band inside which a lattice crossing is re-decided
    // with the reference's own expression (seg_first_sample).  Two steps reach the rounding floor (< 1e-11 sample).
    st.inv_step = newton_rcp(st.code_step, newton_rcp(st.code_step, st.inv_step));
    st.n_req = ceil_to_int(ddiv_by(dsub((double)kCodeChips, st.rem_code), st.code_step, st.inv_step));  // L429
    st.nco_code_err = code_err;
    st.nco_code = nco_code;
    if (errflag != 0.0) status = SYDR_ERR_STATE;
    if (rec != nullptr) {                        // lanes 0-5 correlators, 6 dll, 9 code_freq, 10 code_err, 12-14
        double v = ck;
        v = (lane == 6) ? nco_code : v;
        v = (lane == 9) ? st.code_freq : v;
        v = (lane == 10) ? code_err : v;
        v = (lane == 12) ? (double)e_start : v;
        v = (lane == 13) ? n : v;
        v = (lane == 14) ? st.rem_code : v;
        if ((0x767Fu >> lane) & 1u) reinterpret_cast<double*>(rec)[lane] = v;
    }
}

// Warp 1: close the CARRIER loop of epoch e (remaining carrier phase, PLL_costa + Borre filter +
// carrier NCO, channel_l1ca_borre.py:364-365, 391-396, 423), store its share of the record.
// The remaining carrier phase after the epoch does not depend on the correlator sums either
// (L364-365, the old carrier frequency): evaluated while the all-gather is in flight.
template <class SH>
__device__ __forceinline__ double carrier_pre(const SH& sh, const CarrierState& st, int n_epoch) {
    const double n = i2d(n_epoch);
    // L364-365: rem' = (rem - ((fc*2)*pi*n)/fs) mod 2 pi   (Python float %: result in [0, 2 pi))
    const double twopi = 2.0 * kPi;
    double rc = dsub(st.rem_carrier, ddiv_by(dmul(dmul(dmul(st.carrier_freq, 2.0), kPi), n), sh.K.fs, sh.K.inv_fs));
    const double q = floor(rc * 0.15915494309189535);
    rc = fma(-q, twopi, rc);
    if (rc < 0.0) rc += twopi;
    if (rc >= twopi) rc -= twopi;
    return rc;
}

template <class SH>
__device__ __forceinline__ void carrier_close(SH& sh, CarrierState& st, double ck, double rc,
                                              sydr_trk_epoch* rec, int lane) {
    const unsigned full = 0xffffffffu;
    const double ip = __shfl_sync(full, ck, 2), qp = __shfl_sync(full, ck, 3);
    st.rem_carrier = rc;
    const double ph_err = ddiv_by(atan(ddiv(qp, ip)), kGpsPi * 2.0, 1.0 / (kGpsPi * 2.0));   // PLL_costa
    double nco_car = dmul(sh.K.pll_c1, dsub(ph_err, st.nco_carrier_err));        // BorreLoopFilter
    nco_car = dadd(nco_car, dmul(sh.K.pll_c2, ph_err));
    st.carrier_freq = dadd(st.carrier_freq, nco_car);                            // L423
    st.nco_carrier_err = ph_err;
    st.nco_carrier = nco_car;
    if (rec != nullptr) {                        // lane 7 pll, 8 carrier_freq, 11 carrier_err, 15 rem_carrier
        double v = nco_car;
        v = (lane == 8) ? st.carrier_freq : v;
        v = (lane == 11) ? ph_err : v;
        v = (lane == 15) ? rc : v;
        if ((0x8980u >> lane) & 1u) reinterpret_cast<double*>(rec)[lane] = v;
    }
}

// ---- Kaplan loop closure (warp 1) ----------------------------------------------------------------
// channel_l1ca_kaplan.py:342-619 after the correlators: discriminators, FLL-assisted PLL filter, lock
// indicators, C/N0, NCO, code lock / bit synchronisation, lock-state machine.  IEEE operations in the
// reference's order, no contraction: the same values as the Python floats (atan to 2 ulp).
struct KaplanRegs {                 // mutable members, registers of warp 1
    double ip_prev, qp_prev, fll_lock, pll_lock, cn0, pdpn, vel_memory, fll_bw, pll_bw;
    int accum_counter, lock_state, flags;
    long long code_counter;
    double atan_prev;               // atan(qp_prev / ip_prev): the previous epoch's value of the one arctangent per epoch
};
__device__ __forceinline__ double np_sign(double x) { return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : x); }   // 0 -> 0, nan -> nan

// remainingCarrier after the epoch, channel_l1ca_kaplan.py:529-530 (GPS value of two pi, Python float %)
template <class SH>
__device__ __forceinline__ double carrier_pre_kaplan(const SH& sh, const CarrierState& st, int n_epoch) {
    const double twopi = kGpsPi * 2.0;
    double rc = dsub(st.rem_carrier, __ddiv_rn(dmul(dmul(st.carrier_freq, twopi), i2d(n_epoch)), sh.K.fs));
    const double q = floor(rc * (1.0 / (kGpsPi * 2.0)));
    rc = fma(-q, twopi, rc);
    if (rc < 0.0) rc += twopi;
    if (rc >= twopi) rc -= twopi;
    return rc;
}

// FLL_ATAN with the two arctangents handed in: atan(qp/ip) of this epoch also serves PLL_costa, and is the
// next epoch's atan(qpPrev/ipPrev) (same operands, same value): one arctangent per epoch instead of three.
__device__ __forceinline__ double fll_atan(double at_now, double at_prev) {                     // tracking.py:156-176
    double e = at_now - at_prev;
    if (isnan(e)) e = 0.0;
    const double half_pi = kGpsPi / 2.0;
    if (e >= half_pi) e = dsub(e, kGpsPi);
    else if (e <= -half_pi) e = dadd(e, kGpsPi);
    return ddiv(ddiv(e, 1e-3), kGpsPi * 2.0);
}

template <class SH>
__device__ __forceinline__ void carrier_close_kaplan(SH& sh, CarrierState& st, KaplanRegs& k, double ck, double rc,
                                                     sydr_trk_epoch* rec, sydr_kaplan_epoch* krec, int lane) {
    const unsigned full = 0xffffffffu;
    const double ip = __shfl_sync(full, ck, 2), qp = __shfl_sync(full, ck, 3);
    const sydr_kaplan_state& c = sh.kcfg;
    // runCorrelators: the 20 ms accumulator counter (L387-393)
    if (k.accum_counter == 20) k.accum_counter = 0;
    k.accum_counter += 1;
    // runDiscriminators (L407-432)
    const double at_now = atan(__ddiv_rn(qp, ip));
    double fll = 0.0, pll = 0.0;
    if (k.lock_state == 1) {
        if (k.code_counter > 1) fll = fll_atan(at_now, k.atan_prev);
    } else {
        fll = fll_atan(at_now, k.atan_prev);
        pll = ddiv(at_now, kGpsPi * 2.0);                                                // PLL_costa
    }
    k.atan_prev = at_now;
    // FLLassistedPLL_2ndOrder (tracking.py:246-279), w0f = B_fll / 0.25, w0p = B_pll / 0.53
    const double w0f = dmul(k.fll_bw, 4.0), w0p = ddiv(k.pll_bw, 0.53);                   // x / 0.25 = 4 x exactly
    const double update = dmul(dadd(dmul(pll, dmul(w0p, w0p)), dmul(fll, w0f)), dmul(1.0, 1e-3));
    double cerr = dadd(update, k.vel_memory);
    k.vel_memory = update;
    cerr = dadd(cerr, dmul(dmul(pll, 1.414), w0p));
    // runLoopIndicators (L460-508)
    if (k.code_counter != 0) {
        const double i2 = dmul(ip, ip), q2 = dmul(qp, qp);
        double lock = dsub(dmul(ip, k.ip_prev), dmul(qp, k.qp_prev));
        lock = dmul(lock, np_sign(dadd(dmul(ip, k.ip_prev), dmul(qp, k.qp_prev))));
        lock = fabs(ddiv(lock, dadd(i2, q2)));
        k.fll_lock = dadd(dmul(dsub(1.0, 0.005), k.fll_lock), dmul(0.005, lock));
        if (k.lock_state > 1)
            k.pll_lock = dadd(dmul(dsub(1.0, 0.005), k.pll_lock), dmul(0.005, ddiv(dsub(i2, q2), dadd(i2, q2))));
        const double d = dsub(fabs(ip), fabs(qp));
        k.pdpn = dadd(k.pdpn, __ddiv_rn(dadd(i2, q2), dmul(d, d)));       // (|ip| - |qp|)^2 may be 0: IEEE division
        if (k.accum_counter == 20) {                                                     // CN0_Beaulieu, alpha = 0.1
            const double lam = __ddiv_rn(1.0, __ddiv_rn(k.pdpn, 20.0));
            const double neu = dmul(lam, __ddiv_rn(1.0, dmul(20.0, 1e-3)));
            k.cn0 = dadd(dmul(dsub(1.0, 0.1), k.cn0), dmul(0.1, neu));
            k.pdpn = 0.0;
        }
    }
    // postTrackingUpdate (L512-541): carrier part
    k.code_counter += 1;
    st.rem_carrier = rc;
    st.carrier_freq = dadd(st.carrier_freq, cerr);
    st.nco_carrier_err = pll;
    st.nco_carrier = cerr;
    // trackingStateUpdate (L545-619)
    if (k.lock_state != 1 && k.cn0 > c.dll_threshold && !(k.flags & 1)) k.flags |= 1;
    else if (k.cn0 < c.dll_threshold && (k.flags & 1)) k.flags ^= 1;
    if ((k.flags & 1) && !(k.flags & 2) && np_sign(k.ip_prev) != np_sign(ip)) {
        k.flags |= 2;
        k.accum_counter = 1;
        k.pdpn = 0.0;
    }
    k.ip_prev = ip;
    k.qp_prev = qp;
    if (k.lock_state != 3 && k.fll_lock >= c.fll_thr_narrow && k.pll_lock >= c.pll_thr_narrow) {
        k.lock_state = 3; k.fll_bw = c.fll_bw_narrow; k.pll_bw = c.pll_bw_narrow;
    } else if (k.lock_state != 2 && k.fll_lock >= c.fll_thr_wide && k.fll_lock < c.fll_thr_narrow) {
        k.lock_state = 2; k.fll_bw = c.fll_bw_wide; k.pll_bw = c.pll_bw_wide;
    } else if (k.lock_state != 1 && k.fll_lock <= c.fll_thr_wide) {
        k.lock_state = 1; k.fll_bw = c.fll_bw_pullin; k.pll_bw = 0.0;
    }
    if (rec != nullptr) {                        // lane 7 carrier filter output, 8 carrier_freq, 11 PLL discriminator, 15 rem_carrier
        double v = cerr;
        v = (lane == 8) ? st.carrier_freq : v;
        v = (lane == 11) ? pll : v;
        v = (lane == 15) ? rc : v;
        if ((0x8980u >> lane) & 1u) reinterpret_cast<double*>(rec)[lane] = v;
        if (lane >= 16 && lane < 20) {
            double w = fll;
            w = (lane == 17) ? k.cn0 : w;
            w = (lane == 18) ? k.fll_lock : w;
            w = (lane == 19) ? k.pll_lock : w;
            reinterpret_cast<double*>(krec)[lane - 16] = w;
        }
        if (lane == 20) { krec->lock_state = k.lock_state; krec->flags = k.flags; }
    }
}


// trkm.cu: the prefix-moment formulation (throughput shape).  `group` consecutive channels share a CTA and a recording.
int launch_trkm(const TrkParams& P, int n_channels, int rec_channels, int group, cudaStream_t s);

}  // namespace sydr
