// K-NAV: bit synchronisation and navigation-bit accumulation on the device.
//
// Reference semantics (file:line in /root/reference):
//   bit-sync test                     sydr/channel/channel_l1ca_borre.py:398-413
//   prompt history / resetPrompt      sydr/channel/channel_l1ca_borre.py:367-373, 577-591, 626-627
//   20-epoch prompt sum -> bit        sydr/channel/channel_l1ca_borre.py:455-491
//   Prompt2Bit                        sydr/dsp/decoding.py:16-27
//
// The closed-loop tracking kernel leaves one 128-byte record per channel-epoch in HBM.  The
// step the reference performs next on every one of them (per channel process, per ms) is a
// scalar state machine: wait MIN_CONVERGENCE_TIME epochs, declare bit synchronisation at the
// first sign change of the prompt in-phase sum, then add 20 prompts per bit and take the sign.
// Doing it here means 50 bit/s per channel cross PCIe instead of 128 kB/s.
//
// One warp owns one channel.  Before synchronisation the lanes test 32 epochs per round and a
// ballot picks the first sign change; after it every lane owns one bit and adds its 20 prompts
// in the reference's order (sequential FP64 adds starting from 0.0), so the sums - not only
// their signs - equal the reference's navPromptSum bit for bit (nav.npz: every tick of the live reference channel).
// One prompt per EPOCH goes into a bit.  The reference calls runDecoding on every millisecond TICK; on a tick without a
// completed epoch (the channel's backlog is below one epoch) it adds the previous epoch's prompt again.  Such ticks do
// not occur while the epoch length stays within a sample of 1 ms (all the recordings tested); a recording whose code
// Doppler makes them occur would see that one prompt counted twice by the reference and once here.
//
// Quirk kept on purpose: on the synchronisation epoch resetPrompt() zeroes nbPrompt *before*
// runDecoding reads correlatorsBuffer[nbPrompt - 1], so the first addend of the first bit is
// row 19 of the 20-row prompt history (the prompt of the latest epoch whose index is 19 modulo
// 20), not the synchronisation epoch's own prompt.
#include <string.h>

#include "common.cuh"

namespace sydr {

constexpr int kMsPerBit = 20;                 // LNAV_MS_PER_BIT, sydr/utils/constants.py
constexpr long long kMinConvergence = 100;    // ChannelL1CA.MIN_CONVERGENCE_TIME (L30)
constexpr int kIPromptSlot = 2;               // corr[2] of sydr_trk_epoch

__device__ __forceinline__ int np_sign(double x) { return (x > 0.0) - (x < 0.0); }   // np.sign of a finite value

// kflags != NULL selects the Kaplan channel's rule (channel_l1ca_kaplan.py:555-566, 725-758): BIT_SYNC is the
// flag its trackingStateUpdate raised (bit 2 of the device records' `flags`), and the 20-epoch sums start with
// the synchronisation epoch's own prompt.
__global__ void __launch_bounds__(32) nav_bits_kernel(const sydr_trk_epoch* __restrict__ epochs, int max_epochs,
                                                      const sydr_kaplan_epoch* __restrict__ kepochs,
                                                      const int* __restrict__ nepochs, int first,
                                                      sydr_nav_state* __restrict__ nav, int8_t* __restrict__ bits,
                                                      double* __restrict__ bit_sums, int max_bits,
                                                      int* __restrict__ nbits_out) {
    const int ch = blockIdx.x, lane = threadIdx.x;
    const unsigned full = 0xffffffffu;
    const double* row = reinterpret_cast<const double*>(epochs + (long long)ch * max_epochs + first) + kIPromptSlot;
    auto ip = [&](int k) { return row[(long long)k * (sizeof(sydr_trk_epoch) / sizeof(double))]; };
    const int n = max(nepochs[ch] - first, 0);
    sydr_nav_state s = nav[ch];
    int k = 0;                                 // next local epoch to consume
    int produced = 0;

    if (s.sync_epoch < 0) {
        // ---- L401-407: CODE_LOCK (set by the first epoch) and codeCounter > 100 and a sign change
        int ks = -1;
        for (int base = 0; base < n && ks < 0; base += 32) {
            const int kk = base + lane;
            bool hit = false;
            if (kk < n) {
                if (kepochs != nullptr) {
                    hit = (kepochs[(long long)ch * max_epochs + first + kk].flags & 2) != 0;
                } else {
                    const long long cc = s.code_counter + kk;             // codeCounter when epoch kk is ingested
                    const double prev = (kk == 0) ? s.prev_iprompt : ip(kk - 1);
                    hit = (cc >= 1) && (cc > kMinConvergence) && (np_sign(prev) != np_sign(ip(kk)));
                }
            }
            const unsigned m = __ballot_sync(full, hit);
            if (m) ks = base + __ffs(m) - 1;
        }
        const int upto = (ks >= 0) ? ks : n - 1;                          // last epoch written to the prompt history
        if (upto >= 0) {
            // row 19 of correlatorsBuffer: the latest epoch <= upto with (code_counter + k) % 20 == 19
            const long long g = s.code_counter + upto;
            const long long k19 = upto - ((g + 1) % kMsPerBit);
            if (k19 >= 0) s.row19 = ip((int)k19);
        }
        if (ks < 0) {
            if (n > 0) s.prev_iprompt = ip(n - 1);
            s.code_counter += n;
            if (lane == 0) { nav[ch] = s; nbits_out[ch] = 0; }
            return;
        }
        s.sync_epoch = s.code_counter + ks;
        if (kepochs != nullptr) {                                         // Kaplan: the sync epoch's own prompt is the first addend
            s.nav_sum = 0.0;
            s.nav_count = 0;
            k = ks;
        } else {
            s.nav_sum = s.row19;                                          // L471: correlatorsBuffer[-1]
            s.nav_count = 1;
            k = ks + 1;
        }
    }

    // ---- L471-483: 20 prompts per bit.  Bit j of this call ends at local epoch e_j (exclusive).
    const int to_first = kMsPerBit - s.nav_count;                         // epochs the pending bit still needs
    const int avail = n - k;
    const int nb = (avail >= to_first) ? 1 + (avail - to_first) / kMsPerBit : 0;
    for (int b0 = 0; b0 < nb; b0 += 32) {
        const int j = b0 + lane;
        if (j < nb) {
            double sum = (j == 0) ? s.nav_sum : 0.0;
            const int lo = (j == 0) ? k : k + to_first + (j - 1) * kMsPerBit;
            const int cnt = (j == 0) ? to_first : kMsPerBit;
            for (int u = 0; u < cnt; ++u) sum += ip(lo + u);              // same order as navPromptSum +=
            if (j < max_bits) {
                bits[(long long)ch * max_bits + j] = (sum > 0.0) ? 1 : 0; // Prompt2Bit, decoding.py:27
                if (bit_sums) bit_sums[(long long)ch * max_bits + j] = sum;
            }
        }
    }
    produced = min(nb, max_bits);
    // the bit still being accumulated when the records end
    int rest_lo = k, rest_cnt = avail;
    double rest = s.nav_sum;
    if (nb > 0) {
        rest_lo = k + to_first + (nb - 1) * kMsPerBit;
        rest_cnt = n - rest_lo;
        rest = 0.0;
        s.nav_count = 0;
    }
    if (lane == 0) {
        for (int u = 0; u < rest_cnt; ++u) rest += ip(rest_lo + u);
        s.nav_sum = rest;
        s.nav_count += rest_cnt;
        s.n_bits += nb;
        if (n > 0) s.prev_iprompt = ip(n - 1);
        s.code_counter += n;
        nav[ch] = s;
        nbits_out[ch] = produced;
    }
}

}  // namespace sydr

using namespace sydr;

extern "C" {

int sydr_nav_state_init(sydr_nav_state* h_state) {
    SYDR_REQUIRE(h_state != nullptr, SYDR_ERR_ARG, "state pointer is NULL");
    memset(h_state, 0, sizeof(*h_state));
    h_state->sync_epoch = -1;
    return SYDR_OK;
}

int sydr_nav_bits(const sydr_trk_epoch* d_epochs, int max_epochs, const int* d_nepochs, int first_epoch,
                  sydr_nav_state* d_nav, int n_channels, signed char* d_bits, double* d_bit_sums, int max_bits,
                  int* d_nbits, void* stream) {
    SYDR_REQUIRE(d_epochs && d_nepochs && d_nav && d_bits && d_nbits, SYDR_ERR_ARG, "NULL pointer");
    SYDR_REQUIRE(max_epochs > 0 && max_bits > 0 && first_epoch >= 0, SYDR_ERR_ARG, "sizes must be positive");
    if (n_channels <= 0) return SYDR_OK;
    nav_bits_kernel<<<n_channels, 32, 0, (cudaStream_t)stream>>>(d_epochs, max_epochs, nullptr, d_nepochs, first_epoch, d_nav,
                                                                 reinterpret_cast<int8_t*>(d_bits), d_bit_sums,
                                                                 max_bits, d_nbits);
    count_launch();
    SYDR_CUDA_CHECK(cudaGetLastError());
    return SYDR_OK;
}

}  // extern "C"

extern "C" int sydr_nav_bits_kaplan(const sydr_trk_epoch* d_epochs, const sydr_kaplan_epoch* d_kepochs, int max_epochs,
                                    const int* d_nepochs, int first_epoch, sydr_nav_state* d_nav, int n_channels,
                                    signed char* d_bits, double* d_bit_sums, int max_bits, int* d_nbits, void* stream) {
    SYDR_REQUIRE(d_epochs && d_kepochs && d_nepochs && d_nav && d_bits && d_nbits, SYDR_ERR_ARG, "NULL pointer");
    SYDR_REQUIRE(max_epochs > 0 && max_bits > 0 && first_epoch >= 0, SYDR_ERR_ARG, "sizes must be positive");
    if (n_channels <= 0) return SYDR_OK;
    sydr::nav_bits_kernel<<<n_channels, 32, 0, (cudaStream_t)stream>>>(d_epochs, max_epochs, d_kepochs, d_nepochs, first_epoch,
                                                                       d_nav, reinterpret_cast<int8_t*>(d_bits), d_bit_sums,
                                                                       max_bits, d_nbits);
    sydr::count_launch();
    SYDR_CUDA_CHECK(cudaGetLastError());
    return SYDR_OK;
}

// ------------------------------------------------------------------------------------------
// K-HAND: acquisition -> tracking hand-off on the device.
//
// Reference semantics: sydr/channel/channel_l1ca_borre.py:301-311 (per channel, host scalars):
//   doppler = -((-dopplerRange) + dopplerSteps * freq_idx);  carrierFrequency = IF + doppler;
//   currentSample += acq_requiredSamples - track_requiredSamples + codeOffset + 1.
// Channel selection is the receiver's (best peak ratios above the threshold, at most
// max_channels, then ordered by PRN).  Running it here removes the only host round trip between
// the acquisition and the tracking launch: the tracking kernel is enqueued right behind.
// ------------------------------------------------------------------------------------------
namespace sydr {

__global__ void __launch_bounds__(64) acq_handoff_kernel(const sydr_acq_peak* __restrict__ peaks, int n_prn,
                                                         double inter_freq, double doppler_range, double doppler_step,
                                                         long long required, long long track_required,
                                                         long long current_sample, double threshold,
                                                         const sydr_trk_state* __restrict__ tmpl, long long iq_len,
                                                         sydr_trk_state* __restrict__ states, int max_channels,
                                                         int* __restrict__ n_selected) {
    __shared__ float s_ratio[64];
    __shared__ int s_prn[64];
    __shared__ int s_sel[64];
    const int i = threadIdx.x;
    sydr_acq_peak pk = {};
    bool pass = false;
    if (i < n_prn) {
        pk = peaks[i];
        pass = (double)pk.ratio > threshold;
    }
    s_ratio[i] = pass ? pk.ratio : -1.f;
    s_prn[i] = pk.prn;
    __syncthreads();
    // rank among the passing peaks: larger ratio first, lower slot first on ties (stable argsort of -ratio)
    int rank = 0;
    if (pass)
        for (int j = 0; j < n_prn; ++j) rank += (s_ratio[j] > pk.ratio) || (s_ratio[j] == pk.ratio && j < i);
    const bool sel = pass && rank < max_channels;
    s_sel[i] = sel ? 1 : 0;
    __syncthreads();
    int pos = 0, total = 0;
    for (int j = 0; j < n_prn; ++j) {
        total += s_sel[j];
        pos += s_sel[j] && (s_prn[j] < pk.prn || (s_prn[j] == pk.prn && j < i));
    }
    if (sel) {
        sydr_trk_state st = *tmpl;
        const double doppler = -__dadd_rn(-doppler_range, __dmul_rn(doppler_step, (double)pk.freq_idx));   // L301
        st.prn = pk.prn;
        st.carrier_freq = __dadd_rn(inter_freq, doppler);                                                  // L303
        st.cur = current_sample + required - track_required + (long long)pk.code_idx + 1;                  // L306-311
        // (iq_base stays the template's: where this recording lies inside the tracking launch's buffer, ColdStartBatch)
        st.iq_len = iq_len;
        st.epochs_done = 0;
        st.status = 0;
        states[pos] = st;
    }
    if (i >= total && i < max_channels) {        // unused slots: idle (the tracking kernel leaves them untouched)
        sydr_trk_state st = *tmpl;
        st.prn = 1;
        st.status = 1;
        st.iq_len = 0;
        states[i] = st;
    }
    if (i == 0 && n_selected) *n_selected = total;
}

}  // namespace sydr

extern "C" int sydr_acq_handoff(const sydr_acq_peak* d_peaks, int n_prn, double inter_freq, double doppler_range,
                                double doppler_step, long long required_samples, long long track_required,
                                long long current_sample, double threshold, const sydr_trk_state* d_template,
                                long long iq_len, sydr_trk_state* d_states, int max_channels, int* d_n_selected,
                                void* stream) {
    SYDR_REQUIRE(d_peaks && d_template && d_states, SYDR_ERR_ARG, "NULL pointer");
    SYDR_REQUIRE(n_prn >= 1 && n_prn <= 64 && max_channels >= 1 && max_channels <= 64, SYDR_ERR_ARG,
                 "n_prn and max_channels must be in [1, 64]");
    sydr::acq_handoff_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(d_peaks, n_prn, inter_freq, doppler_range, doppler_step,
                                                                 required_samples, track_required, current_sample,
                                                                 threshold, d_template, iq_len, d_states, max_channels,
                                                                 d_n_selected);
    sydr::count_launch();
    SYDR_CUDA_CHECK(cudaGetLastError());
    return SYDR_OK;
}
