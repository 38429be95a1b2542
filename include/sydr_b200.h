/*
 * libsydr_b200.so -- C ABI of the B200-native SyDR DSP hot paths.
 *
 * Drop-in boundary for the two data-parallel paths of aproposorg/sydr:
 *   PCPS acquisition   sydr/dsp/acquisition.py:9-115      (live NumPy path)
 *                      sydr/c_functions/acquisition.c:82-244 (legacy C ABI)
 *   E/P/L tracking     sydr/dsp/tracking.py:92-186, sydr/channel/channel_l1ca_borre.py:333-451
 *                      sydr/c_functions/tracking.c:31-212   (legacy C ABI)
 *
 * Plain C: pointers, sizes and scalars only; no torch/CUDA types.  "d_" pointers are
 * CUDA device pointers (e.g. torch.Tensor.data_ptr()), "h_" pointers are host pointers.
 * `stream` is a cudaStream_t passed as void* (NULL = default stream).  Every entry point
 * returns SYDR_OK (0) or a negative error code; sydr_last_error() gives the message.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * SYDR_ERR_CUDA.
 *
 * The Python binding is sydr_b200/_lib.py (ctypes); the binding a reference maintainer
 * would add is shown in INTEGRATION.md.
 */
#ifndef SYDR_B200_H
#define SYDR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SYDR_ABI_VERSION 1

#define SYDR_OK               0
#define SYDR_ERR_CUDA        -1
#define SYDR_ERR_ARG         -2
#define SYDR_ERR_UNSUPPORTED -3
#define SYDR_ERR_STATE       -4

/* IQ sample formats.  I8/I16 are the interleaved I,Q file formats RFSignal.readFile parses
 * (sydr/signal/rfsignal.py:107-130) and stay integer all the way into the kernels; F32/F64
 * are interleaved complex64/complex128, what the reference's Python functions receive. */
#define SYDR_IQ_I8   0
#define SYDR_IQ_I16  1
#define SYDR_IQ_F32  2
#define SYDR_IQ_F64  3

/* ------------------------------------------------------------------ library --------- */
int         sydr_abi_version(void);
const char* sydr_last_error(void);
void        sydr_clear_error(void);             /* the legacy void entry points (section 9) report through the message only */
int         sydr_device_count(void);            /* 0 when no CUDA device is usable      */
int         sydr_set_device(int device);
/* Measured FP32 FMA-chain peak of the current device (roofline denominator). */
int         sydr_measure_fp32_peak(double* h_tflops, double* h_sm_clock_mhz);
/* Measured FP64 FMA throughput and dependent-FMA latencies (cycles) of the current device. */
int         sydr_measure_fp64_peak(double* h_tflops, double* h_dfma_latency_cycles, double* h_ffma_latency_cycles);
/* Kernel-launch counter (all kernels launched by this library since load / reset). */
long long   sydr_launch_count(void);
void        sydr_reset_launch_count(void);

/* ------------------------------------------------------------------ code tables ------ */
/* GenerateGPSGoldCode (sydr/signal/gnsssignal.py:9-31, sydr/signal/ca.py:70-112):
 * 1023 chips of +-1.0 for PRN 1..37, generated on the device. */
int sydr_ca_code(int prn, double* h_code1023);
/* UpsampleCode + conj(fft(.)) (sydr/channel/channel_l1ca_borre.py:281-282,
 * sydr/signal/gnsssignal.py:35-58): n_code interleaved complex128 values.
 * Replaces setSatellite of sydr/c_functions/acquisition.c:82-97. */
int sydr_code_spectrum(int prn, double fs, double* h_spectrum_c128, long long n_code);

/* ------------------------------------------------------------------ acquisition ------ */
typedef struct sydr_acq_plan sydr_acq_plan;      /* opaque */

typedef struct {
    int32_t prn;
    int32_t freq_idx;      /* Doppler-bin row of the global maximum (np.argmax C order)   */
    int32_t code_idx;      /* code-phase sample of the global maximum                     */
    float   peak1;         /* correlationMap[freq_idx, code_idx]                          */
    float   peak2;         /* second peak, same row, +-chip excluded (acquisition.py:103) */
    float   ratio;         /* peak1 / peak2 = TwoCorrelationPeakComparison metric         */
} sydr_acq_peak;           /* 24 bytes: the record all-gathered over NCCL                  */

typedef struct {
    float   peak1;
    int32_t code_idx;
    float   peak2;
    int32_t reserved;
} sydr_acq_row;            /* per (PRN, Doppler-bin) row summary, 16 bytes                 */

/* Plan = everything PCPS derives from its scalar arguments (acquisition.py:9-34) plus the
 * precomputed conj(FFT(code)) tables of the PRNs to search.  bins = arange(-range,
 * range+1, step).  [bin_lo, bin_hi) selects a sub-range of Doppler rows (multi-GPU
 * PRN x bin sharding); pass 0, -1 for all rows. */
int sydr_acq_plan_create(double fs, double inter_freq, double doppler_range, double doppler_step,
                         int coh, int noncoh, const int* h_prns, int n_prn,
                         int bin_lo, int bin_hi, sydr_acq_plan** out_plan);
int sydr_acq_plan_destroy(sydr_acq_plan* plan);
int sydr_acq_plan_info(const sydr_acq_plan* plan, int* n_code, int* n_bins_total, int* n_rows_local,
                       int* samples_per_chip, long long* required_samples);
/* Replace a PRN's code spectrum by a caller-supplied one (the `codeFFT` argument of PCPS,
 * acquisition.py:9): n_code interleaved complex128 host values. */
int sydr_acq_plan_set_spectrum(sydr_acq_plan* plan, int prn_slot, const double* h_spectrum_c128);

/* Batched PCPS + TwoCorrelationPeakComparison for every PRN of the plan on one dwell of
 * coh*noncoh*n_code samples starting at d_iq (device).  Outputs (device pointers, any may
 * be NULL): d_peaks[n_prn], d_rows[n_prn * n_rows_local], d_maps[n_prn * n_rows_local *
 * n_code] float32 correlation maps (row-major, the reference's (bins, N) layout).
 * Replaces PCPS + twoCorrelationPeakComparison of sydr/c_functions/acquisition.c:109-244. */
int sydr_acq_run(sydr_acq_plan* plan, const void* d_iq, int iq_dtype, long long n_samples,
                 sydr_acq_peak* d_peaks, sydr_acq_row* d_rows, float* d_maps, void* stream);
/* Reduce row summaries (possibly gathered from several GPUs) to peak records:
 * rows laid out [n_prn][n_bins]; first-maximum-in-C-order tie-break. Host-side helper for
 * the multi-GPU path; takes host pointers. */
int sydr_acq_reduce_rows(const sydr_acq_row* h_rows, const int* h_prns, int n_prn, int n_bins,
                         sydr_acq_peak* h_peaks);

/* TwoCorrelationPeakComparison on a caller-supplied map (acquisition.py:78-115): float64
 * host map (n_bins x n_code).  Runs the same device peak search as sydr_acq_run. */
int sydr_peak_compare(const double* h_map, int n_bins, int n_code, int samples_per_chip,
                      int* h_freq_idx, int* h_code_idx, double* h_ratio);

/* ------------------------------------------------------------------ tracking --------- */
/* One open-loop E/P/L correlation = one call of EPL (sydr/dsp/tracking.py:92-116). */
typedef struct {
    int64_t start;          /* first sample (index into d_iq, in complex samples)          */
    int32_t n;              /* nbSamples                                                  */
    int32_t prn;
    double  carrier_freq;   /* carrierFrequency                                           */
    double  rem_carrier;    /* remainingCarrier                                           */
    double  rem_code;       /* remainingCode                                              */
    double  code_step;      /* codeStep                                                   */
    double  spacing[3];     /* correlatorsSpacing (early, prompt, late)                   */
} sydr_epl_args;            /* 72 bytes */

/* n_calls independent EPL evaluations; d_out receives 6 doubles each
 * [IE, QE, IP, QP, IL, QL].  d_iq must hold iq_len complex samples (16-byte aligned,
 * with at least 64 bytes of readable padding after the last sample).  A call whose window
 * [start, start+n) leaves the recording, whose n <= 0 or whose PRN has no code reads nothing
 * and returns six NaNs (the reference's numpy slice would come up short and raise). */
int sydr_epl_batch(const void* d_iq, int iq_dtype, long long iq_len, double fs,
                   const sydr_epl_args* d_args, int n_calls, double* d_out, void* stream);

/* Per-channel closed-loop state: the NCO / loop-filter members of ChannelL1CA
 * (sydr/channel/channel_l1ca_borre.py:110-120, 231-251). */
typedef struct {
    int64_t iq_base;        /* sample offset of this channel's recording inside d_iq        */
    int64_t iq_len;         /* samples of that recording currently valid on the device      */
    int64_t cur;            /* currentSample: first sample of the next epoch (rec-relative) */
    int64_t n_req;          /* track_requiredSamples                                        */
    int64_t epochs_done;    /* epochs processed so far (all launches)                       */
    int32_t prn;            /* 1..37; anything else aborts the channel (status SYDR_ERR_STATE) */
    int32_t status;         /* 0 ok; <0 = channel aborted (SYDR_ERR_STATE); >0 = idle slot:
                               the kernel leaves the channel untouched                      */
    double  carrier_freq;   /* carrierFrequency                                             */
    double  code_freq;      /* codeFrequency                                                */
    double  code_step;      /* codeStep                                                     */
    double  rem_carrier;    /* NCO_remainingCarrier                                         */
    double  rem_code;       /* NCO_remainingCode                                            */
    double  nco_code, nco_code_err;        /* NCO_code, NCO_codeError                       */
    double  nco_carrier, nco_carrier_err;  /* NCO_carrier, NCO_carrierError                 */
    double  dll_tau1, dll_tau2, dll_pdi;   /* LoopFiltersCoefficients (tracking.py:39-61)   */
    double  pll_tau1, pll_tau2, pll_pdi;
    double  spacing[3];
} sydr_trk_state;           /* 192 bytes */

typedef struct {
    double corr[6];         /* i_early, q_early, i_prompt, q_prompt, i_late, q_late         */
    double dll, pll;        /* NCO_code, NCO_carrier after this epoch                       */
    double carrier_freq;    /* carrierFrequency after update                                */
    double code_freq;       /* codeFrequency after update                                   */
    double code_err;        /* DLL discriminator (code_frequency_error)                     */
    double carrier_err;     /* PLL discriminator (carrier_frequency_error)                  */
    double start;           /* epoch start sample (rec-relative)                            */
    double n;               /* samples in the epoch                                         */
    double rem_code;        /* NCO_remainingCode after update                               */
    double rem_carrier;     /* NCO_remainingCarrier after update                            */
} sydr_trk_epoch;           /* 128 bytes */

/* Launch configuration of the closed-loop kernel.  cluster = CTAs cooperating on one
 * channel (1,2,4,8); threads = threads per CTA (multiple of 32, 64..384).  0 = auto. */
typedef struct {
    int32_t cluster;
    int32_t threads;
    int32_t use_tma;        /* 1 = cp.async.bulk staged smem windows (default), 0 = LDG     */
    int32_t append;         /* 1 = records go to d_out[ch][state.epochs_done ...] so that
                               successive calls on a growing recording fill one array;
                               d_nepochs then reports the cumulative count                  */
    int64_t iq_len;         /* > 0: valid samples per recording for this call (overrides
                               the states' iq_len; streaming upload)                       */
    double  min_tap_gap;    /* smallest distance, in chips, between the chip-boundary positions
                               of two correlators (0 = 0.5, the -0.5/0/+0.5 spacing); sizes
                               the per-thread chunk for the split-sum path                  */
    int64_t iq_base;        /* with use_iq_base != 0: sample offset of the recording inside
                               d_iq for this call, overriding the states' iq_base.  May be
                               negative: a sliding window that holds samples [w0, w0 + len)
                               of the recording at d_iq passes iq_base = -w0 (w0 a multiple
                               of 8 samples) and iq_len = w0 + len, so that `cur` and the
                               records' `start` stay recording-relative (streaming ingest)   */
    int32_t use_iq_base;
    int32_t dense;          /* 1 = register-lean instantiation of the latency kernel (3 CTAs per SM): a few
                               per cent slower alone, denser when several launches share the GPU;
                               2 = PACK, the throughput shape of the staged kernel (with cluster = 1, use_tma = 1,
                               integer IQ whose 1 ms window fits 108 KB): one CTA and ONE staged window per
                               channel, two channels per SM, one's loop closure under the other's correlation
                               (falls back to 1 where the shape does not apply)                        */
    int32_t kernel;         /* 0 = automatic; 1 = prefix-moment kernel (throughput shape: the samples of a recording are
                               turned into prefix moments once, every channel gathers from them; int16 IQ, Borre loops);
                               2 = per-channel kernels only                                          */
    int32_t group;          /* prefix-moment kernel: consecutive channels per CTA (1 .. 4, same recording); 0 = the
                               smallest group whose CTAs fit one wave of the SMs                     */
    int32_t rec_channels;   /* prefix-moment kernel: channel slots per recording -- channels
                               [r*rec_channels, (r+1)*rec_channels) share recording r; 0 = all the channels are on
                               one recording (a channel that is not on its slot's recording is served by the
                               per-channel kernel)                                                   */
    int32_t reserved;
} sydr_trk_config;

/* Closed-loop Borre tracking (runTracking, channel_l1ca_borre.py:333-451: EPL +
 * DLL_NNEML + PLL_costa + BorreLoopFilter + NCO update) for n_channels channels, each
 * advancing epoch by epoch until fewer than n_req samples remain (cur + n_req > iq_len)
 * or max_epochs epochs were written.  d_out is [n_channels][max_epochs]; d_nepochs[ch]
 * receives the number written by this call.  States are updated in place. */
int sydr_trk_run(const void* d_iq, int iq_dtype, long long iq_alloc_samples, double fs,
                 sydr_trk_state* d_states, int n_channels,
                 sydr_trk_epoch* d_out, int max_epochs, int* d_nepochs,
                 const sydr_trk_config* cfg, void* stream);

/* Diagnostics: per-phase clock64 counters of the loop-closing thread, d_buf = n_channels*16
 * int64 on the device ([0..7] = epoch constants, barrier, window wait, correlate, warp
 * reduction + send, gather wait, loop closure, gather totals; [15] = epochs).  NULL = off. */
int sydr_trk_profile_buffer(long long* d_buf);

/* Diagnostics of the prefix-moment kernel (cfg.kernel = 1): debug flags (bit 1: skip the exact re-evaluation of ambiguous
 * samples); sydr_trkm_shape is kept for the ABI and does nothing (the launch shape is fixed). */
int sydr_trkm_debug(int flags);
int sydr_trkm_shape(int correlating_warps, int producer_warps);

/* Diagnostics / tests: which correlator formulation K-TRK uses.  0 (default) = automatic: the
 * half-chip segment path wherever the spacings are multiples of half a chip and the sampling
 * rate fits, else the chunk paths; 1 = chunk paths only.  Results agree to rounding. */
int sydr_trk_set_mode(int mode);

/* Host helper: initial state exactly as ChannelL1CA leaves it after acquisition
 * (channel_l1ca_borre.py:110-120, 250-251, 301-311). */
int sydr_trk_state_init(sydr_trk_state* h_state, int prn, double fs, double carrier_freq,
                        long long start_sample,
                        double dll_bw, double dll_damp, double dll_gain, double dll_pdi,
                        double pll_bw, double pll_damp, double pll_gain, double pll_pdi,
                        double sp_early, double sp_prompt, double sp_late);

/* int8/int16/complex128 -> complex64 conversion on the device (K-CVT). */
int sydr_convert_to_f32(const void* d_in, int iq_dtype, long long n_samples, float* d_out_c64,
                        void* stream);

/* Acquisition -> tracking hand-off on the device (K-HAND), channel_l1ca_borre.py:301-311:
 * the peaks whose ratio exceeds `threshold` (best first, at most max_channels) become tracking
 * states ordered by PRN: carrierFrequency = IF - (-range + step * freq_idx), currentSample =
 * current_sample + required_samples - track_required + code_idx + 1; every other member is
 * copied from *d_template (loop coefficients, spacings, code NCO at nominal, see
 * sydr_trk_state_init; iq_base too: a template per recording places its channels at the recording's
 * offset inside the buffer of a many-recording tracking launch).  Slots beyond the selected count are marked idle (status 1).
 * d_n_selected (may be NULL) receives the count.  No host synchronisation. */
int sydr_acq_handoff(const sydr_acq_peak* d_peaks, int n_prn, double inter_freq, double doppler_range,
                     double doppler_step, long long required_samples, long long track_required,
                     long long current_sample, double threshold, const sydr_trk_state* d_template,
                     long long iq_len, sydr_trk_state* d_states, int max_channels, int* d_n_selected,
                     void* stream);

/* ------------------------------------------------------------------ Kaplan loops ------ */
/* Loop closure of ChannelL1CA_Kaplan on the device (sydr/channel/channel_l1ca_kaplan.py:342-619):
 * FLL_ATAN + PLL_costa discriminators, FLLassistedPLL_2ndOrder (sydr/dsp/tracking.py:156-176,
 * 246-279), FLL_Lock_Borre / PLL_Lock_Borre / CN0_Beaulieu (sydr/dsp/lockindicator.py:6-99), code
 * lock, bit synchronisation and the PULL_IN / WIDE_TRACK / NARROW_TRACK machine with its
 * bandwidth switching.  The code loop (DLL_NNEML + BorreLoopFilter) and the NCO members are those
 * of sydr_trk_state.  One record per channel next to its sydr_trk_state. */
typedef struct {
    /* [TRACKING] of channel_GPS_L1CA_kaplan.ini */
    double  fll_bw_pullin, fll_bw_wide, fll_bw_narrow;
    double  pll_bw_wide, pll_bw_narrow;
    double  fll_thr_wide, fll_thr_narrow, pll_thr_narrow, dll_threshold;
    /* channel members */
    double  ip_prev, qp_prev;           /* iPromptPrev, qPromptPrev                              */
    double  fll_lock, pll_lock;         /* fllLockIndicator, pllLockIndicator                    */
    double  cn0, pdpn;                  /* cn0 (= dllLockIndicator), cn0_PdPnRatio               */
    double  vel_memory;                 /* fll_vel_memory                                        */
    double  fll_bw, pll_bw;             /* fllBandwidth, pllBandwidth (current)                  */
    int32_t accum_counter;              /* correlatorsAccumCounter                               */
    int32_t lock_state;                 /* LoopLockState: 1 PULL_IN, 2 WIDE_TRACK, 3 NARROW_TRACK */
    int32_t flags;                      /* TrackingFlags bits: 1 CODE_LOCK, 2 BIT_SYNC           */
    int32_t reserved;
    int64_t code_counter;               /* codeCounter                                           */
} sydr_kaplan_state;        /* 168 bytes */

/* What the Kaplan TRACKING_UPDATE packet carries beyond sydr_trk_epoch (there: dll = code filter
 * output = code_frequency_error, code_err = DLL discriminator = "dll", pll = carrier filter output =
 * carrier_frequency_error, carrier_err = PLL discriminator = "pll"). */
typedef struct {
    double  fll;                        /* FLL discriminator                                     */
    double  cn0, fll_lock, pll_lock;
    int32_t lock_state, flags;
} sydr_kaplan_epoch;        /* 40 bytes */

/* sydr_trk_run with the Kaplan loop closure: d_kstates[n_channels] in/out,
 * d_kout[n_channels][max_epochs] indexed like d_out.  Both correlator spacings of the channel
 * configuration (wide / narrow) must be equal, as in the reference's ini. */
int sydr_trk_run_kaplan(const void* d_iq, int iq_dtype, long long iq_alloc_samples, double fs,
                        sydr_trk_state* d_states, sydr_kaplan_state* d_kstates, int n_channels,
                        sydr_trk_epoch* d_out, sydr_kaplan_epoch* d_kout, int max_epochs, int* d_nepochs,
                        const sydr_trk_config* cfg, void* stream);

/* ------------------------------------------------------------------ nav bits --------- */
/* Bit synchronisation + navigation-bit accumulation on the device (K-NAV): the scalar state
 * machine ChannelL1CA runs on every tracking result (channel_l1ca_borre.py:398-413 bit-sync
 * test, L367-373 / L577-591 / L626-627 prompt history, L455-491 20 prompts -> Prompt2Bit,
 * sydr/dsp/decoding.py:16-27).  The members are the channel attributes of the same meaning. */
typedef struct {
    int64_t code_counter;   /* codeCounter: tracking epochs ingested so far                  */
    int64_t sync_epoch;     /* epoch index at which BIT_SYNC was raised; -1 = not yet         */
    double  prev_iprompt;   /* iPrompt of the previous epoch                                  */
    double  row19;          /* correlatorsBuffer[19, IDX_I_PROMPT] (read once, at sync)       */
    double  nav_sum;        /* navPromptSum                                                   */
    int32_t nav_count;      /* navPromptSumCounter                                            */
    int32_t n_bits;         /* bits emitted so far (all calls)                                */
} sydr_nav_state;           /* 48 bytes */

int sydr_nav_state_init(sydr_nav_state* h_state);

/* Consume the tracking records d_epochs[ch][first_epoch .. d_nepochs[ch]) of n_channels
 * channels (the arrays sydr_trk_run fills) and emit the navigation bits they complete:
 * d_bits[ch][0 .. d_nbits[ch]) (0/1, Prompt2Bit with bit0 = 0) and, if d_bit_sums is not
 * NULL, the 20-epoch prompt sums they are the signs of (bit-identical to navPromptSum).  At
 * most max_bits bits per channel are stored per call.  d_nav carries the state across calls,
 * so a recording can be consumed in pieces of any length. */
int sydr_nav_bits(const sydr_trk_epoch* d_epochs, int max_epochs, const int* d_nepochs, int first_epoch,
                  sydr_nav_state* d_nav, int n_channels, signed char* d_bits, double* d_bit_sums,
                  int max_bits, int* d_nbits, void* stream);

/* The same for records of sydr_trk_run_kaplan: the Kaplan channel's rule (channel_l1ca_kaplan.py:555-566,
 * 725-758) - BIT_SYNC as raised by its trackingStateUpdate (d_kepochs[..].flags & 2), sums starting
 * with the synchronisation epoch's own prompt. */
int sydr_nav_bits_kaplan(const sydr_trk_epoch* d_epochs, const sydr_kaplan_epoch* d_kepochs, int max_epochs,
                         const int* d_nepochs, int first_epoch, sydr_nav_state* d_nav, int n_channels,
                         signed char* d_bits, double* d_bit_sums, int max_bits, int* d_nbits, void* stream);

/* ------------------------------------------------------------------ legacy C ABI ----- */
/* The per-call entry points the reference's ctypes callers bind
 * (sydr/old/tracking/tracking_epl_c.py:31-96, sydr/old/acquisition/acquisition_pcps_c.py:32-66),
 * same names, argument order and meaning as sydr/c_functions/tracking.c and acquisition.c,
 * host pointers in and out, void return.  Each one stages its arguments to the device and
 * runs the CUDA path above; errors are reported through sydr_last_error(). */
void generateReplica(double* time, size_t size, double carrierFrequency, double remCarrierPhase,
                     double* r_remCarrierPhase, double* r_replica_c128);          /* tracking.c:31  */
void getCorrelator(double* iSignal, double* qSignal, int* code, size_t size, double codeStep,
                   double remCodePhase, double correlatorSpacing,
                   double* r_iCorr, double* r_qCorr);                             /* tracking.c:69  */
void generateCarrier(double* rfData_c128, double* replica_c128, size_t size,
                     double* r_iSignal, double* r_qSignal);                       /* tracking.c:105 */
void delayLockLoop(double iEarly, double qEarly, double iLate, double qLate, double dllTau1,
                   double dllTau2, double pdiCode, double codeNCO, double codeError,
                   double codeFrequency, double* r_codeNCO, double* r_codeError,
                   double* r_codeFrequency);                                      /* tracking.c:131 */
void phaseLockLoop(double iPrompt, double qPrompt, double pllTau1, double pllTau2,
                   double pdiCarrier, double carrierNCO, double carrierError,
                   double initialFrequency, double* r_carrierNCO, double* r_carrierError,
                   double* r_carrierFrequency);                                   /* tracking.c:168 */
void getLoopCoefficients(double loopNoiseBandwidth, double dumpingRatio, double loopGain,
                         double* r_tau1, double* r_tau2);                         /* tracking.c:200 */
void setSatellite(const double* code, size_t size, double* codeFFT_c128);         /* acquisition.c:82  */
void PCPS(const double* rfData_c128, const double* codeFFT_c128, long long cohIntegration,
          long long nonCohIntegration, long long samplesPerCode, double samplingPeriod,
          double interFrequency, const double* frequencyBins, size_t s_frequencyBins,
          double* r_correlationMap);                                              /* acquisition.c:109 */
void twoCorrelationPeakComparison(const double* correlationMap, size_t s_correlationMap,
                                  const double* frequencyBins, size_t s_frequencyBins,
                                  long long samplesPerCode, long long samplesPerCodeChip,
                                  double interFrequency, double* r_acquisitionMetric,
                                  double* r_estimatedDoppler, double* r_estimatedFrequency,
                                  long long* r_estimatedCode, long long* r_idxEstimatedFrequency,
                                  long long* r_idxEstimatedCode);                 /* acquisition.c:181 */

#ifdef __cplusplus
}
#endif
#endif /* SYDR_B200_H */
