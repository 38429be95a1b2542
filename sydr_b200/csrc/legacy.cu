// Legacy per-call C ABI: the entry points the reference's ctypes callers bind
// (sydr/old/tracking/tracking_epl_c.py:31-96, sydr/old/acquisition/acquisition_pcps_c.py:32-66),
// same names / argument order as sydr/c_functions/tracking.c and acquisition.c.  Every function
// stages its host arguments to the device, runs a CUDA kernel and copies the result back; there
// is no host arithmetic path.  Errors are reported through sydr_last_error().
//
// Deliberate deviation: acquisition.c's PCPS/setSatellite use a real-input FFT and return a
// half-width, numerically wrong map.  The shims here follow the live Python
// semantics instead (full complex spectra, full-width map), which is what the callers'
// Python twins (acquisition_pcps.py) compute.
#include <vector>

#include "common.cuh"

namespace sydr {

__global__ void legacy_replica_kernel(const double* __restrict__ time, size_t size, double fc, double rem,
                                      double* __restrict__ r_rem, double2* __restrict__ replica) {
    // tracking.c:40-49 (GPS pi)
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i <= size; i += stride) {
        const double temp = __dadd_rn(-__dmul_rn(__dmul_rn(__dmul_rn(fc, 2.0), kGpsPi), time[i]), rem);
        if (i < size) {
            double s, c;
            sincos(temp, &s, &c);
            replica[i] = make_double2(c, s);
        } else {
            *r_rem = fmod(temp, 2 * kGpsPi);
        }
    }
}

__global__ void legacy_carrier_kernel(const double2* __restrict__ rf, const double2* __restrict__ rep, size_t size,
                                      double* __restrict__ ri, double* __restrict__ rq) {
    // tracking.c:113-119
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += stride) {
        const double2 a = rf[i], b = rep[i];
        ri[i] = __dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y));
        rq[i] = __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x));
    }
}

__global__ void legacy_correlator_kernel(const double* __restrict__ is, const double* __restrict__ qs,
                                         const int* __restrict__ code, int code_len, size_t size, double codeStep,
                                         double rem, double spacing, double* __restrict__ out2) {
    // tracking.c:79-93: start/stop/step then ceil(start + step*idx)
    __shared__ double si[256], sq[256];
    const double start = __dadd_rn(rem, spacing);
    const double stop = __dadd_rn(__dadd_rn(__dmul_rn((double)size, codeStep), rem), spacing);
    const double step = __ddiv_rn(__dsub_rn(stop, start), (double)size);
    double ai = 0.0, aq = 0.0;
    for (size_t i = threadIdx.x; i < size; i += blockDim.x) {
        long long k = (long long)ceil(__dadd_rn(start, __dmul_rn(step, (double)i)));
        if (k < 0) k = 0;
        if (k >= code_len) k = code_len - 1;
        const double c = (double)code[k];
        ai += c * is[i];
        aq += c * qs[i];
    }
    si[threadIdx.x] = ai; sq[threadIdx.x] = aq;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) { si[threadIdx.x] += si[threadIdx.x + o]; sq[threadIdx.x] += sq[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out2[0] = si[0]; out2[1] = sq[0]; }
}

// op 0: delayLockLoop (tracking.c:144-156); 1: phaseLockLoop (181-188); 2: getLoopCoefficients (206-209)
__global__ void legacy_scalar_kernel(int op, const double* __restrict__ a, double* __restrict__ r) {
    if (op == 0) {
        const double me = sqrt(a[0] * a[0] + a[1] * a[1]), ml = sqrt(a[2] * a[2] + a[3] * a[3]);
        const double err = __ddiv_rn(__dsub_rn(me, ml), __dadd_rn(me, ml));
        double nco = a[7];
        nco = __dadd_rn(nco, __dmul_rn(__ddiv_rn(a[5], a[4]), __dsub_rn(err, a[8])));
        nco = __dadd_rn(nco, __dmul_rn(__ddiv_rn(a[6], a[4]), err));
        r[0] = nco; r[1] = err; r[2] = __dsub_rn(a[9], nco);
    } else if (op == 1) {
        const double err = __ddiv_rn(__ddiv_rn(atan(__ddiv_rn(a[1], a[0])), 2.0), kGpsPi);
        double nco = a[5];
        nco = __dadd_rn(nco, __dmul_rn(__ddiv_rn(a[3], a[2]), __dsub_rn(err, a[6])));
        nco = __dadd_rn(nco, __dmul_rn(__ddiv_rn(a[4], a[2]), err));
        r[0] = nco; r[1] = err; r[2] = __dadd_rn(a[7], nco);
    } else {
        const double wn = __ddiv_rn(__dmul_rn(__dmul_rn(a[0], 8.0), a[1]), __dadd_rn(__dmul_rn(4.0, __dmul_rn(a[1], a[1])), 1.0));
        r[0] = __ddiv_rn(a[2], __dmul_rn(wn, wn));
        r[1] = __ddiv_rn(__dmul_rn(2.0, a[1]), wn);
    }
}

__global__ void widen_map_kernel(const float* __restrict__ in, size_t n, double* __restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (double)in[i];
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    bool alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1) == cudaSuccess; }
    template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

static bool h2d(DevBuf& b, const void* h, size_t bytes) {
    return b.alloc(bytes) && cudaMemcpy(b.p, h, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
}
static bool d2h(void* h, const void* d, size_t bytes) { return cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost) == cudaSuccess; }
static void fail(const char* who) { set_error("%s: CUDA staging failed: %s", who, cudaGetErrorString(cudaGetLastError())); }

static int scalar_call(int op, const double* args, int nargs, double* res, int nres) {
    DevBuf a, r;
    if (!h2d(a, args, sizeof(double) * nargs) || !r.alloc(sizeof(double) * 3)) { fail("scalar"); return SYDR_ERR_CUDA; }
    legacy_scalar_kernel<<<1, 1>>>(op, a.as<double>(), r.as<double>());
    count_launch();
    if (!d2h(res, r.p, sizeof(double) * nres)) { fail("scalar"); return SYDR_ERR_CUDA; }
    return SYDR_OK;
}

int fft64_host_rows(const double* h_in_c128, int n, int batch, double* h_out_c128);   // acq.cu

}  // namespace sydr

using namespace sydr;

extern "C" {

void generateReplica(double* time, size_t size, double carrierFrequency, double remCarrierPhase,
                     double* r_remCarrierPhase, double* r_replica) {
    DevBuf t, rem, rep;
    if (!h2d(t, time, sizeof(double) * (size + 1)) || !rem.alloc(8) || !rep.alloc(sizeof(double2) * size)) return fail("generateReplica");
    legacy_replica_kernel<<<148, 256>>>(t.as<double>(), size, carrierFrequency, remCarrierPhase, rem.as<double>(), rep.as<double2>());
    count_launch();
    if (!d2h(r_remCarrierPhase, rem.p, 8) || !d2h(r_replica, rep.p, sizeof(double2) * size)) fail("generateReplica");
}

void generateCarrier(double* rfData, double* replica, size_t size, double* r_iSignal, double* r_qSignal) {
    DevBuf a, b, ri, rq;
    if (!h2d(a, rfData, sizeof(double2) * size) || !h2d(b, replica, sizeof(double2) * size) || !ri.alloc(8 * size) || !rq.alloc(8 * size))
        return fail("generateCarrier");
    legacy_carrier_kernel<<<148, 256>>>(a.as<double2>(), b.as<double2>(), size, ri.as<double>(), rq.as<double>());
    count_launch();
    if (!d2h(r_iSignal, ri.p, 8 * size) || !d2h(r_qSignal, rq.p, 8 * size)) fail("generateCarrier");
}

void getCorrelator(double* iSignal, double* qSignal, int* code, size_t size, double codeStep, double remCodePhase,
                   double correlatorSpacing, double* r_iCorr, double* r_qCorr) {
    // highest code index the C loop can touch (tracking.c:89)
    long long code_len = (long long)ceil(size * codeStep + remCodePhase + correlatorSpacing) + 1;
    if (code_len < 1) code_len = 1;
    DevBuf a, b, c, o;
    if (!h2d(a, iSignal, 8 * size) || !h2d(b, qSignal, 8 * size) || !h2d(c, code, sizeof(int) * code_len) || !o.alloc(16))
        return fail("getCorrelator");
    legacy_correlator_kernel<<<1, 256>>>(a.as<double>(), b.as<double>(), c.as<int>(), (int)code_len, size, codeStep,
                                         remCodePhase, correlatorSpacing, o.as<double>());
    count_launch();
    double r[2];
    if (!d2h(r, o.p, 16)) return fail("getCorrelator");
    *r_iCorr = r[0];
    *r_qCorr = r[1];
}

void delayLockLoop(double iEarly, double qEarly, double iLate, double qLate, double dllTau1, double dllTau2,
                   double pdiCode, double codeNCO, double codeError, double codeFrequency, double* r_codeNCO,
                   double* r_codeError, double* r_codeFrequency) {
    const double a[10] = {iEarly, qEarly, iLate, qLate, dllTau1, dllTau2, pdiCode, codeNCO, codeError, codeFrequency};
    double r[3];
    if (scalar_call(0, a, 10, r, 3) != SYDR_OK) return;
    *r_codeNCO = r[0]; *r_codeError = r[1]; *r_codeFrequency = r[2];
}

void phaseLockLoop(double iPrompt, double qPrompt, double pllTau1, double pllTau2, double pdiCarrier,
                   double carrierNCO, double carrierError, double initialFrequency, double* r_carrierNCO,
                   double* r_carrierError, double* r_carrierFrequency) {
    const double a[8] = {iPrompt, qPrompt, pllTau1, pllTau2, pdiCarrier, carrierNCO, carrierError, initialFrequency};
    double r[3];
    if (scalar_call(1, a, 8, r, 3) != SYDR_OK) return;
    *r_carrierNCO = r[0]; *r_carrierError = r[1]; *r_carrierFrequency = r[2];
}

void getLoopCoefficients(double loopNoiseBandwidth, double dumpingRatio, double loopGain, double* r_tau1, double* r_tau2) {
    const double a[3] = {loopNoiseBandwidth, dumpingRatio, loopGain};
    double r[2];
    if (scalar_call(2, a, 3, r, 2) != SYDR_OK) return;
    *r_tau1 = r[0]; *r_tau2 = r[1];
}

void setSatellite(const double* code, size_t size, double* codeFFT) {
    // conj(fft(code)), full complex spectrum of `size` points
    std::vector<double> in(2 * size), out(2 * size);
    for (size_t i = 0; i < size; ++i) { in[2 * i] = code[i]; in[2 * i + 1] = 0.0; }
    if (fft64_host_rows(in.data(), (int)size, 1, out.data()) != SYDR_OK) return;
    for (size_t i = 0; i < size; ++i) { codeFFT[2 * i] = out[2 * i]; codeFFT[2 * i + 1] = -out[2 * i + 1]; }
}

void PCPS(const double* rfData, const double* codeFFT, long long cohIntegration, long long nonCohIntegration,
          long long samplesPerCode, double samplingPeriod, double interFrequency, const double* frequencyBins,
          size_t s_frequencyBins, double* r_correlationMap) {
    if (s_frequencyBins < 1) { set_error("PCPS: no frequency bins"); return; }
    const double fs = 1.0 / samplingPeriod;
    const double range = -frequencyBins[0];
    const double step = s_frequencyBins > 1 ? frequencyBins[1] - frequencyBins[0] : 1.0;
    const int prn = 1;      // spectrum is replaced below
    sydr_acq_plan* pl = nullptr;
    // the plan's own bin count may exceed the caller's (arange with or without the +1): take a prefix
    int rc = sydr_acq_plan_create(fs, interFrequency, range, step, (int)cohIntegration, (int)nonCohIntegration, &prn, 1,
                                  0, -1, &pl);
    if (rc != SYDR_OK) return;
    int n_code = 0, n_bins = 0;
    sydr_acq_plan_info(pl, &n_code, &n_bins, nullptr, nullptr, nullptr);
    sydr_acq_plan_destroy(pl);
    if (n_code != samplesPerCode) { set_error("PCPS: samplesPerCode %lld != round(fs/1000) = %d", samplesPerCode, n_code); return; }
    rc = sydr_acq_plan_create(fs, interFrequency, range, step, (int)cohIntegration, (int)nonCohIntegration, &prn, 1, 0,
                              (int)s_frequencyBins <= n_bins ? (int)s_frequencyBins : -1, &pl);
    if (rc != SYDR_OK) return;
    if ((int)s_frequencyBins > n_bins) { set_error("PCPS: %zu bins exceed the plan's %d", s_frequencyBins, n_bins); sydr_acq_plan_destroy(pl); return; }
    const size_t n_iq = (size_t)samplesPerCode * cohIntegration * nonCohIntegration;
    const size_t n_map = s_frequencyBins * (size_t)samplesPerCode;
    DevBuf iq, m32, m64;
    if (sydr_acq_plan_set_spectrum(pl, 0, codeFFT) == SYDR_OK && h2d(iq, rfData, sizeof(double2) * n_iq) &&
        m32.alloc(sizeof(float) * n_map) && m64.alloc(sizeof(double) * n_map)) {
        if (sydr_acq_run(pl, iq.p, SYDR_IQ_F64, (long long)n_iq, nullptr, nullptr, m32.as<float>(), nullptr) == SYDR_OK) {
            widen_map_kernel<<<148 * 4, 256>>>(m32.as<float>(), n_map, m64.as<double>());
            count_launch();
            if (!d2h(r_correlationMap, m64.p, sizeof(double) * n_map)) fail("PCPS");
        }
    } else {
        fail("PCPS");
    }
    sydr_acq_plan_destroy(pl);
}

void twoCorrelationPeakComparison(const double* correlationMap, size_t s_correlationMap, const double* frequencyBins,
                                  size_t s_frequencyBins, long long samplesPerCode, long long samplesPerCodeChip,
                                  double interFrequency, double* r_acquisitionMetric, double* r_estimatedDoppler,
                                  double* r_estimatedFrequency, long long* r_estimatedCode,
                                  long long* r_idxEstimatedFrequency, long long* r_idxEstimatedCode) {
    int fi = 0, ci = 0;
    double ratio = 0.0;
    (void)samplesPerCode;
    if (sydr_peak_compare(correlationMap, (int)s_frequencyBins, (int)s_correlationMap, (int)samplesPerCodeChip, &fi, &ci,
                          &ratio) != SYDR_OK)
        return;
    *r_estimatedDoppler = -frequencyBins[fi];                 // acquisition.c:235
    *r_estimatedCode = ci;
    *r_acquisitionMetric = ratio;
    *r_idxEstimatedFrequency = fi;
    *r_estimatedFrequency = interFrequency + *r_estimatedDoppler;
    *r_idxEstimatedCode = ci;
}

}  // extern "C"
