// In-register forward DFT butterflies (radix 2, 3, 4, 5, prime-factor 10 and Cooley-Tukey composites 8, 16,
// 20, 25) for the shared-memory mixed-radix FFT of the acquisition kernels.
//
// Only *forward* butterflies exist: the inverse transform is run as a forward transform on
// re/im-swapped data (ifft(z) = swap(fft(swap(z)))/N); the magnitude taken afterwards is
// invariant under the swap, so the swap back is never materialised.
#pragma once
#include <cuda_runtime.h>

namespace sydr {

// Complex arithmetic on the packed FP32x2 pipe of sm_100 (FADD2 / FMUL2 / FFMA2: both components of a
// complex number in one instruction, half the issue slots of the scalar forms).
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    // (a.x b.x - a.y b.y, a.x b.y + a.y b.x) = a.x * (b.x, b.y) + a.y * (-b.y, b.x)
    const float2 t = __fmul2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x));
    return __ffma2_rn(make_float2(a.x, a.x), b, t);
}
// s * a and a + s * b for a real scalar s
__device__ __forceinline__ float2 cscale(float s, float2 a) { return __fmul2_rn(make_float2(s, s), a); }
__device__ __forceinline__ float2 caxpy(float s, float2 b, float2 a) { return __ffma2_rn(make_float2(s, s), b, a); }
// multiply by -j
__device__ __forceinline__ float2 mul_mj(float2 a) { return make_float2(a.y, -a.x); }

// exp(-2*pi*i*m/R) for compile-time m (switch tables fold to immediates when unrolled).
template <int R>
__device__ __forceinline__ float2 root(int m);
#include "fft_roots.inc"

template <int R> struct Dft;

template <> struct Dft<1> {
    __device__ static __forceinline__ void run(float2*) {}
};
template <> struct Dft<2> {
    __device__ static __forceinline__ void run(float2* u) {
        const float2 a = u[0], b = u[1];
        u[0] = cadd(a, b);
        u[1] = csub(a, b);
    }
};
template <> struct Dft<3> {
    __device__ static __forceinline__ void run(float2* u) {
        const float s = 0.86602540378443864676f;
        const float2 t = cadd(u[1], u[2]);
        const float2 d = csub(u[1], u[2]);
        const float2 m = caxpy(-0.5f, t, u[0]);
        const float2 n = mul_mj(cscale(s, d));
        u[0] = cadd(u[0], t);
        u[1] = cadd(m, n);
        u[2] = csub(m, n);
    }
};
template <> struct Dft<4> {
    __device__ static __forceinline__ void run(float2* u) {
        const float2 a = cadd(u[0], u[2]), b = csub(u[0], u[2]);
        const float2 c = cadd(u[1], u[3]), d = mul_mj(csub(u[1], u[3]));
        u[0] = cadd(a, c);
        u[1] = cadd(b, d);
        u[2] = csub(a, c);
        u[3] = csub(b, d);
    }
};
template <> struct Dft<5> {
    __device__ static __forceinline__ void run(float2* u) {
        const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
        const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
        const float2 t1 = cadd(u[1], u[4]), t2 = cadd(u[2], u[3]);
        const float2 t3 = csub(u[1], u[4]), t4 = csub(u[2], u[3]);
        const float2 m1 = caxpy(c2, t2, caxpy(c1, t1, u[0]));
        const float2 m2 = caxpy(c1, t2, caxpy(c2, t1, u[0]));
        const float2 n1 = mul_mj(caxpy(s2, t4, cscale(s1, t3)));
        const float2 n2 = mul_mj(caxpy(-s1, t4, cscale(s2, t3)));
        u[0] = cadd(cadd(u[0], t1), t2);
        u[1] = cadd(m1, n1);
        u[4] = csub(m1, n1);
        u[2] = cadd(m2, n2);
        u[3] = csub(m2, n2);
    }
};

// Cooley-Tukey composite R = R1*R2:  n = n1*R2 + n2,  k = k1 + R1*k2.
template <int R1, int R2>
struct DftCT {
    static constexpr int R = R1 * R2;
    __device__ static __forceinline__ void run(float2* u) {
        float2 t[R];
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) {
            float2 a[R1];
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1) a[n1] = u[n1 * R2 + n2];
            Dft<R1>::run(a);
#pragma unroll
            for (int k1 = 0; k1 < R1; ++k1)
                t[k1 * R2 + n2] = (k1 * n2 == 0) ? a[k1] : cmul(a[k1], root<R>(k1 * n2));
        }
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) {
            float2 b[R2];
#pragma unroll
            for (int n2 = 0; n2 < R2; ++n2) b[n2] = t[k1 * R2 + n2];
            Dft<R2>::run(b);
#pragma unroll
            for (int k2 = 0; k2 < R2; ++k2) u[k1 + R1 * k2] = b[k2];
        }
    }
};
template <> struct Dft<8>  { __device__ static __forceinline__ void run(float2* u) { DftCT<2, 4>::run(u); } };
// Radix 10 by the prime-factor (Good-Thomas) map, 2 and 5 being coprime: n = (5 n1 + 2 n2) mod 10,
// k = (5 k1 + 6 k2) mod 10 turns the length-10 DFT into a 2 x 5 two-dimensional DFT with no twiddle
// factors in between (the Cooley-Tukey split spends four complex multiplications on them).
template <> struct Dft<10> {
    __device__ static __forceinline__ void run(float2* u) {
        float2 a[5], b[5];
#pragma unroll
        for (int n2 = 0; n2 < 5; ++n2) {
            const float2 x0 = u[(2 * n2) % 10], x1 = u[(5 + 2 * n2) % 10];
            a[n2] = cadd(x0, x1);                      // k1 = 0
            b[n2] = csub(x0, x1);                      // k1 = 1
        }
        Dft<5>::run(a);
        Dft<5>::run(b);
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2) {
            u[(6 * k2) % 10] = a[k2];
            u[(5 + 6 * k2) % 10] = b[k2];
        }
    }
};
template <> struct Dft<16> { __device__ static __forceinline__ void run(float2* u) { DftCT<4, 4>::run(u); } };
template <> struct Dft<20> { __device__ static __forceinline__ void run(float2* u) { DftCT<4, 5>::run(u); } };
template <> struct Dft<25> { __device__ static __forceinline__ void run(float2* u) { DftCT<5, 5>::run(u); } };

// u[r] *= w^r for r = 1..R-1 with log-depth power products (w = base twiddle of this butterfly).
template <int R>
__device__ __forceinline__ void twiddle_powers(float2* u, float2 w) {
    float2 pw[R];
    pw[0] = make_float2(1.f, 0.f);
    if (R > 1) pw[1] = w;
#pragma unroll
    for (int r = 2; r < R; ++r) pw[r] = cmul(pw[r / 2], pw[r - r / 2]);
#pragma unroll
    for (int r = 1; r < R; ++r) u[r] = cmul(u[r], pw[r]);
}

}  // namespace sydr
