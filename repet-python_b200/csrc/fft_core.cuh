// Complex FFT of N = 512 / 1024 / 2048 points for one group of N/16 threads, three radix stages, fp32.
//
// REPET's window length follows the sampling rate (N = 2^ceil(log2(0.04 fs)), repet.py:130): 512 up to
// 12.8 kHz, 1024 up to 25.6 kHz, 2048 up to 51.2 kHz.  The STFT and ISTFT frames (two real channels
// packed into one complex transform) use Fft<N>; the time-axis transforms of the beat spectrum (two
// frequency rows packed into one complex transform, zero-padded) always use Fft<2048>.
//
// Decomposition (Cooley-Tukey) with T = N/16 threads, each holding 16 points, R1 = 16, T = R2 * R3:
//   n = n1*T + n2*R3 + m2,   k = k1 + 16*k2 + 16*R2*k3
//   stage 1  thread m = t : DFT-16 over n1 of x[n1*T + m], times W_N^(m*k1)
//   stage 2  thread does 16/R2 butterflies (k1 = t%16, m2 = t/16 + (T/16) q): DFT-R2 over n2 of
//            y1[k1][n2*R3 + m2], times W_T^(m2*k2)
//   stage 3  thread owns the columns pi in {t, 2T - t} (k1 = pi%16, k2 = pi/16): DFT-8 over m2 of
//            y2[k2][m2][k1]  ->  Z[pi + 2T*k3]
//   plans: 512 = 16 x 4 x 8, 1024 = 16 x 8 x 8, 2048 = 16 x 16 x 8  (R3 = 8 always)
// The two exchanges go through shared memory:
//   y1 at [k1*(T+1) + m]          (row pad 1: stage-2 reads have an odd stride)
//   y2 at [(k2*8 + m2)*16 + k1]   (k1 fastest: stage-2 writes and stage-3 reads are contiguous)
// A thread owns the same residues mod T on input (n1*T + t) and, mod 2T, on output (columns t and
// 2T - t, see out_column), which lets the STFT keep the overlapping half frame, split the two packed
// channels, and the ISTFT do its overlap-add, all in registers.
#pragma once
#include <cuda_runtime.h>

namespace repet {

// Complex add/sub are one packed-fp32 instruction each on sm_100 (FADD2 / FFMA2): the two halves
// of a float2 live in an aligned register pair.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// multiply by -i
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }
// float64 flavour of the same plan (k_frames64: the similarity operand of REPET-SIM)
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cmul_mi(double2 a) { return make_double2(a.y, -a.x); }

// 4-point DFT, natural order in and out.  V = float2 or double2.
template <typename V>
__device__ __forceinline__ void dft4(V& a0, V& a1, V& a2, V& a3) {
    const V t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), d = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    // t1 +- (-i d): the cross terms stay scalar (no half swap needed)
    a1 = V{t1.x + d.y, t1.y - d.x};
    a3 = V{t1.x - d.y, t1.y + d.x};
}

// constants in the component type of V (the float values are the correctly rounded doubles)
template <typename V>
struct Real;
template <>
struct Real<float2> {
    using type = float;
};
template <>
struct Real<double2> {
    using type = double;
};
#define REPET_SQRT1_2 ((typename Real<V>::type)0.70710678118654752440)
#define REPET_COS_PI_8 ((typename Real<V>::type)0.92387953251128675613)
#define REPET_SIN_PI_8 ((typename Real<V>::type)0.38268343236508977173)

// multiply by W_16^j for the j that occur in a 4x4 split (j = m*k1, m,k1 in 0..3)
template <int J, typename V>
__device__ __forceinline__ V mul_w16(V a) {
    if (J == 0) return a;
    if (J == 1) return cmul(a, V{REPET_COS_PI_8, -REPET_SIN_PI_8});
    if (J == 2) return V{(a.x + a.y) * REPET_SQRT1_2, (a.y - a.x) * REPET_SQRT1_2};
    if (J == 3) return cmul(a, V{REPET_SIN_PI_8, -REPET_COS_PI_8});
    if (J == 4) return cmul_mi(a);
    if (J == 6) return V{(a.y - a.x) * REPET_SQRT1_2, -(a.x + a.y) * REPET_SQRT1_2};
    if (J == 9) return cmul(a, V{-REPET_COS_PI_8, REPET_SIN_PI_8});
    return a;
}

// 16-point DFT in registers, natural order in and out (n = n1*4 + m, k = k1 + 4*k').
template <typename V>
__device__ __forceinline__ void dft16(V (&a)[16]) {
#pragma unroll
    for (int m = 0; m < 4; ++m) dft4(a[m], a[4 + m], a[8 + m], a[12 + m]);
    // a[k1*4 + m] now holds u[m][k1]; twiddle by W_16^(m*k1)
    a[5] = mul_w16<1>(a[5]);
    a[6] = mul_w16<2>(a[6]);
    a[7] = mul_w16<3>(a[7]);
    a[9] = mul_w16<2>(a[9]);
    a[10] = mul_w16<4>(a[10]);
    a[11] = mul_w16<6>(a[11]);
    a[13] = mul_w16<3>(a[13]);
    a[14] = mul_w16<6>(a[14]);
    a[15] = mul_w16<9>(a[15]);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);
    // a[k1*4 + k'] holds X[k1 + 4*k']: transpose the 4x4 index to natural order
    V t;
#define REPET_SWAP(i, j) t = a[i]; a[i] = a[j]; a[j] = t;
    REPET_SWAP(1, 4) REPET_SWAP(2, 8) REPET_SWAP(3, 12) REPET_SWAP(6, 9) REPET_SWAP(7, 13) REPET_SWAP(11, 14)
#undef REPET_SWAP
}

// 8-point DFT in registers, natural order in and out (n = n1*2 + m, k = k1 + 4*k').
template <typename V>
__device__ __forceinline__ void dft8(V (&a)[8]) {
    dft4(a[0], a[2], a[4], a[6]);  // m = 0: u0[k1] in a[2*k1]
    dft4(a[1], a[3], a[5], a[7]);  // m = 1: u1[k1] in a[2*k1 + 1]
    V u1_1 = V{(a[3].x + a[3].y) * REPET_SQRT1_2, (a[3].y - a[3].x) * REPET_SQRT1_2};   // W_8^1
    V u1_2 = cmul_mi(a[5]);                                                             // W_8^2
    V u1_3 = V{(a[7].y - a[7].x) * REPET_SQRT1_2, -(a[7].x + a[7].y) * REPET_SQRT1_2};  // W_8^3
    V u0_0 = a[0], u0_1 = a[2], u0_2 = a[4], u0_3 = a[6], u1_0 = a[1];
    a[0] = cadd(u0_0, u1_0);
    a[4] = csub(u0_0, u1_0);
    a[1] = cadd(u0_1, u1_1);
    a[5] = csub(u0_1, u1_1);
    a[2] = cadd(u0_2, u1_2);
    a[6] = csub(u0_2, u1_2);
    a[3] = cadd(u0_3, u1_3);
    a[7] = csub(u0_3, u1_3);
}

template <int N>
struct FftPlan;
template <>
struct FftPlan<512> {
    static constexpr int R2 = 4;
};
template <>
struct FftPlan<1024> {
    static constexpr int R2 = 8;
};
template <>
struct FftPlan<2048> {
    static constexpr int R2 = 16;
};

// DFT of R consecutive registers r[OFF .. OFF+R) of the 16-element array (natural order in and out)
template <int R, int OFF, typename V>
__device__ __forceinline__ void dft_slice(V (&r)[16]) {
    if constexpr (R == 16) {
        dft16(r);
    } else if constexpr (R == 8) {
        V c[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = r[OFF + i];
        dft8(c);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[OFF + i] = c[i];
    } else {
        dft4(r[OFF], r[OFF + 1], r[OFF + 2], r[OFF + 3]);
    }
}

template <int N, typename V = float2>
struct Fft {
    static constexpr int SIZE = N;
    static constexpr int THREADS = N / 16;
    static constexpr int R2 = FftPlan<N>::R2;
    static constexpr int R3 = 8;
    static constexpr int NB2 = 16 / R2;           // stage-2 butterflies per thread
    static constexpr int PITCH1 = THREADS + 1;    // y1 row pitch
    static constexpr int BUF = 16 * PITCH1;       // V elements per exchange buffer (>= N)
    static constexpr int CCOLS = 2 * THREADS;     // output columns (residues mod 2T)
    static constexpr int TW1 = 15 * THREADS;      // entries of the stage-1 twiddle table
    static constexpr int TW2 = THREADS;           // entries of the stage-2 twiddle table [k2][m2]
    static_assert(R2 * R3 == THREADS, "plan must factor N/16");

    // Per-thread constant twiddles of stage 1: W_N^(m*k1), k1 = 1..15, from the table
    // tw1g[(k1-1)*T + m] (built in double precision on the host, repet_abi.cu).
    struct Twiddle1 {
        V w[15];
        __device__ __forceinline__ void load(const V* __restrict__ tw1g, int t) {
#pragma unroll
            for (int k1 = 1; k1 < 16; ++k1) w[k1 - 1] = __ldg(&tw1g[(k1 - 1) * THREADS + t]);
        }
    };

    // stage 1: r[n1] = x[n1*T + t] on entry; writes y1 to dst.
    static __device__ __forceinline__ void stage1(V (&r)[16], const Twiddle1& tw, V* __restrict__ dst, int t) {
        dft16(r);
        dst[t] = r[0];
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) dst[k1 * PITCH1 + t] = cmul(r[k1], tw.w[k1 - 1]);
    }

    template <int Q>
    static __device__ __forceinline__ void stage2_one(V (&r)[16], const V* __restrict__ src,
                                                      V* __restrict__ dst, const V* __restrict__ s_tw2,
                                                      int k1, int mb) {
        const int m2 = mb + (THREADS / 16) * Q;
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) r[Q * R2 + n2] = src[k1 * PITCH1 + n2 * R3 + m2];
        dft_slice<R2, Q * R2>(r);
        dst[m2 * 16 + k1] = r[Q * R2];
#pragma unroll
        for (int k2 = 1; k2 < R2; ++k2) dst[(k2 * R3 + m2) * 16 + k1] = cmul(r[Q * R2 + k2], s_tw2[k2 * R3 + m2]);
    }

    // stage 2: reads y1 from src, writes y2 to dst.  s_tw2[k2*8 + m2] = W_T^(m2*k2).
    static __device__ __forceinline__ void stage2(V (&r)[16], const V* __restrict__ src,
                                                  V* __restrict__ dst, const V* __restrict__ s_tw2, int t) {
        const int k1 = t & 15, mb = t >> 4;
        stage2_one<0>(r, src, dst, s_tw2, k1, mb);
        if constexpr (NB2 > 1) stage2_one<(NB2 > 1 ? 1 : 0)>(r, src, dst, s_tw2, k1, mb);
        if constexpr (NB2 > 2) {
            stage2_one<(NB2 > 2 ? 2 : 0)>(r, src, dst, s_tw2, k1, mb);
            stage2_one<(NB2 > 2 ? 3 : 0)>(r, src, dst, s_tw2, k1, mb);
        }
    }

    // The two output columns (residues mod 2T) a thread owns after stage 3: t and 2T - t, except
    // thread 0 which owns the two self-mirrored columns 0 and T.  Column c and column 2T - c hold each
    // other's mirror bins (N - (c + 2T k3) = (2T - c) + 2T (7 - k3)), so a thread has Z[k] AND
    // Z[N - k] in registers: the Hermitian split of two packed real channels needs no further
    // exchange.  Time-domain use: samples n and n + N/2 sit in the same column (k3, k3 + 4).
    static __device__ __forceinline__ int out_column(int t, int h) {
        return t == 0 ? THREADS * h : (h == 0 ? t : CCOLS - t);
    }

    // stage 3: reads y2 from src; on exit r[h*8 + k3] = Z[out_column(t, h) + 2T*k3].
    static __device__ __forceinline__ void stage3(V (&r)[16], const V* __restrict__ src, int t) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int pi = out_column(t, h);
            const V* col = src + (pi >> 4) * (R3 * 16) + (pi & 15);
            V c[8];
#pragma unroll
            for (int m2 = 0; m2 < 8; ++m2) c[m2] = col[m2 * 16];
            dft8(c);
#pragma unroll
            for (int k3 = 0; k3 < 8; ++k3) r[h * 8 + k3] = c[k3];
        }
    }

    // ---- transposed plan (decimation in frequency): the same three radix stages run backwards ----
    // Input  r[h*8 + k3] = Z[out_column(t, h) + 2T*k3]   (the forward plan's output ownership: a thread
    //                                                      holds Z[k] and Z[N-k], so Hermitian spectra
    //                                                      are assembled in registers)
    // Output r[n1] = x[n1*T + t]                          (samples n and n + N/2 in one thread: the
    //                                                      overlap-add needs no exchange either)
    //   tstage3  DFT-8 over k3 of the thread's two columns, times W_T^(m2*k2)        -> y2 layout
    //   tstage2  DFT-R2 over k2 for (k1 = t%16, m2), times W_N^((n2*8+m2)*k1)        -> y1 layout
    //   tstage1  DFT-16 over k1 for m = t
    // Per-thread constant twiddles of tstage2, from the stage-1 table tw1g[(k1-1)*T + m].
    struct TwiddleT {
        V w[16];
        __device__ __forceinline__ void load(const V* __restrict__ tw1g, int t) {
            const int k1 = t & 15, mb = t >> 4;
#pragma unroll
            for (int q = 0; q < NB2; ++q)
#pragma unroll
                for (int n2 = 0; n2 < R2; ++n2) {
                    const int m = n2 * R3 + mb + (THREADS / 16) * q;
                    w[q * R2 + n2] = k1 ? __ldg(&tw1g[(k1 - 1) * THREADS + m]) : V{1, 0};
                }
        }
    };

    static __device__ __forceinline__ void tstage3(V (&r)[16], V* __restrict__ dst,
                                                   const V* __restrict__ s_tw2, int t) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int pi = out_column(t, h);
            V c[8];
#pragma unroll
            for (int k3 = 0; k3 < 8; ++k3) c[k3] = r[h * 8 + k3];
            dft8(c);
            V* col = dst + (pi >> 4) * (R3 * 16) + (pi & 15);
            const V* twc = s_tw2 + (pi >> 4) * R3;
            col[0] = c[0];
#pragma unroll
            for (int m2 = 1; m2 < 8; ++m2) col[m2 * 16] = cmul(c[m2], twc[m2]);
        }
    }

    template <int Q>
    static __device__ __forceinline__ void tstage2_one(V (&r)[16], const V* __restrict__ src,
                                                       V* __restrict__ dst, const TwiddleT& tw, int k1, int mb) {
        const int m2 = mb + (THREADS / 16) * Q;
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) r[Q * R2 + k2] = src[(k2 * R3 + m2) * 16 + k1];
        dft_slice<R2, Q * R2>(r);
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) dst[k1 * PITCH1 + n2 * R3 + m2] = cmul(r[Q * R2 + n2], tw.w[Q * R2 + n2]);
    }

    static __device__ __forceinline__ void tstage2(V (&r)[16], const V* __restrict__ src,
                                                   V* __restrict__ dst, const TwiddleT& tw, int t) {
        const int k1 = t & 15, mb = t >> 4;
        tstage2_one<0>(r, src, dst, tw, k1, mb);
        if constexpr (NB2 > 1) tstage2_one<(NB2 > 1 ? 1 : 0)>(r, src, dst, tw, k1, mb);
        if constexpr (NB2 > 2) {
            tstage2_one<(NB2 > 2 ? 2 : 0)>(r, src, dst, tw, k1, mb);
            tstage2_one<(NB2 > 2 ? 3 : 0)>(r, src, dst, tw, k1, mb);
        }
    }

    static __device__ __forceinline__ void tstage1(V (&r)[16], const V* __restrict__ src, int t) {
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) r[k1] = src[k1 * PITCH1 + t];
        dft16(r);
    }
};

// One squared-magnitude / magnitude definition for every kernel, so that |X| is bit-identical
// wherever it is recomputed.  The square root is the hardware approximation (<= 2 ulp): the
// magnitudes only feed medians, soft masks and the beat spectrum, all far above that error.
__device__ __forceinline__ float cmag2(float2 v) { return __fmaf_rn(v.x, v.x, __fmul_rn(v.y, v.y)); }
__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_rsqrt(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float cmag(float2 v) { return fast_sqrt(cmag2(v)); }

// one 128-byte line towards L2, no register or scoreboard cost
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

}  // namespace repet
