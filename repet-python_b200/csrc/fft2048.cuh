// 2048-point complex FFT for one 128-thread group, radix 16 x 16 x 8, fp32.
//
// Every FFT on the REPET path at 32-48 kHz is a 2048-point transform (window length
// N = 2^ceil(log2(0.04 fs)), repet.py:130): the STFT and ISTFT frames (two real channels
// packed into one complex transform) and the time-axis transforms of the beat spectrum
// (two frequency rows packed into one complex transform, zero-padded to 2048).
//
// Decomposition (Cooley-Tukey, n = n1*128 + n2*8 + m2, k = k1 + 16*k2 + 256*k3):
//   stage 1  thread m = t            : DFT-16 over n1 of x[n1*128 + m], times W_2048^(m*k1)
//   stage 2  thread (k1 = t%16, m2 = t/16): DFT-16 over n2 of y1[k1][n2*8 + m2], times W_128^(m2*k2)
//   stage 3  thread owns columns pi in {t, 256-t} (k1 = pi%16, k2 = pi/16): DFT-8 over m2 of y2[k2][m2][k1]
//            -> Z[pi + 256*k3]
// The two exchanges go through shared memory with layouts chosen so that every access of
// a half-warp (8-byte elements) is bank-conflict free:
//   y1 at [k1*129 + m]           (row pad 1: stage-2 reads have stride 129 float2 = 258 words)
//   y2 at [(k2*8 + m2)*16 + k1]  (k1 fastest: stage-2 writes and stage-3 reads are contiguous)
// A thread owns the same residues mod 128 on input (n1*128 + t) and, mod 256, on output
// (columns t and 256 - t, see fft_out_column), which lets the STFT keep the overlapping half
// frame, split the two packed channels, and the ISTFT do its overlap-add, all in registers.
#pragma once
#include <cuda_runtime.h>

namespace repet {

constexpr int FFT_N = 2048;
constexpr int FFT_THREADS = 128;
constexpr int FFT_BUF = 16 * 129;  // float2 elements per exchange buffer (>= 2048)

// Complex add/sub are one packed-fp32 instruction each on sm_100 (FADD2 / FFMA2): the two halves
// of a float2 live in an aligned register pair.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// multiply by -i
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }

// 4-point DFT, natural order in and out.
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    const float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), d = csub(a1, a3);
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    // t1 +- (-i d): the cross terms stay scalar (no half swap needed)
    a1 = make_float2(t1.x + d.y, t1.y - d.x);
    a3 = make_float2(t1.x - d.y, t1.y + d.x);
}

#define REPET_SQRT1_2 0.70710678118654752440f
#define REPET_COS_PI_8 0.92387953251128675613f
#define REPET_SIN_PI_8 0.38268343236508977173f

// multiply by W_16^j for the j that occur in a 4x4 split (j = m*k1, m,k1 in 0..3)
template <int J>
__device__ __forceinline__ float2 mul_w16(float2 a) {
    if (J == 0) return a;
    if (J == 1) return cmul(a, make_float2(REPET_COS_PI_8, -REPET_SIN_PI_8));
    if (J == 2) return make_float2((a.x + a.y) * REPET_SQRT1_2, (a.y - a.x) * REPET_SQRT1_2);
    if (J == 3) return cmul(a, make_float2(REPET_SIN_PI_8, -REPET_COS_PI_8));
    if (J == 4) return cmul_mi(a);
    if (J == 6) return make_float2((a.y - a.x) * REPET_SQRT1_2, -(a.x + a.y) * REPET_SQRT1_2);
    if (J == 9) return cmul(a, make_float2(-REPET_COS_PI_8, REPET_SIN_PI_8));
    return a;
}

// 16-point DFT in registers, natural order in and out (n = n1*4 + m, k = k1 + 4*k').
__device__ __forceinline__ void dft16(float2 (&a)[16]) {
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
    float2 t;
#define REPET_SWAP(i, j) t = a[i]; a[i] = a[j]; a[j] = t;
    REPET_SWAP(1, 4) REPET_SWAP(2, 8) REPET_SWAP(3, 12) REPET_SWAP(6, 9) REPET_SWAP(7, 13) REPET_SWAP(11, 14)
#undef REPET_SWAP
}

// 8-point DFT in registers, natural order in and out (n = n1*2 + m, k = k1 + 4*k').
__device__ __forceinline__ void dft8(float2 (&a)[8]) {
    dft4(a[0], a[2], a[4], a[6]);  // m = 0: u0[k1] in a[2*k1]
    dft4(a[1], a[3], a[5], a[7]);  // m = 1: u1[k1] in a[2*k1 + 1]
    float2 u1_1 = make_float2((a[3].x + a[3].y) * REPET_SQRT1_2, (a[3].y - a[3].x) * REPET_SQRT1_2);   // W_8^1
    float2 u1_2 = cmul_mi(a[5]);                                                                        // W_8^2
    float2 u1_3 = make_float2((a[7].y - a[7].x) * REPET_SQRT1_2, -(a[7].x + a[7].y) * REPET_SQRT1_2);  // W_8^3
    float2 u0_0 = a[0], u0_1 = a[2], u0_2 = a[4], u0_3 = a[6], u1_0 = a[1];
    a[0] = cadd(u0_0, u1_0);
    a[4] = csub(u0_0, u1_0);
    a[1] = cadd(u0_1, u1_1);
    a[5] = csub(u0_1, u1_1);
    a[2] = cadd(u0_2, u1_2);
    a[6] = csub(u0_2, u1_2);
    a[3] = cadd(u0_3, u1_3);
    a[7] = csub(u0_3, u1_3);
}

// Per-thread constant twiddles of stage 1: W_2048^(m*k1), k1 = 1..15, from the table
// tw1g[(k1-1)*128 + m] (built in double precision on the host, repet_abi.cu).
struct Twiddle1 {
    float2 w[15];
    __device__ __forceinline__ void load(const float2* __restrict__ tw1g, int t) {
#pragma unroll
        for (int k1 = 1; k1 < 16; ++k1) w[k1 - 1] = __ldg(&tw1g[(k1 - 1) * FFT_THREADS + t]);
    }
};

// stage 1: r[n1] = x[n1*128 + t] on entry; writes y1 to dst.
__device__ __forceinline__ void fft_stage1(float2 (&r)[16], const Twiddle1& tw, float2* __restrict__ dst, int t) {
    dft16(r);
    dst[t] = r[0];
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) dst[k1 * 129 + t] = cmul(r[k1], tw.w[k1 - 1]);
}

// stage 2: reads y1 from src, writes y2 to dst.  s_tw2[k2*8 + m2] = W_128^(m2*k2).
__device__ __forceinline__ void fft_stage2(float2 (&r)[16], const float2* __restrict__ src, float2* __restrict__ dst,
                                           const float2* __restrict__ s_tw2, int t) {
    const int k1 = t & 15, m2 = t >> 4;
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) r[n2] = src[k1 * 129 + n2 * 8 + m2];
    dft16(r);
    dst[m2 * 16 + k1] = r[0];
#pragma unroll
    for (int k2 = 1; k2 < 16; ++k2) dst[(k2 * 8 + m2) * 16 + k1] = cmul(r[k2], s_tw2[k2 * 8 + m2]);
}

// The two output columns (residues mod 256) a thread owns after stage 3: t and 256 - t, except
// thread 0 which owns the two self-mirrored columns 0 and 128.  Column c and column 256 - c hold
// each other's mirror bins (2048 - (c + 256 k3) = (256 - c) + 256 (7 - k3)), so a thread has
// Z[k] AND Z[2048 - k] in registers: the Hermitian split of two packed real channels needs no
// further exchange.  Time-domain use: samples n and n + 1024 sit in the same column (k3, k3 + 4).
__device__ __forceinline__ int fft_out_column(int t, int h) { return t == 0 ? 128 * h : (h == 0 ? t : 256 - t); }

// stage 3: reads y2 from src; on exit r[h*8 + k3] = Z[fft_out_column(t, h) + 256*k3].
__device__ __forceinline__ void fft_stage3(float2 (&r)[16], const float2* __restrict__ src, int t) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int pi = fft_out_column(t, h);
        const float2* col = src + (pi >> 4) * 128 + (pi & 15);
        float2 c[8];
#pragma unroll
        for (int m2 = 0; m2 < 8; ++m2) c[m2] = col[m2 * 16];
        dft8(c);
#pragma unroll
        for (int k3 = 0; k3 < 8; ++k3) r[h * 8 + k3] = c[k3];
    }
}

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
