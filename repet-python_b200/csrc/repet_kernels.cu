// Hand-written sm_100a kernels of the REPET separation path (fp32 arithmetic, fp64 for the
// one-per-clip period search).  Reference semantics: /root/reference/repet.py, cited per kernel.
//
// Data layout in HBM (per batch item = clip or 10 s segment):
//   audio  [clip][channel][sample]  fp32 planar
//   X      [item][frame][channel][1024] float2  half spectrum; bin 0 = (DC.re, Nyquist.re)
//   P      [item][frame][1032] fp32   (mean_c |X|)^2 (beat spectrum) or mean_c |X| (similarity)
//   model  [item][channel][phase][1032] fp32   median over the period-strided frames
// Frames are the slow axis, bins the fast one: every gather along time (period-strided
// medians, similar-frame medians) is a coalesced row read.
#include "repet_kernels.cuh"
#include "fft2048.cuh"

namespace repet {

// np.finfo(float).eps of the reference's soft mask (repet.py:1446); representable in fp32.
#define REPET_EPS 2.220446049250313e-16f

// ------------------------------------------------------------------------------------------
// k_stft  --  _stft + abs + channel mean (+ square)      repet.py:1001-1060, 158, 162, 667
// One CTA of 128 threads walks K consecutive frames of one item.  Both channels ride in one
// complex transform (z = w*(xL + i xR)) and are separated by Hermitian symmetry afterwards.
// The second half of frame j is the first half of frame j+1 and lands in the SAME thread
// (n = n1*128 + t), so the overlap is carried in registers: 16 new samples per thread-frame.
// Algorithmic bytes per frame: 2*4 KB audio in, 16 KB X out, 4.1 KB P out.
// ------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(FFT_THREADS)
k_stft(const float* __restrict__ audio, Geom g, const float* __restrict__ window, FftTables tb,
       float2* __restrict__ X, float* __restrict__ P, int pmode, int K) {
    __shared__ float2 s_buf[2][FFT_BUF];
    __shared__ float2 s_tw2[128];
    const int t = threadIdx.x;
    const int item = blockIdx.y;
    const int j0 = blockIdx.x * K;
    const int j1 = min(j0 + K, g.T);
    s_tw2[t] = tb.tw2[t];
    Twiddle1 tw;
    tw.load(tb.tw1, t);
    float wv[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) wv[n1] = __ldg(&window[n1 * 128 + t]);
    const int clip = item / g.seg_per_clip, sg = item - clip * g.seg_per_clip;
    const float* __restrict__ a0 = audio + g.first_offset + (long long)clip * g.clip_stride + (long long)sg * g.seg_stride;
    const float* __restrict__ a1 = a0 + g.chan_stride;
    float cl[8], cr[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cl[i] = cr[i] = 0.f;
    __syncthreads();
    int par = 0;
    for (int j = j0; j < j1; ++j) {
        float2 r[16];
        const long long base = (long long)(j - 1) * HOP + t;
        const bool carry = j > j0;
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            float xl, xr = 0.f;
            if (n1 < 8 && carry) {
                xl = cl[n1];
                xr = cr[n1];
            } else {
                const long long idx = base + n1 * 128;
                const bool ok = idx >= 0 && idx < g.S;
                xl = ok ? __ldg(a0 + idx) : 0.f;
                if (NCH == 2) xr = ok ? __ldg(a1 + idx) : 0.f;
            }
            if (n1 >= 8) {
                cl[n1 - 8] = xl;
                cr[n1 - 8] = xr;
            }
            r[n1] = make_float2(wv[n1] * xl, wv[n1] * xr);
        }
        float2* A = s_buf[par];
        float2* B = s_buf[par ^ 1];
        fft_stage1(r, tw, A, t);
        __syncthreads();
        fft_stage2(r, A, B, s_tw2, t);
        __syncthreads();
        fft_stage3(r, B, t);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k3 = 0; k3 < 8; ++k3) A[t + 128 * h + 256 * k3] = r[h * 8 + k3];
        __syncthreads();
        const size_t frame = (size_t)item * g.T + j;
        float2* __restrict__ xrow = X + frame * (size_t)(NCH * XPITCH);
        float* __restrict__ prow = P ? P + frame * (size_t)PPITCH : nullptr;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = t + 128 * i;
            const float2 a = A[k];
            float2 xl, xr;
            float ml, mr;
            if (k == 0) {
                const float2 ny = A[1024];
                xl = make_float2(a.x, ny.x);
                xr = make_float2(a.y, ny.y);
                ml = fabsf(a.x);
                mr = fabsf(a.y);
                if (prow) {
                    const float mn = NCH == 2 ? 0.5f * (fabsf(ny.x) + fabsf(ny.y)) : fabsf(ny.x);
                    prow[1024] = pmode == P_POWER ? mn * mn : mn;
                }
            } else {
                const float2 b = A[2048 - k];
                xl = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y - b.y));
                xr = make_float2(0.5f * (a.y + b.y), 0.5f * (b.x - a.x));
                ml = cmag(xl);
                mr = cmag(xr);
            }
            xrow[k] = xl;
            if (NCH == 2) xrow[XPITCH + k] = xr;
            if (prow) {
                const float mean = NCH == 2 ? 0.5f * (ml + mr) : ml;
                prow[k] = pmode == P_POWER ? mean * mean : mean;
            }
        }
        par ^= 1;
    }
}

void launch_stft(cudaStream_t st, const float* audio, Geom g, int nch, const float* window, FftTables tb, float2* X,
                 float* P, int pmode, int frames_per_cta) {
    dim3 grid((g.T + frames_per_cta - 1) / frames_per_cta, g.n_items);
    if (nch == 2)
        k_stft<2><<<grid, FFT_THREADS, 0, st>>>(audio, g, window, tb, X, P, pmode, frames_per_cta);
    else
        k_stft<1><<<grid, FFT_THREADS, 0, st>>>(audio, g, window, tb, X, P, pmode, frames_per_cta);
}

// ------------------------------------------------------------------------------------------
// k_beat  --  _acorr / _beatspectrum in the FFT domain        repet.py:1108-1158, 1161-1206
// b[l] = mean_f (1/(R-l)) sum_t P[f,t] P[f,t+l].  Per frequency row the autocorrelation is
// the inverse transform of |FFT_L(P_f)|^2 (L = 2048 >= R + max lag, so no circular alias for
// the lags consumed); by linearity the PSDs are SUMMED over f first and inverted once per clip
// (k_periods).  Two rows f, f+1 ride in one complex transform: |Fa|^2 + |Fb|^2 =
// (|Z[k]|^2 + |Z[-k]|^2)/2, so the kernel only accumulates |Z[k]|^2 per thread-owned bin, in
// registers, over all the rows the CTA owns -- no exchange, no atomics, deterministic.
// The CTA stages an 8-row x R tile of P (32 B per frame row: full sectors) transposed in
// shared memory.  Algorithmic bytes: P read once (4.1 KB per frame).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FFT_THREADS)
k_beat(const float* __restrict__ P, int T, int t_first, int t_len, int seg_step, int n_seg, FftTables tb,
       float* __restrict__ psd_part, int n_parts, int f_per_part, int TP) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    float2* s_bufA = reinterpret_cast<float2*>(s_raw);
    float2* s_bufB = s_bufA + FFT_BUF;
    float2* s_tw2 = s_bufB + FFT_BUF;
    float2* s_tile = s_tw2 + 128;  // [4][TP]
    const int t = threadIdx.x;
    const int bi = blockIdx.y;
    const int item = bi / n_seg, sg = bi - item * n_seg;
    const int ts = t_first + sg * seg_step;
    const int f_begin = blockIdx.x * f_per_part;
    const int f_end = min(NBIN, f_begin + f_per_part);
    s_tw2[t] = tb.tw2[t];
    Twiddle1 tw;
    tw.load(tb.tw1, t);
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    const float* __restrict__ Pitem = P + (size_t)item * T * PPITCH;
    for (int f0 = f_begin; f0 < f_end; f0 += 8) {
        for (int idx = t; idx < 2 * t_len; idx += FFT_THREADS) {
            const int row = idx >> 1, half = idx & 1;
            const int frame = ts + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (frame >= 0 && frame < T)
                v = __ldg(reinterpret_cast<const float4*>(Pitem + (size_t)frame * PPITCH + f0 + 4 * half));
            const int f = f0 + 4 * half;
            if (f + 0 >= f_end) v.x = 0.f;
            if (f + 1 >= f_end) v.y = 0.f;
            if (f + 2 >= f_end) v.z = 0.f;
            if (f + 3 >= f_end) v.w = 0.f;
            s_tile[(2 * half) * TP + row] = make_float2(v.x, v.y);
            s_tile[(2 * half + 1) * TP + row] = make_float2(v.z, v.w);
        }
        __syncthreads();
        const int npairs = (min(8, f_end - f0) + 1) >> 1;
        for (int pr = 0; pr < npairs; ++pr) {
            float2 r[16];
            const float2* __restrict__ col = s_tile + pr * TP;
#pragma unroll
            for (int n1 = 0; n1 < 16; ++n1) {
                const int n = n1 * 128 + t;
                r[n1] = n < t_len ? col[n] : make_float2(0.f, 0.f);
            }
            fft_stage1(r, tw, s_bufA, t);
            __syncthreads();
            fft_stage2(r, s_bufA, s_bufB, s_tw2, t);
            __syncthreads();
            fft_stage3(r, s_bufB, t);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(r[i].x, r[i].x, fmaf(r[i].y, r[i].y, acc[i]));
        }
    }
    float* __restrict__ out = psd_part + ((size_t)bi * n_parts + blockIdx.x) * BEAT_L;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int k3 = 0; k3 < 8; ++k3) out[t + 128 * h + 256 * k3] = acc[h * 8 + k3];
}

static size_t beat_smem_bytes(int t_len, int* TP_out) {
    int TP = ((t_len + 7) / 8) * 8 + 4;
    *TP_out = TP;
    return (size_t)(2 * FFT_BUF + 128 + 4 * TP) * sizeof(float2);
}

void launch_beat(cudaStream_t st, const float* P, int n_items, int T, int t_first, int t_len, int seg_step, int n_seg,
                 FftTables tb, float* psd_part, int n_parts, int f_per_part) {
    int TP;
    size_t smem = beat_smem_bytes(t_len, &TP);
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(k_beat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    dim3 grid(n_parts, n_items * n_seg);
    k_beat<<<grid, FFT_THREADS, smem, st>>>(P, T, t_first, t_len, seg_step, n_seg, tb, psd_part, n_parts, f_per_part, TP);
}

// ------------------------------------------------------------------------------------------
// k_periods  --  inverse transform of the summed PSD, unbiased normalisation, argmax
//                                                             repet.py:1129-1137, 1156, 1249-1291
// One CTA per clip (or per adaptive segment), fp64: the summed PSD is dominated by its DC
// bin, and doing this single small transform in double keeps the fp32 front end's error out
// of the argmax.  period = first argmax over lags [lo, hi) + 1 (quirks Q1, Q2).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_periods(const float* __restrict__ psd_part, int n_parts, int t_len, double norm_rows, int lag_lo, int lag_hi,
          int out_lo, int out_hi, double* __restrict__ beat_out, int beat_pitch, int* __restrict__ period,
          double* __restrict__ stats) {
    __shared__ double s_psd[BEAT_L / 2 + 1];
    __shared__ double s_cos[BEAT_L];
    __shared__ double s_b[BEAT_L];
    const int t = threadIdx.x;
    const int bi = blockIdx.x;
    for (int k = t; k < BEAT_L; k += 128) {
        double a = 0.0;
        const float* __restrict__ src = psd_part + (size_t)bi * n_parts * BEAT_L + k;
        for (int part = 0; part < n_parts; ++part) a += (double)src[(size_t)part * BEAT_L];
        s_b[k] = a;
        s_cos[k] = cospi((double)k / (double)(BEAT_L / 2));
    }
    __syncthreads();
    for (int k = t; k <= BEAT_L / 2; k += 128) s_psd[k] = 0.5 * (s_b[k] + s_b[(BEAT_L - k) & (BEAT_L - 1)]);
    __syncthreads();
    int l0 = out_lo, l1 = out_hi;
    if (period) {
        l0 = (out_hi > out_lo) ? min(out_lo, lag_lo) : lag_lo;
        l1 = (out_hi > out_lo) ? max(out_hi, lag_hi) : lag_hi;
    }
    for (int l = l0 + t; l < l1; l += 128) {
        double s = 0.0;
        for (int k = 1; k < BEAT_L / 2; ++k) s = fma(s_psd[k], s_cos[(k * l) & (BEAT_L - 1)], s);
        s = 2.0 * s + s_psd[0] + ((l & 1) ? -s_psd[BEAT_L / 2] : s_psd[BEAT_L / 2]);
        s_b[l] = s / (double)BEAT_L / ((double)(t_len - l) * norm_rows);
    }
    __syncthreads();
    if (beat_out)
        for (int l = out_lo + t; l < out_hi; l += 128) beat_out[(size_t)bi * beat_pitch + (l - out_lo)] = s_b[l];
    if (period && t == 0) {
        double best = s_b[lag_lo], second = -1.0e300;
        int arg = lag_lo, arg2 = -1;
        for (int l = lag_lo + 1; l < lag_hi; ++l) {
            const double v = s_b[l];
            if (v > best) {
                second = best;
                arg2 = arg;
                best = v;
                arg = l;
            } else if (v > second) {
                second = v;
                arg2 = l;
            }
        }
        period[bi] = arg + 1;
        if (stats) {
            stats[4 * bi + 0] = best;
            stats[4 * bi + 1] = second;
            stats[4 * bi + 2] = (double)arg;
            stats[4 * bi + 3] = (double)arg2;
        }
    }
}

void launch_periods(cudaStream_t st, const float* psd_part, int n_beat_items, int n_parts, int t_len, double norm_rows,
                    int lag_lo, int lag_hi, int out_lo, int out_hi, double* beat_out, int beat_pitch, int* period,
                    double* stats) {
    k_periods<<<n_beat_items, 128, 0, st>>>(psd_part, n_parts, t_len, norm_rows, lag_lo, lag_hi, out_lo, out_hi,
                                            beat_out, beat_pitch, period, stats);
}

// ------------------------------------------------------------------------------------------
// Median of n <= NP values held in registers: pad with -inf below and +inf above so that the
// median always sits at sorted positions NP/2-1 (odd n) or NP/2-1, NP/2 (even n: mean of the
// two middle values, as np.median does), then run a bitonic network with static indices.
// ------------------------------------------------------------------------------------------
template <int NP>
__device__ __forceinline__ float median_network(float (&v)[NP], int n) {
#pragma unroll
    for (int k = 2; k <= NP; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const float lo = fminf(v[i], v[l]), hi = fmaxf(v[i], v[l]);
                    if ((i & k) == 0) {
                        v[i] = lo;
                        v[l] = hi;
                    } else {
                        v[i] = hi;
                        v[l] = lo;
                    }
                }
            }
        }
    }
    return (n & 1) ? v[NP / 2 - 1] : 0.5f * (v[NP / 2 - 1] + v[NP / 2]);
}

// value to load into slot s of an NP-wide network holding n real values: slots
// [lo_pad, lo_pad + n) are data, below -inf, above +inf
__device__ __forceinline__ int median_lo_pad(int NP, int n) { return (NP - n) >> 1; }

// Fallback for n > 32: rank selection straight from memory (O(n^2), exact).  `fetch(s)` returns
// value s.  Returns the median with np.median's even-count rule.
template <typename Fetch>
__device__ float median_by_rank(int n, Fetch fetch) {
    const int k_lo = (n - 1) >> 1, k_hi = n >> 1;
    float v_lo = 0.f, v_hi = 0.f;
    bool got_lo = false, got_hi = false;
    for (int i = 0; i < n && !(got_lo && got_hi); ++i) {
        const float vi = fetch(i);
        int less = 0, equal = 0;
        for (int j = 0; j < n; ++j) {
            const float vj = fetch(j);
            less += vj < vi;
            equal += vj == vi;
        }
        if (!got_lo && less <= k_lo && k_lo < less + equal) {
            v_lo = vi;
            got_lo = true;
        }
        if (!got_hi && less <= k_hi && k_hi < less + equal) {
            v_hi = vi;
            got_hi = true;
        }
    }
    return 0.5f * (v_lo + v_hi);
}

// magnitude of bin k (0..1024) of one (frame, channel) row of X; bin 0 packs (DC, Nyquist)
__device__ __forceinline__ float row_mag(const float2* __restrict__ row, int k) {
    if (k == 0) return fabsf(__ldg(&row[0]).x);
    if (k == XPITCH) return fabsf(__ldg(&row[0]).y);
    return cmag(__ldg(&row[k]));
}

template <int NP>
__device__ __forceinline__ float strided_median(const float2* __restrict__ base, size_t stride, int n, int k) {
    float v[NP];
    const int lo = median_lo_pad(NP, n);
#pragma unroll
    for (int s = 0; s < NP; ++s) {
        const int d = s - lo;
        v[s] = d < 0 ? -INFINITY : (d < n ? row_mag(base + (size_t)d * stride, k) : INFINITY);
    }
    return median_network<NP>(v, n);
}

// ------------------------------------------------------------------------------------------
// k_model  --  the repeating segment of _mask                      repet.py:1398-1438
// model[f, q] = median over s of V[f, s*p + q]; r = ceil(T/p) values for phases
// q < T-(r-1)p, r-1 otherwise (the reference's zero padding is excluded, quirk Q9).
// Grid (9 bin blocks, pmax phases, items*channels); CTAs of phases q >= p[item] exit.
// Threads run along bins, so each of the n gathers is a coalesced row read of X.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_model(const float2* __restrict__ X, int T, int nch, const int* __restrict__ period, int pmax,
        float* __restrict__ model) {
    const int item = blockIdx.z / nch, c = blockIdx.z - item * nch;
    const int p = period[item];
    const int q = blockIdx.y;
    if (q >= p || p <= 0) return;
    const int k = blockIdx.x * 128 + threadIdx.x;
    if (k > XPITCH) return;
    const int r = (T + p - 1) / p;
    const int k0 = T - (r - 1) * p;
    const int n = q < k0 ? r : r - 1;
    const size_t stride = (size_t)p * nch * XPITCH;
    const float2* __restrict__ base = X + ((size_t)item * T + q) * (size_t)(nch * XPITCH) + (size_t)c * XPITCH;
    float med;
    if (n <= 0)
        med = nanf("");
    else if (n <= 4)
        med = strided_median<4>(base, stride, n, k);
    else if (n <= 8)
        med = strided_median<8>(base, stride, n, k);
    else if (n <= 16)
        med = strided_median<16>(base, stride, n, k);
    else if (n <= 32)
        med = strided_median<32>(base, stride, n, k);
    else
        med = median_by_rank(n, [&](int s) { return row_mag(base + (size_t)s * stride, k); });
    model[(((size_t)item * nch + c) * pmax + q) * PPITCH + k] = med;
}

void launch_model(cudaStream_t st, const float2* X, int n_items, int T, int nch, const int* period, int pmax,
                  float* model) {
    dim3 grid(9, pmax, n_items * nch);
    k_model<<<grid, 128, 0, st>>>(X, T, nch, period, pmax, model);
}

// soft mask of one bin: min with the mixture, (W+eps)/(V+eps)     repet.py:1441-1448 (quirk Q10)
__device__ __forceinline__ float soft_mask(float model, float v) {
    return __fdiv_rn(fminf(model, v) + REPET_EPS, v + REPET_EPS);
}

// ------------------------------------------------------------------------------------------
// k_mask_istft  --  rest of _mask + high-pass + mirror + apply + _istft
//                                                       repet.py:1441-1456, 185-200, 1063-1105
// One CTA of 128 threads produces `nblk` consecutive output blocks of H samples of one item;
// block b = second half of frame b + first half of frame b+1, so the CTA walks nblk+1 frames.
// Per frame: each bin's mask is computed once (both channels), the masked spectra are packed
// Z = YL + i YR with their Hermitian mirror, and ONE inverse complex transform (forward FFT of
// the re/im-swapped input) returns both channels' frames.  The output index n = pi + 256*k3
// keeps both halves of the overlap in the same thread: overlap-add happens in registers.
// MASKED = false is the plain _istft helper.
// Algorithmic bytes per frame: 16 KB X in, 8 KB audio out (+ model rows, L2 resident).
// ------------------------------------------------------------------------------------------
template <int NCH, bool MASKED>
__global__ void __launch_bounds__(FFT_THREADS)
k_mask_istft(const float2* __restrict__ X, Geom g, const int* __restrict__ period, int pmax,
             const float* __restrict__ model, int cutoff, float scale, FftTables tb, float* __restrict__ out,
             int nblk) {
    __shared__ float2 s_buf[2][FFT_BUF];
    __shared__ float2 s_tw2[128];
    const int t = threadIdx.x;
    const int item = blockIdx.y;
    const int b0 = blockIdx.x * nblk;               // first output block
    const int b1 = min(b0 + nblk, g.T - 1);         // one past the last output block
    s_tw2[t] = tb.tw2[t];
    Twiddle1 tw;
    tw.load(tb.tw1, t);
    const int clip = item / g.seg_per_clip, sg = item - clip * g.seg_per_clip;
    float* __restrict__ o0 = out + g.first_offset + (long long)clip * g.clip_stride + (long long)sg * g.seg_stride;
    float* __restrict__ o1 = o0 + g.chan_stride;
    int p = 1;
    if (MASKED) p = period[item];
    float carry_l[8], carry_r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) carry_l[i] = carry_r[i] = 0.f;
    __syncthreads();
    int par = 0;
    for (int j = b0; j <= b1; ++j) {
        float2* A = s_buf[par];
        float2* B = s_buf[par ^ 1];
        const float2* __restrict__ xl_row = X + ((size_t)item * g.T + j) * (size_t)(NCH * XPITCH);
        const float2* __restrict__ xr_row = xl_row + XPITCH;
        const float* __restrict__ ml_row = nullptr;
        const float* __restrict__ mr_row = nullptr;
        if (MASKED) {
            const int q = j % p;
            ml_row = model + (((size_t)item * NCH + 0) * pmax + q) * PPITCH;
            mr_row = model + (((size_t)item * NCH + (NCH - 1)) * pmax + q) * PPITCH;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = t + 128 * i;
            const float2 xl = __ldg(&xl_row[k]);
            float2 xr = make_float2(0.f, 0.f);
            if (NCH == 2) xr = __ldg(&xr_row[k]);
            if (k == 0) {
                // DC (mask kept, quirk Q11) and Nyquist, both purely real
                float m_dc_l = 1.f, m_ny_l = 1.f, m_dc_r = 1.f, m_ny_r = 1.f;
                if (MASKED) {
                    m_dc_l = soft_mask(__ldg(&ml_row[0]), fabsf(xl.x));
                    m_ny_l = (XPITCH <= cutoff) ? 1.f : soft_mask(__ldg(&ml_row[XPITCH]), fabsf(xl.y));
                    if (NCH == 2) {
                        m_dc_r = soft_mask(__ldg(&mr_row[0]), fabsf(xr.x));
                        m_ny_r = (XPITCH <= cutoff) ? 1.f : soft_mask(__ldg(&mr_row[XPITCH]), fabsf(xr.y));
                    }
                }
                // swapped storage: (im, re) of Z = YL + i YR
                A[0] = make_float2(m_dc_r * xr.x, m_dc_l * xl.x);
                A[1024] = make_float2(m_ny_r * xr.y, m_ny_l * xl.y);
            } else {
                float m_l = 1.f, m_r = 1.f;
                if (MASKED) {
                    if (k > cutoff) {
                        m_l = soft_mask(__ldg(&ml_row[k]), cmag(xl));
                        if (NCH == 2) m_r = soft_mask(__ldg(&mr_row[k]), cmag(xr));
                    }
                }
                const float2 yl = make_float2(m_l * xl.x, m_l * xl.y);
                const float2 yr = make_float2(m_r * xr.x, m_r * xr.y);
                // Z[k] = YL + i YR ; Z[N-k] = conj(YL) + i conj(YR) ; stored re/im swapped
                A[k] = make_float2(yl.y + yr.x, yl.x - yr.y);
                A[2048 - k] = make_float2(yr.x - yl.y, yl.x + yr.y);
            }
        }
        __syncthreads();
        float2 r[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) r[n1] = A[n1 * 128 + t];
        fft_stage1(r, tw, B, t);
        __syncthreads();
        fft_stage2(r, B, A, s_tw2, t);
        __syncthreads();
        fft_stage3(r, A, t);
        // r[h*8+k3] = (N*yR, N*yL) at frame sample n = (t + 128h) + 256*k3
        if (j > b0) {
            const long long blk = (long long)(j - 1) * HOP;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int k3 = 0; k3 < 4; ++k3) {
                    const long long m = blk + t + 128 * h + 256 * k3;
                    if (m < g.S) {
                        o0[m] = (carry_l[h * 4 + k3] + r[h * 8 + k3].y) * scale;
                        if (NCH == 2) o1[m] = (carry_r[h * 4 + k3] + r[h * 8 + k3].x) * scale;
                    }
                }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int k3 = 0; k3 < 4; ++k3) {
                carry_l[h * 4 + k3] = r[h * 8 + 4 + k3].y;
                carry_r[h * 4 + k3] = r[h * 8 + 4 + k3].x;
            }
        par ^= 1;
    }
}

void launch_mask_istft(cudaStream_t st, const float2* X, Geom g, int nch, const int* period, int pmax,
                       const float* model, int cutoff, float scale, FftTables tb, float* out, int blocks_per_cta) {
    const int nblocks = g.T - 1;
    dim3 grid((nblocks + blocks_per_cta - 1) / blocks_per_cta, g.n_items);
    if (nch == 2)
        k_mask_istft<2, true><<<grid, FFT_THREADS, 0, st>>>(X, g, period, pmax, model, cutoff, scale, tb, out, blocks_per_cta);
    else
        k_mask_istft<1, true><<<grid, FFT_THREADS, 0, st>>>(X, g, period, pmax, model, cutoff, scale, tb, out, blocks_per_cta);
}

void launch_istft(cudaStream_t st, const float2* X, Geom g, int nch, float scale, FftTables tb, float* out,
                  int blocks_per_cta) {
    const int nblocks = g.T - 1;
    dim3 grid((nblocks + blocks_per_cta - 1) / blocks_per_cta, g.n_items);
    if (nch == 2)
        k_mask_istft<2, false><<<grid, FFT_THREADS, 0, st>>>(X, g, nullptr, 0, nullptr, 0, scale, tb, out, blocks_per_cta);
    else
        k_mask_istft<1, false><<<grid, FFT_THREADS, 0, st>>>(X, g, nullptr, 0, nullptr, 0, scale, tb, out, blocks_per_cta);
}

// ------------------------------------------------------------------------------------------
// k_mask_only  --  _mask as a helper: M[item][c][frame][bin]          repet.py:1441-1456
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_mask_only(const float2* __restrict__ X, int T, int nch, const int* __restrict__ period, int pmax,
            const float* __restrict__ model, float* __restrict__ mask_out) {
    const int item = blockIdx.z / nch, c = blockIdx.z - item * nch;
    const int j = blockIdx.y;
    const int k = blockIdx.x * 128 + threadIdx.x;
    if (k > XPITCH) return;
    const int p = period[item];
    const float2* __restrict__ row = X + ((size_t)item * T + j) * (size_t)(nch * XPITCH) + (size_t)c * XPITCH;
    const float m = model[(((size_t)item * nch + c) * pmax + (j % p)) * PPITCH + k];
    mask_out[(((size_t)item * nch + c) * T + j) * PPITCH + k] = soft_mask(m, row_mag(row, k));
}

void launch_mask_only(cudaStream_t st, const float2* X, int n_items, int T, int nch, const int* period, int pmax,
                      const float* model, float* mask_out) {
    dim3 grid(9, T, n_items * nch);
    k_mask_only<<<grid, 128, 0, st>>>(X, T, nch, period, pmax, model, mask_out);
}

// ------------------------------------------------------------------------------------------
// float64 (S, C) interleaved  <->  fp32 planar [C][S]: the reference API's NumPy convention
// (repet.py:73-77) converted on the device, so the host never touches the samples.
// ------------------------------------------------------------------------------------------
__global__ void k_f64_to_planar(const double* __restrict__ in, long long S, int C, float* __restrict__ out) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= S * C) return;
    const long long s = n / C;
    const int c = (int)(n - s * C);
    out[(long long)c * S + s] = (float)in[n];
}
__global__ void k_planar_to_f64(const float* __restrict__ in, long long S, int C, double* __restrict__ out) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= S * C) return;
    const long long s = n / C;
    const int c = (int)(n - s * C);
    out[n] = (double)in[(long long)c * S + s];
}
void launch_f64_interleaved_to_planar(cudaStream_t st, const double* in, long long S, int C, float* out) {
    const long long n = S * C;
    k_f64_to_planar<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, S, C, out);
}
void launch_planar_to_f64_interleaved(cudaStream_t st, const float* in, long long S, int C, double* out) {
    const long long n = S * C;
    k_planar_to_f64<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, S, C, out);
}

}  // namespace repet
