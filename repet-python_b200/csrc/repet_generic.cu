// General-size float64 paths of the transform helpers: _stft / _istft for ANY window length and step
// (repet.py:1001-1105), _acorr / _beatspectrum for ANY number of rows (repet.py:1108-1158).
//
// The drivers and the fast helper entry points run register-blocked fp32 transforms of 512 / 1024 / 2048 points
// with a hop of half a window; the reference's private helpers take any window, any step and any matrix size, and
// its examples call them directly (README.md:79-81).  This file is the catch-all behind those helpers: batched
// complex float64 FFTs of any length -- Stockham radix-2 passes through global memory for powers of two,
// Bluestein's chirp-z on top of them for everything else -- compiled once, independent of REPET_WIN_N.
// Throughput is not the point here (a 30 s clip takes ~1 ms); float64 accuracy and generality are.
#include "repet_internal.h"

#include <cmath>

namespace {

using repet::fail;
using repet::align_up;
using repet::Bump;

__device__ __forceinline__ double2 zmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}

// One radix-2 Stockham pass (autosort: natural order in, natural order out after log2(L) passes):
// butterfly j of row `row` combines src[j] and src[j + L/2] with the twiddle of its position inside the current
// sub-transform of length 2 * ns and writes to dst[(j / ns) * 2 ns + j % ns (+ ns)].
__global__ void __launch_bounds__(256)
k_stockham_pass(const double2* __restrict__ src, double2* __restrict__ dst, int L, long long n_rows, int ns, int inverse) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int half = L >> 1;
    const long long row = gid / half;
    if (row >= n_rows) return;
    const int j = (int)(gid - row * half);
    const int k = j & (ns - 1);
    double sn, cs;
    sincospi((inverse ? 1.0 : -1.0) * (double)k / (double)ns, &sn, &cs);
    const double2* __restrict__ s = src + row * L;
    double2* __restrict__ d = dst + row * L;
    const double2 a = s[j];
    const double2 b = zmul(s[j + half], make_double2(cs, sn));
    const int o = ((j - k) << 1) + k;
    d[o] = make_double2(a.x + b.x, a.y + b.y);
    d[o + ns] = make_double2(a.x - b.x, a.y - b.y);
}

// in-place-looking batched FFT of power-of-two length L: data and tmp are [n_rows][L]; the result ends in `data`
// (unnormalised in both directions)
void fft_pow2(cudaStream_t st, double2* data, double2* tmp, int L, long long n_rows, bool inverse) {
    if (L < 2) return;
    int passes = 0;
    for (int v = L; v > 1; v >>= 1) ++passes;
    double2* src = data;
    double2* dst = tmp;
    if (passes & 1) {  // odd number of passes: start from tmp so that the last pass writes `data`
        cudaMemcpyAsync(tmp, data, (size_t)n_rows * L * sizeof(double2), cudaMemcpyDeviceToDevice, st);
        src = tmp;
        dst = data;
    }
    const long long threads = n_rows * (L >> 1);
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    for (int ns = 1; ns < L; ns <<= 1) {
        k_stockham_pass<<<blocks, 256, 0, st>>>(src, dst, L, n_rows, ns, inverse ? 1 : 0);
        double2* t = src;
        src = dst;
        dst = t;
    }
}

int next_pow2(long long v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---- Bluestein: DFT of any length N as a circular convolution of length M = pow2 >= 2N - 1 ----------------
// X[k] = conj(c[k]) * sum_n (x[n] conj(c[n])) c[k - n],  c[n] = exp(i pi n^2 / N)   (forward; the inverse
// conjugates c).  n^2 is reduced mod 2N in integers so that the chirp stays exact for long transforms.
__device__ __forceinline__ double2 chirp(long long n, int N, int inverse) {
    const long long r = (n * n) % (2LL * N);
    double sn, cs;
    sincospi((double)r / (double)N, &sn, &cs);
    return make_double2(cs, inverse ? -sn : sn);
}
__global__ void k_bluestein_filter(double2* __restrict__ filt, int N, int M, int inverse) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    double2 v = make_double2(0.0, 0.0);
    if (m < N) v = chirp(m, N, inverse);
    else if (M - m < N) v = chirp(M - m, N, inverse);
    filt[m] = v;
}
__global__ void k_bluestein_pre(const double2* __restrict__ x, double2* __restrict__ a, int N, int M, long long n_rows,
                                int inverse) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = gid / M;
    if (row >= n_rows) return;
    const int m = (int)(gid - row * M);
    double2 v = make_double2(0.0, 0.0);
    if (m < N) {
        const double2 c = chirp(m, N, inverse);
        v = zmul(x[row * N + m], make_double2(c.x, -c.y));
    }
    a[gid] = v;
}
__global__ void k_pointwise_mul(double2* __restrict__ a, const double2* __restrict__ f, int M, long long n_rows) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_rows * M) return;
    a[gid] = zmul(a[gid], f[gid % M]);
}
__global__ void k_bluestein_post(const double2* __restrict__ a, double2* __restrict__ x, int N, int M, long long n_rows,
                                 int inverse) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = gid / N;
    if (row >= n_rows) return;
    const int k = (int)(gid - row * N);
    const double2 c = chirp(k, N, inverse);
    const double2 v = a[row * M + k];
    const double s = 1.0 / (double)M;  // the inverse transform of the convolution
    x[gid] = zmul(make_double2(v.x * s, v.y * s), make_double2(c.x, -c.y));
}

// workspace (in double2 elements) fft_any needs besides `data` [n_rows][N]
size_t fft_any_workspace(int N, long long n_rows) {
    if ((N & (N - 1)) == 0) return (size_t)n_rows * N;
    const int M = next_pow2(2LL * N - 1);
    return 2 * (size_t)n_rows * M + (size_t)2 * M;
}

// batched DFT of any length N, unnormalised, result in `data`
void fft_any(cudaStream_t st, double2* data, double2* ws, int N, long long n_rows, bool inverse) {
    if ((N & (N - 1)) == 0) {
        fft_pow2(st, data, ws, N, n_rows, inverse);
        return;
    }
    const int M = next_pow2(2LL * N - 1);
    double2* a = ws;
    double2* tmp = a + (size_t)n_rows * M;
    double2* filt = tmp + (size_t)n_rows * M;
    double2* filt_tmp = filt + M;
    const int inv = inverse ? 1 : 0;
    k_bluestein_filter<<<(M + 255) / 256, 256, 0, st>>>(filt, N, M, inv);
    fft_pow2(st, filt, filt_tmp, M, 1, false);
    const long long total = n_rows * M;
    k_bluestein_pre<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(data, a, N, M, n_rows, inv);
    fft_pow2(st, a, tmp, M, n_rows, false);
    k_pointwise_mul<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, filt, M, n_rows);
    fft_pow2(st, a, tmp, M, n_rows, true);
    const long long outs = n_rows * N;
    k_bluestein_post<<<(unsigned)((outs + 255) / 256), 256, 0, st>>>(a, data, N, M, n_rows, inv);
}

// ---- _stft ---------------------------------------------------------------------------------------------------
// frames[j][n] = w[n] * padded[j*step + n], padded = floor(N/2) zeros + signal + zeros     repet.py:1018-1055
__global__ void k_frames(const double* __restrict__ signal, long long S, const double* __restrict__ window, int N,
                         int step, int pad, long long T, double2* __restrict__ frames) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= T * N) return;
    const long long j = gid / N;
    const int n = (int)(gid - j * N);
    const long long m = j * step + n - pad;
    frames[gid] = make_double2((m >= 0 && m < S) ? window[n] * signal[m] : 0.0, 0.0);
}
// [T][N] -> the reference's (N, T) C-order layout, through a 32 x 32 shared-memory tile
__global__ void k_transpose(const double2* __restrict__ in, long long rows, int cols, double2* __restrict__ out) {
    __shared__ double2 tile[32][33];
    const long long r0 = (long long)blockIdx.y * 32;
    const int c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const long long r = r0 + i;
        const int c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = in[r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i;
        const long long r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[(long long)c * rows + r] = tile[threadIdx.x][i];
    }
}
void transpose(cudaStream_t st, const double2* in, long long rows, int cols, double2* out) {
    dim3 grid((cols + 31) / 32, (unsigned)((rows + 31) / 32));
    k_transpose<<<grid, dim3(32, 8), 0, st>>>(in, rows, cols, out);
}

// ---- _istft --------------------------------------------------------------------------------------------------
// y[m] = (1/g) sum_j real(ifft(Y[:, j]))[m + (N - step) - j*step] over the frames covering it  repet.py:1085-1103
__global__ void k_overlap_add(const double2* __restrict__ frames, int N, int step, long long T, long long n_out,
                              double inv_norm, double* __restrict__ out) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_out) return;
    const long long pos = m + (N - step);  // index into the untrimmed overlap-add buffer
    long long j_hi = pos / step;
    if (j_hi > T - 1) j_hi = T - 1;
    // smallest j with pos - j*step < N: ceil((pos - N + 1) / step), clipped at 0
    const long long j_lo = pos - N + 1 <= 0 ? 0 : (pos - N + step) / step;
    double acc = 0.0;
    for (long long j = j_lo; j <= j_hi; ++j) acc += frames[j * N + (pos - j * step)].x;  // frame order, as the reference adds
    out[m] = acc * inv_norm;
}

// ---- _acorr / _beatspectrum -------------------------------------------------------------------------------------
// data [n_rows][n_cols] (row-major, float64) -> columns as zero-padded complex rows [n_cols][L]
__global__ void k_load_columns(const double* __restrict__ data, int n_rows, int n_cols, int L, double2* __restrict__ z) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_cols * L) return;
    const int c = (int)(gid / L), l = (int)(gid - (long long)c * L);
    z[gid] = make_double2(l < n_rows ? data[(size_t)l * n_cols + c] : 0.0, 0.0);
}
__global__ void k_power(double2* __restrict__ z, long long n) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n) return;
    const double2 v = z[gid];
    z[gid] = make_double2(fma(v.x, v.x, v.y * v.y), 0.0);
}
// autocorrelation[l][c] = real(ifft)[l] / L / (n_rows - l)          repet.py:1129-1137
__global__ void k_store_acorr(const double2* __restrict__ z, int n_rows, int n_cols, int L, double* __restrict__ out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_rows * n_cols) return;
    const int l = (int)(gid / n_cols), c = (int)(gid - (long long)l * n_cols);
    out[gid] = z[(size_t)c * L + l].x / (double)L / (double)(n_rows - l);
}
// contiguous sequences [n_seq][n] -> zero-padded complex rows [n_seq][L]
__global__ void k_load_rows(const double* __restrict__ data, int n, int n_seq, int L, double2* __restrict__ z) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n_seq * L) return;
    const int c = (int)(gid / L), l = (int)(gid - (long long)c * L);
    z[gid] = make_double2(l < n ? data[(size_t)c * n + l] : 0.0, 0.0);
}
// beat[l] += sum over this chunk's rows of their unbiased autocorrelation at lag l (repet.py:1135-1137, 1156):
// fixed-order sum, one thread per lag
__global__ void k_accumulate_beat(const double2* __restrict__ z, int n, int n_seq, int L, double* __restrict__ beat) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    double acc = beat[l];
    const double scale = 1.0 / (double)L / (double)(n - l);
    for (int c = 0; c < n_seq; ++c) acc += z[(size_t)c * L + l].x * scale;
    beat[l] = acc;
}
__global__ void k_scale_beat(double* __restrict__ beat, int n, int n_seq) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < n) beat[l] /= (double)n_seq;
}

int acorr_device(repet_handle* h, const double* d_data, int n_rows, int n_cols, double* d_out, unsigned char* ws) {
    // any L >= 2 n_rows - 1 gives the same linear autocorrelation as the reference's length 2 n_rows (quirk Q13)
    const int L = next_pow2(2LL * n_rows);
    cudaStream_t st = h->stream;
    Bump bump(ws);
    double2* z = bump.take<double2>((size_t)n_cols * L);
    double2* tmp = bump.take<double2>((size_t)n_cols * L);
    const long long total = (long long)n_cols * L;
    k_load_columns<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_data, n_rows, n_cols, L, z);
    fft_pow2(st, z, tmp, L, n_cols, false);
    k_power<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(z, total);
    fft_pow2(st, z, tmp, L, n_cols, true);
    const long long outs = (long long)n_rows * n_cols;
    k_store_acorr<<<(unsigned)((outs + 255) / 256), 256, 0, st>>>(z, n_rows, n_cols, L, d_out);
    h->launches += 3 + 2 * 20;
    return REPET_OK;
}

}  // namespace

extern "C" {

int repet_stft_frames(int64_t n_samples, int window_length, int step_length) {
    if (window_length < 1 || step_length < 1 || n_samples < 0) return 0;
    const int64_t pad = window_length / 2;
    const int64_t span = n_samples + 2 * pad - window_length;
    // ceil(span / step) + 1 with Python's rounding for negative spans (repet.py:1021-1028)
    const int64_t q = span >= 0 ? (span + step_length - 1) / step_length : -((-span) / step_length);
    return (int)(q + 1);
}

int repet_stft_f64(repet_handle* h, const double* signal, int64_t n_samples, const double* window, int window_length,
                   int step_length, double* spectrum, int32_t* n_frames_out) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!signal || !window || !spectrum || n_samples < 0 || window_length < 1 || step_length < 1)
        return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    const int N = window_length;
    const long long T = repet_stft_frames(n_samples, window_length, step_length);
    if (n_frames_out) *n_frames_out = (int32_t)T;
    if (T < 1) return fail(h, REPET_E_INVALID_ARG, "signal too short for one frame");
    CU(cudaSetDevice(h->device));
    const size_t elems = (size_t)T * N;
    const size_t need = align_up((size_t)std::max<int64_t>(n_samples, 1) * sizeof(double)) + align_up((size_t)N * sizeof(double)) +
                        2 * align_up(elems * sizeof(double2)) + align_up(fft_any_workspace(N, T) * sizeof(double2)) + 1024;
    int rc = repet::ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    double* d_sig = bump.take<double>((size_t)std::max<int64_t>(n_samples, 1));
    double* d_win = bump.take<double>(N);
    double2* frames = bump.take<double2>(elems);
    double2* out = bump.take<double2>(elems);
    double2* ws = bump.take<double2>(fft_any_workspace(N, T));
    cudaStream_t st = h->stream;
    if (n_samples > 0) CU(cudaMemcpyAsync(d_sig, signal, (size_t)n_samples * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_win, window, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, st));
    k_frames<<<(unsigned)((elems + 255) / 256), 256, 0, st>>>(d_sig, n_samples, d_win, N, step_length, N / 2, T, frames);
    fft_any(st, frames, ws, N, T, false);
    transpose(st, frames, T, N, out);
    h->launches += 2 + 12;
    CU(cudaMemcpyAsync(spectrum, out, elems * sizeof(double2), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

int repet_istft_f64(repet_handle* h, const double* spectrum, int window_length, int n_frames, const double* window,
                    int step_length, double* signal, int64_t* n_samples_out) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!spectrum || !window || !signal || window_length < 1 || step_length < 1 || n_frames < 1)
        return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    const int N = window_length;
    const long long T = n_frames;
    // untrimmed length T*step + (N - step), minus N - step at both ends (repet.py:1079, 1098-1100)
    const long long n_out = T * step_length - (long long)(N - step_length);
    if (n_samples_out) *n_samples_out = std::max<long long>(n_out, 0);
    if (n_out <= 0) return REPET_OK;
    double gain = 0.0;
    for (int i = 0; i < N; i += step_length) gain += window[i];  // sum(window[0:N:step]), repet.py:1103
    CU(cudaSetDevice(h->device));
    const size_t elems = (size_t)T * N;
    const size_t need = 2 * align_up(elems * sizeof(double2)) + align_up(fft_any_workspace(N, T) * sizeof(double2)) +
                        align_up((size_t)n_out * sizeof(double)) + 1024;
    int rc = repet::ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    double2* in = bump.take<double2>(elems);
    double2* frames = bump.take<double2>(elems);
    double2* ws = bump.take<double2>(fft_any_workspace(N, T));
    double* out = bump.take<double>((size_t)n_out);
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(in, spectrum, elems * sizeof(double2), cudaMemcpyHostToDevice, st));
    transpose(st, in, N, (int)T, frames);  // (N, T) -> [T][N]
    fft_any(st, frames, ws, N, T, true);
    k_overlap_add<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(frames, N, step_length, T, n_out, 1.0 / ((double)N * gain), out);
    h->launches += 2 + 12;
    CU(cudaMemcpyAsync(signal, out, (size_t)n_out * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

int repet_acorr_f64(repet_handle* h, const double* data, int n_rows, int n_columns, double* autocorrelation) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!data || !autocorrelation || n_rows < 1 || n_columns < 1) return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    CU(cudaSetDevice(h->device));
    const int L = next_pow2(2LL * n_rows);
    const size_t n = (size_t)n_rows * n_columns;
    // columns in chunks so that the two transform buffers stay below ~4 GB
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_columns, ((size_t)2 << 30) / ((size_t)L * sizeof(double2))));
    const size_t need = 2 * align_up(n * sizeof(double)) + 2 * align_up((size_t)chunk * L * sizeof(double2)) +
                        2 * align_up((size_t)n_rows * chunk * sizeof(double)) + 1024;
    int rc = repet::ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    double* d_in = bump.take<double>(n);
    double* d_out = bump.take<double>(n);
    double* c_in = bump.take<double>((size_t)n_rows * chunk);
    double* c_out = bump.take<double>((size_t)n_rows * chunk);
    unsigned char* ws = h->arena + bump.off;
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(d_in, data, n * sizeof(double), cudaMemcpyHostToDevice, st));
    for (int c0 = 0; c0 < n_columns; c0 += chunk) {
        const int g = std::min(chunk, n_columns - c0);
        if (g == n_columns) {
            if ((rc = acorr_device(h, d_in, n_rows, n_columns, d_out, ws))) return rc;
        } else {  // a block of columns: gather it into a dense [n_rows][g] matrix, scatter the result back
            CU(cudaMemcpy2DAsync(c_in, (size_t)g * sizeof(double), d_in + c0, (size_t)n_columns * sizeof(double),
                                 (size_t)g * sizeof(double), n_rows, cudaMemcpyDeviceToDevice, st));
            if ((rc = acorr_device(h, c_in, n_rows, g, c_out, ws))) return rc;
            CU(cudaMemcpy2DAsync(d_out + c0, (size_t)n_columns * sizeof(double), c_out, (size_t)g * sizeof(double),
                                 (size_t)g * sizeof(double), n_rows, cudaMemcpyDeviceToDevice, st));
        }
    }
    CU(cudaMemcpyAsync(autocorrelation, d_out, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

int repet_beatspectrum_f64(repet_handle* h, const double* spectrogram, int n_frequencies, int n_times, double* beat) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!spectrogram || !beat || n_frequencies < 1 || n_times < 1) return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    CU(cudaSetDevice(h->device));
    // _beatspectrum(V) = mean(_acorr(V.T), axis=1) (repet.py:1153-1156): every frequency row of V is one contiguous
    // sequence of n_times samples; rows go through the transforms in chunks, the mean is accumulated row by row in a
    // fixed order
    const int L = next_pow2(2LL * n_times);
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_frequencies, ((size_t)2 << 30) / ((size_t)L * sizeof(double2))));
    const size_t n = (size_t)n_frequencies * n_times;
    const size_t need = align_up(n * sizeof(double)) + 2 * align_up((size_t)chunk * L * sizeof(double2)) +
                        align_up((size_t)n_times * sizeof(double)) + 1024;
    int rc = repet::ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    double* d_v = bump.take<double>(n);
    double2* z = bump.take<double2>((size_t)chunk * L);
    double2* tmp = bump.take<double2>((size_t)chunk * L);
    double* d_beat = bump.take<double>(n_times);
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(d_v, spectrogram, n * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d_beat, 0, (size_t)n_times * sizeof(double), st));
    for (int f0 = 0; f0 < n_frequencies; f0 += chunk) {
        const int g = std::min(chunk, n_frequencies - f0);
        const long long total = (long long)g * L;
        k_load_rows<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_v + (size_t)f0 * n_times, n_times, g, L, z);
        fft_pow2(st, z, tmp, L, g, false);
        k_power<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(z, total);
        fft_pow2(st, z, tmp, L, g, true);
        k_accumulate_beat<<<(n_times + 127) / 128, 128, 0, st>>>(z, n_times, g, L, d_beat);
        h->launches += 3 + 2 * 20;
    }
    k_scale_beat<<<(n_times + 127) / 128, 128, 0, st>>>(d_beat, n_times, n_frequencies);
    CU(cudaMemcpyAsync(beat, d_beat, (size_t)n_times * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

}  // extern "C"
