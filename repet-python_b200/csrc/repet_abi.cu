// C ABI of librepet_b200.so (declared in include/repet_b200.h): handle lifetime, workspace arena,
// transform tables, profiling, and the helper-level entry points.  The batch drivers live in
// repet_drivers.cu.
#include "repet_internal.h"

#include <cmath>
#include <cstdio>
#include <cstring>

using namespace repet;

namespace {

const double kPi = 3.14159265358979323846264338327950288;
const int FFT_N_HOST = 2048;

const char* kKernelNames[REPET_NUM_KERNELS] = {"k_stft",    "k_beat",  "k_periods",   "k_model",   "k_mask_istft", "k_convert",
                                               "k_xfade",   "k_normalize", "k_simgemm", "k_topk",       "k_other10", "k_other11"};

}  // namespace

namespace repet {

int ensure_arena(repet_handle* h, size_t bytes) {
    if (bytes <= h->arena_bytes) return REPET_OK;
    CU(cudaStreamSynchronize(h->stream));
    if (h->arena) CU(cudaFree(h->arena));
    h->arena = nullptr;
    h->arena_bytes = 0;
    CU(cudaMalloc(&h->arena, bytes));
    h->arena_bytes = bytes;
    return REPET_OK;
}

int check_common(repet_handle* h, const repet_params* p, int n_channels) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!p) return fail(h, REPET_E_INVALID_ARG, "params is null");
    if (p->window_length != WIN_N || p->step_length != HOP)
        return fail(h, REPET_E_UNSUPPORTED,
                    "this build transforms window_length 2048 only (sampling frequencies in (25.6, 51.2] kHz)");
    if (n_channels < 1 || n_channels > 2)
        return fail(h, REPET_E_UNSUPPORTED, "1 or 2 channels supported");
    if (h->window_n != WIN_N) return fail(h, REPET_E_INVALID_ARG, "repet_set_window has not been called");
    return REPET_OK;
}

size_t default_ws_limit(repet_handle* h) {
    if (h->ws_limit) return (size_t)h->ws_limit;
    if (h->ws_auto) return h->ws_auto;  // cudaMemGetInfo is a slow driver call: ask once per handle
    // big chunks win (launch tails and the per-clip period kernel amortise): up to 24 GB, but never
    // more than 40 % of what is free on the device
    size_t free_b = 0, total_b = 0;
    size_t limit = (size_t)24 << 30;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
        limit = std::min(limit, (size_t)((double)(free_b + h->arena_bytes) * 0.4));
    h->ws_auto = std::max(limit, (size_t)256 << 20);
    return h->ws_auto;
}

}  // namespace repet

extern "C" {

const char* repet_version(void) { return "repet_b200 0.1 (sm_100a)"; }

int repet_create(int device, repet_handle** out) {
    if (!out) return REPET_E_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return REPET_E_CUDA;
    repet_handle* h = new repet_handle();
    h->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_compute[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0)
            h->sm_count = sms;
    }
    // transform tables in double precision, rounded once to fp32
    std::vector<float2> tw1(15 * 128), tw2(16 * 8);
    for (int k1 = 1; k1 < 16; ++k1)
        for (int m = 0; m < 128; ++m) {
            const double a = -2.0 * kPi * (double)((m * k1) % FFT_N_HOST) / (double)FFT_N_HOST;
            tw1[(k1 - 1) * 128 + m] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    for (int k2 = 0; k2 < 16; ++k2)
        for (int m2 = 0; m2 < 8; ++m2) {
            const double a = -2.0 * kPi * (double)((m2 * k2) % 128) / 128.0;
            tw2[k2 * 8 + m2] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    if (e == cudaSuccess) e = cudaMalloc(&h->tw1, tw1.size() * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&h->tw2, tw2.size() * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc(&h->window, WIN_N * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(h->tw1, tw1.data(), tw1.size() * sizeof(float2), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->tw2, tw2.data(), tw2.size() * sizeof(float2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        repet_destroy(h);
        return REPET_E_CUDA;
    }
    h->stream = h->own_stream;
    *out = h;
    return REPET_OK;
}

int repet_destroy(repet_handle* h) {
    if (!h) return REPET_OK;
    cudaSetDevice(h->device);
    if (h->own_stream) cudaStreamSynchronize(h->own_stream);
    cudaFree(h->arena);
    cudaFree(h->tw1);
    cudaFree(h->tw2);
    cudaFree(h->window);
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
        if (h->ev_compute[i]) cudaEventDestroy(h->ev_compute[i]);
        if (h->ev_d2h[i]) cudaEventDestroy(h->ev_d2h[i]);
    }
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    delete h;
    return REPET_OK;
}

const char* repet_last_error(repet_handle* h) { return h ? h->err.c_str() : "null handle"; }

int repet_set_stream(repet_handle* h, void* cuda_stream) {
    if (!h) return REPET_E_INVALID_ARG;
    h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
    return REPET_OK;
}

int repet_set_window(repet_handle* h, const double* window, int n) {
    if (!h || !window) return REPET_E_INVALID_ARG;
    if (n != WIN_N) return fail(h, REPET_E_UNSUPPORTED, "this build transforms window_length 2048 only");
    CU(cudaSetDevice(h->device));
    std::vector<float> w(n);
    for (int i = 0; i < n; ++i) w[i] = (float)window[i];
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(h->window, w.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    h->window_n = n;
    h->window_gain = window[0] + window[n / 2];
    return REPET_OK;
}

int repet_set_workspace_limit(repet_handle* h, uint64_t bytes) {
    if (!h) return REPET_E_INVALID_ARG;
    h->ws_limit = bytes;
    return REPET_OK;
}

int repet_synchronize(repet_handle* h) {
    if (!h) return REPET_E_INVALID_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    return REPET_OK;
}

uint64_t repet_launch_count(repet_handle* h) { return h ? h->launches : 0; }

int repet_set_tuning(const char* name, int value) {
    if (!name) return REPET_E_INVALID_ARG;
    const std::string key(name);
    if (key == "stft_minb") g_tuning.stft_minb = value;
    else if (key == "mask_minb") g_tuning.mask_minb = value;
    else if (key == "frames_per_cta") g_tuning.frames_per_cta = value;
    else if (key == "beat_parts") g_tuning.beat_parts = value;
    else if (key == "simgemm_tc") g_tuning.simgemm_tc = value;
    else if (key == "cert_rel_ppm") g_tuning.cert_rel_ppm = value;
    else return REPET_E_INVALID_ARG;
    return REPET_OK;
}

int repet_set_profiling(repet_handle* h, int on) {
    if (!h) return REPET_E_INVALID_ARG;
    h->profiling = on != 0;
    return REPET_OK;
}

int repet_profile_read(repet_handle* h, double* ms, uint64_t* counts, int reset) {
    if (!h) return REPET_E_INVALID_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, h->prof_events[i], h->prof_events[i + 1]));
        const int id = h->prof_ids[i / 2];
        h->prof_ms[id] += (double)t;
        h->prof_count[id] += 1;
    }
    h->prof_used = 0;
    h->prof_ids.clear();
    for (int k = 0; k < REPET_NUM_KERNELS; ++k) {
        if (ms) ms[k] = h->prof_ms[k];
        if (counts) counts[k] = h->prof_count[k];
        if (reset) {
            h->prof_ms[k] = 0.0;
            h->prof_count[k] = 0;
        }
    }
    return REPET_OK;
}

const char* repet_kernel_name(int id) { return (id >= 0 && id < REPET_NUM_KERNELS) ? kKernelNames[id] : ""; }

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
int repet_stft(repet_handle* h, const float* signal, int n_channels, int64_t n_samples, float* spectrum, float* power,
               int32_t* n_frames_out) {
    if (!h) return REPET_E_INVALID_ARG;
    if (n_channels < 1 || n_channels > 2) return fail(h, REPET_E_UNSUPPORTED, "1 or 2 channels supported");
    if (h->window_n != WIN_N) return fail(h, REPET_E_INVALID_ARG, "repet_set_window has not been called");
    if (!signal || !spectrum || n_samples < 0) return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    CU(cudaSetDevice(h->device));
    const int T = frames_of(n_samples);
    if (n_frames_out) *n_frames_out = T;
    const size_t n = (size_t)n_samples * n_channels;
    const size_t x_elems = (size_t)T * n_channels * XPITCH;
    const size_t need = align_up(n * sizeof(float)) + align_up(x_elems * sizeof(float2)) + align_up((size_t)T * PPITCH * sizeof(float));
    int rc = ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    float* in = bump.take<float>(n);
    float2* X = bump.take<float2>(x_elems);
    float* P = bump.take<float>((size_t)T * PPITCH);
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(in, signal, n * sizeof(float), cudaMemcpyHostToDevice, st));
    Geom g{1, 1, 0, 0, n_samples, 0, (int)n_samples, T};
    launch_stft(st, in, g, n_channels, h->window, tables(h), X, power ? P : nullptr, P_POWER, 8);
    h->launches += 1;
    CU(cudaMemcpyAsync(spectrum, X, x_elems * sizeof(float2), cudaMemcpyDeviceToHost, st));
    if (power)
        CU(cudaMemcpy2DAsync(power, NBIN * sizeof(float), P, PPITCH * sizeof(float), NBIN * sizeof(float), T,
                             cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

int repet_istft(repet_handle* h, const float* spectrum, int n_channels, int n_frames, double cola_gain, float* signal) {
    if (!h) return REPET_E_INVALID_ARG;
    if (n_channels < 1 || n_channels > 2) return fail(h, REPET_E_UNSUPPORTED, "1 or 2 channels supported");
    if (!spectrum || !signal || n_frames < 2) return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    CU(cudaSetDevice(h->device));
    const long long S = (long long)(n_frames - 1) * HOP;
    const size_t x_elems = (size_t)n_frames * n_channels * XPITCH;
    const size_t need = align_up(x_elems * sizeof(float2)) + align_up((size_t)S * n_channels * sizeof(float));
    int rc = ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    float2* X = bump.take<float2>(x_elems);
    float* out = bump.take<float>((size_t)S * n_channels);
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(X, spectrum, x_elems * sizeof(float2), cudaMemcpyHostToDevice, st));
    Geom g{1, 1, 0, 0, S, 0, (int)S, n_frames};
    launch_istft(st, X, g, n_channels, (float)(1.0 / ((double)WIN_N * cola_gain)), tables(h), out, 8);
    h->launches += 1;
    CU(cudaMemcpyAsync(signal, out, (size_t)S * n_channels * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

static int beat_common(repet_handle* h, const float* spectrogram, int n_frames, int n_rows, int lag_lo, int lag_hi,
                       double* beat, int32_t* period) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!spectrogram || n_frames < 1 || n_rows < 1 || n_rows > NBIN)
        return fail(h, REPET_E_INVALID_ARG, "bad buffer or size (n_rows <= 1025)");
    const int max_lag = beat ? n_frames - 1 : lag_hi - 1;
    if (n_frames + max_lag > BEAT_L)
        return fail(h, REPET_E_UNSUPPORTED, "n_frames + max lag exceeds the 2048-point beat transform");
    CU(cudaSetDevice(h->device));
    const int n_parts = 17, f_per_part = 64;
    const size_t need = align_up((size_t)n_frames * PPITCH * sizeof(float)) + align_up((size_t)n_parts * BEAT_L * sizeof(float)) +
                        align_up((size_t)BEAT_L * sizeof(double)) + 512;
    int rc = ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    float* P = bump.take<float>((size_t)n_frames * PPITCH);
    float* psd = bump.take<float>((size_t)n_parts * BEAT_L);
    double* b = bump.take<double>(BEAT_L);
    int32_t* per = bump.take<int32_t>(1);
    cudaStream_t st = h->stream;
    CU(cudaMemsetAsync(P, 0, (size_t)n_frames * PPITCH * sizeof(float), st));
    CU(cudaMemcpy2DAsync(P, PPITCH * sizeof(float), spectrogram, n_rows * sizeof(float), n_rows * sizeof(float),
                         n_frames, cudaMemcpyHostToDevice, st));
    launch_beat(st, P, 1, n_frames, 0, n_frames, 0, 1, tables(h), psd, n_parts, f_per_part);
    launch_periods(st, psd, nullptr, 1, n_parts, n_frames, (double)n_rows, lag_lo, lag_hi, 0, beat ? n_frames : 0,
                   beat ? b : nullptr, BEAT_L, period ? per : nullptr, nullptr, nullptr);
    h->launches += 2;
    if (beat) CU(cudaMemcpyAsync(beat, b, (size_t)n_frames * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (period) CU(cudaMemcpyAsync(period, per, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

int repet_beatspectrum(repet_handle* h, const float* spectrogram, int n_frames, int n_rows, double* beat) {
    if (!beat) return fail(h, REPET_E_INVALID_ARG, "beat is null");
    return beat_common(h, spectrogram, n_frames, n_rows, 0, 0, beat, nullptr);
}

int repet_period(repet_handle* h, const float* spectrogram, int n_frames, int n_rows, int period_lo, int period_hi,
                 int32_t* period) {
    if (!period) return fail(h, REPET_E_INVALID_ARG, "period is null");
    const int lag_hi = std::min(period_hi, n_frames / 3);
    if (period_lo < 0 || lag_hi <= period_lo)
        return fail(h, REPET_E_TOO_SHORT, "attempt to get argmax of an empty sequence");
    return beat_common(h, spectrogram, n_frames, n_rows, period_lo, lag_hi, nullptr, period);
}

// magnitudes [n_frames][1025] as purely real spectra; bin 0 packs (DC, Nyquist)
static void pack_magnitudes(const float* magnitude, int T, std::vector<float2>& host) {
    host.resize((size_t)T * XPITCH);
    for (int j = 0; j < T; ++j) {
        const float* row = magnitude + (size_t)j * NBIN;
        host[(size_t)j * XPITCH] = make_float2(row[0], row[XPITCH]);
        for (int k = 1; k < XPITCH; ++k) host[(size_t)j * XPITCH + k] = make_float2(row[k], 0.f);
    }
}

int repet_mask(repet_handle* h, const float* magnitude, int n_frames, int period, float* mask) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!magnitude || !mask || n_frames < 1 || period < 1) return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    CU(cudaSetDevice(h->device));
    const int T = n_frames;
    const size_t x_elems = (size_t)T * XPITCH;
    const size_t need = align_up(x_elems * sizeof(float2)) + align_up((size_t)period * PPITCH * sizeof(float)) +
                        align_up((size_t)T * PPITCH * sizeof(float)) + 512;
    int rc = ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    float2* X = bump.take<float2>(x_elems);
    float* model = bump.take<float>((size_t)period * PPITCH);
    float* M = bump.take<float>((size_t)T * PPITCH);
    int32_t* per = bump.take<int32_t>(1);
    std::vector<float2> host;
    pack_magnitudes(magnitude, T, host);
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(X, host.data(), x_elems * sizeof(float2), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(per, &period, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    launch_model(st, X, 1, T, 1, per, period, model);
    launch_mask_only(st, X, 1, T, 1, per, period, model, M);
    h->launches += 2;
    CU(cudaMemcpy2DAsync(mask, NBIN * sizeof(float), M, PPITCH * sizeof(float), NBIN * sizeof(float), T,
                         cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

int repet_adaptivemask(repet_handle* h, const float* magnitude, int n_frames, const int32_t* periods, int filter_order,
                       float* mask) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!magnitude || !mask || !periods || n_frames < 1 || filter_order < 1)
        return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    CU(cudaSetDevice(h->device));
    const int T = n_frames;
    const size_t x_elems = (size_t)T * XPITCH;
    const size_t need = align_up(x_elems * sizeof(float2)) + 2 * align_up((size_t)T * PPITCH * sizeof(float)) +
                        align_up((size_t)T * sizeof(int32_t)) + 512;
    int rc = ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    float2* X = bump.take<float2>(x_elems);
    float* model = bump.take<float>((size_t)T * PPITCH);
    float* M = bump.take<float>((size_t)T * PPITCH);
    int32_t* per = bump.take<int32_t>(T);
    std::vector<float2> host;
    pack_magnitudes(magnitude, T, host);
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(X, host.data(), x_elems * sizeof(float2), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(per, periods, (size_t)T * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    launch_adaptive_model(st, X, 1, T, 1, per, filter_order, model);
    launch_mask_only(st, X, 1, T, 1, nullptr, T, model, M);
    h->launches += 2;
    CU(cudaMemcpy2DAsync(mask, NBIN * sizeof(float), M, PPITCH * sizeof(float), NBIN * sizeof(float), T,
                         cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

int repet_beatspectrogram(repet_handle* h, const float* spectrogram, int n_frames, int n_rows, int segment_length,
                          int segment_step, double* beat, int32_t* n_segments_out) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!spectrogram || !beat || n_frames < 1 || n_rows < 1 || n_rows > NBIN || segment_length < 1 || segment_step < 1)
        return fail(h, REPET_E_INVALID_ARG, "bad buffer or size (n_rows <= 1025)");
    if (2 * segment_length - 1 > BEAT_L)
        return fail(h, REPET_E_UNSUPPORTED, "segment_length exceeds the 2048-point beat transform");
    CU(cudaSetDevice(h->device));
    const int n_seg = (n_frames + segment_step - 1) / segment_step;
    if (n_segments_out) *n_segments_out = n_seg;
    const int n_parts = 9, f_per_part = 120;
    const size_t need = align_up((size_t)n_frames * PPITCH * sizeof(float)) +
                        align_up((size_t)n_seg * n_parts * BEAT_L * sizeof(float)) +
                        align_up((size_t)n_seg * segment_length * sizeof(double)) + 512;
    int rc = ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    float* P = bump.take<float>((size_t)n_frames * PPITCH);
    float* psd = bump.take<float>((size_t)n_seg * n_parts * BEAT_L);
    double* b = bump.take<double>((size_t)n_seg * segment_length);
    cudaStream_t st = h->stream;
    CU(cudaMemsetAsync(P, 0, (size_t)n_frames * PPITCH * sizeof(float), st));
    CU(cudaMemcpy2DAsync(P, PPITCH * sizeof(float), spectrogram, n_rows * sizeof(float), n_rows * sizeof(float), n_frames,
                         cudaMemcpyHostToDevice, st));
    const int left = segment_length / 2;  // ceil((L-1)/2), repet.py:1182
    launch_beat(st, P, 1, n_frames, -left, segment_length, segment_step, n_seg, tables(h), psd, n_parts, f_per_part);
    launch_periods(st, psd, nullptr, n_seg, n_parts, segment_length, (double)n_rows, 0, 0, 0, segment_length, b, segment_length,
                   nullptr, nullptr, nullptr);
    h->launches += 2;
    CU(cudaMemcpyAsync(beat, b, (size_t)n_seg * segment_length * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

int repet_selfsimilarity(repet_handle* h, const float* magnitude, int n_frames, int n_rows, float* similarity) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!magnitude || !similarity || n_frames < 1 || n_rows < 1 || n_rows > NBIN)
        return fail(h, REPET_E_INVALID_ARG, "bad buffer or size (n_rows <= 1025)");
    CU(cudaSetDevice(h->device));
    const int T = n_frames;
    const size_t need = align_up((size_t)T * PPITCH * sizeof(float)) + 2 * align_up((size_t)T * KPAD * sizeof(float)) +
                        align_up((size_t)T * T * sizeof(float)) + 512;
    int rc = ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    float* V = bump.take<float>((size_t)T * PPITCH);
    float* An32 = bump.take<float>((size_t)T * KPAD);
    float* An32lo = bump.take<float>((size_t)T * KPAD);
    float* S = bump.take<float>((size_t)T * T);
    cudaStream_t st = h->stream;
    CU(cudaMemsetAsync(V, 0, (size_t)T * PPITCH * sizeof(float), st));
    CU(cudaMemcpy2DAsync(V, PPITCH * sizeof(float), magnitude, n_rows * sizeof(float), n_rows * sizeof(float), T,
                         cudaMemcpyHostToDevice, st));
    const bool split = g_tuning.simgemm_tc >= 2;
    launch_normalize(st, V, T, nullptr, An32, split ? An32lo : nullptr, g_tuning.simgemm_tc ? 1 : 0);
    if (g_tuning.simgemm_tc) {
        if (launch_selfsim_tc(st, An32, split ? An32lo : nullptr, 1, T, S, h->sm_count))
            return fail(h, REPET_E_CUDA, "tensor-map encode failed");
    } else {
        launch_selfsim_simt(st, An32, 1, T, S);
    }
    h->launches += 2;
    CU(cudaMemcpyAsync(similarity, S, (size_t)T * T * sizeof(float), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

int repet_periods(repet_handle* h, const double* beat, int n_lags, int n_columns, int period_lo, int period_hi,
                  int32_t* periods) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!beat || !periods || n_lags < 1 || n_columns < 1) return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    const int lag_hi = std::min(period_hi, n_lags / 3);  // repet.py:1265-1267
    if (period_lo < 0 || lag_hi <= period_lo)
        return fail(h, REPET_E_TOO_SHORT, "attempt to get argmax of an empty sequence");
    CU(cudaSetDevice(h->device));
    const size_t n = (size_t)n_lags * n_columns;
    int rc = ensure_arena(h, align_up(n * sizeof(double)) + align_up((size_t)n_columns * sizeof(int32_t)));
    if (rc) return rc;
    Bump bump(h->arena);
    double* b = bump.take<double>(n);
    int32_t* per = bump.take<int32_t>(n_columns);
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(b, beat, n * sizeof(double), cudaMemcpyHostToDevice, st));
    launch_argmax_columns(st, b, n_lags, n_columns, period_lo, lag_hi, per);
    h->launches += 1;
    CU(cudaMemcpyAsync(periods, per, (size_t)n_columns * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

// upload a [n][n_rows] fp32 matrix into rows of PPITCH floats (zero padded)
static int upload_rows(repet_handle* h, const float* host, int n, int n_rows, float* dev) {
    CU(cudaMemsetAsync(dev, 0, (size_t)n * PPITCH * sizeof(float), h->stream));
    CU(cudaMemcpy2DAsync(dev, PPITCH * sizeof(float), host, n_rows * sizeof(float), n_rows * sizeof(float), n,
                         cudaMemcpyHostToDevice, h->stream));
    return REPET_OK;
}

int repet_similarity(repet_handle* h, const float* magnitude1, int n_frames1, const float* magnitude2, int n_frames2,
                     int n_rows, double* similarity) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!magnitude1 || !magnitude2 || !similarity || n_frames1 < 1 || n_frames2 < 1 || n_rows < 1 || n_rows > NBIN)
        return fail(h, REPET_E_INVALID_ARG, "bad buffer or size (n_rows <= 1025)");
    CU(cudaSetDevice(h->device));
    const size_t n1 = n_frames1, n2 = n_frames2;
    const size_t need = align_up((n1 + n2) * PPITCH * sizeof(float)) + align_up((n1 + n2) * APITCH64 * sizeof(double)) +
                        align_up(n1 * n2 * sizeof(double)) + 1024;
    int rc = ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    float* V = bump.take<float>((n1 + n2) * PPITCH);
    double* An = bump.take<double>((n1 + n2) * APITCH64);
    double* out = bump.take<double>(n1 * n2);
    cudaStream_t st = h->stream;
    if ((rc = upload_rows(h, magnitude1, n_frames1, n_rows, V))) return rc;
    if ((rc = upload_rows(h, magnitude2, n_frames2, n_rows, V + n1 * PPITCH))) return rc;
    launch_normalize(st, V, n_frames1 + n_frames2, An, nullptr, nullptr, 0);
    launch_cosine64(st, An, n_frames1, An + n1 * APITCH64, n_frames2, out);
    h->launches += 2;
    CU(cudaMemcpyAsync(similarity, out, n1 * n2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

int repet_localmaxima(repet_handle* h, const double* data, int n, int n_columns, double minimum_value,
                      int minimum_distance, int number_values, int32_t* indices, int32_t* counts, double* values) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!data || !indices || !counts || n < 1 || n_columns < 1 || number_values < 1 || minimum_distance < 0)
        return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    CU(cudaSetDevice(h->device));
    const size_t total = (size_t)n * n_columns, lists = (size_t)n_columns * number_values;
    const size_t need = align_up(total * sizeof(double)) + align_up(lists * sizeof(int32_t)) +
                        align_up((size_t)n_columns * sizeof(int32_t)) + align_up(lists * sizeof(double)) + 1024;
    int rc = ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    double* d = bump.take<double>(total);
    int32_t* idx = bump.take<int32_t>(lists);
    int32_t* cnt = bump.take<int32_t>(n_columns);
    double* val = bump.take<double>(lists);
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(d, data, total * sizeof(double), cudaMemcpyHostToDevice, st));
    if (launch_localmaxima64(st, d, n, n_columns, minimum_value, minimum_distance, number_values, idx, cnt, val))
        return fail(h, REPET_E_UNSUPPORTED, "vector too long for the shared-memory local-maximum scan");
    h->launches += 1;
    CU(cudaMemcpyAsync(indices, idx, lists * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(counts, cnt, (size_t)n_columns * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (values) CU(cudaMemcpyAsync(values, val, lists * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

int repet_simmask(repet_handle* h, const float* magnitude, int n_frames, const int32_t* indices, const int32_t* counts,
                  int number, float* mask) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!magnitude || !indices || !counts || !mask || n_frames < 1 || number < 1)
        return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    CU(cudaSetDevice(h->device));
    const int T = n_frames;
    const size_t x_elems = (size_t)T * XPITCH;
    const size_t need = align_up(x_elems * sizeof(float2)) + 3 * align_up((size_t)T * PPITCH * sizeof(float)) +
                        align_up((size_t)T * number * sizeof(int32_t)) + align_up((size_t)T * sizeof(int32_t)) + 1024;
    int rc = ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    float2* X = bump.take<float2>(x_elems);
    float* model = bump.take<float>((size_t)T * PPITCH);
    float* M = bump.take<float>((size_t)T * PPITCH);
    float* Vsq = bump.take<float>((size_t)T * PPITCH);
    int32_t* idx = bump.take<int32_t>((size_t)T * number);
    int32_t* cnt = bump.take<int32_t>(T);
    std::vector<float2> host;
    pack_magnitudes(magnitude, T, host);
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(X, host.data(), x_elems * sizeof(float2), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(idx, indices, (size_t)T * number * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(cnt, counts, (size_t)T * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    launch_sqmag(st, X, T, Vsq);
    if (launch_simmodel(st, X, Vsq, 1, T, 1, idx, cnt, number, 0, model))
        return fail(h, REPET_E_UNSUPPORTED, "lists too long for the shared-memory median");
    launch_mask_only(st, X, 1, T, 1, nullptr, T, model, M);
    h->launches += 2;
    CU(cudaMemcpy2DAsync(mask, NBIN * sizeof(float), M, PPITCH * sizeof(float), NBIN * sizeof(float), T,
                         cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(cudaGetLastError());
    return REPET_OK;
}

int repet_acorr(repet_handle* h, const float* data, int n_rows, int n_columns, double* autocorrelation) {
    if (!h) return REPET_E_INVALID_ARG;
    if (!data || !autocorrelation || n_rows < 1 || n_columns < 1) return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    if (2 * n_rows - 1 > BEAT_L) return fail(h, REPET_E_UNSUPPORTED, "more rows than the 2048-point transform holds");
    CU(cudaSetDevice(h->device));
    // every column becomes one beat item whose only non-zero frequency row is that column
    const int chunk = std::max(1, std::min(n_columns, (int)(((size_t)512 << 20) / ((size_t)n_rows * PPITCH * sizeof(float)))));
    const size_t need = align_up((size_t)chunk * n_rows * PPITCH * sizeof(float)) + align_up((size_t)chunk * BEAT_L * sizeof(float)) +
                        align_up((size_t)chunk * n_rows * sizeof(double)) + 1024;
    int rc = ensure_arena(h, need);
    if (rc) return rc;
    Bump bump(h->arena);
    float* P = bump.take<float>((size_t)chunk * n_rows * PPITCH);
    float* psd = bump.take<float>((size_t)chunk * BEAT_L);
    double* b = bump.take<double>((size_t)chunk * n_rows);
    cudaStream_t st = h->stream;
    std::vector<double> host((size_t)chunk * n_rows);
    for (int c0 = 0; c0 < n_columns; c0 += chunk) {
        const int g = std::min(chunk, n_columns - c0);
        CU(cudaMemsetAsync(P, 0, (size_t)g * n_rows * PPITCH * sizeof(float), st));
        for (int c = 0; c < g; ++c)  // column c0+c -> item c, frequency row 0
            CU(cudaMemcpy2DAsync(P + (size_t)c * n_rows * PPITCH, PPITCH * sizeof(float), data + c0 + c,
                                 n_columns * sizeof(float), sizeof(float), n_rows, cudaMemcpyHostToDevice, st));
        launch_beat(st, P, g, n_rows, 0, n_rows, 0, 1, tables(h), psd, 1, 8);
        launch_periods(st, psd, nullptr, g, 1, n_rows, 1.0, 0, 0, 0, n_rows, b, n_rows, nullptr, nullptr, nullptr);
        h->launches += 2;
        CU(cudaMemcpyAsync(host.data(), b, (size_t)g * n_rows * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (int c = 0; c < g; ++c)
            for (int l = 0; l < n_rows; ++l) autocorrelation[(size_t)l * n_columns + c0 + c] = host[(size_t)c * n_rows + l];
    }
    CU(cudaGetLastError());
    return REPET_OK;
}

}  // extern "C"
