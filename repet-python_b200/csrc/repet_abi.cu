// C ABI of librepet_b200.so (declared in include/repet_b200.h): handle, workspace arena,
// transform tables, chunked batch drivers and the helper-level entry points.
#include "../../include/repet_b200.h"
#include "repet_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace repet;

struct repet_handle {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t h2d_stream = nullptr;
    cudaStream_t d2h_stream = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr};
    cudaEvent_t ev_compute[2] = {nullptr, nullptr};
    cudaEvent_t ev_d2h[2] = {nullptr, nullptr};
    std::string err;
    float2* tw1 = nullptr;
    float2* tw2 = nullptr;
    float* window = nullptr;
    int window_n = 0;
    double window_gain = 0.0;  // sum(window[0:N:H])
    unsigned char* arena = nullptr;
    size_t arena_bytes = 0;
    uint64_t ws_limit = 0;
    uint64_t launches = 0;
    int sm_count = 148;
    // optional per-kernel timing (bench.py's roofline): one event pair per launch
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;
    std::vector<int> prof_ids;
    size_t prof_used = 0;
    double prof_ms[REPET_NUM_KERNELS] = {0};
    uint64_t prof_count[REPET_NUM_KERNELS] = {0};
};

namespace {

const double kPi = 3.14159265358979323846264338327950288;
const int FFT_N_HOST = 2048;
const int MAX_ITEMS_PER_LAUNCH = 16384;  // grid.y / grid.z stay far below 65535

int fail(repet_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(h, e__ == cudaErrorMemoryAllocation ? REPET_E_OOM : REPET_E_CUDA,          \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                      \
    } while (0)

size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

const char* kKernelNames[REPET_NUM_KERNELS] = {"k_stft", "k_beat", "k_periods", "k_model", "k_mask_istft",
                                               "k_convert", "k_other6", "k_other7"};

// Brackets one launch with an event pair when profiling is on; always counts the launch.
struct Timed {
    repet_handle* h;
    Timed(repet_handle* handle, int id) : h(handle) {
        h->launches += 1;
        if (!h->profiling) return;
        if (h->prof_used + 2 > h->prof_events.size()) {
            for (int i = 0; i < 2; ++i) {
                cudaEvent_t e = nullptr;
                cudaEventCreate(&e);
                h->prof_events.push_back(e);
            }
        }
        h->prof_ids.push_back(id);
        cudaEventRecord(h->prof_events[h->prof_used], h->stream);
    }
    ~Timed() {
        if (!h->profiling) return;
        cudaEventRecord(h->prof_events[h->prof_used + 1], h->stream);
        h->prof_used += 2;
    }
};

struct Bump {
    unsigned char* base;
    size_t off = 0;
    explicit Bump(unsigned char* b) : base(b) {}
    template <typename T>
    T* take(size_t count) {
        T* p = reinterpret_cast<T*>(base + off);
        off = align_up(off + count * sizeof(T));
        return p;
    }
};

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

int frames_of(int64_t n_samples) {  // repet.py:1018-1028 with N = 2H
    return (int)((n_samples + HOP - 1) / HOP) + 1;
}

FftTables tables(repet_handle* h) { return FftTables{h->tw1, h->tw2}; }

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

// per-item workspace of the `original` pipeline
struct OriginalPlan {
    int T, nch, pmax, lag_hi, n_parts, f_per_part;
    size_t bytes_per_item;
};

int plan_original(repet_handle* h, const repet_params* p, int n_channels, int64_t n_samples, int chunk_hint,
                  OriginalPlan* plan) {
    const int T = frames_of(n_samples);
    const int lag_hi = std::min(p->period_hi, T / 3);  // repet.py:1265-1267 (quirk Q2)
    if (p->period_lo < 0 || lag_hi <= p->period_lo)
        return fail(h, REPET_E_TOO_SHORT, "attempt to get argmax of an empty sequence (clip too short for the period range)");
    if (T + lag_hi - 1 > BEAT_L)
        return fail(h, REPET_E_UNSUPPORTED, "clip longer than the single-block beat transform (T + max lag > 2048 frames)");
    plan->T = T;
    plan->nch = n_channels;
    plan->lag_hi = lag_hi;
    plan->pmax = lag_hi;  // period = lag + 1 <= lag_hi
    // measured on B200 (profiles/r1_sweeps.md): ~64 partitions of the 1025 rows per clip is the sweet spot
    int want_parts = std::max(64, std::min(129, (h->sm_count * 6 + chunk_hint - 1) / std::max(1, chunk_hint)));
    if (g_tuning.beat_parts > 0) want_parts = std::min(129, g_tuning.beat_parts);
    int f_per_part = ((NBIN + want_parts - 1) / want_parts + 7) / 8 * 8;
    plan->f_per_part = f_per_part;
    plan->n_parts = (NBIN + f_per_part - 1) / f_per_part;
    size_t b = 0;
    b += align_up((size_t)T * n_channels * XPITCH * sizeof(float2));
    b += align_up((size_t)T * PPITCH * sizeof(float));
    b += align_up((size_t)plan->n_parts * BEAT_L * sizeof(float));
    b += align_up((size_t)n_channels * plan->pmax * PPITCH * sizeof(float));
    b += 512;
    plan->bytes_per_item = b;
    return REPET_OK;
}

size_t default_ws_limit(repet_handle* h) {
    if (h->ws_limit) return (size_t)h->ws_limit;
    // big chunks win (launch tails and the per-clip period kernel amortise): up to 24 GB, but never
    // more than 40 % of what is free on the device
    size_t free_b = 0, total_b = 0;
    size_t limit = (size_t)24 << 30;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
        limit = std::min(limit, (size_t)((double)(free_b + h->arena_bytes) * 0.4));
    return std::max(limit, (size_t)256 << 20);
}

int pick_frames_per_cta(repet_handle* h, long long total_frames) {
    if (g_tuning.frames_per_cta > 0) return g_tuning.frames_per_cta;
    long long k = total_frames / ((long long)h->sm_count * 8);
    return (int)std::max(4LL, std::min(16LL, k));
}

// `original` on device-resident clips, in workspace-sized chunks.  `ws` is arena space after
// whatever the caller reserved.
int original_dev_chunked(repet_handle* h, const float* audio, int n_clips, int n_channels, int64_t n_samples,
                         const repet_params* p, float* background, int32_t* periods_dev, unsigned char* ws,
                         size_t ws_bytes, const OriginalPlan& plan) {
    const int G = (int)std::min<size_t>(std::min<size_t>((size_t)n_clips, MAX_ITEMS_PER_LAUNCH),
                                        std::max<size_t>(1, ws_bytes / plan.bytes_per_item));
    const float scale = (float)(1.0 / ((double)WIN_N * p->cola_gain));
    cudaStream_t st = h->stream;
    for (int first = 0; first < n_clips; first += G) {
        const int g_items = std::min(G, n_clips - first);
        Bump bump(ws);
        float2* X = bump.take<float2>((size_t)g_items * plan.T * n_channels * XPITCH);
        float* P = bump.take<float>((size_t)g_items * plan.T * PPITCH);
        float* psd = bump.take<float>((size_t)g_items * plan.n_parts * BEAT_L);
        float* model = bump.take<float>((size_t)g_items * n_channels * plan.pmax * PPITCH);
        Geom g;
        g.n_items = g_items;
        g.seg_per_clip = 1;
        g.clip_stride = (long long)n_channels * n_samples;
        g.seg_stride = 0;
        g.chan_stride = n_samples;
        g.first_offset = (long long)first * g.clip_stride;
        g.S = (int)n_samples;
        g.T = plan.T;
        const int K = pick_frames_per_cta(h, (long long)g_items * plan.T);
        {
            Timed timed(h, REPET_K_STFT);
            launch_stft(st, audio, g, n_channels, h->window, tables(h), X, P, P_POWER, K);
        }
        {
            Timed timed(h, REPET_K_BEAT);
            launch_beat(st, P, g_items, plan.T, 0, plan.T, 0, 1, tables(h), psd, plan.n_parts, plan.f_per_part);
        }
        {
            Timed timed(h, REPET_K_PERIODS);
            launch_periods(st, psd, g_items, plan.n_parts, plan.T, (double)NBIN, p->period_lo, plan.lag_hi, 0, 0,
                           nullptr, 0, periods_dev + first, nullptr);
        }
        {
            Timed timed(h, REPET_K_MODEL);
            launch_model(st, X, g_items, plan.T, n_channels, periods_dev + first, plan.pmax, model);
        }
        {
            Timed timed(h, REPET_K_MASK_ISTFT);
            launch_mask_istft(st, X, g, n_channels, periods_dev + first, plan.pmax, model, p->cutoff_bins, scale,
                              tables(h), background, K);
        }
    }
    CU(cudaGetLastError());
    return REPET_OK;
}

}  // namespace

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
// drivers
// ---------------------------------------------------------------------------------------------
int repet_original_batch_dev(repet_handle* h, const float* audio, int n_clips, int n_channels, int64_t n_samples,
                             const repet_params* p, float* background, int32_t* periods_dev, int32_t* periods_host) {
    int rc = check_common(h, p, n_channels);
    if (rc) return rc;
    if (!audio || !background || n_clips < 0 || n_samples < 0) return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    if (n_clips == 0) return REPET_OK;
    CU(cudaSetDevice(h->device));
    OriginalPlan plan;
    const size_t limit = default_ws_limit(h);
    // size the chunk first with a neutral hint, then re-plan the beat partition for it
    if ((rc = plan_original(h, p, n_channels, n_samples, 64, &plan))) return rc;
    int G = (int)std::min<size_t>((size_t)n_clips, std::max<size_t>(1, limit / plan.bytes_per_item));
    if ((rc = plan_original(h, p, n_channels, n_samples, G, &plan))) return rc;
    G = (int)std::min<size_t>((size_t)n_clips, std::max<size_t>(1, limit / plan.bytes_per_item));
    const size_t periods_bytes = periods_dev ? 0 : align_up((size_t)n_clips * sizeof(int32_t));
    const size_t need = periods_bytes + (size_t)G * plan.bytes_per_item;
    if ((rc = ensure_arena(h, need))) return rc;
    int32_t* per = periods_dev ? periods_dev : reinterpret_cast<int32_t*>(h->arena);
    rc = original_dev_chunked(h, audio, n_clips, n_channels, n_samples, p, background, per, h->arena + periods_bytes,
                              (size_t)G * plan.bytes_per_item, plan);
    if (rc) return rc;
    if (periods_host) {
        CU(cudaMemcpyAsync(periods_host, per, (size_t)n_clips * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return REPET_OK;
}

int repet_original_batch(repet_handle* h, const float* audio, int n_clips, int n_channels, int64_t n_samples,
                         const repet_params* p, float* background, int32_t* periods_host) {
    int rc = check_common(h, p, n_channels);
    if (rc) return rc;
    if (!audio || !background || n_clips < 0 || n_samples < 0) return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    if (n_clips == 0) return REPET_OK;
    CU(cudaSetDevice(h->device));
    const size_t clip_elems = (size_t)n_channels * (size_t)n_samples;
    const size_t clip_bytes = clip_elems * sizeof(float);
    // copy granularity: about 256 MB per slot, two slots in flight
    int Gc = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_clips, ((size_t)256 << 20) / std::max<size_t>(1, clip_bytes)));
    OriginalPlan plan;
    if ((rc = plan_original(h, p, n_channels, n_samples, Gc, &plan))) return rc;
    const size_t limit = default_ws_limit(h);
    const int Gw = (int)std::min<size_t>((size_t)Gc, std::max<size_t>(1, limit / plan.bytes_per_item));
    const size_t slot_bytes = align_up((size_t)Gc * clip_bytes);
    const size_t periods_bytes = align_up((size_t)n_clips * sizeof(int32_t));
    const size_t need = periods_bytes + 4 * slot_bytes + (size_t)Gw * plan.bytes_per_item;
    if ((rc = ensure_arena(h, need))) return rc;
    int32_t* per = reinterpret_cast<int32_t*>(h->arena);
    float* in_slot[2] = {reinterpret_cast<float*>(h->arena + periods_bytes),
                         reinterpret_cast<float*>(h->arena + periods_bytes + slot_bytes)};
    float* out_slot[2] = {reinterpret_cast<float*>(h->arena + periods_bytes + 2 * slot_bytes),
                          reinterpret_cast<float*>(h->arena + periods_bytes + 3 * slot_bytes)};
    unsigned char* ws = h->arena + periods_bytes + 4 * slot_bytes;
    // the caller's stream must see the arena idle before the copy streams touch it
    CU(cudaStreamSynchronize(h->stream));
    int n_chunks = 0;
    for (int first = 0; first < n_clips; first += Gc, ++n_chunks) {
        const int s = n_chunks & 1;
        const int g = std::min(Gc, n_clips - first);
        if (n_chunks >= 2) CU(cudaStreamWaitEvent(h->h2d_stream, h->ev_compute[s], 0));  // slot's input consumed
        CU(cudaMemcpyAsync(in_slot[s], audio + (size_t)first * clip_elems, (size_t)g * clip_bytes,
                           cudaMemcpyHostToDevice, h->h2d_stream));
        CU(cudaEventRecord(h->ev_h2d[s], h->h2d_stream));
        CU(cudaStreamWaitEvent(h->stream, h->ev_h2d[s], 0));
        if (n_chunks >= 2) CU(cudaStreamWaitEvent(h->stream, h->ev_d2h[s], 0));  // slot's output drained
        rc = original_dev_chunked(h, in_slot[s], g, n_channels, n_samples, p, out_slot[s], per + first, ws,
                                  (size_t)Gw * plan.bytes_per_item, plan);
        if (rc) return rc;
        CU(cudaEventRecord(h->ev_compute[s], h->stream));
        CU(cudaStreamWaitEvent(h->d2h_stream, h->ev_compute[s], 0));
        CU(cudaMemcpyAsync(background + (size_t)first * clip_elems, out_slot[s], (size_t)g * clip_bytes,
                           cudaMemcpyDeviceToHost, h->d2h_stream));
        CU(cudaEventRecord(h->ev_d2h[s], h->d2h_stream));
    }
    if (periods_host)
        CU(cudaMemcpyAsync(periods_host, per, (size_t)n_clips * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaStreamSynchronize(h->d2h_stream));
    return REPET_OK;
}

int repet_original_f64(repet_handle* h, const double* audio, int64_t n_samples, int n_channels,
                       const repet_params* p, double* background, int32_t* period_host) {
    int rc = check_common(h, p, n_channels);
    if (rc) return rc;
    if (!audio || !background || n_samples < 0) return fail(h, REPET_E_INVALID_ARG, "bad buffer or size");
    CU(cudaSetDevice(h->device));
    OriginalPlan plan;
    if ((rc = plan_original(h, p, n_channels, n_samples, 1, &plan))) return rc;
    const size_t n = (size_t)n_samples * n_channels;
    const size_t f64_bytes = align_up(n * sizeof(double));
    const size_t f32_bytes = align_up(n * sizeof(float));
    const size_t need = 256 + f64_bytes + 2 * f32_bytes + plan.bytes_per_item;
    if ((rc = ensure_arena(h, need))) return rc;
    Bump bump(h->arena);
    int32_t* per = bump.take<int32_t>(1);
    double* d64 = bump.take<double>(n);
    float* in32 = bump.take<float>(n);
    float* out32 = bump.take<float>(n);
    unsigned char* ws = h->arena + bump.off;
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(d64, audio, n * sizeof(double), cudaMemcpyHostToDevice, st));
    {
        Timed timed(h, REPET_K_CONVERT);
        launch_f64_interleaved_to_planar(st, d64, n_samples, n_channels, in32);
    }
    rc = original_dev_chunked(h, in32, 1, n_channels, n_samples, p, out32, per, ws, plan.bytes_per_item, plan);
    if (rc) return rc;
    {
        Timed timed(h, REPET_K_CONVERT);
        launch_planar_to_f64_interleaved(st, out32, n_samples, n_channels, d64);
    }
    CU(cudaMemcpyAsync(background, d64, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (period_host) CU(cudaMemcpyAsync(period_host, per, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return REPET_OK;
}

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
    launch_periods(st, psd, 1, n_parts, n_frames, (double)n_rows, lag_lo, lag_hi, 0, beat ? n_frames : 0,
                   beat ? b : nullptr, BEAT_L, period ? per : nullptr, nullptr);
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
    // magnitudes as purely real spectra; bin 0 packs (DC, Nyquist)
    std::vector<float2> host(x_elems);
    for (int j = 0; j < T; ++j) {
        const float* row = magnitude + (size_t)j * NBIN;
        host[(size_t)j * XPITCH] = make_float2(row[0], row[XPITCH]);
        for (int k = 1; k < XPITCH; ++k) host[(size_t)j * XPITCH + k] = make_float2(row[k], 0.f);
    }
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

}  // extern "C"
