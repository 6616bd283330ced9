// Internals shared by repet_abi.cu (handle, helpers) and repet_drivers.cu (batch drivers).
#pragma once
#include "../../include/repet_b200.h"
#include "repet_kernels.cuh"

#include <algorithm>
#include <string>
#include <vector>

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
    size_t ws_auto = 0;  // cached automatic workspace limit
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

namespace repet {

inline int fail(repet_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return repet::fail(h, e__ == cudaErrorMemoryAllocation ? REPET_E_OOM : REPET_E_CUDA,   \
                               std::string(#call) + ": " + cudaGetErrorString(e__));               \
    } while (0)

inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

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

constexpr int MAX_ITEMS_PER_LAUNCH = 16384;  // grid.y / grid.z stay far below 65535

int ensure_arena(repet_handle* h, size_t bytes);
size_t default_ws_limit(repet_handle* h);
inline int frames_of(int64_t n_samples) {  // repet.py:1018-1028 with N = 2H
    return (int)((n_samples + HOP - 1) / HOP) + 1;
}
inline FftTables tables(repet_handle* h) { return FftTables{h->tw1, h->tw2}; }
int check_common(repet_handle* h, const repet_params* p, int n_channels);

}  // namespace repet
