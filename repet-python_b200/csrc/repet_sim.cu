// REPET-SIM kernels: column normalisation, self-similarity, similar-frame selection with exact
// float64 certification, and the similar-frame median model.   Reference: repet.py:1209-1246
// (_selfsimilaritymatrix / _similaritymatrix), 1294-1383 (_localmaxima / _indices), 1511-1545
// (_simmask), 834-901 (the online loop).
#include "repet_kernels.cuh"
#include "fft2048.cuh"
#include "median_networks.cuh"

namespace repet {

// ------------------------------------------------------------------------------------------
// k_normalize  --  A[:, t] = V[:, t] / ||V[:, t]||_2                       repet.py:1220, 1240
// V = mean_c |X| comes from k_stft in fp32; the norm and the quotient are taken in float64.
// Outputs: An64[row][APITCH64] (exact dot products) and An32[row][KPAD] (fast GEMM operand,
// zero padded to a multiple of 32 floats = one 128-byte swizzle atom per K block).
// An all-zero frame gives 0/0 = NaN as in the reference (quirk Q18).   One warp per frame.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_normalize(const float* __restrict__ V, int n_rows, double* __restrict__ An64, float* __restrict__ An32, int round_tf32) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_rows) return;
    const float* __restrict__ v = V + (size_t)warp * PPITCH;
    double sum = 0.0;
    for (int k = lane; k < NBIN; k += 32) {
        const double x = (double)v[k];
        sum = fma(x, x, sum);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const double norm = sqrt(sum);
    for (int k = lane; k < KPAD; k += 32) {
        const double a = k < NBIN ? (double)v[k] / norm : 0.0;
        if (An64 && k < APITCH64) An64[(size_t)warp * APITCH64 + k] = a;
        if (An32) {
            float f = (float)a;
            if (round_tf32) {  // round-to-nearest TF32 here, so the tensor core's operand truncation is exact
                uint32_t bits;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(bits) : "f"(f));
                f = __uint_as_float(bits);
            }
            An32[(size_t)warp * KPAD + k] = f;
        }
    }
}

void launch_normalize(cudaStream_t st, const float* V, int n_rows, double* An64, float* An32, int round_tf32) {
    k_normalize<<<(n_rows * 32 + 255) / 256, 256, 0, st>>>(V, n_rows, An64, An32, round_tf32);
}

// ------------------------------------------------------------------------------------------
// k_selfsim_simt  --  S = A^T A in fp32 on the CUDA cores (fast pass, error bound tau)
// 64x64 output tile per CTA, 16x16 threads x 4x4 accumulators, K in chunks of 16 through smem.
// The products are summed in the same k order for (i, j) and (j, i): S is bitwise symmetric.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_selfsim_simt(const float* __restrict__ An32, int T, float* __restrict__ S) {
    __shared__ float sa[16][65];
    __shared__ float sb[16][65];
    const int item = blockIdx.z;
    const float* __restrict__ A = An32 + (size_t)item * T * KPAD;
    float* __restrict__ Sout = S + (size_t)item * T * T;
    const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
    const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int k0 = 0; k0 < KPAD; k0 += 16) {
        // 64 rows x 16 k per operand: 1024 elements, 4 per thread
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = threadIdx.x + 256 * e;
            const int row = idx >> 4, kk = idx & 15;
            const int ri = i0 + row, rj = j0 + row;
            sa[kk][row] = ri < T ? A[(size_t)ri * KPAD + k0 + kk] : 0.f;
            sb[kk][row] = rj < T ? A[(size_t)rj * KPAD + k0 + kk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = sa[kk][ti + 16 * a];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = sb[kk][tj + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = i0 + ti + 16 * a, j = j0 + tj + 16 * b;
            if (i < T && j < T) Sout[(size_t)j * T + i] = acc[a][b];
        }
}

void launch_selfsim_simt(cudaStream_t st, const float* An32, int n_items, int T, float* S) {
    dim3 grid((T + 63) / 64, (T + 63) / 64, n_items);
    k_selfsim_simt<<<grid, 256, 0, st>>>(An32, T, S);
}

// exact float64 dot product of two normalised frames, one warp (result in every lane)
__device__ __forceinline__ double warp_dot64(const double* __restrict__ a, const double* __restrict__ b, int lane) {
    double s = 0.0;
    for (int k = lane; k < NBIN; k += 32) s = fma(a[k], b[k], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// ------------------------------------------------------------------------------------------
// k_topk  --  _localmaxima / _indices with certification            repet.py:1294-1383
// One CTA per column c of the similarity matrix.  The fast values v~ (row c of the fp32 matrix,
// |v~ - v| <= tau) only PROPOSE candidates:
//   dropped for sure   : v~[i] < thr - tau, or some neighbour within +-d has v~[u] >= v~[i] + 2 tau
//   candidate          : otherwise; "certain" when no neighbour is within 2 tau and v~[i] >= thr + tau
// Every candidate's similarity is then recomputed EXACTLY (float64 dot of the float64-normalised
// frames), uncertain candidates are re-decided against the exact values of their near-tie
// neighbours, and the survivors are ranked by exact value (descending; ties by descending index,
// which is what argsort()[::-1] gives).  The first `number` indices are written.
// With tau = 0 this is the reference's rule verbatim (strict >, windows clipped at the ends,
// NaN never a maximum: quirk Q7).
// ------------------------------------------------------------------------------------------
struct TopkSmem {
    // dynamic layout: float v[T]; double col[APITCH64]; int cand[cap]; double exact[cap]; int flags...
};

__global__ void __launch_bounds__(256)
k_topk(const float* __restrict__ S, const double* __restrict__ An64, int T, float tau, double thr, int d, int number,
       int cap, int* __restrict__ idx_out, int* __restrict__ cnt_out, int* __restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* s_col = reinterpret_cast<double*>(smem);                 // [APITCH64]
    double* s_exact = s_col + APITCH64;                              // [cap]
    float* s_v = reinterpret_cast<float*>(s_exact + cap);            // [T]
    int* s_cand = reinterpret_cast<int*>(s_v + ((T + 3) & ~3));      // [cap]  index | uncertain << 30
    __shared__ int s_count, s_kept;
    const int item = blockIdx.y, c = blockIdx.x;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarp = blockDim.x >> 5;
    const float* __restrict__ row = S + ((size_t)item * T + c) * (size_t)T;
    const double* __restrict__ A = An64 + (size_t)item * T * APITCH64;
    if (t == 0) {
        s_count = 0;
        s_kept = 0;
    }
    for (int i = t; i < T; i += blockDim.x) s_v[i] = row[i];
    for (int k = t; k < APITCH64; k += blockDim.x) s_col[k] = A[(size_t)c * APITCH64 + k];
    __syncthreads();
    const float thr_f = (float)thr;
    // ---- propose -----------------------------------------------------------------------------
    for (int i = t; i < T; i += blockDim.x) {
        const float v = s_v[i];
        if (!(v >= thr_f - tau)) continue;  // also drops NaN
        bool dropped = false, uncertain = !(v >= thr_f + tau);
        const int lo = max(i - d, 0), hi = min(i + d, T - 1);
        for (int off = 1; off <= d && !dropped; ++off) {
            const int ul = i - off, ur = i + off;
            if (ul >= lo) {
                const float w = s_v[ul];
                if (!(w < v + 2.f * tau)) dropped = true;  // w >= v + 2 tau, or NaN
                else if (w > v - 2.f * tau) uncertain = true;
            }
            if (ur <= hi) {
                const float w = s_v[ur];
                if (!(w < v + 2.f * tau)) dropped = true;
                else if (w > v - 2.f * tau) uncertain = true;
            }
        }
        if (tau == 0.f) {
            // exact fast values: ties (w == v) are decided right here by the strict rule
            // (handled above: w >= v drops), nothing is uncertain
            uncertain = false;
        }
        if (!dropped) {
            const int slot = atomicAdd(&s_count, 1);
            if (slot < cap) s_cand[slot] = i | (uncertain ? (1 << 30) : 0);
        }
    }
    __syncthreads();
    int count = s_count;
    if (count > cap) {
        if (t == 0) atomicExch(overflow, 1);
        count = cap;
    }
    // ---- exact values of every candidate -------------------------------------------------------
    for (int q = warp; q < count; q += nwarp) {
        const int i = s_cand[q] & 0x3fffffff;
        const double e = warp_dot64(s_col, A + (size_t)i * APITCH64, lane);
        if (lane == 0) s_exact[q] = e;
    }
    __syncthreads();
    // ---- re-decide the uncertain ones against exact neighbour values ---------------------------
    for (int q = warp; q < count; q += nwarp) {
        const int code = s_cand[q];
        if (!(code & (1 << 30))) continue;
        const int i = code & 0x3fffffff;
        const double e = s_exact[q];
        const float v = s_v[i];
        bool keep = e >= thr;
        const int lo = max(i - d, 0), hi = min(i + d, T - 1);
        for (int u0 = lo; u0 <= hi && keep; u0 += 32) {
            const int u = u0 + lane;
            const bool near_tie = u <= hi && u != i && (s_v[u] > v - 2.f * tau);
            unsigned mask = __ballot_sync(0xffffffffu, near_tie);
            while (mask && keep) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                const double eu = warp_dot64(s_col, A + (size_t)(u0 + src) * APITCH64, lane);
                if (!(e > eu)) keep = false;  // strict >, NaN never passes
            }
        }
        if (lane == 0 && !keep) s_cand[q] = -1;
    }
    __syncthreads();
    // ---- rank the survivors by exact value ------------------------------------------------------
    for (int q = t; q < count; q += blockDim.x) {
        const int code = s_cand[q];
        if (code < 0) continue;
        const int i = code & 0x3fffffff;
        const double e = s_exact[q];
        int rank = 0;
        for (int r = 0; r < count; ++r) {
            const int other = s_cand[r];
            if (other < 0 || r == q) continue;
            const double eo = s_exact[r];
            const int io = other & 0x3fffffff;
            rank += (eo > e) || (eo == e && io > i);
        }
        atomicAdd(&s_kept, 1);
        if (rank < number) idx_out[((size_t)item * T + c) * (size_t)number + rank] = i;
    }
    __syncthreads();
    if (t == 0) cnt_out[(size_t)item * T + c] = min(s_kept, number);
}

size_t topk_smem_bytes(int T, int cap) {
    return (size_t)APITCH64 * 8 + (size_t)cap * 8 + (size_t)((T + 3) & ~3) * 4 + (size_t)cap * 4;
}

int launch_topk(cudaStream_t st, const float* S, const double* An64, int n_items, int T, float tau, double thr, int d,
                int number, int* idx_out, int* cnt_out, int* overflow) {
    int cap = T;
    size_t smem = topk_smem_bytes(T, cap);
    const size_t limit = 220 * 1024;
    if (smem > limit) {
        // keep the fast row resident, give the rest to candidates
        const size_t fixed = (size_t)APITCH64 * 8 + (size_t)((T + 3) & ~3) * 4;
        if (fixed + 12 * 256 > limit) return -1;  // row does not fit: the caller reports UNSUPPORTED
        cap = (int)((limit - fixed) / 12);
        smem = topk_smem_bytes(T, cap);
    }
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(k_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    dim3 grid(T, n_items);
    k_topk<<<grid, 256, smem, st>>>(S, An64, T, tau, thr, d, number, cap, idx_out, cnt_out, overflow);
    return 0;
}

// ------------------------------------------------------------------------------------------
// k_online_select  --  the per-frame selection of the online REPET-SIM     repet.py:834-866
// Frame j (>= B-1) is compared with the B frames in its ring buffer, visited in SLOT order:
// slot b holds frame j-(j0-b) for b <= j0 and j-(j0-b)-B for b > j0, j0 = j mod B (quirk Q6).
// Similarities are exact float64 dots of the float64-normalised frames, so the local-maximum
// rule and the ranking need no certification.  Writes FRAME indices.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_online_select(const double* __restrict__ An64, int T, int B, double thr, int d, int number, int* __restrict__ idx_out,
                int* __restrict__ cnt_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    double* s_col = reinterpret_cast<double*>(smem);  // [APITCH64]
    double* s_sim = s_col + APITCH64;                 // [B]
    int* s_keep = reinterpret_cast<int*>(s_sim + B);  // [B]
    __shared__ int s_kept;
    const int item = blockIdx.y;
    const int j = blockIdx.x + (B - 1);
    if (j >= T) return;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, nwarp = blockDim.x >> 5;
    const double* __restrict__ A = An64 + (size_t)item * T * APITCH64;
    const int j0 = j % B;
    if (t == 0) s_kept = 0;
    for (int k = t; k < APITCH64; k += blockDim.x) s_col[k] = A[(size_t)j * APITCH64 + k];
    __syncthreads();
    for (int b = warp; b < B; b += nwarp) {
        const int frame = b <= j0 ? j - (j0 - b) : j - (j0 - b) - B;
        const double e = warp_dot64(s_col, A + (size_t)frame * APITCH64, lane);
        if (lane == 0) s_sim[b] = e;
    }
    __syncthreads();
    for (int b = t; b < B; b += blockDim.x) {
        const double v = s_sim[b];
        bool keep = v >= thr;
        const int lo = max(b - d, 0), hi = min(b + d, B - 1);
        for (int u = lo; u <= hi && keep; ++u)
            if (u != b && !(v > s_sim[u])) keep = false;
        s_keep[b] = keep ? 1 : 0;
    }
    __syncthreads();
    for (int b = t; b < B; b += blockDim.x) {
        if (!s_keep[b]) continue;
        const double v = s_sim[b];
        int rank = 0;
        for (int u = 0; u < B; ++u)
            if (u != b && s_keep[u]) rank += (s_sim[u] > v) || (s_sim[u] == v && u > b);
        atomicAdd(&s_kept, 1);
        if (rank < number) {
            const int frame = b <= j0 ? j - (j0 - b) : j - (j0 - b) - B;
            idx_out[((size_t)item * T + j) * (size_t)number + rank] = frame;
        }
    }
    __syncthreads();
    if (t == 0) cnt_out[(size_t)item * T + j] = min(s_kept, number);
}

void launch_online_select(cudaStream_t st, const double* An64, int n_items, int T, int B, double thr, int d, int number,
                          int* idx_out, int* cnt_out) {
    if (T < B) return;
    const size_t smem = (size_t)APITCH64 * 8 + (size_t)B * 12;
    static size_t configured = 0;
    if (smem > configured && smem > 48 * 1024) {
        cudaFuncSetAttribute(k_online_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    dim3 grid(T - (B - 1), n_items);
    k_online_select<<<grid, 256, smem, st>>>(An64, T, B, thr, d, number, idx_out, cnt_out);
}

// ------------------------------------------------------------------------------------------
// k_simmodel  --  the similar-frame median of _simmask               repet.py:1529-1535, 872
// model[j][c][bin] = median over u in list(j) of |X_c[u][bin]| (an empty list gives NaN, as
// np.median of an empty selection does).  One CTA per (frame, item*channel).  Lists of up to 32
// frames go through the register selection networks on bin pairs; longer lists are staged as
// squared magnitudes in shared memory [n][256 bins] and selected by counting ranks.
// ------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ float2 gather_median_pair(const float2* __restrict__ chan, size_t row, const int* __restrict__ list,
                                                     int k, bool first_is_dc) {
    float v0[N], v1[N];
#pragma unroll
    for (int s = 0; s < N; ++s) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(chan + (size_t)list[s] * row + k));
        v0[s] = first_is_dc ? __fmul_rn(x.x, x.x) : __fmaf_rn(x.x, x.x, __fmul_rn(x.y, x.y));
        v1[s] = __fmaf_rn(x.z, x.z, __fmul_rn(x.w, x.w));
    }
    median_select<N>(v0);
    median_select<N>(v1);
    if (N & 1) return make_float2(fast_sqrt(v0[(N - 1) / 2]), fast_sqrt(v1[(N - 1) / 2]));
    return make_float2(0.5f * (fast_sqrt(v0[(N - 1) / 2]) + fast_sqrt(v0[N / 2])),
                       0.5f * (fast_sqrt(v1[(N - 1) / 2]) + fast_sqrt(v1[N / 2])));
}

template <int N>
__device__ __forceinline__ void simmodel_small(const float2* __restrict__ chan, size_t row, const int* __restrict__ list,
                                               float* __restrict__ out, int t) {
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const int k = 2 * (t + 128 * i);
        *reinterpret_cast<float2*>(out + k) = gather_median_pair<N>(chan, row, list, k, k == 0);
    }
    if (t == 0) {
        // Nyquist rides in bin 0's imaginary slot
        float v[N];
#pragma unroll
        for (int s = 0; s < N; ++s) {
            const float y = __ldg(&chan[(size_t)list[s] * row]).y;
            v[s] = __fmul_rn(y, y);
        }
        median_select<N>(v);
        out[XPITCH] = (N & 1) ? fast_sqrt(v[(N - 1) / 2]) : 0.5f * (fast_sqrt(v[(N - 1) / 2]) + fast_sqrt(v[N / 2]));
    }
}

// median of column `col` of a [n][pitch] shared-memory tile of squared magnitudes, by rank counting
__device__ __forceinline__ float tile_median(const float* __restrict__ tile, int n, int pitch, int col) {
    const int k_lo = (n - 1) >> 1, k_hi = n >> 1;
    float v_lo = 0.f, v_hi = 0.f;
    bool got_lo = false, got_hi = false;
    for (int a = 0; a < n && !(got_lo && got_hi); ++a) {
        const float va = tile[a * pitch + col];
        int less = 0, equal = 0;
        for (int b = 0; b < n; ++b) {
            const float vb = tile[b * pitch + col];
            less += vb < va;
            equal += vb == va;
        }
        if (!got_lo && less <= k_lo && k_lo < less + equal) {
            v_lo = va;
            got_lo = true;
        }
        if (!got_hi && less <= k_hi && k_hi < less + equal) {
            v_hi = va;
            got_hi = true;
        }
    }
    return 0.5f * (fast_sqrt(v_lo) + fast_sqrt(v_hi));
}

__global__ void __launch_bounds__(128)
k_simmodel(const float2* __restrict__ X, int T, int nch, const int* __restrict__ idx, const int* __restrict__ cnt,
           int number, int first_frame, float* __restrict__ model) {
    extern __shared__ __align__(16) unsigned char smem[];
    int* s_list = reinterpret_cast<int*>(smem);                       // [number]
    float* s_tile = reinterpret_cast<float*>(s_list + ((number + 3) & ~3));  // [n][256] when n > 32
    const int j = blockIdx.x + first_frame;
    const int item = blockIdx.y / nch, c = blockIdx.y - item * nch;
    const int t = threadIdx.x;
    const size_t row = (size_t)nch * XPITCH;
    const float2* __restrict__ chan = X + (size_t)item * T * row + (size_t)c * XPITCH;
    float* __restrict__ out = model + (((size_t)item * nch + c) * (size_t)T + j) * PPITCH;
    const int n = cnt[(size_t)item * T + j];
    for (int s = t; s < n; s += 128) s_list[s] = idx[((size_t)item * T + j) * (size_t)number + s];
    __syncthreads();
    switch (n) {
        case 0:
            for (int k = t; k <= XPITCH; k += 128) out[k] = nanf("");
            break;
#define REPET_CASE(N) case N: simmodel_small<N>(chan, row, s_list, out, t); break;
            REPET_CASE(1) REPET_CASE(2) REPET_CASE(3) REPET_CASE(4) REPET_CASE(5) REPET_CASE(6) REPET_CASE(7)
            REPET_CASE(8) REPET_CASE(9) REPET_CASE(10) REPET_CASE(11) REPET_CASE(12) REPET_CASE(13) REPET_CASE(14)
            REPET_CASE(15) REPET_CASE(16) REPET_CASE(17) REPET_CASE(18) REPET_CASE(19) REPET_CASE(20) REPET_CASE(21)
            REPET_CASE(22) REPET_CASE(23) REPET_CASE(24) REPET_CASE(25) REPET_CASE(26) REPET_CASE(27) REPET_CASE(28)
            REPET_CASE(29) REPET_CASE(30) REPET_CASE(31) REPET_CASE(32)
#undef REPET_CASE
        default: {
            // long lists: stage squared magnitudes [n][256 bins] per pass, then count ranks
            for (int pass = 0; pass < 4; ++pass) {
                const int k = 2 * t + 256 * pass;
                for (int s = 0; s < n; ++s) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(chan + (size_t)s_list[s] * row + k));
                    s_tile[s * 256 + 2 * t] = (k == 0) ? __fmul_rn(x.x, x.x) : __fmaf_rn(x.x, x.x, __fmul_rn(x.y, x.y));
                    s_tile[s * 256 + 2 * t + 1] = __fmaf_rn(x.z, x.z, __fmul_rn(x.w, x.w));
                }
                // each thread reads back only what it wrote: no barrier needed
                out[k] = tile_median(s_tile, n, 256, 2 * t);
                out[k + 1] = tile_median(s_tile, n, 256, 2 * t + 1);
            }
            if (t == 0) {
                for (int s = 0; s < n; ++s) {
                    const float y = __ldg(&chan[(size_t)s_list[s] * row]).y;
                    s_tile[s * 256] = __fmul_rn(y, y);
                }
                out[XPITCH] = tile_median(s_tile, n, 256, 0);
            }
        }
    }
}

int launch_simmodel(cudaStream_t st, const float2* X, int n_items, int T, int nch, const int* idx, const int* cnt,
                    int number, int first_frame, float* model) {
    const size_t smem = (size_t)((number + 3) & ~3) * 4 + (number > 32 ? (size_t)number * 256 * 4 : 0);
    if (smem > 220 * 1024) return -1;
    static size_t configured = 0;
    if (smem > configured && smem > 48 * 1024) {
        cudaFuncSetAttribute(k_simmodel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    if (T <= first_frame) return 0;
    dim3 grid(T - first_frame, n_items * nch);
    k_simmodel<<<grid, 128, smem, st>>>(X, T, nch, idx, cnt, number, first_frame, model);
    return 0;
}

}  // namespace repet
