// Optimistic exploration (optimistic_exploration.py:14-196) and the batch-1 inference
// calls around it (policy.get_action, trainer.predict) as fused per-observation kernels:
// one CTA per observation runs policy forward -> critics forward -> closed-form
// dQ_UB/d(pre-tanh mean) (no autograd tape) -> KL-constrained mean shift -> sample, with
// every activation in shared memory.  The reference issues ~120 launches for the same
// 1.3 MFLOP; at batch 1 the work is bounded by reading the ~2 MB
// of weights once.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "device_util.cuh"
#include "../../include/oac_b200.h"
#include "oac_error.h"

namespace oac {

constexpr int EX_THREADS = 512;
constexpr int EX_WARPS = EX_THREADS / 32;
constexpr int EX_MAX_H = 512;

// out[n] = act(sum_k W[n*ld + k] x[k] + b[n]) for n in [0,N): one warp per output row, FOUR rows at a time -- the
// loads of four rows (12 x 16 B per lane at K = 376) are in flight before the first reduction, so a layer costs ~4
// L2 round trips per warp instead of 16 (the batch-1 exploration kernel is pure load latency: 2 MB of weights, 1.3 MFLOP)
template <int NW = 16>
__device__ __forceinline__ void gemv_rows(const float* __restrict__ W, int ld, const float* __restrict__ b,
                                          const float* x, int K, int N, float* out, bool relu_) {
    constexpr int RB = 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    const int K4 = vec ? (K & ~3) : 0;
    for (int n0 = warp * RB; n0 < N; n0 += NW * RB) {
        float acc[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r] = 0.f;
        for (int k = lane * 4; k < K4; k += 128) {
            float4 wv[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r)
                wv[r] = (n0 + r < N) ? __ldg(reinterpret_cast<const float4*>(W + (long long)(n0 + r) * ld + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float x0 = x[k], x1 = x[k + 1], x2 = x[k + 2], x3 = x[k + 3];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                acc[r] = fmaf(wv[r].x, x0, acc[r]); acc[r] = fmaf(wv[r].y, x1, acc[r]);
                acc[r] = fmaf(wv[r].z, x2, acc[r]); acc[r] = fmaf(wv[r].w, x3, acc[r]);
            }
        }
        for (int k = K4 + lane; k < K; k += 32) {
            const float xk = x[k];
#pragma unroll
            for (int r = 0; r < RB; ++r)
                if (n0 + r < N) acc[r] = fmaf(__ldg(W + (long long)(n0 + r) * ld + k), xk, acc[r]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int r = 0; r < RB; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
        }
        if (lane < RB && n0 + lane < N) {
            float sel = acc[0];
#pragma unroll
            for (int r = 1; r < RB; ++r) sel = (lane == r) ? acc[r] : sel;
            const float v = sel + __ldg(b + n0 + lane);
            out[n0 + lane] = relu_ ? fmaxf(v, 0.f) : v;
        }
    }
}

// out[k] (+)= mask[k] > 0 ? sum_n v[n] W[n*ld + col0 + k] : 0  for k in [0,K); rows n in [0,N)
// part: scratch [EX_WARPS][K]
__device__ __forceinline__ void gemv_cols(const float* __restrict__ W, int ld, int col0, const float* v, int N, int K,
                                          const float* mask, float* out, float* part, bool accumulate) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[EX_MAX_H / 32];
#pragma unroll
    for (int c = 0; c < EX_MAX_H / 32; ++c) acc[c] = 0.f;
    // (throughput variant: many observations per launch keep the SMs busy, so rows are walked one at a time with few
    // registers -- two CTAs per SM; the latency variant below batches its loads instead)
    for (int n = warp; n < N; n += EX_WARPS) {
        const float vn = v[n];
        if (vn == 0.f) continue;                       // ReLU-masked rows contribute nothing
        const float* w = W + (long long)n * ld + col0;
#pragma unroll
        for (int c = 0; c < EX_MAX_H / 32; ++c) {
            int k = c * 32 + lane;
            if (k < K) acc[c] = fmaf(vn, __ldg(w + k), acc[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < EX_MAX_H / 32; ++c) {
        int k = c * 32 + lane;
        if (k < K) part[warp * K + k] = acc[c];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += EX_THREADS) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < EX_WARPS; ++w) s += part[w * K + k];
        if (mask && !(mask[k] > 0.f)) s = 0.f;
        out[k] = accumulate ? out[k] + s : s;
    }
    __syncthreads();
}

struct NetPtrs {
    const float *w0, *b0, *w1, *b1, *w2, *b2;
    int in_dim, in_ld, H, n_out;
};
__host__ __device__ inline NetPtrs net_ptrs(const float* base, const OacNetLayout& l) {
    NetPtrs p;
    p.w0 = base + l.off_w0; p.b0 = base + l.off_b0; p.w1 = base + l.off_w1; p.b1 = base + l.off_b1;
    p.w2 = base + l.off_w2; p.b2 = base + l.off_b2;
    p.in_dim = l.in_dim; p.in_ld = l.in_ld; p.H = l.hidden; p.n_out = l.n_out;
    return p;
}

constexpr int EX_MAX_Q = 16;

struct ExploreParams {
    NetPtrs policy;
    NetPtrs q[EX_MAX_Q];
    int n_q, mode, deterministic, quantile_index;
    unsigned exp_mask;
    float beta, sqrt_2delta;
    int O, A, H;
    const float* obs; const float* eps;
    unsigned long long rng_seed, rng_offset;
    float *action, *mu_E, *grad;
    const int* obs_group;          // per observation: which parameter arena (seed of a group) serves it, or NULL
    long long group_stride;        // floats between consecutive arenas
};

// dyn smem layout (floats): x[O+A pad] | h1[H] | h2[H] | head[2A] | qh1[n_q][H] | qh2[n_q][H] |
//                           qv[64] | cf[64] | g2[H] | g1[H] | da[A] | part[EX_WARPS*max(H,A)]
__global__ void __launch_bounds__(EX_THREADS) explore_kernel(ExploreParams p) {
    extern __shared__ __align__(16) float sm[];
    const int O = p.O, A = p.A, H = p.H;
    const int ob = blockIdx.x;
    const long long goff = p.obs_group ? (long long)p.obs_group[ob] * p.group_stride : 0;   // this observation's weights
    float* x = sm;
    float* h1 = x + ((O + A + 3) & ~3);
    float* h2 = h1 + H;
    float* head = h2 + H;
    float* qh1 = head + ((2 * A + 3) & ~3);
    float* qh2 = qh1 + p.n_q * H;
    float* qv = qh2 + p.n_q * H;
    float* cf = qv + 64;
    float* g2 = cf + 64;
    float* g1 = g2 + H;
    float* da = g1 + H;
    float* part = da + ((A + 3) & ~3);
    const int tid = threadIdx.x;

    for (int i = tid; i < O; i += EX_THREADS) x[i] = p.obs[(long long)ob * O + i];
    __syncthreads();
    // ---- policy forward (trainer/policies.py:260-283), deterministic head: a = tanh(mean) ----
    gemv_rows((p.policy.w0 + goff), p.policy.in_ld, (p.policy.b0 + goff), x, O, H, h1, true);
    __syncthreads();
    gemv_rows((p.policy.w1 + goff), H, (p.policy.b1 + goff), h1, H, H, h2, true);
    __syncthreads();
    gemv_rows((p.policy.w2 + goff), H, (p.policy.b2 + goff), h2, H, 2 * A, head, false);
    __syncthreads();
    for (int j = tid; j < A; j += EX_THREADS) x[O + j] = tanhf(head[j]);
    __syncthreads();
    // ---- critics forward ----
    int n_vals = 0;
    for (int q = 0; q < p.n_q; ++q) {
        const NetPtrs& N = p.q[q];
        gemv_rows((N.w0 + goff), N.in_ld, (N.b0 + goff), x, O + A, H, qh1 + q * H, true);
        __syncthreads();
        gemv_rows((N.w1 + goff), H, (N.b1 + goff), qh1 + q * H, H, H, qh2 + q * H, true);
        __syncthreads();
        gemv_rows((N.w2 + goff), H, (N.b2 + goff), qh2 + q * H, H, N.n_out, qv + n_vals, false);
        n_vals += N.n_out;
    }
    __syncthreads();
    // ---- dQ_UB / d(head outputs) ----
    if (tid == 0) {
        const int n_heads = p.q[0].n_out;
        for (int i = 0; i < n_vals; ++i)
            if (i < 32 && ((p.exp_mask >> i) & 1u)) qv[i] = expf(qv[i]);       // networks.py:69-75
        if (p.mode == OAC_EXPLORE_TWIN) {
            // Q_UB = (Q1+Q2)/2 + beta |Q1-Q2|/2     (optimistic_exploration.py:42-46,60)
            float dlt = qv[0] - qv[n_heads];
            float sg = dlt > 0.f ? 1.f : (dlt < 0.f ? -1.f : 0.f);
            for (int i = 0; i < n_vals; ++i) cf[i] = 0.f;
            cf[0] = 0.5f + 0.5f * p.beta * sg;
            cf[n_heads] = 0.5f - 0.5f * p.beta * sg;
        } else if (p.mode == OAC_EXPLORE_ENSEMBLE) {
            // mean + beta * std (unbiased)           (:47-58)
            float mean = 0.f;
            for (int i = 0; i < n_vals; ++i) mean += qv[i];
            mean /= (float)n_vals;
            float var = 0.f;
            for (int i = 0; i < n_vals; ++i) var += (qv[i] - mean) * (qv[i] - mean);
            var /= (float)(n_vals - 1);
            float sd = sqrtf(var);
            for (int i = 0; i < n_vals; ++i)
                cf[i] = 1.f / (float)n_vals + p.beta * (qv[i] - mean) / ((float)(n_vals - 1) * sd);
        } else {
            // sorted particle at quantile_index (ParticleTrainer.predict, particle_trainer_oac.py:147-167)
            for (int i = 0; i < n_vals; ++i) {
                int r = 0;
                for (int j = 0; j < n_vals; ++j) r += (qv[j] < qv[i]) || (qv[j] == qv[i] && j < i);
                cf[i] = (r == p.quantile_index) ? 1.f : 0.f;
            }
        }
        for (int i = 0; i < n_vals; ++i)
            if (i < 32 && ((p.exp_mask >> i) & 1u)) cf[i] *= qv[i];            // d exp(u)/du
    }
    for (int j = tid; j < A; j += EX_THREADS) da[j] = 0.f;
    __syncthreads();
    // ---- backward to the action, per critic ----
    int v0 = 0;
    for (int q = 0; q < p.n_q; ++q) {
        const NetPtrs& N = p.q[q];
        for (int n = tid; n < H; n += EX_THREADS) {
            float s = 0.f;
            for (int hd = 0; hd < N.n_out; ++hd) s = fmaf(cf[v0 + hd], __ldg((N.w2 + goff) + (long long)hd * H + n), s);
            g2[n] = qh2[q * H + n] > 0.f ? s : 0.f;
        }
        __syncthreads();
        gemv_cols((N.w1 + goff), H, 0, g2, H, H, qh1 + q * H, g1, part, false);
        gemv_cols((N.w0 + goff), N.in_ld, O, g1, H, A, nullptr, da, part, true);
        v0 += N.n_out;
    }
    // ---- shift + sample (one warp) ----
    if (tid < 32) {
        float num = 0.f;
        for (int j = tid; j < A; j += 32) {
            float a = x[O + j];
            float g = da[j] * (1.f - a * a);                // d tanh(mu)/d mu
            float sd = expf(fminf(fmaxf(head[A + j], LOG_SIG_MIN_F), LOG_SIG_MAX_F));
            float S = p.deterministic ? 1.f : sd * sd;      // :71 / L2 variant :160
            num += g * g * S;
        }
        num = warp_sum(num);
        const float denom = sqrtf(num) + 10e-6f;            // :76-80
        for (int j = tid; j < A; j += 32) {
            float a = x[O + j];
            float g = da[j] * (1.f - a * a);
            float sd = expf(fminf(fmaxf(head[A + j], LOG_SIG_MIN_F), LOG_SIG_MAX_F));
            float S = p.deterministic ? 1.f : sd * sd;
            float muE = head[j] + (p.sqrt_2delta * (S * g)) / denom;      // :83-87
            float out;
            if (p.deterministic) {
                out = muE;                                  // un-squashed (:181)
            } else {
                float e = p.eps ? p.eps[(long long)ob * A + j]
                                : philox_normal(p.rng_seed, 7u, (uint32_t)p.rng_offset, (uint32_t)(p.rng_offset >> 32) + ob, j);
                out = tanhf(fmaf(sd, e, muE));              // TanhNormal(mu_E, std).sample() (:92-94)
            }
            p.action[(long long)ob * A + j] = out;
            if (p.mu_E) p.mu_E[(long long)ob * A + j] = muE;
            if (p.grad) p.grad[(long long)ob * A + j] = g;
        }
    }
}

// =====================================================================================
// Latency variant: one thread-block CLUSTER of 8 CTAs per observation.
// A single CTA reads the ~2 MB of weights through one SM and walks ~13 dependent GEMV layers, each a few L2 round
// trips deep (~60 us).  Here every layer is split eight ways: forward layers by OUTPUT ROWS (each CTA computes H/8
// neurons and stores them into all eight CTAs' shared memory through DSMEM), the backward products by OUTPUT COLUMNS
// (no cross-CTA reduction), with one cluster barrier per exchanged layer; the tiny head layers, the coefficient math
// and dQ/da are computed redundantly by every CTA instead of being exchanged.  Used for up to EXC_MAX_OBS observations
// per call (the per-environment-step case); larger batches keep one CTA per observation.
// =====================================================================================
namespace cg = cooperative_groups;
constexpr int EXC_CS = 8;              // CTAs per cluster (portable maximum)
constexpr int EXC_THREADS = 256;
constexpr int EXC_WARPS = EXC_THREADS / 32;
constexpr int EXC_MAX_OBS = 16;

// rows [r0, r1) of  out = act(W x + b), stored into `out` of EVERY CTA of the cluster (same shared-memory offset)
__device__ __forceinline__ void gemv_rows_bcast(cg::cluster_group& cl, const float* __restrict__ W, int ld, const float* __restrict__ b,
                                                const float* x, int K, int r0, int r1, float* out, bool relu_) {
    constexpr int RB = 4;              // RB * EXC_CS == 32: lane -> (row r = lane & 3, destination rank = lane >> 2)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    const int K4 = vec ? (K & ~3) : 0;
    for (int n0 = r0 + warp * RB; n0 < r1; n0 += EXC_WARPS * RB) {
        float acc[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r] = 0.f;
        for (int k = lane * 4; k < K4; k += 128) {
            float4 wv[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r)
                wv[r] = (n0 + r < r1) ? __ldg(reinterpret_cast<const float4*>(W + (long long)(n0 + r) * ld + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float x0 = x[k], x1 = x[k + 1], x2 = x[k + 2], x3 = x[k + 3];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                acc[r] = fmaf(wv[r].x, x0, acc[r]); acc[r] = fmaf(wv[r].y, x1, acc[r]);
                acc[r] = fmaf(wv[r].z, x2, acc[r]); acc[r] = fmaf(wv[r].w, x3, acc[r]);
            }
        }
        for (int k = K4 + lane; k < K; k += 32) {
            const float xk = x[k];
#pragma unroll
            for (int r = 0; r < RB; ++r)
                if (n0 + r < r1) acc[r] = fmaf(__ldg(W + (long long)(n0 + r) * ld + k), xk, acc[r]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int r = 0; r < RB; ++r) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
        }
        const int r = lane & (RB - 1), rank = lane >> 2;
        if (n0 + r < r1) {
            float sel = acc[0];
#pragma unroll
            for (int q = 1; q < RB; ++q) sel = (r == q) ? acc[q] : sel;
            float v = sel + __ldg(b + n0 + r);
            v = relu_ ? fmaxf(v, 0.f) : v;
            cl.map_shared_rank(out, rank)[n0 + r] = v;
        }
    }
}

// columns [k0, k1) of  out[k] = mask[k] > 0 ? sum_n v[n] W[n*ld + col0 + k] : 0  over rows n in [0, N); `bcast`: stored
// into every CTA's `out`, else only locally (accumulating when `accumulate`).  part: scratch [EXC_WARPS][k1 - k0]
__device__ __forceinline__ void gemv_cols_slice(cg::cluster_group& cl, const float* __restrict__ W, int ld, int col0, const float* v, int N,
                                                int k0, int k1, const float* mask, float* out, float* part, bool bcast, bool accumulate) {
    constexpr int RB = 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int width = k1 - k0;
    for (int kb = 0; kb < width; kb += 32) {
        const int k = k0 + kb + lane;
        float acc = 0.f;
        for (int n0 = warp * RB; n0 < N; n0 += EXC_WARPS * RB) {
            float vn[RB], wv[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) vn[r] = (n0 + r < N) ? v[n0 + r] : 0.f;
#pragma unroll
            for (int r = 0; r < RB; ++r)
                wv[r] = (k < k1 && vn[r] != 0.f) ? __ldg(W + (long long)(n0 + r) * ld + col0 + k) : 0.f;   // ReLU-masked rows are skipped
#pragma unroll
            for (int r = 0; r < RB; ++r) acc = fmaf(vn[r], wv[r], acc);
        }
        if (kb + lane < width) part[warp * width + kb + lane] = acc;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < width; t += EXC_THREADS) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < EXC_WARPS; ++w) sum += part[w * width + t];
        const int k = k0 + t;
        if (mask && !(mask[k] > 0.f)) sum = 0.f;
        if (bcast) {
#pragma unroll
            for (int rk = 0; rk < EXC_CS; ++rk) cl.map_shared_rank(out, rk)[k] = sum;
        } else {
            out[k] = accumulate ? out[k] + sum : sum;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(EXC_THREADS) explore_cluster_kernel(ExploreParams p) {
    extern __shared__ __align__(16) float sm[];
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const int O = p.O, A = p.A, H = p.H;
    const int ob = blockIdx.x / EXC_CS;
    const long long goff = p.obs_group ? (long long)p.obs_group[ob] * p.group_stride : 0;   // this observation's weights
    float* x = sm;
    float* h1 = x + ((O + A + 3) & ~3);
    float* h2 = h1 + H;
    float* head = h2 + H;
    float* qh1 = head + ((2 * A + 3) & ~3);
    float* qh2 = qh1 + p.n_q * H;
    float* qv = qh2 + p.n_q * H;
    float* cf = qv + 64;
    float* g2 = cf + 64;                       // [n_q][H]
    float* g1 = g2 + p.n_q * H;                // [n_q][H]
    float* da = g1 + p.n_q * H;
    float* part = da + ((A + 3) & ~3);
    const int tid = threadIdx.x;
    const int hs = (H + EXC_CS - 1) / EXC_CS;  // this CTA's neurons / columns of a hidden layer: [s0, s1)
    const int s0 = min(H, rank * hs), s1 = min(H, s0 + hs);

    for (int i = tid; i < O; i += EXC_THREADS) x[i] = p.obs[(long long)ob * O + i];
    cl.sync();                                 // also: every CTA of the cluster is running before the first remote store
    // ---- policy forward: hidden layers split by rows, the 2A head rows on every CTA ----
    gemv_rows_bcast(cl, (p.policy.w0 + goff), p.policy.in_ld, (p.policy.b0 + goff), x, O, s0, s1, h1, true);
    cl.sync();
    gemv_rows_bcast(cl, (p.policy.w1 + goff), H, (p.policy.b1 + goff), h1, H, s0, s1, h2, true);
    cl.sync();
    gemv_rows<EXC_WARPS>((p.policy.w2 + goff), H, (p.policy.b2 + goff), h2, H, 2 * A, head, false);     // 2A short rows, redundantly per CTA
    __syncthreads();
    for (int j = tid; j < A; j += EXC_THREADS) x[O + j] = tanhf(head[j]);
    __syncthreads();
    // ---- critics forward ----
    for (int q = 0; q < p.n_q; ++q) gemv_rows_bcast(cl, (p.q[q].w0 + goff), p.q[q].in_ld, (p.q[q].b0 + goff), x, O + A, s0, s1, qh1 + q * H, true);
    cl.sync();
    for (int q = 0; q < p.n_q; ++q) gemv_rows_bcast(cl, (p.q[q].w1 + goff), H, (p.q[q].b1 + goff), qh1 + q * H, H, s0, s1, qh2 + q * H, true);
    cl.sync();
    int n_vals = 0;
    {
        const int warp = tid >> 5, lane = tid & 31;
        for (int q = 0; q < p.n_q; ++q) {
            const NetPtrs& N = p.q[q];
            for (int n = warp; n < N.n_out; n += EXC_WARPS) {
                const float* w = (N.w2 + goff) + (long long)n * H;
                float acc = 0.f;
                for (int k = lane; k < H; k += 32) acc = fmaf(__ldg(w + k), qh2[q * H + k], acc);
                acc = warp_sum(acc);
                if (lane == 0) qv[n_vals + n] = acc + __ldg((N.b2 + goff) + n);
            }
            n_vals += N.n_out;
        }
    }
    __syncthreads();
    // ---- dQ_UB / d(head outputs): same arithmetic as explore_kernel, on every CTA ----
    if (tid == 0) {
        const int n_heads = p.q[0].n_out;
        for (int i = 0; i < n_vals; ++i)
            if (i < 32 && ((p.exp_mask >> i) & 1u)) qv[i] = expf(qv[i]);
        if (p.mode == OAC_EXPLORE_TWIN) {
            float dlt = qv[0] - qv[n_heads];
            float sg = dlt > 0.f ? 1.f : (dlt < 0.f ? -1.f : 0.f);
            for (int i = 0; i < n_vals; ++i) cf[i] = 0.f;
            cf[0] = 0.5f + 0.5f * p.beta * sg;
            cf[n_heads] = 0.5f - 0.5f * p.beta * sg;
        } else if (p.mode == OAC_EXPLORE_ENSEMBLE) {
            float mean = 0.f;
            for (int i = 0; i < n_vals; ++i) mean += qv[i];
            mean /= (float)n_vals;
            float var = 0.f;
            for (int i = 0; i < n_vals; ++i) var += (qv[i] - mean) * (qv[i] - mean);
            var /= (float)(n_vals - 1);
            float sd = sqrtf(var);
            for (int i = 0; i < n_vals; ++i)
                cf[i] = 1.f / (float)n_vals + p.beta * (qv[i] - mean) / ((float)(n_vals - 1) * sd);
        } else {
            for (int i = 0; i < n_vals; ++i) {
                int r = 0;
                for (int j = 0; j < n_vals; ++j) r += (qv[j] < qv[i]) || (qv[j] == qv[i] && j < i);
                cf[i] = (r == p.quantile_index) ? 1.f : 0.f;
            }
        }
        for (int i = 0; i < n_vals; ++i)
            if (i < 32 && ((p.exp_mask >> i) & 1u)) cf[i] *= qv[i];
    }
    for (int j = tid; j < A; j += EXC_THREADS) da[j] = 0.f;
    __syncthreads();
    // ---- backward to the action: g2 locally, g1 split by columns and exchanged, dQ/da on every CTA ----
    {
        int v0 = 0;
        for (int q = 0; q < p.n_q; ++q) {
            const NetPtrs& N = p.q[q];
            for (int n = tid; n < H; n += EXC_THREADS) {
                float sacc = 0.f;
                for (int hd = 0; hd < N.n_out; ++hd) sacc = fmaf(cf[v0 + hd], __ldg((N.w2 + goff) + (long long)hd * H + n), sacc);
                g2[q * H + n] = qh2[q * H + n] > 0.f ? sacc : 0.f;
            }
            v0 += N.n_out;
        }
    }
    __syncthreads();
    for (int q = 0; q < p.n_q; ++q)
        gemv_cols_slice(cl, (p.q[q].w1 + goff), H, 0, g2 + q * H, H, s0, s1, qh1 + q * H, g1 + q * H, part, true, false);
    cl.sync();
    if (rank == 0) {
        for (int q = 0; q < p.n_q; ++q)
            gemv_cols_slice(cl, (p.q[q].w0 + goff), p.q[q].in_ld, O, g1 + q * H, H, 0, A, nullptr, da, part, false, true);
        // ---- shift + sample (one warp) ----
        if (tid < 32) {
            float num = 0.f;
            for (int j = tid; j < A; j += 32) {
                float a = x[O + j];
                float g = da[j] * (1.f - a * a);
                float sd = expf(fminf(fmaxf(head[A + j], LOG_SIG_MIN_F), LOG_SIG_MAX_F));
                float S = p.deterministic ? 1.f : sd * sd;
                num += g * g * S;
            }
            num = warp_sum(num);
            const float denom = sqrtf(num) + 10e-6f;
            for (int j = tid; j < A; j += 32) {
                float a = x[O + j];
                float g = da[j] * (1.f - a * a);
                float sd = expf(fminf(fmaxf(head[A + j], LOG_SIG_MIN_F), LOG_SIG_MAX_F));
                float S = p.deterministic ? 1.f : sd * sd;
                float muE = head[j] + (p.sqrt_2delta * (S * g)) / denom;
                float outv;
                if (p.deterministic) {
                    outv = muE;
                } else {
                    float e = p.eps ? p.eps[(long long)ob * A + j]
                                    : philox_normal(p.rng_seed, 7u, (uint32_t)p.rng_offset, (uint32_t)(p.rng_offset >> 32) + ob, j);
                    outv = tanhf(fmaf(sd, e, muE));
                }
                p.action[(long long)ob * A + j] = outv;
                if (p.mu_E) p.mu_E[(long long)ob * A + j] = muE;
                if (p.grad) p.grad[(long long)ob * A + j] = g;
            }
        }
    }
    cl.sync();          // no CTA exits while a sibling may still store into its shared memory
}

// ---- TanhGaussianPolicy.forward on n rows (one CTA per row) ----
struct PolicyFwdParams {
    NetPtrs net;
    const float* obs; int obs_ld; const float* eps;
    float *action, *mean, *log_std, *std, *pre_tanh, *log_prob;
};
__global__ void __launch_bounds__(EX_THREADS) policy_forward_kernel(PolicyFwdParams p) {
    extern __shared__ __align__(16) float sm[];
    const int O = p.net.in_dim, H = p.net.H, A = p.net.n_out / 2;
    const int r = blockIdx.x, tid = threadIdx.x;
    float* x = sm;
    float* h1 = x + ((O + 3) & ~3);
    float* h2 = h1 + H;
    float* head = h2 + H;
    for (int i = tid; i < O; i += EX_THREADS) x[i] = p.obs[(long long)r * p.obs_ld + i];
    __syncthreads();
    gemv_rows(p.net.w0, p.net.in_ld, p.net.b0, x, O, H, h1, true);
    __syncthreads();
    gemv_rows(p.net.w1, H, p.net.b1, h1, H, H, h2, true);
    __syncthreads();
    gemv_rows(p.net.w2, H, p.net.b2, h2, H, 2 * A, head, false);
    __syncthreads();
    if (tid < 32) {
        float lp = 0.f;
        for (int j = tid; j < A; j += 32) {
            float mean = head[j];
            float ls = fminf(fmaxf(head[A + j], LOG_SIG_MIN_F), LOG_SIG_MAX_F);
            float sd = expf(ls);
            float z = mean, a;
            if (p.eps) {
                z = fmaf(sd, p.eps[(long long)r * A + j], mean);
                a = tanhf(z);
                float d = z - mean;
                lp += -(d * d) / (2.f * sd * sd) - logf(sd) - 0.91893853320467274178f - logf(1.f - a * a + TANH_EPS_F);
            } else {
                a = tanhf(mean);
            }
            long long o = (long long)r * A + j;
            if (p.action) p.action[o] = a;
            if (p.mean) p.mean[o] = mean;
            if (p.log_std) p.log_std[o] = ls;
            if (p.std) p.std[o] = sd;
            if (p.pre_tanh) p.pre_tanh[o] = z;
        }
        lp = warp_sum(lp);
        if (tid == 0 && p.log_prob) p.log_prob[r] = lp;
    }
}

// ---- FlattenMlp.forward on n rows ----
struct QFwdParams {
    NetPtrs net;
    const float* x; int x_ld; unsigned exp_mask; float* out;
};
__global__ void __launch_bounds__(EX_THREADS) q_forward_kernel(QFwdParams p) {
    extern __shared__ __align__(16) float sm[];
    const int K = p.net.in_dim, H = p.net.H, NO = p.net.n_out;
    const int r = blockIdx.x, tid = threadIdx.x;
    float* x = sm;
    float* h1 = x + ((K + 3) & ~3);
    float* h2 = h1 + H;
    float* out = h2 + H;
    for (int i = tid; i < K; i += EX_THREADS) x[i] = p.x[(long long)r * p.x_ld + i];
    __syncthreads();
    gemv_rows(p.net.w0, p.net.in_ld, p.net.b0, x, K, H, h1, true);
    __syncthreads();
    gemv_rows(p.net.w1, H, p.net.b1, h1, H, H, h2, true);
    __syncthreads();
    gemv_rows(p.net.w2, H, p.net.b2, h2, H, NO, out, false);
    __syncthreads();
    for (int i = tid; i < NO; i += EX_THREADS) {
        float v = out[i];
        if ((p.exp_mask >> i) & 1u) v = expf(v);
        p.out[(long long)r * NO + i] = v;
    }
}

template <typename K>
static int ensure_smem(K kernel, size_t bytes) {
    if (bytes > 200 * 1024) return set_error(OAC_E_UNSUPPORTED, "network too large for the per-observation kernels");
    if (bytes > 48 * 1024) OAC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

}  // namespace oac

using namespace oac;

extern "C" int oac_explore(const OacExploreArgs* a, void* stream) {
    if (!a || !a->policy || !a->obs || !a->action || a->n_obs < 1 || a->n_q < 1 || a->n_q > EX_MAX_Q)
        return set_error(OAC_E_INVALID, "oac_explore: bad argument");
    const int O = a->policy_lay.in_dim, A = a->policy_lay.n_out / 2, H = a->policy_lay.hidden;
    if (H > EX_MAX_H || A > EX_MAX_H) return set_error(OAC_E_UNSUPPORTED, "oac_explore: hidden > 512");
    if (a->q_lay.in_dim != O + A || a->q_lay.hidden != H) return set_error(OAC_E_INVALID, "oac_explore: critic shape");
    const int n_vals = a->n_q * a->q_lay.n_out;
    if (n_vals > 64) return set_error(OAC_E_UNSUPPORTED, "oac_explore: more than 64 critic outputs");
    if (a->mode == OAC_EXPLORE_TWIN && a->n_q < 2) return set_error(OAC_E_INVALID, "oac_explore: twin mode needs 2 critics");
    if (a->mode == OAC_EXPLORE_ENSEMBLE && n_vals < 2) return set_error(OAC_E_INVALID, "oac_explore: ensemble of one");
    ExploreParams p;
    p.policy = net_ptrs(a->policy, a->policy_lay);
    p.n_q = a->mode == OAC_EXPLORE_TWIN ? 2 : a->n_q;      // twin: only q[0], q[1] (dispatch quirk :42-46)
    for (int i = 0; i < p.n_q; ++i) {
        if (!a->q[i]) return set_error(OAC_E_INVALID, "oac_explore: null critic");
        p.q[i] = net_ptrs(a->q[i], a->q_lay);
    }
    p.mode = a->mode; p.deterministic = a->deterministic; p.quantile_index = a->quantile_index;
    p.exp_mask = a->exp_mask; p.beta = a->beta_UB;
    p.sqrt_2delta = (float)sqrt(2.0 * (double)a->delta);
    p.O = O; p.A = A; p.H = H; p.obs = a->obs; p.eps = a->eps;
    p.rng_seed = a->rng_seed; p.rng_offset = a->rng_offset;
    p.action = a->action; p.mu_E = a->mu_E; p.grad = a->grad;
    p.obs_group = a->obs_group; p.group_stride = a->group_stride;
    size_t fl = ((O + A + 3) & ~3) + 2 * H + ((2 * A + 3) & ~3) + 2 * (size_t)p.n_q * H + 128 + 2 * H +
                ((A + 3) & ~3) + (size_t)EX_WARPS * (H > A ? H : A);
    static const bool no_cluster = getenv("OAC_NO_CLUSTER") && getenv("OAC_NO_CLUSTER")[0] == '1';
    if (a->n_obs <= EXC_MAX_OBS && !no_cluster) {
        // latency variant: a cluster of 8 CTAs per observation (layers split eight ways, activations through DSMEM)
        const int wmax = (H + EXC_CS - 1) / EXC_CS > A ? (H + EXC_CS - 1) / EXC_CS : A;
        size_t flc = ((O + A + 3) & ~3) + 2 * H + ((2 * A + 3) & ~3) + 2 * (size_t)p.n_q * H + 128 + 2 * (size_t)p.n_q * H +
                     ((A + 3) & ~3) + (size_t)EXC_WARPS * wmax;
        if (int e = ensure_smem(explore_cluster_kernel, flc * sizeof(float))) return e;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(a->n_obs * EXC_CS); cfg.blockDim = dim3(EXC_THREADS);
        cfg.dynamicSmemBytes = flc * sizeof(float); cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = EXC_CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        OAC_CUDA(cudaLaunchKernelEx(&cfg, explore_cluster_kernel, p));
        return 0;
    }
    if (int e = ensure_smem(explore_kernel, fl * sizeof(float))) return e;
    explore_kernel<<<a->n_obs, EX_THREADS, fl * sizeof(float), (cudaStream_t)stream>>>(p);
    OAC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int oac_policy_forward(const float* net, const OacNetLayout* lay, const float* obs, int32_t obs_ld,
                                  int32_t n, const float* eps, float* action, float* mean, float* log_std,
                                  float* std, float* pre_tanh, float* log_prob, void* stream) {
    if (!net || !lay || !obs || n < 1) return set_error(OAC_E_INVALID, "oac_policy_forward: bad argument");
    if (lay->hidden > EX_MAX_H) return set_error(OAC_E_UNSUPPORTED, "oac_policy_forward: hidden > 512");
    PolicyFwdParams p{net_ptrs(net, *lay), obs, obs_ld, eps, action, mean, log_std, std, pre_tanh, log_prob};
    size_t fl = ((lay->in_dim + 3) & ~3) + 2 * (size_t)lay->hidden + lay->n_out + 4;
    if (int e = ensure_smem(policy_forward_kernel, fl * sizeof(float))) return e;
    policy_forward_kernel<<<n, EX_THREADS, fl * sizeof(float), (cudaStream_t)stream>>>(p);
    OAC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int oac_q_forward(const float* net, const OacNetLayout* lay, const float* x, int32_t x_ld, int32_t n,
                             uint32_t exp_mask, float* out, void* stream) {
    if (!net || !lay || !x || !out || n < 1) return set_error(OAC_E_INVALID, "oac_q_forward: bad argument");
    if (lay->hidden > EX_MAX_H) return set_error(OAC_E_UNSUPPORTED, "oac_q_forward: hidden > 512");
    QFwdParams p{net_ptrs(net, *lay), x, x_ld, exp_mask, out};
    size_t fl = ((lay->in_dim + 3) & ~3) + 2 * (size_t)lay->hidden + lay->n_out + 4;
    if (int e = ensure_smem(q_forward_kernel, fl * sizeof(float))) return e;
    q_forward_kernel<<<n, EX_THREADS, fl * sizeof(float), (cudaStream_t)stream>>>(p);
    OAC_CUDA(cudaGetLastError());
    return 0;
}
