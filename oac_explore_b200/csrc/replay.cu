// GPU-resident replay buffer: ReplayBuffer.random_batch as one coalesced row-gather
// (replay_buffer.py:106-115), ReplayBufferCount counts (:180-197) and the ring append
// (:88-104).  HBM-bound byte movement: one CTA per sampled transition, 128-bit loads and
// stores when the row bases are 16-byte aligned, scalar otherwise.  The f64->f32 cast of
// utils/core.py:45 is done once at insert time, so gathered batches are bit-exact copies
// of float32(store rows).
#include <cuda_runtime.h>

#include "oac_error.h"
#include "../../include/oac_b200.h"

namespace oac {

constexpr int GATHER_THREADS = 128;

__device__ __forceinline__ void copy_row(float* __restrict__ dst, const float* __restrict__ src, int n,
                                         int tid, int nthreads) {
    const bool vec = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0;
    if (vec) {
        const int n4 = n >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = tid; i < n4; i += nthreads) d4[i] = __ldg(s4 + i);
        for (int i = (n4 << 2) + tid; i < n; i += nthreads) dst[i] = __ldg(src + i);
    } else {
        for (int i = tid; i < n; i += nthreads) dst[i] = __ldg(src + i);
    }
}

// one source row feeding up to three destinations (the obs copies of X blocks 0/1/2)
__device__ __forceinline__ void copy_row3(float* d0, float* d1, float* d2, const float* __restrict__ src, int n,
                                          int tid, int nthreads) {
    uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(d0);
    if (d1) al |= reinterpret_cast<uintptr_t>(d1);
    if (d2) al |= reinterpret_cast<uintptr_t>(d2);
    if ((al & 15) == 0) {
        const int n4 = n >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        for (int i = tid; i < n4; i += nthreads) {
            float4 v = __ldg(s4 + i);
            reinterpret_cast<float4*>(d0)[i] = v;
            if (d1) reinterpret_cast<float4*>(d1)[i] = v;
            if (d2) reinterpret_cast<float4*>(d2)[i] = v;
        }
        for (int i = (n4 << 2) + tid; i < n; i += nthreads) {
            float v = __ldg(src + i);
            d0[i] = v; if (d1) d1[i] = v; if (d2) d2[i] = v;
        }
    } else {
        for (int i = tid; i < n; i += nthreads) {
            float v = __ldg(src + i);
            d0[i] = v; if (d1) d1[i] = v; if (d2) d2[i] = v;
        }
    }
}

struct GatherArgs {
    OacReplayStore st;
    OacBatchDst dst;
    const int64_t* idx;
    int batch;
};

__global__ void __launch_bounds__(GATHER_THREADS) replay_gather_kernel(GatherArgs a) {
    const int i = blockIdx.x;                 // sample within the batch
    const int s = blockIdx.y;                 // seed
    const int B = a.batch, O = a.st.obs_dim, A = a.st.act_dim;
    const long long r = a.idx[(long long)s * B + i];
    const long long so = (long long)s * a.dst.seed_stride;
    float* x = a.dst.x + so;
    const int ld = a.dst.x_ld;
    auto xrow = [&](int block) -> float* { return block < 0 ? nullptr : x + ((long long)block * B + i) * ld; };
    float* d0 = xrow(a.dst.obs_blocks[0]);
    float* d1 = xrow(a.dst.obs_blocks[1]);
    float* d2 = xrow(a.dst.obs_blocks[2]);
    copy_row3(d0, d1, d2, a.st.obs + r * O, O, threadIdx.x, GATHER_THREADS);
    copy_row(xrow(a.dst.next_block), a.st.next_obs + r * O, O, threadIdx.x, GATHER_THREADS);
    {
        float* da = xrow(a.dst.act_block) + O;
        const float* sa = a.st.actions + r * A;
        for (int j = threadIdx.x; j < A; j += GATHER_THREADS) da[j] = __ldg(sa + j);
    }
    if (threadIdx.x == 0) {
        a.dst.rewards[so + i] = __ldg(a.st.rewards + r);
        a.dst.terminals[so + i] = __ldg(a.st.terminals + r);
        if (a.st.counts && a.dst.counts) a.dst.counts[so + i] = a.st.counts[r];
    }
}

// Many sampled rows per launch (seed groups: 16 384 rows for 64 seeds): one WARP per transition, every 16-byte piece of its
// obs and next_obs rows requested before the first store (six loads in flight per lane at O = 376), 8 transitions per CTA.
// The one-CTA-per-row kernel above leaves each CTA with two dependent memory latencies (index, then row) and ~3 KB in
// flight: at 16 resident CTAs per SM it moved 3.8 TB/s of read + written bytes (ncu: r02_64seeds_tf32_full.txt).
constexpr int GW_WARPS = 8, GW_MAX_IT = 4;      // rows of up to 32 * 4 float4 = 512 floats
__global__ void __launch_bounds__(GW_WARPS * 32) replay_gather_warp_kernel(GatherArgs a) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * GW_WARPS + (threadIdx.x >> 5);       // sample within the batch
    const int s = blockIdx.y;                                       // seed
    const int B = a.batch, O = a.st.obs_dim, A = a.st.act_dim;
    if (i >= B) return;
    const long long r = __ldg(a.idx + (long long)s * B + i);
    const long long so = (long long)s * a.dst.seed_stride;
    float* x = a.dst.x + so;
    const int ld = a.dst.x_ld;
    auto xrow = [&](int block) -> float* { return block < 0 ? nullptr : x + ((long long)block * B + i) * ld; };
    const int n4 = O >> 2;
    const float4* so4 = reinterpret_cast<const float4*>(a.st.obs + r * O);
    const float4* sn4 = reinterpret_cast<const float4*>(a.st.next_obs + r * O);
    float4 vo[GW_MAX_IT], vn[GW_MAX_IT];
#pragma unroll
    for (int k = 0; k < GW_MAX_IT; ++k) {
        const int j = lane + 32 * k;
        if (j < n4) { vo[k] = __ldg(so4 + j); vn[k] = __ldg(sn4 + j); }
    }
    float av = 0.f, rw = 0.f, tm = 0.f, cn = 0.f;
    if (lane < A) av = __ldg(a.st.actions + r * A + lane);           // A <= 32 (checked by the host)
    if (lane == 0) {
        rw = __ldg(a.st.rewards + r); tm = __ldg(a.st.terminals + r);
        if (a.st.counts && a.dst.counts) cn = a.st.counts[r];
    }
    float4* d0 = reinterpret_cast<float4*>(xrow(a.dst.obs_blocks[0]));
    float4* d1 = reinterpret_cast<float4*>(xrow(a.dst.obs_blocks[1]));
    float4* d2 = reinterpret_cast<float4*>(xrow(a.dst.obs_blocks[2]));
    float4* dn = reinterpret_cast<float4*>(xrow(a.dst.next_block));
#pragma unroll
    for (int k = 0; k < GW_MAX_IT; ++k) {
        const int j = lane + 32 * k;
        if (j < n4) {
            d0[j] = vo[k];
            if (d1) d1[j] = vo[k];
            if (d2) d2[j] = vo[k];
            dn[j] = vn[k];
        }
    }
    if (lane < A) xrow(a.dst.act_block)[O + lane] = av;
    if (lane == 0) {
        a.dst.rewards[so + i] = rw;
        a.dst.terminals[so + i] = tm;
        if (a.st.counts && a.dst.counts) a.dst.counts[so + i] = cn;
    }
}

// counts[idx] += 1 once per DISTINCT index (numpy fancy-index "+=", replay_buffer.py:195).  The index array may live in
// mapped pinned HOST memory (the single-seed path reads it in place): it is staged into shared memory with ONE coalesced
// pass, and the earlier-duplicate scan of every sample then runs on shared memory (it used to re-read idx[] from its home
// O(B^2) times -- up to 32k PCIe reads per batch).  Every CTA stages the prefix it needs: [0, end of its own range).
constexpr int BUMP_THREADS = 256;
constexpr int BUMP_MAX_SMEM_BATCH = 6144;      // 48 KB of int64
__global__ void __launch_bounds__(BUMP_THREADS) replay_counts_bump_kernel(float* counts, const int64_t* idx, int batch) {
    extern __shared__ long long s_idx[];
    const int i = blockIdx.x * BUMP_THREADS + threadIdx.x;
    const int need = min(batch, (int)(blockIdx.x + 1) * BUMP_THREADS);
    for (int j = threadIdx.x; j < need; j += BUMP_THREADS) s_idx[j] = idx[j];
    __syncthreads();
    if (i >= batch) return;
    const long long r = s_idx[i];
    for (int j = 0; j < i; ++j)
        if (s_idx[j] == r) return;        // an earlier occurrence owns the increment
    counts[r] += 1.0f;
}
// batches too large for the shared-memory staging (never the reference's 256): the plain scan
__global__ void replay_counts_bump_global_kernel(float* counts, const int64_t* idx, int batch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const long long r = idx[i];
    for (int j = 0; j < i; ++j)
        if (idx[j] == r) return;
    counts[r] += 1.0f;
}
static inline void launch_counts_bump(float* counts, const int64_t* idx, int batch, cudaStream_t st) {
    if (batch <= BUMP_MAX_SMEM_BATCH)
        replay_counts_bump_kernel<<<(batch + BUMP_THREADS - 1) / BUMP_THREADS, BUMP_THREADS, sizeof(long long) * (size_t)batch, st>>>(counts, idx, batch);
    else
        replay_counts_bump_global_kernel<<<(batch + 127) / 128, 128, 0, st>>>(counts, idx, batch);
}

struct DenseArgs {
    OacReplayStore st;
    const int64_t* idx;
    float *obs, *actions, *rewards, *terminals, *next_obs, *counts_out;
    int batch;
};

__global__ void __launch_bounds__(GATHER_THREADS) replay_gather_dense_kernel(DenseArgs a) {
    const int i = blockIdx.x;
    const int O = a.st.obs_dim, A = a.st.act_dim;
    const long long r = a.idx[i];
    copy_row(a.obs + (long long)i * O, a.st.obs + r * O, O, threadIdx.x, GATHER_THREADS);
    copy_row(a.next_obs + (long long)i * O, a.st.next_obs + r * O, O, threadIdx.x, GATHER_THREADS);
    copy_row(a.actions + (long long)i * A, a.st.actions + r * A, A, threadIdx.x, GATHER_THREADS);
    if (threadIdx.x == 0) {
        a.rewards[i] = __ldg(a.st.rewards + r);
        a.terminals[i] = __ldg(a.st.terminals + r);
        if (a.st.counts && a.counts_out) a.counts_out[i] = a.st.counts[r];
    }
}

struct AddArgs {
    float *obs, *next_obs, *actions, *rewards, *terminals, *counts;
    const float* rows;
    long long capacity, top;
    int O, A, n;
};

__global__ void __launch_bounds__(GATHER_THREADS) replay_add_kernel(AddArgs a) {
    const int i = blockIdx.x;
    const long long slot = (a.top + i) % a.capacity;
    const int W = 2 * a.O + a.A + 2;
    const float* row = a.rows + (long long)i * W;
    // n > capacity: a later row overwrites an earlier one; only the last writer of a slot may run
    if ((long long)i + a.capacity < (long long)a.n) return;
    for (int j = threadIdx.x; j < a.O; j += GATHER_THREADS) {
        a.obs[slot * a.O + j] = row[j];
        a.next_obs[slot * a.O + j] = row[a.O + a.A + 2 + j];
    }
    for (int j = threadIdx.x; j < a.A; j += GATHER_THREADS) a.actions[slot * a.A + j] = row[a.O + j];
    if (threadIdx.x == 0) {
        a.rewards[slot] = row[a.O + a.A];
        a.terminals[slot] = row[a.O + a.A + 1];
        if (a.counts) a.counts[slot] = 0.f;
    }
}

}  // namespace oac

using namespace oac;

extern "C" int oac_replay_gather(const OacReplayStore* store, const int64_t* indices, int32_t batch,
                                 const OacBatchDst* dst, void* stream) {
    if (!store || !indices || !dst || batch < 1) return set_error(OAC_E_INVALID, "oac_replay_gather: bad argument");
    if (dst->obs_blocks[0] < 0 || dst->act_block < 0 || dst->next_block < 0)
        return set_error(OAC_E_INVALID, "oac_replay_gather: destination blocks unset");
    const int seeds = dst->n_seeds > 0 ? dst->n_seeds : 1;
    if (store->counts && seeds > 1) return set_error(OAC_E_UNSUPPORTED, "counts with a store shared by several seeds");
    GatherArgs a{*store, *dst, indices, batch};
    cudaStream_t st = (cudaStream_t)stream;
    // seed groups (thousands of rows per launch) take the warp-per-row kernel when every row is made of whole, aligned
    // 16-byte pieces; a single trainer's 256 rows keep one CTA per row (latency: more SMs per row)
    const bool al16 = ((reinterpret_cast<uintptr_t>(store->obs) | reinterpret_cast<uintptr_t>(store->next_obs) |
                        reinterpret_cast<uintptr_t>(dst->x)) & 15) == 0;
    const bool warp_rows = (long long)batch * seeds >= 2048 && (store->obs_dim & 3) == 0 && (dst->x_ld & 3) == 0 &&
                           (dst->seed_stride & 3) == 0 && al16 && store->obs_dim <= 128 * GW_MAX_IT && store->act_dim <= 32;
    if (warp_rows) replay_gather_warp_kernel<<<dim3((batch + GW_WARPS - 1) / GW_WARPS, seeds), GW_WARPS * 32, 0, st>>>(a);
    else replay_gather_kernel<<<dim3(batch, seeds), GATHER_THREADS, 0, st>>>(a);
    OAC_CUDA(cudaGetLastError());
    if (store->counts) {
        launch_counts_bump(store->counts, indices, batch, st);
        OAC_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int oac_replay_gather_dense(const OacReplayStore* store, const int64_t* indices, int32_t batch,
                                       float* obs, float* actions, float* rewards, float* terminals,
                                       float* next_obs, float* counts_out, void* stream) {
    if (!store || !indices || batch < 1 || !obs || !actions || !rewards || !terminals || !next_obs)
        return set_error(OAC_E_INVALID, "oac_replay_gather_dense: bad argument");
    DenseArgs a{*store, indices, obs, actions, rewards, terminals, next_obs, counts_out, batch};
    cudaStream_t st = (cudaStream_t)stream;
    replay_gather_dense_kernel<<<batch, GATHER_THREADS, 0, st>>>(a);
    OAC_CUDA(cudaGetLastError());
    if (store->counts) {
        launch_counts_bump(store->counts, indices, batch, st);
        OAC_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int oac_replay_add(float* obs, float* next_obs, float* actions, float* rewards, float* terminals,
                              float* counts, int64_t capacity, int32_t obs_dim, int32_t act_dim,
                              const float* rows, int32_t n, int64_t top, void* stream) {
    if (!obs || !next_obs || !actions || !rewards || !terminals || !rows || n < 1 || capacity < 1)
        return set_error(OAC_E_INVALID, "oac_replay_add: bad argument");
    AddArgs a{obs, next_obs, actions, rewards, terminals, counts, rows, capacity, top, obs_dim, act_dim, n};
    replay_add_kernel<<<n, GATHER_THREADS, 0, (cudaStream_t)stream>>>(a);
    OAC_CUDA(cudaGetLastError());
    return 0;
}
