// Two dependent forward layers in ONE launch (latency regime):  h1 = relu(X W1^T + b1),  h2 = relu(h1 W2^T + b2).
//
// A single seed's step is a chain of small dependent kernels and every link costs ~6 us of launch / drain floor whatever
// the kernel does.  Layer 2 of a 32-row strip only needs layer 1 of the SAME strip, so the strip is
// given to a thread-block cluster: CTA c of the cluster computes the 32 x 32 tile (strip, columns 32c..) of layer 1 with
// the arithmetic of gemm_sk_kernel (same k-group split, same summation order: results are bit-identical to the two-launch
// schedule), stores it (h1 is needed by the backward pass anyway), the cluster meets at one barrier, and the CTA goes on
// to its tile of layer 2, whose A operand is the strip of h1 just written (read back through L2 by TMA) and whose W2
// slice was requested before the barrier.  One kernel boundary per two layers is gone from the critical chain.
//
// Task table: n pairs stored as [layer-1 tasks ..., layer-2 tasks ...] (blockIdx.y = pair); tensor maps as sk_tma_plan
// lays them out (A, B per task).  Requirements (checked by the host): K-contiguous operands, EPI_BIAS / EPI_BIAS_RELU,
// both layers with the same M and the same number of column tiles (= the cluster size, <= 8).
#pragma once
#include "gemm_simt.cuh"
#include "tma_util.cuh"

namespace oac {

__device__ __forceinline__ void mbar_expect_tx_only(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_acq_rel() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// the K-contiguous x K-contiguous FFMA tile of gemm_sk_body (TMA / SWIZZLE_128B operand atoms), this thread's k-group
__device__ __forceinline__ void fwd2_tile(const float* As, const float* Bs, int K, int kg, int r0, int c0, float (&acc)[SK_T][SK_T]) {
    auto kmaj = [](const float* base, int row, int k) {
        return reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(base) + (k >> 5) * 4096 + row * 128 +
                                               ((((k >> 2) & 7) ^ (row & 7)) << 4));
    };
    const int kn4 = (K + 3) & ~3;
    const int ks = ((kn4 >> 2) + SK_KS - 1) / SK_KS * 4;
    const int kb = kg * ks, ke = min(kb + ks, kn4);
#pragma unroll 2
    for (int k = kb; k < ke; k += 4) {
        float4 a[SK_T], b[SK_T];
#pragma unroll
        for (int i = 0; i < SK_T; ++i) a[i] = *kmaj(As, r0 + i * 8, k);
#pragma unroll
        for (int j = 0; j < SK_T; ++j) b[j] = *kmaj(Bs, c0 + j * 8, k);
#pragma unroll
        for (int i = 0; i < SK_T; ++i)
#pragma unroll
            for (int j = 0; j < SK_T; ++j) {
                acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
            }
    }
}

// the four k-groups' partial tiles meet in shared memory; every thread finishes one row x 4 adjacent columns
__device__ __forceinline__ void fwd2_finish(const StageParams& sp, const GemmTask& T, int seed, float* part, int m0, int n0,
                                            int kg, int r0, int c0, const float (&acc)[SK_T][SK_T]) {
    const int tid = threadIdx.x;
    const int er = tid >> 3, ec = (tid & 7) << 2;
#pragma unroll
    for (int i = 0; i < SK_T; ++i)
#pragma unroll
        for (int j = 0; j < SK_T; ++j) part[(kg * SK_BM + r0 + i * 8) * SK_PLD + c0 + j * 8] = acc[i][j];
    __syncthreads();
    const int m = m0 + er;
    if (m < T.M) {
        float* __restrict__ C = resolve(sp.as, T.C, seed);
        const float* bias = resolve(sp.as, T.bias, seed);
#pragma unroll
        for (int j = 0; j < SK_T; ++j) {
            const int n = n0 + ec + j;
            if (n >= T.N) continue;
            float x = part[er * SK_PLD + ec + j];
#pragma unroll
            for (int g = 1; g < SK_KS; ++g) x += part[(g * SK_BM + er) * SK_PLD + ec + j];
            x += bias[n];
            if (T.epi == EPI_BIAS_RELU) x = relu(x);
            C[(long long)m * T.ldc + n] = x;
        }
    }
}

__global__ void __launch_bounds__(SK_THREADS) gemm_fwd2_kernel(StageParams sp, int n_pairs) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t s_bar;
    const int pair = blockIdx.y, seed = blockIdx.z;
    const GemmTask& T1 = sp.tasks[pair];
    const GemmTask& T2 = sp.tasks[n_pairs + pair];
    const int tile = blockIdx.x;
    if (tile >= T1.tiles_m * T1.tiles_n) return;          // whole clusters leave together: tiles_n is the cluster size
    const int tm = tile / T1.tiles_n, tn = tile - tm * T1.tiles_n;
    const int m0 = tm * SK_BM, n0 = tn * SK_BN;
    const int tid = threadIdx.x;
    const int kg = tid >> 6, t64 = tid & 63;
    const int r0 = t64 >> 3, c0 = t64 & 7;
    const int K1 = T1.K, K2 = T2.K;

    float* base = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(smem) + ((1024u - (smem_u32(smem) & 1023u)) & 1023u));
    const int a1_bytes = sk_tile_bytes(false, K1), a2_bytes = sk_tile_bytes(false, K2);
    float* As1 = base;
    float* Bs1 = base + a1_bytes / 4;
    float* As2 = base;
    float* Bs2 = base + a2_bytes / 4;
    const uint32_t bar = smem_u32(&s_bar);
    const CUtensorMap* ta1 = sp.tmaps + 2 * pair;
    const CUtensorMap* ta2 = sp.tmaps + 2 * (n_pairs + pair);
    // the layer-2 weight slice may be requested while layer 1 finishes if it lands clear of the partial-tile area
    const bool early_b2 = a2_bytes >= (int)(sizeof(float) * SK_KS * SK_BM * SK_PLD);

    // ---- layer 1 ----
    if (tid == 0) {
        mbar_init(&s_bar, 1); fence_mbar_init();
        mbar_expect_tx(bar, (uint32_t)(2 * a1_bytes));
        const uint32_t sa = smem_u32(As1), sb = smem_u32(Bs1);
        for (int a = 0; a < ((K1 + 31) >> 5); ++a) tma_load_3d(sa + a * 4096, ta1, a * 32, m0, seed, bar);
        for (int a = 0; a < ((K1 + 31) >> 5); ++a) tma_load_3d(sb + a * 4096, ta1 + 1, a * 32, n0, seed, bar);
    }
    float acc[SK_T][SK_T];
#pragma unroll
    for (int i = 0; i < SK_T; ++i)
#pragma unroll
        for (int j = 0; j < SK_T; ++j) acc[i][j] = 0.f;
    __syncthreads();                                       // the mbarrier is initialised
    mbar_wait(&s_bar, 0);
    fwd2_tile(As1, Bs1, K1, kg, r0, c0, acc);
    __syncthreads();                                       // layer-1 operand tiles are dead
    if (tid == 0 && early_b2) {
        mbar_expect_tx_only(bar, (uint32_t)a2_bytes);      // no arrival yet: phase 1 completes with the A tile below
        const uint32_t sb = smem_u32(Bs2);
        for (int a = 0; a < ((K2 + 31) >> 5); ++a) tma_load_3d(sb + a * 4096, ta2 + 1, a * 32, n0, seed, bar);
    }
    fwd2_finish(sp, T1, seed, base, m0, n0, kg, r0, c0, acc);
    // h1 strip: generic-proxy stores of 8 CTAs -> async-proxy (TMA) reads of all of them
    fence_proxy_async();
    cluster_sync_acq_rel();
    fence_proxy_async();

    // ---- layer 2 ----
    if (tid == 0) {
        mbar_expect_tx(bar, (uint32_t)(early_b2 ? a2_bytes : 2 * a2_bytes));
        const uint32_t sa = smem_u32(As2), sb = smem_u32(Bs2);
        for (int a = 0; a < ((K2 + 31) >> 5); ++a) tma_load_3d(sa + a * 4096, ta2, a * 32, m0, seed, bar);
        if (!early_b2)
            for (int a = 0; a < ((K2 + 31) >> 5); ++a) tma_load_3d(sb + a * 4096, ta2 + 1, a * 32, n0, seed, bar);
    }
#pragma unroll
    for (int i = 0; i < SK_T; ++i)
#pragma unroll
        for (int j = 0; j < SK_T; ++j) acc[i][j] = 0.f;
    mbar_wait(&s_bar, 1);
    fwd2_tile(As2, Bs2, K2, kg, r0, c0, acc);
    __syncthreads();
    fwd2_finish(sp, T2, seed, base, m0, n0, kg, r0, c0, acc);
}

}  // namespace oac
