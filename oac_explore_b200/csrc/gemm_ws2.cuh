// CTA-PAIR variant of the warp-specialised persistent tcgen05 kernel (gemm_ws.cuh): tcgen05.mma.cta_group::2.
//
// gemm_ws stages sit at ~48 GB/s of operand ingest per SM (7.1 TB/s chip-wide) whatever the tensor pipe could do: with
// N = 256 a 128 x 256 tile already spans the whole layer, so every 128-row strip re-streams the full weight matrix (405 KB
// of 608 KB per critic_l1 tile).  A CTA pair (two SMs of one TPC) computes a 256 x bn tile with ONE instruction stream:
// each CTA loads its own 128 rows of A and only HALF of the B tile (bn / 2 rows), the leader's tcgen05.mma.cta_group::2
// reads both halves, and each CTA keeps its 128 x bn accumulator in its own TMEM (two buffers, as before).  Operand bytes
// per SM and output drop by a third (608 -> 405 KB per 128 x 256 outputs), the ring slot shrinks from 48 to 32 KB (one more
// slot), and the two M-tiles of a B = 256 batch share every weight byte on chip.
//   * cluster of 2 CTAs; work items are (seed, task, 256-row tile pair, n-tile); CTA rank r owns rows [256 tm + 128 r, +128)
//     and B rows [n0 + r bn / 2, + bn / 2);
//   * both producers issue cp.async.bulk.tensor.cta_group::2 loads that complete on the LEADER's full barrier (expect_tx =
//     the bytes of both CTAs); the leader's MMA lane issues the pair's MMAs and frees ring slots / publishes accumulators in
//     BOTH CTAs with tcgen05.commit.cta_group::2 ... multicast::cluster; both epilogues arrive on the leader's
//     accumulator-empty barrier (remote mbarrier.arrive through mapa).
// Used for stages whose tasks all have M a multiple of 256 (the trunk layers, their dX and dW products); the others keep
// the single-CTA kernel.
#pragma once
#include <cuda.h>
#include "gemm_ws.cuh"

namespace oac {

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;\n" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t n_clusters_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;\n" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `local_smem_addr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are counted on a barrier that may live in the PEER CTA (cluster address)
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(WS_THREADS, 1) gemm_ws2_kernel(WsParams wp) {
    extern __shared__ __align__(1024) uint8_t ws_smem[];
    __shared__ __align__(8) uint64_t s_full[WS_MAX_SLOTS], s_empty[WS_MAX_SLOTS], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem;
    __shared__ int s_tile0[WS_MAX_TASKS + 1];

    pdl_wait();
    const StageParams& sp = wp.sp;
    const GemmTask* __restrict__ tasks = sp.tasks;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();              // 0: leader (issues the pair's MMAs)
    const int cl_id = (int)cluster_id_x(), n_cl = (int)n_clusters_x();
    uint8_t* ring = ws_smem + ((1024u - (smem_u32(ws_smem) & 1023u)) & 1023u);
    uint8_t* ones = ring + (size_t)wp.n_slots * wp.slot_bytes;
    float* slabs = reinterpret_cast<float*>(ones + WS_ONES_BYTES);

    // ---- one-time setup ----
    for (int i = tid; i <= wp.n_tasks; i += WS_THREADS) s_tile0[i] = (i < wp.n_tasks) ? tasks[i].tile0 : wp.tiles_per_seed;
    for (int i = tid; i < (int)(WS_ONES_BYTES / 4); i += WS_THREADS) reinterpret_cast<float*>(ones)[i] = 1.0f;
    fence_async_smem();
    if (tid == 0) {
        for (int i = 0; i < wp.n_slots; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
        mbar_init(&s_tfull[0], 1); mbar_init(&s_tfull[1], 1);
        // the accumulator-empty barriers are used in the leader only: the epilogue warps of BOTH CTAs arrive there
        mbar_init(&s_tempty[0], 2 * WS_EPI_WARPS); mbar_init(&s_tempty[1], 2 * WS_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                   // the peer's barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    auto decode = [&](int g, int& seed, int& j, int& tm, int& tn) {
        // seed is the FAST index: the CTAs of a round work on the same tile of different seeds, so every round costs
        // every CTA the same; tasks are sorted by cost, so the ragged last round is made of the cheapest tiles
        const int r = g / wp.n_seeds;
        seed = g - r * wp.n_seeds;
        j = 0;
        while (r >= s_tile0[j + 1]) ++j;
        const int t = r - s_tile0[j];
        const int tnc = tasks[j].tiles_n;
        tm = t / tnc; tn = t - tm * tnc;
    };

    if (warp == 0) {
        // =========================== TMA producer (both CTAs) ===========================
        int slot = 0; uint32_t ph = 0;
        for (int g = cl_id; g < wp.total_tiles; g += n_cl) {
            int seed, j, tm, tn;
            decode(g, seed, j, tm, tn);
            const GemmTask& T = tasks[j];
            const int bn = T.bn, bnh = bn >> 1;                                // this CTA loads bn / 2 rows of B
            const int m0 = tm * (2 * WS_BM) + (int)rank * WS_BM, n0 = tn * bn + (int)rank * bnh;
            if (lane == 0) {
                const CUtensorMap* ta = wp.tmaps + 2 * j;
                const CUtensorMap* tb = ta + 1;
                const int nch = (T.K + WS_KC - 1) / WS_KC;
                const uint32_t bytes = WS_A_BYTES + (uint32_t)bnh * (WS_KC * 4);   // per CTA
                for (int c = 0; c < nch; ++c) {
                    mbar_wait_relaxed(&s_empty[slot], ph ^ 1u);
                    const uint32_t sa = smem_u32(ring + (size_t)slot * wp.slot_bytes), sb = sa + WS_A_BYTES;
                    const uint32_t bar = mapa_u32(smem_u32(&s_full[slot]), 0);     // the LEADER's full barrier
                    if (rank == 0) mbar_expect_tx(smem_u32(&s_full[slot]), 2 * bytes);
                    if (!A_MN) tma_load_3d_2sm(sa, ta, c * WS_KC, m0, seed, bar);
                    else tma_load_4d_2sm(sa, ta, 0, c * WS_KC, m0 >> 5, seed, bar);      // all 32-wide atoms of the tile in one box
                    if (!B_MN) tma_load_3d_2sm(sb, tb, c * WS_KC, n0, seed, bar);
                    else tma_load_4d_2sm(sb, tb, 0, c * WS_KC, n0 >> 5, seed, bar);
                    if (++slot == wp.n_slots) { slot = 0; ph ^= 1u; }
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // =========================== MMA issuer (leader CTA only) ===========================
        if (lane == 0 && rank == 0) {
            int slot = 0; uint32_t ph = 0;
            int tl = 0;
            const uint64_t ones_desc = umma_desc(smem_u32(ones), 4096, 512, 1);
            const uint32_t idesc_bias = umma_idesc_tf32(2 * WS_BM, 32, true, true);
            for (int g = cl_id; g < wp.total_tiles; g += n_cl, ++tl) {
                int seed, j, tm, tn;
                decode(g, seed, j, tm, tn);
                const GemmTask& T = tasks[j];
                const int bn = T.bn, K = T.K;
                const int nch = (K + WS_KC - 1) / WS_KC;
                const int buf = tl & 1;
                mbar_wait_relaxed(&s_tempty[buf], (((uint32_t)tl >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_main = tmem + (uint32_t)(buf * 256), d_bias = d_main + WS_BIAS_COL;
                const uint32_t idesc = umma_idesc_tf32(2 * WS_BM, bn, A_MN, B_MN);
                const bool bias_mma = A_MN && B_MN && (T.epi == EPI_ADAM || T.epi == EPI_GRAD) && T.has_bias && tn == 0;
                for (int c = 0; c < nch; ++c) {
                    mbar_wait_relaxed(&s_full[slot], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(ring + (size_t)slot * wp.slot_bytes), sb = sa + WS_A_BYTES;
                    const int ksteps = (min(WS_KC, K - c * WS_KC) + 7) >> 3;
                    // K-major (SWIZZLE_128B): 8-row groups 1024 B apart (SBO), a k-step is 32 B inside the swizzle row.
                    // MN-major (SWIZZLE_128B_ATOM_32B): slot holds [mn-atom (32)][k (32 rows)][128 B]: atoms 4096 B apart
                    // (LBO), 4-row k-groups 512 B apart (SBO), a k-step is 8 rows = 1024 B.
                    uint64_t ad = A_MN ? umma_desc(sa, 4096, 512, 1) : umma_desc(sa, 16, 1024, 2);
                    uint64_t bd = B_MN ? umma_desc(sb, 4096, 512, 1) : umma_desc(sb, 16, 1024, 2);
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t acc = (c > 0 || ks > 0) ? 1u : 0u;
                        umma_tf32_2sm(d_main, ad, bd, idesc, acc);
                        if (bias_mma) umma_tf32_2sm(d_bias, ad, ones_desc, idesc_bias, acc);
                        ad += A_MN ? 64u : 2u;
                        bd += B_MN ? 64u : 2u;
                    }
                    umma_commit_2sm(&s_empty[slot]);             // frees the slot in BOTH CTAs once these MMAs have read it
                    if (++slot == wp.n_slots) { slot = 0; ph ^= 1u; }
                }
                umma_commit_2sm(&s_tfull[buf]);                  // both CTAs' accumulator halves are complete
            }
        }
    } else {
        // =========================== epilogue warps ===========================
        const int e = warp - 2;
        const int q = warp & 3;                                  // TMEM lane quarter this warp may read
        const int hsel = e >> 2;                                 // two warps per quarter alternate over the slabs
        float* slab = slabs + e * (32 * WS_SLAB_LD);
        int tl = 0;
        const uint32_t tempty_leader[2] = {mapa_u32(smem_u32(&s_tempty[0]), 0), mapa_u32(smem_u32(&s_tempty[1]), 0)};
        for (int g = cl_id; g < wp.total_tiles; g += n_cl, ++tl) {
            int seed, j, tm, tn;
            decode(g, seed, j, tm, tn);
            const GemmTask& T = tasks[j];
            const int m0 = tm * (2 * WS_BM) + (int)rank * WS_BM, n0 = tn * T.bn;      // this CTA's 128 rows, all bn columns
            const int buf = tl & 1;
            mbar_wait_relaxed(&s_tfull[buf], ((uint32_t)tl >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256);
            ws_tile_epilogue<A_MN, B_MN>(sp, T, seed, m0, n0, tn, t_base, slab, q, hsel, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader[buf]);      // the leader's MMA lane waits for both CTAs' epilogues
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                   // nobody frees TMEM / exits while the peer may still signal it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
    }
}

}  // namespace oac
