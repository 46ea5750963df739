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
        float* __restrict__ m1 = sp.as.base[AR_ADAM_M];
        float* __restrict__ m2 = sp.as.base[AR_ADAM_V];
        float* __restrict__ pb0 = sp.as.base[AR_PARAM];
        const int rsub = lane >> 4, c4 = (lane & 15) << 2;
        int tl = 0;
        const uint32_t tempty_leader[2] = {mapa_u32(smem_u32(&s_tempty[0]), 0), mapa_u32(smem_u32(&s_tempty[1]), 0)};
        for (int g = cl_id; g < wp.total_tiles; g += n_cl, ++tl) {
            int seed, j, tm, tn;
            decode(g, seed, j, tm, tn);
            const GemmTask& T = tasks[j];
            const int bn = T.bn, M = T.M, N = T.N, epi = T.epi, ldc = T.ldc;
            const int m0 = tm * (2 * WS_BM) + (int)rank * WS_BM, n0 = tn * bn;        // this CTA's 128 rows, all bn columns
            const int nlim = min(N, n0 + bn);
            const int buf = tl & 1;
            float* __restrict__ C = resolve(sp.as, T.C, seed);
            // the operand layouts pin the epilogue class (dW products are the only (MN, MN) tasks, masked dX products
            // the only (K, MN) ones): dead epilogues are compiled out, which keeps their registers out of the live set
            constexpr bool CAN_ADAM = A_MN && B_MN, CAN_MASK = !A_MN && B_MN;
            const bool is_adam = CAN_ADAM && epi == EPI_ADAM;
            const bool is_grad = CAN_ADAM && epi == EPI_GRAD;        // plain store of dW; Adam streams later (adam_stream.cuh)
            AdamScalars s;
            float inv_bc2 = 1.f;
            float* __restrict__ am = nullptr; float* __restrict__ av = nullptr; float* __restrict__ tg = nullptr;
            if (is_adam) {
                const int32_t* cnt = sp.as.counters + seed * sp.as.n_counters;
                s = make_adam_scalars_fast(sp.hyper, T.lr, cnt[T.counter], cnt[CNT_TRAIN_STEPS]);
                inv_bc2 = 1.0f / s.bc2_sqrt;
                am = m1 + (long long)seed * sp.as.stride[AR_ADAM_M] + T.adam_off;
                av = m2 + (long long)seed * sp.as.stride[AR_ADAM_V] + T.adam_off;
                if (T.target_off >= 0 && s.do_polyak) tg = pb0 + (long long)seed * sp.as.stride[AR_PARAM] + T.target_off;
            }
            const float* __restrict__ bias = (epi == EPI_BIAS || epi == EPI_BIAS_RELU) ? resolve(sp.as, T.bias, seed) : nullptr;
            const float* __restrict__ mask = (CAN_MASK && epi == EPI_MASK) ? resolve(sp.as, T.mask, seed) : nullptr;
            const int ldmask = T.ldmask;

            mbar_wait_relaxed(&s_tfull[buf], ((uint32_t)tl >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256);
            const bool rows_live = m0 + q * 32 < M;              // warp-uniform: nothing to write for this quarter
            for (int sl = hsel; sl * WS_SLAB < nlim - n0 && rows_live; sl += 2) {
                const int c0 = sl * WS_SLAB;
                // ReLU-mask epilogue: all 16 mask loads of this lane fly while the accumulator is read back and
                // transposed (one at a time they cost a DRAM latency each: 21 us per K=1 tile before this)
                float4 k4[16];
                const bool mask_vec = mask != nullptr && (n0 + c0 + c4 + 3 < nlim);
                if (mask_vec) {
#pragma unroll
                    for (int rp = 0; rp < 16; ++rp) {
                        const int m = m0 + q * 32 + 2 * rp + rsub;
                        k4[rp] = (m < M) ? __ldg(reinterpret_cast<const float4*>(mask + (long long)m * ldmask + n0 + c0 + c4))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {                 // 32 columns at a time: 32 live registers
                    if (c0 + 32 * hf >= bn) break;
                    float v[32];
                    tmem_ld16_nowait(t_base + (uint32_t)(c0 + 32 * hf), &v[0]);
                    if (c0 + 32 * hf + 16 < bn) tmem_ld16_nowait(t_base + (uint32_t)(c0 + 32 * hf + 16), &v[16]);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *reinterpret_cast<float4*>(slab + lane * WS_SLAB_LD + 32 * hf + 4 * i) =
                            make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
                __syncwarp();
                const int n = n0 + c0 + c4;
                if (n < nlim) {
                    const bool vec = n + 3 < nlim;
                    if (is_adam) {
                        constexpr int RB = 4;                    // row pairs per batch: 4 x 4 float4 loads in flight per lane
                        for (int rp0 = 0; rp0 < 16; rp0 += RB) {
                            float4 x[RB], p4[RB], a4[RB], v4[RB], t4[RB];
                            long long eo[RB];
#pragma unroll
                            for (int r = 0; r < RB; ++r) {
                                const int row = 2 * (rp0 + r) + rsub, m = m0 + q * 32 + row;
                                eo[r] = (m < M) ? (long long)m * ldc + n : -1;
                                x[r] = *reinterpret_cast<const float4*>(slab + row * WS_SLAB_LD + c4);
                                if (eo[r] >= 0 && vec) {
                                    p4[r] = *reinterpret_cast<const float4*>(C + eo[r]);
                                    a4[r] = *reinterpret_cast<const float4*>(am + eo[r]);
                                    v4[r] = *reinterpret_cast<const float4*>(av + eo[r]);
                                    if (tg) t4[r] = *reinterpret_cast<const float4*>(tg + eo[r]);
                                }
                            }
#pragma unroll
                            for (int r = 0; r < RB; ++r) {
                                if (eo[r] < 0) continue;
                                if (vec) {
                                    const bool ht = tg != nullptr;
                                    adam_core(x[r].x, p4[r].x, a4[r].x, v4[r].x, t4[r].x, ht, s, inv_bc2);
                                    adam_core(x[r].y, p4[r].y, a4[r].y, v4[r].y, t4[r].y, ht, s, inv_bc2);
                                    adam_core(x[r].z, p4[r].z, a4[r].z, v4[r].z, t4[r].z, ht, s, inv_bc2);
                                    adam_core(x[r].w, p4[r].w, a4[r].w, v4[r].w, t4[r].w, ht, s, inv_bc2);
                                    *reinterpret_cast<float4*>(C + eo[r]) = p4[r];
                                    *reinterpret_cast<float4*>(am + eo[r]) = a4[r];
                                    *reinterpret_cast<float4*>(av + eo[r]) = v4[r];
                                    if (ht) *reinterpret_cast<float4*>(tg + eo[r]) = t4[r];
                                } else {
#pragma unroll
                                    for (int jj = 0; jj < 3; ++jj) {             // a partial float4 holds at most 3 live columns
                                        if (n + jj >= nlim) continue;
                                        const float xj = jj == 0 ? x[r].x : (jj == 1 ? x[r].y : x[r].z);
                                        const long long ee = eo[r] + jj;
                                        float pp = C[ee], mm = am[ee], vv = av[ee], tt = tg ? tg[ee] : 0.f;
                                        adam_core(xj, pp, mm, vv, tt, tg != nullptr, s, inv_bc2);
                                        C[ee] = pp; am[ee] = mm; av[ee] = vv;
                                        if (tg) tg[ee] = tt;
                                    }
                                }
                            }
                        }
                    } else {
                        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (bias != nullptr) {
                            if (vec) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
                            else { b4.x = __ldg(bias + n); if (n + 1 < nlim) b4.y = __ldg(bias + n + 1); if (n + 2 < nlim) b4.z = __ldg(bias + n + 2); }
                        }
#pragma unroll
                        for (int rp = 0; rp < 16; ++rp) {
                            const int row = 2 * rp + rsub, m = m0 + q * 32 + row;
                            if (m >= M) continue;
                            float4 x = *reinterpret_cast<const float4*>(slab + row * WS_SLAB_LD + c4);
                            x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
                            if (epi == EPI_BIAS_RELU) { x.x = relu(x.x); x.y = relu(x.y); x.z = relu(x.z); x.w = relu(x.w); }
                            float* dst = C + (long long)m * ldc + n;
                            if (vec) {
                                if (mask != nullptr) {
                                    x.x = k4[rp].x > 0.f ? x.x : 0.f; x.y = k4[rp].y > 0.f ? x.y : 0.f;
                                    x.z = k4[rp].z > 0.f ? x.z : 0.f; x.w = k4[rp].w > 0.f ? x.w : 0.f;
                                }
                                *reinterpret_cast<float4*>(dst) = x;
                            } else {
#pragma unroll
                                for (int jj = 0; jj < 3; ++jj) {
                                    if (n + jj >= nlim) continue;
                                    float y = jj == 0 ? x.x : (jj == 1 ? x.y : x.z);
                                    if (mask != nullptr) y = __ldg(mask + (long long)m * ldmask + n + jj) > 0.f ? y : 0.f;
                                    dst[jj] = y;
                                }
                            }
                        }
                    }
                }
                __syncwarp();                                    // slab is rewritten by the next pass
            }
            // bias block of a dW task: column sums of dY sit in the spare TMEM columns (every column is the row sum)
            if (is_grad && T.has_bias && tn == 0 && hsel == 0 && rows_live) {
                const float gsum = tmem_ld1(t_base + WS_BIAS_COL);
                const int m = m0 + q * 32 + lane;
                if (m < M) resolve(sp.as, T.bias, seed)[m] = T.train_bias ? gsum : 0.f;      // a frozen bias gets a zero gradient
            }
            if (is_adam && T.has_bias && tn == 0 && hsel == 0 && rows_live) {
                const float gsum = tmem_ld1(t_base + WS_BIAS_COL);
                const int m = m0 + q * 32 + lane;
                if (m < M) {
                    float* pb = resolve(sp.as, T.bias, seed) + m;
                    float* tgb = T.target_bias_off >= 0 ? pb0 + (long long)seed * sp.as.stride[AR_PARAM] + T.target_bias_off + m : nullptr;
                    if (T.train_bias) {
                        adam_update(gsum, pb, m1 + (long long)seed * sp.as.stride[AR_ADAM_M] + T.adam_bias_off + m,
                                    m2 + (long long)seed * sp.as.stride[AR_ADAM_V] + T.adam_bias_off + m, tgb, s);
                    } else if (tgb != nullptr && s.do_polyak) {
                        *tgb = __fadd_rn(__fmul_rn(*tgb, s.one_m_tau), __fmul_rn(*pb, s.tau));
                    }
                }
            }
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
