// Strip-fused forward chain on the tensor pipe: l1 -> ReLU -> l2 -> ReLU (-> head) of one 128-row strip on ONE persistent CTA,
// with the hidden activations handed from layer to layer through TENSOR MEMORY instead of HBM.
//
//   L1  D1[128 x 256] (TMEM columns 0..255)   = X_strip[128 x K0] W0^T      A and B chunks through the TMA ring (gemm_ws.cuh)
//   E1  the epilogue warps read D1 (tcgen05.ld, lane = row), add the bias, apply ReLU, store h1 to global memory (the
//       backward pass needs it) and write the tf32-rounded values BACK into the same TMEM columns (tcgen05.st): an
//       accumulator converted in place has exactly the layout tcgen05.mma wants for an A operand in tensor memory
//       (lane = row m, one 32-bit column per k)
//   L2  D2[128 x 256] (TMEM columns 256..511) = h1 W1^T                     A from TMEM, only the W1 chunks go through the ring
//   E2  D2 -> bias, ReLU -> h2 to global memory (and, when a head layer follows, back into TMEM in place)
//   L3  D3[128 x bn3] (TMEM columns 0..bn3-1) = h2 Wh^T                     A from TMEM (policy: mean | log_std rows)
//   E3  D3 -> bias -> head outputs
// What this removes per strip: the store -> kernel boundary -> TMA re-load of h1 (and of h2 before a head layer), two of
// the three tile life cycles (descriptor fetch, first TMA round trip, pipeline fill, drain) and two launches.  What it does
// not remove: the weight tiles are re-streamed per strip as before.  L1 of the next strip overlaps E2 of the current one
// (D1 is free once L2's MMAs, which read it, have been issued: MMAs of one CTA execute in issue order).
// Numerics: h1 / h2 reach the next layer rounded to tf32 (cvt.rna) exactly as the TFLOAT32 tensor maps round them in the
// unfused path.  The copies stored to global memory for the backward pass are the ROUNDED values (the store reads the
// converted accumulator back while the next layer's MMAs run): every consumer of a stored hidden activation either rounds
// it to tf32 itself (TFLOAT32 tensor maps of the dW products: rounding is idempotent) or only tests its sign (ReLU masks),
// except the critic head's fp32 dot product, which then sees tf32-rounded h2 rows (covered by the oracle's tf32 model).
//
// The same machine runs two BACKWARD chains of the many-seed program (B_MN = true: the weights are read [out, in], i.e. M/N-
// contiguous, through one 4-d TMA box per chunk; epilogues apply ReLU-derivative masks from the sign bytes instead of bias /
// ReLU):
//   policy-loss gradient through a critic:  dh1 = (dh2 W2) * 1[h1 > 0]  ->  da = dh1 W0[:, O:O+A]     (dh1 never leaves the SM)
//   policy backward:                        dh2 = (dhead Wh) * 1[h2 > 0] -> dh1 = (dh2 W2) * 1[h1 > 0]  (both stored: dW operands)
#pragma once
#include "gemm_ws.cuh"

namespace oac {

constexpr int CH_MAX_LAYERS = 3;

struct ChainParams {
    StageParams sp;               // sp.tasks: [n_layers][n_chains] GemmTask (layer-major, like the fused2 FFMA stages)
    const CUtensorMap* tmaps;     // [2 * n_layers * n_chains]: A and B map of every task (A maps of layers >= 1 unused)
    int n_layers, n_chains;
    int strips0[WS_MAX_TASKS + 1];// first strip of every chain in the per-seed item list; [n_chains] = strips per seed
    int n_seeds, total_items;
    int n_slots, slot_bytes;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ uint32_t to_tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
    return r;
}
// D[tmem] (+)= A[tmem] B[smem]:  A = 128 lanes x 8 columns (one k-step of tf32) starting at a_tmem
__device__ __forceinline__ void umma_tf32_ta(uint32_t tmem_d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
        ::"r"(tmem_d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

template <bool B_MN>
__global__ void __launch_bounds__(WS_THREADS, 1) gemm_chain_kernel(ChainParams cp) {
    extern __shared__ __align__(1024) uint8_t ws_smem[];
    __shared__ __align__(8) uint64_t s_full[WS_MAX_SLOTS], s_empty[WS_MAX_SLOTS];
    __shared__ __align__(8) uint64_t s_dfull[CH_MAX_LAYERS];     // accumulator of layer l complete (MMA lane -> epilogue)
    __shared__ __align__(8) uint64_t s_aready[2];                // D1 / D2 converted in place: A operand of the next layer
    __shared__ __align__(8) uint64_t s_d1empty, s_d2empty, s_d3empty;   // the epilogue has drained D1 / D2 / D3 (may be overwritten)
    __shared__ uint32_t s_tmem;

    const StageParams& sp = cp.sp;
    const GemmTask* __restrict__ tasks = sp.tasks;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* ring = ws_smem + ((1024u - (smem_u32(ws_smem) & 1023u)) & 1023u);
    float* slabs = reinterpret_cast<float*>(ring + (size_t)cp.n_slots * cp.slot_bytes);
    const int NL = cp.n_layers, NC = cp.n_chains;

    if (tid == 0) {
        for (int i = 0; i < cp.n_slots; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 1); }
        for (int l = 0; l < CH_MAX_LAYERS; ++l) mbar_init(&s_dfull[l], 1);
        mbar_init(&s_aready[0], WS_EPI_WARPS); mbar_init(&s_aready[1], WS_EPI_WARPS);
        mbar_init(&s_d1empty, WS_EPI_WARPS); mbar_init(&s_d2empty, WS_EPI_WARPS); mbar_init(&s_d3empty, WS_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int strips_per_seed = cp.strips0[NC];

    auto decode = [&](int g, int& seed, int& ch, int& tm) {
        const int r = g / cp.n_seeds;                 // seed is the fast index (gemm_ws.cuh)
        seed = g - r * cp.n_seeds;
        ch = 0;
        while (r >= cp.strips0[ch + 1]) ++ch;
        tm = r - cp.strips0[ch];
    };
    (void)strips_per_seed;

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            int slot = 0; uint32_t ph = 0;
            for (int g = blockIdx.x; g < cp.total_items; g += gridDim.x) {
                int seed, ch, tm;
                decode(g, seed, ch, tm);
                const int m0 = tm * WS_BM;
                for (int l = 0; l < NL; ++l) {
                    const int j = l * NC + ch;
                    const GemmTask& T = tasks[j];
                    const CUtensorMap* ta = cp.tmaps + 2 * j;
                    const CUtensorMap* tb = ta + 1;
                    const int nch = (T.K + WS_KC - 1) / WS_KC;
                    const uint32_t bytes = (l == 0 ? WS_A_BYTES : 0u) + (uint32_t)T.bn * (WS_KC * 4);
                    for (int c = 0; c < nch; ++c) {
                        mbar_wait_relaxed(&s_empty[slot], ph ^ 1u);
                        const uint32_t sa = smem_u32(ring + (size_t)slot * cp.slot_bytes), sb = sa + WS_A_BYTES;
                        const uint32_t bar = smem_u32(&s_full[slot]);
                        mbar_expect_tx(bar, bytes);
                        if (l == 0) tma_load_3d(sa, ta, c * WS_KC, m0, seed, bar);
                        if (!B_MN) tma_load_3d(sb, tb, c * WS_KC, 0, seed, bar);
                        else tma_load_4d(sb, tb, 0, c * WS_KC, 0, seed, bar);            // all 32-wide atoms of the chunk in one box
                        if (++slot == cp.n_slots) { slot = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            int slot = 0; uint32_t ph = 0;
            int it = 0;
            for (int g = blockIdx.x; g < cp.total_items; g += gridDim.x, ++it) {
                int seed, ch, tm;
                decode(g, seed, ch, tm);
                const uint32_t par = (uint32_t)it & 1u;
                for (int l = 0; l < NL; ++l) {
                    const int j = l * NC + ch;
                    const GemmTask& T = tasks[j];
                    const int bn = T.bn, K = T.K;
                    const int nch = (K + WS_KC - 1) / WS_KC;
                    // accumulator of layer l: D1 = columns [0,256), D2 = [256,512), D3 = D1's first columns
                    const uint32_t d = tmem + (l == 1 ? 256u : 0u);
                    const uint32_t a_t = tmem + (l == 2 ? 256u : 0u);          // A operand in TMEM (layers 1, 2)
                    if (l == 0 && NL == 3 && it > 0) mbar_wait_relaxed(&s_d3empty, par ^ 1u);   // D3 of the previous item drained
                    if (l == 0 && NL == 2 && it > 0) mbar_wait_relaxed(&s_d1empty, par ^ 1u);   // h1 of the previous item stored
                    if (l == 2) mbar_wait_relaxed(&s_d1empty, par);                             // h1 of this item stored: D3 may overwrite D1
                    if (l == 1) {
                        mbar_wait_relaxed(&s_aready[0], par);                                   // h1 sits in D1 as an A operand
                        if (it > 0) mbar_wait_relaxed(&s_d2empty, par ^ 1u);                    // D2 of the previous item drained
                    }
                    if (l == 2) mbar_wait_relaxed(&s_aready[1], par);
                    tc_fence_after();
                    const uint32_t idesc = umma_idesc_tf32(WS_BM, bn, false, B_MN);
                    for (int c = 0; c < nch; ++c) {
                        mbar_wait_relaxed(&s_full[slot], ph);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(ring + (size_t)slot * cp.slot_bytes), sb = sa + WS_A_BYTES;
                        const int ksteps = (min(WS_KC, K - c * WS_KC) + 7) >> 3;
                        uint64_t ad = umma_desc(sa, 16, 1024, 2);
                        uint64_t bd = B_MN ? umma_desc(sb, 4096, 512, 1) : umma_desc(sb, 16, 1024, 2);
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint32_t acc = (c > 0 || ks > 0) ? 1u : 0u;
                            if (l == 0) umma_tf32(d, ad, bd, idesc, acc);
                            else umma_tf32_ta(d, a_t + (uint32_t)(c * WS_KC + ks * 8), bd, idesc, acc);
                            ad += 2u; bd += B_MN ? 64u : 2u;
                        }
                        umma_commit(&s_empty[slot]);
                        if (++slot == cp.n_slots) { slot = 0; ph ^= 1u; }
                    }
                    umma_commit(&s_dfull[l]);
                }
            }
        }
    } else {
        // =========================== epilogue warps ===========================
        const int e = warp - 2;
        const int q = warp & 3;                                  // TMEM lane quarter this warp may access
        const int hsel = e >> 2;                                 // two warps per quarter alternate over the 64-column slabs
        float* slab = slabs + e * (32 * WS_SLAB_LD);
        const int rsub = lane >> 4, c4 = (lane & 15) << 2;
        int it = 0;
        for (int g = blockIdx.x; g < cp.total_items; g += gridDim.x, ++it) {
            int seed, ch, tm;
            decode(g, seed, ch, tm);
            const uint32_t par = (uint32_t)it & 1u;
            const int m0 = tm * WS_BM;
            for (int l = 0; l < NL; ++l) {
                const int j = l * NC + ch;
                const GemmTask& T = tasks[j];
                const int bn = T.bn, M = T.M, N = T.N, ldc = T.ldc;
                const bool relu_ = T.epi == EPI_BIAS_RELU;
                const bool to_tmem = l + 1 < NL;                 // the next layer reads this one from tensor memory
                float* __restrict__ C = resolve(sp.as, T.C, seed);
                const float* __restrict__ bias = (!B_MN && (T.epi == EPI_BIAS_RELU || T.epi == EPI_BIAS)) ? resolve(sp.as, T.bias, seed) : nullptr;
                uint8_t* __restrict__ bits_out = (!B_MN && T.ldbits > 0 && relu_) ? reinterpret_cast<uint8_t*>(resolve(sp.as, T.bits, seed)) : nullptr;
                const bool rows_live = m0 + q * 32 < M;
                // backward chains: the sign bytes of this thread's row (lane = row), 8 bytes per 32 columns, requested before
                // the accumulator is waited for: km[2 * i + hf] covers columns 64 * (hsel + 2 i) + 32 hf .. + 31
                uint2 km[4];
                const bool masked = B_MN && T.epi == EPI_MASK;
                if (masked) {
                    const int m = m0 + q * 32 + lane;
                    const uint8_t* mrow = reinterpret_cast<const uint8_t*>(resolve(sp.as, T.mask, seed)) + (long long)m * T.ldmask;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int cb = 64 * (hsel + 2 * (i >> 1)) + 32 * (i & 1);
                        km[i] = (rows_live && m < M && cb < bn) ? __ldg(reinterpret_cast<const uint2*>(mrow + (cb >> 2))) : make_uint2(0u, 0u);
                    }
                }
                auto masked_val = [&](float x, const uint2& k, int i) {      // column i (0..31) of the group: byte i >> 2, bit i & 3
                    const uint32_t w = (i < 16) ? k.x : k.y;
                    return ((w >> (((i & 15) >> 2) * 8 + (i & 3))) & 1u) ? x : 0.f;
                };
                mbar_wait_relaxed(&s_dfull[l], par);
                tc_fence_after();
                const uint32_t t_base = tmem + ((uint32_t)(q * 32) << 16) + (l == 1 ? 256u : 0u);
                if (to_tmem) {
                    // ---- phase A: convert the accumulator in place (bias, ReLU, tf32 rounding): the next layer's MMAs can start ----
                    auto convert_slab = [&](int si) {                    // this warp's slabs: hsel, hsel + 2
                        const int sl = hsel + 2 * si;
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const int cb = sl * WS_SLAB + 32 * hf;
                            float v[32];
                            tmem_ld16_nowait(t_base + (uint32_t)cb, &v[0]);
                            tmem_ld16_nowait(t_base + (uint32_t)(cb + 16), &v[16]);
                            tmem_ld_wait();
                            uint32_t r[32];
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                float x = v[i];
                                if (!B_MN) {
                                    x += __ldg(bias + cb + i);                            // uniform address: one broadcast load
                                    if (relu_) x = relu(x);
                                } else if (masked) x = masked_val(x, km[2 * si + hf], i);
                                r[i] = to_tf32_rna(x);
                            }
                            tmem_st16(t_base + (uint32_t)cb, &r[0]);
                            tmem_st16(t_base + (uint32_t)(cb + 16), &r[16]);
                        }
                    };
                    if (B_MN) {                                          // (static slab index: the mask registers are not addressable)
                        if (hsel * WS_SLAB < bn) convert_slab(0);
                        if ((hsel + 2) * WS_SLAB < bn) convert_slab(1);
                    } else {
                        for (int si = 0; (hsel + 2 * si) * WS_SLAB < bn; ++si) convert_slab(si);
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_aready[l]);        // l = 0 -> D1 is an A operand, l = 1 -> D2 is
                }
                // ---- phase B: accumulator (converted: the stored activation is the tf32-rounded one every consumer would round it
                // to anyway -- TFLOAT32 tensor maps, sign tests) -> slab -> coalesced global store; overlaps the next layer's MMAs ----
                auto store_slab = [&](int si) {
                    const int sl = hsel + 2 * si;
                    const int c0 = sl * WS_SLAB;
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const int cb = c0 + 32 * hf;
                        if (cb >= bn) break;
                        float v[32];
                        tmem_ld16_nowait(t_base + (uint32_t)cb, &v[0]);
                        if (cb + 16 < bn) tmem_ld16_nowait(t_base + (uint32_t)(cb + 16), &v[16]);
                        tmem_ld_wait();
                        if (!to_tmem) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                const int n = cb + i;
                                float x = v[i];
                                if (!B_MN) {
                                    x += (n < N) ? __ldg(bias + n) : 0.f;
                                    if (relu_) x = relu(x);
                                } else if (masked) x = masked_val(x, km[2 * si + hf], i);
                                v[i] = x;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            *reinterpret_cast<float4*>(slab + lane * WS_SLAB_LD + 32 * hf + 4 * i) =
                                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    }
                    __syncwarp();
                    const int n = c0 + c4;
                    if (n < N && rows_live) {
                        const bool vec = n + 3 < N;
#pragma unroll
                        for (int rp = 0; rp < 16; ++rp) {
                            const int row = 2 * rp + rsub, m = m0 + q * 32 + row;
                            if (m >= M) continue;
                            const float4 x = *reinterpret_cast<const float4*>(slab + row * WS_SLAB_LD + c4);
                            float* dst = C + (long long)m * ldc + n;
                            // sign bits of these four columns (one byte) for the masked dX epilogues (gemm_ws.cuh): written here,
                            // off the path the next layer's MMAs wait on
                            if (bits_out != nullptr && vec)
                                bits_out[(long long)m * T.ldbits + (n >> 2)] =
                                    (uint8_t)((x.x > 0.f ? 1u : 0u) | (x.y > 0.f ? 2u : 0u) | (x.z > 0.f ? 4u : 0u) | (x.w > 0.f ? 8u : 0u));
                            if (vec) *reinterpret_cast<float4*>(dst) = x;
                            else {
                                dst[0] = x.x;
                                if (n + 1 < N) dst[1] = x.y;
                                if (n + 2 < N) dst[2] = x.z;
                            }
                        }
                    }
                    __syncwarp();
                };
                if (!T.no_store) {
                    if (B_MN) {
                        if (hsel * WS_SLAB < bn) store_slab(0);
                        if ((hsel + 2) * WS_SLAB < bn) store_slab(1);
                    } else {
                        for (int si = 0; (hsel + 2 * si) * WS_SLAB < bn; ++si) store_slab(si);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (l == 0) mbar_arrive(&s_d1empty);         // h1 is in global memory: D1 may be overwritten
                    if (l == 1) mbar_arrive(&s_d2empty);         // D2 has been read (a following head layer reads it by MMA, in order)
                    if (l == 2) mbar_arrive(&s_d3empty);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory");
    }
}

}  // namespace oac
