// Warp-specialised, persistent tcgen05 (kind::tf32) grouped-GEMM stage: the throughput path of the
// batched-seed configuration (BASELINE config 5: 64 independent OAC seeds as grouped GEMMs).
//
// Same task table and fused epilogues as gemm_simt.cuh / gemm_tc.cuh, different machine mapping:
//   * one persistent CTA per SM walks a static list of (seed, task, tile) work items;
//   * warp 0 (one lane) is the TMA producer: every operand tile is fetched with
//     cp.async.bulk.tensor.3d through a per-task tensor map ([seed][row][col] view of the arena), 32 k per
//     pipeline slot, straight into the canonical UMMA layouts -- SWIZZLE_128B for K-contiguous operands,
//     SWIZZLE_128B_ATOM_32B for the M/N-contiguous ones (dX and dW products), so transposed copies never
//     exist (an M/N-contiguous operand is described as [seed][32-wide atom][k][32]: ONE 4-d box fetches all atoms of a
//     tile's chunk instead of one box instruction per atom; ragged last atoms read the rest of the row pitch: ws_plan).
//     The maps use the TFLOAT32 element type: the TMA unit rounds fp32 -> tf32 to nearest while it copies (measured:
//     tools/probes/tma_probe.cu; as fast as plain fp32 maps), which removes the MMA's truncation bias without any
//     rounding pass over shared memory or rounded copies of the weights.  Ragged M / K edges are zero-filled by the TMA
//     bounds check.
//   * warp 1 (one lane) issues tcgen05.mma into one of TWO TMEM accumulators (2 x 256 columns) and
//     signals slot reuse / accumulator completion with tcgen05.commit -> mbarrier;
//   * warps 2..9 are the epilogue: tcgen05.ld (lane = row) -> per-warp shared slab -> float4 accesses with
//     consecutive lanes on consecutive columns, so bias/ReLU, the ReLU mask and the Adam (+Polyak) update
//     of the weight block (param, two moments, target: 32 B per element) are fully coalesced and keep
//     ~64 KB of loads in flight per SM.  The epilogue of tile i overlaps the TMA + MMA of tile i+1.
//   * the bias gradient of a dW task (column sums of dY) is one extra N=32 MMA per k-step against a
//     constant all-ones tile into spare TMEM columns.
#pragma once
#include <cuda.h>
#include "gemm_tc.cuh"

namespace oac {

constexpr int WS_BM = 128;
constexpr int WS_KC = 32;                        // k per pipeline slot = one 128-byte swizzle row
constexpr int WS_EPI_WARPS = 8;
constexpr int WS_THREADS = (2 + WS_EPI_WARPS) * 32;
// Register budget: 10 warps = 3 warps on two of the four SM sub-partitions (16384 registers each), so a thread gets at most
// 16384 / 96 = 168 registers (a kernel compiled for more fails to launch): the epilogues below are written to stay under it.
constexpr int WS_SLAB = 64;                      // columns per epilogue slab
constexpr int WS_SLAB_LD = 68;                   // floats; 16-byte aligned rows, conflict-free v4 stores (lane = row)
constexpr int WS_MAX_TASKS = 64;
constexpr int WS_MAX_SLOTS = 8;
constexpr int WS_BIAS_COL = 224;                 // TMEM column (inside an accumulator buffer) of the bias-gradient MMA
constexpr int WS_BN_MAX_BIAS = 224;              // tile width limit for tasks that carry the bias MMA
constexpr uint32_t WS_A_BYTES = WS_BM * WS_KC * 4;
constexpr uint32_t WS_ONES_BYTES = 32 * WS_KC * 4;
constexpr uint32_t WS_SLAB_BYTES = WS_EPI_WARPS * 32 * WS_SLAB_LD * 4;

struct WsParams {
    StageParams sp;
    const CUtensorMap* tmaps;     // device, [2 * n_tasks]: A and B operand of every task
    int n_tasks;
    int tiles_per_seed;
    int n_seeds;
    int total_tiles;              // tiles_per_seed * n_seeds
    int n_slots;                  // operand ring depth
    int slot_bytes;               // WS_A_BYTES + max_bn * 128
};

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

__device__ __forceinline__ AdamScalars make_adam_scalars_fast(const AdamHyper& h, float lr, int t, int train_steps_done) {
    AdamScalars s;
    const double b1 = rint((double)h.beta1 * 1e6) * 1e-6, b2 = rint((double)h.beta2 * 1e6) * 1e-6;
    s.step_size = (float)((double)lr / (1.0 - ipow(b1, t)));
    s.bc2_sqrt = (float)sqrt(1.0 - ipow(b2, t));
    s.beta1 = (float)b1; s.beta2 = (float)b2;
    s.one_m_beta1 = (float)(1.0 - b1); s.one_m_beta2 = (float)(1.0 - b2);
    s.eps = h.eps; s.tau = h.tau; s.one_m_tau = h.one_minus_tau;
    s.do_polyak = ((train_steps_done - 1) % (h.target_period > 0 ? h.target_period : 1)) == 0;
    return s;
}
// Adam (+Polyak) on register operands for the TF32 path.  The weight gradient g already carries ~1e-3 of tf32
// rounding, so the IEEE-exact divide / square root of adam_update (gemm_simt.cuh: ~40 dependent instructions per
// element, 40 % of this kernel's instruction stream in a mid-round ncu capture) buy nothing here: MUFU sqrt / reciprocal
// (2 ulp) and fused multiply-adds instead.  inv_bc2 = 1 / sqrt(1 - beta2^t).
__device__ __forceinline__ void adam_core(float g, float& p, float& m, float& v, float& tgt, bool has_tgt, const AdamScalars& s,
                                          float inv_bc2) {
    m = fmaf(m, s.beta1, s.one_m_beta1 * g);
    v = fmaf(v, s.beta2, s.one_m_beta2 * g * g);
    float sq, rc;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(v));
    const float denom = fmaf(sq, inv_bc2, s.eps);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(denom));
    p = fmaf(-s.step_size * m, rc, p);
    if (has_tgt) tgt = fmaf(tgt, s.one_m_tau, p * s.tau);
}

// Sign bits of a hidden activation ("mask bits"): the forward epilogues (here and in gemm_chain.cuh) also store, for every
// row and every group of 4 columns, one BYTE whose bit j says h[row][4 g + j] > 0 (what a lane of the coalesced epilogue
// domain holds: a float4).  The masked dX epilogues and the rank-1 pass read these bytes (64 B per 256-wide row) instead
// of the fp32 activation (1 KB): 16 registers of loads in flight per lane instead of 64, and 1/16 of the mask traffic.
// Epilogue of ONE accumulator tile (rows m0 + 32 q .. + 31 of this warp's TMEM lane quarter, the 64-column slabs sl = hsel,
// hsel + 2, ..): tcgen05.ld -> per-warp slab -> coalesced float4 rows -> bias / ReLU (+ mask bits) | ReLU-derivative mask |
// gradient store | Adam (+ Polyak).  Shared by gemm_ws_kernel and its CTA-pair variant (gemm_ws2.cuh).
// VARIANT: (K, MN) stages -- 1: the masks are sign bits; (MN, MN) stages -- 1: the fused Adam (+ Polyak) epilogue is compiled in
// (few-seed programs; the many-seed program stores gradients, and keeping the Adam path's ~90 registers out of that
// instantiation keeps its store loop free of spills).
template <bool A_MN, bool B_MN, bool VARIANT = false>
__device__ __forceinline__ void ws_tile_epilogue(const StageParams& sp, const GemmTask& T, int seed, int m0, int n0, int tn,
                                                 uint32_t t_base, float* slab, int q, int hsel, int lane) {
    float* __restrict__ m1 = sp.as.base[AR_ADAM_M];
    float* __restrict__ m2 = sp.as.base[AR_ADAM_V];
    float* __restrict__ pb0 = sp.as.base[AR_PARAM];
    const int rsub = lane >> 4, c4 = (lane & 15) << 2;
    const int bn = T.bn, M = T.M, N = T.N, epi = T.epi, ldc = T.ldc;
    const int nlim = min(N, n0 + bn);
    float* __restrict__ C = resolve(sp.as, T.C, seed);
    // the operand layouts pin the epilogue class (dW products are the only (MN, MN) tasks, masked dX products
    // the only (K, MN) ones): dead epilogues are compiled out, which keeps their registers out of the live set
    constexpr bool CAN_ADAM = A_MN && B_MN && VARIANT, CAN_GRAD = A_MN && B_MN, CAN_MASK = !A_MN && B_MN, CAN_BITS = !A_MN && !B_MN;
    constexpr bool MASK_BITS = VARIANT;
    const bool is_adam = CAN_ADAM && epi == EPI_ADAM;
    const bool is_grad = CAN_GRAD && epi == EPI_GRAD;        // plain store of dW; Adam streams later (adam_stream.cuh)
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
    constexpr bool mask_bits = CAN_MASK && MASK_BITS;     // a stage's tasks all carry the same kind of mask (ws_plan)
    const int ldmask = T.ldmask;
    // forward layer whose sign bits a later stage wants
    uint8_t* __restrict__ bits_out = (CAN_BITS && T.ldbits > 0 && epi == EPI_BIAS_RELU)
                                         ? reinterpret_cast<uint8_t*>(resolve(sp.as, T.bits, seed)) : nullptr;
    const bool rows_live = m0 + q * 32 < M;              // warp-uniform: nothing to write for this quarter
    for (int sl = hsel; sl * WS_SLAB < nlim - n0 && rows_live; sl += 2) {
        const int c0 = sl * WS_SLAB;
        const int n = n0 + c0 + c4;
        // ReLU-mask epilogue.  One mask load at a time costs a DRAM latency each (21 us per K=1 tile, round 1), so the
        // loads are requested before the accumulator read-back.  Mask BITS: 16 words per lane (8 lanes share a word).
        // fp32 masks (activations of a stage that stores no bits): all 16 float4 of a lane together with the 32
        // accumulator registers and the prefetched slab rows exceed the 168-register budget -- ptxas spilled into the
        // store loop and this instantiation spent 12 us per 128 x 256 tile in its epilogue whatever K (5 us in the
        // forward kernel; ncu source page: stores waiting on local-memory reloads) -- hence two batches of 8 row pairs,
        // each reduced to sign bits as soon as it is consumed: batch A flies during the read-back, batch B during A's stores.
        uint32_t kw[mask_bits ? 16 : 1];
        float4 kq[mask_bits ? 1 : 8];
        const bool mask_vec = !mask_bits && mask != nullptr && (n + 3 < nlim);
        auto load_masks = [&](int rp0) {
            if constexpr (!mask_bits) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int m = m0 + q * 32 + 2 * (rp0 + r) + rsub;
                    kq[r] = (m < M) ? __ldg(reinterpret_cast<const float4*>(mask + (long long)m * ldmask + n))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        };
        auto pack_masks = [&]() {
            uint32_t b = 0;
            if constexpr (!mask_bits) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    b |= ((kq[r].x > 0.f ? 1u : 0u) | (kq[r].y > 0.f ? 2u : 0u) | (kq[r].z > 0.f ? 4u : 0u) | (kq[r].w > 0.f ? 8u : 0u)) << (4 * r);
            }
            return b;
        };
        if (mask_bits) {
            if (mask != nullptr) {
                const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(mask) + (n >> 2);
#pragma unroll
                for (int rp = 0; rp < 16; ++rp) {
                    const int m = m0 + q * 32 + 2 * rp + rsub;
                    kw[mask_bits ? rp : 0] = (m < M && n < nlim) ? (uint32_t)__ldg(wsrc + (long long)m * ldmask) : 0u;
                }
            }
        } else if (mask_vec) load_masks(0);
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
        const bool vec = n + 3 < nlim;
        if (is_adam) {
            if (n < nlim) {
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
            }
        } else {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias != nullptr && n < nlim) {
                if (vec) b4 = __ldg(reinterpret_cast<const float4*>(bias + n));
                else { b4.x = __ldg(bias + n); if (n + 1 < nlim) b4.y = __ldg(bias + n + 1); if (n + 2 < nlim) b4.z = __ldg(bias + n + 2); }
            }
            uint32_t mbits = 0xffffffffu;
            if (mask_vec) {
                mbits = pack_masks();
                asm volatile("" ::: "memory");           // batch B is requested only now: its registers replace batch A's
                load_masks(8);
            }
            auto store_rows = [&](int rp0) {
                if (n >= nlim) return;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int row = 2 * (rp0 + r) + rsub, m = m0 + q * 32 + row;
                    if (m >= M) continue;
                    float4 x = *reinterpret_cast<const float4*>(slab + row * WS_SLAB_LD + c4);
                    x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
                    if (epi == EPI_BIAS_RELU) { x.x = relu(x.x); x.y = relu(x.y); x.z = relu(x.z); x.w = relu(x.w); }
                    float* dst = C + (long long)m * ldc + n;
                    if (vec) {
                        if (CAN_MASK && mask != nullptr) {
                            const uint32_t nb = mask_bits ? kw[mask_bits ? rp0 + r : 0] : mbits >> (4 * r);
                            x.x = (nb & 1u) ? x.x : 0.f; x.y = (nb & 2u) ? x.y : 0.f;
                            x.z = (nb & 4u) ? x.z : 0.f; x.w = (nb & 8u) ? x.w : 0.f;
                        }
                        *reinterpret_cast<float4*>(dst) = x;
                        if (CAN_BITS && bits_out != nullptr)
                            bits_out[(long long)m * T.ldbits + (n >> 2)] =
                                (uint8_t)((x.x > 0.f ? 1u : 0u) | (x.y > 0.f ? 2u : 0u) | (x.z > 0.f ? 4u : 0u) | (x.w > 0.f ? 8u : 0u));
                    } else {
                        uint32_t nib = 0;
#pragma unroll
                        for (int jj = 0; jj < 3; ++jj) {
                            if (n + jj >= nlim) continue;
                            float y = jj == 0 ? x.x : (jj == 1 ? x.y : x.z);
                            if (CAN_MASK && mask != nullptr) {
                                const bool on = mask_bits ? ((kw[mask_bits ? rp0 + r : 0] >> jj) & 1u) != 0u
                                                          : __ldg(mask + (long long)m * ldmask + n + jj) > 0.f;
                                y = on ? y : 0.f;
                            }
                            nib |= (y > 0.f ? 1u : 0u) << jj;
                            dst[jj] = y;
                        }
                        if (CAN_BITS && bits_out != nullptr) bits_out[(long long)m * T.ldbits + (n >> 2)] = (uint8_t)nib;
                    }
                }
            };
            store_rows(0);
            if (mask_vec) mbits = pack_masks();
            store_rows(8);
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
}

template <bool A_MN, bool B_MN, bool VARIANT = false>
__global__ void __launch_bounds__(WS_THREADS, 1) gemm_ws_kernel(WsParams wp) {
    extern __shared__ __align__(1024) uint8_t ws_smem[];
    __shared__ __align__(8) uint64_t s_full[WS_MAX_SLOTS], s_empty[WS_MAX_SLOTS], s_tfull[2], s_tempty[2];
    __shared__ uint32_t s_tmem;
    __shared__ int s_tile0[WS_MAX_TASKS + 1];

    pdl_wait();
    const StageParams& sp = wp.sp;
    const GemmTask* __restrict__ tasks = sp.tasks;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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
        mbar_init(&s_tempty[0], WS_EPI_WARPS); mbar_init(&s_tempty[1], WS_EPI_WARPS);
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
        // =========================== TMA producer ===========================
        int slot = 0; uint32_t ph = 0;
        for (int g = blockIdx.x; g < wp.total_tiles; g += gridDim.x) {
            int seed, j, tm, tn;
            decode(g, seed, j, tm, tn);
            const GemmTask& T = tasks[j];
            const int bn = T.bn;
            const int m0 = tm * WS_BM, n0 = tn * bn;
            if (lane == 0) {
                const CUtensorMap* ta = wp.tmaps + 2 * j;
                const CUtensorMap* tb = ta + 1;
                const int nch = (T.K + WS_KC - 1) / WS_KC;
                const uint32_t bytes = WS_A_BYTES + (uint32_t)bn * (WS_KC * 4);
                for (int c = 0; c < nch; ++c) {
                    mbar_wait_relaxed(&s_empty[slot], ph ^ 1u);
                    const uint32_t sa = smem_u32(ring + (size_t)slot * wp.slot_bytes), sb = sa + WS_A_BYTES;
                    const uint32_t bar = smem_u32(&s_full[slot]);
                    mbar_expect_tx(bar, bytes);
                    if (!A_MN) tma_load_3d(sa, ta, c * WS_KC, m0, seed, bar);
                    else tma_load_4d(sa, ta, 0, c * WS_KC, m0 >> 5, seed, bar);          // all 32-wide atoms of the tile in one box
                    if (!B_MN) tma_load_3d(sb, tb, c * WS_KC, n0, seed, bar);
                    else tma_load_4d(sb, tb, 0, c * WS_KC, n0 >> 5, seed, bar);
                    if (++slot == wp.n_slots) { slot = 0; ph ^= 1u; }
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            int slot = 0; uint32_t ph = 0;
            int tl = 0;
            const uint64_t ones_desc = umma_desc(smem_u32(ones), 4096, 512, 1);
            const uint32_t idesc_bias = umma_idesc_tf32(WS_BM, 32, true, true);
            for (int g = blockIdx.x; g < wp.total_tiles; g += gridDim.x, ++tl) {
                int seed, j, tm, tn;
                decode(g, seed, j, tm, tn);
                const GemmTask& T = tasks[j];
                const int bn = T.bn, K = T.K;
                const int nch = (K + WS_KC - 1) / WS_KC;
                const int buf = tl & 1;
                mbar_wait_relaxed(&s_tempty[buf], (((uint32_t)tl >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_main = tmem + (uint32_t)(buf * 256), d_bias = d_main + WS_BIAS_COL;
                const uint32_t idesc = umma_idesc_tf32(WS_BM, bn, A_MN, B_MN);
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
                        umma_tf32(d_main, ad, bd, idesc, acc);
                        if (bias_mma) umma_tf32(d_bias, ad, ones_desc, idesc_bias, acc);
                        ad += A_MN ? 64u : 2u;
                        bd += B_MN ? 64u : 2u;
                    }
                    umma_commit(&s_empty[slot]);                 // slot reusable once these MMAs have read it
                    if (++slot == wp.n_slots) { slot = 0; ph ^= 1u; }
                }
                umma_commit(&s_tfull[buf]);                      // accumulator complete
            }
        }
    } else {
        // =========================== epilogue warps ===========================
        const int e = warp - 2;
        const int q = warp & 3;                                  // TMEM lane quarter this warp may read
        const int hsel = e >> 2;                                 // two warps per quarter alternate over the slabs
        float* slab = slabs + e * (32 * WS_SLAB_LD);
        int tl = 0;
        for (int g = blockIdx.x; g < wp.total_tiles; g += gridDim.x, ++tl) {
            int seed, j, tm, tn;
            decode(g, seed, j, tm, tn);
            const GemmTask& T = tasks[j];
            const int buf = tl & 1;
            mbar_wait_relaxed(&s_tfull[buf], ((uint32_t)tl >> 1) & 1u);
            tc_fence_after();
            const uint32_t t_base = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256);
            ws_tile_epilogue<A_MN, B_MN, VARIANT>(sp, T, seed, tm * WS_BM, tn * T.bn, tn, t_base, slab, q, hsel, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_tempty[buf]);
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
