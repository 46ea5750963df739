// tcgen05 (5th-gen tensor core) batched GEMM stage, kind::tf32, fp32 accumulation in TMEM.
//
// Same task table and epilogues as gemm_simt.cuh (one launch = one stage of grouped GEMMs over
// tasks x seeds), but the contraction runs on the tensor pipe:
//   * CTA tile 128 x BN (BN = 16..256 chosen per stage), UMMA M=128, N=BN, K=8 (tf32);
//   * operands are staged global -> shared with cp.async straight into the canonical
//     SWIZZLE_128B UMMA layouts -- K-major for K-contiguous operands, MN-major for the
//     M/N-contiguous ones (dX and dW products), so no transposed copies exist anywhere;
//   * K is consumed in chunks of 128 through a 2-deep shared-memory ring: the elected thread
//     issues the chunk's tcgen05.mma and commits to the ring slot's mbarrier while all threads
//     already stage the next chunk;
//   * the accumulator lives in TMEM (128 lanes x BN columns) and is read back with tcgen05.ld
//     by the 8 epilogue warps (lane quarter = warp % 4, column half = warp / 4), which apply
//     bias/ReLU, the ReLU mask, or Adam (+Polyak) directly on the weight block;
//   * the bias gradient of a dW task comes for free from an extra all-ones B row (column N of
//     the accumulator = sum_k A(m,k)).
// TF32 keeps 10 mantissa bits of each operand: results agree with the fp32 path to ~1e-3
// relative (BASELINE.json north_star tolerance for the tensor-core path).
#pragma once
#include "gemm_simt.cuh"

namespace oac {

constexpr int TC_THREADS = 256;
constexpr int TC_BM = 128;
// k elements per ring slot are a stage parameter (kc: 32 / 64 / 128, a multiple of the 32-element atom)

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) (2 = SWIZZLE_128B)
// layout_type: 2 = SWIZZLE_128B (K-major operands), 1 = SWIZZLE_128B_BASE32B (the only layout
// tcgen05 accepts for MN-major 32-bit operands: 32-byte units XOR-ed with the k-row, 4-row atoms)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout_type << 61;
    return d;
}
// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) [4,6), a/b_format TF32 (2)
// [7,10)/[10,13), a_major [15], b_major [16] (1 = MN-major), N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= 2u << 7;
    d |= 2u << 10;
    d |= (a_mn ? 1u : 0u) << 15;
    d |= (b_mn ? 1u : 0u) << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                 ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- operand staging into the canonical SWIZZLE_128B layouts -----------------------------------
// K-major tile: [k-atom (32 elems)][R rows][128 B]; 16B unit j of row r sits at j ^ (r & 7).
// Element (row, k) of the global operand is src[row*ld + k].
__device__ __forceinline__ void stage_kmajor(uint8_t* sm, const float* __restrict__ src, int ld, int row0, int R,
                                             int row_max, int k0, int kn8, int k_max, bool vec_ok) {
    const int tid = threadIdx.x;
    if (vec_ok) {
        // lane = 16-byte unit of the 128-float chunk row (coalesced 512 B per warp), warp w takes rows w, w+8, ...
        // -> the swizzle term (u&7)^(r&7) is loop-invariant and both addresses advance by constants.
        const int warp = tid >> 5, u = tid & 31;
        if (u * 4 < kn8) {
            const int gk = k0 + u * 4;
            const int kbytes = gk < k_max ? min(16, (k_max - gk) * 4) : 0;
            uint8_t* dst = sm + (u >> 3) * (R * 128) + warp * 128 + (((u & 7) ^ (warp & 7)) << 4);
            const float* p = src + (long long)(row0 + warp) * ld + gk;
            for (int r = warp; r < R; r += 8, dst += 1024, p += 8 * (long long)ld) {
                const bool ok = (row0 + r < row_max) && kbytes > 0;
                cp_async16_zfill(reinterpret_cast<float*>(dst), ok ? p : src, ok ? kbytes : 0);
            }
        }
    } else {
        const int total = R * kn8;
        for (int c = tid; c < total; c += TC_THREADS) {
            const int r = c / kn8, k = c - r * kn8;
            const int gr = row0 + r, gk = k0 + k;
            const int u = k >> 2;
            float* dst = reinterpret_cast<float*>(sm + (u >> 3) * (R * 128) + r * 128 + (((u & 7) ^ (r & 7)) << 4)) + (k & 3);
            if (gr < row_max && gk < k_max) cp_async4(dst, src + (long long)gr * ld + gk);
            else *dst = 0.f;
        }
    }
}
// MN-major tile (SWIZZLE_128B_BASE32B): [k-group (4 rows)][mn-atom (32 elems)][4 k-rows][128 B]; the 32-byte
// unit j of k-row kk sits at j ^ (kk & 3).  Element (mn, k) of the global operand is src[k*ld + mn].
__device__ __forceinline__ uint32_t mn_offset(int k, int m, int atoms) {
    const int u = (m & 31) >> 2;                        // 16-byte unit inside the 128-byte row
    return (uint32_t)(((k >> 2) * atoms + (m >> 5)) * 512 + (k & 3) * 128 + ((((u >> 1) ^ (k & 3)) << 5) | ((u & 1) << 4)) +
                      (m & 3) * 4);
}
__device__ __forceinline__ void stage_mnmajor(uint8_t* sm, const float* __restrict__ src, int ld, int mn0, int R,
                                              int mn_max, int k0, int kn8, int k_max, bool vec_ok) {
    const int tid = threadIdx.x;
    const int atoms = R >> 5;
    if (vec_ok) {
        // a k-row holds R/4 16-byte units (R = 32..128 -> 8..32 units): lanes cover units first, then k sub-rows;
        // a thread's k advances by a multiple of 4, so (k & 3) and the swizzle term stay fixed.
        const int upr = R >> 2;                         // power of two (R is 32, 64 or 128)
        const int u = tid & (upr - 1);
        const int kstart = tid / upr, kstep = TC_THREADS / upr;      // kstep = 8, 16 or 32
        const int gm = mn0 + u * 4;
        const int mbytes = gm < mn_max ? min(16, (mn_max - gm) * 4) : 0;
        const float* p = src + (long long)(k0 + kstart) * ld + gm;
        for (int k = kstart; k < kn8; k += kstep, p += (long long)kstep * ld) {
            const bool ok = (k0 + k < k_max) && mbytes > 0;
            cp_async16_zfill(reinterpret_cast<float*>(sm + mn_offset(k, u * 4, atoms)), ok ? p : src, ok ? mbytes : 0);
        }
    } else {
        const int total = kn8 * R;
        for (int c = tid; c < total; c += TC_THREADS) {
            const int k = c / R, m = c - k * R;
            const int gk = k0 + k, gm = mn0 + m;
            float* dst = reinterpret_cast<float*>(sm + mn_offset(k, m, atoms));
            if (gk < k_max && gm < mn_max) cp_async4(dst, src + (long long)gk * ld + gm);
            else *dst = 0.f;
        }
    }
}
// all-ones B row at n = n_local (dW tasks: accumulator column = sum_k A(m,k) = bias gradient); MN-major B only
__device__ __forceinline__ void stage_ones_row_mn(uint8_t* sm, int R, int n_local, int k0, int kn8, int k_max) {
    const int atoms = R >> 5;
    for (int k = threadIdx.x; k < kn8; k += TC_THREADS)
        *reinterpret_cast<float*>(sm + mn_offset(k, n_local, atoms)) = (k0 + k < k_max) ? 1.0f : 0.f;
}

struct TcStageParams {
    StageParams sp;
    int bn;             // tile N (multiple of 16 for K-major B, of 32 for MN-major B)
    int kc;             // k elements per ring slot (32 / 64 / 128)
    int tmem_cols;      // power of two >= max(32, n_acc * bn)
    int n_main;         // X3: number of main accumulators the K chunks rotate over (+1 for the corrections)
    long long* dbg;     // optional [gridDim.x][8] clock64 phase stamps of task 0 (measurement aid)
};

// X3: 3xTF32 split (a = a_hi + a_lo with both parts exact in tf32; D += a_hi b_hi + a_hi b_lo + a_lo b_hi):
// fp32-grade accuracy on the tensor pipe (dropped term a_lo b_lo ~ 2^-22) at three MMAs per k-step.
template <bool A_MN, bool B_MN, bool X3>
__global__ void __launch_bounds__(TC_THREADS, 2) gemm_tc_kernel(TcStageParams tp) {
    extern __shared__ __align__(1024) uint8_t tc_smem[];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_tmem;
    __shared__ AdamScalars s_adam;

    pdl_wait();
    const StageParams& sp = tp.sp;
    const GemmTask& T = sp.tasks[blockIdx.y];
    const int tile = blockIdx.x;
    if (tile >= T.tiles_m * T.tiles_n) return;
    long long* dbg = (tp.dbg && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) ? tp.dbg + 8 * tile : nullptr;
    if (dbg) dbg[0] = clock64();
    const int seed = blockIdx.z;
    const int BN = tp.bn;
    const int tm = tile / T.tiles_n, tn = tile - tm * T.tiles_n;
    const int m0 = tm * TC_BM, n0 = tn * BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const float* __restrict__ A = resolve(sp.as, T.A, seed);
    const float* __restrict__ B = resolve(sp.as, T.B, seed);
    const int M = T.M, N = T.N, K = T.K;
    const bool a_vec = ((T.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool b_vec = ((T.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    const bool is_adam = T.epi == EPI_ADAM;
    // bias gradient through the all-ones B row: it sits at global column N
    const bool ones_here = B_MN && is_adam && T.has_bias && (N >= n0) && (N < n0 + BN);

    // ---- one-time setup: TMEM, mbarriers ----
    const int KC = tp.kc;
    const uint32_t a_bytes = TC_BM * KC * 4, b_bytes = (uint32_t)BN * KC * 4;
    // SWIZZLE_128B atoms need 1024-byte alignment in the shared window: align by hand (host adds the slack)
    uint8_t* ring = tc_smem + ((1024u - (smem_u32(tc_smem) & 1023u)) & 1023u);   // slot: A | B [| A_lo | B_lo]
    const uint32_t slot_bytes = (a_bytes + b_bytes) * (X3 ? 2u : 1u);
    const int nchunks = (K + KC - 1) / KC;
    auto stage_chunk = [&](int c) {
        const int slot = c & 1;
        const int k0 = c * KC;
        const int kn8 = (min(KC, K - k0) + 7) & ~7;
        uint8_t* sa = ring + slot * slot_bytes;
        uint8_t* sb = sa + a_bytes;
        if (c >= 2) mbar_wait(&s_bar[slot], (uint32_t)((c >> 1) - 1) & 1u);   // MMAs of chunk c-2 are done with this slot
        if (!A_MN) stage_kmajor(sa, A, T.lda, m0, TC_BM, M, k0, kn8, K, a_vec);
        else       stage_mnmajor(sa, A, T.lda, m0, TC_BM, M, k0, kn8, K, a_vec);
        if (!B_MN) stage_kmajor(sb, B, T.ldb, n0, BN, N, k0, kn8, K, b_vec);
        else       stage_mnmajor(sb, B, T.ldb, n0, BN, N, k0, kn8, K, b_vec);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    stage_chunk(0);           // the first loads fly while TMEM is allocated and the barriers are set up
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                     ::"r"(smem_u32(&s_tmem)), "r"((uint32_t)tp.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (is_adam && tid == 64) {
        int t = sp.as.counters[seed * sp.as.n_counters + T.counter];
        int ts = sp.as.counters[seed * sp.as.n_counters + CNT_TRAIN_STEPS];
        s_adam = make_adam_scalars(sp.hyper, T.lr, t, ts);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = s_tmem;
    if (dbg) dbg[1] = clock64();
    const uint32_t idesc = umma_idesc_tf32(TC_BM, BN, A_MN, B_MN);

    // ---- main loop over K chunks: chunk c+1 is in flight (cp.async group) while chunk c is rounded and multiplied ----
    for (int c = 0; c < nchunks; ++c) {
        const int slot = c & 1;
        const int k0 = c * KC;
        const int kn = min(KC, K - k0);
        const int kn8 = (kn + 7) & ~7;
        uint8_t* sa = ring + slot * slot_bytes;
        uint8_t* sb = sa + a_bytes;
        if (c + 1 < nchunks) {
            stage_chunk(c + 1);
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");     // chunk c has landed (this thread's part)
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        __syncthreads();                          // everybody's cp.async / zero-fill of chunk c has landed
        if (dbg && c == 0) dbg[2] = clock64();
        if (ones_here) {
            stage_ones_row_mn(sb, BN, N - n0, k0, kn8, K);
            __syncthreads();
        }
        {
            // one pass over the STAGED part of the slot (generic proxy, before the async-proxy fence):
            //  X3 : hi = top 19 bits (exactly what kind::tf32 keeps), lo = x - hi (exact), stored behind the slot
            //  TF32: round to nearest tf32 in place -- the MMA itself truncates, which biases every product
            //        towards zero (measured 7.7e-4 relative, independent of K); rounding makes the error zero-mean
            const int a_used = A_MN ? (kn8 >> 2) * (TC_BM / 32) * 512 : ((kn8 + 31) >> 5) * (TC_BM * 128);
            const int b_used = B_MN ? (kn8 >> 2) * (BN / 32) * 512 : ((kn8 + 31) >> 5) * (BN * 128);
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                float4* raw = reinterpret_cast<float4*>(part == 0 ? sa : sb);
                float4* lo = reinterpret_cast<float4*>((part == 0 ? sa : sb) + a_bytes + b_bytes);
                const int n4 = (part == 0 ? a_used : b_used) >> 4;
                for (int i = tid; i < n4; i += TC_THREADS) {
                    float4 x = raw[i], h;
                    if (X3) {
                        float4 l;
                        h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
                        h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
                        h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
                        h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
                        lo[i] = l;
                    } else {
                        h.x = __uint_as_float((__float_as_uint(x.x) + 0x1000u) & 0xFFFFE000u);
                        h.y = __uint_as_float((__float_as_uint(x.y) + 0x1000u) & 0xFFFFE000u);
                        h.z = __uint_as_float((__float_as_uint(x.z) + 0x1000u) & 0xFFFFE000u);
                        h.w = __uint_as_float((__float_as_uint(x.w) + 0x1000u) & 0xFFFFE000u);
                    }
                    raw[i] = h;
                }
            }
        }
        fence_async_smem();
        __syncthreads();
        if (dbg && c == 0) dbg[3] = clock64();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sa), b_addr = smem_u32(sb);
            const int ksteps = kn8 >> 3;
            // K-major: 8-row groups 1024 B apart (SBO); a k-step is 32 B inside the 128-byte swizzle row and every
            //          4th step moves to the next 32-element atom (R*128 B further).
            // MN-major: one k-step = two 4-row k-groups (SBO apart), mn-atoms 512 B apart (LBO).
            // Descriptors only differ in the start-address field (units of 16 B): build once, add per step.
            const uint64_t ad0 = A_MN ? umma_desc(a_addr, 512, (TC_BM / 32) * 512, 1) : umma_desc(a_addr, 16, 1024, 2);
            const uint64_t bd0 = B_MN ? umma_desc(b_addr, 512, (BN / 32) * 512, 1) : umma_desc(b_addr, 16, 1024, 2);
            const uint32_t a_step = A_MN ? (2 * (TC_BM / 32) * 512) >> 4 : 2u;
            const uint32_t b_step = B_MN ? (uint32_t)(2 * (BN / 32) * 512) >> 4 : 2u;
            const uint32_t a_atom = A_MN ? 0u : (uint32_t)((TC_BM * 128) >> 4) - 8u;      // extra jump every 4th step
            const uint32_t b_atom = B_MN ? 0u : (uint32_t)((BN * 128) >> 4) - 8u;
            uint64_t ad = ad0, bd = bd0;
            for (int ks = 0; ks < ksteps; ++ks) {
                if (!X3) {
                    umma_tf32(tmem_d, ad, bd, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                } else {
                    // The TMEM accumulate truncates, so the error grows with the number of accumulations into one
                    // accumulator (measured ~K * 1e-8).  Chunks therefore rotate over n_main accumulators and the
                    // small correction terms get their own; the epilogue adds them in fp32 round-to-nearest.
                    const uint32_t d_main = tmem_d + (uint32_t)((c % tp.n_main) * BN);
                    const uint32_t d_corr = tmem_d + (uint32_t)(tp.n_main * BN);
                    umma_tf32(d_main, ad, bd, idesc, (c >= tp.n_main || ks > 0) ? 1u : 0u);
                    // the lo copies sit (a_bytes + b_bytes) further in the slot: descriptors differ by that offset
                    const uint64_t off = (uint64_t)((a_bytes + b_bytes) >> 4);
                    umma_tf32(d_corr, ad, bd + off, idesc, (c > 0 || ks > 0) ? 1u : 0u);   // a_hi * b_lo
                    umma_tf32(d_corr, ad + off, bd, idesc, 1u);                              // a_lo * b_hi
                }
                ad += a_step; bd += b_step;
                if ((ks & 3) == 3) { ad += a_atom; bd += b_atom; }
            }
            umma_commit(&s_bar[slot]);
        }
    }
    if (dbg) dbg[4] = clock64();
    // all MMAs retire in order: waiting for the last commit covers every chunk
    mbar_wait(&s_bar[(nchunks - 1) & 1], (uint32_t)((nchunks - 1) >> 1) & 1u);
    tc_fence_after();
    if (dbg) dbg[5] = clock64();

    pdl_trigger();
    // ---- epilogue: TMEM -> registers -> shared (row-major slab) -> coalesced global ----
    // TMEM hands each thread one accumulator ROW (lane = row); writing rows straight to global would touch 32
    // different rows per store.  The tile is therefore parked in shared memory, SLAB columns at a time (the
    // operand ring is idle once the last MMA retired), and re-read with consecutive threads on consecutive
    // columns, so the bias / mask / Adam traffic (param, two moments, Polyak target) is fully coalesced.
    float* stile = reinterpret_cast<float*>(ring);
    const int SLAB = BN < 64 ? BN : 64;
    const int sld = SLAB + 1;                               // odd row stride: conflict-free lane=row stores
    float* __restrict__ C = resolve(sp.as, T.C, seed);
    const int ldc = T.ldc, epi = T.epi;
    const float* __restrict__ bias = (epi == EPI_BIAS || epi == EPI_BIAS_RELU) ? resolve(sp.as, T.bias, seed) : nullptr;
    const float* __restrict__ mask = (epi == EPI_MASK) ? resolve(sp.as, T.mask, seed) : nullptr;
    float* __restrict__ m1 = sp.as.base[AR_ADAM_M] + (long long)seed * sp.as.stride[AR_ADAM_M];
    float* __restrict__ m2 = sp.as.base[AR_ADAM_V] + (long long)seed * sp.as.stride[AR_ADAM_V];
    float* __restrict__ pbase = sp.as.base[AR_PARAM] + (long long)seed * sp.as.stride[AR_PARAM];
    // task fields live in global memory next to buffers this loop stores to: copy them to registers once
    const long long t_adam = T.adam_off, t_adam_b = T.adam_bias_off, t_tgt = T.target_off, t_tgt_b = T.target_bias_off;
    const int t_has_bias = T.has_bias, t_train_bias = T.train_bias, t_ldmask = T.ldmask;
    float* __restrict__ pb_base = is_adam ? resolve(sp.as, T.bias, seed) : nullptr;
    const int sl_shift = 31 - __clz(SLAB);                  // SLAB is a power of two
    const int rows = min(TC_BM, M - m0);
    const int total = rows << sl_shift;
    for (int c0 = 0; c0 < BN; c0 += SLAB) {
        if (n0 + c0 >= N + (is_adam && t_has_bias ? 1 : 0)) break;          // nothing live further right (CTA-uniform)
        {
            const int q = warp & 3, half = warp >> 2;
            const int cols_per_half = SLAB >> 1;            // SLAB >= 16 -> >= 8
            const int row = q * 32 + lane;
            for (int cg = 0; cg < cols_per_half; cg += 8) {
                const int nloc = half * cols_per_half + cg;
                float v[8];
                tmem_ld8(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(c0 + nloc), v);
                if (X3) {
                    const int n_acc = min(tp.n_main, nchunks);
                    for (int a = 1; a <= tp.n_main; ++a) {
                        if (a < tp.n_main && a >= n_acc) continue;          // accumulator never written
                        float w[8];
                        tmem_ld8(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c0 + nloc), w);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] += w[j];
                    }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) stile[row * sld + nloc + j] = v[j];
            }
        }
        __syncthreads();
        if (is_adam) {
            const AdamScalars s = s_adam;
            for (int idx = tid; idx < total; idx += TC_THREADS) {
                const int r = idx >> sl_shift, nl = idx & (SLAB - 1);
                const int m = m0 + r, n = n0 + c0 + nl;
                const float x = stile[r * sld + nl];
                if (n < N) {
                    const long long e = (long long)m * ldc + n;
                    adam_update(x, C + e, m1 + t_adam + e, m2 + t_adam + e, t_tgt >= 0 ? pbase + t_tgt + e : nullptr, s);
                } else if (n == N && t_has_bias) {
                    float* pb = pb_base + m;
                    float* tgt = t_tgt_b >= 0 ? pbase + t_tgt_b + m : nullptr;
                    if (t_train_bias) adam_update(x, pb, m1 + t_adam_b + m, m2 + t_adam_b + m, tgt, s);
                    else if (tgt != nullptr && s.do_polyak) *tgt = __fadd_rn(__fmul_rn(*tgt, s.one_m_tau), __fmul_rn(*pb, s.tau));
                }
            }
        } else {
#pragma unroll 4
            for (int idx = tid; idx < total; idx += TC_THREADS) {
                const int r = idx >> sl_shift, nl = idx & (SLAB - 1);
                const int m = m0 + r, n = n0 + c0 + nl;
                if (n >= N) continue;
                float x = stile[r * sld + nl];
                if (epi == EPI_BIAS) x += __ldg(bias + n);
                else if (epi == EPI_BIAS_RELU) x = relu(x + __ldg(bias + n));
                else if (epi == EPI_MASK) x = __ldg(mask + (long long)m * t_ldmask + n) > 0.f ? x : 0.f;
                C[(long long)m * ldc + n] = x;
            }
        }
        __syncthreads();                                   // the slab buffer is reused
    }
    if (dbg) dbg[6] = clock64();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_d), "r"((uint32_t)tp.tmem_cols) : "memory");
    }
    if (dbg) dbg[7] = clock64();
}

}  // namespace oac
