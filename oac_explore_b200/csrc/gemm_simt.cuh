// fp32 SIMT (FFMA) batched GEMM stage with fused epilogues: the <=1e-5 parity path.
//
// One launch executes a *stage*: grid = (max tiles, tasks, seeds).  Each task is
//     C[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
// with either operand stored K-contiguous or M/N-contiguous, so the same kernel
// runs forward (X W^T), dX (dY W) and dW (dY^T X) products without transposed
// copies.  The contraction lengths on this path are short (K <= 400), so a CTA
// stages the WHOLE K extent of its A and B tiles in shared memory with one burst of
// cp.async (every byte in flight at once: one DRAM/L2 latency per stage instead of
// one per k-tile), then runs the FFMA loop without further barriers.  Epilogues fuse
// bias+ReLU, the ReLU-derivative mask and, for dW tasks, the Adam update of the
// weight block (+ its bias from the column sums of A) and the Polyak update of the
// target copy, so every weight / moment is touched once per step (reference:
// 26 addmm + 29 mm + ~150 Adam + 48 Polyak launches).
#pragma once
#include "oac_internal.h"
#include "tma_util.cuh"

namespace oac {

struct StageParams {
    const GemmTask* tasks;    // device
    ArenaSet as;
    AdamHyper hyper;
    int kc;                   // K chunk staged per pass (multiple of 4; >= K when it fits)
    const CUtensorMap* tmaps; // gemm_sk_kernel<.., .., true>: [2 * n_tasks] fp32 tensor maps of the A and B operands
};

__device__ __forceinline__ float relu(float x) { return x > 0.f ? x : 0.f; }

// Plain (coherent) global load.  The glue kernels read activations / weights that an earlier stage wrote; inside the
// single-launch step those writes happen in the SAME kernel, where ld.global.nc (__ldg, or what nvcc derives from
// const __restrict__) may return stale lines.  volatile: never hoisted above a grid barrier.
__device__ __forceinline__ float ld_g(const float* p) {
    float v;
    asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// Programmatic dependent launch (both are no-ops for a kernel launched without the PDL attribute).
//   pdl_wait   : first thing in every kernel -- blocks until the previous stage's grid has completed and its writes
//                are visible.  Everything above it (none here) could overlap the predecessor.
//   pdl_trigger: placed AFTER a kernel's main loop.  Once every CTA of the grid has executed it (or exited), the next
//                stage's grid may be scheduled, so its launch latency and CTA set-up overlap this kernel's epilogue
//                and drain instead of following them.  Triggering at the top of the kernel (first attempt, r01) let the
//                dependent CTAs take SM slots from this kernel's own not-yet-started CTAs and was slower than no PDL.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

// torch.optim.Adam (1.4 formula order, trainer/trainer.py:75-91) on one element; no FMA
// contraction so the rounding sequence is the reference's: mul, add; mul, addcmul; sqrt,
// div, add; div, mul, add.
struct AdamScalars {
    float step_size;       // (float)(lr / (1 - beta1^t))
    float bc2_sqrt;        // (float)sqrt(1 - beta2^t)
    float beta1, beta2, one_m_beta1, one_m_beta2, eps;
    float tau, one_m_tau;
    int do_polyak;
};

// beta^t for integer t by squaring (double): agrees with pow() to ~1e-15 relative, far below the fp32 rounding of
// the two scalars derived from it, at a few dozen DP multiplies instead of two DP pow() calls -- which cost ~4 us on
// one thread and sat on the critical path of every Adam stage of the single-seed step.
__device__ __forceinline__ double ipow(double b, int t) {
    double r = 1.0;
    while (t > 0) { if (t & 1) r *= b; b *= b; t >>= 1; }
    return r;
}

__device__ __forceinline__ AdamScalars make_adam_scalars(const AdamHyper& h, float lr, int t, int train_steps_done) {
    AdamScalars s;
    double b1 = (double)h.beta1, b2 = (double)h.beta2;
    // beta is stored as float; the reference holds the Python doubles 0.9 / 0.999.
    // Recover them by rounding to 6 decimals (exact for the defaults and any sane setting).
    b1 = rint(b1 * 1e6) * 1e-6;
    b2 = rint(b2 * 1e6) * 1e-6;
    double bc1 = 1.0 - ipow(b1, t);
    double bc2 = 1.0 - ipow(b2, t);
    s.step_size = (float)((double)lr / bc1);
    s.bc2_sqrt = (float)sqrt(bc2);
    s.beta1 = (float)b1; s.beta2 = (float)b2;
    s.one_m_beta1 = (float)(1.0 - b1); s.one_m_beta2 = (float)(1.0 - b2);
    s.eps = h.eps;
    s.tau = h.tau; s.one_m_tau = h.one_minus_tau;
    // Polyak fires when (_n_train_steps_total % period == 0) with the PRE-increment count
    s.do_polyak = ((train_steps_done - 1) % (h.target_period > 0 ? h.target_period : 1)) == 0;
    return s;
}

__device__ __forceinline__ void adam_update(float g, float* __restrict__ p, float* __restrict__ m,
                                            float* __restrict__ v, float* __restrict__ tgt,
                                            const AdamScalars& s) {
    float mm = __fadd_rn(__fmul_rn(*m, s.beta1), __fmul_rn(s.one_m_beta1, g));
    float vv = __fadd_rn(__fmul_rn(*v, s.beta2), __fmul_rn(__fmul_rn(s.one_m_beta2, g), g));
    float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), s.bc2_sqrt), s.eps);
    float pp = __fadd_rn(*p, __fmul_rn(-s.step_size, __fdiv_rn(mm, denom)));
    *m = mm; *v = vv; *p = pp;
    if (tgt != nullptr && s.do_polyak) {
        // utils/pytorch_util.py:5-9: target*(1-tau) + param*tau
        *tgt = __fadd_rn(__fmul_rn(*tgt, s.one_m_tau), __fmul_rn(pp, s.tau));
    }
}

// ---- cp.async helpers ----
__device__ __forceinline__ void cp_async16_zfill(float* smem_dst, const float* gsrc, int src_bytes) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// row stride (floats) of a K-contiguous operand tile: a multiple of 4 that is 4 mod 8, so that 8
// consecutive rows' float4 at the same k fall into 8 distinct 4-bank groups (conflict-free LDS.128)
__host__ __device__ inline int kpad_of(int kc) { int p = (kc + 3) & ~3; return (p & 4) ? p : p + 4; }

// Stages rows [r0, r0+R) x contiguous [c0, c0+CW) of a strided matrix (element (r,c) at src[r*ld+c])
// into smem (row stride sld).  Out-of-range elements are zero-filled.
template <int NT>
__device__ __forceinline__ void stage_tile(float* __restrict__ sm, int sld, const float* __restrict__ src, int ld,
                                           int r0, int R, int rmax, int c0, int CW, int cmax, bool vec_ok) {
    const int tid = threadIdx.x;
    if (vec_ok) {
        // lanes over a row's 16-byte chunks (coalesced); short rows share a warp (lpr lanes per row, a power of
        // two, so the index math is shifts only)
        const int cpr = (CW + 3) >> 2;
        const int lpr_shift = cpr <= 8 ? 3 : (cpr <= 16 ? 4 : 5);
        const int lpr = 1 << lpr_shift;
        const int rows_per_pass = NT >> lpr_shift;
        const int q0 = tid & (lpr - 1);
        for (int r = tid >> lpr_shift; r < R; r += rows_per_pass) {
            const int gr = r0 + r;
            const float* prow = src + (long long)gr * ld + c0;
            float* srow = sm + r * sld;
            for (int q = q0; q < cpr; q += lpr) {
                const int gc = c0 + (q << 2);
                int bytes = 0;
                const float* p = src;
                if (gr < rmax && gc < cmax) { bytes = min(16, (cmax - gc) * 4); p = prow + (q << 2); }
                cp_async16_zfill(srow + (q << 2), p, bytes);
            }
        }
    } else {
        const int cw4 = (CW + 3) & ~3;
        const int total = R * cw4;
        for (int c = tid; c < total; c += NT) {
            const int r = c / cw4, q = c - r * cw4;
            const int gr = r0 + r, gc = c0 + q;
            if (gr < rmax && gc < cmax) cp_async4(sm + r * sld + q, src + (long long)gr * ld + gc);
            else sm[r * sld + q] = 0.f;
        }
    }
}

// N adjacent floats (N = 2 or 4, pointer aligned to N * 4 bytes) with one shared-memory load
template <int N>
__device__ __forceinline__ void load_adjacent(const float* p, float (&v)[N]) {
    if (N == 2) { const float2 x = *reinterpret_cast<const float2*>(p); v[0] = x.x; v[1] = x.y; }
    else if (N == 4) { const float4 x = *reinterpret_cast<const float4*>(p); v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; }
    else {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = p[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The FFMA tile: 32 x 32 outputs per CTA, 256 threads = 4 k-groups x 64 threads, 4 x 4 outputs per thread.  A device
// function (gemm_sk_body) so that the same code runs as its own kernel and as a phase of the single-launch step
// (mega.cuh); data produced by earlier stages is read with cp.async / TMA / plain loads only (no ld.global.nc).
// A single seed's stages have too few outputs to give every thread a large register tile AND fill the chip, and the
// 2 x 2-per-thread tile this replaces was bound by shared-memory wavefronts (4 LDS.128 per 16 FFMA; 37 % of its stall
// samples at the first FFMA after the loads, profiles/r01_gemm_simt).  Splitting K four ways inside the CTA keeps
// 256 threads per tile with 4 x 4 register tiles (8 LDS.128 per 64 FFMA); the four partial tiles meet in shared
// memory and every thread finishes one row x 4 adjacent columns (128-byte rows of global traffic in the epilogue).
// ---------------------------------------------------------------------------------------------------------------
constexpr int SK_BM = 32, SK_BN = 32, SK_T = 4, SK_KS = 4, SK_THREADS = 256, SK_PLD = 33;

// TMA = true: the whole-K operand tiles are fetched by ONE thread with cp.async.bulk.tensor (fp32 tensor maps, bounds
// zero-filled by the TMA unit) and everybody waits on one mbarrier, instead of every thread computing addresses and
// predicates for its cp.async chunks (a fifth of this kernel's instructions at K = 393).  K-contiguous operands land in
// SWIZZLE_128B atoms ([k-atom of 32][32 rows][128 B], 16-byte unit u of row r at u ^ (r & 7): conflict-free LDS.128 for
// lanes on different rows), M/N-contiguous ones as dense [k][32] rows -- the layout the cp.async path uses as well.
__host__ __device__ inline int sk_rows_per_box(int K) { const int k4 = (K + 3) & ~3; return k4 < 256 ? k4 : 256; }
__host__ __device__ inline int sk_tile_bytes(bool mn_major, int K) {
    if (!mn_major) return ((K + 31) >> 5) * 4096;
    const int rb = sk_rows_per_box(K);
    return ((K + rb - 1) / rb) * rb * 128;
}

template <bool AT, bool BT, bool TMA = false>
__device__ __forceinline__ void gemm_sk_body(const StageParams& sp, int bx, int by, int bz) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t s_tma_bar;
    __shared__ AdamScalars s_adam;
    __shared__ float s_bsum[8][SK_BM];

    const GemmTask& T = sp.tasks[by];
    const int tile = bx;
    if (tile >= T.tiles_m * T.tiles_n) return;
    const int seed = bz;
    const int tm = tile / T.tiles_n, tn = tile - tm * T.tiles_n;
    const int m0 = tm * SK_BM, n0 = tn * SK_BN;
    const int tid = threadIdx.x;
    const int kg = tid >> 6, t64 = tid & 63;
    const int tx = t64 & 7, ty = t64 >> 3;

    const float* A = resolve(sp.as, T.A, seed);
    const float* B = resolve(sp.as, T.B, seed);
    const int M = T.M, N = T.N, K = T.K;
    const int lda = T.lda, ldb = T.ldb;
    const bool a_vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool b_vec = ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

    const int kc = sp.kc;
    const int kp = kpad_of(kc);
    const int a_ld = AT ? SK_BM : kp;
    const int b_ld = BT ? SK_BN : kp;
    // TMA: 1024-byte aligned tiles (swizzle atoms), A then B, sized by this task's K
    float* tma_base = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(smem) + ((1024u - (smem_u32(smem) & 1023u)) & 1023u));
    float* As = TMA ? tma_base : smem;
    float* Bs = TMA ? tma_base + sk_tile_bytes(AT, K) / 4 : smem + (AT ? kc * SK_BM : SK_BM * kp);
    float* red = TMA ? tma_base : smem;            // the k-groups' partial tiles reuse the operand area

    const bool is_adam = T.epi == EPI_ADAM;
    float acc[SK_T][SK_T];
#pragma unroll
    for (int i = 0; i < SK_T; ++i)
#pragma unroll
        for (int j = 0; j < SK_T; ++j) acc[i][j] = 0.f;
    const bool bias_on = AT && is_adam && T.has_bias && tn == 0;
    // rows / columns of this thread: adjacent for an M/N-contiguous operand (one LDS.128 per k), interleaved by 8 for a
    // K-contiguous one (conflict-free LDS.128 along k)
    const int r0 = AT ? ty * SK_T : ty, rs = AT ? 1 : 8;
    const int c0 = BT ? tx * SK_T : tx, cs = BT ? 1 : 8;

    // epilogue mapping (thread -> one row, 4 adjacent columns), known up front so that the Adam operands can be prefetched
    const int er = tid >> 3, ec = (tid & 7) << 2;
    float* ep_C = nullptr; float* ep_m = nullptr; float* ep_v = nullptr; float* ep_t = nullptr;
    bool ep_vec = false;
    float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f), a4 = p4, v4 = p4, t4 = p4;
    if (is_adam && m0 + er < M) {
        const long long e0 = (long long)(m0 + er) * T.ldc + n0 + ec;
        ep_C = resolve(sp.as, T.C, seed) + e0;
        ep_m = sp.as.base[AR_ADAM_M] + (long long)seed * sp.as.stride[AR_ADAM_M] + T.adam_off + e0;
        ep_v = sp.as.base[AR_ADAM_V] + (long long)seed * sp.as.stride[AR_ADAM_V] + T.adam_off + e0;
        ep_t = T.target_off >= 0 ? sp.as.base[AR_PARAM] + (long long)seed * sp.as.stride[AR_PARAM] + T.target_off + e0 : nullptr;
        ep_vec = (n0 + ec + 3 < N) && ((T.ldc & 3) == 0) &&
                 (((reinterpret_cast<uintptr_t>(ep_C) | reinterpret_cast<uintptr_t>(ep_m) | reinterpret_cast<uintptr_t>(ep_v) |
                    reinterpret_cast<uintptr_t>(ep_t)) & 15) == 0);
    }

    // the ReLU mask of a dX stage (EPI_MASK) is an activation of an earlier stage: like the Adam operands, this thread's
    // four mask values are requested before the main loop instead of costing an L2 round trip in the epilogue
    const float* ep_mask = nullptr;
    bool mask_vec = false;
    float4 k4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (T.epi == EPI_MASK && m0 + er < M) {
        ep_mask = resolve(sp.as, T.mask, seed) + (long long)(m0 + er) * T.ldmask + n0 + ec;
        mask_vec = (n0 + ec + 3 < N) && ((T.ldmask & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep_mask) & 15) == 0);
    }

    for (int k0 = 0; k0 < K; k0 += (TMA ? K : kc)) {
        const int kn = TMA ? K : min(kc, K - k0);
        if (k0 > 0) __syncthreads();
        const int kn4 = (kn + 3) & ~3;            // k tails are zero-filled
        if (TMA) {
            if (tid == 0) {                          // init, arm and issue in one go; the others meet the barrier below
                mbar_init(&s_tma_bar, 1); fence_mbar_init();
                const CUtensorMap* ta = sp.tmaps + 2 * by;
                const CUtensorMap* tb = ta + 1;
                const uint32_t bar = smem_u32(&s_tma_bar);
                mbar_expect_tx(bar, (uint32_t)(sk_tile_bytes(AT, K) + sk_tile_bytes(BT, K)));
                const uint32_t sa = smem_u32(As), sb = smem_u32(Bs);
                const int rb = sk_rows_per_box(K);
                if (!AT) for (int a = 0; a < ((K + 31) >> 5); ++a) tma_load_3d(sa + a * 4096, ta, a * 32, m0, seed, bar);
                else     for (int b = 0; b * rb < K; ++b) tma_load_3d(sa + b * rb * 128, ta, m0, b * rb, seed, bar);
                if (!BT) for (int a = 0; a < ((K + 31) >> 5); ++a) tma_load_3d(sb + a * 4096, tb, a * 32, n0, seed, bar);
                else     for (int b = 0; b * rb < K; ++b) tma_load_3d(sb + b * rb * 128, tb, n0, b * rb, seed, bar);
            }
        } else {
            if (!AT) stage_tile<SK_THREADS>(As, a_ld, A, lda, m0, SK_BM, M, k0, kn, K, a_vec);
            else     stage_tile<SK_THREADS>(As, a_ld, A, lda, k0, kn4, K, m0, SK_BM, M, a_vec);
            if (!BT) stage_tile<SK_THREADS>(Bs, b_ld, B, ldb, n0, SK_BN, N, k0, kn, K, b_vec);
            else     stage_tile<SK_THREADS>(Bs, b_ld, B, ldb, k0, kn4, K, n0, SK_BN, N, b_vec);
        }
        if (k0 == 0 && ep_vec) {
            // this thread's parameter / moment / target float4s are requested now and consumed in the epilogue: their
            // L2 round trip overlaps the operand staging and the FFMA loop
            p4 = *reinterpret_cast<const float4*>(ep_C);
            a4 = *reinterpret_cast<const float4*>(ep_m);
            v4 = *reinterpret_cast<const float4*>(ep_v);
            if (ep_t) t4 = *reinterpret_cast<const float4*>(ep_t);
        }
        if (k0 == 0 && mask_vec) k4 = *reinterpret_cast<const float4*>(ep_mask);
        if (is_adam && k0 == 0 && tid == 0) {      // the step's Adam scalars: computed while the operand loads fly
            int t = sp.as.counters[seed * sp.as.n_counters + T.counter];
            int ts = sp.as.counters[seed * sp.as.n_counters + CNT_TRAIN_STEPS];
            s_adam = make_adam_scalars(sp.hyper, T.lr, t, ts);
        }
        if (TMA) {
            __syncthreads();                       // the mbarrier is initialised; s_adam is written
            mbar_wait(&s_tma_bar, 0);
        } else {
            cp_async_wait_all();
            __syncthreads();
        }
        // K-contiguous operand element (row, k) under TMA: SWIZZLE_128B atoms
        auto kmaj = [](const float* base, int row, int k) {
            return reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(base) + (k >> 5) * 4096 + row * 128 +
                                                   ((((k >> 2) & 7) ^ (row & 7)) << 4));
        };
        const int ks = ((kn4 >> 2) + SK_KS - 1) / SK_KS * 4;          // this chunk's k per group (multiple of 4)
        const int kb = kg * ks, ke = min(kb + ks, kn4);
        if (!AT && !BT) {
#pragma unroll 2
            for (int k = kb; k < ke; k += 4) {
                float4 a[SK_T], b[SK_T];
#pragma unroll
                for (int i = 0; i < SK_T; ++i) a[i] = TMA ? *kmaj(As, r0 + i * rs, k) : *reinterpret_cast<const float4*>(As + (r0 + i * rs) * a_ld + k);
#pragma unroll
                for (int j = 0; j < SK_T; ++j) b[j] = TMA ? *kmaj(Bs, c0 + j * cs, k) : *reinterpret_cast<const float4*>(Bs + (c0 + j * cs) * b_ld + k);
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
        } else if (!AT && BT) {
#pragma unroll 2
            for (int k = kb; k < ke; k += 4) {
                float4 a[SK_T];
#pragma unroll
                for (int i = 0; i < SK_T; ++i) a[i] = TMA ? *kmaj(As, r0 + i * rs, k) : *reinterpret_cast<const float4*>(As + (r0 + i * rs) * a_ld + k);
                float b[4][SK_T];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) load_adjacent<SK_T>(Bs + (k + kk) * b_ld + c0, b[kk]);
#pragma unroll
                for (int i = 0; i < SK_T; ++i)
#pragma unroll
                    for (int j = 0; j < SK_T; ++j) {
                        acc[i][j] = fmaf(a[i].x, b[0][j], acc[i][j]);
                        acc[i][j] = fmaf(a[i].y, b[1][j], acc[i][j]);
                        acc[i][j] = fmaf(a[i].z, b[2][j], acc[i][j]);
                        acc[i][j] = fmaf(a[i].w, b[3][j], acc[i][j]);
                    }
            }
        } else {
#pragma unroll 4
            for (int k = kb; k < ke; ++k) {
                float a[SK_T], b[SK_T];
                load_adjacent<SK_T>(As + k * a_ld + r0, a);
                if (BT) load_adjacent<SK_T>(Bs + k * b_ld + c0, b);
                else {
#pragma unroll
                    for (int j = 0; j < SK_T; ++j) b[j] = Bs[(c0 + j * cs) * b_ld + k];
                }
#pragma unroll
                for (int i = 0; i < SK_T; ++i)
#pragma unroll
                    for (int j = 0; j < SK_T; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            // bias gradient = column sums of the staged dY^T tile: 8 k-slices x 32 rows, one thread each (kept out of the
            // FFMA loop, where it cost 4 FADD per k on every thread)
            if (bias_on) {
                const int br = tid & 31, bs_ = tid >> 5;
                float sum = 0.f;
                for (int k = bs_; k < kn; k += 8) sum += As[k * a_ld + br];
                s_bsum[bs_][br] = (k0 == 0 ? 0.f : s_bsum[bs_][br]) + sum;
            }
        }
    }
    pdl_trigger();
    // ---- the four k-groups' partial tiles meet in shared memory (the operand tiles are dead) ----
    __syncthreads();
    float* part = red;                                            // [SK_KS][32][SK_PLD]
#pragma unroll
    for (int i = 0; i < SK_T; ++i)
#pragma unroll
        for (int j = 0; j < SK_T; ++j) part[(kg * SK_BM + r0 + i * rs) * SK_PLD + c0 + j * cs] = acc[i][j];
    __syncthreads();
    // ---- epilogue: thread -> one row, 4 adjacent columns ----
    float v[SK_T];
#pragma unroll
    for (int j = 0; j < SK_T; ++j) {
        float x = part[er * SK_PLD + ec + j];
#pragma unroll
        for (int g = 1; g < SK_KS; ++g) x += part[(g * SK_BM + er) * SK_PLD + ec + j];
        v[j] = x;
    }
    const int m = m0 + er;
    if (m >= M) return;
    float* __restrict__ C = resolve(sp.as, T.C, seed);
    const int ldc = T.ldc;
    const int epi = T.epi;
    if (is_adam) {
        float* __restrict__ m1 = sp.as.base[AR_ADAM_M] + (long long)seed * sp.as.stride[AR_ADAM_M];
        float* __restrict__ m2 = sp.as.base[AR_ADAM_V] + (long long)seed * sp.as.stride[AR_ADAM_V];
        float* __restrict__ pbase = sp.as.base[AR_PARAM] + (long long)seed * sp.as.stride[AR_PARAM];
        const AdamScalars s = s_adam;
        float* pt = s.do_polyak ? ep_t : nullptr;
        if (ep_vec) {
            // one 16-byte access per array (issued before the main loop) instead of four dependent scalar round trips
            adam_update(v[0], &p4.x, &a4.x, &v4.x, pt ? &t4.x : nullptr, s);
            adam_update(v[1], &p4.y, &a4.y, &v4.y, pt ? &t4.y : nullptr, s);
            adam_update(v[2], &p4.z, &a4.z, &v4.z, pt ? &t4.z : nullptr, s);
            adam_update(v[3], &p4.w, &a4.w, &v4.w, pt ? &t4.w : nullptr, s);
            *reinterpret_cast<float4*>(ep_C) = p4;
            *reinterpret_cast<float4*>(ep_m) = a4;
            *reinterpret_cast<float4*>(ep_v) = v4;
            if (pt) *reinterpret_cast<float4*>(pt) = t4;
        } else {
#pragma unroll
            for (int j = 0; j < SK_T; ++j) {
                if (n0 + ec + j >= N) continue;
                adam_update(v[j], ep_C + j, ep_m + j, ep_v + j, pt ? pt + j : nullptr, s);
            }
        }
        if (bias_on && ec == 0) {
            float bs_ = s_bsum[0][er];
#pragma unroll
            for (int g = 1; g < 8; ++g) bs_ += s_bsum[g][er];
            float* pb = resolve(sp.as, T.bias, seed) + m;
            float* tgt = T.target_bias_off >= 0 ? pbase + T.target_bias_off + m : nullptr;
            if (T.train_bias) {
                adam_update(bs_, pb, m1 + T.adam_bias_off + m, m2 + T.adam_bias_off + m, tgt, s);
            } else if (tgt != nullptr && s.do_polyak) {
                *tgt = __fadd_rn(__fmul_rn(*tgt, s.one_m_tau), __fmul_rn(*pb, s.tau));   // a frozen bias still takes part in soft_update_from_to
            }
        }
        return;
    }
    const float* bias = (epi == EPI_BIAS || epi == EPI_BIAS_RELU) ? resolve(sp.as, T.bias, seed) : nullptr;
    const float* mask = (epi == EPI_MASK) ? resolve(sp.as, T.mask, seed) : nullptr;
    const unsigned mbits = (k4.x > 0.f ? 1u : 0u) | (k4.y > 0.f ? 2u : 0u) | (k4.z > 0.f ? 4u : 0u) | (k4.w > 0.f ? 8u : 0u);
#pragma unroll
    for (int j = 0; j < SK_T; ++j) {
        const int n = n0 + ec + j;
        if (n >= N) continue;
        float x = v[j];
        if (epi == EPI_BIAS) x += bias[n];
        else if (epi == EPI_BIAS_RELU) x = relu(x + bias[n]);
        else if (epi == EPI_MASK) x = (mask_vec ? ((mbits >> j) & 1u) != 0u : mask[(long long)m * T.ldmask + n] > 0.f) ? x : 0.f;
        C[(long long)m * ldc + n] = x;
    }
}

template <bool AT, bool BT, bool TMA = false>
__global__ void __launch_bounds__(SK_THREADS) gemm_sk_kernel(StageParams sp) {
    pdl_wait();
    gemm_sk_body<AT, BT, TMA>(sp, blockIdx.x, blockIdx.y, blockIdx.z);
}

}  // namespace oac
