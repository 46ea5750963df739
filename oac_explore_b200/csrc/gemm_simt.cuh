// fp32 SIMT (FFMA) batched GEMM stage with fused epilogues: the <=1e-5 parity path.
//
// One launch executes a *stage*: grid = (max tiles, tasks, seeds).  Each task is
//     C[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
// with either operand stored K-contiguous or M/N-contiguous, so the same kernel
// runs forward (X W^T), dX (dY W) and dW (dY^T X) products without transposed
// copies.  Epilogues fuse bias+ReLU, the ReLU-derivative mask and, for dW tasks,
// the Adam update of the weight block (+ its bias from the column sums of A) and
// the Polyak update of the target copy, so every weight / moment is touched once
// per step (reference: 26 addmm + 29 mm + ~150 Adam + 48 Polyak launches,
// SURVEY.md section 2.2).
#pragma once
#include "oac_internal.h"

namespace oac {

struct StageParams {
    const GemmTask* tasks;    // device
    ArenaSet as;
    AdamHyper hyper;
};

__device__ __forceinline__ float relu(float x) { return x > 0.f ? x : 0.f; }

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

__device__ __forceinline__ AdamScalars make_adam_scalars(const AdamHyper& h, float lr, int t, int train_steps_done) {
    AdamScalars s;
    double b1 = (double)h.beta1, b2 = (double)h.beta2;
    // beta is stored as float; the reference holds the Python doubles 0.9 / 0.999.
    // Recover them by rounding to 6 decimals (exact for the defaults and any sane setting).
    b1 = rint(b1 * 1e6) * 1e-6;
    b2 = rint(b2 * 1e6) * 1e-6;
    double bc1 = 1.0 - pow(b1, (double)t);
    double bc2 = 1.0 - pow(b2, (double)t);
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

// Loads a ROWSxCOLS tile slice as float4 along the contiguous dimension.
// contiguous index c in [0,C), strided index r in [0,R): element at src[r*ld + c].
__device__ __forceinline__ float4 load4_guard(const float* __restrict__ src, int ld, int r, int c,
                                              int R, int C, bool vec_ok) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < R) {
        const float* p = src + (long long)r * ld + c;
        if (vec_ok && c + 3 < C) {
            v = *reinterpret_cast<const float4*>(p);
        } else {
            if (c + 0 < C) v.x = p[0];
            if (c + 1 < C) v.y = p[1];
            if (c + 2 < C) v.z = p[2];
            if (c + 3 < C) v.w = p[3];
        }
    }
    return v;
}

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_stage_kernel(StageParams sp) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int TX = BN / TN;
    static_assert(BK % 4 == 0 && BM % 4 == 0 && BN % 4 == 0, "tile dims");
    constexpr int A_VECS = BM * BK / 4;     // float4 per A tile
    constexpr int B_VECS = BN * BK / 4;
    static_assert(A_VECS + B_VECS <= 2 * NT && A_VECS <= NT && B_VECS <= NT, "one float4 per thread per operand");
    constexpr int PADM = BM + 4, PADN = BN + 4;

    __shared__ __align__(16) float As[2][BK][PADM];
    __shared__ __align__(16) float Bs[2][BK][PADN];
    __shared__ AdamScalars s_adam;

    const GemmTask& T = sp.tasks[blockIdx.y];
    const int tile = blockIdx.x;
    if (tile >= T.tiles_m * T.tiles_n) return;
    const int seed = blockIdx.z;
    const int tm = tile / T.tiles_n, tn = tile % T.tiles_n;
    const int m0 = tm * BM, n0 = tn * BN;
    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;

    const float* __restrict__ A = resolve(sp.as, T.A, seed);
    const float* __restrict__ B = resolve(sp.as, T.B, seed);
    const int M = T.M, N = T.N, K = T.K;
    const int lda = T.lda, ldb = T.ldb;
    const bool a_trans = T.a_trans != 0, b_trans = T.b_trans != 0;
    const bool a_vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool b_vec = ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

    if (T.epi == EPI_ADAM && tid == 0) {
        int t = sp.as.counters[seed * sp.as.n_counters + T.counter];
        int ts = sp.as.counters[seed * sp.as.n_counters + CNT_TRAIN_STEPS];
        s_adam = make_adam_scalars(sp.hyper, T.lr, t, ts);
    }

    // per-thread load coordinates
    // A, K-contiguous:  thread -> (m = v / (BK/4), k4 = (v % (BK/4))*4)
    // A, M-contiguous:  thread -> (k = v / (BM/4), m4 = (v % (BM/4))*4)
    const bool loads_a = tid < A_VECS;
    const int vb = (A_VECS + B_VECS <= NT) ? tid - A_VECS : tid;   // B loader index
    const bool loads_b = (A_VECS + B_VECS <= NT) ? (tid >= A_VECS && vb < B_VECS) : (tid < B_VECS);

    float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
    auto fetch = [&](int k0) {
        if (loads_a) {
            if (!a_trans) {
                int m = tid / (BK / 4), k4 = (tid % (BK / 4)) * 4;
                ra = load4_guard(A, lda, m0 + m, k0 + k4, M, K, a_vec);
            } else {
                int k = tid / (BM / 4), m4 = (tid % (BM / 4)) * 4;
                ra = load4_guard(A, lda, k0 + k, m0 + m4, K, M, a_vec);
            }
        }
        if (loads_b) {
            if (!b_trans) {
                int n = vb / (BK / 4), k4 = (vb % (BK / 4)) * 4;
                rb = load4_guard(B, ldb, n0 + n, k0 + k4, N, K, b_vec);
            } else {
                int k = vb / (BN / 4), n4 = (vb % (BN / 4)) * 4;
                rb = load4_guard(B, ldb, k0 + k, n0 + n4, K, N, b_vec);
            }
        }
    };
    auto stash = [&](int buf) {
        if (loads_a) {
            if (!a_trans) {
                int m = tid / (BK / 4), k4 = (tid % (BK / 4)) * 4;
                As[buf][k4 + 0][m] = ra.x; As[buf][k4 + 1][m] = ra.y;
                As[buf][k4 + 2][m] = ra.z; As[buf][k4 + 3][m] = ra.w;
            } else {
                int k = tid / (BM / 4), m4 = (tid % (BM / 4)) * 4;
                *reinterpret_cast<float4*>(&As[buf][k][m4]) = ra;
            }
        }
        if (loads_b) {
            if (!b_trans) {
                int n = vb / (BK / 4), k4 = (vb % (BK / 4)) * 4;
                Bs[buf][k4 + 0][n] = rb.x; Bs[buf][k4 + 1][n] = rb.y;
                Bs[buf][k4 + 2][n] = rb.z; Bs[buf][k4 + 3][n] = rb.w;
            } else {
                int k = vb / (BN / 4), n4 = (vb % (BN / 4)) * 4;
                *reinterpret_cast<float4*>(&Bs[buf][k][n4]) = rb;
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    float bsum[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) bsum[i] = 0.f;
    const float bias_on = (T.epi == EPI_ADAM && T.has_bias && tn == 0) ? 1.f : 0.f;

    const int nk = (K + BK - 1) / BK;
    fetch(0);
    stash(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) fetch((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[buf][k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[buf][k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                bsum[i] = fmaf(bias_on, a[i], bsum[i]);
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
        if (kt + 1 < nk) stash(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue ----
    float* __restrict__ C = resolve(sp.as, T.C, seed);
    const int ldc = T.ldc;
    const int epi = T.epi;
    if (epi == EPI_ADAM) {
        float* __restrict__ m1 = sp.as.base[AR_ADAM_M] + (long long)seed * sp.as.stride[AR_ADAM_M];
        float* __restrict__ m2 = sp.as.base[AR_ADAM_V] + (long long)seed * sp.as.stride[AR_ADAM_V];
        float* __restrict__ pbase = sp.as.base[AR_PARAM] + (long long)seed * sp.as.stride[AR_PARAM];
        const AdamScalars s = s_adam;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int m = m0 + ty * TM + i;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                int n = n0 + tx * TN + j;
                if (n >= N) continue;
                long long e = (long long)m * ldc + n;
                float* tgt = T.target_off >= 0 ? pbase + T.target_off + e : nullptr;
                adam_update(acc[i][j], C + e, m1 + T.adam_off + e, m2 + T.adam_off + e, tgt, s);
            }
            if (bias_on != 0.f && tx == 0 && T.train_bias) {
                float* pb = resolve(sp.as, T.bias, seed) + m;
                float* tgt = T.target_bias_off >= 0 ? pbase + T.target_bias_off + m : nullptr;
                adam_update(bsum[i], pb, m1 + T.adam_bias_off + m, m2 + T.adam_bias_off + m, tgt, s);
            } else if (bias_on != 0.f && tx == 0 && !T.train_bias && T.target_bias_off >= 0 && s.do_polyak) {
                // frozen bias still takes part in soft_update_from_to (it is a parameter)
                float* pb = resolve(sp.as, T.bias, seed) + m;
                float* tgt = pbase + T.target_bias_off + m;
                *tgt = __fadd_rn(__fmul_rn(*tgt, s.one_m_tau), __fmul_rn(*pb, s.tau));
            }
        }
        return;
    }
    const float* __restrict__ bias = (epi == EPI_BIAS || epi == EPI_BIAS_RELU) ? resolve(sp.as, T.bias, seed) : nullptr;
    const float* __restrict__ mask = (epi == EPI_MASK) ? resolve(sp.as, T.mask, seed) : nullptr;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (epi == EPI_BIAS) v += bias[n];
            else if (epi == EPI_BIAS_RELU) v = relu(v + bias[n]);
            else if (epi == EPI_MASK) v = mask[(long long)m * T.ldmask + n] > 0.f ? v : 0.f;
            C[(long long)m * ldc + n] = v;
        }
    }
}

}  // namespace oac
