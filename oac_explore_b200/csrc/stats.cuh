// On-device diagnostics: the reference fills `eval_statistics` once per epoch with ~40 host-side numpy reductions
// after seven device-to-host copies (trainer/trainer.py:230-279, particle_trainer_oac.py:332-362,
// gaussian_trainer.py:397-436).  Here ONE kernel reduces the step's per-sample outputs (already sitting in the IO
// slice) into one small fp32 vector per seed -- the same quantities, in the reference's key order -- which is also the
// payload of the per-seed statistics all-gather.  `out` may be mapped pinned host memory.
//
// Vector layouts (create_stats_ordered_dict = Mean, Std (population), Max, Min):
//   SAC   [32]: QF mean, QF std, QF1 Loss, QF2 Loss, Q Loss, Policy Loss, Q1 Predictions x4, Q2 Predictions x4,
//               Q Targets x4, Log Pis x4, Policy mu x4, Policy log std x4, Alpha, Alpha Loss
//   P-OAC [11 + 9P]: QF mean, QF std, then per particle i: QFi Loss, QiPredictions x4, QiTargets x4; Policy Loss,
//               Policy mu x4, Policy log std x4
//   G-OAC [29]: QF mean, QF std, QF Loss, Q Predictions x4, Q Target x4, STD Loss, Q STD Predictions x4,
//               Q STD Target x4, Policy Loss, Policy mu x4, Policy log std x4   (mu / log std of the target policy, :361-363)
#pragma once
#include "oac_internal.h"

namespace oac {

constexpr int STATS_THREADS = 256;
constexpr int STATS_MAX = 160;                 // 11 + 9 * 16

struct StatsParams {
    ArenaSet as;
    int algo, B, A, P, nq, deterministic, auto_alpha;
    float standard_bound;
    long long off_q_pred, off_q_target, off_q_new, off_log_pi, off_mean, off_log_std, off_scalars;
    float* out;
    int out_ld;
};

__device__ __forceinline__ double stats_block_sum(double v) {
    __shared__ double red[STATS_THREADS / 32];
    __shared__ double total;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < STATS_THREADS / 32; ++w) t += red[w];
        total = t;
    }
    __syncthreads();
    const double r = total;
    __syncthreads();
    return r;
}
__device__ __forceinline__ float stats_block_max(float v) {
    __shared__ float red[STATS_THREADS / 32];
    __shared__ float total;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = red[0];
        for (int w = 1; w < STATS_THREADS / 32; ++w) t = fmaxf(t, red[w]);
        total = t;
    }
    __syncthreads();
    const float r = total;
    __syncthreads();
    return r;
}
// Mean, Std (population, like np.std), Max, Min of x[i * stride], i < n
__device__ __forceinline__ void stats_quad(const float* x, int n, int stride, float* out4) {
    double s = 0.0, ss = 0.0;
    float mx = -INFINITY, mn = INFINITY;
    for (int i = threadIdx.x; i < n; i += STATS_THREADS) {
        const float v = x[(long long)i * stride];
        s += v; ss += (double)v * v; mx = fmaxf(mx, v); mn = fminf(mn, v);
    }
    s = stats_block_sum(s); ss = stats_block_sum(ss);
    mx = stats_block_max(mx); mn = -stats_block_max(-mn);
    if (threadIdx.x == 0) {
        const double mean = s / n, var = ss / n - mean * mean;
        out4[0] = (float)mean; out4[1] = (float)sqrt(var > 0.0 ? var : 0.0); out4[2] = mx; out4[3] = mn;
    }
}
// mean over i < n of (a[i*sa] - b[i*sb])^2
__device__ __forceinline__ float stats_mse(const float* a, int sa, const float* b, int sb, int n) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += STATS_THREADS) {
        const double d = (double)a[(long long)i * sa] - (double)b[(long long)i * sb];
        s += d * d;
    }
    return (float)(stats_block_sum(s) / n);
}

__global__ void __launch_bounds__(STATS_THREADS) trainer_stats_kernel(StatsParams p) {
    const int seed = blockIdx.x, tid = threadIdx.x;
    const float* io = p.as.base[AR_IO] + (long long)seed * p.as.stride[AR_IO];
    float* out = p.out + (long long)seed * p.out_ld;
    const int B = p.B, A = p.A, nq = p.nq;
    const float* qp = io + p.off_q_pred;
    const float* qt = io + p.off_q_target;
    const float* qn = io + p.off_q_new;
    const float* lp = io + p.off_log_pi;
    const float* sc = io + p.off_scalars;
    if (p.algo == OAC_ALGO_SAC) {
        // trainer/trainer.py:230-279
        double m = 0.0, sd = 0.0, pl = 0.0;
        for (int b = tid; b < B; b += STATS_THREADS) {
            const float q1 = qp[2 * b], q2 = qp[2 * b + 1];
            m += 0.5 * ((double)q1 + q2);
            sd += 0.5 * fabs((double)q1 - q2);                       // np.std of two values
            pl += (double)lp[b] - fminf(qn[2 * b], qn[2 * b + 1]);  // logged without alpha (:236)
        }
        m = stats_block_sum(m); sd = stats_block_sum(sd); pl = stats_block_sum(pl);
        const float l1 = stats_mse(qp, 2, qt, 2, B), l2 = stats_mse(qp + 1, 2, qt, 2, B);
        if (tid == 0) {
            out[0] = (float)(m / B); out[1] = (float)(sd / B); out[2] = l1; out[3] = l2; out[4] = l1 + l2;
            out[5] = (float)(pl / B);
            out[30] = p.auto_alpha ? sc[SC_ALPHA] : 0.f; out[31] = p.auto_alpha ? sc[SC_ALPHA_LOSS] : 0.f;
        }
        stats_quad(qp, B, 2, out + 6);
        stats_quad(qp + 1, B, 2, out + 10);
        stats_quad(qt, B, 2, out + 14);
        stats_quad(lp, B, 1, out + 18);
        stats_quad(io + p.off_mean, B * A, 1, out + 22);
        stats_quad(io + p.off_log_std, B * A, 1, out + 26);
    } else if (p.algo == OAC_ALGO_POAC) {
        // trainer/particle_trainer_oac.py:332-362; q_pred / q_target are [B, P] (sorted particles / per-rank targets)
        const int P = p.P;
        const float alpha = sc[SC_ALPHA];
        double m = 0.0, sd = 0.0, pl = 0.0;
        for (int b = tid; b < B; b += STATS_THREADS) {
            double s = 0.0, ss = 0.0;
            float mn = INFINITY;
            for (int i = 0; i < P; ++i) {
                const double v = qp[(long long)b * nq + i];
                s += v; ss += v * v;
                mn = fminf(mn, qn[(long long)b * nq + i]);
            }
            const double mu = s / P, var = ss / P - mu * mu;
            m += mu; sd += sqrt(var > 0.0 ? var : 0.0);
            pl += (p.deterministic ? 0.0 : (double)alpha * lp[b]) - mn;
        }
        m = stats_block_sum(m); sd = stats_block_sum(sd); pl = stats_block_sum(pl);
        if (tid == 0) { out[0] = (float)(m / B); out[1] = (float)(sd / B); out[2 + 9 * P] = (float)(pl / B); }
        for (int i = 0; i < P; ++i) {
            const float l = stats_mse(qp + i, nq, qt + i, nq, B);
            if (tid == 0) out[2 + 9 * i] = l;
            stats_quad(qp + i, B, nq, out + 2 + 9 * i + 1);
            stats_quad(qt + i, B, nq, out + 2 + 9 * i + 5);
        }
        stats_quad(io + p.off_mean, B * A, 1, out + 3 + 9 * P);
        stats_quad(io + p.off_log_std, B * A, 1, out + 7 + 9 * P);
    } else {
        // trainer/gaussian_trainer.py:397-436; q_pred / q_target / q_new are [B, 2] = (mean, std)
        double m = 0.0, sd = 0.0, pl = 0.0;
        for (int b = tid; b < B; b += STATS_THREADS) {
            m += qp[2 * b]; sd += qp[2 * b + 1];
            pl += (double)qn[2 * b] + (double)p.standard_bound * qn[2 * b + 1];
        }
        m = stats_block_sum(m); sd = stats_block_sum(sd); pl = stats_block_sum(pl);
        const float lq = stats_mse(qp, 2, qt, 2, B), ls = stats_mse(qp + 1, 2, qt + 1, 2, B);
        if (tid == 0) { out[0] = (float)(m / B); out[1] = (float)(sd / B); out[2] = lq; out[11] = ls; out[20] = (float)(pl / B); }
        stats_quad(qp, B, 2, out + 3);
        stats_quad(qt, B, 2, out + 7);
        stats_quad(qp + 1, B, 2, out + 12);
        stats_quad(qt + 1, B, 2, out + 16);
        stats_quad(io + p.off_mean + 2ll * B * A, B * A, 1, out + 21);       // the target policy's rows [2B, 3B)
        stats_quad(io + p.off_log_std + 2ll * B * A, B * A, 1, out + 25);
    }
}

}  // namespace oac
