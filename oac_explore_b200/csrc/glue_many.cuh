// Many-seed specialisations of the per-sample glue kernels (glue.cuh) for the batched-seed tensor-core program, where the
// head layers / dQ/da / the policy's dh2 are GEMM stages and the glue is pure per-sample math over 16 k .. 32 k rows.
// ncu on the generic kernels at 64 seeds (profiles/r02_glue64_before.txt): policy_head<1> 614 warp instructions per row
// (50 us, issue-bound), critic_head<1> 2 212 per sample (74 us for 122 MB of traffic), policy_grad<1> 26 us at 19 % issue
// utilisation (120 registers, two CTAs per SM).  The same arithmetic, restated with one self-contained warp per row, no
// shared-memory staging, no block barriers and every load of a row requested before the first use:
//   policy_head_many_kernel    TanhNormal rsample / log-prob from the head GEMM's output (bit-identical to the generic kernel)
//   critic_head_sac256_kernel  SAC, H = 256, one head per critic: six 1 KB hidden rows + four head rows as float4, six dots,
//                              targets / dLoss/dq, and dh2 of the Q-loss rows written from the registers already loaded
//   policy_grad_many_kernel    chain rule from dQ/da (GEMM output) to d(mean), d(raw log_std)  (bit-identical)
//   rank1_mask_kernel          dh2 = (dq W3) * relu'(h2) as an elementwise pass: the policy-loss first backward step used to
//                              be a K = n_heads "GEMM" (35 us at 64 seeds for 64 MB of traffic; the product is exact here)
#pragma once
#include "glue.cuh"

namespace oac {

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GLUE_THREADS) policy_head_many_kernel(PolicyHeadParams p, int use_external_eps) {
    const PolicyHeadTask& T = p.tasks[blockIdx.y];
    const int seed = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int A = p.A, B = p.B;
    float* io = p.as.base[AR_IO] + (long long)seed * p.as.stride[AR_IO];
    int32_t* cnt = p.as.counters + seed * p.as.n_counters;
    const int step = cnt[CNT_TRAIN_STEPS];
    const unsigned long long seed_key = ((unsigned long long)(uint32_t)cnt[CNT_RNG_HI] << 32) | (uint32_t)cnt[CNT_RNG_LO];
    const int row = blockIdx.x * GLUE_WARPS + warp;
    if (row < T.rows) {
        const float* head_in = resolve(p.as, T.head_in, seed) + (long long)row * T.head_ld;
        float* save = resolve(p.as, T.save, seed) + (long long)row * 4 * A;
        const int blk = row / B, b = row - blk * B;
        float lp_acc = 0.f;
        for (int j = lane; j < A; j += 32) {
            const float mean_j = ld_g(head_in + j), raw_j = ld_g(head_in + A + j);
            float log_std = fminf(fmaxf(raw_j, LOG_SIG_MIN_F), LOG_SIG_MAX_F);
            float std = expf(log_std);
            float action, eps = 0.f, lp = 0.f;
            if (p.deterministic) {
                action = tanhf(mean_j);
            } else {
                if (use_external_eps)
                    eps = io[p.off_eps + ((long long)T.eps_slot[blk] * B + b) * A + j];
                else
                    eps = philox_normal(p.rng_seed + 0x9E3779B97F4A7C15ull * seed_key,
                                        (uint32_t)T.eps_slot[blk], (uint32_t)step, (uint32_t)b, (uint32_t)j);
                float z = fmaf(std, eps, mean_j);
                action = tanhf(z);
                float d = z - mean_j;
                float var = std * std;
                lp = -(d * d) / (2.f * var) - logf(std) - 0.91893853320467274178f
                     - logf(1.f - action * action + TANH_EPS_F);
            }
            lp_acc += lp;
            const int orow = T.out_row0 + row;
            io[p.off_mean + (long long)orow * A + j] = mean_j;
            io[p.off_log_std + (long long)orow * A + j] = log_std;
            io[p.off_x + ((long long)T.dst_block[blk] * B + b) * p.x_ld + p.O + j] = action;
            save[0 * A + j] = action; save[1 * A + j] = std;
            save[2 * A + j] = raw_j; save[3 * A + j] = eps;
        }
        lp_acc = warp_sum(lp_acc);
        if (lane == 0) io[p.off_log_pi + T.out_row0 + row] = lp_acc;
    }
    if (p.tail_in_own_kernel) return;
    // ---- last CTA of this seed: bump step counters, entropy-temperature Adam step (as in policy_head_body) ----
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int total = gridDim.x * gridDim.y;
        const int prev = atomicAdd(&cnt[CNT_TICKET0], 1);
        s_last = (prev == total - 1);
        if (s_last) cnt[CNT_TICKET0] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    step_tail(p, seed);
}

// ------------------------------------------------------------------------------------------------------------------
// sources (glue.cuh CM_SAC): 0 qf1(a_pi) 1 qf2(a_pi) 2 qf1(data) 3 qf2(data) 4 target_qf1(next) 5 target_qf2(next)
__global__ void __launch_bounds__(GLUE_THREADS) critic_head_sac256_kernel(const CriticHeadParams* __restrict__ pp) {
    const CriticHeadParams& p = *pp;
    constexpr int H = 256;
    const int seed = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int B = p.B;
    const int b = blockIdx.x * GLUE_WARPS + warp;
    if (b >= B) return;
    float* io = p.as.base[AR_IO] + (long long)seed * p.as.stride[AR_IO];
    // lane l owns columns 4l .. 4l+3 and 128+4l .. 128+4l+3 of every row
    float4 h[6][2], w[4][2];
#pragma unroll
    for (int s = 0; s < 6; ++s) {
        const float4* r = reinterpret_cast<const float4*>(resolve(p.as, p.src[s].h2, seed) + (long long)(p.src[s].row0 + b) * H);
        h[s][0] = __ldg(r + lane); h[s][1] = __ldg(r + 32 + lane);
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {                       // head rows of qf1, qf2, target_qf1, target_qf2 (sources 2..5)
        const float4* r = reinterpret_cast<const float4*>(resolve(p.as, p.src[2 + n].w3, seed));
        w[n][0] = __ldg(r + lane); w[n][1] = __ldg(r + 32 + lane);
    }
    float r_pre = 0.f, d_pre = 0.f, alpha = 0.f, lp_next = 0.f, b3[4] = {0.f, 0.f, 0.f, 0.f};
    if (lane == 0) {
        r_pre = io[p.off_rewards + b]; d_pre = io[p.off_terminals + b];
        alpha = io[p.off_scalars + SC_ALPHA]; lp_next = io[p.off_log_pi + B + b];
#pragma unroll
        for (int n = 0; n < 4; ++n) b3[n] = ld_g(resolve(p.as, p.src[2 + n].b3, seed));
    }
    float v[6];
#pragma unroll
    for (int s = 0; s < 6; ++s) {
        const int n = s < 2 ? s : s - 2;               // sources 0 / 2 use qf1's head, 1 / 3 qf2's, 4 / 5 the targets'
        float a = h[s][0].x * w[n][0].x;
        a = fmaf(h[s][0].y, w[n][0].y, a); a = fmaf(h[s][0].z, w[n][0].z, a); a = fmaf(h[s][0].w, w[n][0].w, a);
        a = fmaf(h[s][1].x, w[n][1].x, a); a = fmaf(h[s][1].y, w[n][1].y, a);
        a = fmaf(h[s][1].z, w[n][1].z, a); a = fmaf(h[s][1].w, w[n][1].w, a);
        v[s] = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int s = 0; s < 6; ++s) v[s] += __shfl_xor_sync(0xffffffffu, v[s], o);
    }
    float dq1 = 0.f, dq2 = 0.f, dp1 = 0.f, dp2 = 0.f;
    if (lane == 0) {
        const float invB = 1.0f / (float)B;
        const float nd = (1.f - d_pre) * p.discount;
        const float q1n = v[0] + b3[0], q2n = v[1] + b3[1], q1 = v[2] + b3[0], q2 = v[3] + b3[1];
        const float t1 = v[4] + b3[2], t2 = v[5] + b3[3];
        const float tq = fminf(t1, t2) - alpha * lp_next;                 // trainer.py:178-184
        const float y = p.reward_scale * r_pre + nd * tq;
        io[p.off_q_pred + b * 2 + 0] = q1; io[p.off_q_pred + b * 2 + 1] = q2;
        io[p.off_q_target + b * 2 + 0] = y; io[p.off_q_target + b * 2 + 1] = y;
        io[p.off_q_new + b * 2 + 0] = q1n; io[p.off_q_new + b * 2 + 1] = q2n;
        dq1 = 2.f * (q1 - y) * invB; dq2 = 2.f * (q2 - y) * invB;         // MSELoss mean over B (trainer.py:194-195)
        const bool sel1 = q1n <= q2n;                                      // -mean(min(q1, q2)): first argument on ties
        dp1 = sel1 ? -invB : 0.f; dp2 = sel1 ? 0.f : -invB;
        resolve(p.as, p.src[2].dq, seed)[(long long)b * p.src[2].dq_ld] = dq1;
        resolve(p.as, p.src[3].dq, seed)[(long long)b * p.src[3].dq_ld] = dq2;
        resolve(p.as, p.src[0].dq, seed)[(long long)b * p.src[0].dq_ld] = dp1;
        resolve(p.as, p.src[1].dq, seed)[(long long)b * p.src[1].dq_ld] = dp2;
    }
    dq1 = __shfl_sync(0xffffffffu, dq1, 0); dq2 = __shfl_sync(0xffffffffu, dq2, 0);
    dp1 = __shfl_sync(0xffffffffu, dp1, 0); dp2 = __shfl_sync(0xffffffffu, dp2, 0);
    // ---- dh2 = (dq W3) * relu'(h2) for the sources whose backward starts with the current head weights ----
    const float dqs[4] = {dp1, dp2, dq1, dq2};
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (!p.src[s].write_dh2) continue;
        const int n = s & 1;
        float4* out = reinterpret_cast<float4*>(resolve(p.as, p.src[s].dh2, seed) + (long long)b * H);
        const float g = dqs[s];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            float4 o;
            o.x = h[s][hf].x > 0.f ? g * w[n][hf].x : 0.f; o.y = h[s][hf].y > 0.f ? g * w[n][hf].y : 0.f;
            o.z = h[s][hf].z > 0.f ? g * w[n][hf].z : 0.f; o.w = h[s][hf].w > 0.f ? g * w[n][hf].w : 0.f;
            out[hf * 32 + lane] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GLUE_THREADS) policy_grad_many_kernel(PolicyGradParams p) {
    const PolicyGradTask& T = p.tasks[blockIdx.y];
    const int seed = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int A = p.A, B = p.B;
    const int b = blockIdx.x * GLUE_WARPS + warp;
    if (b >= B) return;
    const float* io = p.as.base[AR_IO] + (long long)seed * p.as.stride[AR_IO];
    const float alpha = io[p.off_scalars + SC_ALPHA];
    const float invB = 1.0f / (float)B;
    const float* save = resolve(p.as, T.save, seed) + (long long)(T.save_row0 + b) * 4 * A;
    float* dhead = resolve(p.as, T.dhead, seed) + (long long)b * T.dhead_ld;
    for (int j = lane; j < A; j += 32) {
        float gaj = 0.f;
        for (int s = 0; s < T.n_src; ++s) gaj += ld_g(resolve(p.as, T.da[s], seed) + (long long)b * T.da_ld + j);
        const float a = save[0 * A + j];
        const float one_m_a2 = 1.f - a * a;
        float dmean, draw;
        if (T.entropy) {
            const float std = save[1 * A + j], raw = save[2 * A + j], eps = save[3 * A + j];
            const float u = one_m_a2 + TANH_EPS_F;
            dmean = alpha * invB * (2.f * a * one_m_a2 / u) + gaj * one_m_a2;
            const float dstd = dmean * eps - alpha * invB / std;
            const bool inside = (raw >= LOG_SIG_MIN_F) && (raw <= LOG_SIG_MAX_F);
            draw = inside ? dstd * std : 0.f;
        } else {
            dmean = gaj * one_m_a2;
            draw = 0.f;
        }
        dhead[j] = dmean; dhead[A + j] = draw;
    }
}

// ------------------------------------------------------------------------------------------------------------------
struct Rank1Task {
    Ref dq;  int dq_ld;        // [rows, dq_ld]: gradient w.r.t. the head outputs
    Ref w3;  int n_heads;      // [n_heads, cols]
    Ref mask; int ldmask;      // [rows, ldmask]: the activation whose ReLU is differentiated
    int mask_bits;             // 1: `mask` holds its sign bits instead (one byte per 4 columns, gemm_ws.cuh), ldmask = bytes per row
    Ref out; int ldo;          // [rows, ldo]
    int rows, cols;            // cols % 4 == 0, every leading dimension % 4 == 0, 16-byte aligned bases (checked by the host)
};
constexpr int RANK1_THREADS = 256;
constexpr int RANK1_ROWS = 4;      // rows per thread: four independent 16-byte mask loads in flight per thread

// one thread = one 4-column group of RANK1_ROWS consecutive rows (the head row slice w is shared by them)
// (the task is read in place: a by-value copy of the struct lived in local memory -- 96 bytes of stack and 94 registers, two
// CTAs per SM -- and the pass ran at 1.6 TB/s of stores, latency-bound: ncu r02b_64seeds_tf32_full.txt)
__global__ void __launch_bounds__(RANK1_THREADS, 3) rank1_mask_kernel(const Rank1Task* __restrict__ tasks, ArenaSet as) {
    const Rank1Task& T = tasks[blockIdx.y];
    const int seed = blockIdx.z;
    const int c4n = T.cols >> 2;
    const int idx = blockIdx.x * RANK1_THREADS + threadIdx.x;
    const int rg = idx / c4n, c = (idx - rg * c4n) << 2;
    const int row0 = rg * RANK1_ROWS;
    if (row0 >= T.rows) return;
    const float* mask = as.base[T.mask.arena] + (long long)seed * as.stride[T.mask.arena] + T.mask.off + (T.mask_bits ? 0 : c);
    const float* dq = as.base[T.dq.arena] + (long long)seed * as.stride[T.dq.arena] + T.dq.off;
    const float* w = as.base[T.w3.arena] + (long long)seed * as.stride[T.w3.arena] + T.w3.off + c;
    float* out = as.base[T.out.arena] + (long long)seed * as.stride[T.out.arena] + T.out.off + c;
    float4 m[RANK1_ROWS];
    if (T.mask_bits) {
        // one byte per (row, 4 columns): bit j = column c + j
        const uint8_t* mw = reinterpret_cast<const uint8_t*>(mask) + (c >> 2);
#pragma unroll
        for (int r = 0; r < RANK1_ROWS; ++r) {
            const uint32_t w = (row0 + r < T.rows) ? (uint32_t)__ldg(mw + (long long)(row0 + r) * T.ldmask) : 0u;
            m[r] = make_float4((w & 1u) ? 1.f : 0.f, (w & 2u) ? 1.f : 0.f, (w & 4u) ? 1.f : 0.f, (w & 8u) ? 1.f : 0.f);
        }
    } else {
#pragma unroll
        for (int r = 0; r < RANK1_ROWS; ++r)
            m[r] = (row0 + r < T.rows) ? __ldg(reinterpret_cast<const float4*>(mask + (long long)(row0 + r) * T.ldmask))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float4 acc[RANK1_ROWS];
#pragma unroll
    for (int r = 0; r < RANK1_ROWS; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int hd = 0; hd < T.n_heads; ++hd) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (long long)hd * T.cols));
#pragma unroll
        for (int r = 0; r < RANK1_ROWS; ++r) {
            const float g = (row0 + r < T.rows) ? __ldg(dq + (long long)(row0 + r) * T.dq_ld + hd) : 0.f;
            acc[r].x = fmaf(g, wv.x, acc[r].x); acc[r].y = fmaf(g, wv.y, acc[r].y);
            acc[r].z = fmaf(g, wv.z, acc[r].z); acc[r].w = fmaf(g, wv.w, acc[r].w);
        }
    }
#pragma unroll
    for (int r = 0; r < RANK1_ROWS; ++r) {
        if (row0 + r >= T.rows) break;
        float4 o;
        o.x = m[r].x > 0.f ? acc[r].x : 0.f; o.y = m[r].y > 0.f ? acc[r].y : 0.f;
        o.z = m[r].z > 0.f ? acc[r].z : 0.f; o.w = m[r].w > 0.f ? acc[r].w : 0.f;
        *reinterpret_cast<float4*>(out + (long long)(row0 + r) * T.ldo) = o;
    }
}

}  // namespace oac
