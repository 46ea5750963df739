// Per-sample "glue" kernels between the GEMM stages: policy heads + TanhNormal
// rsample/log_prob (+ the entropy-temperature Adam step), critic heads + the
// algorithm-specific targets / loss gradients (SAC twin-min, P-OAC sorted
// particles, G-OAC mean/std), and the policy-loss gradient w.r.t. the policy
// outputs.  One warp per sample; everything is fp32.
#pragma once
#include <type_traits>
#include "gemm_simt.cuh"
#include "device_util.cuh"

namespace oac {

// =====================================================================================
// policy heads: mean / log_std GEMV + TanhNormal.rsample + log_prob  (trainer/policies.py:260-316)
// =====================================================================================
constexpr int GLUE_MAX_HR = 16;     // hidden <= 512: one hidden row = <= 16 registers per lane

// contiguous global -> shared copy with every load in flight at once (cp.async); caller waits + syncs
__device__ __forceinline__ void stage_contig(float* dst, const float* __restrict__ src, int n) {
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)__cvta_generic_to_shared(dst)) & 15) == 0;
    if (vec) {
        const int n4 = n >> 2;
        for (int i = threadIdx.x; i < n4; i += blockDim.x) cp_async16_zfill(dst + 4 * i, src + 4 * i, 16);
        for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) cp_async4(dst + i, src + i);
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) cp_async4(dst + i, src + i);
    }
}

// Row groups: G warps cooperate on one row and synchronise through shared memory.  G = 4 spreads a single seed's
// 256..512 rows over enough CTAs (latency); G = 1 makes every warp self-contained (no block barriers inside the
// row loop), which is what the many-seed launches want.
template <int G>
__device__ __forceinline__ void group_sync() {
    if (G == 1) __syncwarp(); else __syncthreads();
}

struct PolicyHeadTask {
    Ref h2;             // [rows, H] last hidden activation
    Ref w, b;           // heads [2A, H] (mean rows then log_std rows), [2A]
    int rows;           // B or 2B
    int out_row0;       // first row in the io outputs (log_pi [.], mean/log_std [., A])
    int dst_block[2];   // X row block (units of B rows) receiving tanh actions, per B-row block
    int eps_slot[2];    // io eps slot per block
    Ref save;           // [rows, 4, A]: action, std, raw log_std, eps  (for the backward glue)
    Ref head_in;        // PolicyHeadParams::head_from_gemm: [rows, head_ld] head outputs (mean | raw log_std) of a GEMM stage
    int head_ld;
};

struct AlphaUpdate {
    int enabled;        // use_automatic_entropy_tuning
    int task, block;    // which policy-head rows feed it (out rows out_row0 + block*B .. +B)
    Ref log_alpha;      // AR_PARAM
    long long adam_off; // in the Adam arenas
    float lr, target_entropy;
    int counter;
};

struct PolicyHeadParams {
    const PolicyHeadTask* tasks;
    ArenaSet as;
    AdamHyper hyper;
    AlphaUpdate alpha;
    long long off_x, off_eps, off_log_pi, off_mean, off_log_std, off_scalars;
    int x_ld, O, A, H, B;
    int deterministic, use_external_eps;
    unsigned long long rng_seed;
    int n_opt_counters;      // counters CNT_OPT0 .. CNT_OPT0+n-1 are bumped once per step here
    int iters;               // row groups (of GLUE_SPC rows) per CTA
    int head_from_gemm;      // 1: the head layer ran as a tensor-core GEMM stage; this kernel only samples
    float* host_scalars;     // optional mapped pinned host copy of the step's scalars, [2][n_seeds, SC_COUNT] (slot = step parity)
    int n_seeds;
    int tail_in_own_kernel;  // 1: policy_head ends at its last store; step_tail_kernel follows on a side lane
};

// Once per seed and step, after every row's log-prob is written: bump the step counters, take the entropy-temperature
// Adam step (trainer/trainer.py:140-149) and publish the step's scalars.  Runs in the last policy_head CTA to finish, or
// as a one-CTA-per-seed kernel of its own on a side lane (the single-seed critical chain then ends policy_head at its
// last store: the fence / ticket / reduction tail cost ~2 us there).
__device__ __forceinline__ void step_tail(const PolicyHeadParams& p, int seed) {
    __shared__ float s_red[GLUE_THREADS];
    const int B = p.B;
    float* io = p.as.base[AR_IO] + (long long)seed * p.as.stride[AR_IO];
    int32_t* cnt = p.as.counters + seed * p.as.n_counters;
    const int step = cnt[CNT_TRAIN_STEPS];
    float part = 0.f;
    if (p.alpha.enabled) {
        const PolicyHeadTask& TA = p.tasks[p.alpha.task];
        const volatile float* lp = io + p.off_log_pi + TA.out_row0 + (long long)p.alpha.block * B;
        for (int i = threadIdx.x; i < B; i += GLUE_THREADS) part += lp[i];
    }
    s_red[threadIdx.x] = part;
    __syncthreads();
    for (int s = GLUE_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) s_red[threadIdx.x] += s_red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        cnt[CNT_TRAIN_STEPS] = step + 1;
        for (int i = 0; i < p.n_opt_counters; ++i) cnt[CNT_OPT0 + i] += 1;
        float* sc = io + p.off_scalars;
        if (p.alpha.enabled) {
            // alpha_loss = -(log_alpha * (log_pi + target_entropy).detach()).mean()  (trainer.py:140-146)
            float mean_lp = s_red[0] / (float)B;
            float* la = resolve(p.as, p.alpha.log_alpha, seed);
            float* m1 = p.as.base[AR_ADAM_M] + (long long)seed * p.as.stride[AR_ADAM_M] + p.alpha.adam_off;
            float* m2 = p.as.base[AR_ADAM_V] + (long long)seed * p.as.stride[AR_ADAM_V] + p.alpha.adam_off;
            float tgt = mean_lp + p.alpha.target_entropy;
            sc[SC_ALPHA_LOSS] = -(la[0] * tgt);
            sc[SC_MEAN_LOGPI] = mean_lp;
            AdamScalars s = make_adam_scalars(p.hyper, p.alpha.lr, cnt[CNT_OPT0 + p.alpha.counter], step + 1);
            adam_update(-tgt, la, m1, m2, nullptr, s);
            sc[SC_ALPHA] = expf(la[0]);      // POST-step alpha (trainer.py:147)
        } else {
            sc[SC_ALPHA] = 0.f;              // the fork's choice (trainer.py:148-149)
            sc[SC_ALPHA_LOSS] = 0.f;
        }
        if (p.host_scalars != nullptr) {     // zero-copy device -> host: the caller only waits for the step
            // two slots, chosen by the parity of the step index, each stamped with the number of the step that wrote it: a
            // caller that runs one step ahead of the device reads step i's scalars while step i + 1 is in flight
            float* hs = p.host_scalars + ((long long)(step & 1) * p.n_seeds + seed) * SC_COUNT;
            hs[SC_ALPHA] = sc[SC_ALPHA]; hs[SC_ALPHA_LOSS] = sc[SC_ALPHA_LOSS]; hs[SC_MEAN_LOGPI] = sc[SC_MEAN_LOGPI];
            hs[SC_STEP_STAMP] = (float)(step + 1);
        }
    }
}

// Fused head layer + sampling: the [2A, H] head weights are staged in shared memory once per CTA; GLUE_G warps
// share one row (each keeps the hidden row in registers and reduces every GLUE_G-th pair of dot products with
// shuffles), then one of them runs the per-action-dim sampling math.  dyn smem: (2A*H + 2A + SPC*2A) floats.
template <int G>
__device__ __forceinline__ void policy_head_body(const PolicyHeadParams& p, int use_external_eps, int bx, int by, int bz,
                                                 int gdx, int gdy) {
    constexpr int SPC = GLUE_WARPS / G;                    // rows per group (G warps share a row)
    extern __shared__ __align__(16) float s_ph[];
    const PolicyHeadTask& T = p.tasks[by];
    const int seed = bz;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sl = warp / G, g = warp % G;
    const int A = p.A, B = p.B, H = p.H;
    float* io = p.as.base[AR_IO] + (long long)seed * p.as.stride[AR_IO];
    int32_t* cnt = p.as.counters + seed * p.as.n_counters;
    const int step = cnt[CNT_TRAIN_STEPS];
    // the seed's own key (the real seed id of a batched seed): noise streams follow the seed, not its slot in the group
    const unsigned long long seed_key = ((unsigned long long)(uint32_t)cnt[CNT_RNG_HI] << 32) | (uint32_t)cnt[CNT_RNG_LO];

    const bool from_gemm = p.head_from_gemm != 0;           // dyn smem is then only SPC * 2A floats
    float* Ws = s_ph;
    float* bs = s_ph + (from_gemm ? 0 : 2 * A * H);
    float* outv = bs + (from_gemm ? 0 : 2 * A) + sl * 2 * A;  // this row's head outputs (mean | raw log_std)
    if (!from_gemm) {
        stage_contig(Ws, resolve(p.as, T.w, seed), 2 * A * H);
        stage_contig(bs, resolve(p.as, T.b, seed), 2 * A);
    }
    const float* h2 = resolve(p.as, T.h2, seed);
    const float* head_in = from_gemm ? resolve(p.as, T.head_in, seed) : nullptr;
    // the staged head weights serve p.iters row groups (many-seed launches: 16 -> 1/16 of the staging traffic)
    float hreg[GLUE_MAX_HR];
    auto load_row = [&](int row) {
        if (row < T.rows && !from_gemm) {
            const float* h = h2 + (long long)row * H;
#pragma unroll
            for (int c = 0; c < GLUE_MAX_HR; ++c) hreg[c] = (lane + 32 * c < H) ? ld_g(h + lane + 32 * c) : 0.f;
        }
    };
    load_row(bx * p.iters * SPC + sl);
    cp_async_wait_all();
    __syncthreads();
    for (int it = 0; it < p.iters; ++it) {
    const int row = (bx * p.iters + it) * SPC + sl;
    if (row < T.rows && from_gemm) {
        if (g == 0) for (int j = lane; j < 2 * A; j += 32) outv[j] = ld_g(head_in + (long long)row * T.head_ld + j);
    } else if (row < T.rows) {
        for (int j = g; j < A; j += G) {
            const float* wm = Ws + j * H;
            const float* ws = Ws + (A + j) * H;
            float sm = 0.f, ss = 0.f;
            if (H == 256) {                                 // the common width: 8 unconditional steps
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int k = lane + 32 * c;
                    sm = fmaf(hreg[c], wm[k], sm); ss = fmaf(hreg[c], ws[k], ss);
                }
            } else {
#pragma unroll
                for (int c = 0; c < GLUE_MAX_HR; ++c) {
                    const int k = lane + 32 * c;
                    if (k < H) { sm = fmaf(hreg[c], wm[k], sm); ss = fmaf(hreg[c], ws[k], ss); }
                }
            }
            // the two reductions interleaved: one shuffle latency chain instead of two
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sm += __shfl_xor_sync(0xffffffffu, sm, o);
                ss += __shfl_xor_sync(0xffffffffu, ss, o);
            }
            if (lane == 0) { outv[j] = sm + bs[j]; outv[A + j] = ss + bs[A + j]; }
        }
    }
    if (it + 1 < p.iters) load_row(row + SPC);        // next group's hidden row flies during the sampling math
    group_sync<G>();
    if (row < T.rows && g == 0) {
        float* save = resolve(p.as, T.save, seed) + (long long)row * 4 * A;
        const int blk = row / B, b = row % B;
        float lp_acc = 0.f;
        for (int j = lane; j < A; j += 32) {
            const float mean_j = outv[j], raw_j = outv[A + j];
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
                // Normal(mean,std).log_prob(z) - log(1 - a^2 + eps)   (policies.py:147-160)
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
    if (it + 1 < p.iters) group_sync<G>();                 // outv is rewritten by the next group
    }

    pdl_trigger();
    if (p.tail_in_own_kernel) return;          // step_tail_kernel runs on a side lane (multi-lane schedule)
    // ---- last CTA of this seed: bump step counters, entropy-temperature Adam step ----
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int total = gdx * gdy;
        int prev = atomicAdd(&cnt[CNT_TICKET0], 1);
        s_last = (prev == total - 1);
        if (s_last) cnt[CNT_TICKET0] = 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    step_tail(p, seed);
}

template <int G>
__global__ void __launch_bounds__(GLUE_THREADS) policy_head_kernel(PolicyHeadParams p) {
    pdl_wait();
    policy_head_body<G>(p, p.use_external_eps, blockIdx.x, blockIdx.y, blockIdx.z, gridDim.x, gridDim.y);
}

__global__ void __launch_bounds__(GLUE_THREADS) step_tail_kernel(PolicyHeadParams p) {     // grid = seeds
    pdl_wait();
    step_tail(p, blockIdx.x);
}

// =====================================================================================
// critic heads + per-algorithm targets / loss gradients
// =====================================================================================
constexpr int MAX_HEAD_SRC = 40;
constexpr int MAX_VALS = 40;

struct HeadSrc {
    Ref h2;          // [*, H] hidden activations; sample b uses row row0 + b
    int row0;
    Ref w3, b3;      // [n_heads, H], [n_heads]
    int n_heads;
    Ref dq;          // [B, dq_ld] gradient w.r.t. the (pre-exp) head outputs, written here
    int dq_ld;       // row stride of dq (n_heads rounded up to 4)
    int write_dh2;   // also emit dh2[b,:] = (dq[b,:] W3) * (h2 > 0) with the CURRENT W3
    Ref dh2;         // [B, H]
};

enum CriticMode {
    CM_SAC = 0,         // srcs: qf1(a_pi) qf2(a_pi) qf1(data) qf2(data) tq1 tq2
    CM_POAC_Q = 1,      // srcs: qfs(data) x n, tfs(next) x n
    CM_POAC_PI = 2,     // srcs: qfs(a_pi) x n
    CM_GOAC_Q = 3,      // srcs: q(data) [std(data)] q_target(next) [std_target(next)]
    CM_GOAC_PI = 4      // srcs: q(a_pi) [std(a_pi)] q(a_tp) [std(a_tp)]
};

struct CriticHeadParams {
    HeadSrc src[MAX_HEAD_SRC];
    int n_src;
    int mode;
    ArenaSet as;
    long long off_rewards, off_terminals, off_counts, off_log_pi, off_q_pred, off_q_target, off_q_new, off_scalars;
    int B, H, P;          // P: particles (P-OAC) / heads per sample
    int nq;               // io row width of q_pred / q_target / q_new
    int n_nets;           // P-OAC / G-OAC: critic nets per group (1 shared, P or 2 separate)
    int share_layers, counts;
    float discount, reward_scale, standard_bound, std_init;
    int std_soft_update;  // P-OAC (:210-219)
    float std_soft_prob;
    int iters;            // sample groups (of GLUE_SPC samples) per CTA
    // (critic, head) pairs in evaluation order, built by the host: pair i is head pair_hd[i] of source pair_src[i];
    // goff[s] = first pair of source s
    int n_pairs;
    short pair_src[MAX_VALS], pair_hd[MAX_VALS];
    int goff[MAX_HEAD_SRC];
};

// Head dot products of one sample: the loads of up to four (critic, head) pairs are issued before the first
// reduction, so a warp keeps 4 x 2 x HR loads in flight (the kernel is bound by the latency of the h2 rows).
template <int HR, int G>
__device__ __forceinline__ void head_dots(const CriticHeadParams& p, int seed, int b, int g, int lane, int npairs,
                                          const short* pair_src, const short* pair_hd, float* vals) {
    constexpr int PB = (G == 1 ? 16 : 32) / HR;       // load registers: 64 (G = 4: few warps, deep batches), 32 (G = 1: occupancy)
    const int H = p.H;
    for (int base = g; base < npairs; base += G * PB) {
        float hv[PB][HR], wv[PB][HR];
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int pi = base + u * G;
            if (pi < npairs) {
                const HeadSrc& S = p.src[pair_src[pi]];
                const float* __restrict__ h = resolve(p.as, S.h2, seed) + (long long)(S.row0 + b) * H;
                const float* __restrict__ w = resolve(p.as, S.w3, seed) + (long long)pair_hd[pi] * H;
#pragma unroll
                for (int c = 0; c < HR; ++c) {
                    const int k = lane + 32 * c;
                    hv[u][c] = k < H ? ld_g(h + k) : 0.f;
                    wv[u][c] = k < H ? ld_g(w + k) : 0.f;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PB; ++u) {
            const int pi = base + u * G;
            if (pi < npairs) {
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < HR; ++c) acc = fmaf(hv[u][c], wv[u][c], acc);
                acc = warp_sum(acc);
                if (lane == 0) vals[pi] = acc + ld_g(resolve(p.as, p.src[pair_src[pi]].b3, seed) + pair_hd[pi]);
            }
        }
    }
}

// Fused critic head layer + targets / loss gradients + first backward step.  GLUE_G warps share one sample:
// the (critic, head) dot products are dealt round-robin to them (hidden row in registers, shuffle reduction),
// one lane then evaluates the algorithm's targets and dLoss/dq, and all GLUE_G warps emit
// dh2 = (dq W3) * relu'(h2) for the critics whose backward starts here.
template <int G>
__device__ __forceinline__ void critic_head_body(const CriticHeadParams& p, int bx, int by) {
    constexpr int SPC = GLUE_WARPS / G;
    __shared__ float s_vals[SPC][MAX_VALS];
    __shared__ float s_dq[SPC][MAX_VALS];
    const int* s_goff = p.goff;
    const short* s_pair_src = p.pair_src;
    const short* s_pair_hd = p.pair_hd;
    const int s_npairs = p.n_pairs;
    const int seed = by;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sl = warp / G, g = warp % G;
    const int B = p.B, H = p.H;
    for (int it = 0; it < p.iters; ++it) {                 // p.iters sample groups per CTA (many-seed launches)
    if (it > 0) group_sync<G>();                           // s_vals / s_dq are rewritten
    const int b = (bx * p.iters + it) * SPC + sl;
    const bool live = b < B;
    float* io = p.as.base[AR_IO] + (long long)seed * p.as.stride[AR_IO];
    float* vals = s_vals[sl];
    float* dqv = s_dq[sl];
    // per-sample scalars: issue the loads now, they are consumed after the head reductions
    float r_pre = 0.f, d_pre = 0.f, alpha_pre = 0.f, lpn_pre = 0.f, cnt_pre = 0.f;
    if (live && g == 0 && lane == 0) {
        r_pre = io[p.off_rewards + b]; d_pre = io[p.off_terminals + b];
        alpha_pre = io[p.off_scalars + SC_ALPHA]; lpn_pre = io[p.off_log_pi + B + b];
        cnt_pre = p.counts ? io[p.off_counts + b] : 0.f;
    }
    if (live) {
        if (H <= 256) head_dots<8, G>(p, seed, b, g, lane, s_npairs, s_pair_src, s_pair_hd, vals);
        else head_dots<GLUE_MAX_HR, G>(p, seed, b, g, lane, s_npairs, s_pair_src, s_pair_hd, vals);
    }
    group_sync<G>();
    // every dq goes to global (GEMM operand of the weight-gradient stage) and to the sample's shared copy
    auto put = [&](int s, int i, float v) {
        resolve(p.as, p.src[s].dq, seed)[(long long)b * p.src[s].dq_ld + i] = v;
        dqv[s_goff[s] + i] = v;
    };
    if (live && g == 0 && lane == 0) {
    const float invB = 1.0f / (float)B;
    const float r = r_pre;
    const float d = d_pre;
    const float nd = (1.f - d) * p.discount;       // (1 - terminals) * discount

    if (p.mode == CM_SAC) {
        const float alpha = alpha_pre;
        const float q1n = vals[0], q2n = vals[1], q1 = vals[2], q2 = vals[3], t1 = vals[4], t2 = vals[5];
        const float lp_next = lpn_pre;
        // trainer.py:178-184
        float tq = fminf(t1, t2) - alpha * lp_next;
        float y = p.reward_scale * r + nd * tq;
        io[p.off_q_pred + b * 2 + 0] = q1; io[p.off_q_pred + b * 2 + 1] = q2;
        io[p.off_q_target + b * 2 + 0] = y; io[p.off_q_target + b * 2 + 1] = y;
        io[p.off_q_new + b * 2 + 0] = q1n; io[p.off_q_new + b * 2 + 1] = q2n;
        // MSELoss mean over B: d/dq = 2 (q - y) / B          (trainer.py:194-195)
        put(2, 0, 2.f * (q1 - y) * invB);
        put(3, 0, 2.f * (q2 - y) * invB);
        // policy loss -mean(min(q1,q2)): gradient to the smaller one (first arg on ties)
        const bool sel1 = q1n <= q2n;
        put(0, 0, sel1 ? -invB : 0.f);
        put(1, 0, sel1 ? 0.f : -invB);
    } else if (p.mode == CM_POAC_Q) {
        // vals[0..P) current particles, vals[P..2P) target particles (particle_trainer_oac.py:185-208)
        const int P = p.P;
        const float* q = vals;
        const float* tq = vals + P;
        float sq[16], st[16];
        int rank_q[16];
        for (int i = 0; i < P; ++i) {
            int rq = 0, rt = 0;
            for (int j = 0; j < P; ++j) {
                rq += (q[j] < q[i]) || (q[j] == q[i] && j < i);
                rt += (tq[j] < tq[i]) || (tq[j] == tq[i] && j < i);
            }
            rank_q[i] = rq; sq[rq] = q[i]; st[rt] = tq[i];
        }
        float T[16];
        float mean_sq = 0.f, mean_T = 0.f;
        for (int i = 0; i < P; ++i) {
            T[i] = p.reward_scale * r + nd * st[i];
            mean_sq += sq[i]; mean_T += T[i];
        }
        if (p.std_soft_update) {      // :210-219: keep the current spread, move the mean (never together with counts, :97)
            const float msq = mean_sq / (float)P, mT = mean_T / (float)P;
            for (int i = 0; i < P; ++i) T[i] = p.std_soft_prob * T[i] + (1.f - p.std_soft_prob) * (sq[i] - msq + mT);
        }
        if (p.counts) {      // :220-224
            mean_sq /= (float)P; mean_T /= (float)P;
            const float f = cnt_pre == 0.f ? 1.f : 0.f;
            for (int i = 0; i < P; ++i) T[i] = T[i] * f + (1.f - f) * (sq[i] - mean_sq + mean_T);
        }
        for (int i = 0; i < P; ++i) {
            io[p.off_q_pred + (long long)b * p.nq + i] = sq[i];
            io[p.off_q_target + (long long)b * p.nq + i] = T[i];
        }
        // each head regresses to the target at its current rank; with separate nets a net only
        // receives gradient where it sits at its own rank (:247-264)
        for (int i = 0; i < P; ++i) {
            float gq = 2.f * (q[i] - T[rank_q[i]]) * invB;
            if (p.share_layers) put(0, i, gq);
            else put(i, 0, (rank_q[i] == i) ? gq : 0.f);
        }
    } else if (p.mode == CM_POAC_PI) {
        // policy loss uses the lowest particle (:291-295)
        const int P = p.P;
        int best = 0;
        for (int i = 1; i < P; ++i) if (vals[i] < vals[best]) best = i;
        for (int i = 0; i < P; ++i) {
            io[p.off_q_new + (long long)b * p.nq + i] = vals[i];
            float gq = (i == best) ? -invB : 0.f;
            if (p.share_layers) put(0, i, gq);
            else put(i, 0, gq);
        }
    } else if (p.mode == CM_GOAC_Q) {
        // shared: vals = q0, raw1 | tq0, traw1 ; separate: q0 | raw1 | tq0 | traw1  (same order)
        const float q0 = vals[0], sig = expf(vals[1]), t0 = vals[2], tsig = expf(vals[3]);
        float std_t = nd * tsig;                                    // gaussian_trainer.py:217
        if (p.counts) {
            const float f = cnt_pre == 0.f ? 1.f : 0.f;
            std_t = std_t * f + (1.f - f) * sig;                    // :224-228
        }
        const float y = p.reward_scale * r + nd * t0;               // :231-232
        std_t = fminf(fmaxf(std_t, 0.f), p.std_init);               // :233
        io[p.off_q_pred + b * 2 + 0] = q0; io[p.off_q_pred + b * 2 + 1] = sig;
        io[p.off_q_target + b * 2 + 0] = y; io[p.off_q_target + b * 2 + 1] = std_t;
        const float g0 = 2.f * (q0 - y) * invB;
        const float g1 = 2.f * (sig - std_t) * invB * sig;          // through exp
        if (p.share_layers) { put(0, 0, g0); put(0, 1, g1); }
        else { put(0, 0, g0); put(1, 0, g1); }
    } else {   // CM_GOAC_PI
        // policy: -(q + z*sigma).mean() (:344-356); target policy: -(q).mean() (:361-373)
        const float q0 = vals[0], sig = expf(vals[1]);
        io[p.off_q_new + b * 2 + 0] = q0; io[p.off_q_new + b * 2 + 1] = sig;
        const float g0 = -invB, g1 = -invB * p.standard_bound * sig;
        if (p.share_layers) { put(0, 0, g0); put(0, 1, g1); put(1, 0, -invB); put(1, 1, 0.f); }
        else { put(0, 0, g0); put(1, 0, g1); put(2, 0, -invB); put(3, 0, 0.f); }
    }
    }   // lane 0
    group_sync<G>();
    if (it + 1 == p.iters) pdl_trigger();
    // ---- dh2 = (dq W3) * relu'(h2) for the critics whose backward starts with the current weights ----
    for (int s = 0; s < p.n_src && live; ++s) {
        const HeadSrc& S = p.src[s];
        if (!S.write_dh2) continue;
        const float* __restrict__ h = resolve(p.as, S.h2, seed) + (long long)(S.row0 + b) * H;
        const float* __restrict__ w = resolve(p.as, S.w3, seed);
        float* __restrict__ out = resolve(p.as, S.dh2, seed) + (long long)b * H;
        const int g0 = s_goff[s];
#pragma unroll 4
        for (int k = g * 32 + lane; k < H; k += 32 * G) {
            const float hv = ld_g(h + k);
            float acc = 0.f;
            for (int hd = 0; hd < S.n_heads; ++hd) acc = fmaf(dqv[g0 + hd], ld_g(w + (long long)hd * H + k), acc);
            out[k] = hv > 0.f ? acc : 0.f;
        }
    }
    }   // it
}

template <int G>
__global__ void __launch_bounds__(GLUE_THREADS, G == 1 ? 4 : 2) critic_head_kernel(const CriticHeadParams* __restrict__ pp) {
    pdl_wait();
    critic_head_body<G>(*pp, blockIdx.x, blockIdx.y);
}

// =====================================================================================
// policy-loss gradient w.r.t. the policy head outputs
// =====================================================================================
struct PolicyGradTask {
    Ref dh1[20];      // per critic: [B, H] gradient at its first hidden layer (rows of this policy's actions)
    Ref w1[20];       // per critic: fc0.weight [H, ld]; the action columns start at O
    int ld[20];
    int n_src;
    Ref save;         // [., 4, A] from policy_head (rows of this policy start at save_row0)
    int save_row0;
    Ref dhead;        // out: [B, dhead_ld]  d loss / d(mean), d loss / d(raw log_std)
    int dhead_ld;     // row stride (2A rounded up to 4)
    int entropy;      // 1: alpha*log_pi term present (stochastic policy)
    Ref wh;           // policy head weights [2A, H]
    Ref h2;           // policy second hidden activation [*, H]; rows of this batch start at h2_row0
    int h2_row0;
    Ref dhp2;         // out: [B, H] = (dhead Wh) * relu'(h2)
    Ref da[20];       // PolicyGradParams::da_from_gemm: per critic [B, da_ld] = dh1 W0[:, O:O+A] from a GEMM stage
    int da_ld;
};

struct PolicyGradParams {
    const PolicyGradTask* tasks;
    ArenaSet as;
    long long off_scalars;
    int O, A, H, B;
    int iters;        // sample groups (of GLUE_SPC samples) per CTA
    int da_from_gemm; // 1: dQ/da and the policy's dh2 run as tensor-core GEMM stages; this kernel is the chain rule only
    int wa_sources;   // critics whose fc0 action columns are staged together (1 or 2)
};

// Fused: dLoss/d(action) through every critic's first layer (only the A action columns of fc0.weight are
// needed, not the full 393-wide dX the reference's autograd computes), the TanhNormal / entropy chain rule
// to the policy head outputs, and the policy's first backward step dh2 = (dhead Wh) * relu'(h2).
// GLUE_G warps share one sample.  dyn smem: Wa [H*AS] | Wh [2A*H] | ga [SPC][A] | dhead [SPC][2A]  (AS = A | 1)
template <int G>
__device__ __forceinline__ void policy_grad_body(const PolicyGradParams& p, int bx, int by, int bz) {
    constexpr int SPC = GLUE_WARPS / G;
    extern __shared__ __align__(16) float s_pg[];
    const PolicyGradTask& T = p.tasks[by];
    const int seed = bz;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sl = warp / G, g = warp % G;
    const int A = p.A, H = p.H, B = p.B, O = p.O;
    const int AS = A | 1;                                   // odd stride: conflict-free column reads
    const int iters = p.iters;                              // sample groups per CTA: the staged weights serve all of them
    const int b0 = bx * iters * SPC + sl;      // this warp's sample in group it: b0 + it*SPC
    const bool from_gemm = p.da_from_gemm != 0;             // dyn smem is then only the sga / sdh rows
    float* Wa = s_pg;
    float* Whs = Wa + (from_gemm ? 0 : p.wa_sources * H * AS);
    float* sga_all = Whs + (from_gemm ? 0 : 2 * A * H);     // [iters][SPC][A]
    float* sdh_all = sga_all + iters * SPC * A;        // [iters][SPC][2A]
    float* io = p.as.base[AR_IO] + (long long)seed * p.as.stride[AR_IO];
    if (from_gemm) {
        for (int it = 0; it < iters; ++it) {
            const int b = b0 + it * SPC;
            if (b < B && g == 0) {
                float* sga = sga_all + (it * SPC + sl) * A;
                for (int j = lane; j < A; j += 32) {
                    float acc = 0.f;
                    for (int s = 0; s < T.n_src; ++s) acc += ld_g(resolve(p.as, T.da[s], seed) + (long long)b * T.da_ld + j);
                    sga[j] = acc;
                }
            }
        }
    } else stage_contig(Whs, resolve(p.as, T.wh, seed), 2 * A * H);       // consumed in the last phase
    // dQ/da: the action columns of up to two critics' fc0.weight are staged together (the twin-critic case: one L2 round
    // trip instead of two), more sources take turns.  HR = hidden / 32 registers per lane hold a dh1 row.
    const int n_stage = p.wa_sources;                       // 1 or 2 (host: min(n_src, 2) when the shared memory fits)
    auto dq_da = [&](auto hr_tag) {
        constexpr int HR = decltype(hr_tag)::value;
        for (int s0 = 0; s0 < T.n_src; s0 += n_stage) {
            const int ns = min(n_stage, T.n_src - s0);
            __syncthreads();
            for (int u = 0; u < ns; ++u) {
                const float* w1 = resolve(p.as, T.w1[s0 + u], seed);
                const int ld = T.ld[s0 + u];
                float* Wu = Wa + u * H * AS;
                // element (n, j) -> Wu[n*AS + j], flat over the H*A elements with every lane busy; (n, j) advance
                // incrementally (one division per thread, none per element)
                const int dn = GLUE_THREADS / A, dj = GLUE_THREADS - dn * A;
                int n = (int)threadIdx.x / A, j = (int)threadIdx.x - n * A;
                for (; n < H; n += dn, j += dj) {
                    if (j >= A) { j -= A; ++n; if (n >= H) break; }
                    cp_async4(Wu + n * AS + j, w1 + (long long)n * ld + O + j);
                }
            }
            float dreg[2][HR];
            auto load_rows = [&](int b) {
                if (b < B) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        if (u < ns) {
                            const float* dh = resolve(p.as, T.dh1[s0 + u], seed) + (long long)b * H;
#pragma unroll
                            for (int c = 0; c < HR; ++c) dreg[u][c] = (lane + 32 * c < H) ? ld_g(dh + lane + 32 * c) : 0.f;
                        }
                    }
                }
            };
            load_rows(b0);
            cp_async_wait_all();
            __syncthreads();
            for (int it = 0; it < iters; ++it) {
                const int b = b0 + it * SPC;
                float* sga = sga_all + (it * SPC + sl) * A;
                if (b < B) {
                    for (int j = g; j < A; j += G) {
                        float acc = 0.f;
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            if (u < ns) {
                                const float* Wu = Wa + u * H * AS + j;
                                float au = 0.f;                 // per-critic sums are added in source order, as before
#pragma unroll
                                for (int c = 0; c < HR; ++c) {
                                    const int n = lane + 32 * c;
                                    if (HR * 32 <= 256 || n < H) au = fmaf(dreg[u][c], Wu[n * AS], au);
                                }
                                au = warp_sum(au);
                                acc = (u == 0) ? au : acc + au;
                            }
                        }
                        if (lane == 0) sga[j] = (s0 == 0 ? 0.f : sga[j]) + acc;  // column j belongs to warp g only
                    }
                }
                if (it + 1 < iters) load_rows(b + SPC);
            }
        }
    };
    if (!from_gemm) {
        if (H == 256) dq_da(std::integral_constant<int, 8>());
        else dq_da(std::integral_constant<int, GLUE_MAX_HR>());
    }
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        const int b = b0 + it * SPC;
        if (b < B && g == 0) {
            const float* sga = sga_all + (it * SPC + sl) * A;
            float* sdh = sdh_all + (it * SPC + sl) * 2 * A;
            const float alpha = io[p.off_scalars + SC_ALPHA];
            const float invB = 1.0f / (float)B;
            const float* save = resolve(p.as, T.save, seed) + (long long)(T.save_row0 + b) * 4 * A;
            float* dhead = resolve(p.as, T.dhead, seed) + (long long)b * T.dhead_ld;
            for (int j = lane; j < A; j += 32) {
                const float a = save[0 * A + j];
                const float one_m_a2 = 1.f - a * a;
                const float gaj = sga[j];
                float dmean, draw;
                if (T.entropy) {
                    const float std = save[1 * A + j], raw = save[2 * A + j], eps = save[3 * A + j];
                    const float u = one_m_a2 + TANH_EPS_F;
                    // dL/dz: alpha/B * d(-log(1-a^2+eps))/dz  +  dL/da * (1-a^2)      
                    dmean = alpha * invB * (2.f * a * one_m_a2 / u) + gaj * one_m_a2;
                    const float dstd = dmean * eps - alpha * invB / std;
                    const bool inside = (raw >= LOG_SIG_MIN_F) && (raw <= LOG_SIG_MAX_F);
                    draw = inside ? dstd * std : 0.f;
                } else {
                    dmean = gaj * one_m_a2;    // a = tanh(mean); log_std head receives no gradient
                    draw = 0.f;
                }
                dhead[j] = dmean; dhead[A + j] = draw;
                sdh[j] = dmean; sdh[A + j] = draw;
            }
        }
    }
    pdl_trigger();
    if (from_gemm) return;                                  // dh2 is a GEMM stage of its own
    __syncthreads();
    // ---- policy backward, first step: dh2 = (dhead Wh) * relu'(h2) ----
    for (int it = 0; it < iters; ++it) {
        const int b = b0 + it * SPC;
        if (b >= B) break;
        const float* sdh = sdh_all + (it * SPC + sl) * 2 * A;
        const float* __restrict__ hp2 = resolve(p.as, T.h2, seed) + (long long)(T.h2_row0 + b) * H;
        float* __restrict__ out = resolve(p.as, T.dhp2, seed) + (long long)b * H;
#pragma unroll 2
        for (int n = g * 32 + lane; n < H; n += 32 * G) {
            const float hv = ld_g(hp2 + n);
            float acc = 0.f;
            for (int j = 0; j < 2 * A; ++j) acc = fmaf(sdh[j], Whs[j * H + n], acc);
            out[n] = hv > 0.f ? acc : 0.f;
        }
    }
}

template <int G>
__global__ void __launch_bounds__(GLUE_THREADS) policy_grad_kernel(PolicyGradParams p) {
    pdl_wait();
    policy_grad_body<G>(p, blockIdx.x, blockIdx.y, blockIdx.z);
}

}  // namespace oac
