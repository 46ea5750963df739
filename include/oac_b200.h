/*
 * oac_b200.h -- C ABI of liboac_b200.so: the B200 (sm_100a) implementation of the
 * per-gradient-step hot path of amarildolikmeta/oac-explore.
 *
 * Every entry point replaces one reference interface (paths relative to the
 * reference root):
 *
 *   oac_replay_gather / oac_replay_add      replay_buffer.py:88-115, 167-197
 *                                           (ReplayBuffer.random_batch / add_sample,
 *                                            ReplayBufferCount counts) + utils/core.py:40-61
 *   oac_trainer_create / _step / _destroy   trainer/trainer.py:126-224 (SACTrainer.train_from_torch),
 *                                           trainer/particle_trainer_oac.py:169-324,
 *                                           trainer/gaussian_trainer.py:177-388,
 *                                           torch.optim.Adam as used at trainer/trainer.py:75-91,
 *                                           utils/pytorch_util.py:5-9 (soft_update_from_to)
 *   oac_explore                             optimistic_exploration.py:14-196
 *   oac_policy_forward / oac_q_forward      trainer/policies.py:260-316, networks.py:62-79,154-161
 *                                           (inference calls: policy.get_action, trainer.predict)
 *
 * Conventions
 *   - plain C, no torch types.  All pointers are DEVICE pointers unless named host_*.
 *   - ownership: the caller (PyTorch on the Python side) allocates and owns every
 *     buffer: weights, Adam moments, counters, replay store, batches, workspace.
 *     The library borrows them for the duration of a call (or of a trainer handle)
 *     and allocates only small descriptor tables at *_create time, never on the
 *     step path.
 *   - every function returns 0 on success, a cudaError_t value or a negative
 *     OAC_E_* code otherwise; oac_last_error_string() describes the last failure
 *     of the calling thread.  No C++ exception crosses the boundary.
 *   - calls are asynchronous on the caller-supplied stream (a cudaStream_t passed
 *     as void*); a handle must not be used from two streams at once.
 *   - all arithmetic is fp32 (the reference casts every batch with .float(),
 *     utils/pytorch_util.py:76-77); OacConfig.gemm_path selects SIMT fp32 or the tcgen05 kind::tf32 paths.
 */
#ifndef OAC_B200_H
#define OAC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OAC_ABI_VERSION 4
#define OAC_HOST_SCALAR_SLOTS 2
#define OAC_MAX_NETS 48

enum { OAC_E_INVALID = -1, OAC_E_UNSUPPORTED = -2, OAC_E_NOMEM = -3 };

enum OacAlgo { OAC_ALGO_SAC = 0, OAC_ALGO_POAC = 1, OAC_ALGO_GOAC = 2 };
/* FP32: SIMT FFMA.  TF32: tcgen05 kind::tf32, one MMA per product (rel ~1e-3).  TF32X3: tcgen05 with the
 * 3xTF32 operand split (a_hi b_hi + a_hi b_lo + a_lo b_hi), fp32-grade accuracy on the tensor pipe. */
enum OacGemmPath { OAC_GEMM_FP32 = 0, OAC_GEMM_TF32 = 1, OAC_GEMM_TF32X3 = 2 };
enum OacNetKind { OAC_NET_POLICY = 0, OAC_NET_Q = 1, OAC_NET_SCALAR = 2 };

/* ---- trainer configuration (mirrors the reference constructors' kwargs) ---- */
typedef struct OacConfig {
    int32_t algo;                /* OacAlgo */
    int32_t obs_dim, act_dim;    /* O, A */
    int32_t hidden;              /* H: two hidden layers of width H (main.py num_layers=2) */
    int32_t batch;               /* B */
    int32_t n_seeds;             /* independent trainers batched in one handle (>=1) */
    int32_t n_particles;         /* P-OAC: n_estimators; others: ignored */
    int32_t share_layers;        /* P-OAC / G-OAC: one trunk with P (resp. 2) heads */
    int32_t deterministic;       /* policy acts tanh(mean): no sampling, log_pi = 0 */
    int32_t auto_alpha;          /* use_automatic_entropy_tuning */
    int32_t counts;              /* batch carries ReplayBufferCount counts */
    int32_t train_bias;          /* last-layer critic bias trainable (q_producer train_bias) */
    int32_t stale_graph_mode;    /* SAC only: 0 = "A" torch-1.4-literal (post-step Q weights in
                                    the policy-loss dX), 1 = "B" pre-step weights */
    int32_t target_update_period;
    int32_t gemm_path;           /* OacGemmPath */
    int32_t std_soft_update;     /* P-OAC: targets = p * next + (1 - p) * (current - mean(current) + mean(next))
                                    over the particle axis (trainer/particle_trainer_oac.py:210-219) */
    float discount, reward_scale, soft_target_tau;
    float policy_lr, qf_lr, std_lr;
    float target_entropy;
    float standard_bound;        /* G-OAC: norm.ppf(delta) */
    float std_init;              /* G-OAC: (q_max-q_min)/sqrt(12) */
    float adam_beta1, adam_beta2, adam_eps;
    float std_soft_update_prob;  /* P-OAC std_soft_update: p */
    float reserved2;
    uint64_t rng_seed;           /* device Philox stream when eps == NULL */
} OacConfig;

/* One MLP inside the per-seed parameter arena.  Offsets are in floats from the start
 * of the seed's arena slice.  Weights are [out, ld] row-major (y = x W^T + b) exactly
 * like nn.Linear; fc0.weight rows are padded from in_dim to in_ld floats (pad = 0). */
typedef struct OacNetLayout {
    int32_t kind;                /* OacNetKind */
    int32_t in_dim, in_ld;       /* fc0: [hidden, in_ld], first in_dim columns live */
    int32_t hidden;
    int32_t n_out;               /* head rows: 2A for a policy (mean rows, then log_std rows),
                                    n heads for a critic */
    int32_t trainable;           /* has Adam state (arena prefix) */
    int64_t off_w0, off_b0, off_w1, off_b1, off_w2, off_b2;
    int64_t size;                /* floats, multiple of 4 */
} OacNetLayout;

/* Memory plan for one configuration; filled by oac_trainer_layout().
 * Net order: SAC   policy, qf1, qf2, log_alpha | target_qf1, target_qf2
 *            P-OAC policy, qf[0..n), log_alpha | tf[0..n)      (n = 1 if share_layers else P)
 *            G-OAC policy, target_policy, q, [std], log_alpha | q_target, [std_target]
 * Nets before '|' are trainable and form the arena prefix covered by the Adam arenas. */
typedef struct OacLayout {
    int32_t n_nets;
    int32_t n_trainable;
    OacNetLayout nets[OAC_MAX_NETS];
    int64_t param_floats;        /* per seed */
    int64_t adam_floats;         /* per seed: trainable prefix */
    int64_t work_floats;         /* per seed: activations / gradients scratch */
    int64_t io_floats;           /* per seed: batch rows + rewards/terminals/counts + eps + outputs */
    int32_t n_counters;          /* per seed int32 counters (optimizer steps, train steps, ...) */
    int32_t x_rows, x_ld;        /* batch-row matrix X [x_rows, x_ld] at io offset off_x */
    /* offsets into the per-seed IO slice (floats) */
    int64_t off_x;               /* X [4B, x_ld]: block 0 [obs|a_tp] 1 [obs|a_pi] 2 [obs|actions] 3 [next_obs|a_next] */
    int64_t off_rewards, off_terminals, off_counts;      /* [B] each */
    int64_t off_eps;             /* [2, B, A]: slot 0 -> policy(obs) draw, slot 1 -> policy(next_obs) draw */
    int64_t off_log_pi;          /* [3B]: rows [0,B) policy(obs), [B,2B) policy(next_obs), [2B,3B) target_policy(obs) */
    int64_t off_mean, off_log_std;   /* [3B, A] each, same row order */
    int64_t off_q_pred;          /* [B, nq]  critic outputs on (obs, actions) (sorted for P-OAC) */
    int64_t off_q_target;        /* [B, nq]  regression targets */
    int64_t off_q_new;           /* [B, nq]  critic outputs on (obs, a_pi) */
    int64_t off_scalars;         /* [16]: 0 alpha, 1 alpha_loss, 2 mean log_pi(obs) ... */
    int32_t nq;                  /* critic outputs per sample (2 SAC, P P-OAC, 2 G-OAC) */
    int32_t reserved1;
} OacLayout;

typedef struct OacBuffers {
    float* params;    /* [n_seeds, param_floats] */
    float* adam_m;    /* [n_seeds, adam_floats]  */
    float* adam_v;    /* [n_seeds, adam_floats]  */
    float* work;      /* [n_seeds, work_floats]  */
    float* io;        /* [n_seeds, io_floats]    */
    int32_t* counters;/* [n_seeds, n_counters]   */
    float* host_scalars; /* optional (may be NULL): [OAC_HOST_SCALAR_SLOTS, n_seeds, 16] MAPPED PINNED HOST memory; every
                          * step also stores its scalars (0 alpha, 1 alpha loss, 2 mean log pi, 3 the 1-based number of the
                          * step that wrote them) into slot (steps taken before it) % 2, so a caller that wants the step's
                          * result on the host only has to wait for the step (no device-to-host copy call) and may keep
                          * one further step in flight while it reads (the training loop of rl_algorithm.py:159-165 never
                          * blocks on a step) */
} OacBuffers;

typedef struct OacTrainer OacTrainer;   /* opaque */

const char* oac_last_error_string(void);
int oac_abi_version(void);

/* ---- replay buffer: replay_buffer.py ---- */
/* Device store, struct-of-arrays fp32 (cast once at insert time instead of per batch,
 * utils/core.py:45): obs [N,O], next_obs [N,O], actions [N,A], rewards [N], terminals [N]
 * (0/1 as float), counts [N] or NULL.  oac_replay_gather writes the batch straight into a
 * trainer's X matrix / IO slice (X is [4B, x_ld], four blocks of B rows:
 *   block 0 [obs|a_target_policy] (G-OAC), 1 [obs|a_pi], 2 [obs|actions], 3 [next_obs|a_next];
 * the action columns of blocks 0, 1, 3 are filled by the step itself) and, when
 * counts != NULL, returns counts[idx] then increments each DISTINCT index once
 * (numpy fancy "+=" semantics, replay_buffer.py:193-195). */
typedef struct OacReplayStore {
    const float* obs; const float* next_obs; const float* actions;
    const float* rewards; const float* terminals; float* counts;
    int64_t capacity; int32_t obs_dim, act_dim;
} OacReplayStore;

typedef struct OacBatchDst {
    float* x; int32_t x_ld;      /* seed 0's X matrix (io + off_x) */
    int32_t obs_blocks[3];       /* X row blocks (units of `batch` rows) receiving obs[idx]; -1 = unused */
    int32_t act_block;           /* block whose columns [O,O+A) receive actions[idx] */
    int32_t next_block;          /* block receiving next_obs[idx] */
    float* rewards; float* terminals; float* counts;   /* seed 0's [B] slots (io + off_*) */
    int32_t n_seeds;             /* indices is [n_seeds, batch]; seed s writes at + s*seed_stride floats */
    int32_t reserved;
    int64_t seed_stride;
} OacBatchDst;

int oac_replay_gather(const OacReplayStore* store, const int64_t* indices, int32_t batch,
                      const OacBatchDst* dst, void* stream);
/* Plain gather into five dense tensors (ReplayBuffer.random_batch "fast" return). */
int oac_replay_gather_dense(const OacReplayStore* store, const int64_t* indices, int32_t batch,
                            float* obs, float* actions, float* rewards, float* terminals,
                            float* next_obs, float* counts_out, void* stream);
/* add_sample for n consecutive transitions starting at ring slot `top` (wraps at capacity).
 * rows: packed [n, 2*O + A + 2] = obs | action | reward | terminal | next_obs (fp32, device).
 * Zeroes counts of the written slots (replay_buffer.py:177). */
int oac_replay_add(float* obs, float* next_obs, float* actions, float* rewards, float* terminals,
                   float* counts, int64_t capacity, int32_t obs_dim, int32_t act_dim,
                   const float* rows, int32_t n, int64_t top, void* stream);

/* ---- trainers ---- */
int oac_trainer_layout(const OacConfig* cfg, OacLayout* out);
int oac_trainer_create(const OacConfig* cfg, const OacBuffers* buf, OacTrainer** out);
int oac_trainer_destroy(OacTrainer* t);
/* One train_from_torch on every seed of the handle.  The batch must already sit in the IO
 * slice (oac_replay_gather or a caller copy).  use_external_eps != 0: the N(0,1) draws are
 * read from io[off_eps] (parity mode); 0: generated on device (Philox, rng_seed, step). */
int oac_trainer_step(OacTrainer* t, int32_t use_external_eps, void* stream);
/* Number of kernel launches one step issues (for bench.py's gpu_launches). */
int oac_trainer_launches_per_step(const OacTrainer* t);
/* Number of GEMM stages of the step that run on the warp-specialised TMA + tcgen05 kernel (tests assert the
 * tensor-core path is the one measured). */
int oac_trainer_ws_stages(const OacTrainer* t);
/* On-device diagnostics (trainer/trainer.py:230-279, particle_trainer_oac.py:332-362, gaussian_trainer.py:397-436): one
 * kernel reduces the last step's per-sample outputs into the `eval_statistics` vector of every seed, out[seed * out_ld + i],
 * in the reference's key order:
 *   SAC   [32]      QF mean, QF std, QF1 Loss, QF2 Loss, Q Loss, Policy Loss, {Q1 Predictions, Q2 Predictions, Q Targets,
 *                   Log Pis, Policy mu, Policy log std} x {Mean, Std, Max, Min}, Alpha, Alpha Loss
 *   P-OAC [11 + 9P] QF mean, QF std, per particle {QFi Loss, QiPredictions x4, QiTargets x4}, Policy Loss, Policy mu x4,
 *                   Policy log std x4
 *   G-OAC [29]      QF mean, QF std, QF Loss, Q Predictions x4, Q Target x4, STD Loss, Q STD Predictions x4,
 *                   Q STD Target x4, Policy Loss, Policy mu x4, Policy log std x4
 * `out` is device memory or mapped pinned host memory (then a stream synchronisation is all the caller needs); it is
 * also the payload of the per-seed statistics all-gather.  oac_trainer_stats_count returns the vector length. */
int oac_trainer_stats_count(const OacTrainer* t);
int oac_trainer_stats(OacTrainer* t, float* out, int32_t out_ld, void* stream);
/* Measurement aid: runs `iters` steps stage by stage (no graph) with a CUDA event between
 * stages and returns the mean duration of each stage in milliseconds (ms[n_stages]) and, per
 * stage, whether it is a GEMM stage (is_gemm) and its algorithmic FLOPs per seed (flops).
 * names: n_stages pointers to static strings.  Synchronises the stream. */
int oac_trainer_profile(OacTrainer* t, int32_t iters, int32_t max_stages, float* ms, int32_t* is_gemm,
                        double* flops, const char** names, int32_t* n_stages, void* stream);

/* Test / measurement aid: one GEMM C[M,N] = act(sum_k A(m,k) B(n,k) + bias[n]) through the fp32 SIMT stage
 * kernel (gemm_path 0) or the tcgen05 kind::tf32 one (1).  a_trans: A(m,k) = A[k*lda+m] else A[m*lda+k];
 * b_trans: B(n,k) = B[k*ldb+n] else B[n*ldb+k].  Synchronises the stream. */
int oac_gemm_debug(int32_t gemm_path, int32_t a_trans, int32_t b_trans, int32_t M, int32_t N, int32_t K,
                   const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc,
                   const float* bias, int32_t relu, void* stream);
/* Which kernel the last oac_gemm_debug call ran: 0 SIMT, 1 per-tile tcgen05, 2 warp-specialised TMA + tcgen05,
 * 3 its CTA-pair variant (tcgen05.mma.cta_group::2). */
int oac_gemm_debug_kernel(void);

/* ---- inference ---- */
/* TanhGaussianPolicy.forward on n rows: obs [n, obs_ld]; eps [n,A] or NULL (deterministic).
 * Outputs (any may be NULL): action, mean, log_std, std, pre_tanh [n,A]; log_prob [n]. */
int oac_policy_forward(const float* net, const OacNetLayout* lay, const float* obs, int32_t obs_ld,
                       int32_t n, const float* eps, float* action, float* mean, float* log_std,
                       float* std, float* pre_tanh, float* log_prob, void* stream);
/* FlattenMlp.forward on n rows of x = [obs|act] with leading dim x_ld -> out [n, n_out].
 * exp_mask bit i set: out[:, i] = exp(out[:, i]) (networks.py:69-75 "positive"). */
int oac_q_forward(const float* net, const OacNetLayout* lay, const float* x, int32_t x_ld, int32_t n,
                  uint32_t exp_mask, float* out, void* stream);

/* ---- optimistic exploration: optimistic_exploration.py:14-196 ---- */
enum OacExploreMode {
    OAC_EXPLORE_TWIN = 0,        /* (Q1+Q2)/2 + beta*|Q1-Q2|/2 on q[0], q[1]   (:42-46, trainer.predict :105-123) */
    OAC_EXPLORE_ENSEMBLE = 1,    /* mean + beta*std(unbiased) over all heads of all nets (:47-58) */
    OAC_EXPLORE_QUANTILE = 2     /* sorted particle `quantile_index` (ParticleTrainer.predict) */
};
typedef struct OacExploreArgs {
    const float* policy; OacNetLayout policy_lay;
    const float* q[OAC_MAX_NETS]; OacNetLayout q_lay; int32_t n_q;
    int32_t mode;                /* OacExploreMode */
    int32_t deterministic;       /* L2-normalised shift, returns un-squashed mu_E (:111-196) */
    int32_t quantile_index;
    uint32_t exp_mask;           /* bit v set: critic OUTPUT v passes through exp (networks.py:69-75 "positive"); outputs are
                                    numbered net-major, v = net * n_heads + head, so every critic carries its own flags
                                    (G-OAC shared net: 0b10; separate mean / std nets: 0b10 as well) */
    float beta_UB, delta;
    int32_t n_obs;               /* independent observations (1 in the reference's rollout) */
    const float* obs;            /* [n_obs, obs_dim] */
    const float* eps;            /* [n_obs, A] N(0,1) draws for the final sample, or NULL: Philox */
    uint64_t rng_seed, rng_offset;
    float* action;               /* [n_obs, A]: tanh(N(mu_E, std)) or mu_E (deterministic) */
    float* mu_E;                 /* [n_obs, A] or NULL */
    float* grad;                 /* [n_obs, A] dQ_UB/d(pre-tanh mean) or NULL */
    /* Per-seed batched exploration (path_collector.py:214-232 vectorised over the seeds of one GPU): when
     * obs_group != NULL observation i is served by the parameter arena obs_group[i] -- `policy` and `q[]` point into arena 0
     * and arena g starts group_stride floats further (the [n_seeds, param_floats] arena of a seed group) -- so ONE launch
     * computes every seed's action with that seed's own policy and critics. */
    const int32_t* obs_group;    /* [n_obs] device, or NULL: every observation uses `policy` / `q[]` as given */
    int64_t group_stride;        /* floats */
} OacExploreArgs;
int oac_explore(const OacExploreArgs* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OAC_B200_H */
