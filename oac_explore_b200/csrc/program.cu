// Host-side "program" builder: turns an OacConfig into the stage list of one
// train_from_torch (GEMM stages + glue kernels), uploads the task tables once and
// replays the whole step as one CUDA graph.
//
// Stage order follows the reference's update order where it defines the numerics
// (SURVEY.md section 3.3 / 8c):
//   SAC   : policy fwd(obs,next) -> alpha step -> Q fwd x6 -> targets/dq -> Q bwd ->
//           Q Adam(+Polyak) -> policy-loss dX through the (mode A: post-step) Q weights ->
//           policy bwd -> policy Adam.                         trainer/trainer.py:126-224
//   P-OAC : Q fwd(data), policy(next) , target fwd, sort, per-rank targets -> Q bwd/Adam ->
//           policy(obs), alpha step, fresh Q fwd with updated weights, min particle ->
//           policy bwd/Adam.                                   trainer/particle_trainer_oac.py:169-324
//   G-OAC : mean/std critic regression -> Adam -> upper-bound policy and mean target-policy
//           updates through the updated critic.               trainer/gaussian_trainer.py:177-388
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "glue.cuh"
#include "glue_many.cuh"
#include "gemm_tc.cuh"
#include "gemm_ws.cuh"
#include "gemm_ws2.cuh"
#include "gemm_chain.cuh"
#include "gemm_fwd2.cuh"
#include "adam_stream.cuh"
#include "stats.cuh"
#include "oac_error.h"

namespace oac {

static int sm_count();
constexpr int OAC_E_CHAIN_UNAVAILABLE = -101;     // internal: finalize() asks for a rebuild without strip-fused forward chains
constexpr int OAC_E_BITS_UNAVAILABLE = -102;      // internal: finalize() asks for a rebuild without sign-bit masks
constexpr int OAC_E_SPLIT_UNAVAILABLE = -100;      // internal: finalize() asks for a rebuild without gradient-store stages
static inline int pad4(int x) { return (x + 3) & ~3; }
static inline long long pad4ll(long long x) { return (x + 3) & ~3ll; }

// ------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------
static void net_layout(OacNetLayout& n, int kind, int in_dim, int hidden, int n_out, int trainable,
                       long long& cursor) {
    memset(&n, 0, sizeof(n));
    n.kind = kind; n.in_dim = in_dim; n.in_ld = pad4(in_dim); n.hidden = hidden; n.n_out = n_out;
    n.trainable = trainable;
    long long start = cursor;
    if (kind == OAC_NET_SCALAR) {
        n.off_w0 = cursor; cursor += 4;
        n.off_b0 = n.off_w1 = n.off_b1 = n.off_w2 = n.off_b2 = n.off_w0;
    } else {
        n.off_w0 = cursor; cursor += pad4ll((long long)hidden * n.in_ld);
        n.off_b0 = cursor; cursor += pad4(hidden);
        n.off_w1 = cursor; cursor += pad4ll((long long)hidden * hidden);
        n.off_b1 = cursor; cursor += pad4(hidden);
        n.off_w2 = cursor; cursor += pad4ll((long long)n_out * hidden);
        n.off_b2 = cursor; cursor += pad4(n_out);
    }
    n.size = cursor - start;
}

struct NetIds {       // indices into OacLayout::nets
    int policy = -1, target_policy = -1, log_alpha = -1;
    std::vector<int> qf, tf;          // critics and their Polyak targets
};

static int validate(const OacConfig& c) {
    if (c.obs_dim < 1 || c.act_dim < 1 || c.hidden < 1 || c.batch < 1 || c.n_seeds < 1)
        return set_error(OAC_E_INVALID, "dims must be positive");
    if (c.act_dim > 128) return set_error(OAC_E_UNSUPPORTED, "act_dim > 128");
    if (c.hidden > 512) return set_error(OAC_E_UNSUPPORTED, "hidden > 512 (glue kernels keep a hidden row in registers)");
    if ((size_t)2 * c.hidden * (c.act_dim | 1) + (size_t)2 * c.act_dim * c.hidden + 16 * GLUE_WARPS * 3 * c.act_dim > 50000)
        return set_error(OAC_E_UNSUPPORTED, "hidden*act_dim too large for the fused glue kernels' shared memory");
    if (c.algo == OAC_ALGO_POAC && (c.n_particles < 2 || c.n_particles > 16))
        return set_error(OAC_E_UNSUPPORTED, "P-OAC needs 2 <= n_particles <= 16");
    if (c.algo < 0 || c.algo > OAC_ALGO_GOAC) return set_error(OAC_E_INVALID, "unknown algo");
    if (c.std_soft_update && (c.algo != OAC_ALGO_POAC || c.counts))
        return set_error(OAC_E_INVALID, "std_soft_update is a P-OAC option and excludes counts (particle_trainer_oac.py:97)");
    return 0;
}

static int build_layout(const OacConfig& c, OacLayout& L, NetIds& ids) {
    if (int e = validate(c)) return e;
    memset(&L, 0, sizeof(L));
    const int O = c.obs_dim, A = c.act_dim, H = c.hidden, B = c.batch;
    long long cur = 0;
    int n = 0;
    auto add = [&](int kind, int in_dim, int n_out, int trainable) {
        net_layout(L.nets[n], kind, in_dim, H, n_out, trainable, cur);
        return n++;
    };
    int n_crit = 1, heads = 1;
    if (c.algo == OAC_ALGO_SAC) { n_crit = 2; heads = 1; L.nq = 2; }
    else if (c.algo == OAC_ALGO_POAC) {
        n_crit = c.share_layers ? 1 : c.n_particles; heads = c.share_layers ? c.n_particles : 1;
        L.nq = c.n_particles;
    } else { n_crit = c.share_layers ? 1 : 2; heads = c.share_layers ? 2 : 1; L.nq = 2; }
    ids.policy = add(OAC_NET_POLICY, O, 2 * A, 1);
    if (c.algo == OAC_ALGO_GOAC) ids.target_policy = add(OAC_NET_POLICY, O, 2 * A, 1);
    for (int i = 0; i < n_crit; ++i) ids.qf.push_back(add(OAC_NET_Q, O + A, heads, 1));
    ids.log_alpha = add(OAC_NET_SCALAR, 1, 1, 1);
    L.n_trainable = n;
    L.adam_floats = cur;
    for (int i = 0; i < n_crit; ++i) ids.tf.push_back(add(OAC_NET_Q, O + A, heads, 0));
    L.n_nets = n;
    L.param_floats = cur;

    // IO slice
    long long io = 0;
    L.x_rows = 4 * B; L.x_ld = pad4(O + A);
    L.off_x = io; io += (long long)L.x_rows * L.x_ld;
    L.off_rewards = io; io += pad4(B);
    L.off_terminals = io; io += pad4(B);
    L.off_counts = io; io += pad4(B);
    L.off_eps = io; io += pad4ll(2ll * B * A);
    L.off_log_pi = io; io += pad4(3 * B);
    L.off_mean = io; io += pad4ll(3ll * B * A);
    L.off_log_std = io; io += pad4ll(3ll * B * A);
    L.off_q_pred = io; io += pad4ll((long long)B * L.nq);
    L.off_q_target = io; io += pad4ll((long long)B * L.nq);
    L.off_q_new = io; io += pad4ll((long long)B * L.nq);
    L.off_scalars = io; io += SC_COUNT;
    L.io_floats = io;
    L.n_counters = CNT_TOTAL;
    return 0;
}

// ------------------------------------------------------------------------------------
// program
// ------------------------------------------------------------------------------------
enum StageKind { ST_GEMM = 0, ST_POLICY_HEAD = 1, ST_CRITIC_HEAD = 2, ST_POLICY_GRAD = 3, ST_ADAM = 4, ST_STEP_TAIL = 5, ST_RANK1 = 6 };

struct Stage {
    int kind;
    std::vector<GemmTask> gemm;
    std::vector<PolicyHeadTask> ph;
    PolicyHeadParams php;
    CriticHeadParams chp;
    std::vector<PolicyGradTask> pg;
    PolicyGradParams pgp;
    AdamStreamParams asp;       // ST_ADAM
    std::vector<Rank1Task> r1;  // ST_RANK1
    int many = 0;               // glue stage runs its many-seed specialisation (glue_many.cuh)
    // launch
    void* dev = nullptr;        // task table / params on the device
    int a_trans = 0, b_trans = 0;   // all tasks of a GEMM stage share the operand layouts
    int kc = 0;
    size_t smem = 0;
    int use_tc = 0, bn = 0, tmem_cols = 0, n_main = 1;     // tcgen05 path
    // warp-specialised persistent tcgen05 path (gemm_ws.cuh)
    int use_ws = 0, ws_tiles_per_seed = 0, ws_slots = 0, ws_slot_bytes = 0, ws_grid = 0;
    int ws_adam = 0;            // a dW stage with fused Adam epilogues (gemm_ws_kernel<true, true, true>)
    int ws_mask_bits = 0;       // every task of this masked dX stage reads sign-bit words (gemm_ws_kernel<.., .., true>)
    int ws_pair = 0;            // CTA-pair kernel (gemm_ws2.cuh: tcgen05.mma.cta_group::2, 256-row tiles)
    int chain_bwd = 0;          // the chain is a backward one: M/N-contiguous weights, mask / store epilogues (gemm_chain_kernel<true>)
    int chain = 0;              // > 0: strip-fused forward chain of this many layers (gemm_chain.cuh); tasks are layer-major
    int chain_strips0[WS_MAX_TASKS + 1];
    int sk_tma = 0;             // latency-regime FFMA tile with TMA-staged operands (tensor maps in ws_tmaps)
    int fused2 = 0;             // tasks are [layer-1 ..., layer-2 ...] pairs of two dependent forward layers: one cluster launch
    int fwd2_cluster = 0;       // (gemm_fwd2.cuh) when the plan allows it, else two plain launches
    void* ws_tmaps = nullptr;
    int max_tiles = 0;
    int max_rows = 0;
    int glue_iters = 1;         // glue kernels: sample groups per CTA (> 1 when many seeds share a launch)
    int glue_g = 4;             // ... and warps cooperating on one sample
    // multi-lane schedule (latency regime): lane 1 / 2 stages run on side streams; each waits for the main lane's
    // position where it is listed (and for its own lane's earlier stages).  `join` is a bit mask of the side lanes the
    // main lane waits for before this stage (bit 0: lane 1, bit 1: lane 2); everything is joined at the end of the step.
    int lane = 0, join = 0;
    int sm_budget = 0;          // > 0: the persistent tcgen05 grid of this stage is planned for this many SMs (it shares the chip with a side lane)
    const char* name = "";
};

}  // namespace oac

using namespace oac;

struct OacTrainer {
    OacConfig cfg;
    OacLayout lay;
    NetIds ids;
    ArenaSet as;
    AdamHyper hyper;
    std::vector<Stage> stages;
    std::vector<void*> dev_allocs;
    long long work_cursor = 0;
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    cudaStream_t side[2] = {nullptr, nullptr};            // lanes 1, 2
    cudaEvent_t ev_fork[2] = {nullptr, nullptr}, ev_join[2] = {nullptr, nullptr};
    bool use_graph = true;
    int n_opt = 0;
    long long* tc_dbg = nullptr;
    float* host_scalars = nullptr;   // OacBuffers::host_scalars
    bool allow_ws = true;      // OAC_NO_WS=1 forces the per-tile tcgen05 kernel (A/B measurement aid)
    bool allow_ws2 = true;     // OAC_NO_WS2=1: single-CTA tiles only, no CTA pairs (A/B measurement aid)
    bool allow_bwd_chain = true;   // OAC_NO_BWD_CHAIN=1: one launch per dX layer instead of the two strip-fused backward chains (A/B measurement aid)
    bool allow_bits = true;    // OAC_NO_MASK_BITS=1: masked dX epilogues read the fp32 activations instead of sign-bit words (A/B measurement aid)
    bool allow_chain = true;   // OAC_NO_CHAIN=1: one launch per forward layer instead of strip-fused chains (A/B measurement aid)
    bool allow_split = true;   // many-seed tensor-core path: gradient store + streaming Adam instead of the fused epilogue
    bool allow_lanes = true;   // OAC_NO_LANES=1: strictly linear stage order (A/B measurement aid)
    bool allow_sk_tma = true;  // OAC_NO_SK_TMA=1: cp.async staging in the latency-regime FFMA tile
};

namespace oac {

// floats of slack after the last work-arena buffer: a TMA box over a ragged 32-column atom of an M/N-contiguous operand
// reads up to 31 floats past the operand's last row (ws_plan)
constexpr long long WORK_SLACK = 32;

struct Builder {
    OacTrainer& t;
    const OacConfig& c;
    const OacLayout& L;
    int O, A, H, B;
    // Many seeds on the tensor-core path: the head layers, dQ/da and the policy's first backward step run as
    // GEMM stages (the fused glue kernels are instruction-bound there: ncu: 2 800 warp instructions per row, issue-bound); a single seed keeps
    // them fused into the glue kernels (fewer launches: latency).
    bool tensor_glue;
    explicit Builder(OacTrainer& tr) : t(tr), c(tr.cfg), L(tr.lay) {
        O = c.obs_dim; A = c.act_dim; H = c.hidden; B = c.batch;
        static const long long tg_rows = getenv("OAC_TENSOR_GLUE_ROWS") ? atoll(getenv("OAC_TENSOR_GLUE_ROWS")) : 2048;   // measurement aid
        tensor_glue = c.gemm_path == OAC_GEMM_TF32 && (long long)c.n_seeds * c.batch >= tg_rows;
        // (OAC_SPLIT_MIN_SEEDS: measurement aid -- groups below it keep the Adam update fused into the dW epilogue)
        static const int split_min = getenv("OAC_SPLIT_MIN_SEEDS") ? atoi(getenv("OAC_SPLIT_MIN_SEEDS")) : 0;
        split_adam = tensor_glue && tr.allow_split && tr.allow_ws && c.n_seeds >= split_min;
        use_bits = tensor_glue && tr.allow_ws && tr.allow_bits && (H % 64) == 0;
        if (split_adam) grad = work(L.adam_floats);
    }
    // Same regime: the weight-gradient GEMMs store plain gradients (laid out like the trainable prefix of the
    // parameter arena) and a streaming kernel applies Adam + Polyak to whole nets (adam_stream.cuh).
    bool split_adam;
    // Sign bits of the hidden activations (gemm_ws.cuh): forward stages on the TMA + tcgen05 path store one
    // byte per row and 4 columns next to h1 / h2, the masked dX stages and the rank-1 pass read those instead of the fp32
    // activation.  Registered per activation buffer; fwd() / dx() translate a Ref into the buffer to the matching words.
    bool use_bits = false;
    struct BitMap { Ref h; long long len; Ref bits; };
    std::vector<BitMap> bitmaps;
    void add_bits(Ref h, long long rows) {
        if (!use_bits) return;
        BitMap b; b.h = h; b.len = rows * H; b.bits = work(rows * (H / 16));      // H / 4 bytes per row
        bitmaps.push_back(b);
    }
    // words of the rows starting at `r` (a row-aligned Ref into a registered activation), or false
    bool bits_of(Ref r, Ref* out) const {
        for (const BitMap& b : bitmaps)
            if (r.arena == b.h.arena && r.off >= b.h.off && r.off < b.h.off + b.len && ((r.off - b.h.off) % H) == 0) {
                *out = Ref{b.bits.arena, b.bits.off + (r.off - b.h.off) / H * (H / 16)};
                return true;
            }
        return false;
    }
    Ref grad{AR_WORK, 0};
    std::vector<AdamSeg> pending_segs;
    void adam_seg(int ni, int ti, float lr, int counter) {
        if (!split_adam) return;
        const OacNetLayout& n = net(ni);
        AdamSeg sg;
        sg.off = n.off_w0; sg.len = n.size; sg.grad_off = grad.off + n.off_w0;
        sg.target_off = ti >= 0 ? net(ti).off_w0 : -1;
        sg.lr = lr; sg.counter = CNT_OPT0 + counter;
        pending_segs.push_back(sg);
    }
    // part of a net: fc0 (weight + bias) or the rest (fc1 .. last bias) -- both are contiguous blocks of the arena
    void adam_seg_part(int ni, int ti, float lr, int counter, bool fc0) {
        if (!split_adam) return;
        const OacNetLayout& n = net(ni);
        AdamSeg sg;
        const long long lo = fc0 ? n.off_w0 : n.off_w1, hi = fc0 ? n.off_w1 : n.off_w0 + n.size;
        sg.off = lo; sg.len = hi - lo; sg.grad_off = grad.off + lo;
        sg.target_off = ti >= 0 ? net(ti).off_w0 + (lo - n.off_w0) : -1;
        sg.lr = lr; sg.counter = CNT_OPT0 + counter;
        pending_segs.push_back(sg);
    }
    Stage* flush_adam(const char* name) {
        if (!split_adam || pending_segs.empty()) return nullptr;
        Stage& s = add_stage(ST_ADAM, name);
        memset(&s.asp, 0, sizeof(s.asp));
        s.asp.n_seg = (int)pending_segs.size();
        long long tot = 0;
        for (size_t i = 0; i < pending_segs.size(); ++i) { s.asp.seg[i] = pending_segs[i]; tot += pending_segs[i].len >> 2; }
        s.asp.total4 = tot;
        pending_segs.clear();
        return &s;
    }
    Ref work(long long n) {
        Ref r{AR_WORK, t.work_cursor};
        t.work_cursor += pad4ll(n);
        return r;
    }
    Ref P(long long off) const { return Ref{AR_PARAM, off}; }
    Ref X(int block, int col = 0) const { return Ref{AR_IO, L.off_x + (long long)block * B * L.x_ld + col}; }
    const OacNetLayout& net(int i) const { return L.nets[i]; }

    Stage& add_stage(int kind, const char* name) {
        t.stages.emplace_back();
        Stage& s = t.stages.back();
        s.kind = kind; s.name = name;
        return s;
    }
    GemmTask base_task() {
        GemmTask g;
        memset(&g, 0, sizeof(g));
        g.target_off = g.target_bias_off = -1;
        g.train_bias = 1;
        return g;
    }
    // Y[M,N] = act(X[M,K] W[N,K]^T + b)
    void fwd(Stage& s, Ref x, int ldx, int M, int K, Ref w, int ldw, Ref b, int N, Ref y, int ldy, bool relu_) {
        GemmTask g = base_task();
        g.A = x; g.lda = ldx; g.a_trans = 0;
        g.B = w; g.ldb = ldw; g.b_trans = 0;
        g.C = y; g.ldc = ldy; g.M = M; g.N = N; g.K = K;
        g.epi = relu_ ? EPI_BIAS_RELU : EPI_BIAS; g.bias = b;
        Ref bw;
        if (relu_ && N == H && ldy == H && bits_of(y, &bw)) { g.bits = bw; g.ldbits = H / 4; }
        s.gemm.push_back(g);
    }
    // dX[M,Kin] = (dY[M,Nout] W[Nout,Kin]) * (mask > 0)      (mask.arena < 0: no mask)
    void dx(Stage& s, Ref dy, int ldy, int M, int Nout, Ref w, int ldw, int Kin, Ref out, int ldo, Ref mask, int ldmask, bool use_mask) {
        GemmTask g = base_task();
        g.A = dy; g.lda = ldy; g.a_trans = 0;
        g.B = w; g.ldb = ldw; g.b_trans = 1;
        g.C = out; g.ldc = ldo; g.M = M; g.N = Kin; g.K = Nout;
        g.epi = use_mask ? EPI_MASK : EPI_STORE; g.mask = mask; g.ldmask = ldmask;
        Ref bw;
        if (use_mask && Kin == H && ldmask == H && bits_of(mask, &bw)) { g.mask = bw; g.ldmask = H / 4; g.mask_bits = 1; }
        s.gemm.push_back(g);
    }
    // W[Nout,Kin] <- Adam(dY[rows,Nout]^T X[rows,Kin]); bias <- Adam(colsum dY)
    void dw(Stage& s, Ref dy, int ldy, Ref x, int ldx, int rows, int Nout, int Kin, long long w_off, int ldw,
            long long b_off, long long tw_off, long long tb_off, float lr, int counter, int train_bias) {
        GemmTask g = base_task();
        g.A = dy; g.lda = ldy; g.a_trans = 1;
        g.B = x; g.ldb = ldx; g.b_trans = 1;
        g.C = P(w_off); g.ldc = ldw; g.M = Nout; g.N = Kin; g.K = rows;
        g.epi = EPI_ADAM; g.bias = P(b_off); g.has_bias = 1; g.train_bias = train_bias;
        if (split_adam) {
            g.epi = EPI_GRAD;
            g.C = Ref{AR_WORK, grad.off + w_off}; g.bias = Ref{AR_WORK, grad.off + b_off};
        }
        g.adam_off = w_off; g.adam_bias_off = b_off;
        g.target_off = tw_off; g.target_bias_off = tb_off;
        g.lr = lr; g.counter = CNT_OPT0 + counter;
        s.gemm.push_back(g);
    }

    // ---- policy forward (trunk) for rows of X blocks [blk0, blk0+nblk) ----
    struct PolAct { Ref h1, h2, head, save; int rows; };
    PolAct alloc_pol(int nblk) {
        PolAct a; a.rows = nblk * B;
        a.h1 = work((long long)a.rows * H); a.h2 = work((long long)a.rows * H);
        add_bits(a.h1, a.rows); add_bits(a.h2, a.rows);
        a.head = work((long long)a.rows * pad4(2 * A));
        a.save = work((long long)a.rows * 4 * A);
        return a;
    }
    struct CritAct { Ref h1, h2, q, dq, dh2, dh1, da; int rows; };
    CritAct alloc_crit(int nblk, int heads) {
        CritAct a; a.rows = nblk * B;
        a.h1 = work((long long)a.rows * H); a.h2 = work((long long)a.rows * H);
        add_bits(a.h1, a.rows); add_bits(a.h2, a.rows);
        a.q = work((long long)a.rows * heads);
        a.dq = work((long long)a.rows * pad4(heads));         // leading dimension pad4(heads): TMA needs 16-byte row strides
        a.dh2 = work((long long)a.rows * H); a.dh1 = work((long long)a.rows * H);
        a.da = work((long long)a.rows * pad4(A));
        return a;
    }
    void pol_l3(Stage& s, int ni, const PolAct& a) {
        const OacNetLayout& n = net(ni);
        fwd(s, a.h2, H, a.rows, H, P(n.off_w2), H, P(n.off_b2), 2 * A, a.head, pad4(2 * A), false);
    }
    void crit_l3(Stage& s, int ni, const CritAct& a) {
        const OacNetLayout& n = net(ni);
        fwd(s, a.h2, H, a.rows, H, P(n.off_w2), H, P(n.off_b2), n.n_out, a.q, n.n_out, false);
    }
    // da[B,A] = dh1[B,H] W0[:, O:O+A] for rows [row0,row0+B)
    void crit_da(Stage& s, int ni, const CritAct& a, int row0) {
        const OacNetLayout& n = net(ni);
        dx(s, Ref{a.dh1.arena, a.dh1.off + (long long)row0 * H}, H, B, H, P(n.off_w0 + O), n.in_ld, A,
           Ref{a.da.arena, a.da.off + (long long)row0 * pad4(A)}, pad4(A), Ref{0, 0}, 0, false);
    }
    void pol_l1(Stage& s, int ni, int blk0, const PolAct& a) {
        const OacNetLayout& n = net(ni);
        fwd(s, X(blk0), L.x_ld, a.rows, O, P(n.off_w0), n.in_ld, P(n.off_b0), H, a.h1, H, true);
    }
    void pol_l2(Stage& s, int ni, const PolAct& a) {
        const OacNetLayout& n = net(ni);
        fwd(s, a.h1, H, a.rows, H, P(n.off_w1), H, P(n.off_b1), H, a.h2, H, true);
    }
    void crit_l1(Stage& s, int ni, int blk0, const CritAct& a) {
        const OacNetLayout& n = net(ni);
        fwd(s, X(blk0), L.x_ld, a.rows, O + A, P(n.off_w0), n.in_ld, P(n.off_b0), H, a.h1, H, true);
    }
    void crit_l2(Stage& s, int ni, const CritAct& a) {
        const OacNetLayout& n = net(ni);
        fwd(s, a.h1, H, a.rows, H, P(n.off_w1), H, P(n.off_b1), H, a.h2, H, true);
    }
    // the same two layers restricted to rows [row0, row0+B) of the activation set (X block blk)
    void crit_l1_rows(Stage& s, int ni, int blk, const CritAct& a, int row0) {
        const OacNetLayout& n = net(ni);
        fwd(s, X(blk), L.x_ld, B, O + A, P(n.off_w0), n.in_ld, P(n.off_b0), H, Ref{a.h1.arena, a.h1.off + (long long)row0 * H}, H, true);
    }
    void crit_l2_rows(Stage& s, int ni, const CritAct& a, int row0) {
        const OacNetLayout& n = net(ni);
        const long long ro = (long long)row0 * H;
        fwd(s, Ref{a.h1.arena, a.h1.off + ro}, H, B, H, P(n.off_w1), H, P(n.off_b1), H, Ref{a.h2.arena, a.h2.off + ro}, H, true);
    }
    HeadSrc head_src(int ni, const CritAct& a, int row0) {
        const OacNetLayout& n = net(ni);
        HeadSrc h; memset(&h, 0, sizeof(h));
        h.h2 = a.h2; h.row0 = row0; h.w3 = P(n.off_w2); h.b3 = P(n.off_b2); h.n_heads = n.n_out;
        h.write_dh2 = 0; h.dh2 = Ref{a.dh2.arena, a.dh2.off + (long long)row0 * H};
        h.dq = Ref{a.dq.arena, a.dq.off + (long long)row0 * pad4(n.n_out)}; h.dq_ld = pad4(n.n_out);
        return h;
    }
    // dh2 = (dq W3) * (h2>0) for rows [row0, row0+B) of a critic activation set
    void crit_dh2(Stage& s, int ni, const CritAct& a, int row0) {
        const OacNetLayout& n = net(ni);
        long long ro = (long long)row0 * H;
        dx(s, Ref{a.dq.arena, a.dq.off + (long long)row0 * pad4(n.n_out)}, pad4(n.n_out), B, n.n_out, P(n.off_w2), H, H,
           Ref{a.dh2.arena, a.dh2.off + ro}, H, Ref{a.h2.arena, a.h2.off + ro}, H, true);
    }
    // the same product as an elementwise pass (many-seed regime: K = n_heads is no GEMM)
    void crit_dh2_rank1(Stage& s, int ni, const CritAct& a, int row0) {
        const OacNetLayout& n = net(ni);
        long long ro = (long long)row0 * H;
        Rank1Task r; memset(&r, 0, sizeof(r));
        r.dq = Ref{a.dq.arena, a.dq.off + (long long)row0 * pad4(n.n_out)}; r.dq_ld = pad4(n.n_out);
        r.w3 = P(n.off_w2); r.n_heads = n.n_out;
        r.mask = Ref{a.h2.arena, a.h2.off + ro}; r.ldmask = H;
        Ref bw;
        if (bits_of(r.mask, &bw)) { r.mask = bw; r.ldmask = H / 4; r.mask_bits = 1; }
        r.out = Ref{a.dh2.arena, a.dh2.off + ro}; r.ldo = H;
        r.rows = B; r.cols = H;
        s.r1.push_back(r);
    }
    void crit_dh1(Stage& s, int ni, const CritAct& a, int row0) {
        const OacNetLayout& n = net(ni);
        long long ro = (long long)row0 * H;
        dx(s, Ref{a.dh2.arena, a.dh2.off + ro}, H, B, H, P(n.off_w1), H, H,
           Ref{a.dh1.arena, a.dh1.off + ro}, H, Ref{a.h1.arena, a.h1.off + ro}, H, true);
    }
    // Adam on all three layers of a critic from rows [row0,row0+B) (inputs: X block xblk)
    // layers: bit 0 fc0, bit 1 fc1, bit 2 head
    void crit_adam(Stage& s, int ni, int ti, const CritAct& a, int row0, int xblk, float lr, int counter, int layers = 7) {
        const OacNetLayout& n = net(ni);
        const OacNetLayout* tn = ti >= 0 ? &net(ti) : nullptr;
        long long ro = (long long)row0 * H;
        Ref dh1{a.dh1.arena, a.dh1.off + ro}, dh2{a.dh2.arena, a.dh2.off + ro};
        Ref h1{a.h1.arena, a.h1.off + ro}, h2{a.h2.arena, a.h2.off + ro};
        Ref dq{a.dq.arena, a.dq.off + (long long)row0 * pad4(n.n_out)};
        if (layers & 1) dw(s, dh1, H, X(xblk), L.x_ld, B, H, O + A, n.off_w0, n.in_ld, n.off_b0,
                           tn ? tn->off_w0 : -1, tn ? tn->off_b0 : -1, lr, counter, 1);
        if (layers & 2) dw(s, dh2, H, h1, H, B, H, H, n.off_w1, H, n.off_b1, tn ? tn->off_w1 : -1, tn ? tn->off_b1 : -1, lr, counter, 1);
        if (layers & 4) dw(s, dq, pad4(n.n_out), h2, H, B, n.n_out, H, n.off_w2, H, n.off_b2, tn ? tn->off_w2 : -1, tn ? tn->off_b2 : -1,
                           lr, counter, c.train_bias);
        if (layers == 7) adam_seg(ni, ti, lr, counter);
    }
    // policy backward from dhead [B,2A] at rows [row0,row0+B) of a policy activation set
    struct PolGrad { Ref dhead, dh2, dh1; };
    PolGrad alloc_polgrad() {
        PolGrad g; g.dhead = work((long long)B * pad4(2 * A)); g.dh2 = work((long long)B * H); g.dh1 = work((long long)B * H);
        return g;
    }
    void pol_dh2(Stage& s, int ni, const PolAct& a, int row0, const PolGrad& g) {
        const OacNetLayout& n = net(ni);
        dx(s, g.dhead, pad4(2 * A), B, 2 * A, P(n.off_w2), H, H, g.dh2, H, Ref{a.h2.arena, a.h2.off + (long long)row0 * H}, H, true);
    }
    void pol_dh1(Stage& s, int ni, const PolAct& a, int row0, const PolGrad& g) {
        const OacNetLayout& n = net(ni);
        dx(s, g.dh2, H, B, H, P(n.off_w1), H, H, g.dh1, H, Ref{a.h1.arena, a.h1.off + (long long)row0 * H}, H, true);
    }
    void pol_adam(Stage& s, int ni, const PolAct& a, int row0, int xblk, const PolGrad& g, float lr, int counter) {
        const OacNetLayout& n = net(ni);
        long long ro = (long long)row0 * H;
        dw(s, g.dh1, H, X(xblk), L.x_ld, B, H, O, n.off_w0, n.in_ld, n.off_b0, -1, -1, lr, counter, 1);
        dw(s, g.dh2, H, Ref{a.h1.arena, a.h1.off + ro}, H, B, H, H, n.off_w1, H, n.off_b1, -1, -1, lr, counter, 1);
        dw(s, g.dhead, pad4(2 * A), Ref{a.h2.arena, a.h2.off + ro}, H, B, 2 * A, H, n.off_w2, H, n.off_b2, -1, -1, lr, counter, 1);
        adam_seg(ni, -1, lr, counter);
    }
    // policy-loss gradient task: critics (dh1 rows, fc0 weights) -> dhead, and the policy's own dh2
    // crit_da_refs (tensor_glue): per critic, where the pi_da GEMM stage left dh1 W0[:, O:O+A] for these rows
    PolicyGradTask pg_task(const std::vector<std::pair<int, Ref>>& crit_dh1, int pol, const PolAct& a, int row0,
                           const PolGrad& g, bool entropy, const std::vector<Ref>& crit_da_refs = std::vector<Ref>()) {
        PolicyGradTask t_; memset(&t_, 0, sizeof(t_));
        int i = 0;
        for (auto& c_ : crit_dh1) {
            const OacNetLayout& qn = net(c_.first);
            t_.dh1[i] = c_.second; t_.w1[i] = P(qn.off_w0); t_.ld[i] = qn.in_ld;
            if (i < (int)crit_da_refs.size()) t_.da[i] = crit_da_refs[i];
            ++i;
        }
        t_.da_ld = pad4(A);
        t_.n_src = i;
        t_.save = a.save; t_.save_row0 = row0; t_.dhead = g.dhead; t_.dhead_ld = pad4(2 * A); t_.entropy = entropy ? 1 : 0;
        t_.wh = P(net(pol).off_w2); t_.h2 = a.h2; t_.h2_row0 = row0; t_.dhp2 = g.dh2;
        return t_;
    }
    PolicyHeadTask ph_task(int ni, const PolAct& a, int out_row0, int dst0, int dst1, int eps0, int eps1) {
        const OacNetLayout& n = net(ni);
        PolicyHeadTask p;
        memset(&p, 0, sizeof(p));
        p.h2 = a.h2; p.w = P(n.off_w2); p.b = P(n.off_b2); p.rows = a.rows; p.out_row0 = out_row0;
        p.dst_block[0] = dst0; p.dst_block[1] = dst1; p.eps_slot[0] = eps0; p.eps_slot[1] = eps1;
        p.save = a.save;
        p.head_in = a.head; p.head_ld = pad4(2 * A);
        return p;
    }
    void fill_php(Stage& s, int alpha_task, int alpha_block, int alpha_counter) {
        PolicyHeadParams& p = s.php;
        memset(&p, 0, sizeof(p));
        p.off_x = L.off_x; p.off_eps = L.off_eps; p.off_log_pi = L.off_log_pi; p.off_mean = L.off_mean;
        p.off_log_std = L.off_log_std; p.off_scalars = L.off_scalars;
        p.x_ld = L.x_ld; p.O = O; p.A = A; p.H = H; p.B = B;
        p.deterministic = c.deterministic; p.rng_seed = c.rng_seed;
        p.n_opt_counters = t.n_opt;
        p.alpha.enabled = c.auto_alpha; p.alpha.task = alpha_task; p.alpha.block = alpha_block;
        const OacNetLayout& la = net(t.ids.log_alpha);
        p.alpha.log_alpha = P(la.off_w0); p.alpha.adam_off = la.off_w0;
        p.alpha.lr = c.policy_lr; p.alpha.target_entropy = c.target_entropy; p.alpha.counter = alpha_counter;
        p.head_from_gemm = tensor_glue ? 1 : 0;
        p.host_scalars = t.host_scalars; p.n_seeds = c.n_seeds;
    }
    void fill_chp(Stage& s, int mode, int n_nets) {
        CriticHeadParams& p = s.chp;
        p.mode = mode;
        p.off_rewards = L.off_rewards; p.off_terminals = L.off_terminals; p.off_counts = L.off_counts;
        p.off_log_pi = L.off_log_pi; p.off_q_pred = L.off_q_pred; p.off_q_target = L.off_q_target;
        p.off_q_new = L.off_q_new; p.off_scalars = L.off_scalars;
        p.B = B; p.H = H; p.P = c.n_particles; p.nq = L.nq; p.n_nets = n_nets;
        p.share_layers = c.share_layers; p.counts = c.counts;
        p.discount = c.discount; p.reward_scale = c.reward_scale;
        p.standard_bound = c.standard_bound; p.std_init = c.std_init;
        p.std_soft_update = c.std_soft_update; p.std_soft_prob = c.std_soft_update_prob;
    }
    void fill_pgp(Stage& s) {
        PolicyGradParams& p = s.pgp;
        memset(&p, 0, sizeof(p));
        p.off_scalars = L.off_scalars; p.O = O; p.A = A; p.H = H; p.B = B;
        p.da_from_gemm = tensor_glue ? 1 : 0;
    }

    // X blocks: 0 [obs|a_tp] (G-OAC), 1 [obs|a_pi], 2 [obs|actions], 3 [next_obs|a_next]
    // Latency regime (few seeds): a step is a chain of small kernels that leave most SMs idle, so independent work runs
    // on side lanes (launch_stages).
    bool latency_lanes() const {
        static const long long lane_rows = getenv("OAC_LANE_ROWS") ? atoll(getenv("OAC_LANE_ROWS")) : 1024;              // measurement aid
        return !tensor_glue && c.n_seeds * (long long)B <= lane_rows && t.allow_lanes;
    }
    // The step counters / entropy-temperature Adam step are needed by the first critic_head at the earliest: they leave
    // the policy_head kernel just added (fence + ticket + reduction in its last CTA) for a one-CTA kernel on lane 2
    // (measured per step: SAC 123.6 -> 119.4 us, P-OAC 129.6 -> 127.0, G-OAC 131.0 -> 128.9).
    void split_step_tail() {
        // many-seed program: the ticket + last-CTA tail costs policy_head 6 us with 8 seeds (12 us with 64); as a kernel of its
        // own on lane 2 it overlaps the critics' forward.  Pays for small groups (8 seeds: 217.1 -> 213.7 us per step); with
        // 64 seeds the extra launch and the fork / join cancel it (0.798 vs 0.800 ms), so: up to 16 seeds.
        const bool small_group = tensor_glue && t.allow_lanes && c.n_seeds <= 16;
        if (!(latency_lanes() || small_group) || getenv("OAC_NO_TAIL_SPLIT")) return;                      // (env: A/B measurement aid)
        t.stages.back().php.tail_in_own_kernel = 1;
        Stage head = t.stages.back();
        Stage& s = add_stage(ST_STEP_TAIL, "alpha+step_counters"); s.lane = 2; s.ph = head.ph; s.php = head.php;
    }
    // two dependent forward layers as one cluster launch (gemm_fwd2.cuh): the FFMA latency regime only
    // many-seed tensor-core regime: l1 -> l2 (-> head) of a 128-row strip in one launch, activations handed on in TMEM
    bool fuse_chain() const { return tensor_glue && t.allow_chain && t.allow_ws && H == 256 && (B % 128) == 0; }
    // builds the forward stages of a group of nets: `layers` = 2 (critics) or 3 (policies: trunk + head GEMM)
    template <typename F1, typename F2, typename F3>
    void forward_stages(const char* n1, const char* n2, const char* n3, const char* nchain, int layers, F1 l1, F2 l2, F3 l3) {
        if (fuse_chain()) {
            // Measured (B200): a chain pays when its strips fit ONE round of the persistent grid (8 seeds: policy 33.3 -> 27.1 us,
            // critics 31.3 -> 26.9 us: two tile life cycles and two launches gone); with several rounds the unfused stages
            // overlap every epilogue with the next tile's mainloop and win (64 seeds: 0.796 vs 0.814 ms per step).
            Stage probe; l1(probe);
            long long strips = 0;
            for (auto& g : probe.gemm) strips += (g.M + 127) / 128;
            const bool all = getenv("OAC_CHAIN_ALL") && getenv("OAC_CHAIN_ALL")[0] == '1';       // (tests, A/B runs)
            // (re-measured on the final build: up to ~1.3 rounds the chain still wins -- 16 seeds, critics' 192 strips: step 0.2755
            // -> 0.2699 ms; neutral at 2.6 rounds, 32 seeds)
            if (all || strips * c.n_seeds <= sm_count() + sm_count() / 2) {
                Stage& s = add_stage(ST_GEMM, nchain); s.chain = layers;
                l1(s); l2(s); if (layers == 3) l3(s);
                return;
            }
        }
        { Stage& s = add_stage(ST_GEMM, n1); l1(s); }
        // several rounds of strips: only the cheap head layer rides on its predecessor (l2 > head: h2 goes on through tensor
        // memory, the head's 34-column MMAs run while h2 is being stored) -- OAC_TAIL_CHAIN=0 keeps the three launches
        static const bool tail_chain = !(getenv("OAC_TAIL_CHAIN") && getenv("OAC_TAIL_CHAIN")[0] == '0');
        if (layers == 3 && fuse_chain() && tail_chain) {
            Stage& s = add_stage(ST_GEMM, "policy_l2>l3"); s.chain = 2;
            l2(s); l3(s);
            return;
        }
        { Stage& s = add_stage(ST_GEMM, n2); l2(s); }
        if (layers == 3 && tensor_glue) { Stage& s = add_stage(ST_GEMM, n3); l3(s); }
    }
    bool fuse_fwd2() const {
        return latency_lanes() && c.gemm_path == OAC_GEMM_FP32 && !getenv("OAC_NO_FWD2");               // (env: A/B measurement aid)
    }
    void build_sac();
    void build_poac();
    void build_goac();
};

void Builder::build_sac() {
    // optimizers: 0 policy, 1 qf1, 2 qf2, 3 alpha
    t.n_opt = 4;
    const int pol = t.ids.policy, q1 = t.ids.qf[0], q2 = t.ids.qf[1], t1 = t.ids.tf[0], t2 = t.ids.tf[1];
    PolAct pa = alloc_pol(2);                     // rows [0,B) obs (block 2), [B,2B) next_obs (block 3)
    CritAct ca1 = alloc_crit(2, 1), ca2 = alloc_crit(2, 1);   // rows [0,B) a_pi (block 1), [B,2B) data (block 2)
    CritAct ta1 = alloc_crit(1, 1), ta2 = alloc_crit(1, 1);   // block 3
    PolGrad pg = alloc_polgrad();
    const bool mode_b = c.stale_graph_mode == 1;
    // second lane: the critics' forward on the DATA rows does not depend on the policy and overlaps its forward
    const bool two_lanes = latency_lanes();
    // layer pairs as one cluster launch each (gemm_fwd2.cuh).  Measured per step: none 119.4 us, all three 117.3 us; any
    // subset is between 118.7 and 123.5 us (the lanes' co-scheduling, not the sum of the parts, decides).
    const bool fwd2 = fuse_fwd2();
    if (two_lanes && fwd2) {
        Stage& s = add_stage(ST_GEMM, "critic_l1+l2_data"); s.lane = 1; s.fused2 = 1;
        crit_l1_rows(s, q1, 2, ca1, B); crit_l1_rows(s, q2, 2, ca2, B); crit_l2_rows(s, q1, ca1, B); crit_l2_rows(s, q2, ca2, B);
    } else if (two_lanes) {
        { Stage& s = add_stage(ST_GEMM, "critic_l1_data"); s.lane = 1; crit_l1_rows(s, q1, 2, ca1, B); crit_l1_rows(s, q2, 2, ca2, B); }
        { Stage& s = add_stage(ST_GEMM, "critic_l2_data"); s.lane = 1; crit_l2_rows(s, q1, ca1, B); crit_l2_rows(s, q2, ca2, B); }
    }
    if (fwd2) { Stage& s = add_stage(ST_GEMM, "policy_l1+l2"); s.fused2 = 1; pol_l1(s, pol, 2, pa); pol_l2(s, pol, pa); }
    else forward_stages("policy_l1", "policy_l2", "policy_l3", "policy_l1>l2>l3", 3,
                        [&](Stage& s) { pol_l1(s, pol, 2, pa); }, [&](Stage& s) { pol_l2(s, pol, pa); }, [&](Stage& s) { pol_l3(s, pol, pa); });
    { Stage& s = add_stage(ST_POLICY_HEAD, "policy_head+sample+alpha");
      s.ph.push_back(ph_task(pol, pa, 0, 1, 3, 0, 1)); fill_php(s, 0, 0, 3); }
    split_step_tail();
    if (two_lanes && fwd2) {
        Stage& s = add_stage(ST_GEMM, "critic_l1+l2_pi"); s.fused2 = 1;
        crit_l1_rows(s, q1, 1, ca1, 0); crit_l1_rows(s, q2, 1, ca2, 0); crit_l1(s, t1, 3, ta1); crit_l1(s, t2, 3, ta2);
        crit_l2_rows(s, q1, ca1, 0); crit_l2_rows(s, q2, ca2, 0); crit_l2(s, t1, ta1); crit_l2(s, t2, ta2);
    } else if (two_lanes) {
        { Stage& s = add_stage(ST_GEMM, "critic_l1_pi");
          crit_l1_rows(s, q1, 1, ca1, 0); crit_l1_rows(s, q2, 1, ca2, 0); crit_l1(s, t1, 3, ta1); crit_l1(s, t2, 3, ta2); }
        { Stage& s = add_stage(ST_GEMM, "critic_l2_pi");
          crit_l2_rows(s, q1, ca1, 0); crit_l2_rows(s, q2, ca2, 0); crit_l2(s, t1, ta1); crit_l2(s, t2, ta2); }
    } else {
        forward_stages("critic_l1", "critic_l2", "", "critic_l1>l2", 2,
                       [&](Stage& s) { crit_l1(s, q1, 1, ca1); crit_l1(s, q2, 1, ca2); crit_l1(s, t1, 3, ta1); crit_l1(s, t2, 3, ta2); },
                       [&](Stage& s) { crit_l2(s, q1, ca1); crit_l2(s, q2, ca2); crit_l2(s, t1, ta1); crit_l2(s, t2, ta2); },
                       [&](Stage&) {});
    }
    { Stage& s = add_stage(ST_CRITIC_HEAD, "critic_head+targets+dh2"); s.join = 3;
      memset(&s.chp, 0, sizeof(s.chp));
      s.chp.src[0] = head_src(q1, ca1, 0); s.chp.src[1] = head_src(q2, ca2, 0);
      s.chp.src[2] = head_src(q1, ca1, B); s.chp.src[3] = head_src(q2, ca2, B);
      s.chp.src[4] = head_src(t1, ta1, 0); s.chp.src[5] = head_src(t2, ta2, 0);
      s.chp.src[2].write_dh2 = s.chp.src[3].write_dh2 = 1;              // Q-loss backward starts here
      // mode B: the policy-loss dX uses the PRE-step critic weights, so its dh2 can be formed here as well;
      // mode A (torch 1.4) multiplies by the POST-step W3 -> separate stage after the critic Adam
      if (mode_b) s.chp.src[0].write_dh2 = s.chp.src[1].write_dh2 = 1;
      s.chp.n_src = 6; fill_chp(s, CM_SAC, 2); }
    // Small groups on the tensor path, mode A (every stage is ONE round of tiles, i.e. pure per-tile latency: DESIGN 6): the
    // policy-loss backward through the critics needs only the POST-step fc1 / head weights for its first two stages, and those
    // gradients are complete after critic_head.  Lane 1: dW(fc1, head) -> their Adam -> pi_dh2 -> pi_dh1, next to the main
    // lane's qloss_dh1 -> dW(fc0) -> its Adam; the lanes meet at pi_da.  The GEMM stages of the two branches are planned
    // for half of the SMs each so that they really run side by side.  OAC_GROUP_LANES=0 restores the linear order (A/B aid).
    // Strip-fused BACKWARD chains on the tensor pipe (gemm_chain.cuh, B_MN): pi_dh1 > pi_da (the policy-loss gradient through a
    // critic: dh1 never leaves the SM) and policy_dh2 > policy_dh1; the masks are the sign bytes of the forward pass.
    const bool bwd_chain = fuse_chain() && use_bits && t.allow_bwd_chain && !mode_b;
    // (measured, B200: 64 seeds pi_dh1 + pi_da 43.4 -> 31.0 us, policy_dh2 + policy_dh1 26.4 -> 17.3 us, step 0.727 -> 0.709 ms.
    // In a small group pi_dh1 already hides on a side lane -- group_lanes below -- and the fused pair would put its 19 us on the
    // critical path instead of pi_da's 12.6: there only the policy pair is fused.)
    static const int gl_max = getenv("OAC_GROUP_LANES") ? atoi(getenv("OAC_GROUP_LANES")) : 16;
    const bool group_lanes = tensor_glue && t.allow_lanes && !mode_b && (H & 3) == 0 && c.n_seeds <= gl_max;
    const int half = sm_count() / 2;
    const bool pi_chain = bwd_chain && !group_lanes;
    if (two_lanes && !mode_b) {
        // mode A: pi_dh2 needs the POST-step head weights, and the head gradient (dq^T h2) is complete after critic_head:
        // head Adam + pi_dh2 leave the critical chain for lane 2 and run next to qloss_dh1 / the fc1 Adam
        { Stage& s = add_stage(ST_GEMM, "critic_adam_head"); s.lane = 2;
          crit_adam(s, q1, t1, ca1, B, 2, c.qf_lr, 1, 4); crit_adam(s, q2, t2, ca2, B, 2, c.qf_lr, 2, 4); }
        { Stage& s = add_stage(ST_GEMM, "pi_dh2"); s.lane = 2; crit_dh2(s, q1, ca1, 0); crit_dh2(s, q2, ca2, 0); }
    }
    if (group_lanes) {
        // (listed before qloss_dh1: a side-lane stage waits for the main lane's position where it is listed)
        { Stage& s = add_stage(ST_GEMM, "critic_adam_fc1+head"); s.lane = 1; s.sm_budget = half;
          crit_adam(s, q1, t1, ca1, B, 2, c.qf_lr, 1, 6); crit_adam(s, q2, t2, ca2, B, 2, c.qf_lr, 2, 6);
          adam_seg_part(q1, t1, c.qf_lr, 1, false); adam_seg_part(q2, t2, c.qf_lr, 2, false); }
        if (Stage* a = flush_adam("critic_adam_apply_fc1+head")) a->lane = 1;
        { Stage& s = add_stage(ST_RANK1, "pi_dh2"); s.lane = 1; crit_dh2_rank1(s, q1, ca1, 0); crit_dh2_rank1(s, q2, ca2, 0); }
        { Stage& s = add_stage(ST_GEMM, "pi_dh1"); s.lane = 1; s.sm_budget = half; crit_dh1(s, q1, ca1, 0); crit_dh1(s, q2, ca2, 0); }
    }
    { Stage& s = add_stage(ST_GEMM, "qloss_dh1"); crit_dh1(s, q1, ca1, B); crit_dh1(s, q2, ca2, B);
      if (group_lanes) s.sm_budget = half;
      if (mode_b) { crit_dh1(s, q1, ca1, 0); crit_dh1(s, q2, ca2, 0); } }
    auto policy_grad_stage = [&]() {
        if (pi_chain) {
            Stage& sd = add_stage(ST_GEMM, "pi_dh1>da"); sd.chain = 2; sd.chain_bwd = 1;
            crit_dh1(sd, q1, ca1, 0); sd.gemm.back().no_store = 1; crit_dh1(sd, q2, ca2, 0); sd.gemm.back().no_store = 1;
            crit_da(sd, q1, ca1, 0); crit_da(sd, q2, ca2, 0);
        } else if (tensor_glue) { Stage& sd = add_stage(ST_GEMM, "pi_da"); sd.join = group_lanes ? 1 : 0; crit_da(sd, q1, ca1, 0); crit_da(sd, q2, ca2, 0); }
        Stage& s = add_stage(ST_POLICY_GRAD, tensor_glue ? "policy_grad" : "policy_grad+da+dh2");
        s.join = 3;
        s.pg.push_back(pg_task({{q1, ca1.dh1}, {q2, ca2.dh1}}, pol, pa, 0, pg, !c.deterministic, {ca1.da, ca2.da}));
        fill_pgp(s);
        if (bwd_chain) {
            Stage& s2 = add_stage(ST_GEMM, "policy_dh2>dh1"); s2.chain = 2; s2.chain_bwd = 1;
            pol_dh2(s2, pol, pa, 0, pg); pol_dh1(s2, pol, pa, 0, pg);
        } else if (tensor_glue) { Stage& s2 = add_stage(ST_GEMM, "policy_dh2"); pol_dh2(s2, pol, pa, 0, pg); }
    };
    // NB mode B reads fc0.weight's action columns too: its policy_grad stage runs before the critic Adam
    if (mode_b) policy_grad_stage();
    if (two_lanes && !mode_b) {
        // mode A: pi_dh2 / pi_dh1 read the POST-step head and fc1 weights, policy_grad the post-step fc0 action columns:
        // the (largest) fc0 Adam runs on lane 1 next to the two dX stages
        { Stage& s = add_stage(ST_GEMM, "critic_adam_fc0"); s.lane = 1;
          crit_adam(s, q1, t1, ca1, B, 2, c.qf_lr, 1, 1); crit_adam(s, q2, t2, ca2, B, 2, c.qf_lr, 2, 1); }
        { Stage& s = add_stage(ST_GEMM, "critic_adam_fc1");
          crit_adam(s, q1, t1, ca1, B, 2, c.qf_lr, 1, 2); crit_adam(s, q2, t2, ca2, B, 2, c.qf_lr, 2, 2); }
    } else if (group_lanes) {
        { Stage& s = add_stage(ST_GEMM, "critic_adam_fc0"); s.sm_budget = half;
          crit_adam(s, q1, t1, ca1, B, 2, c.qf_lr, 1, 1); crit_adam(s, q2, t2, ca2, B, 2, c.qf_lr, 2, 1);
          adam_seg_part(q1, t1, c.qf_lr, 1, true); adam_seg_part(q2, t2, c.qf_lr, 2, true); }
        flush_adam("critic_adam_apply_fc0");
    } else {
        { Stage& s = add_stage(ST_GEMM, "critic_adam");
          crit_adam(s, q1, t1, ca1, B, 2, c.qf_lr, 1); crit_adam(s, q2, t2, ca2, B, 2, c.qf_lr, 2); }
        flush_adam("critic_adam_apply");
    }
    if (!mode_b && group_lanes) {
        policy_grad_stage();                   // pi_da joins lane 1
    } else if (!mode_b) {
        if (!two_lanes && tensor_glue && (H & 3) == 0) {
            Stage& s = add_stage(ST_RANK1, "pi_dh2"); crit_dh2_rank1(s, q1, ca1, 0); crit_dh2_rank1(s, q2, ca2, 0);
        } else if (!two_lanes) { Stage& s = add_stage(ST_GEMM, "pi_dh2"); crit_dh2(s, q1, ca1, 0); crit_dh2(s, q2, ca2, 0); }
        if (!pi_chain) { Stage& s = add_stage(ST_GEMM, "pi_dh1"); s.join = 2; crit_dh1(s, q1, ca1, 0); crit_dh1(s, q2, ca2, 0); }
        policy_grad_stage();
    }
    if (!bwd_chain) { Stage& s = add_stage(ST_GEMM, "policy_dh1"); pol_dh1(s, pol, pa, 0, pg); }
    { Stage& s = add_stage(ST_GEMM, "policy_adam"); pol_adam(s, pol, pa, 0, 2, pg, c.policy_lr, 0); }
    flush_adam("policy_adam_apply");
}

void Builder::build_poac() {
    // optimizers: 0 policy, 1 alpha, 2.. critics
    const int n = (int)t.ids.qf.size();
    t.n_opt = 2 + n;
    const int pol = t.ids.policy;
    const int heads = net(t.ids.qf[0]).n_out;
    PolAct pa = alloc_pol(2);
    std::vector<CritAct> qa, ta, pa_q;      // data rows (block 2), next rows (block 3), a_pi rows (block 1)
    for (int i = 0; i < n; ++i) { qa.push_back(alloc_crit(1, heads)); ta.push_back(alloc_crit(1, heads)); pa_q.push_back(alloc_crit(1, heads)); }
    PolGrad pg = alloc_polgrad();
    // (fusing the layer pairs of this linear chain into cluster launches, gemm_fwd2.cuh, measured slower: 131.3 vs 127.0 us)
    { Stage& s = add_stage(ST_GEMM, "policy_l1"); pol_l1(s, pol, 2, pa); }
    { Stage& s = add_stage(ST_GEMM, "policy_l2"); pol_l2(s, pol, pa); }
    if (tensor_glue) { Stage& s = add_stage(ST_GEMM, "policy_l3"); pol_l3(s, pol, pa); }
    // eps slots are named by meaning (0: obs draw, 1: next_obs draw); the reference draws the
    // next_obs noise FIRST here (:193 then :271) -- the host wrapper maps call order to slots
    { Stage& s = add_stage(ST_POLICY_HEAD, "policy_head+sample+alpha");
      s.ph.push_back(ph_task(pol, pa, 0, 1, 3, 0, 1)); fill_php(s, 0, 0, 1); }
    // Of the SAC program's lanes only the step tail pays here (measured, 127.0 us per step): the data-row critic forward
    // on lane 1 (134.2 us) and the head / fc1 Adam stages on side lanes (129.0 us) cost more in forks and joins than the
    // two critic phases of this step can hide.
    split_step_tail();
    { Stage& s = add_stage(ST_GEMM, "critic_l1");
      for (int i = 0; i < n; ++i) { crit_l1(s, t.ids.qf[i], 2, qa[i]); crit_l1(s, t.ids.tf[i], 3, ta[i]); } }
    { Stage& s = add_stage(ST_GEMM, "critic_l2");
      for (int i = 0; i < n; ++i) { crit_l2(s, t.ids.qf[i], qa[i]); crit_l2(s, t.ids.tf[i], ta[i]); } }
    { Stage& s = add_stage(ST_CRITIC_HEAD, "critic_head+sort+targets+dh2"); s.join = 3;
      memset(&s.chp, 0, sizeof(s.chp));
      for (int i = 0; i < n; ++i) { s.chp.src[i] = head_src(t.ids.qf[i], qa[i], 0); s.chp.src[i].write_dh2 = 1; }
      for (int i = 0; i < n; ++i) s.chp.src[n + i] = head_src(t.ids.tf[i], ta[i], 0);
      s.chp.n_src = 2 * n; fill_chp(s, CM_POAC_Q, n); }
    { Stage& s = add_stage(ST_GEMM, "qloss_dh1"); for (int i = 0; i < n; ++i) crit_dh1(s, t.ids.qf[i], qa[i], 0); }
    { Stage& s = add_stage(ST_GEMM, "critic_adam");
      for (int i = 0; i < n; ++i) crit_adam(s, t.ids.qf[i], t.ids.tf[i], qa[i], 0, 2, c.qf_lr, 2 + i); }
    flush_adam("critic_adam_apply");
    // policy phase through the UPDATED critics
    { Stage& s = add_stage(ST_GEMM, "pi_critic_l1"); for (int i = 0; i < n; ++i) crit_l1(s, t.ids.qf[i], 1, pa_q[i]); }
    { Stage& s = add_stage(ST_GEMM, "pi_critic_l2"); for (int i = 0; i < n; ++i) crit_l2(s, t.ids.qf[i], pa_q[i]); }
    { Stage& s = add_stage(ST_CRITIC_HEAD, "critic_head+min_particle+dh2");
      memset(&s.chp, 0, sizeof(s.chp));
      for (int i = 0; i < n; ++i) { s.chp.src[i] = head_src(t.ids.qf[i], pa_q[i], 0); s.chp.src[i].write_dh2 = 1; }
      s.chp.n_src = n; fill_chp(s, CM_POAC_PI, n); }
    { Stage& s = add_stage(ST_GEMM, "pi_dh1"); for (int i = 0; i < n; ++i) crit_dh1(s, t.ids.qf[i], pa_q[i], 0); }
    if (tensor_glue) { Stage& s = add_stage(ST_GEMM, "pi_da"); for (int i = 0; i < n; ++i) crit_da(s, t.ids.qf[i], pa_q[i], 0); }
    { Stage& s = add_stage(ST_POLICY_GRAD, tensor_glue ? "policy_grad" : "policy_grad+da+dh2");
      std::vector<std::pair<int, Ref>> cr;
      std::vector<Ref> das;
      for (int i = 0; i < n; ++i) { cr.push_back({t.ids.qf[i], pa_q[i].dh1}); das.push_back(pa_q[i].da); }
      s.pg.push_back(pg_task(cr, pol, pa, 0, pg, !c.deterministic, das)); fill_pgp(s); }
    if (tensor_glue) { Stage& s = add_stage(ST_GEMM, "policy_dh2"); pol_dh2(s, pol, pa, 0, pg); }
    { Stage& s = add_stage(ST_GEMM, "policy_dh1"); pol_dh1(s, pol, pa, 0, pg); }
    { Stage& s = add_stage(ST_GEMM, "policy_adam"); pol_adam(s, pol, pa, 0, 2, pg, c.policy_lr, 0); }
    flush_adam("policy_adam_apply");
}

void Builder::build_goac() {
    // optimizers: 0 policy, 1 target_policy, 2 q, 3 std (separate nets), alpha never steps
    const int n = (int)t.ids.qf.size();       // 1 shared, 2 separate (q, std)
    t.n_opt = 2 + n;
    const int pol = t.ids.policy, tpol = t.ids.target_policy;
    const int heads = net(t.ids.qf[0]).n_out;
    PolAct pa = alloc_pol(2);                  // policy on obs (block 2), next_obs (block 3)
    PolAct tpa = alloc_pol(1);                 // target policy on obs (block 2)
    std::vector<CritAct> qa, ta, pq;           // data (block 2) | next (block 3) | policy phase blocks 0,1 (2B rows)
    for (int i = 0; i < n; ++i) { qa.push_back(alloc_crit(1, heads)); ta.push_back(alloc_crit(1, heads)); pq.push_back(alloc_crit(2, heads)); }
    PolGrad pg = alloc_polgrad(), tpg = alloc_polgrad();
    { Stage& s = add_stage(ST_GEMM, "policy_l1"); pol_l1(s, pol, 2, pa); pol_l1(s, tpol, 2, tpa); }
    { Stage& s = add_stage(ST_GEMM, "policy_l2"); pol_l2(s, pol, pa); pol_l2(s, tpol, tpa); }
    if (tensor_glue) { Stage& s = add_stage(ST_GEMM, "policy_l3"); pol_l3(s, pol, pa); pol_l3(s, tpol, tpa); }
    { Stage& s = add_stage(ST_POLICY_HEAD, "policy_head");
      s.ph.push_back(ph_task(pol, pa, 0, 1, 3, 0, 1));
      s.ph.push_back(ph_task(tpol, tpa, 2 * B, 0, 0, 0, 0));
      fill_php(s, 0, 0, 0); s.php.alpha.enabled = 0; s.php.deterministic = 1; }
    split_step_tail();                           // as in P-OAC: the only side-lane stage that pays (131.0 -> 128.9 us)
    { Stage& s = add_stage(ST_GEMM, "critic_l1");
      for (int i = 0; i < n; ++i) { crit_l1(s, t.ids.qf[i], 2, qa[i]); crit_l1(s, t.ids.tf[i], 3, ta[i]); } }
    { Stage& s = add_stage(ST_GEMM, "critic_l2");
      for (int i = 0; i < n; ++i) { crit_l2(s, t.ids.qf[i], qa[i]); crit_l2(s, t.ids.tf[i], ta[i]); } }
    { Stage& s = add_stage(ST_CRITIC_HEAD, "critic_head+targets+dh2"); s.join = 3;
      memset(&s.chp, 0, sizeof(s.chp));
      for (int i = 0; i < n; ++i) { s.chp.src[i] = head_src(t.ids.qf[i], qa[i], 0); s.chp.src[i].write_dh2 = 1; }
      for (int i = 0; i < n; ++i) s.chp.src[n + i] = head_src(t.ids.tf[i], ta[i], 0);
      s.chp.n_src = 2 * n; fill_chp(s, CM_GOAC_Q, n); }
    { Stage& s = add_stage(ST_GEMM, "qloss_dh1"); for (int i = 0; i < n; ++i) crit_dh1(s, t.ids.qf[i], qa[i], 0); }
    { Stage& s = add_stage(ST_GEMM, "critic_adam");
      for (int i = 0; i < n; ++i)
          crit_adam(s, t.ids.qf[i], t.ids.tf[i], qa[i], 0, 2, i == 0 ? c.qf_lr : c.std_lr, 2 + i); }
    flush_adam("critic_adam_apply");
    // policy (rows [B,2B) = block 1) and target policy (rows [0,B) = block 0) through the updated critic
    { Stage& s = add_stage(ST_GEMM, "pi_critic_l1"); for (int i = 0; i < n; ++i) crit_l1(s, t.ids.qf[i], 0, pq[i]); }
    { Stage& s = add_stage(ST_GEMM, "pi_critic_l2"); for (int i = 0; i < n; ++i) crit_l2(s, t.ids.qf[i], pq[i]); }
    { Stage& s = add_stage(ST_CRITIC_HEAD, "critic_head+upper_bound+dh2");
      memset(&s.chp, 0, sizeof(s.chp));
      for (int i = 0; i < n; ++i) { s.chp.src[i] = head_src(t.ids.qf[i], pq[i], B); s.chp.src[i].write_dh2 = 1; }          // a_pi rows
      for (int i = 0; i < n; ++i) { s.chp.src[n + i] = head_src(t.ids.qf[i], pq[i], 0); s.chp.src[n + i].write_dh2 = 1; }  // a_tp rows
      s.chp.n_src = 2 * n; fill_chp(s, CM_GOAC_PI, n); }
    { Stage& s = add_stage(ST_GEMM, "pi_dh1");
      for (int i = 0; i < n; ++i) { crit_dh1(s, t.ids.qf[i], pq[i], 0); crit_dh1(s, t.ids.qf[i], pq[i], B); } }
    if (tensor_glue) { Stage& s = add_stage(ST_GEMM, "pi_da");
      for (int i = 0; i < n; ++i) { crit_da(s, t.ids.qf[i], pq[i], 0); crit_da(s, t.ids.qf[i], pq[i], B); } }
    { Stage& s = add_stage(ST_POLICY_GRAD, tensor_glue ? "policy_grad" : "policy_grad+da+dh2");
      std::vector<std::pair<int, Ref>> cr, crt;
      std::vector<Ref> das, dast;
      for (int i = 0; i < n; ++i) {
          cr.push_back({t.ids.qf[i], Ref{pq[i].dh1.arena, pq[i].dh1.off + (long long)B * H}});
          crt.push_back({t.ids.qf[i], pq[i].dh1});
          das.push_back(Ref{pq[i].da.arena, pq[i].da.off + (long long)B * pad4(A)});
          dast.push_back(pq[i].da);
      }
      s.pg.push_back(pg_task(cr, pol, pa, 0, pg, false, das));
      s.pg.push_back(pg_task(crt, tpol, tpa, 0, tpg, false, dast));
      fill_pgp(s); }
    if (tensor_glue) { Stage& s = add_stage(ST_GEMM, "policy_dh2"); pol_dh2(s, pol, pa, 0, pg); pol_dh2(s, tpol, tpa, 0, tpg); }
    { Stage& s = add_stage(ST_GEMM, "policy_dh1"); pol_dh1(s, pol, pa, 0, pg); pol_dh1(s, tpol, tpa, 0, tpg); }
    { Stage& s = add_stage(ST_GEMM, "policy_adam");
      pol_adam(s, pol, pa, 0, 2, pg, c.policy_lr, 0); pol_adam(s, tpol, tpa, 0, 2, tpg, c.policy_lr, 1); }
    flush_adam("policy_adam_apply");
}

// ------------------------------------------------------------------------------------
// finalize: tile counts, device tables
// ------------------------------------------------------------------------------------
template <typename T>
static int upload(OacTrainer& t, const T* host, size_t count, void** dev) {
    OAC_CUDA(cudaMalloc(dev, sizeof(T) * count));
    t.dev_allocs.push_back(*dev);
    OAC_CUDA(cudaMemcpy(*dev, host, sizeof(T) * count, cudaMemcpyHostToDevice));
    return 0;
}

// ------------------------------------------------------------------------------------
// warp-specialised tcgen05 path: eligibility, tile plan, tensor maps
// ------------------------------------------------------------------------------------
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
    static TensorMapEncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        cudaDriverEntryPointQueryResult q;
        void* p = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (TensorMapEncodeFn)p;
        cudaGetLastError();
    }
    return fn;
}
static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    }
    return n;
}
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Every operand must be expressible as a [seed][row][col] fp32 tensor with 16-byte aligned base and strides, and the
// float4 epilogue needs the same of the outputs.
static bool ws_eligible(const OacTrainer& t, const Stage& s) {
    const ArenaSet& as = t.as;
    const int seeds = t.cfg.n_seeds;
    if ((int)s.gemm.size() > WS_MAX_TASKS) return false;
    for (int a = 0; a < AR_COUNT; ++a)
        if (seeds > 1 && (as.stride[a] & 3)) return false;
    for (const GemmTask& g : s.gemm) {
        if ((g.lda & 3) || (g.ldb & 3) || (g.ldc & 3)) return false;
        if (!al16(resolve(as, g.A, 0)) || !al16(resolve(as, g.B, 0)) || !al16(resolve(as, g.C, 0))) return false;
        if ((g.epi == EPI_BIAS || g.epi == EPI_BIAS_RELU) && !al16(resolve(as, g.bias, 0))) return false;
        if (g.epi == EPI_MASK && ((g.ldmask & 3) || !al16(resolve(as, g.mask, 0)) || g.a_trans || !g.b_trans)) return false;
        if (g.epi == EPI_GRAD && !(g.a_trans && g.b_trans)) return false;
        if (g.epi == EPI_ADAM) {
            if ((g.adam_off & 3) || (g.target_off >= 0 && (g.target_off & 3))) return false;
            if (!al16(as.base[AR_ADAM_M]) || !al16(as.base[AR_ADAM_V]) || !al16(as.base[AR_PARAM])) return false;
            if (!(g.a_trans && g.b_trans)) return false;       // the bias-gradient MMA assumes the dW operand layouts
        }
    }
    return tensor_map_encoder() != nullptr;
}

// L2 promotion of the operand maps: 256 bytes (64 seeds: critic_l1 82.4 -> 80.8 us, critic_l2 69.9 -> 66.7 us, step 0.708 ->
// 0.699 ms; nothing else moves; 8 seeds unchanged).  OAC_WS_L2PROMO=64|128|256 overrides (measurement aid).
static CUtensorMapL2promotion ws_l2_promotion() {
    const char* e = getenv("OAC_WS_L2PROMO");
    const int v = e ? atoi(e) : 256;
    return v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
}
// element type of the operand maps (measurement aid: OAC_WS_F32MAPS=1 copies plain fp32 -- the MMA then truncates -- to
// see what the in-flight fp32 -> tf32 rounding of the TFLOAT32 type costs)
static CUtensorMapDataType ws_map_dtype() {
    const char* e = getenv("OAC_WS_F32MAPS");
    return (e && e[0] == '1') ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
}
static int ws_plan(OacTrainer& t, Stage& s) {
    const int seeds = t.cfg.n_seeds;
    // CTA pairs (gemm_ws2.cuh) when every task is made of whole 256-row tiles: each CTA then loads half of the B tile
    // Measured (64 seeds, B200): the first-layer forward stages gain 13 % (critic_l1 90 -> 77 us: K = 393, the weight matrix is
    // two thirds of a tile's operand bytes), K = 256 stages and the MN-major products gain nothing, and with few seeds the pair
    // costs time (one round of tiles either way, half as many tiles in flight: 8 seeds 0.222 -> 0.235 ms per step).  So: K-major
    // stages with K >= 320 and at least two rounds of work.  OAC_WS2_ALL=1 lifts the restriction (tests, A/B runs).
    const bool ws2_all = getenv("OAC_WS2_ALL") && getenv("OAC_WS2_ALL")[0] == '1';
    bool pair = t.allow_ws2;
    {
        int kmax = 0;
        long long tiles256 = 0;
        for (auto& g : s.gemm) {
            pair = pair && (g.M % (2 * WS_BM) == 0) && g.epi != EPI_ADAM;
            kmax = std::max(kmax, g.K);
            tiles256 += (long long)(g.M / (2 * WS_BM)) * ((g.N + 255) / 256);
        }
        if (!ws2_all) pair = pair && !s.a_trans && !s.b_trans && kmax >= 320 && tiles256 * seeds >= 2ll * (sm_count() / 2);
    }
    const int BM = pair ? 2 * WS_BM : WS_BM;
    const int sms = s.sm_budget > 0 ? std::min(s.sm_budget, sm_count()) : sm_count();
    const int n_units = pair ? sms / 2 : sms;                  // CTAs or CTA pairs working at once
    const int gran = (s.b_trans ? 32 : 16) * (pair ? 2 : 1);   // MN-major B tiles come in 32-column TMA boxes (per CTA)
    auto width = [&](const GemmTask& g, int cap) {             // tile width for a column cap: even split, rounded up
        const int lim = ((g.epi == EPI_ADAM || g.epi == EPI_GRAD) && g.has_bias) ? std::min(cap, pair ? 192 : (int)WS_BN_MAX_BIAS) : cap;   // room for the bias MMA's columns
        const int tn = (g.N + lim - 1) / lim;
        int bn = (((g.N + tn - 1) / tn) + gran - 1) / gran * gran;
        return std::max(bn, gran);
    };
    auto tiles = [&](int cap) {
        long long n = 0;
        for (auto& g : s.gemm) { const int bn = width(g, cap); n += (long long)((g.M + BM - 1) / BM) * ((g.N + bn - 1) / bn); }
        return n;
    };
    // Tile width.  With many rounds of the persistent grid (>= 4 at the widest tile) the widest tiles that still give every
    // SM work win: operand re-reads fall with the width.  With fewer, what counts is the number of rounds times the cost
    // of a tile, which has a fixed part (pipeline fill, TMEM round trip, epilogue set-up: worth ~64 columns) next to the
    // one that grows with the width (measured: 8 seeds 27.9k -> 31.4k seed-updates/s, 16 seeds 43.8k -> 48.7k, 32 seeds
    // 60.9k -> 63.0k, 64 seeds unchanged).
    int cap = 256;
    const int cap_min = pair ? 64 : 32;
    while (cap > cap_min && tiles(cap) * seeds < n_units) cap >>= 1;
    if (tiles(256) * seeds < 4ll * n_units) {
        double best = 1e30;
        for (int cnd = 256; cnd >= cap_min; cnd >>= 1) {
            const long long n = tiles(cnd) * seeds;
            const long long rounds = (n + n_units - 1) / n_units;
            const double cost = rounds * (64.0 + cnd);
            if (cost < best) { best = cost; cap = cnd; }
        }
    }
    // heaviest tiles first (epilogue elements dominate; an Adam element moves 8x the bytes of a stored one)
    auto tile_cost = [&](const GemmTask& g) {
        const double rows = std::min(g.M, BM), cols = std::min(g.N, width(g, cap));
        return rows * cols * (g.epi == EPI_ADAM ? 8.0 : (g.epi == EPI_MASK ? 2.0 : 1.0)) + 0.05 * WS_BM * cols * g.K / 32.0;
    };
    std::stable_sort(s.gemm.begin(), s.gemm.end(), [&](const GemmTask& a, const GemmTask& b) { return tile_cost(a) > tile_cost(b); });
    int t0 = 0, bn_max = 0;
    for (auto& g : s.gemm) {
        g.bn = width(g, cap);
        g.tiles_m = (g.M + BM - 1) / BM; g.tiles_n = (g.N + g.bn - 1) / g.bn;
        g.tile0 = t0; t0 += g.tiles_m * g.tiles_n;
        bn_max = std::max(bn_max, g.bn);
    }
    int n_mb = 0;
    for (auto& g : s.gemm) {
        if ((g.ldbits > 0 || g.mask_bits) && (g.bn & 31)) return set_error(OAC_E_BITS_UNAVAILABLE, "sign-bit masks need 32-column aligned tiles");
        n_mb += g.mask_bits ? 1 : 0;
    }
    if (n_mb != 0 && (n_mb != (int)s.gemm.size() || pair)) return set_error(OAC_E_BITS_UNAVAILABLE, "sign-bit masks: mixed stage");
    s.ws_mask_bits = n_mb != 0;
    for (auto& g : s.gemm) if (g.epi == EPI_ADAM) s.ws_adam = 1;
    s.ws_tiles_per_seed = t0;
    s.ws_pair = pair ? 1 : 0;
    s.ws_slot_bytes = (int)WS_A_BYTES + (pair ? bn_max / 2 : bn_max) * (WS_KC * 4);
    const int budget = 224 * 1024 - 1024 - (int)WS_ONES_BYTES - (int)WS_SLAB_BYTES;
    s.ws_slots = std::min((int)WS_MAX_SLOTS, budget / s.ws_slot_bytes);
    if (const char* ms = getenv("OAC_WS_MAX_SLOTS")) s.ws_slots = std::max(2, std::min(s.ws_slots, atoi(ms)));     // measurement aid: ring depth
    if (s.ws_slots < 2) return set_error(OAC_E_INVALID, "internal: ws ring does not fit");
    s.smem = (size_t)s.ws_slots * s.ws_slot_bytes + WS_ONES_BYTES + WS_SLAB_BYTES + 1024;
    s.ws_grid = pair ? 2 * (int)std::min<long long>((long long)t0 * seeds, n_units)
                     : (int)std::min<long long>((long long)t0 * seeds, sms);
    // tensor maps
    TensorMapEncodeFn enc = tensor_map_encoder();
    std::vector<CUtensorMap> maps(2 * s.gemm.size());
    for (size_t i = 0; i < s.gemm.size(); ++i) {
        const GemmTask& g = s.gemm[i];
        for (int op = 0; op < 2; ++op) {
            const Ref r = op == 0 ? g.A : g.B;
            const bool mn = op == 0 ? g.a_trans != 0 : g.b_trans != 0;
            const int ld = op == 0 ? g.lda : g.ldb;
            const int ext = op == 0 ? g.M : g.N;               // M / N extent of this operand
            const long long sstride = std::max<long long>(t.as.stride[r.arena], 4);
            cuuint64_t dims[4], strides[3];
            cuuint32_t box[4], es[4] = {1, 1, 1, 1};
            const cuuint32_t box_mn = op == 0 ? WS_BM : (cuuint32_t)(pair ? g.bn / 2 : g.bn);   // M / N extent of this CTA's box
            int rank = 3;
            if (!mn) {
                dims[0] = (cuuint64_t)g.K; dims[1] = (cuuint64_t)ext; box[0] = WS_KC; box[1] = box_mn;
                dims[2] = (cuuint64_t)seeds; box[2] = 1;
                strides[0] = (cuuint64_t)ld * 4; strides[1] = (cuuint64_t)sstride * 4;
            } else {
                // M/N-contiguous operand: [seed][atom of 32 columns][k][32], so ONE box {32, WS_KC, atoms} lands as the
                // UMMA layout [atom][k][128 B].  A ragged last atom reads up to 31 floats past the extent (the rest of the
                // row pitch / the next row: finite data inside the arena, see the slack in oac_trainer_layout); those
                // columns only feed output rows / columns that the epilogue does not store.
                rank = 4;
                dims[0] = 32; dims[1] = (cuuint64_t)g.K; dims[2] = (cuuint64_t)((ext + 31) / 32); dims[3] = (cuuint64_t)seeds;
                box[0] = 32; box[1] = WS_KC; box[2] = box_mn / 32; box[3] = 1;
                strides[0] = (cuuint64_t)ld * 4; strides[1] = 128; strides[2] = (cuuint64_t)sstride * 4;
            }
            // TFLOAT32: the TMA unit rounds fp32 -> tf32 (nearest) on the way into shared memory
            CUresult rc = enc(&maps[2 * i + op], ws_map_dtype(), rank, (void*)resolve(t.as, r, 0), dims, strides,
                              box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                              ws_l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rc != CUDA_SUCCESS) {
                char msg[160];
                snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (%d) for stage %s task %d operand %d", (int)rc, s.name, (int)i, op);
                return set_error(OAC_E_INVALID, msg);
            }
        }
    }
    if (int e = upload(t, maps.data(), maps.size(), &s.ws_tmaps)) return e;
    if (int e = upload(t, s.gemm.data(), s.gemm.size(), &s.dev)) return e;
    s.use_ws = 1;
    return 0;
}

// Strip-fused forward chain (gemm_chain.cuh): tasks are [layer][chain]; every chain's layers share M, hidden layers are H = 256
// wide (one 256-column tile = the whole A operand of the next layer), a head layer is rounded up to 16 columns.
static int chain_plan(OacTrainer& t, Stage& s) {
    const int seeds = t.cfg.n_seeds;
    const int NL = s.chain, NC = (int)s.gemm.size() / NL;
    const bool bwd = s.chain_bwd != 0;
    if (NL < 2 || NL > CH_MAX_LAYERS || (bwd && NL != 2) || NC < 1 || NC * NL != (int)s.gemm.size() || NC > WS_MAX_TASKS || !ws_eligible(t, s))
        return set_error(OAC_E_CHAIN_UNAVAILABLE, "chain: shape");
    int strips = 0;
    for (int c = 0; c < NC; ++c) {
        const GemmTask& g0 = s.gemm[c];
        for (int l = 0; l < NL; ++l) {
            GemmTask& g = s.gemm[l * NC + c];
            const bool last = l == NL - 1;
            if (g.a_trans || (g.b_trans != 0) != bwd || g.M != g0.M || (g.M % WS_BM) != 0) return set_error(OAC_E_CHAIN_UNAVAILABLE, "chain: layout");
            if (l > 0) {      // reads the previous layer's output in full
                const GemmTask& pv = s.gemm[(l - 1) * NC + c];
                if (g.K != pv.N || g.A.arena != pv.C.arena || g.A.off != pv.C.off || g.K != 256) return set_error(OAC_E_CHAIN_UNAVAILABLE, "chain: link");
            }
            if (g.N > 256) return set_error(OAC_E_CHAIN_UNAVAILABLE, "chain: width");
            if (!bwd) {
                if (!last && (g.N != 256 || g.epi != EPI_BIAS_RELU)) return set_error(OAC_E_CHAIN_UNAVAILABLE, "forward chain: hidden layer");
                if (g.epi != EPI_BIAS_RELU && g.epi != EPI_BIAS) return set_error(OAC_E_CHAIN_UNAVAILABLE, "forward chain: epilogue");
                g.bn = last ? std::max(16, (g.N + 15) / 16 * 16) : 256;
            } else {
                // masks are sign bytes (one uint2 per row and 32 columns, lane = row), stores are plain
                if (!last && (g.N != 256 || g.epi != EPI_MASK)) return set_error(OAC_E_CHAIN_UNAVAILABLE, "backward chain: hidden layer");
                if (g.epi != EPI_MASK && g.epi != EPI_STORE) return set_error(OAC_E_CHAIN_UNAVAILABLE, "backward chain: epilogue");
                if (g.epi == EPI_MASK && (!g.mask_bits || (g.ldmask & 7) || g.N != 256)) return set_error(OAC_E_CHAIN_UNAVAILABLE, "backward chain: mask");
                if (last && g.no_store) return set_error(OAC_E_CHAIN_UNAVAILABLE, "backward chain: nothing stored");
                g.bn = last ? std::max(32, (g.N + 31) / 32 * 32) : 256;       // M/N-contiguous B: whole 32-column atoms
            }
            g.tiles_m = g.M / WS_BM; g.tiles_n = 1; g.tile0 = strips;
        }
        s.chain_strips0[c] = strips;
        strips += g0.M / WS_BM;
    }
    s.chain_strips0[NC] = strips;
    s.ws_tiles_per_seed = strips;
    s.ws_slot_bytes = (int)WS_A_BYTES + 256 * (WS_KC * 4);
    const int budget = 224 * 1024 - 1024 - (int)WS_SLAB_BYTES;
    s.ws_slots = std::min((int)WS_MAX_SLOTS, budget / s.ws_slot_bytes);
    if (s.ws_slots < 2) return set_error(OAC_E_CHAIN_UNAVAILABLE, "chain: ring");
    s.smem = (size_t)s.ws_slots * s.ws_slot_bytes + WS_SLAB_BYTES + 1024;
    s.ws_grid = (int)std::min<long long>((long long)strips * seeds, sm_count());
    TensorMapEncodeFn enc = tensor_map_encoder();
    std::vector<CUtensorMap> maps(2 * s.gemm.size());
    for (size_t i = 0; i < s.gemm.size(); ++i) {
        const GemmTask& g = s.gemm[i];
        for (int op = 0; op < 2; ++op) {
            const Ref r = op == 0 ? g.A : g.B;
            const int ld = op == 0 ? g.lda : g.ldb;
            const int ext = op == 0 ? g.M : g.N;
            const long long sstride = std::max<long long>(t.as.stride[r.arena], 4);
            CUresult rc;
            if (op == 1 && bwd) {
                // M/N-contiguous weights: [seed][atom of 32 columns][k][32], one box = all atoms of the tile (ws_plan)
                cuuint64_t dims[4] = {32, (cuuint64_t)g.K, (cuuint64_t)((ext + 31) / 32), (cuuint64_t)seeds};
                cuuint64_t strides[3] = {(cuuint64_t)ld * 4, 128, (cuuint64_t)sstride * 4};
                cuuint32_t box[4] = {32, WS_KC, (cuuint32_t)(g.bn / 32), 1}, es[4] = {1, 1, 1, 1};
                rc = enc(&maps[2 * i + op], ws_map_dtype(), 4, (void*)resolve(t.as, r, 0), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, ws_l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            } else {
                cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)ext, (cuuint64_t)seeds};
                cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)sstride * 4};
                cuuint32_t box[3] = {WS_KC, op == 0 ? (cuuint32_t)WS_BM : (cuuint32_t)g.bn, 1}, es[3] = {1, 1, 1};
                rc = enc(&maps[2 * i + op], ws_map_dtype(), 3, (void*)resolve(t.as, r, 0), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, ws_l2_promotion(), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            }
            if (rc != CUDA_SUCCESS) return set_error(OAC_E_CHAIN_UNAVAILABLE, "chain: tensor map");
        }
    }
    if (int e = upload(t, maps.data(), maps.size(), &s.ws_tmaps)) return e;
    if (int e = upload(t, s.gemm.data(), s.gemm.size(), &s.dev)) return e;
    s.use_ws = 1;
    return 0;
}

// Glue kernels: G warps per row (4: a single seed's few hundred rows still fill the chip; 1: self-contained warps
// for the many-seed launches) and, since every CTA stages head / action-column weights once, several row groups
// per CTA when there are enough rows to keep ~4 CTAs on every SM anyway.
static void glue_plan(Stage& s, long long rows_total, bool stages_weights) {
    s.glue_g = rows_total >= 4096 ? 1 : 4;
    const int spc = GLUE_WARPS / s.glue_g;
    const long long groups = (rows_total + spc - 1) / spc;
    const long long want = groups / (4ll * sm_count());
    s.glue_iters = stages_weights ? (int)std::max<long long>(1, std::min<long long>(16, want)) : 1;
}

// Latency-regime FFMA tile with TMA staging: fp32 tensor maps (no conversion) of both operands of every task, 32 x 32
// boxes in SWIZZLE_128B for K-contiguous operands, 32-wide row boxes for M/N-contiguous ones.  Falls back silently
// (cp.async staging) when an operand cannot be described (row stride not a multiple of 16 B, unaligned base) or the
// whole-K tiles of a task do not fit.
static int sk_tma_plan(OacTrainer& t, Stage& s) {
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc) return 0;
    const int seeds = t.cfg.n_seeds;
    size_t smem = 0;
    for (int a = 0; a < AR_COUNT; ++a)
        if (seeds > 1 && (t.as.stride[a] & 3)) return 0;
    for (const GemmTask& g : s.gemm) {
        if ((g.lda & 3) || (g.ldb & 3) || !al16(resolve(t.as, g.A, 0)) || !al16(resolve(t.as, g.B, 0))) return 0;
        smem = std::max(smem, (size_t)sk_tile_bytes(g.a_trans != 0, g.K) + sk_tile_bytes(g.b_trans != 0, g.K));
    }
    smem = std::max(smem, sizeof(float) * SK_KS * SK_BM * SK_PLD) + 1024;
    if (smem > 110 * 1024) return 0;                           // keep two CTAs per SM
    std::vector<CUtensorMap> maps(2 * s.gemm.size());
    for (size_t i = 0; i < s.gemm.size(); ++i) {
        const GemmTask& g = s.gemm[i];
        for (int op = 0; op < 2; ++op) {
            const Ref r = op == 0 ? g.A : g.B;
            const bool mn = op == 0 ? g.a_trans != 0 : g.b_trans != 0;
            const int ld = op == 0 ? g.lda : g.ldb;
            const int ext = op == 0 ? g.M : g.N;
            const long long sstride = std::max<long long>(t.as.stride[r.arena], 4);
            cuuint64_t dims[3], strides[2];
            cuuint32_t box[3], es[3] = {1, 1, 1};
            if (!mn) { dims[0] = (cuuint64_t)g.K; dims[1] = (cuuint64_t)ext; box[0] = 32; box[1] = 32; }
            else     { dims[0] = (cuuint64_t)ext; dims[1] = (cuuint64_t)g.K; box[0] = 32; box[1] = (cuuint32_t)sk_rows_per_box(g.K); }
            dims[2] = (cuuint64_t)seeds; box[2] = 1;
            strides[0] = (cuuint64_t)ld * 4; strides[1] = (cuuint64_t)sstride * 4;
            CUresult rc = enc(&maps[2 * i + op], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)resolve(t.as, r, 0), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, mn ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rc != CUDA_SUCCESS) return 0;
        }
    }
    if (int e = upload(t, maps.data(), maps.size(), &s.ws_tmaps)) return e;
    s.sk_tma = 1;
    s.smem = smem;
    return 0;
}

static int finalize(OacTrainer& t) {
    const int seeds = t.cfg.n_seeds;
    for (Stage& s : t.stages) {
        if (s.kind == ST_GEMM) {
            int kmax = 0;
            s.a_trans = s.gemm[0].a_trans; s.b_trans = s.gemm[0].b_trans;
            for (auto& g : s.gemm) {
                kmax = std::max(kmax, g.K);
                if (g.a_trans != s.a_trans || g.b_trans != s.b_trans || (g.a_trans && !g.b_trans))
                    return set_error(OAC_E_INVALID, "internal: mixed operand layouts in one GEMM stage");
            }
            if (s.fused2 && t.cfg.gemm_path != OAC_GEMM_FP32)
                return set_error(OAC_E_INVALID, "internal: fused layer pairs exist on the FFMA path only");
            if (s.chain) {
                if (t.cfg.gemm_path != OAC_GEMM_TF32) return set_error(OAC_E_CHAIN_UNAVAILABLE, "forward chain: gemm path");
                if (int e = chain_plan(t, s)) return e;
                continue;
            }
            if (t.cfg.gemm_path == OAC_GEMM_TF32 || t.cfg.gemm_path == OAC_GEMM_TF32X3) {
                const bool x3 = t.cfg.gemm_path == OAC_GEMM_TF32X3;
                if (!x3 && t.allow_ws && ws_eligible(t, s)) {
                    if (int e = ws_plan(t, s)) return e;
                    continue;
                }
                for (auto& g : s.gemm)
                    if (g.ldbits > 0 || g.mask_bits) return set_error(OAC_E_BITS_UNAVAILABLE, "sign-bit masks need the TMA path");
                for (auto& g : s.gemm)
                    if (g.epi == EPI_GRAD) return set_error(OAC_E_SPLIT_UNAVAILABLE, "gradient-store stages need the TMA path");
                // tcgen05 path: 128 x BN tiles.  Pick the largest BN that still fills the chip.
                s.use_tc = 1;
                const int bn_min = s.b_trans ? 32 : 16;
                int nmax = 0;
                for (auto& g : s.gemm) nmax = std::max(nmax, g.N + ((g.epi == EPI_ADAM && g.has_bias) ? 1 : 0));
                auto ctas = [&](int bn) {
                    long long n = 0;
                    for (auto& g : s.gemm) {
                        int ncols = g.N + ((g.epi == EPI_ADAM && g.has_bias) ? 1 : 0);
                        n += (long long)((g.M + 127) / 128) * ((ncols + bn - 1) / bn);
                    }
                    return n * seeds;
                };
                // Two regimes.  Latency (the grid fits ~2 waves even with 64-wide tiles: single seed): one CTA per
                // SM, the narrowest tile whose grid still fits ONE wave of 148 CTAs (a partial second wave doubles
                // the stage time), 128-deep K chunks.  Throughput (many seeds): the widest tile the N extent allows
                // (one A-tile staging feeds up to 256 output columns), 32-deep K chunks so the 2-slot ring is <= 96 KB
                // and TWO CTAs share an SM (one stages / rounds while the other's MMAs and epilogue run).
                int bn = 64;
                const bool throughput = !x3 && ctas(64) > 2 * 148;
                if (throughput) {
                    bn = bn_min;
                    while (bn < 256 && bn < nmax) bn <<= 1;
                    s.kc = 32;
                } else {
                    while (bn > bn_min) {
                        if (nmax <= bn / 2) { bn >>= 1; continue; }        // narrower tile is free
                        if (ctas(bn / 2) <= 148) { bn >>= 1; continue; }   // more CTAs, still one wave
                        break;
                    }
                    s.kc = x3 ? 64 : 128;
                }
                s.bn = bn;
                if (x3) {
                    const int nchunks = (kmax + s.kc - 1) / s.kc;
                    s.n_main = std::max(1, std::min(nchunks, 512 / bn - 1));
                }
                int cols = (x3 ? (s.n_main + 1) : 1) * bn;
                s.tmem_cols = 32;
                while (s.tmem_cols < cols) s.tmem_cols <<= 1;
                s.smem = 2 * (size_t)(128 + bn) * s.kc * sizeof(float) * (x3 ? 2 : 1) + 1024;
                s.max_tiles = 0;
                for (auto& g : s.gemm) {
                    int ncols = g.N + ((g.epi == EPI_ADAM && g.has_bias) ? 1 : 0);
                    g.tiles_m = (g.M + 127) / 128; g.tiles_n = (ncols + bn - 1) / bn;
                    s.max_tiles = std::max(s.max_tiles, g.tiles_m * g.tiles_n);
                }
                if (int e = upload(t, s.gemm.data(), s.gemm.size(), &s.dev)) return e;
                continue;
            }
            // one FFMA tile shape: 32 x 32 outputs, 4-way split-K (gemm_sk_kernel).  A 64 x 64 / 4 x 4-per-thread variant
            // for large grids was measured slower at every seed count (64 seeds: 4.32 vs 3.26 ms per step) and is gone.
            const int bm = 32;
            // largest K chunk (multiple of 4) whose A+B tiles fit the shared-memory budget
            auto bytes_of = [&](int kc) {
                size_t a = s.a_trans ? (size_t)kc * bm : (size_t)bm * kpad_of(kc);
                size_t b = s.b_trans ? (size_t)kc * bm : (size_t)bm * kpad_of(kc);
                return (a + b) * sizeof(float);
            };
            int kc = (kmax + 3) & ~3;
            const size_t budget = 100 * 1024;                          // keep 2 CTAs / SM
            while (kc > 16 && bytes_of(kc) > budget) kc = ((kc / 2) + 3) & ~3;
            s.kc = kc; s.smem = bytes_of(kc);
            s.smem = std::max(s.smem, sizeof(float) * SK_KS * SK_BM * SK_PLD);   // the k-groups' partial tiles
            if (t.allow_sk_tma) { if (int e = sk_tma_plan(t, s)) return e; }
            s.max_tiles = 0;
            for (auto& g : s.gemm) {
                g.tiles_m = (g.M + bm - 1) / bm; g.tiles_n = (g.N + bm - 1) / bm;
                s.max_tiles = std::max(s.max_tiles, g.tiles_m * g.tiles_n);
            }
            if (s.fused2) {
                // one cluster launch when every pair is (K-contiguous, bias[/relu]) x 2 with the same strip structure
                const size_t np = s.gemm.size() / 2;
                bool ok = s.sk_tma && !s.a_trans && !s.b_trans && s.gemm.size() % 2 == 0 && np > 0;
                const int cl = ok ? s.gemm[0].tiles_n : 0;
                ok = ok && cl >= 1 && cl <= 8;
                for (size_t i = 0; ok && i < np; ++i) {
                    const GemmTask& a = s.gemm[i]; const GemmTask& b = s.gemm[np + i];
                    ok = a.M == b.M && a.tiles_n == cl && b.tiles_n == cl && b.K == a.N && b.A.arena == a.C.arena && b.A.off == a.C.off &&
                         (a.epi == EPI_BIAS || a.epi == EPI_BIAS_RELU) && (b.epi == EPI_BIAS || b.epi == EPI_BIAS_RELU);
                }
                s.fwd2_cluster = ok ? cl : 0;
            }
            if (int e = upload(t, s.gemm.data(), s.gemm.size(), &s.dev)) return e;
        } else if (s.kind == ST_POLICY_HEAD) {
            s.max_rows = 0;
            for (auto& p : s.ph) s.max_rows = std::max(s.max_rows, p.rows);
            glue_plan(s, (long long)s.max_rows * s.ph.size() * seeds, !s.php.head_from_gemm);
            s.many = s.php.head_from_gemm && !getenv("OAC_NO_GLUE_MANY");      // (env: A/B measurement aid)
            if (int e = upload(t, s.ph.data(), s.ph.size(), &s.dev)) return e;
            s.php.tasks = (const PolicyHeadTask*)s.dev;
            s.php.as = t.as; s.php.hyper = t.hyper;
        } else if (s.kind == ST_STEP_TAIL) {
            if (int e = upload(t, s.ph.data(), s.ph.size(), &s.dev)) return e;
            s.php.tasks = (const PolicyHeadTask*)s.dev;
            s.php.as = t.as; s.php.hyper = t.hyper;
        } else if (s.kind == ST_CRITIC_HEAD) {
            glue_plan(s, (long long)t.cfg.batch * seeds, false);
            s.chp.iters = s.glue_iters;
            {
                int n = 0;
                for (int i = 0; i < s.chp.n_src; ++i) {
                    s.chp.goff[i] = n;
                    for (int hd = 0; hd < s.chp.src[i].n_heads; ++hd, ++n) {
                        if (n >= MAX_VALS) return set_error(OAC_E_UNSUPPORTED, "too many critic heads for one critic_head stage");
                        s.chp.pair_src[n] = (short)i; s.chp.pair_hd[n] = (short)hd;
                    }
                }
                s.chp.n_pairs = n;
            }
            s.chp.as = t.as;
            {
                // the lean SAC kernel (one self-contained warp per sample: glue_many.cuh) also wins in the single-seed latency
                // regime: critic_head 8.3 -> 6.2 us, step 118.9 -> 115.8 us (OAC_CH_LEAN_ALL=0 keeps the generic kernel there)
                static const bool lean_all = !(getenv("OAC_CH_LEAN_ALL") && getenv("OAC_CH_LEAN_ALL")[0] == '0');
                bool ok = s.chp.mode == CM_SAC && t.cfg.hidden == 256 &&
                          (lean_all || (t.cfg.gemm_path == OAC_GEMM_TF32 && (long long)seeds * t.cfg.batch >= 2048)) &&
                          s.chp.n_src == 6 && !getenv("OAC_NO_GLUE_MANY");
                for (int i = 0; i < s.chp.n_src && ok; ++i) ok = s.chp.src[i].n_heads == 1;
                for (int a = 0; a < AR_COUNT && ok; ++a) ok = (t.as.stride[a] & 3) == 0 && al16(t.as.base[a]);
                s.many = ok;
            }
            if (int e = upload(t, &s.chp, 1, &s.dev)) return e;
        } else if (s.kind == ST_RANK1) {
            if (int e = upload(t, s.r1.data(), s.r1.size(), &s.dev)) return e;
        } else if (s.kind == ST_ADAM) {
            s.asp.as = t.as; s.asp.hyper = t.hyper;
            if (int e = upload(t, &s.asp, 1, &s.dev)) return e;
        } else if (s.kind == ST_POLICY_GRAD) {
            glue_plan(s, (long long)t.cfg.batch * s.pg.size() * seeds, !s.pgp.da_from_gemm);
            s.many = s.pgp.da_from_gemm && !getenv("OAC_NO_GLUE_MANY");
            if (int e = upload(t, s.pg.data(), s.pg.size(), &s.dev)) return e;
            s.pgp.tasks = (const PolicyGradTask*)s.dev;
            s.pgp.as = t.as;
        }
    }
    return 0;
}

// Launch with Programmatic Dependent Launch: each kernel waits for its predecessor first thing (griddepcontrol.wait)
// and releases its successor after its main loop (griddepcontrol.launch_dependents), so the successor's launch latency
// and CTA set-up could overlap the epilogue and drain of the current stage.  Measured on B200 inside the CUDA graph it is
// a LOSS either way (trigger at kernel entry: 209 vs 181 us per step; trigger after the main loop: 179 vs 164 us), so
// the attribute is only set with OAC_PDL=1.
static bool g_use_pdl = false;
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// critics whose fc0 action columns the policy_grad kernel stages together: both of a twin pair, else one at a time
static int pg_wa_sources(const OacTrainer& t, const Stage& s) {
    int n = 0;
    for (const auto& g : s.pg) n = std::max(n, g.n_src);
    (void)t;
    return n >= 2 ? 2 : 1;
}
static size_t glue_smem(const OacTrainer& t, const Stage& s) {
    const int A_ = t.cfg.act_dim, H_ = t.cfg.hidden;
    const int spc = GLUE_WARPS / s.glue_g, per_cta = spc * s.glue_iters;
    if (s.kind == ST_POLICY_HEAD)
        return sizeof(float) * (s.php.head_from_gemm ? (size_t)2 * A_ * spc : (size_t)2 * A_ * H_ + 2 * A_ * (1 + spc));
    if (s.kind == ST_POLICY_GRAD)
        return sizeof(float) * ((s.pgp.da_from_gemm ? 0 : (size_t)pg_wa_sources(t, s) * H_ * (A_ | 1) + (size_t)2 * A_ * H_) +
                                (size_t)per_cta * 3 * A_);
    return 0;
}

static int launch_stage(OacTrainer& t, Stage& s, int use_external_eps, cudaStream_t st);

// Runs the stage list.  Lane-1 stages go to the side stream: it is forked from the main stream at the first lane-1
// stage after a join (event record / wait, which also works under stream capture: the side stream joins the capture),
// and joined back before a stage that asks for it and at the end of the step.
static int launch_stages(OacTrainer& t, int use_external_eps, cudaStream_t main_st) {
    bool busy[2] = {false, false};
    auto join = [&](int mask) -> int {
        for (int l = 0; l < 2; ++l) {
            if (!(mask & (1 << l)) || !busy[l]) continue;
            OAC_CUDA(cudaEventRecord(t.ev_join[l], t.side[l]));
            OAC_CUDA(cudaStreamWaitEvent(main_st, t.ev_join[l], 0));
            busy[l] = false;
        }
        return 0;
    };
    for (Stage& s : t.stages) {
        cudaStream_t st = main_st;
        if (s.lane >= 1 && t.side[s.lane - 1] != nullptr) {
            const int l = s.lane - 1;
            // the side lane picks up the main lane's current position (under capture: a dependency edge)
            OAC_CUDA(cudaEventRecord(t.ev_fork[l], main_st));
            OAC_CUDA(cudaStreamWaitEvent(t.side[l], t.ev_fork[l], 0));
            busy[l] = true;
            st = t.side[l];
        } else if (s.join) {
            if (int e = join(s.join)) return e;
        }
        if (int e = launch_stage(t, s, use_external_eps, st)) return e;
    }
    return join(3);
}

static int launch_stage(OacTrainer& t, Stage& s, int use_external_eps, cudaStream_t st) {
    const int seeds = t.cfg.n_seeds;
    {
        if (s.kind == ST_GEMM) {
            StageParams sp; sp.tasks = (const GemmTask*)s.dev; sp.as = t.as; sp.hyper = t.hyper; sp.kc = s.kc;
            sp.tmaps = s.sk_tma ? (const CUtensorMap*)s.ws_tmaps : nullptr;
            dim3 grid(s.max_tiles, (unsigned)s.gemm.size(), seeds);
            if (s.chain) {
                ChainParams cp; cp.sp = sp; cp.sp.tmaps = nullptr; cp.tmaps = (const CUtensorMap*)s.ws_tmaps;
                cp.n_layers = s.chain; cp.n_chains = (int)s.gemm.size() / s.chain;
                for (int i = 0; i <= WS_MAX_TASKS; ++i) cp.strips0[i] = i <= cp.n_chains ? s.chain_strips0[i] : 0;
                cp.n_seeds = seeds; cp.total_items = s.ws_tiles_per_seed * seeds; cp.n_slots = s.ws_slots; cp.slot_bytes = s.ws_slot_bytes;
                if (s.chain_bwd) launch_pdl(gemm_chain_kernel<true>, dim3(s.ws_grid), dim3(WS_THREADS), s.smem, st, cp);
                else launch_pdl(gemm_chain_kernel<false>, dim3(s.ws_grid), dim3(WS_THREADS), s.smem, st, cp);
                OAC_CUDA(cudaGetLastError());
                return 0;
            }
            if (s.use_ws) {
                WsParams wp; wp.sp = sp; wp.tmaps = (const CUtensorMap*)s.ws_tmaps; wp.n_tasks = (int)s.gemm.size();
                wp.tiles_per_seed = s.ws_tiles_per_seed; wp.n_seeds = seeds; wp.total_tiles = s.ws_tiles_per_seed * seeds;
                wp.n_slots = s.ws_slots; wp.slot_bytes = s.ws_slot_bytes;
                const dim3 wg(s.ws_grid), wb(WS_THREADS);
                if (s.ws_pair) {
                    cudaLaunchConfig_t cfg;
                    memset(&cfg, 0, sizeof(cfg));
                    cfg.gridDim = wg; cfg.blockDim = wb; cfg.dynamicSmemBytes = s.smem; cfg.stream = st;
                    cudaLaunchAttribute attr[1];
                    attr[0].id = cudaLaunchAttributeClusterDimension;
                    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                    cfg.attrs = attr; cfg.numAttrs = 1;
                    if (!s.a_trans && !s.b_trans) OAC_CUDA(cudaLaunchKernelEx(&cfg, gemm_ws2_kernel<false, false>, wp));
                    else if (!s.a_trans) OAC_CUDA(cudaLaunchKernelEx(&cfg, gemm_ws2_kernel<false, true>, wp));
                    else OAC_CUDA(cudaLaunchKernelEx(&cfg, gemm_ws2_kernel<true, true>, wp));
                    return 0;
                }
                if (!s.a_trans && !s.b_trans) launch_pdl(gemm_ws_kernel<false, false>, wg, wb, s.smem, st, wp);
                else if (!s.a_trans && s.ws_mask_bits) launch_pdl(gemm_ws_kernel<false, true, true>, wg, wb, s.smem, st, wp);
                else if (!s.a_trans) launch_pdl(gemm_ws_kernel<false, true>, wg, wb, s.smem, st, wp);
                else if (s.ws_adam) launch_pdl(gemm_ws_kernel<true, true, true>, wg, wb, s.smem, st, wp);
                else launch_pdl(gemm_ws_kernel<true, true>, wg, wb, s.smem, st, wp);
                OAC_CUDA(cudaGetLastError());
                return 0;
            }
            if (s.use_tc) {
                TcStageParams tp; tp.sp = sp; tp.bn = s.bn; tp.kc = s.kc; tp.tmem_cols = s.tmem_cols; tp.n_main = s.n_main; tp.dbg = t.tc_dbg;
                const bool x3 = t.cfg.gemm_path == OAC_GEMM_TF32X3;
                if (!s.a_trans && !s.b_trans) {
                    if (x3) launch_pdl(gemm_tc_kernel<false, false, true>, grid, dim3(TC_THREADS), s.smem, st, tp);
                    else launch_pdl(gemm_tc_kernel<false, false, false>, grid, dim3(TC_THREADS), s.smem, st, tp);
                } else if (!s.a_trans) {
                    if (x3) launch_pdl(gemm_tc_kernel<false, true, true>, grid, dim3(TC_THREADS), s.smem, st, tp);
                    else launch_pdl(gemm_tc_kernel<false, true, false>, grid, dim3(TC_THREADS), s.smem, st, tp);
                } else {
                    if (x3) launch_pdl(gemm_tc_kernel<true, true, true>, grid, dim3(TC_THREADS), s.smem, st, tp);
                    else launch_pdl(gemm_tc_kernel<true, true, false>, grid, dim3(TC_THREADS), s.smem, st, tp);
                }
                OAC_CUDA(cudaGetLastError());
                return 0;
            }
            if (s.fused2) {
                const int np = (int)s.gemm.size() / 2;
                grid.y = np;
                if (s.fwd2_cluster) {                  // both layers of every 32-row strip in one cluster launch
                    cudaLaunchConfig_t cfg;
                    memset(&cfg, 0, sizeof(cfg));
                    cfg.gridDim = grid; cfg.blockDim = dim3(SK_THREADS); cfg.dynamicSmemBytes = s.smem; cfg.stream = st;
                    cudaLaunchAttribute attr[1];
                    attr[0].id = cudaLaunchAttributeClusterDimension;
                    attr[0].val.clusterDim.x = s.fwd2_cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                    cfg.attrs = attr; cfg.numAttrs = 1;
                    OAC_CUDA(cudaLaunchKernelEx(&cfg, gemm_fwd2_kernel, sp, np));
                } else {                               // layer 1, then layer 2
                    for (int l = 0; l < 2; ++l) {
                        StageParams sl = sp;
                        sl.tasks = sp.tasks + l * np;
                        if (sl.tmaps) sl.tmaps = sp.tmaps + 2 * l * np;
                        if (s.sk_tma) launch_pdl(gemm_sk_kernel<false, false, true>, grid, dim3(SK_THREADS), s.smem, st, sl);
                        else launch_pdl(gemm_sk_kernel<false, false>, grid, dim3(SK_THREADS), s.smem, st, sl);
                    }
                }
                OAC_CUDA(cudaGetLastError());
                return 0;
            }
            const int sel = s.a_trans ? 2 : (s.b_trans ? 1 : 0);
            switch (sel) {
                case 0: if (s.sk_tma) launch_pdl(gemm_sk_kernel<false, false, true>, grid, dim3(SK_THREADS), s.smem, st, sp);
                        else launch_pdl(gemm_sk_kernel<false, false>, grid, dim3(SK_THREADS), s.smem, st, sp);
                        break;
                case 1: if (s.sk_tma) launch_pdl(gemm_sk_kernel<false, true, true>, grid, dim3(SK_THREADS), s.smem, st, sp);
                        else launch_pdl(gemm_sk_kernel<false, true>, grid, dim3(SK_THREADS), s.smem, st, sp);
                        break;
                case 2: if (s.sk_tma) launch_pdl(gemm_sk_kernel<true, true, true>, grid, dim3(SK_THREADS), s.smem, st, sp);
                        else launch_pdl(gemm_sk_kernel<true, true>, grid, dim3(SK_THREADS), s.smem, st, sp);
                        break;
                default: break;
            }
        } else if (s.kind == ST_RANK1) {
            int cells = 0;
            for (auto& r : s.r1) cells = std::max(cells, ((r.rows + RANK1_ROWS - 1) / RANK1_ROWS) * (r.cols >> 2));
            dim3 grid((cells + RANK1_THREADS - 1) / RANK1_THREADS, (unsigned)s.r1.size(), seeds);
            launch_pdl(rank1_mask_kernel, grid, dim3(RANK1_THREADS), 0, st, (const Rank1Task*)s.dev, t.as);
        } else if (s.kind == ST_POLICY_HEAD && s.many) {
            PolicyHeadParams p = s.php; p.use_external_eps = use_external_eps; p.iters = 1;
            dim3 grid((s.max_rows + GLUE_WARPS - 1) / GLUE_WARPS, (unsigned)s.ph.size(), seeds);
            launch_pdl(policy_head_many_kernel, grid, dim3(GLUE_THREADS), 0, st, p, use_external_eps);
        } else if (s.kind == ST_CRITIC_HEAD && s.many) {
            dim3 grid((t.cfg.batch + GLUE_WARPS - 1) / GLUE_WARPS, seeds, 1);
            launch_pdl(critic_head_sac256_kernel, grid, dim3(GLUE_THREADS), 0, st, (const CriticHeadParams*)s.dev);
        } else if (s.kind == ST_POLICY_GRAD && s.many) {
            PolicyGradParams p = s.pgp; p.iters = 1; p.wa_sources = 1;
            dim3 grid((t.cfg.batch + GLUE_WARPS - 1) / GLUE_WARPS, (unsigned)s.pg.size(), seeds);
            launch_pdl(policy_grad_many_kernel, grid, dim3(GLUE_THREADS), 0, st, p);
        } else if (s.kind == ST_POLICY_HEAD) {
            PolicyHeadParams p = s.php; p.use_external_eps = use_external_eps; p.iters = s.glue_iters;
            const int spc = GLUE_WARPS / s.glue_g, per_cta = spc * s.glue_iters;
            dim3 grid((s.max_rows + per_cta - 1) / per_cta, (unsigned)s.ph.size(), seeds);
            const size_t smem = glue_smem(t, s);
            if (s.glue_g == 1) launch_pdl(policy_head_kernel<1>, grid, dim3(GLUE_THREADS), smem, st, p);
            else launch_pdl(policy_head_kernel<4>, grid, dim3(GLUE_THREADS), smem, st, p);
        } else if (s.kind == ST_STEP_TAIL) {
            launch_pdl(step_tail_kernel, dim3(seeds), dim3(GLUE_THREADS), 0, st, s.php);
        } else if (s.kind == ST_ADAM) {
            const long long per_cta = (long long)ADAM_THREADS * ADAM_UNROLL;
            dim3 grid((unsigned)((s.asp.total4 + per_cta - 1) / per_cta), seeds, 1);
            launch_pdl(adam_stream_kernel<ADAM_UNROLL, false>, grid, dim3(ADAM_THREADS), 0, st, (const AdamStreamParams*)s.dev);
        } else if (s.kind == ST_CRITIC_HEAD) {
            const int per_cta = (GLUE_WARPS / s.glue_g) * s.glue_iters;       // iters is part of the uploaded CriticHeadParams
            dim3 grid((t.cfg.batch + per_cta - 1) / per_cta, seeds, 1);
            if (s.glue_g == 1) launch_pdl(critic_head_kernel<1>, grid, dim3(GLUE_THREADS), 0, st, (const CriticHeadParams*)s.dev);
            else launch_pdl(critic_head_kernel<4>, grid, dim3(GLUE_THREADS), 0, st, (const CriticHeadParams*)s.dev);
        } else {
            PolicyGradParams p = s.pgp; p.iters = s.glue_iters; p.wa_sources = pg_wa_sources(t, s);
            const int per_cta = (GLUE_WARPS / s.glue_g) * s.glue_iters;
            dim3 grid((t.cfg.batch + per_cta - 1) / per_cta, (unsigned)s.pg.size(), seeds);
            const size_t smem = glue_smem(t, s);
            if (s.glue_g == 1) launch_pdl(policy_grad_kernel<1>, grid, dim3(GLUE_THREADS), smem, st, p);
            else launch_pdl(policy_grad_kernel<4>, grid, dim3(GLUE_THREADS), smem, st, p);
        }
        OAC_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // namespace oac

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" int oac_trainer_layout(const OacConfig* cfg, OacLayout* out) {
    if (!cfg || !out) return set_error(OAC_E_INVALID, "null argument");
    OacTrainer tmp;
    tmp.cfg = *cfg;
    NetIds ids;
    if (int e = build_layout(*cfg, tmp.lay, ids)) return e;
    tmp.ids = ids;
    Builder b(tmp);
    if (cfg->algo == OAC_ALGO_SAC) b.build_sac();
    else if (cfg->algo == OAC_ALGO_POAC) b.build_poac();
    else b.build_goac();
    tmp.lay.work_floats = tmp.work_cursor + WORK_SLACK;
    *out = tmp.lay;
    return 0;
}

extern "C" int oac_trainer_create(const OacConfig* cfg, const OacBuffers* buf, OacTrainer** out) {
    if (!cfg || !buf || !out) return set_error(OAC_E_INVALID, "null argument");
    if (!buf->params || !buf->adam_m || !buf->adam_v || !buf->work || !buf->io || !buf->counters)
        return set_error(OAC_E_INVALID, "null buffer");
    if (cfg->gemm_path < OAC_GEMM_FP32 || cfg->gemm_path > OAC_GEMM_TF32X3) return set_error(OAC_E_INVALID, "gemm_path");
    OacTrainer* t = new OacTrainer();
    t->cfg = *cfg;
    t->host_scalars = buf->host_scalars;
    if (int e = build_layout(*cfg, t->lay, t->ids)) { delete t; return e; }
    { const char* nw = getenv("OAC_NO_WS"); t->allow_ws = !(nw && nw[0] == '1'); }
    { const char* nl = getenv("OAC_NO_LANES"); t->allow_lanes = !(nl && nl[0] == '1'); }
    { const char* n2 = getenv("OAC_NO_WS2"); t->allow_ws2 = !(n2 && n2[0] == '1'); }
    { const char* nc = getenv("OAC_NO_CHAIN"); t->allow_chain = !(nc && nc[0] == '1'); }
    { const char* nb = getenv("OAC_NO_MASK_BITS"); t->allow_bits = !(nb && nb[0] == '1'); }
    { const char* nb = getenv("OAC_NO_BWD_CHAIN"); t->allow_bwd_chain = !(nb && nb[0] == '1'); }
    { const char* nk = getenv("OAC_NO_SK_TMA"); t->allow_sk_tma = !(nk && nk[0] == '1'); }
    Builder b(*t);
    if (cfg->algo == OAC_ALGO_SAC) b.build_sac();
    else if (cfg->algo == OAC_ALGO_POAC) b.build_poac();
    else b.build_goac();
    t->lay.work_floats = t->work_cursor + WORK_SLACK;
    const OacLayout& L = t->lay;
    t->as.base[AR_PARAM] = buf->params;  t->as.stride[AR_PARAM] = L.param_floats;
    t->as.base[AR_ADAM_M] = buf->adam_m; t->as.stride[AR_ADAM_M] = L.adam_floats;
    t->as.base[AR_ADAM_V] = buf->adam_v; t->as.stride[AR_ADAM_V] = L.adam_floats;
    t->as.base[AR_WORK] = buf->work;     t->as.stride[AR_WORK] = L.work_floats;
    t->as.base[AR_IO] = buf->io;         t->as.stride[AR_IO] = L.io_floats;
    t->as.counters = buf->counters;      t->as.n_counters = L.n_counters;
    t->hyper.beta1 = cfg->adam_beta1; t->hyper.beta2 = cfg->adam_beta2; t->hyper.eps = cfg->adam_eps;
    t->hyper.tau = cfg->soft_target_tau;
    t->hyper.one_minus_tau = (float)(1.0 - (double)cfg->soft_target_tau);
    t->hyper.target_period = cfg->target_update_period > 0 ? cfg->target_update_period : 1;
    {   // the Python double tau (e.g. 5e-3) is not representable in fp32: recover it like the betas
        double tau = rint((double)cfg->soft_target_tau * 1e9) * 1e-9;
        t->hyper.one_minus_tau = (float)(1.0 - tau);
    }
    {
        const int big = 212 * 1024;
        cudaError_t e = cudaSuccess;
        auto opt_in = [&](const void* f) { if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, big); };
        opt_in((const void*)gemm_sk_kernel<false, false>);
        opt_in((const void*)gemm_sk_kernel<false, false, true>);
        opt_in((const void*)gemm_sk_kernel<false, true, true>);
        opt_in((const void*)gemm_sk_kernel<true, true, true>);
        opt_in((const void*)gemm_sk_kernel<false, true>);
        opt_in((const void*)gemm_sk_kernel<true, true>);
        opt_in((const void*)gemm_fwd2_kernel);
        opt_in((const void*)gemm_tc_kernel<false, false, false>);
        opt_in((const void*)gemm_tc_kernel<false, true, false>);
        opt_in((const void*)gemm_tc_kernel<true, true, false>);
        opt_in((const void*)gemm_tc_kernel<false, false, true>);
        opt_in((const void*)gemm_tc_kernel<false, true, true>);
        opt_in((const void*)gemm_tc_kernel<true, true, true>);
        const int ws_big = 224 * 1024;
        auto opt_ws = [&](const void* f) { if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, ws_big); };
        opt_ws((const void*)gemm_ws_kernel<false, false>);
        opt_ws((const void*)gemm_ws_kernel<false, true>);
        opt_ws((const void*)gemm_ws_kernel<true, true>);
        opt_ws((const void*)gemm_ws_kernel<false, true, true>);
        opt_ws((const void*)gemm_ws_kernel<true, true, true>);
        opt_ws((const void*)gemm_ws2_kernel<false, false>);
        opt_ws((const void*)gemm_ws2_kernel<false, true>);
        opt_ws((const void*)gemm_ws2_kernel<true, true>);
        opt_ws((const void*)gemm_chain_kernel<false>);
        opt_ws((const void*)gemm_chain_kernel<true>);
        opt_in((const void*)policy_head_kernel<1>);
        opt_in((const void*)policy_head_kernel<4>);
        opt_in((const void*)policy_grad_kernel<1>);
        opt_in((const void*)policy_grad_kernel<4>);
        if (e != cudaSuccess) { delete t; return set_cuda_error(e, "cudaFuncSetAttribute"); }
    }
    int fe = finalize(*t);
    // A feature of the tensor-core program that this configuration cannot have (tensor maps unavailable / misaligned buffers /
    // shapes) is switched off and the program is rebuilt: strip-fused forward chains -> one launch per layer, sign-bit masks
    // -> fp32 activations as masks, gradient-store stages + streaming Adam -> fused Adam epilogues.
    for (int attempt = 0; attempt < 3 && (fe == OAC_E_CHAIN_UNAVAILABLE || fe == OAC_E_BITS_UNAVAILABLE || fe == OAC_E_SPLIT_UNAVAILABLE); ++attempt) {
        if (fe == OAC_E_CHAIN_UNAVAILABLE) t->allow_chain = false;
        else if (fe == OAC_E_BITS_UNAVAILABLE) t->allow_bits = false;
        else t->allow_split = false;
        for (void* p : t->dev_allocs) cudaFree(p);
        t->dev_allocs.clear(); t->stages.clear(); t->work_cursor = 0;
        Builder b2(*t);
        if (cfg->algo == OAC_ALGO_SAC) b2.build_sac();
        else if (cfg->algo == OAC_ALGO_POAC) b2.build_poac();
        else b2.build_goac();
        if (t->work_cursor + WORK_SLACK > t->lay.work_floats) { oac_trainer_destroy(t); return set_error(OAC_E_INVALID, "internal: work arena"); }
        fe = finalize(*t);
    }
    if (fe) { oac_trainer_destroy(t); return fe; }
    { const char* np_ = getenv("OAC_PDL"); g_use_pdl = (np_ && np_[0] == '1'); }
    const char* ng = getenv("OAC_NO_GRAPH");
    t->use_graph = !(ng && ng[0] == '1');
    {
        int lanes = 0;
        for (const Stage& s : t->stages) lanes = std::max(lanes, s.lane);
        for (int l = 0; l < lanes && l < 2; ++l) {
            cudaError_t e = cudaStreamCreateWithFlags(&t->side[l], cudaStreamNonBlocking);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_fork[l], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->ev_join[l], cudaEventDisableTiming);
            if (e != cudaSuccess) { oac_trainer_destroy(t); return set_cuda_error(e, "side stream"); }
        }
    }
    *out = t;
    return 0;
}

extern "C" int oac_trainer_destroy(OacTrainer* t) {
    if (!t) return 0;
    for (int i = 0; i < 2; ++i) if (t->graph[i]) cudaGraphExecDestroy(t->graph[i]);
    for (int l = 0; l < 2; ++l) {
        if (t->ev_fork[l]) cudaEventDestroy(t->ev_fork[l]);
        if (t->ev_join[l]) cudaEventDestroy(t->ev_join[l]);
        if (t->side[l]) cudaStreamDestroy(t->side[l]);
    }
    for (void* p : t->dev_allocs) cudaFree(p);
    delete t;
    return 0;
}

extern "C" int oac_trainer_launches_per_step(const OacTrainer* t) {
    if (!t) return 0;
    int n = 0;
    for (const Stage& s : t->stages) n += (s.fused2 && !s.fwd2_cluster) ? 2 : 1;      // a fused layer pair without its cluster: two launches
    return n;
}

static int stats_len(const OacConfig& c) {
    return c.algo == OAC_ALGO_SAC ? 32 : (c.algo == OAC_ALGO_POAC ? 11 + 9 * c.n_particles : 29);
}

extern "C" int oac_trainer_stats_count(const OacTrainer* t) { return t ? stats_len(t->cfg) : 0; }

extern "C" int oac_trainer_stats(OacTrainer* t, float* out, int32_t out_ld, void* stream) {
    if (!t || !out) return set_error(OAC_E_INVALID, "oac_trainer_stats: null argument");
    if (out_ld < stats_len(t->cfg)) return set_error(OAC_E_INVALID, "oac_trainer_stats: out_ld too small");
    StatsParams p;
    memset(&p, 0, sizeof(p));
    const OacLayout& L = t->lay;
    p.as = t->as; p.algo = t->cfg.algo; p.B = t->cfg.batch; p.A = t->cfg.act_dim; p.P = t->cfg.n_particles; p.nq = L.nq;
    p.deterministic = t->cfg.deterministic; p.auto_alpha = t->cfg.auto_alpha; p.standard_bound = t->cfg.standard_bound;
    p.off_q_pred = L.off_q_pred; p.off_q_target = L.off_q_target; p.off_q_new = L.off_q_new; p.off_log_pi = L.off_log_pi;
    p.off_mean = L.off_mean; p.off_log_std = L.off_log_std; p.off_scalars = L.off_scalars;
    p.out = out; p.out_ld = out_ld;
    trainer_stats_kernel<<<t->cfg.n_seeds, STATS_THREADS, 0, (cudaStream_t)stream>>>(p);
    OAC_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int oac_trainer_ws_stages(const OacTrainer* t) {
    int n = 0;
    if (t) for (const Stage& s : t->stages) n += s.chain ? s.chain : s.use_ws;
    return n;
}

extern "C" int oac_trainer_step(OacTrainer* t, int32_t use_external_eps, void* stream) {
    if (!t) return set_error(OAC_E_INVALID, "null trainer");
    cudaStream_t st = (cudaStream_t)stream;
    const int gi = use_external_eps ? 1 : 0;
    if (!t->use_graph) return launch_stages(*t, gi, st);
    if (!t->graph[gi]) {
        // capture on a private stream so the caller's stream state is untouched
        cudaStream_t cs;
        OAC_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) { cudaStreamDestroy(cs); return set_cuda_error(e, "cudaStreamBeginCapture"); }
        int rc = launch_stages(*t, gi, cs);
        e = cudaStreamEndCapture(cs, &g);
        cudaStreamDestroy(cs);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return set_cuda_error(e, "cudaStreamEndCapture");
        e = cudaGraphInstantiate(&t->graph[gi], g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return set_cuda_error(e, "cudaGraphInstantiate");
    }
    OAC_CUDA(cudaGraphLaunch(t->graph[gi], st));
    return 0;
}

extern "C" int oac_trainer_profile(OacTrainer* t, int32_t iters, int32_t max_stages, float* ms, int32_t* is_gemm,
                                   double* flops, const char** names, int32_t* n_stages, void* stream) {
    if (!t || !ms || !n_stages || iters < 1) return set_error(OAC_E_INVALID, "oac_trainer_profile: bad argument");
    const int n = (int)t->stages.size();
    if (n > max_stages) return set_error(OAC_E_INVALID, "oac_trainer_profile: max_stages too small");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) OAC_CUDA(cudaEventCreate(&e));
    std::vector<double> acc(n, 0.0);
    std::vector<Stage> all;
    all.swap(t->stages);
    int rc = 0;
    // each stage is launched `iters` times back to back between two events, so the figure is the
    // warm, launch-gap-free kernel duration (the state drifts: call this on a scratch trainer)
    for (int i = 0; i < n && !rc; ++i) {
        t->stages.clear();
        t->stages.push_back(all[i]);
        rc = launch_stages(*t, 0, st);          // warm-up
        cudaEventRecord(ev[0], st);
        for (int it = 0; it < iters && !rc; ++it) rc = launch_stages(*t, 0, st);
        cudaEventRecord(ev[1], st);
        cudaStreamSynchronize(st);
        float f = 0.f;
        cudaEventElapsedTime(&f, ev[0], ev[1]);
        acc[i] = f;
    }
    t->stages.swap(all);
    for (auto& e : ev) cudaEventDestroy(e);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) {
        ms[i] = (float)(acc[i] / iters);
        const Stage& s = t->stages[i];
        if (is_gemm) is_gemm[i] = s.kind == ST_GEMM;
        if (names) names[i] = s.name;
        if (flops) {
            double f = 0.0;
            if (s.kind == ST_GEMM) for (auto& g : s.gemm) f += 2.0 * g.M * (double)g.N * g.K;
            flops[i] = f;
        }
    }
    *n_stages = n;
    return 0;
}

static int g_debug_kernel = -1;
// Which kernel the last oac_gemm_debug call ran: 0 SIMT, 1 per-tile tcgen05, 2 warp-specialised TMA + tcgen05, 3 its CTA-pair variant.
extern "C" int oac_gemm_debug_kernel(void) { return g_debug_kernel; }

// Test / measurement aid: one GEMM  C[M,N] = epi(sum_k A(m,k) B(n,k))  through either stage kernel.
extern "C" int oac_gemm_debug(int32_t gemm_path, int32_t a_trans, int32_t b_trans, int32_t M, int32_t N, int32_t K,
                              const float* A, int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc,
                              const float* bias, int32_t relu_, void* stream) {
    if (!A || !B || !C || M < 1 || N < 1 || K < 1) return set_error(OAC_E_INVALID, "oac_gemm_debug: bad argument");
    if (a_trans && !b_trans) return set_error(OAC_E_UNSUPPORTED, "oac_gemm_debug: (a_trans, !b_trans) is unused");
    OacTrainer t;
    memset(&t.cfg, 0, sizeof(t.cfg));
    t.cfg.n_seeds = 1; t.cfg.gemm_path = gemm_path;
    memset(&t.as, 0, sizeof(t.as));          // null arena bases: offsets are absolute addresses / 4
    memset(&t.hyper, 0, sizeof(t.hyper));
    auto ref = [](const float* p) { return Ref{AR_WORK, (long long)(reinterpret_cast<uintptr_t>(p) / 4)}; };
    t.stages.emplace_back();
    Stage& s = t.stages.back();
    s.kind = ST_GEMM; s.name = "debug";
    GemmTask g; memset(&g, 0, sizeof(g));
    g.A = ref(A); g.B = ref(B); g.C = ref(C); g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.a_trans = a_trans; g.b_trans = b_trans; g.M = M; g.N = N; g.K = K;
    g.epi = bias ? (relu_ ? EPI_BIAS_RELU : EPI_BIAS) : EPI_STORE;
    if (bias) g.bias = ref(bias);
    g.target_off = g.target_bias_off = -1;
    s.gemm.push_back(g);
    {
        const int big = 212 * 1024;
        cudaFuncSetAttribute((const void*)gemm_sk_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute((const void*)gemm_sk_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute((const void*)gemm_sk_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute((const void*)gemm_sk_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute((const void*)gemm_sk_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute((const void*)gemm_sk_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute((const void*)gemm_tc_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute((const void*)gemm_tc_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute((const void*)gemm_tc_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute((const void*)gemm_tc_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute((const void*)gemm_tc_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute((const void*)gemm_tc_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
        cudaFuncSetAttribute((const void*)gemm_ws_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        cudaFuncSetAttribute((const void*)gemm_ws_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        cudaFuncSetAttribute((const void*)gemm_ws_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        cudaFuncSetAttribute((const void*)gemm_ws_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        cudaFuncSetAttribute((const void*)gemm_ws_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        cudaFuncSetAttribute((const void*)gemm_ws2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        cudaFuncSetAttribute((const void*)gemm_ws2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        cudaFuncSetAttribute((const void*)gemm_ws2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    }
    { const char* nw = getenv("OAC_NO_WS"); t.allow_ws = !(nw && nw[0] == '1'); }
    int rc = finalize(t);
    const char* dbg_env = getenv("OAC_TC_DEBUG");
    const int dbg_n = 64;
    if (dbg_env && dbg_env[0] == '1') { cudaMalloc(&t.tc_dbg, sizeof(long long) * 8 * dbg_n * 8); cudaMemset(t.tc_dbg, 0, sizeof(long long) * 8 * dbg_n * 8); }
    g_debug_kernel = rc ? -1 : (s.use_ws ? (s.ws_pair ? 3 : 2) : (s.use_tc ? 1 : 0));
    if (!rc) rc = launch_stages(t, 0, (cudaStream_t)stream);
    if (!rc && t.tc_dbg) rc = launch_stages(t, 0, (cudaStream_t)stream);      // second (warm) run is the one reported
    if (const char* reps_env = getenv("OAC_GEMM_DEBUG_REPS")) {                // measurement aid: mean time of this one stage
        const int reps = atoi(reps_env);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, (cudaStream_t)stream);
        for (int i = 0; i < reps && !rc; ++i) rc = launch_stages(t, 0, (cudaStream_t)stream);
        cudaEventRecord(e1, (cudaStream_t)stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        fprintf(stderr, "[gemm debug] path %d layout (%d,%d) M %d N %d K %d: %.2f us per launch, grid %d, bn %d, slots %d, %.1f TFLOP/s\n", gemm_path,
                a_trans, b_trans, M, N, K, 1e3 * ms / reps, s.ws_grid, s.gemm[0].bn, s.ws_slots, 2.0 * M * N * K / (1e9 * ms / reps));
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
    cudaStreamSynchronize((cudaStream_t)stream);
    if (t.tc_dbg) {
        std::vector<long long> h(8 * dbg_n);
        cudaMemcpy(h.data(), t.tc_dbg, sizeof(long long) * 8 * dbg_n, cudaMemcpyDeviceToHost);
        for (int i = 0; i < 3; ++i) {
            fprintf(stderr, "[tc dbg] cta %d:", i);
            for (int j = 1; j < 8; ++j) fprintf(stderr, " %lld", h[8 * i + j] - h[8 * i]);
            fprintf(stderr, "  (alloc | chunk0 landed | pass done | mma issued | mma done | epilogue | dealloc)\n");
        }
        cudaFree(t.tc_dbg);
    }
    for (void* p : t.dev_allocs) cudaFree(p);
    return rc;
}
