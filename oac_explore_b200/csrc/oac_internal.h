// Internal structures shared by the host-side program builder and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/oac_b200.h"

namespace oac {

// ---- arenas: every device pointer a kernel touches is (arena base + seed*stride + offset) ----
enum Arena { AR_PARAM = 0, AR_ADAM_M = 1, AR_ADAM_V = 2, AR_WORK = 3, AR_IO = 4, AR_COUNT = 5 };

struct ArenaSet {
    float* base[AR_COUNT];
    long long stride[AR_COUNT];   // floats per seed
    int32_t* counters;
    int n_counters;
};

struct Ref {          // a location inside an arena (per seed)
    int arena;
    long long off;    // floats
};

__host__ __device__ inline float* resolve(const ArenaSet& as, Ref r, int seed) {
    return as.base[r.arena] + (long long)seed * as.stride[r.arena] + r.off;
}

// ---- counters (int32, per seed) ----
enum Counter {
    CNT_TRAIN_STEPS = 0,   // _n_train_steps_total
    CNT_OPT0 = 1,          // Adam step counts, one per optimizer (see builder)
    CNT_MAX_OPT = 16,
    CNT_TICKET0 = 20,      // last-CTA-done tickets for the glue kernels (reset by the owner)
    CNT_TICKET1 = 21,
    CNT_RNG_LO = 22,       // per-seed 64-bit key (the REAL seed id of a batched seed, 0 for a single trainer) mixed into the
    CNT_RNG_HI = 23,       // Philox key of the step's rsample noise: streams follow the seed, not its slot in the group
    CNT_TOTAL = 24
};

// ---- scalars slot (io[off_scalars + i]) ----
enum Scalar { SC_ALPHA = 0, SC_ALPHA_LOSS = 1, SC_MEAN_LOGPI = 2, SC_STEP_STAMP = 3, SC_COUNT = 16 };

// ---- GEMM stage ----
enum Epilogue {
    EPI_STORE = 0,
    EPI_BIAS = 1,        // C = acc + bias[n]
    EPI_BIAS_RELU = 2,   // C = max(acc + bias[n], 0)
    EPI_MASK = 3,        // C = acc * (mask[m, n] > 0)
    EPI_ADAM = 4,        // acc is dW for the parameter block at C: Adam step (+ Polyak of `target`)
    EPI_GRAD = 5         // acc is dW, stored to the gradient block at C (bias gradient to `bias`); Adam streams later
};

struct GemmTask {
    Ref A, B, C;
    int lda, ldb, ldc;
    int a_trans;       // 0: A(m,k) = A[m*lda + k]   1: A(m,k) = A[k*lda + m]
    int b_trans;       // 0: B(n,k) = B[n*ldb + k]   1: B(n,k) = B[k*ldb + n]
    int M, N, K;
    int epi;
    Ref bias;          // EPI_BIAS*: bias[n].  EPI_ADAM: bias parameter block [M] (db = sum_k A(m,k))
    Ref mask;          // EPI_MASK
    int ldmask;
    int mask_bits;     // EPI_MASK: `mask` holds sign bits (one byte per 4 columns, gemm_ws.cuh), ldmask = bytes per row
    Ref bits;          // EPI_BIAS_RELU on the TMA + tcgen05 path: also store the sign bits of C here (ldbits > 0)
    int ldbits;        // bytes per row (N / 4), 0 = none
    int no_store;      // strip-fused chains: the output only feeds the next layer (tensor memory), nothing is written
    // EPI_ADAM
    long long adam_off;       // offset of C inside the Adam arenas (same as C.off: trainable prefix)
    long long adam_bias_off;
    long long target_off;     // Polyak target of C inside AR_PARAM, or -1
    long long target_bias_off;
    float lr;
    int counter;              // index of the optimizer's step counter
    int has_bias;             // update the bias block too
    int train_bias;           // 0: bias frozen
    int tiles_m, tiles_n;     // filled by the builder for the chosen tile shape
    int bn;                   // warp-specialised tcgen05 path: this task's tile width
    int tile0;                // ... and its first tile in the per-seed work list 
};

struct AdamHyper {
    float beta1, beta2, eps;
    float tau;              // soft_target_tau
    float one_minus_tau;    // (float)(1.0 - tau) evaluated in double like the reference's Python
    int target_period;
};

}  // namespace oac
