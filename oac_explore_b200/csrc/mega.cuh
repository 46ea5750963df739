// Single-launch step for the latency regime (one or a few seeds, fp32 FFMA path).
//
// A single-seed update is a dependent chain of ~13 small stages, none of which can use more than a fraction of the
// chip; as separate kernels (even replayed from a CUDA graph) each stage pays ~4 us of launch, ramp and drain on top
// of its work (profiles/r01: 6.2 us for a K = 1 stage).  Here the WHOLE step is one cooperative kernel: a persistent grid
// of 2 CTAs per SM walks the phases of the step, every CTA taking virtual blocks of the phase's stages
// (gemm_sk_body / policy_head_body / critic_head_body / policy_grad_body -- the same device code the stand-alone
// kernels run), and phases are separated by a grid-wide barrier (release arrive + acquire spin, ~1.5 us) instead of a
// kernel boundary.  Stages the two-lane schedule marks as independent (lane 1) share a phase with the critical-chain
// stage they overlap, so the idle SMs of that phase do their blocks.
//
// Memory model: a stage reads what earlier phases wrote.  All such reads are cp.async or plain ld.global (ld_g), never
// ld.global.nc, and the barrier's acquire load invalidates the SM's L1 before the next phase starts.
#pragma once
#include "glue.cuh"

namespace oac {

constexpr int MEGA_MAX_PHASES = 40;
constexpr int MEGA_MAX_PER_PHASE = 3;
constexpr int MEGA_THREADS = 256;

enum MegaKind { MK_GEMM_NN = 0, MK_GEMM_NT = 1, MK_GEMM_TT = 2, MK_POLICY_HEAD = 3, MK_CRITIC_HEAD = 4, MK_POLICY_GRAD = 5 };

struct MegaStage {
    int kind;
    int gx, gy, gz;          // virtual grid of the stage (what the stand-alone launch would use)
    int nvb;                 // gx * gy * gz
    const void* params;      // device copy of StageParams / PolicyHeadParams / CriticHeadParams / PolicyGradParams
};
struct MegaPhase {
    MegaStage st[MEGA_MAX_PER_PHASE];
    int n;
    int total_vb;
};
struct MegaProgram {
    MegaPhase ph[MEGA_MAX_PHASES];
    int n_phases;
    unsigned* barrier;       // one counter, 0 between launches (the last CTA to leave resets it)
    unsigned long long* dbg; // optional [2 * MEGA_MAX_PHASES + 1] globaltimer stamps of CTA 0 (OAC_MEGA_DEBUG=1)
};

__device__ __forceinline__ void mega_grid_barrier(unsigned* ctr, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        // poll with relaxed loads (an acquire load invalidates the SM's L1 every time, which also hits the other CTA of
        // this SM while it is still working) and acquire once at the end
        unsigned v = 0;
        for (unsigned it = 0; ; ++it) {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= target) break;
            if (it > (1u << 23)) __trap();           // a lost CTA must not hang the GPU
            __nanosleep(40);
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

__global__ void __launch_bounds__(MEGA_THREADS, 2) step_mega_kernel(const MegaProgram* __restrict__ prog, int use_external_eps) {
    unsigned target = 0;
    const int n_phases = prog->n_phases;
    unsigned long long* dbg = (blockIdx.x == 0 && threadIdx.x == 0) ? prog->dbg : nullptr;
    auto stamp = [&](int i) {
        if (dbg) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); dbg[i] = t_; }
    };
    stamp(0);
    for (int ph = 0; ph < n_phases; ++ph) {
        const MegaPhase& P = prog->ph[ph];
        for (int vb = blockIdx.x; vb < P.total_vb; vb += gridDim.x) {
            int si = 0, v = vb;
            while (v >= P.st[si].nvb) { v -= P.st[si].nvb; ++si; }
            const MegaStage& S = P.st[si];
            const int bx = v % S.gx, r = v / S.gx, by = r % S.gy, bz = r / S.gy;
            switch (S.kind) {
                case MK_GEMM_NN: gemm_sk_body<false, false>(*static_cast<const StageParams*>(S.params), bx, by, bz); break;
                case MK_GEMM_NT: gemm_sk_body<false, true>(*static_cast<const StageParams*>(S.params), bx, by, bz); break;
                case MK_GEMM_TT: gemm_sk_body<true, true>(*static_cast<const StageParams*>(S.params), bx, by, bz); break;
                case MK_POLICY_HEAD:
                    policy_head_body<4>(*static_cast<const PolicyHeadParams*>(S.params), use_external_eps, bx, by, bz, S.gx, S.gy);
                    break;
                case MK_CRITIC_HEAD: critic_head_body<4>(*static_cast<const CriticHeadParams*>(S.params), bx, by); break;
                default: policy_grad_body<4>(*static_cast<const PolicyGradParams*>(S.params), bx, by, bz); break;
            }
            __syncthreads();                         // shared memory is reused by the next virtual block
        }
        stamp(2 * ph + 1);                           // this CTA's blocks of the phase are done
        if (ph + 1 < n_phases) mega_grid_barrier(prog->barrier, target);
        stamp(2 * ph + 2);                           // ... and so are everybody's
    }
    // leave: the last CTA of the launch puts the counter back to 0 for the next step
    if (threadIdx.x == 0) {
        const unsigned total = (unsigned)n_phases * gridDim.x;       // (n_phases - 1) barriers + this arrival
        const unsigned old = atomicAdd(prog->barrier, 1u);
        if (old + 1u == total) atomicExch(prog->barrier, 0u);
    }
}

}  // namespace oac
