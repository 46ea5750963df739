// Streaming Adam (+Polyak) over whole parameter blocks: the optimizer half of the many-seed tensor-core path.
//
// With one or a few seeds the Adam update is fused into the epilogue of the weight-gradient GEMM (every weight
// and moment is touched once, no gradient buffer exists).  With many seeds that epilogue is the slowest part of
// the step: its 256 threads per SM cannot keep more than ~64 KB of loads in flight, which caps the stage at
// ~3.6 TB/s however the loads are arranged (ncu, mid-round capture).  Here the weight-gradient GEMMs store plain gradients
// (EPI_GRAD, 4 B per element, laid out like the parameter arena) and this kernel streams
//     grad, param, exp_avg, exp_avg_sq [, Polyak target]  ->  param, exp_avg, exp_avg_sq [, target]
// with 1536 threads per SM, one float4 of every stream per thread.  The arithmetic is adam_update()'s
// (torch.optim.Adam 1.4 operation order, IEEE divide / square root), so the optimizer step itself is exact.
#pragma once
#include "gemm_simt.cuh"

namespace oac {

constexpr int ADAM_MAX_SEGS = 24;
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_UNROLL = 1;

struct AdamSeg {
    long long off;          // first float of the block in the parameter / moment arenas (multiple of 4)
    long long len;          // floats (multiple of 4): a whole net, fc0.weight .. last bias, pads included
    long long grad_off;     // where the gradient of `off` sits in the work arena
    long long target_off;   // Polyak target block in the parameter arena, or -1
    float lr;
    int counter;            // CNT_OPT0 + optimizer index
};

struct AdamStreamParams {
    AdamSeg seg[ADAM_MAX_SEGS];
    int n_seg;
    long long total4;       // float4 elements per seed over all segments
    ArenaSet as;
    AdamHyper hyper;
};

__device__ __forceinline__ void adam_elem(float g, float& p, float& m, float& v, float& t, bool polyak, const AdamScalars& s) {
    m = __fadd_rn(__fmul_rn(m, s.beta1), __fmul_rn(s.one_m_beta1, g));
    v = __fadd_rn(__fmul_rn(v, s.beta2), __fmul_rn(__fmul_rn(s.one_m_beta2, g), g));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), s.bc2_sqrt), s.eps);
    p = __fadd_rn(p, __fmul_rn(-s.step_size, __fdiv_rn(m, denom)));
    if (polyak) t = __fadd_rn(__fmul_rn(t, s.one_m_tau), __fmul_rn(p, s.tau));
}

// UN float4 elements per thread.  Measured at 64 seeds (767 MB per critic launch): UN = 4 (121 registers, 2 CTAs / SM)
// 4.4 TB/s, UN = 2 (4 CTAs / SM) 5.6 TB/s, UN = 1 (6 CTAs / SM) 6.0 TB/s = 91 % of the copy peak -- occupancy, not
// per-thread unrolling, hides the latency here; cache-streaming hints (STREAM) made no difference.
template <int UN, bool STREAM>
__global__ void __launch_bounds__(ADAM_THREADS, UN == 1 ? 6 : (UN == 2 ? 4 : 2)) adam_stream_kernel(const AdamStreamParams* __restrict__ pp) {
    pdl_wait();
    const AdamStreamParams& P = *pp;
    __shared__ AdamScalars s_sc[ADAM_MAX_SEGS];
    __shared__ long long s_start4[ADAM_MAX_SEGS + 1];
    const int seed = blockIdx.y;
    if (threadIdx.x < P.n_seg) {
        const AdamSeg& S = P.seg[threadIdx.x];
        const int32_t* cnt = P.as.counters + seed * P.as.n_counters;
        s_sc[threadIdx.x] = make_adam_scalars(P.hyper, S.lr, cnt[S.counter], cnt[CNT_TRAIN_STEPS]);
    }
    if (threadIdx.x == 0) {
        long long a = 0;
        for (int i = 0; i < P.n_seg; ++i) { s_start4[i] = a; a += P.seg[i].len >> 2; }
        s_start4[P.n_seg] = a;
    }
    __syncthreads();
    float* par = P.as.base[AR_PARAM] + (long long)seed * P.as.stride[AR_PARAM];
    float* m1 = P.as.base[AR_ADAM_M] + (long long)seed * P.as.stride[AR_ADAM_M];
    float* m2 = P.as.base[AR_ADAM_V] + (long long)seed * P.as.stride[AR_ADAM_V];
    const float* wrk = P.as.base[AR_WORK] + (long long)seed * P.as.stride[AR_WORK];
    auto ld4 = [](const float* p_) { return STREAM ? __ldcs(reinterpret_cast<const float4*>(p_)) : *reinterpret_cast<const float4*>(p_); };
    auto st4 = [](float* p_, const float4& v_) { if (STREAM) __stcs(reinterpret_cast<float4*>(p_), v_); else *reinterpret_cast<float4*>(p_) = v_; };
    float4 g4[UN], p4[UN], a4[UN], v4[UN], t4[UN];
    long long po[UN], to[UN];
    int sg[UN];
    const long long base = (long long)blockIdx.x * (ADAM_THREADS * UN) + threadIdx.x;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
        const long long i4 = base + (long long)u * ADAM_THREADS;
        po[u] = -1; to[u] = -1; sg[u] = 0;
        if (i4 < P.total4) {
            int k = 0;
            while (i4 >= s_start4[k + 1]) ++k;
            const AdamSeg& S = P.seg[k];
            const long long e = (i4 - s_start4[k]) << 2;
            sg[u] = k; po[u] = S.off + e;
            g4[u] = ld4(wrk + S.grad_off + e);
            p4[u] = ld4(par + po[u]);
            a4[u] = ld4(m1 + po[u]);
            v4[u] = ld4(m2 + po[u]);
            if (S.target_off >= 0 && s_sc[k].do_polyak) {
                to[u] = S.target_off + e;
                t4[u] = ld4(par + to[u]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
        if (po[u] < 0) continue;
        const AdamScalars s = s_sc[sg[u]];
        const bool pk = to[u] >= 0;
        adam_elem(g4[u].x, p4[u].x, a4[u].x, v4[u].x, t4[u].x, pk, s);
        adam_elem(g4[u].y, p4[u].y, a4[u].y, v4[u].y, t4[u].y, pk, s);
        adam_elem(g4[u].z, p4[u].z, a4[u].z, v4[u].z, t4[u].z, pk, s);
        adam_elem(g4[u].w, p4[u].w, a4[u].w, v4[u].w, t4[u].w, pk, s);
        st4(par + po[u], p4[u]);
        st4(m1 + po[u], a4[u]);
        st4(m2 + po[u], v4[u]);
        if (pk) st4(par + to[u], t4[u]);
    }
}

}  // namespace oac
