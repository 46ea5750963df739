// Small device helpers shared by the glue and per-observation kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oac {

constexpr int GLUE_WARPS = 8;
constexpr int GLUE_THREADS = GLUE_WARPS * 32;
constexpr float LOG_SIG_MAX_F = 2.0f;     // trainer/policies.py:10
constexpr float LOG_SIG_MIN_F = -20.0f;   // trainer/policies.py:11
constexpr float TANH_EPS_F = 1e-6f;       // trainer/policies.py:127

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- Philox4x32-10 -> N(0,1) (device-side noise when the caller injects none) ----
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t stream, uint32_t step,
                                               uint32_t row, uint32_t col) {
    uint4 r = philox4x32(make_uint4(row, col, step, stream),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    float u1 = ((float)r.x + 1.0f) * 2.3283064365386963e-10f;   // (0,1]
    float u2 = (float)r.y * 2.3283064365386963e-10f;
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

}  // namespace oac
