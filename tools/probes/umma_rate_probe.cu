// Probe: issue rate of tcgen05.mma kind::tf32 as a function of the operand layouts (K-major SWIZZLE_128B vs MN-major
// SWIZZLE_128B_ATOM_32B in shared memory, A in tensor memory) and of N.  One CTA, operands are whatever is in shared
// memory (timing only).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../oac_explore_b200/csrc -I ../../include
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "gemm_chain.cuh"
using namespace oac;

__global__ void __launch_bounds__(128, 1) probe(long long* out, int n_mma) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar, bar2[8], bar3;
    __shared__ uint32_t s_tmem;
    uint8_t* base = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 48 * 1024; i += 128) reinterpret_cast<float*>(base)[i] = 1.0f;
    fence_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&bar2[i], 1); mbar_init(&bar3, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        uint32_t ph = 0;
        int slot = 0;
        for (int cfg = 0; cfg < 16; ++cfg) {
            const int N = (cfg & 1) ? 256 : 64;
            const int mode = cfg >= 10 ? 0 : cfg >> 1;
            const int cmode = cfg >= 10 ? (cfg - 10) / 2 + 1 : 0;   // 1: commit per 4 MMAs, 2: commit per 8 MMAs, 3: commit per 4 + probe of a completed barrier                 // 0 (K,K)  1 (K,MN)  2 (MN,MN)  3 (TMEM,K)  4 (TMEM,MN)
            const bool a_mn = mode == 2, b_mn = mode == 1 || mode == 2 || mode == 4, a_t = mode >= 3;
            const uint32_t idesc = umma_idesc_tf32(128, N, a_mn, b_mn);
            const uint32_t sa = smem_u32(base), sb = sa + 16384;
            const long long t0 = clock64();
            for (int i = 0; i < n_mma; ++i) {
                uint64_t ad = a_mn ? umma_desc(sa, 4096, 512, 1) : umma_desc(sa, 16, 1024, 2);
                uint64_t bd = b_mn ? umma_desc(sb, 4096, 512, 1) : umma_desc(sb, 16, 1024, 2);
                for (int ks = 0; ks < 4; ++ks) {
                    if (a_t) umma_tf32_ta(tmem + 256, tmem + (uint32_t)(ks * 8), bd, idesc, 1u);
                    else umma_tf32(tmem + 256, ad, bd, idesc, 1u);
                    ad += a_mn ? 64u : 2u; bd += b_mn ? 64u : 2u;
                }
                if (cmode == 1 || cmode == 3 || (cmode == 2 && (i & 1))) umma_commit(&bar2[i & 7]);
                if (cmode == 3) { (void)mbar_try_wait(&bar3, 1u); tc_fence_after(); }
            }
            umma_commit(&bar);
            mbar_wait(&bar, ph); ph ^= 1u;
            const long long t1 = clock64();
            out[slot++] = t1 - t0;
        }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(512u) : "memory"); }
}

int main() {
    long long* d; cudaMalloc(&d, 64 * sizeof(long long));
    cudaFuncSetAttribute((const void*)probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int n = 64;
    for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, 200 * 1024>>>(d, n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long h[16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    const char* cn[3] = {"(K,K) + commit every 4 MMAs", "(K,K) + commit every 8 MMAs", "(K,K) + commit per 4 + barrier probe"};
    const char* names[5] = {"A K-major smem, B K-major", "A K-major smem, B MN-major", "A MN-major smem, B MN-major", "A tmem, B K-major", "A tmem, B MN-major"};
    for (int c = 0; c < 10; ++c)
        printf("%-30s N %3d: %.1f cycles per 128xNx8 MMA (nominal %d)\n", names[c >> 1], (c & 1) ? 256 : 64, (double)h[c] / (n * 4), (c & 1) ? 128 : 32);
    for (int c = 10; c < 16; ++c)
        printf("%-38s N %3d: %.1f cycles per MMA\n", cn[(c - 10) / 2], (c & 1) ? 256 : 64, (double)h[c] / (n * 4));
    return 0;
}
