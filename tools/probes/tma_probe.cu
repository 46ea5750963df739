// Hardware probe (measurement aid, not product code): what does TMA write to shared memory for
//   (a) FLOAT32 vs TFLOAT32 element types (does TFLOAT32 round, and how?),
//   (b) SWIZZLE_128B and SWIZZLE_128B_ATOM_32B,
// with the tensor map held in GLOBAL memory.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cmath>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const CUtensorMap* tm, int c0, int c1, int c2, int bytes, float* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    uint8_t* dst = sm + ((1024u - ((uint32_t)__cvta_generic_to_shared(sm) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) reinterpret_cast<float*>(dst)[i] = -777.f;
    uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
                     ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(b) : "memory");
    }
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(b) : "memory");
    if (!done) { if (threadIdx.x == 0) out[0] = -12345.f; return; }
    for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(dst)[i];
}

int main() {
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    const int S = 2, R = 64, C = 100;            // [seed][row][col], ld = 100 floats (400 B, multiple of 16)
    std::vector<float> h((size_t)S * R * C);
    for (int s = 0; s < S; ++s) for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c)
        h[((size_t)s * R + r) * C + c] = (float)(s * 10000 + r * 100 + c) + 0.123456789f * (float)((r * 7 + c * 13) % 11);
    float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    float* out; cudaMalloc(&out, 64 * 1024);
    CUtensorMap* dtm; cudaMalloc(&dtm, sizeof(CUtensorMap));
    struct Case { const char* name; CUtensorMapDataType dt; CUtensorMapSwizzle sw; int box_rows; };
    Case cases[] = {{"f32 sw128", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B, 16},
                    {"tf32 sw128", CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, CU_TENSOR_MAP_SWIZZLE_128B, 16},
                    {"f32 sw128_atom32", CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 16}};
    for (auto& cs : cases) {
        CUtensorMap tm;
        cuuint64_t dims[3] = {(cuuint64_t)C - 3, R, S};             // inner extent 97: columns 97.. are out of bounds
        cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)R * C * 4};
        cuuint32_t box[3] = {32, (cuuint32_t)cs.box_rows, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tm, cs.dt, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, cs.sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("== %s: encode rc=%d\n", cs.name, (int)r);
        if (r) continue;
        cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
        const int bytes = 32 * cs.box_rows * 4;
        const int c0 = 64, c1 = 8, c2 = 1;       // covers columns 64..95 ; second probe at 96 tests OOB fill
        probe<<<1, 128, bytes + 2048>>>(dtm, c0, c1, c2, bytes, out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("   kernel: %s\n", cudaGetErrorString(e));
        std::vector<float> o(bytes / 4);
        cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
        // where did element (row r, col k) land?
        int bad_k = 0, bad_mn = 0, exact = 0, zeroed_low = 0, rn = 0, rz = 0;
        for (int rr = 0; rr < cs.box_rows; ++rr) for (int k = 0; k < 32; ++k) {
            float want = h[((size_t)c2 * R + (c1 + rr)) * C + c0 + k];
            // K-major SW128 prediction: 16-byte unit (k/4) ^ (rr & 7)
            int off_k = rr * 32 + (((k >> 2) ^ (rr & 7)) << 2) + (k & 3);
            // ATOM_32B prediction: 32-byte unit (k/8) ^ (rr & 3)
            int off_mn = rr * 32 + (((k >> 3) ^ (rr & 3)) << 3) + (k & 7);
            float gk = o[off_k], gm = o[off_mn];
            uint32_t wb, gb; memcpy(&wb, &want, 4);
            float g = (cs.sw == CU_TENSOR_MAP_SWIZZLE_128B) ? gk : gm;
            memcpy(&gb, &g, 4);
            if (cs.sw == CU_TENSOR_MAP_SWIZZLE_128B) { if (fabsf(gk - want) > 1e-2f * fabsf(want) + 1e-3f) ++bad_k; }
            else { if (fabsf(gm - want) > 1e-2f * fabsf(want) + 1e-3f) ++bad_mn; }
            if (gb == wb) ++exact;
            if ((gb & 0x1FFFu) == 0) ++zeroed_low;
            if (gb == ((wb + 0x1000u) & 0xFFFFE000u)) ++rn;
            if (gb == (wb & 0xFFFFE000u)) ++rz;
        }
        printf("   layout mismatches: kmajor-pred %d, atom32-pred %d (of %d)\n", bad_k, bad_mn, cs.box_rows * 32);
        printf("   bits: exact %d, low13 zero %d, == round-half-up %d, == truncate %d\n", exact, zeroed_low, rn, rz);
        // OOB probe
        probe<<<1, 128, bytes + 2048>>>(dtm, 96, 60, 1, bytes, out);
        cudaDeviceSynchronize();
        cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
        int nz = 0, neg = 0;
        for (float v : o) { if (v == 0.f) ++nz; if (v == -777.f) ++neg; }
        printf("   OOB box (col 96.., rows 60..): zeros %d, untouched %d of %d (in-bounds: 1 col x 4 rows)\n", nz, neg, (int)o.size());
    }
    return 0;
}
