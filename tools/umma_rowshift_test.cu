// Micro-test: can a tcgen05.mma A operand (K-major, SWIZZLE_128B, written by TMA) start at an arbitrary ROW of a
// 1024B-aligned smem slab — i.e. descriptor start address = slab + r*128 with r not a multiple of 8 — and if so, does
// the descriptor's "matrix base offset" field (bits 49-51) have to carry (start >> 7) & 7?  This decides whether the
// 9 filter taps of a 3x3 conv can be served from ONE smem copy of the activations (shifted descriptors).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/umma_rowshift_test tools/umma_rowshift_test.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define AROWS 192
#define NB 16
__device__ __forceinline__ uint32_t su32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool wait_bar(uint64_t *bar, uint32_t parity)
{
    for (int i = 0; i < 4000000; i++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(su32(bar)), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float *out,
                                            int shift, int use_base_offset, int *err)
{
    extern __shared__ uint8_t raw[];
    uint8_t *sm = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = sm, *sB = sm + AROWS * 128;              // 24 KB slab, then B (16 rows x 128 B), both 1024-aligned
    __shared__ uint64_t full, done;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&done)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(su32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tmem_base;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&full)), "r"(AROWS * 128 + NB * 128) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(su32(sA)), "l"((uint64_t)&mapA), "r"(su32(&full)), "r"(0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(su32(sB)), "l"((uint64_t)&mapB), "r"(su32(&full)), "r"(0), "r"(0) : "memory");
        if (!wait_bar(&full, 0)) atomicOr(err, 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((128u >> 4) << 24);
        for (int kk = 0; kk < 4; kk++) {
            uint32_t a_addr = su32(sA) + shift * 128 + kk * 32, b_addr = su32(sB) + kk * 32;
            uint64_t bo = use_base_offset ? (uint64_t)((a_addr >> 7) & 7) : 0ull;
            uint64_t adesc = (uint64_t)((a_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (bo << 49) | (2ull << 61);
            uint64_t bdesc = (uint64_t)((b_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
            uint32_t acc = kk ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tb), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(&done)) : "memory");
    }
    if (!wait_bar(&done, 0)) atomicOr(err, 2);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tb + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; j++) out[(warp * 32 + lane) * NB + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tb) : "memory");
}
int main()
{
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    PFN enc = (PFN)fp;
    std::vector<__nv_bfloat16> hA(AROWS * 64), hB(NB * 64);
    std::vector<float> fA(AROWS * 64), fB(NB * 64);
    for (int i = 0; i < AROWS; i++) for (int j = 0; j < 64; j++) { float x = (float)(((i * 7 + j * 3) % 13) - 6); fA[i * 64 + j] = x; hA[i * 64 + j] = __float2bfloat16(x); }
    for (int i = 0; i < NB; i++) for (int j = 0; j < 64; j++) { float x = (float)(((i * 5 + j * 11) % 7) - 3); fB[i * 64 + j] = x; hB[i * 64 + j] = __float2bfloat16(x); }
    __nv_bfloat16 *dA, *dB; float *dO; int *dE;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * NB * 4); cudaMalloc(&dE, 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap mA, mB; cuuint64_t str[1] = {128}; cuuint32_t es[2] = {1, 1};
    cuuint64_t dimsA[2] = {64, AROWS}; cuuint32_t boxA[2] = {64, AROWS};
    cuuint64_t dimsB[2] = {64, NB}; cuuint32_t boxB[2] = {64, NB};
    int r1 = enc(&mA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dimsA, str, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int r2 = enc(&mB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsB, str, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d %d\n", r1, r2);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, AROWS * 128 + NB * 128 + 2048);
    for (int mode = 0; mode < 2; mode++)
        for (int shift : {0, 8, 1, 2, 3, 5, 7, 9, 18, 19, 20, 37, 38}) {
            cudaMemset(dO, 0, 128 * NB * 4); cudaMemset(dE, 0, 4);
            k<<<1, 128, AROWS * 128 + NB * 128 + 2048>>>(mA, mB, dO, shift, mode, dE);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> o(128 * NB); int herr = 0;
            cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&herr, dE, 4, cudaMemcpyDeviceToHost);
            double maxerr = 0; int bad = 0;
            for (int m = 0; m < 128; m++) for (int n = 0; n < NB; n++) {
                float ref = 0; for (int kx = 0; kx < 64; kx++) ref += fA[(m + shift) * 64 + kx] * fB[n * 64 + kx];
                double d = fabs((double)ref - o[m * NB + n]); if (d > maxerr) maxerr = d; if (d > 1e-3) bad++;
            }
            printf("base_offset_field=%d shift=%2d cuda=%d flags=%d mismatches=%4d max_err=%g -> %s\n", mode, shift, (int)e, herr, bad, maxerr, bad ? "WRONG" : "ok");
            if (e != cudaSuccess) return 1;
        }
    return 0;
}
