// Micro-test: is TMA's 128B swizzle a function of the absolute smem address (so that a box
// landing at a 128B-aligned but not 1024B-aligned offset continues the pattern), or box-relative?
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t su32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap map, uint16_t *out, int row_off)
{
    extern __shared__ uint8_t raw[];
    uint8_t *sm = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 8192 / 2; i += blockDim.x) ((uint16_t *)sm)[i] = 0xFFFF;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar)), "r"(2 * 8 * 128) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(su32(sm)), "l"((uint64_t)&map), "r"(su32(&bar)), "r"(0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(su32(sm + row_off * 128)), "l"((uint64_t)&map), "r"(su32(&bar)), "r"(0), "r"(8) : "memory");
        uint32_t ok = 0;
        for (int i = 0; i < 1000000 && !ok; i++)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(su32(&bar)) : "memory");
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8192 / 2; i += blockDim.x) out[i] = ((uint16_t *)sm)[i];
}
int main()
{
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    PFN enc = (PFN)fp;
    std::vector<uint16_t> h(64 * 64);
    for (int i = 0; i < 64 * 64; i++) h[i] = (uint16_t)i;            // value = row*64 + col
    uint16_t *d, *o; cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 8192);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap map; cuuint64_t dims[2] = {64, 64}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, 8}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    for (int row_off : {8, 11, 23}) {
        k<<<1, 128, 8192 + 1024>>>(map, o, row_off);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<uint16_t> s(4096); cudaMemcpy(s.data(), o, 8192, cudaMemcpyDeviceToHost);
        int abs_ok = 1, rel_ok = 1;
        for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) for (int el = 0; el < 8; el++) {
            int R = row_off + i; uint16_t want = (uint16_t)((8 + i) * 64 + j * 8 + el);
            if (s[R * 64 + ((j ^ (R & 7)) * 8) + el] != want) abs_ok = 0;
            if (s[R * 64 + ((j ^ (i & 7)) * 8) + el] != want) rel_ok = 0;
        }
        int first_ok = 1;
        for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) if (s[i * 64 + ((j ^ i) * 8)] != (uint16_t)(i * 64 + j * 8)) first_ok = 0;
        printf("row_off=%d err=%d first_box_ok=%d address_based=%d box_relative=%d\n", row_off, (int)e, first_ok, abs_ok, rel_ok);
    }
    return 0;
}
