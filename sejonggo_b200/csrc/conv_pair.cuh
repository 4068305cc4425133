// conv_pair.cuh — k_conv3x3_pair: the tower's 3x3 convolution on CTA PAIRS (cta_group::2).
//
// Why (measured, profiles/r01_conv_variants.json): in the single-CTA kernel the two fp32
// accumulators fill TMEM, so the epilogue (TMEM -> bias/skip/ReLU -> HBM, latency bound on the
// skip reads) cannot overlap the next tile's MMAs: 1,935 TFLOP/s without the epilogue,
// 1,124 with it.  Pairing two SMs halves the accumulator per SM (128 rows x 256 fp32 columns),
// so TMEM holds TWO accumulator stages and the epilogue of tile i runs under the MMAs of tile
// i+1; each SM stages only its own 128 A rows and half of the weights (B) per k-block, which
// also doubles the smem pipeline depth (6 x 32 KB).
//
//   cluster (2,1,1); per CTA 192 threads: warp 0 TMA producer (both CTAs), warp 1 MMA issuer
//   (leader CTA only) + TMEM alloc (both), warps 2-5 epilogue (both).
//   pair tile = RT image rows of W pixels (RT*W <= 256): rank 0 owns tile rows [0,128), rank 1
//   the rest.  128 is not a multiple of W=17, so each rank's A operand is two TMA boxes (full
//   rows + a partial row); TMA's 128B swizzle is a function of the absolute smem address
//   (tools/tma_swizzle_test.cu, verified on B200), so boxes landing at any 128B-aligned row
//   offset continue the K-major SW128 pattern the UMMA descriptor expects.
#pragma once

#define PR_STAGES 6
#define PR_A_BYTES (128 * 128)
#define PR_B_BYTES (128 * 128)
#define PR_STAGE_BYTES (PR_A_BYTES + PR_B_BYTES)

#define PR_MAXH 8
struct PairMaps {
    CUtensorMap full0, part0, part1, full1;   // activation boxes: (64, W, f0), (64, p0, 1), (64, W-p0, 1), (64, W, f1)
    CUtensorMap w;                            // weights box (64 ci, 128 co)
    CUtensorMap fullh[PR_MAXH];               // real-row tiling: (64, W, h) for h = 1..PR_MAXH (a run of rows inside ONE position)
    CUtensorMap pf;                           // L2 prefetch box (C, W, pf_rows), unswizzled: half of the NEXT tile's input rows
};

struct PairArgs {
    int W, RT, rows_per_pos, YB, n_tiles, w_row0, relu;
    // rr != 0: a tile is RT consecutive REAL pixel rows (pad rows are skipped, no MAC is spent on them): tile row i of tile t
    // is real row R = t*RT + i = (pos, y) at padded row pos*rows_per_pos + 1 + y; a tile crosses at most one position
    // boundary (needs RT <= W), so each CTA's full rows are one or two row runs.  n_real = positions * W.
    int rr, n_real;
    int dbg;                  // timing ablations only (SGO_CONV_DEBUG): 1 = no epilogue global traffic, 2 = no A loads, 4 = no B loads,
                              // 8 = tap-major k order (old), 16 = no next-tile L2 prefetch
    int pf_rows;              // rows per prefetch box (0 = off): 2 * pf_rows >= RT + 3
    int n_taps, kchunks;      // 9 x 4 for the tower convs; 1 x 3 for the stem GEMM over the im2col tensor
    int f0, p0, f1;           // rank 0: f0 full rows + p0 pixels of row f0; rank 1: (W-p0) pixels of row f0 (if p0) + f1 full rows
    const float *bias;
    const __nv_bfloat16 *skip;
    __nv_bfloat16 *out;       // nullptr: do not store the activations (last layer feeding only the heads)
    int32_t *err;
    // fused 1x1 head convolutions (model.py:73,83) on the last layer: feat[(pos*W*W + pix)*4 + {p0,p1,v0,v1}]
    const float *head_w4;     // [C][4] folded weights, nullptr = off
    const float *head_b4;     // [4]
    float *feat;
};

struct PairSmemTail {
    uint64_t full[PR_STAGES], empty[PR_STAGES], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
    float bias[TW_C];
    float4 w4[TW_C];
};
#define PR_SMEM_BYTES (PR_STAGES * PR_STAGE_BYTES + (int)sizeof(PairSmemTail) + 1024)

#define PR_PEER_MASK 0xFEFFFFFFu      // cute::Sm100MmaPeerBitMask: clear the CTA-rank bit -> leader CTA's smem

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-SM TMA loads: data lands in THIS CTA's smem, the transaction bytes are credited to the
// LEADER CTA's mbarrier (copy_sm100_tma.hpp SM100_TMA_2SM_LOAD_*)
__device__ __forceinline__ void tma2_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & PR_PEER_MASK), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & PR_PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
// bring a box into L2 only (no smem, no barrier): the next tile's activations, so that its first-touch DRAM reads do not all
// land in the first k-blocks of the tile (every CTA reaches a tile boundary at about the same time)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at this smem offset in BOTH CTAs once the pair's MMAs so far retire
__device__ __forceinline__ void umma2_commit_mc(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the LEADER CTA's barrier (local for rank 0, remote for rank 1)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t *bar)
{
    asm volatile("{\n\t.reg .b32 ra;\n\t"
                 "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
                 "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
                 ::"r"(smem_u32(bar)) : "memory");
}
// M=256 (pair), N=256, bf16 x bf16 -> f32, both operands K-major
#define PR_IDESC ((1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24))

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TW_THREADS, 1)
k_conv3x3_pair(const __grid_constant__ PairMaps maps, PairArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    PairSmemTail *tail = reinterpret_cast<PairSmemTail *>(smem + PR_STAGES * PR_STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.full0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.full1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.w) : "memory");
        if (a.rr)
            for (int h = 0; h < PR_MAXH; h++) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.fullh[h]) : "memory");
        for (int s = 0; s < PR_STAGES; s++) { mbar_init(&tail->full[s], 1); mbar_init(&tail->empty[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&tail->tmem_full[s], 1); mbar_init(&tail->tmem_empty[s], 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tail->tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < TW_C; i += blockDim.x) {
        tail->bias[i] = a.bias[i];
        if (a.head_w4) tail->w4[i] = reinterpret_cast<const float4 *>(a.head_w4)[i];
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                       // peer barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;
    const int rows_total = a.RT * a.W;
    const uint32_t a_bytes0 = 128u * 128u, a_bytes1 = (uint32_t)(rows_total - 128) * 128u;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            bool ok = true;
            for (int tile = pair; tile < a.n_tiles && ok; tile += n_pairs) {
                int yb0 = tile * a.RT;
                int h1 = a.RT;                                    // tile rows inside the first position
                if (a.rr) {
                    int pos0 = yb0 / a.W, y0 = yb0 - pos0 * a.W;
                    h1 = a.W - y0 < a.RT ? a.W - y0 : a.RT;
                    yb0 = pos0 * a.rows_per_pos + 1 + y0;         // padded row of tile row 0; tile row i -> yb0 + i + (i >= h1)
                }
                const int r1s = a.f0 + (a.p0 ? 1 : 0);            // first full tile row of rank 1
                if (a.pf_rows && !(a.dbg & 16) && tile + n_pairs < a.n_tiles) {
                    int nyb = (tile + n_pairs) * a.RT;            // padded row of the next tile's row 0
                    if (a.rr) nyb += nyb / a.W + 1;
                    tma_prefetch_3d(&maps.pf, 0, 0, nyb - 1 + (int)rank * a.pf_rows);
                }
                const int nkb = a.n_taps * a.kchunks;
                // k order: channel chunk outer, filter tap inner — the 9 taps of one chunk re-read (shifted) the same activation
                // slice, so first-touch traffic is spread over the tile instead of filling its first 4 k-blocks
                for (int kb = 0; kb < nkb && ok; kb++) {
                    int tap, kc;
                    if (a.dbg & 8) { tap = kb / a.kchunks; kc = kb - tap * a.kchunks; }
                    else { kc = kb / a.n_taps; tap = kb - kc * a.n_taps; }
                    int dy = a.n_taps == 1 ? 0 : tap / 3 - 1, dx = a.n_taps == 1 ? 0 : tap % 3 - 1;
                    {
                        ok = mbar_wait(&tail->empty[stage], phase ^ 1, a.err);
                        if (!ok) break;
                        uint8_t *sa = smem + (size_t)stage * PR_STAGE_BYTES, *sb = sa + PR_A_BYTES;
                        uint64_t *fb = &tail->full[stage];
                        if (a.dbg & 6) {                                                  // ablation: drop operand loads
                            if (rank == 0) mbar_expect_tx(fb, ((a.dbg & 2) ? 0u : a_bytes0 + a_bytes1) + ((a.dbg & 4) ? 0u : 2u * PR_B_BYTES));
                            if (!(a.dbg & 4)) tma2_load_2d(sb, &maps.w, kc * TW_KCH, a.w_row0 + tap * TW_C + (int)rank * 128, fb);
                            if (!(a.dbg & 2)) {
                                if (rank == 0) { tma2_load_3d(sa, &maps.full0, kc * TW_KCH, dx, yb0 + dy, fb);
                                                 if (a.p0) tma2_load_3d(sa + (size_t)a.f0 * a.W * 128, &maps.part0, kc * TW_KCH, dx, yb0 + a.f0 + dy, fb); }
                                else { if (a.p0) tma2_load_3d(sa, &maps.part1, kc * TW_KCH, dx + a.p0, yb0 + a.f0 + dy, fb);
                                       if (a.f1) tma2_load_3d(sa + (size_t)(a.p0 ? a.W - a.p0 : 0) * 128, &maps.full1, kc * TW_KCH, dx, yb0 + r1s + dy, fb); }
                            }
                            if (++stage == PR_STAGES) { stage = 0; phase ^= 1; }
                            continue;
                        }
                        if (rank == 0) {
                            mbar_expect_tx(fb, a_bytes0 + a_bytes1 + 2 * PR_B_BYTES);     // bytes of BOTH CTAs
                            if (h1 >= a.f0) {
                                tma2_load_3d(sa, &maps.full0, kc * TW_KCH, dx, yb0 + dy, fb);
                            } else {                                                      // the position boundary cuts this rank's rows
                                tma2_load_3d(sa, &maps.fullh[h1 - 1], kc * TW_KCH, dx, yb0 + dy, fb);
                                tma2_load_3d(sa + (size_t)h1 * a.W * 128, &maps.fullh[a.f0 - h1 - 1], kc * TW_KCH, dx, yb0 + h1 + 1 + dy, fb);
                            }
                            if (a.p0) tma2_load_3d(sa + (size_t)a.f0 * a.W * 128, &maps.part0, kc * TW_KCH, dx,
                                                   yb0 + a.f0 + (a.f0 >= h1 ? 1 : 0) + dy, fb);
                        } else {
                            uint8_t *sf = sa + (size_t)(a.p0 ? a.W - a.p0 : 0) * 128;
                            if (a.p0) tma2_load_3d(sa, &maps.part1, kc * TW_KCH, dx + a.p0, yb0 + a.f0 + (a.f0 >= h1 ? 1 : 0) + dy, fb);
                            if (a.f1) {
                                if (h1 <= r1s || h1 >= r1s + a.f1) {
                                    tma2_load_3d(sf, &maps.full1, kc * TW_KCH, dx, yb0 + r1s + (r1s >= h1 ? 1 : 0) + dy, fb);
                                } else {
                                    tma2_load_3d(sf, &maps.fullh[h1 - r1s - 1], kc * TW_KCH, dx, yb0 + r1s + dy, fb);
                                    tma2_load_3d(sf + (size_t)(h1 - r1s) * a.W * 128, &maps.fullh[r1s + a.f1 - h1 - 1], kc * TW_KCH, dx,
                                                 yb0 + h1 + 1 + dy, fb);
                                }
                            }
                        }
                        tma2_load_2d(sb, &maps.w, kc * TW_KCH, a.w_row0 + tap * TW_C + (int)rank * 128, fb);
                        if (++stage == PR_STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0 && lane == 0) {
            uint32_t stage = 0, phase = 0;
            bool ok = true;
            int it = 0;
            for (int tile = pair; tile < a.n_tiles && ok; tile += n_pairs, it++) {
                const int acc = it & 1;
                ok = mbar_wait(&tail->tmem_empty[acc], ((it >> 1) & 1) ^ 1, a.err);    // both epilogues drained this stage
                if (!ok) break;
                tc_fence_after();
                for (int kb = 0; kb < a.n_taps * a.kchunks; kb++) {
                    ok = mbar_wait(&tail->full[stage], phase, a.err);
                    if (!ok) break;
                    tc_fence_after();
                    uint32_t sa = smem_u32(smem + (size_t)stage * PR_STAGE_BYTES), sb = sa + PR_A_BYTES;
#pragma unroll
                    for (int k = 0; k < TW_KCH / 16; k++)
                        umma2_bf16(tmem_base + acc * 256, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), PR_IDESC,
                                   (kb | k) ? 1u : 0u);
                    umma2_commit_mc(&tail->empty[stage]);                               // frees the slot in both CTAs
                    if (++stage == PR_STAGES) { stage = 0; phase ^= 1; }
                }
                if (ok) umma2_commit_mc(&tail->tmem_full[acc]);
            }
        }
    } else {
        const int q = warp & 3;
        const int r = (int)rank * 128 + q * 32 + lane;             // row of the pair tile owned by this thread
        const int ry = r / a.W, x = r - ry * a.W;
        bool ok = true;
        int it = 0;
        for (int tile = pair; tile < a.n_tiles; tile += n_pairs, it++) {
            const int acc = it & 1;
            int yb = tile * a.RT + ry;
            bool valid;
            if (a.rr) {                                              // yb = real row -> padded row
                valid = r < rows_total && yb < a.n_real;
                int pos = yb / a.W;
                yb += pos + 1;
            } else {
                valid = r < rows_total && yb < a.YB && (yb % a.rows_per_pos) != 0;
            }
            const size_t gofs = ((size_t)yb * a.W + x) * TW_C;
            if (a.dbg & 1) valid = false;                            // ablation: no epilogue global traffic
            if (valid && a.skip) {                                   // idle until the MMAs finish: pull the skip row towards L2
#pragma unroll
                for (int j = 0; j < 4; j++) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.skip + gofs + j * 64));
            }
            if (ok) ok = mbar_wait(&tail->tmem_full[acc], (it >> 1) & 1, a.err);
            ok = __all_sync(SGO_FULL, ok);
            if (!ok) break;
            tc_fence_after();
            float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f;
#pragma unroll 1
            for (int c = 0; c < TW_C / 32; c++) {
                uint4 sk[4];
                if (valid && a.skip) {
                    const uint4 *sp = reinterpret_cast<const uint4 *>(a.skip + gofs + c * 32);
#pragma unroll
                    for (int j = 0; j < 4; j++) sk[j] = sp[j];
                }
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256 + c * 32, v);
                if (valid) {
                    uint4 o[4];
                    uint32_t *ow = reinterpret_cast<uint32_t *>(o);
                    const uint32_t *sw = reinterpret_cast<const uint32_t *>(sk);
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        float f0 = __uint_as_float(v[2 * j]) + tail->bias[c * 32 + 2 * j];
                        float f1 = __uint_as_float(v[2 * j + 1]) + tail->bias[c * 32 + 2 * j + 1];
                        if (a.skip) {
                            __nv_bfloat162 s2 = *reinterpret_cast<const __nv_bfloat162 *>(&sw[j]);
                            f0 += __bfloat162float(s2.x);
                            f1 += __bfloat162float(s2.y);
                        }
                        if (a.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
                        __nv_bfloat162 p = __floats2bfloat162_rn(f0, f1);
                        ow[j] = *reinterpret_cast<uint32_t *>(&p);
                        if (a.head_w4) {
                            float4 wa = tail->w4[c * 32 + 2 * j], wb = tail->w4[c * 32 + 2 * j + 1];
                            h0 = fmaf(f0, wa.x, h0); h1 = fmaf(f0, wa.y, h1); h2 = fmaf(f0, wa.z, h2); h3 = fmaf(f0, wa.w, h3);
                            h0 = fmaf(f1, wb.x, h0); h1 = fmaf(f1, wb.y, h1); h2 = fmaf(f1, wb.z, h2); h3 = fmaf(f1, wb.w, h3);
                        }
                    }
                    if (a.out) {
                        uint4 *op = reinterpret_cast<uint4 *>(a.out + gofs + c * 32);
#pragma unroll
                        for (int j = 0; j < 4; j++) op[j] = o[j];
                    }
                }
            }
            tc_fence_before();
            mbar_arrive_leader(&tail->tmem_empty[acc]);
            if (valid && a.head_w4) {
                int pos = (yb - 1) / a.rows_per_pos, y = (yb - 1) - pos * a.rows_per_pos;
                float4 r4;
                r4.x = fmaxf(h0 + a.head_b4[0], 0.f); r4.y = fmaxf(h1 + a.head_b4[1], 0.f);
                r4.z = fmaxf(h2 + a.head_b4[2], 0.f); r4.w = fmaxf(h3 + a.head_b4[3], 0.f);
                reinterpret_cast<float4 *>(a.feat)[((size_t)pos * a.W + y) * a.W + x] = r4;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                       // neither CTA frees TMEM / exits while its peer still uses the pair
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}
