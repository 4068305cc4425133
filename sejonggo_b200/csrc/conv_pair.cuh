// conv_pair.cuh — k_conv3x3_pair: the tower's 3x3 convolution as a SHIFTED implicit GEMM on CTA pairs.
//
// What bounds this kernel (measured, profiles/r01_conv_ablations_rowtile.json): not the tensor pipe and not the
// epilogue but the bytes the SMs pull out of L2.  A 256x256x64 k-block needs 32 KB of operands per SM per 512 tensor
// cycles = 64 B/clk/SM, and the chip delivers ~6,300 B/clk L2->SM in total (42.6 B/clk/SM): a kernel that re-loads its
// A operand for each of the 9 filter taps cannot pass ~62% tensor-pipe utilisation (ncu: 61.6%).  So the A operand is
// loaded ONCE per 64-channel chunk and the 9 taps are 9 MMAs on the same shared-memory slab with shifted descriptors:
//
//   activations in HBM:  bf16 [Q][C], DENSE: q = (position * W + y) * W + x, no padding of any kind.  Filter tap
//   (dy, dx) of output pixel q reads input pixel q + dy*W + dx — a uniform row shift of the GEMM's A operand — which is
//   the right neighbour except where it crosses an image edge (x + dx or y + dy outside 0..W-1: the shifted index lands
//   in the neighbouring row / position).  Those (pixel, tap) products must contribute zero ("same" padding,
//   model.py:39,42), and tcgen05.mma has exactly the operand for it: DISABLE-OUTPUT-LANE, a 256-bit mask of accumulator
//   rows the instruction must not update.  Each tap's MMAs carry the mask of the rows whose neighbour is off the board;
//   the centre tap is issued first (it is valid for every row, and the tile's first MMA overwrites the accumulator).
//   The masks depend only on (first q of the tile) mod W*W: a table of W*W x 9 x 8 words built once per network
//   (k_lane_masks); the issuing warp fetches the next tile's row while it issues the current tile.
//   Round 1 padded instead (one zero pixel per row, one zero row per position: 324 q per 289 pixels), so 10.8% of the
//   MMAs multiplied zeros; an ideal pad-free kernel was measured 12.8% faster (profiles/r02_conv_pad_bound.json).
//   A pair tile = 256 consecutive q; per channel chunk each CTA loads its 128 rows plus a halo of W+1 rows on either
//   side (ONE 2-D TMA box, 164 x 128 B, rows outside the tensor zero-filled) into a 128B-swizzled slab; tap (dy, dx) is
//   the UMMA descriptor starting at slab + (halo + dy*W + dx) * 128 B.  SWIZZLE_128B is a function of the absolute smem
//   address on both the TMA and the MMA side, so a descriptor may start at any 128-B row (tools/umma_rowshift_test.cu,
//   verified on B200 with the descriptor's base-offset field left 0).  Operand traffic per tile: 1,322 KB instead of the
//   2,304 KB of one A load per tap.
//
//   cluster (2,1,1); per CTA 192 threads: warp 0 TMA producer (both CTAs), warp 1 MMA issuer (leader CTA only) +
//   TMEM alloc (both), warps 2-5 epilogue (both).  cta_group::2: each SM stages its own 128 A rows and half of the
//   weights; TMEM holds TWO 128x256 fp32 accumulator stages so the epilogue of tile i (tcgen05.ld, bias, skip, ReLU,
//   bf16, fused 1x1 head convs on the last layer) runs under the MMAs of tile i+1.
//   k order: channel chunk outer (one A slab), filter tap inner (nine 64x256 weight blocks through an 8-stage ring).
#pragma once

#ifndef PR_WIDE_LDST
#define PR_WIDE_LDST 1                  // 256-bit epilogue loads/stores (one full 32-B sector per lane per instruction): +2% on the convs
#endif
#ifdef SGO_CONV_ABLATE
#define PR_DBG(bit) ((a.dbg & (bit)) != 0)
#else
#define PR_DBG(bit) false                // the shipped library has no switch that makes the timed kernel skip work
#endif
#ifndef PR_COALESCED_STORE
#define PR_COALESCED_STORE 0            // 1: epilogue stores staged through shared memory so that a warp writes whole 128-B lines.  Measured:
#endif                                  // convs +0.6%, stem -13%, forward +0.4% (profiles/r02_conv_coalesced_store_ab.json) — not adopted
#define PR_STAGE_PITCH 144              // bytes per staged row: 128 + 16, so that 8 lanes writing / reading 16 B each hit 32 distinct banks
#define PR_STAGE_BYTES (PR_COALESCED_STORE ? 4 * 32 * PR_STAGE_PITCH : 0)   // one 32-row x 128-B staging tile per epilogue warp
#define PR_SLABS 3
#define PR_SLAB_BYTES (21 * 1024)                  // >= (128 + 2 * (W + 1)) * 128 for W <= 19
#define PR_MASK_WORDS 72                           // 9 taps (issue order) x 8 words of disable-output-lane mask per tile alignment
#define PR_BSTAGES 8
#define PR_B_BYTES (128 * 128)
#define PR_MAX_HALO ((PR_SLAB_BYTES / 128 - 128) / 2)

struct PairMaps {
    CUtensorMap act;                          // activations [Q][C] (or the stem's im2col [Q][192]): box (64, 128 + 2*halo)
    CUtensorMap w;                            // weights box (64 ci, 128 co)
};

struct PairArgs {
    int W, PX, Q, n_tiles, w_row0, relu;      // PX = W*W pixels per position, Q = positions * PX rows in all
    int n_taps, kchunks, halo;                // 9 x 4, halo W+1 for the tower convs; 1 x 3, halo 0 for the stem GEMM over the im2col tensor
    const uint32_t *masks;                    // [PX][PR_MASK_WORDS] lane masks by tile alignment (nullptr: one tap, nothing to mask)
    int dbg;                                  // timing ablations, compiled in only with -DSGO_CONV_ABLATE (tools/conv_variants.py): 1 = no epilogue
                                              // global traffic, 2 = no A loads, 4 = no B loads, 32 = no activation stores, 64 = no skip loads,
                                              // 16 = all lane masks zero (what the masks cost)
    const float *bias;
    const __nv_bfloat16 *skip;
    __nv_bfloat16 *out;                       // nullptr: do not store the activations (last layer feeding only the heads)
    int32_t *err;
    // fused 1x1 head convolutions (model.py:73,83) on the last layer: the post-ReLU features go out as the two fp32
    // A matrices of the dense-head GEMMs, featp / featv [pos][feat_ld], column = pix*2 + channel (HWC flatten, model.py:76,86)
    const float *head_w4;                     // [C][4] folded weights, nullptr = off
    const float *head_b4;                     // [4]
    float *featp, *featv;
    int feat_ld;
    // MODE 1 (dense heads as a TF32 GEMM, rows = positions): fp32 output [row][ldo], columns [0, ncols) of this N tile, rows < n_valid
    float *outf;
    int ldo, ncols, n_valid;
};

struct PairSmemTail {
    uint64_t a_full[PR_SLABS], a_empty[PR_SLABS], b_full[PR_BSTAGES], b_empty[PR_BSTAGES], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
    alignas(16) uint32_t lane_masks[2][PR_MASK_WORDS + 8];       // this tile's / the next tile's masks (written and read by the MMA warp only)
    float bias[TW_C];
    float4 w4[TW_C];
};
#define PR_TAIL_BYTES (((int)sizeof(PairSmemTail) + 15) / 16 * 16)
#define PR_SMEM_BYTES (PR_SLABS * PR_SLAB_BYTES + PR_BSTAGES * PR_B_BYTES + PR_TAIL_BYTES + PR_STAGE_BYTES + 1024)

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
// 2-SM TMA load: data lands in THIS CTA's smem, the transaction bytes are credited to the
// LEADER CTA's mbarrier (copy_sm100_tma.hpp SM100_TMA_2SM_LOAD_*)
__device__ __forceinline__ void tma2_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar) & PR_PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
// the same with the two 64-bit smem descriptors given as (low word, shared high word)
__device__ __forceinline__ void umma2_bf16_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                 "mov.b64 da, {%1, %3};\n\t"
                 "mov.b64 db, {%2, %3};\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// bf16 MMA that leaves the accumulator rows flagged in the 256-bit mask m0..m7 untouched (bit i of word w = row 32w + i of
// the pair tile: rows 0-127 live in the leader CTA's TMEM lanes, 128-255 in its peer's)
__device__ __forceinline__ void umma2_bf16_masked(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate,
                                                  const uint4 &ma, const uint4 &mb)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                 "mov.b64 da, {%1, %3};\n\t"
                 "mov.b64 db, {%2, %3};\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, {%6, %7, %8, %9, %10, %11, %12, %13}, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate),
                   "r"(ma.x), "r"(ma.y), "r"(ma.z), "r"(ma.w), "r"(mb.x), "r"(mb.y), "r"(mb.z), "r"(mb.w) : "memory");
}
// fp32 operands read as TF32 (10-bit mantissa), K = 8 per instruction = the same 32 bytes per k-step as bf16 K = 16
__device__ __forceinline__ void umma2_tf32_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
                 "mov.b64 da, {%1, %3};\n\t"
                 "mov.b64 db, {%2, %3};\n\t"
                 "setp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %4, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
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
// issue order of the 9 filter taps: the centre tap first (valid for every pixel: the first MMA of a tile overwrites the
// accumulator, so it must not skip any row), then the rest in raster order.  tap = (dy + 1) * 3 + (dx + 1).
__device__ __forceinline__ int tap_of(int i) { return i == 0 ? 4 : (i <= 4 ? i - 1 : i); }
// M=256 (pair), N=256, bf16 x bf16 -> f32, both operands K-major
#define PR_IDESC ((1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24))
// the same shape with TF32 operands (format code 2)
#define PR_IDESC_TF32 ((1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24))

// MODE 0: the tower convolutions / stem GEMM (bf16 operands, bf16 activations out).
// MODE 1: the dense heads (model.py:77-92) as plain GEMMs on the same pipeline: rows = positions, A = the fp32 feature
//         matrix, B = the transposed dense weights, TF32 tensor-core math, fp32 rows out (+bias, ReLU on the value half).
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TW_THREADS, 1)
k_conv3x3_pair(const __grid_constant__ PairMaps maps, PairArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_b = smem + PR_SLABS * PR_SLAB_BYTES;
    PairSmemTail *tail = reinterpret_cast<PairSmemTail *>(smem_b + PR_BSTAGES * PR_B_BYTES);
    uint8_t *stage_base = reinterpret_cast<uint8_t *>(tail) + PR_TAIL_BYTES;      // epilogue store staging (PR_STAGE_BYTES)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    constexpr int KEL = MODE == 0 ? TW_KCH : TW_KCH / 2;          // elements per 128-byte K chunk: 64 bf16 / 32 fp32

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.act) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.w) : "memory");
        for (int s = 0; s < PR_SLABS; s++) { mbar_init(&tail->a_full[s], 1); mbar_init(&tail->a_empty[s], 1); }
        for (int s = 0; s < PR_BSTAGES; s++) { mbar_init(&tail->b_full[s], 1); mbar_init(&tail->b_empty[s], 1); }
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

    if (warp == 0) {
        {
            // ---- producer: one A slab per channel chunk, one weight block per (chunk, tap) -------------------------
            // (whole warp in the control flow, one elected lane issues: see the MMA issuer below)
            const bool leader = elect_one();
            uint32_t slab = 0, sphase = 0, stage = 0, phase = 0;
            const uint32_t slab_bytes = (uint32_t)(128 + 2 * a.halo) * 128u;
            const bool no_a = PR_DBG(2), no_b = PR_DBG(4);
            bool ok = true;
            for (int tile = pair; tile < a.n_tiles && ok; tile += n_pairs) {
                const int q_lo = tile * 256 + (int)rank * 128 - a.halo;          // first slab row (may be < 0: zero filled)
                for (int kc = 0; kc < a.kchunks && ok; kc++) {
                    ok = mbar_wait(&tail->a_empty[slab], sphase ^ 1, a.err);
                    if (!ok) break;
                    if (leader) {
                        if (rank == 0) mbar_expect_tx(&tail->a_full[slab], no_a ? 0u : 2u * slab_bytes);  // bytes of BOTH CTAs
                        if (!no_a) tma2_load_2d(smem + (size_t)slab * PR_SLAB_BYTES, &maps.act, kc * KEL, q_lo, &tail->a_full[slab]);
                    }
                    if (++slab == PR_SLABS) { slab = 0; sphase ^= 1; }
                    for (int tap = 0; tap < a.n_taps; tap++) {
                        ok = mbar_wait(&tail->b_empty[stage], phase ^ 1, a.err);
                        if (!ok) break;
                        if (leader) {
                            if (rank == 0) mbar_expect_tx(&tail->b_full[stage], no_b ? 0u : 2u * PR_B_BYTES);
                            if (!no_b) tma2_load_2d(smem_b + (size_t)stage * PR_B_BYTES, &maps.w, kc * KEL,
                                                    a.w_row0 + (a.n_taps == 1 ? 0 : tap_of(tap)) * TW_C + (int)rank * 128, &tail->b_full[stage]);
                        }
                        if (++stage == PR_BSTAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ---- MMA issuer: 9 taps x 4 K-steps on one slab, descriptor start shifted by (dy*W + dx) rows, rows whose
            // neighbour is off the board masked out of the update ---------------------------------------------------------
            // The WHOLE warp runs the (warp-uniform) control flow and the barrier waits; one elected lane issues.  Written
            // this way — and with the descriptors kept as a constant high word plus a running 32-bit low word — the loop
            // is ~45 SASS instructions per k-block instead of ~130 (a lane==0 branch makes nvcc wrap every tcgen05
            // instruction in an ELECT/R2UR.BROADCAST waterfall loop); at ~130 the single issuing thread needed about as
            // long as the 512 tensor cycles of the k-block itself (ncu: pipe 73% active with no barrier stalls).
            const bool leader = elect_one();
            uint32_t slab = 0, sphase = 0, stage = 0, phase = 0;
            const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);       // SBO 1024 B, version 1, SWIZZLE_128B
            const uint32_t a_lo0 = ((smem_u32(smem) & 0x3FFFF) >> 4) | (1u << 16);          // + slab * (PR_SLAB_BYTES >> 4)
            const uint32_t b_lo0 = ((smem_u32(smem_b) & 0x3FFFF) >> 4) | (1u << 16);        // + stage * (PR_B_BYTES >> 4)
            // lane masks: lanes 0-17 move one 16-byte piece each of a tile's 288-byte table row; the row of the NEXT tile
            // is requested at the start of a tile and parked in shared memory once the first channel chunk has been issued
            const bool masked = MODE == 0 && a.masks != nullptr && !PR_DBG(16);
            auto mask_row = [&](int tile) {
                const int al = (int)(((long long)tile * 256) % a.PX);
                return reinterpret_cast<const uint4 *>(a.masks + (size_t)al * PR_MASK_WORDS)[lane];
            };
            for (int i = lane; i < 2 * (PR_MASK_WORDS + 8); i += 32) (&tail->lane_masks[0][0])[i] = 0u;
            __syncwarp();
            if (masked && pair < a.n_tiles && lane < PR_MASK_WORDS / 4) reinterpret_cast<uint4 *>(tail->lane_masks[0])[lane] = mask_row(pair);
            __syncwarp();
            bool ok = true;
            int it = 0;
            for (int tile = pair; tile < a.n_tiles && ok; tile += n_pairs, it++) {
                const uint32_t d_tmem = tmem_base + (uint32_t)(it & 1) * 256u;
                const bool fetch = masked && tile + n_pairs < a.n_tiles && lane < PR_MASK_WORDS / 4;
                uint4 next_masks = make_uint4(0u, 0u, 0u, 0u);
                if (fetch) next_masks = mask_row(tile + n_pairs);
                const uint32_t mk = smem_u32(tail->lane_masks[it & 1]);                 // shared-space address of this tile's masks
                ok = mbar_wait(&tail->tmem_empty[it & 1], ((it >> 1) & 1) ^ 1, a.err);  // both epilogues drained this stage
                if (!ok) break;
                tc_fence_after();
                uint32_t accum = 0;
                for (int kc = 0; kc < a.kchunks && ok; kc++) {
                    ok = mbar_wait(&tail->a_full[slab], sphase, a.err);
                    if (!ok) break;
                    const uint32_t a_lo = a_lo0 + slab * (PR_SLAB_BYTES >> 4) + (uint32_t)a.halo * 8u;
                    // one tap = wait for its weight block, 4 MMAs (K = 4 x 16) with its lane mask, free the block.  The 9-tap
                    // loop is unrolled so that the tap's row shift and mask offsets are immediates: the issuing thread has
                    // ~500 tensor cycles per tap, and every instruction between two UTCHMMA groups counts against them
                    auto issue_tap = [&](const int tap, const int shift) {
                        uint4 ma, mb;                                                   // (zeros when nothing is masked)
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ma.x), "=r"(ma.y), "=r"(ma.z), "=r"(ma.w) : "r"(mk + 32u * tap));
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(mb.x), "=r"(mb.y), "=r"(mb.z), "=r"(mb.w) : "r"(mk + 32u * tap + 16u));
                        ok = mbar_wait(&tail->b_full[stage], phase, a.err);
                        if (!ok) return;
                        tc_fence_after();
                        const uint32_t alo = a_lo + (uint32_t)(shift * 8), blo = b_lo0 + stage * (PR_B_BYTES >> 4);
                        if (leader) {
#pragma unroll
                            for (int k = 0; k < TW_KCH / 16; k++) {
                                if constexpr (MODE == 0) umma2_bf16_masked(d_tmem, alo + 2 * k, blo + 2 * k, desc_hi, PR_IDESC, accum, ma, mb);
                                else umma2_tf32_lohi(d_tmem, alo + 2 * k, blo + 2 * k, desc_hi, PR_IDESC_TF32, accum);
                                accum = 1;
                            }
                            umma2_commit_mc(&tail->b_empty[stage]);                     // frees the weight slot in both CTAs
                        }
                        accum = 1;
                        if (++stage == PR_BSTAGES) { stage = 0; phase ^= 1; }
                    };
                    if (MODE == 0 && a.n_taps == 9) {
#pragma unroll
                        for (int tap = 0; tap < 9; tap++) {
                            const int tp = tap == 0 ? 4 : (tap <= 4 ? tap - 1 : tap);   // tap_of(tap), folded at compile time
                            if (ok) issue_tap(tap, (tp / 3 - 1) * a.W + (tp % 3 - 1));
                        }
                    } else {
                        issue_tap(0, 0);
                    }
                    if (ok && leader) umma2_commit_mc(&tail->a_empty[slab]);            // frees the slab in both CTAs
                    if (++slab == PR_SLABS) { slab = 0; sphase ^= 1; }
                    if (kc == 0) {                                                      // the next tile's masks have arrived by now
                        if (fetch) reinterpret_cast<uint4 *>(tail->lane_masks[(it + 1) & 1])[lane] = next_masks;
                        __syncwarp();
                    }
                }
                if (ok && leader) umma2_commit_mc(&tail->tmem_full[it & 1]);
            }
        }
    } else {
        // ---- epilogue: thread = one q (pixel) of the pair tile -------------------------------------------------------------
        const int qw = warp & 3;
        const int r = (int)rank * 128 + qw * 32 + lane;
        bool ok = true;
        int it = 0;
        if constexpr (MODE == 1) {
            // ---- dense-head GEMM epilogue: thread = one position; fp32 row out (+bias, ReLU for the value hidden layer)
            for (int tile = pair; tile < a.n_tiles; tile += n_pairs, it++) {
                const int acc = it & 1;
                const int q = tile * 256 + r;
                const bool valid = q < a.n_valid;
                if (ok) ok = mbar_wait(&tail->tmem_full[acc], (it >> 1) & 1, a.err);
                ok = __all_sync(SGO_FULL, ok);
                if (!ok) break;
                tc_fence_after();
                const uint32_t t_addr = tmem_base + ((uint32_t)(qw * 32) << 16) + acc * 256;
                float *orow = a.outf + (size_t)q * a.ldo;
#pragma unroll 1
                for (int c = 0; c < TW_C / 32; c++) {
                    uint32_t v[32];
                    tmem_ld32(t_addr + c * 32, v);
                    if (valid && c * 32 < a.ncols) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            float f = __uint_as_float(v[j]) + tail->bias[c * 32 + j];
                            if (a.relu) f = fmaxf(f, 0.f);
                            v[j] = __float_as_uint(f);
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) stg256(orow + c * 32 + j * 8, v + j * 8);     // (the buffer is padded to whole 32-column groups)
                    }
                }
                tc_fence_before();
                mbar_arrive_leader(&tail->tmem_empty[acc]);
            }
        } else
        for (int tile = pair; tile < a.n_tiles; tile += n_pairs, it++) {
            const int acc = it & 1;
            const int q = tile * 256 + r;
            bool valid = q < a.Q;                                    // (only the last tile has rows past the end)
            if (PR_DBG(1)) valid = false;
            const size_t gofs = (size_t)q * TW_C;
            if (valid && a.skip) {                                   // idle until the MMAs finish: pull the skip row towards L2
#pragma unroll
                for (int j = 0; j < 4; j++) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.skip + gofs + j * 64));
            }
            if (ok) ok = mbar_wait(&tail->tmem_full[acc], (it >> 1) & 1, a.err);
            ok = __all_sync(SGO_FULL, ok);
            if (!ok) break;
            tc_fence_after();
            float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f;
            const uint32_t t_addr = tmem_base + ((uint32_t)(qw * 32) << 16) + acc * 256;
#pragma unroll 1
            for (int c = 0; c < TW_C / 32; c++) {                    // (software-pipelining the TMEM loads over a fully
                uint32_t sk[16];                                     //  unrolled loop was measured slower: profiles/r01_conv_epilogue_ab.json)
                const bool do_skip = valid && a.skip && !PR_DBG(64);
                if (do_skip) {
#if PR_WIDE_LDST
                    ldg256(a.skip + gofs + c * 32, sk);
                    ldg256(a.skip + gofs + c * 32 + 16, sk + 8);
#else
                    const uint4 *sp = reinterpret_cast<const uint4 *>(a.skip + gofs + c * 32);
#pragma unroll
                    for (int j = 0; j < 4; j++) reinterpret_cast<uint4 *>(sk)[j] = sp[j];
#endif
                }
                uint32_t v[32];
                tmem_ld32(t_addr + c * 32, v);
                uint32_t ow[16];
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        float f0 = __uint_as_float(v[2 * j]) + tail->bias[c * 32 + 2 * j];
                        float f1 = __uint_as_float(v[2 * j + 1]) + tail->bias[c * 32 + 2 * j + 1];
                        if (a.skip) {
                            __nv_bfloat162 s2 = *reinterpret_cast<const __nv_bfloat162 *>(&sk[j]);
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
#if !PR_COALESCED_STORE
                    if (a.out && !PR_DBG(32)) {
                        stg256(a.out + gofs + c * 32, ow);
                        stg256(a.out + gofs + c * 32 + 16, ow + 8);
                    }
#endif
                }
#if PR_COALESCED_STORE
                // A lane holds 64 B of ITS row: stored directly, a warp instruction touches 32 different 128-B lines with one
                // 32-B sector each.  Staged through shared memory (two iterations = 128 B per row), eight lanes then write
                // one whole line: a quarter of the L1 / L2 transactions for the same bytes.
                if (a.out && !PR_DBG(32)) {                          // (warp-uniform)
                    uint8_t *st = stage_base + (size_t)qw * (32 * PR_STAGE_PITCH);
                    uint4 *mine = reinterpret_cast<uint4 *>(st + lane * PR_STAGE_PITCH + (c & 1) * 64);
#pragma unroll
                    for (int j = 0; j < 4; j++) mine[j] = reinterpret_cast<const uint4 *>(ow)[j];
                    if (c & 1) {
                        __syncwarp();
                        const int q_warp = tile * 256 + (int)rank * 128 + qw * 32;
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            const int row = 4 * k + (lane >> 3);
                            const uint4 val = *reinterpret_cast<const uint4 *>(st + row * PR_STAGE_PITCH + (lane & 7) * 16);
                            const int qq = q_warp + row;
                            if (qq < a.Q && !PR_DBG(1))
                                *reinterpret_cast<uint4 *>(a.out + (size_t)qq * TW_C + (c - 1) * 32 + (lane & 7) * 8) = val;
                        }
                        __syncwarp();
                    }
                }
#endif
            }
            tc_fence_before();
            mbar_arrive_leader(&tail->tmem_empty[acc]);
            if (valid && a.head_w4) {
                const int pos = q / a.PX, pix = q - pos * a.PX;
                const size_t fo = (size_t)pos * a.feat_ld + (size_t)pix * 2;
                *reinterpret_cast<float2 *>(a.featp + fo) = make_float2(fmaxf(h0 + a.head_b4[0], 0.f), fmaxf(h1 + a.head_b4[1], 0.f));
                *reinterpret_cast<float2 *>(a.featv + fo) = make_float2(fmaxf(h2 + a.head_b4[2], 0.f), fmaxf(h3 + a.head_b4[3], 0.f));
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
