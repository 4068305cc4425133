// stem_fused.cuh — k_stem_fused: the stem (Conv3x3 "valid" 17 -> 256 + BN + ReLU, model.py:57-61) with the im2col built
// INSIDE the GEMM's producer, straight from the packed bitboards into the 128B-swizzled shared-memory tile the MMA reads:
// the [Q][192] im2col tensor of k_stem_im2col (1.77 GB written and read back per 16,384 positions) never exists.
//
//   CTA pairs as in conv_pair.cuh, 448 threads per CTA: warp 0 loads this CTA's half of the stem weights once (they stay
//   in shared memory: 128 co x 192 k = 48 KB), warp 1 issues the MMAs (leader CTA; M = 256, N = 256, K = 192 = 12 MMAs per
//   tile), warps 2-5 are the epilogue (bias, ReLU, bf16 rows out), warps 6-13 the producers: per tile they gather the
//   16-plane cell masks of the positions (two at 19x19) the CTA's 128 rows touch — input-plane construction
//   (play.py:295-299) and the symmetry gather (symmetry.py:45-114) as in k_stem_im2col — and each thread writes half a row
//   of the A tile (96 bf16 values 0 / 1 / +-1 in 12 swizzled 16-byte stores).  A tiles, cell masks and accumulators are
//   double-buffered.
#pragma once

#define SF_EPI_WARPS 4                             // warps 2-5, one per TMEM lane quarter (8 warps, two per quarter, measured 8% slower: r02_stem_fused_ab.json)
#define SF_PRODUCERS 256                           // warps 6-13: two threads per A-tile row
#define SF_THREADS (32 * (2 + SF_EPI_WARPS) + SF_PRODUCERS)
#define SF_EPI_CHUNKS (TW_C / 32 / (SF_EPI_WARPS / 4))   // 32-column accumulator chunks per epilogue warp
#define SF_CHUNK_BYTES (128 * 128)                 // one K chunk (64 bf16) of 128 rows
#define SF_ABUF_BYTES (3 * SF_CHUNK_BYTES)         // one A tile: K = 192
// positions under one CTA's 128 rows: 127 / PX + 2 at most (PX = (S-2)^2 >= 1): 2 at 19x19, 4 at 9x9, 16 at 5x5, 129 at 3x3;
// their cell masks take (127 / PX + 2) * S * S entries: 722 at 19x19, 1161 at 3x3 (the maximum over 3 <= S <= SGO_MAXS)
#define SF_MAX_POS 132
#define SF_CELLS 1168

struct StemArgs {
    const Board *boards;
    const int32_t *index, *syms;
    int n, S, W, PX, Q, n_tiles;
    const float *bias;
    __nv_bfloat16 *out;
    int32_t *err;
};

struct StemSmemTail {
    uint64_t w_full, a_full[2], a_empty[2], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
    float bias[TW_C];
    int4 hdr[2][SF_MAX_POS];                       // (board index or -1, symmetry, to_move, history head) of the positions under the tile
    uint16_t cell[2][SF_CELLS];                    // their 16 stone planes per cell as one mask; double-buffered like the A tiles
};
#define SF_SMEM_BYTES (SF_ABUF_BYTES /* weights */ + 2 * SF_ABUF_BYTES + (int)sizeof(StemSmemTail) + 1024)

// 16 consecutive k of one im2col row (sector SEC of 12) as two swizzled 16-byte shared-memory stores
template <int SEC>
__device__ __forceinline__ void im2col_sector_smem(const uint32_t (&m)[9], uint32_t tmv, uint32_t row_addr, uint32_t r7)
{
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t v2[2];
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
            const int k = SEC * 16 + 2 * j + hh, tap = k / 17, p = k - tap * 17;
            v2[hh] = k >= 153 ? 0u : (p == 16 ? tmv : ((m[tap < 9 ? tap : 0] >> p) & 1u) * 0x3F80u);
        }
        w[j] = v2[0] | (v2[1] << 16);
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
        constexpr int dummy = 0; (void)dummy;
        const int g = 2 * SEC + h, kc = g >> 3, jj = g & 7;            // 16-byte piece g of the row: K chunk kc, slot jj before the swizzle
        const uint32_t addr = row_addr + (uint32_t)kc * SF_CHUNK_BYTES + (((uint32_t)jj ^ r7) << 4);
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[4 * h]), "r"(w[4 * h + 1]), "r"(w[4 * h + 2]), "r"(w[4 * h + 3]) : "memory");
    }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SF_THREADS, 1)
k_stem_fused(const __grid_constant__ PairMaps maps /* .w = stem weights [256][192], box (64, 128) */, StemArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *wsm = smem;                                   // [3 chunks][128 co][64 k]
    uint8_t *abuf = smem + SF_ABUF_BYTES;                  // [2][3 chunks][128 rows][64 k]
    StemSmemTail *tail = reinterpret_cast<StemSmemTail *>(abuf + 2 * SF_ABUF_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.w) : "memory");
        mbar_init(&tail->w_full, 1);
        for (int s = 0; s < 2; s++) {
            mbar_init(&tail->a_full[s], 2 * SF_PRODUCERS); mbar_init(&tail->a_empty[s], 1);
            mbar_init(&tail->tmem_full[s], 1); mbar_init(&tail->tmem_empty[s], 2 * 32 * SF_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tail->tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < TW_C; i += blockDim.x) tail->bias[i] = a.bias[i];
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_base;

    if (warp == 0) {
        // ---- this CTA's 128 output channels of the stem weights, once ------------------------------------------------------
        if (elect_one()) {
            if (rank == 0) mbar_expect_tx(&tail->w_full, 2u * SF_ABUF_BYTES);               // bytes of BOTH CTAs
            for (int kc = 0; kc < 3; kc++)
                tma2_load_2d(wsm + (size_t)kc * SF_CHUNK_BYTES, &maps.w, kc * TW_KCH, (int)rank * 128, &tail->w_full);
        }
    } else if (warp == 1) {
        if (rank == 0) {
            const bool leader = elect_one();
            const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);       // SBO 1024 B, version 1, SWIZZLE_128B
            const uint32_t a_lo0 = ((smem_u32(abuf) & 0x3FFFF) >> 4) | (1u << 16);
            const uint32_t b_lo0 = ((smem_u32(wsm) & 0x3FFFF) >> 4) | (1u << 16);
            bool ok = mbar_wait(&tail->w_full, 0, a.err);
            int it = 0;
            for (int tile = pair; tile < a.n_tiles && ok; tile += n_pairs, it++) {
                const int b = it & 1;
                const uint32_t d_tmem = tmem_base + (uint32_t)b * 256u;
                ok = mbar_wait(&tail->tmem_empty[b], ((it >> 1) & 1) ^ 1, a.err);
                if (!ok) break;
                ok = mbar_wait(&tail->a_full[b], (it >> 1) & 1, a.err);                      // all 256 producer rows of the pair are in place
                if (!ok) break;
                tc_fence_after();
                if (leader) {
#pragma unroll
                    for (int kc = 0; kc < 3; kc++)
#pragma unroll
                        for (int k = 0; k < TW_KCH / 16; k++)
                            umma2_bf16_lohi(d_tmem, a_lo0 + (uint32_t)b * (SF_ABUF_BYTES >> 4) + kc * (SF_CHUNK_BYTES >> 4) + 2 * k,
                                            b_lo0 + kc * (SF_CHUNK_BYTES >> 4) + 2 * k, desc_hi, PR_IDESC, (kc | k) ? 1u : 0u);
                    umma2_commit_mc(&tail->a_empty[b]);                                     // the A tile may be rebuilt (both CTAs)
                    umma2_commit_mc(&tail->tmem_full[b]);
                }
            }
        }
    } else if (warp < 2 + SF_EPI_WARPS) {
        // ---- epilogue: thread = one pixel row; + folded-BN bias, ReLU, bf16, 64-B stores -----------------------------------
        const int qw = warp & 3, c0 = ((warp - 2) >> 2) * SF_EPI_CHUNKS;
        const int r = (int)rank * 128 + qw * 32 + lane;
        bool ok = true;
        int it = 0;
        for (int tile = pair; tile < a.n_tiles; tile += n_pairs, it++) {
            const int acc = it & 1;
            const int q = tile * 256 + r;
            const bool valid = q < a.Q;
            if (ok) ok = mbar_wait(&tail->tmem_full[acc], (it >> 1) & 1, a.err);
            ok = __all_sync(SGO_FULL, ok);
            if (!ok) break;
            tc_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(qw * 32) << 16) + acc * 256;
            __nv_bfloat16 *orow = a.out + (size_t)q * TW_C;
#pragma unroll 1
            for (int c = c0; c < c0 + SF_EPI_CHUNKS; c++) {
                uint32_t v[32];
                tmem_ld32(t_addr + c * 32, v);
                if (valid) {
                    uint32_t ow[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const float f0 = fmaxf(__uint_as_float(v[2 * j]) + tail->bias[c * 32 + 2 * j], 0.f);
                        const float f1 = fmaxf(__uint_as_float(v[2 * j + 1]) + tail->bias[c * 32 + 2 * j + 1], 0.f);
                        __nv_bfloat162 p = __floats2bfloat162_rn(f0, f1);
                        ow[j] = *reinterpret_cast<uint32_t *>(&p);
                    }
                    stg256(orow + c * 32, ow);
                    stg256(orow + c * 32 + 16, ow + 8);
                }
            }
            tc_fence_before();
            mbar_arrive_leader(&tail->tmem_empty[acc]);
        }
    } else {
        // ---- producers: cell masks of the positions under this CTA's rows, then half an im2col row per thread -------------------
        const int pt = threadIdx.x - 32 * (2 + SF_EPI_WARPS);                                   // 0..255: row pt & 127 of the CTA's tile, K half pt >> 7
        const int S = a.S, W = a.W, SS = S * S;
        bool ok = true;
        int it = 0;
        for (int tile = pair; tile < a.n_tiles; tile += n_pairs, it++) {
            const int b = it & 1;
            const int q_cta = tile * 256 + (int)rank * 128;
            const int p0 = q_cta / a.PX;
            int np = (q_cta + 127) / a.PX - p0 + 1;
            if (p0 + np > a.n) np = a.n - p0;
            int4 *hdr = tail->hdr[b];
            uint16_t *cell = tail->cell[b];
            for (int pq = pt; pq < np; pq += SF_PRODUCERS) {
                int4 h;
                h.x = a.index ? a.index[p0 + pq] : p0 + pq;
                h.y = a.syms ? (a.syms[p0 + pq] & 7) : 0;
                h.z = 0; h.w = 0;
                if (h.x >= 0) { h.z = a.boards[h.x].to_move; h.w = a.boards[h.x].head; }
                hdr[pq] = h;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int c = pt; c < np * SS; c += SF_PRODUCERS) {
                const int pq = c / SS, cc = c - pq * SS;
                const int4 h = hdr[pq];
                uint32_t m = 0;
                if (h.x >= 0) {
                    const Board *bd = a.boards + h.x;
                    int y = cc / S, x = cc - y * S, sy, sx;
                    sym_src_t(S, h.y, y, x, sy, sx);
#pragma unroll
                    for (int k = 0; k < SGO_HIST; k++) {
                        const int slot = (h.w + SGO_HIST - k) & (SGO_HIST - 1);
                        const uint32_t bl = (bd->st[slot][0][sy] >> sx) & 1u, wh = (bd->st[slot][1][sy] >> sx) & 1u;
                        const uint32_t own = h.z == 1 ? bl : wh, opp = h.z == 1 ? wh : bl;
                        m |= (own << (2 * k)) | (opp << (2 * k + 1));
                    }
                }
                cell[c] = (uint16_t)m;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");                   // the cell masks are complete (producer warps only)
            if (ok) ok = mbar_wait(&tail->a_empty[b], ((it >> 1) & 1) ^ 1, a.err);           // the MMAs of the tile before last are done with this buffer
            const int row = pt & 127, q = q_cta + row;
            if (ok && q < a.Q) {
                const int pos = q / a.PX, pix = q - pos * a.PX, pq = pos - p0;
                const int y = pix / W, x = pix - y * W;
                uint32_t m[9];
#pragma unroll
                for (int tap = 0; tap < 9; tap++) m[tap] = cell[pq * SS + (y + tap / 3) * S + x + tap % 3];
                const int tmq = hdr[pq].z;
                const uint32_t tmv = tmq == 1 ? 0x3F80u : (tmq == 0 ? 0u : 0xBF80u);        // bf16 +1 / -1 (plane 16); 0 for an empty slot
                const uint32_t row_addr = smem_u32(abuf) + (uint32_t)b * SF_ABUF_BYTES + (uint32_t)row * 128u, r7 = (uint32_t)row & 7u;
                if (pt < 128) {                                                             // warp-uniform: k 0..95 / 96..191
                    im2col_sector_smem<0>(m, tmv, row_addr, r7); im2col_sector_smem<1>(m, tmv, row_addr, r7);
                    im2col_sector_smem<2>(m, tmv, row_addr, r7); im2col_sector_smem<3>(m, tmv, row_addr, r7);
                    im2col_sector_smem<4>(m, tmv, row_addr, r7); im2col_sector_smem<5>(m, tmv, row_addr, r7);
                } else {
                    im2col_sector_smem<6>(m, tmv, row_addr, r7); im2col_sector_smem<7>(m, tmv, row_addr, r7);
                    im2col_sector_smem<8>(m, tmv, row_addr, r7); im2col_sector_smem<9>(m, tmv, row_addr, r7);
                    im2col_sector_smem<10>(m, tmv, row_addr, r7); im2col_sector_smem<11>(m, tmv, row_addr, r7);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core's async proxy
            mbar_arrive_leader(&tail->a_full[b]);
            // no third barrier: hdr / cell are double-buffered, and the next use of buffer b lies behind the two barriers of tile it + 1
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}
