// conv_wide.cuh — k_conv3x3_wide: the tower convolutions on 512-row super-tiles.  A MEASURED ALTERNATIVE to the 256-row
// pair tile of conv_pair.cuh, compiled into the tower only with -DSGO_CONV_WIDE_TILES: it is correct (it passes
// tests/test_gpu_tower.py) and 2% faster at burst clocks, but 0.5% slower sustained under the power cap
// (profiles/r02_conv_wide_tiles_ab.json): what it saves in L2->SM weight bytes it spends on shared-memory operand reads.
//
// Why: under the board power cap the kernel's clock is set by the energy it spends per useful MAC, and the sustained
// ablations (profiles/r02_conv_power_ablation.json) put 13% of the time under the cap on the WEIGHT loads alone: with a
// 256 x 256 pair tile every CTA pulls all 9 x 256 x 256 weights out of L2 once per 256 output pixels (1,152 KB per
// pair tile, against 168 KB of activations).  TMEM holds 512 accumulator columns per SM: instead of two stages of
// 128 rows x 256 channels, this kernel keeps two stages of (2 x 128 rows) x 128 channels — a CTA owns 256 pixels
// and walks the 256 output channels in two passes of 128.  A weight block (64 ci x 64 co per CTA) now feeds
// 2 x 128 rows per CTA, so the weight traffic per MAC halves; the activation slabs of all four channel chunks
// (256 rows + halo, 37 KB each) stay resident in shared memory for both passes, so they are still loaded once.
// Operand traffic per 512 x 256 outputs: 726 KB per CTA instead of 1,320 KB.
//
//   cluster (2,1,1), 192 threads as in conv_pair.cuh: warp 0 TMA producer, warp 1 MMA issuer (leader CTA) + TMEM
//   alloc, warps 2-5 epilogue.  tcgen05.mma.cta_group::2 with M = 256 (128 rows per CTA), N = 128, K = 16; per
//   (channel chunk, tap): one weight block, 2 row blocks x 4 K-steps = 8 MMAs with the row block's lane mask.
//   Epilogue of pass p (channels 128p .. 128p+127 of both row blocks) runs under the MMAs of the next pass.
#pragma once

#define WD_SLABS 4                                  // = channel chunks: every chunk's slab stays resident for both passes
#define WD_SLAB_BYTES (37 * 1024)                   // >= (256 + 2 * halo) * 128 for halo <= 20
#define WD_BSTAGES 6
#define WD_B_BYTES (64 * 128)                       // 64 output channels (this CTA's half of a 128-channel pass) x 64 input channels
#define WD_MASK_WORDS 144                           // 2 row blocks x 9 taps x 8 words per super-tile alignment
// M=256 (pair), N=128, bf16 x bf16 -> f32, both operands K-major
#define WD_IDESC ((1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24))

struct WideMaps {
    CUtensorMap act;                          // activations [Q][C]: box (64, 128 + halo) — half a slab
    CUtensorMap w;                            // weights: box (64 ci, 64 co)
};

struct WideSmemTail {
    uint64_t a_full[WD_SLABS], a_empty[WD_SLABS], b_full[WD_BSTAGES], b_empty[WD_BSTAGES], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad[3];
    alignas(16) uint32_t lane_masks[2][WD_MASK_WORDS];   // this super-tile's / the next one's masks (MMA warp only)
    float bias[TW_C];
    float4 w4[TW_C];
    float4 hsum[2][128];                                  // fused head convs: partial sums of pass 0, per row block and epilogue thread
};
#define WD_SMEM_BYTES (WD_SLABS * WD_SLAB_BYTES + WD_BSTAGES * WD_B_BYTES + (int)sizeof(WideSmemTail) + 1024)

// PairArgs is reused: W, PX, Q, n_tiles (= super-tiles), w_row0, relu, halo, masks ([PX][WD_MASK_WORDS]), bias, skip, out, err, head_*, feat*.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TW_THREADS, 1)
k_conv3x3_wide(const __grid_constant__ WideMaps maps, PairArgs a)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_b = smem + WD_SLABS * WD_SLAB_BYTES;
    WideSmemTail *tail = reinterpret_cast<WideSmemTail *>(smem_b + WD_BSTAGES * WD_B_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.act) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&maps.w) : "memory");
        for (int s = 0; s < WD_SLABS; s++) { mbar_init(&tail->a_full[s], 1); mbar_init(&tail->a_empty[s], 1); }
        for (int s = 0; s < WD_BSTAGES; s++) { mbar_init(&tail->b_full[s], 1); mbar_init(&tail->b_empty[s], 1); }
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
    const int half_rows = 128 + a.halo;       // rows of one activation box; a slab = two of them = 256 + 2 * halo rows

    if (warp == 0) {
        // ---- producer: per super-tile the four slabs (pass 0 only), per (pass, chunk, tap) one weight block --------------
        const bool leader = elect_one();
        uint32_t stage = 0, phase = 0;
        const uint32_t slab_bytes = (uint32_t)(2 * half_rows) * 128u;
        bool ok = true;
        int si = 0;
        for (int st = pair; st < a.n_tiles && ok; st += n_pairs, si++) {
            const int q_lo = st * 512 + (int)rank * 256 - a.halo;               // first slab row (may be < 0: zero filled)
            for (int pass = 0; pass < 2 && ok; pass++) {
                for (int kc = 0; kc < WD_SLABS && ok; kc++) {
                    if (pass == 0) {
                        ok = mbar_wait(&tail->a_empty[kc], (si & 1) ^ 1, a.err);
                        if (!ok) break;
                        if (leader) {
                            if (rank == 0) mbar_expect_tx(&tail->a_full[kc], 2u * slab_bytes);          // bytes of BOTH CTAs
                            uint8_t *dst = smem + (size_t)kc * WD_SLAB_BYTES;
                            tma2_load_2d(dst, &maps.act, kc * TW_KCH, q_lo, &tail->a_full[kc]);
                            tma2_load_2d(dst + (size_t)half_rows * 128, &maps.act, kc * TW_KCH, q_lo + half_rows, &tail->a_full[kc]);
                        }
                    }
                    for (int tap = 0; tap < 9; tap++) {
                        ok = mbar_wait(&tail->b_empty[stage], phase ^ 1, a.err);
                        if (!ok) break;
                        if (leader) {
                            if (rank == 0) mbar_expect_tx(&tail->b_full[stage], 2u * WD_B_BYTES);
                            tma2_load_2d(smem_b + (size_t)stage * WD_B_BYTES, &maps.w, kc * TW_KCH,
                                         a.w_row0 + tap_of(tap) * TW_C + pass * 128 + (int)rank * 64, &tail->b_full[stage]);
                        }
                        if (++stage == WD_BSTAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ---- MMA issuer (see conv_pair.cuh for why the whole warp runs the control flow and one elected lane issues) ---
            const bool leader = elect_one();
            uint32_t stage = 0, phase = 0;
            const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);       // SBO 1024 B, version 1, SWIZZLE_128B
            const uint32_t a_lo0 = ((smem_u32(smem) & 0x3FFFF) >> 4) | (1u << 16);          // + kc * (WD_SLAB_BYTES >> 4)
            const uint32_t b_lo0 = ((smem_u32(smem_b) & 0x3FFFF) >> 4) | (1u << 16);        // + stage * (WD_B_BYTES >> 4)
            auto fetch_masks = [&](int st, uint4 &m0, uint4 &m1) {                           // 144 words = 36 x 16 bytes over 32 lanes
                const int al = (int)(((long long)st * 512) % a.PX);
                const uint4 *src = reinterpret_cast<const uint4 *>(a.masks + (size_t)al * WD_MASK_WORDS);
                m0 = src[lane];
                m1 = lane < 4 ? src[32 + lane] : make_uint4(0u, 0u, 0u, 0u);
            };
            auto park_masks = [&](int buf, const uint4 &m0, const uint4 &m1) {
                uint4 *dst = reinterpret_cast<uint4 *>(tail->lane_masks[buf]);
                dst[lane] = m0;
                if (lane < 4) dst[32 + lane] = m1;
                __syncwarp();
            };
            {
                uint4 m0 = make_uint4(0u, 0u, 0u, 0u), m1 = m0;
                if (pair < a.n_tiles) fetch_masks(pair, m0, m1);
                park_masks(0, m0, m1);
            }
            bool ok = true;
            int si = 0, pc = 0;
            for (int st = pair; st < a.n_tiles && ok; st += n_pairs, si++) {
                const bool fetch = st + n_pairs < a.n_tiles;
                uint4 nm0 = make_uint4(0u, 0u, 0u, 0u), nm1 = nm0;
                if (fetch) fetch_masks(st + n_pairs, nm0, nm1);
                const uint32_t mk = smem_u32(tail->lane_masks[si & 1]);
                for (int pass = 0; pass < 2 && ok; pass++, pc++) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(pc & 1) * 256u;
                    ok = mbar_wait(&tail->tmem_empty[pc & 1], ((pc >> 1) & 1) ^ 1, a.err);      // both epilogues drained this stage
                    if (!ok) break;
                    tc_fence_after();
                    for (int kc = 0; kc < WD_SLABS && ok; kc++) {
                        if (pass == 0) {
                            ok = mbar_wait(&tail->a_full[kc], si & 1, a.err);
                            if (!ok) break;
                        }
                        const uint32_t a_lo = a_lo0 + (uint32_t)kc * (WD_SLAB_BYTES >> 4) + (uint32_t)a.halo * 8u;
#pragma unroll
                        for (int tap = 0; tap < 9; tap++) {
                            if (!ok) break;
                            const int tp = tap == 0 ? 4 : (tap <= 4 ? tap - 1 : tap);          // tap_of(tap), folded at compile time
                            const int shift = (tp / 3 - 1) * a.W + (tp % 3 - 1);
                            uint4 ma0, mb0, ma1, mb1;                                           // lane masks of row block 0 / 1 for this tap
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ma0.x), "=r"(ma0.y), "=r"(ma0.z), "=r"(ma0.w) : "r"(mk + 32u * tap));
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(mb0.x), "=r"(mb0.y), "=r"(mb0.z), "=r"(mb0.w) : "r"(mk + 32u * tap + 16u));
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ma1.x), "=r"(ma1.y), "=r"(ma1.z), "=r"(ma1.w) : "r"(mk + 288u + 32u * tap));
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(mb1.x), "=r"(mb1.y), "=r"(mb1.z), "=r"(mb1.w) : "r"(mk + 288u + 32u * tap + 16u));
                            ok = mbar_wait(&tail->b_full[stage], phase, a.err);
                            if (!ok) break;
                            tc_fence_after();
                            const uint32_t alo = a_lo + (uint32_t)(shift * 8), blo = b_lo0 + stage * (WD_B_BYTES >> 4);
                            const uint32_t first = (kc == 0 && tap == 0) ? 0u : 1u;             // the pass's first MMA per row block overwrites
                            if (leader) {
#pragma unroll
                                for (int k = 0; k < TW_KCH / 16; k++)
                                    umma2_bf16_masked(d_tmem, alo + 2 * k, blo + 2 * k, desc_hi, WD_IDESC, k == 0 ? first : 1u, ma0, mb0);
#pragma unroll
                                for (int k = 0; k < TW_KCH / 16; k++)
                                    umma2_bf16_masked(d_tmem + 128u, alo + 128u * 8u + 2 * k, blo + 2 * k, desc_hi, WD_IDESC, k == 0 ? first : 1u, ma1, mb1);
                                umma2_commit_mc(&tail->b_empty[stage]);                         // frees the weight slot in both CTAs
                            }
                            if (++stage == WD_BSTAGES) { stage = 0; phase ^= 1; }
                        }
                        if (ok && leader && pass == 1) umma2_commit_mc(&tail->a_empty[kc]);     // the slab is done with for this super-tile
                        if (pass == 0 && kc == 0) {                                             // the next super-tile's masks have arrived by now
                            if (fetch) park_masks((si + 1) & 1, nm0, nm1);
                            else __syncwarp();
                        }
                    }
                    if (ok && leader) umma2_commit_mc(&tail->tmem_full[pc & 1]);
                }
            }
        }
    } else {
        // ---- epilogue: thread = one pixel of each of the CTA's two row blocks; pass p covers channels 128p .. 128p + 127 ------
        const int qw = warp & 3, et = qw * 32 + lane;
        bool ok = true;
        int pc = 0;
        for (int st = pair; st < a.n_tiles; st += n_pairs) {
            for (int pass = 0; pass < 2; pass++, pc++) {
                const int acc = pc & 1;
                const int q0 = st * 512 + (int)rank * 256 + et;
                if (a.skip) {                                            // idle until the MMAs finish: pull the skip rows towards L2
#pragma unroll
                    for (int j = 0; j < 2; j++)
                        if (q0 + j * 128 < a.Q) {
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.skip + (size_t)(q0 + j * 128) * TW_C + pass * 128));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(a.skip + (size_t)(q0 + j * 128) * TW_C + pass * 128 + 64));
                        }
                }
                if (ok) ok = mbar_wait(&tail->tmem_full[acc], (pc >> 1) & 1, a.err);
                ok = __all_sync(SGO_FULL, ok);
                if (!ok) break;
                tc_fence_after();
#pragma unroll 1
                for (int j = 0; j < 2; j++) {
                    const int q = q0 + j * 128;
                    const bool valid = q < a.Q;                          // (only the last super-tile has rows past the end)
                    const size_t gofs = (size_t)q * TW_C + pass * 128;
                    float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f;
                    const uint32_t t_addr = tmem_base + ((uint32_t)(qw * 32) << 16) + acc * 256 + j * 128;
#pragma unroll 1
                    for (int c = 0; c < 4; c++) {
                        uint32_t sk[16];
                        const bool do_skip = valid && a.skip;
                        if (do_skip) {
                            ldg256(a.skip + gofs + c * 32, sk);
                            ldg256(a.skip + gofs + c * 32 + 16, sk + 8);
                        }
                        uint32_t v[32];
                        tmem_ld32(t_addr + c * 32, v);
                        if (valid) {
                            uint32_t ow[16];
                            const int ch = pass * 128 + c * 32;
#pragma unroll
                            for (int i = 0; i < 16; i++) {
                                float f0 = __uint_as_float(v[2 * i]) + tail->bias[ch + 2 * i];
                                float f1 = __uint_as_float(v[2 * i + 1]) + tail->bias[ch + 2 * i + 1];
                                if (a.skip) {
                                    __nv_bfloat162 s2 = *reinterpret_cast<const __nv_bfloat162 *>(&sk[i]);
                                    f0 += __bfloat162float(s2.x);
                                    f1 += __bfloat162float(s2.y);
                                }
                                if (a.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
                                __nv_bfloat162 p = __floats2bfloat162_rn(f0, f1);
                                ow[i] = *reinterpret_cast<uint32_t *>(&p);
                                if (a.head_w4) {
                                    float4 wa = tail->w4[ch + 2 * i], wb = tail->w4[ch + 2 * i + 1];
                                    h0 = fmaf(f0, wa.x, h0); h1 = fmaf(f0, wa.y, h1); h2 = fmaf(f0, wa.z, h2); h3 = fmaf(f0, wa.w, h3);
                                    h0 = fmaf(f1, wb.x, h0); h1 = fmaf(f1, wb.y, h1); h2 = fmaf(f1, wb.z, h2); h3 = fmaf(f1, wb.w, h3);
                                }
                            }
                            if (a.out) {
                                stg256(a.out + gofs + c * 32, ow);
                                stg256(a.out + gofs + c * 32 + 16, ow + 8);
                            }
                        }
                    }
                    if (a.head_w4) {                                     // 1x1 head convs: channels 0-127 in pass 0, 128-255 in pass 1
                        if (pass == 0) tail->hsum[j][et] = make_float4(h0, h1, h2, h3);
                        else if (valid) {
                            const float4 p0 = tail->hsum[j][et];
                            const int pos = q / a.PX, pix = q - pos * a.PX;
                            const size_t fo = (size_t)pos * a.feat_ld + (size_t)pix * 2;
                            *reinterpret_cast<float2 *>(a.featp + fo) = make_float2(fmaxf(p0.x + h0 + a.head_b4[0], 0.f), fmaxf(p0.y + h1 + a.head_b4[1], 0.f));
                            *reinterpret_cast<float2 *>(a.featv + fo) = make_float2(fmaxf(p0.z + h2 + a.head_b4[2], 0.f), fmaxf(p0.w + h3 + a.head_b4[3], 0.f));
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive_leader(&tail->tmem_empty[acc]);
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
