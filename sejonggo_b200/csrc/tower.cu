// tower.cu — the residual policy/value tower of model.py:55-96 on sm_100a.
//
//   stem   Conv3x3 "valid" 17->C on the SxS input planes + BN + ReLU (model.py:57-61) as ONE
//          tensor-core GEMM (K = 153 -> 192) whose A tiles are the im2col of the packed bitboards,
//          built in shared memory by the kernel's producer warps (stem_fused.cuh; input-plane
//          construction play.py:295-299 and the symmetry gather symmetry.py:45-114 fused in; values
//          0/1/+-1 are exact in bf16).  Block-less towers (tests) use k_stem_im2col + the pair kernel.
//   tower  2*N_RESIDUAL_BLOCKS Conv3x3 "same" C->C (+BN folded, +skip, ReLU): implicit GEMMs on
//          CTA pairs — TMA boxes per filter tap, tcgen05.mma.cta_group::2, double-buffered TMEM
//          accumulators (conv_pair.cuh).
//   heads  the two 1x1 convs are fused into the last conv's epilogue (its activations never
//          reach HBM) and leave the two fp32 feature matrices; the dense layers Dense(578->362) and
//          Dense(578->256) (model.py:80,90) are TF32 GEMMs on the same CTA-pair pipeline (conv_pair.cuh
//          MODE 1); k_heads_finish does softmax, the "reverse" symmetry policy gather, Dense(256->1) and tanh.
//
// Activation layout in HBM: bf16 [Q][C], dense, q = (position * W + y) * W + x with W = S-2 (Q11).  The conv is a shifted
// GEMM over q with the image edges handled by tcgen05's disable-output-lane masks (conv_pair.cuh).  History of the
// layout (profiles/): row tiles with per-tap TMA boxes (r01_tower_bench_pair*.json), real-row tiles
// (r01_conv_real_row_tiling_ab.json) and a dense 4-D layout (r01_experiment_dense_*.json) all re-loaded the A operand
// once per tap and were bound by L2->SM bytes; round 1 ended on the shifted GEMM over a zero-PADDED q (18 x 18 slots per
// 17 x 17 image: 10.8% of the MMAs on zeros, r02_conv_pad_bound.json); the lane masks remove the padding.
#include "engine.h"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstring>
#include <cstdlib>
#include <new>

#define TW_C 256
#define TW_KCH 64
#define TW_THREADS 192
#define TW_SPIN_LIMIT (1u << 21)
#define TW_STEM_K 192                      // 9 taps x 17 planes = 153, padded to 3 x 64

struct sgo_tower {
    int n_blocks, S, W, PX, max_pos;         // PX = W*W pixels per position
    int n_layers;
    float *stem_w, *stem_b;                 // [9][17][C], [C]
    __nv_bfloat16 *conv_w;                  // [n_layers][9][C co][C ci]
    float *conv_b;                          // [n_layers][C]
    float *pol_conv_w, *pol_conv_b, *pol_fc_w, *pol_fc_b;
    float *val_conv_w, *val_conv_b, *val_fc1_w, *val_fc1_b, *val_fc2_w, *val_fc2_b;
    __nv_bfloat16 *act[3];                  // [max_pos*PX][C]
    uint32_t *lane_masks;                   // [PX][PR_MASK_WORDS] disable-output-lane masks per tile alignment (conv_pair.cuh)
    uint32_t *lane_masks_wide;              // [PX][WD_MASK_WORDS] the same per 512-row super-tile alignment (conv_wide.cuh)
    struct WideMaps *wmaps;                 // [0..2] tensor maps of k_conv3x3_wide per activation buffer
    struct PairMaps *pmaps;                 // [0..2] tensor maps of the CTA-pair kernel per activation buffer (conv_pair.cuh); [3] = stem im2col;
                                            // [4], [5] = dense heads: policy / value feature matrix + the transposed dense weights (fp32)
    __nv_bfloat16 *stem_col;                // [max_pos*PX][TW_STEM_K] im2col of the input planes (0/1, +-1)
    __nv_bfloat16 *stem_wb;                 // [C co][TW_STEM_K] bf16 stem weights, k = tap*17 + plane
    float *head_w4, *head_b4;               // fused 1x1 head convs: [C][4], [4]
    float *featp, *featv;                   // their post-ReLU outputs = A matrices of the dense heads: fp32 [max_pos][feat_ld], col = pix*2 + ch
    int feat_ld;                            // 2*W*W rounded up to 32 (TF32 K chunks of 128 bytes); the pad columns stay zero
    int pol_tiles;                          // 256-column N tiles of the policy layer: ceil(A / 256)
    float *head_wt;                         // [(pol_tiles + 1) * 256][feat_ld] transposed dense weights: policy outputs, then the 256 value hidden units
    float *head_bias;                       // [(pol_tiles + 1) * 256]
    float *hbuf;                            // [max_pos][(pol_tiles + 1) * 256] policy logits | value hidden (post ReLU)
    int32_t *err;
    int sm_count;
    // optional live profiling (bench.py roofline): 4 events per forward call
    int prof_on, prof_n;
    cudaEvent_t *prof_ev;          // [TW_PROF_MAX][4]: start, after stem, after convs, after heads
    int *prof_pos;                 // positions in each profiled forward
};
#define TW_PROF_MAX 4096

static inline cudaStream_t S_(void *s) { return (cudaStream_t)s; }

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug sets an error flag instead of hanging the GPU
__device__ __forceinline__ bool mbar_wait(uint64_t *bar, uint32_t parity, int32_t *err)
{
    uint32_t addr = smem_u32(bar);
    for (uint32_t i = 0; i < TW_SPIN_LIMIT; i++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return true;
    }
    atomicOr(err, 16);
    return false;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t v[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one full 32-B sector per lane per instruction
__device__ __forceinline__ void ldg256(const void *p, uint32_t v[8])
{
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(void *p, const uint32_t v[8])
{
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

#include "conv_pair.cuh"
#include "conv_wide.cuh"

// ------------------------------------------------------------------ stem
__device__ __forceinline__ void sym_src_t(int S, int sym, int y, int x, int &sy, int &sx)
{
    int m = S - 1;
    switch (sym) {
    case 0: sy = y; sx = x; break;
    case 1: sy = x; sx = y; break;
    case 2: sy = y; sx = m - x; break;
    case 3: sy = m - y; sx = x; break;
    case 4: sy = x; sx = m - y; break;
    case 5: sy = m - y; sx = m - x; break;
    case 6: sy = m - x; sx = y; break;
    default: sy = m - x; sx = m - y; break;
    }
}

// The stem's im2col as a tensor in HBM (block-less towers and the -DSGO_STEM_SEPARATE baseline; the product path builds
// the same rows in shared memory, stem_fused.cuh): the (symmetry-transformed) input planes straight from
// the packed bitboards: col[row(y), x][tap*17 + p] = plane p at (y+ky, x+kx)  (valid conv, Q11).
// Every value is 0, 1 or +-1, exact in bf16.  A block takes IM_NP positions at a time: first the
// 16 stone planes of every cell as one 16-bit mask in shared memory, then one thread per output
// pixel writes the pixel's whole 384-B row — the 192 k's are unrolled, so tap and plane are
// compile-time constants (3 ALU ops per element) and every store is a full 32-B sector.
#define IM_NP 8
#define IM_THREADS 256

template <int SEC>
__device__ __forceinline__ void im2col_sector(const uint32_t (&m)[9], uint32_t tmv, __nv_bfloat16 *dst)
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
    stg256(dst + SEC * 16, w);
}

__global__ void __launch_bounds__(IM_THREADS)
k_stem_im2col(const Board *boards, const int32_t *index, const int32_t *syms, int n, int S, __nv_bfloat16 *col /* [n*W*W][TW_STEM_K] */)
{
    __shared__ uint16_t cell[IM_NP][SGO_MAXS * SGO_MAXS];
    __shared__ int s_tm[IM_NP];
    const int W = S - 2, PX = W * W;
    for (int i0 = blockIdx.x * IM_NP; i0 < n; i0 += gridDim.x * IM_NP) {
        const int np = min(IM_NP, n - i0);
        __syncthreads();
        for (int c = threadIdx.x; c < np * S * S; c += blockDim.x) {
            const int q = c / (S * S), cc = c - q * S * S;
            const int src = index ? index[i0 + q] : i0 + q;
            uint32_t m = 0;
            if (src >= 0) {
                const Board *bd = boards + src;
                const int sym = syms ? (syms[i0 + q] & 7) : 0;
                const int tm = bd->to_move, head = bd->head;
                int y = cc / S, x = cc - y * S, sy, sx;
                sym_src_t(S, sym, y, x, sy, sx);
#pragma unroll
                for (int k = 0; k < SGO_HIST; k++) {
                    int slot = (head + SGO_HIST - k) & (SGO_HIST - 1);
                    uint32_t bl = (bd->st[slot][0][sy] >> sx) & 1u, wh = (bd->st[slot][1][sy] >> sx) & 1u;
                    uint32_t own = tm == 1 ? bl : wh, opp = tm == 1 ? wh : bl;
                    m |= (own << (2 * k)) | (opp << (2 * k + 1));
                }
                if (cc == 0) s_tm[q] = tm;
            } else if (cc == 0) s_tm[q] = 0;
            cell[q][cc] = (uint16_t)m;
        }
        __syncthreads();
        for (int it = threadIdx.x; it < np * PX; it += blockDim.x) {
            const int q = it / PX, px = it - q * PX;
            const int src = index ? index[i0 + q] : i0 + q;
            if (src < 0) continue;
            const int y = px / W, x = px - y * W;
            uint32_t m[9];
#pragma unroll
            for (int tap = 0; tap < 9; tap++) m[tap] = cell[q][(y + tap / 3) * S + x + tap % 3];
            const uint32_t tmv = s_tm[q] == 1 ? 0x3F80u : 0xBF80u;           // bf16 +1 / -1 (plane 16)
            __nv_bfloat16 *dst = col + ((size_t)(i0 + q) * PX + px) * TW_STEM_K;
            im2col_sector<0>(m, tmv, dst); im2col_sector<1>(m, tmv, dst); im2col_sector<2>(m, tmv, dst);
            im2col_sector<3>(m, tmv, dst); im2col_sector<4>(m, tmv, dst); im2col_sector<5>(m, tmv, dst);
            im2col_sector<6>(m, tmv, dst); im2col_sector<7>(m, tmv, dst); im2col_sector<8>(m, tmv, dst);
            im2col_sector<9>(m, tmv, dst); im2col_sector<10>(m, tmv, dst); im2col_sector<11>(m, tmv, dst);
        }
    }
}

#include "stem_fused.cuh"

// disable-output-lane masks of the 3x3 "same" convolution over the dense q layout (conv_pair.cuh): for a pair tile whose first
// row is pixel `al` of a position, bit i of word w of tap-slot t is set when row 32w + i must NOT receive tap tap_of(t),
// i.e. when its (dy, dx) neighbour lies off the W x W board.  One thread per (alignment, tap slot, word).
__global__ void k_lane_masks(int W, uint32_t *masks)
{
    const int PX = W * W;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= PX * PR_MASK_WORDS) return;
    const int al = i / PR_MASK_WORDS, r = i - al * PR_MASK_WORDS, t = r >> 3, w = r & 7;
    const int tp = t == 0 ? 4 : (t <= 4 ? t - 1 : t);        // tap_of(t)
    const int dy = tp / 3 - 1, dx = tp % 3 - 1;
    uint32_t m = 0;
    for (int b = 0; b < 32; b++) {
        const int pix = (al + 32 * w + b) % PX, y = pix / W, x = pix - y * W;
        if (x + dx < 0 || x + dx >= W || y + dy < 0 || y + dy >= W) m |= 1u << b;
    }
    masks[i] = m;
}

// the same for the 512-row super-tiles of conv_wide.cuh: [alignment][row block j][tap slot][8 words]; words 0-3 are the
// leader CTA's rows (super-tile rows 128j .. 128j+127), words 4-7 its peer's (256 + 128j ..)
__global__ void k_lane_masks_wide(int W, uint32_t *masks)
{
    const int PX = W * W;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= PX * WD_MASK_WORDS) return;
    const int al = i / WD_MASK_WORDS, r = i - al * WD_MASK_WORDS, j = r / 72, t = (r % 72) >> 3, w = r & 7;
    const int tp = t == 0 ? 4 : (t <= 4 ? t - 1 : t);        // tap_of(t)
    const int dy = tp / 3 - 1, dx = tp % 3 - 1;
    const int row0 = (w < 4 ? 0 : 256) + 128 * j + 32 * (w & 3);
    uint32_t m = 0;
    for (int b = 0; b < 32; b++) {
        const int pix = (al + row0 + b) % PX, y = pix / W, x = pix - y * W;
        if (x + dx < 0 || x + dx >= W || y + dy < 0 || y + dy >= W) m |= 1u << b;
    }
    masks[i] = m;
}

// 1x1 head conv weights [C][2] + [C][2] -> interleaved [C][4] (p0,p1,v0,v1) and biases [4]
__global__ void k_head_w4(const float *pcw, const float *pcb, const float *vcw, const float *vcb, float *w4, float *b4)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < TW_C) {
        w4[c * 4 + 0] = pcw[c * 2]; w4[c * 4 + 1] = pcw[c * 2 + 1];
        w4[c * 4 + 2] = vcw[c * 2]; w4[c * 4 + 3] = vcw[c * 2 + 1];
    }
    if (c < 2) { b4[c] = pcb[c]; b4[2 + c] = vcb[c]; }
}

// fp32 [9][17][C] stem weights (BN folded) -> bf16 [C][TW_STEM_K]
__global__ void k_stem_weights(const float *w, __nv_bfloat16 *wb)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TW_C * TW_STEM_K) return;
    int co = i / TW_STEM_K, k = i - co * TW_STEM_K;
    wb[i] = __float2bfloat16(k < 153 ? w[(size_t)k * TW_C + co] : 0.f);
}

// ------------------------------------------------------------------ heads
// What is left after the two dense GEMMs (hbuf[pos] = policy logits [0, A) | value hidden units, ReLU'd, at column
// vcol): softmax over the A logits, the "reverse" symmetry gather of the policy (symmetry.py:90-114; it re-uses the
// forward map, Q8), Dense(256 -> 1) + tanh (model.py:91-92).  One warp per position; ~2.5 KB read, 1.5 KB written.
#define HF_WARPS 8

struct FinishArgs {
    int n, S, A, ld, vcol, scatter;
    const int32_t *index, *syms;
    const float *hbuf, *v2w, *v2b;
    float *policy, *value;
};

__global__ void __launch_bounds__(HF_WARPS * 32)
k_heads_finish(FinishArgs h)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * HF_WARPS + (threadIdx.x >> 5);
    if (i >= h.n) return;
    const float *row = h.hbuf + (size_t)i * h.ld;
    float lg[SGO_AWORDS];
    float mx = -3.4e38f;
#pragma unroll
    for (int it = 0; it < SGO_AWORDS; it++) {
        const int a0 = it * 32 + lane;
        lg[it] = a0 < h.A ? row[a0] : -3.4e38f;
        mx = fmaxf(mx, lg[it]);
    }
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(SGO_FULL, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int it = 0; it < SGO_AWORDS; it++) {
        const int a0 = it * 32 + lane;
        lg[it] = a0 < h.A ? expf(lg[it] - mx) : 0.f;
        sum += lg[it];
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(SGO_FULL, sum, o);
    const float inv = 1.f / sum;
    const size_t orow = h.scatter ? (size_t)(h.index ? h.index[i] : i) : (size_t)i;
    const int sym = h.syms ? (h.syms[i] & 7) : 0;
    float *pout = h.policy + orow * h.A;
    if (sym == 0) {
#pragma unroll
        for (int it = 0; it < SGO_AWORDS; it++) {
            const int a0 = it * 32 + lane;
            if (a0 < h.A) pout[a0] = lg[it] * inv;
        }
    } else {
        // out[a] = softmax[g(a)]: the value sits in lane g(a) % 32, register g(a) / 32 -> go through shared memory
        __shared__ float sp[HF_WARPS][SGO_APAD];
        float *mine = sp[threadIdx.x >> 5];
#pragma unroll
        for (int it = 0; it < SGO_AWORDS; it++) mine[it * 32 + lane] = lg[it] * inv;
        __syncwarp();
        for (int a0 = lane; a0 < h.A; a0 += 32) {
            int src = a0;
            if (a0 < h.S * h.S) {
                int y = a0 / h.S, x = a0 - y * h.S, sy, sx;
                sym_src_t(h.S, sym, y, x, sy, sx);
                src = sy * h.S + sx;
            }
            pout[a0] = mine[src];
        }
    }
    float vs = 0.f;
    for (int k = lane; k < 256; k += 32) vs = fmaf(row[h.vcol + k], h.v2w[k], vs);
    for (int o = 16; o; o >>= 1) vs += __shfl_xor_sync(SGO_FULL, vs, o);
    if (lane == 0) h.value[orow] = tanhf(vs + h.v2b[0]);
}

// dense weights [F][A] and [F][256] (Keras: y = x @ W) -> one transposed, padded fp32 matrix [(pol_tiles+1)*256][ld]
// (row = output unit, K-major) + the padded bias vector
__global__ void k_head_wt(const float *pfw, const float *pfb, const float *v1w, const float *v1b, int F, int A, int ld, int pol_tiles,
                          float *wt, float *bias)
{
    const int rows = (pol_tiles + 1) * 256;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * ld) return;
    const int r = (int)(i / ld), k = (int)(i - (size_t)r * ld);
    float v = 0.f;
    if (k < F) {
        if (r < pol_tiles * 256) { if (r < A) v = pfw[(size_t)k * A + r]; }
        else v = v1w[(size_t)k * 256 + (r - pol_tiles * 256)];
    }
    wt[i] = v;
    if (k == 0) bias[r] = r < pol_tiles * 256 ? (r < A ? pfb[r] : 0.f) : v1b[r - pol_tiles * 256];
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

static int make_maps(sgo_engine *e, sgo_tower *t)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) return sgo_fail(e, "cuTensorMapEncodeTiled unavailable");
    cuuint32_t es[2] = {1, 1};
    t->pmaps = new PairMaps[6];
    memset(t->pmaps, 0, sizeof(PairMaps) * 6);
    const cuuint64_t Q = (cuuint64_t)t->max_pos * t->PX;
    for (int i = 0; i < 4; i++) {                          // activation buffers [0..2], stem im2col [3]
        const int kw = i < 3 ? TW_C : TW_STEM_K, halo = i < 3 ? t->W + 1 : 0;
        cuuint64_t dims[2] = {(cuuint64_t)kw, Q};
        cuuint64_t strides[1] = {(cuuint64_t)kw * 2};
        cuuint32_t box[2] = {TW_KCH, (cuuint32_t)(128 + 2 * halo)};      // one CTA's 128 rows + the (dy, dx) halo on both sides
        CUresult rr = enc(&t->pmaps[i].act, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, i < 3 ? (void *)t->act[i] : (void *)t->stem_col, dims, strides,
                          box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rr != CUDA_SUCCESS) return sgo_fail(e, "cuTensorMapEncodeTiled(activation slab) failed");
        cuuint64_t wdims[2] = {(cuuint64_t)kw, i < 3 ? (cuuint64_t)(t->n_layers ? t->n_layers : 1) * 9 * TW_C : (cuuint64_t)TW_C};
        cuuint32_t boxw[2] = {TW_KCH, 128};
        rr = enc(&t->pmaps[i].w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, i < 3 ? (void *)t->conv_w : (void *)t->stem_wb, wdims, strides, boxw, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rr != CUDA_SUCCESS) return sgo_fail(e, "cuTensorMapEncodeTiled(weights) failed");
    }
    t->wmaps = new WideMaps[3];
    memset(t->wmaps, 0, sizeof(WideMaps) * 3);
    for (int i = 0; i < 3; i++) {                          // k_conv3x3_wide: half-slab activation boxes, 64 x 64 weight boxes
        cuuint64_t dims[2] = {(cuuint64_t)TW_C, Q};
        cuuint64_t strides[1] = {(cuuint64_t)TW_C * 2};
        cuuint32_t box[2] = {TW_KCH, (cuuint32_t)(128 + t->W + 1)};
        CUresult rr = enc(&t->wmaps[i].act, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)t->act[i], dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rr != CUDA_SUCCESS) return sgo_fail(e, "cuTensorMapEncodeTiled(wide activation box) failed");
        cuuint64_t wdims[2] = {(cuuint64_t)TW_C, (cuuint64_t)(t->n_layers ? t->n_layers : 1) * 9 * TW_C};
        cuuint32_t boxw[2] = {TW_KCH, 64};
        rr = enc(&t->wmaps[i].w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)t->conv_w, wdims, strides, boxw, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rr != CUDA_SUCCESS) return sgo_fail(e, "cuTensorMapEncodeTiled(wide weights) failed");
    }
    for (int i = 4; i < 6; i++) {                          // dense heads: fp32, 32 elements = 128 bytes per K chunk
        cuuint64_t dims[2] = {(cuuint64_t)t->feat_ld, (cuuint64_t)t->max_pos};
        cuuint64_t strides[1] = {(cuuint64_t)t->feat_ld * 4};
        cuuint32_t box[2] = {TW_KCH / 2, 128};
        CUresult rr = enc(&t->pmaps[i].act, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, i == 4 ? (void *)t->featp : (void *)t->featv, dims, strides,
                          box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rr != CUDA_SUCCESS) return sgo_fail(e, "cuTensorMapEncodeTiled(head features) failed");
        cuuint64_t wdims[2] = {(cuuint64_t)t->feat_ld, (cuuint64_t)(t->pol_tiles + 1) * 256};
        rr = enc(&t->pmaps[i].w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)t->head_wt, wdims, strides, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rr != CUDA_SUCCESS) return sgo_fail(e, "cuTensorMapEncodeTiled(head weights) failed");
    }
    return 0;
}

static void tower_free(sgo_tower *t)
{
    if (!t) return;
    cudaFree(t->stem_w); cudaFree(t->stem_b); cudaFree(t->conv_w); cudaFree(t->conv_b);
    cudaFree(t->pol_conv_w); cudaFree(t->pol_conv_b); cudaFree(t->pol_fc_w); cudaFree(t->pol_fc_b);
    cudaFree(t->val_conv_w); cudaFree(t->val_conv_b); cudaFree(t->val_fc1_w); cudaFree(t->val_fc1_b);
    cudaFree(t->val_fc2_w); cudaFree(t->val_fc2_b);
    for (int i = 0; i < 3; i++) cudaFree(t->act[i]);
    cudaFree(t->err); cudaFree(t->lane_masks); cudaFree(t->lane_masks_wide);
    delete[] t->wmaps;
    cudaFree(t->stem_col); cudaFree(t->stem_wb); cudaFree(t->head_w4); cudaFree(t->head_b4);
    cudaFree(t->featp); cudaFree(t->featv); cudaFree(t->head_wt); cudaFree(t->head_bias); cudaFree(t->hbuf);
    delete[] t->pmaps;
    if (t->prof_ev) {
        for (int i = 0; i < TW_PROF_MAX * 4; i++) cudaEventDestroy(t->prof_ev[i]);
        delete[] t->prof_ev;
        delete[] t->prof_pos;
    }
    delete t;
}

extern "C" int sgo_tower_free(sgo_engine *e, int32_t slot)
{
    if (slot < 0 || slot > 1) return sgo_fail(e, "tower slot must be 0 or 1");
    tower_free(e->tower[slot]);
    e->tower[slot] = nullptr;
    return 0;
}

#define DUP(dst, src, count, type)                                                                     \
    do {                                                                                               \
        SGO_CUDA_OK(e, cudaMalloc(&(dst), sizeof(type) * (size_t)(count)));                            \
        SGO_CUDA_OK(e, cudaMemcpyAsync((dst), (src), sizeof(type) * (size_t)(count), cudaMemcpyDeviceToDevice, S_(stream))); \
    } while (0)

extern "C" int sgo_tower_load_weights(sgo_engine *e, int32_t slot, const sgo_tower_weights *w, int32_t max_positions, void *stream)
{
    if (slot < 0 || slot > 1) return sgo_fail(e, "tower slot must be 0 or 1");
    if (w->channels != TW_C) return sgo_fail(e, "tower kernels are built for 256 channels");
    if (w->size != e->S || e->S < 5) return sgo_fail(e, "tower board size mismatch (needs S >= 5)");
    if (max_positions < 1) return sgo_fail(e, "max_positions must be positive");
    sgo_tower_free(e, slot);
    sgo_tower *t = new (std::nothrow) sgo_tower();
    if (!t) return sgo_fail(e, "out of host memory");
    memset(t, 0, sizeof(*t));
    e->tower[slot] = t;
    t->n_blocks = w->n_blocks; t->n_layers = 2 * w->n_blocks; t->S = e->S; t->W = e->S - 2;
    t->PX = t->W * t->W;
    if (t->W + 1 > PR_MAX_HALO) return sgo_fail(e, "board too wide for the conv slab");
    t->max_pos = max_positions;
    int P = t->W * t->W, A = e->A;
    DUP(t->stem_w, w->stem_w, 9 * 17 * TW_C, float);
    DUP(t->stem_b, w->stem_b, TW_C, float);
    DUP(t->conv_w, (const __nv_bfloat16 *)w->conv_w, (size_t)t->n_layers * 9 * TW_C * TW_C, __nv_bfloat16);
    DUP(t->conv_b, w->conv_b, t->n_layers * TW_C, float);
    DUP(t->pol_conv_w, w->pol_conv_w, TW_C * 2, float);
    DUP(t->pol_conv_b, w->pol_conv_b, 2, float);
    DUP(t->pol_fc_w, w->pol_fc_w, (size_t)2 * P * A, float);
    DUP(t->pol_fc_b, w->pol_fc_b, A, float);
    DUP(t->val_conv_w, w->val_conv_w, TW_C * 2, float);
    DUP(t->val_conv_b, w->val_conv_b, 2, float);
    DUP(t->val_fc1_w, w->val_fc1_w, (size_t)2 * P * 256, float);
    DUP(t->val_fc1_b, w->val_fc1_b, 256, float);
    DUP(t->val_fc2_w, w->val_fc2_w, 256, float);
    DUP(t->val_fc2_b, w->val_fc2_b, 1, float);
    size_t act_bytes = (size_t)max_positions * t->PX * TW_C * sizeof(__nv_bfloat16);
    for (int i = 0; i < 3; i++) {
        SGO_CUDA_OK(e, cudaMalloc(&t->act[i], act_bytes));
        SGO_CUDA_OK(e, cudaMemsetAsync(t->act[i], 0, act_bytes, S_(stream)));
    }
    SGO_CUDA_OK(e, cudaMalloc(&t->lane_masks, sizeof(uint32_t) * (size_t)t->PX * PR_MASK_WORDS));
    k_lane_masks<<<(t->PX * PR_MASK_WORDS + 255) / 256, 256, 0, S_(stream)>>>(t->W, t->lane_masks);
    SGO_CUDA_OK(e, cudaGetLastError());
    SGO_CUDA_OK(e, cudaMalloc(&t->lane_masks_wide, sizeof(uint32_t) * (size_t)t->PX * WD_MASK_WORDS));
    k_lane_masks_wide<<<(t->PX * WD_MASK_WORDS + 255) / 256, 256, 0, S_(stream)>>>(t->W, t->lane_masks_wide);
    SGO_CUDA_OK(e, cudaGetLastError());
    size_t col_bytes = (size_t)max_positions * t->PX * TW_STEM_K * sizeof(__nv_bfloat16);
    SGO_CUDA_OK(e, cudaMalloc(&t->stem_col, col_bytes));
    SGO_CUDA_OK(e, cudaMemsetAsync(t->stem_col, 0, col_bytes, S_(stream)));
    SGO_CUDA_OK(e, cudaMalloc(&t->stem_wb, sizeof(__nv_bfloat16) * TW_C * TW_STEM_K));
    k_stem_weights<<<(TW_C * TW_STEM_K + 255) / 256, 256, 0, S_(stream)>>>(t->stem_w, t->stem_wb);
    SGO_CUDA_OK(e, cudaGetLastError());
    SGO_CUDA_OK(e, cudaMalloc(&t->head_w4, sizeof(float) * TW_C * 4));
    SGO_CUDA_OK(e, cudaMalloc(&t->head_b4, sizeof(float) * 4));
    k_head_w4<<<1, TW_C, 0, S_(stream)>>>(t->pol_conv_w, t->pol_conv_b, t->val_conv_w, t->val_conv_b, t->head_w4, t->head_b4);
    SGO_CUDA_OK(e, cudaGetLastError());
    // dense heads as TF32 GEMMs: feature matrices (pad columns zero for ever), transposed weights, output rows
    t->feat_ld = (2 * P + 31) / 32 * 32;
    t->pol_tiles = (A + 255) / 256;
    {
        const size_t fb = sizeof(float) * (size_t)t->feat_ld * max_positions, rows = (size_t)(t->pol_tiles + 1) * 256;
        SGO_CUDA_OK(e, cudaMalloc(&t->featp, fb));
        SGO_CUDA_OK(e, cudaMalloc(&t->featv, fb));
        SGO_CUDA_OK(e, cudaMemsetAsync(t->featp, 0, fb, S_(stream)));
        SGO_CUDA_OK(e, cudaMemsetAsync(t->featv, 0, fb, S_(stream)));
        SGO_CUDA_OK(e, cudaMalloc(&t->head_wt, sizeof(float) * rows * t->feat_ld));
        SGO_CUDA_OK(e, cudaMalloc(&t->head_bias, sizeof(float) * rows));
        SGO_CUDA_OK(e, cudaMalloc(&t->hbuf, sizeof(float) * rows * max_positions));
        const size_t tot = rows * t->feat_ld;
        k_head_wt<<<(unsigned)((tot + 255) / 256), 256, 0, S_(stream)>>>(t->pol_fc_w, t->pol_fc_b, t->val_fc1_w, t->val_fc1_b, 2 * P, A,
                                                                        t->feat_ld, t->pol_tiles, t->head_wt, t->head_bias);
        SGO_CUDA_OK(e, cudaGetLastError());
    }
    SGO_CUDA_OK(e, cudaMalloc(&t->err, sizeof(int32_t)));
    SGO_CUDA_OK(e, cudaMemsetAsync(t->err, 0, sizeof(int32_t), S_(stream)));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, dev);
    SGO_CUDA_OK(e, cudaFuncSetAttribute(k_conv3x3_pair<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_SMEM_BYTES));
    SGO_CUDA_OK(e, cudaFuncSetAttribute(k_conv3x3_pair<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_SMEM_BYTES));
    SGO_CUDA_OK(e, cudaFuncSetAttribute(k_conv3x3_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, WD_SMEM_BYTES));
    SGO_CUDA_OK(e, cudaFuncSetAttribute(k_stem_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, SF_SMEM_BYTES));
    int rc = make_maps(e, t);
    if (rc) return rc;
    SGO_CUDA_OK(e, cudaStreamSynchronize(S_(stream)));
    return 0;
}

extern "C" int sgo_tower_max_positions(sgo_engine *e, int32_t slot)
{
    if (slot < 0 || slot > 1 || !e->tower[slot]) return 0;
    return e->tower[slot]->max_pos;
}

// one launch of the pair kernel: a tower conv layer (layer >= 0) or the stem GEMM (layer < 0)
static int launch_conv(sgo_engine *e, sgo_tower *t, int n, int layer, int in, int out, int skip, void *stream, bool heads = false)
{
    PairArgs pa;
    pa.W = t->W; pa.PX = t->PX; pa.Q = n * t->PX;
#ifdef SGO_CONV_ABLATE
    { const char *d = getenv("SGO_CONV_DEBUG"); pa.dbg = d ? atoi(d) : 0; }
#else
    pa.dbg = 0;
#endif
    pa.n_tiles = (pa.Q + 255) / 256;
    pa.relu = 1; pa.err = t->err;
    if (layer >= 0) {
        pa.n_taps = 9; pa.kchunks = TW_C / TW_KCH; pa.w_row0 = layer * 9 * TW_C; pa.halo = t->W + 1;
        pa.bias = t->conv_b + (size_t)layer * TW_C; pa.masks = t->lane_masks;
    } else {
        pa.n_taps = 1; pa.kchunks = TW_STEM_K / TW_KCH; pa.w_row0 = 0; pa.halo = 0;
        pa.bias = t->stem_b; pa.masks = nullptr;
    }
    pa.skip = skip >= 0 ? t->act[skip] : nullptr;
    pa.out = heads ? nullptr : t->act[out];      // the last layer's activations are consumed by the fused 1x1 head convs only
    pa.head_w4 = heads ? t->head_w4 : nullptr; pa.head_b4 = t->head_b4;
    pa.featp = t->featp; pa.featv = t->featv; pa.feat_ld = t->feat_ld;
    pa.outf = nullptr; pa.ldo = pa.ncols = pa.n_valid = 0;
    int pairs = t->sm_count / 2;
#ifdef SGO_CONV_WIDE_TILES
    if (layer >= 0) {                           // A/B only (profiles/r02_conv_wide_tiles_ab.json): 512-row super-tiles, weight traffic halved (conv_wide.cuh)
        pa.n_tiles = (pa.Q + 511) / 512;
        pa.masks = t->lane_masks_wide;
        if (pairs > pa.n_tiles) pairs = pa.n_tiles;
        k_conv3x3_wide<<<2 * pairs, TW_THREADS, WD_SMEM_BYTES, S_(stream)>>>(t->wmaps[in], pa);
        SGO_LAUNCHED(e);
        return 0;
    }
#endif
    if (pairs > pa.n_tiles) pairs = pa.n_tiles;
    k_conv3x3_pair<0><<<2 * pairs, TW_THREADS, PR_SMEM_BYTES, S_(stream)>>>(t->pmaps[layer >= 0 ? in : 3], pa);
    SGO_LAUNCHED(e);
    return 0;
}

// the dense heads of n positions: (pol_tiles + 1) launches of the pair kernel in MODE 1 (one 256-column N tile each:
// the policy outputs, then the value hidden layer), then k_heads_finish
static int launch_heads(sgo_engine *e, sgo_tower *t, int n, const int32_t *d_index, const int32_t *d_sym, int scatter,
                        float *d_policy, float *d_value, void *stream)
{
    const int ld = (t->pol_tiles + 1) * 256;
    for (int nt = 0; nt <= t->pol_tiles; nt++) {
        PairArgs pa;
        memset(&pa, 0, sizeof(pa));
        pa.W = t->W; pa.PX = t->PX; pa.Q = n;
        pa.n_tiles = (n + 255) / 256;
        pa.n_taps = 1; pa.kchunks = t->feat_ld / 32; pa.halo = 0; pa.w_row0 = nt * 256;
        pa.relu = nt == t->pol_tiles;                         // Dense(256, relu) of the value head; the policy logits go to the softmax raw
        pa.bias = t->head_bias + nt * 256;
        pa.err = t->err;
        pa.outf = t->hbuf + nt * 256; pa.ldo = ld; pa.n_valid = n;
        pa.ncols = nt < t->pol_tiles ? (e->A - nt * 256 < 256 ? e->A - nt * 256 : 256) : 256;
        int pairs = t->sm_count / 2;
        if (pairs > pa.n_tiles) pairs = pa.n_tiles;
        k_conv3x3_pair<1><<<2 * pairs, TW_THREADS, PR_SMEM_BYTES, S_(stream)>>>(t->pmaps[nt < t->pol_tiles ? 4 : 5], pa);
        SGO_LAUNCHED(e);
    }
    FinishArgs f;
    f.n = n; f.S = t->S; f.A = e->A; f.ld = ld; f.vcol = t->pol_tiles * 256; f.scatter = scatter; f.index = d_index; f.syms = d_sym;
    f.hbuf = t->hbuf; f.v2w = t->val_fc2_w; f.v2b = t->val_fc2_b; f.policy = d_policy; f.value = d_value;
    k_heads_finish<<<(n + HF_WARPS - 1) / HF_WARPS, HF_WARPS * 32, 0, S_(stream)>>>(f);
    SGO_LAUNCHED(e);
    return 0;
}

// model.predict_on_batch for n positions (games: which=0, leaf slots: which=1) gathered by
// d_index (NULL = 0..n-1), symmetry ids d_sym per position (NULL = identity).  Outputs are
// compact [n] rows, or scattered to row d_index[i] when scatter != 0.
extern "C" int sgo_tower_forward(sgo_engine *e, int32_t slot, int32_t which, const int32_t *d_index, int32_t n,
                                 const int32_t *d_sym, int32_t scatter, float *d_policy, float *d_value, void *stream)
{
    if (slot < 0 || slot > 1 || !e->tower[slot]) return sgo_fail(e, "tower slot has no weights");
    sgo_tower *t = e->tower[slot];
    if (n < 0 || n > t->max_pos) return sgo_fail(e, "n exceeds the tower's max_positions");
    if (n == 0) return 0;
    const Board *boards = which ? e->leaf_boards : e->boards;
    int P = t->W * t->W;
    const bool prof = t->prof_on && t->prof_n < TW_PROF_MAX;
    cudaEvent_t *pe = prof ? t->prof_ev + (size_t)t->prof_n * 4 : nullptr;
    if (prof) cudaEventRecord(pe[0], S_(stream));
    // stem = one GEMM (K = 192) over the im2col of the bitboards, bias/ReLU epilogue
#ifndef SGO_STEM_SEPARATE                           // -DSGO_STEM_SEPARATE: the A/B baseline of profiles/r02_stem_fused_ab.json
    if (t->n_blocks > 0) {                          // im2col built inside the GEMM's producer (stem_fused.cuh)
        StemArgs sa;
        sa.boards = boards; sa.index = d_index; sa.syms = d_sym; sa.n = n; sa.S = t->S; sa.W = t->W; sa.PX = t->PX;
        sa.Q = n * t->PX; sa.n_tiles = (sa.Q + 255) / 256; sa.bias = t->stem_b; sa.out = t->act[0]; sa.err = t->err;
        int pairs = t->sm_count / 2;
        if (pairs > sa.n_tiles) pairs = sa.n_tiles;
        k_stem_fused<<<2 * pairs, SF_THREADS, SF_SMEM_BYTES, S_(stream)>>>(t->pmaps[3], sa);
        SGO_LAUNCHED(e);
    } else
#endif
    {                                               // towers without blocks: the stem GEMM's epilogue carries the 1x1 head convs
        int g2 = (n + IM_NP - 1) / IM_NP;
        if (g2 > 8 * t->sm_count) g2 = 8 * t->sm_count;
        k_stem_im2col<<<g2, IM_THREADS, 0, S_(stream)>>>(boards, d_index, d_sym, n, t->S, t->stem_col);
        SGO_LAUNCHED(e);
        int rc0 = launch_conv(e, t, n, -1, 3, 0, -1, stream, t->n_blocks == 0);
        if (rc0) return rc0;
    }
    if (prof) cudaEventRecord(pe[1], S_(stream));
    int x = 0;                                     // act[x] holds the block input
    for (int b = 0; b < t->n_blocks; b++) {
        int tmp = (x + 1) % 3, y = (x + 2) % 3;
        int rc = launch_conv(e, t, n, 2 * b, x, tmp, -1, stream);           // conv1 + BN + ReLU   (model.py:39-41)
        if (rc) return rc;
        rc = launch_conv(e, t, n, 2 * b + 1, tmp, y, x, stream, b == t->n_blocks - 1);   // conv2 + BN + skip + ReLU (model.py:42-45)
        if (rc) return rc;
        x = y;
    }
    if (prof) cudaEventRecord(pe[2], S_(stream));
    {
        int rc = launch_heads(e, t, n, d_index, d_sym, scatter, d_policy, d_value, stream);
        if (rc) return rc;
    }
    if (prof) { cudaEventRecord(pe[3], S_(stream)); t->prof_pos[t->prof_n++] = n; }
    return 0;
}

// live kernel timing for the roofline figures: enable, run, then read (synchronises).
// h_out[0..2] = total ms in stem / conv layers / heads, h_out[3] = conv kernel launches,
// h_out[4] = positions evaluated, h_out[5] = forward calls profiled
extern "C" int sgo_tower_profile(sgo_engine *e, int32_t slot, int32_t enable)
{
    if (slot < 0 || slot > 1 || !e->tower[slot]) return sgo_fail(e, "tower slot has no weights");
    sgo_tower *t = e->tower[slot];
    if (enable && !t->prof_ev) {
        t->prof_ev = new cudaEvent_t[TW_PROF_MAX * 4];
        t->prof_pos = new int[TW_PROF_MAX];
        for (int i = 0; i < TW_PROF_MAX * 4; i++) SGO_CUDA_OK(e, cudaEventCreate(&t->prof_ev[i]));
    }
    t->prof_on = enable;
    t->prof_n = 0;
    return 0;
}

extern "C" int sgo_tower_profile_read_sync(sgo_engine *e, int32_t slot, double *h_out)
{
    if (slot < 0 || slot > 1 || !e->tower[slot]) return sgo_fail(e, "tower slot has no weights");
    sgo_tower *t = e->tower[slot];
    SGO_CUDA_OK(e, cudaDeviceSynchronize());
    for (int i = 0; i < 6; i++) h_out[i] = 0;
    for (int i = 0; i < t->prof_n; i++) {
        cudaEvent_t *pe = t->prof_ev + (size_t)i * 4;
        for (int k = 0; k < 3; k++) {
            float ms = 0;
            SGO_CUDA_OK(e, cudaEventElapsedTime(&ms, pe[k], pe[k + 1]));
            h_out[k] += ms;
        }
        h_out[3] += t->n_layers;
        h_out[4] += t->prof_pos[i];
    }
    h_out[5] = t->prof_n;
    t->prof_n = 0;
    return 0;
}

// debugging / parity hook: copy the activations after `layer` convs... kept minimal: returns
// the sticky tower error flag (bit 16 = mbarrier wait timed out) and clears it.
extern "C" int sgo_tower_check_sync(sgo_engine *e, int32_t slot, int32_t *h_flags, void *stream)
{
    if (slot < 0 || slot > 1 || !e->tower[slot]) return sgo_fail(e, "tower slot has no weights");
    SGO_CUDA_OK(e, cudaMemcpyAsync(e->h_pinned + 6, e->tower[slot]->err, sizeof(int32_t), cudaMemcpyDeviceToHost, S_(stream)));
    SGO_CUDA_OK(e, cudaMemsetAsync(e->tower[slot]->err, 0, sizeof(int32_t), S_(stream)));
    SGO_CUDA_OK(e, cudaStreamSynchronize(S_(stream)));
    if (h_flags) *h_flags = e->h_pinned[6];
    return 0;
}

// tensor-core conv in isolation (tests / profiling): runs conv layer `layer` from act[in] to act[out]
extern "C" int sgo_tower_debug_conv(sgo_engine *e, int32_t slot, int32_t n, int32_t layer, int32_t in, int32_t out, int32_t skip, void *stream)
{
    if (slot < 0 || slot > 1 || !e->tower[slot]) return sgo_fail(e, "tower slot has no weights");
    sgo_tower *t = e->tower[slot];
    if (n < 1 || n > t->max_pos || layer < 0 || layer >= t->n_layers || in < 0 || in > 2 || out < 0 || out > 2 || skip > 2)
        return sgo_fail(e, "debug_conv arguments out of range");
    return launch_conv(e, t, n, layer, in, out, skip, stream);
}

// raw activation buffer access (bf16 [n*W*W][C], dense) for tests
extern "C" int sgo_tower_act_copy(sgo_engine *e, int32_t slot, int32_t buf, int32_t n, void *d_data, int32_t to_tower, void *stream)
{
    if (slot < 0 || slot > 1 || !e->tower[slot]) return sgo_fail(e, "tower slot has no weights");
    sgo_tower *t = e->tower[slot];
    if (buf < 0 || buf > 2 || n < 1 || n > t->max_pos) return sgo_fail(e, "act_copy arguments out of range");
    size_t bytes = (size_t)n * t->PX * TW_C * sizeof(__nv_bfloat16);
    if (to_tower) SGO_CUDA_OK(e, cudaMemcpyAsync(t->act[buf], d_data, bytes, cudaMemcpyDeviceToDevice, S_(stream)));
    else SGO_CUDA_OK(e, cudaMemcpyAsync(d_data, t->act[buf], bytes, cudaMemcpyDeviceToDevice, S_(stream)));
    return 0;
}
