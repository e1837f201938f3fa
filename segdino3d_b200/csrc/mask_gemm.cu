// mask_gemm.cu -- query x superpoint mask logits: out[n,S] = q[n,d] . mf[S,d]^T  (+ fused attention mask).
//
// Replaces   pred_mask = torch.einsum('nd,md->nm', norm_query, mask_feats[i])
//            attn_mask = (pred_mask.sigmoid() < thr); all-true rows reset to all-false
// at /root/reference/segdino3d/models/decoder/instance_seg_3d_decoder.py:567-571 (and :339-343), which the
// reference runs as cuBLAS SGEMM + ~5 elementwise/reduce kernels, 7 times per forward.
//
// Two precisions behind one entry point:
//   SD3D_F32  : fp32 FFMA register-tiled kernel, K summed in ascending order (<=1e-5 rel.).
//   SD3D_BF16 : the one dense contraction of the path -> 5th-gen tensor cores. Both operands are K-major
//               ("TN"), so each CTA converts its fp32 operand tiles to bf16 straight into the canonical
//               K-major SWIZZLE_128B shared-memory layout (no transpose, no extra HBM pass), one elected
//               thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=64, K=16 x d/16) with the fp32
//               accumulator in TMEM, tcgen05.commit signals an mbarrier, and the four warps read their
//               32 TMEM lanes back with tcgen05.ld for the epilogue (logits + sigmoid threshold).
// Tensor-pipe roofline note: 2*n*S*d = 51 MFLOP at the ScanNet200 shape -> launch/latency bound; see DESIGN.md.
#include <cmath>

#include "common.cuh"

namespace sd3d {

// ------------------------------------------------------------------------------------------------
// fp32 path
// ------------------------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
    mask_logits_f32_kernel(const float* __restrict__ q, const float* __restrict__ mf, int n, int S, int d,
                           float* __restrict__ out, float thr, uint8_t* __restrict__ attn) {
    constexpr int BK = 32;
    constexpr int NT = (BM / TM) * (BN / TN);
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < d; k0 += BK) {
        for (int e = tid; e < BM * BK; e += NT) {
            const int r = e / BK, k = e % BK;
            const int gm = m0 + r, gk = k0 + k;
            As[k][r] = (gm < n && gk < d) ? __ldg(q + (int64_t)gm * d + gk) : 0.f;
        }
        for (int e = tid; e < BN * BK; e += NT) {
            const int r = e / BK, k = e % BK;
            const int gn = n0 + r, gk = k0 + k;
            Bs[k][r] = (gn < S && gk < d) ? __ldg(mf + (int64_t)gn * d + gk) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gm = m0 + ty * TM + i;
        if (gm >= n) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gn = n0 + tx * TN + j;
            if (gn >= S) continue;
            const float v = acc[i][j];
            out[(int64_t)gm * S + gn] = v;
            if (attn) attn[(int64_t)gm * S + gn] = v < thr ? 1 : 0;  // thr is already a logit
        }
    }
}

// rows whose mask is all-true are reset to all-false (instance_seg_3d_decoder.py:570-571); warp per row
__global__ void attn_mask_fix_kernel(uint8_t* __restrict__ attn, int n, int S) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const int lane = lane_id();
    uint8_t* r = attn + (int64_t)row * S;
    int all_true = 1;
    for (int c = lane; c < S; c += 32) all_true &= (r[c] != 0);
    all_true = __all_sync(kFull, all_true);
    if (all_true)
        for (int c = lane; c < S; c += 32) r[c] = 0;
}

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM path (bf16 operands, fp32 accumulate)
// ------------------------------------------------------------------------------------------------
constexpr int kTcBM = 128;      // UMMA_M
constexpr int kTcBN = 64;       // UMMA_N (TMEM columns per accumulator stage)
constexpr int kTcThreads = 256; // 8 warps: warp w reads TMEM lanes 32*(w&3).. and column half (w>>2)
constexpr int kTcKBlock = 64;   // bf16 elements per 128-byte swizzle row
constexpr int kTcMaxTiles = 8;  // N tiles one CTA walks with its A operand resident in shared memory

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 format: version=1 at bit 46, layout type 2 at 61)
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t desc = 0;
    desc |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);        // start address, 16-byte units
    desc |= (uint64_t)1u << 16;                            // leading byte offset (unused for swizzled K-major)
    desc |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;      // stride byte offset: 8 rows x 128 B
    desc |= (uint64_t)1u << 46;                            // descriptor version (Blackwell)
    desc |= (uint64_t)2u << 61;                            // LayoutType::SWIZZLE_128B
    return desc;
}

// instruction descriptor: kind::f16, A=B=bf16, D=f32, both K-major, M=128, N=kTcBN
__device__ __forceinline__ uint32_t make_idesc_bf16_f32(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 24)) __trap();  // never hang the GPU: a lost commit becomes a reported error
    }
}

// Convert a [rows x d] fp32 row-major tile (rows beyond `valid_rows` are zero) to bf16 in the canonical
// K-major SWIZZLE_128B layout: K-block kb (64 elements) is a [ROWS x 128 B] slab; 16-byte chunk j of row r
// lands at r*128 + ((j ^ (r & 7)) << 4). Four chunks (8 x 128-bit loads) are in flight per thread.
template <int ROWS>
__device__ __forceinline__ void stage_operand(const float* __restrict__ g, int64_t row0, int valid_rows, int d,
                                              uint8_t* smem_tile) {
    const int chunks_per_row = d >> 3;
    const int total = ROWS * chunks_per_row;
    constexpr int kU = 4;
    for (int e0 = threadIdx.x; e0 < total; e0 += kTcThreads * kU) {
        float4 lo[kU], hi[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int e = e0 + u * kTcThreads;
            const int r = e / chunks_per_row, kc = e % chunks_per_row;
            lo[u] = hi[u] = f4_zero();
            if (e < total && r < valid_rows) {
                const float* src = g + (row0 + r) * (int64_t)d + kc * 8;
                lo[u] = ldg_f4(src);
                hi[u] = ldg_f4(src + 4);
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int e = e0 + u * kTcThreads;
            if (e >= total) continue;
            const int r = e / chunks_per_row, kc = e % chunks_per_row;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(lo[u].x, lo[u].y), p1 = __floats2bfloat162_rn(lo[u].z, lo[u].w);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(hi[u].x, hi[u].y), p3 = __floats2bfloat162_rn(hi[u].z, hi[u].w);
            uint4 packed;
            packed.x = *reinterpret_cast<uint32_t*>(&p0);
            packed.y = *reinterpret_cast<uint32_t*>(&p1);
            packed.z = *reinterpret_cast<uint32_t*>(&p2);
            packed.w = *reinterpret_cast<uint32_t*>(&p3);
            const int kb = kc >> 3, j = kc & 7;
            uint8_t* dst = smem_tile + (size_t)kb * (ROWS * 128) + r * 128 + ((j ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = packed;
        }
    }
}

// epilogue of one 128 x 64 accumulator stage: warp w reads TMEM lanes 32*(w&3)..+31 (= accumulator rows) and the
// 32 columns of half (w>>2); thread = one row x 32 columns -> 128-bit stores
__device__ __forceinline__ bool tc_epilogue(uint32_t tmem_stage, int m0, int n0, int n, int S, float* __restrict__ out,
                                            float thr, uint8_t* __restrict__ attn) {
    bool all_true = true;  // of this thread's (row, 32-column) part, columns < S only
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int c0 = (warp >> 2) * 32;
    const int gm = m0 + (warp & 3) * 32 + lane;
    uint32_t v[32];
    const uint32_t taddr = tmem_stage + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (gm < n) {
        float* orow = out + (int64_t)gm * S + n0 + c0;
        const bool full = (n0 + c0 + 32 <= S);
        if (full && (S & 3) == 0) {  // 16-byte aligned row chunks: 128-bit stores
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(orow + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            if (attn) {
                uint32_t* arow = reinterpret_cast<uint32_t*>(attn + (int64_t)gm * S + n0 + c0);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    uint32_t w = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) w |= (__uint_as_float(v[j + k]) < thr ? 1u : 0u) << (8 * k);
                    arow[j >> 2] = w;
                    all_true &= w == 0x01010101u;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int gn = n0 + c0 + j;
                if (gn < S) {
                    const float val = __uint_as_float(v[j]);
                    out[(int64_t)gm * S + gn] = val;
                    if (attn) attn[(int64_t)gm * S + gn] = val < thr ? 1 : 0;  // thr is already a logit
                    all_true &= val < thr;
                }
            }
        }
    }
    return all_true;
}

// One CTA = one 128-row block of queries x `tiles` consecutive 64-column tiles of superpoints. The A operand is
// converted to bf16 once and stays in shared memory; B tiles and TMEM accumulators are STAGES-deep, so with
// STAGES == 2 the tensor core works on tile t (async, tcgen05.commit -> mbarrier) while all warps write out
// tile t-1 and then stage tile t+1: one __syncthreads per tile.
// If the CTA owns ALL tiles of its rows (fuse_reset), the all-true-row reset of instance_seg_3d_decoder.py:570-571 is
// done here: every thread tracks whether its part of its row was all-true, the two column halves meet in shared
// memory, and the (rare) all-true rows are rewritten with zeros -- no second pass over the mask.
template <int STAGES>
__device__ __forceinline__ void mask_logits_tc_body(const float* __restrict__ q, const float* __restrict__ mf, int n,
                                                    int S, int d, int tiles, float* __restrict__ out, float thr,
                                                    uint8_t* __restrict__ attn, int m_block, int tile_group,
                                                    bool fuse_reset) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment required by SWIZZLE_128B atoms
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kblocks = d / kTcKBlock;
    const size_t a_bytes = (size_t)kblocks * (kTcBM * 128), b_bytes = (size_t)kblocks * (kTcBN * 128);
    uint8_t* sA = smem;
    uint8_t* sB = smem + a_bytes;
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ uint32_t s_tmem;
    __shared__ uint8_t s_alltrue[2][kTcBM];
    bool row_all_true = true;

    const int warp = threadIdx.x >> 5;
    const int m0 = m_block * kTcBM;
    const int nt0 = tile_group * tiles;
    const int n_tiles_total = (S + kTcBN - 1) / kTcBN;
    const int my_tiles = min(tiles, n_tiles_total - nt0);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "r"((uint32_t)(kTcBN * STAGES))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    stage_operand<kTcBM>(q, m0, min(kTcBM, n - m0), d, sA);
    stage_operand<kTcBN>(mf, (int64_t)nt0 * kTcBN, min(kTcBN, S - nt0 * kTcBN), d, sB);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    const uint32_t idesc = make_idesc_bf16_f32(kTcBM, kTcBN);
    const uint64_t descA0 = make_kmajor_sw128_desc(smem_u32(sA));

    for (int t = 0; t < my_tiles; ++t) {
        const int st = (STAGES == 2) ? (t & 1) : 0;
        // B tile t is staged and visible here (prologue / previous iteration + barrier): issue its MMAs
        if (threadIdx.x == 0) {
            const uint64_t descB0 = make_kmajor_sw128_desc(smem_u32(sB + (size_t)st * b_bytes));
            const uint32_t tmem_d = tmem_base + (uint32_t)(st * kTcBN);
            for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll
                for (int ks = 0; ks < kTcKBlock / 16; ++ks) {
                    const uint64_t da = descA0 + (uint64_t)(((uint32_t)kb * (kTcBM * 128) + ks * 32) >> 4);
                    const uint64_t db = descB0 + (uint64_t)(((uint32_t)kb * (kTcBN * 128) + ks * 32) >> 4);
                    const uint32_t accumulate = (kb | ks) ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "setp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                        :
                        : "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
                        : "memory");
                }
            }
            // commit: arrives on the stage's mbarrier when the MMAs above have completed
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_u32(&s_bar[st]))
                         : "memory");
        }
        __syncwarp();
        if (STAGES == 2) {
            // while the tensor core works on tile t: write out tile t-1, then stage tile t+1
            if (t > 0) {
                mbar_wait_parity(smem_u32(&s_bar[st ^ 1]), (uint32_t)(((t - 1) >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                row_all_true &= tc_epilogue(tmem_base + (uint32_t)((st ^ 1) * kTcBN), m0, (nt0 + t - 1) * kTcBN, n, S, out, thr, attn);
            }
            if (t + 1 < my_tiles) {  // stage st^1 is free: its MMAs (tile t-1) were waited for above
                stage_operand<kTcBN>(mf, (int64_t)(nt0 + t + 1) * kTcBN, min(kTcBN, S - (nt0 + t + 1) * kTcBN), d,
                                     sB + (size_t)(st ^ 1) * b_bytes);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
        } else {
            mbar_wait_parity(smem_u32(&s_bar[0]), (uint32_t)(t & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            row_all_true &= tc_epilogue(tmem_base, m0, (nt0 + t) * kTcBN, n, S, out, thr, attn);
            if (t + 1 < my_tiles) {
                stage_operand<kTcBN>(mf, (int64_t)(nt0 + t + 1) * kTcBN, min(kTcBN, S - (nt0 + t + 1) * kTcBN), d, sB);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // TMEM stage read out + next B tile staged, for every warp
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (STAGES == 2 && my_tiles > 0) {  // drain: the last tile
        const int t = my_tiles - 1, st = t & 1;
        mbar_wait_parity(smem_u32(&s_bar[st]), (uint32_t)((t >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        row_all_true &= tc_epilogue(tmem_base + (uint32_t)(st * kTcBN), m0, (nt0 + t) * kTcBN, n, S, out, thr, attn);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (fuse_reset && attn != nullptr) {  // uniform over the CTA
        s_alltrue[warp >> 2][(warp & 3) * 32 + lane_id()] = row_all_true ? 1 : 0;
        __syncthreads();  // (also orders every warp's mask stores before the rewrite below)
        if (warp < 4) {
            const int r = warp * 32 + lane_id(), gm = m0 + r;
            if (gm < n && s_alltrue[0][r] && s_alltrue[1][r]) {
                uint8_t* arow = attn + (int64_t)gm * S;
                for (int c = 0; c < S; ++c) arow[c] = 0;
            }
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)(kTcBN * STAGES))
                     : "memory");
    }
}


template <int STAGES>
__global__ void __launch_bounds__(kTcThreads)
    mask_logits_tc_kernel(const float* __restrict__ q, const float* __restrict__ mf, int n, int S, int d, int tiles,
                          float* __restrict__ out, float thr, uint8_t* __restrict__ attn, int fuse_reset) {
    mask_logits_tc_body<STAGES>(q, mf, n, S, d, tiles, out, thr, attn, blockIdx.y, blockIdx.x, fuse_reset != 0);
}

// ---- batched over the scenes of a batch (the per-scene python loop of _forward_head, instance_seg_3d_decoder.py:557):
// one launch for all (scene, row block, tile group) CTAs
constexpr int kMaxMaskProblems = 32;
struct MaskProblem {
    const float* q;
    const float* mf;
    float* out;
    uint8_t* attn;
    int n, S, cta_begin, tiles, groups, fuse_reset;
};
struct MaskBatch {
    MaskProblem p[kMaxMaskProblems];
    int count;
};

template <int STAGES>
__global__ void __launch_bounds__(kTcThreads)
    mask_logits_tc_batched_kernel(const __grid_constant__ MaskBatch b, int d, float thr) {
    int i = 0;
    while (i + 1 < b.count && (int)blockIdx.x >= b.p[i + 1].cta_begin) ++i;
    const MaskProblem& P = b.p[i];
    const int local = (int)blockIdx.x - P.cta_begin;
    mask_logits_tc_body<STAGES>(P.q, P.mf, P.n, P.S, d, P.tiles, P.out, thr, P.attn, local / P.groups, local % P.groups,
                                P.fuse_reset != 0);
}

__global__ void attn_mask_fix_batched_kernel(const __grid_constant__ MaskBatch b) {
    const MaskProblem& P = b.p[blockIdx.y];
    if (P.attn == nullptr || P.fuse_reset) return;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= P.n) return;
    const int lane = lane_id();
    uint8_t* r = P.attn + (int64_t)row * P.S;
    int all_true = 1;
    for (int c = lane; c < P.S; c += 32) all_true &= (r[c] != 0);
    all_true = __all_sync(kFull, all_true);
    if (all_true)
        for (int c = lane; c < P.S; c += 32) r[c] = 0;
}

// fp32 path, batched: one 32 x 32 tile per CTA, problems concatenated along blockIdx.x
__global__ void __launch_bounds__(256) mask_logits_f32_batched_kernel(const __grid_constant__ MaskBatch b, int d, float thr) {
    int i = 0;
    while (i + 1 < b.count && (int)blockIdx.x >= b.p[i + 1].cta_begin) ++i;
    const MaskProblem& P = b.p[i];
    const int local = (int)blockIdx.x - P.cta_begin;
    constexpr int BM = 32, BN = 32, BK = 32, TM = 2, TN = 2, NT = 256;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = (local / P.groups) * BM, n0 = (local % P.groups) * BN;
    float acc[TM][TN] = {{0.f, 0.f}, {0.f, 0.f}};
    for (int k0 = 0; k0 < d; k0 += BK) {
        for (int e = tid; e < BM * BK; e += NT) {
            const int r = e / BK, k = e % BK;
            As[k][r] = (m0 + r < P.n && k0 + k < d) ? __ldg(P.q + (int64_t)(m0 + r) * d + k0 + k) : 0.f;
        }
        for (int e = tid; e < BN * BK; e += NT) {
            const int r = e / BK, k = e % BK;
            Bs[k][r] = (n0 + r < P.S && k0 + k < d) ? __ldg(P.mf + (int64_t)(n0 + r) * d + k0 + k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
#pragma unroll
            for (int a = 0; a < TM; ++a)
#pragma unroll
                for (int c = 0; c < TN; ++c) acc[a][c] = fmaf(As[k][ty * TM + a], Bs[k][tx * TN + c], acc[a][c]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int c = 0; c < TN; ++c) {
            const int gm = m0 + ty * TM + a, gn = n0 + tx * TN + c;
            if (gm < P.n && gn < P.S) {
                P.out[(int64_t)gm * P.S + gn] = acc[a][c];
                if (P.attn) P.attn[(int64_t)gm * P.S + gn] = acc[a][c] < thr ? 1 : 0;
            }
        }
}
}  // namespace sd3d

using namespace sd3d;

static int set_tc_attributes();

extern "C" int sd3d_mask_logits(const float* q, const float* mf, int n, int S, int d, int precision, float* out,
                                float thr, uint8_t* attn_mask, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || S < 0 || d <= 0) {
        set_error("sd3d_mask_logits: bad shape n=%d S=%d d=%d", n, S, d);
        return SD3D_ERR_ARG;
    }
    if (n == 0 || S == 0) return SD3D_OK;
    // sigmoid(x) < thr  <=>  x < logit(thr): the epilogue compares logits (no expf per element)
    if (attn_mask != nullptr) {
        const double t = (double)thr;
        thr = t <= 0.0 ? -INFINITY : (t >= 1.0 ? INFINITY : (float)log(t / (1.0 - t)));
    }
    if (q == nullptr || mf == nullptr || out == nullptr) {
        set_error("sd3d_mask_logits: null buffer");
        return SD3D_ERR_ARG;
    }
    bool fused_reset = false;
    if (precision == SD3D_F32) {
        const bool big = ((int64_t)((n + 63) / 64) * ((S + 63) / 64)) >= 2 * (int64_t)num_sms();
        if (big) {
            dim3 grid((S + 63) / 64, (n + 63) / 64);
            mask_logits_f32_kernel<64, 64, 4, 4><<<grid, 256, 0, stream>>>(q, mf, n, S, d, out, thr, attn_mask);
        } else {
            dim3 grid((S + 31) / 32, (n + 31) / 32);
            mask_logits_f32_kernel<32, 32, 2, 2><<<grid, 256, 0, stream>>>(q, mf, n, S, d, out, thr, attn_mask);
        }
    } else if (precision == SD3D_BF16) {
        if (d % kTcKBlock != 0 || d > 512 || !aligned16(q) || !aligned16(mf)) {
            set_error("sd3d_mask_logits: tcgen05 path needs d %% 64 == 0, d <= 512 and 16-byte aligned operands (d=%d)",
                      d);
            return SD3D_ERR_UNSUPPORTED;
        }
        const int stages = d <= 256 ? 2 : 1;  // two B / TMEM stages fit beside a 128 x 256 bf16 A tile
        const size_t smem = (size_t)(d / kTcKBlock) * (kTcBM + stages * kTcBN) * 128 + 1024;
        const int rc_attr = set_tc_attributes();
        if (rc_attr != SD3D_OK) return rc_attr;
        // N tiles per CTA: keep >= ~2 CTAs per SM in the grid, at most kTcMaxTiles per CTA
        const int n_tiles = (S + kTcBN - 1) / kTcBN, m_blocks = (n + kTcBM - 1) / kTcBM;
        int tiles = (int)(((int64_t)n_tiles * m_blocks) / (2 * (int64_t)num_sms()));
        if (tiles < 1) tiles = 1;
        if (tiles > kTcMaxTiles) tiles = kTcMaxTiles;
        dim3 grid((n_tiles + tiles - 1) / tiles, m_blocks);
        fused_reset = attn_mask != nullptr && grid.x == 1;  // one CTA per row block sees whole rows
        if (stages == 2)
            mask_logits_tc_kernel<2><<<grid, kTcThreads, smem, stream>>>(q, mf, n, S, d, tiles, out, thr, attn_mask, fused_reset);
        else
            mask_logits_tc_kernel<1><<<grid, kTcThreads, smem, stream>>>(q, mf, n, S, d, tiles, out, thr, attn_mask, fused_reset);
    } else {
        set_error("sd3d_mask_logits: precision code %d unsupported (SD3D_F32 | SD3D_BF16)", precision);
        return SD3D_ERR_UNSUPPORTED;
    }
    if (attn_mask != nullptr && !fused_reset) {
        attn_mask_fix_kernel<<<(n + 7) / 8, 256, 0, stream>>>(attn_mask, n, S);
    }
    return check_launch("sd3d_mask_logits");
}

static int set_tc_attributes() {
    static std::atomic<uint64_t> attr_set{0};  // per-device flag (the attribute is per device)
    if (!first_on_device(&attr_set)) return SD3D_OK;
    const int b1 = (int)(512 / kTcKBlock * (kTcBM + kTcBN) * 128 + 1024), b2 = (int)(256 / kTcKBlock * (kTcBM + 2 * kTcBN) * 128 + 1024);
    cudaError_t e = cudaFuncSetAttribute(mask_logits_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, b1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mask_logits_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, b2);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mask_logits_tc_batched_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, b1);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mask_logits_tc_batched_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, b2);
    if (e != cudaSuccess) {
        attr_set.store(0);
        set_error("sd3d_mask_logits: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return SD3D_ERR_CUDA;
    }
    return SD3D_OK;
}

extern "C" int sd3d_mask_logits_batched(const float* const* q_host, const float* const* mf_host, const int* n_host,
                                        const int* S_host, int count, int d, int precision, float* const* out_host,
                                        float thr, uint8_t* const* attn_host, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (count < 0 || d <= 0 || (count > 0 && (q_host == nullptr || mf_host == nullptr || n_host == nullptr ||
                                              S_host == nullptr || out_host == nullptr))) {
        set_error("sd3d_mask_logits_batched: bad argument (count=%d d=%d)", count, d);
        return SD3D_ERR_ARG;
    }
    if (precision != SD3D_F32 && precision != SD3D_BF16) {
        set_error("sd3d_mask_logits_batched: precision code %d unsupported (SD3D_F32 | SD3D_BF16)", precision);
        return SD3D_ERR_UNSUPPORTED;
    }
    const bool want_attn = attn_host != nullptr;
    if (want_attn) {
        const double t = (double)thr;
        thr = t <= 0.0 ? -INFINITY : (t >= 1.0 ? INFINITY : (float)log(t / (1.0 - t)));
    }
    const int stages = d <= 256 ? 2 : 1;
    if (precision == SD3D_BF16) {
        if (d % kTcKBlock != 0 || d > 512) {
            set_error("sd3d_mask_logits_batched: tcgen05 path needs d %% 64 == 0 and d <= 512 (d=%d)", d);
            return SD3D_ERR_UNSUPPORTED;
        }
        const int rc = set_tc_attributes();
        if (rc != SD3D_OK) return rc;
    }
    for (int base = 0; base < count; base += kMaxMaskProblems) {  // (a batch has a handful of scenes: one pass)
        MaskBatch b;
        b.count = 0;
        int ctas = 0, max_n = 0;
        bool any_unfused = false;
        for (int i = base; i < count && b.count < kMaxMaskProblems; ++i) {
            const int n = n_host[i], S = S_host[i];
            if (n < 0 || S < 0) {
                set_error("sd3d_mask_logits_batched: problem %d has a negative size", i);
                return SD3D_ERR_ARG;
            }
            if (n == 0 || S == 0) continue;
            if (q_host[i] == nullptr || mf_host[i] == nullptr || out_host[i] == nullptr ||
                (precision == SD3D_BF16 && (!aligned16(q_host[i]) || !aligned16(mf_host[i])))) {
                set_error("sd3d_mask_logits_batched: problem %d has a null or misaligned buffer", i);
                return SD3D_ERR_ARG;
            }
            MaskProblem& P = b.p[b.count++];
            P.q = q_host[i];
            P.mf = mf_host[i];
            P.out = out_host[i];
            P.attn = want_attn ? attn_host[i] : nullptr;
            P.n = n;
            P.S = S;
            P.cta_begin = ctas;
            if (precision == SD3D_BF16) {
                const int n_tiles = (S + kTcBN - 1) / kTcBN, m_blocks = (n + kTcBM - 1) / kTcBM;
                P.tiles = n_tiles <= kTcMaxTiles ? n_tiles : kTcMaxTiles;  // whole rows per CTA when they fit
                P.groups = (n_tiles + P.tiles - 1) / P.tiles;
                P.fuse_reset = (P.attn != nullptr && P.groups == 1) ? 1 : 0;
                ctas += m_blocks * P.groups;
            } else {
                P.tiles = 1;
                P.groups = (S + 31) / 32;
                P.fuse_reset = 0;
                ctas += ((n + 31) / 32) * P.groups;
            }
            any_unfused |= P.attn != nullptr && !P.fuse_reset;
            max_n = n > max_n ? n : max_n;
        }
        if (b.count == 0) continue;
        if (precision == SD3D_BF16) {
            const size_t smem = (size_t)(d / kTcKBlock) * (kTcBM + stages * kTcBN) * 128 + 1024;
            if (stages == 2) mask_logits_tc_batched_kernel<2><<<ctas, kTcThreads, smem, stream>>>(b, d, thr);
            else mask_logits_tc_batched_kernel<1><<<ctas, kTcThreads, smem, stream>>>(b, d, thr);
        } else {
            mask_logits_f32_batched_kernel<<<ctas, 256, 0, stream>>>(b, d, thr);
        }
        if (any_unfused) attn_mask_fix_batched_kernel<<<dim3((max_n + 7) / 8, b.count), 256, 0, stream>>>(b);
    }
    return check_launch("sd3d_mask_logits_batched");
}
