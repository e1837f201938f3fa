// mask_gemm_tma.cu -- the mask-logit contraction with bf16 operands fed by TMA tensor loads, and the operand producers.
//
//   out[n,S] = q[n,d] . mf[S,d]^T (+ attention mask)          instance_seg_3d_decoder.py:567-573
//   q  = LayerNorm(queries)  (self.out_norm, :558)             -> sd3d_layernorm_cast writes fp32 (for the cls / sem /
//   mf = x_mask MLP output   (:261-263, :645)                     score heads) AND bf16 (for this kernel) in one pass
//
// mask_gemm.cu stages fp32 operands through registers (LDG + cvt + swizzled STS in every CTA): at 5000 x 5000 x 256 that
// staging and the 100 MB fp32 output bound it (9 % of the bf16 peak). Here the operands are bf16 in HBM, so
//   * warp 0   issues cp.async.bulk.tensor.2d (TMA, SWIZZLE_128B boxes of 128 / 256 rows x 64 bf16) straight into the
//              canonical K-major layout the UMMA descriptors expect: no load / convert / store instructions at all;
//   * warp 1   issues tcgen05.mma.cta_group::1.kind::f16 M=128, N=256 (bf16) or 128 (bf16x3), K=16 into one of two TMEM
//              accumulator stages and commits to the mbarriers that free the B ring slot and publish the accumulator;
//   * warps 2-9 are two epilogue groups, one per TMEM stage (tiles alternate between them): each warp reads its lane
//              quarter with tcgen05.ld, stages 32 x 32 fp32 blocks in shared memory and stores them with
//              cp.async.bulk.tensor (full 128-byte lines), writes the thresholded mask bytes, tracks all-true rows and -- the warp that
//              finishes the last tile of its 32 rows -- resets them in place (flags + progress counters in `ws`).
// Persistent CTAs (one per SM, 225 KB of shared memory): CTA c owns a contiguous range of 128-column half tiles
// (row block major), reloading the 128 x d A block only when the row block changes.
//
// What was measured while shaping it (5000 x 5000 x 256, ncu kernel times, debug builds that switched parts off):
//   * the synchronisation skeleton alone (no loads, MMAs, tcgen05.ld, stores) cost 1.1 us per 128 x 128 tile: the MMA
//     thread pays ~60 cycles per tcgen05.mma issue and ~200 per tcgen05.commit (clock64 around the issue block), i.e.
//     more than the 64 tensor cycles of an N = 128 instruction -> N = 256 instructions (half the issues per flop):
//     front end 24.6 -> 17.6 us; ring depth (2..8 K blocks in flight), polling vs try_wait, converged vs single-lane
//     producer / issuer warps, 4 vs 8 epilogue warps made no difference;
//   * TMA loads of B are not the limit (skipping them: -0.4 us); the output path adds ~8 us on top of the front end
//     (shared-memory bandwidth: MMA operand reads + TMA-in + staging write + TMA-store read ~ 4.5 k cycles per 128 x 256
//     tile); plain 16-byte-per-lane stores instead of the TMA store: 46 us.
#include <cuda.h>

#include <cmath>

#include "common.cuh"

namespace sd3d {

constexpr int kTmBM = 128, kTmKB = 64;  // tile rows, bf16 elements per 128-byte swizzle row
// tile columns = N of one tcgen05.mma: 256 for bf16 (one MMA issue + commit costs the issuing thread ~60 / ~200 cycles,
// measured; at N = 128 that alone exceeds the 64 tensor cycles of the instruction), 128 for bf16x3 (shared memory)
constexpr int tm_bn(bool split) { return split ? 128 : 256; }
constexpr int kTmThreads = 320;                       // warp 0 TMA, warp 1 MMA, warps 2..5 / 6..9 epilogue of TMEM stage 0 / 1
constexpr int kTmStages = 2;
constexpr int kTmSlab = kTmBM * 128;                  // one K block of a 128-row operand tile: 16 KB

__device__ __forceinline__ uint32_t tm_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// bounded wait: a lost arrival becomes a reported error (trap) instead of a hung GPU
__device__ __forceinline__ void tm_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void tm_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tm_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tm_commit(uint32_t bar) {  // arrives when all previously issued MMAs have completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tm_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100: version 1 at bit 46, layout type 2 at 61)
__device__ __forceinline__ uint64_t tm_desc(uint32_t smem_addr) {
    uint64_t desc = 0;
    desc |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    desc |= (uint64_t)1u << 16;
    desc |= (uint64_t)((1024u >> 4) & 0x3FFFu) << 32;  // 8 rows x 128 B between row groups
    desc |= (uint64_t)1u << 46;
    desc |= (uint64_t)2u << 61;
    return desc;
}

// A block (128 rows x d) resident in shared memory; B streams through a ring of K-block steps (one step = the 64-wide
// K block kb of a tile, ring_steps steps in flight), so that shared memory is left for double-buffered epilogue staging.
// SPLIT = false: operands are [rows, d] bf16; a step is 256 rows x 64 of mf (32 KB, one TMA box).
// SPLIT = true ("bf16x3", fp32-level accuracy): operands are [rows, 2d] bf16 = (hi | mid) with hi = bf16(x),
// mid = bf16(x - hi); out = hi.mid + mid.hi + hi.hi accumulated in the fp32 TMEM accumulator (the dropped terms are
// <= 2^-16 relative per product). Both halves of the A block stay resident; a step is the (hi, mid) slab pair of 128 rows.
//
// Work is dealt out in 128-column half tiles (row-block major): CTA c owns half tiles [c * halves_per_cta, ...). An even
// half tile and its successor (same row block, same CTA) are processed as ONE 128 x 256 tile (N = 256 MMAs), the
// rest as 128 x 128 tiles (the same 256-row TMA box, N = 128 in the instruction descriptor): full-width MMAs where
// possible and a makespan quantised in half tiles.
constexpr int kTmMaxRing = 4;

struct TmTile {
    int mb, col0, width;
};
__device__ __forceinline__ TmTile tm_next_tile(int64_t& h, int64_t h_end, int n_half, int bn) {
    const int mb = (int)(h / n_half), hb = (int)(h % n_half);
    const int halves = (bn == 256 && (hb & 1) == 0 && h + 1 < h_end && hb + 1 < n_half) ? 2 : 1;
    h += halves;
    return TmTile{mb, hb * 128, halves * 128};
}

template <bool SPLIT>
__global__ void __launch_bounds__(kTmThreads, 1)
    mask_logits_tma_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_mf,
                           const __grid_constant__ CUtensorMap map_out, int tma_store, int n, int S, int d,
                           int halves_per_cta, int ring_steps, int epi_bufs, float* __restrict__ out, float thr,
                           uint8_t* __restrict__ attn, int32_t* __restrict__ row_false) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kblocks = d / kTmKB;
    constexpr int kTmBN = tm_bn(SPLIT);
    constexpr int kStepSlabs = 2;  // bf16: 256 rows of one K block; bf16x3: the (hi, mid) slabs of 128 rows
    const int a_slabs = SPLIT ? 2 * kblocks : kblocks;    // resident A block
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)a_slabs * kTmSlab;       // ring_steps steps
    uint8_t* sEpi = sB + (size_t)ring_steps * kStepSlabs * kTmSlab;  // 8 epilogue warps x epi_bufs x [32 rows][128 B] (swizzled)
    // a_full[kb], a_free[kb] (one pair per K block of the A block), b_full[ring], b_free[ring], acc_full[2], acc_free[2]
    constexpr int kMaxKb = 256 / kTmKB;
    __shared__ __align__(8) uint64_t s_bar[2 * kMaxKb + 2 * kTmMaxRing + 2 * kTmStages];
    __shared__ uint32_t s_tmem;
    const uint32_t a_full = tm_smem(&s_bar[0]), a_free = tm_smem(&s_bar[kMaxKb]);
    const uint32_t b_full = tm_smem(&s_bar[2 * kMaxKb]), b_free = tm_smem(&s_bar[2 * kMaxKb + kTmMaxRing]);
    const uint32_t acc_full = tm_smem(&s_bar[2 * kMaxKb + 2 * kTmMaxRing]),
                   acc_free = tm_smem(&s_bar[2 * kMaxKb + 2 * kTmMaxRing + kTmStages]);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_half = (S + 127) / 128, m_blocks = (n + kTmBM - 1) / kTmBM;
    const int64_t total = (int64_t)n_half * m_blocks;
    const int64_t h_begin = (int64_t)blockIdx.x * halves_per_cta;
    const int64_t h_end = h_begin + halves_per_cta < total ? h_begin + halves_per_cta : total;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tm_smem(&s_tmem)),
                     "r"((uint32_t)(kTmBN * kTmStages))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 * kMaxKb + 2 * kTmMaxRing + kTmStages; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tm_smem(&s_bar[i])) : "memory");
        for (int i = 0; i < kTmStages; ++i)  // accumulator stage freed by the 4 epilogue warps of its group
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(acc_free + 8 * i) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    const uint32_t a_kb_bytes = (uint32_t)(SPLIT ? 2 : 1) * kTmSlab, step_bytes = (uint32_t)kStepSlabs * kTmSlab;

    if (warp == 0) {
        // ---------------- TMA producer: the whole warp walks the loop (converged), lane 0 issues ----------------
        int cur_m = -1, a_loads = 0;
        int bs = 0;                 // ring slot of the current step
        uint32_t lap = 0;           // step / ring_steps
        for (int64_t h = h_begin; h < h_end;) {
            const TmTile tile = tm_next_tile(h, h_end, n_half, kTmBN);
            if (tile.mb != cur_m) {
                // the A block is replaced K block by K block: slab kb is refilled as soon as the last tile of the old row
                // block has consumed it, while that tile's MMAs on the later K blocks are still running
                for (int kb = 0; kb < kblocks; ++kb) {
                    if (a_loads > 0) tm_wait(a_free + 8 * kb, (uint32_t)((a_loads - 1) & 1));
                    if (lane == 0) {
                        tm_expect_tx(a_full + 8 * kb, a_kb_bytes);
                        tma_load_2d(tm_smem(sA + (size_t)kb * kTmSlab), &map_q, kb * kTmKB, tile.mb * kTmBM, a_full + 8 * kb);
                        if constexpr (SPLIT)  // slab kblocks + kb is the mid half (columns d + kb * 64)
                            tma_load_2d(tm_smem(sA + (size_t)(kblocks + kb) * kTmSlab), &map_q, d + kb * kTmKB, tile.mb * kTmBM,
                                        a_full + 8 * kb);
                    }
                    __syncwarp();
                }
                cur_m = tile.mb;
                ++a_loads;
            }
            for (int kb = 0; kb < kblocks; ++kb) {
                if (lap > 0) tm_wait(b_free + 8 * bs, (lap - 1) & 1);
                if (lane == 0) {
                    // the box is always the full step (256 rows, or 128 rows x (hi, mid)); rows past S arrive as zeros
                    tm_expect_tx(b_full + 8 * bs, step_bytes);
                    uint8_t* dst = sB + (size_t)bs * kStepSlabs * kTmSlab;
                    tma_load_2d(tm_smem(dst), &map_mf, kb * kTmKB, tile.col0, b_full + 8 * bs);
                    if constexpr (SPLIT) tma_load_2d(tm_smem(dst + kTmSlab), &map_mf, d + kb * kTmKB, tile.col0, b_full + 8 * bs);
                }
                __syncwarp();
                if (++bs == ring_steps) { bs = 0; ++lap; }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: converged warp, lane 0 issues tcgen05.mma / commit ----------------
        // instruction descriptor: kind::f16, A = B = bf16, D = f32, both K-major, M = 128, N = tile width
        const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTmBM >> 4) << 24);
        const uint64_t descA0 = tm_desc(tm_smem(sA));
        int cur_m = -1, a_loads = 0;
        int64_t i = 0;
        int bs = 0;
        uint32_t lap = 0;
        for (int64_t h = h_begin; h < h_end; ++i) {
            const TmTile tile = tm_next_tile(h, h_end, n_half, kTmBN);
            const uint32_t idesc = idesc0 | ((uint32_t)(tile.width >> 3) << 17);
            const bool fresh_a = tile.mb != cur_m;  // first tile of a row block: wait for each A slab before its first use
            if (fresh_a) {
                cur_m = tile.mb;
                ++a_loads;
            }
            // last tile of the row block (and more work follows): hand each A slab back right after its last use
            const bool release_a = h < h_end && (int)(h / n_half) != tile.mb;
            const int st = (int)(i % kTmStages);  // accumulator stage
            if (i >= kTmStages) tm_wait(acc_free + 8 * st, (uint32_t)(((i / kTmStages) - 1) & 1));
            const uint32_t tmem_d = tmem_base + (uint32_t)(st * kTmBN);
            for (int kb = 0; kb < kblocks; ++kb) {
                if (fresh_a) tm_wait(a_full + 8 * kb, (uint32_t)((a_loads - 1) & 1));
                tm_wait(b_full + 8 * bs, lap & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint64_t a_hi = descA0 + (uint64_t)(((uint32_t)kb * kTmSlab) >> 4);
                    const uint64_t b_hi = tm_desc(tm_smem(sB + (size_t)bs * kStepSlabs * kTmSlab));
#pragma unroll
                    for (int ks = 0; ks < kTmKB / 16; ++ks) {
                        const uint64_t o = (uint64_t)((ks * 32) >> 4);
                        if constexpr (SPLIT) {
                            const uint64_t a_mid = descA0 + (uint64_t)(((uint32_t)(kblocks + kb) * kTmSlab) >> 4);
                            const uint64_t b_mid = b_hi + (uint64_t)(kTmSlab >> 4);
                            tm_mma(tmem_d, a_hi + o, b_mid + o, idesc, (kb | ks) ? 1u : 0u);  // small terms first
                            tm_mma(tmem_d, a_mid + o, b_hi + o, idesc, 1u);
                            tm_mma(tmem_d, a_hi + o, b_hi + o, idesc, 1u);
                        } else {
                            tm_mma(tmem_d, a_hi + o, b_hi + o, idesc, (kb | ks) ? 1u : 0u);
                        }
                    }
                    tm_commit(b_free + 8 * bs);  // the ring slot may be overwritten once these MMAs are done
                    if (release_a) tm_commit(a_free + 8 * kb);
                    if (kb == kblocks - 1) tm_commit(acc_full + 8 * st);  // the accumulator is complete
                }
                __syncwarp();
                if (++bs == ring_steps) { bs = 0; ++lap; }
            }
        }
    } else {
        // ---------------- epilogue: group g = warps 2+4g .. 5+4g owns TMEM stage g, lane quarter = warp % 4 ----------------
        const int quarter = warp & 3, group = (warp - 2) >> 2;
        int chunk_no = 0;
        bool row_has_false = false;
        int32_t* progress = row_false + n;  // [m_blocks][4]: half tiles finished per (row block, lane quarter)
        int64_t i = 0;
        for (int64_t h = h_begin; h < h_end; ++i) {
            const TmTile tile = tm_next_tile(h, h_end, n_half, kTmBN);
            if ((int)(i % kTmStages) != group) continue;
            const int mb = tile.mb;
            const int st = group;
            tm_wait(acc_full + 8 * st, (uint32_t)((i / kTmStages) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int gm = mb * kTmBM + quarter * 32 + lane;
            const int n0 = tile.col0;
#pragma unroll 1
            for (int c0 = 0; c0 < tile.width; c0 += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(st * kTmBN + c0);
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
                if (tma_store && n0 + c0 < S) {
                    // the warp's 32 x 32 fp32 block goes through shared memory (128-byte rows, 16-byte chunks XOR-swizzled
                    // like the tensor map) and leaves as ONE bulk tensor store of full 128-byte lines; rows >= n and
                    // columns >= S are clipped by the tensor map
                    const uint32_t buf_s = tm_smem(sEpi + (size_t)((warp - 2) * epi_bufs + (epi_bufs > 1 ? (chunk_no & 1) : 0)) * 4096);
                    if (lane == 0) {  // this buffer's previous store has been read
                        if (epi_bufs > 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(buf_s + lane * 128 + ((j ^ (lane & 7)) << 4)),
                                     "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                                     : "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map_out),
                                     "r"(n0 + c0), "r"(mb * kTmBM + quarter * 32), "r"(buf_s)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    ++chunk_no;
                    if (attn != nullptr) {
                        // mask bytes from the staged block, transposed: 8 lanes cover the 32 columns of one row, so a store
                        // instruction writes 4 rows x 32 contiguous bytes (whole sectors) instead of 32 rows x 16 bytes
                        const int sub = lane & 7, grp = lane >> 3;
                        const int col = n0 + c0 + 4 * sub;
                        uint32_t valid = 0;  // byte lanes of columns < S
#pragma unroll
                        for (int b = 0; b < 4; ++b) valid |= (col + b < S ? 1u : 0u) << (8 * b);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int r = 4 * k + grp;
                            float4 f;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                         : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w)
                                         : "r"(buf_s + r * 128 + ((sub ^ (r & 7)) << 4)));
                            const uint32_t w = (f.x < thr ? 1u : 0u) | (f.y < thr ? 0x100u : 0u) | (f.z < thr ? 0x10000u : 0u) |
                                               (f.w < thr ? 0x1000000u : 0u);
                            const int grow = mb * kTmBM + quarter * 32 + r;
                            if (grow < n && valid != 0) {
                                uint8_t* dst = attn + (int64_t)grow * S + col;
                                if (valid == 0x01010101u) {
                                    *reinterpret_cast<uint32_t*>(dst) = w;  // S % 4 == 0 on this path: 4-byte aligned
                                } else {
#pragma unroll
                                    for (int b = 0; b < 4; ++b)
                                        if (col + b < S) dst[b] = (uint8_t)((w >> (8 * b)) & 1u);
                                }
                            }
                            const uint32_t bal = __ballot_sync(kFull, ((w ^ 0x01010101u) & valid) != 0);
                            if ((lane >> 2) == k) row_has_false |= ((bal >> (8 * (lane & 3))) & 0xFFu) != 0;  // lane = row 4k + grp
                        }
                    }
                } else if (!tma_store && gm < n && n0 + c0 < S) {
                    float* orow = out + (int64_t)gm * S + n0 + c0;
                    if (n0 + c0 + 32 <= S && (S & 3) == 0) {
                        if (!tma_store) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                *reinterpret_cast<float4*>(orow + j) =
                                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                                __uint_as_float(v[j + 3]));
                        }
                        if (attn) {
                            uint8_t* arow = attn + (int64_t)gm * S + n0 + c0;
                            uint32_t w[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                w[j] = 0;
#pragma unroll
                                for (int k = 0; k < 4; ++k) w[j] |= (__uint_as_float(v[4 * j + k]) < thr ? 1u : 0u) << (8 * k);
                                row_has_false |= w[j] != 0x01010101u;
                            }
                            if ((S & 15) == 0) {  // the widest store the row alignment allows: fewer, fuller L2 write requests
                                reinterpret_cast<uint4*>(arow)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                                reinterpret_cast<uint4*>(arow)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                            } else if ((S & 7) == 0) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) reinterpret_cast<uint2*>(arow)[j] = make_uint2(w[2 * j], w[2 * j + 1]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) reinterpret_cast<uint32_t*>(arow)[j] = w[j];
                            }
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int gn = n0 + c0 + j;
                            if (gn < S) {
                                const float val = __uint_as_float(v[j]);
                                if (!tma_store) out[(int64_t)gm * S + gn] = val;
                                if (attn) {
                                    attn[(int64_t)gm * S + gn] = val < thr ? 1 : 0;
                                    row_has_false |= !(val < thr);
                                }
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) tm_arrive(acc_free + 8 * st);
            if (attn != nullptr) {
                // publish this warp's 32 rows of the tile: the "row has a false entry" flags, then (behind a fence) the
                // progress counter of its (row block, quarter). The warp that completes the counter knows every mask byte
                // of those rows is written and resets the all-true rows (instance_seg_3d_decoder.py:570-571) right here:
                // no second pass over the mask, no extra launch. Flags and counters are handed back zeroed.
                if (row_has_false && gm < n) atomicOr(row_false + gm, 1);
                row_has_false = false;
                __threadfence();
                __syncwarp();
                const int halves = tile.width >> 7;
                int done = 0;
                if (lane == 0) done = atomicAdd(progress + mb * 4 + quarter, halves) + halves;
                done = __shfl_sync(kFull, done, 0);
                if (done == n_half) {
                    __threadfence();
                    const int flag = gm < n ? atomicExch(row_false + gm, 0) : 1;
                    uint32_t reset = __ballot_sync(kFull, flag == 0);
                    while (reset) {
                        const int r = __ffs(reset) - 1;
                        reset &= reset - 1;
                        uint8_t* arow = attn + (int64_t)(mb * kTmBM + quarter * 32 + r) * S;
                        if ((S & 3) == 0)
                            for (int c = lane; c < (S >> 2); c += 32) reinterpret_cast<uint32_t*>(arow)[c] = 0u;
                        else
                            for (int c = lane; c < S; c += 32) arow[c] = 0;
                    }
                    if (lane == 0) progress[mb * 4 + quarter] = 0;
                }
            }
        }
        if (tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory stays valid until read
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)(kTmBN * kTmStages))
                     : "memory");
    }
}

// y = LayerNorm(x) * w + b over the last dimension (w, b nullable: plain copy / cast), written as fp32 and / or bf16.
// One warp per row; mean and variance as two passes over the row held in registers (d <= 1024).
__global__ void __launch_bounds__(256) layernorm_cast_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ b, int n, int d, float eps,
                                                             int normalize, float* __restrict__ y32,
                                                             __nv_bfloat16* __restrict__ y16) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const float* xr = x + (int64_t)row * d;
    float v[32];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int c = lane + 32 * j;
        v[j] = c < d ? xr[c] : 0.f;
        sum += v[j];
    }
    float mean = 0.f, rstd = 1.f;
    if (normalize) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
        mean = sum / (float)d;
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int c = lane + 32 * j;
            const float t = c < d ? v[j] - mean : 0.f;
            sq += t * t;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(kFull, sq, o);
        rstd = rsqrtf(sq / (float)d + eps);
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int c = lane + 32 * j;
        if (c < d) {
            float y = normalize ? (v[j] - mean) * rstd : v[j];
            if (w) y = y * w[c];
            if (b) y = y + b[c];
            if (y32) y32[(int64_t)row * d + c] = y;
            if (y16) y16[(int64_t)row * d + c] = __float2bfloat16_rn(y);
        }
    }
}

// x[n,d] fp32 -> y[n,2d] bf16 = (hi | mid): hi = bf16(x), mid = bf16(x - hi); x = hi + mid up to 2^-17 relative.
// One thread per 4 consecutive elements (d % 4 == 0).
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, int64_t quads, int d,
                                                         __nv_bfloat16* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= quads) return;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const int64_t e = i * 4, row = e / d;
    const int c = (int)(e - row * d);
    const float f[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 hi[4], mid[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        hi[k] = __float2bfloat16_rn(f[k]);
        mid[k] = __float2bfloat16_rn(f[k] - __bfloat162float(hi[k]));
    }
    __nv_bfloat16* yr = y + row * 2 * d + c;
    *reinterpret_cast<uint2*>(yr) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(yr + d) = *reinterpret_cast<const uint2*>(mid);
}

// x[n*d] fp32 -> bf16, 4 elements per thread (the plain operand cast: no normalisation, no affine)
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, int64_t quads, __nv_bfloat16* __restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= quads) return;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    __nv_bfloat16 h[4] = {__float2bfloat16_rn(v.x), __float2bfloat16_rn(v.y), __float2bfloat16_rn(v.z), __float2bfloat16_rn(v.w)};
    reinterpret_cast<uint2*>(y)[i] = *reinterpret_cast<const uint2*>(h);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;  // benign race: same value from every thread
    if (fn != nullptr) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess || p == nullptr) {
        cudaGetLastError();
        return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// out[n, S] fp32 row-major -> boxes of 32 rows x 32 columns (128-byte rows), 128-byte swizzle
static bool make_out_map(CUtensorMap* map, const void* base, int n, int S) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (enc == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)S, (cuuint64_t)n};
    const cuuint64_t strides[1] = {(cuuint64_t)S * 4};
    const cuuint32_t box[2] = {32, 32};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// [rows, d] bf16 row-major -> boxes of box_rows (128 / 256) rows x 64 elements, 128-byte swizzle, zero fill outside
static bool make_operand_map(CUtensorMap* map, const void* base, int rows, int d, int box_rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (enc == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)d * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kTmKB, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace sd3d

using namespace sd3d;

extern "C" int sd3d_layernorm_cast(const float* x, const float* weight, const float* bias, int n, int d, float eps,
                                   int normalize, float* y_f32, void* y_bf16, void* stream_) {
    if (n < 0 || d <= 0 || d > 1024) {
        set_error("sd3d_layernorm_cast: bad shape n=%d d=%d (d <= 1024)", n, d);
        return SD3D_ERR_ARG;
    }
    if (n == 0) return SD3D_OK;
    if (x == nullptr || (y_f32 == nullptr && y_bf16 == nullptr)) {
        set_error("sd3d_layernorm_cast: null input or no output");
        return SD3D_ERR_ARG;
    }
    if (!normalize && weight == nullptr && bias == nullptr && y_f32 == nullptr && d % 4 == 0 && aligned16(x) &&
        (reinterpret_cast<uintptr_t>(y_bf16) & 7) == 0) {  // the plain operand cast: vectorised
        const int64_t quads = (int64_t)n * d / 4;
        cast_bf16_kernel<<<(unsigned)ceil_div64(quads, 256), 256, 0, (cudaStream_t)stream_>>>(x, quads,
                                                                                             static_cast<__nv_bfloat16*>(y_bf16));
        return check_launch("sd3d_layernorm_cast");
    }
    layernorm_cast_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(x, weight, bias, n, d, eps, normalize, y_f32,
                                                                          static_cast<__nv_bfloat16*>(y_bf16));
    return check_launch("sd3d_layernorm_cast");
}

// n row flags + 4 progress counters per 128-row block
extern "C" size_t sd3d_mask_logits_bf16_workspace_bytes(int n) {
    return n > 0 ? ((size_t)n + 4 * (size_t)((n + kTmBM - 1) / kTmBM)) * sizeof(int32_t) : 0;
}

// shared launcher of the two operand formats: split = 0 -> [rows, d] bf16, split = 1 -> [rows, 2d] bf16 (hi | mid)
static int launch_tma(const char* name, const void* q, const void* mf, int n, int S, int d, int split, float* out, float thr,
                      uint8_t* attn_mask, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (n < 0 || S < 0 || d <= 0) {
        set_error("%s: bad shape n=%d S=%d d=%d", name, n, S, d);
        return SD3D_ERR_ARG;
    }
    if (n == 0 || S == 0) return SD3D_OK;
    if (d % kTmKB != 0 || d > 256) {
        set_error("%s: needs d %% 64 == 0 and d <= 256 (d=%d)", name, d);
        return SD3D_ERR_UNSUPPORTED;
    }
    if (q == nullptr || mf == nullptr || out == nullptr || !aligned16(q) || !aligned16(mf) || !aligned16(out)) {
        set_error("%s: null or misaligned buffer", name);
        return SD3D_ERR_ARG;
    }
    if (attn_mask != nullptr && (reinterpret_cast<uintptr_t>(attn_mask) & 15) != 0) {
        set_error("%s: attn_mask must be 16-byte aligned", name);
        return SD3D_ERR_ARG;
    }
    if (attn_mask != nullptr && (ws == nullptr || ws_bytes < sd3d_mask_logits_bf16_workspace_bytes(n))) {
        set_error("%s: the attention mask needs a zero-filled workspace of sd3d_mask_logits_bf16_workspace_bytes(n) bytes", name);
        return SD3D_ERR_ARG;
    }
    CUtensorMap map_q, map_mf, map_out;
    const int tma_store = (S % 4 == 0) ? 1 : 0;  // the tensor map needs 16-byte row strides; otherwise plain stores
    const int kcols = split ? 2 * d : d;
    if (!make_operand_map(&map_q, q, n, kcols, kTmBM) || !make_operand_map(&map_mf, mf, S, kcols, tm_bn(split != 0)) ||
        !make_out_map(&map_out, tma_store ? (const void*)out : q, tma_store ? n : 32, tma_store ? S : 32)) {
        set_error("%s: cuTensorMapEncodeTiled failed", name);
        return SD3D_ERR_CUDA;
    }
    if (attn_mask != nullptr) {
        const double t = (double)thr;  // sigmoid(x) < thr  <=>  x < logit(thr)
        thr = t <= 0.0 ? -INFINITY : (t >= 1.0 ? INFINITY : (float)log(t / (1.0 - t)));
    }
    // shared memory: resident A block + B ring + epilogue staging. bf16: A <= 64 KB, ring of 3 steps x 32 KB (256 rows of
    // one K block), staging double-buffered (64 KB); bf16x3: A <= 128 KB, ring of 2 steps x 32 KB ((hi, mid) of 128
    // rows), staging single-buffered (32 KB)
    const int kblocks = d / kTmKB;
    const int ring_steps = split ? 2 : 3, epi_bufs = split ? 1 : 2;
    const size_t smem = ((size_t)(split ? 2 : 1) * kblocks + (size_t)ring_steps * 2) * kTmSlab + (size_t)8 * epi_bufs * 4096 + 1024;
    static std::atomic<uint64_t> attr_set{0};
    if (first_on_device(&attr_set)) {
        const int max_smem = 12 * kTmSlab + 8 * 4096 + 1024;  // bf16x3 at d = 256 (bf16: 9 slabs + 64 KB)
        cudaError_t e = cudaFuncSetAttribute(mask_logits_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(mask_logits_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
        if (e != cudaSuccess) {
            attr_set.store(0);
            set_error("%s: cudaFuncSetAttribute: %s", name, cudaGetErrorString(e));
            return SD3D_ERR_CUDA;
        }
    }
    const int64_t total = (int64_t)((S + 127) / 128) * ((n + kTmBM - 1) / kTmBM);  // 128-column half tiles
    const int ctas = (int)imin64(total, num_sms());
    const int halves_per_cta = (int)ceil_div64(total, ctas);
    const int grid = (int)ceil_div64(total, halves_per_cta);
    if (split)
        mask_logits_tma_kernel<true><<<grid, kTmThreads, smem, stream>>>(map_q, map_mf, map_out, tma_store, n, S, d, halves_per_cta,
                                                                         ring_steps, epi_bufs, out, thr, attn_mask,
                                                                         static_cast<int32_t*>(ws));
    else
        mask_logits_tma_kernel<false><<<grid, kTmThreads, smem, stream>>>(map_q, map_mf, map_out, tma_store, n, S, d, halves_per_cta,
                                                                          ring_steps, epi_bufs, out, thr, attn_mask,
                                                                          static_cast<int32_t*>(ws));
    return check_launch(name);
}

extern "C" int sd3d_mask_logits_bf16(const void* q_bf16, const void* mf_bf16, int n, int S, int d, float* out, float thr,
                                     uint8_t* attn_mask, void* ws, size_t ws_bytes, void* stream_) {
    return launch_tma("sd3d_mask_logits_bf16", q_bf16, mf_bf16, n, S, d, 0, out, thr, attn_mask, ws, ws_bytes,
                      (cudaStream_t)stream_);
}

extern "C" int sd3d_mask_logits_bf16x3(const void* q_split, const void* mf_split, int n, int S, int d, float* out, float thr,
                                       uint8_t* attn_mask, void* ws, size_t ws_bytes, void* stream_) {
    return launch_tma("sd3d_mask_logits_bf16x3", q_split, mf_split, n, S, d, 1, out, thr, attn_mask, ws, ws_bytes,
                      (cudaStream_t)stream_);
}

extern "C" int sd3d_split_bf16(const float* x, int n, int d, void* y_split, void* stream_) {
    if (n < 0 || d <= 0 || d % 4 != 0) {
        set_error("sd3d_split_bf16: bad shape n=%d d=%d (d %% 4 == 0)", n, d);
        return SD3D_ERR_ARG;
    }
    if (n == 0) return SD3D_OK;
    if (x == nullptr || y_split == nullptr || !aligned16(x) || !aligned16(y_split)) {
        set_error("sd3d_split_bf16: null or misaligned buffer");
        return SD3D_ERR_ARG;
    }
    const int64_t quads = (int64_t)n * d / 4;
    split_bf16_kernel<<<(unsigned)ceil_div64(quads, 256), 256, 0, (cudaStream_t)stream_>>>(
        x, quads, d, static_cast<__nv_bfloat16*>(y_split));
    return check_launch("sd3d_split_bf16");
}

// fp32 operands in, eval-scale problem: the operand conversion (plain bf16 cast, or the (hi | mid) split for the
// fp32-tolerance path) and the TMA-fed GEMM behind ONE host call. scratch holds the converted operands (any contents),
// flags is the zero-filled, self-cleaning workspace of sd3d_mask_logits_bf16 (only used with attn_mask).
static size_t tm_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" size_t sd3d_mask_logits_large_scratch_bytes(int n, int S, int d, int precision) {
    if (n <= 0 || S <= 0 || d <= 0) return 0;
    const size_t per_elem = precision == SD3D_BF16 ? 2 : 4;  // bf16, or two bf16 halves
    return tm_align_up((size_t)n * d * per_elem, 256) + tm_align_up((size_t)S * d * per_elem, 256);
}

extern "C" int sd3d_mask_logits_large(const float* q, const float* mf, int n, int S, int d, int precision, float* out,
                                      float thr, uint8_t* attn_mask, void* scratch, size_t scratch_bytes, void* flags,
                                      size_t flags_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (precision != SD3D_BF16 && precision != SD3D_F32) {
        set_error("sd3d_mask_logits_large: precision %d unknown", precision);
        return SD3D_ERR_UNSUPPORTED;
    }
    if (n < 0 || S < 0 || d <= 0) {
        set_error("sd3d_mask_logits_large: bad shape n=%d S=%d d=%d", n, S, d);
        return SD3D_ERR_ARG;
    }
    if (n == 0 || S == 0) return SD3D_OK;
    if (d % kTmKB != 0 || d > 256) {
        set_error("sd3d_mask_logits_large: needs d %% 64 == 0 and d <= 256 (d=%d)", d);
        return SD3D_ERR_UNSUPPORTED;
    }
    if (q == nullptr || mf == nullptr || scratch == nullptr || !aligned16(q) || !aligned16(mf) || !aligned16(scratch) ||
        scratch_bytes < sd3d_mask_logits_large_scratch_bytes(n, S, d, precision)) {
        set_error("sd3d_mask_logits_large: null / misaligned operand or scratch < sd3d_mask_logits_large_scratch_bytes()");
        return SD3D_ERR_ARG;
    }
    const bool split = precision == SD3D_F32;
    const size_t per_elem = split ? 4 : 2;
    uint8_t* q_conv = static_cast<uint8_t*>(scratch);
    uint8_t* mf_conv = q_conv + tm_align_up((size_t)n * d * per_elem, 256);
    if (split) {
        const int64_t qq = (int64_t)n * d / 4, mq = (int64_t)S * d / 4;
        split_bf16_kernel<<<(unsigned)ceil_div64(qq, 256), 256, 0, stream>>>(q, qq, d, reinterpret_cast<__nv_bfloat16*>(q_conv));
        split_bf16_kernel<<<(unsigned)ceil_div64(mq, 256), 256, 0, stream>>>(mf, mq, d, reinterpret_cast<__nv_bfloat16*>(mf_conv));
    } else {
        const int64_t qq = (int64_t)n * d / 4, mq = (int64_t)S * d / 4;
        cast_bf16_kernel<<<(unsigned)ceil_div64(qq, 256), 256, 0, stream>>>(q, qq, reinterpret_cast<__nv_bfloat16*>(q_conv));
        cast_bf16_kernel<<<(unsigned)ceil_div64(mq, 256), 256, 0, stream>>>(mf, mq, reinterpret_cast<__nv_bfloat16*>(mf_conv));
    }
    return launch_tma("sd3d_mask_logits_large", q_conv, mf_conv, n, S, d, split ? 1 : 0, out, thr, attn_mask, flags, flags_bytes,
                      stream);
}
