// mask_gemm_tma.cu -- the mask-logit contraction with bf16 operands fed by TMA tensor loads, and the operand producers.
//
//   out[n,S] = q[n,d] . mf[S,d]^T (+ attention mask)          instance_seg_3d_decoder.py:567-573
//   q  = LayerNorm(queries)  (self.out_norm, :558)             -> sd3d_layernorm_cast writes fp32 (for the cls / sem /
//   mf = x_mask MLP output   (:261-263, :645)                     score heads) AND bf16 (for this kernel) in one pass
//
// mask_gemm.cu stages fp32 operands through registers (LDG + cvt + swizzled STS in every CTA): at 5000 x 5000 x 256 that
// staging and the 100 MB fp32 output bound it (9 % of the bf16 peak). Here the operands are bf16 in HBM, so
//   * warp 0   issues cp.async.bulk.tensor.2d (TMA, SWIZZLE_128B boxes of 128 rows x 64 bf16) straight into the canonical
//              K-major layout the UMMA descriptors expect: no load / convert / store instructions at all;
//   * warp 1   issues tcgen05.mma.cta_group::1.kind::f16 M=128, N=128, K=16 (d/16 per tile) into one of two TMEM stages
//              and commits to the mbarriers that free the B stage and publish the accumulator;
//   * warps 2-5 read the accumulator with tcgen05.ld (each warp its lane quarter), write the fp32 logits with 128-bit
//              stores and the thresholded mask bytes, and track all-true rows.
// Persistent CTAs (one per SM, 192 KB of shared memory): CTA c owns a contiguous range of the linear tile index
// (row block major), reloading the 128 x d A block only when the row block changes.
#include <cuda.h>

#include <cmath>

#include "common.cuh"

namespace sd3d {

constexpr int kTmBM = 128, kTmBN = 128, kTmKB = 64;  // tile rows / columns, bf16 elements per 128-byte swizzle row
constexpr int kTmThreads = 192;                       // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
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

__global__ void __launch_bounds__(kTmThreads, 1)
    mask_logits_tma_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_mf,
                           const __grid_constant__ CUtensorMap map_out, int tma_store, int n, int S, int d,
                           int tiles_per_cta, float* __restrict__ out, float thr, uint8_t* __restrict__ attn,
                           int32_t* __restrict__ row_false) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int kblocks = d / kTmKB;
    uint8_t* sA = smem;                                   // kblocks slabs
    uint8_t* sB = smem + (size_t)kblocks * kTmSlab;       // kTmStages x kblocks slabs
    uint8_t* sEpi = sB + (size_t)kTmStages * kblocks * kTmSlab;  // 4 epilogue warps x 2 buffers x [32 rows][128 B] (swizzled)
    __shared__ __align__(8) uint64_t s_bar[2 + 4 * kTmStages];  // a_full, a_free, b_full[2], b_free[2], acc_full[2], acc_free[2]
    __shared__ uint32_t s_tmem;
    const uint32_t a_full = tm_smem(&s_bar[0]), a_free = tm_smem(&s_bar[1]);
    const uint32_t b_full = tm_smem(&s_bar[2]), b_free = tm_smem(&s_bar[2 + kTmStages]);
    const uint32_t acc_full = tm_smem(&s_bar[2 + 2 * kTmStages]), acc_free = tm_smem(&s_bar[2 + 3 * kTmStages]);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = (S + kTmBN - 1) / kTmBN, m_blocks = (n + kTmBM - 1) / kTmBM;
    const int64_t total = (int64_t)n_tiles * m_blocks;
    const int64_t t_begin = (int64_t)blockIdx.x * tiles_per_cta;
    const int64_t t_end = t_begin + tiles_per_cta < total ? t_begin + tiles_per_cta : total;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tm_smem(&s_tmem)),
                     "r"((uint32_t)(kTmBN * kTmStages))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 + 3 * kTmStages; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tm_smem(&s_bar[i])) : "memory");
        for (int i = 0; i < kTmStages; ++i)  // accumulator stage freed by the 4 epilogue warps
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(acc_free + 8 * i) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    const uint32_t slab_bytes = (uint32_t)kblocks * kTmSlab;

    if (warp == 0) {
        // ---------------- TMA producer (one elected lane) ----------------
        if (lane == 0) {
            int cur_m = -1, a_loads = 0;
            int64_t i = 0;
            for (int64_t t = t_begin; t < t_end; ++t, ++i) {
                const int mb = (int)(t / n_tiles), nt = (int)(t % n_tiles);
                if (mb != cur_m) {
                    if (a_loads > 0) tm_wait(a_free, (uint32_t)((a_loads - 1) & 1));  // MMAs of the previous row block are done
                    tm_expect_tx(a_full, slab_bytes);
                    for (int kb = 0; kb < kblocks; ++kb)
                        tma_load_2d(tm_smem(sA + (size_t)kb * kTmSlab), &map_q, kb * kTmKB, mb * kTmBM, a_full);
                    cur_m = mb;
                    ++a_loads;
                }
                const int st = (int)(i % kTmStages);
                if (i >= kTmStages) tm_wait(b_free + 8 * st, (uint32_t)(((i / kTmStages) - 1) & 1));
                tm_expect_tx(b_full + 8 * st, slab_bytes);
                for (int kb = 0; kb < kblocks; ++kb)
                    tma_load_2d(tm_smem(sB + ((size_t)st * kblocks + kb) * kTmSlab), &map_mf, kb * kTmKB, nt * kTmBN,
                                b_full + 8 * st);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (one elected lane) ----------------
        if (lane == 0) {
            // instruction descriptor: kind::f16, A = B = bf16, D = f32, both K-major, M = 128, N = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTmBN >> 3) << 17) | ((uint32_t)(kTmBM >> 4) << 24);
            const uint64_t descA0 = tm_desc(tm_smem(sA));
            int cur_m = -1, a_loads = 0;
            int64_t i = 0;
            for (int64_t t = t_begin; t < t_end; ++t, ++i) {
                const int mb = (int)(t / n_tiles);
                if (mb != cur_m) {
                    if (a_loads > 0) tm_commit(a_free);  // every MMA that read the old A block
                    tm_wait(a_full, (uint32_t)(a_loads & 1));
                    cur_m = mb;
                    ++a_loads;
                }
                const int st = (int)(i % kTmStages);
                tm_wait(b_full + 8 * st, (uint32_t)((i / kTmStages) & 1));
                if (i >= kTmStages) tm_wait(acc_free + 8 * st, (uint32_t)(((i / kTmStages) - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t descB0 = tm_desc(tm_smem(sB + (size_t)st * slab_bytes));
                const uint32_t tmem_d = tmem_base + (uint32_t)(st * kTmBN);
                for (int kb = 0; kb < kblocks; ++kb) {
#pragma unroll
                    for (int ks = 0; ks < kTmKB / 16; ++ks) {
                        const uint64_t da = descA0 + (uint64_t)(((uint32_t)kb * kTmSlab + ks * 32) >> 4);
                        const uint64_t db = descB0 + (uint64_t)(((uint32_t)kb * kTmSlab + ks * 32) >> 4);
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
                tm_commit(b_free + 8 * st);    // the B stage may be overwritten
                tm_commit(acc_full + 8 * st);  // the accumulator is complete
            }
        }
    } else {
        // ---------------- epilogue: warps 2..5, TMEM lane quarter = warp % 4 ----------------
        const int quarter = warp & 3;
        int cur_m = -1, chunk_no = 0;
        bool row_has_false = false;
        int64_t i = 0;
        for (int64_t t = t_begin; t < t_end; ++t, ++i) {
            const int mb = (int)(t / n_tiles), nt = (int)(t % n_tiles);
            const int gm_old = cur_m * kTmBM + quarter * 32 + lane;
            if (mb != cur_m) {
                if (cur_m >= 0 && attn != nullptr && row_has_false && gm_old < n) atomicOr(row_false + gm_old, 1);
                cur_m = mb;
                row_has_false = false;
            }
            const int st = (int)(i % kTmStages);
            tm_wait(acc_full + 8 * st, (uint32_t)((i / kTmStages) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int gm = mb * kTmBM + quarter * 32 + lane;
            const int n0 = nt * kTmBN;
#pragma unroll 1
            for (int c0 = 0; c0 < kTmBN; c0 += 32) {
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
                    uint8_t* buf = sEpi + (size_t)((warp - 2) * 2 + (chunk_no & 1)) * 4096;
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // this buffer's previous store has been read
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                            make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map_out),
                                     "r"(n0 + c0), "r"(mb * kTmBM + quarter * 32), "r"(tm_smem(buf))
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    ++chunk_no;
                }
                if (gm < n && n0 + c0 < S) {
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
        }
        const int gm_last = cur_m * kTmBM + quarter * 32 + lane;
        if (cur_m >= 0 && attn != nullptr && row_has_false && gm_last < n) atomicOr(row_false + gm_last, 1);
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

// rows with no false entry (all-true) are reset to all-false (instance_seg_3d_decoder.py:570-571); warp per row
__global__ void attn_reset_flagged_kernel(uint8_t* __restrict__ attn, const int32_t* __restrict__ row_false, int n, int S) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n || row_false[row] != 0) return;
    uint8_t* r = attn + (int64_t)row * S;
    for (int c = threadIdx.x & 31; c < S; c += 32) r[c] = 0;
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

// [rows, d] bf16 row-major -> boxes of 128 rows x 64 elements, 128-byte swizzle, zero fill outside
static bool make_operand_map(CUtensorMap* map, const void* base, int rows, int d) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (enc == nullptr) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)d * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kTmKB, (cuuint32_t)kTmBM};
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
    layernorm_cast_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(x, weight, bias, n, d, eps, normalize, y_f32,
                                                                          static_cast<__nv_bfloat16*>(y_bf16));
    return check_launch("sd3d_layernorm_cast");
}

extern "C" size_t sd3d_mask_logits_bf16_workspace_bytes(int n) { return n > 0 ? (size_t)n * sizeof(int32_t) : 0; }

extern "C" int sd3d_mask_logits_bf16(const void* q_bf16, const void* mf_bf16, int n, int S, int d, float* out, float thr,
                                     uint8_t* attn_mask, void* ws, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || S < 0 || d <= 0) {
        set_error("sd3d_mask_logits_bf16: bad shape n=%d S=%d d=%d", n, S, d);
        return SD3D_ERR_ARG;
    }
    if (n == 0 || S == 0) return SD3D_OK;
    if (d % kTmKB != 0 || d > 256) {
        set_error("sd3d_mask_logits_bf16: needs d %% 64 == 0 and d <= 256 (d=%d)", d);
        return SD3D_ERR_UNSUPPORTED;
    }
    if (q_bf16 == nullptr || mf_bf16 == nullptr || out == nullptr || !aligned16(q_bf16) || !aligned16(mf_bf16) ||
        !aligned16(out)) {
        set_error("sd3d_mask_logits_bf16: null or misaligned buffer");
        return SD3D_ERR_ARG;
    }
    if (attn_mask != nullptr && (ws == nullptr || ws_bytes < sd3d_mask_logits_bf16_workspace_bytes(n))) {
        set_error("sd3d_mask_logits_bf16: the attention mask needs a workspace of n int32 row flags");
        return SD3D_ERR_ARG;
    }
    CUtensorMap map_q, map_mf, map_out;
    const int tma_store = (S % 4 == 0) ? 1 : 0;  // the tensor map needs 16-byte row strides; otherwise plain stores
    if (!make_operand_map(&map_q, q_bf16, n, d) || !make_operand_map(&map_mf, mf_bf16, S, d) ||
        !make_out_map(&map_out, tma_store ? out : q_bf16, tma_store ? n : 32, tma_store ? S : 32)) {
        set_error("sd3d_mask_logits_bf16: cuTensorMapEncodeTiled failed");
        return SD3D_ERR_CUDA;
    }
    if (attn_mask != nullptr) {
        const double t = (double)thr;  // sigmoid(x) < thr  <=>  x < logit(thr)
        thr = t <= 0.0 ? -INFINITY : (t >= 1.0 ? INFINITY : (float)log(t / (1.0 - t)));
        cudaMemsetAsync(ws, 0, (size_t)n * sizeof(int32_t), stream);
    }
    const size_t smem = (size_t)(d / kTmKB) * kTmSlab * (1 + kTmStages) + 4 * 2 * 4096 + 1024;
    static std::atomic<uint64_t> attr_set{0};
    if (first_on_device(&attr_set)) {
        const cudaError_t e = cudaFuncSetAttribute(mask_logits_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   (int)((256 / kTmKB) * kTmSlab * (1 + kTmStages) + 4 * 2 * 4096 + 1024));
        if (e != cudaSuccess) {
            attr_set.store(0);
            set_error("sd3d_mask_logits_bf16: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return SD3D_ERR_CUDA;
        }
    }
    const int64_t total = (int64_t)((S + kTmBN - 1) / kTmBN) * ((n + kTmBM - 1) / kTmBM);
    const int ctas = (int)imin64(total, num_sms());
    const int tiles_per_cta = (int)ceil_div64(total, ctas);
    const int grid = (int)ceil_div64(total, tiles_per_cta);
    mask_logits_tma_kernel<<<grid, kTmThreads, smem, stream>>>(map_q, map_mf, map_out, tma_store, n, S, d, tiles_per_cta, out,
                                                               thr, attn_mask, static_cast<int32_t*>(ws));
    if (attn_mask != nullptr)
        attn_reset_flagged_kernel<<<(n + 7) / 8, 256, 0, stream>>>(attn_mask, static_cast<const int32_t*>(ws), n, S);
    return check_launch("sd3d_mask_logits_bf16");
}
