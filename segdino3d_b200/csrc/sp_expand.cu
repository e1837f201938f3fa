// sp_expand.cu -- rows "next" of SURVEY.md section 8(f): superpoint -> point mask expansion, and the backward
// of the superpoint mean.
//
// (1) sd3d_sp_expand_mask replaces, in Baseline3D.predict_by_feat_instance
//     (/root/reference/segdino3d/models/architecture/baseline3d.py:453-454, :463),
//         mask_pred_sigmoid = mask_pred_sigmoid[:, superpoints]          # [K,S] -> [K,N] fp32 gather (240 MB at K=600)
//         mask_pred = mask_pred_sigmoid > self.test_cfg.sp_score_thr
//         mask_pointnum = mask_pred.sum(1)
//     with ONE pass that never materialises the fp32 [K,N] tensor: the comparison is done per superpoint, the
//     boolean row is expanded with 64-bit stores from a bit-matrix transpose + byte look-up table, the per-instance
//     point count is accumulated with integer atomics (exact). Write-bound: K*N bytes out, N*8 bytes of ids in (once
//     per 32 rows).
// (2) sd3d_sp_mean_backward: grad_src[p,:] = grad_out[idx[p],:] / max(|idx[p]|,1), the gradient of
//     scatter_mean(src, idx, dim=0) (spconvunet.py:390 is called under autograd in training,
//     engine/train_engine_3d.py:99-105). Read N*8 + S*C*4 (L2 resident), write N*C*4.
#include "common.cuh"

namespace sd3d {

constexpr int kExpThreads = 256;
constexpr int kExpPtsPerThread = 8;   // one 64-bit store of mask bytes per row
constexpr int kExpRows = 32;          // instances per CTA: the ids are loaded once for 32 output rows
constexpr int kExpMaxChunks = 31;     // chunks per CTA: the per-thread row counters are bytes (8 points per chunk)

// 8 x 8 bit-matrix transpose of a 64-bit word (byte j = row j): byte r of the result holds bit r of every input byte
__device__ __forceinline__ uint64_t transpose8x8(uint64_t x) {
    uint64_t t;
    t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull;
    x = x ^ t ^ (t << 7);
    t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull;
    x = x ^ t ^ (t << 14);
    t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull;
    x = x ^ t ^ (t << 28);
    return x;
}
// per-byte population count of a 64-bit word
__device__ __forceinline__ uint64_t popcount_bytes(uint64_t x) {
    x = x - ((x >> 1) & 0x5555555555555555ull);
    x = (x & 0x3333333333333333ull) + ((x >> 2) & 0x3333333333333333ull);
    return (x + (x >> 4)) & 0x0F0F0F0F0F0F0F0Full;
}

__global__ void __launch_bounds__(kExpThreads, 4)
    sp_expand_mask_kernel(const float* __restrict__ mask_sig, const int64_t* __restrict__ superpoints, int K, int S,
                          int64_t N, float thr, uint8_t* __restrict__ out, int32_t* __restrict__ pointnum, int vec_io) {
    // s_word[s] bit r = (mask_sig[k0+r, s] > thr): ONE shared-memory lookup per point yields the bits of all 32 instance
    // rows this CTA writes. A thread owns 8 consecutive points: their 8 words are an 8 x 32 bit matrix whose transpose
    // (four 8 x 8 blocks, 64-bit SWAR) gives, per row, the 8 points' bits as ONE byte; s_lut turns that byte into the 8
    // output bytes (one LDS.64 + one 64-bit store per row instead of 8 extract / shift / or steps), and a per-byte
    // population count of the transposed blocks feeds the per-instance point counts (bytes in registers, reduced once
    // per CTA).
    extern __shared__ uint32_t s_word[];
    __shared__ uint2 s_lut[256];
    __shared__ int32_t s_cnt[kExpRows];
    const int k0 = blockIdx.y * kExpRows;
    const int rows = min(kExpRows, K - k0);
    if (rows == kExpRows) {  // full row group: 16 independent loads in flight per thread (the table build is latency-bound)
        for (int s = threadIdx.x; s < S; s += kExpThreads) {
            const float* col = mask_sig + (int64_t)k0 * S + s;
            uint32_t w = 0u;
#pragma unroll
            for (int r0 = 0; r0 < kExpRows; r0 += 16) {
                float v[16];
#pragma unroll
                for (int r = 0; r < 16; ++r) v[r] = __ldg(col + (int64_t)(r0 + r) * S);
#pragma unroll
                for (int r = 0; r < 16; ++r) w |= (v[r] > thr ? 1u : 0u) << (r0 + r);
            }
            s_word[s] = w;
        }
    } else {
        for (int s = threadIdx.x; s < S; s += kExpThreads) {
            uint32_t w = 0u;
            for (int r = 0; r < rows; ++r) w |= (__ldg(mask_sig + (int64_t)(k0 + r) * S + s) > thr ? 1u : 0u) << r;
            s_word[s] = w;
        }
    }
    {
        const uint32_t t = threadIdx.x;  // kExpThreads == 256: byte value -> its 8 bits as bytes
        s_lut[t] = make_uint2(((t & 15u) * 0x00204081u) & 0x01010101u, ((t >> 4) * 0x00204081u) & 0x01010101u);
    }
    if (threadIdx.x < kExpRows) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t n_chunks = ceil_div64(N, (int64_t)kExpThreads * kExpPtsPerThread);
    uint64_t cnt[4] = {0ull, 0ull, 0ull, 0ull};  // byte i of cnt[k]: this thread's points in row 8k + i (<= 8 per chunk)
    const bool n_vec = vec_io && (N & 7) == 0;  // 64-bit stores: rows and the base pointer are 8-byte aligned
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {  // the tables are built once per CTA
        const int64_t p0 = (chunk * kExpThreads + threadIdx.x) * kExpPtsPerThread;
        uint32_t pw[kExpPtsPerThread];
        if (vec_io && p0 + kExpPtsPerThread <= N) {
            const longlong2* ids = reinterpret_cast<const longlong2*>(superpoints + p0);  // 16-byte aligned: p0 % 8 == 0
#pragma unroll
            for (int j = 0; j < kExpPtsPerThread / 2; ++j) {
                const longlong2 id = __ldg(ids + j);
                pw[2 * j] = (id.x >= 0 && id.x < S) ? s_word[id.x] : 0u;  // ids outside [0,S) expand to False
                pw[2 * j + 1] = (id.y >= 0 && id.y < S) ? s_word[id.y] : 0u;
            }
        } else {
#pragma unroll
            for (int j = 0; j < kExpPtsPerThread; ++j) {
                const int64_t pt = p0 + j;
                const int64_t id = pt < N ? __ldg(superpoints + pt) : -1;
                pw[j] = (id >= 0 && id < S) ? s_word[id] : 0u;
            }
        }
        const bool vec_ok = n_vec && (p0 + kExpPtsPerThread <= N);
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // rows 8k .. 8k+7
            // byte k of the 8 words, point j in byte j
            const uint32_t lo = __byte_perm(__byte_perm(pw[0], pw[1], 0x0040 + 0x11 * k), __byte_perm(pw[2], pw[3], 0x0040 + 0x11 * k), 0x5410);
            const uint32_t hi = __byte_perm(__byte_perm(pw[4], pw[5], 0x0040 + 0x11 * k), __byte_perm(pw[6], pw[7], 0x0040 + 0x11 * k), 0x5410);
            const uint64_t t = transpose8x8(((uint64_t)hi << 32) | lo);  // byte i: bit j = point j of row 8k + i
            cnt[k] += popcount_bytes(t);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = 8 * k + i;
                if (r < rows) {
                    const uint2 bytes = s_lut[(uint32_t)(t >> (8 * i)) & 0xFFu];
                    uint8_t* dst = out + (int64_t)(k0 + r) * N + p0;
                    if (vec_ok) {
                        __stcs(reinterpret_cast<uint2*>(dst), bytes);
                    } else {
#pragma unroll
                        for (int j = 0; j < kExpPtsPerThread; ++j)
                            if (p0 + j < N) dst[j] = (uint8_t)(((j < 4 ? bytes.x : bytes.y) >> (8 * (j & 3))) & 1u);
                    }
                }
            }
        }
    }
    // per-instance point counts: bytes -> one REDUX per row per warp, one shared atomic per warp, one global per CTA
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int c = (int)((cnt[k] >> (8 * i)) & 0xFFull);
            c = __reduce_add_sync(kFull, c);
            if (lane_id() == 0 && c) atomicAdd(&s_cnt[8 * k + i], c);
        }
    }
    __syncthreads();
    if (threadIdx.x < rows && s_cnt[threadIdx.x]) atomicAdd(pointnum + k0 + threadIdx.x, s_cnt[threadIdx.x]);
}

// grad_src[p, :] = grad_out[idx[p], :] / max(n_idx[p], 1); seg sizes from seg_offsets (sd3d_sp_sort)
__global__ void __launch_bounds__(256)
    sp_mean_backward_kernel(const float* __restrict__ grad_out, const int64_t* __restrict__ idx,
                            const int32_t* __restrict__ seg_offsets, int64_t N, int32_t S, int C,
                            float* __restrict__ grad_src) {
    const int vec_per_row = C >> 2;  // vectorised path: C % 4 == 0
    const int64_t total = N * vec_per_row;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = g / vec_per_row;
        const int c = (int)(g % vec_per_row) * 4;
        const int64_t s = __ldg(idx + p);
        float4 v = f4_zero();
        if (s >= 0 && s < S) {
            const float n = (float)max(__ldg(seg_offsets + s + 1) - __ldg(seg_offsets + s), 1);
            v = f4_div(ldg_f4(grad_out + s * (int64_t)C + c), n);
        }
        st_cs_f4(grad_src + p * (int64_t)C + c, v);
    }
}

__global__ void __launch_bounds__(256)
    sp_mean_backward_scalar_kernel(const float* __restrict__ grad_out, const int64_t* __restrict__ idx,
                                   const int32_t* __restrict__ seg_offsets, int64_t N, int32_t S, int C,
                                   float* __restrict__ grad_src) {
    const int64_t total = N * C;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = g / C;
        const int c = (int)(g % C);
        const int64_t s = __ldg(idx + p);
        float v = 0.f;
        if (s >= 0 && s < S) {
            const float n = (float)max(__ldg(seg_offsets + s + 1) - __ldg(seg_offsets + s), 1);
            v = __fdiv_rn(__ldg(grad_out + s * (int64_t)C + c), n);
        }
        grad_src[g] = v;
    }
}

}  // namespace sd3d

using namespace sd3d;

extern "C" int sd3d_sp_expand_mask(const float* mask_sig, const int64_t* superpoints, int K, int64_t S, int64_t N,
                                   float thr, uint8_t* out, int32_t* pointnum, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (K < 0 || S < 0 || N < 0 || S > 50000) {
        set_error("sd3d_sp_expand_mask: bad shape K=%d S=%lld N=%lld (S <= 50000)", K, (long long)S, (long long)N);
        return SD3D_ERR_ARG;
    }
    if (K == 0) return SD3D_OK;
    if (pointnum == nullptr || (N > 0 && (out == nullptr || superpoints == nullptr)) || (S > 0 && mask_sig == nullptr)) {
        set_error("sd3d_sp_expand_mask: null buffer");
        return SD3D_ERR_ARG;
    }
    cudaMemsetAsync(pointnum, 0, (size_t)K * sizeof(int32_t), stream);
    if (N == 0) return check_launch("sd3d_sp_expand_mask(empty)");
    const size_t smem = sizeof(uint32_t) * (size_t)(S > 0 ? S : 1);
    if (smem > 200 * 1024) {
        set_error("sd3d_sp_expand_mask: S=%lld too large for the shared-memory mask tile", (long long)S);
        return SD3D_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(sp_expand_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   200 * 1024);
        if (e != cudaSuccess) {
            set_error("sd3d_sp_expand_mask: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return SD3D_ERR_CUDA;
        }
    }
    const int64_t n_chunks = ceil_div64(N, (int64_t)kExpThreads * kExpPtsPerThread);
    const int row_groups = (K + kExpRows - 1) / kExpRows;
    // ~4 CTAs per SM in total: every CTA builds its 32-row word table once and then loops over point chunks (at most
    // kExpMaxChunks of them: the per-thread row counters are bytes)
    const int64_t gx = imin64(n_chunks, imax64(imax64(1, (int64_t)4 * num_sms() / row_groups), ceil_div64(n_chunks, kExpMaxChunks)));
    dim3 grid((unsigned)gx, (unsigned)row_groups);
    const int vec_io = aligned16(superpoints) && (reinterpret_cast<uintptr_t>(out) & 7) == 0;
    sp_expand_mask_kernel<<<grid, kExpThreads, smem, stream>>>(mask_sig, superpoints, K, (int)S, N, thr, out, pointnum, vec_io);
    return check_launch("sd3d_sp_expand_mask");
}

extern "C" int sd3d_sp_mean_backward(const float* grad_out, const int64_t* idx, const int32_t* seg_offsets, int64_t N,
                                     int64_t S, int C, float* grad_src, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N < 0 || S < 0 || C <= 0 || S >= (int64_t(1) << 30)) {
        set_error("sd3d_sp_mean_backward: bad shape N=%lld S=%lld C=%d", (long long)N, (long long)S, C);
        return SD3D_ERR_ARG;
    }
    if (N == 0) return SD3D_OK;
    if (idx == nullptr || seg_offsets == nullptr || grad_src == nullptr || (S > 0 && grad_out == nullptr)) {
        set_error("sd3d_sp_mean_backward: null buffer");
        return SD3D_ERR_ARG;
    }
    const bool vec = (C % 4 == 0) && aligned16(grad_out) && aligned16(grad_src);
    const int64_t total = vec ? N * (C / 4) : N * (int64_t)C;
    const unsigned grid = (unsigned)imin64(ceil_div64(total, 256), (int64_t)num_sms() * 32);
    if (vec)
        sp_mean_backward_kernel<<<grid, 256, 0, stream>>>(grad_out, idx, seg_offsets, N, (int32_t)S, C, grad_src);
    else
        sp_mean_backward_scalar_kernel<<<grid, 256, 0, stream>>>(grad_out, idx, seg_offsets, N, (int32_t)S, C, grad_src);
    return check_launch("sd3d_sp_mean_backward");
}

// (3) sd3d_sp_label_vote: the superpoint-level ground truth of the dataset loaders
//     (/root/reference/segdino3d/datasets/dataset/scannet200.py:243-253, scannet.py:204-211):
//         onehot = F.one_hot(labels)[:, :K]                    # labels outside [0,K) (the background) drop out
//         sp = scatter_mean(onehot.float(), super_point_masks, dim=0) > 0.5
//         [semantic variant]  sp[sp.sum(-1) == 0, -1] = True
//     without the [N,K] one-hot tensor, its fp32 scatter and the [S,K] fp32 means: one warp per superpoint counts
//     its points' labels in a shared-memory histogram (integer: exact) and writes the boolean row. mean > 0.5 in
//     fp32 <=> 2 * count > size for superpoints below 2^23 points (the correctly rounded quotient of count / size
//     with 2 * count > size is at least 0.5 + 2^-24).
namespace sd3d {

constexpr int kVoteWarps = 4;
constexpr int kVoteBins = 1024;  // labels per pass of the histogram

__global__ void __launch_bounds__(kVoteWarps * 32)
    sp_label_vote_kernel(const int64_t* __restrict__ labels, const int32_t* __restrict__ perm,
                         const int32_t* __restrict__ seg_offsets, int32_t S, int K, int background_if_none,
                         uint8_t* __restrict__ out) {
    __shared__ int32_t s_hist[kVoteWarps][kVoteBins];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int s = blockIdx.x * kVoteWarps + warp;
    if (s >= S) return;
    const int begin = seg_offsets[s], end = seg_offsets[s + 1];
    const int n = end - begin;
    int32_t* hist = s_hist[warp];
    bool any_true = false;
    for (int k0 = 0; k0 < K; k0 += kVoteBins) {
        const int kb = min(kVoteBins, K - k0);
        for (int k = lane; k < kb; k += 32) hist[k] = 0;
        __syncwarp();
        for (int i = begin + lane; i < end; i += 32) {
            const int64_t l = __ldg(labels + perm[i]) - k0;
            if (l >= 0 && l < kb) atomicAdd(hist + (int)l, 1);
        }
        __syncwarp();
        for (int k = lane; k < kb; k += 32) {
            const bool t = 2 * hist[k] > n;
            any_true |= t;
            out[(int64_t)s * K + k0 + k] = t ? 1 : 0;
        }
        __syncwarp();
    }
    if (background_if_none && K > 0) {
        any_true = __any_sync(kFull, any_true);
        if (!any_true && lane == 0) out[(int64_t)s * K + K - 1] = 1;
    }
}

}  // namespace sd3d

extern "C" int sd3d_sp_label_vote(const int64_t* labels, const int32_t* perm, const int32_t* seg_offsets, int64_t N,
                                  int64_t S, int K, int background_if_none, uint8_t* out, void* stream_) {
    if (N < 0 || S < 0 || K < 0 || S >= (int64_t(1) << 30) || N >= (int64_t(1) << 31)) {
        set_error("sd3d_sp_label_vote: bad shape N=%lld S=%lld K=%d", (long long)N, (long long)S, K);
        return SD3D_ERR_ARG;
    }
    if (S == 0 || K == 0) return SD3D_OK;
    if (seg_offsets == nullptr || out == nullptr || (N > 0 && (labels == nullptr || perm == nullptr))) {
        set_error("sd3d_sp_label_vote: null buffer");
        return SD3D_ERR_ARG;
    }
    sp_label_vote_kernel<<<(unsigned)ceil_div64(S, kVoteWarps), kVoteWarps * 32, 0, (cudaStream_t)stream_>>>(
        labels, perm, seg_offsets, (int32_t)S, K, background_if_none, out);
    return check_launch("sd3d_sp_label_vote");
}
