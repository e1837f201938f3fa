// sp_mean.cu -- superpoint segmented mean: out[s,:] = sum_{p in s} src[p,:] / max(|s|,1).
//
// Drop-in arithmetic for torch_scatter.scatter_mean(src, index, dim=0) at
//   segdino3d/models/backbone/spconvunet.py:325,350,390,392 ; minkunet.py:639,641,653,674
// (torch-scatter 2.1.2: zeros.scatter_add_ + count clamp(min=1) + true_divide_). The reference's CUDA
// path issues one global fp32 atomicAdd per element (N*C atomics on S*C addresses, run-to-run
// non-deterministic); here the points are pre-sorted by superpoint (sp_sort.cu) and each (run, 128-channel
// slab) is reduced by one warp with coalesced 128-bit row reads and NO atomics:
//   EXACT: one warp walks the whole superpoint in ascending point index -> bit-identical to the aten CPU
//          scatter_add_ order; loads are issued 8 rows ahead, adds stay in order.
//   FAST : superpoints are split into runs of `run` rows (sp_tasks); partial rows are combined in run
//          order by sp_combine_rows_kernel -> deterministic, <= 1e-5 relative to the oracle.
// HBM-bound: algorithmic bytes = N*C*4 (src) + N*4 (perm) + S*C*4 (out).
#include "common.cuh"

namespace sd3d {

constexpr int kMeanThreads = 128;
constexpr int kMeanWarps = kMeanThreads / 32;
constexpr int kMeanUnroll = 8;

// vectorised path: C % 4 == 0, 16-byte aligned rows. One warp per (task, slab of 128 channels).
template <bool EXACT, bool HAS_COUNT>
__global__ void __launch_bounds__(kMeanThreads)
    sp_mean_vec_kernel(const float* __restrict__ src, const int32_t* __restrict__ perm,
                       const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ task_offsets,
                       const int32_t* __restrict__ task_seg, int run, int32_t S, int C, int nslabs,
                       const int32_t* __restrict__ point_count, float* __restrict__ dst) {
    const int lane = lane_id();
    const int64_t gw = (int64_t)blockIdx.x * kMeanWarps + (threadIdx.x >> 5);
    const int slab = (int)(gw % nslabs);
    const int64_t task = gw / nslabs;
    int64_t start, end;
    if (EXACT) {
        if (task >= S) return;
        start = seg_offsets[task];
        end = seg_offsets[task + 1];
    } else {
        if (task >= task_offsets[S + 1]) return;
        const int seg = task_seg[task];
        if (seg >= S) return;  // runs of the trash segment (invalid ids) are not pooled
        start = (int64_t)seg_offsets[seg] + (task - task_offsets[seg]) * (int64_t)run;
        end = min(start + (int64_t)run, (int64_t)seg_offsets[seg + 1]);
    }
    const int c = slab * 128 + lane * 4;
    const bool cok = c < C;
    float4 acc = f4_zero();
    for (int64_t i0 = start; i0 < end; i0 += 32) {
        const int nrow = (int)imin64(32, end - i0);
        int32_t my_p = 0, my_cnt = 1;
        if (lane < nrow) {
            my_p = perm[i0 + lane];
            if (HAS_COUNT) my_cnt = max(__ldg(point_count + my_p), 1);
        }
        for (int r0 = 0; r0 < nrow; r0 += kMeanUnroll) {
            float4 v[kMeanUnroll];
            float den[kMeanUnroll];
#pragma unroll
            for (int k = 0; k < kMeanUnroll; ++k) {
                const int r = r0 + k;
                const int32_t p = __shfl_sync(kFull, my_p, r & 31);
                den[k] = (float)__shfl_sync(kFull, my_cnt, r & 31);
                v[k] = (cok && r < nrow) ? ldg_f4(src + (int64_t)p * C + c) : f4_zero();
            }
#pragma unroll
            for (int k = 0; k < kMeanUnroll; ++k) {
                if (r0 + k < nrow) acc = f4_add(acc, HAS_COUNT ? f4_div(v[k], den[k]) : v[k]);
            }
        }
    }
    if (!cok) return;
    if (EXACT) {
        const int64_t n = end - start;
        *reinterpret_cast<float4*>(dst + task * (int64_t)C + c) = f4_div(acc, (float)imax64(n, 1));
    } else {
        *reinterpret_cast<float4*>(dst + task * (int64_t)C + c) = acc;
    }
}

// scalar path for any C (e.g. the C=3 superpoint-centre pooling, spconvunet.py:325): one warp per task,
// lanes stride over channels.
template <bool EXACT, bool HAS_COUNT>
__global__ void __launch_bounds__(kMeanThreads)
    sp_mean_scalar_kernel(const float* __restrict__ src, const int32_t* __restrict__ perm,
                          const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ task_offsets,
                          const int32_t* __restrict__ task_seg, int run, int32_t S, int C,
                          const int32_t* __restrict__ point_count, float* __restrict__ dst) {
    const int lane = lane_id();
    const int64_t task = (int64_t)blockIdx.x * kMeanWarps + (threadIdx.x >> 5);
    int64_t start, end;
    if (EXACT) {
        if (task >= S) return;
        start = seg_offsets[task];
        end = seg_offsets[task + 1];
    } else {
        if (task >= task_offsets[S + 1]) return;
        const int seg = task_seg[task];
        if (seg >= S) return;
        start = (int64_t)seg_offsets[seg] + (task - task_offsets[seg]) * (int64_t)run;
        end = min(start + (int64_t)run, (int64_t)seg_offsets[seg + 1]);
    }
    for (int c = lane; c < C; c += 32) {
        float acc = 0.f;
        for (int64_t i = start; i < end; ++i) {
            const int32_t p = perm[i];
            float v = __ldg(src + (int64_t)p * C + c);
            if (HAS_COUNT) v = __fdiv_rn(v, (float)max(__ldg(point_count + p), 1));
            acc = __fadd_rn(acc, v);
        }
        if (EXACT) acc = __fdiv_rn(acc, (float)imax64(end - start, 1));
        dst[task * (int64_t)C + c] = acc;
    }
}

// out[s,c] = (P[t0,c] + P[t0+1,c] + ...) / max(n_s,1), any C
__global__ void sp_combine_rows_kernel(const float* __restrict__ partials, const int32_t* __restrict__ task_offsets,
                                       const int32_t* __restrict__ seg_offsets, int32_t S, int C, int run,
                                       float* __restrict__ out) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)S * C) return;
    const int s = (int)(gid / C);
    const int c = (int)(gid % C);
    const int n = seg_offsets[s + 1] - seg_offsets[s];
    const int t0 = task_offsets[s], t1 = t0 + (n + run - 1) / run;
    float acc = 0.f;
    for (int t = t0; t < t1; ++t) acc = __fadd_rn(acc, partials[(int64_t)t * C + c]);
    out[gid] = __fdiv_rn(acc, (float)max(n, 1));
}

}  // namespace sd3d

using namespace sd3d;

extern "C" int sd3d_sp_mean(const float* src, const int32_t* perm, const int32_t* seg_offsets, int64_t N, int64_t S,
                            int C, const int32_t* point_count, int mode, const int32_t* task_offsets,
                            const int32_t* task_seg, int64_t max_tasks, int run, void* ws, size_t ws_bytes, float* out,
                            void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N < 0 || S < 0 || C <= 0 || S >= (int64_t(1) << 30) || N >= (int64_t(1) << 31) - 64) {
        set_error("sd3d_sp_mean: bad shape N=%lld S=%lld C=%d", (long long)N, (long long)S, C);
        return SD3D_ERR_ARG;
    }
    if (S == 0) return SD3D_OK;
    if (out == nullptr || seg_offsets == nullptr || (N > 0 && (src == nullptr || perm == nullptr))) {
        set_error("sd3d_sp_mean: null buffer");
        return SD3D_ERR_ARG;
    }
    if (mode != SD3D_POOL_FAST && mode != SD3D_POOL_EXACT) {
        set_error("sd3d_sp_mean: mode %d unknown", mode);
        return SD3D_ERR_ARG;
    }
    const bool exact = mode == SD3D_POOL_EXACT;
    float* partials = reinterpret_cast<float*>(ws);
    if (!exact) {
        if (run <= 0 || task_offsets == nullptr || task_seg == nullptr || ws == nullptr ||
            max_tasks < sd3d_sp_max_tasks(N, S, run) || ws_bytes < (size_t)max_tasks * C * sizeof(float) ||
            !aligned16(ws)) {
            set_error("sd3d_sp_mean: FAST mode needs run>0, task tables and ws >= max_tasks*C*4 bytes");
            return SD3D_ERR_ARG;
        }
    }
    const bool vec = (C % 4 == 0) && aligned16(src) && aligned16(out);
    const bool has_count = point_count != nullptr;
    const int64_t n_tasks = exact ? S : max_tasks;
    float* dst = exact ? out : partials;
    if (vec) {
        const int nslabs = (C + 127) / 128;
        const unsigned grid = (unsigned)ceil_div64(n_tasks * nslabs, kMeanWarps);
#define SD3D_LAUNCH_VEC(E, H)                                                                                       \
    sp_mean_vec_kernel<E, H><<<grid, kMeanThreads, 0, stream>>>(src, perm, seg_offsets, task_offsets, task_seg, run, \
                                                                (int32_t)S, C, nslabs, point_count, dst)
        if (exact && has_count) SD3D_LAUNCH_VEC(true, true);
        else if (exact) SD3D_LAUNCH_VEC(true, false);
        else if (has_count) SD3D_LAUNCH_VEC(false, true);
        else SD3D_LAUNCH_VEC(false, false);
#undef SD3D_LAUNCH_VEC
    } else {
        const unsigned grid = (unsigned)ceil_div64(n_tasks, kMeanWarps);
#define SD3D_LAUNCH_SCALAR(E, H)                                                                                \
    sp_mean_scalar_kernel<E, H><<<grid, kMeanThreads, 0, stream>>>(src, perm, seg_offsets, task_offsets, task_seg, \
                                                                   run, (int32_t)S, C, point_count, dst)
        if (exact && has_count) SD3D_LAUNCH_SCALAR(true, true);
        else if (exact) SD3D_LAUNCH_SCALAR(true, false);
        else if (has_count) SD3D_LAUNCH_SCALAR(false, true);
        else SD3D_LAUNCH_SCALAR(false, false);
#undef SD3D_LAUNCH_SCALAR
    }
    if (!exact) {
        const int64_t threads = S * (int64_t)C;
        sp_combine_rows_kernel<<<(unsigned)ceil_div64(threads, 256), 256, 0, stream>>>(partials, task_offsets,
                                                                                      seg_offsets, (int32_t)S, C, run, out);
    }
    return check_launch("sd3d_sp_mean");
}
