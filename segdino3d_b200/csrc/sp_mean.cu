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
// Three EXACT kernels, same summation order: sp_mean_cta_kernel (one CTA per superpoint, cp.async ring: sizeable
// superpoints, narrow rows or few superpoints), sp_mean_small_kernel (several narrow rows per warp load), and the
// warp-per-(superpoint, slab) kernel below (wide rows, thousands of superpoints: at the HBM copy peak).
#include <atomic>

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

// narrow rows (the reference's live pooling widths: C = 32 backbone features, C = 3 coordinates; spconvunet.py:390,325):
// a row occupies only L = C/4 (float4) or C (scalar) lanes, so G = 32 / L rows are loaded per instruction (4 x G rows in
// flight per warp instead of 8 rows on a quarter of the lanes) and then added in ascending row order through a shuffle
// chain -- the summation order, hence every bit of the result, is the one of the wide kernel and of the CPU reference.
template <typename V>
struct RowVec;
template <>
struct RowVec<float4> {
    static constexpr int kWidth = 4;
    static __device__ __forceinline__ float4 zero() { return f4_zero(); }
    static __device__ __forceinline__ float4 load(const float* p) { return ldg_f4(p); }
    static __device__ __forceinline__ float4 add(float4 a, float4 b) { return f4_add(a, b); }
    static __device__ __forceinline__ float4 div(float4 a, float d) { return f4_div(a, d); }
    static __device__ __forceinline__ float4 shfl(float4 a, int src) {
        return make_float4(__shfl_sync(kFull, a.x, src), __shfl_sync(kFull, a.y, src), __shfl_sync(kFull, a.z, src),
                           __shfl_sync(kFull, a.w, src));
    }
    static __device__ __forceinline__ void store(float* p, float4 a) { *reinterpret_cast<float4*>(p) = a; }
};
template <>
struct RowVec<float> {
    static constexpr int kWidth = 1;
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ float load(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float d) { return __fdiv_rn(a, d); }
    static __device__ __forceinline__ float shfl(float a, int src) { return __shfl_sync(kFull, a, src); }
    static __device__ __forceinline__ void store(float* p, float a) { *p = a; }
};

constexpr int kSmallUnroll = 4;

template <typename V, bool EXACT, bool HAS_COUNT>
__global__ void __launch_bounds__(kMeanThreads)
    sp_mean_small_kernel(const float* __restrict__ src, const int32_t* __restrict__ perm,
                         const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ task_offsets,
                         const int32_t* __restrict__ task_seg, int run, int32_t S, int C, int L, int G,
                         const int32_t* __restrict__ point_count, float* __restrict__ dst) {
    using RV = RowVec<V>;
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
    const int g = lane / L, j = lane - g * L;  // row slot of this lane, position inside the row
    V acc = RV::zero();
    for (int64_t i0 = start; i0 < end; i0 += 32) {
        const int nrow = (int)imin64(32, end - i0);
        int32_t my_p = 0, my_cnt = 1;
        if (lane < nrow) {
            my_p = perm[i0 + lane];
            if (HAS_COUNT) my_cnt = max(__ldg(point_count + my_p), 1);
        }
        for (int r0 = 0; r0 < nrow; r0 += G * kSmallUnroll) {
            V v[kSmallUnroll];
#pragma unroll
            for (int u = 0; u < kSmallUnroll; ++u) {
                const int r = r0 + u * G + g;
                const int32_t p = __shfl_sync(kFull, my_p, r & 31);
                const float den = (float)__shfl_sync(kFull, my_cnt, r & 31);
                v[u] = (g < G && r < nrow) ? RV::load(src + (int64_t)p * C + j * RV::kWidth) : RV::zero();
                if (HAS_COUNT) v[u] = RV::div(v[u], den);
            }
#pragma unroll
            for (int u = 0; u < kSmallUnroll; ++u) {
                for (int k = 0; k < G; ++k) {  // rows r0 + u*G + k in ascending order; the bound is warp-uniform
                    const V x = RV::shfl(v[u], k * L + j);
                    if (r0 + u * G + k < nrow) acc = RV::add(acc, x);
                }
            }
        }
    }
    if (lane >= L) return;
    if (EXACT) acc = RV::div(acc, (float)imax64(end - start, 1));
    RV::store(dst + task * (int64_t)C + j * RV::kWidth, acc);
}

// EXACT mode on realistically sized superpoints (hundreds to thousands of rows each, a few room-sized ones): the adds of a
// superpoint are an inherently serial fp32 chain, but its LOADS are not. One CTA per superpoint: all 256 threads stream
// the rows (in perm order) into a 3-stage shared-memory ring with cp.async (up to 64 KB in flight per CTA instead of the
// 8 rows one warp keeps in flight), and C/4 (or C) threads add each landed stage row by row in ascending order from
// shared memory -- the same sequence of rounded adds as the warp kernels and the CPU reference, so the result is
// bit-identical; the largest superpoint no longer walks its rows at one warp's latency-bound rate.
constexpr int kCtaThreads = 256;
constexpr int kCtaStages = 3;
constexpr int kCtaStageBytes = 32768;
constexpr int kCtaMaxRows = 512;  // rows per stage (narrow rows)

template <bool VEC>
__device__ __forceinline__ void cta_copy(uint32_t dst, const float* src) {
    if (VEC) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

template <bool VEC, bool HAS_COUNT>
__global__ void __launch_bounds__(kCtaThreads)
    sp_mean_cta_kernel(const float* __restrict__ src, const int32_t* __restrict__ perm,
                       const int32_t* __restrict__ seg_offsets, int C, int rows_per_stage,
                       const int32_t* __restrict__ point_count, float* __restrict__ dst) {
    extern __shared__ __align__(16) uint8_t cta_smem[];
    constexpr int kSlots = kCtaStages + 1;  // row-id / divisor slots run one chunk ahead of the data ring
    int32_t* s_perm = reinterpret_cast<int32_t*>(cta_smem + (size_t)kCtaStages * kCtaStageBytes);  // [kSlots][kCtaMaxRows]
    float* s_den = reinterpret_cast<float*>(s_perm + kSlots * kCtaMaxRows);                         // [kSlots][kCtaMaxRows]
    const int task = blockIdx.x, tid = threadIdx.x;
    const int64_t start = seg_offsets[task], end = seg_offsets[task + 1];
    const int n = (int)(end - start);
    constexpr int W = VEC ? 4 : 1;
    const int epr = C / W;  // copy / accumulate elements per row
    const int nchunks = (n + rows_per_stage - 1) / rows_per_stage;
    const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(cta_smem);

    // row ids (and 1 / count divisors) of chunk k -> slot k % kSlots; visible to the CTA after the next __syncthreads
    auto fetch_rows = [&](int k) {
        if (k >= nchunks) return;
        const int64_t row0 = start + (int64_t)k * rows_per_stage;
        const int rows = (int)imin64(rows_per_stage, end - row0);
        for (int r = tid; r < rows; r += kCtaThreads) {
            const int32_t p = __ldg(perm + row0 + r);
            s_perm[(k % kSlots) * kCtaMaxRows + r] = p;
            if (HAS_COUNT) s_den[(k % kSlots) * kCtaMaxRows + r] = (float)max(__ldg(point_count + p), 1);
        }
    };
    // rows of chunk k -> data stage k % kCtaStages (cp.async; the row ids come from shared memory: no dependent global load)
    auto issue = [&](int k) {
        if (k < nchunks) {
            const int rows = min(rows_per_stage, n - k * rows_per_stage);
            const uint32_t sdst = smem_base + (uint32_t)(k % kCtaStages) * kCtaStageBytes;
            const int32_t* rowid = s_perm + (k % kSlots) * kCtaMaxRows;
            for (int i = tid; i < rows * epr; i += kCtaThreads) {
                const int row = i / epr, j = i - row * epr;
                cta_copy<VEC>(sdst + (uint32_t)i * (W * 4), src + (int64_t)rowid[row] * C + j * W);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // (an empty group keeps the wait count uniform)
    };

    float4 acc4 = f4_zero();
    float acc1 = 0.f;
    for (int k = 0; k < kCtaStages; ++k) fetch_rows(k);
    __syncthreads();
    for (int k = 0; k < kCtaStages - 1; ++k) issue(k);
    for (int k = 0; k < nchunks; ++k) {
        issue(k + kCtaStages - 1);       // its row ids were fetched one iteration (and one barrier) ago
        fetch_rows(k + kCtaStages);      // slot (k + 3) % 4 = slot of chunk k - 1, consumed before the last barrier
        asm volatile("cp.async.wait_group %0;" ::"n"(kCtaStages - 1) : "memory");
        __syncthreads();  // every thread's copies of chunk k are visible
        const int rows = min(rows_per_stage, n - k * rows_per_stage);
        if (tid < epr) {
            const uint8_t* base = cta_smem + (size_t)(k % kCtaStages) * kCtaStageBytes;
            const float* den = s_den + (k % kSlots) * kCtaMaxRows;
            if (VEC) {
                const float4* rowp = reinterpret_cast<const float4*>(base) + tid;
#pragma unroll 4
                for (int r = 0; r < rows; ++r) {
                    float4 v = rowp[(size_t)r * epr];
                    if (HAS_COUNT) v = f4_div(v, den[r]);
                    acc4 = f4_add(acc4, v);
                }
            } else {
                const float* rowp = reinterpret_cast<const float*>(base) + tid;
#pragma unroll 4
                for (int r = 0; r < rows; ++r) {
                    float v = rowp[(size_t)r * epr];
                    if (HAS_COUNT) v = __fdiv_rn(v, den[r]);
                    acc1 = __fadd_rn(acc1, v);
                }
            }
        }
        __syncthreads();  // the data stage and the row-id slot of chunk k may be refilled
    }
    if (tid < epr) {
        const float d = (float)max(n, 1);
        if (VEC) *reinterpret_cast<float4*>(dst + (int64_t)task * C + tid * 4) = f4_div(acc4, d);
        else dst[(int64_t)task * C + tid] = __fdiv_rn(acc1, d);
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
    const int small_l = vec ? C / 4 : C;  // lanes one row occupies
    // EXACT on sizeable superpoints (>= 64 rows on average), rows of at most 256 accumulating threads and 32 KB per
    // stage: one CTA per superpoint with a cp.async ring
    // (wide rows with thousands of superpoints already fill the machine with one warp per (superpoint, slab): 1 M x 256 /
    // 5000 runs at the HBM copy peak that way, 1.6x faster than with CTAs)
    const bool cta_pays = (int64_t)C * 4 <= 512 || S * (int64_t)((C + 127) / 128) <= 2400;
    if (exact && cta_pays && N >= 64 * S && small_l <= kCtaThreads && (int64_t)C * 4 <= kCtaStageBytes / 8) {
        const int rows_per_stage = (int)imin64(kCtaMaxRows, kCtaStageBytes / (C * 4));
        const size_t smem = (size_t)kCtaStages * kCtaStageBytes + (size_t)2 * (kCtaStages + 1) * kCtaMaxRows * sizeof(float);
        static std::atomic<uint64_t> cta_attr{0};
        if (first_on_device(&cta_attr)) {
            cudaError_t e = cudaFuncSetAttribute(sp_mean_cta_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(sp_mean_cta_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(sp_mean_cta_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(sp_mean_cta_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) {
                cta_attr.store(0);
                set_error("sd3d_sp_mean: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                return SD3D_ERR_CUDA;
            }
        }
        const unsigned grid = (unsigned)S;
        if (vec && has_count) sp_mean_cta_kernel<true, true><<<grid, kCtaThreads, smem, stream>>>(src, perm, seg_offsets, C, rows_per_stage, point_count, out);
        else if (vec) sp_mean_cta_kernel<true, false><<<grid, kCtaThreads, smem, stream>>>(src, perm, seg_offsets, C, rows_per_stage, point_count, out);
        else if (has_count) sp_mean_cta_kernel<false, true><<<grid, kCtaThreads, smem, stream>>>(src, perm, seg_offsets, C, rows_per_stage, point_count, out);
        else sp_mean_cta_kernel<false, false><<<grid, kCtaThreads, smem, stream>>>(src, perm, seg_offsets, C, rows_per_stage, point_count, out);
    } else if (small_l <= 16) {
        const int G = 32 / small_l;
        const unsigned grid = (unsigned)ceil_div64(n_tasks, kMeanWarps);
#define SD3D_LAUNCH_SMALL(V, E, H)                                                                                    \
    sp_mean_small_kernel<V, E, H><<<grid, kMeanThreads, 0, stream>>>(src, perm, seg_offsets, task_offsets, task_seg, run, \
                                                                     (int32_t)S, C, small_l, G, point_count, dst)
        if (vec) {
            if (exact && has_count) SD3D_LAUNCH_SMALL(float4, true, true);
            else if (exact) SD3D_LAUNCH_SMALL(float4, true, false);
            else if (has_count) SD3D_LAUNCH_SMALL(float4, false, true);
            else SD3D_LAUNCH_SMALL(float4, false, false);
        } else {
            if (exact && has_count) SD3D_LAUNCH_SMALL(float, true, true);
            else if (exact) SD3D_LAUNCH_SMALL(float, true, false);
            else if (has_count) SD3D_LAUNCH_SMALL(float, false, true);
            else SD3D_LAUNCH_SMALL(float, false, false);
        }
#undef SD3D_LAUNCH_SMALL
    } else if (vec) {
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
