// lift.cu -- fused projection + depth-visibility + bilinear gather + view mean (+ superpoint run partials).
//
// Steps a-1..a-3 of SURVEY.md section 8(a), frozen spec = SURVEY.md Appendix A. The reference ships no
// code for these steps (features are loaded precomputed: segdino3d/datasets/dataset/scannet200.py:219-226),
// so the arithmetic below follows Appendix A op for op: every mul/add/div is a separately rounded fp32
// operation (__fmul_rn/__fadd_rn/__fdiv_rn are never contracted into FMA), views are summed in ascending
// order per point -> pix_idx / vis / count AND the fp32 sums are bit-identical to the oracle.
//
// Mapping (HBM/L2-bound gather, no tensor cores):
//   * one warp owns a run of <= `run` consecutive points of the processing order (points sorted by
//     superpoint => spatially coherent => neighbouring samples hit the same feature-map rows in L1/L2);
//   * points are handled G at a time; for each chunk of 32 views LANE = VIEW computes projection,
//     depth lookup and the visibility predicate; __ballot_sync turns that into per-point view masks;
//   * the warp then walks the set bits in ascending view order: LANE = CHANNEL VECTOR, each of the 4
//     bilinear taps is a fully coalesced row read (128-bit per lane, 512 B per request);
//   * per-point accumulators stay in registers across all views; one coalesced streaming store per point;
//   * optional fused superpoint pooling: the finalised rows of the run are summed in registers and written
//     as ONE partial row per run (no atomics); sp_combine_kernel adds the partials in run order.
#include "common.cuh"

namespace sd3d {

constexpr int kLiftThreads = 128;
constexpr int kLiftWarps = kLiftThreads / 32;

struct LiftParams {
    const float* xyz;
    int64_t N;
    const float* K4;
    const float* w2c;
    int v_begin, v_end;
    const void* depth;
    int depth_u16;
    int Hd, Wd;
    const void* fmap;
    int Hf, Wf, C;
    float stride, tau, z_near;
    int accumulate, finalize;
    const int32_t* order;
    float* out;
    int32_t* count;
    int32_t* pix_idx;
    uint8_t* vis;
    // plan (pool != 0)
    int pool;
    const int32_t* seg_offsets;
    const int32_t* task_offsets;
    const int32_t* task_seg;
    int32_t S;
    int run;
    float* partials;
};

template <typename FT>
__device__ __forceinline__ float4 load_tap(const FT* p);
template <>
__device__ __forceinline__ float4 load_tap<float>(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 load_tap<__half>(const __half* p) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
    const __half2 a = *reinterpret_cast<const __half2*>(&raw.x);
    const __half2 b = *reinterpret_cast<const __half2*>(&raw.y);
    const float2 fa = __half22float2(a), fb = __half22float2(b);
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <>
__device__ __forceinline__ float4 load_tap<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
    return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u),
                       __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
}

// f = ((w00*t00 + w01*t01) + w10*t10) + w11*t11 ; acc = acc + f      (Appendix A, unfused)
__device__ __forceinline__ float blend1(float acc, float w00, float w01, float w10, float w11, float t00, float t01,
                                        float t10, float t11) {
    float f = __fadd_rn(__fmul_rn(w00, t00), __fmul_rn(w01, t01));
    f = __fadd_rn(f, __fmul_rn(w10, t10));
    f = __fadd_rn(f, __fmul_rn(w11, t11));
    return __fadd_rn(acc, f);
}

template <int NV, typename FT>
__device__ __forceinline__ void gather_sample(float4 (&acc)[NV], const FT* __restrict__ fmap_v, float u, float w,
                                              float stride, int Hf, int Wf, int C, int lane) {
    const float uf = __fsub_rn(__fdiv_rn(__fadd_rn(u, 0.5f), stride), 0.5f);
    const float wf = __fsub_rn(__fdiv_rn(__fadd_rn(w, 0.5f), stride), 0.5f);
    const float x0f = floorf(uf), y0f = floorf(wf);
    const float ax = __fsub_rn(uf, x0f), ay = __fsub_rn(wf, y0f);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float omx = __fsub_rn(1.0f, ax), omy = __fsub_rn(1.0f, ay);
    const float w00 = __fmul_rn(omx, omy), w01 = __fmul_rn(ax, omy);
    const float w10 = __fmul_rn(omx, ay), w11 = __fmul_rn(ax, ay);
    const bool okx0 = (x0 >= 0) && (x0 < Wf), okx1 = (x0 + 1 >= 0) && (x0 + 1 < Wf);
    const bool oky0 = (y0 >= 0) && (y0 < Hf), oky1 = (y0 + 1 >= 0) && (y0 + 1 < Hf);
    const bool ok00 = oky0 && okx0, ok01 = oky0 && okx1, ok10 = oky1 && okx0, ok11 = oky1 && okx1;
    const int64_t o00 = ((int64_t)y0 * Wf + x0) * C;
    const int64_t o01 = o00 + C, o10 = o00 + (int64_t)Wf * C, o11 = o10 + C;
    float4 t00[NV], t01[NV], t10[NV], t11[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int c = (k * 32 + lane) * 4;
        const bool cok = c < C;
        t00[k] = (cok && ok00) ? load_tap<FT>(fmap_v + o00 + c) : f4_zero();
        t01[k] = (cok && ok01) ? load_tap<FT>(fmap_v + o01 + c) : f4_zero();
        t10[k] = (cok && ok10) ? load_tap<FT>(fmap_v + o10 + c) : f4_zero();
        t11[k] = (cok && ok11) ? load_tap<FT>(fmap_v + o11 + c) : f4_zero();
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        acc[k].x = blend1(acc[k].x, w00, w01, w10, w11, t00[k].x, t01[k].x, t10[k].x, t11[k].x);
        acc[k].y = blend1(acc[k].y, w00, w01, w10, w11, t00[k].y, t01[k].y, t10[k].y, t11[k].y);
        acc[k].z = blend1(acc[k].z, w00, w01, w10, w11, t00[k].z, t01[k].z, t10[k].z, t11[k].z);
        acc[k].w = blend1(acc[k].w, w00, w01, w10, w11, t00[k].w, t01[k].w, t10[k].w, t11[k].w);
    }
}

template <int NV, typename FT, int G>
__global__ void __launch_bounds__(kLiftThreads) lift_kernel(const LiftParams p) {
    const int lane = lane_id();
    const int64_t task = (int64_t)blockIdx.x * kLiftWarps + (threadIdx.x >> 5);
    int64_t start, end;
    int seg = -1;
    if (p.pool) {
        const int32_t n_tasks = p.task_offsets[p.S + 1];
        if (task >= n_tasks) return;
        seg = p.task_seg[task];
        start = (int64_t)p.seg_offsets[seg] + (task - p.task_offsets[seg]) * (int64_t)p.run;
        end = min(start + (int64_t)p.run, (int64_t)p.seg_offsets[seg + 1]);
    } else {
        start = task * (int64_t)p.run;
        if (start >= p.N) return;
        end = min(start + (int64_t)p.run, p.N);
    }
    const FT* __restrict__ fmap = reinterpret_cast<const FT*>(p.fmap);
    const int64_t view_elems = (int64_t)p.Hf * p.Wf * p.C;
    const int64_t depth_elems = (int64_t)p.Hd * p.Wd;
    const float wd_f = (float)p.Wd, hd_f = (float)p.Hd;

    float4 sp_acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) sp_acc[k] = f4_zero();

    for (int64_t g0 = start; g0 < end; g0 += G) {
        int32_t pid[G];
        float px[G], py[G], pz[G];
        float4 acc[G][NV];
        int cnt[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            pid[j] = -1;
            px[j] = py[j] = pz[j] = 0.f;
            cnt[j] = 0;
            if (g0 + j < end) {
                pid[j] = p.order ? p.order[g0 + j] : (int32_t)(g0 + j);
                px[j] = __ldg(p.xyz + 3 * (int64_t)pid[j]);
                py[j] = __ldg(p.xyz + 3 * (int64_t)pid[j] + 1);
                pz[j] = __ldg(p.xyz + 3 * (int64_t)pid[j] + 2);
            }
#pragma unroll
            for (int k = 0; k < NV; ++k) acc[j][k] = f4_zero();
            if (p.accumulate && pid[j] >= 0) {
                cnt[j] = p.count[pid[j]];
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    const int c = (k * 32 + lane) * 4;
                    if (c < p.C) acc[j][k] = *reinterpret_cast<const float4*>(p.out + (int64_t)pid[j] * p.C + c);
                }
            }
        }

        for (int v0 = p.v_begin; v0 < p.v_end; v0 += 32) {
            const int v = v0 + lane;
            const bool vok = v < p.v_end;
            float4 k4 = f4_zero(), r0 = f4_zero(), r1 = f4_zero(), r2 = f4_zero();
            if (vok) {
                k4 = ldg_f4(p.K4 + 4 * (int64_t)v);
                r0 = ldg_f4(p.w2c + 12 * (int64_t)v);
                r1 = ldg_f4(p.w2c + 12 * (int64_t)v + 4);
                r2 = ldg_f4(p.w2c + 12 * (int64_t)v + 8);
            }
            float us[G], ws[G];
            unsigned mask[G];
            unsigned any = 0u;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                bool visible = false;
                float uu = 0.f, ww = 0.f;
                int pix = -1;
                if (vok && pid[j] >= 0) {
                    const float xc = __fadd_rn(
                        __fadd_rn(__fadd_rn(__fmul_rn(r0.x, px[j]), __fmul_rn(r0.y, py[j])), __fmul_rn(r0.z, pz[j])),
                        r0.w);
                    const float yc = __fadd_rn(
                        __fadd_rn(__fadd_rn(__fmul_rn(r1.x, px[j]), __fmul_rn(r1.y, py[j])), __fmul_rn(r1.z, pz[j])),
                        r1.w);
                    const float zc = __fadd_rn(
                        __fadd_rn(__fadd_rn(__fmul_rn(r2.x, px[j]), __fmul_rn(r2.y, py[j])), __fmul_rn(r2.z, pz[j])),
                        r2.w);
                    if (zc > p.z_near) {
                        uu = __fadd_rn(__fdiv_rn(__fmul_rn(k4.x, xc), zc), k4.z);
                        ww = __fadd_rn(__fdiv_rn(__fmul_rn(k4.y, yc), zc), k4.w);
                        const float uif = floorf(__fadd_rn(uu, 0.5f));
                        const float wif = floorf(__fadd_rn(ww, 0.5f));
                        if (uif >= 0.f && uif < wd_f && wif >= 0.f && wif < hd_f) {
                            const int cand = (int)wif * p.Wd + (int)uif;
                            float d;
                            if (p.depth_u16)
                                d = __fmul_rn((float)__ldg(reinterpret_cast<const uint16_t*>(p.depth) +
                                                           (int64_t)v * depth_elems + cand),
                                              0.001f);
                            else
                                d = __ldg(reinterpret_cast<const float*>(p.depth) + (int64_t)v * depth_elems + cand);
                            if (d > 0.f && fabsf(__fsub_rn(d, zc)) <= p.tau) {
                                visible = true;
                                pix = cand;
                            }
                        }
                    }
                    if (p.pix_idx) p.pix_idx[(int64_t)v * p.N + pid[j]] = pix;
                    if (p.vis) p.vis[(int64_t)v * p.N + pid[j]] = visible ? 1 : 0;
                }
                us[j] = uu;
                ws[j] = ww;
                mask[j] = __ballot_sync(kFull, visible);
                any |= mask[j];
            }
            while (any) {
                const int b = __ffs(any) - 1;
                any &= any - 1u;
                const FT* __restrict__ fmap_v = fmap + (int64_t)(v0 + b) * view_elems;
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    if ((mask[j] >> b) & 1u) {
                        const float uu = __shfl_sync(kFull, us[j], b);
                        const float ww = __shfl_sync(kFull, ws[j], b);
                        gather_sample<NV, FT>(acc[j], fmap_v, uu, ww, p.stride, p.Hf, p.Wf, p.C, lane);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < G; ++j) cnt[j] += __popc(mask[j]);
        }

#pragma unroll
        for (int j = 0; j < G; ++j) {
            if (pid[j] >= 0) {
                const float denom = (float)max(cnt[j], 1);
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    const int c = (k * 32 + lane) * 4;
                    if (c < p.C) {
                        const float4 o = p.finalize ? f4_div(acc[j][k], denom) : acc[j][k];
                        st_cs_f4(p.out + (int64_t)pid[j] * p.C + c, o);
                        sp_acc[k] = f4_add(sp_acc[k], o);
                    }
                }
                if (lane == 0) p.count[pid[j]] = cnt[j];
            }
        }
    }
    if (p.pool && seg < p.S) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = (k * 32 + lane) * 4;
            if (c < p.C) *reinterpret_cast<float4*>(p.partials + task * (int64_t)p.C + c) = sp_acc[k];
        }
    }
}

// out[s,:] = (P[t0] + P[t0+1] + ... in task order) / max(n_s,1)
__global__ void sp_combine_kernel(const float* __restrict__ partials, const int32_t* __restrict__ task_offsets,
                                  const int32_t* __restrict__ seg_offsets, int32_t S, int C,
                                  float* __restrict__ out) {
    const int vec_per_row = C >> 2;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)S * vec_per_row) return;
    const int s = (int)(gid / vec_per_row);
    const int c = (int)(gid % vec_per_row) * 4;
    const int t0 = task_offsets[s], t1 = task_offsets[s + 1];
    float4 acc = f4_zero();
    int t = t0;
    for (; t + 4 <= t1; t += 4) {
        const float4 a = *reinterpret_cast<const float4*>(partials + (int64_t)t * C + c);
        const float4 b = *reinterpret_cast<const float4*>(partials + (int64_t)(t + 1) * C + c);
        const float4 d = *reinterpret_cast<const float4*>(partials + (int64_t)(t + 2) * C + c);
        const float4 e = *reinterpret_cast<const float4*>(partials + (int64_t)(t + 3) * C + c);
        acc = f4_add(f4_add(f4_add(f4_add(acc, a), b), d), e);
    }
    for (; t < t1; ++t) acc = f4_add(acc, *reinterpret_cast<const float4*>(partials + (int64_t)t * C + c));
    const int n = seg_offsets[s + 1] - seg_offsets[s];
    *reinterpret_cast<float4*>(out + (int64_t)s * C + c) = f4_div(acc, (float)max(n, 1));
}

__global__ void finalize_kernel(float* __restrict__ sum, const int32_t* __restrict__ count, int64_t N, int C) {
    const int vec_per_row = C >> 2;
    const int64_t total = N * vec_per_row;
    for (int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
         gid += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = gid / vec_per_row;
        const float denom = (float)max(__ldg(count + row), 1);
        float4* ptr = reinterpret_cast<float4*>(sum) + gid;
        *ptr = f4_div(*ptr, denom);
    }
}

constexpr int kMaxScales = 8;
struct ScalePtrs {
    const float* p[kMaxScales];
};
// torch.stack(list).mean(0): sequential sum over scales then divide by L (scannet200.py:233-234)
__global__ void scale_mean_kernel(ScalePtrs ptrs, int L, int64_t numel, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
        float a = ptrs.p[0][i];
        for (int l = 1; l < L; ++l) a = __fadd_rn(a, ptrs.p[l][i]);
        out[i] = __fdiv_rn(a, (float)L);
    }
}

template <int NV, typename FT, int G>
static void launch_lift(const LiftParams& p, int64_t n_tasks, cudaStream_t stream) {
    const unsigned grid = (unsigned)ceil_div64(n_tasks, kLiftWarps);
    lift_kernel<NV, FT, G><<<grid, kLiftThreads, 0, stream>>>(p);
}

template <typename FT>
static int dispatch_lift(const LiftParams& p, int64_t n_tasks, int variant, cudaStream_t stream) {
    const int nv = (p.C + 127) / 128;
    if (nv == 1) {
        launch_lift<1, FT, 8>(p, n_tasks, stream);
    } else if (nv == 2) {
        if constexpr (sizeof(FT) == 4) {
            switch (variant) {
                case 1: launch_lift<2, FT, 1>(p, n_tasks, stream); break;
                case 2: launch_lift<2, FT, 2>(p, n_tasks, stream); break;
                case 8: launch_lift<2, FT, 8>(p, n_tasks, stream); break;
                default: launch_lift<2, FT, 4>(p, n_tasks, stream); break;
            }
        } else {
            launch_lift<2, FT, 4>(p, n_tasks, stream);
        }
    } else if (nv <= 4) {
        launch_lift<4, FT, 2>(p, n_tasks, stream);
    } else if (nv <= 8) {
        launch_lift<8, FT, 1>(p, n_tasks, stream);
    } else {
        set_error("sd3d_lift: C=%d > 1024 unsupported", p.C);
        return SD3D_ERR_UNSUPPORTED;
    }
    return SD3D_OK;
}

}  // namespace sd3d

using namespace sd3d;

extern "C" int sd3d_lift(const float* xyz, int64_t N, const float* K4, const float* w2c, int V, int view_begin,
                         int view_end, const void* depth, int depth_dtype, int Hd, int Wd, const void* fmap,
                         int fmap_dtype, int Hf, int Wf, int C, float stride, float tau, float z_near, int accumulate,
                         int finalize, const int32_t* order, float* out_feat, int32_t* count, int32_t* pix_idx,
                         uint8_t* vis, const int32_t* seg_offsets, int64_t S, const int32_t* task_offsets,
                         const int32_t* task_seg, int64_t max_tasks, int run, void* ws, size_t ws_bytes, int pool_,
                         int variant, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N < 0 || V < 0 || view_begin < 0 || view_end > V || view_begin > view_end || Hd <= 0 || Wd <= 0 || Hf <= 0 ||
        Wf <= 0 || C <= 0 || !(stride > 0.f) || N >= (int64_t(1) << 31) - 64 || (int64_t)Hd * Wd >= (int64_t(1) << 31)) {
        set_error("sd3d_lift: bad shape N=%lld V=%d views=[%d,%d) depth=%dx%d fmap=%dx%dx%d stride=%g", (long long)N,
                  V, view_begin, view_end, Hd, Wd, Hf, Wf, C, (double)stride);
        return SD3D_ERR_ARG;
    }
    if (C % 4 != 0 || C > 1024) {
        set_error("sd3d_lift: C=%d must be a multiple of 4 and <= 1024", C);
        return SD3D_ERR_UNSUPPORTED;
    }
    if (depth_dtype != SD3D_F32 && depth_dtype != SD3D_U16) {
        set_error("sd3d_lift: depth dtype code %d unsupported", depth_dtype);
        return SD3D_ERR_UNSUPPORTED;
    }
    if (N > 0 && (xyz == nullptr || out_feat == nullptr || count == nullptr)) {
        set_error("sd3d_lift: null xyz/out_feat/count");
        return SD3D_ERR_ARG;
    }
    if (view_end > view_begin && N > 0 && (K4 == nullptr || w2c == nullptr || depth == nullptr || fmap == nullptr)) {
        set_error("sd3d_lift: null camera/depth/fmap");
        return SD3D_ERR_ARG;
    }
    if (!aligned16(K4) || !aligned16(w2c) || !aligned16(fmap) || !aligned16(out_feat)) {
        set_error("sd3d_lift: K4/w2c/fmap/out_feat must be 16-byte aligned");
        return SD3D_ERR_ARG;
    }
    const bool pool = pool_ != 0;
    if (run <= 0) run = 32;
    if (pool) {
        if (!finalize || order == nullptr || seg_offsets == nullptr || task_offsets == nullptr ||
            task_seg == nullptr || ws == nullptr || S < 0 || max_tasks < sd3d_sp_max_tasks(N, S, run) ||
            ws_bytes < (size_t)max_tasks * C * sizeof(float) || !aligned16(ws)) {
            set_error("sd3d_lift: fused pooling needs finalize=1, order, seg_offsets, task tables and ws >= max_tasks*C*4");
            return SD3D_ERR_ARG;
        }
    }
    if (N == 0) return SD3D_OK;  // task_offsets are all zero -> sd3d_sp_combine writes zero rows
    LiftParams p;
    p.xyz = xyz; p.N = N; p.K4 = K4; p.w2c = w2c; p.v_begin = view_begin; p.v_end = view_end;
    p.depth = depth; p.depth_u16 = depth_dtype == SD3D_U16; p.Hd = Hd; p.Wd = Wd;
    p.fmap = fmap; p.Hf = Hf; p.Wf = Wf; p.C = C; p.stride = stride; p.tau = tau; p.z_near = z_near;
    p.accumulate = accumulate; p.finalize = finalize; p.order = order; p.out = out_feat; p.count = count;
    p.pix_idx = pix_idx; p.vis = vis; p.pool = pool ? 1 : 0; p.seg_offsets = seg_offsets;
    p.task_offsets = task_offsets; p.task_seg = task_seg; p.S = (int32_t)S; p.run = run;
    p.partials = reinterpret_cast<float*>(ws);
    const int64_t n_tasks = pool ? max_tasks : ceil_div64(N, run);
    int rc;
    switch (fmap_dtype) {
        case SD3D_F32: rc = dispatch_lift<float>(p, n_tasks, variant, stream); break;
        case SD3D_F16: rc = dispatch_lift<__half>(p, n_tasks, variant, stream); break;
        case SD3D_BF16: rc = dispatch_lift<__nv_bfloat16>(p, n_tasks, variant, stream); break;
        default:
            set_error("sd3d_lift: fmap dtype code %d unsupported", fmap_dtype);
            return SD3D_ERR_UNSUPPORTED;
    }
    if (rc != SD3D_OK) return rc;
    return check_launch("sd3d_lift");
}

extern "C" int sd3d_sp_combine(const void* partials, const int32_t* task_offsets, const int32_t* seg_offsets,
                               int64_t S, int C, float* sp_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (S < 0 || C <= 0 || C % 4 != 0 || S >= (int64_t(1) << 30)) {
        set_error("sd3d_sp_combine: bad shape S=%lld C=%d (C must be a multiple of 4)", (long long)S, C);
        return SD3D_ERR_ARG;
    }
    if (S == 0) return SD3D_OK;
    if (partials == nullptr || task_offsets == nullptr || seg_offsets == nullptr || sp_out == nullptr ||
        !aligned16(partials) || !aligned16(sp_out)) {
        set_error("sd3d_sp_combine: null or misaligned buffer");
        return SD3D_ERR_ARG;
    }
    const int64_t threads = S * (int64_t)(C / 4);
    sp_combine_kernel<<<(unsigned)ceil_div64(threads, 256), 256, 0, stream>>>(
        reinterpret_cast<const float*>(partials), task_offsets, seg_offsets, (int32_t)S, C, sp_out);
    return check_launch("sd3d_sp_combine");
}

extern "C" int sd3d_lift_finalize(float* sum_inout, const int32_t* count, int64_t N, int C, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N < 0 || C <= 0 || C % 4 != 0 || (N > 0 && (sum_inout == nullptr || count == nullptr)) ||
        !aligned16(sum_inout)) {
        set_error("sd3d_lift_finalize: bad argument (C must be a multiple of 4, buffers 16-byte aligned)");
        return SD3D_ERR_ARG;
    }
    if (N == 0) return SD3D_OK;
    const int64_t total = N * (C / 4);
    const unsigned grid = (unsigned)imin64(ceil_div64(total, 256), (int64_t)num_sms() * 16);
    finalize_kernel<<<grid, 256, 0, stream>>>(sum_inout, count, N, C);
    return check_launch("sd3d_lift_finalize");
}

extern "C" int sd3d_scale_mean(const float* const* feats_host, int L, int64_t numel, float* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (feats_host == nullptr || L <= 0 || L > kMaxScales || numel < 0 || (numel > 0 && out == nullptr)) {
        set_error("sd3d_scale_mean: bad argument (1 <= L <= %d)", kMaxScales);
        return SD3D_ERR_ARG;
    }
    if (numel == 0) return SD3D_OK;
    ScalePtrs ptrs;
    for (int l = 0; l < kMaxScales; ++l) ptrs.p[l] = l < L ? feats_host[l] : nullptr;
    const unsigned grid = (unsigned)imin64(ceil_div64(numel, 256), (int64_t)num_sms() * 16);
    scale_mean_kernel<<<grid, 256, 0, stream>>>(ptrs, L, numel, out);
    return check_launch("sd3d_scale_mean");
}
