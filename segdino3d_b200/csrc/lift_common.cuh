// lift_common.cuh -- device helpers shared by the lifting kernels (lift.cu: projection + direct gather,
// lift_staged.cu: shared-memory staged gather). Arithmetic follows SURVEY.md Appendix A op for op.
#pragma once
#include <cstring>

#include "common.cuh"

namespace sd3d {

constexpr int kLiftThreads = 128;
constexpr int kLiftWarps = kLiftThreads / 32;

#ifndef SD3D_SCALAR_BLEND
#define SD3D_SCALAR_BLEND 0  // 1 = scalar FMUL/FADD blend (the pre-FMUL2 code, kept for A/B timing)
#endif

constexpr int kMaxPeers = 16;

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
    float stride, inv_stride, tau, z_near;  // inv_stride = 1/stride when stride is a power of two, else 0
    int accumulate, finalize;
    int by_pos;  // rows of out / count are indexed by processing position instead of point id
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
    // K1 -> K2 hand-off (workspace): per point, one record per visible view of this call, in ascending view order
    // push mode (view-sharded multi-GPU, sd3d_lift_push): the un-normalised row of processing position i goes to rank
    // i / rows_per_rank, slot [src_rank][i % rows_per_rank] of that rank's staging buffers (peer-mapped device memory)
    int n_peers, src_rank, task_rot;
    int64_t rows_per_rank;
    float* peer_out[kMaxPeers];
    int32_t* peer_cnt[kMaxPeers];
    int k_views;      // > 0: nearest-view sampling -- only the k_views visible views with the smallest camera depth count
    int4* recs;       // [N][n_views]; only the first nvis[pid] entries of a row are written
    int32_t* nvis;    // [N]
    int n_views;
};

// One 128-bit load per lane per tap row: 4 fp32 channels, or 8 fp16 / bf16 channels (= two float4 registers once
// decoded). The loaded word stays RAW while the load is in flight (`Raw`); it is decoded to fp32 when the sample is
// blended, so that the conversion instructions never wait on a load that was only just issued.
// chan_of(k, lane) = first channel held by float4 register k of this lane.
template <typename FT>
struct Tap;
template <>
struct Tap<float> {
    static constexpr int kElems = 4, kRegs = 1;
    typedef float4 Raw;
    __device__ __forceinline__ static Raw zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ static Raw load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    __device__ __forceinline__ static Raw from_smem(const uint8_t* p) { return *reinterpret_cast<const float4*>(p); }
    __device__ __forceinline__ static void decode(float4* dst, const Raw raw) { dst[0] = raw; }
};
template <>
struct Tap<__half> {
    static constexpr int kElems = 8, kRegs = 2;
    typedef uint4 Raw;
    __device__ __forceinline__ static Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
    __device__ __forceinline__ static Raw load(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ static Raw from_smem(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ static void decode(float4* dst, const Raw raw) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
        const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&raw.z));
        const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&raw.w));
        dst[0] = make_float4(a.x, a.y, b.x, b.y);
        dst[1] = make_float4(c.x, c.y, d.x, d.y);
    }
};
template <>
struct Tap<__nv_bfloat16> {
    static constexpr int kElems = 8, kRegs = 2;
    typedef uint4 Raw;
    __device__ __forceinline__ static Raw zero() { return make_uint4(0u, 0u, 0u, 0u); }
    __device__ __forceinline__ static Raw load(const __nv_bfloat16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ static Raw from_smem(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ static void decode(float4* dst, const Raw raw) {
        dst[0] = make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u),
                             __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
        dst[1] = make_float4(__uint_as_float(raw.z << 16), __uint_as_float(raw.z & 0xffff0000u),
                             __uint_as_float(raw.w << 16), __uint_as_float(raw.w & 0xffff0000u));
    }
};
template <typename FT>
__device__ __forceinline__ int chan_of(int k, int lane) {
    return ((k / Tap<FT>::kRegs) * 32 + lane) * Tap<FT>::kElems + (k % Tap<FT>::kRegs) * 4;
}

// f = ((w00*t00 + w01*t01) + w10*t10) + w11*t11 ; acc = acc + f      (Appendix A, unfused)
template <bool FAST>
__device__ __forceinline__ float blend1(float acc, float w00, float w01, float w10, float w11, float t00, float t01,
                                        float t10, float t11) {
    if (FAST) {  // contracted: 4 FFMA, differs from the spec order by O(1 ulp) per sample
        return fmaf(w11, t11, fmaf(w10, t10, fmaf(w01, t01, fmaf(w00, t00, acc))));
    }
    float f = __fadd_rn(__fmul_rn(w00, t00), __fmul_rn(w01, t01));
    f = __fadd_rn(f, __fmul_rn(w10, t10));
    f = __fadd_rn(f, __fmul_rn(w11, t11));
    return __fadd_rn(acc, f);
}

// ---- packed fp32x2 arithmetic (sm_100 FMUL2 / FADD2 / FFMA2): two channels per instruction, each half rounded exactly
// like the scalar __fmul_rn / __fadd_rn, so the blend issues half as many FP instructions and stays bit-exact.
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (single rounding; unlike the scalar .rn forms, and
// even through the __fmul2_rn/__fadd2_rn intrinsics), which would break Appendix A. The adds that consume a product
// are therefore written as fma(product, 1.0, addend) with the 1.0 read from constant memory, which ptxas cannot see
// through: product*1 is exact, so the result is rn(product + addend), the same as the unfused add.
typedef unsigned long long u64;
static __constant__ float2 c_one2 = {1.0f, 1.0f};
__device__ __forceinline__ u64 pk2(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpk2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
struct Weights2 {
    u64 w00, w01, w10, w11, one;
};
// two channels of blend1: f = ((w00*t00 + w01*t01) + w10*t10) + w11*t11 ; acc = acc + f
template <bool FAST>
__device__ __forceinline__ void blend2(float& ax, float& ay, const Weights2& w, float t00x, float t00y, float t01x,
                                       float t01y, float t10x, float t10y, float t11x, float t11y) {
    u64 r;
    if (FAST) {
        r = fma2(w.w11, pk2(t11x, t11y),
                 fma2(w.w10, pk2(t10x, t10y), fma2(w.w01, pk2(t01x, t01y), fma2(w.w00, pk2(t00x, t00y), pk2(ax, ay)))));
    } else {
        u64 f = fma2(mul2(w.w01, pk2(t01x, t01y)), w.one, mul2(w.w00, pk2(t00x, t00y)));
        f = fma2(mul2(w.w10, pk2(t10x, t10y)), w.one, f);
        f = fma2(mul2(w.w11, pk2(t11x, t11y)), w.one, f);
        r = add2(pk2(ax, ay), f);  // neither operand is a bare product: nothing to contract
    }
    unpk2(r, ax, ay);
}

// One bilinear sample in flight: the four tap rows (this lane's channel vectors, still raw) + the weights.
// NV = float4 accumulator registers per lane; a raw tap word decodes to Tap<FT>::kRegs of them.
template <int NV, typename FT>
struct Sample {
    float w00, w01, w10, w11;
    typename Tap<FT>::Raw t00[NV / Tap<FT>::kRegs], t01[NV / Tap<FT>::kRegs], t10[NV / Tap<FT>::kRegs], t11[NV / Tap<FT>::kRegs];
};

// Per-sample scalars, computed ONCE by the lane that owns the sample (Appendix A lines `uf = ...` ..
// `w11 = ...`) and broadcast with shuffles: in-view element offset of tap (y0,x0), the four weights, and
// which taps fall inside the map (bits 0..3 = t00,t01,t10,t11).
struct SampleScalars {
    // byte address of tap (y0,x0), channel 0, of the sample's view (16-byte aligned), with the tap-valid bits
    // (bit0..3 = t00,t01,t10,t11) packed into the 4 free low bits. The address may lie outside the map when
    // its tap is outside (then that bit is clear and the address is never dereferenced).
    uint32_t addr_lo, addr_hi;
    float w00, w01, w10, w11;
};

__device__ __forceinline__ void scalars_clear(SampleScalars& r) {
    r.addr_lo = r.addr_hi = 0u;
    r.w00 = r.w01 = r.w10 = r.w11 = 0.f;
}

// Appendix A lines `uf = ...` .. `w11 = ...`: tap origin, bilinear weights, which taps lie inside the map
struct TapGeom {
    int x0, y0;
    float ax, ay;
    float w00, w01, w10, w11;
    uint32_t flags;  // bit0..3 = t00,t01,t10,t11 inside the map
};
// inv_stride > 0: stride is a power of two, so x / stride == x * inv_stride bit for bit (a pure exponent shift)
__device__ __forceinline__ TapGeom tap_geometry(float u, float w, float stride, float inv_stride, int Hf, int Wf) {
    TapGeom r;
    float uf, wf;
    if (inv_stride > 0.f) {
        uf = __fsub_rn(__fmul_rn(__fadd_rn(u, 0.5f), inv_stride), 0.5f);
        wf = __fsub_rn(__fmul_rn(__fadd_rn(w, 0.5f), inv_stride), 0.5f);
    } else {
        uf = __fsub_rn(__fdiv_rn(__fadd_rn(u, 0.5f), stride), 0.5f);
        wf = __fsub_rn(__fdiv_rn(__fadd_rn(w, 0.5f), stride), 0.5f);
    }
    const float x0f = floorf(uf), y0f = floorf(wf);
    const float ax = __fsub_rn(uf, x0f), ay = __fsub_rn(wf, y0f);
    r.x0 = (int)x0f;
    r.y0 = (int)y0f;
    r.ax = ax;
    r.ay = ay;
    const float omx = __fsub_rn(1.0f, ax), omy = __fsub_rn(1.0f, ay);
    r.w00 = __fmul_rn(omx, omy);
    r.w01 = __fmul_rn(ax, omy);
    r.w10 = __fmul_rn(omx, ay);
    r.w11 = __fmul_rn(ax, ay);
    const bool okx0 = (r.x0 >= 0) && (r.x0 < Wf), okx1 = (r.x0 + 1 >= 0) && (r.x0 + 1 < Wf);
    const bool oky0 = (r.y0 >= 0) && (r.y0 < Hf), oky1 = (r.y0 + 1 >= 0) && (r.y0 + 1 < Hf);
    r.flags = (uint32_t)(oky0 && okx0) | ((uint32_t)(oky0 && okx1) << 1) | ((uint32_t)(oky1 && okx0) << 2) |
              ((uint32_t)(oky1 && okx1) << 3);
    return r;
}

template <typename FT>
__device__ __forceinline__ SampleScalars make_scalars(const FT* fmap, int64_t view_elems, int view, float u, float w,
                                                      float stride, float inv_stride, int Hf, int Wf, int C) {
    SampleScalars r;
    const TapGeom g = tap_geometry(u, w, stride, inv_stride, Hf, Wf);
    r.w00 = g.w00;
    r.w01 = g.w01;
    r.w10 = g.w10;
    r.w11 = g.w11;
    const int64_t elem = (int64_t)view * view_elems + ((int64_t)g.y0 * Wf + g.x0) * C;
    const uint64_t addr = (uint64_t)(reinterpret_cast<uintptr_t>(fmap) + elem * (int64_t)sizeof(FT));
    r.addr_lo = (uint32_t)addr | g.flags;
    r.addr_hi = (uint32_t)(addr >> 32);
    return r;
}

// issue the 4*NV 128-bit loads of sample `src_lane` (no use of the data here -> they stay in flight).
// cmask: bit l set = this lane's l-th 128-bit vector lies inside the C channels (hoisted out of the sample loop).
// Sample record written by K1 for every visible (point, view): {pixel index of tap (y0, x0) in the [V, Hf, Wf] map
// (may lie outside the map by one row/column), tap-valid flags, ax, ay}. K2 rebuilds the packed tap address and the
// four weights from it with the same operations as make_scalars (w = products of ax, ay, 1-ax, 1-ay).
// Word 1 = tap-valid flags (bits 0..3) | (x0 + 1) << 4 (14 bits) | (y0 + 1) << 18 (14 bits): the tap origin again, in
// the form the stage planner of lift_staged.cu wants (x0, y0 >= -1; Hf, Wf <= kMaxMapDim).
constexpr int kMaxMapDim = 16382;
__device__ __forceinline__ int4 make_record(const TapGeom& g, float ax, float ay, int view, int Hf, int Wf) {
    const uint32_t w1 = g.flags | ((uint32_t)(g.x0 + 1) << 4) | ((uint32_t)(g.y0 + 1) << 18);
    return make_int4((view * Hf + g.y0) * Wf + g.x0, (int)w1, __float_as_int(ax), __float_as_int(ay));
}
__device__ __forceinline__ int rec_x0(int w1) { return (int)(((uint32_t)w1 >> 4) & 0x3fffu) - 1; }
__device__ __forceinline__ int rec_y0(int w1) { return (int)(((uint32_t)w1 >> 18) & 0x3fffu) - 1; }
template <typename FT>
__device__ __forceinline__ SampleScalars scalars_from_record(const int4 rec, const FT* fmap, int C) {
    SampleScalars r;
    const float ax = __int_as_float(rec.z), ay = __int_as_float(rec.w);
    const float omx = __fsub_rn(1.0f, ax), omy = __fsub_rn(1.0f, ay);
    r.w00 = __fmul_rn(omx, omy);
    r.w01 = __fmul_rn(ax, omy);
    r.w10 = __fmul_rn(omx, ay);
    r.w11 = __fmul_rn(ax, ay);
    const uint64_t addr = (uint64_t)(reinterpret_cast<uintptr_t>(fmap) + (int64_t)rec.x * C * (int64_t)sizeof(FT));
    r.addr_lo = (uint32_t)addr | ((uint32_t)rec.y & 0xFu);
    r.addr_hi = (uint32_t)(addr >> 32);
    return r;
}

template <int NV, typename FT>
__device__ __forceinline__ void sample_issue_loads(Sample<NV, FT>& s, uint32_t lo, uint32_t hi, int C, int row_elems, int lane,
                                                   unsigned cmask);
template <int NV, typename FT>
__device__ __forceinline__ void sample_issue(Sample<NV, FT>& s, const SampleScalars& mine, int src_lane, int C,
                                             int row_elems, int lane, unsigned cmask) {
    const uint32_t lo = __shfl_sync(kFull, mine.addr_lo, src_lane);
    const uint32_t hi = __shfl_sync(kFull, mine.addr_hi, src_lane);
    s.w00 = __shfl_sync(kFull, mine.w00, src_lane);
    s.w01 = __shfl_sync(kFull, mine.w01, src_lane);
    s.w10 = __shfl_sync(kFull, mine.w10, src_lane);
    s.w11 = __shfl_sync(kFull, mine.w11, src_lane);
    sample_issue_loads<NV, FT>(s, lo, hi, C, row_elems, lane, cmask);
}
// the loads of one sample given its (warp-uniform) packed tap address + flags
template <int NV, typename FT>
__device__ __forceinline__ void sample_issue_loads(Sample<NV, FT>& s, uint32_t lo, uint32_t hi, int C, int row_elems, int lane,
                                                   unsigned cmask) {
    constexpr int kE = Tap<FT>::kElems, kR = Tap<FT>::kRegs;
    static_assert(NV % kR == 0, "register vectors per tap must be a multiple of the registers one load fills");
    const uint32_t flags = lo & 0xFu;
    const FT* __restrict__ p00 =
        reinterpret_cast<const FT*>((uintptr_t)(((uint64_t)hi << 32) | (uint64_t)(lo & ~0xFu))) + lane * kE;
    const FT* __restrict__ p10 = p00 + row_elems;
    if (flags == 0xFu) {  // interior sample (the common case): unpredicated-on-validity loads
#pragma unroll
        for (int l = 0; l < NV / kR; ++l) {
            if (cmask & (1u << l)) {
                s.t00[l] = Tap<FT>::load(p00 + l * 32 * kE);
                s.t01[l] = Tap<FT>::load(p00 + C + l * 32 * kE);
                s.t10[l] = Tap<FT>::load(p10 + l * 32 * kE);
                s.t11[l] = Tap<FT>::load(p10 + C + l * 32 * kE);
            }
        }
    } else {  // border sample: taps outside the map read as zero (Appendix A `tap(y,x)`)
#pragma unroll
        for (int l = 0; l < NV / kR; ++l) {
            const bool cok = (cmask >> l) & 1u;
            s.t00[l] = s.t01[l] = s.t10[l] = s.t11[l] = Tap<FT>::zero();
            if (cok && (flags & 1)) s.t00[l] = Tap<FT>::load(p00 + l * 32 * kE);
            if (cok && (flags & 2)) s.t01[l] = Tap<FT>::load(p00 + C + l * 32 * kE);
            if (cok && (flags & 4)) s.t10[l] = Tap<FT>::load(p10 + l * 32 * kE);
            if (cok && (flags & 8)) s.t11[l] = Tap<FT>::load(p10 + C + l * 32 * kE);
        }
    }
}

template <typename FT, int NV>
__device__ __forceinline__ unsigned channel_mask(int C, int lane) {
    unsigned m = 0u;
#pragma unroll
    for (int l = 0; l < NV / Tap<FT>::kRegs; ++l)
        if ((l * 32 + lane) * Tap<FT>::kElems < C) m |= 1u << l;
    return m;
}

template <int NV, typename FT>
__device__ __forceinline__ void sample_clear(Sample<NV, FT>& s) {
#pragma unroll
    for (int l = 0; l < NV / Tap<FT>::kRegs; ++l) s.t00[l] = s.t01[l] = s.t10[l] = s.t11[l] = Tap<FT>::zero();
}

template <int NV, bool FAST, typename FT>
__device__ __forceinline__ void sample_accum(float4 (&acc)[NV], const Sample<NV, FT>& s) {
    constexpr int kR = Tap<FT>::kRegs;
#if !SD3D_SCALAR_BLEND
    Weights2 w;
    w.w00 = pk2(s.w00, s.w00);
    w.w01 = pk2(s.w01, s.w01);
    w.w10 = pk2(s.w10, s.w10);
    w.w11 = pk2(s.w11, s.w11);
    w.one = pk2(c_one2.x, c_one2.y);
#endif
#pragma unroll
    for (int l = 0; l < NV / kR; ++l) {
        float4 t00[kR], t01[kR], t10[kR], t11[kR];  // decoded here, when the loads have long been issued
        Tap<FT>::decode(t00, s.t00[l]);
        Tap<FT>::decode(t01, s.t01[l]);
        Tap<FT>::decode(t10, s.t10[l]);
        Tap<FT>::decode(t11, s.t11[l]);
#pragma unroll
        for (int r = 0; r < kR; ++r) {
            float4& a = acc[l * kR + r];
#if SD3D_SCALAR_BLEND
            a.x = blend1<FAST>(a.x, s.w00, s.w01, s.w10, s.w11, t00[r].x, t01[r].x, t10[r].x, t11[r].x);
            a.y = blend1<FAST>(a.y, s.w00, s.w01, s.w10, s.w11, t00[r].y, t01[r].y, t10[r].y, t11[r].y);
            a.z = blend1<FAST>(a.z, s.w00, s.w01, s.w10, s.w11, t00[r].z, t01[r].z, t10[r].z, t11[r].z);
            a.w = blend1<FAST>(a.w, s.w00, s.w01, s.w10, s.w11, t00[r].w, t01[r].w, t10[r].w, t11[r].w);
#else
            blend2<FAST>(a.x, a.y, w, t00[r].x, t00[r].y, t01[r].x, t01[r].y, t10[r].x, t10[r].y, t11[r].x, t11[r].y);
            blend2<FAST>(a.z, a.w, w, t00[r].z, t00[r].w, t01[r].z, t01[r].w, t10[r].z, t10[r].w, t11[r].z, t11[r].w);
#endif
        }
    }
}

// Appendix A lines `xc = ...` .. `w = ...` for one (point, view): returns zc, writes u / w
__device__ __forceinline__ float project_point(const float4 k4, const float4 r0, const float4 r1, const float4 r2,
                                               float px, float py, float pz, float z_near, float& u, float& w) {
    const float xc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r0.x, px), __fmul_rn(r0.y, py)), __fmul_rn(r0.z, pz)), r0.w);
    const float yc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r1.x, px), __fmul_rn(r1.y, py)), __fmul_rn(r1.z, pz)), r1.w);
    const float zc = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r2.x, px), __fmul_rn(r2.y, py)), __fmul_rn(r2.z, pz)), r2.w);
    u = 0.f;
    w = 0.f;
    if (zc > z_near) {
        u = __fadd_rn(__fdiv_rn(__fmul_rn(k4.x, xc), zc), k4.z);
        w = __fadd_rn(__fdiv_rn(__fmul_rn(k4.y, yc), zc), k4.w);
    }
    return zc;
}

}  // namespace sd3d
