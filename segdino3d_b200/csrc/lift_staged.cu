// lift_staged.cu -- bilinear gather + view sum with the tap rows STAGED IN SHARED MEMORY by the bulk-copy (TMA) engine.
//
// Steps a-2/a-3 (+ run partials of a-4) of SURVEY.md section 8(a), spec = SURVEY.md Appendix A; the output slot is
// extra_features["points_2dfeats"] (segdino3d/datasets/dataset/scannet200.py:219-234). Same arithmetic, same per-point
// view order and the same unfused blend as gather_kernel in lift.cu -> bit-identical results.
//
// Why: the direct gather pulls 4 tap rows per visible (point, view) through the L1 load path (2.9 GB per cfg2
// launch against 0.30 GB of compulsory bytes) and is bound by that path (~64 B/clk/SM for LDG.128). The 32 points
// of a run are spatial neighbours, so through ONE view their 2x2 footprints overlap: ~24 distinct feature-map pixels
// for ~80 taps. Here every distinct pixel of a (run, view) is copied ONCE from L2/HBM into shared memory by
// cp.async.bulk (no LSU instructions, no registers, completion on an mbarrier) and the blend reads its taps with
// LDS.128 at the 128 B/clk/SM shared-memory rate.
//
//   stage_plan_kernel   one warp per run (lane = point): walks the run's views in ascending order and cuts the visible
//                       samples of (run, view) into STAGES = boxes of <= 16x16 tap pixels with <= cap distinct pixels.
//                       Per stage it emits a 64-byte header {view, lane mask, box origin, 256-bit pixel bitmap}; the
//                       rank of a pixel inside the bitmap is its shared-memory slot, and the four slot ranks of every
//                       sample are written back into its 16-byte record (K1's hand-off, lift.cu). At most kStMaxSub
//                       box stages per (run, view); what is left over (samples far from the others) is marked DIRECT
//                       and fetched from global memory by the consumer, exactly like gather_kernel does.
//   gather_staged_kernel persistent CTAs (2 per SM), warp-specialised:
//       producer warp    takes runs from an atomic counter, and per stage: allocates slots in the CTA's ring of
//                        row buffers (FIFO, freed by the consumers' `empty` mbarriers), turns the bitmap rows into
//                        runs of consecutive pixels = one cp.async.bulk each (global row segment -> consecutive
//                        slots), copies the lanes' sample records next to the stage header (slot ranks translated to
//                        ring slots) and arms the stage's `full` mbarrier with the byte count.
//       consumer warps   each owns 32/WARPS points of the run, accumulators in registers across all views of the run;
//                        per stage: wait `full`, for each owned point in the stage mask: one broadcast LDS.128 of the
//                        record, 4 x C/4 LDS.128 tap vectors (lane = channel vector), unfused FMUL2/FFMA2 blend;
//                        arrive on `empty`. After the run's last stage: mean, streaming row stores, run partial.
//   Taps outside the map read a zero row that lives behind the ring (Appendix A `tap(y,x) = 0`).
#include "lift_common.cuh"
#include "lift_staged.cuh"

namespace sd3d {

constexpr int kStQ = 8;         // stage queue entries per CTA
constexpr uint32_t kStFirst = 1u, kStLast = 2u, kStExit = 4u;
constexpr uint32_t kRecDirect = 0x80000000u;  // record word 1, bit 31: taps come from global memory
constexpr uint32_t kRankZero = 0x7Fu;         // slot rank of a tap outside the map

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
// same, for waits that are expected to be long (the producer waiting for ring space): the hardware suspends the warp
// for up to the hinted time instead of spinning on the issue slot
__device__ __forceinline__ void mbar_wait_sleepy(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity), "r"(2000u)
        : "memory");
}
// global -> shared bulk copy through the TMA engine; completion is signalled as `bytes` of transaction on `bar`
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// run of task `task`: segment, first processing position, number of points (same mapping as gather_kernel)
__device__ __forceinline__ void task_range(const LiftParams& p, int64_t task, int& seg, int64_t& start, int& npts) {
    if (p.pool) {
        seg = p.task_seg[task];
        start = (int64_t)p.seg_offsets[seg] + (task - p.task_offsets[seg]) * (int64_t)p.run;
        npts = (int)(min(start + (int64_t)p.run, (int64_t)p.seg_offsets[seg + 1]) - start);
    } else {
        seg = -1;
        start = task * (int64_t)p.run;
        npts = (int)(min(start + (int64_t)p.run, p.N) - start);
    }
}

// ---------------------------------------------------------------------------------------------------
// stage planner: one warp per (run, chunk of 8 views), lane = point of the run
// ---------------------------------------------------------------------------------------------------
constexpr int kPlanWarps = 4;
constexpr int kPlanViews = 8;                       // views per planner warp
constexpr int kPlanSlots = kPlanViews * kStMaxSub;  // header slots of a chunk

__global__ void __launch_bounds__(kPlanWarps * 32) stage_plan_kernel(const LiftParams p, const StagedParams sp) {
    const int lane = lane_id();
    const int nck = (p.n_views + kPlanViews - 1) / kPlanViews;  // view chunks per run
    const int64_t gw = (int64_t)blockIdx.x * kPlanWarps + (threadIdx.x >> 5);
    const int64_t task = gw / nck;
    const int ck = (int)(gw - task * nck);
    if (blockIdx.x == 0 && threadIdx.x == 0) *sp.counter = 0;  // the gather's run dispenser
    const int64_t n_tasks = p.pool ? (int64_t)p.task_offsets[p.S + 1] : sp.n_tasks;
    if (task >= n_tasks) return;
    int seg, npts;
    int64_t start;
    task_range(p, task, seg, start, npts);
    const bool has = lane < npts;
    const int32_t pid = has ? (p.order ? p.order[start + lane] : (int32_t)(start + lane)) : -1;
    // this lane's visible views inside the chunk (bits) and before it (= index of its first record of the chunk)
    const int v_lo = ck * kPlanViews;  // relative to p.v_begin; a chunk never straddles a 32-bit mask word
    uint32_t bits = 0u;
    int below = 0;
    if (has) {
        const uint32_t* __restrict__ mrow = sp.masks + (int64_t)pid * sp.nchunks;
        const int wi = v_lo >> 5;
        for (int j = 0; j < wi; ++j) below += __popc(__ldg(mrow + j));
        const uint32_t mw = __ldg(mrow + wi);
        below += __popc(mw & ((1u << (v_lo & 31)) - 1u));
        bits = (mw >> (v_lo & 31)) & ((1u << kPlanViews) - 1u);
    }
    int4* __restrict__ recs = p.recs + (int64_t)max(pid, 0) * p.n_views + below;
    int4 recv[kPlanViews];  // the lane's record of chunk view i (if visible): all loads in flight together
#pragma unroll
    for (int i = 0; i < kPlanViews; ++i) {
        recv[i] = make_int4(0, 0, 0, 0);
        if (bits & (1u << i)) recv[i] = recs[__popc(bits & ((1u << i) - 1u))];
    }
    uint32_t* __restrict__ H = sp.hdrs + (task * (int64_t)sp.cap_stages + (int64_t)ck * kPlanSlots) * 16;
    int nst = 0;
    const uint32_t any = __reduce_or_sync(kFull, bits);
#pragma unroll 1
    for (int i = 0; i < kPlanViews; ++i) {
        if (!((any >> i) & 1u)) continue;
        int4 rec = make_int4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < kPlanViews; ++j)
            if (j == i) rec = recv[j];
        const int v = p.v_begin + v_lo + i;
        const int x0 = rec_x0(rec.y), y0 = rec_y0(rec.y);
        const uint32_t flags = (uint32_t)rec.y & 15u;
        bool cand = (bits >> i) & 1u;
        const int ridx = __popc(bits & ((1u << i) - 1u));
        for (int sub = 0; sub < kStMaxSub; ++sub) {
            if (!__any_sync(kFull, cand)) break;
            uint32_t bm[8];
            int xmin = 0, ymin = 0, npix = 0;
            bool sel = false;
            for (int box = 16; box >= 2; box >>= 1) {
                xmin = __reduce_min_sync(kFull, cand ? x0 : 0x3fffffff);
                const bool fitx = cand && (x0 - xmin <= box - 2);
                ymin = __reduce_min_sync(kFull, fitx ? y0 : 0x3fffffff);
                sel = fitx && (y0 - ymin <= box - 2);
                const int cx = x0 - xmin, cy = y0 - ymin;  // 0..14 when sel
                const uint32_t top = sel ? (flags & 3u) << cx : 0u, bot = sel ? ((flags >> 2) & 3u) << cx : 0u;
                npix = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                    uint32_t c = 0u;
                    if ((cy >> 1) == w) c |= top << ((cy & 1) * 16);
                    if (((cy + 1) >> 1) == w) c |= bot << (((cy + 1) & 1) * 16);
                    bm[w] = __reduce_or_sync(kFull, c);
                    npix += __popc(bm[w]);
                }
                if (npix <= sp.cap_pix) break;
            }
            // slot rank of a tap = number of bitmap bits below it
            uint32_t ranks = 0u;
            if (sel) {
                const int cx = x0 - xmin, cy = y0 - ymin;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    uint32_t r = kRankZero;
                    if (flags & (1u << t)) {
                        const int b = (cy + (t >> 1)) * 16 + cx + (t & 1);
                        int rk = 0;
#pragma unroll
                        for (int w = 0; w < 8; ++w) {
                            if (w < (b >> 5)) rk += __popc(bm[w]);
                            if (w == (b >> 5)) rk += __popc(bm[w] & ((1u << (b & 31)) - 1u));
                        }
                        r = (uint32_t)rk;
                    }
                    ranks |= r << (8 * t);
                }
            }
            const bool direct = (sub == kStMaxSub - 1) && cand && !sel;  // the view's leftovers ride on its last stage
            const bool member = sel || direct;
            const uint32_t mask = __ballot_sync(kFull, member);
            if (lane < 16) {
                uint32_t wv = 0u;
                if (lane == 0) wv = (uint32_t)v;
                if (lane == 1) wv = mask;
                if (lane == 2) wv = ((uint32_t)xmin & 0xffffu) | ((uint32_t)ymin << 16);
                if (lane == 3) wv = (uint32_t)npix;
#pragma unroll
                for (int w = 0; w < 8; ++w)
                    if (lane == 4 + w) wv = bm[w];
                H[nst * 16 + lane] = wv;
            }
            if (member) {
                recs[ridx].y = sel ? (int)ranks : (int)(flags | kRecDirect);
                cand = false;
            }
            ++nst;
        }
    }
    // the last warp of the run to finish packs the chunks' headers to the front of the run's header array
    __syncwarp();
    int prev = 0;
    if (lane == 0) {
        sp.chunk_cnt[task * nck + ck] = nst;
        __threadfence();
        prev = atomicAdd(sp.done + task, 1);
    }
    prev = __shfl_sync(kFull, prev, 0);
    if (prev != nck - 1) return;
    __threadfence();
    uint4* __restrict__ HV = reinterpret_cast<uint4*>(sp.hdrs + task * (int64_t)sp.cap_stages * 16);
    int total = 0;
    for (int c0 = 0; c0 < nck; c0 += 32) {
        const int c = c0 + lane;
        const int cnt = c < nck ? __ldcg(sp.chunk_cnt + task * nck + c) : 0;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += up;
        }
        const int dst0 = total + incl - cnt;
        total += __shfl_sync(kFull, incl, 31);
        for (int l = 0; l < 32; ++l) {  // chunk c0 + l: its headers move from slot (c0+l)*kPlanSlots to dst
            const int n_l = __shfl_sync(kFull, cnt, l), d_l = __shfl_sync(kFull, dst0, l);
            const int s_l = (c0 + l) * kPlanSlots;
            if (n_l == 0 || d_l == s_l) continue;
            for (int j0 = 0; j0 < n_l; j0 += 8) {  // 8 headers (4 x 16 bytes each) per pass; dst < src: load all, then store
                const int j = j0 + (lane >> 2);
                uint4 val = make_uint4(0u, 0u, 0u, 0u);
                if (j < n_l) val = __ldcg(HV + (int64_t)(s_l + j) * 4 + (lane & 3));
                __syncwarp();
                if (j < n_l) HV[(int64_t)(d_l + j) * 4 + (lane & 3)] = val;
            }
        }
    }
    if (total == 0) {  // nothing of this run is visible: one empty stage carries the run through the pipeline
        if (lane < 4) HV[lane] = make_uint4(0u, 0u, 0u, 0u);
        total = 1;
    }
    sp.runpts[task * 32 + lane] = make_int2(pid, has ? p.nvis[pid] : 0);
    if (lane == 0) sp.runinfo[task] = make_int4((int)start, npts, seg, total);
}

// ---------------------------------------------------------------------------------------------------
// staged gather
// ---------------------------------------------------------------------------------------------------
struct __align__(16) StageSlot {
    uint4 hdr;        // mask, first ring slot, flags, task
    uint4 run;        // first processing position, points, segment, -
    uint4 rec[32];    // the lanes' sample records, translated by the producer:
                      //   staged: x = off00 | off01 << 16, y = off10 | off11 << 16 (tap row offsets in the ring, / 16)
                      //   direct: x = pixel index of tap (y0, x0), y = tap-valid flags;  z = ax (sign bit = direct), w = ay
    int32_t pid[32];  // kStFirst only
    int32_t cnt[32];  // kStFirst only: visible views of the point in this call
};

// the rare sample whose taps are not staged: same loads as gather_kernel, kept out of line so that the unrolled
// consumer loop stays small (instruction cache)
template <int NV, typename FT>
__device__ __noinline__ Sample<NV> direct_taps(const FT* __restrict__ fmap, uint32_t pix, uint32_t flags, int C,
                                               int row_elems, int lane, unsigned cmask) {
    constexpr int kE = Tap<FT>::kElems, kR = Tap<FT>::kRegs;
    Sample<NV> s;
    sample_clear<NV>(s);
    const FT* __restrict__ p00 = fmap + (int64_t)(int)pix * C + lane * kE;
    const FT* __restrict__ p10 = p00 + row_elems;
#pragma unroll
    for (int l = 0; l < NV / kR; ++l) {
        if (!((cmask >> l) & 1u)) continue;
        if (flags & 1u) Tap<FT>::load(&s.t00[l * kR], p00 + l * 32 * kE);
        if (flags & 2u) Tap<FT>::load(&s.t01[l * kR], p00 + C + l * 32 * kE);
        if (flags & 4u) Tap<FT>::load(&s.t10[l * kR], p10 + l * 32 * kE);
        if (flags & 8u) Tap<FT>::load(&s.t11[l * kR], p10 + C + l * 32 * kE);
    }
    return s;
}

template <int NV, typename FT>
__device__ __forceinline__ void staged_issue(Sample<NV>& s, const uint4 r, const uint8_t* __restrict__ lane_base,
                                             const FT* __restrict__ fmap, int C, int row_elems, int lane,
                                             unsigned cmask) {
    constexpr int kR = Tap<FT>::kRegs;
    const float ax = __uint_as_float(r.z & 0x7fffffffu), ay = __uint_as_float(r.w);  // ax >= 0: its sign bit = direct
    const float omx = __fsub_rn(1.0f, ax), omy = __fsub_rn(1.0f, ay);
    s.w00 = __fmul_rn(omx, omy);
    s.w01 = __fmul_rn(ax, omy);
    s.w10 = __fmul_rn(omx, ay);
    s.w11 = __fmul_rn(ax, ay);
    if (r.z & kRecDirect) {  // warp-uniform, rare
        const Sample<NV> d = direct_taps<NV, FT>(fmap, r.x, r.y & 0xFu, C, row_elems, lane, cmask);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            s.t00[k] = d.t00[k];
            s.t01[k] = d.t01[k];
            s.t10[k] = d.t10[k];
            s.t11[k] = d.t11[k];
        }
    } else {
        const uint8_t* __restrict__ p00 = lane_base + ((r.x & 0xffffu) << 4);
        const uint8_t* __restrict__ p01 = lane_base + ((r.x >> 16) << 4);
        const uint8_t* __restrict__ p10 = lane_base + ((r.y & 0xffffu) << 4);
        const uint8_t* __restrict__ p11 = lane_base + ((r.y >> 16) << 4);
#pragma unroll
        for (int l = 0; l < NV / kR; ++l) {
            if (cmask & (1u << l)) {
                if constexpr (kR == 1) {
                    s.t00[l] = *reinterpret_cast<const float4*>(p00 + l * 512);
                    s.t01[l] = *reinterpret_cast<const float4*>(p01 + l * 512);
                    s.t10[l] = *reinterpret_cast<const float4*>(p10 + l * 512);
                    s.t11[l] = *reinterpret_cast<const float4*>(p11 + l * 512);
                } else {
                    Tap<FT>::decode(&s.t00[l * kR], *reinterpret_cast<const uint4*>(p00 + l * 512));
                    Tap<FT>::decode(&s.t01[l * kR], *reinterpret_cast<const uint4*>(p01 + l * 512));
                    Tap<FT>::decode(&s.t10[l * kR], *reinterpret_cast<const uint4*>(p10 + l * 512));
                    Tap<FT>::decode(&s.t11[l * kR], *reinterpret_cast<const uint4*>(p11 + l * 512));
                }
            }
        }
    }
}

// FULL: C fills all NV register vectors of every lane (C = 128 * NV fp32 / 256 * NV / 2 16-bit): no channel predicates
template <int NV, typename FT, bool FAST, int PTS, bool DB, bool FULL>
__global__ void __launch_bounds__((32 / PTS + 1) * 32, 2)
    gather_staged_kernel(const __grid_constant__ LiftParams p, const __grid_constant__ StagedParams sp) {
    constexpr int WARPS = 32 / PTS;  // consumer warps; warp WARPS is the producer
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int rowb = p.C * (int)sizeof(FT);
    const int ring = sp.ring_slots;
    // shared memory: ring rows | zero row | stage slots | barriers | producer scratch | reduce scratch
    uint8_t* const ring_ptr = smem;
    StageSlot* const slots = reinterpret_cast<StageSlot*>(smem + (size_t)(ring + 1) * rowb);
    uint64_t* const bars = reinterpret_cast<uint64_t*>(slots + kStQ);  // full[kStQ], empty[kStQ]
    uint32_t* const hch = reinterpret_cast<uint32_t*>(bars + 2 * kStQ);  // 8 headers x 16 words
    int32_t* const ext = reinterpret_cast<int32_t*>(hch + 128);          // ring slots held by stage entry q
    float4* const sred = reinterpret_cast<float4*>(ext + 16);            // [2][WARPS][NV*32]
    const uint32_t full0 = smem_addr(bars), empty0 = smem_addr(bars + kStQ);

    for (int i = threadIdx.x; i < rowb / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(ring_ptr + (size_t)ring * rowb)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) {
        for (int q = 0; q < kStQ; ++q) {
            mbar_init(full0 + 8 * q, 1);
            mbar_init(empty0 + 8 * q, WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const FT* __restrict__ fmap = reinterpret_cast<const FT*>(p.fmap);
    const int64_t n_tasks = p.pool ? (int64_t)p.task_offsets[p.S + 1] : sp.n_tasks;

    if (warp == WARPS) {
        // ------------------------------------------------ producer ------------------------------------------------
        const uint32_t ring_s = smem_addr(ring_ptr);
        const uint32_t zero_off = (uint32_t)(ring * rowb) >> 4;
        int n = 0, tail_n = 0, head = 0, free_slots = ring;
        // entry / ring allocation: FIFO over the consumers' `empty` barriers (warp-uniform)
        auto acquire = [&](int npix) -> int {
            const bool wrap = head + npix > ring;  // a stage's slots are contiguous: skip the ring's tail
            const int skip = wrap ? ring - head : 0;
            while (tail_n + kStQ <= n || free_slots < npix + skip) {
                mbar_wait_sleepy(empty0 + 8 * (tail_n % kStQ), (uint32_t)((tail_n / kStQ) & 1));
                free_slots += ext[tail_n % kStQ];
                ++tail_n;
            }
            if (wrap) head = 0;
            const int base = head;
            head += npix;
            free_slots -= npix + skip;
            __syncwarp();
            if (lane == 0) ext[n % kStQ] = npix + skip;
            __syncwarp();
            return base;
        };
        // run pipeline: c_next = ticket of the run after the current one, its descriptors are loaded one run ahead
        int c_cur = 0, c_next = 0;
        if (lane == 0) {
            c_cur = atomicAdd(sp.counter, 1);
            c_next = atomicAdd(sp.counter, 1);
        }
        c_cur = __shfl_sync(kFull, c_cur, 0);
        c_next = __shfl_sync(kFull, c_next, 0);
        auto task_of = [&](int c) -> int64_t { return ((int64_t)c + sp.task_rot) % n_tasks; };
        const int hlanes = min(8, sp.cap_stages) * 4;
        int4 ri = make_int4(0, 0, 0, 0);
        int2 rp = make_int2(-1, 0);
        uint4 hc = make_uint4(0u, 0u, 0u, 0u);
        if (c_cur < n_tasks) {
            const int64_t t = task_of(c_cur);
            ri = sp.runinfo[t];
            rp = sp.runpts[t * 32 + lane];
            if (lane < hlanes) hc = reinterpret_cast<const uint4*>(sp.hdrs + t * (int64_t)sp.cap_stages * 16)[lane];
        }
        while (c_cur < n_tasks) {
            const int64_t task = task_of(c_cur);
            const int nst = ri.w, pid = rp.x, nv = rp.y;
            const int4* __restrict__ recs = p.recs + (int64_t)max(pid, 0) * p.n_views;
            const uint4* __restrict__ hsrc = reinterpret_cast<const uint4*>(sp.hdrs + task * (int64_t)sp.cap_stages * 16);
            int k = 0;
            int4 rec = make_int4(0, 0, 0, 0);
            if (k < nv) rec = recs[0];
            // prefetch the next run's descriptors and the ticket after it
            const int c_after_l = (lane == 0) ? atomicAdd(sp.counter, 1) : 0;
            int4 ri_n = make_int4(0, 0, 0, 0);
            int2 rp_n = make_int2(-1, 0);
            uint4 hc_n = make_uint4(0u, 0u, 0u, 0u);
            if (c_next < n_tasks) {
                const int64_t t = task_of(c_next);
                ri_n = sp.runinfo[t];
                rp_n = sp.runpts[t * 32 + lane];
                if (lane < hlanes) hc_n = reinterpret_cast<const uint4*>(sp.hdrs + t * (int64_t)sp.cap_stages * 16)[lane];
            }
            for (int s0 = 0; s0 < nst; s0 += 8) {
                const int cnt8 = min(8, nst - s0);
                __syncwarp();
                reinterpret_cast<uint4*>(hch)[lane] = hc;
                __syncwarp();
                if (s0 + 8 < nst)  // next chunk of headers: in flight while this one is turned into copies
                    hc = (lane < min(8, nst - s0 - 8) * 4) ? hsrc[(s0 + 8) * 4 + lane] : make_uint4(0u, 0u, 0u, 0u);
                for (int s = 0; s < cnt8; ++s) {
                    const uint32_t* Hd = hch + s * 16;
                    const uint32_t view = Hd[0], mask = Hd[1], xy = Hd[2];
                    const int npix = (int)Hd[3];
                    const int xmin = (int)(int16_t)(xy & 0xffffu), ymin = (int)(int16_t)(xy >> 16);
                    const int q = n % kStQ;
                    const int base = acquire(npix);
                    StageSlot& S = slots[q];
                    const uint32_t flags = ((s0 + s == 0) ? kStFirst : 0u) | ((s0 + s == nst - 1) ? kStLast : 0u);
                    if (flags & kStFirst) {
                        S.pid[lane] = pid;
                        S.cnt[lane] = nv;
                        if (lane == 0) S.run = make_uint4((uint32_t)ri.x, (uint32_t)ri.y, (uint32_t)ri.z, 0u);
                    }
                    if ((mask >> lane) & 1u) {
                        uint4 r = make_uint4((uint32_t)rec.x, (uint32_t)rec.y, (uint32_t)rec.z, (uint32_t)rec.w);
                        if (r.y & kRecDirect) {
                            r.z |= kRecDirect;
                        } else {  // slot ranks -> byte offsets / 16 inside the ring (the zero row sits behind it)
                            const uint32_t rk = r.y;
                            uint32_t o[4];
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const uint32_t b = (rk >> (8 * t)) & 0xffu;
                                o[t] = b == kRankZero ? zero_off : ((uint32_t)(base + (int)b) * (uint32_t)rowb) >> 4;
                            }
                            r.x = o[0] | (o[1] << 16);
                            r.y = o[2] | (o[3] << 16);
                        }
                        S.rec[lane] = r;
                        ++k;
                        if (k < nv) rec = recs[k];
                    }
                    __syncwarp();
                    const uint32_t fullb = full0 + 8 * q;
                    if (lane == 0) {
                        S.hdr = make_uint4(mask, (uint32_t)base, flags, (uint32_t)task);
                        mbar_arrive_expect_tx(fullb, (uint32_t)npix * (uint32_t)rowb);
                    }
                    __syncwarp();
                    if (npix > 0) {  // bitmap rows -> runs of consecutive pixels -> one bulk copy each
                        const int r = lane & 15;
                        uint32_t bits = lane < 16 ? (Hd[4 + (r >> 1)] >> ((r & 1) * 16)) & 0xffffu : 0u;
                        int rowbase = __popc(bits);
#pragma unroll
                        for (int o = 1; o < 16; o <<= 1) {
                            const int up = __shfl_up_sync(kFull, rowbase, o);
                            if (lane >= o) rowbase += up;
                        }
                        rowbase -= __popc(bits);  // exclusive
                        const uint8_t* src_row = reinterpret_cast<const uint8_t*>(fmap) +
                                                 (((int64_t)view * p.Hf + (ymin + r)) * p.Wf + xmin) * (int64_t)rowb;
                        const uint32_t dst_row = ring_s + (uint32_t)(base + rowbase) * (uint32_t)rowb;
                        const uint32_t orig = bits;
                        while (bits) {
                            const int st = __ffs(bits) - 1;
                            const int len = __ffs(~(bits >> st)) - 1;
                            bulk_copy_g2s(dst_row + (uint32_t)__popc(orig & ((1u << st) - 1u)) * (uint32_t)rowb,
                                          src_row + (int64_t)st * rowb, (uint32_t)len * (uint32_t)rowb, fullb);
                            bits &= ~(((1u << len) - 1u) << st);
                        }
                    }
                    ++n;
                }
            }
            c_cur = c_next;
            c_next = __shfl_sync(kFull, c_after_l, 0);
            ri = ri_n;
            rp = rp_n;
            hc = hc_n;
        }
        // exit marker
        const int q = n % kStQ;
        acquire(0);
        if (lane == 0) {
            slots[q].hdr = make_uint4(0u, 0u, kStExit, 0u);
            mbar_arrive(full0 + 8 * q);
        }
        return;
    }

    // ------------------------------------------------ consumers ------------------------------------------------
    const int row_elems = p.Wf * p.C;
    const unsigned cmask = FULL ? 0xffffffffu : channel_mask<FT, NV>(p.C, lane);
    const uint8_t* const lane_base = ring_ptr + lane * 16;
    float4 acc[PTS][NV];
    Sample<NV> sa, sb;
    sample_clear<NV>(sa);
    if (DB) sample_clear<NV>(sb);
    int my_pid = -1, my_cnt = 0;  // lane t < PTS: the t-th point of this warp = point warp + WARPS * t of the run
    int64_t run_start = 0;
    int run_seg = -1, run_parity = 0;
    int64_t run_task = 0;
    for (int n = 0;; ++n) {
        const int q = n % kStQ;
        mbar_wait(full0 + 8 * q, (uint32_t)((n / kStQ) & 1));
        const StageSlot& S = slots[q];
        const uint4 h = S.hdr;
        if (h.z & kStExit) break;
        if (h.z & kStFirst) {
            const uint4 hr = S.run;
            run_start = (int64_t)(int)hr.x;
            run_seg = (int)hr.z;
            run_task = (int64_t)h.w;
            const int j = warp + WARPS * lane;
            my_pid = (lane < PTS && j < (int)hr.y) ? S.pid[j] : -1;
            my_cnt = (lane < PTS && j < (int)hr.y) ? S.cnt[j] : 0;
#pragma unroll
            for (int t = 0; t < PTS; ++t) {
#pragma unroll
                for (int k = 0; k < NV; ++k) acc[t][k] = f4_zero();
            }
            if (p.accumulate) {
                if (my_pid >= 0) {
                    const int64_t orow = p.by_pos ? run_start + warp + WARPS * lane : (int64_t)my_pid;
                    my_cnt += p.count[orow];
                }
#pragma unroll
                for (int t = 0; t < PTS; ++t) {
                    const int pt = __shfl_sync(kFull, my_pid, t);
                    if (pt >= 0) {
                        const int64_t orow = p.by_pos ? run_start + warp + WARPS * t : (int64_t)pt;
#pragma unroll
                        for (int k = 0; k < NV; ++k) {
                            const int c = chan_of<FT>(k, lane);
                            if (c < p.C) acc[t][k] = *reinterpret_cast<const float4*>(p.out + orow * p.C + c);
                        }
                    }
                }
            }
        }
        const uint32_t wbits = h.x >> warp;  // bit WARPS * t = this warp's t-th point is in the stage
        if (DB) {
            // two samples in flight: the taps of the next owned point are requested before the current one is blended
            if (wbits & 1u) staged_issue<NV, FT>(sa, S.rec[warp], lane_base, fmap, p.C, row_elems, lane, cmask);
#pragma unroll
            for (int t = 0; t < PTS; ++t) {
                if (t + 1 < PTS && ((wbits >> (WARPS * (t + 1))) & 1u))
                    staged_issue<NV, FT>((t & 1) ? sa : sb, S.rec[warp + WARPS * (t + 1)], lane_base, fmap, p.C, row_elems,
                                         lane, cmask);
                if ((wbits >> (WARPS * t)) & 1u) sample_accum<NV, FAST>(acc[t], (t & 1) ? sb : sa);
            }
        } else {
#pragma unroll
            for (int t = 0; t < PTS; ++t) {
                if ((wbits >> (WARPS * t)) & 1u) {
                    staged_issue<NV, FT>(sa, S.rec[warp + WARPS * t], lane_base, fmap, p.C, row_elems, lane, cmask);
                    sample_accum<NV, FAST>(acc[t], sa);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * q);
        if (h.z & kStLast) {
            // ---- run epilogue: mean, row stores, run partial (same order of operations as gather_kernel) ----
            float4 sp_acc[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) sp_acc[k] = f4_zero();
#pragma unroll
            for (int t = 0; t < PTS; ++t) {
                const int pt = __shfl_sync(kFull, my_pid, t);
                const int cnt = __shfl_sync(kFull, my_cnt, t);
                if (pt < 0) continue;
                const int64_t i = run_start + warp + WARPS * t;
                const int64_t orow = p.by_pos ? i : (int64_t)pt;
                float* out_row = p.out + orow * p.C;
                int32_t* cnt_dst = p.count + orow;
                bool store_row = true;
                if (p.n_peers > 0) {
                    const int owner = (int)(i / p.rows_per_rank);
                    const int64_t slot = (int64_t)p.src_rank * p.rows_per_rank + (i - (int64_t)owner * p.rows_per_rank);
                    out_row = p.peer_out[owner] + slot * p.C;
                    cnt_dst = p.peer_cnt[owner] + slot;
                    store_row = cnt > 0;
                }
                const float denom = (float)max(cnt, 1);
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    const int c = chan_of<FT>(k, lane);
                    if ((FULL || c < p.C) && store_row) {
                        const float4 o = (p.finalize && cnt > 1) ? f4_div(acc[t][k], denom) : acc[t][k];  // x / 1 = x
                        st_cs_f4(out_row + c, o);
                        sp_acc[k] = f4_add(sp_acc[k], o);
                    }
                }
                if (lane == 0) *cnt_dst = cnt;
            }
            if (p.pool && run_seg < p.S) {  // uniform over the consumer warps
                float4* red = sred + (size_t)run_parity * WARPS * NV * 32;
#pragma unroll
                for (int k = 0; k < NV; ++k) red[(warp * NV + k) * 32 + lane] = sp_acc[k];
                asm volatile("bar.sync 1, %0;" ::"r"(WARPS * 32) : "memory");
                if (warp == 0) {
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        float4 tsum = red[k * 32 + lane];
#pragma unroll
                        for (int w = 1; w < WARPS; ++w) tsum = f4_add(tsum, red[(w * NV + k) * 32 + lane]);
                        const int c = chan_of<FT>(k, lane);
                        if (FULL || c < p.C) *reinterpret_cast<float4*>(p.partials + run_task * (int64_t)p.C + c) = tsum;
                    }
                }
                run_parity ^= 1;
            }
        }
    }
}

size_t staged_smem_bytes(int ring_slots, int rowb, int warps, int nv) {
    return (size_t)(ring_slots + 1) * rowb + sizeof(StageSlot) * kStQ + 2 * kStQ * 8 + 128 * 4 + 16 * 4 +
           (size_t)2 * warps * nv * 32 * 16;
}

template <int NV, typename FT, int PTS>
static int launch_staged(const LiftParams& p, StagedParams sp, int variant, bool do_plan, bool do_gather,
                         cudaStream_t stream) {
    constexpr int WARPS = 32 / PTS;
    const int rowb = p.C * (int)sizeof(FT);
    // two CTAs per SM share 228 KB (1 KB of each is reserved by the system)
    const size_t budget = (size_t)(233472 - 2 * 1024) / 2;
    const size_t fixed = staged_smem_bytes(0, rowb, WARPS, NV);
    const int ring = (int)((budget - fixed) / rowb);
    if (ring < 8) return SD3D_ERR_UNSUPPORTED;
    sp.ring_slots = ring;
    sp.cap_pix = min(ring / 2, 126);  // slot ranks travel as bytes inside the records (0x7F = outside the map)
    const size_t smem = staged_smem_bytes(ring, rowb, WARPS, NV);
    const bool fast = (variant & 1) != 0, db = (variant & 4) != 0;
    const bool full = p.C == 128 * NV;  // every lane's NV register vectors hold channels (fp32: 4 each, 16-bit: 8 per 2)
    const int64_t plan_tasks = p.pool ? sp.max_tasks : sp.n_tasks;
    const int nck = (p.n_views + kPlanViews - 1) / kPlanViews;
    if (do_plan)
        stage_plan_kernel<<<(unsigned)ceil_div64(plan_tasks * nck, kPlanWarps), kPlanWarps * 32, 0, stream>>>(p, sp);
    if (!do_gather) return SD3D_OK;
    const unsigned grid = (unsigned)imin64(plan_tasks, (int64_t)2 * num_sms());
    const unsigned threads = (WARPS + 1) * 32;
#define SD3D_LAUNCH_STAGED(FASTV, DBV, FULLV)                                                                     \
    do {                                                                                                          \
        auto kern = gather_staged_kernel<NV, FT, FASTV, PTS, DBV, FULLV>;                                         \
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                       \
        kern<<<grid, threads, smem, stream>>>(p, sp);                                                             \
    } while (0)
#define SD3D_LAUNCH_STAGED2(FASTV, DBV)            \
    do {                                           \
        if (full) SD3D_LAUNCH_STAGED(FASTV, DBV, true); \
        else SD3D_LAUNCH_STAGED(FASTV, DBV, false);     \
    } while (0)
    if (fast && db) SD3D_LAUNCH_STAGED2(true, true);
    else if (fast) SD3D_LAUNCH_STAGED2(true, false);
    else if (db) SD3D_LAUNCH_STAGED2(false, true);
    else SD3D_LAUNCH_STAGED2(false, false);
#undef SD3D_LAUNCH_STAGED2
#undef SD3D_LAUNCH_STAGED
    return SD3D_OK;
}

template <typename FT>
static int dispatch_staged_t(const LiftParams& p, const StagedParams& sp, int variant, bool do_plan, bool do_gather,
                             cudaStream_t stream) {
    constexpr int kR = Tap<FT>::kRegs;
    const int nv = kR == 1 ? (p.C + 127) / 128 : 2 * ((p.C + 255) / 256);
    const bool narrow = (variant & 8) != 0;  // 4 consumer warps x 8 points instead of 8 x 4
    if (nv == 1) {
        if constexpr (kR == 1)
            return narrow ? launch_staged<1, FT, 8>(p, sp, variant, do_plan, do_gather, stream)
                          : launch_staged<1, FT, 4>(p, sp, variant, do_plan, do_gather, stream);
        return SD3D_ERR_UNSUPPORTED;
    }
    if (nv == 2)
        return narrow ? launch_staged<2, FT, 8>(p, sp, variant, do_plan, do_gather, stream)
                      : launch_staged<2, FT, 4>(p, sp, variant, do_plan, do_gather, stream);
    if (nv <= 4) return launch_staged<4, FT, 4>(p, sp, variant, do_plan, do_gather, stream);
    return SD3D_ERR_UNSUPPORTED;
}

bool staged_supported(const LiftParams& p, int fmap_dtype, int n_views) {
    const int sz = fmap_dtype == SD3D_F32 ? 4 : 2;
    const int rowb = p.C * sz;
    return p.order != nullptr && p.run == 32 && n_views > 0 && n_views <= 32767 && rowb % 16 == 0 && rowb <= 2048 &&
           p.C <= 512 && p.Hf <= kMaxMapDim && p.Wf <= kMaxMapDim && p.N < (int64_t(1) << 31) - 64;
}

static size_t staged_cap_stages(int n_views) { return (size_t)((n_views + kPlanViews - 1) / kPlanViews) * kPlanSlots; }

size_t staged_workspace_bytes(int64_t tasks, int n_views) {
    const size_t nck = (size_t)(n_views + kPlanViews - 1) / kPlanViews;
    return (size_t)tasks * staged_cap_stages(n_views) * 64 + (size_t)tasks * 16 + (size_t)tasks * 32 * 8 +
           (size_t)tasks * nck * 4 + (size_t)tasks * 4 + 256;
}

void staged_carve(void* base, int64_t tasks, int n_views, StagedParams& sp) {
    uint8_t* b = reinterpret_cast<uint8_t*>(base);
    const size_t nck = (size_t)(n_views + kPlanViews - 1) / kPlanViews;
    sp.cap_stages = (int)staged_cap_stages(n_views);
    sp.hdrs = reinterpret_cast<uint32_t*>(b);
    b += (size_t)tasks * sp.cap_stages * 64;
    sp.runinfo = reinterpret_cast<int4*>(b);
    b += (size_t)tasks * 16;
    sp.runpts = reinterpret_cast<int2*>(b);
    b += (size_t)tasks * 32 * 8;
    sp.chunk_cnt = reinterpret_cast<int32_t*>(b);
    b += (size_t)tasks * nck * 4;
    sp.done = reinterpret_cast<int32_t*>(b);
    b += (size_t)tasks * 4;
    sp.counter = reinterpret_cast<int32_t*>(b);
    sp.n_done = tasks;
}

int dispatch_staged(const LiftParams& p, const StagedParams& sp, int fmap_dtype, int variant, bool do_plan,
                    bool do_gather, cudaStream_t stream) {
    switch (fmap_dtype) {
        case SD3D_F32: return dispatch_staged_t<float>(p, sp, variant, do_plan, do_gather, stream);
        case SD3D_F16: return dispatch_staged_t<__half>(p, sp, variant, do_plan, do_gather, stream);
        case SD3D_BF16: return dispatch_staged_t<__nv_bfloat16>(p, sp, variant, do_plan, do_gather, stream);
    }
    return SD3D_ERR_UNSUPPORTED;
}

}  // namespace sd3d
