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

constexpr int kStQ = 8;  // stage queue entries per CTA
constexpr uint32_t kStFirst = 1u, kStLast = 2u, kStExit = 4u;
// Sample record after planning (16 bytes, in place of K1's record):
//   staged: x = off00 | off01 << 16, y = off10 | off11 << 16: offsets of the four tap rows from the stage's first ring
//           slot, in 16-byte units (0xFFFF = tap outside the map); z = ax; w = ay, sign bit set if any tap is outside
//   direct: x = pixel index of tap (y0, x0), y = tap-valid flags, z = ax with the sign bit set, w = ay
// (ax, ay are in [0, 1): their sign bits are free)
constexpr uint32_t kSignBit = 0x80000000u;
constexpr uint32_t kOffZero = 0xFFFFu;

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
// stage building (shared by the fused projection kernel and the stand-alone planner)
// ---------------------------------------------------------------------------------------------------
// Stage header, 16 words: [0] view  [1] lane mask  [2] box origin (x | y << 16, int16 each)  [3] distinct pixels
// [4..11] pixel bitmap, row r of the box = bits [16 r, 16 r + 16)  [12..15] 16 bytes: pixels in the rows before row r
//
// One view of one run, lane = point. `cand`: the lane's point is visible in view v with tap origin (x0, y0), tap-valid
// flags, fractional offsets ax / ay and pixel index pix of tap (y0, x0). Cuts the candidates into at most kStMaxSub
// stages (greedy boxes: leftmost candidate, then the topmost candidate within `box` columns of it; the box shrinks
// until its distinct pixels fit `cap_pix`), appends their headers at H and returns their number. Every candidate lane
// gets its final sample record in `out_rec` (see the format above).
__device__ __forceinline__ int build_view_stages(bool cand, int x0, int y0, uint32_t flags, int ax_bits, int ay_bits,
                                                 int pix, int v, uint32_t* __restrict__ H, int cap_pix, uint32_t rowb16,
                                                 int lane, int4& out_rec) {
    int nst = 0;
    for (int sub = 0; sub < kStMaxSub; ++sub) {
        if (!__any_sync(kFull, cand)) break;
        uint32_t bm[8];
        int xmin = 0, ymin = 0, npix = 0, cx = 0, cy = 0;
        bool sel = false, small = false;
        for (int box = 16; box >= 2; box >>= 1) {
            xmin = __reduce_min_sync(kFull, cand ? x0 : 0x3fffffff);
            const bool fitx = cand && (x0 - xmin <= box - 2);
            ymin = __reduce_min_sync(kFull, fitx ? y0 : 0x3fffffff);
            sel = fitx && (y0 - ymin <= box - 2);
            cx = x0 - xmin;
            cy = y0 - ymin;  // 0..14 when sel
            small = !__any_sync(kFull, sel && (cx > 6 || cy > 6));
            const uint32_t top = sel ? (flags & 3u) << cx : 0u, bot = sel ? ((flags >> 2) & 3u) << cx : 0u;
            if (small) {  // the box fits 8 x 8: 64-bit bitmap, two reductions
                const unsigned long long c64 = ((unsigned long long)top << (8 * cy)) | ((unsigned long long)bot << (8 * cy + 8));
                bm[0] = __reduce_or_sync(kFull, (uint32_t)c64);
                bm[1] = __reduce_or_sync(kFull, (uint32_t)(c64 >> 32));
                npix = __popc(bm[0]) + __popc(bm[1]);
            } else {
                npix = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                    uint32_t c = 0u;
                    if ((cy >> 1) == w) c |= top << ((cy & 1) * 16);
                    if (((cy + 1) >> 1) == w) c |= bot << (((cy + 1) & 1) * 16);
                    bm[w] = __reduce_or_sync(kFull, c);
                    npix += __popc(bm[w]);
                }
            }
            if (npix <= cap_pix) break;
        }
        // number of bitmap bits below bit (row, col) = shared-memory slot of that pixel inside the stage
        auto rank_of = [&](int row, int col) -> uint32_t {
            if (small) {
                const int b = row * 8 + col;
                const uint32_t lo_m = b >= 32 ? 0xffffffffu : ((1u << b) - 1u);
                const uint32_t hi_m = b >= 32 ? ((1u << (b - 32)) - 1u) : 0u;
                return (uint32_t)(__popc(bm[0] & lo_m) + __popc(bm[1] & hi_m));
            }
            const int b = row * 16 + col;
            int rk = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                if (w < (b >> 5)) rk += __popc(bm[w]);
                if (w == (b >> 5)) rk += __popc(bm[w] & ((1u << (b & 31)) - 1u));
            }
            return (uint32_t)rk;
        };
        const bool direct = (sub == kStMaxSub - 1) && cand && !sel;  // the view's leftovers ride on its last stage
        const bool member = sel || direct;
        const uint32_t mask = __ballot_sync(kFull, member);
        if (sel) {
            uint32_t off[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
                off[t] = (flags & (1u << t)) ? rank_of(cy + (t >> 1), cx + (t & 1)) * rowb16 : kOffZero;
            out_rec = make_int4((int)(off[0] | (off[1] << 16)), (int)(off[2] | (off[3] << 16)), ax_bits,
                                flags != 15u ? (ay_bits | (int)kSignBit) : ay_bits);
        } else if (direct) {
            out_rec = make_int4(pix, (int)flags, ax_bits | (int)kSignBit, ay_bits);
        }
        // header: lanes 0..11 write one word each, lanes 16..31 the 16 row-prefix bytes
        uint32_t* __restrict__ Hs = H + nst * 16;
        if (lane < 12) {
            uint32_t wv = 0u;
            if (lane == 0) wv = (uint32_t)v;
            if (lane == 1) wv = mask;
            if (lane == 2) wv = ((uint32_t)xmin & 0xffffu) | ((uint32_t)ymin << 16);
            if (lane == 3) wv = (uint32_t)npix;
            if (lane >= 4) {
                const int w = lane - 4;
                if (small) {  // rows 2w, 2w+1 of the 8 x 8 bitmap -> two 16-bit rows
                    const uint32_t two = w < 4 ? ((w < 2 ? bm[0] : bm[1]) >> (16 * (w & 1))) & 0xffffu : 0u;
                    wv = (two & 0xffu) | ((two >> 8) << 16);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (k == w) wv = bm[k];
                }
            }
            Hs[lane] = wv;
        } else if (lane >= 16) {
            const int r = lane - 16;
            reinterpret_cast<uint8_t*>(Hs + 12)[r] = (uint8_t)(r < (small ? 8 : 16) ? rank_of(r, 0) : (uint32_t)npix);
        }
        if (member) cand = false;
        ++nst;
    }
    return nst;
}

// ---------------------------------------------------------------------------------------------------
// K1 + stage planner in one pass (the default when a plan exists): one warp per run, lane = point of the run.
// Projection and depth test exactly as project_kernel (lift.cu: same operations in the same order -> the same
// pix_idx / vis / count bits); the warp then cuts each view's visible samples into stages while their tap geometry is
// still in registers, so the records are written once, in their final form, and the run's headers are appended in
// view order by the warp that owns the run (no atomics, no second pass over the records).
// ---------------------------------------------------------------------------------------------------
constexpr int kPsWarps = 4;
constexpr int kPsViews = 4;  // views per round: 4 independent depth reads in flight per lane

__global__ void __launch_bounds__(kPsWarps * 32) project_stage_kernel(const LiftParams p, uint32_t* __restrict__ masks,
                                                                      int nchunks, const StagedParams sp) {
    const int lane = lane_id();
    const int64_t task = (int64_t)blockIdx.x * kPsWarps + (threadIdx.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x == 0) *sp.counter = 0;  // the gather's run dispenser
    const int64_t n_tasks = p.pool ? (int64_t)p.task_offsets[p.S + 1] : sp.n_tasks;
    if (task >= n_tasks) return;
    int seg, npts;
    int64_t start;
    task_range(p, task, seg, start, npts);
    const bool has = lane < npts;
    const int64_t pid = has ? (int64_t)p.order[start + lane] : 0;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (has) {
        px = __ldg(p.xyz + 3 * pid);
        py = __ldg(p.xyz + 3 * pid + 1);
        pz = __ldg(p.xyz + 3 * pid + 2);
    }
    const int64_t depth_elems = (int64_t)p.Hd * p.Wd;
    const float wd_f = (float)p.Wd, hd_f = (float)p.Hd;
    int4* __restrict__ recs = p.recs + pid * p.n_views;
    uint32_t* __restrict__ mrow = masks + pid * nchunks;
    uint32_t* __restrict__ H = sp.hdrs + task * (int64_t)sp.cap_stages * 16;
    const uint32_t rowb16 = (uint32_t)sp.rowb >> 4;
    int nv = 0, nst = 0;
    uint32_t m = 0u;
    int4 rec0 = make_int4(0, 0, 0, 0);
    for (int v0 = p.v_begin; v0 < p.v_end; v0 += kPsViews) {
        float zc[kPsViews], d[kPsViews], us[kPsViews], ws[kPsViews];
        int cand[kPsViews];
#pragma unroll
        for (int t = 0; t < kPsViews; ++t) {  // phase 1: project, issue the depth reads (camera loads are uniform)
            const int v = v0 + t;
            cand[t] = -1;
            d[t] = zc[t] = us[t] = ws[t] = 0.f;
            if (v < p.v_end && has) {
                const float4 k4 = ldg_f4(p.K4 + 4 * (int64_t)v);
                const float4 r0 = ldg_f4(p.w2c + 12 * (int64_t)v);
                const float4 r1 = ldg_f4(p.w2c + 12 * (int64_t)v + 4);
                const float4 r2 = ldg_f4(p.w2c + 12 * (int64_t)v + 8);
                zc[t] = project_point(k4, r0, r1, r2, px, py, pz, p.z_near, us[t], ws[t]);
                if (zc[t] > p.z_near) {
                    const float uif = floorf(__fadd_rn(us[t], 0.5f));
                    const float wif = floorf(__fadd_rn(ws[t], 0.5f));
                    if (uif >= 0.f && uif < wd_f && wif >= 0.f && wif < hd_f) {
                        cand[t] = (int)wif * p.Wd + (int)uif;
                        if (p.depth_u16)
                            d[t] = __fmul_rn((float)__ldg(reinterpret_cast<const uint16_t*>(p.depth) +
                                                          (int64_t)v * depth_elems + cand[t]),
                                             0.001f);
                        else
                            d[t] = __ldg(reinterpret_cast<const float*>(p.depth) + (int64_t)v * depth_elems + cand[t]);
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < kPsViews; ++t) {  // phase 2: depth test, mask bit, stages + sample record
            const int v = v0 + t;
            if (v >= p.v_end) break;
            const bool visible = has && cand[t] >= 0 && d[t] > 0.f && fabsf(__fsub_rn(d[t], zc[t])) <= p.tau;
            if (has) {
                if (p.pix_idx) p.pix_idx[(int64_t)v * p.N + pid] = visible ? cand[t] : -1;
                if (p.vis) p.vis[(int64_t)v * p.N + pid] = visible ? 1 : 0;
            }
            const int bit = (v - p.v_begin) & 31;
            if (visible) m |= 1u << bit;
            if (__any_sync(kFull, visible)) {
                TapGeom g;
                g.x0 = g.y0 = 0;
                g.ax = g.ay = 0.f;
                g.flags = 0u;
                if (visible) g = tap_geometry(us[t], ws[t], p.stride, p.inv_stride, p.Hf, p.Wf);
                int4 rec = make_int4(0, 0, 0, 0);
                nst += build_view_stages(visible, g.x0, g.y0, g.flags, __float_as_int(g.ax), __float_as_int(g.ay),
                                         (v * p.Hf + g.y0) * p.Wf + g.x0, v, H + nst * 16, sp.cap_pix, rowb16, lane, rec);
                if (visible) {
                    recs[nv] = rec;
                    if (nv == 0) rec0 = rec;
                    ++nv;
                }
            }
            if (has && (bit == 31 || v == p.v_end - 1)) {
                mrow[(v - p.v_begin) >> 5] = m;
                m = 0u;
            }
        }
    }
    if (nst == 0) {  // nothing of this run is visible: one empty stage carries the run through the pipeline
        if (lane < 16) H[lane] = 0u;
        nst = 1;
    }
    if (has) p.nvis[pid] = nv;
    int4* __restrict__ rp = sp.runpts + (task * 32 + lane) * 2;
    rp[0] = make_int4(has ? (int)pid : -1, nv, 0, 0);
    rp[1] = rec0;
    if (lane == 0) sp.runinfo[task] = make_int4((int)start, npts, seg, nst);
}

// ---------------------------------------------------------------------------------------------------
// stand-alone stage planner (used when the projection ran without a plan, e.g. concurrently with the plan kernels):
// one warp per (run, chunk of 8 views), lane = point of the run; re-reads K1's records
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
    int4* __restrict__ recs = p.recs + (int64_t)max(pid, 0) * p.n_views;
    int nst = 0;
    const uint32_t any = __reduce_or_sync(kFull, bits);
    if (any) {
        int4 recv[kPlanViews];  // the lane's record of chunk view i (if visible): all loads in flight together
#pragma unroll
        for (int i = 0; i < kPlanViews; ++i) {
            recv[i] = make_int4(0, 0, 0, 0);
            if (bits & (1u << i)) recv[i] = recs[below + __popc(bits & ((1u << i) - 1u))];
        }
        uint32_t* __restrict__ H = sp.hdrs + (task * (int64_t)sp.cap_stages + (int64_t)ck * kPlanSlots) * 16;
        const uint32_t rowb16 = (uint32_t)sp.rowb >> 4;
#pragma unroll 1
        for (int i = 0; i < kPlanViews; ++i) {
            if (!((any >> i) & 1u)) continue;
            int4 rec = make_int4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < kPlanViews; ++j)
                if (j == i) rec = recv[j];
            const bool cand = (bits >> i) & 1u;
            int4 out = rec;
            nst += build_view_stages(cand, rec_x0(rec.y), rec_y0(rec.y), (uint32_t)rec.y & 15u, rec.z, rec.w, rec.x,
                                     p.v_begin + v_lo + i, H + nst * 16, sp.cap_pix, rowb16, lane, out);
            if (cand) recs[below + __popc(bits & ((1u << i) - 1u))] = out;
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
        uint32_t todo = __ballot_sync(kFull, cnt > 0 && dst0 != c * kPlanSlots);
        while (todo) {  // chunk c0 + l: its headers move from slot (c0 + l) * kPlanSlots to dst (dst < src)
            const int l = __ffs(todo) - 1;
            todo &= todo - 1u;
            const int n_l = __shfl_sync(kFull, cnt, l), d_l = __shfl_sync(kFull, dst0, l);
            const int s_l = (c0 + l) * kPlanSlots;
            for (int j0 = 0; j0 < n_l; j0 += 8) {  // 8 headers (4 x 16 bytes each) per pass: load all, then store
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
    const int nv = has ? p.nvis[pid] : 0;
    int4 rec0 = make_int4(0, 0, 0, 0);
    if (nv > 0) rec0 = __ldcg(recs);  // as rewritten by the planner warp of view chunk 0
    int4* __restrict__ rp = sp.runpts + (task * 32 + lane) * 2;
    rp[0] = make_int4(pid, nv, 0, 0);
    rp[1] = rec0;
    if (lane == 0) sp.runinfo[task] = make_int4((int)start, npts, seg, total);
}

// ---------------------------------------------------------------------------------------------------
// staged gather
// ---------------------------------------------------------------------------------------------------
struct __align__(16) StageSlot {
    uint4 hdr;        // mask, first ring slot, flags, task
    uint4 run;        // kStFirst only: first processing position, points, segment, -
    uint4 rec[32];    // kStFirst only: first sample record of every point of the run
    int32_t pid[32];  // kStFirst only
    int32_t cnt[32];  // kStFirst only: visible views of the point in this call
};

template <int NV, typename FT>
__device__ __forceinline__ void staged_issue(Sample<NV, FT>& s, const uint4 r, const uint8_t* __restrict__ stage_ptr,
                                             const uint8_t* __restrict__ zero_ptr, const FT* __restrict__ fmap, int C,
                                             int row_elems, int lane, unsigned cmask) {
    constexpr int kR = Tap<FT>::kRegs;
    const float ax = __uint_as_float(r.z & ~kSignBit), ay = __uint_as_float(r.w & ~kSignBit);
    const float omx = __fsub_rn(1.0f, ax), omy = __fsub_rn(1.0f, ay);
    s.w00 = __fmul_rn(omx, omy);
    s.w01 = __fmul_rn(ax, omy);
    s.w10 = __fmul_rn(omx, ay);
    s.w11 = __fmul_rn(ax, ay);
    if (r.z & kSignBit) {  // warp-uniform, rare: taps from global memory (predicated loads, zero outside the map)
        constexpr int kE = Tap<FT>::kElems;
        const FT* __restrict__ g00 = fmap + (int64_t)(int)r.x * C + lane * kE;
        const FT* __restrict__ g10 = g00 + row_elems;
#pragma unroll
        for (int l = 0; l < NV / kR; ++l) {
            s.t00[l] = s.t01[l] = s.t10[l] = s.t11[l] = Tap<FT>::zero();
            const bool cok = (cmask >> l) & 1u;
            if (cok && (r.y & 1u)) s.t00[l] = Tap<FT>::load(g00 + l * 32 * kE);
            if (cok && (r.y & 2u)) s.t01[l] = Tap<FT>::load(g00 + C + l * 32 * kE);
            if (cok && (r.y & 4u)) s.t10[l] = Tap<FT>::load(g10 + l * 32 * kE);
            if (cok && (r.y & 8u)) s.t11[l] = Tap<FT>::load(g10 + C + l * 32 * kE);
        }
        return;
    }
    const uint8_t* __restrict__ p00 = stage_ptr + ((r.x & 0xffffu) << 4);
    const uint8_t* __restrict__ p01 = stage_ptr + ((r.x >> 16) << 4);
    const uint8_t* __restrict__ p10 = stage_ptr + ((r.y & 0xffffu) << 4);
    const uint8_t* __restrict__ p11 = stage_ptr + ((r.y >> 16) << 4);
    if (r.w & kSignBit) {  // warp-uniform, map border only: taps outside the map read the zero row
        if ((r.x & 0xffffu) == kOffZero) p00 = zero_ptr;
        if ((r.x >> 16) == kOffZero) p01 = zero_ptr;
        if ((r.y & 0xffffu) == kOffZero) p10 = zero_ptr;
        if ((r.y >> 16) == kOffZero) p11 = zero_ptr;
    }
#pragma unroll
    for (int l = 0; l < NV / kR; ++l) {
        if (cmask & (1u << l)) {
            s.t00[l] = Tap<FT>::from_smem(p00 + l * 512);
            s.t01[l] = Tap<FT>::from_smem(p01 + l * 512);
            s.t10[l] = Tap<FT>::from_smem(p10 + l * 512);
            s.t11[l] = Tap<FT>::from_smem(p11 + l * 512);
        }
    }
}

// FULL: C fills all NV register vectors of every lane (C = 128 * NV): no channel predicates
template <int NV, typename FT, bool FAST, int PTS, bool DB, bool FULL>
__global__ void __launch_bounds__((32 / PTS + 1) * 32, 2)
    gather_staged_kernel(const __grid_constant__ LiftParams p, const __grid_constant__ StagedParams sp) {
    constexpr int WARPS = 32 / PTS;  // consumer warps; warp WARPS is the producer
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int rowb = p.C * (int)sizeof(FT);
    const int ring = sp.ring_slots;
    // shared memory: ring rows | zero row | stage slots | barriers | producer scratch | record scratch | reduce scratch
    uint8_t* const ring_ptr = smem;
    StageSlot* const slots = reinterpret_cast<StageSlot*>(smem + (size_t)(ring + 1) * rowb);
    uint64_t* const bars = reinterpret_cast<uint64_t*>(slots + kStQ);        // full[kStQ], empty[kStQ], red_full, red_empty
    uint32_t* const hch = reinterpret_cast<uint32_t*>(bars + 2 * kStQ + 2);  // 8 headers x 16 words
    int32_t* const ext = reinterpret_cast<int32_t*>(hch + 128);              // ring slots held by stage entry q
    float4* const sred = reinterpret_cast<float4*>(ext + 16);                // [WARPS][NV*32]
    const uint32_t full0 = smem_addr(bars), empty0 = smem_addr(bars + kStQ);
    const uint32_t red_full = smem_addr(bars + 2 * kStQ), red_empty = red_full + 8;

    for (int i = threadIdx.x; i < rowb / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(ring_ptr + (size_t)ring * rowb)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (threadIdx.x == 0) {
        for (int q = 0; q < kStQ; ++q) {
            mbar_init(full0 + 8 * q, 1);
            mbar_init(empty0 + 8 * q, WARPS);
        }
        mbar_init(red_full, WARPS - 1);
        mbar_init(red_empty, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const FT* __restrict__ fmap = reinterpret_cast<const FT*>(p.fmap);
    const int64_t n_tasks = p.pool ? (int64_t)p.task_offsets[p.S + 1] : sp.n_tasks;

    if (warp == WARPS) {
        // ------------------------------------------------ producer ------------------------------------------------
        const uint32_t ring_s = smem_addr(ring_ptr);
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
        // run pipeline: c_next = ticket of the run after the current one; its descriptors are loaded one run ahead
        int c_cur = 0, c_next = 0;
        if (lane == 0) {
            c_cur = atomicAdd(sp.counter, 1);
            c_next = atomicAdd(sp.counter, 1);
        }
        c_cur = __shfl_sync(kFull, c_cur, 0);
        c_next = __shfl_sync(kFull, c_next, 0);
        auto task_of = [&](int c) -> int64_t { return ((int64_t)c + sp.task_rot) % n_tasks; };
        const int hlanes = min(8, sp.cap_stages) * 4;
        int4 ri = make_int4(0, 0, 0, 0), rpa = make_int4(-1, 0, 0, 0), rec0 = make_int4(0, 0, 0, 0);
        uint4 hc = make_uint4(0u, 0u, 0u, 0u);
        if (c_cur < n_tasks) {
            const int64_t t = task_of(c_cur);
            ri = sp.runinfo[t];
            rpa = sp.runpts[(t * 32 + lane) * 2];
            rec0 = sp.runpts[(t * 32 + lane) * 2 + 1];
            if (lane < hlanes) hc = reinterpret_cast<const uint4*>(sp.hdrs + t * (int64_t)sp.cap_stages * 16)[lane];
        }
        while (c_cur < n_tasks) {
            const int64_t task = task_of(c_cur);
            const int nst = ri.w, nv = rpa.y;
            const uint4* __restrict__ hsrc = reinterpret_cast<const uint4*>(sp.hdrs + task * (int64_t)sp.cap_stages * 16);
            // lane = point of the run: its next sample record (the first one came with the run descriptor)
            const int4* __restrict__ recs = p.recs + (int64_t)max(rpa.x, 0) * p.n_views;
            int4 rec = rec0;
            int k = 0;
            // prefetch the next run's descriptors and the ticket after it
            const int c_after_l = (lane == 0) ? atomicAdd(sp.counter, 1) : 0;
            int4 ri_n = make_int4(0, 0, 0, 0), rpa_n = make_int4(-1, 0, 0, 0), rec0_n = make_int4(0, 0, 0, 0);
            uint4 hc_n = make_uint4(0u, 0u, 0u, 0u);
            if (c_next < n_tasks) {
                const int64_t t = task_of(c_next);
                ri_n = sp.runinfo[t];
                rpa_n = sp.runpts[(t * 32 + lane) * 2];
                rec0_n = sp.runpts[(t * 32 + lane) * 2 + 1];
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
                    const uint4 h0 = *reinterpret_cast<const uint4*>(Hd);  // view, mask, box origin, pixels
                    const int npix = (int)h0.w;
                    const int xmin = (int)(int16_t)(h0.z & 0xffffu), ymin = (int)(int16_t)(h0.z >> 16);
                    const int q = n % kStQ;
                    const int base = acquire(npix);
                    StageSlot& S = slots[q];
                    const uint32_t flags = ((s0 + s == 0) ? kStFirst : 0u) | ((s0 + s == nst - 1) ? kStLast : 0u);
                    if (flags & kStFirst) {
                        S.pid[lane] = rpa.x;
                        S.cnt[lane] = nv;
                        if (lane == 0) S.run = make_uint4((uint32_t)ri.x, (uint32_t)ri.y, (uint32_t)ri.z, 0u);
                    }
                    if ((h0.y >> lane) & 1u) {
                        S.rec[lane] = make_uint4((uint32_t)rec.x, (uint32_t)rec.y, (uint32_t)rec.z, (uint32_t)rec.w);
                        if (++k < nv) rec = recs[k];
                    }
                    __syncwarp();
                    const uint32_t fullb = full0 + 8 * q;
                    if (lane == 0) {
                        S.hdr = make_uint4(h0.y, (uint32_t)base, flags, (uint32_t)task);
                        mbar_arrive_expect_tx(fullb, (uint32_t)npix * (uint32_t)rowb);
                    }
                    __syncwarp();
                    if (npix > 0 && lane < 16) {  // bitmap rows -> runs of consecutive pixels -> one bulk copy each
                        const int r = lane;
                        uint32_t bits = (Hd[4 + (r >> 1)] >> ((r & 1) * 16)) & 0xffffu;
                        if (bits) {
                            const uint32_t rowbase = (Hd[12 + (r >> 2)] >> (8 * (r & 3))) & 0xffu;
                            const uint8_t* src_row = reinterpret_cast<const uint8_t*>(fmap) +
                                                     (((int64_t)h0.x * p.Hf + (ymin + r)) * p.Wf + xmin) * (int64_t)rowb;
                            uint32_t dst = ring_s + ((uint32_t)base + rowbase) * (uint32_t)rowb;
                            while (bits) {
                                const int st = __ffs(bits) - 1;
                                const int len = __ffs(~(bits >> st)) - 1;
                                bulk_copy_g2s(dst, src_row + (int64_t)st * rowb, (uint32_t)len * (uint32_t)rowb, fullb);
                                dst += (uint32_t)len * (uint32_t)rowb;
                                bits &= ~(((1u << len) - 1u) << st);
                            }
                        }
                    }
                    ++n;
                }
            }
            c_cur = c_next;
            c_next = __shfl_sync(kFull, c_after_l, 0);
            ri = ri_n;
            rpa = rpa_n;
            rec0 = rec0_n;
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
    const uint8_t* const zero_ptr = lane_base + (size_t)ring * rowb;
    float4 acc[PTS][NV];
    Sample<NV, FT> sa, sb;
    sample_clear<NV, FT>(sa);
    if (DB) sample_clear<NV, FT>(sb);
    int my_pid = -1, my_cnt = 0;  // lane t < PTS: the t-th point of this warp = point warp + WARPS * t of the run
    int run_start = 0, run_seg = -1, run_task = 0, red_n = 0;
    for (int n = 0;; ++n) {
        const int q = n % kStQ;
        mbar_wait(full0 + 8 * q, (uint32_t)((n / kStQ) & 1));
        const StageSlot& S = slots[q];
        const uint4 h = S.hdr;
        if (h.z & kStExit) break;
        if (h.z & kStFirst) {
            const uint4 hr = S.run;
            run_start = (int)hr.x;
            run_seg = (int)hr.z;
            run_task = (int)h.w;
            const int j = warp + WARPS * lane;
            const bool mine = lane < PTS && j < (int)hr.y;
            my_pid = mine ? S.pid[j] : -1;
            my_cnt = mine ? S.cnt[j] : 0;
#pragma unroll
            for (int t = 0; t < PTS; ++t) {
#pragma unroll
                for (int k = 0; k < NV; ++k) acc[t][k] = f4_zero();
            }
            if (p.accumulate) {
                if (my_pid >= 0) {
                    const int64_t orow = p.by_pos ? (int64_t)run_start + warp + WARPS * lane : (int64_t)my_pid;
                    my_cnt += p.count[orow];
                }
#pragma unroll
                for (int t = 0; t < PTS; ++t) {
                    const int pt = __shfl_sync(kFull, my_pid, t);
                    if (pt >= 0) {
                        const int64_t orow = p.by_pos ? (int64_t)run_start + warp + WARPS * t : (int64_t)pt;
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
        const uint8_t* const stage_ptr = lane_base + h.y * (uint32_t)rowb;
        if (DB) {
            // two samples in flight: the taps of the next owned point are requested before the current one is blended
            if (wbits & 1u) staged_issue<NV, FT>(sa, S.rec[warp], stage_ptr, zero_ptr, fmap, p.C, row_elems, lane, cmask);
#pragma unroll
            for (int t = 0; t < PTS; ++t) {
                if (t + 1 < PTS && ((wbits >> (WARPS * (t + 1))) & 1u))
                    staged_issue<NV, FT>((t & 1) ? sa : sb, S.rec[warp + WARPS * (t + 1)], stage_ptr, zero_ptr, fmap, p.C,
                                         row_elems, lane, cmask);
                if ((wbits >> (WARPS * t)) & 1u) sample_accum<NV, FAST, FT>(acc[t], (t & 1) ? sb : sa);
            }
        } else {
#pragma unroll
            for (int t = 0; t < PTS; ++t) {
                if ((wbits >> (WARPS * t)) & 1u) {
                    staged_issue<NV, FT>(sa, S.rec[warp + WARPS * t], stage_ptr, zero_ptr, fmap, p.C, row_elems, lane, cmask);
                    sample_accum<NV, FAST, FT>(acc[t], sa);
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
                const int64_t i = (int64_t)run_start + warp + WARPS * t;
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
                // warp 0 adds the warps' rows in warp order. Hand-off through two mbarriers, so that a warp only ever
                // waits for what it really depends on: warp 0 for the seven rows, the others for warp 0 having read
                // their previous row before they overwrite it.
                if (warp != 0 && red_n > 0) mbar_wait(red_empty, (uint32_t)((red_n - 1) & 1));
#pragma unroll
                for (int k = 0; k < NV; ++k) sred[(warp * NV + k) * 32 + lane] = sp_acc[k];
                __syncwarp();
                if (warp != 0) {
                    if (lane == 0) mbar_arrive(red_full);
                } else {
                    mbar_wait(red_full, (uint32_t)(red_n & 1));
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        float4 tsum = sred[k * 32 + lane];
#pragma unroll
                        for (int w = 1; w < WARPS; ++w) tsum = f4_add(tsum, sred[(w * NV + k) * 32 + lane]);
                        const int c = chan_of<FT>(k, lane);
                        if (FULL || c < p.C) *reinterpret_cast<float4*>(p.partials + run_task * (int64_t)p.C + c) = tsum;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(red_empty);
                }
                ++red_n;
            }
        }
    }
}

size_t staged_smem_bytes(int ring_slots, int rowb, int warps, int nv) {
    return (size_t)(ring_slots + 1) * rowb + sizeof(StageSlot) * kStQ + (2 * kStQ + 2) * 8 + 128 * 4 + 16 * 4 +
           (size_t)warps * nv * 32 * 16;
}

template <int NV, typename FT, int PTS>
static int launch_staged(const LiftParams& p, StagedParams sp, int variant, int plan_mode, uint32_t* masks,
                         bool do_gather, cudaStream_t stream) {
    constexpr int WARPS = 32 / PTS;
    const int rowb = p.C * (int)sizeof(FT);
    // two CTAs per SM share 228 KB (1 KB of each is reserved by the system)
    const size_t budget = (size_t)(233472 - 2 * 1024) / 2;
    const size_t fixed = staged_smem_bytes(0, rowb, WARPS, NV);
    const int ring = (int)((budget - fixed) / rowb);
    if (ring < 8) return SD3D_ERR_UNSUPPORTED;
    sp.ring_slots = ring;
    sp.cap_pix = min(ring / 2, 255);  // the header's row-prefix counts are bytes
    sp.rowb = rowb;
    const size_t smem = staged_smem_bytes(ring, rowb, WARPS, NV);
    const bool fast = (variant & 1) != 0, db = (variant & 4) != 0;
    const bool full = p.C == 128 * NV;  // every lane's NV register vectors hold channels (fp32: 4 each, 16-bit: 8 per 2)
    const int64_t plan_tasks = p.pool ? sp.max_tasks : sp.n_tasks;
    const int nck = (p.n_views + kPlanViews - 1) / kPlanViews;
    if (plan_mode == 2)  // projection + stage planning in one pass over the runs
        project_stage_kernel<<<(unsigned)ceil_div64(plan_tasks, kPsWarps), kPsWarps * 32, 0, stream>>>(p, masks, sp.nchunks, sp);
    else if (plan_mode == 1)  // records of an earlier projection-only call
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
static int dispatch_staged_t(const LiftParams& p, const StagedParams& sp, int variant, int plan_mode, uint32_t* masks,
                             bool do_gather, cudaStream_t stream) {
    constexpr int kR = Tap<FT>::kRegs;
    const int nv = kR == 1 ? (p.C + 127) / 128 : 2 * ((p.C + 255) / 256);
    const bool narrow = (variant & 8) != 0;  // 4 consumer warps x 8 points instead of 8 x 4
    if (nv == 1) {
        if constexpr (kR == 1)
            return narrow ? launch_staged<1, FT, 8>(p, sp, variant, plan_mode, masks, do_gather, stream)
                          : launch_staged<1, FT, 4>(p, sp, variant, plan_mode, masks, do_gather, stream);
        return SD3D_ERR_UNSUPPORTED;
    }
    if (nv == 2)
        return narrow ? launch_staged<2, FT, 8>(p, sp, variant, plan_mode, masks, do_gather, stream)
                      : launch_staged<2, FT, 4>(p, sp, variant, plan_mode, masks, do_gather, stream);
    if (nv <= 4) return launch_staged<4, FT, 4>(p, sp, variant, plan_mode, masks, do_gather, stream);
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
    return (size_t)tasks * staged_cap_stages(n_views) * 64 + (size_t)tasks * 16 + (size_t)tasks * 32 * 32 +
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
    sp.runpts = reinterpret_cast<int4*>(b);
    b += (size_t)tasks * 32 * 32;
    sp.chunk_cnt = reinterpret_cast<int32_t*>(b);
    b += (size_t)tasks * nck * 4;
    sp.done = reinterpret_cast<int32_t*>(b);
    b += (size_t)tasks * 4;
    sp.counter = reinterpret_cast<int32_t*>(b);
    sp.n_done = tasks;
}

int dispatch_staged(const LiftParams& p, const StagedParams& sp, int fmap_dtype, int variant, int plan_mode,
                    uint32_t* masks, bool do_gather, cudaStream_t stream) {
    switch (fmap_dtype) {
        case SD3D_F32: return dispatch_staged_t<float>(p, sp, variant, plan_mode, masks, do_gather, stream);
        case SD3D_F16: return dispatch_staged_t<__half>(p, sp, variant, plan_mode, masks, do_gather, stream);
        case SD3D_BF16: return dispatch_staged_t<__nv_bfloat16>(p, sp, variant, plan_mode, masks, do_gather, stream);
    }
    return SD3D_ERR_UNSUPPORTED;
}

}  // namespace sd3d
