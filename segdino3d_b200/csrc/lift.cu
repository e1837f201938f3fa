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
#include "lift_common.cuh"
#include "lift_staged.cuh"

namespace sd3d {

// ---------------------------------------------------------------------------------------------------
// K1: projection + depth-visibility (step a-1). LANE = POINT: a thread walks the views of its point, four at a time
// (four independent depth reads in flight), so every lane is busy whatever the number of views, camera matrices are
// warp-uniform loads, and no shuffles are needed: the thread builds its point's visibility mask words and appends
// one sample record per visible view (ascending view order) for K2. Integer outputs + records only.
// ---------------------------------------------------------------------------------------------------
constexpr int kProjThreads = 64;
constexpr int kProjViews = 4;  // views per round: 4 independent depth reads in flight per lane
constexpr int kMaxNearest = 8;  // largest k of the nearest-view sampling variant

template <bool NEAREST>
__global__ void __launch_bounds__(kProjThreads) project_kernel(const LiftParams p, uint32_t* __restrict__ masks,
                                                               int nchunks, int32_t* __restrict__ plan_done,
                                                               int64_t n_done) {
    const int64_t pos = (int64_t)blockIdx.x * kProjThreads + threadIdx.x;
    // arrival counters of the stage planner (lift_staged.cu), which runs after this kernel
    for (int64_t i = pos; i < n_done; i += (int64_t)gridDim.x * kProjThreads) plan_done[i] = 0;
    if (pos >= p.N) return;
    // consecutive lanes = consecutive points of the processing order: with a plan, spatial neighbours read
    // neighbouring depth pixels (same 32-byte sectors); the optional [V, N] maps are written coalesced
    const int64_t pid = p.order ? (int64_t)p.order[pos] : pos;
    const float px = __ldg(p.xyz + 3 * pid), py = __ldg(p.xyz + 3 * pid + 1), pz = __ldg(p.xyz + 3 * pid + 2);
    const int64_t depth_elems = (int64_t)p.Hd * p.Wd;
    const float wd_f = (float)p.Wd, hd_f = (float)p.Hd;
    int4* __restrict__ recs = p.recs + pid * p.n_views;
    uint32_t* __restrict__ mrow = masks + pid * nchunks;
    int nv = 0;        // visible views so far = slot of the next record
    uint32_t m = 0u;   // mask word being filled
    // nearest-view sampling (k_views > 0): the k visible views with the smallest camera depth, ties to the lower view
    float best_z[kMaxNearest];
    int best_v[kMaxNearest];
#pragma unroll
    for (int j = 0; j < kMaxNearest; ++j) {
        best_z[j] = 0.f;
        best_v[j] = -1;
    }
    for (int v0 = p.v_begin; v0 < p.v_end; v0 += kProjViews) {
        float zc[kProjViews], d[kProjViews], us[kProjViews], ws[kProjViews];
        int cand[kProjViews];
#pragma unroll
        for (int t = 0; t < kProjViews; ++t) {  // phase 1: project, issue the depth reads (camera loads are uniform)
            const int v = v0 + t;
            cand[t] = -1;
            d[t] = zc[t] = us[t] = ws[t] = 0.f;
            if (v < p.v_end) {
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
        for (int t = 0; t < kProjViews; ++t) {  // phase 2: depth test, mask bit, sample record
            const int v = v0 + t;
            if (v < p.v_end) {
                const bool visible = cand[t] >= 0 && d[t] > 0.f && fabsf(__fsub_rn(d[t], zc[t])) <= p.tau;
                if (p.pix_idx) p.pix_idx[(int64_t)v * p.N + pid] = visible ? cand[t] : -1;
                if (p.vis) p.vis[(int64_t)v * p.N + pid] = visible ? 1 : 0;
                const int bit = (v - p.v_begin) & 31;
                if (visible) {
                    const TapGeom g = tap_geometry(us[t], ws[t], p.stride, p.inv_stride, p.Hf, p.Wf);
                    recs[nv++] = make_record(g, g.ax, g.ay, v, p.Hf, p.Wf);
                    m |= 1u << bit;
                    if (NEAREST) {  // keep (zc, v) if it is among the k smallest so far: evict the largest (zc, v)
                        int worst = -1, wv = -2;  // slot to replace; wv = its view (-1: a free slot, always taken)
                        float wz = 0.f;
#pragma unroll
                        for (int j = 0; j < kMaxNearest; ++j) {
                            if (j < p.k_views && wv != -1) {
                                if (best_v[j] < 0) {
                                    worst = j;
                                    wv = -1;
                                } else if (worst < 0 || best_z[j] > wz || (best_z[j] == wz && best_v[j] > wv)) {
                                    worst = j;
                                    wz = best_z[j];
                                    wv = best_v[j];
                                }
                            }
                        }
                        const bool take = wv == -1 || zc[t] < wz;
#pragma unroll
                        for (int j = 0; j < kMaxNearest; ++j)
                            if (j == worst && take) {
                                best_z[j] = zc[t];
                                best_v[j] = v;
                            }
                    }
                }
                if (bit == 31 || v == p.v_end - 1) {
                    mrow[(v - p.v_begin) >> 5] = m;
                    m = 0u;
                }
            }
        }
    }
    if (NEAREST && nv > p.k_views) {
        // keep only the selected views' records (the i-th set mask bit is the i-th record), still in ascending view order
        int i = 0, j = 0;
        for (int c = 0; c < nchunks; ++c) {
            uint32_t w = mrow[c];
            while (w) {
                const int v = p.v_begin + c * 32 + __ffs(w) - 1;
                w &= w - 1u;
                bool sel = false;
#pragma unroll
                for (int q = 0; q < kMaxNearest; ++q) sel |= best_v[q] == v;
                if (sel) {
                    if (j != i) recs[j] = recs[i];
                    ++j;
                }
                ++i;
            }
        }
        nv = j;
    }
    p.nvis[pid] = nv;
}

// ---------------------------------------------------------------------------------------------------
// K2 (default): bilinear gather + view sum + mean (+ run partial for the fused superpoint pooling), steps a-2/a-3.
// CTA = 4 warps = one run of <= `run` consecutive points of the processing order; warp w takes points
// w, w+4, ... of the run (neighbouring points at the same time -> shared L1 lines). Per point, lane r
// re-projects the r-th visible view (cheap ALU, no depth read), then the warp walks the samples in
// ascending view order with LANE = CHANNEL VECTOR and a two-deep register prefetch (sample i+1's tap
// rows are in flight while sample i is blended).
// ---------------------------------------------------------------------------------------------------
template <int NV, typename FT, bool FAST, bool PREFETCH, int MINB>
__global__ void __launch_bounds__(kLiftThreads, MINB) gather_kernel(const __grid_constant__ LiftParams p, const uint32_t* __restrict__ masks,
                                                              int nchunks) {
    __shared__ float4 s_red[kLiftWarps][NV * 32];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int64_t task = blockIdx.x;
    int64_t start, end;
    int seg = -1;
    if (p.pool) {
        const int32_t n_tasks = p.task_offsets[p.S + 1];
        if (task >= n_tasks) return;
        seg = p.task_seg[task];
        start = (int64_t)p.seg_offsets[seg] + (task - p.task_offsets[seg]) * (int64_t)p.run;
        end = min(start + (int64_t)p.run, (int64_t)p.seg_offsets[seg + 1]);
    } else {
        // push mode rotates the chunk order by rank (task_rot) so that at any moment the ranks store into
        // DIFFERENT owners: a balanced all-to-all instead of an incast on one rank's NVLink ingress
        start = (int64_t)((blockIdx.x + (unsigned)p.task_rot) % gridDim.x) * p.run;
        if (start >= p.N) return;
        end = min(start + (int64_t)p.run, p.N);
    }
    const FT* __restrict__ fmap = reinterpret_cast<const FT*>(p.fmap);
    const int row_elems = p.Wf * p.C;
    const unsigned cmask = channel_mask<FT, NV>(p.C, lane);

    float4 sp_acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) sp_acc[k] = f4_zero();
    Sample<NV, FT> sa, sb;  // lanes beyond C never load: give their registers a defined value once
    sample_clear<NV, FT>(sa);
    sample_clear<NV, FT>(sb);

    for (int64_t i = start + warp; i < end; i += kLiftWarps) {
        const int32_t pid = p.order ? p.order[i] : (int32_t)i;
        const int64_t orow = p.by_pos ? i : (int64_t)pid;  // output row
        const int n_total = nchunks > 0 ? p.nvis[pid] : 0;
        const int4* __restrict__ recs = p.recs + (int64_t)pid * p.n_views;
        float4 acc[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] = f4_zero();
        int cnt = n_total;
        if (p.accumulate) {
            cnt += p.count[orow];
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const int c = chan_of<FT>(k, lane);
                if (c < p.C) acc[k] = *reinterpret_cast<const float4*>(p.out + orow * p.C + c);
            }
        }
        for (int r0 = 0; r0 < n_total; r0 += 32) {
            // lane r owns the (r0+r)-th visible view of this point (ascending view order): K1 left its sample
            // record (tap pixel, flags, fractional offsets) in the workspace
            SampleScalars mine;
            scalars_clear(mine);
            if (r0 + lane < n_total) mine = scalars_from_record<FT>(recs[r0 + lane], fmap, p.C);
            const int n_round = min(32, n_total - r0);
            if (PREFETCH) {
                sample_issue<NV, FT>(sa, mine, 0, p.C, row_elems, lane, cmask);
                int sidx = 0;
                while (true) {
                    if (sidx + 1 < n_round) sample_issue<NV, FT>(sb, mine, sidx + 1, p.C, row_elems, lane, cmask);
                    sample_accum<NV, FAST, FT>(acc, sa);
                    if (++sidx >= n_round) break;
                    if (sidx + 1 < n_round) sample_issue<NV, FT>(sa, mine, sidx + 1, p.C, row_elems, lane, cmask);
                    sample_accum<NV, FAST, FT>(acc, sb);
                    if (++sidx >= n_round) break;
                }
            } else {  // wide rows (C > 512): one sample in flight, the tap rows alone fill the register file
                for (int sidx = 0; sidx < n_round; ++sidx) {
                    sample_issue<NV, FT>(sa, mine, sidx, p.C, row_elems, lane, cmask);
                    sample_accum<NV, FAST, FT>(acc, sa);
                }
            }
        }
        float* out_row = p.out + orow * p.C;
        int32_t* cnt_dst = p.count + orow;
        bool store_row = true;
        if (p.n_peers > 0) {  // push mode: store straight into the owner rank's staging slot (NVLink peer store)
            const int owner = (int)(i / p.rows_per_rank);
            const int64_t slot = (int64_t)p.src_rank * p.rows_per_rank + (i - (int64_t)owner * p.rows_per_rank);
            out_row = p.peer_out[owner] + slot * p.C;
            cnt_dst = p.peer_cnt[owner] + slot;
            store_row = cnt > 0;  // a point none of this rank's views sees sends its count (0) only: the reducer
                                  // skips the slot, and a contiguous view shard sees a fraction of the scene
        }
        const float denom = (float)max(cnt, 1);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int c = chan_of<FT>(k, lane);
            if (c < p.C && store_row) {
                const float4 o = p.finalize ? f4_div(acc[k], denom) : acc[k];
                st_cs_f4(out_row + c, o);
                sp_acc[k] = f4_add(sp_acc[k], o);
            }
        }
        if (lane == 0) *cnt_dst = cnt;
    }
    if (p.pool && seg < p.S) {  // uniform per CTA
#pragma unroll
        for (int k = 0; k < NV; ++k) s_red[warp][k * 32 + lane] = sp_acc[k];
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                float4 t = s_red[0][k * 32 + lane];
#pragma unroll
                for (int w = 1; w < kLiftWarps; ++w) t = f4_add(t, s_red[w][k * 32 + lane]);
                const int c = chan_of<FT>(k, lane);
                if (c < p.C) *reinterpret_cast<float4*>(p.partials + task * (int64_t)p.C + c) = t;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K2 (alternative, variant bit 1): view-synchronous tile gather.
// The L2->SM link is the bound of the plain gather (every sample pulls 4 rows = 4*C*sizeof bytes; measured
// ~10.5 TB/s = the L2 roof), while the points of a spatially compact tile share most of their tap rows
// when looked at through ONE view (2.8-4x fewer distinct rows per (tile, view), tools/analyze_reuse.py).
// So a CTA (8 warps) owns a tile of 32 consecutive points of the refined order, every warp keeps the
// accumulators of its 4 points in registers, and the whole CTA walks the views of the tile IN LOCKSTEP
// (one __syncthreads per view): all rows of (tile, view) are requested within a short window, the first
// request of a row goes to L2, the repeats hit L1. Per view, lanes 0..3 re-project their point and build
// the sample scalars; samples are issued/blended two at a time (register double buffer).
// Per point the views are still summed in ascending order with the same unfused arithmetic -> the result
// is bit-identical to gather_kernel and to the oracle.
// ---------------------------------------------------------------------------------------------------
constexpr int kTilePts = 32;    // points per tile step
constexpr int kMaxChunks = 32;  // <= 1024 views per launch

template <int NV, typename FT, bool FAST, bool DB, int kTileG>
__global__ void __launch_bounds__(32 * (kTilePts / kTileG), 512 / (32 * (kTilePts / kTileG)))
    gather_tile_kernel(const LiftParams p, const uint32_t* __restrict__ masks, int nchunks, int flags) {
    constexpr int kTileWarps = kTilePts / kTileG;
    const bool do_sync = (flags & 1) == 0, do_prefetch = (flags & 2) == 0;
    __shared__ uint32_t s_union[kMaxChunks];
    __shared__ float4 s_red[kTileWarps][NV * 32];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int64_t task = blockIdx.x;
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
    const int row_elems = p.Wf * p.C;
    const unsigned cmask = channel_mask<FT, NV>(p.C, lane);

    float4 sp_acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) sp_acc[k] = f4_zero();
    Sample<NV, FT> sa, sb;
    sample_clear<NV, FT>(sa);
    if (DB) sample_clear<NV, FT>(sb);

    for (int64_t t0 = start; t0 < end; t0 += kTilePts) {  // one pass when run == 32
        // lanes 0..3 of every warp own one point each
        const int64_t my_pos = t0 + warp * kTileG + lane;
        const bool own = lane < kTileG && my_pos < end;
        int32_t my_pid = -1;
        float mx = 0.f, my = 0.f, mz = 0.f;
        if (own) {
            my_pid = p.order ? p.order[my_pos] : (int32_t)my_pos;
            mx = __ldg(p.xyz + 3 * (int64_t)my_pid);
            my = __ldg(p.xyz + 3 * (int64_t)my_pid + 1);
            mz = __ldg(p.xyz + 3 * (int64_t)my_pid + 2);
        }
        const uint32_t* __restrict__ mw = masks + (int64_t)max(my_pid, 0) * nchunks;
        // union of the tile's view masks
        if (threadIdx.x < kMaxChunks) s_union[threadIdx.x] = 0u;
        __syncthreads();
        int my_cnt = 0;
        for (int c = 0; c < nchunks; ++c) {
            uint32_t m = own ? __ldg(mw + c) : 0u;
            my_cnt += __popc(m);
#pragma unroll
            for (int o = 1; o < kTileG; o <<= 1) m |= __shfl_xor_sync(kFull, m, o);
            if (lane == 0 && m) atomicOr(&s_union[c], m);
        }
        __syncthreads();

        float4 acc[kTileG][NV];
#pragma unroll
        for (int j = 0; j < kTileG; ++j) {
#pragma unroll
            for (int k = 0; k < NV; ++k) acc[j][k] = f4_zero();
        }
        if (p.accumulate) {
            if (own) my_cnt += p.count[my_pid];
#pragma unroll
            for (int j = 0; j < kTileG; ++j) {
                const int32_t pj = __shfl_sync(kFull, my_pid, j);
                if (pj >= 0) {
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        const int c = chan_of<FT>(k, lane);
                        if (c < p.C) acc[j][k] = *reinterpret_cast<const float4*>(p.out + (int64_t)pj * p.C + c);
                    }
                }
            }
        }

        // view iterator over the set bits of s_union (uniform over the CTA), one view of look-ahead:
        // while view i is blended, the rows of view i+1 are already being prefetched into L1.
        int it_c = 0;
        uint32_t it_um = nchunks > 0 ? s_union[0] : 0u;
        auto next_view = [&](int& chunk) -> int {
            while (it_um == 0u) {
                if (++it_c >= nchunks) return -1;
                it_um = s_union[it_c];
            }
            const int b = __ffs(it_um) - 1;
            it_um &= it_um - 1u;
            chunk = it_c;
            return it_c * 32 + b;
        };
        auto make_view = [&](int vrel, int chunk, SampleScalars& sc, unsigned& vm) {
            const bool vis = own && ((__ldg(mw + chunk) >> (vrel & 31)) & 1u);
            scalars_clear(sc);
            if (vis) {
                const int v = p.v_begin + vrel;
                const float4 k4 = ldg_f4(p.K4 + 4 * (int64_t)v);
                const float4 q0 = ldg_f4(p.w2c + 12 * (int64_t)v);
                const float4 q1 = ldg_f4(p.w2c + 12 * (int64_t)v + 4);
                const float4 q2 = ldg_f4(p.w2c + 12 * (int64_t)v + 8);
                float uu, ww;
                project_point(k4, q0, q1, q2, mx, my, mz, p.z_near, uu, ww);
                sc = make_scalars<FT>(fmap, view_elems, v, uu, ww, p.stride, p.inv_stride, p.Hf, p.Wf, p.C);
            }
            vm = __ballot_sync(kFull, vis) & ((1u << kTileG) - 1u);
            // L1 prefetch of the (up to) 4 rows of each visible sample: lane = (row, 128-byte line)
            const int prow = lane >> 3, pline = lane & 7;
            const int line_elems = 128 / (int)sizeof(FT);
            const bool line_ok = do_prefetch && pline * line_elems < p.C;
#pragma unroll
            for (int j = 0; j < kTileG; ++j) {
                if (!do_prefetch) break;
                if (vm & (1u << j)) {
                    const uint32_t alo = __shfl_sync(kFull, sc.addr_lo, j);
                    const uint32_t ahi = __shfl_sync(kFull, sc.addr_hi, j);
                    if (line_ok && ((alo >> prow) & 1u)) {
                        const FT* a = reinterpret_cast<const FT*>((uintptr_t)(((uint64_t)ahi << 32) | (uint64_t)(alo & ~0xFu))) +
                                      (prow & 1) * p.C + (prow >> 1) * row_elems + pline * line_elems;
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
                    }
                }
            }
        };

        SampleScalars cur, nxt;
        unsigned vm_cur = 0u, vm_nxt = 0u;
        int chunk_tmp = 0;
        int v_cur = next_view(chunk_tmp);
        if (v_cur >= 0) make_view(v_cur, chunk_tmp, cur, vm_cur);
        while (v_cur >= 0) {
            const int v_nxt = next_view(chunk_tmp);
            if (v_nxt >= 0) make_view(v_nxt, chunk_tmp, nxt, vm_nxt);
            if (DB && kTileG == 4) {
                // static schedule, two samples in flight: A<-0, B<-1, use A, A<-2, use B, B<-3, use A, use B
                if (vm_cur & 1u) sample_issue<NV, FT>(sa, cur, 0, p.C, row_elems, lane, cmask);
                if (vm_cur & 2u) sample_issue<NV, FT>(sb, cur, 1, p.C, row_elems, lane, cmask);
                if (vm_cur & 1u) sample_accum<NV, FAST, FT>(acc[0], sa);
                if (vm_cur & 4u) sample_issue<NV, FT>(sa, cur, 2, p.C, row_elems, lane, cmask);
                if (vm_cur & 2u) sample_accum<NV, FAST, FT>(acc[1], sb);
                if (vm_cur & 8u) sample_issue<NV, FT>(sb, cur, 3, p.C, row_elems, lane, cmask);
                if (vm_cur & 4u) sample_accum<NV, FAST, FT>(acc[2], sa);
                if (vm_cur & 8u) sample_accum<NV, FAST, FT>(acc[3], sb);
            } else {  // rows were prefetched into L1 one view ahead: a single register buffer suffices
#pragma unroll
                for (int j = 0; j < kTileG; ++j) {
                    if (vm_cur & (1u << j)) {
                        sample_issue<NV, FT>(sa, cur, j, p.C, row_elems, lane, cmask);
                        sample_accum<NV, FAST, FT>(acc[j], sa);
                    }
                }
            }
            if (do_sync) __syncthreads();  // keep the tile's warps on the same view: their rows meet in L1
            v_cur = v_nxt;
            cur = nxt;
            vm_cur = vm_nxt;
        }

#pragma unroll
        for (int j = 0; j < kTileG; ++j) {
            const int32_t pj = __shfl_sync(kFull, my_pid, j);
            const int cj = __shfl_sync(kFull, my_cnt, j);
            if (pj >= 0) {
                const float denom = (float)max(cj, 1);
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    const int c = chan_of<FT>(k, lane);
                    if (c < p.C) {
                        const float4 o = p.finalize ? f4_div(acc[j][k], denom) : acc[j][k];
                        st_cs_f4(p.out + (int64_t)pj * p.C + c, o);
                        sp_acc[k] = f4_add(sp_acc[k], o);
                    }
                }
                if (lane == 0) p.count[pj] = cj;
            }
        }
        __syncthreads();  // s_union is rewritten by the next tile step
    }
    if (p.pool && seg < p.S) {  // uniform per CTA
#pragma unroll
        for (int k = 0; k < NV; ++k) s_red[warp][k * 32 + lane] = sp_acc[k];
        __syncthreads();
        if (warp == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                float4 t = s_red[0][k * 32 + lane];
#pragma unroll
                for (int w = 1; w < kTileWarps; ++w) t = f4_add(t, s_red[w][k * 32 + lane]);
                const int c = chan_of<FT>(k, lane);
                if (c < p.C) *reinterpret_cast<float4*>(p.partials + task * (int64_t)p.C + c) = t;
            }
        }
    }
}

// out[s,:] = (sum of the run partials of s) / max(n_s,1). Eight lanes share one (superpoint, channel
// vector): lane j adds partials t0+j, t0+j+8, ... in order, then the eight lane sums are added in lane
// order -> a fixed summation tree (deterministic), 8x shorter dependent-load chains than a serial walk.
__global__ void __launch_bounds__(256) sp_combine_kernel(const float* __restrict__ partials,
                                                         const int32_t* __restrict__ task_offsets,
                                                         const int32_t* __restrict__ seg_offsets, int32_t S, int C,
                                                         int run, float* __restrict__ out) {
    const int vec_per_row = C >> 2;
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int sub = threadIdx.x & 7;
    const bool active = gid < (int64_t)S * vec_per_row;
    float4 acc = f4_zero();
    int n = 0, s = 0, c = 0;
    if (active) {
        s = (int)(gid / vec_per_row);
        c = (int)(gid % vec_per_row) * 4;
        n = seg_offsets[s + 1] - seg_offsets[s];
        const int t0 = task_offsets[s], t1 = t0 + (n + run - 1) / run;
        for (int t = t0 + sub; t < t1; t += 8)
            acc = f4_add(acc, *reinterpret_cast<const float4*>(partials + (int64_t)t * C + c));
    }
    // fixed-order combine of the 8 sub-lane sums: ((0+1)+(2+3)) + ((4+5)+(6+7))
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        float4 other;
        other.x = __shfl_xor_sync(kFull, acc.x, o);
        other.y = __shfl_xor_sync(kFull, acc.y, o);
        other.z = __shfl_xor_sync(kFull, acc.z, o);
        other.w = __shfl_xor_sync(kFull, acc.w, o);
        // both partners compute lo + hi in the same operand order -> identical bits
        acc = (sub & o) ? f4_add(other, acc) : f4_add(acc, other);
    }
    if (active && sub == 0) *reinterpret_cast<float4*>(out + (int64_t)s * C + c) = f4_div(acc, (float)max(n, 1));
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

template <int NV, typename FT>
static void launch_gather(const LiftParams& p, const uint32_t* masks, int nchunks, int64_t n_tasks, bool fast,
                          bool tile_kernel, bool double_buffer, int tile_flags, cudaStream_t stream) {
    const unsigned grid = (unsigned)n_tasks;
    constexpr bool kPrefetch = NV <= 4;
    if (!tile_kernel || NV > 2) {  // point-streaming gather (default; also the only path for wide rows, C > 256)
        // registers: 4 CTAs/SM (<=128 regs) is the default; the occupancy variants are for C<=256 only
        constexpr bool kVar = NV <= 2;
        const int occ = kVar ? (tile_flags >> 2) & 3 : 0;  // 0: 4 CTAs/SM, 1: 5 CTAs/SM (spills), 2: 3 CTAs/SM
        if (occ == 1) {
            if (fast) gather_kernel<NV, FT, true, kPrefetch, (kVar ? 5 : 1)><<<grid, kLiftThreads, 0, stream>>>(p, masks, nchunks);
            else gather_kernel<NV, FT, false, kPrefetch, (kVar ? 5 : 1)><<<grid, kLiftThreads, 0, stream>>>(p, masks, nchunks);
        } else if (occ == 2) {
            if (fast) gather_kernel<NV, FT, true, kPrefetch, (kVar ? 3 : 1)><<<grid, kLiftThreads, 0, stream>>>(p, masks, nchunks);
            else gather_kernel<NV, FT, false, kPrefetch, (kVar ? 3 : 1)><<<grid, kLiftThreads, 0, stream>>>(p, masks, nchunks);
        } else {
            if (fast) gather_kernel<NV, FT, true, kPrefetch, (kVar ? 4 : 1)><<<grid, kLiftThreads, 0, stream>>>(p, masks, nchunks);
            else gather_kernel<NV, FT, false, kPrefetch, (kVar ? 4 : 1)><<<grid, kLiftThreads, 0, stream>>>(p, masks, nchunks);
        }
    } else {
        constexpr int kNV = NV > 2 ? 2 : NV;
        const int tflags = tile_flags & 3;
        if (fast && double_buffer)
            gather_tile_kernel<kNV, FT, true, true, 4><<<grid, 256, 0, stream>>>(p, masks, nchunks, tflags);
        else if (fast)
            gather_tile_kernel<kNV, FT, true, false, 4><<<grid, 256, 0, stream>>>(p, masks, nchunks, tflags);
        else if (double_buffer)
            gather_tile_kernel<kNV, FT, false, true, 4><<<grid, 256, 0, stream>>>(p, masks, nchunks, tflags);
        else
            gather_tile_kernel<kNV, FT, false, false, 4><<<grid, 256, 0, stream>>>(p, masks, nchunks, tflags);
    }
}

template <typename FT>
static int dispatch_gather(const LiftParams& p, const uint32_t* masks, int nchunks, int64_t n_tasks, int variant,
                           cudaStream_t stream) {
    // float4 registers per tap per lane: C/128 for fp32 rows, 2*ceil(C/256) for 16-bit rows (8 channels per load)
    constexpr int kR = Tap<FT>::kRegs;
    const int nv = kR == 1 ? (p.C + 127) / 128 : 2 * ((p.C + 255) / 256);
    const bool fast = (variant & 1) != 0;
    const bool tk = (variant & 2) != 0 && nchunks <= kMaxChunks && !p.by_pos;
    const bool db = (variant & 4) != 0;
    const int tf = (variant >> 3) & 15;  // experiment bits: 8 = tile: no per-view barrier, 16 = tile: no L1 prefetch,
                                         // 32 / 64 = streaming kernel compiled for 5 / 3 CTAs per SM
    if (nv == 1) launch_gather<kR, FT>(p, masks, nchunks, n_tasks, fast, tk, db, tf, stream);  // (kR == 1 only)
    else if (nv == 2) launch_gather<2, FT>(p, masks, nchunks, n_tasks, fast, tk, db, tf, stream);
    else if (nv <= 4) launch_gather<4, FT>(p, masks, nchunks, n_tasks, fast, tk, db, tf, stream);
    else if (nv <= 8) launch_gather<8, FT>(p, masks, nchunks, n_tasks, fast, tk, db, tf, stream);
    else {
        set_error("sd3d_lift: C=%d > 1024 unsupported", p.C);
        return SD3D_ERR_UNSUPPORTED;
    }
    return SD3D_OK;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
// runs the staged gather may see: the plan's bound, or ceil(N / 32) consecutive runs when there is no run table
static int64_t staged_task_bound(int64_t N, int64_t max_tasks) { return imax64(max_tasks, ceil_div64(N, 32)); }
// 1/stride if stride is a (normal) power of two, else 0: then dividing and multiplying by the reciprocal agree bit for bit
static float pow2_reciprocal(float stride) {
    int e = 0;
    const float m = frexpf(stride, &e);
    return (m == 0.5f && e > -100 && e < 100) ? 1.0f / stride : 0.f;
}

}  // namespace sd3d

using namespace sd3d;

extern "C" size_t sd3d_lift_workspace_bytes(int64_t N, int n_views, int C, int64_t max_tasks) {
    if (N < 0 || n_views < 0 || C < 0 || max_tasks < 0) return 0;
    const size_t nchunks = (size_t)(n_views + 31) / 32;
    // run partials | visibility masks | visible-view counts | sample records (16 B per (point, view) slot) |
    // stage headers + run descriptors of the staged gather (lift_staged.cu)
    return align_up((size_t)max_tasks * C * sizeof(float), 256) + align_up((size_t)N * nchunks * sizeof(uint32_t), 256) +
           align_up((size_t)N * sizeof(int32_t), 256) + align_up((size_t)N * n_views * sizeof(int4), 256) + 256 +
           staged_workspace_bytes(staged_task_bound(N, max_tasks), n_views);
}

struct PushTargets {
    int n_ranks, src_rank;
    int64_t rows_per_rank;
    void* const* outs;  // [n_ranks] float* staging [n_ranks][rows_per_rank][C] of every rank (peer-mapped)
    void* const* cnts;  // [n_ranks] int32* staging [n_ranks][rows_per_rank]
};

static int lift_impl(const float* xyz, int64_t N, const float* K4, const float* w2c, int V, int view_begin,
                     int view_end, const void* depth, int depth_dtype, int Hd, int Wd, const void* fmap,
                     int fmap_dtype, int Hf, int Wf, int C, float stride, float tau, float z_near, int accumulate,
                     int finalize, const int32_t* order, float* out_feat, int32_t* count, int32_t* pix_idx,
                     uint8_t* vis, const int32_t* seg_offsets, int64_t S, const int32_t* task_offsets,
                     const int32_t* task_seg, int64_t max_tasks, int run, void* ws, size_t ws_bytes, int pool_,
                     int variant, void* stream_, const PushTargets* push) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (push != nullptr) {
        if (push->n_ranks < 1 || push->n_ranks > kMaxPeers || push->src_rank < 0 || push->src_rank >= push->n_ranks ||
            push->rows_per_rank <= 0 || push->rows_per_rank * push->n_ranks < N || push->outs == nullptr ||
            push->cnts == nullptr || accumulate || finalize || pool_ || (order == nullptr && (variant & 256) == 0) ||
            (variant & 2)) {
            set_error("sd3d_lift_push: needs 1..%d ranks, rows_per_rank * ranks >= N, a processing order, and "
                      "accumulate = finalize = pool = 0 (default gather kernel)", kMaxPeers);
            return SD3D_ERR_ARG;
        }
        for (int r = 0; r < push->n_ranks; ++r)
            if (push->outs[r] == nullptr || push->cnts[r] == nullptr || !aligned16(push->outs[r])) {
                set_error("sd3d_lift_push: staging pointer of rank %d is null or misaligned", r);
                return SD3D_ERR_ARG;
            }
    }
    if (N < 0 || V < 0 || view_begin < 0 || view_end > V || view_begin > view_end || Hd <= 0 || Wd <= 0 || Hf <= 0 ||
        Wf <= 0 || C <= 0 || !(stride > 0.f) || N >= (int64_t(1) << 31) - 64 || (int64_t)Hd * Wd >= (int64_t(1) << 31) ||
        ((int64_t)V + 1) * Hf * Wf >= (int64_t(1) << 31)) {  // sample records hold 32-bit pixel indices
        set_error("sd3d_lift: bad shape N=%lld V=%d views=[%d,%d) depth=%dx%d fmap=%dx%dx%d stride=%g", (long long)N,
                  V, view_begin, view_end, Hd, Wd, Hf, Wf, C, (double)stride);
        return SD3D_ERR_ARG;
    }
    if (C % 4 != 0 || C > 1024) {
        set_error("sd3d_lift: C=%d must be a multiple of 4 and <= 1024", C);
        return SD3D_ERR_UNSUPPORTED;
    }
    if (fmap_dtype != SD3D_F32 && C % 8 != 0) {
        set_error("sd3d_lift: C=%d must be a multiple of 8 for 16-bit feature maps", C);
        return SD3D_ERR_UNSUPPORTED;
    }
    if (depth_dtype != SD3D_F32 && depth_dtype != SD3D_U16) {
        set_error("sd3d_lift: depth dtype code %d unsupported", depth_dtype);
        return SD3D_ERR_UNSUPPORTED;
    }
    if (N > 0 && (xyz == nullptr || (push == nullptr && (out_feat == nullptr || count == nullptr)))) {
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
    // variant bits 8..13 select which of the three launches run: projection | stage planner | gather
    const bool do_project = (variant & (512 | 4096 | 8192)) == 0, do_gather = (variant & (256 | 4096)) == 0;
    const bool do_stage_plan = (variant & (256 | 8192)) == 0;
    if (pool) {
        if (!finalize || S < 0 || max_tasks < sd3d_sp_max_tasks(N, S, run) ||
            (do_gather && N > 0 && (order == nullptr || seg_offsets == nullptr || task_offsets == nullptr || task_seg == nullptr))) {
            set_error("sd3d_lift: fused pooling needs finalize=1, order, seg_offsets and the task tables");
            return SD3D_ERR_ARG;
        }
    } else {
        max_tasks = 0;
    }
    if (N == 0) return SD3D_OK;  // task_offsets are all zero -> sd3d_sp_combine writes zero rows
    const int n_views = view_end - view_begin;
    const int nchunks = (n_views + 31) / 32;
    if (ws == nullptr || !aligned16(ws) || ws_bytes < sd3d_lift_workspace_bytes(N, n_views, C, max_tasks)) {
        set_error("sd3d_lift: workspace missing/misaligned/too small (%zu < %zu bytes)", ws_bytes,
                  sd3d_lift_workspace_bytes(N, n_views, C, max_tasks));
        return SD3D_ERR_ARG;
    }
    uint32_t* masks = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(ws) +
                                                  align_up((size_t)max_tasks * C * sizeof(float), 256));
    LiftParams p;
    p.nvis = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(masks) + align_up((size_t)N * nchunks * sizeof(uint32_t), 256));
    p.recs = reinterpret_cast<int4*>(reinterpret_cast<uint8_t*>(p.nvis) + align_up((size_t)N * sizeof(int32_t), 256));
    p.n_views = n_views;
    p.k_views = (variant >> 16) & 0xff;
    if (p.k_views > kMaxNearest || (p.k_views > 0 && ((variant & 2) || accumulate))) {
        set_error("sd3d_lift: nearest-view sampling supports k <= %d, the default gather and accumulate = 0", kMaxNearest);
        return SD3D_ERR_UNSUPPORTED;
    }
    p.n_peers = 0;
    p.src_rank = 0;
    p.task_rot = 0;
    p.rows_per_rank = 1;
    for (int r = 0; r < kMaxPeers; ++r) {
        p.peer_out[r] = nullptr;
        p.peer_cnt[r] = nullptr;
    }
    if (push != nullptr) {
        p.n_peers = push->n_ranks;
        p.src_rank = push->src_rank;
        p.rows_per_rank = push->rows_per_rank;
        p.task_rot = (int)(((int64_t)((push->src_rank + 1) % push->n_ranks) * push->rows_per_rank) / run);
        for (int r = 0; r < push->n_ranks; ++r) {
            p.peer_out[r] = static_cast<float*>(push->outs[r]);
            p.peer_cnt[r] = static_cast<int32_t*>(push->cnts[r]);
        }
    }
    p.xyz = xyz; p.N = N; p.K4 = K4; p.w2c = w2c; p.v_begin = view_begin; p.v_end = view_end;
    p.depth = depth; p.depth_u16 = depth_dtype == SD3D_U16; p.Hd = Hd; p.Wd = Wd;
    p.fmap = fmap; p.Hf = Hf; p.Wf = Wf; p.C = C; p.stride = stride; p.inv_stride = pow2_reciprocal(stride); p.tau = tau; p.z_near = z_near;
    p.accumulate = accumulate; p.finalize = finalize; p.by_pos = ((variant & 1024) || push != nullptr) ? 1 : 0; p.order = order; p.out = out_feat; p.count = count;
    p.pix_idx = pix_idx; p.vis = vis; p.pool = pool ? 1 : 0; p.seg_offsets = seg_offsets;
    p.task_offsets = task_offsets; p.task_seg = task_seg; p.S = (int32_t)S; p.run = run;
    p.partials = reinterpret_cast<float*>(ws);
    const int64_t n_tasks = pool ? max_tasks : ceil_div64(N, run);
    StagedParams sp;
    staged_carve(reinterpret_cast<uint8_t*>(p.recs) + align_up((size_t)N * n_views * sizeof(int4), 256),
                 staged_task_bound(N, max_tasks), n_views, sp);
    sp.n_tasks = ceil_div64(N, run);
    sp.max_tasks = max_tasks;
    sp.task_rot = p.task_rot;
    sp.cap_pix = sp.ring_slots = 0;
    sp.masks = masks;
    sp.nchunks = nchunks;
    const bool staged = (variant & 32768) != 0 && (variant & 2) == 0 && p.k_views == 0 && staged_supported(p, fmap_dtype, n_views);
    if (staged) {
        // variant bit 15: tap rows staged in shared memory by the bulk-copy engine (lift_staged.cu); the projection
        // kernel of that path also plans the stages
        const int plan_mode = do_project ? 2 : (do_stage_plan ? 1 : 0);
        const int rc = dispatch_staged(p, sp, fmap_dtype, variant, plan_mode, masks, do_gather, stream);
        if (rc != SD3D_OK) {
            set_error("sd3d_lift: no staged gather for C=%d dtype=%d", C, fmap_dtype);
            return rc;
        }
        return check_launch("sd3d_lift(staged)");
    }
    if (do_project && nchunks > 0) {
        const unsigned pgrid = (unsigned)ceil_div64(N, kProjThreads);
        if (p.k_views > 0) project_kernel<true><<<pgrid, kProjThreads, 0, stream>>>(p, masks, nchunks, sp.done, sp.n_done);
        else project_kernel<false><<<pgrid, kProjThreads, 0, stream>>>(p, masks, nchunks, sp.done, sp.n_done);
    }
    int rc;
    if (!do_gather) return check_launch("sd3d_lift(project)");
    switch (fmap_dtype) {
        case SD3D_F32: rc = dispatch_gather<float>(p, masks, nchunks, n_tasks, variant, stream); break;
        case SD3D_F16: rc = dispatch_gather<__half>(p, masks, nchunks, n_tasks, variant, stream); break;
        case SD3D_BF16: rc = dispatch_gather<__nv_bfloat16>(p, masks, nchunks, n_tasks, variant, stream); break;
        default:
            set_error("sd3d_lift: fmap dtype code %d unsupported", fmap_dtype);
            return SD3D_ERR_UNSUPPORTED;
    }
    if (rc != SD3D_OK) return rc;
    return check_launch("sd3d_lift");
}

extern "C" int sd3d_sp_combine(const void* partials, const int32_t* task_offsets, const int32_t* seg_offsets,
                               int64_t S, int C, int run, float* sp_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (S < 0 || C <= 0 || C % 4 != 0 || S >= (int64_t(1) << 30) || run <= 0) {
        set_error("sd3d_sp_combine: bad shape S=%lld C=%d run=%d (C must be a multiple of 4)", (long long)S, C, run);
        return SD3D_ERR_ARG;
    }
    if (S == 0) return SD3D_OK;
    if (partials == nullptr || task_offsets == nullptr || seg_offsets == nullptr || sp_out == nullptr ||
        !aligned16(partials) || !aligned16(sp_out)) {
        set_error("sd3d_sp_combine: null or misaligned buffer");
        return SD3D_ERR_ARG;
    }
    const int64_t threads = S * (int64_t)(C / 4) * 8;
    sp_combine_kernel<<<(unsigned)ceil_div64(threads, 256), 256, 0, stream>>>(
        reinterpret_cast<const float*>(partials), task_offsets, seg_offsets, (int32_t)S, C, run, sp_out);
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

extern "C" int sd3d_lift(const float* xyz, int64_t N, const float* K4, const float* w2c, int V, int view_begin,
                         int view_end, const void* depth, int depth_dtype, int Hd, int Wd, const void* fmap,
                         int fmap_dtype, int Hf, int Wf, int C, float stride, float tau, float z_near, int accumulate,
                         int finalize, const int32_t* order, float* out_feat, int32_t* count, int32_t* pix_idx,
                         uint8_t* vis, const int32_t* seg_offsets, int64_t S, const int32_t* task_offsets,
                         const int32_t* task_seg, int64_t max_tasks, int run, void* ws, size_t ws_bytes, int pool_,
                         int variant, void* stream_) {
    return lift_impl(xyz, N, K4, w2c, V, view_begin, view_end, depth, depth_dtype, Hd, Wd, fmap, fmap_dtype, Hf, Wf, C,
                     stride, tau, z_near, accumulate, finalize, order, out_feat, count, pix_idx, vis, seg_offsets, S,
                     task_offsets, task_seg, max_tasks, run, ws, ws_bytes, pool_, variant, stream_, nullptr);
}

extern "C" int sd3d_lift_push(const float* xyz, int64_t N, const float* K4, const float* w2c, int V, int view_begin,
                              int view_end, const void* depth, int depth_dtype, int Hd, int Wd, const void* fmap,
                              int fmap_dtype, int Hf, int Wf, int C, float stride, float tau, float z_near,
                              const int32_t* order, int run, void* ws, size_t ws_bytes, int n_ranks, int src_rank,
                              int64_t rows_per_rank, void* const* peer_sum, void* const* peer_count, int variant,
                              void* stream_) {
    PushTargets t;
    t.n_ranks = n_ranks;
    t.src_rank = src_rank;
    t.rows_per_rank = rows_per_rank;
    t.outs = peer_sum;
    t.cnts = peer_count;
    return lift_impl(xyz, N, K4, w2c, V, view_begin, view_end, depth, depth_dtype, Hd, Wd, fmap, fmap_dtype, Hf, Wf, C,
                     stride, tau, z_near, 0, 0, order, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr,
                     0, run, ws, ws_bytes, 0, variant & ~2, stream_, &t);
}

namespace sd3d {
// rows of this rank after all ranks have pushed: feat[l] = (sum over source ranks, ascending) / max(total count, 1)
__global__ void __launch_bounds__(256) push_reduce_kernel(const float* __restrict__ stage_sum,
                                                          const int32_t* __restrict__ stage_cnt, int n_ranks,
                                                          int64_t rows_per_rank, int64_t rows, int C,
                                                          float* __restrict__ feat, int32_t* __restrict__ count) {
    const int vec_per_row = C >> 2;
    const int64_t total = rows * vec_per_row;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = e / vec_per_row;
        const int cv = (int)(e - l * vec_per_row);
        int32_t cnt = 0;
        float4 acc = f4_zero();
        for (int r = 0; r < n_ranks; ++r) {
            const int64_t slot = (int64_t)r * rows_per_rank + l;
            const int32_t cr = stage_cnt[slot];
            cnt += cr;
            if (cr > 0)  // ranks that saw nothing of this point did not send a row (the slot holds stale data)
                acc = f4_add(acc, *reinterpret_cast<const float4*>(stage_sum + slot * C + 4 * cv));
        }
        *reinterpret_cast<float4*>(feat + l * C + 4 * cv) = f4_div(acc, (float)max(cnt, 1));
        if (cv == 0) count[l] = cnt;
    }
}
}  // namespace sd3d

extern "C" int sd3d_push_reduce(const float* stage_sum, const int32_t* stage_count, int n_ranks, int64_t rows_per_rank,
                                int64_t rows, int C, float* feat, int32_t* count, void* stream_) {
    if (n_ranks < 1 || rows_per_rank < rows || rows < 0 || C <= 0 || C % 4 != 0) {
        set_error("sd3d_push_reduce: bad shape ranks=%d rows_per_rank=%lld rows=%lld C=%d", n_ranks,
                  (long long)rows_per_rank, (long long)rows, C);
        return SD3D_ERR_ARG;
    }
    if (rows == 0) return SD3D_OK;
    if (stage_sum == nullptr || stage_count == nullptr || feat == nullptr || count == nullptr || !aligned16(stage_sum) ||
        !aligned16(feat)) {
        set_error("sd3d_push_reduce: null or misaligned buffer");
        return SD3D_ERR_ARG;
    }
    const int64_t total = rows * (C / 4);
    const unsigned grid = (unsigned)imin64(ceil_div64(total, 256), (int64_t)num_sms() * 8);
    push_reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(stage_sum, stage_count, n_ranks, rows_per_rank, rows, C,
                                                                feat, count);
    return check_launch("sd3d_push_reduce");
}

// ---- peer-mapped staging memory (one process per GPU: CUDA IPC) ----
extern "C" int sd3d_peer_alloc(size_t bytes, void** ptr) {
    if (ptr == nullptr || bytes == 0) {
        set_error("sd3d_peer_alloc: bad argument");
        return SD3D_ERR_ARG;
    }
    const cudaError_t e = cudaMalloc(ptr, bytes);
    if (e != cudaSuccess) {
        set_error("sd3d_peer_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
        return SD3D_ERR_CUDA;
    }
    return SD3D_OK;
}
extern "C" int sd3d_peer_free(void* ptr) {
    if (ptr != nullptr && cudaFree(ptr) != cudaSuccess) {
        set_error("sd3d_peer_free: %s", cudaGetErrorString(cudaGetLastError()));
        return SD3D_ERR_CUDA;
    }
    return SD3D_OK;
}
extern "C" int sd3d_ipc_export(void* ptr, uint8_t* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaIpcMemHandle_t h;
    if (ptr == nullptr || handle64 == nullptr || cudaIpcGetMemHandle(&h, ptr) != cudaSuccess) {
        set_error("sd3d_ipc_export: %s", cudaGetErrorString(cudaGetLastError()));
        return SD3D_ERR_CUDA;
    }
    memcpy(handle64, &h, 64);
    return SD3D_OK;
}
extern "C" int sd3d_ipc_import(const uint8_t* handle64, void** ptr) {
    cudaIpcMemHandle_t h;
    if (handle64 == nullptr || ptr == nullptr) {
        set_error("sd3d_ipc_import: null argument");
        return SD3D_ERR_ARG;
    }
    memcpy(&h, handle64, 64);
    const cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        set_error("sd3d_ipc_import: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return SD3D_ERR_CUDA;
    }
    return SD3D_OK;
}
extern "C" int sd3d_ipc_close(void* ptr) {
    if (ptr != nullptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) {
        set_error("sd3d_ipc_close: %s", cudaGetErrorString(cudaGetLastError()));
        return SD3D_ERR_CUDA;
    }
    return SD3D_OK;
}
