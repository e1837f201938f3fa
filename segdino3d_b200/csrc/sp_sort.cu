// sp_sort.cu -- the superpoint "plan": stable sort of points by superpoint id, spatial refinement of the
// processing order, and the run ("task") table.
//
// Replaces the grouping that torch_scatter.scatter_mean performs with one global atomicAdd per element
// (reference call sites: segdino3d/models/backbone/spconvunet.py:390,392; minkunet.py:639,641). Sorting
// once turns the pooling into segmented reductions with no atomics on fp32 data, and gives the lifting
// kernels a spatially coherent processing order.
//
// Integer-only work, bit-exact by construction: perm is the unique stable permutation; `order` is a
// deterministic function of (ids, xyz).
//
// Pipeline (all kernels tiny and latency-bound at ScanNet sizes, so the count of launches is what matters):
//   radix pass = ONE cooperative kernel (radix_pass_fused_kernel): per-block digit histogram (smem int atomics)
//              -> one grid.sync -> EVERY block reads the L2-resident [blocks][bins] matrix and derives its own column
//                 prefix + the bin totals (no serial scan CTA, no second grid barrier)
//              -> stable scatter (__match_any_sync ranks keep equal keys in input order; the last pass also emits a
//                 27-bit Morton cell key per sorted point).
//                 Fallback when the grid cannot be co-resident: three kernels hist / scan (one CTA) / scatter.
//   one pass for S < 1024, ceil(log2(S+1)/10) passes otherwise (+ a boundary-search kernel).
//   refine = per-superpoint stable LSD counting sort by the 27-bit cell key (one CTA per superpoint, <= 3 x 512 bins)
//   tasks  = superpoints ranked along the world Morton curve, ceil(n_s/run) runs each (ONE CTA), launched on a
//            library-owned side stream (plan_side_stream) so that it overlaps the refinement.
// Keys outside [0,S) are mapped to the extra key S ("trash"), which sorts last.
#include <cooperative_groups.h>

#include <mutex>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace sd3d {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kMaxDigitBits = 10;
constexpr int kMaxSortBlocks = 256;
constexpr int kScanCap = 48 * 1024;  // ints of the [bins][blocks] matrix that fit the scan CTA's smem (192 KB)

struct SortGeom {
    int key_bits, passes, bits_per_pass, items_per_block, nb;
};

static SortGeom sort_geom(int64_t N, int64_t S) {
    SortGeom g;
    int kb = 1;
    while ((int64_t(1) << kb) < S + 1) ++kb;
    g.key_bits = kb;
    g.passes = (kb + kMaxDigitBits - 1) / kMaxDigitBits;
    g.bits_per_pass = (kb + g.passes - 1) / g.passes;
    const int max_bins = 1 << (g.passes == 1 ? kb : g.bits_per_pass);
    int nb_cap = kScanCap / max_bins;
    if (nb_cap > kMaxSortBlocks) nb_cap = kMaxSortBlocks;
    // the passes are latency-bound at ScanNet sizes: ~2048 items per block keeps the per-warp serial chain of
    // the stable scatter short (8 steps) while the [bins][blocks] matrix still fits one CTA's shared memory;
    // <= kFusedMaxBlocks blocks so that the pass can run as one cooperative kernel
    const int64_t by_size = N / 1024 > 8 ? N / 1024 : 8;
    if (nb_cap > by_size) nb_cap = (int)by_size;
    if (nb_cap > 128) nb_cap = 128;
    int64_t t = ceil_div64(N > 0 ? N : 1, nb_cap);
    t = ceil_div64(t, kSortThreads) * kSortThreads;
    if (t < 1024) t = 1024;
    g.items_per_block = (int)t;
    g.nb = (int)ceil_div64(N > 0 ? N : 1, t);
    return g;
}

__device__ __forceinline__ int32_t clamp_key(int64_t id, int32_t S) { return (id < 0 || id >= S) ? S : (int32_t)id; }

// 10-bit-per-axis Morton code
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t morton30(uint32_t x, uint32_t y, uint32_t z) {
    return spread10(x) | (spread10(y) << 1) | (spread10(z) << 2);
}
// position of a point inside its (512 cells)^3 block of the world grid as a 27-bit Morton code (9 bits per
// axis; 41 m blocks at the default 8 cm cell -- larger than any indoor scene). Points of one superpoint are
// ordered along the curve; no bounding boxes needed. Digit 0 (low 9 bits) orders points inside a (8 cells)^3
// sub-block, digits 1 and 2 order the sub-blocks; a superpoint only pays for the digits it actually spans.
__device__ __forceinline__ uint32_t cell_key27(float x, float y, float z, float inv_cell) {
    const int cx = (int)floorf(x * inv_cell) & 511, cy = (int)floorf(y * inv_cell) & 511,
              cz = (int)floorf(z * inv_cell) & 511;
    return morton30(cx, cy, cz) & 0x7FFFFFFu;
}

template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const int64_t* __restrict__ idx,
                                                                  const int32_t* __restrict__ keys_in, int64_t N,
                                                                  int32_t S, int shift, int bits, int items_per_block,
                                                                  int nb, int32_t* __restrict__ hist) {
    extern __shared__ int32_t s_hist[];
    const int bins = 1 << bits;
    for (int i = threadIdx.x; i < bins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int64_t beg = (int64_t)blockIdx.x * items_per_block;
    const int64_t end = min(beg + (int64_t)items_per_block, N);
    for (int64_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
        const int32_t key = FIRST ? clamp_key(idx[i], S) : keys_in[i];
        atomicAdd(&s_hist[(key >> shift) & (bins - 1)], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x) hist[(int64_t)i * nb + blockIdx.x] = s_hist[i];
}

// In-place exclusive scan of the bin-major matrix hist[bins][nb] (bins*nb <= kScanCap) by ONE CTA: the whole
// matrix is staged in shared memory with coalesced loads that are all in flight at once (the kernel is
// latency-bound), rows are scanned by warps, row totals by the CTA. Optionally emits seg_offsets[s] = global
// start of key s for s in [0,S] (valid when a single pass covers all key bits).
__global__ void __launch_bounds__(1024) radix_scan_kernel(int32_t* __restrict__ hist, int bins, int nb,
                                                          int32_t* __restrict__ seg_offsets, int32_t S, int64_t N) {
    extern __shared__ int32_t s_mat[];  // [bins*nb]
    __shared__ int32_t s_rowbase[1 << kMaxDigitBits];
    __shared__ int32_t s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int total = bins * nb;
    for (int i = tid; i < total; i += 1024) s_mat[i] = hist[i];
    __syncthreads();
    // rows: exclusive scan of each row in place, row total to s_rowbase
    for (int row = warp; row < bins; row += 32) {
        int32_t* r = s_mat + row * nb;
        int32_t carry = 0;
        for (int c0 = 0; c0 < nb; c0 += 32) {
            const int col = c0 + lane;
            const int32_t v = col < nb ? r[col] : 0;
            int32_t inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t n = __shfl_up_sync(kFull, inc, o);
                if (lane >= o) inc += n;
            }
            if (col < nb) r[col] = carry + inc - v;
            carry += __shfl_sync(kFull, inc, 31);
        }
        if (lane == 0) s_rowbase[row] = carry;
    }
    __syncthreads();
    // exclusive scan of the row totals (bins <= 1024 = one value per thread)
    {
        const int32_t t = tid < bins ? s_rowbase[tid] : 0;
        int32_t inc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t n = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t n = __shfl_up_sync(kFull, w, o);
                if (lane >= o) w += n;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        if (tid < bins) s_rowbase[tid] = inc - t + (warp > 0 ? s_warp[warp - 1] : 0);
    }
    __syncthreads();
    for (int i = tid; i < total; i += 1024) hist[i] = s_mat[i] + s_rowbase[i / nb];
    if (seg_offsets != nullptr) {
        for (int s = tid; s <= S; s += 1024) seg_offsets[s] = s_rowbase[s];
        if (tid == 0) seg_offsets[S + 1] = (int32_t)N;  // end of the trash segment
    }
}

template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads)
    radix_scatter_kernel(const int64_t* __restrict__ idx, const int32_t* __restrict__ keys_in,
                         const int32_t* __restrict__ vals_in, int64_t N, int32_t S, int shift, int bits,
                         int items_per_block, int nb, const int32_t* __restrict__ base, int32_t* __restrict__ keys_out,
                         int32_t* __restrict__ vals_out, const float* __restrict__ xyz, float inv_cell,
                         uint32_t* __restrict__ cell_out) {
    extern __shared__ int32_t s_cnt[];  // [kSortWarps][bins]
    const int bins = 1 << bits;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < bins * kSortWarps; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    const int64_t beg = (int64_t)blockIdx.x * items_per_block;
    const int64_t end = min(beg + (int64_t)items_per_block, N);
    const int per_warp = items_per_block / kSortWarps;  // multiple of 32
    const int64_t wbeg = min(beg + (int64_t)warp * per_warp, end);
    const int64_t wend = min(wbeg + (int64_t)per_warp, end);
    int32_t* my_cnt = s_cnt + warp * bins;
    // phase 1: per-warp digit counts over the warp's contiguous sub-chunk
    for (int64_t i = wbeg + lane; i < wend; i += 32) {
        const int32_t key = FIRST ? clamp_key(idx[i], S) : keys_in[i];
        atomicAdd(&my_cnt[(key >> shift) & (bins - 1)], 1);
    }
    __syncthreads();
    // phase 2: per bin, exclusive scan over warps on top of this block's global base
    for (int b = threadIdx.x; b < bins; b += blockDim.x) {
        int32_t run = base[(int64_t)b * nb + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const int32_t c = s_cnt[w * bins + b];
            s_cnt[w * bins + b] = run;
            run += c;
        }
    }
    __syncthreads();
    // phase 3: stable ranks, 32 items at a time in input order
    for (int64_t i0 = wbeg; i0 < wend; i0 += 32) {
        const int64_t i = i0 + lane;
        const bool active = i < wend;
        int32_t key = 0, digit = bins;  // inactive lanes share a sentinel digit
        if (active) {
            key = FIRST ? clamp_key(idx[i], S) : keys_in[i];
            digit = (key >> shift) & (bins - 1);
        }
        const unsigned peers = __match_any_sync(kFull, digit);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        int32_t dst = 0;
        if (active) dst = my_cnt[digit] + rank;
        __syncwarp();
        if (active && rank == 0) my_cnt[digit] += __popc(peers);
        __syncwarp();
        if (active) {
            const int32_t val = FIRST ? (int32_t)i : vals_in[i];
            if (keys_out != nullptr) keys_out[dst] = key;
            vals_out[dst] = val;
            if (cell_out != nullptr) {
                const float x = __ldg(xyz + 3 * (int64_t)val), y = __ldg(xyz + 3 * (int64_t)val + 1),
                            z = __ldg(xyz + 3 * (int64_t)val + 2);
                cell_out[dst] = cell_key27(x, y, z, inv_cell);
            }
        }
    }
}

// One radix pass as ONE cooperative kernel: hist -> grid.sync -> per-block column prefix + bin scan -> scatter.
// At ScanNet sizes every phase is a few microseconds of latency, so the two kernel boundaries (launch gap +
// tail + ramp) cost more than the work; the grid barrier replaces them. Same arithmetic as the three
// separate kernels (which remain the fallback when the grid cannot be co-resident).
constexpr int kFusedMaxBlocks = 128;

template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads)
    radix_pass_fused_kernel(const int64_t* __restrict__ idx, const int32_t* __restrict__ keys_in,
                            const int32_t* __restrict__ vals_in, int64_t N, int32_t S, int shift, int bits,
                            int items_per_block, int nb, int32_t* __restrict__ hist, int32_t* __restrict__ keys_out,
                            int32_t* __restrict__ vals_out, const float* __restrict__ xyz, float inv_cell,
                            uint32_t* __restrict__ cell_out, int32_t* __restrict__ seg_offsets) {
    extern __shared__ int32_t s_dyn[];
    __shared__ int32_t s_rowbase[1 << kMaxDigitBits], s_pre[1 << kMaxDigitBits];
    __shared__ int32_t s_warp[kSortWarps];
    cg::grid_group grid = cg::this_grid();
    const int bins = 1 << bits;
    const int lane = lane_id(), warp = threadIdx.x >> 5, tid = threadIdx.x;
    const int64_t beg = (int64_t)blockIdx.x * items_per_block;
    const int64_t end = min(beg + (int64_t)items_per_block, N);
    // ---- phase 1: per-block digit histogram
    for (int i = tid; i < bins; i += kSortThreads) s_dyn[i] = 0;
    __syncthreads();
    for (int64_t i = beg + tid; i < end; i += kSortThreads) {
        const int32_t key = FIRST ? clamp_key(idx[i], S) : keys_in[i];
        atomicAdd(&s_dyn[(key >> shift) & (bins - 1)], 1);
    }
    __syncthreads();
    for (int i = tid; i < bins; i += kSortThreads) hist[(int64_t)blockIdx.x * bins + i] = s_dyn[i];  // [nb][bins]
    grid.sync();
    // ---- phase 2: EVERY block derives what it needs from the [nb][bins] matrix (L2-resident, coalesced over bins):
    // per bin, the count in the blocks before it (its offset inside the bin) and the bin total; then a block-local
    // exclusive scan of the totals gives the bin bases. No serial scan CTA, no second grid-wide barrier.
    for (int b = tid; b < bins; b += kSortThreads) {
        int32_t pre = 0, tot = 0;
        for (int k = 0; k < nb; ++k) {
            const int32_t v = hist[(int64_t)k * bins + b];
            tot += v;
            if (k < (int)blockIdx.x) pre += v;
        }
        s_pre[b] = pre;
        s_rowbase[b] = tot;
    }
    __syncthreads();
    {  // exclusive scan of the bin totals: 4 consecutive bins per thread (bins <= 1024)
        int32_t t[4], tsum = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int b = tid * 4 + k;
            t[k] = b < bins ? s_rowbase[b] : 0;
            tsum += t[k];
        }
        int32_t inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t n = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        int32_t wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += s_warp[w];
        int32_t excl = wbase + inc - tsum;
        __syncthreads();  // everyone has read its totals before they are overwritten with the bases
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int b = tid * 4 + k;
            if (b < bins) s_rowbase[b] = excl;
            excl += t[k];
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && seg_offsets != nullptr) {
        for (int sg = tid; sg <= S; sg += kSortThreads) seg_offsets[sg] = s_rowbase[sg];
        if (tid == 0) seg_offsets[S + 1] = (int32_t)N;
    }
    // ---- phase 3: stable scatter (identical to radix_scatter_kernel)
    int32_t* s_cnt = s_dyn;  // [kSortWarps][bins]
    for (int i = tid; i < bins * kSortWarps; i += kSortThreads) s_cnt[i] = 0;
    __syncthreads();
    const int per_warp = items_per_block / kSortWarps;
    const int64_t wbeg = min(beg + (int64_t)warp * per_warp, end);
    const int64_t wend = min(wbeg + (int64_t)per_warp, end);
    int32_t* my_cnt = s_cnt + warp * bins;
    for (int64_t i = wbeg + lane; i < wend; i += 32) {
        const int32_t key = FIRST ? clamp_key(idx[i], S) : keys_in[i];
        atomicAdd(&my_cnt[(key >> shift) & (bins - 1)], 1);
    }
    __syncthreads();
    for (int b = tid; b < bins; b += kSortThreads) {
        int32_t run = s_rowbase[b] + s_pre[b];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const int32_t c = s_cnt[w * bins + b];
            s_cnt[w * bins + b] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int64_t i0 = wbeg; i0 < wend; i0 += 32) {
        const int64_t i = i0 + lane;
        const bool active = i < wend;
        int32_t key = 0, digit = bins;
        if (active) {
            key = FIRST ? clamp_key(idx[i], S) : keys_in[i];
            digit = (key >> shift) & (bins - 1);
        }
        const unsigned peers = __match_any_sync(kFull, digit);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        int32_t dst = 0;
        if (active) dst = my_cnt[digit] + rank;
        __syncwarp();
        if (active && rank == 0) my_cnt[digit] += __popc(peers);
        __syncwarp();
        if (active) {
            const int32_t val = FIRST ? (int32_t)i : vals_in[i];
            if (keys_out != nullptr) keys_out[dst] = key;
            vals_out[dst] = val;
            if (cell_out != nullptr) {
                const float x = __ldg(xyz + 3 * (int64_t)val), y = __ldg(xyz + 3 * (int64_t)val + 1),
                            z = __ldg(xyz + 3 * (int64_t)val + 2);
                cell_out[dst] = cell_key27(x, y, z, inv_cell);
            }
        }
    }
}

// seg_offsets[s] = first sorted position whose key >= s, for s in [0,S]  (multi-pass case)
__global__ void seg_bounds_kernel(const int32_t* __restrict__ sorted_keys, int64_t N, int32_t S,
                                  int32_t* __restrict__ seg_offsets) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > N) return;
    const int32_t kp = i > 0 ? sorted_keys[i - 1] : -1;
    const int32_t kc = i < N ? sorted_keys[i] : S;
    for (int32_t s = kp + 1; s <= kc && s <= S; ++s) seg_offsets[s] = (int32_t)i;
    if (i == N) seg_offsets[S + 1] = (int32_t)N;  // end of the trash segment
}

// Spatial refinement: inside every superpoint the points are re-ordered by their 27-bit Morton cell key with
// up to three stable counting-sort passes of 9 bits (LSD; one CTA per superpoint; same three phases as
// radix_scatter_kernel; passes over digits in which the superpoint's keys do not differ are skipped), so that
// the points one CTA lifts together project to neighbouring pixels. `order` is a permutation of `perm` inside
// each segment; lifting results do not depend on it (only cache behaviour does).
constexpr int kRefineBins = 512;

// one stable counting-sort pass over [beg,end) of a segment by ((key >> shift) & 511); whole CTA
// (no __restrict__: the second pass reads what the first one wrote -- keep these off the non-coherent load path)
__device__ __forceinline__ void seg_sort_pass(const int32_t* vals_in, const uint32_t* keys_in, int shift, int beg,
                                              int end, int32_t* vals_out, uint32_t* keys_out,
                                              int32_t (*s_cnt)[kRefineBins],
                                              int32_t* s_tot, int32_t* s_warp) {
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * kRefineBins; i += blockDim.x) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const int n = end - beg;
    const int per_warp = ((n + kSortWarps - 1) / kSortWarps + 31) / 32 * 32;
    const int wbeg = min(beg + warp * per_warp, end), wend = min(wbeg + per_warp, end);
    for (int i = wbeg + lane; i < wend; i += 32) atomicAdd(&s_cnt[warp][(keys_in[i] >> shift) & (kRefineBins - 1)], 1);
    __syncthreads();
    // bin totals -> exclusive scan over bins (2 bins per thread) -> running offsets per (warp, bin)
    {
        const int b0 = threadIdx.x * 2;
        int32_t t0 = 0, t1 = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            t0 += s_cnt[w][b0];
            t1 += s_cnt[w][b0 + 1];
        }
        int32_t inc = t0 + t1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t v = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        int32_t wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += s_warp[w];
        const int32_t excl = wbase + inc - (t0 + t1);
        s_tot[b0] = excl;
        s_tot[b0 + 1] = excl + t0;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kRefineBins; b += blockDim.x) {
        int32_t run = beg + s_tot[b];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const int32_t c = s_cnt[w][b];
            s_cnt[w][b] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int i0 = wbeg; i0 < wend; i0 += 32) {
        const int i = i0 + lane;
        const bool active = i < wend;
        const uint32_t key = active ? keys_in[i] : 0u;
        const int digit = active ? (int)((key >> shift) & (kRefineBins - 1)) : kRefineBins;
        const unsigned peers = __match_any_sync(kFull, digit);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        int32_t dst = 0;
        if (active) dst = s_cnt[warp][digit] + rank;
        __syncwarp();
        if (active && rank == 0) s_cnt[warp][digit] += __popc(peers);
        __syncwarp();
        if (active) {
            vals_out[dst] = vals_in[i];
            if (keys_out != nullptr) keys_out[dst] = key;
        }
    }
    __syncthreads();  // the CTA's global writes are visible to the CTA's next pass
}

__global__ void __launch_bounds__(kSortThreads)
    sp_refine_kernel(const int32_t* perm, const uint32_t* cell, const int32_t* __restrict__ seg_offsets,
                     int32_t* tmp_perm, uint32_t* tmp_key, int64_t N, int32_t* order) {
    __shared__ int32_t s_cnt[kSortWarps][kRefineBins];
    __shared__ int32_t s_tot[kRefineBins];
    __shared__ int32_t s_warp[kSortWarps];
    __shared__ uint32_t s_span;  // OR of (key ^ first key): which digits differ inside the superpoint
    const int seg = blockIdx.x;
    const int beg = seg_offsets[seg], end = seg_offsets[seg + 1];
    if (end <= beg) return;
    if (threadIdx.x == 0) s_span = 0u;
    __syncthreads();
    const uint32_t k0 = cell[beg];
    uint32_t diff = 0u;
    for (int i = beg + threadIdx.x; i < end; i += blockDim.x) diff |= cell[i] ^ k0;
    diff = __reduce_or_sync(kFull, diff);
    if (lane_id() == 0 && diff) atomicOr(&s_span, diff);
    __syncthreads();
    const uint32_t span = s_span;
    const int passes = (span >> 18) ? 3 : ((span >> 9) ? 2 : 1);
    int32_t* pa = tmp_perm;
    uint32_t* ka = tmp_key;
    int32_t* pb = tmp_perm + N;
    uint32_t* kb = tmp_key + N;
    if (passes == 1) {
        seg_sort_pass(perm, cell, 0, beg, end, order, nullptr, s_cnt, s_tot, s_warp);
    } else if (passes == 2) {
        seg_sort_pass(perm, cell, 0, beg, end, pa, ka, s_cnt, s_tot, s_warp);
        seg_sort_pass(pa, ka, 9, beg, end, order, nullptr, s_cnt, s_tot, s_warp);
    } else {
        seg_sort_pass(perm, cell, 0, beg, end, pa, ka, s_cnt, s_tot, s_warp);
        seg_sort_pass(pa, ka, 9, beg, end, pb, kb, s_cnt, s_tot, s_warp);
        seg_sort_pass(pb, kb, 18, beg, end, order, nullptr, s_cnt, s_tot, s_warp);
    }
}

// in-smem bitonic sort of n2 (power of two) 64-bit items (key << 32 | value) by the whole CTA
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* items, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n2; t += blockDim.x) {
                const int partner = t ^ j;
                if (partner > t) {
                    const uint64_t a = items[t], b = items[partner];
                    const bool up = (t & k) == 0;
                    if ((a > b) == up) {
                        items[t] = b;
                        items[partner] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Task table. Segment s (n_s points) gets ceil(n_s/run) consecutive tasks starting at task_offsets[s];
// task_seg[t] = segment of task t; task_offsets[nseg] = total number of tasks (nseg counts the trash
// segment). With `anchor` the segments are laid out along the world Morton curve (spatially coherent
// CTAs run at the same time -> the feature-map regions they touch stay in L1/L2), else in id order.
// ONE CTA of 1024 threads; nseg <= 1024: rank by counting (no barriers); <= 8192: bitonic; else id order.
constexpr int kMaxOrderedSegs = 8192;
constexpr int kMaxRankedSegs = 1024;

__device__ __forceinline__ uint32_t seg_anchor(const int32_t* __restrict__ seg_offsets, const int32_t* __restrict__ perm,
                                               const float* __restrict__ xyz, int seg) {
    // world-grid Morton key (0.25 m cells, origin -128 m) of the first point of the segment
    const int beg = seg_offsets[seg], end = seg_offsets[seg + 1];
    if (end <= beg) return 0x7FFFFFFFu;
    const int32_t p0 = perm[beg];
    const int gx = min(max((int)floorf(__ldg(xyz + 3 * (int64_t)p0) * 4.0f) + 512, 0), 1023);
    const int gy = min(max((int)floorf(__ldg(xyz + 3 * (int64_t)p0 + 1) * 4.0f) + 512, 0), 1023);
    const int gz = min(max((int)floorf(__ldg(xyz + 3 * (int64_t)p0 + 2) * 4.0f) + 512, 0), 1023);
    return morton30(gx, gy, gz);
}

__global__ void __launch_bounds__(1024) sp_tasks_kernel(const int32_t* __restrict__ seg_offsets,
                                                        const int32_t* __restrict__ perm,
                                                        const float* __restrict__ xyz, int32_t nseg, int run,
                                                        int32_t* __restrict__ task_offsets,
                                                        int32_t* __restrict__ task_seg, int64_t max_tasks) {
    extern __shared__ uint64_t s_sorted[];  // [n2] (anchor << 32 | seg) when ordering is on
    __shared__ int32_t s_warp[32];
    __shared__ int32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool ordered = xyz != nullptr && nseg <= kMaxOrderedSegs;
    if (ordered) {
        // the trash segment (last) gets the largest key so that it stays at the end
        if (nseg <= kMaxRankedSegs) {
            uint64_t mine = ~uint64_t(0);
            if (tid < nseg)
                mine = ((uint64_t)(tid == nseg - 1 ? 0xFFFFFFFEu : seg_anchor(seg_offsets, perm, xyz, tid)) << 32) | (uint32_t)tid;
            s_sorted[kMaxRankedSegs + tid] = mine;  // staging half
            __syncthreads();
            if (tid < nseg) {
                int rank = 0;
                for (int t = 0; t < nseg; ++t) rank += s_sorted[kMaxRankedSegs + t] < mine ? 1 : 0;  // keys are distinct
                s_sorted[rank] = mine;
            }
            __syncthreads();
        } else {
            int n2 = 1;
            while (n2 < nseg) n2 <<= 1;
            for (int i = tid; i < n2; i += blockDim.x) {
                uint64_t item = ~uint64_t(0);
                if (i < nseg)
                    item = ((uint64_t)(i == nseg - 1 ? 0xFFFFFFFEu : seg_anchor(seg_offsets, perm, xyz, i)) << 32) | (uint32_t)i;
                s_sorted[i] = item;
            }
            __syncthreads();
            bitonic_sort_smem(s_sorted, n2);
        }
    }
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nseg; base += 1024) {
        const int k = base + tid;
        int s = -1;
        int32_t nt = 0;
        if (k < nseg) {
            s = ordered ? (int)(s_sorted[k] & 0xFFFFFFFFu) : k;
            nt = (seg_offsets[s + 1] - seg_offsets[s] + run - 1) / run;
        }
        int32_t inc = nt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t n = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t n = __shfl_up_sync(kFull, w, o);
                if (lane >= o) w += n;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int32_t carry = s_carry;
        const int32_t excl = carry + (inc - nt) + (warp > 0 ? s_warp[warp - 1] : 0);
        if (s >= 0) {
            task_offsets[s] = excl;
            for (int32_t t = 0; t < nt; ++t)
                if ((int64_t)excl + t < max_tasks) task_seg[excl + t] = s;
        }
        __syncthreads();
        if (tid == 0) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (tid == 0) task_offsets[nseg] = s_carry;
}

static size_t align_up_sz(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct SortWs {
    int32_t* hist;
    int32_t* bufs;      // 4 * N ints (ping-pong keys / values of multi-pass sorts)
    uint32_t* cell;     // N   27-bit Morton cell keys of the sorted points
    int32_t* tmp_perm;  // 2N  ping-pong buffers of the multi-pass refinement
    uint32_t* tmp_key;  // 2N
    uint32_t* anchor;   // S + 1
};

static size_t sort_ws_layout(int64_t N, int64_t S, void* ws, SortWs* out) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = ws ? static_cast<uint8_t*>(ws) + off : nullptr;
        off += align_up_sz(bytes, 256);
        return p;
    };
    SortWs w;
    w.hist = static_cast<int32_t*>(take((size_t)kScanCap * sizeof(int32_t)));
    w.bufs = static_cast<int32_t*>(take(4 * (size_t)(N > 0 ? N : 0) * sizeof(int32_t)));
    w.cell = static_cast<uint32_t*>(take((size_t)(N > 0 ? N : 0) * sizeof(uint32_t)));
    w.tmp_perm = static_cast<int32_t*>(take(2 * (size_t)(N > 0 ? N : 0) * sizeof(int32_t)));
    w.tmp_key = static_cast<uint32_t*>(take(2 * (size_t)(N > 0 ? N : 0) * sizeof(uint32_t)));
    w.anchor = static_cast<uint32_t*>(take((size_t)(S + 1) * sizeof(uint32_t)));
    if (out) *out = w;
    return off + 256;
}

static int set_smem_attr_once(const void* fn, int bytes, std::atomic<uint64_t>* done, const char* what) {
    if (!first_on_device(done)) return SD3D_OK;  // the attribute is per device
    const cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
        done->store(0);
        set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
        return SD3D_ERR_CUDA;
    }
    return SD3D_OK;
}

// sort (+ optional cell keys); the shared body of sd3d_sp_sort and sd3d_sp_plan
static int run_sort(const int64_t* idx, const float* xyz, float inv_cell, int64_t N, int64_t S, int32_t* perm,
                    int32_t* seg_offsets, const SortWs& w, cudaStream_t stream) {
    static std::atomic<uint64_t> scan_attr{0};
    int rc = set_smem_attr_once(reinterpret_cast<const void*>(radix_scan_kernel), kScanCap * (int)sizeof(int32_t),
                                &scan_attr, "sd3d_sp_sort");
    if (rc != SD3D_OK) return rc;
    const SortGeom g = sort_geom(N, S);
    int32_t* keysA = w.bufs;
    int32_t* valsA = w.bufs + N;
    int32_t* keysB = w.bufs + 2 * N;
    int32_t* valsB = w.bufs + 3 * N;
    const int32_t* kin = nullptr;
    const int32_t* vin = nullptr;
    for (int p = 0; p < g.passes; ++p) {
        const int shift = p * g.bits_per_pass;
        const int bits = (p == g.passes - 1) ? (g.key_bits - shift) : g.bits_per_pass;
        const int bins = 1 << bits;
        const bool first = (p == 0), last = (p == g.passes - 1);
        int32_t* kout = last ? (g.passes > 1 ? ((p & 1) ? keysB : keysA) : nullptr) : ((p & 1) ? keysB : keysA);
        int32_t* vout = last ? perm : ((p & 1) ? valsB : valsA);
        uint32_t* cell_out = (last && xyz != nullptr) ? w.cell : nullptr;
        const size_t sm_hist = (size_t)bins * sizeof(int32_t);
        const size_t sm_scat = (size_t)bins * kSortWarps * sizeof(int32_t);
        const size_t sm_scan = (size_t)bins * g.nb * sizeof(int32_t);
        const size_t sm_fused = sm_scat;  // [kSortWarps][bins] counters (>= the [bins] histogram of phase 1)
        bool fused_done = false;
        if (g.nb <= kFusedMaxBlocks) {
            int32_t S32 = (int32_t)S;
            int shift_ = shift, bits_ = bits, ipb = g.items_per_block, nb_ = g.nb;
            int32_t* hist_ = w.hist;
            int32_t* seg_ = (g.passes == 1) ? seg_offsets : nullptr;
            const int64_t* idx_ = first ? idx : nullptr;
            const int32_t* kin_ = first ? nullptr : kin;
            const int32_t* vin_ = first ? nullptr : vin;
            void* args[] = {(void*)&idx_, (void*)&kin_, (void*)&vin_, (void*)&N, (void*)&S32, (void*)&shift_,
                            (void*)&bits_, (void*)&ipb, (void*)&nb_, (void*)&hist_, (void*)&kout, (void*)&vout,
                            (void*)&xyz, (void*)&inv_cell, (void*)&cell_out, (void*)&seg_};
            const void* fn = first ? reinterpret_cast<const void*>(radix_pass_fused_kernel<true>)
                                   : reinterpret_cast<const void*>(radix_pass_fused_kernel<false>);
            static std::atomic<uint64_t> attr_t{0}, attr_f{0};
            rc = set_smem_attr_once(fn, kScanCap * (int)sizeof(int32_t), first ? &attr_t : &attr_f, "sd3d_sp_sort");
            if (rc != SD3D_OK) return rc;
            const cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(g.nb), dim3(kSortThreads), args, sm_fused, stream);
            if (e == cudaSuccess) fused_done = true;
            else cudaGetLastError();  // not co-resident / unsupported: fall back to the three-kernel pass
        }
        if (!fused_done) {
            if (first)
                radix_hist_kernel<true><<<g.nb, kSortThreads, sm_hist, stream>>>(idx, nullptr, N, (int32_t)S, shift, bits,
                                                                                g.items_per_block, g.nb, w.hist);
            else
                radix_hist_kernel<false><<<g.nb, kSortThreads, sm_hist, stream>>>(nullptr, kin, N, (int32_t)S, shift,
                                                                                 bits, g.items_per_block, g.nb, w.hist);
            radix_scan_kernel<<<1, 1024, sm_scan, stream>>>(w.hist, bins, g.nb, (g.passes == 1) ? seg_offsets : nullptr,
                                                            (int32_t)S, N);
            if (first)
                radix_scatter_kernel<true><<<g.nb, kSortThreads, sm_scat, stream>>>(
                    idx, nullptr, nullptr, N, (int32_t)S, shift, bits, g.items_per_block, g.nb, w.hist, kout, vout, xyz,
                    inv_cell, cell_out);
            else
                radix_scatter_kernel<false><<<g.nb, kSortThreads, sm_scat, stream>>>(
                    nullptr, kin, vin, N, (int32_t)S, shift, bits, g.items_per_block, g.nb, w.hist, kout, vout, xyz,
                    inv_cell, cell_out);
        }
        kin = kout;
        vin = vout;
    }
    if (g.passes > 1) {
        const int64_t threads = N + 1;
        seg_bounds_kernel<<<(unsigned)ceil_div64(threads, 256), 256, 0, stream>>>(kin, N, (int32_t)S, seg_offsets);
    }
    return SD3D_OK;
}

static int run_tasks(const int32_t* seg_offsets, const int32_t* perm, const float* xyz, int64_t S, int run,
                     int32_t* task_offsets, int32_t* task_seg, int64_t max_tasks, cudaStream_t stream) {
    // S+1 segments: the superpoints plus the trash segment [seg_offsets[S], seg_offsets[S+1])
    const int32_t nseg = (int32_t)S + 1;
    size_t smem = 0;
    if (xyz != nullptr && nseg <= kMaxOrderedSegs) {
        int n2 = 1;
        while (n2 < nseg) n2 <<= 1;
        smem = (size_t)(nseg <= kMaxRankedSegs ? 2 * kMaxRankedSegs : n2) * sizeof(uint64_t);
        static std::atomic<uint64_t> attr{0};
        const int rc = set_smem_attr_once(reinterpret_cast<const void*>(sp_tasks_kernel),
                                          kMaxOrderedSegs * (int)sizeof(uint64_t), &attr, "sd3d_sp_tasks");
        if (rc != SD3D_OK) return rc;
    }
    sp_tasks_kernel<<<1, 1024, smem, stream>>>(seg_offsets, perm, xyz, nseg, run, task_offsets, task_seg, max_tasks);
    return SD3D_OK;
}

// Side streams owned by the library (per device, a few of them so that plans of different caller streams do not
// serialise behind each other); never destroyed.
cudaStream_t plan_side_stream() {
    constexpr int kMaxDev = 64, kPerDev = 4;
    static std::mutex mu;
    static cudaStream_t pool[kMaxDev][kPerDev] = {};
    static unsigned next[kMaxDev] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    const unsigned k = next[dev]++ % kPerDev;
    if (pool[dev][k] == nullptr && cudaStreamCreateWithFlags(&pool[dev][k], cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        pool[dev][k] = nullptr;
    }
    return pool[dev][k];
}

}  // namespace sd3d

using namespace sd3d;

extern "C" size_t sd3d_sp_sort_workspace_bytes(int64_t N, int64_t S) {
    if (N < 0) N = 0;
    if (S < 0) S = 0;
    return sort_ws_layout(N, S, nullptr, nullptr);
}

static int check_sort_args(const char* what, const int64_t* idx, int64_t N, int64_t S, const int32_t* perm,
                           const int32_t* seg_offsets, const void* ws, size_t ws_bytes) {
    if (N < 0 || S < 0 || N >= (int64_t(1) << 31) - 64 || S >= (int64_t(1) << 30)) {
        set_error("%s: N=%lld S=%lld out of range", what, (long long)N, (long long)S);
        return SD3D_ERR_ARG;
    }
    if (seg_offsets == nullptr || (N > 0 && (idx == nullptr || perm == nullptr)) || ws == nullptr ||
        !aligned16(ws) || ws_bytes < sd3d_sp_sort_workspace_bytes(N, S)) {
        set_error("%s: null buffer or workspace too small (%zu < %zu)", what, ws_bytes,
                  sd3d_sp_sort_workspace_bytes(N, S));
        return SD3D_ERR_ARG;
    }
    return SD3D_OK;
}

extern "C" int sd3d_sp_sort(const int64_t* idx, int64_t N, int64_t S, int32_t* perm, int32_t* seg_offsets, void* ws,
                            size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_sort_args("sd3d_sp_sort", idx, N, S, perm, seg_offsets, ws, ws_bytes);
    if (rc != SD3D_OK) return rc;
    if (N == 0) {
        cudaMemsetAsync(seg_offsets, 0, (size_t)(S + 2) * sizeof(int32_t), stream);
        return check_launch("sd3d_sp_sort(memset)");
    }
    SortWs w;
    sort_ws_layout(N, S, ws, &w);
    rc = run_sort(idx, nullptr, 0.f, N, S, perm, seg_offsets, w, stream);
    if (rc != SD3D_OK) return rc;
    return check_launch("sd3d_sp_sort");
}

extern "C" int64_t sd3d_sp_max_tasks(int64_t N, int64_t S, int run) {
    if (run <= 0 || N < 0 || S < 0) return -1;
    return ceil_div64(N, run) + S + 1;
}

extern "C" int sd3d_sp_tasks(const int32_t* seg_offsets, int64_t S, int run, int32_t* task_offsets, int32_t* task_seg,
                             int64_t max_tasks, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (seg_offsets == nullptr || task_offsets == nullptr || (task_seg == nullptr && max_tasks > 0) || run <= 0 ||
        S < 0 || S >= (int64_t(1) << 30)) {
        set_error("sd3d_sp_tasks: bad argument");
        return SD3D_ERR_ARG;
    }
    const int rc = run_tasks(seg_offsets, nullptr, nullptr, S, run, task_offsets, task_seg, max_tasks, stream);
    if (rc != SD3D_OK) return rc;
    return check_launch("sd3d_sp_tasks");
}

extern "C" int sd3d_sp_plan(const int64_t* idx, const float* xyz, int64_t N, int64_t S, int run, float cell,
                            int32_t* perm, int32_t* order, int32_t* seg_offsets, int32_t* task_offsets,
                            int32_t* task_seg, int64_t max_tasks, void* ws, size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = check_sort_args("sd3d_sp_plan", idx, N, S, perm, seg_offsets, ws, ws_bytes);
    if (rc != SD3D_OK) return rc;
    if (task_offsets == nullptr || (task_seg == nullptr && max_tasks > 0) || run <= 0 || !(cell > 0.f) ||
        max_tasks < sd3d_sp_max_tasks(N, S, run) || (N > 0 && (xyz == nullptr || order == nullptr))) {
        set_error("sd3d_sp_plan: bad argument (run=%d cell=%g max_tasks=%lld)", run, (double)cell, (long long)max_tasks);
        return SD3D_ERR_ARG;
    }
    SortWs w;
    sort_ws_layout(N, S, ws, &w);
    if (N == 0) {
        cudaMemsetAsync(seg_offsets, 0, (size_t)(S + 2) * sizeof(int32_t), stream);
        rc = run_tasks(seg_offsets, nullptr, nullptr, S, run, task_offsets, task_seg, max_tasks, stream);
        if (rc != SD3D_OK) return rc;
        return check_launch("sd3d_sp_plan(empty)");
    }
    rc = run_sort(idx, xyz, 1.0f / cell, N, S, perm, seg_offsets, w, stream);
    if (rc != SD3D_OK) return rc;
    // The task table (one latency-bound CTA) and the refinement (one CTA per superpoint) both depend only on the
    // sort: fork the task table onto a library-owned side stream and join before returning, so they overlap.
    cudaStream_t side = plan_side_stream();
    cudaEvent_t e_fork = nullptr, e_join = nullptr;
    if (side != nullptr && cudaEventCreateWithFlags(&e_fork, cudaEventDisableTiming) == cudaSuccess &&
        cudaEventCreateWithFlags(&e_join, cudaEventDisableTiming) == cudaSuccess) {
        cudaEventRecord(e_fork, stream);
        cudaStreamWaitEvent(side, e_fork, 0);
        rc = run_tasks(seg_offsets, perm, xyz, S, run, task_offsets, task_seg, max_tasks, side);
        cudaEventRecord(e_join, side);
        sp_refine_kernel<<<(unsigned)(S + 1), kSortThreads, 0, stream>>>(perm, w.cell, seg_offsets, w.tmp_perm, w.tmp_key, N, order);
        cudaStreamWaitEvent(stream, e_join, 0);
    } else {
        cudaGetLastError();
        rc = run_tasks(seg_offsets, perm, xyz, S, run, task_offsets, task_seg, max_tasks, stream);
        sp_refine_kernel<<<(unsigned)(S + 1), kSortThreads, 0, stream>>>(perm, w.cell, seg_offsets, w.tmp_perm, w.tmp_key, N, order);
    }
    if (e_fork) cudaEventDestroy(e_fork);  // released by the runtime once the recorded work has completed
    if (e_join) cudaEventDestroy(e_join);
    if (rc != SD3D_OK) return rc;
    return check_launch("sd3d_sp_plan");
}
