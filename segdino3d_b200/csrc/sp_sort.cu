// sp_sort.cu -- stable counting/radix sort of points by superpoint id, and the run ("task") table.
//
// Replaces the grouping that torch_scatter.scatter_mean performs with one global atomicAdd per
// element (reference call sites: segdino3d/models/backbone/spconvunet.py:390,392; minkunet.py:639,641).
// Sorting once turns the pooling into segmented reductions with no atomics on fp32 data, and gives the
// lifting kernel a spatially coherent processing order (superpoints are compact in space).
//
// Integer-only work, bit-exact by construction: perm is the unique stable permutation.
//
// Algorithm: LSD radix passes over ceil(log2(S+1)) key bits, <=10 bits per pass (one pass for S<1024).
//   pass = hist (per-block digit histogram, smem int atomics)
//        -> scan (single CTA exclusive scan of the bin-major [bins][blocks] matrix = global bases)
//        -> scatter (per-warp contiguous sub-chunks; __match_any_sync ranks keep equal keys in input order)
// Keys outside [0,S) are mapped to the extra key S ("trash"), which sorts last.
#include "common.cuh"

namespace sd3d {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kMaxDigitBits = 10;
constexpr int kMaxSortBlocks = 256;

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
    int64_t t = ceil_div64(N > 0 ? N : 1, kMaxSortBlocks);
    t = ceil_div64(t, kSortThreads) * kSortThreads;
    if (t < 1024) t = 1024;
    g.items_per_block = (int)t;
    g.nb = (int)ceil_div64(N > 0 ? N : 1, t);
    return g;
}

__device__ __forceinline__ int32_t clamp_key(int64_t id, int32_t S) { return (id < 0 || id >= S) ? S : (int32_t)id; }

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

// In-place exclusive scan of the bin-major matrix hist[bins][nb] (nb <= 256) by ONE CTA of 1024 threads:
// each warp scans whole rows (8 values per lane, all loads of a row in flight at once), the 32 warps'
// row totals are scanned through shared memory, then the row bases are added. Optionally emits
// seg_offsets[s] = scanned[s*nb] for s in [0,S] (valid when a single pass covers all key bits).
__global__ void __launch_bounds__(1024) radix_scan_kernel(int32_t* __restrict__ hist, int bins, int nb,
                                                          int32_t* __restrict__ seg_offsets, int32_t S, int64_t N) {
    __shared__ int32_t s_rowbase[1 << kMaxDigitBits];
    __shared__ int32_t s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kPer = kMaxSortBlocks / 32;  // 8 values per lane cover nb <= 256
    // pass 1: row-local exclusive scan in registers, row totals to smem
    for (int row = warp; row < bins; row += 32) {
        int32_t v[kPer];
        int32_t tsum = 0;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int col = lane * kPer + k;
            v[k] = col < nb ? hist[(int64_t)row * nb + col] : 0;
            tsum += v[k];
        }
        int32_t inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t n = __shfl_up_sync(kFull, inc, o);
            if (lane >= o) inc += n;
        }
        int32_t excl = inc - tsum;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int col = lane * kPer + k;
            if (col < nb) hist[(int64_t)row * nb + col] = excl;
            excl += v[k];
        }
        if (lane == 31) s_rowbase[row] = inc;  // row total
    }
    __syncthreads();
    // pass 2: exclusive scan of the row totals (bins <= 1024 = one value per thread)
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
    // pass 3: add the row bases
    for (int row = warp; row < bins; row += 32) {
        const int32_t base = s_rowbase[row];
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            const int col = lane * kPer + k;
            if (col < nb) hist[(int64_t)row * nb + col] += base;
        }
    }
    if (seg_offsets != nullptr) {
        for (int s = tid; s <= S; s += blockDim.x) seg_offsets[s] = s_rowbase[s];
        if (tid == 0) seg_offsets[S + 1] = (int32_t)N;  // end of the trash segment
    }
}

template <bool FIRST>
__global__ void __launch_bounds__(kSortThreads)
    radix_scatter_kernel(const int64_t* __restrict__ idx, const int32_t* __restrict__ keys_in,
                         const int32_t* __restrict__ vals_in, int64_t N, int32_t S, int shift, int bits,
                         int items_per_block, int nb, const int32_t* __restrict__ base, int32_t* __restrict__ keys_out,
                         int32_t* __restrict__ vals_out) {
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
            if (keys_out != nullptr) keys_out[dst] = key;
            vals_out[dst] = FIRST ? (int32_t)i : vals_in[i];
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

// Spatial refinement of the processing order: inside every superpoint the points are re-ordered along a
// Morton curve (chunks of <= kRefineChunk points, keys relative to the chunk's bounding box), so that the
// points one CTA lifts together project to neighbouring pixels. `order` is a permutation of `perm` inside
// each segment; results of the lifting do not depend on it (only cache behaviour does). Also emits one
// world-grid Morton key per superpoint (`anchor`), used to walk the superpoints in a spatially coherent order.
// Integer/compare-only work: deterministic.
constexpr int kRefineChunk = 2048;
constexpr int kRefineThreads = 256;

__global__ void __launch_bounds__(kRefineThreads)
    sp_refine_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ perm,
                     const int32_t* __restrict__ seg_offsets, int32_t* __restrict__ order,
                     uint32_t* __restrict__ anchor) {
    __shared__ uint64_t s_items[kRefineChunk];
    __shared__ float s_red[6][kRefineThreads / 32];
    __shared__ float s_box[6];
    const int seg = blockIdx.x;
    const int beg = seg_offsets[seg], end = seg_offsets[seg + 1];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    if (end <= beg) {
        if (threadIdx.x == 0) anchor[seg] = 0x7FFFFFFFu;
        return;
    }
    for (int c0 = beg; c0 < end; c0 += kRefineChunk) {
        const int n = min(kRefineChunk, end - c0);
        float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int32_t pid = perm[c0 + i];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float v = __ldg(xyz + 3 * (int64_t)pid + a);
                if (v == v) {  // NaNs do not move the box
                    lo[a] = fminf(lo[a], v);
                    hi[a] = fmaxf(hi[a], v);
                }
            }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                lo[a] = fminf(lo[a], __shfl_xor_sync(kFull, lo[a], o));
                hi[a] = fmaxf(hi[a], __shfl_xor_sync(kFull, hi[a], o));
            }
            if (lane == 0) {
                s_red[a][warp] = lo[a];
                s_red[3 + a][warp] = hi[a];
            }
        }
        __syncthreads();
        if (threadIdx.x < 6) {
            float v = s_red[threadIdx.x][0];
            for (int w = 1; w < kRefineThreads / 32; ++w)
                v = threadIdx.x < 3 ? fminf(v, s_red[threadIdx.x][w]) : fmaxf(v, s_red[threadIdx.x][w]);
            s_box[threadIdx.x] = v;
        }
        __syncthreads();
        const float bx = s_box[0], by = s_box[1], bz = s_box[2];
        const float ext = fmaxf(fmaxf(s_box[3] - bx, s_box[4] - by), fmaxf(s_box[5] - bz, 1e-6f));
        const float scale = 1023.0f / ext;
        if (c0 == beg && threadIdx.x == 0) {
            // world-grid key of the box centre: 0.25 m cells, grid origin at -128 m
            const float cx = 0.5f * (s_box[0] + s_box[3]), cy = 0.5f * (s_box[1] + s_box[4]),
                        cz = 0.5f * (s_box[2] + s_box[5]);
            const int gx = min(max((int)floorf(cx * 4.0f) + 512, 0), 1023);
            const int gy = min(max((int)floorf(cy * 4.0f) + 512, 0), 1023);
            const int gz = min(max((int)floorf(cz * 4.0f) + 512, 0), 1023);
            anchor[seg] = morton30(gx, gy, gz);
        }
        int n2 = 1;
        while (n2 < n) n2 <<= 1;
        for (int i = threadIdx.x; i < n2; i += blockDim.x) {
            uint64_t item = ~uint64_t(0);  // padding sorts last
            if (i < n) {
                const int32_t pid = perm[c0 + i];
                const float x = __ldg(xyz + 3 * (int64_t)pid), y = __ldg(xyz + 3 * (int64_t)pid + 1),
                            z = __ldg(xyz + 3 * (int64_t)pid + 2);
                const int qx = min(max((int)((x - bx) * scale), 0), 1023);
                const int qy = min(max((int)((y - by) * scale), 0), 1023);
                const int qz = min(max((int)((z - bz) * scale), 0), 1023);
                item = ((uint64_t)morton30(qx, qy, qz) << 32) | (uint32_t)pid;
            }
            s_items[i] = item;
        }
        __syncthreads();
        bitonic_sort_smem(s_items, n2);
        for (int i = threadIdx.x; i < n; i += blockDim.x) order[c0 + i] = (int32_t)(s_items[i] & 0xFFFFFFFFu);
        __syncthreads();
    }
}

// Task table. Segment s (n_s points) gets ceil(n_s/run) consecutive tasks starting at task_offsets[s];
// task_seg[t] = segment of task t; task_offsets[nseg] = total number of tasks (nseg counts the trash
// segment). With `anchor` the segments are laid out along the world Morton curve (spatially coherent
// CTAs run at the same time -> the feature-map regions they touch stay in L2), else in id order.
// ONE CTA of 1024 threads; ordering is skipped when nseg exceeds the in-smem sort capacity.
constexpr int kMaxOrderedSegs = 8192;

__global__ void __launch_bounds__(1024) sp_tasks_kernel(const int32_t* __restrict__ seg_offsets,
                                                        const uint32_t* __restrict__ anchor, int32_t nseg, int run,
                                                        int32_t* __restrict__ task_offsets,
                                                        int32_t* __restrict__ task_seg, int64_t max_tasks) {
    extern __shared__ uint64_t s_sorted[];  // [n2] (anchor << 32 | seg) when ordering is on
    __shared__ int32_t s_warp[32];
    __shared__ int32_t s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool ordered = anchor != nullptr && nseg <= kMaxOrderedSegs;
    if (ordered) {
        int n2 = 1;
        while (n2 < nseg) n2 <<= 1;
        for (int i = tid; i < n2; i += blockDim.x) {
            uint64_t item = ~uint64_t(0);
            // the trash segment (last) keeps the largest real key so that it stays at the end
            if (i < nseg) item = ((uint64_t)(i == nseg - 1 ? 0xFFFFFFFEu : anchor[i]) << 32) | (uint32_t)i;
            s_sorted[i] = item;
        }
        __syncthreads();
        bitonic_sort_smem(s_sorted, n2);
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

}  // namespace sd3d

using namespace sd3d;

extern "C" size_t sd3d_sp_sort_workspace_bytes(int64_t N, int64_t S) {
    (void)S;
    if (N < 0) N = 0;
    return (size_t)(1 << kMaxDigitBits) * kMaxSortBlocks * sizeof(int32_t) + 4 * (size_t)N * sizeof(int32_t) + 1024;
}

extern "C" int sd3d_sp_sort(const int64_t* idx, int64_t N, int64_t S, int32_t* perm, int32_t* seg_offsets, void* ws,
                            size_t ws_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N < 0 || S < 0 || N >= (int64_t(1) << 31) - 64 || S >= (int64_t(1) << 30)) {
        set_error("sd3d_sp_sort: N=%lld S=%lld out of range", (long long)N, (long long)S);
        return SD3D_ERR_ARG;
    }
    if (seg_offsets == nullptr || (N > 0 && (idx == nullptr || perm == nullptr)) || ws == nullptr ||
        ws_bytes < sd3d_sp_sort_workspace_bytes(N, S)) {
        set_error("sd3d_sp_sort: null buffer or workspace too small (%zu < %zu)", ws_bytes,
                  sd3d_sp_sort_workspace_bytes(N, S));
        return SD3D_ERR_ARG;
    }
    if (N == 0) {
        cudaMemsetAsync(seg_offsets, 0, (size_t)(S + 2) * sizeof(int32_t), stream);
        return check_launch("sd3d_sp_sort(memset)");
    }
    const SortGeom g = sort_geom(N, S);
    int32_t* hist = reinterpret_cast<int32_t*>(ws);
    int32_t* bufs = hist + (size_t)(1 << kMaxDigitBits) * kMaxSortBlocks;
    int32_t* keysA = bufs;
    int32_t* valsA = bufs + N;
    int32_t* keysB = bufs + 2 * N;
    int32_t* valsB = bufs + 3 * N;
    const int32_t* kin = nullptr;
    const int32_t* vin = nullptr;
    for (int p = 0; p < g.passes; ++p) {
        const int shift = p * g.bits_per_pass;
        const int bits = (p == g.passes - 1) ? (g.key_bits - shift) : g.bits_per_pass;
        const int bins = 1 << bits;
        const bool first = (p == 0), last = (p == g.passes - 1);
        int32_t* kout = last ? (g.passes > 1 ? ((p & 1) ? keysB : keysA) : nullptr) : ((p & 1) ? keysB : keysA);
        int32_t* vout = last ? perm : ((p & 1) ? valsB : valsA);
        const size_t sm_hist = (size_t)bins * sizeof(int32_t);
        const size_t sm_scat = (size_t)bins * kSortWarps * sizeof(int32_t);
        if (first)
            radix_hist_kernel<true><<<g.nb, kSortThreads, sm_hist, stream>>>(idx, nullptr, N, (int32_t)S, shift, bits,
                                                                            g.items_per_block, g.nb, hist);
        else
            radix_hist_kernel<false><<<g.nb, kSortThreads, sm_hist, stream>>>(nullptr, kin, N, (int32_t)S, shift,
                                                                             bits, g.items_per_block, g.nb, hist);
        radix_scan_kernel<<<1, 1024, 0, stream>>>(hist, bins, g.nb, (g.passes == 1) ? seg_offsets : nullptr,
                                                  (int32_t)S, N);
        if (first)
            radix_scatter_kernel<true><<<g.nb, kSortThreads, sm_scat, stream>>>(
                idx, nullptr, nullptr, N, (int32_t)S, shift, bits, g.items_per_block, g.nb, hist, kout, vout);
        else
            radix_scatter_kernel<false><<<g.nb, kSortThreads, sm_scat, stream>>>(
                nullptr, kin, vin, N, (int32_t)S, shift, bits, g.items_per_block, g.nb, hist, kout, vout);
        kin = kout;
        vin = vout;
    }
    if (g.passes > 1) {
        const int64_t threads = N + 1;
        seg_bounds_kernel<<<(unsigned)ceil_div64(threads, 256), 256, 0, stream>>>(kin, N, (int32_t)S, seg_offsets);
    }
    return check_launch("sd3d_sp_sort");
}

extern "C" int64_t sd3d_sp_max_tasks(int64_t N, int64_t S, int run) {
    if (run <= 0 || N < 0 || S < 0) return -1;
    return ceil_div64(N, run) + S + 1;
}

extern "C" int sd3d_sp_refine(const float* xyz, const int32_t* perm, const int32_t* seg_offsets, int64_t N, int64_t S,
                              int32_t* order, uint32_t* anchor, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N < 0 || S < 0 || S >= (int64_t(1) << 30) || seg_offsets == nullptr || anchor == nullptr ||
        (N > 0 && (xyz == nullptr || perm == nullptr || order == nullptr))) {
        set_error("sd3d_sp_refine: bad argument");
        return SD3D_ERR_ARG;
    }
    sp_refine_kernel<<<(unsigned)(S + 1), kRefineThreads, 0, stream>>>(xyz, perm, seg_offsets, order, anchor);
    return check_launch("sd3d_sp_refine");
}

extern "C" int sd3d_sp_tasks(const int32_t* seg_offsets, const uint32_t* anchor, int64_t S, int run,
                             int32_t* task_offsets, int32_t* task_seg, int64_t max_tasks, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (seg_offsets == nullptr || task_offsets == nullptr || (task_seg == nullptr && max_tasks > 0) || run <= 0 ||
        S < 0 || S >= (int64_t(1) << 30)) {
        set_error("sd3d_sp_tasks: bad argument");
        return SD3D_ERR_ARG;
    }
    // S+1 segments: the superpoints plus the trash segment [seg_offsets[S], seg_offsets[S+1])
    const int32_t nseg = (int32_t)S + 1;
    size_t smem = 0;
    if (anchor != nullptr && nseg <= kMaxOrderedSegs) {
        int n2 = 1;
        while (n2 < nseg) n2 <<= 1;
        smem = (size_t)n2 * sizeof(uint64_t);
        static bool attr_set = false;
        if (!attr_set) {
            const cudaError_t e = cudaFuncSetAttribute(sp_tasks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       kMaxOrderedSegs * (int)sizeof(uint64_t));
            if (e != cudaSuccess) {
                set_error("sd3d_sp_tasks: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                return SD3D_ERR_CUDA;
            }
            attr_set = true;
        }
    }
    sp_tasks_kernel<<<1, 1024, smem, stream>>>(seg_offsets, anchor, nseg, run, task_offsets, task_seg, max_tasks);
    return check_launch("sd3d_sp_tasks");
}
