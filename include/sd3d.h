/* sd3d.h -- C ABI of libsd3d.so: the B200 (sm_100a) implementation of SegDINO3D's 2D->3D feature
 * lifting + superpoint pooling + mask-logit path.
 *
 * Boundary rules:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller (PyTorch) owns and allocates every buffer, including outputs and workspaces; the compute
 *     entries never allocate, free or retain a pointer past return. Two documented exceptions, neither on the
 *     per-scene path: sd3d_peer_alloc / sd3d_ipc_import hand out device memory the caller releases with
 *     sd3d_peer_free / sd3d_ipc_close, and the library keeps a small pool of non-blocking side streams per device
 *     (created on first use, never destroyed) on which sd3d_sp_plan and sd3d_lift_and_pool overlap independent
 *     kernels; work on them is forked from and joined back into `stream` with events before the call returns;
 *   - thread safety: entries are re-entrant; the only shared state is that stream pool (mutex) and per-device
 *     once-flags for kernel attributes (atomics); sd3d_last_error() is thread-local;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - every entry returns SD3D_OK or a negative code; sd3d_last_error() gives the thread-local text;
 *   - there is no CPU fallback: without a CUDA device every compute entry returns SD3D_ERR_CUDA.
 *
 * Each entry cites the reference interface it replaces (paths relative to /root/reference).
 */
#ifndef SD3D_H_
#define SD3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD3D_VERSION 100

#define SD3D_OK 0
#define SD3D_ERR_ARG (-1)         /* null pointer, negative size, misaligned buffer, workspace too small */
#define SD3D_ERR_UNSUPPORTED (-2) /* shape / dtype outside what the kernels are specialised for */
#define SD3D_ERR_CUDA (-3)        /* CUDA launch / runtime error (text in sd3d_last_error) */

/* element type codes */
#define SD3D_F32 0
#define SD3D_F16 1
#define SD3D_BF16 2
#define SD3D_U16 3 /* depth only: ScanNet uint16 millimetres, d = (float)raw * 0.001f */

/* sd3d_sp_mean / sd3d_lift modes */
#define SD3D_POOL_FAST 0  /* run-partials + ordered combine: deterministic, <=1e-5 rel. of the oracle */
#define SD3D_POOL_EXACT 1 /* ascending-point-index summation: bit-identical to aten CPU scatter_add_ */

int sd3d_version(void);
const char* sd3d_last_error(void);
/* number of SMs of the current device (grid sizing of the host mirror); <0 on error */
int sd3d_device_sms(void);

/* ---------------------------------------------------------------------------------------------
 * a-4 prerequisite: stable counting sort of points by superpoint id.
 * Replaces the implicit grouping inside torch_scatter.scatter_mean(src, index, dim=0)
 *   segdino3d/models/backbone/spconvunet.py:390,392 ; minkunet.py:639,641 (index = int64 ids).
 * idx[N] int64 ids in [0,S) (ids outside that range are parked after seg_offsets[S] and ignored).
 * perm[N] int32: point indices, superpoint by superpoint, ascending point index inside each.
 * seg_offsets[S+2] int32: superpoint s owns perm[seg_offsets[s] .. seg_offsets[s+1]); the last pair
 *   [seg_offsets[S], seg_offsets[S+1]=N) holds the points whose id was outside [0,S).
 * --------------------------------------------------------------------------------------------- */
size_t sd3d_sp_sort_workspace_bytes(int64_t N, int64_t S);
int sd3d_sp_sort(const int64_t* idx, int64_t N, int64_t S, int32_t* perm, int32_t* seg_offsets, void* ws,
                 size_t ws_bytes, void* stream);

/* Splits every superpoint into runs of <= `run` consecutive sorted points (the unit one warp reduces).
 * task_offsets[S+2]: superpoint s owns the ceil(n_s/run) tasks starting at task_offsets[s]; task_seg[t] = s
 * (segment S = the invalid-id points, which are lifted but never pooled); task_offsets[S+1] = task count.
 * max_tasks = sd3d_sp_max_tasks(N,S,run) is the size the caller must give task_seg. */
int64_t sd3d_sp_max_tasks(int64_t N, int64_t S, int run);
int sd3d_sp_tasks(const int32_t* seg_offsets, int64_t S, int run, int32_t* task_offsets, int32_t* task_seg,
                  int64_t max_tasks, void* stream);

/* The whole plan of the lifting path in one call: sd3d_sp_sort + spatial refinement + run table.
 * order[N] = perm with the points of every superpoint re-ordered along a Morton curve (world grid of
 * `cell` metres, e.g. 0.08), and the runs of task_seg laid out superpoint by superpoint along the world
 * Morton curve -- cache locality only, results of sd3d_lift never depend on it. xyz[N,3] f32 as given to
 * sd3d_lift. ws: sd3d_sp_sort_workspace_bytes(N,S). No reference counterpart (the reference gathers nothing). */
int sd3d_sp_plan(const int64_t* idx, const float* xyz, int64_t N, int64_t S, int run, float cell, int32_t* perm,
                 int32_t* order, int32_t* seg_offsets, int32_t* task_offsets, int32_t* task_seg, int64_t max_tasks,
                 void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a-4: out[s,:] = sum_{p in s} src[p,:] / max(|s|,1)    == scatter_mean(src, idx, dim=0)
 *   spconvunet.py:325,350,390,392 ; minkunet.py:639,641,653,674 ; torch-scatter 2.1.2 scatter_mean.
 * src[N,C] f32 row-major; perm/seg_offsets from sd3d_sp_sort; out[S,C] f32 fully overwritten.
 * point_count (nullable): fused lift finalize -- row p is divided by (float)max(point_count[p],1) first.
 * mode FAST needs task_offsets/task_seg (sd3d_sp_tasks) and ws >= max_tasks*C*4 bytes; EXACT needs none.
 * --------------------------------------------------------------------------------------------- */
int sd3d_sp_mean(const float* src, const int32_t* perm, const int32_t* seg_offsets, int64_t N, int64_t S, int C,
                 const int32_t* point_count, int mode, const int32_t* task_offsets, const int32_t* task_seg,
                 int64_t max_tasks, int run, void* ws, size_t ws_bytes, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a-1..a-3 (+ optional fused a-4): projection + depth-visibility test, bilinear gather of the
 * channels-last feature maps, sum over visible views, mean, superpoint pooling.
 * The reference has NO code for this (features are loaded precomputed,
 *   segdino3d/datasets/dataset/scannet200.py:219-226 ; scannet.py:177-184); the contract is the frozen
 *   spec of SURVEY.md Appendix A, whose output feeds extra_features["points_2dfeats"]
 *   (spconvunet.py:378 ; minkunet.py:612-618).
 *
 * xyz[N,3] f32; K4[V,4] f32 (fx,fy,cx,cy); w2c[V,3,4] f32; depth[V,Hd,Wd] (SD3D_F32 metres | SD3D_U16 mm);
 * fmap[V,Hf,Wf,C] channels-last (SD3D_F32 | SD3D_F16 | SD3D_BF16), C % 4 == 0 (C % 8 for 16-bit), C <= 1024.
 * view_begin/view_end: the ascending view range [view_begin, view_end) this call sums (a multi-GPU view
 *   shard); pix_idx/vis rows outside it are not touched.
 * accumulate != 0: start from the (sum,count) already in out_feat / count (continues the view loop of a
 *   previous call bit-exactly); finalize != 0: out_feat = sum / (float)max(count,1), else the raw sum.
 * order (nullable) int32[N]: processing order (perm from sd3d_sp_sort; spatially coherent order makes
 *   the gather cache-friendly). Results do not depend on it.
 * pix_idx[V,N] int32 (wi*Wd+ui or -1) and vis[V,N] u8: nullable parity outputs.
 * Fused pooling (pool != 0): needs order/seg_offsets/task_offsets/task_seg/run from the plan entries
 *   above and finalize != 0; the kernel leaves one partial row per run at the start of ws
 *   and sd3d_sp_combine(ws, ...) then yields sp_out[S,C] = scatter_mean(out_feat, idx).
 * ws: sd3d_lift_workspace_bytes(N, view_end-view_begin, C, pool ? max_tasks : 0) bytes: what the projection kernel
 *   hands to the gather kernel -- per-point view bit-masks, visible-view counts and one 16-byte sample record per
 *   visible (point, view) in [N][views] slots -- plus the run partials when pool != 0 (at offset 0). A
 *   projection-only call and the matching gather-only call must pass the same N, view range, C and max_tasks.
 * variant: bit 0 = contract the bilinear blend into FFMA (not bit-exact to Appendix A, <= 1e-6 rel.);
 *   bit 1 = view-synchronous tile gather kernel instead of the point-streaming one (same results; bits 2-4
 *   tune it: 4 = register double buffer, 8 = no per-view barrier, 16 = no L1 prefetch);
 *   bit 8 (256) = run the projection kernel only, bit 9 (512) = run the gather kernel only on the masks a
 *   previous projection-only call left in the same ws (per-kernel timing, stream overlap);
 *   bit 10 (1024) = rows of out_feat / count are indexed by processing position i (point order[i]) instead
 *   of by point id (what sd3d_lift_push does for its staging rows);
 *   bits 5/6 (32/64) = default gather compiled for 5 / 3 resident CTAs per SM (tuning points);
 *   bit 15 (32768) = shared-memory STAGED gather (needs order, run = 32, C * elemsize <= 2048 and a multiple of 16,
 *   else the direct gather runs): every distinct tap pixel of a (run, view) is copied once into shared memory by
 *   cp.async.bulk and blended from there; same bits as the direct gather. With it, the projection call (which then
 *   needs `order`) also plans the stages; after a projection WITHOUT order, bit 9 (512) runs a stand-alone stage
 *   planner before the gather; bit 12 (4096) = that planner only, bit 13 (8192) = gather only, stages planned.
 *   Bits 2 / 3 then mean: two samples in flight per consumer warp / 4 consumer warps x 8 points instead of 8 x 4.
 *   bits 16..23 = k (0..8): nearest-view sampling (the "Nearest View Sampling" box of the paper's overview figure,
 *   assets/overview.png; no code in the reference): of the views that see a point only the k with the smallest camera
 *   depth zc are summed and counted (ties to the lower view index); pix_idx / vis still report plain visibility.
 * --------------------------------------------------------------------------------------------- */
size_t sd3d_lift_workspace_bytes(int64_t N, int n_views, int C, int64_t max_tasks);
int sd3d_lift(const float* xyz, int64_t N, const float* K4, const float* w2c, int V, int view_begin, int view_end,
              const void* depth, int depth_dtype, int Hd, int Wd, const void* fmap, int fmap_dtype, int Hf, int Wf,
              int C, float stride, float tau, float z_near, int accumulate, int finalize, const int32_t* order,
              float* out_feat, int32_t* count, int32_t* pix_idx, uint8_t* vis, const int32_t* seg_offsets,
              int64_t S, const int32_t* task_offsets, const int32_t* task_seg, int64_t max_tasks, int run, void* ws,
              size_t ws_bytes, int pool, int variant, void* stream);

/* The whole path of one scene in ONE call: sd3d_sp_plan + sd3d_lift (projection on a library-owned side stream,
 * concurrent with the plan; fused pooling) + sd3d_sp_combine, all enqueued on `stream` from C++ into caller-owned
 * buffers. Replaces the per-scene python loop + scatter_mean of SpConvUNet.forward_wrapper
 * (segdino3d/models/backbone/spconvunet.py:365-395) fed by the lifted features
 * (segdino3d/datasets/dataset/scannet200.py:219-234). Outputs: out_feat[N,C] = mean over visible views, count[N],
 * sp_out[S,C] = scatter_mean(out_feat, sp_ids), and the plan (perm / order [N], seg_offsets / task_offsets [S+2],
 * task_seg [max_tasks = sd3d_sp_max_tasks(N,S,run)]). ws: sd3d_lift_and_pool_workspace_bytes(...) bytes, reusable
 * across scenes of the same (or smaller) size: the host side of a step is this one call, no allocation.
 * variant: as sd3d_lift (its stage-selection bits 8..13 are ignored). */
size_t sd3d_lift_and_pool_workspace_bytes(int64_t N, int64_t S, int n_views, int C, int run);
int sd3d_lift_and_pool(const float* xyz, int64_t N, const float* K4, const float* w2c, int V, const void* depth,
                       int depth_dtype, int Hd, int Wd, const void* fmap, int fmap_dtype, int Hf, int Wf, int C,
                       float stride, float tau, float z_near, const int64_t* sp_ids, int64_t S, int run, float cell,
                       int32_t* perm, int32_t* order, int32_t* seg_offsets, int32_t* task_offsets, int32_t* task_seg,
                       int64_t max_tasks, float* out_feat, int32_t* count, float* sp_out, void* ws, size_t ws_bytes,
                       int variant, void* stream);

/* second half of the fused pooling: sp_out[s,:] = (sum of the run partials of s, in run order) / max(|s|,1) */
int sd3d_sp_combine(const void* partials, const int32_t* task_offsets, const int32_t* seg_offsets, int64_t S, int C,
                    int run, float* sp_out, void* stream);

/* ---- view-sharded multi-GPU lifting with the exchange fused into the gather ("fuse the exchange with
 * the producing kernel"; the reference is single-GPU, evaluation/evaluate_3d.py:45). Rank `src_rank` lifts its views
 * [view_begin, view_end) for ALL points and stores the un-normalised row of processing position i (= order[i]) and
 * its visible count straight into the staging buffers of the rank that owns the position:
 *     owner = i / rows_per_rank,  slot = src_rank * rows_per_rank + (i % rows_per_rank)
 *     peer_sum[owner][slot, :] = sum over this rank's visible views,  peer_count[owner][slot] = their number
 * (the row is only sent when that number is > 0; sd3d_push_reduce ignores the rows of slots whose count is 0)
 * peer_sum[r] / peer_count[r] are device pointers valid on THIS device for rank r's staging buffers
 * ([n_ranks][rows_per_rank][C] f32 and [n_ranks][rows_per_rank] i32; peer-mapped with sd3d_ipc_import below, the
 * rank's own buffer directly). Stores to other ranks travel over NVLink while the kernel keeps gathering; no
 * collective moves feature rows. Same workspace as sd3d_lift. After a cross-rank barrier, sd3d_push_reduce on every
 * rank sums its slots in ascending rank order and divides by max(total count, 1). */
int sd3d_lift_push(const float* xyz, int64_t N, const float* K4, const float* w2c, int V, int view_begin, int view_end,
                   const void* depth, int depth_dtype, int Hd, int Wd, const void* fmap, int fmap_dtype, int Hf, int Wf,
                   int C, float stride, float tau, float z_near, const int32_t* order, int run, void* ws, size_t ws_bytes,
                   int n_ranks, int src_rank, int64_t rows_per_rank, void* const* peer_sum /*host array*/,
                   void* const* peer_count /*host array*/, int variant, void* stream);
int sd3d_push_reduce(const float* stage_sum, const int32_t* stage_count, int n_ranks, int64_t rows_per_rank,
                     int64_t rows /*<= rows_per_rank: rows this rank owns*/, int C, float* feat /*[rows,C]*/,
                     int32_t* count /*[rows]*/, void* stream);

/* staging memory that other processes can map (one process per GPU): plain cudaMalloc + CUDA IPC handles (64 bytes,
 * exchanged by the caller, e.g. torch.distributed.all_gather). The library never frees imported mappings itself. */
int sd3d_peer_alloc(size_t bytes, void** ptr);
int sd3d_peer_free(void* ptr);
int sd3d_ipc_export(void* ptr, uint8_t* handle64);
int sd3d_ipc_import(const uint8_t* handle64, void** ptr);
int sd3d_ipc_close(void* ptr);

/* feat = sum / (float)max(count,1) in place (Appendix A `feat_l`); used after the multi-GPU all-reduce */
int sd3d_lift_finalize(float* sum_inout, const int32_t* count, int64_t N, int C, void* stream);

/* out = (a + b + ...)/L over L scales: points_2dfeats = stack(list).mean(0), scannet200.py:233-234 */
int sd3d_scale_mean(const float* const* feats_host /*host array of L device ptrs*/, int L, int64_t numel,
                    float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * a-5: pred_mask[n,m] = sum_d q[n,d] * mf[m,d]     == torch.einsum('nd,md->nm', q, mf)
 *   segdino3d/models/decoder/instance_seg_3d_decoder.py:567 (ScanNetQueryDecoder._forward_head), :339.
 * q[n,d], mf[S,d] f32 row-major (K-major); out[n,S] f32 row-major.
 * precision: SD3D_F32  -> fp32 FFMA tiles (<=1e-5 rel.),
 *            SD3D_BF16 -> operands rounded to bf16, tcgen05.mma kind::f16 with fp32 TMEM accumulators (<=1e-2).
 * attn_mask (nullable) u8[n,S]: fused epilogue of :568-570 -- sigmoid(pred) < thr; rows that are
 *   all-true are reset to all-false by sd3d_attn_mask_fix (second tiny launch, done inside this call).
 * --------------------------------------------------------------------------------------------- */
int sd3d_mask_logits(const float* q, const float* mf, int n, int S, int d, int precision, float* out,
                     float thr, uint8_t* attn_mask, void* stream);

/* The same contraction for every scene of a batch in ONE launch (the per-scene python loop of _forward_head,
 * instance_seg_3d_decoder.py:557): problem i is q_host[i][n_host[i], d] x mf_host[i][S_host[i], d] -> out_host[i]
 * (+ attn_host[i] when attn_host != NULL). The *_host arrays are HOST arrays of `count` device pointers / sizes.
 * On the tensor-core path a CTA owns whole rows whenever S <= 512, and then the all-true-row reset of :570-571 is
 * done in the same kernel (no second pass over the mask). */
int sd3d_mask_logits_batched(const float* const* q_host, const float* const* mf_host, const int* n_host,
                             const int* S_host, int count, int d, int precision, float* const* out_host, float thr,
                             uint8_t* const* attn_host /*nullable*/, void* stream);

/* Operand producer of the mask head: y = LayerNorm(x) * weight + bias over the last dimension
 *   (self.out_norm(queries[i]), instance_seg_3d_decoder.py:558), written in ONE pass as fp32 (y_f32: what the cls /
 *   sem / score heads read, :561-566) and / or bf16 (y_bf16: the tensor-core operand). normalize == 0: no
 *   normalisation (weight / bias still applied when given) = the cast of the x_mask MLP output (:261-263, :645).
 *   x[n,d] f32 row-major, d <= 1024; weight / bias [d] nullable; y_f32 [n,d] / y_bf16 [n,d] nullable (not both). */
int sd3d_layernorm_cast(const float* x, const float* weight, const float* bias, int n, int d, float eps, int normalize,
                        float* y_f32, void* y_bf16, void* stream);

/* The same contraction on bf16 operands fed by TMA tensor loads (cp.async.bulk.tensor, 128-byte swizzle) into
 *   tcgen05.mma 128 x 128 x 16 tiles, persistent CTAs, warp-specialised (TMA / MMA / epilogue):
 *   q_bf16[n,d], mf_bf16[S,d] row-major bf16 (d % 64 == 0, d <= 256, 16-byte aligned), out[n,S] f32,
 *   attn_mask nullable as in sd3d_mask_logits (16-byte aligned); ws: sd3d_mask_logits_bf16_workspace_bytes(n) bytes of
 *   row flags + progress counters, only used when attn_mask != NULL. CONTRACT: ws must be ZERO-FILLED before the first
 *   call that uses it; every successful call hands it back zero-filled (the kernel that finishes a row's last tile
 *   resets the all-true rows of :570-571 in place and clears its flags), so a persistent ws needs no memset per call.
 *   One ws per concurrently running call. Logits within 1e-2 of the fp32 einsum (bf16 operands, fp32 accumulate). */
size_t sd3d_mask_logits_bf16_workspace_bytes(int n);
int sd3d_mask_logits_bf16(const void* q_bf16, const void* mf_bf16, int n, int S, int d, float* out, float thr,
                          uint8_t* attn_mask, void* ws, size_t ws_bytes, void* stream);

/* fp32-level accuracy on the same tensor-core kernel ("bf16x3"): every fp32 operand element is split into
 *   hi = bf16(x), mid = bf16(x - hi)  (sd3d_split_bf16: x[n,d] f32 -> y[n,2d] bf16 = (hi | mid), d % 4 == 0), and
 *   out = hi.hi + hi.mid + mid.hi accumulated in fp32. The dropped terms are <= 2^-16 relative per product; measured
 *   against the float64 product the result is within 2e-6 of |q||mf| (the cuBLAS fp32 einsum: 3e-7), i.e. inside the
 *   1e-5 tolerance of the path. Same shapes / workspace / attention-mask contract as sd3d_mask_logits_bf16, with
 *   q_split[n,2d], mf_split[S,2d]. Replaces the FFMA path of sd3d_mask_logits for large problems
 *   (instance_seg_3d_decoder.py:567 at eval scale: the reference runs this einsum in fp32, amp=False). */
int sd3d_split_bf16(const float* x, int n, int d, void* y_split, void* stream);

/* fp32 operands in, ONE host call: converts q / mf (plain bf16 cast for precision == SD3D_BF16, the (hi | mid) split for
 *   SD3D_F32 = fp32 tolerance) into `scratch` (sd3d_mask_logits_large_scratch_bytes(); contents irrelevant) and runs the
 *   TMA-fed kernel; `flags` is the zero-filled self-cleaning workspace of sd3d_mask_logits_bf16 (used with attn_mask
 *   only). The einsum + epilogue of instance_seg_3d_decoder.py:567-573 at eval scale; what ops.mask_logits calls for
 *   problems of >= 64 output tiles. */
size_t sd3d_mask_logits_large_scratch_bytes(int n, int S, int d, int precision);
int sd3d_mask_logits_large(const float* q, const float* mf, int n, int S, int d, int precision, float* out, float thr,
                           uint8_t* attn_mask, void* scratch, size_t scratch_bytes, void* flags, size_t flags_bytes,
                           void* stream);
int sd3d_mask_logits_bf16x3(const void* q_split, const void* mf_split, int n, int S, int d, float* out, float thr,
                            uint8_t* attn_mask, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * "next" rows of SURVEY 8(f)
 * --------------------------------------------------------------------------------------------- */

/* Superpoint -> point mask expansion with threshold and per-instance point count, one pass:
 *   mask_pred = mask_pred_sigmoid[:, superpoints] > thr ; mask_pointnum = mask_pred.sum(1)
 *   segdino3d/models/architecture/baseline3d.py:453-454,463.
 * mask_sig[K,S] f32 (already sigmoid / NMS-decayed, as at :453), superpoints[N] int64, out[K,N] u8 (0/1),
 * pointnum[K] int32 (fully overwritten). Ids outside [0,S) expand to 0. */
int sd3d_sp_expand_mask(const float* mask_sig, const int64_t* superpoints, int K, int64_t S, int64_t N, float thr,
                        uint8_t* out, int32_t* pointnum, void* stream);

/* Superpoint-level ground truth, one pass: out[s,k] = (mean over the points of superpoint s of [labels == k]) > 0.5
 *   == scatter_mean(F.one_hot(labels)[:, :K].float(), super_point_masks, dim=0) > 0.5
 *   segdino3d/datasets/dataset/scannet200.py:243-253 ; scannet.py:204-211 (instance and semantic variants).
 * labels[N] int64 (values outside [0,K) -- the background the reference drops -- vote for nobody), perm / seg_offsets
 * from sd3d_sp_sort of the superpoint ids, out[S,K] u8 fully overwritten. background_if_none != 0: a row without a
 * winner gets its LAST column set (scannet200.py:251). Integer counting: 2 * count > size, exact. */
int sd3d_sp_label_vote(const int64_t* labels, const int32_t* perm, const int32_t* seg_offsets, int64_t N, int64_t S,
                       int K, int background_if_none, uint8_t* out, void* stream);

/* Gradient of scatter_mean(src, idx, dim=0) w.r.t. src: grad_src[p,:] = grad_out[idx[p],:] / max(|idx[p]|,1)
 *   (the pooling runs under autograd in training: engine/train_engine_3d.py:99-105, spconvunet.py:390).
 * seg_offsets from sd3d_sp_sort (superpoint sizes); ids outside [0,S) get zero gradient. */
int sd3d_sp_mean_backward(const float* grad_out, const int64_t* idx, const int32_t* seg_offsets, int64_t N, int64_t S,
                          int C, float* grad_src, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SD3D_H_ */
