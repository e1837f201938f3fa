// lift_pool.cu -- the whole hot path of one scene behind ONE C-ABI call: plan (sort by superpoint + Morton refinement +
// run table), projection, gather + view mean + run partials, ordered combine.
//
// Call pattern it replaces: the per-scene python loop around scatter_mean in SpConvUNet.forward_wrapper
// (segdino3d/models/backbone/spconvunet.py:365-395) fed by the offline lifter's output slot
// (segdino3d/datasets/dataset/scannet200.py:219-234). Everything is enqueued from C++ into caller-owned buffers and a
// caller-owned workspace that can be kept across scenes: the host spends one ctypes call per scene (no allocation, no
// python-side stream fork/join), which is what bounds the multi-GPU replica throughput otherwise.
#include "common.cuh"

using namespace sd3d;

namespace {
size_t align256(size_t x) { return (x + 255) / 256 * 256; }
}  // namespace

extern "C" size_t sd3d_lift_and_pool_workspace_bytes(int64_t N, int64_t S, int n_views, int C, int run) {
    if (N < 0 || S < 0 || n_views < 0 || C < 0) return 0;
    if (run <= 0) run = 32;
    return align256(sd3d_sp_sort_workspace_bytes(N, S)) +
           sd3d_lift_workspace_bytes(N, n_views, C, sd3d_sp_max_tasks(N, S, run)) + 256;
}

extern "C" int sd3d_lift_and_pool(const float* xyz, int64_t N, const float* K4, const float* w2c, int V,
                                  const void* depth, int depth_dtype, int Hd, int Wd, const void* fmap, int fmap_dtype,
                                  int Hf, int Wf, int C, float stride, float tau, float z_near, const int64_t* sp_ids,
                                  int64_t S, int run, float cell, int32_t* perm, int32_t* order, int32_t* seg_offsets,
                                  int32_t* task_offsets, int32_t* task_seg, int64_t max_tasks, float* out_feat,
                                  int32_t* count, float* sp_out, void* ws, size_t ws_bytes, int variant, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (run <= 0) run = 32;
    if (N < 0 || S < 0 || V < 0 || ws == nullptr || !aligned16(ws) ||
        ws_bytes < sd3d_lift_and_pool_workspace_bytes(N, S, V, C, run) || max_tasks < sd3d_sp_max_tasks(N, S, run)) {
        set_error("sd3d_lift_and_pool: bad size, or workspace missing / too small (%zu < %zu bytes), or max_tasks %lld < %lld",
                  ws_bytes, sd3d_lift_and_pool_workspace_bytes(N, S, V, C, run), (long long)max_tasks,
                  (long long)sd3d_sp_max_tasks(N, S, run));
        return SD3D_ERR_ARG;
    }
    if (sp_out == nullptr && S > 0) {
        set_error("sd3d_lift_and_pool: null sp_out");
        return SD3D_ERR_ARG;
    }
    const size_t sort_bytes = sd3d_sp_sort_workspace_bytes(N, S);
    uint8_t* lift_ws = reinterpret_cast<uint8_t*>(ws) + align256(sort_bytes);
    const size_t lift_bytes = ws_bytes - align256(sort_bytes);
    variant &= ~(256 | 512 | 1024 | 4096 | 8192);  // the stage selection bits are this function's business
    auto lift = [&](int bits, const int32_t* ord, void* st) {
        return sd3d_lift(xyz, N, K4, w2c, V, 0, V, depth, depth_dtype, Hd, Wd, fmap, fmap_dtype, Hf, Wf, C, stride, tau,
                         z_near, 0, 1, ord, out_feat, count, nullptr, nullptr, seg_offsets, S, task_offsets, task_seg,
                         max_tasks, run, lift_ws, lift_bytes, 1, variant | bits, st);
    };
    auto plan = [&]() {
        return sd3d_sp_plan(sp_ids, xyz, N, S, run, cell, perm, order, seg_offsets, task_offsets, task_seg, max_tasks, ws,
                            sort_bytes, stream_);
    };
    int rc;
    cudaStream_t side = ((variant & 32768) || N == 0) ? nullptr : plan_side_stream();
    cudaEvent_t e_fork = nullptr, e_join = nullptr;
    if (side != nullptr && (cudaEventCreateWithFlags(&e_fork, cudaEventDisableTiming) != cudaSuccess ||
                            cudaEventCreateWithFlags(&e_join, cudaEventDisableTiming) != cudaSuccess)) {
        cudaGetLastError();
        if (e_fork) cudaEventDestroy(e_fork);
        e_fork = e_join = nullptr;
        side = nullptr;
    }
    if (side != nullptr) {
        // the projection needs no plan: it runs on a library-owned side stream concurrently with the (latency-bound)
        // plan kernels; the gather waits for both
        cudaEventRecord(e_fork, stream);
        cudaStreamWaitEvent(side, e_fork, 0);
        rc = lift(256, nullptr, side);
        cudaEventRecord(e_join, side);
        const int rc2 = plan();
        cudaStreamWaitEvent(stream, e_join, 0);
        cudaEventDestroy(e_fork);  // released by the runtime once the recorded work has completed
        cudaEventDestroy(e_join);
        if (rc != SD3D_OK) return rc;
        if (rc2 != SD3D_OK) return rc2;
        rc = lift(512, order, stream_);
    } else {
        // staged gather (variant bit 15): its projection kernel also cuts the stages and needs the plan first
        rc = plan();
        if (rc != SD3D_OK) return rc;
        rc = lift(0, order, stream_);
    }
    if (rc != SD3D_OK) return rc;
    return sd3d_sp_combine(lift_ws, task_offsets, seg_offsets, S, C, run, sp_out, stream_);
}
