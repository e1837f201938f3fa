// lift_staged.cuh -- interface between lift.cu (entry points, workspace) and lift_staged.cu (stage planner + gather).
#pragma once
#include "lift_common.cuh"

namespace sd3d {

constexpr int kStMaxSub = 2;  // box stages per (run, view); further samples of that view are fetched from global memory

struct StagedParams {
    uint32_t* hdrs;    // [tasks][cap_stages][16 words]: view, lane mask, box origin, pixel count, 256-bit bitmap
    int4* runinfo;     // [tasks]: first processing position, points, segment, stages
    int4* runpts;      // [tasks][32][2]: {point id, visible views, -, -}, the point's first sample record
    int32_t* counter;  // run dispenser of the persistent gather
    int32_t* chunk_cnt;  // [tasks][view chunks]: stages the planner warp of (run, chunk) produced
    int32_t* done;       // [tasks]: planner warps of the run that have finished (zeroed by the projection kernel)
    int64_t n_done;      // entries of `done`
    const uint32_t* masks;  // [N][nchunks] visible-view bit masks (projection kernel)
    int nchunks;
    int cap_stages;    // header slots per run = views * kStMaxSub
    int cap_pix;       // distinct pixels per stage (<= half the ring)
    int ring_slots;    // row buffers in a CTA's shared-memory ring
    int rowb;          // bytes of one feature-map pixel (C * element size)
    int64_t n_tasks;   // runs when there is no plan table (pool == 0): ceil(N / run)
    int64_t max_tasks; // bound of the run count with a plan table (pool != 0)
    int task_rot;      // push mode: rank-dependent rotation of the run order (balanced all-to-all)
};

bool staged_supported(const LiftParams& p, int fmap_dtype, int n_views);
size_t staged_workspace_bytes(int64_t tasks, int n_views);
void staged_carve(void* base, int64_t tasks, int n_views, StagedParams& sp);
// plan_mode 2: project_stage_kernel (projection + depth test + stage planning, needs `masks`), 1: stage_plan_kernel on the
// records of an earlier projection-only call, 0: stages are already planned; then gather_staged_kernel if do_gather.
// SD3D_ERR_UNSUPPORTED if the shape has no staged specialisation.
int dispatch_staged(const LiftParams& p, const StagedParams& sp, int fmap_dtype, int variant, int plan_mode,
                    uint32_t* masks, bool do_gather, cudaStream_t stream);

}  // namespace sd3d
