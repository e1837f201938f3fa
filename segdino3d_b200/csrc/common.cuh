// common.cuh -- shared helpers of libsd3d (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/sd3d.h"

namespace sd3d {

// thread-local error text behind sd3d_last_error()
void set_error(const char* fmt, ...);
// returns SD3D_OK or SD3D_ERR_CUDA (recording cudaGetLastError text prefixed by `what`)
int check_launch(const char* what);
int num_sms();
// a library-owned non-blocking side stream of the current device (small pool, round robin; never destroyed); nullptr on error
cudaStream_t plan_side_stream();

// Per-DEVICE once flag for function attributes (cudaFuncSetAttribute is per device; a process may drive several):
// returns true the first time it is called on the current device (always true for device ordinals >= 64).
inline bool first_on_device(std::atomic<uint64_t>* seen) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    const uint64_t bit = uint64_t(1) << dev;
    return (seen->fetch_or(bit) & bit) == 0;
}

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ __forceinline__ int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }

// 128-bit read-only loads
__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming (evict-first) 128-bit store: written once, not re-read by this kernel
__device__ __forceinline__ void st_cs_f4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// separately rounded a + b per component (never contracted)
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float4 f4_div(float4 a, float d) {
    return make_float4(__fdiv_rn(a.x, d), __fdiv_rn(a.y, d), __fdiv_rn(a.z, d), __fdiv_rn(a.w, d));
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace sd3d
