// Shared device/host helpers for libmvg_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/mvg_b200.h"

namespace mvg {

// ---- error plumbing ---------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MVG_ELAUNCH;
  }
  return MVG_OK;
}

#define MVG_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::mvg::set_error(__VA_ARGS__);  \
      return MVG_EINVAL;              \
    }                                 \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- programmatic dependent launch (PDL) --------------------------------------------------------
// The kernels of a decoder step form one dependent chain of ~60 launches, most of them 3-40 us long.
// Launched with the programmatic-stream-serialization attribute, kernel i+1's CTAs are scheduled as the
// CTAs of kernel i drain (every kernel calls pdl_enter() before its first global access: wait for the
// whole previous grid + its memory flush, then allow the NEXT kernel's early launch), so launch latency
// and CTA start-up overlap the previous kernel's tail instead of following it.  MVG_PDL=0 disables the
// attribute (plain stream order; pdl_enter() is then a no-op pair of instructions).
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);     // errors surface in check_launch()
}

// ---- packed camera record: MVG_CAM_FLOATS = 64 fp32 per (frame, view) --------------------
// Filled by mvgformer_b200/cameras.py::pack_cameras (host-side mirror of
// unfold_camera_param_batch / get_affine_transform / get_proj_matricies_batch).
struct __align__(16) MvgCamera {
  float R[9];        // 0   world->camera rotation, row-major; x_cam = R (x - T)
  float T[3];        // 9   camera centre, world mm
  float f[2];        // 12  fx, fy
  float c[2];        // 14  cx, cy
  float k[3];        // 16  k1, k2, k3 (radial)
  float p[2];        // 19  p1, p2 (tangential)
  float aff[6];      // 21  original px -> network px (2x3 row-major)
  float inv_aff[6];  // 27  network px -> original px (2x3)
  float P[12];       // 33  K [R | -R T] (3x4 row-major)
  float Kinv[9];     // 45  inverse calibration matrix
  float wh[2];       // 54  original image size = 2 * center
  float clamp_max;   // 56  max(wh) over the whole batch tensor (dq_decoder.py:383)
  float pad[7];
};
static_assert(sizeof(MvgCamera) == MVG_CAM_FLOATS * sizeof(float), "camera record size");

// ---- bf16 helpers -------------------------------------------------------------------------
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  f[0] = bf16lo(q.x); f[1] = bf16hi(q.x); f[2] = bf16lo(q.y); f[3] = bf16hi(q.y);
  f[4] = bf16lo(q.z); f[5] = bf16hi(q.z); f[6] = bf16lo(q.w); f[7] = bf16hi(q.w);
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}

// fp32 ops without FMA contraction: the geometry / index path mirrors the op-by-op
// rounding of the reference's eager PyTorch code.
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

}  // namespace mvg
