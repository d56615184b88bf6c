// mvg_pack_cameras: raw `meta` camera tensors -> packed (B, V, 64) fp32 records, ONE launch.
//
// Replaces the ~350 eager tensor ops (and the address-keyed cache) the host mirror needed.
// What it restates, per (frame b, view v):
//   * unfold_camera_param_batch            lib/utils/cameras.py:118-133   float32 casts of R, T, f, c, k, p
//   * get_affine_transform(center, scale, 0, img_size)   lib/utils/transforms.py:72-112, which the
//     reference evaluates on the HOST with numpy + cv2.getAffineTransform per (view, frame, layer)
//     (lib/models/dq_decoder.py:361-372): the three float32 point pairs are rebuilt here and the
//     3-point affine is solved exactly in float64 (adjugate), then rounded to float32
//   * meta['inv_affine_trans'][:, :2, :]   lib/models/dq_decoder.py:414-418
//   * get_calib_matrix / K.inverse() / get_proj_matricies_batch(inv_trans=True)
//                                          lib/models/dq_decoder.py:207-246, :171
//   * wh = 2 * center, clamp_max = max over the view's whole (B,2) tensor  lib/models/dq_decoder.py:383
// Record layout = struct MvgCamera (common.cuh).
#include "common.cuh"

namespace mvg {

struct CamViewPtrs {
  const void* R; const void* T; const void* fx; const void* fy; const void* cx; const void* cy;
  const void* k; const void* p; const void* center; const void* scale; const void* inv_aff;
};
struct CamPackArgs {
  CamViewPtrs v[MVG_MAX_VIEWS];
  uint32_t f64_mask[MVG_MAX_VIEWS];   // bit i set: field i of the view is float64 (else float32)
};

__device__ __forceinline__ double ld_any(const void* p, int i, bool f64) {
  return f64 ? static_cast<const double*>(p)[i] : static_cast<double>(static_cast<const float*>(p)[i]);
}
__device__ __forceinline__ double f32r(double x) { return static_cast<double>(static_cast<float>(x)); }

__global__ void pack_cameras_kernel(const CamPackArgs args, int batch, int views, float out_w, float out_h,
                                    float* __restrict__ out) {
  const int v = blockIdx.x;
  const CamViewPtrs& q = args.v[v];
  const uint32_t m = args.f64_mask[v];
  __shared__ float s_max[32];
  float local_max = -INFINITY;
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    MvgCamera c;
    float R[9], T[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = c.R[i] = static_cast<float>(ld_any(q.R, b * 9 + i, m & 1u));
#pragma unroll
    for (int i = 0; i < 3; ++i) T[i] = c.T[i] = static_cast<float>(ld_any(q.T, b * 3 + i, m & 2u));
    const float fx = static_cast<float>(ld_any(q.fx, b, m & 4u)), fy = static_cast<float>(ld_any(q.fy, b, m & 8u));
    const float cx = static_cast<float>(ld_any(q.cx, b, m & 16u)), cy = static_cast<float>(ld_any(q.cy, b, m & 32u));
    c.f[0] = fx; c.f[1] = fy; c.c[0] = cx; c.c[1] = cy;
#pragma unroll
    for (int i = 0; i < 3; ++i) c.k[i] = static_cast<float>(ld_any(q.k, b * 3 + i, m & 64u));
#pragma unroll
    for (int i = 0; i < 2; ++i) c.p[i] = static_cast<float>(ld_any(q.p, b * 2 + i, m & 128u));
    // ---- affine original px -> network px (rot = 0), transforms.py:84-110
    const double cen0 = ld_any(q.center, b * 2, m & 256u), cen1 = ld_any(q.center, b * 2 + 1, m & 256u);
    const double src_w = ld_any(q.scale, b * 2, m & 512u) * 200.0, src_h = ld_any(q.scale, b * 2 + 1, m & 512u) * 200.0;
    const double dst_w = static_cast<double>(out_w), dst_h = static_cast<double>(out_h);
    const bool wide = src_w >= src_h;
    const double sdx = wide ? 0.0 : src_h * -0.5, sdy = wide ? src_w * -0.5 : 0.0;
    const double ddx = wide ? 0.0 : f32r(dst_h * -0.5), ddy = wide ? f32r(dst_w * -0.5) : 0.0;
    double sx[3], sy[3], dx[3], dy[3];
    sx[0] = f32r(cen0); sy[0] = f32r(cen1);
    sx[1] = f32r(cen0 + sdx); sy[1] = f32r(cen1 + sdy);
    dx[0] = f32r(dst_w * 0.5); dy[0] = f32r(dst_h * 0.5);
    dx[1] = f32r(dst_w * 0.5 + ddx); dy[1] = f32r(dst_h * 0.5 + ddy);
    // get_3rd_point(a, b) = b + (-(a-b).y, (a-b).x), stored as float32
    sx[2] = f32r(sx[1] + f32r(-(sy[0] - sy[1]))); sy[2] = f32r(sy[1] + f32r(sx[0] - sx[1]));
    dx[2] = f32r(dx[1] + f32r(-(dy[0] - dy[1]))); dy[2] = f32r(dy[1] + f32r(dx[0] - dx[1]));
    // inverse of [[x0,y0,1],[x1,y1,1],[x2,y2,1]] = adj / det (what cv2.getAffineTransform solves)
    const double det = sx[0] * (sy[1] - sy[2]) - sy[0] * (sx[1] - sx[2]) + (sx[1] * sy[2] - sx[2] * sy[1]);
    const double inv[3][3] = {
        {(sy[1] - sy[2]) / det, (sy[2] - sy[0]) / det, (sy[0] - sy[1]) / det},
        {(sx[2] - sx[1]) / det, (sx[0] - sx[2]) / det, (sx[1] - sx[0]) / det},
        {(sx[1] * sy[2] - sx[2] * sy[1]) / det, (sx[2] * sy[0] - sx[0] * sy[2]) / det,
         (sx[0] * sy[1] - sx[1] * sy[0]) / det}};
#pragma unroll
    for (int r = 0; r < 3; ++r) {      // aff[0][r] = sum_j inv[r][j] * dx[j], aff[1][r] = ... dy[j]
      c.aff[r] = static_cast<float>(inv[r][0] * dx[0] + inv[r][1] * dx[1] + inv[r][2] * dx[2]);
      c.aff[3 + r] = static_cast<float>(inv[r][0] * dy[0] + inv[r][1] * dy[1] + inv[r][2] * dy[2]);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) c.inv_aff[i] = static_cast<float>(ld_any(q.inv_aff, b * 9 + i, m & 1024u));
    // ---- P = K [R | -R T]  (fp32, un-contracted multiply-adds in the reference's matmul order)
    float Rt[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      Rt[r][0] = R[3 * r]; Rt[r][1] = R[3 * r + 1]; Rt[r][2] = R[3 * r + 2];
      Rt[r][3] = -fadd(fadd(fmul(R[3 * r], T[0]), fmul(R[3 * r + 1], T[1])), fmul(R[3 * r + 2], T[2]));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      c.P[j] = fadd(fmul(fx, Rt[0][j]), fmul(cx, Rt[2][j]));
      c.P[4 + j] = fadd(fmul(fy, Rt[1][j]), fmul(cy, Rt[2][j]));
      c.P[8 + j] = Rt[2][j];
    }
    c.Kinv[0] = fdiv(1.f, fx); c.Kinv[1] = 0.f; c.Kinv[2] = fdiv(-cx, fx);
    c.Kinv[3] = 0.f; c.Kinv[4] = fdiv(1.f, fy); c.Kinv[5] = fdiv(-cy, fy);
    c.Kinv[6] = 0.f; c.Kinv[7] = 0.f; c.Kinv[8] = 1.f;
    c.wh[0] = static_cast<float>(cen0 * 2.0); c.wh[1] = static_cast<float>(cen1 * 2.0);
    local_max = fmaxf(local_max, fmaxf(c.wh[0], c.wh[1]));
    c.clamp_max = 0.f;
#pragma unroll
    for (int i = 0; i < 7; ++i) c.pad[i] = 0.f;
    float* dst = out + (static_cast<int64_t>(b) * views + v) * MVG_CAM_FLOATS;
    const float* src = reinterpret_cast<const float*>(&c);
#pragma unroll
    for (int i = 0; i < MVG_CAM_FLOATS; ++i) dst[i] = src[i];
  }
  // clamp bound of dq_decoder.py:383: max over the whole (B,2) image-size tensor of this view
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = local_max;
  __syncthreads();
  float mx = -INFINITY;
  for (int w = 0; w < (blockDim.x + 31) / 32; ++w) mx = fmaxf(mx, s_max[w]);
  for (int b = threadIdx.x; b < batch; b += blockDim.x)
    out[(static_cast<int64_t>(b) * views + v) * MVG_CAM_FLOATS + 56] = mx;
}

}  // namespace mvg

extern "C" int mvg_pack_cameras(const void* const* fields, const int* dtypes, int batch, int views, float img_w,
                                float img_h, float* cams, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(fields && dtypes && cams, "mvg_pack_cameras: null pointer");
  MVG_REQUIRE(batch > 0 && views > 0 && views <= MVG_MAX_VIEWS, "mvg_pack_cameras: batch %d views %d (max %d views)",
              batch, views, MVG_MAX_VIEWS);
  CamPackArgs args;
  for (int v = 0; v < views; ++v) {
    const void* const* f = fields + v * MVG_CAM_FIELDS;
    uint32_t mask = 0;
    for (int i = 0; i < MVG_CAM_FIELDS; ++i) {
      MVG_REQUIRE(f[i] != nullptr, "mvg_pack_cameras: view %d field %d is null", v, i);
      const int dt = dtypes[v * MVG_CAM_FIELDS + i];
      MVG_REQUIRE(dt == MVG_F32 || dt == MVG_F64, "mvg_pack_cameras: view %d field %d dtype %d (float32 / float64 only)",
                  v, i, dt);
      MVG_REQUIRE((reinterpret_cast<uintptr_t>(f[i]) & (dt == MVG_F64 ? 7 : 3)) == 0,
                  "mvg_pack_cameras: view %d field %d is misaligned", v, i);
      if (dt == MVG_F64) mask |= 1u << i;
    }
    args.v[v] = CamViewPtrs{f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8], f[9], f[10]};
    args.f64_mask[v] = mask;
  }
  const int threads = batch < 32 ? 32 : (batch > 256 ? 256 : ((batch + 31) / 32) * 32);
  pack_cameras_kernel<<<views, threads, 0, static_cast<cudaStream_t>(stream)>>>(args, batch, views, img_w, img_h, cams);
  return check_launch("mvg_pack_cameras");
}
