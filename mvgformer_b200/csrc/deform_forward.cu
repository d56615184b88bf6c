// Multi-scale deformable-attention sampling, forward (drop-in for Deformable.deform_forward).
//
// Semantics: lib/models/ops/src/cuda/deform_im2col_cuda.cuh:247-309 (kernel) and :41-93
// (bilinear) of the reference.  The reference maps one thread to one output scalar
// (b,q,m,c) and issues 96 scalar loads per thread; here a group of G = D*sizeof(T)/16
// lanes owns one (b,q,m) and every lane gathers 16 B (4 fp32 / 8 bf16 channels) per
// corner, so a warp-wide LDG.128 serves 32/G sample corners at once.  Accumulation is fp32.
// The integer path (level start, floor, corner validity tests) is evaluated exactly as the
// reference does, without FMA contraction on h_im / w_im.
#include "common.cuh"

namespace mvg {

template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int kElems = 4;
  __device__ static void load(const float* p, float* f) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
  __device__ static void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
  __device__ static float scalar(const float* p) { return __ldg(p); }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int kElems = 8;
  __device__ static void load(const __nv_bfloat16* p, float* f) { unpack8(ldg_nc_v4(p), f); }
  __device__ static void store(__nv_bfloat16* p, const float* f) {
    uint4 q;
    q.x = pack_bf16x2(f[0], f[1]); q.y = pack_bf16x2(f[2], f[3]);
    q.z = pack_bf16x2(f[4], f[5]); q.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = q;
  }
  __device__ static float scalar(const __nv_bfloat16* p) {
    return __bfloat162float(__ldg(p));
  }
};

struct LevelInfo {
  int h[MVG_MAX_LEVELS];
  int w[MVG_MAX_LEVELS];
  int start[MVG_MAX_LEVELS];
};

// Reads the int64 (H,W) / start tensors once per thread block (the reference re-reads them
// from global memory in every thread, deform_im2col_cuda.cuh:283-286).
__device__ __forceinline__ void load_levels(const int64_t* shapes, const int64_t* lsi,
                                            int num_levels, LevelInfo* s) {
  if (threadIdx.x < num_levels) {
    s->h[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x]);
    s->w[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x + 1]);
    s->start[threadIdx.x] = static_cast<int>(lsi[threadIdx.x]);
  }
  __syncthreads();
}

template <typename T, int D>
__global__ void __launch_bounds__(256)
deform_forward_kernel(const T* __restrict__ value, const int64_t* __restrict__ shapes,
                      const int64_t* __restrict__ lsi, const T* __restrict__ loc,
                      const T* __restrict__ attn, int64_t n_groups, int spatial_size,
                      int num_heads, int num_levels, int num_query, int num_point,
                      T* __restrict__ out) {
  constexpr int E = Vec16<T>::kElems;   // channels per lane
  constexpr int G = D / E;              // lanes per (b,q,m)
  __shared__ LevelInfo lv;
  load_levels(shapes, lsi, num_levels, &lv);

  const int64_t gid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / G;
  const int sub = threadIdx.x % G;
  if (gid >= n_groups) return;
  const int m = static_cast<int>(gid % num_heads);
  const int64_t bq = gid / num_heads;
  const int b = static_cast<int>(bq / num_query);

  const int row_stride = num_heads * D;  // elements per spatial position
  const T* vbase = value + static_cast<int64_t>(b) * spatial_size * row_stride + m * D + sub * E;
  const int64_t samp0 = gid * num_levels * num_point;

  float acc[E];
#pragma unroll
  for (int i = 0; i < E; ++i) acc[i] = 0.f;

  for (int l = 0; l < num_levels; ++l) {
    const int H = lv.h[l], W = lv.w[l];
    const T* vl = vbase + static_cast<int64_t>(lv.start[l]) * row_stride;
    const float fH = static_cast<float>(H), fW = static_cast<float>(W);
#pragma unroll 2
    for (int p = 0; p < num_point; ++p) {
      const int64_t s = samp0 + l * num_point + p;
      const float loc_w = Vec16<T>::scalar(loc + 2 * s);
      const float loc_h = Vec16<T>::scalar(loc + 2 * s + 1);
      const float wgt = Vec16<T>::scalar(attn + s);
      const float h_im = fsub(fmul(loc_h, fH), 0.5f);
      const float w_im = fsub(fmul(loc_w, fW), 0.5f);
      if (h_im > -1.f && w_im > -1.f && h_im < fH && w_im < fW) {
        const int h_low = static_cast<int>(floorf(h_im));
        const int w_low = static_cast<int>(floorf(w_im));
        const int h_high = h_low + 1, w_high = w_low + 1;
        const float lh = h_im - static_cast<float>(h_low);
        const float lw = w_im - static_cast<float>(w_low);
        const float hh = 1.f - lh, hw = 1.f - lw;
        float v1[E], v2[E], v3[E], v4[E];
        const bool ok1 = h_low >= 0 && w_low >= 0;
        const bool ok2 = h_low >= 0 && w_high <= W - 1;
        const bool ok3 = h_high <= H - 1 && w_low >= 0;
        const bool ok4 = h_high <= H - 1 && w_high <= W - 1;
#pragma unroll
        for (int i = 0; i < E; ++i) v1[i] = v2[i] = v3[i] = v4[i] = 0.f;
        if (ok1) Vec16<T>::load(vl + (static_cast<int64_t>(h_low) * W + w_low) * row_stride, v1);
        if (ok2) Vec16<T>::load(vl + (static_cast<int64_t>(h_low) * W + w_high) * row_stride, v2);
        if (ok3) Vec16<T>::load(vl + (static_cast<int64_t>(h_high) * W + w_low) * row_stride, v3);
        if (ok4) Vec16<T>::load(vl + (static_cast<int64_t>(h_high) * W + w_high) * row_stride, v4);
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
#pragma unroll
        for (int i = 0; i < E; ++i) {
          const float val = w1 * v1[i] + w2 * v2[i] + w3 * v3[i] + w4 * v4[i];
          acc[i] += val * wgt;
        }
      }
    }
  }
  Vec16<T>::store(out + gid * D + sub * E, acc);
}

// float64 (the reference dispatches AT_DISPATCH_FLOATING_TYPES, deform_cuda.cu:75): used for gradient
// checks only, so this is the plain formulation - one lane per (b, q, head, channel), a warp = the 32
// channels of one head (128-byte coalesced corner reads), double arithmetic in the reference's order
// (deform_im2col_cuda.cuh:247-309, :41-93).
__global__ void __launch_bounds__(256)
deform_forward_f64_kernel(const double* __restrict__ value, const int64_t* __restrict__ shapes,
                          const int64_t* __restrict__ lsi, const double* __restrict__ loc_all,
                          const double* __restrict__ attn_all, int64_t n, int spatial_size, int num_heads,
                          int num_levels, int num_query, int num_point, double* __restrict__ out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int c = static_cast<int>(idx % 32);
  const int64_t g = idx / 32;                              // (b, q, m)
  const int m = static_cast<int>(g % num_heads);
  const int64_t b = g / (static_cast<int64_t>(num_heads) * num_query);
  const double* loc = loc_all + g * num_levels * num_point * 2;
  const double* attn = attn_all + g * num_levels * num_point;
  const double* vb = value + b * spatial_size * num_heads * 32;
  double acc = 0.0;
  for (int l = 0; l < num_levels; ++l) {
    const int H = static_cast<int>(shapes[2 * l]), W = static_cast<int>(shapes[2 * l + 1]);
    const double* vl = vb + lsi[l] * num_heads * 32 + m * 32 + c;
    for (int p = 0; p < num_point; ++p) {
      const double loc_w = loc[(l * num_point + p) * 2], loc_h = loc[(l * num_point + p) * 2 + 1];
      const double wgt = attn[l * num_point + p];
      const double h_im = loc_h * H - 0.5, w_im = loc_w * W - 0.5;
      if (h_im > -1 && w_im > -1 && h_im < H && w_im < W) {
        const int h_low = static_cast<int>(floor(h_im)), w_low = static_cast<int>(floor(w_im));
        const double lh = h_im - h_low, lw = w_im - w_low, hh = 1 - lh, hw = 1 - lw;
        const int64_t rs = static_cast<int64_t>(num_heads) * 32;
        double v1 = 0, v2 = 0, v3 = 0, v4 = 0;
        if (h_low >= 0 && w_low >= 0) v1 = vl[(static_cast<int64_t>(h_low) * W + w_low) * rs];
        if (h_low >= 0 && w_low + 1 <= W - 1) v2 = vl[(static_cast<int64_t>(h_low) * W + w_low + 1) * rs];
        if (h_low + 1 <= H - 1 && w_low >= 0) v3 = vl[(static_cast<int64_t>(h_low + 1) * W + w_low) * rs];
        if (h_low + 1 <= H - 1 && w_low + 1 <= W - 1) v4 = vl[(static_cast<int64_t>(h_low + 1) * W + w_low + 1) * rs];
        acc += (hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4) * wgt;
      }
    }
  }
  out[idx] = acc;
}

}  // namespace mvg

extern "C" int mvg_deform_forward(const void* value, const int64_t* spatial_shapes,
                                  const int64_t* level_start_index, const void* sampling_loc,
                                  const void* attn_weight, int dtype, int batch,
                                  int spatial_size, int num_heads, int channels, int num_levels,
                                  int num_query, int num_point, int im2col_step, void* out,
                                  void* stream) {
  using namespace mvg;
  MVG_REQUIRE(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && out,
              "mvg_deform_forward: null pointer");
  MVG_REQUIRE(channels == 32, "mvg_deform_forward: channels per head must be 32, got %d", channels);
  MVG_REQUIRE(num_levels >= 1 && num_levels <= MVG_MAX_LEVELS,
              "mvg_deform_forward: num_levels %d out of range", num_levels);
  MVG_REQUIRE(batch > 0 && num_query > 0 && num_heads > 0 && num_point > 0 && spatial_size > 0,
              "mvg_deform_forward: empty shape");
  const int step = batch < im2col_step ? batch : im2col_step;
  MVG_REQUIRE(step > 0 && batch % step == 0,
              "batch(%d) must divide im2col_step(%d)", batch, step);  // deform_cuda.cu:63
  const int64_t n_groups = static_cast<int64_t>(batch) * num_query * num_heads;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int threads = 256;
  if (dtype == MVG_F32) {
    const int64_t blocks = (n_groups * 8 + threads - 1) / threads;
    deform_forward_kernel<float, 32><<<static_cast<unsigned>(blocks), threads, 0, st>>>(
        static_cast<const float*>(value), spatial_shapes, level_start_index,
        static_cast<const float*>(sampling_loc), static_cast<const float*>(attn_weight),
        n_groups, spatial_size, num_heads, num_levels, num_query, num_point,
        static_cast<float*>(out));
  } else if (dtype == MVG_BF16) {
    const int64_t blocks = (n_groups * 4 + threads - 1) / threads;
    deform_forward_kernel<__nv_bfloat16, 32><<<static_cast<unsigned>(blocks), threads, 0, st>>>(
        static_cast<const __nv_bfloat16*>(value), spatial_shapes, level_start_index,
        static_cast<const __nv_bfloat16*>(sampling_loc),
        static_cast<const __nv_bfloat16*>(attn_weight), n_groups, spatial_size, num_heads,
        num_levels, num_query, num_point, static_cast<__nv_bfloat16*>(out));
  } else if (dtype == MVG_F64) {
    const int64_t n = n_groups * 32;
    deform_forward_f64_kernel<<<static_cast<unsigned>((n + threads - 1) / threads), threads, 0, st>>>(
        static_cast<const double*>(value), spatial_shapes, level_start_index,
        static_cast<const double*>(sampling_loc), static_cast<const double*>(attn_weight), n, spatial_size,
        num_heads, num_levels, num_query, num_point, static_cast<double*>(out));
  } else {
    set_error("mvg_deform_forward: unsupported dtype %d", dtype);
    return MVG_EUNSUPPORTED;
  }
  return check_launch("mvg_deform_forward");
}
