// Multi-scale deformable-attention sampling, backward (drop-in for Deformable.deform_backward).
//
// Semantics: lib/models/ops/src/cuda/deform_im2col_cuda.cuh:96-203 (col2im bilinear) and the
// col2im kernels :311-930; host wrapper lib/models/ops/src/cuda/deform_cuda.cu:94-164.
// The reference launches one thread per (b,q,m,c) and tree-reduces grad_sampling_loc /
// grad_attn_weight over the channel threads of a block in shared memory.  Here a group of 8
// lanes owns one (b,q,m): each lane holds 4 channels, the channel reduction is three
// warp-shuffle steps, and grad_value is scattered with 16-byte vector reductions
// (red.global.add.v4.f32) instead of 4 scalar atomics.
#include "common.cuh"

namespace mvg {

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b),
               "f"(c), "f"(d)
               : "memory");
}

__global__ void __launch_bounds__(256)
deform_backward_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                       const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                       const float* __restrict__ attn, const float* __restrict__ grad_out,
                       int64_t n_groups, int spatial_size, int num_heads, int num_levels,
                       int num_query, int num_point, float* __restrict__ grad_value,
                       float* __restrict__ grad_loc, float* __restrict__ grad_attn) {
  constexpr int D = 32, E = 4, G = D / E;
  __shared__ int s_h[MVG_MAX_LEVELS], s_w[MVG_MAX_LEVELS], s_start[MVG_MAX_LEVELS];
  if (threadIdx.x < num_levels) {
    s_h[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x]);
    s_w[threadIdx.x] = static_cast<int>(shapes[2 * threadIdx.x + 1]);
    s_start[threadIdx.x] = static_cast<int>(lsi[threadIdx.x]);
  }
  __syncthreads();
  int64_t gid = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / G;
  const int sub = threadIdx.x % G;
  const bool active = gid < n_groups;
  if (!active) gid = n_groups - 1;            // keep the warp converged for the shuffles
  const int m = static_cast<int>(gid % num_heads);
  const int b = static_cast<int>(gid / num_heads / num_query);
  const int row_stride = num_heads * D;
  const int64_t voff = static_cast<int64_t>(b) * spatial_size * row_stride + m * D + sub * E;
  const float4 go = __ldg(reinterpret_cast<const float4*>(grad_out + gid * D + sub * E));
  const int64_t samp0 = gid * num_levels * num_point;

  for (int l = 0; l < num_levels; ++l) {
    const int H = s_h[l], W = s_w[l];
    const float fH = static_cast<float>(H), fW = static_cast<float>(W);
    const int64_t lvl_off = voff + static_cast<int64_t>(s_start[l]) * row_stride;
    for (int p = 0; p < num_point; ++p) {
      const int64_t s = samp0 + l * num_point + p;
      const float loc_w = __ldg(loc + 2 * s), loc_h = __ldg(loc + 2 * s + 1);
      const float wgt = __ldg(attn + s);
      const float h_im = fsub(fmul(loc_h, fH), 0.5f);
      const float w_im = fsub(fmul(loc_w, fW), 0.5f);
      float g_w = 0.f, g_h = 0.f, g_a = 0.f;
      if (h_im > -1.f && w_im > -1.f && h_im < fH && w_im < fW) {
        const int h_low = static_cast<int>(floorf(h_im)), w_low = static_cast<int>(floorf(w_im));
        const int h_high = h_low + 1, w_high = w_low + 1;
        const float lh = h_im - static_cast<float>(h_low), lw = w_im - static_cast<float>(w_low);
        const float hh = 1.f - lh, hw = 1.f - lw;
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const float tg[4] = {go.x * wgt, go.y * wgt, go.z * wgt, go.w * wgt};
        float val[4] = {0.f, 0.f, 0.f, 0.f}, gh[4] = {0.f, 0.f, 0.f, 0.f}, gw[4] = {0.f, 0.f, 0.f, 0.f};
        auto corner = [&](bool ok, int hy, int wx, float wc, float ch, float cw) {
          if (!ok) return;
          const int64_t o = lvl_off + (static_cast<int64_t>(hy) * W + wx) * row_stride;
          const float4 v = __ldg(reinterpret_cast<const float4*>(value + o));
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            val[i] += wc * vv[i];
            gh[i] += ch * vv[i];
            gw[i] += cw * vv[i];
          }
          if (active) red_add_v4(grad_value + o, wc * tg[0], wc * tg[1], wc * tg[2], wc * tg[3]);
        };
        corner(h_low >= 0 && w_low >= 0, h_low, w_low, w1, -hw, -hh);
        corner(h_low >= 0 && w_high <= W - 1, h_low, w_high, w2, -lw, hh);
        corner(h_high <= H - 1 && w_low >= 0, h_high, w_low, w3, hw, -lh);
        corner(h_high <= H - 1 && w_high <= W - 1, h_high, w_high, w4, lw, lh);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float gi = (i == 0 ? go.x : i == 1 ? go.y : i == 2 ? go.z : go.w);
          g_a += gi * val[i];
          g_h += tg[i] * gh[i];
          g_w += tg[i] * gw[i];
        }
      }
#pragma unroll
      for (int o = 1; o < G; o <<= 1) {       // reduce over the 8 lanes (= 32 channels) of the group
        g_a += __shfl_xor_sync(0xffffffffu, g_a, o);
        g_h += __shfl_xor_sync(0xffffffffu, g_h, o);
        g_w += __shfl_xor_sync(0xffffffffu, g_w, o);
      }
      if (active && sub == 0) {
        grad_attn[s] = g_a;
        grad_loc[2 * s] = fW * g_w;
        grad_loc[2 * s + 1] = fH * g_h;
      }
    }
  }
}

}  // namespace mvg

extern "C" int mvg_deform_backward(const float* value, const int64_t* spatial_shapes,
                                   const int64_t* level_start_index, const float* sampling_loc,
                                   const float* attn_weight, const float* grad_output, int batch,
                                   int spatial_size, int num_heads, int channels, int num_levels,
                                   int num_query, int num_point, int im2col_step,
                                   float* grad_value, float* grad_sampling_loc,
                                   float* grad_attn_weight, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(value && spatial_shapes && level_start_index && sampling_loc && attn_weight &&
                  grad_output && grad_value && grad_sampling_loc && grad_attn_weight,
              "mvg_deform_backward: null pointer");
  MVG_REQUIRE(channels == 32, "mvg_deform_backward: channels per head must be 32, got %d", channels);
  MVG_REQUIRE(num_levels >= 1 && num_levels <= MVG_MAX_LEVELS, "mvg_deform_backward: num_levels %d", num_levels);
  MVG_REQUIRE(batch > 0 && num_query > 0 && num_heads > 0 && num_point > 0, "mvg_deform_backward: empty shape");
  const int step = batch < im2col_step ? batch : im2col_step;
  MVG_REQUIRE(step > 0 && batch % step == 0, "batch(%d) must divide im2col_step(%d)", batch, step);
  const int64_t n_groups = static_cast<int64_t>(batch) * num_query * num_heads;
  const int64_t blocks = (n_groups * 8 + 255) / 256;
  deform_backward_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, n_groups,
      spatial_size, num_heads, num_levels, num_query, num_point, grad_value, grad_sampling_loc,
      grad_attn_weight);
  return check_launch("mvg_deform_backward");
}
