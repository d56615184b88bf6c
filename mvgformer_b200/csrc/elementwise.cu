// Small HBM-bound stages of DQDecoderLayer.update_feature and the class head
// (lib/models/dq_decoder.py:763-778, :845-848, :889-893; lib/models/mvp_decoder.py:94-98).
// Each is one pass over its operands with 16-byte accesses; they exist so that the layer does
// not bounce through a dozen eager elementwise launches.
#include "common.cuh"

namespace mvg {

// aver[b,n,:] = (1/V) sum_v bounding[b,v,n] * x[b,v,n,:]       (:585-586 mask, :770 mean)
__global__ void __launch_bounds__(256)
masked_view_mean_kernel(const __nv_bfloat16* __restrict__ x, const uint8_t* __restrict__ bounding,
                        int B, int V, int64_t N, int C, __nv_bfloat16* __restrict__ out) {
  pdl_enter();
  const int vec_per_row = C / 8;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t total = static_cast<int64_t>(B) * N * vec_per_row;
  if (idx >= total) return;
  const int cv = static_cast<int>(idx % vec_per_row);
  const int64_t bn = idx / vec_per_row;
  const int b = static_cast<int>(bn / N);
  const int64_t n = bn % N;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int v = 0; v < V; ++v) {
    const int64_t row = (static_cast<int64_t>(b) * V + v) * N + n;
    if (bounding[row]) {
      float t[8];
      unpack8(ldg_nc_v4(x + row * C + cv * 8), t);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += t[i];
    }
  }
  const float inv = 1.f / static_cast<float>(V);
  uint4 o;
  o.x = pack_bf16x2(acc[0] * inv, acc[1] * inv); o.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
  o.z = pack_bf16x2(acc[4] * inv, acc[5] * inv); o.w = pack_bf16x2(acc[6] * inv, acc[7] * inv);
  *reinterpret_cast<uint4*>(out + bn * C + cv * 8) = o;
}

// out = LayerNorm(a + b) * gamma + beta over 256 channels; one warp per row, 8 channels / lane.
template <bool B_BF16>
__global__ void __launch_bounds__(256)
add_layernorm256_kernel(const float* __restrict__ a, const void* __restrict__ bptr,
                        const float* __restrict__ gamma, const float* __restrict__ beta,
                        int64_t rows, float eps, float* __restrict__ out_f32,
                        __nv_bfloat16* __restrict__ out_bf16) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float x[8];
  {
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(a + row * 256 + lane * 8));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(a + row * 256 + lane * 8 + 4));
    float t[8];
    if (B_BF16) {
      unpack8(ldg_nc_v4(static_cast<const __nv_bfloat16*>(bptr) + row * 256 + lane * 8), t);
    } else {
      const float* bf = static_cast<const float*>(bptr) + row * 256 + lane * 8;
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bf));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bf + 4));
      t[0] = b0.x; t[1] = b0.y; t[2] = b0.z; t[3] = b0.w;
      t[4] = b1.x; t[5] = b1.y; t[6] = b1.z; t[7] = b1.w;
    }
    x[0] = a0.x + t[0]; x[1] = a0.y + t[1]; x[2] = a0.z + t[2]; x[3] = a0.w + t[3];
    x[4] = a1.x + t[4]; x[5] = a1.y + t[5]; x[6] = a1.z + t[6]; x[7] = a1.w + t[7];
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / 256.f);
  float vs = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float d = x[i] - mean; vs += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, o);
  const float rstd = rsqrtf(vs * (1.f / 256.f) + eps);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + lane * 8));
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + lane * 8 + 4));
  const float4 e0 = __ldg(reinterpret_cast<const float4*>(beta + lane * 8));
  const float4 e1 = __ldg(reinterpret_cast<const float4*>(beta + lane * 8 + 4));
  const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float e[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
  float y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] = (x[i] - mean) * rstd * g[i] + e[i];
  if (out_f32) {
    float4* o = reinterpret_cast<float4*>(out_f32 + row * 256 + lane * 8);
    o[0] = make_float4(y[0], y[1], y[2], y[3]);
    o[1] = make_float4(y[4], y[5], y[6], y[7]);
  }
  if (out_bf16) {
    uint4 o;
    o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]);
    o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
    *reinterpret_cast<uint4*>(out_bf16 + row * 256 + lane * 8) = o;
  }
}

// prob[b,q,c] = mean_j sigmoid(cls[b, q*J + j, c])          (:889-893)
__global__ void __launch_bounds__(256)
class_prob_kernel(const float* __restrict__ cls, int64_t BQ, int J, float* __restrict__ prob) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= BQ * 2) return;
  const int64_t bq = idx >> 1;
  const int c = static_cast<int>(idx & 1);
  float s = 0.f;
  for (int j = 0; j < J; ++j) {
    const float z = __ldg(cls + (bq * J + j) * 2 + c);
    s += 1.f / (1.f + expf(-z));
  }
  prob[idx] = s / static_cast<float>(J);
}

// class_embed (256 -> 2) + sigmoid + mean over joints; one warp per (b, q).
__global__ void __launch_bounds__(256)
class_head_kernel(const float* __restrict__ x, const float* __restrict__ w,
                  const float* __restrict__ bias, int64_t BQ, int J, float* __restrict__ prob) {
  const int lane = threadIdx.x & 31;
  const int64_t bq = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (bq >= BQ) return;
  float w0[8], w1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    w0[i] = __ldg(w + lane * 8 + i);
    w1[i] = __ldg(w + 256 + lane * 8 + i);
  }
  const float b0 = __ldg(bias), b1 = __ldg(bias + 1);
  float s0 = 0.f, s1 = 0.f;
  for (int j = 0; j < J; ++j) {
    const float* row = x + (bq * J + j) * 256 + lane * 8;
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(row));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(row + 4));
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { d0 += a[i] * w0[i]; d1 += a[i] * w1[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      d0 += __shfl_xor_sync(0xffffffffu, d0, o);
      d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    }
    s0 += 1.f / (1.f + expf(-(d0 + b0)));
    s1 += 1.f / (1.f + expf(-(d1 + b1)));
  }
  if (lane == 0) {
    prob[bq * 2] = s0 / static_cast<float>(J);
    prob[bq * 2 + 1] = s1 / static_cast<float>(J);
  }
}

// out_bf16 = bf16(a + b)   (with_pos_embed + cast for the tensor-core operand, dq_decoder.py:580)
__global__ void __launch_bounds__(256)
add_cast_bf16_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n8,
                     __nv_bfloat16* __restrict__ out) {
  pdl_enter();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(a) + 2 * i);
  const float4 a1 = __ldg(reinterpret_cast<const float4*>(a) + 2 * i + 1);
  float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
  if (b != nullptr) {
    b0 = __ldg(reinterpret_cast<const float4*>(b) + 2 * i);
    b1 = __ldg(reinterpret_cast<const float4*>(b) + 2 * i + 1);
  }
  uint4 o;
  o.x = pack_bf16x2(a0.x + b0.x, a0.y + b0.y); o.y = pack_bf16x2(a0.z + b0.z, a0.w + b0.w);
  o.z = pack_bf16x2(a1.x + b1.x, a1.y + b1.y); o.w = pack_bf16x2(a1.z + b1.z, a1.w + b1.w);
  reinterpret_cast<uint4*>(out)[i] = o;
}

// class head, one CTA per (b, q), one warp per joint: dot products by warp shuffle, the mean over
// joints through shared memory.
__global__ void __launch_bounds__(512)
class_head_kernel_v2(const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ bias, int J, float* __restrict__ prob) {
  __shared__ float part[16][2];
  pdl_enter();
  const int lane = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int64_t bq = blockIdx.x;
  if (j < J) {
    const float* row = x + (bq * J + j) * 256 + lane * 8;
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(row));
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(row + 4));
    const float4 w00 = __ldg(reinterpret_cast<const float4*>(w + lane * 8));
    const float4 w01 = __ldg(reinterpret_cast<const float4*>(w + lane * 8 + 4));
    const float4 w10 = __ldg(reinterpret_cast<const float4*>(w + 256 + lane * 8));
    const float4 w11 = __ldg(reinterpret_cast<const float4*>(w + 256 + lane * 8 + 4));
    float d0 = a0.x * w00.x + a0.y * w00.y + a0.z * w00.z + a0.w * w00.w + a1.x * w01.x + a1.y * w01.y +
               a1.z * w01.z + a1.w * w01.w;
    float d1 = a0.x * w10.x + a0.y * w10.y + a0.z * w10.z + a0.w * w10.w + a1.x * w11.x + a1.y * w11.y +
               a1.z * w11.z + a1.w * w11.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      d0 += __shfl_xor_sync(0xffffffffu, d0, o);
      d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    }
    if (lane == 0) {
      part[j][0] = 1.f / (1.f + expf(-(d0 + __ldg(bias))));
      part[j][1] = 1.f / (1.f + expf(-(d1 + __ldg(bias + 1))));
    }
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int k = 0; k < J; ++k) s += part[k][threadIdx.x];
    prob[bq * 2 + threadIdx.x] = s / static_cast<float>(J);
  }
}

}  // namespace mvg

extern "C" int mvg_add_cast_bf16(const float* a, const float* b, void* out_bf16, int64_t n, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(a && out_bf16 && n > 0 && n % 8 == 0, "mvg_add_cast_bf16: bad argument (n must be a multiple of 8)");
  const int64_t n8 = n / 8;
  launch_k(add_cast_bf16_kernel, dim3(static_cast<unsigned>((n8 + 255) / 256)), dim3(256), 0,
           static_cast<cudaStream_t>(stream), a, b, n8, static_cast<__nv_bfloat16*>(out_bf16));
  return check_launch("mvg_add_cast_bf16");
}

extern "C" int mvg_class_head(const float* x, const float* w, const float* bias, int batch,
                              int queries, int joints, float* prob, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(x && w && bias && prob && batch > 0 && queries > 0 && joints > 0,
              "mvg_class_head: bad argument");
  const int64_t bq = static_cast<int64_t>(batch) * queries;
  if (joints <= 16)
    launch_k(class_head_kernel_v2, dim3(static_cast<unsigned>(bq)), dim3(joints * 32), 0,
             static_cast<cudaStream_t>(stream), x, w, bias, joints, prob);
  else
    class_head_kernel<<<static_cast<unsigned>((bq + 7) / 8), 256, 0,
                        static_cast<cudaStream_t>(stream)>>>(x, w, bias, bq, joints, prob);
  return check_launch("mvg_class_head");
}

extern "C" int mvg_masked_view_mean(const void* x_bf16, const uint8_t* bounding, int batch,
                                    int views, int points, int channels, void* out_bf16,
                                    void* stream) {
  using namespace mvg;
  MVG_REQUIRE(x_bf16 && bounding && out_bf16, "mvg_masked_view_mean: null pointer");
  MVG_REQUIRE(channels % 8 == 0 && batch > 0 && views > 0 && points > 0, "mvg_masked_view_mean: bad shape");
  const int64_t total = static_cast<int64_t>(batch) * points * (channels / 8);
  launch_k(masked_view_mean_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0,
           static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(x_bf16), bounding, batch, views,
           static_cast<int64_t>(points), channels, static_cast<__nv_bfloat16*>(out_bf16));
  return check_launch("mvg_masked_view_mean");
}

extern "C" int mvg_add_layernorm(const float* a, const void* b, int b_dtype, const float* gamma,
                                 const float* beta, int64_t rows, int channels, float eps,
                                 float* out_f32, void* out_bf16, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(a && b && gamma && beta && (out_f32 || out_bf16), "mvg_add_layernorm: null pointer");
  MVG_REQUIRE(channels == 256 && rows > 0, "mvg_add_layernorm: channels must be 256");
  const unsigned blocks = static_cast<unsigned>((rows + 7) / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (b_dtype == MVG_BF16)
    add_layernorm256_kernel<true><<<blocks, 256, 0, st>>>(a, b, gamma, beta, rows, eps, out_f32,
                                                         static_cast<__nv_bfloat16*>(out_bf16));
  else if (b_dtype == MVG_F32)
    add_layernorm256_kernel<false><<<blocks, 256, 0, st>>>(a, b, gamma, beta, rows, eps, out_f32,
                                                          static_cast<__nv_bfloat16*>(out_bf16));
  else {
    set_error("mvg_add_layernorm: unsupported dtype %d", b_dtype);
    return MVG_EUNSUPPORTED;
  }
  return check_launch("mvg_add_layernorm");
}

extern "C" int mvg_class_prob(const float* cls, int batch, int queries, int joints, float* prob,
                              void* stream) {
  using namespace mvg;
  MVG_REQUIRE(cls && prob && batch > 0 && queries > 0 && joints > 0, "mvg_class_prob: bad argument");
  const int64_t bq = static_cast<int64_t>(batch) * queries;
  class_prob_kernel<<<static_cast<unsigned>((bq * 2 + 255) / 256), 256, 0,
                      static_cast<cudaStream_t>(stream)>>>(cls, bq, joints, prob);
  return check_launch("mvg_class_prob");
}
