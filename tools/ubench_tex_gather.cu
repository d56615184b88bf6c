// Micro-benchmark 4 (round 2): the texture path as the gather's data path.
//   The value map of one (view, level) as a pitch-linear 2D texture of half4 texels: 64 planes
//   (8 heads x 8 four-channel groups) of (H + 1) rows (one zero row between planes), bilinear
//   filtering + zero border done by the texture unit.  A warp gathers one (item, head): lane =
//   sample quarter * 8 + channel group, 6 fetches per lane = 24 samples x 8 groups, all inside a
//   window of `win` x `win` texels per unit of 128 items (the L1-resident tile of the binned gather).
//   MODE 0: float4 result + 4 FFMA;  MODE 1: v2.f16x2 result + 2 HFMA2.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_tex_gather tools/ubench_tex_gather.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int MODE>
__global__ void __launch_bounds__(512) tex_gather(cudaTextureObject_t tex, int W, int H, int win, int units,
                                                   float* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int g = lane & 7, sq = lane >> 3;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  __half2 hacc[2] = {__half2{}, __half2{}};
  const float inv = static_cast<float>(win) / 65536.f;
  for (int unit = blockIdx.x; unit < units; unit += gridDim.x) {
    const uint32_t hu = hash32(unit * 2654435761u + 17u);
    const int head = hu & 7;
    const float ox = static_cast<float>((hu >> 3) % static_cast<uint32_t>(W - win));
    const float oy = static_cast<float>((hu >> 13) % static_cast<uint32_t>(H - win)) +
                     static_cast<float>((head * 8 + g) * (H + 1));
    for (int it = warp; it < 128; it += nwarp) {
      const uint32_t hi = hash32(hu + it * 0x9E3779B9u);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const uint32_t hs = hash32(hi + (sq + 4 * i) * 0x85ebca6bu);
        const float x = ox + static_cast<float>(hs & 0xffffu) * inv;
        const float y = oy + static_cast<float>(hs >> 16) * inv;
        const float w = static_cast<float>(hs & 0xffu) * (1.f / 4096.f);
        if (MODE == 0) {
          const float4 v = tex2D<float4>(tex, x, y);
          acc[0] = fmaf(w, v.x, acc[0]); acc[1] = fmaf(w, v.y, acc[1]);
          acc[2] = fmaf(w, v.z, acc[2]); acc[3] = fmaf(w, v.w, acc[3]);
        } else {
          uint32_t a, b;
          asm volatile("tex.2d.v2.f16x2.f32 {%0, %1}, [%2, {%3, %4}];" : "=r"(a), "=r"(b) : "l"(tex), "f"(x), "f"(y));
          const __half2 wh = __float2half2_rn(w);
          hacc[0] = __hfma2(wh, *reinterpret_cast<__half2*>(&a), hacc[0]);
          hacc[1] = __hfma2(wh, *reinterpret_cast<__half2*>(&b), hacc[1]);
        }
      }
    }
  }
  float s = acc[0] + acc[1] + acc[2] + acc[3] + __low2float(hacc[0]) + __high2float(hacc[0]) +
            __low2float(hacc[1]) + __high2float(hacc[1]);
  if (s == 123.456f) out[0] = s;
}

// accuracy of the 9-bit filter weights: max |tex - exact fp32 bilinear of the fp16 texels| over random points
__global__ void tex_accuracy(cudaTextureObject_t tex, const __half* base, size_t pitch_elems, int W, int H, int n,
                             float* max_err, float* sum_err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t h = hash32(i * 2654435761u + 5u), h2 = hash32(h + 99u);
  const float im_x = static_cast<float>(h & 0xffffffu) * (static_cast<float>(W + 1) / 16777216.f) - 1.f;   // (-1, W)
  const float im_y = static_cast<float>(h2 & 0xffffffu) * (static_cast<float>(H + 1) / 16777216.f) - 1.f;
  const float4 v = tex2D<float4>(tex, im_x + 0.5f, im_y + 0.5f);
  const float fx = floorf(im_x), fy = floorf(im_y);
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
  const float lx = im_x - fx, ly = im_y - fy;
  float ref[4] = {0.f, 0.f, 0.f, 0.f};
  for (int dy = 0; dy < 2; ++dy)
    for (int dx = 0; dx < 2; ++dx) {
      const int x = x0 + dx, y = y0 + dy;
      if (x < 0 || x >= W || y < 0 || y >= H) continue;
      const float wgt = (dx ? lx : 1.f - lx) * (dy ? ly : 1.f - ly);
      for (int c = 0; c < 4; ++c) ref[c] += wgt * __half2float(base[y * pitch_elems + x * 4 + c]);
    }
  const float got[4] = {v.x, v.y, v.z, v.w};
  float e = 0.f;
  for (int c = 0; c < 4; ++c) e = fmaxf(e, fabsf(got[c] - ref[c]));
  atomicMax(reinterpret_cast<int*>(max_err), __float_as_int(e));
  atomicAdd(sum_err, e);
}

int main(int argc, char** argv) {
  const int W = 240, H = 128, planes = 64;
  const int rows = planes * (H + 1);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, texturePitchAlignment %zu, maxTexture2DLinear %d x %d (pitch %d)\n", prop.name, sms,
         prop.texturePitchAlignment, prop.maxTexture2DLinear[0], prop.maxTexture2DLinear[1], prop.maxTexture2DLinear[2]);
  const size_t pitch = W * 8;   // 1920 B, multiple of 32
  std::vector<__half> host(static_cast<size_t>(rows) * W * 4);
  uint32_t s = 12345u;
  for (int r = 0; r < rows; ++r)
    for (int i = 0; i < W * 4; ++i) {
      s = s * 1664525u + 1013904223u;
      const float v = (r % (H + 1) == H) ? 0.f : (static_cast<float>(s >> 8) / 8388608.f - 1.f) * 2.f;
      host[static_cast<size_t>(r) * W * 4 + i] = __float2half(v);
    }
  __half* dmap;
  CK(cudaMalloc(&dmap, host.size() * 2));
  CK(cudaMemcpy(dmap, host.data(), host.size() * 2, cudaMemcpyHostToDevice));
  cudaResourceDesc res = {};
  res.resType = cudaResourceTypePitch2D;
  res.res.pitch2D.devPtr = dmap;
  res.res.pitch2D.desc = cudaCreateChannelDescHalf4();
  res.res.pitch2D.width = W;
  res.res.pitch2D.height = rows;
  res.res.pitch2D.pitchInBytes = pitch;
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
  td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  cudaTextureObject_t tex;
  CK(cudaCreateTextureObject(&tex, &res, &td, nullptr));
  float* dout;
  CK(cudaMalloc(&dout, 64));
  CK(cudaMemset(dout, 0, 64));

  // accuracy (plane 0)
  {
    const int n = 1 << 20;
    tex_accuracy<<<n / 256, 256>>>(tex, dmap, pitch / 2, W, H, n, dout, dout + 1);
    CK(cudaDeviceSynchronize());
    float errs[2];
    CK(cudaMemcpy(errs, dout, 8, cudaMemcpyDeviceToHost));
    printf("filter accuracy vs exact fp32 bilinear (texels U(-2,2)): max %.5f mean %.6f\n", errs[0], errs[1] / n);
    CK(cudaMemset(dout, 0, 64));
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int units = 148 * 32;
  const double fetches = static_cast<double>(units) * 128 * 24 * 8;   // quad fetches
  int clk_khz = prop.clockRate;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("units %d x 128 items x 24 samples x 8 groups = %.1f M fetches; job = 103 M fetches per gather launch\n", units,
         fetches * 1e-6);
  for (int mode = 0; mode < 2; ++mode)
    for (int win : {24, 40, 64, 120})
      for (int threads : {256, 512})
        for (int cps : {1, 2, 4}) {
          if (threads * cps > 2048) continue;
          const int grid = sms * cps;
          auto launch = [&]() {
            if (mode == 0) tex_gather<0><<<grid, threads>>>(tex, W, H, win, units, dout);
            else tex_gather<1><<<grid, threads>>>(tex, W, H, win, units, dout);
          };
          launch();
          CK(cudaDeviceSynchronize());
          CK(cudaEventRecord(e0));
          for (int r = 0; r < 5; ++r) launch();
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          ms /= 5;
          const double per_clk_sm = fetches / (ms * 1e-3) / (static_cast<double>(clk_khz) * 1e3) / sms;
          printf("mode %d (%s) win %3d threads %d x %d CTA/SM: %.3f ms  %.2f fetches/clk/SM (at %d MHz)  -> 103 M fetches in %.1f us\n",
                 mode, mode ? "f16x2 + HFMA2" : "f32 + FFMA", win, threads, cps, ms, per_clk_sm, clk_khz / 1000,
                 103.0e6 / (fetches / (ms * 1e-3)) * 1e6);
        }
  return 0;
}
