// Micro-benchmark 3 (round 2): the data path the tiled gather uses.
//   (a) warp-level LDS.128 gathers of 128-byte segments (two horizontally adjacent 64-byte
//       (texel, head) rows) at random 64-byte boundaries of a shared-memory tile, 4 segments per
//       warp instruction (one per quarter-warp) - alone, with the fp16 HFMA2 blend of the new
//       kernel, and with the bf16 -> fp32 unpack + FFMA2 blend of the round-1 kernel;
//   (b) cp.async.bulk (UBLKCP) staging of tile rows global -> shared memory, all SMs at once,
//       from an L2-resident region, double buffered.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_smem_gather tools/ubench_smem_gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// MODE 0: loads only; 1: + fp16 HFMA2 blend (4 per load); 2: + bf16 unpack + FFMA2 (8 ALU + 4 FFMA2 per load)
template <int MODE>
__global__ void __launch_bounds__(512, 1) lds_gather(const uint8_t* __restrict__ src, uint32_t tile_bytes,
                                                     uint32_t pitch, int iters, float* out) {
  extern __shared__ __align__(128) uint8_t tile[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x < 32) {
    const uint32_t rows = tile_bytes / pitch;
    if (threadIdx.x == 0) mbar_expect_tx(&bar, rows * pitch);
    __syncwarp();
    for (uint32_t r = threadIdx.x; r < rows; r += 32)
      bulk_g2s(tile + r * pitch, src + (static_cast<size_t>(blockIdx.x) * rows + r) * pitch, pitch, &bar);
  }
  mbar_wait(&bar, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t seed = (wid * 9781u + 12345u) ^ ((lane >> 3) * 0x9E3779B9u);
  const uint32_t nseg = (tile_bytes - pitch) / 64 - 2;
  const uint32_t tbase = smem_u32(tile) + (lane & 7) * 16;
  __half2 hacc[4] = {__half2{}, __half2{}, __half2{}, __half2{}};
  float facc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t xacc = 0;
#pragma unroll 4
  for (int it = 0; it < iters; ++it) {
    seed = seed * 1664525u + 1013904223u;
    const uint32_t seg = __umulhi(seed, nseg);
    const uint32_t a = tbase + seg * 64;
    uint4 t, b;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(a));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(a + pitch));
    if (MODE == 0) {
      xacc ^= t.x ^ t.w ^ b.y ^ b.z;
    } else if (MODE == 1) {
      const uint32_t wraw = 0x3c003c00u ^ (seed & 0x03ff03ffu);
      const __half2 wt = *reinterpret_cast<const __half2*>(&wraw);
      const uint32_t wraw2 = wraw ^ 0x00100010u;
      const __half2 wb = *reinterpret_cast<const __half2*>(&wraw2);
      const __half2* th = reinterpret_cast<const __half2*>(&t);
      const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
      for (int i = 0; i < 4; ++i) { hacc[i] = __hfma2(wt, th[i], hacc[i]); hacc[i] = __hfma2(wb, bh[i], hacc[i]); }
    } else {
      const float wt = __uint_as_float(0x3f000000u | (seed & 0xffffu)), wb = 1.f - wt;
      const uint32_t tu[4] = {t.x, t.y, t.z, t.w}, bu[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        facc[2 * i] = fmaf(wt, __uint_as_float(tu[i] << 16), facc[2 * i]);
        facc[2 * i + 1] = fmaf(wt, __uint_as_float(tu[i] & 0xffff0000u), facc[2 * i + 1]);
        facc[2 * i] = fmaf(wb, __uint_as_float(bu[i] << 16), facc[2 * i]);
        facc[2 * i + 1] = fmaf(wb, __uint_as_float(bu[i] & 0xffff0000u), facc[2 * i + 1]);
      }
    }
  }
  float r = __uint_as_float(xacc);
  for (int i = 0; i < 4; ++i) r += __low2float(hacc[i]) + __high2float(hacc[i]);
  for (int i = 0; i < 8; ++i) r += facc[i];
  if (r == 123.456f) out[0] = r;
}

// staging: every CTA streams tiles of `rows` rows x `row_bytes` from its own window of an L2-resident
// region into two shared-memory buffers (one in flight while the other is "consumed" = waited on)
__global__ void __launch_bounds__(128, 1) stage_tiles(const uint8_t* __restrict__ src, size_t region_bytes,
                                                      uint32_t rows, uint32_t row_bytes, uint32_t src_pitch,
                                                      int tiles, float* out) {
  extern __shared__ __align__(128) uint8_t bufs[];
  __shared__ uint64_t bar[2];
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const uint32_t tile_bytes = rows * row_bytes;
  if (threadIdx.x < 32) {
    uint32_t seed = blockIdx.x * 7919u + 17u;
    for (int t = 0; t < tiles + 1; ++t) {
      if (t < tiles) {
        seed = seed * 1664525u + 1013904223u;
        const size_t span = static_cast<size_t>(rows) * src_pitch;
        const size_t off = (static_cast<size_t>(__umulhi(seed, static_cast<uint32_t>((region_bytes - span) >> 7))) << 7);
        if (threadIdx.x == 0) mbar_expect_tx(&bar[t & 1], tile_bytes);
        __syncwarp();
        for (uint32_t r = threadIdx.x; r < rows; r += 32)
          bulk_g2s(bufs + (t & 1) * tile_bytes + r * row_bytes, src + off + static_cast<size_t>(r) * src_pitch, row_bytes, &bar[t & 1]);
      }
      if (t > 0) mbar_wait(&bar[(t - 1) & 1], ((t - 1) >> 1) & 1);
    }
  }
  __syncthreads();
  if (bufs[threadIdx.x] == 77 && out) out[1] = 1.f;
}

int main() {
  uint8_t* buf; float* out;
  const size_t region = 64ull << 20;
  cudaMalloc(&buf, region); cudaMalloc(&out, 64);
  cudaMemset(buf, 0x11, region);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, nominal %.0f MHz\n", prop.name, prop.multiProcessorCount, clk_khz / 1e3);
  const int grid = prop.multiProcessorCount;
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  const uint32_t tile_bytes = 96 * 1024, pitch = 48 * 64;
  const char* names[3] = {"LDS.128 x2 only", "LDS.128 x2 + 8 HFMA2 (fp16 blend)", "LDS.128 x2 + bf16 unpack + FFMA (round-1 blend)"};
  for (int mode = 0; mode < 3; ++mode) {
    auto k = mode == 0 ? lds_gather<0> : mode == 1 ? lds_gather<1> : lds_gather<2>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_bytes);
    const int iters = 4096;
    k<<<grid, 512, tile_bytes>>>(buf, tile_bytes, pitch, 64, out);
    cudaEventRecord(s);
    k<<<grid, 512, tile_bytes>>>(buf, tile_bytes, pitch, iters, out);
    cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e);
    const double bytes = static_cast<double>(grid) * 16 * iters * 1024.0;
    printf("(a) %-52s %8.1f GB/s  = %6.1f B/clk/SM @1.965 GHz  (%.3f ms)\n", names[mode], bytes / (ms * 1e-3) / 1e9,
           bytes / (ms * 1e-3) / grid / 1.965e9, ms);
  }
  // (b) staging
  const uint32_t shapes[4][2] = {{40, 2560}, {24, 1536}, {64, 1024}, {32, 3072}};
  for (int i = 0; i < 4; ++i) {
    const uint32_t rows = shapes[i][0], rb = shapes[i][1];
    const uint32_t smem = 2 * rows * rb;
    cudaFuncSetAttribute(stage_tiles, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int tiles = 400;
    stage_tiles<<<grid, 128, smem>>>(buf, 36ull << 20, rows, rb, 240 * 64, 16, out);
    cudaEventRecord(s);
    stage_tiles<<<grid, 128, smem>>>(buf, 36ull << 20, rows, rb, 240 * 64, tiles, out);
    cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e);
    const double bytes = static_cast<double>(grid) * tiles * rows * rb;
    printf("(b) cp.async.bulk staging, tile %2u rows x %4u B (%3u KB), 2 in flight: %8.1f GB/s = %5.1f B/clk/SM, %.2f us/tile\n",
           rows, rb, rows * rb / 1024, bytes / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / grid / 1.965e9, ms * 1e3 / tiles);
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
