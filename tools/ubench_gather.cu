// Micro-benchmark: L1/L2 gather throughput on B200 for the access patterns candidate layouts of
// the value map would produce.  Footprint per SM-resident set is random texels of a 20 MB map
// (L2 resident).  Reports useful GB/s and "cycles per warp-load".
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_gather tools/ubench_gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// MODE 0: LDG.128, 4 lanes per 64 B segment, 8 random segments per warp-load (current kernel)
// MODE 1: LDG.128, 8 lanes per aligned 128 B line, 4 random lines per warp-load
// MODE 2: LDG.32, 32 lanes in ONE aligned 128 B line
// MODE 3: LDG.32, 32 lanes reading 128 B that start at a random 64 B boundary (straddles 50 %)
// MODE 4: LDG.64, 16 lanes per aligned 128 B line, 2 random lines per warp-load
// MODE 5: LDG.64, 8 lanes per 64 B segment, 4 random segments per warp-load
// MODE 6: LDG.128, 32 lanes fully contiguous 512 B (aligned)
template <int MODE>
__global__ void __launch_bounds__(512, 1) gather_kernel(const uint8_t* __restrict__ base, uint32_t line_mask,
                                                         int iters, float* out, int locality) {
  const int lane = threadIdx.x & 31;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float acc = 0.f;
  // cheap address generation (2-3 integer ops per load) so that the LSU/L1 is what is measured
  const int grp = MODE == 0 ? lane >> 2 : MODE == 1 ? lane >> 3 : MODE == 4 ? lane >> 4 : MODE == 5 ? lane >> 3 : 0;
  uint32_t seed = (wid * 9781u + 12345u) ^ (grp * 0x9E3779B9u);
  const uint32_t mask = locality ? 0xFFFu : line_mask;          // 4096 lines = 512 KB window
  const uint32_t region = locality ? ((wid * 2654435761u) & line_mask & ~0xFFFu) : 0u;
  const uint32_t lane_off = MODE == 0 ? (lane & 3) * 16 : MODE == 1 ? (lane & 7) * 16 : MODE == 2 ? lane * 4
                          : MODE == 3 ? lane * 4 : MODE == 4 ? (lane & 15) * 8 : MODE == 5 ? (lane & 7) * 8 : lane * 16;
  const uint8_t* b2 = base + lane_off;
#pragma unroll 8
  for (int it = 0; it < iters; ++it) {
    seed = seed * 1664525u + 1013904223u;
    const uint32_t r = seed >> 10;
    if (MODE == 0 || MODE == 5 || MODE == 3) {          // 64 B granular
      const uint8_t* p = b2 + (static_cast<uint64_t>(region + ((r >> 1) & mask)) << 7) + ((r & 1u) << 6);
      if (MODE == 0) { uint4 v = __ldg(reinterpret_cast<const uint4*>(p)); acc += __uint_as_float(v.x) + __uint_as_float(v.w); }
      else if (MODE == 5) { uint2 v = __ldg(reinterpret_cast<const uint2*>(p)); acc += __uint_as_float(v.x) + __uint_as_float(v.y); }
      else acc += __uint_as_float(__ldg(reinterpret_cast<const uint32_t*>(p)));
    } else {                                             // 128 B line granular
      const uint8_t* p = b2 + (static_cast<uint64_t>(region + (r & mask)) << 7);
      if (MODE == 1 || MODE == 6) { uint4 v = __ldg(reinterpret_cast<const uint4*>(p)); acc += __uint_as_float(v.x) + __uint_as_float(v.w); }
      else if (MODE == 4) { uint2 v = __ldg(reinterpret_cast<const uint2*>(p)); acc += __uint_as_float(v.x) + __uint_as_float(v.y); }
      else acc += __uint_as_float(__ldg(reinterpret_cast<const uint32_t*>(p)));
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int MODE> void run(const uint8_t* buf, uint32_t n_lines, float* out, int bytes_per_warp_load, const char* name) {
  for (int locality = 0; locality < 2; ++locality) {
    const int iters = 4096, grid = 148, block = 512;
    gather_kernel<MODE><<<grid, block>>>(buf, n_lines - 1, 64, out, locality);
    cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
    cudaEventRecord(s);
    gather_kernel<MODE><<<grid, block>>>(buf, n_lines - 1, iters, out, locality);
    cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e);
    double loads = (double)grid * (block / 32) * iters;
    double gbs = loads * bytes_per_warp_load / (ms * 1e-3) / 1e9;
    double ns_per_load_per_sm = ms * 1e6 / (loads / grid);
    printf("%-58s %s  %8.1f GB/s useful  %6.2f ns/warp-load/SM (%.1f clk @1.9GHz)\n", name,
           locality ? "local 512KB/warp" : "random 16MB     ", gbs, ns_per_load_per_sm, ns_per_load_per_sm * 1.9);
  }
}

int main() {
  const uint32_t n_lines = 1u << 17;   // 16 MB
  uint8_t* buf; float* out;
  cudaMalloc(&buf, (size_t)n_lines * 128 + 4096); cudaMalloc(&out, 4);
  cudaMemset(buf, 0, (size_t)n_lines * 128 + 4096);
  run<0>(buf, n_lines, out, 512, "0 LDG.128 8 x 64B segments/warp (current)");
  run<1>(buf, n_lines, out, 512, "1 LDG.128 4 x aligned 128B lines/warp");
  run<2>(buf, n_lines, out, 128, "2 LDG.32 one aligned 128B line/warp");
  run<3>(buf, n_lines, out, 128, "3 LDG.32 128B at random 64B boundary");
  run<4>(buf, n_lines, out, 256, "4 LDG.64 2 x aligned 128B lines/warp");
  run<5>(buf, n_lines, out, 256, "5 LDG.64 4 x 64B segments/warp");
  run<6>(buf, n_lines, out, 512, "6 LDG.128 contiguous aligned 512B/warp");
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
