// Micro-benchmark 2 (round 1): (a) cost of warp-level gather loads when the working set is
// L1-RESIDENT (the fused gather runs at 73 % L1 hit rate, ubench_gather.cu measured the miss
// path), per access pattern a candidate value-map layout would produce; (b) issue rate of the
// legacy mma.sync.m16n8k16 bf16 pipe, which a tensor-pipe bilinear blend would use.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench_l1_mma tools/ubench_l1_mma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// MODE 0: LDG.128, 4 lanes per 64 B segment, 8 random segments per warp-load (current kernel)
// MODE 1: LDG.128, 8 lanes per aligned 128 B line, 4 random lines
// MODE 2: LDG.32, 32 lanes in ONE aligned 128 B line
// MODE 3: LDG.128, 8 lanes per 128 B chunk at a random 64 B boundary, 4 chunks (x-pair, head-major)
// MODE 4: LDG.64, 16 lanes per aligned 128 B line, 2 random lines
// MODE 5: LDG.128, 16 lanes per aligned 256 B block, 2 random blocks (2x2 quad contiguous)
// MODE 6: LDG.128, 32 lanes contiguous aligned 512 B
// MODE 7: LDG.64, 8 lanes per 64 B segment, 4 random segments
template <int MODE>
__global__ void __launch_bounds__(512, 1) gather_kernel(const uint8_t* __restrict__ base, uint32_t window_lines,
                                                         int iters, float* out) {
  const int lane = threadIdx.x & 31;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float acc = 0.f;
  const int grp = MODE == 0 ? lane >> 2 : (MODE == 1 || MODE == 3) ? lane >> 3 : MODE == 4 ? lane >> 4
                : MODE == 5 ? lane >> 4 : MODE == 7 ? lane >> 3 : 0;
  uint32_t seed = (wid * 9781u + 12345u) ^ (grp * 0x9E3779B9u);
  const uint32_t mask = window_lines - 1;               // window per CTA (shared by its 16 warps)
  const uint8_t* region = base + (static_cast<uint64_t>(blockIdx.x) * window_lines << 7);
  const uint32_t lane_off = MODE == 0 ? (lane & 3) * 16 : (MODE == 1 || MODE == 3) ? (lane & 7) * 16
                          : MODE == 2 ? lane * 4 : MODE == 4 ? (lane & 15) * 8 : MODE == 5 ? (lane & 15) * 16
                          : MODE == 6 ? lane * 16 : (lane & 7) * 8;
  const uint8_t* b2 = region + lane_off;
#pragma unroll 8
  for (int it = 0; it < iters; ++it) {
    seed = seed * 1664525u + 1013904223u;
    const uint32_t r = seed >> 10;
    const uint8_t* p;
    if (MODE == 0 || MODE == 3 || MODE == 7) p = b2 + (static_cast<uint64_t>((r >> 1) & mask) << 7) + ((r & 1u) << 6);
    else if (MODE == 5) p = b2 + (static_cast<uint64_t>(r & mask & ~1u) << 7);
    else if (MODE == 6) p = b2 + (static_cast<uint64_t>(r & mask & ~3u) << 7);
    else p = b2 + (static_cast<uint64_t>(r & mask) << 7);
    if (MODE == 2) acc += __uint_as_float(__ldg(reinterpret_cast<const uint32_t*>(p)));
    else if (MODE == 4 || MODE == 7) { uint2 v = __ldg(reinterpret_cast<const uint2*>(p)); acc += __uint_as_float(v.x) + __uint_as_float(v.y); }
    else { uint4 v = __ldg(reinterpret_cast<const uint4*>(p)); acc += __uint_as_float(v.x) + __uint_as_float(v.w); }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int MODE> void run(const uint8_t* buf, float* out, int bytes_per_warp_load, const char* name) {
  const uint32_t windows[3] = {256u, 1024u, 8192u};     // 32 KB, 128 KB (L1-resident), 1 MB per SM (L2)
  for (int w = 0; w < 3; ++w) {
    const int iters = 8192, grid = 148, block = 512;
    gather_kernel<MODE><<<grid, block>>>(buf, windows[w], 256, out);
    cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
    cudaEventRecord(s);
    gather_kernel<MODE><<<grid, block>>>(buf, windows[w], iters, out);
    cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e);
    double loads = (double)grid * (block / 32) * iters;
    double gbs = loads * bytes_per_warp_load / (ms * 1e-3) / 1e9;
    double ns = ms * 1e6 / (loads / grid);
    printf("%-52s window %5u KB/SM  %8.1f GB/s  %6.2f ns/warp-load/SM (%.1f clk @1.9GHz)\n", name,
           windows[w] / 8, gbs, ns, ns * 1.9);
  }
}

// ---- mma.sync m16n8k16 bf16 rate: NACC independent accumulator chains per warp
template <int NACC>
__global__ void __launch_bounds__(1024, 1) mma_kernel(int iters, float* out, uint32_t a_seed) {
  uint32_t a[4] = {a_seed, a_seed + 1, a_seed + 2, a_seed + 3};
  uint32_t b[2] = {a_seed * 3, a_seed * 5};
  float d[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (s == 123.456f) out[0] = s;
}

template <int NACC> void run_mma(float* out, int warps) {
  const int iters = 4096, grid = 148;
  mma_kernel<NACC><<<grid, warps * 32>>>(64, out, 0);
  cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
  cudaEventRecord(s);
  mma_kernel<NACC><<<grid, warps * 32>>>(iters, out, 0);
  cudaEventRecord(e); cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  double mmas_per_smsp = (double)iters * NACC * warps / 4.0;
  double ns = ms * 1e6 / mmas_per_smsp;
  double tflops = (double)iters * NACC * warps * grid * 4096.0 / (ms * 1e-3) / 1e12;
  printf("mma.sync m16n8k16 bf16  warps/SM %2d  chains/warp %d : %6.2f ns/MMA/SMSP (%.1f clk @1.9GHz)  %.0f TFLOP/s\n",
         warps, NACC, ns, ns * 1.9, tflops);
}

int main() {
  const size_t bytes = (size_t)148 * 8192 * 128 + 4096;
  uint8_t* buf; float* out;
  cudaMalloc(&buf, bytes); cudaMalloc(&out, 4);
  cudaMemset(buf, 0, bytes);
  run<0>(buf, out, 512, "0 LDG.128 8 x 64B segments");
  run<1>(buf, out, 512, "1 LDG.128 4 x aligned 128B lines");
  run<3>(buf, out, 512, "3 LDG.128 4 x 128B at random 64B boundary");
  run<5>(buf, out, 512, "5 LDG.128 2 x aligned 256B blocks");
  run<6>(buf, out, 512, "6 LDG.128 contiguous aligned 512B");
  run<4>(buf, out, 256, "4 LDG.64 2 x aligned 128B lines");
  run<7>(buf, out, 256, "7 LDG.64 4 x 64B segments");
  run<2>(buf, out, 128, "2 LDG.32 one aligned 128B line");
  run_mma<4>(out, 4); run_mma<4>(out, 8); run_mma<4>(out, 16); run_mma<8>(out, 16); run_mma<2>(out, 32);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
