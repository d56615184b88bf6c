// Pyramid hand-off: Lv NCHW feature maps -> one channels-last bf16 matrix (rows, S, C).
//
// Replaces `torch.cat([src.flatten(2) for src in src_views], -1).permute(0, 2, 1)`
// (lib/models/ops/modules/projattn.py:160), which the reference materialises once per view
// per layer in fp32.  Here it runs once per decoder call; every later kernel (tcgen05 value
// projection, gathers) wants the 256 channels of a position contiguous (512 B).
// HBM-bound: reads rows*S*C*sizeof(src) and writes rows*S*C*2 bytes, both fully coalesced
// (32-position x 64-channel tiles transposed through padded shared memory).
#include "common.cuh"

namespace mvg {

struct PyramidParams {
  const void* src[MVG_MAX_LEVELS];
  int hw[MVG_MAX_LEVELS];
  int start[MVG_MAX_LEVELS];
  int tile_begin[MVG_MAX_LEVELS + 1];  // cumulative 32-position tiles per level
  int num_levels;
  int channels;
  int spatial_size;
};

constexpr int kTP = 32;  // positions per tile
constexpr int kTC = 64;  // channels per tile

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}

template <typename T>
__global__ void __launch_bounds__(256)
pyramid_to_cl_kernel(PyramidParams p, __nv_bfloat16* __restrict__ dst) {
  __shared__ float tile[kTC][kTP + 1];
  int l = 0;
#pragma unroll
  for (int i = 1; i < MVG_MAX_LEVELS; ++i)
    if (i < p.num_levels && static_cast<int>(blockIdx.x) >= p.tile_begin[i]) l = i;
  const int pos0 = (blockIdx.x - p.tile_begin[l]) * kTP;
  const int c0 = blockIdx.y * kTC;
  const int row = blockIdx.z;
  const int hw = p.hw[l];
  const T* src = static_cast<const T*>(p.src[l]) +
                 (static_cast<int64_t>(row) * p.channels + c0) * hw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pos = pos0 + lane;
#pragma unroll
  for (int i = 0; i < kTC / 8; ++i) {
    const int c = warp + 8 * i;
    tile[c][lane] = (pos < hw) ? to_f32<T>(src[static_cast<int64_t>(c) * hw + pos]) : 0.f;
  }
  __syncthreads();
  __nv_bfloat16* out = dst + (static_cast<int64_t>(row) * p.spatial_size + p.start[l]) * p.channels + c0;
#pragma unroll
  for (int i = 0; i < kTP / 8; ++i) {
    const int pp = warp + 8 * i;
    if (pos0 + pp < hw) {
      const uint32_t v = pack_bf16x2(tile[2 * lane][pp], tile[2 * lane + 1][pp]);
      *reinterpret_cast<uint32_t*>(out + static_cast<int64_t>(pos0 + pp) * p.channels + 2 * lane) = v;
    }
  }
}

// bf16 -> bf16: 64 channels x 64 positions per CTA.  Loads: one 16-byte vector = 8 positions of a
// channel row (a warp reads 4 full 128-byte rows); shared tile [64 ch][64 pos + 2] bf16 (odd word
// stride); stores: one 16-byte vector = 8 channels of a position, assembled from 8 conflict-free
// 2-byte shared loads.  Requires H*W % 8 == 0 for the level (else the generic kernel runs).
__global__ void __launch_bounds__(256)
pyramid_to_cl_bf16_kernel(PyramidParams p, __nv_bfloat16* __restrict__ dst) {
  constexpr int TP = 64, TCH = 64, LDS_ = TP + 2;
  __shared__ __align__(16) __nv_bfloat16 tile[TCH * LDS_];
  int l = 0;
#pragma unroll
  for (int i = 1; i < MVG_MAX_LEVELS; ++i)
    if (i < p.num_levels && static_cast<int>(blockIdx.x) >= p.tile_begin[i]) l = i;
  const int pos0 = (blockIdx.x - p.tile_begin[l]) * TP;
  const int c0 = blockIdx.y * TCH;
  const int row = blockIdx.z;
  const int hw = p.hw[l];
  const __nv_bfloat16* src = static_cast<const __nv_bfloat16*>(p.src[l]) +
                             (static_cast<int64_t>(row) * p.channels + c0) * hw;
  const int t = threadIdx.x;
  // load: 64 rows x 8 vectors = 512 vectors, 2 per thread
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int v = t + 256 * i;
    const int c = v >> 3, pv = (v & 7) * 8;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (pos0 + pv < hw) q = ldg_nc_v4(src + static_cast<int64_t>(c) * hw + pos0 + pv);
    uint32_t* d = reinterpret_cast<uint32_t*>(&tile[c * LDS_ + pv]);   // 4-byte aligned (LDS_ even)
    d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
  }
  __syncthreads();
  __nv_bfloat16* out = dst + (static_cast<int64_t>(row) * p.spatial_size + p.start[l]) * p.channels + c0;
  // store: 64 positions x 8 channel-vectors = 512 vectors, 2 per thread; a warp covers 32
  // consecutive positions of one channel group
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int v = t + 256 * i;
    const int g = v >> 6, pp = v & 63;          // channel group (8 channels), position
    if (pos0 + pp < hw) {
      const unsigned short* tp = reinterpret_cast<const unsigned short*>(tile) + (g * 8) * LDS_ + pp;
      uint4 o;
      o.x = tp[0] | (static_cast<uint32_t>(tp[LDS_]) << 16);
      o.y = tp[2 * LDS_] | (static_cast<uint32_t>(tp[3 * LDS_]) << 16);
      o.z = tp[4 * LDS_] | (static_cast<uint32_t>(tp[5 * LDS_]) << 16);
      o.w = tp[6 * LDS_] | (static_cast<uint32_t>(tp[7 * LDS_]) << 16);
      *reinterpret_cast<uint4*>(out + static_cast<int64_t>(pos0 + pp) * p.channels + g * 8) = o;
    }
  }
}

}  // namespace mvg

extern "C" int mvg_pyramid_to_channels_last(const void* const* src_levels, int src_dtype,
                                            int num_levels, const int* level_hw, int rows,
                                            int channels, void* dst_bf16, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(src_levels && level_hw && dst_bf16, "mvg_pyramid_to_channels_last: null pointer");
  MVG_REQUIRE(num_levels >= 1 && num_levels <= MVG_MAX_LEVELS, "num_levels %d out of range", num_levels);
  MVG_REQUIRE(channels % kTC == 0 && rows > 0, "channels must be a multiple of %d", kTC);
  PyramidParams p{};
  p.num_levels = num_levels;
  p.channels = channels;
  int start = 0, tiles = 0;
  for (int l = 0; l < num_levels; ++l) {
    MVG_REQUIRE(src_levels[l] != nullptr && level_hw[l] > 0, "level %d: bad pointer/size", l);
    p.src[l] = src_levels[l];
    p.hw[l] = level_hw[l];
    p.start[l] = start;
    p.tile_begin[l] = tiles;
    start += level_hw[l];
    tiles += (level_hw[l] + kTP - 1) / kTP;
  }
  p.tile_begin[num_levels] = tiles;
  p.spatial_size = start;
  dim3 grid(tiles, channels / kTC, rows);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool fast = src_dtype == MVG_BF16;
  for (int l = 0; l < num_levels; ++l)
    fast = fast && (level_hw[l] % 8 == 0) && ((reinterpret_cast<uintptr_t>(src_levels[l]) & 15) == 0);
  if (fast) {
    PyramidParams q = p;
    int t64 = 0;
    for (int l = 0; l < num_levels; ++l) {
      q.tile_begin[l] = t64;
      t64 += (level_hw[l] + 63) / 64;
    }
    q.tile_begin[num_levels] = t64;
    dim3 g64(t64, channels / 64, rows);
    pyramid_to_cl_bf16_kernel<<<g64, 256, 0, st>>>(q, static_cast<__nv_bfloat16*>(dst_bf16));
    return check_launch("mvg_pyramid_to_channels_last");
  }
  if (src_dtype == MVG_F32)
    pyramid_to_cl_kernel<float><<<grid, 256, 0, st>>>(p, static_cast<__nv_bfloat16*>(dst_bf16));
  else if (src_dtype == MVG_BF16)
    pyramid_to_cl_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(p, static_cast<__nv_bfloat16*>(dst_bf16));
  else {
    set_error("mvg_pyramid_to_channels_last: unsupported dtype %d", src_dtype);
    return MVG_EUNSUPPORTED;
  }
  return check_launch("mvg_pyramid_to_channels_last");
}
