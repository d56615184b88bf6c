// Fused projection + projective-attention sampling, round-2 design: spatially binned items,
// TMA-staged shared-memory tiles, fp16 HFMA2 blend.
//
//   a3  project_ref_points      lib/models/dq_decoder.py:331-397, lib/utils/cameras.py:167-207,
//                               lib/utils/transforms.py:135-141
//   a4  ProjAttn (non-GEMM part) lib/models/ops/modules/projattn.py:139-153 (ref-point feature
//                               lookup), :180-191 (offsets, softmax over Lv*P, locations, and the
//                               `.view` layout scramble of the per-level Linear outputs)
//   a5  deformable gather       lib/models/ops/src/cuda/deform_im2col_cuda.cuh:247-309, :41-93
//
// Why tiles.  Every in-view (view, point) item gathers 8 heads x 24 samples x 4 corners x 64 B =
// 49 KB; at Q = 1024 that is 3.8 GB per launch against 0.23 GB of compulsory HBM bytes.  A
// warp-level LDG.128 delivers at most 64 B/clk/SM even on L1 hits (profiles/ubench_l1_mma_r1.txt),
// LDS.128 from shared memory 119 B/clk/SM with the fp16 blend attached
// (profiles/ubench_smem_gather_r2.txt), and the round-1 bf16 -> fp32 unpack + FFMA blend was
// ALU-bound at 79 B/clk/SM whatever the source.  So: (1) the value / offset-logit maps are fp16
// (written by the tcgen05 GEMM epilogue; 8x finer than bf16) and the bilinear x attention blend
// runs in packed HFMA2 - 4 instructions per 16-byte load instead of 12 - with fp32 accumulation
// across the pyramid levels; (2) the gathered rows come from shared-memory tiles staged by
// tensor-map TMA (cp.async.bulk.tensor.5d, UTMALDG.5D; the sample records by cp.async.bulk, UBLKCP).
//
// Pipeline of one call (all on `stream`, no host sync):
//   project_bin_kernel   one thread per (frame, view, point): projection in non-contracted fp32
//                        in the reference's op order (`bounding` bit-exact), ref2d / bounding
//                        outputs, zero `sampled` rows for out-of-view points (the reference
//                        multiplies their attention feature by 0, dq_decoder.py:585-586, and
//                        reads it nowhere else), and a KEY = (frame-view, 24 x 24 level-0 texel
//                        cell of the reference point) with its rank inside the key
//                        (warp-aggregated atomics).
//   bin_scan_kernel      one block: key offsets, and CHUNKS of <= 128 same-key items.
//   bin_scatter_kernel   counting-sort scatter -> item list ordered by key.
//   sample_params_kernel work queue over (chunk, 32 items), one warp per item: phase A samples the
//                        pre-projected 192-channel map G at the reference point (+ qproj), phase B
//                        does the 24-way softmax and, per (head, sample), the clamped 2x2 texel block
//                        and its four bilinear x attention weights (fp16) -> 16-byte records in the
//                        workspace, plus the bounding box of every (chunk, head, level)'s samples
//                        (atomic min / max into the chunk's boxes).
//   gather_tiles_kernel  persistent, one CTA per SM: a producer warp stages, per (chunk, head),
//                        the bounding-box tile of each level (a box of the head-major value
//                        tensor: cp.async.bulk.tensor.5d in 8-row + 1-row boxes at one of three row
//                        pitches, mbarrier complete_tx; one cp.async.bulk per tile row with
//                        MVG_TILE_TMA=0) into a per-level shared-memory region; 16 consumer warps gather level by level
//                        (LDS.128: a quarter-warp reads the two horizontal corners of one head =
//                        128 contiguous bytes, conflict-free), so the level-l region is re-filled
//                        for the next unit while levels l+1.. of this unit are still being read.
//   gather_direct_kernel same arithmetic with LDG from global memory, for the (chunk, head) units
//                        whose boxes do not fit the regions (never on the shipped configurations;
//                        keeps the operator correct for arbitrary offsets).
// Results do not depend on the (non-deterministic) order of items inside a key: every item's
// arithmetic is a fixed sequence over its own records.
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "tcgen05.cuh"

namespace mvg {

constexpr int kQP = 192;           // 128 offset channels + 64 logit channels per level
constexpr int kHeads = 8;
constexpr int kPcThreads = 256;    // project_bin block
#ifndef MVG_CORE
#define MVG_CORE 20
#endif
constexpr int kCore = MVG_CORE;    // key cell edge, level-0 texels
constexpr int kChunk = 128;        // items per chunk (upper bound)
#ifndef MVG_P_WARPS
#define MVG_P_WARPS 8
#endif
#ifndef MVG_P_MINBLK
#define MVG_P_MINBLK 2
#endif
constexpr int kPWarps = MVG_P_WARPS;   // sample_params: warps per CTA
#ifndef MVG_G_WARPS
#define MVG_G_WARPS 16
#endif
constexpr int kGWarps = MVG_G_WARPS;   // gather_tiles: consumer warps per CTA (+ 1 producer warp)
constexpr int kScanThreads = 1024;
constexpr int kIPW = (kChunk + kGWarps - 1) / kGWarps;   // items per consumer warp and unit
constexpr int kPart = 32;          // sample_params work item: kPart items of one chunk
constexpr int kParts = kChunk / kPart;
constexpr int kRecBytes = 128;     // records of one (item, head, level): 2 block columns x 8 points x 8 B

// per-level shared-memory region capacity in texels (64 B each) for the tiled gather
// (227 KB - LV x 16 KB record regions)
template <int LV>
__host__ __device__ constexpr int region_cap(int l) {
  return LV == 1 ? (l == 0 ? 3328 : 0)
       : LV == 2 ? (l == 0 ? 2048 : l == 1 ? 1024 : 0)
       : LV == 3 ? (l == 0 ? 1536 : l == 1 ? 768 : l == 2 ? 544 : 0)
                 : (l == 0 ? 1280 : l == 1 ? 672 : l == 2 ? 416 : l == 3 ? 224 : 0);
}

// Tile staging by tensor map (cp.async.bulk.tensor.5d): a tensor map fixes its box, so a tile is stored with one of
// three row pitches per level (p0 = the smallest even p with p * p >= capacity, p0 - 4, p0 - 8 texels) and loaded
// as floor(rows / 8) boxes of 8 rows + (rows % 8) boxes of one row: <= 11 TMA ops per tile instead of one bulk copy
// per tile row (~38 at level 0; the producer warp needed ~2.8 us per unit for ~70 copies at ~30 ns each and was the
// limiter for small chunks - the whole kernel for the 128-query ranks of an 8-GPU run).
constexpr int kPitchClasses = 3;
constexpr int kPitchStep = 4;      // 6 sent ~7x more units to gather_direct (31 vs 4 us): a 20 x 25 level-2 box needs
                                   // pitch <= 21 to fit its 544-texel region
template <int LV>
__host__ __device__ constexpr int tile_pitch(int l, int c) {
  int p = 2;
  while (p * p < region_cap<LV>(l)) p += 2;
  return p - kPitchStep * c;
}
struct TileMaps {
  CUtensorMap m[MVG_MAX_LEVELS][kPitchClasses][2];   // [level][pitch class][0: 8-row box, 1: 1-row box]
  int enabled;                                       // 0: one cp.async.bulk per tile row (MVG_TILE_TMA=0, or encode failed)
};

struct GatherWs {
  int* counts;        // [BV]            in-view items per (frame, view)       (zeroed per call)
  int* hist;          // [keys]          items per key                          (zeroed per call)
  int* ctrs;          // [8]             0 chunks, 1 in-view items, 2 unit cursor, 3 direct units,
                      //                 5 work-item cursor of sample_params (zeroed)
  int* key_off;       // [keys]
  int* key_chunk0;    // [keys + 1]
  int* item_key;      // [items]
  int* item_rank;     // [items]
  int* sorted;        // [items]
  int4* chunks;       // [max_chunks]    {first, count, frame-view, key}
  int4* bbox;         // [max_chunks * 8 * LV]   {65535 - min x0, 65535 - min y0, max x0 + 1, max y0 + 1} of the
                      //                 2x2 sample blocks: zeroed with the counters, reduced with atomicMax
  int* direct_list;   // [max_chunks * 8]  units whose boxes exceed the shared-memory regions (gather_tiles -> gather_direct)
  uint2* params;      // [8 heads][LV][items][2 block columns][8 slots]   {x0 | y0 << 16, half2(w_top, w_bottom)}
  int keys, kx, ky, max_chunks;
  int64_t items;
  int64_t zero_bytes; // counts | hist | ctrs | bbox: one memset per call
};

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// Lays the workspace out; returns the total size in bytes (base may be null for the size query).
static int64_t make_ws(const MvgSampleParams& p, void* base, GatherWs* w) {
  const int64_t BV = static_cast<int64_t>(p.batch) * p.views;
  const int64_t items = BV * p.points;
  const int kx = (p.level_w[0] + kCore - 1) / kCore, ky = (p.level_h[0] + kCore - 1) / kCore;
  const int64_t keys = BV * kx * ky;
  const int64_t max_chunks = items / kChunk + keys + 1;
  uint8_t* b = static_cast<uint8_t*>(base);
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off = align_up(off + bytes, 256); return b ? b + o : nullptr; };
  w->counts = reinterpret_cast<int*>(take(4 * (BV + keys + 8)));        // counts | hist | ctrs: one memset
  w->hist = w->counts ? w->counts + BV : nullptr;
  w->ctrs = w->counts ? w->hist + keys : nullptr;
  w->bbox = reinterpret_cast<int4*>(take(16 * max_chunks * kHeads * p.num_levels));
  w->zero_bytes = off;
  w->key_off = reinterpret_cast<int*>(take(4 * keys));
  w->key_chunk0 = reinterpret_cast<int*>(take(4 * (keys + 1)));
  w->item_key = reinterpret_cast<int*>(take(4 * items));
  w->item_rank = reinterpret_cast<int*>(take(4 * items));
  w->sorted = reinterpret_cast<int*>(take(4 * items));
  w->chunks = reinterpret_cast<int4*>(take(16 * max_chunks));
  w->direct_list = reinterpret_cast<int*>(take(4 * max_chunks * kHeads));
  w->params = reinterpret_cast<uint2*>(take(8 * items * kHeads * p.num_levels * 16));
  w->keys = static_cast<int>(keys);
  w->kx = kx;
  w->ky = ky;
  w->max_chunks = static_cast<int>(max_chunks);
  w->items = items;
  return off;
}

__device__ __forceinline__ int key_of(float rx, float ry, int bv, const MvgSampleParams& prm, int kx, int ky) {
  // rx, ry: normalised [0, 1] image coordinates (any monotone cell assignment works: binning only)
  const int W0 = prm.level_w[0], H0 = prm.level_h[0];
  const int tx = min(max(static_cast<int>(rx * static_cast<float>(W0)), 0), W0 - 1) / kCore;
  const int ty = min(max(static_cast<int>(ry * static_cast<float>(H0)), 0), H0 - 1) / kCore;
  return (bv * ky + ty) * kx + tx;
}

// rank of this lane's item inside its key: one atomic per distinct key per warp
__device__ __forceinline__ int key_rank(int* hist, int key, uint32_t active) {
  const int lane = threadIdx.x & 31;
  const uint32_t peers = __match_any_sync(active, key);
  const int leader = __ffs(peers) - 1;
  int base = 0;
  if (lane == leader) base = atomicAdd(hist + key, __popc(peers));
  base = __shfl_sync(peers, base, leader);
  return base + __popc(peers & ((1u << lane) - 1u));
}

// a3: one 3D point through one packed camera -> normalised network-image coordinates + the
// `bounding` bit.  Non-contracted fp32 in the reference's op order (cameras.py:167-207,
// dq_decoder.py:374-397, transforms.py:135-141): `bounding` is bit-exact.
__device__ __forceinline__ bool project_point(const MvgCamera* cam, const float* __restrict__ x3, float img_w,
                                              float img_h, float& rx, float& ry) {
  const float dx = fsub(__ldg(x3 + 0), cam->T[0]);
  const float dy = fsub(__ldg(x3 + 1), cam->T[1]);
  const float dz = fsub(__ldg(x3 + 2), cam->T[2]);
  const float xc = fadd(fadd(fmul(cam->R[0], dx), fmul(cam->R[1], dy)), fmul(cam->R[2], dz));
  const float yc = fadd(fadd(fmul(cam->R[3], dx), fmul(cam->R[4], dy)), fmul(cam->R[5], dz));
  const float zc = fadd(fadd(fmul(cam->R[6], dx), fmul(cam->R[7], dy)), fmul(cam->R[8], dz));
  const float zden = fadd(zc, 1e-5f);
  float y0 = fdiv(xc, zden), y1 = fdiv(yc, zden);
  const float r2 = fadd(fmul(y0, y0), fmul(y1, y1));
  const float r4 = fmul(r2, r2), r6 = fmul(fmul(r2, r2), r2);
  const float radial = fadd(1.f, fadd(fadd(fmul(cam->k[0], r2), fmul(cam->k[1], r4)), fmul(cam->k[2], r6)));
  const float tanv = fadd(fmul(cam->p[0], y1), fmul(cam->p[1], y0));
  const float corr = fadd(radial, fmul(2.f, tanv));
  y0 = fadd(fmul(y0, corr), fmul(cam->p[1], r2));
  y1 = fadd(fmul(y1, corr), fmul(cam->p[0], r2));
  float px = fadd(fmul(cam->f[0], y0), cam->c[0]);
  float py = fadd(fmul(cam->f[1], y1), cam->c[1]);
  const bool inb = (px >= 0.f) && (py >= 0.f) && (px < cam->wh[0]) && (py < cam->wh[1]);
  px = fminf(fmaxf(px, -1.f), cam->clamp_max);
  py = fminf(fmaxf(py, -1.f), cam->clamp_max);
  const float ax = fadd(fadd(fmul(px, cam->aff[0]), fmul(py, cam->aff[1])), cam->aff[2]);
  const float ay = fadd(fadd(fmul(px, cam->aff[3]), fmul(py, cam->aff[4])), cam->aff[5]);
  rx = fdiv(ax, img_w);
  ry = fdiv(ay, img_h);
  return inb;
}

// The projection alone (training path, diagnostics): no binning, no gather.
__global__ void __launch_bounds__(kPcThreads)
project_points_kernel(const float* __restrict__ ref3d, const MvgCamera* __restrict__ cams, int V, int N,
                      float img_w, float img_h, float* __restrict__ ref2d_out, uint8_t* __restrict__ bounding_out) {
  const int bv = blockIdx.y;
  const int n = blockIdx.x * kPcThreads + threadIdx.x;
  if (n >= N) return;
  const int64_t item = static_cast<int64_t>(bv) * N + n;
  float rx, ry;
  const bool inb = project_point(cams + bv, ref3d + (static_cast<int64_t>(bv / V) * N + n) * 3, img_w, img_h, rx, ry);
  *reinterpret_cast<float2*>(ref2d_out + 2 * item) = make_float2(rx, ry);
  bounding_out[item] = inb ? 1 : 0;
}

// ------------------------------------------------------------------ projection + binning
// One block = 256 consecutive points of ONE (frame, view) pair bv = blockIdx.y.
__global__ void __launch_bounds__(kPcThreads)
project_bin_kernel(const float* __restrict__ ref3d, const MvgCamera* __restrict__ cams,
                   const MvgSampleParams prm, float* __restrict__ ref2d_out,
                   uint8_t* __restrict__ bounding_out, __nv_bfloat16* __restrict__ sampled,
                   const GatherWs ws) {
  pdl_enter();
  const int N = prm.points, V = prm.views;
  const int bv = blockIdx.y;
  const int n = blockIdx.x * kPcThreads + threadIdx.x;
  const int64_t item = static_cast<int64_t>(bv) * N + n;
  const int lane = threadIdx.x & 31;
  bool inb = false;
  float rx = 0.f, ry = 0.f;
  if (n < N) {
    const int b = bv / V;
    inb = project_point(cams + bv /* (B,V) row-major == item / N */, ref3d + (static_cast<int64_t>(b) * N + n) * 3,
                        prm.img_w, prm.img_h, rx, ry);
    *reinterpret_cast<float2*>(ref2d_out + 2 * item) = make_float2(rx, ry);
    bounding_out[item] = inb ? 1 : 0;
  }
  // out-of-view rows of `sampled` are defined (zeros); a warp writes one 512-byte row at a time
  uint32_t out_mask = __ballot_sync(0xffffffffu, n < N && !inb);
  const int64_t item0 = item - lane;
  while (out_mask) {
    const int src = __ffs(out_mask) - 1;
    out_mask &= out_mask - 1;
    reinterpret_cast<uint4*>(sampled + (item0 + src) * 256)[lane] = make_uint4(0u, 0u, 0u, 0u);
  }
  const uint32_t active = __ballot_sync(0xffffffffu, inb);
  if (inb) {
    const int key = key_of(rx, ry, bv, prm, ws.kx, ws.ky);
    ws.item_key[item] = key;
    ws.item_rank[item] = key_rank(ws.hist, key, active);
  } else if (n < N) {
    ws.item_key[item] = -1;
  }
}

// ProjAttn.forward entry: the per-level reference points are given, every item is gathered.
__global__ void __launch_bounds__(kPcThreads)
bin_refl_kernel(const float* __restrict__ refl_in, const MvgSampleParams prm, const GatherWs ws) {
  const int N = prm.points;
  const int bv = blockIdx.y;
  const int n = blockIdx.x * kPcThreads + threadIdx.x;
  const int64_t item = static_cast<int64_t>(bv) * N + n;
  const bool ok = n < N;
  const uint32_t active = __ballot_sync(0xffffffffu, ok);
  if (ok) {
    const float W0 = static_cast<float>(prm.level_w[0]), H0 = static_cast<float>(prm.level_h[0]);
    const float rx = __ldg(refl_in + item * prm.num_levels * 2) * (W0 - 1.f) / W0;
    const float ry = __ldg(refl_in + item * prm.num_levels * 2 + 1) * (H0 - 1.f) / H0;
    const int key = key_of(rx, ry, bv, prm, ws.kx, ws.ky);
    ws.item_key[item] = key;
    ws.item_rank[item] = key_rank(ws.hist, key, active);
  }
}

// ------------------------------------------------------------------ key offsets + chunk table
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();                       // warp_sums may still be read from the previous call
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int s = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    warp_sums[lane] = s;                 // inclusive over warps
  }
  __syncthreads();
  *total = warp_sums[kScanThreads / 32 - 1];
  return inc - v + (warp > 0 ? warp_sums[warp - 1] : 0);
}

__global__ void __launch_bounds__(kScanThreads)
bin_scan_kernel(const GatherWs ws, int cells_per_bv) {
  __shared__ int warp_sums[32];
  pdl_enter();
  int carry_items = 0, carry_chunks = 0;
  for (int base = 0; base < ws.keys; base += kScanThreads) {
    const int k = base + threadIdx.x;
    const int cnt = k < ws.keys ? ws.hist[k] : 0;
    const int nch = (cnt + kChunk - 1) / kChunk;
    int tot_i, tot_c;
    const int ex_i = block_exclusive_scan(cnt, warp_sums, &tot_i);
    const int ex_c = block_exclusive_scan(nch, warp_sums, &tot_c);
    if (k < ws.keys) {
      ws.key_off[k] = carry_items + ex_i;
      ws.key_chunk0[k] = carry_chunks + ex_c;
      if (cnt > 0) atomicAdd(ws.counts + k / cells_per_bv, cnt);
    }
    carry_items += tot_i;
    carry_chunks += tot_c;
  }
  if (threadIdx.x == 0) {
    ws.key_chunk0[ws.keys] = carry_chunks;
    ws.ctrs[0] = carry_chunks;
    ws.ctrs[1] = carry_items;
  }
  __syncthreads();
  // chunk table: chunk c belongs to the last key whose first chunk is <= c (binary search)
  for (int c = threadIdx.x; c < carry_chunks; c += kScanThreads) {
    int lo = 0, hi = ws.keys;            // key_chunk0[lo] <= c < key_chunk0[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (ws.key_chunk0[mid] <= c) lo = mid; else hi = mid;
    }
    const int j = c - ws.key_chunk0[lo];
    const int cnt = ws.hist[lo];
    const int nch = (cnt + kChunk - 1) / kChunk;
    const int sz = (cnt + nch - 1) / nch;              // the key's items split evenly over its chunks
    ws.chunks[c] = make_int4(ws.key_off[lo] + j * sz, min(sz, cnt - j * sz), lo / cells_per_bv, lo);
  }
}

__global__ void __launch_bounds__(256) bin_scatter_kernel(const GatherWs ws) {
  pdl_enter();
  const int64_t item = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (item >= ws.items) return;
  const int key = ws.item_key[item];
  if (key >= 0) ws.sorted[ws.key_off[key] + ws.item_rank[item]] = static_cast<int>(item);
}

// ------------------------------------------------------------------ per-sample parameters
__device__ __forceinline__ __half2 u32_as_half2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint32_t half2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// Per-warp scratch of sample_params_kernel: the Linear outputs of one item after the `.view` scramble,
// i.e. per head m its LV*16 flat offsets and LV*8 flat logits.  The head stride is padded to 4 banks
// (mod 32), so that phase B's lanes (head m = lane & 7, sample group lane >> 3) read conflict-free
// (the unpadded [level][192] layout cost a 4-way conflict on the offsets and 2-way on the logits).
template <int LV> struct ParamScratch {
  static constexpr int kOffN = LV * 16, kLogN = LV * 8;
  static constexpr int kOffStride = kOffN + (36 - kOffN % 32) % 32;
  static constexpr int kLogStride = kLogN + (36 - kLogN % 32) % 32;
  float off[kHeads * kOffStride];
  float logit[kHeads * kLogStride];
};

__device__ __forceinline__ void fma8_f16(float (&acc)[8], const uint4& c, float w) {
  const float2 f0 = __half22float2(u32_as_half2(c.x)), f1 = __half22float2(u32_as_half2(c.y));
  const float2 f2 = __half22float2(u32_as_half2(c.z)), f3 = __half22float2(u32_as_half2(c.w));
  acc[0] = fmaf(w, f0.x, acc[0]); acc[1] = fmaf(w, f0.y, acc[1]);
  acc[2] = fmaf(w, f1.x, acc[2]); acc[3] = fmaf(w, f1.y, acc[3]);
  acc[4] = fmaf(w, f2.x, acc[4]); acc[5] = fmaf(w, f2.y, acc[5]);
  acc[6] = fmaf(w, f3.x, acc[6]); acc[7] = fmaf(w, f3.y, acc[7]);
}

// Everything downstream of the fp16 map G (sampling offsets, softmax, bilinear weights) is not
// bit-comparable with the reference anyway, so this kernel uses contracted / approximate fp32
// (fmaf, ex2.approx, reciprocal multiplies): the location math costs a handful of instructions
// per sample instead of IEEE divisions.  The bit-exact integer path (`bounding`, selection) does
// not pass through here; mvg_deform_forward keeps the reference's exact arithmetic.
template <int LV>
__global__ void __launch_bounds__(kPWarps * 32, MVG_P_MINBLK)
sample_params_kernel(const __half* __restrict__ gmap, const float* __restrict__ qproj, const MvgSampleParams prm,
                     const float* __restrict__ ref2d, const float* __restrict__ refl_in, const GatherWs ws) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  ParamScratch<LV>* scratch = reinterpret_cast<ParamScratch<LV>*>(smem_dyn);        // [kPWarps]
  __shared__ int s_bb[kHeads][LV][4];     // x0 min, y0 min, x0 max, y0 max
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_enter();
  using PS = ParamScratch<LV>;
  PS& sc = scratch[warp];
  const int N = prm.points, V = prm.views, B = prm.batch;
  const int ldg = prm.ld_g;
  constexpr int NS = LV * 8;             // samples per head
  // phase-B ownership: head m = lane & 7, sample group sub = lane >> 3 (samples r = sub + 4 i)
  const int m = lane & 7, sub = lane >> 3;
  const int nchunks = ws.ctrs[0];
  float fWl[LV], fHl[LV], sxl[LV], syl[LV];
#pragma unroll
  for (int l = 0; l < LV; ++l) {
    fWl[l] = static_cast<float>(prm.level_w[l]);
    fHl[l] = static_cast<float>(prm.level_h[l]);
    sxl[l] = fdiv(fWl[l], fsub(fWl[l], 1.f));        // dq_decoder.py:570-573: r * W / (W - 1)
    syl[l] = fdiv(fHl[l], fsub(fHl[l], 1.f));
  }
  // record pointer of (head m, level 0, position 0), this lane's slot pair
  uint2* const rec_lane = ws.params + static_cast<int64_t>(m) * LV * ws.items * 16 + sub * 2;
  const int64_t lvl_stride = ws.items * 16;
  __shared__ int s_chunk;
  __shared__ int s_item[kPart];
  __shared__ float2 s_ref[kPart];
#pragma unroll 1
  for (;;) {
    __syncthreads();                     // previous work item's boxes were written out, s_chunk was read
    // work queue over (chunk, part of kPart items): a whole chunk per CTA is ~50 us of work and left
    // the SMs idle for a fifth of the kernel (296 CTAs, ~2.5 chunks each)
    if (threadIdx.x == 0) s_chunk = atomicAdd(ws.ctrs + 5, 1);
    __syncthreads();
    const int chunk = s_chunk / kParts, part = s_chunk % kParts;
    if (chunk >= nchunks) break;
    const int4 ch = ws.chunks[chunk];
    const int i0 = part * kPart, i1 = min(i0 + kPart, ch.y);
    if (i0 >= i1) continue;
    // ids + reference points of the work item's items: one dependent global-load chain per CTA and work
    // item instead of one per warp and item (it showed up as 8 % of the stall samples)
    if (threadIdx.x < i1 - i0) {
      const int it = ws.sorted[ch.x + i0 + threadIdx.x];
      s_item[threadIdx.x] = it;
      if (refl_in == nullptr) s_ref[threadIdx.x] = __ldg(reinterpret_cast<const float2*>(ref2d) + it);
    }
    for (int i = threadIdx.x; i < kHeads * LV * 4; i += blockDim.x)
      (&s_bb[0][0][0])[i] = (i & 2) ? INT_MIN : INT_MAX;
    __syncthreads();
    int bbx0[LV], bby0[LV], bbx1[LV], bby1[LV];
#pragma unroll
    for (int l = 0; l < LV; ++l) { bbx0[l] = bby0[l] = INT_MAX; bbx1[l] = bby1[l] = INT_MIN; }
    const int bv = ch.z;
    const int v = bv % V, b = bv / V;
    const __half* grow = gmap + static_cast<int64_t>(v * B + b) * prm.spatial_size * ldg + lane * 8;
    const float* qrow = qproj + static_cast<int64_t>(b) * N * kQP + lane * 8;
    const int item_base = bv * N;
    // (measured and dropped: a cross-item pipeline that keeps item i+1's 12 corner rows in flight during
    //  item i's phase B - 168 registers, 12 warps per SM, 145 us vs 127 us for this version; an L1
    //  prefetch of those rows, 132 vs 127 us)
#pragma unroll 1
    for (int idx = i0 + warp; idx < i1; idx += kPWarps) {
      const int pos = ch.x + idx;
      const int cur = s_item[idx - i0];
      const int n = cur - item_base;
      float refl_x[LV], refl_y[LV];
      if (refl_in != nullptr) {          // ProjAttn.forward entry: reference points are given
#pragma unroll
        for (int l = 0; l < LV; ++l) {
          refl_x[l] = __ldg(refl_in + (static_cast<int64_t>(cur) * LV + l) * 2);
          refl_y[l] = __ldg(refl_in + (static_cast<int64_t>(cur) * LV + l) * 2 + 1);
        }
      } else {
        const float2 rr = s_ref[idx - i0];
#pragma unroll
        for (int l = 0; l < LV; ++l) {
          refl_x[l] = rr.x * sxl[l];
          refl_y[l] = rr.y * syl[l];
        }
      }
      // ---------------- phase A (a4 i+iii): sample the pre-projected map G at the reference point
      if (lane < kQP / 8) {
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(qrow + n * kQP));
        const float4 q1 = __ldg(reinterpret_cast<const float4*>(qrow + n * kQP + 4));
        // one level in flight ahead of the one being blended: 8 corner rows live instead of 12
        uint4 cn[2][4];
        float cwgt[2][4];
        auto issue = [&](int l, uint4 (&c)[4], float (&w)[4]) {
          const int W = prm.level_w[l], H = prm.level_h[l];
          // F.grid_sample(bilinear, zeros, align_corners=False): projattn.py:139-153
          // ix = ((2 r - 1) + 1) * W / 2 - 0.5 = r * W - 0.5 (grid clamped to [-1.1, 1.1])
          const float gx = fminf(fmaxf(fmaf(refl_x[l], 2.f, -1.f), -1.1f), 1.1f);
          const float gy = fminf(fmaxf(fmaf(refl_y[l], 2.f, -1.f), -1.1f), 1.1f);
          const float ix = fmaf(gx + 1.f, fWl[l] * 0.5f, -0.5f);
          const float iy = fmaf(gy + 1.f, fHl[l] * 0.5f, -0.5f);
          const float fx0 = floorf(ix), fy0 = floorf(iy);
          const int x0 = static_cast<int>(fx0), yy0 = static_cast<int>(fy0);
          const float we = ix - fx0, ww = 1.f - we;
          const float wso = iy - fy0, wn = 1.f - wso;
          const bool okx0 = x0 >= 0 && x0 < W, okx1 = x0 + 1 >= 0 && x0 + 1 < W;
          const bool oky0 = yy0 >= 0 && yy0 < H, oky1 = yy0 + 1 >= 0 && yy0 + 1 < H;
          const int xa = min(max(x0, 0), W - 1), xb = min(max(x0 + 1, 0), W - 1);
          const int ya = min(max(yy0, 0), H - 1), yb = min(max(yy0 + 1, 0), H - 1);
          const __half* gl = grow + static_cast<int64_t>(prm.level_start[l]) * ldg;
          c[0] = ldg_nc_v4(gl + (ya * W + xa) * ldg);
          c[1] = ldg_nc_v4(gl + (ya * W + xb) * ldg);
          c[2] = ldg_nc_v4(gl + (yb * W + xa) * ldg);
          c[3] = ldg_nc_v4(gl + (yb * W + xb) * ldg);
          w[0] = (oky0 && okx0) ? wn * ww : 0.f;
          w[1] = (oky0 && okx1) ? wn * we : 0.f;
          w[2] = (oky1 && okx0) ? wso * ww : 0.f;
          w[3] = (oky1 && okx1) ? wso * we : 0.f;
        };
        issue(0, cn[0], cwgt[0]);
#pragma unroll
        for (int l = 0; l < LV; ++l) {
          if (l + 1 < LV) issue(l + 1, cn[(l + 1) & 1], cwgt[(l + 1) & 1]);
          float r8[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) fma8_f16(r8, cn[l & 1][c], cwgt[l & 1][c]);
          // flat index after the `.view`: offsets l*128 + 8 lane .. +7, logits l*64 + 8 (lane - 16) .. +7
          // (8 consecutive flat indices never straddle a head: LV*16 and LV*8 are multiples of 8)
          float4* dst;
          if (lane < 16) {
            const int f0 = l * 128 + lane * 8;
            dst = reinterpret_cast<float4*>(&sc.off[(f0 / PS::kOffN) * PS::kOffStride + f0 % PS::kOffN]);
          } else {
            const int g0 = l * 64 + (lane - 16) * 8;
            dst = reinterpret_cast<float4*>(&sc.logit[(g0 / PS::kLogN) * PS::kLogStride + g0 % PS::kLogN]);
          }
          // lanes 4 apart hit the same banks: they store their two halves in opposite order
          const int h0 = (lane >> 2) & 1;
          const float4 lo4 = make_float4(r8[0], r8[1], r8[2], r8[3]), hi4 = make_float4(r8[4], r8[5], r8[6], r8[7]);
          dst[h0] = h0 ? hi4 : lo4;
          dst[h0 ^ 1] = h0 ? lo4 : hi4;
        }
      }
      __syncwarp();
      // ---------------- phase B (a4 iv+v, a5 index path): per (head, sample) records.
      // 4 lanes per head, lane `sub` owns samples r = sub + 4 i (level i >> 1, point sub + 4 (i & 1)).
      {
        float lg[NS / 4];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < NS / 4; ++i) {
          lg[i] = sc.logit[m * PS::kLogStride + sub + 4 * i];      // flat logit m * NS + sub + 4 i after the `.view`
          mx = fmaxf(mx, lg[i]);
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NS / 4; ++i) {
          lg[i] = __expf(lg[i] - mx);
          sum += lg[i];
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 8);
        sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        const float inv_sum = __fdividef(1.f, sum);
        uint2* rec_item = rec_lane + static_cast<int64_t>(pos) * 16;
#pragma unroll
        for (int l = 0; l < LV; ++l) {
          const int W = prm.level_w[l], H = prm.level_h[l];
          const float fW = fWl[l], fH = fHl[l];
          // projattn.py:186-191 (loc = ref + off / (W, H)), then deform_im2col_cuda.cuh:291-301
          // (im = loc * size - 0.5): im = ref * size + off - 0.5
          const float bx = fmaf(refl_x[l], fW, -0.5f), by = fmaf(refl_y[l], fH, -0.5f);
          // a sample the reference skips keeps weight 0 and is parked on the reference point's
          // texel, so that it does not stretch the tile box
          const int park_x = static_cast<int>(fminf(fmaxf(bx, 0.f), fW - 2.f));
          const int park_y = static_cast<int>(fminf(fmaxf(by, 0.f), fH - 2.f));
          int lx0 = INT_MAX, ly0 = INT_MAX, lx1 = INT_MIN, ly1 = INT_MIN;
          uint32_t rxy[2], rw0[2], rw1[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int i = 2 * l + e;                     // sample r = sub + 4 i of this head
            const int r = sub + 4 * i;
            const float wgt = lg[i] * inv_sum;
            // flat offset index after the `.view`: m * (NS * 2) + 2 r
            const float2 off = *reinterpret_cast<const float2*>(&sc.off[m * PS::kOffStride + 2 * r]);
            const float w_im = bx + off.x, h_im = by + off.y;
            const bool inside = h_im > -1.f && w_im > -1.f && h_im < fH && w_im < fW;
            const float fh = floorf(h_im), fw = floorf(w_im);
            const int h_low = static_cast<int>(fh), w_low = static_cast<int>(fw);
            const float lh = h_im - fh, lw = w_im - fw;
            const float hh = 1.f - lh, hw = 1.f - lw;
            // The 2x2 texel block is clamped into the level ((ha, wa) .. (ha+1, wa+1) always exist,
            // H, W >= 2); a corner the reference skips (deform_im2col_cuda.cuh:57-80) gets weight 0
            // and the surviving row / column moves to the block row / column that holds its texel.
            const int ha = inside ? min(max(h_low, 0), H - 2) : park_y;
            const int wa = inside ? min(max(w_low, 0), W - 2) : park_x;
            const float sw = inside ? wgt : 0.f;
            const float ry0 = (h_low < 0 ? lh : (h_low > H - 2 ? 0.f : hh)) * sw;
            const float ry1 = (h_low < 0 ? 0.f : (h_low > H - 2 ? hh : lh)) * sw;
            const float rx0 = w_low < 0 ? lw : (w_low > W - 2 ? 0.f : hw);
            const float rx1 = w_low < 0 ? 0.f : (w_low > W - 2 ? hw : lw);
            lx0 = min(lx0, wa); lx1 = max(lx1, wa);
            ly0 = min(ly0, ha); ly1 = max(ly1, ha);
            rxy[e] = static_cast<uint32_t>(wa) | (static_cast<uint32_t>(ha) << 16);
            rw0[e] = pack_f16x2(ry0 * rx0, ry1 * rx0);      // left block column: (top, bottom)
            rw1[e] = pack_f16x2(ry0 * rx1, ry1 * rx1);      // right block column
          }
          // slot order inside a block column: points (q, q + 4) adjacent, so that this lane writes - and the
          // gather lane of quarter q reads - both records with one 16-byte access (whole 32-byte sectors
          // per head: the 8-byte scattered stores of the first version cost a third of the kernel)
          uint4* rec = reinterpret_cast<uint4*>(rec_item + l * lvl_stride);
          rec[0] = make_uint4(rxy[0], rw0[0], rxy[1], rw0[1]);
          rec[4] = make_uint4(rxy[0], rw1[0], rxy[1], rw1[1]);
          bbx0[l] = min(bbx0[l], lx0); bbx1[l] = max(bbx1[l], lx1);
          bby0[l] = min(bby0[l], ly0); bby1[l] = max(bby1[l], ly1);
        }
      }
      __syncwarp();   // scratch is reused by the next item
    }
    // chunk-wide bounding boxes of the (head, level) sample blocks
#pragma unroll
    for (int l = 0; l < LV; ++l) {
#pragma unroll
      for (int o = 8; o <= 16; o <<= 1) {
        bbx0[l] = min(bbx0[l], __shfl_xor_sync(0xffffffffu, bbx0[l], o));
        bby0[l] = min(bby0[l], __shfl_xor_sync(0xffffffffu, bby0[l], o));
        bbx1[l] = max(bbx1[l], __shfl_xor_sync(0xffffffffu, bbx1[l], o));
        bby1[l] = max(bby1[l], __shfl_xor_sync(0xffffffffu, bby1[l], o));
      }
      if (sub == 0 && bbx0[l] != INT_MAX) {
        atomicMin(&s_bb[m][l][0], bbx0[l]); atomicMin(&s_bb[m][l][1], bby0[l]);
        atomicMax(&s_bb[m][l][2], bbx1[l]); atomicMax(&s_bb[m][l][3], bby1[l]);
      }
    }
    __syncthreads();
    // fold this work item's boxes into the chunk's (one thread per (head, level))
    if (threadIdx.x < kHeads * LV) {
      const int h = threadIdx.x / LV, l = threadIdx.x % LV;
      if (s_bb[h][l][0] != INT_MAX) {
        int* bb = reinterpret_cast<int*>(ws.bbox + (static_cast<int64_t>(chunk) * kHeads + h) * LV + l);
        atomicMax(bb + 0, 65535 - s_bb[h][l][0]); atomicMax(bb + 1, 65535 - s_bb[h][l][1]);
        atomicMax(bb + 2, s_bb[h][l][2] + 1); atomicMax(bb + 3, s_bb[h][l][3] + 1);
      }
    }
  }
}

#ifdef MVG_GT_TRACE     // phase timestamps of CTA 0, units 6..9 (debug builds: tools/build_variant.sh gtrace -DMVG_GT_TRACE,
                        // read with tools/trace_gather.py).  Slot = (unit - 6) * 32 + event.
__device__ unsigned long long g_gt_trace[160];
__device__ int g_gt_trace_count[4];
__device__ __forceinline__ void gt_stamp(uint32_t seq, int ev) {
  if (blockIdx.x == 0 && seq >= 6 && seq < 10) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_gt_trace[(seq - 6) * 32 + ev] = t;
  }
}
#define GT_STAMP(seq, ev) gt_stamp(seq, ev)
#else
#define GT_STAMP(seq, ev)
#endif

// ------------------------------------------------------------------ the gather proper
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// Two samples of one (item, head, level): lane = quarter q * 8 + block column dx * 4 + 16-byte chunk c
// gathers, for sample points q and q + 4, the top and bottom texel rows of its block column and blends
// them in packed fp16 into the lane's 8-channel accumulator (a0..a3).  All pyramid levels of an
// (item, head) accumulate into the same registers.
//   sbase (shared): region + (lane & 7) * 16 - (box_y0 * bw + box_x0) * 64, so that the address of texel
//   (x0, y0) is sbase + (y0 * bw + x0) * 64;  gbase (global): level base + (lane & 7).
template <bool kShared>
__device__ __forceinline__ void blend_two(const uint4 rc, uint32_t sbase, const uint4* gbase, int bw,
                                          __half2& a0, __half2& a1, __half2& a2, __half2& a3) {
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    const uint32_t rxy = s ? rc.z : rc.x, rw = s ? rc.w : rc.y;
    const uint32_t x0 = rxy & 0xffffu, y0 = rxy >> 16;
    uint4 top, bot;
    if (kShared) {
      const uint32_t a = sbase + (y0 * static_cast<uint32_t>(bw) + x0) * 64u;
      top = lds128(a);
      bot = lds128(a + static_cast<uint32_t>(bw) * 64u);
    } else {
      const uint4* p = gbase + static_cast<int64_t>(y0 * static_cast<uint32_t>(bw) + x0) * 4;
      top = __ldg(p);
      bot = __ldg(p + bw * 4);
    }
    const __half2 w = u32_as_half2(rw);
    const __half2 wt = __low2half2(w), wb = __high2half2(w);
    a0 = __hfma2(wt, u32_as_half2(top.x), a0); a0 = __hfma2(wb, u32_as_half2(bot.x), a0);
    a1 = __hfma2(wt, u32_as_half2(top.y), a1); a1 = __hfma2(wb, u32_as_half2(bot.y), a1);
    a2 = __hfma2(wt, u32_as_half2(top.z), a2); a2 = __hfma2(wb, u32_as_half2(bot.z), a2);
    a3 = __hfma2(wt, u32_as_half2(top.w), a3); a3 = __hfma2(wb, u32_as_half2(bot.w), a3);
  }
}
// Two items at once from shared memory (same arithmetic per item as blend_two).
__device__ __forceinline__ void blend_pair(const uint4 ra, const uint4 rb, uint32_t sbase, int bw,
                                           __half2 (&a)[4], __half2 (&b)[4]) {
  const uint32_t bw64 = static_cast<uint32_t>(bw) * 64u;
  const uint32_t xa0 = ra.x & 0xffffu, ya0 = ra.x >> 16, xa1 = ra.z & 0xffffu, ya1 = ra.z >> 16;
  const uint32_t xb0 = rb.x & 0xffffu, yb0 = rb.x >> 16, xb1 = rb.z & 0xffffu, yb1 = rb.z >> 16;
  const uint32_t pa0 = sbase + (ya0 * static_cast<uint32_t>(bw) + xa0) * 64u;
  const uint32_t pa1 = sbase + (ya1 * static_cast<uint32_t>(bw) + xa1) * 64u;
  const uint32_t pb0 = sbase + (yb0 * static_cast<uint32_t>(bw) + xb0) * 64u;
  const uint32_t pb1 = sbase + (yb1 * static_cast<uint32_t>(bw) + xb1) * 64u;
#ifdef MVG_ABL_NOLDS      // timing experiment only (wrong results): no tile loads
  const uint4 ta0 = make_uint4(pa0, pa0, pa0, pa0), ba0 = make_uint4(bw64, pa0, bw64, pa0), ta1 = make_uint4(pa1, pa1, pa1, pa1), ba1 = ba0;
  const uint4 tb0 = make_uint4(pb0, pb0, pb0, pb0), bb0 = ba0, tb1 = make_uint4(pb1, pb1, pb1, pb1), bb1 = ba0;
#else
  const uint4 ta0 = lds128(pa0), ba0 = lds128(pa0 + bw64), ta1 = lds128(pa1), ba1 = lds128(pa1 + bw64);
  const uint4 tb0 = lds128(pb0), bb0 = lds128(pb0 + bw64), tb1 = lds128(pb1), bb1 = lds128(pb1 + bw64);
#endif
#ifdef MVG_ABL_NOBLEND    // timing experiment only (wrong results): one LOP3 per loaded register pair instead of two HFMA2
  auto blend = [](const uint4& top, const uint4& bot, uint32_t rw, __half2 (&c)[4]) {
    c[0] = u32_as_half2(half2_as_u32(c[0]) ^ top.x ^ bot.x); c[1] = u32_as_half2(half2_as_u32(c[1]) ^ top.y ^ bot.y);
    c[2] = u32_as_half2(half2_as_u32(c[2]) ^ top.z ^ bot.z); c[3] = u32_as_half2(half2_as_u32(c[3]) ^ top.w ^ (bot.w + rw));
  };
  auto blend_unused = [](const uint4& top, const uint4& bot, uint32_t rw, __half2 (&c)[4]) {
#else
  auto blend = [](const uint4& top, const uint4& bot, uint32_t rw, __half2 (&c)[4]) {
#endif
    const __half2 w = u32_as_half2(rw);
    const __half2 wt = __low2half2(w), wb = __high2half2(w);
    c[0] = __hfma2(wt, u32_as_half2(top.x), c[0]); c[0] = __hfma2(wb, u32_as_half2(bot.x), c[0]);
    c[1] = __hfma2(wt, u32_as_half2(top.y), c[1]); c[1] = __hfma2(wb, u32_as_half2(bot.y), c[1]);
    c[2] = __hfma2(wt, u32_as_half2(top.z), c[2]); c[2] = __hfma2(wb, u32_as_half2(bot.z), c[2]);
    c[3] = __hfma2(wt, u32_as_half2(top.w), c[3]); c[3] = __hfma2(wb, u32_as_half2(bot.w), c[3]);
  };
  blend(ta0, ba0, ra.y, a);
  blend(ta1, ba1, ra.w, a);
  blend(tb0, bb0, rb.y, b);
  blend(tb1, bb1, rb.w, b);
}
// Sums the 8 (q, dx) roles with a transposing butterfly; afterwards lane holds channel
// (lane & 3) * 8 + bit4 * 4 + bit3 * 2 + bit2 of the head.
__device__ __forceinline__ float reduce_roles(__half2 a0, __half2 a1, __half2 a2, __half2 a3, int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  const uint32_t u0 = half2_as_u32(a0), u1 = half2_as_u32(a1), u2 = half2_as_u32(a2), u3 = half2_as_u32(a3);
  const uint32_t keep0 = b4 ? u2 : u0, keep1 = b4 ? u3 : u1, send0 = b4 ? u0 : u2, send1 = b4 ? u1 : u3;
  const __half2 k0 = __hadd2(u32_as_half2(keep0), u32_as_half2(__shfl_xor_sync(0xffffffffu, send0, 16)));
  const __half2 k1 = __hadd2(u32_as_half2(keep1), u32_as_half2(__shfl_xor_sync(0xffffffffu, send1, 16)));
  const uint32_t v0 = half2_as_u32(k0), v1 = half2_as_u32(k1);
  __half2 kk = __hadd2(u32_as_half2(b3 ? v1 : v0), u32_as_half2(__shfl_xor_sync(0xffffffffu, b3 ? v0 : v1, 8)));
  kk = __hadd2(kk, u32_as_half2(__shfl_xor_sync(0xffffffffu, half2_as_u32(kk), 4)));
  return b2 ? __high2float(kk) : __low2float(kk);
}
__device__ __forceinline__ int lane_channel(int lane) {
  return (lane & 3) * 8 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
}

struct __align__(16) UnitDesc {
  int first, count, head, unit;
  int4 box[MVG_MAX_LEVELS];
};

template <int LV>
struct TileSmem {
  static constexpr int kTexels = region_cap<LV>(0) + region_cap<LV>(1) + region_cap<LV>(2) + region_cap<LV>(3);
  static constexpr int kRecOff = kTexels * 64;                       // LV record regions of kChunk x 128 B
  static constexpr int kDescOff = kRecOff + LV * kChunk * kRecBytes;
  static constexpr int kBytes = kDescOff + 2 * static_cast<int>(sizeof(UnitDesc)) + 128;
};

template <int LV>
__global__ void __launch_bounds__((kGWarps + 1) * 32, 1)
gather_tiles_kernel(const __half* __restrict__ value_hm, const MvgSampleParams prm,
                    __nv_bfloat16* __restrict__ sampled, const GatherWs ws, const __grid_constant__ TileMaps tm) {
  extern __shared__ __align__(128) uint8_t smem[];
  UnitDesc* sdesc = reinterpret_cast<UnitDesc*>(smem + TileSmem<LV>::kDescOff);      // [2]
  uint64_t* full = reinterpret_cast<uint64_t*>(sdesc + 2);                           // [LV]
  uint64_t* empty = full + MVG_MAX_LEVELS;                                           // [LV]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t region[LV], recs[LV];
  {
    uint32_t a = smem_u32(smem);
#pragma unroll
    for (int l = 0; l < LV; ++l) {
      region[l] = a;
      a += region_cap<LV>(l) * 64;
      recs[l] = smem_u32(smem) + TileSmem<LV>::kRecOff + l * kChunk * kRecBytes;
    }
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int l = 0; l < LV; ++l) { mbar_init(&full[l], 1); mbar_init(&empty[l], kGWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_enter();               // barriers are set up; everything below reads the previous kernels' outputs
  const int V = prm.views, B = prm.batch;

  if (warp == kGWarps) {
    // ===================== producer: unit descriptors, tile rows, record blocks =====================
    // The unit metadata runs through a three-stage pipeline - work-queue ticket (atomicAdd) -> loads of
    // that unit's chunk + boxes -> use - so that the staging loop never waits for a global round trip:
    // with a blocking fetch per unit (atomic, then dependent loads: ~2 us) on top of ~70 bulk copies at
    // ~30 ns each the producer warp needed ~4 us per unit against ~5 us of consumption, and tile staging
    // hardly overlapped the gather (ablation: the kernel without any loads / blends still took 103 us).
    // (Sharing one set of tiles between the chunks of a dense key was measured and dropped: the
    //  union boxes grow and a 4-chunk work item unbalances the tail - 189 -> 209 us.)
    const int n_units = ws.ctrs[0] * kHeads;
    int tk = 0;                                  // lane 0: ticket whose loads are not issued yet
    int lu = 0;                                  // unit whose loads are in flight
    int4 lch = make_int4(0, 0, 0, 0), lmm[LV];
    auto take_ticket = [&]() { if (lane == 0) tk = atomicAdd(ws.ctrs + 2, 1); };
    auto issue_loads = [&]() {
      lu = __shfl_sync(0xffffffffu, tk, 0);
      if (lu < n_units) {
        lch = ws.chunks[lu / kHeads];
#pragma unroll
        for (int l = 0; l < LV; ++l) lmm[l] = ws.bbox[static_cast<int64_t>(lu) * LV + l];
      }
    };
    take_ticket();
    issue_loads();
    take_ticket();
    uint32_t seq = 0;
    for (;;) {
      // ---- use stage: the next unit whose boxes fit the regions
      int unit;
      int4 ch, box[LV];                                      // {x0, y0, row pitch of the staged tile, rows}
      int pcls[LV];                                          // pitch class (tensor-map staging)
      for (;;) {
        unit = lu;
        ch = lch;
        bool fits = true;
        if (unit < n_units) {
#pragma unroll
          for (int l = 0; l < LV; ++l) {
            const int4 mm = lmm[l];                          // encoded min / max of the 2x2 block origins
            const int x0 = 65535 - mm.x, y0 = 65535 - mm.y;
            int bw = mm.z + 1 - x0;
            const int bh = mm.w + 1 - y0;
            pcls[l] = 0;
            if (tm.enabled) {                                // the narrowest pitch class that holds the box
              pcls[l] = bw <= tile_pitch<LV>(l, 2) ? 2 : bw <= tile_pitch<LV>(l, 1) ? 1 : 0;
              fits = fits && bw <= tile_pitch<LV>(l, 0);
              bw = tile_pitch<LV>(l, 0) - kPitchStep * pcls[l];
            }
            box[l] = make_int4(x0, y0, bw, bh);
            fits = fits && bw * bh <= region_cap<LV>(l);
          }
        }
        issue_loads();                                       // loads of the ticket taken one step ago
        take_ticket();
        if (unit >= n_units) { unit = -1; break; }
        if (fits) break;
        // (measured and dropped: also sending chunks of < 16 / 32 / 64 items to gather_direct_kernel to save
        //  their tile staging - the direct kernel costs 4-8x more per item: +70 / +95 / +160 us at Q = 1024)
        if (lane == 0) ws.direct_list[atomicAdd(ws.ctrs + 3, 1)] = unit;      // left to gather_direct_kernel
      }
      const uint32_t par = seq & 1u;
      if (lane == 0) GT_STAMP(seq, 0);
      mbar_wait(&empty[0], par ^ 1u);      // level 0 of unit seq-1 consumed => descriptor slot (seq & 1) is free
      if (lane == 0) GT_STAMP(seq, 1);
      UnitDesc* d = &sdesc[par];
      if (unit < 0) {
        if (lane == 0) { d->count = -1; mbar_arrive(&full[0]); }
        break;
      }
      const int head = unit % kHeads;
      const int first = ch.x, count = ch.y;
      const int v = ch.z % V, b = ch.z / V;
      const int64_t vrow0 = static_cast<int64_t>(v * B + b) * prm.spatial_size;
      if (lane == 0) {
        d->first = first; d->count = count; d->head = head; d->unit = unit;
#pragma unroll
        for (int l = 0; l < LV; ++l) d->box[l] = box[l];
      }
      __syncwarp();
      const __half* vh = value_hm + static_cast<int64_t>(head) * prm.value_head_stride;
      const uint2* rec_src = ws.params + (static_cast<int64_t>(head) * LV * ws.items + first) * 16;
#pragma unroll
      for (int l = 0; l < LV; ++l) {
        if (l > 0) mbar_wait(&empty[l], par ^ 1u);
        if (lane == 0) GT_STAMP(seq, 2 + 2 * l);
        const int bw = box[l].z, bh = box[l].w;
        const uint32_t row_bytes = static_cast<uint32_t>(bw) * 64u;
        const uint32_t rec_bytes = static_cast<uint32_t>(count) * kRecBytes;
        if (lane == 0) {
          mbar_expect_tx(&full[l], row_bytes * static_cast<uint32_t>(bh) + rec_bytes);
          bulk_g2s(recs[l], rec_src + static_cast<int64_t>(l) * ws.items * 16, rec_bytes, &full[l]);
        }
        __syncwarp();
        if (tm.enabled) {
          // floor(bh / 8) boxes of 8 rows, then bh % 8 boxes of one row: lane j issues op j (<= 16 ops)
          const int n8 = bh >> 3, nops = n8 + (bh & 7);
          if (lane < nops) {
            const int row = lane < n8 ? lane * 8 : n8 * 8 + (lane - n8);
            tma_load_5d(&tm.m[l][pcls[l]][lane < n8 ? 0 : 1], &full[l], region[l] + static_cast<uint32_t>(row) * row_bytes,
                        0, box[l].x, box[l].y + row, v * B + b, head);
          }
        } else {
          const __half* src0 = vh + (vrow0 + prm.level_start[l] + static_cast<int64_t>(box[l].y) * prm.level_w[l] + box[l].x) * 32;
          for (int r = lane; r < bh; r += 32)
            bulk_g2s(region[l] + static_cast<uint32_t>(r) * row_bytes, src0 + static_cast<int64_t>(r) * prm.level_w[l] * 32,
                     row_bytes, &full[l]);
        }
        __syncwarp();
        if (lane == 0) GT_STAMP(seq, 3 + 2 * l);
      }
      ++seq;
    }
  } else {
    // ===================== consumers: level-by-level gather =====================
    uint32_t seq = 0;
    const int q = lane >> 3, dx = (lane >> 2) & 1;
    const int chn = lane_channel(lane);
    const uint32_t rec_lane = static_cast<uint32_t>(dx * 64 + q * 16);
    for (;;) {
      const uint32_t par = seq & 1u;
      if (warp == 0 && lane == 0) GT_STAMP(seq, 15);
      mbar_wait(&full[0], par);
      if (warp == 0 && lane == 0) GT_STAMP(seq, 16);
      const UnitDesc* d = &sdesc[par];
      const int count = d->count;
      if (count < 0) break;
      const int first = d->first, head = d->head;
      // ids of the items this warp owns (i = warp + 16 k): lane k keeps item k's id for the final store
      int my_item = 0;
      if (warp + kGWarps * lane < count) my_item = __ldg(ws.sorted + first + warp + kGWarps * lane);
      // per-lane fp16 accumulators of the warp's <= 8 items, carried across the pyramid levels
      // (statically unrolled: the independent load -> blend chains of different items overlap)
      __half2 acc[kIPW][4];
#pragma unroll
      for (int k = 0; k < kIPW; ++k) acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = __float2half2_rn(0.f);
#pragma unroll
      for (int l = 0; l < LV; ++l) {
        if (l > 0) mbar_wait(&full[l], par);
        if (warp == 0 && lane == 0) GT_STAMP(seq, 17 + 3 * l);
        const int4 box = d->box[l];
        const uint32_t tile = region[l] + static_cast<uint32_t>(lane & 7) * 16u -
                              static_cast<uint32_t>(box.y * box.z + box.x) * 64u;
        const uint32_t rcs = recs[l] + rec_lane + static_cast<uint32_t>(warp) * kRecBytes;
#ifdef MVG_ABL_NOREC
        const uint4 rec0 = lds128(rcs);
#endif
        // items in pairs: inside a pair both record loads, then all eight tile loads are issued before
        // the blends, so the shared-memory latency of one item hides behind the other's arithmetic
#pragma unroll
        for (int k = 0; k < kIPW; k += 2) {
          const int ia = warp + kGWarps * k;
          constexpr int kLast = kIPW - 1;                // an odd kIPW leaves the last item without a partner
          if (k < kLast && ia + kGWarps < count) {       // warp-uniform: both items exist
#ifdef MVG_ABL_NOREC      // timing experiment only (wrong results): one record per warp and level
            const uint4 ra = rec0, rb = rec0;
#else
            const uint4 ra = lds128(rcs + static_cast<uint32_t>(k * kGWarps) * kRecBytes);
            const uint4 rb = lds128(rcs + static_cast<uint32_t>((k + 1) * kGWarps) * kRecBytes);
#endif
            blend_pair(ra, rb, tile, box.z, acc[k], acc[k < kLast ? k + 1 : k]);
          } else if (ia < count) {
            const uint4 rc = lds128(rcs + static_cast<uint32_t>(k * kGWarps) * kRecBytes);
            blend_two<true>(rc, tile, nullptr, box.z, acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
          }
        }
        __syncwarp();
        if (warp == 0 && lane == 0) GT_STAMP(seq, 18 + 3 * l);
        if (lane == 0) mbar_arrive(&empty[l]);
      }
#ifdef MVG_GT_TRACE
      if (warp == 0 && lane == 0) { GT_STAMP(seq, 27); if (blockIdx.x == 0 && seq >= 6 && seq < 10) g_gt_trace_count[seq - 6] = count; }
#endif
      // role reduction of all kIPW slots without branches (a missing item's accumulators are zero): the
      // independent shuffle chains overlap, only the store is predicated.  With one branchy block per item
      // this epilogue took 1.5 us of a 4.7 us unit (phase trace, tools/trace_gather.py).
      float vred[kIPW];
#pragma unroll
      for (int k = 0; k < kIPW; ++k) vred[k] = reduce_roles(acc[k][0], acc[k][1], acc[k][2], acc[k][3], lane);
#pragma unroll
      for (int k = 0; k < kIPW; ++k) {
        const int64_t item = __shfl_sync(0xffffffffu, my_item, k);
        if (warp + kGWarps * k < count) sampled[item * 256 + head * 32 + chn] = __float2bfloat16(vred[k]);
      }
      if (warp == 0 && lane == 0) GT_STAMP(seq, 28);
      ++seq;
    }
  }
}

// Units whose sample boxes exceed the shared-memory regions: one warp per (unit, item), same
// arithmetic from global memory.
template <int LV>
__global__ void __launch_bounds__(256)
gather_direct_kernel(const __half* __restrict__ value_hm, const MvgSampleParams prm,
                     __nv_bfloat16* __restrict__ sampled, const GatherWs ws) {
  const int lane = threadIdx.x & 31;
  pdl_enter();
  const int n_direct = ws.ctrs[3];
  const int q = lane >> 3, dx = (lane >> 2) & 1;
  const int chn = lane_channel(lane);
  const int V = prm.views, B = prm.batch;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  const int64_t gw = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
#pragma unroll 1
  for (int64_t w = gw; w < static_cast<int64_t>(n_direct) * kChunk; w += warps_total) {
    const int unit = ws.direct_list[w / kChunk];
    const int i = static_cast<int>(w % kChunk);
    const int chunk = unit / kHeads, head = unit % kHeads;
    const int4 ch = ws.chunks[chunk];
    if (i >= ch.y) continue;
    const int v = ch.z % V, b = ch.z / V;
    const int64_t vrow0 = static_cast<int64_t>(v * B + b) * prm.spatial_size;
    const __half* vh = value_hm + static_cast<int64_t>(head) * prm.value_head_stride;
    const uint2* rec = ws.params + (static_cast<int64_t>(head) * LV * ws.items + ch.x + i) * 16 + dx * 8 + q * 2;
    __half2 a0 = __float2half2_rn(0.f), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
    for (int l = 0; l < LV; ++l) {
      const uint4 rc = __ldg(reinterpret_cast<const uint4*>(rec + l * ws.items * 16));
      const uint4* gb = reinterpret_cast<const uint4*>(vh + (vrow0 + prm.level_start[l]) * 32) + (lane & 7);
      blend_two<false>(rc, 0u, gb, prm.level_w[l], a0, a1, a2, a3);
    }
    const float tot = reduce_roles(a0, a1, a2, a3, lane);
    const int64_t item = ws.sorted[ch.x + i];
    sampled[item * 256 + head * 32 + chn] = __float2bfloat16(tot);
  }
}

// Tensor maps of the head-major value tensor for the tile staging: per level a 5-D view {32 channels, W_l, H_l,
// view-frame row, head} of this layer's 8 heads, boxes {32, pitch, 8 | 1, 1, 1}, no swizzle (the staged tile is a dense
// [rows][pitch][32] array, what the consumers index).  Encoded once per (pointer, shape) and cached: a decoder
// call cycles through its L layers' slices of value_hm.
template <int LV>
static TileMaps get_tile_maps(const __half* vhm, const MvgSampleParams& prm) {       // by value: copied under the lock
  struct Entry { const void* ptr; MvgSampleParams prm; TileMaps maps; bool used; };
  static Entry cache[16];
  static int next = 0;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  for (Entry& e : cache)
    if (e.used && e.ptr == vhm && memcmp(&e.prm, &prm, sizeof(prm)) == 0) return e.maps;
  Entry& e = cache[next];
  next = (next + 1) % 16;
  e.used = true;
  e.ptr = vhm;
  e.prm = prm;
  memset(&e.maps, 0, sizeof(e.maps));
  static const bool want = []() { const char* v = getenv("MVG_TILE_TMA"); return v == nullptr || v[0] != '0'; }();
  EncodeTiledFn enc = get_encode_fn();
  bool ok = want && enc != nullptr;
  const cuuint64_t rows = static_cast<cuuint64_t>(prm.batch) * prm.views;
  for (int l = 0; ok && l < LV; ++l)
    for (int c = 0; ok && c < kPitchClasses; ++c)
      for (int h8 = 0; ok && h8 < 2; ++h8) {
        const cuuint64_t dims[5] = {32, static_cast<cuuint64_t>(prm.level_w[l]), static_cast<cuuint64_t>(prm.level_h[l]), rows,
                                    static_cast<cuuint64_t>(kHeads)};
        const cuuint64_t strides[4] = {64, static_cast<cuuint64_t>(prm.level_w[l]) * 64,
                                       static_cast<cuuint64_t>(prm.spatial_size) * 64,
                                       static_cast<cuuint64_t>(prm.value_head_stride) * 2};
        const cuuint32_t box[5] = {32, static_cast<cuuint32_t>(tile_pitch<LV>(l, c)), h8 == 0 ? 8u : 1u, 1, 1};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        void* base = const_cast<__half*>(vhm) + static_cast<int64_t>(prm.level_start[l]) * 32;
        ok = box[1] >= 2 && enc(&e.maps.m[l][c][h8], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, base, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
      }
  e.maps.enabled = ok ? 1 : 0;       // 0: the kernel stages tile rows with cp.async.bulk as before
  return e.maps;
}

template <int LV>
static int launch_gather(const __half* vhm, const __half* gmp, const float* qproj, const MvgSampleParams& prm,
                         __nv_bfloat16* sp, const float* ref2d, const float* refl_in, const GatherWs& ws,
                         cudaStream_t st) {
  constexpr int smem = TileSmem<LV>::kBytes;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gather_tiles_kernel<LV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gather_tiles, %d B): %s", smem, cudaGetErrorString(e));
      return MVG_ELAUNCH;
    }
    attr_done = true;
  }
  constexpr int psmem = kPWarps * static_cast<int>(sizeof(ParamScratch<LV>));
  static bool pattr_done = false;
  if (!pattr_done) {
    cudaError_t e = cudaFuncSetAttribute(sample_params_kernel<LV>, cudaFuncAttributeMaxDynamicSharedMemorySize, psmem);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(sample_params, %d B): %s", psmem, cudaGetErrorString(e));
      return MVG_ELAUNCH;
    }
    pattr_done = true;
  }
  const int64_t chunks_bound = ws.items / kChunk + ws.keys + 1;
  const int64_t parts_bound = chunks_bound * kParts;
  const int pgrid = static_cast<int>(parts_bound < MVG_P_MINBLK * kNumSMs ? parts_bound : MVG_P_MINBLK * kNumSMs);
  launch_k(sample_params_kernel<LV>, dim3(pgrid), dim3(kPWarps * 32), psmem, st, gmp, qproj, prm, ref2d, refl_in, ws);
  int rc = check_launch("mvg_project_sample_fused(sample_params)");
  if (rc != MVG_OK) return rc;
  const int64_t units_bound = chunks_bound * kHeads;
  const int ggrid = static_cast<int>(units_bound < kNumSMs ? units_bound : kNumSMs);
  launch_k(gather_tiles_kernel<LV>, dim3(ggrid), dim3((kGWarps + 1) * 32), smem, st, vhm, prm, sp, ws,
           get_tile_maps<LV>(vhm, prm));
  rc = check_launch("mvg_project_sample_fused(gather_tiles)");
  if (rc != MVG_OK) return rc;
  launch_k(gather_direct_kernel<LV>, dim3(kNumSMs), dim3(256), 0, st, vhm, prm, sp, ws);
  return check_launch("mvg_project_sample_fused(gather_direct)");
}

}  // namespace mvg

extern "C" int64_t mvg_project_sample_workspace_bytes(const MvgSampleParams* prm) {
  using namespace mvg;
  if (prm == nullptr || prm->batch <= 0 || prm->views <= 0 || prm->points <= 0 || prm->num_levels < 1 ||
      prm->num_levels > MVG_MAX_LEVELS)
    return -1;
  GatherWs w;
  return make_ws(*prm, nullptr, &w);
}

namespace mvg {

// argument checks shared by the three entry points (value_hm / gmap / qproj may be null for the binning half)
static int check_sample_call(const char* who, const MvgSampleParams* prm, const void* sampled, const void* workspace) {
  MVG_REQUIRE(prm && sampled && workspace, "%s: null pointer", who);
  MVG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "%s: workspace must be 256-byte aligned", who);
  MVG_REQUIRE(prm->num_levels >= 1 && prm->num_levels <= MVG_MAX_LEVELS, "%s: num_levels %d out of range", who,
              prm->num_levels);
  MVG_REQUIRE(prm->batch > 0 && prm->views > 0 && prm->points > 0, "%s: empty shape", who);
  MVG_REQUIRE(prm->ld_g >= 192 && prm->ld_g % 8 == 0, "%s: ld_g %d", who, prm->ld_g);
  MVG_REQUIRE(prm->value_head_stride >= static_cast<int64_t>(prm->batch) * prm->views * prm->spatial_size * 32 &&
                  prm->value_head_stride % 8 == 0,
              "%s: value_head_stride %lld", who, static_cast<long long>(prm->value_head_stride));
  int s = 0;
  for (int l = 0; l < prm->num_levels; ++l) {
    MVG_REQUIRE(prm->level_h[l] > 1 && prm->level_w[l] > 1 && prm->level_start[l] == s,
                "%s: level %d shape/start inconsistent", who, l);
    MVG_REQUIRE(prm->level_h[l] < 32768 && prm->level_w[l] < 32768, "%s: level %d larger than 32767 texels per side",
                who, l);
    s += prm->level_h[l] * prm->level_w[l];
  }
  MVG_REQUIRE(s == prm->spatial_size, "%s: spatial_size %d != sum H*W %d", who, prm->spatial_size, s);
  MVG_REQUIRE(static_cast<int64_t>(s) * 4 < (1ll << 31), "%s: per-view map too large for 32-bit texel offsets", who);
  MVG_REQUIRE(static_cast<int64_t>(prm->batch) * prm->views * prm->points < (1ll << 31), "%s: too many items", who);
  return MVG_OK;
}

// projection (or the given per-level reference points) + binning: everything that does not need qproj
static int run_project_bin(const float* ref3d, const float* cams, const MvgSampleParams* prm, void* sampled,
                           float* ref2d, uint8_t* bounding, const float* refl_in, void* workspace, cudaStream_t st) {
  GatherWs ws;
  make_ws(*prm, workspace, &ws);
  const int64_t BV = static_cast<int64_t>(prm->batch) * prm->views;
  const int64_t total = BV * prm->points;
  cudaError_t e = cudaMemsetAsync(ws.counts, 0, static_cast<size_t>(ws.zero_bytes), st);
  if (e != cudaSuccess) {
    set_error("mvg_project_bin: cudaMemsetAsync: %s", cudaGetErrorString(e));
    return MVG_ELAUNCH;
  }
  const dim3 pc_grid((prm->points + kPcThreads - 1) / kPcThreads, static_cast<unsigned>(BV));
  int rc;
  if (refl_in == nullptr) {
    launch_k(project_bin_kernel, pc_grid, dim3(kPcThreads), 0, st, ref3d, reinterpret_cast<const MvgCamera*>(cams), *prm,
             ref2d, bounding, static_cast<__nv_bfloat16*>(sampled), ws);
    rc = check_launch("mvg_project_bin(project_bin)");
  } else {
    bin_refl_kernel<<<pc_grid, kPcThreads, 0, st>>>(refl_in, *prm, ws);
    rc = check_launch("mvg_project_bin(bin_refl)");
  }
  if (rc != MVG_OK) return rc;
  launch_k(bin_scan_kernel, dim3(1), dim3(kScanThreads), 0, st, ws, ws.kx * ws.ky);
  rc = check_launch("mvg_project_bin(bin_scan)");
  if (rc != MVG_OK) return rc;
  launch_k(bin_scatter_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, ws);
  return check_launch("mvg_project_bin(bin_scatter)");
}

static int run_sample_gather(const void* value_hm, const void* gmap, const float* qproj, const MvgSampleParams* prm,
                             void* sampled, const float* ref2d, const float* refl_in, void* workspace,
                             cudaStream_t st) {
  GatherWs ws;
  make_ws(*prm, workspace, &ws);
  const __half* vhm = static_cast<const __half*>(value_hm);
  const __half* gmp = static_cast<const __half*>(gmap);
  __nv_bfloat16* sp = static_cast<__nv_bfloat16*>(sampled);
  switch (prm->num_levels) {
    case 1: return launch_gather<1>(vhm, gmp, qproj, *prm, sp, ref2d, refl_in, ws, st);
    case 2: return launch_gather<2>(vhm, gmp, qproj, *prm, sp, ref2d, refl_in, ws, st);
    case 3: return launch_gather<3>(vhm, gmp, qproj, *prm, sp, ref2d, refl_in, ws, st);
    default: return launch_gather<4>(vhm, gmp, qproj, *prm, sp, ref2d, refl_in, ws, st);
  }
}

}  // namespace mvg

extern "C" int mvg_project_bin(const float* ref3d, const float* cams, const MvgSampleParams* prm, void* sampled,
                               float* ref2d, uint8_t* bounding, const float* refl_in, void* workspace,
                               void* stream) {
  using namespace mvg;
  int rc = check_sample_call("mvg_project_bin", prm, sampled, workspace);
  if (rc != MVG_OK) return rc;
  MVG_REQUIRE(refl_in || (ref3d && cams && ref2d && bounding), "mvg_project_bin: projection inputs/outputs missing");
  return run_project_bin(ref3d, cams, prm, sampled, ref2d, bounding, refl_in, workspace,
                         static_cast<cudaStream_t>(stream));
}

extern "C" int mvg_sample_gather(const void* value_hm, const void* gmap, const float* qproj,
                                 const MvgSampleParams* prm, void* sampled, const float* ref2d,
                                 const float* refl_in, void* workspace, void* stream) {
  using namespace mvg;
  int rc = check_sample_call("mvg_sample_gather", prm, sampled, workspace);
  if (rc != MVG_OK) return rc;
  MVG_REQUIRE(value_hm && gmap && qproj && (refl_in || ref2d), "mvg_sample_gather: null pointer");
  MVG_REQUIRE((reinterpret_cast<uintptr_t>(value_hm) & 15) == 0 && (reinterpret_cast<uintptr_t>(gmap) & 15) == 0,
              "mvg_sample_gather: value / G map must be 16-byte aligned");
  return run_sample_gather(value_hm, gmap, qproj, prm, sampled, ref2d, refl_in, workspace,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int mvg_project_sample_fused(const float* ref3d, const float* cams, const void* value_hm,
                                        const void* gmap, const float* qproj, const MvgSampleParams* prm,
                                        void* sampled, float* ref2d, uint8_t* bounding,
                                        const float* refl_in, void* workspace, void* stream) {
  using namespace mvg;
  int rc = check_sample_call("mvg_project_sample_fused", prm, sampled, workspace);
  if (rc != MVG_OK) return rc;
  MVG_REQUIRE(value_hm && gmap && qproj, "mvg_project_sample_fused: null pointer");
  MVG_REQUIRE((reinterpret_cast<uintptr_t>(value_hm) & 15) == 0 && (reinterpret_cast<uintptr_t>(gmap) & 15) == 0,
              "mvg_project_sample_fused: value / G map must be 16-byte aligned");
  MVG_REQUIRE(refl_in || (ref3d && cams && ref2d && bounding),
              "mvg_project_sample_fused: projection inputs/outputs missing");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = run_project_bin(ref3d, cams, prm, sampled, ref2d, bounding, refl_in, workspace, st);
  if (rc != MVG_OK) return rc;
  return run_sample_gather(value_hm, gmap, qproj, prm, sampled, ref2d, refl_in, workspace, st);
}

extern "C" int mvg_project_points(const float* ref3d, const float* cams, int batch, int views, int points,
                                  float img_w, float img_h, float* ref2d, uint8_t* bounding, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(ref3d && cams && ref2d && bounding, "mvg_project_points: null pointer");
  MVG_REQUIRE(batch > 0 && views > 0 && points > 0, "mvg_project_points: empty shape");
  MVG_REQUIRE(static_cast<int64_t>(batch) * views <= 65535, "mvg_project_points: batch * views > 65535");
  const dim3 grid((points + kPcThreads - 1) / kPcThreads, static_cast<unsigned>(batch * views));
  project_points_kernel<<<grid, kPcThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      ref3d, reinterpret_cast<const MvgCamera*>(cams), views, points, img_w, img_h, ref2d, bounding);
  return check_launch("mvg_project_points");
}

#ifdef MVG_GT_TRACE
extern "C" __attribute__((visibility("default"))) int mvg_debug_gather_trace(unsigned long long* out160, int* counts4) {
  if (cudaMemcpyFromSymbol(out160, mvg::g_gt_trace, sizeof(unsigned long long) * 160) != cudaSuccess) return -1;
  return cudaMemcpyFromSymbol(counts4, mvg::g_gt_trace_count, sizeof(int) * 4) == cudaSuccess ? 0 : -1;
}
#endif
