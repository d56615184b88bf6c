// Fused query-feature update (dq_decoder.py:770-778 'MLP' branch + forward_ffn,
// mvp_decoder.py:94-98) for one 128-row tile of points, all on chip:
//
//   t2  = aver @ Wfu^T + bfu                      tcgen05, fp32 accumulator in TMEM
//   tu  = LayerNorm2(tgt + t2)                    epilogue: row statistics straight from TMEM
//   h_c = relu(tu @ W1[c]^T + b1[c])   c = 0..7   hidden layer in 128-column chunks, bf16 in smem
//   y  += h_c @ W2[:, c]^T                        second accumulator in TMEM
// The hidden chunks are software-pipelined: two 128-column accumulators and two h buffers alternate, so
// that the MMA warp issues chunk c+1's GEMM while the epilogue warps turn chunk c into relu(h_c) (with
// 256-column chunks and one accumulator the tensor pipe idled through every epilogue: 18 % busy).
//   out = LayerNorm3(tu + y + b2)
//
// Before this kernel the chain was 6 launches (3 GEMMs + 2 LayerNorm kernels + their bf16 / fp32
// round trips through HBM, 73 us per layer at M = 15 360); here the activations never leave the
// SM: the A operand of every GEMM is written by the previous epilogue directly in the
// K-major / SWIZZLE_128B layout the UMMA descriptors expect, t2 and y stay fp32 (the unfused
// path rounded both to bf16), and only the weights stream from L2 (1.15 MB per tile, 32 KB
// TMA stages).  One CTA per tile (120 tiles for Q = 1024), 10 warps:
//   warp 0      TMA producer: the aver tile (4 K-panels of 128 x 64) + 36 weight stages
//   warp 1      TMEM owner + single-thread tcgen05.mma issue (M = 128, N = 256, K = 16)
//   warps 2-9   epilogues; warp w owns TMEM lanes 32 (w % 4) .. +31 and one 128-column half,
//               the two warps of a row group exchange LayerNorm partial sums through smem
#include "tcgen05.cuh"

namespace mvg {

constexpr int kFcStages = 3;
constexpr int kFcStageBytes = 256 * kBlockK * 2;        // 32 KB: 256 weight rows x 64 K
constexpr int kFcPanelBytes = kBlockM * kBlockK * 2;    // 16 KB: 128 rows x 64 K
constexpr int kFcActBytes = 4 * kFcPanelBytes;          // 64 KB: a 128 x 256 bf16 activation tile
constexpr int kFcHBytes = 2 * kFcPanelBytes;            // 32 KB: one 128 x 128 bf16 hidden chunk
constexpr int kFcHC = 128;                              // hidden columns per chunk
constexpr int kFcThreads = 10 * 32;
constexpr int kFcSmemBytes = 2 * kFcActBytes + kFcStages * kFcStageBytes + 2048 /*stats*/ + 256 /*barriers*/;
static_assert(kFcSmemBytes <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
constexpr int kFcD = 256;                               // d_model

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
      "%13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void pair_barrier(int id) {   // the two epilogue warps of a row group
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

// 16 fp32 values of row r, columns [col, col + 16) -> bf16 -> K-major SWIZZLE_128B activation tile
__device__ __forceinline__ void store_act16(uint8_t* act, int r, int col, const float* v) {
  uint8_t* rowp = act + (col >> 6) * kFcPanelBytes + r * 128;
  const int j0 = (col & 63) >> 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint4 o;
    o.x = pack_bf16x2(v[8 * h + 0], v[8 * h + 1]); o.y = pack_bf16x2(v[8 * h + 2], v[8 * h + 3]);
    o.z = pack_bf16x2(v[8 * h + 4], v[8 * h + 5]); o.w = pack_bf16x2(v[8 * h + 6], v[8 * h + 7]);
    *reinterpret_cast<uint4*>(rowp + (((j0 + h) ^ (r & 7)) << 4)) = o;
  }
}

#ifdef MVG_FFN_TRACE    // phase timestamps of CTA 0 (debug builds: VARIANT_SRC=ffn_chain tools/build_variant.sh trace -DMVG_FFN_TRACE,
                        // read with tools/trace_ffn.py)
__device__ unsigned long long g_ffn_trace[64];
__device__ __forceinline__ void ffn_stamp(int slot) {
  if (blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_ffn_trace[slot] = t;
  }
}
#define FFN_STAMP(slot) ffn_stamp(slot)
#else
#define FFN_STAMP(slot)
#endif

struct FfnChainParams {
  const float* tgt;        // (M, 256) fp32
  const float* b_fu;       // (256)
  const float* g2;         // LayerNorm2 weight / bias
  const float* e2;
  const float* b1;         // (d_ffn)
  const float* b2;         // (256)
  const float* g3;
  const float* e3;
  float* out;              // (M, 256) fp32; also holds tu between the two LayerNorms
  int M;
  int n_chunks;            // d_ffn / 128
  float eps2, eps3;
};

// ---- coalesced movement of fp32 tile rows between global memory and the row-per-thread epilogues ----
// The TMEM layout gives every epilogue thread one row; reading / writing that row straight from / to global
// memory costs 32 sectors per warp instruction (phase trace of a tile: 18 of 48 us went there).  Instead the
// four warps of a column half move 128-row x 64-column slabs (32 KB) through a staging area in hbuf (idle
// during both LayerNorms): the global side is coalesced (16 lanes per 256-byte row piece), the TMEM side is
// one row per thread, and the 16-byte chunks are XOR-swizzled by the row so that both patterns are free of
// bank conflicts.
constexpr int kFcSlabCols = 64;
constexpr int kFcSlabBytes = kBlockM * kFcSlabCols * 4;          // 32 KB per column half
static_assert(2 * kFcSlabBytes <= kFcActBytes, "the two staging slabs live in hbuf");

__device__ __forceinline__ uint32_t slab_off(int r, int c) {     // row r, 16-byte chunk c (0..15)
  return static_cast<uint32_t>(r * 256 + ((c ^ (r & 15)) << 4));
}
__device__ __forceinline__ void half_barrier(int half) {         // the four epilogue warps of a column half
  asm volatile("bar.sync %0, 128;" ::"r"(5 + half) : "memory");
}
// global rows [0, rows_valid) x columns [col0, col0 + 64) of the tile starting at `base` -> staging
template <bool kLdg>
__device__ __forceinline__ void slab_load(uint8_t* stage, const float* base, int col0, int tid_h, int rows_valid) {
  const int c = tid_h & 15, r0 = tid_h >> 4;
  float4 t[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = r0 + i * 8;
    const float4* src = reinterpret_cast<const float4*>(base + static_cast<int64_t>(r) * kFcD + col0) + c;
    t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < rows_valid) t[i] = kLdg ? __ldg(src) : *src;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) *reinterpret_cast<float4*>(stage + slab_off(r0 + i * 8, c)) = t[i];
}
__device__ __forceinline__ void slab_store(const uint8_t* stage, float* base, int col0, int tid_h, int rows_valid) {
  const int c = tid_h & 15, r0 = tid_h >> 4;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r = r0 + i * 8;
    if (r < rows_valid)
      *(reinterpret_cast<float4*>(base + static_cast<int64_t>(r) * kFcD + col0) + c) =
          *reinterpret_cast<const float4*>(stage + slab_off(r, c));
  }
}

// x = acc + bias + residual for this warp's 128 columns of row r, written back to TMEM; returns the
// partial sum / sum of squares.  `res_tile` = the residual tile's row 0 (rows past `rows_valid` read as 0).
template <bool kLdg>
__device__ __forceinline__ void tmem_add_residual(uint32_t taddr, int half, const float* __restrict__ bias,
                                                  const float* res_tile, int rows_valid, uint8_t* stage, int tid_h,
                                                  int r, float& s, float& ss) {
  s = 0.f; ss = 0.f;
#pragma unroll 1
  for (int c4 = 0; c4 < 2; ++c4) {
    const int col0 = half * 128 + c4 * kFcSlabCols;
    slab_load<kLdg>(stage, res_tile, col0, tid_h, rows_valid);
    float4 bb[16];                                         // the slab's 64 bias values: one batch of loads
#pragma unroll
    for (int i = 0; i < 16; ++i) bb[i] = __ldg(reinterpret_cast<const float4*>(bias + col0) + i);
    half_barrier(half);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + j * 16;
      uint32_t u[16];
      tmem_ld16(taddr + col, u);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 b4 = bb[j * 4 + (i >> 2)];
        const float4 t4 = *reinterpret_cast<const float4*>(stage + slab_off(r, j * 4 + (i >> 2)));
        const float x0 = __uint_as_float(u[i + 0]) + b4.x + t4.x, x1 = __uint_as_float(u[i + 1]) + b4.y + t4.y;
        const float x2 = __uint_as_float(u[i + 2]) + b4.z + t4.z, x3 = __uint_as_float(u[i + 3]) + b4.w + t4.w;
        s += (x0 + x1) + (x2 + x3);
        ss += (x0 * x0 + x1 * x1) + (x2 * x2 + x3 * x3);
        u[i + 0] = __float_as_uint(x0); u[i + 1] = __float_as_uint(x1);
        u[i + 2] = __float_as_uint(x2); u[i + 3] = __float_as_uint(x3);
      }
      tmem_st16(taddr + col, u);
    }
    half_barrier(half);                                    // slab consumed before the next one overwrites it
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// LayerNorm over the 256 columns of a row whose pre-norm values sit in TMEM columns
// [taddr, taddr + 256) of this thread's lane.  `half` selects this warp's 128 columns; the partner
// warp (same lanes, other half) contributes its partial (sum, sum of squares) through `stats`.
// Variance = E[x^2] - mean^2 in fp32 (|mean| <~ std for these rows).  The normalised fp32 rows go to
// `out_tile` through the staging slab; fn(col, y[16]) sees them as well (bf16 activation tile).
template <typename F>
__device__ __forceinline__ void tmem_layernorm(uint32_t taddr, int half, int r, int pair_id, float2* stats,
                                               float s, float ss, const float* __restrict__ gamma,
                                               const float* __restrict__ beta, float eps, float* out_tile,
                                               int rows_valid, uint8_t* stage, int tid_h, F&& fn) {
  stats[half * 128 + r] = make_float2(s, ss);
  pair_barrier(pair_id);
  const float2 o = stats[(half ^ 1) * 128 + r];
  pair_barrier(pair_id);                                   // both reads done before the next use
  const float mean = (s + o.x) * (1.f / 256.f);
  const float var = fmaxf((ss + o.y) * (1.f / 256.f) - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
#pragma unroll 1
  for (int c4 = 0; c4 < 2; ++c4) {
    const int col0 = half * 128 + c4 * kFcSlabCols;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + j * 16;
      float4 gb[4], eb[4];                                  // this group's gamma / beta: eight loads in flight at once
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        gb[i] = __ldg(reinterpret_cast<const float4*>(gamma + col) + i);
        eb[i] = __ldg(reinterpret_cast<const float4*>(beta + col) + i);
      }
      uint32_t u[16];
      tmem_ld16(taddr + col, u);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float y[16];
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 g = gb[i >> 2];
        const float4 e = eb[i >> 2];
        y[i + 0] = (__uint_as_float(u[i + 0]) - mean) * rstd * g.x + e.x;
        y[i + 1] = (__uint_as_float(u[i + 1]) - mean) * rstd * g.y + e.y;
        y[i + 2] = (__uint_as_float(u[i + 2]) - mean) * rstd * g.z + e.z;
        y[i + 3] = (__uint_as_float(u[i + 3]) - mean) * rstd * g.w + e.w;
        *reinterpret_cast<float4*>(stage + slab_off(r, j * 4 + (i >> 2))) = make_float4(y[i], y[i + 1], y[i + 2], y[i + 3]);
      }
      fn(col, y);
    }
    half_barrier(half);
    slab_store(stage, out_tile, col0, tid_h, rows_valid);
    half_barrier(half);                                    // slab written out before the next one overwrites it
  }
}

__global__ void __launch_bounds__(kFcThreads, 1)
ffn_chain_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_fu,
                 const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
                 const FfnChainParams p) {
  // no static shared memory in this kernel: the dynamic window starts 1024-byte aligned (checked),
  // which SWIZZLE_128B tiles need; there is no room left for an alignment slack
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  if (threadIdx.x == 0) FFN_STAMP(41);
  uint8_t* xbuf = smem;                                  // aver tile, then tu (bf16)
  uint8_t* hbuf = smem + kFcActBytes;                    // relu(hidden chunk) (bf16)
  uint8_t* wbuf = smem + 2 * kFcActBytes;                // weight ring
  float2* stats = reinterpret_cast<float2*>(wbuf + kFcStages * kFcStageBytes);   // [2][128] (sum, sumsq)
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stats) + 2048);
  uint64_t* w_full = bars;                   // [kFcStages]
  uint64_t* w_empty = bars + kFcStages;      // [kFcStages]
  uint64_t* x_full = bars + 2 * kFcStages;
  uint64_t* x_free = x_full + 1;
  uint64_t* g0_full = x_full + 2;            // t2 (all 256 columns of acc1)
  uint64_t* acc2_full = x_full + 3;
  uint64_t* acc2_free = x_full + 4;
  uint64_t* tu_ready = x_full + 5;
  uint64_t* a1_full = x_full + 6;            // [2] hidden-chunk accumulator b holds tu W1[c]^T
  uint64_t* h_ready = x_full + 8;            // [2] relu(h_c) is in h buffer b, accumulator b is free
  uint64_t* h_free = x_full + 10;            // [2] the y GEMM has read h buffer b
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_full + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + kBlockM - 1) / kBlockM;
  const int NCH = p.n_chunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_fu)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w1)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w2)) : "memory");
    for (int s = 0; s < kFcStages; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    mbar_init(x_full, 1);
    mbar_init(x_free, 1);
    mbar_init(g0_full, 1);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_free, 8);
    mbar_init(tu_ready, 8);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&a1_full[b], 1);
      mbar_init(&h_ready[b], 8);
      mbar_init(&h_free[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc1 = tmem_base, acc2 = tmem_base + 256;
  pdl_enter();               // barriers / TMEM are set up; everything below touches global memory
  if (threadIdx.x == 0) FFN_STAMP(40);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t ws = 0, it = 0;
      auto load_w = [&](const CUtensorMap* map, int c0, int c1) {
        const int s = ws % kFcStages;
        mbar_wait(&w_empty[s], ((ws / kFcStages) & 1) ^ 1);
        mbar_expect_tx(&w_full[s], kFcStageBytes);
        tma_load_2d(map, &w_full[s], wbuf + s * kFcStageBytes, c0, c1);
        ++ws;
      };
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++it) {
        if (it > 0) mbar_wait(x_free, (it - 1) & 1);      // previous tile's MMAs are done with xbuf
        mbar_expect_tx(x_full, kFcActBytes);
        for (int kb = 0; kb < 4; ++kb)
          tma_load_2d(&tmap_x, x_full, xbuf + kb * kFcPanelBytes, kb * kBlockK, mt * kBlockM);
        for (int kb = 0; kb < 4; ++kb) load_w(&tmap_fu, kb * kBlockK, 0);
        // W1 chunk: 128 rows x 256 K = two stages of two 128 x 64 boxes; W2 chunk: 256 rows x 128 K = two
        // stages.  Same order as the MMA warp issues them: W1(0), then W1(c+1), W2(c).
        auto load_w1 = [&](int c) {
          for (int s2 = 0; s2 < 2; ++s2) {
            const int s = ws % kFcStages;
            mbar_wait(&w_empty[s], ((ws / kFcStages) & 1) ^ 1);
            mbar_expect_tx(&w_full[s], kFcStageBytes);
            tma_load_2d(&tmap_w1, &w_full[s], wbuf + s * kFcStageBytes, (2 * s2) * kBlockK, c * kFcHC);
            tma_load_2d(&tmap_w1, &w_full[s], wbuf + s * kFcStageBytes + kFcPanelBytes, (2 * s2 + 1) * kBlockK,
                        c * kFcHC);
            ++ws;
          }
        };
        load_w1(0);
        for (int c = 0; c < NCH; ++c) {
          if (c + 1 < NCH) load_w1(c + 1);
          for (int j = 0; j < 2; ++j) load_w(&tmap_w2, c * kFcHC + j * kBlockK, 0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(256), idesc_h = make_idesc_bf16(kFcHC);
      uint32_t ws = 0, it = 0, hc = 0;
      // t2: one 128 x 256 x 256 GEMM, A = 4 K-panels of xbuf, B = the next 4 weight stages
      auto gemm0 = [&]() {
        for (int kb = 0; kb < 4; ++kb, ++ws) {
          const int s = ws % kFcStages;
          mbar_wait(&w_full[s], (ws / kFcStages) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_smem_desc_sw128(smem_u32(xbuf + kb * kFcPanelBytes));
          const uint64_t db = make_smem_desc_sw128(smem_u32(wbuf + s * kFcStageBytes));
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_bf16(acc1, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          umma_commit(&w_empty[s]);
        }
      };
      // hidden chunk: 128 x 128 x 256, A = xbuf (tu), B = two stages of two 128 x 64 boxes -> accumulator b
      auto gemm1 = [&](uint32_t b) {
        for (int s2 = 0; s2 < 2; ++s2, ++ws) {
          const int s = ws % kFcStages;
          mbar_wait(&w_full[s], (ws / kFcStages) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t da = make_smem_desc_sw128(smem_u32(xbuf + (2 * s2 + j) * kFcPanelBytes));
            const uint64_t db = make_smem_desc_sw128(smem_u32(wbuf + s * kFcStageBytes + j * kFcPanelBytes));
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              umma_bf16(acc1 + b * kFcHC, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc_h,
                        (s2 | j | k) != 0 ? 1u : 0u);
          }
          umma_commit(&w_empty[s]);
        }
      };
      // y += h_c W2[:, c]^T: 128 x 256 x 128, A = h buffer b, B = two stages
      auto gemm2 = [&](uint32_t b, bool fresh) {
        for (int j = 0; j < 2; ++j, ++ws) {
          const int s = ws % kFcStages;
          mbar_wait(&w_full[s], (ws / kFcStages) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_smem_desc_sw128(smem_u32(hbuf + b * kFcHBytes + j * kFcPanelBytes));
          const uint64_t db = make_smem_desc_sw128(smem_u32(wbuf + s * kFcStageBytes));
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_bf16(acc2, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                      (!fresh || (j | k) != 0) ? 1u : 0u);
          umma_commit(&w_empty[s]);
        }
      };
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++it) {
        FFN_STAMP(0);
        mbar_wait(x_full, it & 1);
        FFN_STAMP(1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        gemm0();                                         // t2
        umma_commit(g0_full);
        FFN_STAMP(2);
        mbar_wait(tu_ready, it & 1);                     // LayerNorm2 wrote tu into xbuf, acc1 is free
        FFN_STAMP(3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        gemm1(hc & 1);                                   // hidden chunk 0
        umma_commit(&a1_full[hc & 1]);
        for (int c = 0; c < NCH; ++c, ++hc) {
          if (c + 1 < NCH) {
            // chunk c+1 into the other accumulator while the epilogue warps work on chunk c (its previous
            // reader, chunk c-1, was waited for below one iteration ago)
            gemm1((hc + 1) & 1);
            umma_commit(&a1_full[(hc + 1) & 1]);
            if (c + 1 == NCH - 1) umma_commit(x_free);   // the last GEMM that reads xbuf
          }
          mbar_wait(&h_ready[hc & 1], (hc >> 1) & 1);    // relu(h_c) is in its h buffer, its accumulator is free
          FFN_STAMP(4 + (c & 7));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (c == 0 && it > 0) {
            mbar_wait(acc2_free, (it - 1) & 1);          // previous tile's LayerNorm3 has drained acc2
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          gemm2(hc & 1, c == 0);                         // y += h_c W2[:, c]^T
          umma_commit(&h_free[hc & 1]);
        }
        umma_commit(acc2_full);
        FFN_STAMP(12);
      }
    }
  } else {
    // ===================== epilogues (warps 2..9) =====================
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                     // column half
    const int r = q * 32 + lane;                          // row inside the tile
    const int pair_id = 1 + q;
    const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
    const int tid_h = ((warp - 2) & 3) * 32 + lane;       // 0..127 inside the column half's four warps
    uint8_t* stage = hbuf + half * kFcSlabBytes;          // fp32 slab staging (hbuf is idle during the LayerNorms)
    uint32_t it = 0, hc = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++it) {
      const int rows_valid = min(kBlockM, p.M - mt * kBlockM);
      const float* ttile = p.tgt + static_cast<int64_t>(mt) * kBlockM * kFcD;
      float* otile = p.out + static_cast<int64_t>(mt) * kBlockM * kFcD;
      // ---- LayerNorm2(tgt + t2): x = acc1 + bfu + tgt written back to TMEM, then normalised
      mbar_wait(g0_full, it & 1);
      if (warp == 2 && lane == 0) FFN_STAMP(16);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float s_, ss_;
      tmem_add_residual<true>(acc1 + lane_sel, half, p.b_fu, ttile, rows_valid, stage, tid_h, r, s_, ss_);
      if (warp == 2 && lane == 0) FFN_STAMP(17);
      tmem_layernorm(acc1 + lane_sel, half, r, pair_id, stats, s_, ss_, p.g2, p.e2, p.eps2, otile, rows_valid,
                     stage, tid_h, [&](int col, const float* y) { store_act16(xbuf, r, col, y); });
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tu_ready);
      if (warp == 2 && lane == 0) FFN_STAMP(18);
      // ---- hidden chunks: relu(accumulator b + b1) -> h buffer b; this warp: 32 rows x 64 of the 128 columns
      for (int c = 0; c < NCH; ++c, ++hc) {
        const uint32_t b = hc & 1u, use = hc >> 1;
        // this thread's 64 bias values, all 16 loads in flight before the waits: read 16 bytes at a time right
        // before use they cost an L2 round trip per 16 columns (phase trace: 0.29 of 0.35 us per group)
        float4 bb[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) bb[i] = __ldg(reinterpret_cast<const float4*>(p.b1 + c * kFcHC + half * 64) + i);
        mbar_wait(&a1_full[b], use & 1);
        if (use > 0) mbar_wait(&h_free[b], (use - 1) & 1);       // the y GEMM of two chunks ago has read this buffer
        if (warp == 2 && lane == 0) FFN_STAMP(20 + (c & 7));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint8_t* hb = hbuf + b * kFcHBytes;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int col = half * 64 + cc * 16;
          uint32_t u[16];
          tmem_ld16(acc1 + b * kFcHC + lane_sel + col, u);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = bb[cc * 4 + (i >> 2)];
            v[i + 0] = fmaxf(__uint_as_float(u[i + 0]) + b4.x, 0.f);
            v[i + 1] = fmaxf(__uint_as_float(u[i + 1]) + b4.y, 0.f);
            v[i + 2] = fmaxf(__uint_as_float(u[i + 2]) + b4.z, 0.f);
            v[i + 3] = fmaxf(__uint_as_float(u[i + 3]) + b4.w, 0.f);
          }
          store_act16(hb, r, col, v);
          if (warp == 2 && lane == 0 && c == 3) FFN_STAMP(49 + cc);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (warp == 2 && lane == 0 && c == 3) FFN_STAMP(53);
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_ready[b]);
        if (warp == 2 && lane == 0 && c == 3) FFN_STAMP(54);
      }
      // ---- LayerNorm3(tu + y + b2)
      if (warp == 2 && lane == 0) FFN_STAMP(29);
      mbar_wait(acc2_full, it & 1);
      if (warp == 2 && lane == 0) FFN_STAMP(30);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // + tu: written to `out` by the LayerNorm2 slab stores of this column half (ordered by half_barrier)
      tmem_add_residual<false>(acc2 + lane_sel, half, p.b2, otile, rows_valid, stage, tid_h, r, s_, ss_);
      tmem_layernorm(acc2 + lane_sel, half, r, pair_id, stats, s_, ss_, p.g3, p.e3, p.eps3, otile, rows_valid,
                     stage, tid_h, [&](int, const float*) {});
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_free);
      if (warp == 2 && lane == 0) FFN_STAMP(31);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace mvg

extern "C" int mvg_ffn_chain(const void* aver_bf16, const float* tgt, const void* w_fu, const float* b_fu,
                             const float* g2, const float* e2, float eps2, const void* w1, const float* b1,
                             const void* w2, const float* b2, const float* g3, const float* e3, float eps3,
                             int64_t M, int d_ffn, float* out, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(aver_bf16 && tgt && w_fu && b_fu && g2 && e2 && w1 && b1 && w2 && b2 && g3 && e3 && out,
              "mvg_ffn_chain: null pointer");
  MVG_REQUIRE(M > 0 && M < (1ll << 31), "mvg_ffn_chain: bad M");
  MVG_REQUIRE(d_ffn >= 256 && d_ffn % 256 == 0, "mvg_ffn_chain: d_ffn=%d must be a multiple of 256", d_ffn);
  const void* ptrs[] = {aver_bf16, tgt, w_fu, b_fu, g2, e2, w1, b1, w2, b2, g3, e3, out};
  for (const void* q : ptrs)
    MVG_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, "mvg_ffn_chain: operands must be 16-byte aligned");
  CUtensorMap tx, tfu, tw1, tw2;
  int rc = make_tmap(&tx, aver_bf16, M, kFcD, kBlockM);
  if (rc) return rc;
  rc = make_tmap(&tfu, w_fu, kFcD, kFcD, 256);
  if (rc) return rc;
  rc = make_tmap(&tw1, w1, d_ffn, kFcD, kFcHC);
  if (rc) return rc;
  rc = make_tmap(&tw2, w2, kFcD, d_ffn, 256);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(ffn_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFcSmemBytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d): %s", kFcSmemBytes, cudaGetErrorString(e));
      return MVG_ELAUNCH;
    }
    attr_set = true;
  }
  FfnChainParams p{tgt, b_fu, g2, e2, b1, b2, g3, e3, out, static_cast<int>(M), d_ffn / kFcHC, eps2, eps3};
  const int m_tiles = static_cast<int>((M + kBlockM - 1) / kBlockM);
  const int grid = m_tiles < kNumSMs ? m_tiles : kNumSMs;
  launch_k(ffn_chain_kernel, dim3(grid), dim3(kFcThreads), kFcSmemBytes, static_cast<cudaStream_t>(stream), tx, tfu, tw1,
           tw2, p);
  return check_launch("mvg_ffn_chain");
}

#ifdef MVG_FFN_TRACE
extern "C" __attribute__((visibility("default"))) int mvg_debug_ffn_trace(unsigned long long* out64) {
  return cudaMemcpyFromSymbol(out64, mvg::g_ffn_trace, sizeof(unsigned long long) * 64) == cudaSuccess ? 0 : -1;
}
#endif
