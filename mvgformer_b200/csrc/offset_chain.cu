// Fused offset_net MLP (lib/models/dq_decoder.py:97-111, 659-717; MLP of
// lib/models/multi_view_pose_transformer.py:81-102) on the rows that need it.
//
//   h1  = relu(attn @ W1^T + b1)          tcgen05, fp32 accumulator in TMEM, bf16 -> smem
//   h2  = relu(h1 @ W2^T + b2)            second accumulator in TMEM, stays fp32
//   out = h2 @ W3^T + b3   (3 columns)    fp32 dot products in the epilogue
//
// The reference runs the MLP on the SELECTED queries only (it gathers them into a padded
// rectangle first, dq_decoder.py:899-932).  Here the 128-row A tile of the first GEMM is gathered
// straight from the rows of the selected queries (all views, all joints) by the epilogue warps
// - the id arrays come from mvg_select_pad, the row count lives on the device, so there is
// no host synchronisation - and (dx, dy, confidence logit) are scattered back to the rows'
// natural positions; rows of unselected queries are never read (mvg_offsets_dlt ignores them).
// Before: three full-size GEMMs (76 800 rows, 0.24 ms per step at Q = 1024) with two bf16 round
// trips of the hidden activations through HBM and a 16-column padded GEMM for the 3-wide head.
//
// One CTA per SM, persistent over the active tiles, 14 warps:
//   warp 0       TMA producer: 8 weight stages (256 rows x 64 K, 32 KB) per tile
//   warp 1       TMEM owner + single-thread tcgen05.mma issue (M = 128, N = 256, K = 16)
//   warps 2-9    the two epilogues (h1 -> smem, head)
//   warps 10-13  row gather of the NEXT tile into the K-major SWIZZLE_128B A tile: xbuf is free as soon as
//                GEMM 1 has read it, so the gather and GEMM 1 of tile i+1 run under the epilogues of tile i
//                (with the gather done by the epilogue warps a tile took ~16 us for 2.2 us of MMA)
#include "tcgen05.cuh"

namespace mvg {

constexpr int kOcStages = 3;
constexpr int kOcStageBytes = 256 * kBlockK * 2;        // 32 KB: 256 weight rows x 64 K
constexpr int kOcPanelBytes = kBlockM * kBlockK * 2;    // 16 KB: 128 rows x 64 K
constexpr int kOcActBytes = 4 * kOcPanelBytes;          // 64 KB: a 128 x 256 bf16 activation tile
constexpr int kOcGatherWarps = 4;
constexpr int kOcThreads = (10 + kOcGatherWarps) * 32;
constexpr int kOcSmemBytes = 2 * kOcActBytes + kOcStages * kOcStageBytes + 2048 /*partials*/ + 256 /*barriers*/;
static_assert(kOcSmemBytes <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");

struct OffsetChainParams {
  const __nv_bfloat16* attn;     // (B*V*N, 256) bf16
  const int32_t* info;           // [0] = number of selected queries, [1] = padded queries per frame (device)
  const int64_t* query_ids_pad;  // (B * info[1]) query ids of the padded rectangle   (mvg_select_pad outputs)
  const int64_t* batch_ids_rev;  // (n_sel) frame of the s-th selected query
  const int64_t* query_ids_rev;  // (n_sel) its position inside the frame's padded row
  const float* b1;               // (256)
  const float* b2;               // (256)
  const float* w3;               // (3, 256) fp32
  const float* b3;               // (3)
  float* out;                    // (B*V*N, out_ld) fp32, columns 0..2 written for the active rows
  int out_ld;
  int views, joints, points;     // V, J, N = Q*J
};

// 16 fp32 values of row r, columns [col, col + 16) -> bf16 -> K-major SWIZZLE_128B activation tile
__device__ __forceinline__ void oc_store_act16(uint8_t* act, int r, int col, const float* v) {
  uint8_t* rowp = act + (col >> 6) * kOcPanelBytes + r * 128;
  const int j0 = (col & 63) >> 3;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint4 o;
    o.x = pack_bf16x2(v[8 * h + 0], v[8 * h + 1]); o.y = pack_bf16x2(v[8 * h + 2], v[8 * h + 3]);
    o.z = pack_bf16x2(v[8 * h + 4], v[8 * h + 5]); o.w = pack_bf16x2(v[8 * h + 6], v[8 * h + 7]);
    *reinterpret_cast<uint4*>(rowp + (((j0 + h) ^ (r & 7)) << 4)) = o;
  }
}

// active row index a -> row of the (B, V, N, .) tensors; -1 past the end
__device__ __forceinline__ int64_t oc_row_of(int64_t a, int64_t n_active, const OffsetChainParams& p) {
  if (a >= n_active) return -1;
  const int vj = p.views * p.joints;
  const int64_t s = a / vj;
  const int rem = static_cast<int>(a - s * vj);
  const int v = rem / p.joints, j = rem - v * p.joints;
  // dq_decoder.py:941-947: valid entry s of the padded (B, max_count) rectangle
  const int64_t b = __ldg(p.batch_ids_rev + s), pos = __ldg(p.query_ids_rev + s);
  const int64_t q = __ldg(p.query_ids_pad + b * __ldg(p.info + 1) + pos);
  return (b * p.views + v) * p.points + q * p.joints + j;
}

__global__ void __launch_bounds__(kOcThreads, 1)
offset_chain_kernel(const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
                    const OffsetChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* xbuf = smem;                                  // gathered attn rows (bf16, swizzled)
  uint8_t* hbuf = smem + kOcActBytes;                    // relu(h1) (bf16, swizzled)
  uint8_t* wbuf = smem + 2 * kOcActBytes;                // weight ring
  float4* partial = reinterpret_cast<float4*>(wbuf + kOcStages * kOcStageBytes);   // [128] head partial sums of half 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(partial) + 2048);
  uint64_t* w_full = bars;                   // [kOcStages]
  uint64_t* w_empty = bars + kOcStages;      // [kOcStages]
  uint64_t* x_ready = bars + 2 * kOcStages;
  uint64_t* x_free = x_ready + 1;
  uint64_t* acc1_full = x_ready + 2;
  uint64_t* h_ready = x_ready + 3;
  uint64_t* acc2_full = x_ready + 4;
  uint64_t* acc2_free = x_ready + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_ready + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_enter();               // the row count below is the previous kernel's (mvg_select_pad) output
  const int64_t n_active = static_cast<int64_t>(__ldg(p.info)) * p.views * p.joints;
  const int m_tiles = static_cast<int>((n_active + kBlockM - 1) / kBlockM);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w1)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w2)) : "memory");
    for (int s = 0; s < kOcStages; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    mbar_init(x_ready, kOcGatherWarps);
    mbar_init(x_free, 1);
    mbar_init(acc1_full, 1);
    mbar_init(h_ready, 8);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_free, 8);
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

  if (warp == 0) {
    // ===================== TMA producer: W1 then W2, four K-panels each, per tile =====================
    if (lane == 0) {
      uint32_t ws = 0;
      auto load_w = [&](const CUtensorMap* map, int c0) {
        const int s = ws % kOcStages;
        mbar_wait(&w_empty[s], ((ws / kOcStages) & 1) ^ 1);
        mbar_expect_tx(&w_full[s], kOcStageBytes);
        tma_load_2d(map, &w_full[s], wbuf + s * kOcStageBytes, c0, 0);
        ++ws;
      };
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
        for (int kb = 0; kb < 4; ++kb) load_w(&tmap_w1, kb * kBlockK);
        for (int kb = 0; kb < 4; ++kb) load_w(&tmap_w2, kb * kBlockK);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(256);
      uint32_t ws = 0, it = 0;
      auto gemm = [&](uint8_t* abuf, uint32_t tmem_d) {
        for (int kb = 0; kb < 4; ++kb, ++ws) {
          const int s = ws % kOcStages;
          mbar_wait(&w_full[s], (ws / kOcStages) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = make_smem_desc_sw128(smem_u32(abuf + kb * kOcPanelBytes));
          const uint64_t db = make_smem_desc_sw128(smem_u32(wbuf + s * kOcStageBytes));
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_bf16(tmem_d, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          umma_commit(&w_empty[s]);
        }
      };
      for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++it) {
        mbar_wait(x_ready, it & 1);                      // A tile gathered; acc1 drained by the previous h1 epilogue
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        gemm(xbuf, acc1);
        umma_commit(acc1_full);
        umma_commit(x_free);
        mbar_wait(h_ready, it & 1);                      // relu(h1) is in hbuf
        if (it > 0) mbar_wait(acc2_free, (it - 1) & 1);  // previous tile's head epilogue has drained acc2
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        gemm(hbuf, acc2);
        umma_commit(acc2_full);
      }
    }
  } else if (warp >= 10) {
    // ===================== row gather (warps 10..13): one row of the tile per thread =====================
    const int gr = threadIdx.x - 320;                     // 0..127
    uint32_t it = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++it) {
      const int64_t row = oc_row_of(static_cast<int64_t>(mt) * kBlockM + gr, n_active, p);
      const uint4* src = reinterpret_cast<const uint4*>(p.attn + (row < 0 ? 0 : row) * 256);
#pragma unroll 1
      for (int gh = 0; gh < 2; ++gh) {
        uint4 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = row < 0 ? make_uint4(0u, 0u, 0u, 0u) : __ldg(src + gh * 16 + j);
        if (gh == 0 && it > 0) mbar_wait(x_free, (it - 1) & 1);     // GEMM 1 of the previous tile has read xbuf
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = gh * 128 + j * 8;
          *reinterpret_cast<uint4*>(xbuf + (col >> 6) * kOcPanelBytes + gr * 128 +
                                    ((((col & 63) >> 3) ^ (gr & 7)) << 4)) = v[j];
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(x_ready);
    }
  } else {
    // ===================== epilogues (warps 2..9) =====================
    const int q = warp & 3;                               // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                     // column half
    const int r = q * 32 + lane;                          // row inside the tile (epilogues)
    const int pair_id = 1 + q;
    const uint32_t lane_sel = static_cast<uint32_t>(q * 32) << 16;
    uint32_t it = 0;
    for (int mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++it) {
      // ---- h1 = relu(acc1 + b1) -> hbuf (bf16).  hbuf is free: this thread waited acc2_full of the previous tile
      // The per-column vectors are fetched in batches ahead of the TMEM reads: read 16 bytes at a time right
      // before use, every group of 16 columns paid an L2 round trip (phase trace of ffn_chain: 0.29 of 0.35 us).
      float4 bb[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) bb[i] = __ldg(reinterpret_cast<const float4*>(p.b1 + half * 128) + i);
      mbar_wait(acc1_full, it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int c4 = 0; c4 < 2; ++c4) {
        if (c4 == 1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) bb[i] = __ldg(reinterpret_cast<const float4*>(p.b1 + half * 128 + 64) + i);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = half * 128 + c4 * 64 + j * 16;
          uint32_t u[16];
          tmem_ld16(acc1 + lane_sel + col, u);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b4 = bb[j * 4 + (i >> 2)];
            v[i + 0] = fmaxf(__uint_as_float(u[i + 0]) + b4.x, 0.f);
            v[i + 1] = fmaxf(__uint_as_float(u[i + 1]) + b4.y, 0.f);
            v[i + 2] = fmaxf(__uint_as_float(u[i + 2]) + b4.z, 0.f);
            v[i + 3] = fmaxf(__uint_as_float(u[i + 3]) + b4.w, 0.f);
          }
          oc_store_act16(hbuf, r, col, v);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(h_ready);
      // ---- head: out = relu(acc2 + b2) @ W3^T + b3, fp32
      mbar_wait(acc2_full, it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll 2
      for (int cc = 0; cc < 8; ++cc) {
        const int col = half * 128 + cc * 16;
        float4 vb[4], va[4], vw[4], vc[4];                  // b2 and the three rows of W3 for this group: 16 loads in flight
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          vb[i] = __ldg(reinterpret_cast<const float4*>(p.b2 + col) + i);
          va[i] = __ldg(reinterpret_cast<const float4*>(p.w3 + col) + i);
          vw[i] = __ldg(reinterpret_cast<const float4*>(p.w3 + 256 + col) + i);
          vc[i] = __ldg(reinterpret_cast<const float4*>(p.w3 + 512 + col) + i);
        }
        uint32_t u[16];
        tmem_ld16(acc2 + lane_sel + col, u);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = vb[i >> 2];
          const float4 wa = va[i >> 2];
          const float4 wb = vw[i >> 2];
          const float4 wc = vc[i >> 2];
          const float h0 = fmaxf(__uint_as_float(u[i + 0]) + b4.x, 0.f), h1 = fmaxf(__uint_as_float(u[i + 1]) + b4.y, 0.f);
          const float h2 = fmaxf(__uint_as_float(u[i + 2]) + b4.z, 0.f), h3 = fmaxf(__uint_as_float(u[i + 3]) + b4.w, 0.f);
          o0 += h0 * wa.x + h1 * wa.y + h2 * wa.z + h3 * wa.w;
          o1 += h0 * wb.x + h1 * wb.y + h2 * wb.z + h3 * wb.w;
          o2 += h0 * wc.x + h1 * wc.y + h2 * wc.z + h3 * wc.w;
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_free);
      // the two column halves of a row sit in two warps: exchange through shared memory
      if (half == 1) partial[r] = make_float4(o0, o1, o2, 0.f);
      asm volatile("bar.sync %0, 64;" ::"r"(pair_id) : "memory");
      if (half == 0) {
        const float4 other = partial[r];
        const int64_t row = oc_row_of(static_cast<int64_t>(mt) * kBlockM + r, n_active, p);
        if (row >= 0) {
          float* dst = p.out + row * p.out_ld;
          dst[0] = o0 + other.x + __ldg(p.b3 + 0);
          dst[1] = o1 + other.y + __ldg(p.b3 + 1);
          dst[2] = o2 + other.z + __ldg(p.b3 + 2);
        }
      }
      asm volatile("bar.sync %0, 64;" ::"r"(pair_id) : "memory");      // partial[] is reused by the next tile
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

extern "C" int mvg_offset_chain(const void* attn_bf16, const int32_t* info, const int64_t* query_ids_pad,
                                const int64_t* batch_ids_rev, const int64_t* query_ids_rev, const void* w1, const float* b1, const void* w2,
                                const float* b2, const float* w3, const float* b3, int batch, int views, int queries,
                                int joints, float* out, int out_ld, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(attn_bf16 && info && query_ids_pad && batch_ids_rev && query_ids_rev && w1 && b1 && w2 && b2 && w3 && b3 && out,
              "mvg_offset_chain: null pointer");
  MVG_REQUIRE(batch > 0 && views > 0 && queries > 0 && joints > 0 && out_ld >= 3, "mvg_offset_chain: bad shape");
  const void* ptrs[] = {attn_bf16, w1, b1, w2, b2, w3};
  for (const void* q : ptrs)
    MVG_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, "mvg_offset_chain: operands must be 16-byte aligned");
  const int64_t rows = static_cast<int64_t>(batch) * views * queries * joints;
  MVG_REQUIRE(rows < (1ll << 31), "mvg_offset_chain: too many rows");
  CUtensorMap t1, t2;
  int rc = make_tmap(&t1, w1, 256, 256, 256);
  if (rc) return rc;
  rc = make_tmap(&t2, w2, 256, 256, 256);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(offset_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kOcSmemBytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d): %s", kOcSmemBytes, cudaGetErrorString(e));
      return MVG_ELAUNCH;
    }
    attr_set = true;
  }
  OffsetChainParams p{static_cast<const __nv_bfloat16*>(attn_bf16), info, query_ids_pad, batch_ids_rev, query_ids_rev, b1, b2, w3,
                      b3, out,
                      out_ld, views, joints, queries * joints};
  const int64_t max_tiles = (rows + kBlockM - 1) / kBlockM;
  const int grid = static_cast<int>(max_tiles < kNumSMs ? max_tiles : kNumSMs);
  launch_k(offset_chain_kernel, dim3(grid), dim3(kOcThreads), kOcSmemBytes, static_cast<cudaStream_t>(stream), t1, t2, p);
  return check_launch("mvg_offset_chain");
}
