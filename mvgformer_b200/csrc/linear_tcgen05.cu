// Dense projection  out = act(A @ W^T + bias)  on 5th-generation tensor cores (sm_100a).
//
//   A (M,K) bf16 row-major (activations / channels-last pyramid) - or the NCHW pyramid levels in
//   place, loaded as an MN-major operand (struct NchwA, mvg_value_proj_gemm_nchw) -, W (Nout,K) bf16
//   row-major (nn.Linear layout), fp32 accumulation in TMEM, bias + ReLU + bf16/fp32 conversion fused
//   in the epilogue.  Used for every nn.Linear on the hot path (kernel table):
//   rayconv / sampling_offsets / attention_weights on the pyramid (one launch for all L layers),
//   the per-point qproj, output_proj, feature_update_mlp, FFN and the offset_net MLP.
//
// Structure (persistent, one CTA per SM, 10 warps, warp-specialised):
//   warp 0  TMA producer : loads its weight chunk W[n0:n0+NC, :] once (<= 128 KB, stationary),
//                          then streams 128 x 64 activation boxes (cp.async.bulk.tensor,
//                          SWIZZLE_128B) through a 5-stage ring, mbarrier expect_tx/complete_tx
//   warp 1  MMA issuer   : allocates all 512 TMEM columns (two NC-wide fp32 accumulators), one
//                          thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=NC<=256,
//                          K=16) from shared-memory descriptors; tcgen05.commit frees ring slots
//                          and hands a finished accumulator to the epilogue
//   warps 2-9 epilogue   : tcgen05.ld (32 lanes x 16 columns) -> +bias -> ReLU -> row mask ->
//                          convert -> 16-byte global stores, overlapping the next tile's MMAs;
//                          warp w may only touch TMEM lanes [32 (w%4), 32 (w%4) + 32)
#include "tcgen05.cuh"

namespace mvg {


constexpr int kStages = 4;                       // A ring: 4 x (128 rows x 64 K) = 64 KB
constexpr int kATileBytes = kBlockM * kBlockK * 2;
constexpr int kMaxWBytes = 128 * 1024;           // stationary weight chunk
constexpr int kEpiWarps = 8;
constexpr int kGemmThreadsV2 = (2 + kEpiWarps) * 32;
constexpr int kAccStride = 256;                  // TMEM columns between the two accumulators
constexpr int kStageOutBytes = 32 * 128;         // per epilogue warp: 32 rows x 128 B, SWIZZLE_128B
constexpr int kSmemBytesV2 = kMaxWBytes + kStages * kATileBytes + kEpiWarps * kStageOutBytes + 1024 + 256;

// Persistent, weight-stationary GEMM.  CTA c keeps the weight chunk  W[n0 : n0+NC, :]  resident in
// shared memory (<= 128 KB) and streams 128-row activation tiles through a 5-stage TMA ring; the
// accumulator is double-buffered in TMEM so the 8 epilogue warps drain tile i while the MMA warp
// already issues tile i+1.  grid = n_chunks x groups; the CTAs of one group sweep the same M tiles
// at the same time, so an A tile is read from HBM once and hit in L2 by the other chunks.
// A operand given as the NCHW pyramid levels themselves (mvg_value_proj_gemm_nchw): per level a 3-D map
// {s, channel, view-frame row}; an M tile of 128 texels is two {64 s, 64 c} boxes = an MN-major operand, so
// the channels-last copy of the pyramid (mvg_pyramid_to_channels_last: 206 MB of traffic) is never made.
struct NchwA {
  CUtensorMap map[MVG_MAX_LEVELS];
  int tile_start[MVG_MAX_LEVELS + 1];    // first 128-texel tile of level l inside one view-frame row
  int num_levels;                        // 0: A is the row-major (M, K) matrix of tmap_a
  int tiles_per_row;                     // S / 128
};

template <typename OutT>
__global__ void __launch_bounds__(kGemmThreadsV2, 1)
linear_tcgen05_ws_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ NchwA nchw,
                         const __grid_constant__ CUtensorMap tmap_w,
                         const __grid_constant__ CUtensorMap tmap_out,
                         const __grid_constant__ CUtensorMap tmap_hm, int hm_period, int use_tma_store,
                         const float* __restrict__ bias,
                         const uint8_t* __restrict__ row_mask, OutT* __restrict__ out, int M, int N,
                         int K, int NC, int n_chunks, int groups, int64_t ldo, int relu) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + kMaxWBytes;
  uint8_t* smem_out = smem_a + kStages * kATileBytes;              // 1024-aligned staging tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + kEpiWarps * kStageOutBytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + kStages;
  uint64_t* acc_full = bars + 2 * kStages;       // [2]
  uint64_t* acc_empty = bars + 2 * kStages + 2;  // [2]
  uint64_t* w_full = bars + 2 * kStages + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x % n_chunks;
  const int group = blockIdx.x / n_chunks;
  const int n0 = chunk * NC;
  const int num_kb = K / kBlockK;
  const int m_tiles = (M + kBlockM - 1) / kBlockM;
  const int w_kb_bytes = NC * kBlockK * 2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kEpiWarps);
    }
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_enter();               // barriers / TMEM are set up; everything below touches global memory

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(w_full, static_cast<uint32_t>(num_kb * w_kb_bytes));
      for (int kb = 0; kb < num_kb; ++kb)
        tma_load_2d(&tmap_w, w_full, smem_w + kb * w_kb_bytes, kb * kBlockK, n0);
      uint32_t it = 0;
      for (int mt = group; mt < m_tiles; mt += groups) {
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&a_empty[s], ph ^ 1);
          mbar_expect_tx(&a_full[s], kATileBytes);
          if (nchw.num_levels > 0) {
            const int row = mt / nchw.tiles_per_row, t = mt % nchw.tiles_per_row;
            int l = 0;
            while (l + 1 < nchw.num_levels && t >= nchw.tile_start[l + 1]) ++l;
            const int s0 = (t - nchw.tile_start[l]) * kBlockM;
            tma_load_3d(&nchw.map[l], &a_full[s], smem_a + s * kATileBytes, s0, kb * kBlockK, row);
            tma_load_3d(&nchw.map[l], &a_full[s], smem_a + s * kATileBytes + kATileBytes / 2, s0 + kBlockM / 2,
                        kb * kBlockK, row);
          } else {
            tma_load_2d(&tmap_a, &a_full[s], smem_a + s * kATileBytes, kb * kBlockK, mt * kBlockM);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const bool a_mn = nchw.num_levels > 0;
      const uint32_t idesc = make_idesc_bf16(NC, a_mn);
      const uint64_t a_step = a_mn ? (2048u >> 4) : 2u;   // start-address advance per K=16 step
      mbar_wait(w_full, 0);
      uint32_t it = 0, i = 0;
      for (int mt = group; mt < m_tiles; mt += groups, ++i) {
        const uint32_t buf = i & 1;
        mbar_wait(&acc_empty[buf], ((i >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + buf * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&a_full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_u32(smem_a + s * kATileBytes);
          const uint64_t da = a_mn ? make_smem_desc_sw128_mn(a_addr) : make_smem_desc_sw128(a_addr);
          const uint64_t db = make_smem_desc_sw128(smem_u32(smem_w + kb * w_kb_bytes));
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k)
            umma_bf16(tmem_d, da + a_step * static_cast<uint64_t>(k), db + static_cast<uint64_t>(2 * k),
                      idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&a_empty[s]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int e = warp - 2;
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    uint32_t i = 0;
    if (use_tma_store) {
      // TMEM -> registers -> swizzled smem tile (32 rows x 128 B) -> cp.async.bulk.tensor store.
      constexpr int kUnitCols = 128 / static_cast<int>(sizeof(OutT));   // 64 bf16 / 32 fp32
      uint8_t* stage = smem_out + e * kStageOutBytes;
      const int n_units = NC / kUnitCols;
      const int u_begin = (e >> 2) == 0 ? 0 : (n_units + 1) / 2;
      const int u_end = (e >> 2) == 0 ? (n_units + 1) / 2 : n_units;
      for (int mt = group; mt < m_tiles; mt += groups, ++i) {
        const uint32_t buf = i & 1;
        mbar_wait(&acc_full[buf], (i >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row0 = mt * kBlockM + q * 32;
        const int row = row0 + lane;
        const bool keep = row < M && (row_mask == nullptr || row_mask[row] != 0);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kAccStride;
        for (int u = u_begin; u < u_end; ++u) {
          const int c0 = u * kUnitCols;
          if (n0 + c0 >= N) break;                 // warp-uniform
          // split mode (value / offset-logit projection of the pyramid): every hm_period units form
          // one decoder layer's 448 columns; its first 4 units (8 heads x 32 channels) go to the
          // head-major value tensor, the remaining 3 to the (rows, layers * 192) map G
          int hm_head = -1, out_col = n0 + c0;
          if (sizeof(OutT) == 2 && hm_period > 0) {
            const int ug = (n0 + c0) / kUnitCols, layer = ug / hm_period, w = ug % hm_period;
            if (w < 4) hm_head = layer * 8 + 2 * w;
            else out_col = layer * (hm_period - 4) * kUnitCols + (w - 4) * kUnitCols;
          }
          // the previous bulk store must have finished READING the staging tile
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int h = 0; h < kUnitCols / 32; ++h) {
            uint32_t r[32];
            tmem_ld16(taddr + static_cast<uint32_t>(c0 + h * 32), r);
            tmem_ld16(taddr + static_cast<uint32_t>(c0 + h * 32 + 16), r + 16);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float v[32];
#pragma unroll
            for (int t = 0; t < 32; ++t) v[t] = __uint_as_float(r[t]);
            if (bias != nullptr) {
#pragma unroll
              for (int t = 0; t < 32; t += 4) {
                const int col = n0 + c0 + h * 32 + t;
                if (col < N) {                     // N % 16 == 0: whole float4 in or out
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col));
                  v[t] += b4.x; v[t + 1] += b4.y; v[t + 2] += b4.z; v[t + 3] += b4.w;
                }
              }
            }
            if (relu) {
#pragma unroll
              for (int t = 0; t < 32; ++t) v[t] = fmaxf(v[t], 0.f);
            }
            if (!keep) {
#pragma unroll
              for (int t = 0; t < 32; ++t) v[t] = 0.f;
            }
            uint8_t* srow = stage + lane * 128;
            if constexpr (sizeof(OutT) == 2) {
#pragma unroll
              for (int t = 0; t < 4; ++t) {        // 4 chunks of 8 bf16 (fp16 in split mode)
                uint4 o;
                if (hm_period > 0) {               // warp-uniform: the gather's maps are fp16
                  o.x = pack_f16x2(v[8 * t], v[8 * t + 1]);     o.y = pack_f16x2(v[8 * t + 2], v[8 * t + 3]);
                  o.z = pack_f16x2(v[8 * t + 4], v[8 * t + 5]); o.w = pack_f16x2(v[8 * t + 6], v[8 * t + 7]);
                } else {
                  o.x = pack_bf16x2(v[8 * t], v[8 * t + 1]);     o.y = pack_bf16x2(v[8 * t + 2], v[8 * t + 3]);
                  o.z = pack_bf16x2(v[8 * t + 4], v[8 * t + 5]); o.w = pack_bf16x2(v[8 * t + 6], v[8 * t + 7]);
                }
                if (hm_head >= 0) {
                  // head-major value rows: one 32-row x 64-byte tile per head, SWIZZLE_64B
                  *reinterpret_cast<uint4*>(stage + h * 2048 + lane * 64 + ((t ^ ((lane >> 1) & 3)) << 4)) = o;
                } else {
                  const int j = h * 4 + t;
                  *reinterpret_cast<uint4*>(srow + ((j ^ (lane & 7)) << 4)) = o;
                }
              }
            } else {
#pragma unroll
              for (int t = 0; t < 8; ++t) {        // 8 chunks of 4 fp32
                const float4 o = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
                *reinterpret_cast<float4*>(srow + ((t ^ (lane & 7)) << 4)) = o;
              }
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (hm_head >= 0) {
              tma_store_3d(&tmap_hm, stage, 0, row0, hm_head);
              tma_store_3d(&tmap_hm, stage + 2048, 0, row0, hm_head + 1);
            } else {
              tma_store_2d(&tmap_out, stage, out_col, row0);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
    const int n_c16 = NC / 16;
    const int c_begin = (e >> 2) == 0 ? 0 : (n_c16 + 1) / 2;
    const int c_end = (e >> 2) == 0 ? (n_c16 + 1) / 2 : n_c16;
    for (int mt = group; mt < m_tiles; mt += groups, ++i) {
      const uint32_t buf = i & 1;
      mbar_wait(&acc_full[buf], (i >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = mt * kBlockM + q * 32 + lane;
      const bool row_ok = row < M;
      const bool keep = row_ok && (row_mask == nullptr || row_mask[row] != 0);
      OutT* orow = out + static_cast<int64_t>(row) * ldo + n0;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kAccStride;
      for (int cc = c_begin; cc < c_end; ++cc) {
        const int c = cc * 16;
        if (n0 + c >= N) break;                    // warp-uniform
        uint32_t r[16];
        tmem_ld16(taddr + static_cast<uint32_t>(c), r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float v[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) v[t] = __uint_as_float(r[t]);
        if (bias != nullptr) {
#pragma unroll
          for (int t = 0; t < 16; t += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c + t));
            v[t] += b4.x; v[t + 1] += b4.y; v[t + 2] += b4.z; v[t + 3] += b4.w;
          }
        }
        if (relu) {
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] = fmaxf(v[t], 0.f);
        }
        if (!keep) {
#pragma unroll
          for (int t = 0; t < 16; ++t) v[t] = 0.f;
        }
        if (row_ok) {
          if constexpr (sizeof(OutT) == 2) {
            uint4 o0, o1;
            o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);
            o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
            o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]);
            o1.z = pack_bf16x2(v[12], v[13]); o1.w = pack_bf16x2(v[14], v[15]);
            uint4* dst = reinterpret_cast<uint4*>(orow + c);
            dst[0] = o0;
            dst[1] = o1;
          } else {
            float4* dst = reinterpret_cast<float4*>(orow + c);
#pragma unroll
            for (int t = 0; t < 4; ++t)
              dst[t] = make_float4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
                 : "memory");
  }
}

// Output tensor map for the epilogue's bulk stores: (rows, cols) row-major with row stride ldo,
// box = (128 bytes of columns, 32 rows), 128-byte swizzle (matches the staging tile).
static int make_out_tmap(CUtensorMap* map, const void* ptr, int64_t rows, int cols, int64_t ldo,
                         int elem_size) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return MVG_ELAUNCH;
  }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ldo) * elem_size};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elem_size), 32};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, elem_size == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                   2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(out) failed (%d) rows=%lld cols=%d ldo=%lld", static_cast<int>(r),
              static_cast<long long>(rows), cols, static_cast<long long>(ldo));
    return MVG_ELAUNCH;
  }
  return MVG_OK;
}

// 3-D bf16 tensor (heads, rows, 32) for the head-major value store: box = (32 ch, 32 rows, 1 head),
// 64-byte swizzle (matches the epilogue's per-head staging tiles).
static int make_hm_tmap(CUtensorMap* map, const void* ptr, int64_t rows, int heads) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return MVG_ELAUNCH;
  }
  const cuuint64_t dims[3] = {32, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(heads)};
  const cuuint64_t strides[2] = {64, static_cast<cuuint64_t>(rows) * 64};
  const cuuint32_t box[3] = {32, 32, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(head-major) failed (%d) rows=%lld heads=%d", static_cast<int>(r),
              static_cast<long long>(rows), heads);
    return MVG_ELAUNCH;
  }
  return MVG_OK;
}

// `value_hm` != nullptr selects the split mode: `out` is then the G map (M, (Nout/448)*192) and the
// value columns go to value_hm ((Nout/448)*8, M, 32).
template <typename OutT>
static int launch_linear(const void* A, const void* W, const float* bias, const uint8_t* row_mask,
                         void* out, int64_t M, int Nout, int K, int64_t ldo, int relu,
                         cudaStream_t st, void* value_hm = nullptr, const NchwA* nchw_a = nullptr) {
  // weight chunk width: as wide as fits 128 KB / one UMMA (256), split evenly over the chunks
  int nc_max = kMaxWBytes / (K * 2);
  nc_max = nc_max > 256 ? 256 : (nc_max / 16) * 16;
  if (nc_max < 16) {
    set_error("mvg_linear_bf16: K=%d too large for the weight-stationary kernel", K);
    return MVG_EUNSUPPORTED;
  }
  const int n_chunks = (Nout + nc_max - 1) / nc_max;
  // bulk-store epilogue works in units of 128 B of columns; narrow outputs use direct stores
  constexpr int kUnit = 128 / static_cast<int>(sizeof(OutT));
  const int use_tma_store = (Nout >= kUnit && nc_max >= kUnit) ? 1 : 0;
  const int gran = use_tma_store ? kUnit : 16;
  int NC = (((Nout + n_chunks - 1) / n_chunks) + gran - 1) / gran * gran;
  if (NC > nc_max) NC = nc_max;
  if (n_chunks > kNumSMs) {
    set_error("mvg_linear_bf16: Nout=%d needs %d weight chunks (> %d SMs)", Nout, n_chunks, kNumSMs);
    return MVG_EUNSUPPORTED;
  }
  const int m_tiles = static_cast<int>((M + kBlockM - 1) / kBlockM);
  int groups = kNumSMs / n_chunks;
  if (groups > m_tiles) groups = m_tiles;
  CUtensorMap ta, tw;
  int rc = make_tmap(&tw, W, Nout, K, NC);
  if (rc) return rc;
  static const NchwA kNoNchw{};                      // num_levels = 0: row-major A
  if (nchw_a == nullptr) {
    rc = make_tmap(&ta, A, M, K, kBlockM);
    if (rc) return rc;
    nchw_a = &kNoNchw;
  } else {
    ta = tw;                                         // unused by the kernel
  }
  CUtensorMap tout, thm;
  int hm_period = 0;
  if (value_hm != nullptr) {
    hm_period = 7;                                   // 448 columns per layer = 7 units of 64
    const int layers = Nout / 448;
    rc = make_out_tmap(&tout, out, M, layers * 192, ldo, 2);
    if (rc) return rc;
    rc = make_hm_tmap(&thm, value_hm, M, layers * 8);
    if (rc) return rc;
  } else if (use_tma_store) {
    rc = make_out_tmap(&tout, out, M, Nout, ldo, static_cast<int>(sizeof(OutT)));
    if (rc) return rc;
    thm = tout;  // unused by the kernel
  } else {
    tout = ta;   // unused by the kernel
    thm = ta;
  }
  auto kern = linear_tcgen05_ws_kernel<OutT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesV2);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d): %s", kSmemBytesV2, cudaGetErrorString(e));
      return MVG_ELAUNCH;
    }
    attr_set = true;
  }
  launch_k(kern, dim3(n_chunks * groups), dim3(kGemmThreadsV2), kSmemBytesV2, st, ta, *nchw_a, tw, tout, thm, hm_period,
           use_tma_store, bias, row_mask, static_cast<OutT*>(out), static_cast<int>(M), Nout, K, NC, n_chunks, groups,
           ldo, relu);
  return check_launch("mvg_linear_bf16");
}

}  // namespace mvg

extern "C" int mvg_linear_bf16(const void* A, const void* W, const float* bias, void* out,
                               int out_dtype, int64_t M, int Nout, int K, int64_t ldo, int relu,
                               const uint8_t* row_mask, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(A && W && out, "mvg_linear_bf16: null pointer");
  MVG_REQUIRE(M > 0 && M < (1ll << 31) && Nout > 0 && K > 0, "mvg_linear_bf16: empty shape");
  MVG_REQUIRE(K % kBlockK == 0, "mvg_linear_bf16: K=%d must be a multiple of %d", K, kBlockK);
  MVG_REQUIRE(Nout % 16 == 0, "mvg_linear_bf16: Nout=%d must be a multiple of 16", Nout);
  MVG_REQUIRE(ldo >= Nout, "mvg_linear_bf16: ldo < Nout");
  MVG_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "mvg_linear_bf16: operands must be 16-byte aligned");
  MVG_REQUIRE(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
              "mvg_linear_bf16: bias must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (out_dtype == MVG_BF16) {
    MVG_REQUIRE(ldo % 8 == 0, "mvg_linear_bf16: ldo must be a multiple of 8 for bf16 output");
    return launch_linear<__nv_bfloat16>(A, W, bias, row_mask, out, M, Nout, K, ldo, relu, st);
  } else if (out_dtype == MVG_F32) {
    MVG_REQUIRE(ldo % 4 == 0, "mvg_linear_bf16: ldo must be a multiple of 4 for fp32 output");
    return launch_linear<float>(A, W, bias, row_mask, out, M, Nout, K, ldo, relu, st);
  }
  set_error("mvg_linear_bf16: unsupported out dtype %d", out_dtype);
  return MVG_EUNSUPPORTED;
}


// 3-D bf16 map over one NCHW pyramid level (rows, 256, HW): dims {HW, 256, rows}, box {64 texels, 64 channels, 1}.
static int make_nchw_level_tmap(CUtensorMap* map, const void* ptr, int64_t hw, int64_t rows) {
  mvg::EncodeTiledFn enc = mvg::get_encode_fn();
  if (!enc) {
    mvg::set_error("cuTensorMapEncodeTiled entry point not available");
    return MVG_ELAUNCH;
  }
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(hw), 256, static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(hw) * 2, static_cast<cuuint64_t>(hw) * 512};
  const cuuint32_t box[3] = {64, static_cast<cuuint32_t>(mvg::kBlockK), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mvg::set_error("cuTensorMapEncodeTiled(nchw level) failed (%d) hw=%lld rows=%lld", static_cast<int>(r),
                   static_cast<long long>(hw), static_cast<long long>(rows));
    return MVG_ELAUNCH;
  }
  return MVG_OK;
}

extern "C" int mvg_value_proj_gemm_nchw_supported(int num_levels, const int* level_hw) {
  if (level_hw == nullptr || num_levels < 1 || num_levels > MVG_MAX_LEVELS) return 0;
  for (int l = 0; l < num_levels; ++l)
    if (level_hw[l] <= 0 || level_hw[l] % mvg::kBlockM != 0) return 0;
  return 1;
}

extern "C" int mvg_value_proj_gemm_nchw(const void* const* src_levels, int num_levels, const int* level_hw,
                                        int rows, const void* W, const float* bias, int layers, void* value_hm,
                                        void* gmap, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(src_levels && level_hw && W && value_hm && gmap, "mvg_value_proj_gemm_nchw: null pointer");
  MVG_REQUIRE(mvg_value_proj_gemm_nchw_supported(num_levels, level_hw),
              "mvg_value_proj_gemm_nchw: every level needs H*W %% %d == 0 (use mvg_pyramid_to_channels_last + "
              "mvg_value_proj_gemm otherwise)", kBlockM);
  MVG_REQUIRE(rows > 0 && layers > 0 && layers * 448 <= 148 * 256, "mvg_value_proj_gemm_nchw: bad shape");
  NchwA na{};
  na.num_levels = num_levels;
  int64_t S = 0;
  for (int l = 0; l < num_levels; ++l) {
    MVG_REQUIRE(src_levels[l] != nullptr && (reinterpret_cast<uintptr_t>(src_levels[l]) & 15) == 0,
                "mvg_value_proj_gemm_nchw: level %d pointer null / not 16-byte aligned", l);
    na.tile_start[l] = static_cast<int>(S / kBlockM);
    int rc = make_nchw_level_tmap(&na.map[l], src_levels[l], level_hw[l], rows);
    if (rc) return rc;
    S += level_hw[l];
  }
  na.tile_start[num_levels] = static_cast<int>(S / kBlockM);
  na.tiles_per_row = static_cast<int>(S / kBlockM);
  const int64_t M = static_cast<int64_t>(rows) * S;
  MVG_REQUIRE(M < (1ll << 31), "mvg_value_proj_gemm_nchw: too many rows");
  const void* ptrs[] = {W, value_hm, gmap};
  for (const void* q : ptrs)
    MVG_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, "mvg_value_proj_gemm_nchw: operands must be 16-byte aligned");
  MVG_REQUIRE(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
              "mvg_value_proj_gemm_nchw: bias must be 16-byte aligned");
  return launch_linear<__nv_bfloat16>(nullptr, W, bias, nullptr, gmap, M, layers * 448, 256,
                                      static_cast<int64_t>(layers) * 192, 0, static_cast<cudaStream_t>(stream),
                                      value_hm, &na);
}

extern "C" int mvg_value_proj_gemm(const void* feat, const void* W, const float* bias, int64_t M, int layers,
                                   void* value_hm, void* gmap, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(feat && W && value_hm && gmap, "mvg_value_proj_gemm: null pointer");
  MVG_REQUIRE(M > 0 && M < (1ll << 31) && layers > 0 && layers * 448 <= 148 * 256,
              "mvg_value_proj_gemm: bad shape (M=%lld layers=%d)", static_cast<long long>(M), layers);
  const void* ptrs[] = {feat, W, value_hm, gmap};
  for (const void* q : ptrs)
    MVG_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0, "mvg_value_proj_gemm: operands must be 16-byte aligned");
  MVG_REQUIRE(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0,
              "mvg_value_proj_gemm: bias must be 16-byte aligned");
  return launch_linear<__nv_bfloat16>(feat, W, bias, nullptr, gmap, M, layers * 448, 256,
                                      static_cast<int64_t>(layers) * 192, 0, static_cast<cudaStream_t>(stream),
                                      value_hm);
}
