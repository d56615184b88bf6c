// Dense projection  out = act(A @ W^T + bias)  on 5th-generation tensor cores (sm_100a).
//
//   A (M,K) bf16 row-major (activations / channels-last pyramid), W (Nout,K) bf16 row-major
//   (nn.Linear layout), fp32 accumulation in TMEM, bias + ReLU + bf16/fp32 conversion fused in
//   the epilogue.  Used for every nn.Linear on the hot path (kernel table):
//   rayconv / sampling_offsets / attention_weights on the pyramid (one launch for all L layers),
//   the per-point qproj, output_proj, feature_update_mlp, FFN and the offset_net MLP.
//
// Structure (one CTA = one 128 x BLOCK_N output tile, 6 warps, warp-specialised):
//   warp 0  TMA producer : cp.async.bulk.tensor (SWIZZLE_128B boxes of 64 K-elements) into a
//                          kStages-deep shared-memory ring, mbarrier expect_tx / complete_tx
//   warp 1  MMA issuer   : allocates TMEM, one thread issues tcgen05.mma.cta_group::1.kind::f16
//                          (M=128, N=BLOCK_N, K=16) from shared-memory descriptors, tcgen05.commit
//                          releases ring slots and finally signals the epilogue
//   warps 2-5 epilogue   : tcgen05.ld (32 lanes x 16 columns) -> +bias -> ReLU -> convert ->
//                          16-byte global stores; each warp owns the TMEM lane quarter warp%4
// Two CTAs fit per SM (<= 98 KB smem, <= 128 TMEM columns each), so one CTA's epilogue overlaps
// the other's main loop.  Tile order is N-fastest so the CTAs that share an A tile run together
// and A is fetched from HBM once.
#include <cuda.h>

#include "common.cuh"

namespace mvg {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;           // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B operand tile: rows of 128 B, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1
// [46,48), layout_type=2 (SWIZZLE_128B) [61,64)).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;                  // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;          // SBO: 8 rows x 128 B
  d |= static_cast<uint64_t>(1) << 46;                  // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                  // SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: D=f32, A=B=bf16, both K-major, M=128, N=n.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(kBlockM >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, "
      "%12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

template <int BLOCK_N> struct GemmCfg {
  static constexpr int kStages = (BLOCK_N >= 128) ? 3 : 4;
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = BLOCK_N < 32 ? 32 : BLOCK_N;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N, typename OutT>
__global__ void __launch_bounds__(kGemmThreads)
linear_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                      const __grid_constant__ CUtensorMap tmap_w, const float* __restrict__ bias,
                      OutT* __restrict__ out, int M, int N, int K, int64_t ldo, int relu) {
  using Cfg = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tmem_full_bar = bars + 2 * Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BLOCK_N;
  const int m0 = blockIdx.y * kBlockM;
  const int num_kb = K / kBlockK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation is warp-collective; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(static_cast<uint32_t>(Cfg::kTmemCols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % Cfg::kStages;
        const uint32_t ph = (kb / Cfg::kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
        uint8_t* sa = smem + s * Cfg::kStageBytes;
        tma_load_2d(&tmap_a, &full_bar[s], sa, kb * kBlockK, m0);
        tma_load_2d(&tmap_w, &full_bar[s], sa + Cfg::kABytes, kb * kBlockK, n0);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % Cfg::kStages;
        const uint32_t ph = (kb / Cfg::kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
        const uint64_t da = make_smem_desc_sw128(sa);
        const uint64_t db = make_smem_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          // advance 16 K-elements = 32 B inside the 128 B swizzle row: +2 in the >>4 field
          umma_bf16(tmem_base, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                    idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);            // frees the ring slot when these MMAs retire
      }
      umma_commit(tmem_full_bar);              // accumulator complete -> epilogue
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    mbar_wait(tmem_full_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int row = m0 + q * 32 + lane;
    OutT* orow = out + static_cast<int64_t>(row) * ldo + n0;
#pragma unroll 2
    for (int c = 0; c < BLOCK_N; c += 16) {
      if (n0 + c >= N) break;                  // warp-uniform
      uint32_t r[16];
      tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c), r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
      if (bias != nullptr) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c + i));
          v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
        }
      }
      if (relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      if (row < M) {
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
          for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(Cfg::kTmemCols))
                 : "memory");
  }
}

// ---- host side -----------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// 2-D bf16 row-major (rows, K) tensor, box = (kBlockK, box_rows), 128-byte swizzle.
static int make_tmap(CUtensorMap* map, const void* ptr, int64_t rows, int K, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return MVG_ELAUNCH;
  }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld K=%d box_rows=%d", static_cast<int>(r),
              static_cast<long long>(rows), K, box_rows);
    return MVG_ELAUNCH;
  }
  return MVG_OK;
}

template <int BLOCK_N, typename OutT>
static int launch_linear(const void* A, const void* W, const float* bias, void* out, int64_t M,
                         int Nout, int K, int64_t ldo, int relu, cudaStream_t st) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap ta, tw;
  int rc = make_tmap(&ta, A, M, K, kBlockM);
  if (rc) return rc;
  rc = make_tmap(&tw, W, Nout, K, BLOCK_N);
  if (rc) return rc;
  auto kern = linear_tcgen05_kernel<BLOCK_N, OutT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return MVG_ELAUNCH;
    }
    attr_set = true;
  }
  dim3 grid((Nout + BLOCK_N - 1) / BLOCK_N, static_cast<unsigned>((M + kBlockM - 1) / kBlockM));
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, st>>>(ta, tw, bias, static_cast<OutT*>(out),
                                                    static_cast<int>(M), Nout, K, ldo, relu);
  return check_launch("mvg_linear_bf16");
}

}  // namespace mvg

extern "C" int mvg_linear_bf16(const void* A, const void* W, const float* bias, void* out,
                               int out_dtype, int64_t M, int Nout, int K, int64_t ldo, int relu,
                               void* stream) {
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
    if (Nout <= 16) return launch_linear<16, __nv_bfloat16>(A, W, bias, out, M, Nout, K, ldo, relu, st);
    if (Nout % 128 != 0 && Nout < 512) return launch_linear<64, __nv_bfloat16>(A, W, bias, out, M, Nout, K, ldo, relu, st);
    return launch_linear<128, __nv_bfloat16>(A, W, bias, out, M, Nout, K, ldo, relu, st);
  } else if (out_dtype == MVG_F32) {
    MVG_REQUIRE(ldo % 4 == 0, "mvg_linear_bf16: ldo must be a multiple of 4 for fp32 output");
    if (Nout <= 16) return launch_linear<16, float>(A, W, bias, out, M, Nout, K, ldo, relu, st);
    if (Nout % 128 != 0 && Nout < 512) return launch_linear<64, float>(A, W, bias, out, M, Nout, K, ldo, relu, st);
    return launch_linear<128, float>(A, W, bias, out, M, Nout, K, ldo, relu, st);
  }
  set_error("mvg_linear_bf16: unsupported out dtype %d", out_dtype);
  return MVG_EUNSUPPORTED;
}
