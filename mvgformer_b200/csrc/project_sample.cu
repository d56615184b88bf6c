// Fused projection + projective-attention sampling.
//
//   a3  project_ref_points      lib/models/dq_decoder.py:331-397, lib/utils/cameras.py:167-207,
//                               lib/utils/transforms.py:135-141
//   a4  ProjAttn (non-GEMM part) lib/models/ops/modules/projattn.py:139-153 (ref-point feature
//                               lookup), :180-191 (offsets, softmax over Lv*P, locations, and the
//                               `.view` layout scramble of the per-level Linear outputs)
//   a5  deformable gather       lib/models/ops/src/cuda/deform_im2col_cuda.cuh:247-309, :41-93
//
// Two kernels per call:
//   project_compact_kernel  one thread per (frame, view, point): projection (non-contracted fp32
//       in the reference's op order -> `bounding` bit-exact), ref2d / bounding outputs, and an
//       ordered-per-block compaction of the IN-VIEW items.  The reference multiplies the
//       attention feature of an out-of-view point by 0 (dq_decoder.py:585-586) and uses it
//       nowhere else, so those items are not gathered at all (their `sampled` row is zeros).
//   gather_kernel           one warp per in-view item, one persistent CTA per SM owning a
//       contiguous slice of the compacted list (an SM works on ~one person at a time and the
//       overlapping sampling footprints share its L1).
//
// Restructuring vs the reference:
//   * The per-level Linear on (grid_sample(feat_l) + query) is split by linearity into
//     bilinear-sampling a pre-projected 192-channel map G = feat @ [W_off; W_attn]^T (written by
//     the same tcgen05 GEMM that produces `value`) plus a per-point term qproj = W (tgt+pos) + b.
//   * Phase B computes, once per (head, sample), the four bilinear*attention weights and the
//     base texel offset and stages them in shared memory; phase C is then a branch-free stream
//     of LDG.128 (lane = head*4 + chunk: the 4 lanes of a head fetch one 64-byte (texel, head)
//     row in ONE L1 request) + fp32 FMAs, with the bf16 -> fp32 conversion done by the tensor
//     pipe (see fma_corner).
//   * A tensor-pipe weighted sum (loaded bytes as the MMA A fragment, weights as a
//     block-diagonal B) was built and measured this round: it needs 4x fewer instructions but
//     its fragment-shaped loads split every 64-byte row over two quarter-warps, i.e. two L1
//     requests and two sector fills per row, and the kernel is L1 data-pipe bound - 478 us vs
//     345 us (profiles/gather_experiments_r1.md).
// Geometry and the sampling index path use non-contracted fp32 ops in the reference's op order.
#include "common.cuh"

namespace mvg {

#ifndef MVG_PS_WARPS
#define MVG_PS_WARPS 16
#endif
#ifndef MVG_PS_UNROLL
#define MVG_PS_UNROLL 4
#endif
#ifndef MVG_PS_MMA_UNPACK
#define MVG_PS_MMA_UNPACK 0        // 1: bf16 -> fp32 through the tensor pipe, 0: shift / mask ALU ops
#endif
constexpr int kWarps = MVG_PS_WARPS;   // warps per CTA; one persistent CTA per SM
constexpr int kPsUnroll = MVG_PS_UNROLL;
constexpr int kQP = 192;           // 128 offset channels + 64 logit channels per level
constexpr int kHeads = 8;
constexpr int kPcThreads = 256;    // project_compact block

// acc (two fp32 packed in a 64-bit register) += {w, w} * {bf16 lo, bf16 hi} of the 32-bit word u
// (Blackwell packed FFMA2: one issue slot for two channels).  Used by phase A only.
__device__ __forceinline__ void fma2_bf16pair(uint64_t& acc, uint32_t u, uint64_t ww) {
  uint64_t v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(u << 16), "r"(u & 0xffff0000u));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(v), "l"(ww));
}
__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b));
  return v;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ void fma2_corner(uint64_t (&acc)[4], const uint4& c, float w) {
  const uint64_t ww = pack2(w, w);
  fma2_bf16pair(acc[0], c.x, ww);
  fma2_bf16pair(acc[1], c.y, ww);
  fma2_bf16pair(acc[2], c.z, ww);
  fma2_bf16pair(acc[3], c.w, ww);
}
// acc[0..3] (8 fp32 channels as 4 packed pairs) += w * the 8 bf16 channels of c.
// MVG_PS_MMA_UNPACK: the bf16 -> fp32 conversion runs on the otherwise idle tensor pipe:
// D(16x8) = A(16x8) * I(8x8) with A's fragment = the lane's own words ({c.x, c.y}, then
// {c.z, c.w}) returns, in the same lane, D[g][2t..2t+1] = A[g][2t..2t+1] - the two halves of
// the first word as fp32 - and D[g+8][2t..2t+1] = those of the second word.  The products are
// exact (x * 1.0 + 0), and the results land in aligned register pairs that feed FFMA2
// directly: 2 HMMA + 4 FFMA2 per 16-byte load instead of 8 ALU + 4 FFMA2.  Measured (r1, layer-0
// call): 21 % fewer warp instructions, issue 56 -> 44 %, but 287 us vs 271 us - the kernel is
// bound by L1 latency / data-pipe wavefronts and the HMMA adds latency to every
// load -> accumulate chain.  Kept as an option, off by default.
__device__ __forceinline__ void fma_corner(uint64_t (&acc)[4], const uint4& c, float w, uint32_t b_ident) {
#if MVG_PS_MMA_UNPACK
  const uint64_t ww = pack2(w, w);
  uint64_t v0, v1, v2, v3;
  asm("{\n\t.reg .f32 d0, d1, d2, d3;\n\t"
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {d0, d1, d2, d3}, {%2, %3}, {%4}, {%5, %5, %5, %5};\n\t"
      "mov.b64 %0, {d0, d1};\n\tmov.b64 %1, {d2, d3};\n\t}"
      : "=l"(v0), "=l"(v1) : "r"(c.x), "r"(c.y), "r"(b_ident), "f"(0.f));
  asm("{\n\t.reg .f32 d0, d1, d2, d3;\n\t"
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {d0, d1, d2, d3}, {%2, %3}, {%4}, {%5, %5, %5, %5};\n\t"
      "mov.b64 %0, {d0, d1};\n\tmov.b64 %1, {d2, d3};\n\t}"
      : "=l"(v2), "=l"(v3) : "r"(c.z), "r"(c.w), "r"(b_ident), "f"(0.f));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[0]) : "l"(v0), "l"(ww));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[1]) : "l"(v1), "l"(ww));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[2]) : "l"(v2), "l"(ww));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[3]) : "l"(v3), "l"(ww));
#else
  fma2_corner(acc, c, w);
#endif
}

// ------------------------------------------------------------------ projection + compaction
// One block = 256 consecutive points of ONE (frame, view) pair bv = blockIdx.y.
// ws[bv] = number of in-view points of that pair (zeroed by the host wrapper before the launch),
// ws[hdr + bv*N ...] = their flat item indices bv*N + n, ordered inside each 256-point block
// (hdr = B*V rounded up to a multiple of 4).
__global__ void __launch_bounds__(kPcThreads)
project_compact_kernel(const float* __restrict__ ref3d, const MvgCamera* __restrict__ cams,
                       const MvgSampleParams prm, float* __restrict__ ref2d_out,
                       uint8_t* __restrict__ bounding_out, __nv_bfloat16* __restrict__ sampled,
                       int* __restrict__ ws) {
  __shared__ int warp_cnt[kPcThreads / 32];
  __shared__ int block_base;
  const int N = prm.points, V = prm.views;
  const int bv = blockIdx.y;
  const int n = blockIdx.x * kPcThreads + threadIdx.x;
  const int64_t item = static_cast<int64_t>(bv) * N + n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  bool inb = false;
  if (n < N) {
    const int b = bv / V;
    const MvgCamera* cam = cams + bv;   // (B,V) row-major == item / N
    const float* x3 = ref3d + (static_cast<int64_t>(b) * N + n) * 3;
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
    const float radial = fadd(1.f, fadd(fadd(fmul(cam->k[0], r2), fmul(cam->k[1], r4)),
                                        fmul(cam->k[2], r6)));
    const float tanv = fadd(fmul(cam->p[0], y1), fmul(cam->p[1], y0));
    const float corr = fadd(radial, fmul(2.f, tanv));
    y0 = fadd(fmul(y0, corr), fmul(cam->p[1], r2));
    y1 = fadd(fmul(y1, corr), fmul(cam->p[0], r2));
    float px = fadd(fmul(cam->f[0], y0), cam->c[0]);
    float py = fadd(fmul(cam->f[1], y1), cam->c[1]);
    inb = (px >= 0.f) && (py >= 0.f) && (px < cam->wh[0]) && (py < cam->wh[1]);
    px = fminf(fmaxf(px, -1.f), cam->clamp_max);
    py = fminf(fmaxf(py, -1.f), cam->clamp_max);
    const float ax = fadd(fadd(fmul(px, cam->aff[0]), fmul(py, cam->aff[1])), cam->aff[2]);
    const float ay = fadd(fadd(fmul(px, cam->aff[3]), fmul(py, cam->aff[4])), cam->aff[5]);
    *reinterpret_cast<float2*>(ref2d_out + 2 * item) =
        make_float2(fdiv(ax, prm.img_w), fdiv(ay, prm.img_h));
    bounding_out[item] = inb ? 1 : 0;
  }
  const uint32_t in_mask = __ballot_sync(0xffffffffu, inb);
  // out-of-view rows of `sampled` are defined (zeros); a warp writes one 512-byte row at a time
  uint32_t out_mask = __ballot_sync(0xffffffffu, n < N && !inb);
  const int64_t item0 = item - lane;
  while (out_mask) {
    const int src = __ffs(out_mask) - 1;
    out_mask &= out_mask - 1;
    reinterpret_cast<uint4*>(sampled + (item0 + src) * 256)[lane] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (lane == 0) warp_cnt[warp] = __popc(in_mask);
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < kPcThreads / 32; ++w) {
      const int c = warp_cnt[w];
      warp_cnt[w] = s;
      s += c;
    }
    block_base = s > 0 ? atomicAdd(ws + bv, s) : 0;
  }
  __syncthreads();
  if (inb) {
    const int pos = block_base + warp_cnt[warp] + __popc(in_mask & ((1u << lane) - 1u));
    const int hdr = (prm.batch * V + 3) & ~3;
    ws[hdr + static_cast<int64_t>(bv) * N + pos] = static_cast<int>(item);
  }
}

// ------------------------------------------------------------------ gather
template <int LV> struct WarpScratch {
  float proj[LV][kQP];                    // per pyramid level: Linear outputs (offsets | logits)
  float4 cw[LV * 8 * kHeads];             // [sample][head]: weights * attention as {w00, w10, w01, w11}
                                          // (block column dx = 0 pair, then dx = 1 pair)
  int base[LV * 8 * kHeads];              // [sample][head]: 16-byte offset of the clamped 2x2 block
};

template <int LV>
__global__ void __launch_bounds__(kWarps * 32, 1)
gather_kernel(const __nv_bfloat16* __restrict__ value_hm, const __nv_bfloat16* __restrict__ gmap,
              const float* __restrict__ qproj, const MvgSampleParams prm, __nv_bfloat16* __restrict__ sampled,
              const float* __restrict__ ref2d, const float* __restrict__ refl_in,
              const int* __restrict__ ws) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  WarpScratch<LV>& sc = reinterpret_cast<WarpScratch<LV>*>(smem_dyn)[warp];
  const int N = prm.points, V = prm.views, B = prm.batch;
  const int ldg = prm.ld_g;            // row stride of the offset / logit map G, elements
  constexpr int NS = LV * 8;           // samples per head
  // phase-B ownership: head m = lane & 7, sample group sub = lane >> 3 (samples r = sub + 4 i):
  // a quarter-warp then stores 8 consecutive float4 slots [r][0..7] (conflict-free).
  // phase-C ownership: lane = hsel * 8 + dx * 4 + chunk.  The value tensor is head-major
  // ([head][row][32 ch], 64 B per (texel, head)), so the two horizontal corners of a bilinear
  // footprint are 128 contiguous bytes: the 8 lanes (dx, chunk) of a quarter-warp fetch them with
  // ONE L1 request, and one LDG.128 covers a block row of 4 heads (hsel).  Two passes (hg) cover
  // the 8 heads, two loads (block rows dy) the footprint.
  const int m = lane & 7, sub = lane >> 3;
  const int hsel = lane >> 3, dxl = (lane >> 2) & 1, subc = lane & 3;
  // identity B fragment of mma.m16n8k8 (tensor-pipe unpack option): B[k][n] = (k == n), lane
  // (g, t) = (lane >> 2, lane & 3) holds k = 2t, 2t+1 of column n = g as a bf16 pair
  const uint32_t b_ident = (lane >> 2) == 2 * (lane & 3) ? 0x00003F80u
                                                         : ((lane >> 2) == 2 * (lane & 3) + 1 ? 0x3F800000u : 0u);
  // View-major sweep: for every (frame, view) pair in turn, each CTA takes an equal slice of that
  // pair's in-view list and its 16 warps walk it together (warp w: first + w, first + w + 16, ...).
  // All SMs therefore read ONE view's value / G maps at a time (36 MB, L2-resident) instead of all
  // views at once (180 MB > 126 MB L2), and inside a slice an SM still works on ~one person's
  // joints, whose overlapping footprints share its L1.
  const int BV = B * V;
  const int hdr = (BV + 3) & ~3;
#pragma unroll 1
  for (int bvi = 0; bvi < BV; ++bvi) {
  const int cnt = ws != nullptr ? __ldg(ws + bvi) : N;
  const int first = static_cast<int>(static_cast<int64_t>(cnt) * blockIdx.x / gridDim.x);
  const int last = static_cast<int>(static_cast<int64_t>(cnt) * (blockIdx.x + 1) / gridDim.x);
#pragma unroll 1
  for (int idx = first + warp; idx < last; idx += kWarps) {
    const int64_t item = ws != nullptr ? static_cast<int64_t>(__ldg(ws + hdr + static_cast<int64_t>(bvi) * N + idx))
                                       : static_cast<int64_t>(bvi) * N + idx;
    const int64_t out_idx = item;
    const int n = static_cast<int>(item % N);
    const int bv = static_cast<int>(item / N);
    const int v = bv % V, b = bv / V;
    const int64_t vrow0 = static_cast<int64_t>(v * B + b) * prm.spatial_size;     // first row of this view
    const __nv_bfloat16* grow = gmap + vrow0 * ldg;
    float refl_x[LV], refl_y[LV];
    if (refl_in != nullptr) {          // ProjAttn.forward entry: reference points are given
#pragma unroll
      for (int l = 0; l < LV; ++l) {
        refl_x[l] = __ldg(refl_in + (item * LV + l) * 2);
        refl_y[l] = __ldg(refl_in + (item * LV + l) * 2 + 1);
      }
    } else {
      const float2 r = __ldg(reinterpret_cast<const float2*>(ref2d + 2 * item));
#pragma unroll
      for (int l = 0; l < LV; ++l) {   // dq_decoder.py:570-573
        const float fW = static_cast<float>(prm.level_w[l]), fH = static_cast<float>(prm.level_h[l]);
        refl_x[l] = fdiv(fmul(r.x, fW), fsub(fW, 1.f));
        refl_y[l] = fdiv(fmul(r.y, fH), fsub(fH, 1.f));
      }
    }

    float inv_w[LV], inv_h[LV];
#pragma unroll
    for (int l = 0; l < LV; ++l) {
      inv_w[l] = 1.f / static_cast<float>(prm.level_w[l]);
      inv_h[l] = 1.f / static_cast<float>(prm.level_h[l]);
    }
    // ---------------- phase A (a4 i+iii): sample the pre-projected map G at the reference point
    if (lane < kQP / 8) {
      const float* qp = qproj + (static_cast<int64_t>(b) * N + n) * kQP + lane * 8;
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(qp));
      const float4 q1 = __ldg(reinterpret_cast<const float4*>(qp + 4));
      const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
      uint4 cn[LV][4];
      float cwgt[LV][4];
#pragma unroll
      for (int l = 0; l < LV; ++l) {
        const int W = prm.level_w[l], H = prm.level_h[l];
        // F.grid_sample(bilinear, zeros, align_corners=False): projattn.py:139-153
        const float gx = fminf(fmaxf(fsub(fmul(refl_x[l], 2.f), 1.f), -1.1f), 1.1f);
        const float gy = fminf(fmaxf(fsub(fmul(refl_y[l], 2.f), 1.f), -1.1f), 1.1f);
        const float ix = fsub(fmul(fadd(gx, 1.f), fmul(static_cast<float>(W), 0.5f)), 0.5f);
        const float iy = fsub(fmul(fadd(gy, 1.f), fmul(static_cast<float>(H), 0.5f)), 0.5f);
        const float fx0 = floorf(ix), fy0 = floorf(iy);
        const int x0 = static_cast<int>(fx0), yy0 = static_cast<int>(fy0);
        const float we = fsub(ix, fx0), ww = fsub(1.f, we);
        const float ws = fsub(iy, fy0), wn = fsub(1.f, ws);
        const bool okx0 = x0 >= 0 && x0 < W, okx1 = x0 + 1 >= 0 && x0 + 1 < W;
        const bool oky0 = yy0 >= 0 && yy0 < H, oky1 = yy0 + 1 >= 0 && yy0 + 1 < H;
        const int xa = min(max(x0, 0), W - 1), xb = min(max(x0 + 1, 0), W - 1);
        const int ya = min(max(yy0, 0), H - 1), yb = min(max(yy0 + 1, 0), H - 1);
        const __nv_bfloat16* gl = grow + static_cast<int64_t>(prm.level_start[l]) * ldg + lane * 8;
        cn[l][0] = ldg_nc_v4(gl + static_cast<int64_t>(ya * W + xa) * ldg);
        cn[l][1] = ldg_nc_v4(gl + static_cast<int64_t>(ya * W + xb) * ldg);
        cn[l][2] = ldg_nc_v4(gl + static_cast<int64_t>(yb * W + xa) * ldg);
        cn[l][3] = ldg_nc_v4(gl + static_cast<int64_t>(yb * W + xb) * ldg);
        cwgt[l][0] = (oky0 && okx0) ? wn * ww : 0.f;
        cwgt[l][1] = (oky0 && okx1) ? wn * we : 0.f;
        cwgt[l][2] = (oky1 && okx0) ? ws * ww : 0.f;
        cwgt[l][3] = (oky1 && okx1) ? ws * we : 0.f;
      }
#pragma unroll
      for (int l = 0; l < LV; ++l) {
        uint64_t a2[4] = {pack2(qv[0], qv[1]), pack2(qv[2], qv[3]), pack2(qv[4], qv[5]), pack2(qv[6], qv[7])};
#pragma unroll
        for (int c = 0; c < 4; ++c) fma2_corner(a2, cn[l][c], cwgt[l][c]);
        float r8[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) unpack2(a2[i], r8[2 * i], r8[2 * i + 1]);
        float4* dst = reinterpret_cast<float4*>(&sc.proj[l][lane * 8]);
        dst[0] = make_float4(r8[0], r8[1], r8[2], r8[3]);
        dst[1] = make_float4(r8[4], r8[5], r8[6], r8[7]);
      }
    }
    __syncwarp();

    // ---------------- phase B (a4 iv+v, a5 index path): per (head, sample) parameters.
    // 4 lanes per head, lane `sub` owns samples r = sub + 4 i.
    {
      float lg[NS / 4];
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < NS / 4; ++i) {
        const int g = m * NS + sub + 4 * i;          // flat logit index after the `.view`
        lg[i] = sc.proj[g >> 6][128 + (g & 63)];
        mx = fmaxf(mx, lg[i]);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < NS / 4; ++i) {
        lg[i] = expf(lg[i] - mx);
        sum += lg[i];
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 8);
      sum += __shfl_xor_sync(0xffffffffu, sum, 16);
      const float inv_sum = 1.f / sum;
#pragma unroll
      for (int i = 0; i < NS / 4; ++i) {
        const int r = sub + 4 * i;
        const int l = i >> 1;                        // == r >> 3 because sub < 4 (static index)
        const float wgt = lg[i] * inv_sum;
        const int f = m * (NS * 2) + 2 * r;          // flat offset index after the `.view`
        const float2 off = *reinterpret_cast<const float2*>(&sc.proj[f >> 7][f & 127]);
        const int W = prm.level_w[l], H = prm.level_h[l], start = prm.level_start[l];
        const float rlx = refl_x[l], rly = refl_y[l];
        const float fW = static_cast<float>(W), fH = static_cast<float>(H);
        // projattn.py:186-191, then deform_im2col_cuda.cuh:291-301 and :41-93
        // (reciprocal instead of the reference's division: the offsets come from the bf16 map,
        //  so this path is not bit-comparable anyway; mvg_deform_forward keeps exact inputs)
        const float loc_x = fadd(rlx, fmul(off.x, inv_w[l]));
        const float loc_y = fadd(rly, fmul(off.y, inv_h[l]));
        const float h_im = fsub(fmul(loc_y, fH), 0.5f);
        const float w_im = fsub(fmul(loc_x, fW), 0.5f);
        const bool inside = h_im > -1.f && w_im > -1.f && h_im < fH && w_im < fW;
        const float fh = floorf(h_im), fw = floorf(w_im);
        const int h_low = inside ? static_cast<int>(fh) : 0;
        const int w_low = inside ? static_cast<int>(fw) : 0;
        const float lh = inside ? h_im - fh : 0.f, lw = inside ? w_im - fw : 0.f;
        const float hh = 1.f - lh, hw = 1.f - lw;
        // The 2x2 texel block is clamped into the level ((ha, wa) .. (ha+1, wa+1) always exist,
        // H, W >= 2); a corner the reference skips (deform_im2col_cuda.cuh:57-80) gets weight 0
        // and the surviving row / column moves to the block row / column that holds its texel.
        const int ha = min(max(h_low, 0), H - 2), wa = min(max(w_low, 0), W - 2);
        const float ry0 = h_low < 0 ? lh : (h_low > H - 2 ? 0.f : hh);
        const float ry1 = h_low < 0 ? 0.f : (h_low > H - 2 ? hh : lh);
        const float rx0 = w_low < 0 ? lw : (w_low > W - 2 ? 0.f : hw);
        const float rx1 = w_low < 0 ? 0.f : (w_low > W - 2 ? hw : lw);
        const float sw = inside ? wgt : 0.f;
        sc.cw[r * kHeads + m] = make_float4(ry0 * rx0 * sw, ry1 * rx0 * sw, ry0 * rx1 * sw, ry1 * rx1 * sw);
        sc.base[r * kHeads + m] = (start + ha * W + wa) * 4;      // 16-byte units, 64 B per (texel, head)
      }
    }
    __syncwarp();

    // ---------------- phase C (a5): gather 2 block rows x NS samples x 2 head groups
    uint64_t acc2[2][4];                             // [head group][8 fp32 channels as 4 pairs]
#pragma unroll
    for (int hg = 0; hg < 2; ++hg) acc2[hg][0] = acc2[hg][1] = acc2[hg][2] = acc2[hg][3] = 0ull;
    const uint4* vl[2];
#pragma unroll
    for (int hg = 0; hg < 2; ++hg)
      vl[hg] = reinterpret_cast<const uint4*>(value_hm + (static_cast<int64_t>(hg * 4 + hsel) * prm.value_head_stride +
                                                          vrow0 * 32)) + dxl * 4 + subc;
#pragma unroll
    for (int l = 0; l < LV; ++l) {
      const uint32_t rowstep16 = static_cast<uint32_t>(prm.level_w[l]) * 4u;
#pragma unroll kPsUnroll
      for (int p = 0; p < 8; ++p) {
        const int r = l * 8 + p;
#pragma unroll
        for (int hg = 0; hg < 2; ++hg) {
          // this lane's block column: (top, bottom) weights
          const float2 cwv = reinterpret_cast<const float2*>(&sc.cw[r * kHeads + hg * 4 + hsel])[dxl];
          const uint4* p0 = vl[hg] + static_cast<uint32_t>(sc.base[r * kHeads + hg * 4 + hsel]);
          const uint4 ct = __ldg(p0);
          const uint4 cb = __ldg(p0 + rowstep16);
          fma_corner(acc2[hg], ct, cwv.x, b_ident);
          fma_corner(acc2[hg], cb, cwv.y, b_ident);
        }
      }
    }
    // left + right block columns: lanes (dx = 0) and (dx = 1) hold the two halves of every sum;
    // afterwards lane (hsel, dx, chunk) owns head 4 dx + hsel
    float acc[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a0, a1, b0, b1;
      unpack2(acc2[0][i], a0, a1);
      unpack2(acc2[1][i], b0, b1);
      a0 += __shfl_xor_sync(0xffffffffu, a0, 4); a1 += __shfl_xor_sync(0xffffffffu, a1, 4);
      b0 += __shfl_xor_sync(0xffffffffu, b0, 4); b1 += __shfl_xor_sync(0xffffffffu, b1, 4);
      acc[2 * i] = dxl ? b0 : a0;
      acc[2 * i + 1] = dxl ? b1 : a1;
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(sampled + out_idx * 256 + (dxl * 4 + hsel) * 32 + subc * 8) = o;
    __syncwarp();   // scratch is reused by the next item
  }
  }
}

}  // namespace mvg

extern "C" int mvg_project_sample_fused(const float* ref3d, const float* cams, const void* value_hm,
                                        const void* gmap, const float* qproj, const MvgSampleParams* prm,
                                        void* sampled, float* ref2d, uint8_t* bounding,
                                        const float* refl_in, void* workspace, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(value_hm && gmap && qproj && prm && sampled, "mvg_project_sample_fused: null pointer");
  MVG_REQUIRE((reinterpret_cast<uintptr_t>(value_hm) & 15) == 0 && (reinterpret_cast<uintptr_t>(gmap) & 15) == 0,
              "mvg_project_sample_fused: value / G map must be 16-byte aligned");
  MVG_REQUIRE(refl_in || (ref3d && cams && ref2d && bounding && workspace),
              "mvg_project_sample_fused: projection inputs/outputs/workspace missing");
  MVG_REQUIRE(prm->num_levels >= 1 && prm->num_levels <= MVG_MAX_LEVELS,
              "mvg_project_sample_fused: num_levels %d out of range", prm->num_levels);
  MVG_REQUIRE(prm->batch > 0 && prm->views > 0 && prm->points > 0, "mvg_project_sample_fused: empty shape");
  MVG_REQUIRE(prm->ld_g >= 192 && prm->ld_g % 8 == 0, "mvg_project_sample_fused: ld_g %d", prm->ld_g);
  MVG_REQUIRE(prm->value_head_stride >= static_cast<int64_t>(prm->batch) * prm->views * prm->spatial_size * 32 &&
                  prm->value_head_stride % 8 == 0,
              "mvg_project_sample_fused: value_head_stride %lld", static_cast<long long>(prm->value_head_stride));
  int s = 0;
  for (int l = 0; l < prm->num_levels; ++l) {
    MVG_REQUIRE(prm->level_h[l] > 1 && prm->level_w[l] > 1 && prm->level_start[l] == s,
                "mvg_project_sample_fused: level %d shape/start inconsistent", l);
    s += prm->level_h[l] * prm->level_w[l];
  }
  MVG_REQUIRE(s == prm->spatial_size, "mvg_project_sample_fused: spatial_size %d != sum H*W %d",
              prm->spatial_size, s);
  MVG_REQUIRE(static_cast<int64_t>(s) * 4 < (1ll << 31),
              "mvg_project_sample_fused: per-view map too large for 32-bit texel offsets");
  const int64_t total = static_cast<int64_t>(prm->batch) * prm->views * prm->points;
  MVG_REQUIRE(total < (1ll << 31), "mvg_project_sample_fused: too many items");
  const int64_t per_cta = kWarps * 4;   // at least ~4 items per warp before spreading further
  int64_t want = (total + per_cta - 1) / per_cta;
  const int grid = static_cast<int>(want < kNumSMs ? (want < 1 ? 1 : want) : kNumSMs);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const MvgCamera* cam = reinterpret_cast<const MvgCamera*>(cams);
  const __nv_bfloat16* vhm = static_cast<const __nv_bfloat16*>(value_hm);
  const __nv_bfloat16* gmp = static_cast<const __nv_bfloat16*>(gmap);
  __nv_bfloat16* sp = static_cast<__nv_bfloat16*>(sampled);
  int* ws = refl_in ? nullptr : static_cast<int*>(workspace);
  if (ws != nullptr) {
    const int hdr = (prm->batch * prm->views + 3) & ~3;
    cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(int) * hdr, st);
    if (e != cudaSuccess) {
      set_error("mvg_project_sample_fused: cudaMemsetAsync: %s", cudaGetErrorString(e));
      return MVG_ELAUNCH;
    }
    const dim3 pc_grid((prm->points + kPcThreads - 1) / kPcThreads, prm->batch * prm->views);
    project_compact_kernel<<<pc_grid, kPcThreads, 0, st>>>(ref3d, cam, *prm, ref2d, bounding, sp, ws);
    int rc = check_launch("mvg_project_sample_fused(project_compact)");
    if (rc != MVG_OK) return rc;
  }
#define MVG_LAUNCH_PS(LV)                                                                       \
  {                                                                                             \
    constexpr int smem = kWarps * static_cast<int>(sizeof(WarpScratch<LV>));                    \
    static bool attr_done = false;                                                              \
    if (!attr_done) {                                                                           \
      cudaError_t e = cudaFuncSetAttribute(gather_kernel<LV>,                                   \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, smem);  \
      if (e != cudaSuccess) {                                                                   \
        set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));                           \
        return MVG_ELAUNCH;                                                                     \
      }                                                                                         \
      attr_done = true;                                                                         \
    }                                                                                           \
    gather_kernel<LV><<<grid, kWarps * 32, smem, st>>>(vhm, gmp, qproj, *prm, sp, ref2d, refl_in, ws); \
  }
  switch (prm->num_levels) {
    case 1: MVG_LAUNCH_PS(1) break;
    case 2: MVG_LAUNCH_PS(2) break;
    case 3: MVG_LAUNCH_PS(3) break;
    default: MVG_LAUNCH_PS(4) break;
  }
#undef MVG_LAUNCH_PS
  return check_launch("mvg_project_sample_fused");
}
