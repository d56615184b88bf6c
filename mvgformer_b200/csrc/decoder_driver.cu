// C-ABI drivers of the whole hot path: one decoder layer, the L-layer decoder, and the final
// all-gather of the query-sharded mode - no torch, no Python, caller-provided workspace.
//
//   mvg_decoder_layer   DQDecoderLayer.forward, eval, indices=None   lib/models/dq_decoder.py:850-1045
//   mvg_decoder         DQDecoder.forward, return_intermediate=True  lib/models/dq_decoder.py:1107-1172
//   mvg_allgather_poses the exchange step of SURVEY.md section 8e (one NCCL all-gather per call)
//
// The launch sequence is the one mvgformer_b200/dq_decoder.py issues through the per-kernel entry
// points (tests/test_cabi_driver.py holds the two to bit-equality):
//   per call : [NCHW pyramid -> channels-last bf16]  value | G maps of all layers (one tcgen05 GEMM)
//   per layer: with_pos_embed + cast, qproj GEMM, mvg_project_sample_fused, output_proj (+ bounding
//              mask), masked view mean, fused feature update (mvg_ffn_chain), class head, query
//              selection, fused offset-net MLP on the selected rows (mvg_offset_chain),
//              offsets -> undistort -> DLT -> zero-fill scatter.
// Nothing is allocated or synchronised here; every kernel is enqueued on `stream`.
#include <dlfcn.h>

#include "common.cuh"

namespace mvg {

static inline int64_t up256(int64_t x) { return (x + 255) / 256 * 256; }

struct DecoderWs {
  uint8_t* feat_cl;      // (V*B, S, 256) bf16 (only when the pyramid arrives as NCHW levels)
  uint8_t* value_hm;     // (L*8, V*B*S, 32) fp16
  uint8_t* gmap;         // (V*B*S, L*192) fp16
  uint8_t* q_bf;         // (B, N, 256) bf16
  float* qproj;          // (B, N, 192)
  uint8_t* sampled;      // (B, V, N, 256) bf16
  float* ref2d;          // (B, V, N, 2)
  uint8_t* bounding;     // (B, V, N)
  uint8_t* attn;         // (B, V, N, 256) bf16
  uint8_t* aver;         // (B, N, 256) bf16
  float* mlp_out;        // (B*V*N, 4)
  int64_t* ids;          // 4 x (B*Q): batch / query ids, padded and reverse (mvg_select_pad)
  uint8_t* selected;     // (B, Q)
  int32_t* counts;       // (B)
  int32_t* info;         // (4)
  uint8_t* gather_ws;    // mvg_project_sample_workspace_bytes
  int64_t total;
};

static int64_t layout(const MvgDecoderConfig& c, bool nchw, void* base, DecoderWs* w, MvgSampleParams* prm) {
  const int64_t B = c.batch, V = c.views, N = static_cast<int64_t>(c.queries) * c.joints, L = c.layers;
  int64_t S = 0;
  prm->batch = c.batch; prm->views = c.views; prm->points = static_cast<int>(N); prm->num_levels = c.num_levels;
  for (int l = 0; l < MVG_MAX_LEVELS; ++l) {
    prm->level_h[l] = l < c.num_levels ? c.level_h[l] : 0;
    prm->level_w[l] = l < c.num_levels ? c.level_w[l] : 0;
    prm->level_start[l] = l < c.num_levels ? static_cast<int>(S) : 0;
    if (l < c.num_levels) S += static_cast<int64_t>(c.level_h[l]) * c.level_w[l];
  }
  prm->spatial_size = static_cast<int>(S);
  prm->ld_g = static_cast<int>(L * 192);
  prm->img_w = c.img_w; prm->img_h = c.img_h;
  prm->value_head_stride = V * B * S * 32;
  uint8_t* b = static_cast<uint8_t*>(base);
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off = up256(off + bytes); return b ? b + o : nullptr; };
  const int64_t rows = V * B * S;
  w->feat_cl = nchw ? take(rows * 256 * 2) : nullptr;
  w->value_hm = take(L * 8 * rows * 32 * 2);
  w->gmap = take(rows * L * 192 * 2);
  w->q_bf = take(B * N * 256 * 2);
  w->qproj = reinterpret_cast<float*>(take(B * N * 192 * 4));
  w->sampled = take(B * V * N * 256 * 2);
  w->ref2d = reinterpret_cast<float*>(take(B * V * N * 2 * 4));
  w->bounding = take(B * V * N);
  w->attn = take(B * V * N * 256 * 2);
  w->aver = take(B * N * 256 * 2);
  w->mlp_out = reinterpret_cast<float*>(take(B * V * N * 4 * 4));
  w->ids = reinterpret_cast<int64_t*>(take(4 * B * c.queries * 8));
  w->selected = take(B * c.queries);
  w->counts = reinterpret_cast<int32_t*>(take(4 * B));
  w->info = reinterpret_cast<int32_t*>(take(16));
  w->gather_ws = take(mvg_project_sample_workspace_bytes(prm));
  w->total = off;
  return off;
}

static int check_config(const MvgDecoderConfig* c, const char* who) {
  MVG_REQUIRE(c != nullptr, "%s: null config", who);
  MVG_REQUIRE(c->batch > 0 && c->views > 0 && c->views <= MVG_MAX_VIEWS && c->queries > 0 && c->joints > 0 &&
                  c->layers > 0, "%s: bad shape (B=%d V=%d Q=%d J=%d L=%d)", who, c->batch, c->views, c->queries,
              c->joints, c->layers);
  MVG_REQUIRE(c->num_levels >= 1 && c->num_levels <= MVG_MAX_LEVELS, "%s: num_levels %d", who, c->num_levels);
  MVG_REQUIRE(c->d_ffn >= 256 && c->d_ffn % 256 == 0, "%s: d_ffn %d must be a multiple of 256", who, c->d_ffn);
  MVG_REQUIRE((static_cast<int64_t>(c->queries) * c->joints * 256) % 8 == 0, "%s: Q*J*256 must be a multiple of 8", who);
  return MVG_OK;
}

// One layer on pre-projected maps.  value_hm / gmap point at THIS layer's 8 heads / 192 columns.
static int run_layer(const MvgDecoderConfig& c, const MvgSampleParams& prm, const MvgLayerWeights& w,
                     const void* value_hm, const void* gmap, const float* cams, const float* tgt,
                     const float* query_pos, const float* ref3d, const DecoderWs& ws, float* tgt_out,
                     float* ref_out, float* refined_out, float* projs_out, float* prob_out, int32_t* count_out,
                     void* stream) {
  const int64_t B = c.batch, V = c.views, N = static_cast<int64_t>(c.queries) * c.joints;
  int rc;
#define MVG_TRY(call) do { rc = (call); if (rc != MVG_OK) return rc; } while (0)
  MVG_TRY(mvg_add_cast_bf16(tgt, query_pos, ws.q_bf, B * N * 256, stream));                        // with_pos_embed
  MVG_TRY(mvg_linear_bf16(ws.q_bf, w.w_q, w.b_q, ws.qproj, MVG_F32, B * N, 192, 256, 192, 0, nullptr, stream));
  MVG_TRY(mvg_project_sample_fused(ref3d, cams, value_hm, gmap, ws.qproj, &prm, ws.sampled, ws.ref2d, ws.bounding,
                                   nullptr, ws.gather_ws, stream));
  MVG_TRY(mvg_linear_bf16(ws.sampled, w.w_o, w.b_o, ws.attn, MVG_BF16, B * V * N, 256, 256, 256, 0, ws.bounding, stream));
  MVG_TRY(mvg_masked_view_mean(ws.attn, ws.bounding, c.batch, c.views, static_cast<int>(N), 256, ws.aver, stream));
  MVG_TRY(mvg_ffn_chain(ws.aver, tgt, w.w_fu, w.b_fu, w.g2, w.e2, w.eps2, w.w1, w.b1, w.w2, w.b2, w.g3, w.e3, w.eps3,
                        B * N, c.d_ffn, tgt_out, stream));
  MVG_TRY(mvg_class_head(tgt_out, w.wc, w.bc, c.batch, c.queries, c.joints, prob_out, stream));
  MVG_TRY(mvg_select_pad(prob_out, c.batch, c.queries, c.threshold, c.filter_query ? 0 : 1, c.local_min_one,
                         ws.selected, ws.counts, ws.info, ws.ids, ws.ids + B * c.queries, ws.ids + 2 * B * c.queries,
                         ws.ids + 3 * B * c.queries, stream));
  MVG_TRY(mvg_offset_chain(ws.attn, ws.info, ws.ids + B * c.queries, ws.ids + 2 * B * c.queries,
                           ws.ids + 3 * B * c.queries, w.w_m1, w.b_m1,
                           w.w_m2, w.b_m2, w.w_m3, w.b_m3, c.batch, c.views, c.queries, c.joints, ws.mlp_out, 4, stream));
  MVG_TRY(mvg_offsets_dlt(ws.mlp_out, 4, ws.ref2d, ws.selected, cams, c.batch, c.views, c.queries, c.joints, c.img_w,
                          c.img_h, ref_out, refined_out, projs_out, stream));
  if (count_out != nullptr) {
    cudaError_t e = cudaMemcpyAsync(count_out, ws.info, sizeof(int32_t), cudaMemcpyDeviceToDevice,
                                    static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) {
      set_error("mvg_decoder_layer: cudaMemcpyAsync: %s", cudaGetErrorString(e));
      return MVG_ELAUNCH;
    }
  }
#undef MVG_TRY
  return MVG_OK;
}

// ---- pack / unpack of the query-sharded result: [poses (B, Qmax*J, 3) | prob (B, Qmax, 2) | counts (L)] fp32
__global__ void pack_result_kernel(const float* __restrict__ poses, const float* __restrict__ prob,
                                   const int32_t* __restrict__ counts, int B, int ql, int ql_max, int J, int L,
                                   float* __restrict__ buf) {
  const int64_t n_pose = static_cast<int64_t>(B) * ql_max * J * 3, n_prob = static_cast<int64_t>(B) * ql_max * 2;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n_pose) {
    const int64_t b = i / (static_cast<int64_t>(ql_max) * J * 3), r = i % (static_cast<int64_t>(ql_max) * J * 3);
    buf[i] = r < static_cast<int64_t>(ql) * J * 3 ? poses[b * ql * J * 3 + r] : 0.f;
  } else if (i < n_pose + n_prob) {
    const int64_t k = i - n_pose, b = k / (ql_max * 2), r = k % (ql_max * 2);
    buf[i] = r < ql * 2 ? prob[b * ql * 2 + r] : 0.f;
  } else if (i < n_pose + n_prob + L) {
    buf[i] = static_cast<float>(counts[i - n_pose - n_prob]);
  }
}
__global__ void unpack_result_kernel(const float* __restrict__ all, int world, int B, int Q, int ql_max, int J, int L,
                                     float* __restrict__ poses, float* __restrict__ prob, float* __restrict__ counts) {
  const int64_t n_pose = static_cast<int64_t>(B) * ql_max * J * 3, n_prob = static_cast<int64_t>(B) * ql_max * 2;
  const int64_t per = n_pose + n_prob + L;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t n_out_pose = static_cast<int64_t>(B) * Q * J * 3, n_out_prob = static_cast<int64_t>(B) * Q * 2;
  const int base = Q / world, rem = Q % world;           // contiguous blocks, sizes differ by at most one
  auto owner = [&](int q, int& q0) {
    int r = q < (base + 1) * rem ? q / (base + 1) : rem + (q - (base + 1) * rem) / max(base, 1);
    q0 = r * base + min(r, rem);
    return r;
  };
  if (i < n_out_pose) {
    const int64_t b = i / (static_cast<int64_t>(Q) * J * 3), r = i % (static_cast<int64_t>(Q) * J * 3);
    const int q = static_cast<int>(r / (J * 3));
    int q0;
    const int rk = owner(q, q0);
    poses[i] = all[rk * per + b * ql_max * J * 3 + (r - static_cast<int64_t>(q0) * J * 3)];
  } else if (i < n_out_pose + n_out_prob) {
    const int64_t k = i - n_out_pose, b = k / (Q * 2), r = k % (Q * 2);
    const int q = static_cast<int>(r / 2);
    int q0;
    const int rk = owner(q, q0);
    prob[k] = all[rk * per + n_pose + b * ql_max * 2 + (r - q0 * 2)];
  } else if (i < n_out_pose + n_out_prob + L) {
    const int l = static_cast<int>(i - n_out_pose - n_out_prob);
    float s = 0.f;
    for (int rk = 0; rk < world; ++rk) s += all[rk * per + n_pose + n_prob + l];
    counts[l] = s;
  }
}

// ncclAllGather of the NCCL library the HOST application already uses (torch's bundled copy when the host
// is PyTorch), resolved at run time: the communicator handle is only meaningful to that library instance.
using NcclAllGatherFn = int (*)(const void*, void*, size_t, int /*ncclDataType_t*/, void* /*ncclComm_t*/, cudaStream_t);
static NcclAllGatherFn resolve_allgather() {
  static NcclAllGatherFn fn = []() -> NcclAllGatherFn {
    void* p = dlsym(RTLD_DEFAULT, "ncclAllGather");
    if (p == nullptr) {
      void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (h != nullptr) p = dlsym(h, "ncclAllGather");
    }
    return reinterpret_cast<NcclAllGatherFn>(p);
  }();
  return fn;
}

}  // namespace mvg

extern "C" int64_t mvg_decoder_workspace_bytes(const MvgDecoderConfig* cfg, int pyramid_is_nchw) {
  using namespace mvg;
  if (check_config(cfg, "mvg_decoder_workspace_bytes") != MVG_OK) return -1;
  DecoderWs w;
  MvgSampleParams prm;
  return layout(*cfg, pyramid_is_nchw != 0, nullptr, &w, &prm);
}

extern "C" int mvg_decoder_layer(const MvgDecoderConfig* cfg, const MvgLayerWeights* w, const void* value_hm,
                                 const void* gmap, int ld_g, int64_t value_head_stride, const float* cams,
                                 const float* tgt, const float* query_pos, const float* ref3d, float* tgt_out,
                                 float* ref_out, float* refined_abs, float* projs_abs, float* class_prob,
                                 int32_t* selected_count, void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace mvg;
  int rc = check_config(cfg, "mvg_decoder_layer");
  if (rc != MVG_OK) return rc;
  MVG_REQUIRE(w && value_hm && gmap && cams && tgt && ref3d && tgt_out && ref_out && refined_abs && projs_abs &&
                  class_prob && workspace, "mvg_decoder_layer: null pointer");
  MVG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "mvg_decoder_layer: workspace must be 256-byte aligned");
  DecoderWs ws;
  MvgSampleParams prm;
  const int64_t need = layout(*cfg, false, workspace, &ws, &prm);
  MVG_REQUIRE(workspace_bytes >= need, "mvg_decoder_layer: workspace %lld < %lld bytes",
              static_cast<long long>(workspace_bytes), static_cast<long long>(need));
  prm.ld_g = ld_g;
  prm.value_head_stride = value_head_stride;
  return run_layer(*cfg, prm, *w, value_hm, gmap, cams, tgt, query_pos, ref3d, ws, tgt_out, ref_out, refined_abs,
                   projs_abs, class_prob, selected_count, stream);
}

extern "C" int mvg_decoder(const MvgDecoderConfig* cfg, const MvgLayerWeights* layers, const void* w_vg_all,
                           const float* b_vg_all, const void* const* pyramid_levels, int pyramid_dtype,
                           const void* pyramid_cl, const float* cams, const float* tgt, const float* query_pos,
                           const float* ref3d, float* hs, float* refs, float* refs2d, float* projs2d,
                           float* class_probs, int32_t* selected_counts, void* workspace, int64_t workspace_bytes,
                           void* stream) {
  using namespace mvg;
  int rc = check_config(cfg, "mvg_decoder");
  if (rc != MVG_OK) return rc;
  MVG_REQUIRE(layers && w_vg_all && cams && tgt && ref3d && hs && refs && refs2d && projs2d && class_probs && workspace,
              "mvg_decoder: null pointer");
  MVG_REQUIRE((pyramid_levels != nullptr) != (pyramid_cl != nullptr),
              "mvg_decoder: pass either the NCHW levels or the channels-last pyramid");
  MVG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "mvg_decoder: workspace must be 256-byte aligned");
  const bool nchw = pyramid_levels != nullptr;
  DecoderWs ws;
  MvgSampleParams prm;
  const int64_t need = layout(*cfg, nchw, workspace, &ws, &prm);
  MVG_REQUIRE(workspace_bytes >= need, "mvg_decoder: workspace %lld < %lld bytes",
              static_cast<long long>(workspace_bytes), static_cast<long long>(need));
  const int64_t B = cfg->batch, V = cfg->views, N = static_cast<int64_t>(cfg->queries) * cfg->joints, L = cfg->layers;
  const int64_t rows = V * B * prm.spatial_size;
  const void* feat = pyramid_cl;
  int hw[MVG_MAX_LEVELS];
  for (int l = 0; l < cfg->num_levels; ++l) hw[l] = cfg->level_h[l] * cfg->level_w[l];
  if (nchw && pyramid_dtype == MVG_BF16 && mvg_value_proj_gemm_nchw_supported(cfg->num_levels, hw)) {
    // bf16 levels of 128-texel multiples: the GEMM reads them in place (MN-major TMA operand)
    rc = mvg_value_proj_gemm_nchw(pyramid_levels, cfg->num_levels, hw, static_cast<int>(V * B), w_vg_all, b_vg_all,
                                  cfg->layers, ws.value_hm, ws.gmap, stream);
  } else {
    if (nchw) {
      rc = mvg_pyramid_to_channels_last(pyramid_levels, pyramid_dtype, cfg->num_levels, hw, static_cast<int>(V * B),
                                        256, ws.feat_cl, stream);
      if (rc != MVG_OK) return rc;
      feat = ws.feat_cl;
    }
    rc = mvg_value_proj_gemm(feat, w_vg_all, b_vg_all, rows, cfg->layers, ws.value_hm, ws.gmap, stream);
  }
  if (rc != MVG_OK) return rc;
  const float* tgt_l = tgt;
  const float* ref_l = ref3d;
  for (int l = 0; l < L; ++l) {
    float* tgt_o = hs + l * B * N * 256;
    float* ref_o = refs + l * B * N * 3;
    rc = run_layer(*cfg, prm, layers[l], ws.value_hm + static_cast<int64_t>(l) * 8 * prm.value_head_stride * 2,
                   ws.gmap + static_cast<int64_t>(l) * 192 * 2, cams, tgt_l, query_pos, ref_l, ws, tgt_o, ref_o,
                   refs2d + l * B * V * N * 2, projs2d + l * B * V * N * 2, class_probs + l * B * cfg->queries * 2,
                   selected_counts ? selected_counts + l : nullptr, stream);
    if (rc != MVG_OK) return rc;
    tgt_l = tgt_o;
    ref_l = ref_o;
  }
  return MVG_OK;
}

extern "C" int64_t mvg_allgather_poses_workspace_bytes(int batch, int queries, int joints, int layers, int world) {
  if (batch <= 0 || queries <= 0 || joints <= 0 || layers < 0 || world <= 0) return -1;
  const int64_t ql_max = (queries + world - 1) / world;
  const int64_t per = static_cast<int64_t>(batch) * ql_max * joints * 3 + static_cast<int64_t>(batch) * ql_max * 2 + layers;
  return (per * (world + 1) * 4 + 255) / 256 * 256;
}

extern "C" int mvg_allgather_poses(void* nccl_comm, int rank, int world, const float* poses_local,
                                   const float* prob_local, const int32_t* counts_local, int batch, int queries,
                                   int joints, int layers, float* poses, float* prob, float* counts,
                                   void* workspace, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(poses_local && prob_local && poses && prob && workspace, "mvg_allgather_poses: null pointer");
  MVG_REQUIRE(world >= 1 && rank >= 0 && rank < world && batch > 0 && queries >= world && joints > 0 && layers >= 0,
              "mvg_allgather_poses: bad shape (rank %d / %d, B=%d Q=%d)", rank, world, batch, queries);
  MVG_REQUIRE(layers == 0 || (counts_local && counts), "mvg_allgather_poses: counts missing");
  MVG_REQUIRE(world == 1 || nccl_comm != nullptr, "mvg_allgather_poses: null communicator");
  const int base = queries / world, rem = queries % world;
  const int ql = base + (rank < rem ? 1 : 0), ql_max = base + (rem ? 1 : 0);
  const int64_t per = static_cast<int64_t>(batch) * ql_max * joints * 3 + static_cast<int64_t>(batch) * ql_max * 2 + layers;
  float* send = static_cast<float*>(workspace);
  float* all = send + per;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  pack_result_kernel<<<static_cast<unsigned>((per + 255) / 256), 256, 0, st>>>(poses_local, prob_local, counts_local,
                                                                               batch, ql, ql_max, joints, layers, send);
  int rc = check_launch("mvg_allgather_poses(pack)");
  if (rc != MVG_OK) return rc;
  if (world == 1) {
    cudaError_t e = cudaMemcpyAsync(all, send, per * 4, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) { set_error("mvg_allgather_poses: %s", cudaGetErrorString(e)); return MVG_ELAUNCH; }
  } else {
    NcclAllGatherFn ag = resolve_allgather();
    MVG_REQUIRE(ag != nullptr, "mvg_allgather_poses: ncclAllGather not found in the process (load NCCL first)");
    const int nccl_rc = ag(send, all, static_cast<size_t>(per), 7 /* ncclFloat32 */, nccl_comm, st);
    if (nccl_rc != 0) { set_error("mvg_allgather_poses: ncclAllGather returned %d", nccl_rc); return MVG_ELAUNCH; }
  }
  const int64_t n_out = static_cast<int64_t>(batch) * queries * joints * 3 + static_cast<int64_t>(batch) * queries * 2 + layers;
  unpack_result_kernel<<<static_cast<unsigned>((n_out + 255) / 256), 256, 0, st>>>(all, world, batch, queries, ql_max,
                                                                                   joints, layers, poses, prob, counts);
  return check_launch("mvg_allgather_poses(unpack)");
}
