// The steps either side of the decoder, on the device so that
// `model.forward` minus the backbone needs no host round trip:
//
//   mvg_init_queries           lib/models/dq_transformer.py:394-432 (person_joint embeddings ->
//                              query_pos | tgt) and :298-323 (`sample_space` roots + T-pose,
//                              norm2absolute lib/models/multi_view_pose_transformer.py:575-580)
//   mvg_assemble_predictions   lib/models/dq_transformer.py:568 (inverse_sigmoid,
//                              lib/models/util/misc.py:608-612) + lib/core/function.py:386-392
//                              pred = [x, y, z, (score > thr) - 1, score]; plus the score filter
//                              `pred[pred[:, 0, 3] >= 0]` of run/validate_3d.py:229 as an ordered
//                              list of surviving query ids
//   mvg_nearby_joints_nms      lib/core/nms.py:210-284 (combined_input=True, max_dets=-1)
//
// The NMS is integer work on top of fp32 distances; every fp32 operation is issued un-contracted
// in numpy's order so that the kept indices are bit-identical to the reference's.
#include "common.cuh"

namespace mvg {

// ------------------------------------------------------------------ query construction
// one block per (query q, joint j) row; thread t adds float4 t of the 2C-wide embeddings
__global__ void __launch_bounds__(128)
init_queries_kernel(const float* __restrict__ joint_emb, const float* __restrict__ inst_emb,
                    const float* __restrict__ lin, const double* __restrict__ tpose, float sx, float sy,
                    float sz, float cx, float cy, float cz, int batch, int queries, int joints,
                    int channels, int grid_n, float* __restrict__ query_pos, float* __restrict__ tgt,
                    float* __restrict__ ref) {
  const int row = blockIdx.x;              // q * J + j
  const int q = row / joints, j = row % joints;
  const int64_t N = static_cast<int64_t>(queries) * joints;
  const int c4 = channels / 4;             // float4 per half
  for (int t = threadIdx.x; t < 2 * c4; t += blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(joint_emb + static_cast<int64_t>(j) * 2 * channels) + t);
    const float4 b = __ldg(reinterpret_cast<const float4*>(inst_emb + static_cast<int64_t>(q) * 2 * channels) + t);
    const float4 s = make_float4(fadd(a.x, b.x), fadd(a.y, b.y), fadd(a.z, b.z), fadd(a.w, b.w));
    float* dst = t < c4 ? query_pos : tgt;  // torch.split(query_embeds, c, dim=1): pos first, tgt second
    const int col = (t < c4 ? t : t - c4) * 4;
    for (int b_ = 0; b_ < batch; ++b_)
      *reinterpret_cast<float4*>(dst + (static_cast<int64_t>(b_) * N + row) * channels + col) = s;
  }
  if (threadIdx.x < 3) {
    const int d = threadIdx.x;
    // roots: meshgrid(x_, x_) 'ij', z = 0.5; only the first `queries` of grid_n^2 are used
    const float rn = d == 0 ? __ldg(lin + q / grid_n) : (d == 1 ? __ldg(lin + q % grid_n) : 0.5f);
    const float gs = d == 0 ? sx : (d == 1 ? sy : sz);
    const float gc = d == 0 ? cx : (d == 1 ? cy : cz);
    const float ra = fsub(fadd(fmul(rn, gs), gc), fdiv(gs, 2.0f));   // norm2absolute
    const float v = static_cast<float>(__dadd_rn(static_cast<double>(ra), __ldg(tpose + j * 3 + d)));
    for (int b_ = 0; b_ < batch; ++b_) ref[(static_cast<int64_t>(b_) * N + row) * 3 + d] = v;
  }
}

// ------------------------------------------------------------------ prediction assembly
// one block per frame: pred rows + ordered compaction of the queries with score > threshold
__global__ void __launch_bounds__(1024)
assemble_pred_kernel(const float* __restrict__ poses, const float* __restrict__ prob, int queries,
                     int joints, float threshold, float* __restrict__ pred,
                     int* __restrict__ valid_ids, int* __restrict__ valid_count) {
  __shared__ int warp_cnt[32];
  __shared__ int running;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  for (int q0 = 0; q0 < queries; q0 += blockDim.x) {
    const int q = q0 + threadIdx.x;
    bool valid = false;
    if (q < queries) {
      // inverse_sigmoid (misc.py:608-612) then sigmoid (function.py:388)
      float x = __ldg(prob + (static_cast<int64_t>(b) * queries + q) * 2 + 1);
      x = fminf(fmaxf(x, 0.f), 1.f);
      const float x1 = fmaxf(x, 1e-5f), x2 = fmaxf(fsub(1.f, x), 1e-5f);
      const float logit = logf(fdiv(x1, x2));
      const float score = fdiv(1.f, fadd(1.f, expf(-logit)));
      valid = score > threshold;
      const float flag = valid ? 0.f : -1.f;
      const float* src = poses + (static_cast<int64_t>(b) * queries + q) * joints * 3;
      float* dst = pred + (static_cast<int64_t>(b) * queries + q) * joints * 5;
      for (int j = 0; j < joints; ++j) {
        dst[j * 5 + 0] = __ldg(src + j * 3 + 0);
        dst[j * 5 + 1] = __ldg(src + j * 3 + 1);
        dst[j * 5 + 2] = __ldg(src + j * 3 + 2);
        dst[j * 5 + 3] = flag;
        dst[j * 5 + 4] = score;
      }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = running;
      for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) {
        const int c = warp_cnt[w];
        warp_cnt[w] = s;
        s += c;
      }
      running = s;
    }
    __syncthreads();
    if (valid) valid_ids[static_cast<int64_t>(b) * queries + warp_cnt[warp] + __popc(m & ((1u << lane) - 1u))] = q;
    __syncthreads();
  }
  if (threadIdx.x == 0) valid_count[b] = running;
}

// ------------------------------------------------------------------ NMS, step 1: close-instance bits
// warp per row i (compact index): bit k of closebits[b][i] = #joints{ |kpt_i - kpt_k| < area_i * thr } > num_thr
constexpr int kNmsMaxJoints = 32;
__global__ void __launch_bounds__(256)
nms_close_kernel(const float* __restrict__ pred, const int* __restrict__ valid_ids,
                 const int* __restrict__ valid_count, int queries, int joints, float dist_thr,
                 int num_nearby_thr, uint32_t* __restrict__ closebits) {
  __shared__ float kp_i[8][kNmsMaxJoints * 3];
  const int b = blockIdx.y;
  const int n = valid_count[b];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 8 + warp;
  if (i >= n) return;                                    // whole warp
  const int words = (queries + 31) >> 5;
  const int* vid = valid_ids + static_cast<int64_t>(b) * queries;
  const float* pb = pred + static_cast<int64_t>(b) * queries * joints * 5;
  const float* pi = pb + static_cast<int64_t>(vid[i]) * joints * 5;
  for (int e = lane; e < joints * 3; e += 32) kp_i[warp][e] = __ldg(pi + (e / 3) * 5 + e % 3);
  __syncwarp();
  // pose "area": diagonal of the joint bounding box (nms.py:249-251), fp32 like numpy
  float mx[3], mn[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) mx[d] = mn[d] = kp_i[warp][d];
  for (int j = 1; j < joints; ++j) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      mx[d] = fmaxf(mx[d], kp_i[warp][j * 3 + d]);
      mn[d] = fminf(mn[d], kp_i[warp][j * 3 + d]);
    }
  }
  const float ax = fsub(mx[0], mn[0]), ay = fsub(mx[1], mn[1]), az = fsub(mx[2], mn[2]);
  const float area = __fsqrt_rn(fadd(fadd(fmul(ax, ax), fmul(ay, ay)), fmul(az, az)));
  const float thr = fmul(area, dist_thr);
  uint32_t* row = closebits + (static_cast<int64_t>(b) * queries + i) * words;
  for (int k0 = 0; k0 < n; k0 += 32) {
    const int k = k0 + lane;
    bool close = false;
    if (k < n) {
      const float* pk = pb + static_cast<int64_t>(vid[k]) * joints * 5;
      int cnt = 0;
      for (int j = 0; j < joints; ++j) {
        const float dx = fsub(kp_i[warp][j * 3 + 0], __ldg(pk + j * 5 + 0));
        const float dy = fsub(kp_i[warp][j * 3 + 1], __ldg(pk + j * 5 + 1));
        const float dz = fsub(kp_i[warp][j * 3 + 2], __ldg(pk + j * 5 + 2));
        const float dist = __fsqrt_rn(fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));
        cnt += dist < thr ? 1 : 0;
      }
      close = cnt > num_nearby_thr;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, close);
    if (lane == 0) row[k0 >> 5] = m;
  }
}

// ------------------------------------------------------------------ NMS, step 2: greedy pass
// one block per frame; descending-score order by counting, then one warp walks it (nms.py:263-272)
constexpr int kNmsMaxQ = 4096;
__global__ void __launch_bounds__(1024)
nms_greedy_kernel(const float* __restrict__ pred, const int* __restrict__ valid_ids,
                  const int* __restrict__ valid_count, int queries, int joints,
                  const uint32_t* __restrict__ closebits, int* __restrict__ keep_compact,
                  int* __restrict__ keep_query, int* __restrict__ keep_count) {
  __shared__ float score[kNmsMaxQ];
  __shared__ int order[kNmsMaxQ];
  __shared__ uint32_t ignored[kNmsMaxQ / 32];
  const int b = blockIdx.x;
  const int n = valid_count[b];
  const int words = (queries + 31) >> 5;
  const int* vid = valid_ids + static_cast<int64_t>(b) * queries;
  const float* pb = pred + static_cast<int64_t>(b) * queries * joints * 5;
  for (int i = threadIdx.x; i < n; i += blockDim.x) score[i] = __ldg(pb + static_cast<int64_t>(vid[i]) * joints * 5 + 4);
  for (int w = threadIdx.x; w < words; w += blockDim.x) ignored[w] = 0u;
  __syncthreads();
  // np.argsort(scores)[::-1]: descending; equal scores (the reference's sort is not stable, so
  // their order is unspecified there) are taken larger index first
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float s = score[i];
    int rank = 0;
    for (int k = 0; k < n; ++k) {
      const float t = score[k];
      rank += (t > s || (t == s && k > i)) ? 1 : 0;
    }
    order[rank] = i;
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  const uint32_t* cb = closebits + static_cast<int64_t>(b) * queries * words;
  int cnt = 0;
  const int nw = (n + 31) >> 5;
  for (int r = 0; r < n; ++r) {
    const int i = order[r];
    if ((ignored[i >> 5] >> (i & 31)) & 1u) continue;            // warp-uniform
    // keep_ind = keep_inds[argmax(scores[keep_inds])]: first maximum in index order
    float best = -INFINITY;
    int best_k = 0x7fffffff;
    for (int w = lane; w < nw; w += 32) {
      uint32_t m = __ldg(cb + static_cast<int64_t>(i) * words + w);
      while (m) {
        const int k = (w << 5) + __ffs(m) - 1;
        m &= m - 1;
        const float s = score[k];
        if (s > best || (s == best && k < best_k)) { best = s; best_k = k; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int ok = __shfl_xor_sync(0xffffffffu, best_k, o);
      if (ob > best || (ob == best && ok < best_k)) { best = ob; best_k = ok; }
    }
    // a pose is always close to itself, so best_k is valid.  The ignored bit is read by ONE lane
    // and broadcast, and the warp re-converges before any lane updates the bit set, so every lane
    // takes the same branch (no read / write race on `ignored` under independent thread scheduling).
    const uint32_t ign_word = __shfl_sync(0xffffffffu, ignored[best_k >> 5], 0);
    __syncwarp();
    if (!((ign_word >> (best_k & 31)) & 1u)) {
      if (lane == 0) {
        keep_compact[static_cast<int64_t>(b) * queries + cnt] = best_k;
        keep_query[static_cast<int64_t>(b) * queries + cnt] = vid[best_k];
      }
      ++cnt;
      for (int w = lane; w < nw; w += 32) ignored[w] |= __ldg(cb + static_cast<int64_t>(i) * words + w);
    }
    __syncwarp();
  }
  if (lane == 0) keep_count[b] = cnt;
}

}  // namespace mvg

extern "C" int mvg_init_queries(const float* joint_emb, const float* inst_emb, const float* lin,
                                const double* tpose, const float* space_size, const float* space_center,
                                int batch, int queries, int joints, int channels, int grid_n,
                                float* query_pos, float* tgt, float* ref, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(joint_emb && inst_emb && lin && tpose && space_size && space_center && query_pos && tgt && ref,
              "mvg_init_queries: null pointer");
  MVG_REQUIRE(batch > 0 && queries > 0 && joints > 0 && channels > 0 && channels % 4 == 0,
              "mvg_init_queries: bad shape (B=%d Q=%d J=%d C=%d)", batch, queries, joints, channels);
  MVG_REQUIRE(static_cast<int64_t>(grid_n) * grid_n >= queries, "mvg_init_queries: grid_n^2 < queries");
  init_queries_kernel<<<queries * joints, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      joint_emb, inst_emb, lin, tpose, space_size[0], space_size[1], space_size[2], space_center[0],
      space_center[1], space_center[2], batch, queries, joints, channels, grid_n, query_pos, tgt, ref);
  return check_launch("mvg_init_queries");
}

extern "C" int mvg_assemble_predictions(const float* poses, const float* prob, int batch, int queries,
                                        int joints, float threshold, float* pred, int32_t* valid_ids,
                                        int32_t* valid_count, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(poses && prob && pred && valid_ids && valid_count, "mvg_assemble_predictions: null pointer");
  MVG_REQUIRE(batch > 0 && queries > 0 && joints > 0, "mvg_assemble_predictions: bad shape");
  assemble_pred_kernel<<<batch, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
      poses, prob, queries, joints, threshold, pred, valid_ids, valid_count);
  return check_launch("mvg_assemble_predictions");
}

extern "C" int mvg_nearby_joints_nms(const float* pred, const int32_t* valid_ids, const int32_t* valid_count,
                                     int batch, int queries, int joints, float dist_thr,
                                     int num_nearby_joints_thr, uint32_t* workspace, int32_t* keep_compact,
                                     int32_t* keep_query, int32_t* keep_count, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(pred && valid_ids && valid_count && workspace && keep_compact && keep_query && keep_count,
              "mvg_nearby_joints_nms: null pointer");
  MVG_REQUIRE(batch > 0 && queries > 0 && queries <= kNmsMaxQ && joints > 0 && joints <= kNmsMaxJoints,
              "mvg_nearby_joints_nms: bad shape (Q <= %d, J <= %d)", kNmsMaxQ, kNmsMaxJoints);
  MVG_REQUIRE(dist_thr > 0.f, "mvg_nearby_joints_nms: `dist_thr` must be greater than 0.");   // nms.py:231
  MVG_REQUIRE(num_nearby_joints_thr < joints,
              "mvg_nearby_joints_nms: `num_nearby_joints_thr` must be less than the number of joints.");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  nms_close_kernel<<<dim3((queries + 7) / 8, batch), 256, 0, st>>>(pred, valid_ids, valid_count, queries, joints,
                                                                   dist_thr, num_nearby_joints_thr, workspace);
  int rc = check_launch("mvg_nearby_joints_nms(close)");
  if (rc != MVG_OK) return rc;
  nms_greedy_kernel<<<batch, 1024, 0, st>>>(pred, valid_ids, valid_count, queries, joints, workspace,
                                            keep_compact, keep_query, keep_count);
  return check_launch("mvg_nearby_joints_nms");
}
