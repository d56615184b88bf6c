// Integer path of the query filter: generate_valid_masks + padding_query_with_mask
// (lib/models/dq_decoder.py:596-656), on device, no host round trip.
//
// The reference does: torch.where (row-major order) -> bincount -> .max() [host sync] ->
// Python-side repeat/cat -> stable sort by frame.  The result is fully determined by the
// per-frame compaction of the mask, so one CTA computes it with ballot/popc scans:
//   batch_ids/query_ids : (B, max_count) row-major; frame b holds its selected query ids in
//                         ascending order, then padding entries with query id 0 (:629-644)
//   *_rev               : frame b contributes count[b] entries (b, 0..count[b]-1) (:646-654)
// "Always one query" rule (:620-623): if nothing is selected, (frame 0, query 0) is.
#include "common.cuh"

namespace mvg {

__global__ void __launch_bounds__(1024)
select_pad_kernel(const float* __restrict__ prob, int B, int Q, float thr, int method, int min_one,
                  uint8_t* __restrict__ selected, int32_t* __restrict__ counts,
                  int32_t* __restrict__ info, int64_t* __restrict__ bids,
                  int64_t* __restrict__ qids, int64_t* __restrict__ brev,
                  int64_t* __restrict__ qrev) {
  extern __shared__ int s_dyn[];      // [0,B): count, [B,2B): prefix
  __shared__ int s_warp[32];
  __shared__ int s_maxc;
  int* s_count = s_dyn;
  int* s_prefix = s_dyn + B;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_enter();

  // pass 1: mask + per-frame counts
  for (int b = 0; b < B; ++b) {
    int cnt = 0;
    for (int q0 = 0; q0 < Q; q0 += blockDim.x) {
      const int q = q0 + tid;
      bool pred = false;
      if (q < Q) {
        const float* pr = prob + (static_cast<int64_t>(b) * Q + q) * 2;
        pred = (method == 0) ? (pr[1] > thr) : (pr[0] > 0.f);
        selected[static_cast<int64_t>(b) * Q + q] = pred ? 1 : 0;
      }
      cnt += __syncthreads_count(pred);
    }
    if (tid == 0) s_count[b] = cnt;
  }
  __syncthreads();
  if (tid == 0) {
    int total = 0;
    for (int b = 0; b < B; ++b) total += s_count[b];
    if (total == 0 && min_one) {    // dq_decoder.py:620-623
      selected[0] = 1;
      s_count[0] = 1;
      total = 1;
    }
    int mx = 0, run = 0;
    for (int b = 0; b < B; ++b) {
      s_prefix[b] = run;
      run += s_count[b];
      mx = max(mx, s_count[b]);
      counts[b] = s_count[b];
    }
    s_maxc = mx;
    info[0] = total;
    info[1] = mx;
    info[2] = 0;
    info[3] = 0;
  }
  __syncthreads();
  if (bids == nullptr) return;
  const int maxc = s_maxc;

  // pass 2: ordered compaction per frame
  for (int b = 0; b < B; ++b) {
    int running = 0;
    for (int q0 = 0; q0 < Q; q0 += blockDim.x) {
      const int q = q0 + tid;
      const bool pred = (q < Q) && selected[static_cast<int64_t>(b) * Q + q];
      const unsigned bal = __ballot_sync(0xffffffffu, pred);
      const int lane_pre = __popc(bal & ((1u << lane) - 1u));
      if (lane == 0) s_warp[warp] = __popc(bal);
      __syncthreads();
      int warp_pre = 0, chunk_total = 0;
      for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) {
        const int c = s_warp[w];
        if (w < warp) warp_pre += c;
        chunk_total += c;
      }
      if (pred) {
        const int pos = running + warp_pre + lane_pre;
        bids[static_cast<int64_t>(b) * maxc + pos] = b;
        qids[static_cast<int64_t>(b) * maxc + pos] = q;
        brev[s_prefix[b] + pos] = b;
        qrev[s_prefix[b] + pos] = pos;
      }
      running += chunk_total;
      __syncthreads();
    }
    for (int i = s_count[b] + tid; i < maxc; i += blockDim.x) {   // padding, query id 0
      bids[static_cast<int64_t>(b) * maxc + i] = b;
      qids[static_cast<int64_t>(b) * maxc + i] = 0;
    }
  }
}

}  // namespace mvg

extern "C" int mvg_select_pad(const float* prob, int batch, int queries, float threshold,
                              int method, int min_one, uint8_t* selected, int32_t* counts, int32_t* info,
                              int64_t* batch_ids, int64_t* query_ids, int64_t* batch_ids_rev,
                              int64_t* query_ids_rev, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(prob && selected && counts && info, "mvg_select_pad: null pointer");
  MVG_REQUIRE(batch > 0 && queries > 0 && batch <= 4096, "mvg_select_pad: bad shape B=%d Q=%d", batch, queries);
  MVG_REQUIRE(method == 0 || method == 1, "mvg_select_pad: method %d", method);
  const bool ids = batch_ids || query_ids || batch_ids_rev || query_ids_rev;
  MVG_REQUIRE(!ids || (batch_ids && query_ids && batch_ids_rev && query_ids_rev),
              "mvg_select_pad: id arrays must be all set or all NULL");
  launch_k(select_pad_kernel, dim3(1), dim3(1024), 2 * batch * sizeof(int), static_cast<cudaStream_t>(stream), prob, batch,
           queries, threshold, method, min_one, selected, counts, info, batch_ids, query_ids, batch_ids_rev,
           query_ids_rev);
  return check_launch("mvg_select_pad");
}
