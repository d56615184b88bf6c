// 2D offsets -> refined 2D points -> DLT triangulation, FOUR lanes per (frame, query, joint): the lanes
// split the views (view softmax, inverse affine, 5-iteration undistort, the two DLT rows per view),
// the 4x4 normal matrices are summed with warp shuffles, every lane then runs the same Jacobi
// eigen-solve (no exchange needed) and lane 0 stores.  One thread per problem left 94 % of the warp
// slots idle (15 360 problems on 148 SMs) with every view's ~400-cycle division chain in series.
//
//   a9  calculate_2d_offsets tail   lib/models/dq_decoder.py:678-707
//   a10 inverse affine + undistort   lib/models/dq_decoder.py:413-422, :119-204, P: :223-246
//   a11 DLT                          lib/mvn/utils/multiview.py:170-228, :72-86
//   scatter / zero-fill              lib/models/dq_decoder.py:1011-1029
//
// The reference gathers the selected queries into a padded rectangle, runs tiny torch ops on
// per-query replicated cameras and calls torch.linalg.svd once per query from Python.  Every
// (query, joint) problem is independent, so the select->pad->gather->compute->unpad->scatter
// sequence collapses to "compute where selected, write zeros elsewhere".
//
// Solver: the DLT solution is the right singular vector of the smallest singular value of the
// confidence-weighted (2V x 4) matrix A, i.e. the eigenvector of the smallest eigenvalue of
// A^T A.  A is built in fp32 exactly like the reference (multiview.py:195-203); A^T A is
// accumulated in fp64 (products of fp32 numbers are exact in fp64) and diagonalised by cyclic
// Jacobi in fp64.  Jacobi's graded backward error keeps the null vector accurate although
// column 4 of A is ~10^4 times larger than columns 1-3: measured 1e-9 mm from an fp64 SVD,
// versus 0.15-0.7 mm mean for the reference's own fp32 LAPACK SVD.
#include "common.cuh"

namespace mvg {

constexpr int kJacobiSweeps = 8;   // upper bound; the graded stopping test usually ends after 4-5 sweeps

// Smallest-eigenvalue eigenvector of the symmetric 4x4 `H` (upper triangle used).
__device__ __forceinline__ void smallest_eigvec4(double H[4][4], double out[4]) {
  double Vm[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Vm[i][j] = (i == j) ? 1.0 : 0.0;
  // parallel (round-robin) ordering: the two rotations of a round touch disjoint index pairs, so
  // their parameter chains (div, sqrt, rsqrt) are independent and overlap in the fp64 pipe
  constexpr int kPairs[6][2] = {{0, 1}, {2, 3}, {0, 2}, {1, 3}, {0, 3}, {1, 2}};
#pragma unroll 1
  for (int sweep = 0; sweep < kJacobiSweeps; ++sweep) {
    // graded stopping test |a_pq| <= eps sqrt(a_pp a_qq) for all pairs (the matrix spans 12 orders
    // of magnitude, a norm-wise test would stop too early for the small eigenvalue)
    bool done = true;
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      const int p = kPairs[e][0], q = kPairs[e][1];
      done = done && (H[p][q] * H[p][q] <= 1e-32 * fabs(H[p][p] * H[q][q]));
    }
    if (done) break;
#pragma unroll
    for (int round = 0; round < 3; ++round) {
      double cs[2], sn[2], tt[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int p = kPairs[2 * round + e][0], q = kPairs[2 * round + e][1];
        const double apq = H[p][q];
        const double theta = (H[q][q] - H[p][p]) / (2.0 * apq);
        double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        if (apq == 0.0) t = 0.0;                   // already diagonal in this plane: identity
        const double c = rsqrt(t * t + 1.0);
        cs[e] = c; sn[e] = t * c; tt[e] = t;
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int p = kPairs[2 * round + e][0], q = kPairs[2 * round + e][1];
        const double c = cs[e], s = sn[e], t = tt[e];
        const double apq = H[p][q];
        H[p][p] = H[p][p] - t * apq;
        H[q][q] = H[q][q] + t * apq;
        H[p][q] = 0.0;
        H[q][p] = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k != p && k != q) {
            const double hkp = H[k][p], hkq = H[k][q];
            const double np_ = c * hkp - s * hkq, nq_ = s * hkp + c * hkq;
            H[k][p] = np_; H[p][k] = np_;
            H[k][q] = nq_; H[q][k] = nq_;
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double vkp = Vm[k][p], vkq = Vm[k][q];
          Vm[k][p] = c * vkp - s * vkq;
          Vm[k][q] = s * vkp + c * vkq;
        }
      }
    }
  }
  int best = 0;
  double ev = H[0][0];
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (H[i][i] < ev) { ev = H[i][i]; best = i; }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double v = Vm[k][0];
#pragma unroll
    for (int i = 1; i < 4; ++i)
      if (best == i) v = Vm[k][i];
    out[k] = v;
  }
}

// Adds the two DLT rows of one view to the normal matrix (upper triangle incl. mirror).
__device__ __forceinline__ void accumulate_rows(const float* P /*3x4*/, float u, float v,
                                                float conf, double H[4][4]) {
  float r0[4], r1[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {   // A = p3 * pt; A -= P[:2]; A *= conf  (multiview.py:199-203)
    r0[c] = fmul(fsub(fmul(P[8 + c], u), P[c]), conf);
    r1[c] = fmul(fsub(fmul(P[8 + c], v), P[4 + c]), conf);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a; b < 4; ++b)
      H[a][b] += static_cast<double>(r0[a]) * static_cast<double>(r0[b]) +
                 static_cast<double>(r1[a]) * static_cast<double>(r1[b]);
}

__device__ __forceinline__ void solve_and_store(double H[4][4], float* out3) {
#pragma unroll
  for (int a = 1; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < a; ++b) H[a][b] = H[b][a];
  double x[4];
  smallest_eigvec4(H, x);
  // X = -Vh[3]; homogeneous_to_euclidean (multiview.py:72-86): the sign cancels
  out3[0] = static_cast<float>(x[0] / x[3]);
  out3[1] = static_cast<float>(x[1] / x[3]);
  out3[2] = static_cast<float>(x[2] / x[3]);
}

constexpr int kDltLanes = 4;       // lanes per (frame, query, joint) problem

__global__ void __launch_bounds__(128)
offsets_dlt_kernel(const float* __restrict__ mlp_out, int mlp_ld, const float* __restrict__ ref2d,
                   const uint8_t* __restrict__ selected, const MvgCamera* __restrict__ cams,
                   int B, int V, int Q, int J, float img_w, float img_h,
                   float* __restrict__ new_ref, float* __restrict__ refined_abs,
                   float* __restrict__ projs_abs) {
  pdl_enter();
  const int64_t N = static_cast<int64_t>(Q) * J;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t idx = gid / kDltLanes;                   // problem
  const int sl = static_cast<int>(gid % kDltLanes);      // this lane's first view
  if (idx >= static_cast<int64_t>(B) * N) return;        // whole 4-lane groups leave together
  // shuffles stay inside the 4-lane group; all of its lanes are alive here
  const uint32_t gmask = 0xfu << ((threadIdx.x & 31) & ~3);
  const int b = static_cast<int>(idx / N);
  const int64_t n = idx % N;
  const int q = static_cast<int>(n / J);
  float* oref = new_ref + idx * 3;
  if (!selected[static_cast<int64_t>(b) * Q + q]) {       // dq_decoder.py:1013-1029
    if (sl == 0) { oref[0] = 0.f; oref[1] = 0.f; oref[2] = 0.f; }
    for (int v = sl; v < V; v += kDltLanes) {
      const int64_t o = ((static_cast<int64_t>(b) * V + v) * N + n) * 2;
      refined_abs[o] = 0.f; refined_abs[o + 1] = 0.f;
      projs_abs[o] = 0.f; projs_abs[o + 1] = 0.f;
    }
    return;
  }
  // confidence = softmax over views of the logits (nn.Softmax(dim=0), :305,:706-707)
  float mx = -INFINITY;
  for (int v = sl; v < V; v += kDltLanes)
    mx = fmaxf(mx, __ldg(mlp_out + ((static_cast<int64_t>(b) * V + v) * N + n) * mlp_ld + 2));
  mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(gmask, mx, 2));
  float den = 0.f;
  for (int v = sl; v < V; v += kDltLanes)
    den += expf(__ldg(mlp_out + ((static_cast<int64_t>(b) * V + v) * N + n) * mlp_ld + 2) - mx);
  den += __shfl_xor_sync(gmask, den, 1);
  den += __shfl_xor_sync(gmask, den, 2);

  double H[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) H[a][c] = 0.0;

  for (int v = sl; v < V; v += kDltLanes) {
    const int64_t row = (static_cast<int64_t>(b) * V + v) * N + n;
    const MvgCamera* cam = cams + static_cast<int64_t>(b) * V + v;
    const float ox = __ldg(mlp_out + row * mlp_ld), oy = __ldg(mlp_out + row * mlp_ld + 1);
    const float lg = __ldg(mlp_out + row * mlp_ld + 2);
    const float conf = expf(lg - mx) / den;
    const float rx = __ldg(ref2d + row * 2), ry = __ldg(ref2d + row * 2 + 1);
    // :678-701  offset / img_size, refined = proj + offset, back to network pixels
    const float fx_abs = fmul(fadd(rx, fdiv(ox, img_w)), img_w);
    const float fy_abs = fmul(fadd(ry, fdiv(oy, img_h)), img_h);
    refined_abs[row * 2] = fx_abs;
    refined_abs[row * 2 + 1] = fy_abs;
    projs_abs[row * 2] = fmul(rx, img_w);
    projs_abs[row * 2 + 1] = fmul(ry, img_h);
    // :413-420 inverse affine to original-image pixels
    const float* ia = cam->inv_aff;
    const float px = fadd(fadd(fmul(fx_abs, ia[0]), fmul(fy_abs, ia[1])), ia[2]);
    const float py = fadd(fadd(fmul(fx_abs, ia[3]), fmul(fy_abs, ia[4])), ia[5]);
    // :119-204 undistort: K^-1, 5 fixed-point iterations, K
    const float* Ki = cam->Kinv;
    const float x0 = fadd(fadd(fmul(Ki[0], px), fmul(Ki[1], py)), Ki[2]);
    const float y0 = fadd(fadd(fmul(Ki[3], px), fmul(Ki[4], py)), Ki[5]);
    const float k1 = cam->k[0], k2 = cam->k[1], k3 = cam->k[2], p1 = cam->p[0], p2 = cam->p[1];
    float x = x0, y = y0;
#pragma unroll
    for (int itn = 0; itn < 5; ++itn) {
      const float r2 = fadd(fmul(x, x), fmul(y, y));
      const float den_r = fadd(1.f, fmul(fadd(fmul(fadd(fmul(k3, r2), k2), r2), k1), r2));
      const float icdist = fdiv(1.f, den_r);
      const float dX = fadd(fmul(fmul(fmul(2.f, p1), x), y), fmul(p2, fadd(r2, fmul(fmul(2.f, x), x))));
      const float dY = fadd(fmul(p1, fadd(r2, fmul(fmul(2.f, y), y))), fmul(fmul(fmul(2.f, p2), x), y));
      x = fmul(fsub(x0, dX), icdist);
      y = fmul(fsub(y0, dY), icdist);
    }
    const float u = fadd(fmul(cam->f[0], x), cam->c[0]);
    const float w = fadd(fmul(cam->f[1], y), cam->c[1]);
    accumulate_rows(cam->P, u, w, conf, H);
  }
  // normal equations of all views: butterfly over the group's 4 lanes (upper triangle, fp64)
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = a; c < 4; ++c) {
      double h = H[a][c];
      h += __shfl_xor_sync(gmask, h, 1);
      h += __shfl_xor_sync(gmask, h, 2);
      H[a][c] = h;
    }
  float o3[3];
  solve_and_store(H, o3);                                 // identical on the 4 lanes
  if (sl == 0) { oref[0] = o3[0]; oref[1] = o3[1]; oref[2] = o3[2]; }
}

__global__ void __launch_bounds__(128)
triangulate_kernel(const float* __restrict__ proj, const float* __restrict__ points,
                   const float* __restrict__ conf, int n, int V, int J, float* __restrict__ out) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(n) * J) return;
  const int i = static_cast<int>(idx / J), j = static_cast<int>(idx % J);
  double H[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) H[a][c] = 0.0;
  for (int v = 0; v < V; ++v) {
    const float* P = proj + (static_cast<int64_t>(i) * V + v) * 12;
    float Pl[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) Pl[k] = __ldg(P + k);
    const float* pt = points + ((static_cast<int64_t>(i) * V + v) * J + j) * 2;
    const float cf = conf ? __ldg(conf + (static_cast<int64_t>(i) * V + v) * J + j) : 1.f;
    accumulate_rows(Pl, __ldg(pt), __ldg(pt + 1), cf, H);
  }
  solve_and_store(H, out + idx * 3);
}

}  // namespace mvg

extern "C" int mvg_offsets_dlt(const float* mlp_out, int mlp_ld, const float* ref2d,
                               const uint8_t* selected, const float* cams, int batch, int views,
                               int queries, int joints, float img_w, float img_h, float* new_ref,
                               float* refined_abs, float* projs_abs, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(mlp_out && ref2d && selected && cams && new_ref && refined_abs && projs_abs,
              "mvg_offsets_dlt: null pointer");
  MVG_REQUIRE(batch > 0 && views > 0 && queries > 0 && joints > 0, "mvg_offsets_dlt: empty shape");
  MVG_REQUIRE(mlp_ld >= 3, "mvg_offsets_dlt: mlp_ld %d < 3", mlp_ld);
  const int64_t total = static_cast<int64_t>(batch) * queries * joints * kDltLanes;
  const int threads = 128;
  launch_k(offsets_dlt_kernel, dim3(static_cast<unsigned>((total + threads - 1) / threads)), dim3(threads), 0,
           static_cast<cudaStream_t>(stream), mlp_out, mlp_ld, ref2d, selected, reinterpret_cast<const MvgCamera*>(cams),
           batch, views, queries, joints, img_w, img_h, new_ref, refined_abs, projs_abs);
  return check_launch("mvg_offsets_dlt");
}

extern "C" int mvg_triangulate(const float* proj, const float* points, const float* conf, int n,
                               int views, int joints, float* out, void* stream) {
  using namespace mvg;
  MVG_REQUIRE(proj && points && out, "mvg_triangulate: null pointer");
  MVG_REQUIRE(n > 0 && views > 0 && joints > 0, "mvg_triangulate: empty shape");
  const int64_t total = static_cast<int64_t>(n) * joints;
  const int threads = 128;
  triangulate_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0,
                       static_cast<cudaStream_t>(stream)>>>(proj, points, conf, n, views, joints, out);
  return check_launch("mvg_triangulate");
}
