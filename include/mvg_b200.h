/*
 * mvg_b200.h - C ABI of the B200-native (sm_100a) projective-attention decoder hot path.
 *
 * Drop-in boundary for XunshanMan/MVGFormer (reference citations are file:line relative to
 * the reference root).  The reference binds its one native extension through pybind11
 * (`Deformable.deform_forward/backward`, lib/models/ops/src/vision.cpp:24-27, declared in
 * lib/models/ops/src/deform.h:31-72) and does everything else in Python.  This library
 * exposes the same op plus the fused stages of `DQDecoderLayer.forward`
 * (lib/models/dq_decoder.py:850-1045) as plain C functions:
 *
 *   - raw device pointers + explicit sizes, no torch types;
 *   - the caller owns every buffer (PyTorch allocates), nothing is allocated or
 *     synchronised inside, each call only enqueues kernels on `stream`
 *     (a `cudaStream_t` passed as void*), re-entrant per stream - the same contract as the
 *     reference wrapper, which launches on the current stream without sync
 *     (lib/models/ops/src/cuda/deform_cuda.cu:76);
 *   - return 0 on success, a negative MVG_E* code otherwise; `mvg_last_error()` returns a
 *     thread-local message.  Unlike the reference (which only printf's launch errors,
 *     deform_im2col_cuda.cuh:958-962) launch errors are returned.
 *
 * Shapes: B frames, V views, Q queries, J joints, N=Q*J points, C=256 channels, M=8 heads,
 * D=32 channels/head, Lv pyramid levels (<= MVG_MAX_LEVELS), P=8 points, S=sum H_l*W_l.
 * "bf16" pointers are `__nv_bfloat16` (uint16 storage).
 */
#ifndef MVG_B200_H_
#define MVG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MVG_API __attribute__((visibility("default")))
#else
#define MVG_API
#endif

#define MVG_MAX_LEVELS 4
#define MVG_MAX_VIEWS 8
#define MVG_CAM_FLOATS 64 /* floats per packed camera record, see MvgCamera in csrc/common.cuh */
#define MVG_CAM_FIELDS 11 /* raw tensors per view consumed by mvg_pack_cameras */

enum {
  MVG_OK = 0,
  MVG_EINVAL = -1,  /* bad argument (null pointer, unsupported shape / dtype) */
  MVG_ELAUNCH = -2, /* CUDA launch / runtime error (message holds cudaGetErrorString) */
  MVG_EUNSUPPORTED = -3
};

enum { MVG_F32 = 0, MVG_BF16 = 1, MVG_F64 = 2, MVG_F16 = 3 };

/* Message for the last non-zero return on this thread. */
MVG_API const char* mvg_last_error(void);
/* ABI version of this header (bumped on any signature change). */
MVG_API int mvg_abi_version(void);
/* Number of kernels launched by this library in this process (for bench accounting). */
MVG_API int64_t mvg_launch_count(void);

/* ---------------------------------------------------------------------------------------
 * Camera packing: the raw `meta[v]` tensors of lib/dataset/JointsDataset.py:197-220 (batch-first
 * after DataLoader collation) -> cams (B, V, MVG_CAM_FLOATS) fp32 records, ONE launch, no cache.
 * Restates unfold_camera_param_batch (lib/utils/cameras.py:118-133), get_affine_transform(center,
 * scale, 0, img_size) (lib/utils/transforms.py:72-112 - host numpy + cv2 per (view, frame, layer) in
 * the reference, lib/models/dq_decoder.py:361-372), inv_affine_trans[:, :2] (:414-418),
 * get_calib_matrix / K^-1 / P = K [R | -R T] (:207-246) and the clamp bound of :383.
 *   fields: HOST array of views * MVG_CAM_FIELDS DEVICE pointers, per view in this order:
 *     R (B,3,3), T (B,3,1), fx, fy, cx, cy (B), k (B,3,1), p (B,2,1), center (B,2), scale (B,2),
 *     inv_affine_trans (B,3,3) - all contiguous; dtypes: matching HOST array of MVG_F32 / MVG_F64.
 *   The pointer table is read at call time (captured by value in a CUDA graph).
 */
MVG_API int mvg_pack_cameras(const void* const* fields, const int* dtypes, int batch, int views,
                     float img_w, float img_h, float* cams, void* stream);

/* ---------------------------------------------------------------------------------------
 * Deformable.deform_forward  (lib/models/ops/src/deform.h:31-50,
 * lib/models/ops/src/cuda/deform_cuda.cu:31-91, kernel deform_im2col_cuda.cuh:247-309)
 *   value (B,S,M,D) dtype `dtype`; spatial_shapes (Lv,2) int64 DEVICE; level_start_index
 *   (Lv) int64 DEVICE; sampling_loc (B,Lq,M,Lv,P,2) and attn_weight (B,Lq,M,Lv,P) dtype
 *   `dtype`; out (B,Lq,M*D) dtype `dtype`, fully overwritten (the reference zero-fills it,
 *   deform_cuda.cu:65).  dtype MVG_F32 / MVG_BF16 (fp32 accumulation) or MVG_F64 (the reference
 *   dispatches all floating types, deform_cuda.cu:75).  D must be 32, M*D <= 1024.
 *   `im2col_step` is accepted for signature fidelity; batch % min(batch, step) == 0 is
 *   enforced like deform_cuda.cu:63, the chunk loop itself is not needed.
 */
MVG_API int mvg_deform_forward(const void* value, const int64_t* spatial_shapes,
                       const int64_t* level_start_index, const void* sampling_loc,
                       const void* attn_weight, int dtype, int batch, int spatial_size,
                       int num_heads, int channels, int num_levels, int num_query,
                       int num_point, int im2col_step, void* out, void* stream);

/* Deformable.deform_backward (deform.h:52-72, deform_cuda.cu:94-164, col2im kernels
 * deform_im2col_cuda.cuh:311-930).  fp32 only.  grad_value (B,S,M,D) must be ZEROED by the
 * caller (the reference allocates it with at::zeros, deform_cuda.cu:129); grad_sampling_loc
 * and grad_attn_weight are fully overwritten. */
MVG_API int mvg_deform_backward(const float* value, const int64_t* spatial_shapes,
                        const int64_t* level_start_index, const float* sampling_loc,
                        const float* attn_weight, const float* grad_output, int batch,
                        int spatial_size, int num_heads, int channels, int num_levels,
                        int num_query, int num_point, int im2col_step, float* grad_value,
                        float* grad_sampling_loc, float* grad_attn_weight, void* stream);

/* ---------------------------------------------------------------------------------------
 * Pyramid hand-off: NCHW maps -> one channels-last bf16 matrix (rows, S, C).
 * Replaces `input_flatten = cat([src.flatten(2)...]).permute(0,2,1)`
 * (lib/models/ops/modules/projattn.py:160).  src_l: (rows, C, H_l*W_l) fp32 or bf16;
 * dst: (rows, S, C) bf16; level l occupies positions [start_l, start_l + H_l*W_l).
 */
MVG_API int mvg_pyramid_to_channels_last(const void* const* src_levels, int src_dtype, int num_levels,
                                 const int* level_hw /* Lv: H_l*W_l */, int rows, int channels,
                                 void* dst_bf16, void* stream);

/* ---------------------------------------------------------------------------------------
 * Dense projection on tcgen05 tensor cores:  out = act(A @ W^T + bias)
 *   A (M,K) bf16 row-major, lda = K;  W (Nout,K) bf16 row-major (nn.Linear layout);
 *   bias (Nout) fp32 or NULL; out (M,Nout) bf16 or fp32 (`out_dtype`), row stride ldo
 *   elements.  relu != 0 applies max(0, .).  K % 64 == 0, Nout % 16 == 0.
 *   row_mask (M) uint8 or NULL: rows with mask 0 are written as zeros (the `bounding` filter
 *   of dq_decoder.py:585-586 fused into output_proj).
 * Used for: value/rayconv + the per-level sampling_offsets/attention_weights projections of
 * the raw pyramid (projattn.py:169,180-181), output_proj (:203), feature_update_mlp, FFN,
 * offset_net MLP (dq_decoder.py:101,284,289-292).
 */
MVG_API int mvg_linear_bf16(const void* A, const void* W, const float* bias, void* out, int out_dtype,
                    int64_t M, int Nout, int K, int64_t ldo, int relu, const uint8_t* row_mask,
                    void* stream);

/* Value / offset / logit projection of the channels-last pyramid for ALL decoder layers in one
 * tcgen05 GEMM (rayconv + the per-level sampling_offsets / attention_weights Linears,
 * projattn.py:169,180-181), written in the layouts the fused gather reads:
 *   feat (M = V*B*S, 256) bf16; W (layers*448, 256) bf16, per layer rows [rayconv 256 |
 *   sampling_offsets 128 | attention_weights 64]; bias (layers*448) fp32 or NULL;
 *   value_hm (layers*8, M, 32) FP16 head-major; gmap (M, layers*192) FP16 (bf16 operands, fp32
 *   accumulation, fp16 stores: the fused gather blends in packed fp16).
 */
MVG_API int mvg_value_proj_gemm(const void* feat, const void* W, const float* bias, int64_t M, int layers,
                        void* value_hm, void* gmap, void* stream);

/* The same projection reading the NCHW pyramid levels IN PLACE (what the reference's backbone hands to
 * ProjAttn before its flatten + permute, projattn.py:160): src_levels = HOST array of num_levels device
 * pointers to (rows = V*B, 256, H_l, W_l) BF16 maps, level_hw[l] = H_l * W_l.  An M tile of 128 texels is
 * loaded by TMA as an MN-major tcgen05 operand, so mvg_pyramid_to_channels_last and its (rows, S, 256)
 * buffer are not needed.  Requires H_l * W_l % 128 == 0 for every level
 * (mvg_value_proj_gemm_nchw_supported returns 1); results are identical to the two-step path. */
MVG_API int mvg_value_proj_gemm_nchw_supported(int num_levels, const int* level_hw);
MVG_API int mvg_value_proj_gemm_nchw(const void* const* src_levels, int num_levels, const int* level_hw, int rows,
                             const void* W, const float* bias, int layers, void* value_hm, void* gmap,
                             void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused projection + projective attention sampling for all (b, v, n):
 *   a3  project_ref_points   (dq_decoder.py:331-397, cameras.py:167-207, transforms.py:135-141)
 *   a4  ProjAttn minus the dense GEMMs (projattn.py:139-153 ref-point feature lookup,
 *       :180-191 offsets / softmax over Lv*P / sampling locations incl. the layout scramble)
 *   a5  multi-scale deformable gather (deform_im2col_cuda.cuh:247-309)
 * Inputs:
 *   ref3d (B,N,3) fp32 world mm;  cams (B,V,MVG_CAM_FLOATS) fp32 packed cameras;
 *   value_hm (8 heads, V*B*S rows, 32) FP16, HEAD-MAJOR: head h of position s of map row r = v*B + b
 *       is the 64 bytes at value_hm + h*value_head_stride + (r*S + s)*32 (rayconv output; the two
 *       horizontal corners of a bilinear footprint are contiguous);
 *   gmap (V*B*S rows, ld_g) FP16: columns [0,128) = sampling_offsets.weight @ feat, [128,192) =
 *       attention_weights.weight @ feat (both WITHOUT bias; bilinear interpolation commutes with
 *       the linear map).  Both are written by mvg_value_proj_gemm;
 *   qproj (B,N,192) fp32 = [sampling_offsets; attention_weights](tgt + query_pos) + bias.
 * Outputs:
 *   sampled (B,V,N,256) bf16 (input of output_proj), ref2d (B,V,N,2) fp32 normalised
 *   network-image coordinates, bounding (B,V,N) uint8.
 * ProjAttn.forward entry (projattn.py:115): when `refl_in` (B,V,N,Lv,2) is non-NULL the
 *   projection is skipped and the per-level normalised reference points are read from it
 *   (ref3d, cams, ref2d, bounding may then be NULL).
 * Out-of-view points (bounding == 0): the reference multiplies their attention feature by 0
 *   (dq_decoder.py:585-586) and reads it nowhere else, so they are not gathered; their
 *   `sampled` rows are zeros.
 * Arithmetic: geometry, softmax and the bilinear x attention weights in fp32; the weights are then
 *   rounded to fp16 and the 96-term blend of an (item, head) runs in packed fp16 (HFMA2; the bf16
 *   rounding of `sampled` is coarser than its accumulation error).  Items are binned by image cell and gathered from shared-memory tiles staged by
 *   cp.async.bulk (csrc/project_sample.cu).
 * `workspace`: device, 256-byte aligned, mvg_project_sample_workspace_bytes(prm) bytes, contents
 *   irrelevant on entry; its first B*V int32 receive the in-view item count of every (frame, view).
 */
typedef struct {
  int batch, views, points;     /* B, V, N */
  int num_levels;               /* Lv (3 for the shipped configs) */
  int level_h[MVG_MAX_LEVELS];
  int level_w[MVG_MAX_LEVELS];
  int level_start[MVG_MAX_LEVELS];
  int spatial_size;             /* S */
  int ld_g;                     /* row stride of the offset/logit map G in elements (>= 192) */
  float img_w, img_h;           /* network image size (NETWORK.IMAGE_SIZE) */
  int64_t value_head_stride;    /* elements between consecutive heads of value_hm (>= V*B*S*32) */
} MvgSampleParams;

MVG_API int64_t mvg_project_sample_workspace_bytes(const MvgSampleParams* prm);
MVG_API int mvg_project_sample_fused(const float* ref3d, const float* cams, const void* value_hm,
                             const void* gmap, const float* qproj, const MvgSampleParams* prm, void* sampled,
                             float* ref2d, uint8_t* bounding, const float* refl_in,
                             void* workspace, void* stream);

/* The two halves of mvg_project_sample_fused as separate calls on the same workspace, so that a host can run
 * the first on a side stream while the per-point projection GEMM (qproj) is still being computed:
 *   mvg_project_bin     projection (or `refl_in`), ref2d / bounding, zero rows for out-of-view points, binning
 *   mvg_sample_gather   per-sample parameters + the tiled gather (needs qproj, value_hm, gmap)
 * mvg_project_sample_fused == mvg_project_bin followed by mvg_sample_gather on one stream. */
MVG_API int mvg_project_bin(const float* ref3d, const float* cams, const MvgSampleParams* prm, void* sampled,
                    float* ref2d, uint8_t* bounding, const float* refl_in, void* workspace, void* stream);
MVG_API int mvg_sample_gather(const void* value_hm, const void* gmap, const float* qproj, const MvgSampleParams* prm,
                      void* sampled, const float* ref2d, const float* refl_in, void* workspace, void* stream);

/* The projection (a3) alone - DQDecoderLayer.project_ref_points (dq_decoder.py:331-397) for all V views in
 * one launch, same arithmetic as inside mvg_project_sample_fused (`bounding` bit-exact): ref3d (B,N,3)
 * world mm, cams (B,V,MVG_CAM_FLOATS) -> ref2d (B,V,N,2) fp32 normalised network-image coordinates,
 * bounding (B,V,N) uint8.  Used by the training-mode layer (mvgformer_b200/training.py), where the
 * sampling runs through mvg_deform_forward / mvg_deform_backward. */
MVG_API int mvg_project_points(const float* ref3d, const float* cams, int batch, int views, int points,
                       float img_w, float img_h, float* ref2d, uint8_t* bounding, void* stream);

/* ---------------------------------------------------------------------------------------
 * Integer path of the query filter (dq_decoder.py:596-656): threshold mask, torch.where
 * order, per-frame counts, padding to the max count with query id 0, stable sort by frame,
 * and the reverse (un-pad) ids.  prob (B,Q,2) fp32.  method 0 = 'threshold'
 * (prob[...,1] > thr), 1 = 'all' (prob[...,0] > 0).
 * min_one != 0 applies the "always one query" rule (:620-623) locally; query-sharded ranks
 * pass 0 and apply it after a global count (mvgformer_b200/sharding.py).
 * Outputs (device): selected (B,Q) uint8 incl. the "always one query" rule (:620-623);
 * counts (B) int32; info (4) int32 = {n_valid, max_count, 0, 0}; batch_ids / query_ids
 * (capacity B*Q, first B*max_count valid) int64; batch_ids_rev / query_ids_rev
 * (capacity B*Q, first n_valid valid) int64.  The id arrays may be NULL.
 */
MVG_API int mvg_select_pad(const float* prob, int batch, int queries, float threshold, int method,
                   int min_one, uint8_t* selected, int32_t* counts, int32_t* info, int64_t* batch_ids,
                   int64_t* query_ids, int64_t* batch_ids_rev, int64_t* query_ids_rev,
                   void* stream);

/* ---------------------------------------------------------------------------------------
 * Offsets -> refined 2D -> DLT triangulation, with the scatter/zero-fill semantics of
 * dq_decoder.py:1011-1029:
 *   a9  calculate_2d_offsets tail (:678-707): offset/img_size, refined = proj + offset,
 *       x img_size, confidence softmax over views
 *   a10 inverse affine, 5-iteration undistort (:119-204), P = K[R|-RT] (:223-246)
 *   a11 DLT (lib/mvn/utils/multiview.py:170-228): smallest right singular vector of the
 *       confidence-weighted 2Vx4 system, solved as a 4x4 symmetric eigenproblem of A^T A by
 *       cyclic Jacobi in fp64 (more accurate than the reference's fp32 LAPACK SVD).
 * mlp_out (B*V*N rows, row stride mlp_ld >= 3 floats) fp32 = offset_net output
 * (dx, dy, conf-logit) in columns 0..2; ref2d (B,V,N,2);
 * selected (B,Q) uint8.  Outputs: new_ref (B,N,3), refined_abs (B,V,N,2), projs_abs
 * (B,V,N,2) fp32 - zeros where the query is not selected.
 */
MVG_API int mvg_offsets_dlt(const float* mlp_out, int mlp_ld, const float* ref2d,
                    const uint8_t* selected, const float* cams, int batch, int views,
                    int queries, int joints, float img_w, float img_h, float* new_ref,
                    float* refined_abs, float* projs_abs, void* stream);

/* multiview.triangulate_batch_of_points_batch_version (lib/mvn/utils/multiview.py:257-269):
 * proj (n,V,3,4) fp32, points (n,V,J,2) fp32, conf (n,V,J) fp32 or NULL -> out (n,J,3). */
MVG_API int mvg_triangulate(const float* proj, const float* points, const float* conf, int n,
                    int views, int joints, float* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused elementwise stages of update_feature (dq_decoder.py:763-778, :845-848):
 *   mvg_masked_view_mean: aver[b,n,:] = mean_v( bounding[b,v,n] * x[b,v,n,:] )   (bf16 io)
 *   mvg_add_layernorm   : out = LayerNorm(a + b) * gamma + beta  (rows x 256; a fp32, b bf16
 *                         or fp32; writes fp32 and optionally a bf16 copy for the next GEMM)
 *   mvg_class_prob      : prob[b,q,:] = mean_j sigmoid(cls[b,q*J+j,:])  (dq_decoder.py:889-893)
 */
MVG_API int mvg_masked_view_mean(const void* x_bf16, const uint8_t* bounding, int batch, int views,
                         int points, int channels, void* out_bf16, void* stream);
MVG_API int mvg_add_layernorm(const float* a, const void* b, int b_dtype, const float* gamma,
                      const float* beta, int64_t rows, int channels, float eps, float* out_f32,
                      void* out_bf16, void* stream);
MVG_API int mvg_class_prob(const float* cls, int batch, int queries, int joints, float* prob,
                   void* stream);
/* out_bf16[i] = bf16(a[i] + b[i]) (b may be NULL): `with_pos_embed` + operand cast,
 * dq_decoder.py:580 / mvp_decoder.py:90-92.  n % 8 == 0. */
MVG_API int mvg_add_cast_bf16(const float* a, const float* b, void* out_bf16, int64_t n, void* stream);
/* class_embed Linear(256,2) + sigmoid + mean over joints in one pass (fp32):
 * x (B, Q*J, 256) fp32, w (2,256), bias (2) -> prob (B,Q,2). */
MVG_API int mvg_class_head(const float* x, const float* w, const float* bias, int batch, int queries,
                   int joints, float* prob, void* stream);

/* Fused query-feature update of one decoder layer (dq_decoder.py:770-778 'MLP' branch +
 * forward_ffn, mvp_decoder.py:94-98), one kernel, activations resident on chip:
 *   tu  = LayerNorm(tgt + aver @ w_fu^T + b_fu; g2, e2, eps2)
 *   out = LayerNorm(tu + relu(tu @ w1^T + b1) @ w2^T + b2; g3, e3, eps3)
 * aver (M,256) bf16 (mvg_masked_view_mean output), tgt (M,256) fp32, w_fu (256,256), w1
 * (d_ffn,256), w2 (256,d_ffn) bf16 row-major (nn.Linear layout), biases / LayerNorm parameters
 * fp32, out (M,256) fp32.  d_model is 256, d_ffn a multiple of 256.  Dropout is identity
 * (inference).  GEMMs on tcgen05 with fp32 accumulation; t2 and the FFN output are NOT rounded
 * to bf16 (the unfused path does round them).
 */
MVG_API int mvg_ffn_chain(const void* aver_bf16, const float* tgt, const void* w_fu, const float* b_fu,
                  const float* g2, const float* e2, float eps2, const void* w1, const float* b1,
                  const void* w2, const float* b2, const float* g3, const float* e3, float eps3,
                  int64_t M, int d_ffn, float* out, void* stream);

/* Fused offset_net MLP (dq_decoder.py:97-111, :659-717; MLP multi_view_pose_transformer.py:81-102) on the
 * rows of the SELECTED queries only (the reference gathers them into a padded rectangle, :899-932):
 *   out[row, 0..2] = relu(relu(attn[row] @ W1^T + b1) @ W2^T + b2) @ W3^T + b3
 * attn (B*V*N, 256) bf16 (output_proj result); info, query_ids_pad (= query_ids), batch_ids_rev,
 * query_ids_rev: the DEVICE outputs of mvg_select_pad (selected query s sits in frame
 * batch_ids_rev[s] at position query_ids_rev[s] of the padded rectangle, :941-947);
 * W1, W2 (256,256) bf16, b1, b2 (256) fp32, W3 (3,256) fp32, b3 (3) fp32; out (B*V*N, out_ld) fp32:
 * columns 0..2 of the rows of selected queries (all views, all joints) are written, nothing else.
 * One tcgen05 kernel, hidden activations stay on chip, no host synchronisation. */
MVG_API int mvg_offset_chain(const void* attn_bf16, const int32_t* info, const int64_t* query_ids_pad,
                     const int64_t* batch_ids_rev, const int64_t* query_ids_rev, const void* w1, const float* b1, const void* w2,
                     const float* b2, const float* w3, const float* b3, int batch, int views, int queries,
                     int joints, float* out, int out_ld, void* stream);

/* ---------------------------------------------------------------------------------------
 * The steps either side of the decoder.
 *
 * mvg_init_queries: query / reference-point construction of DyanmicQueryTransformer.forward
 *   (lib/models/dq_transformer.py:394-432 query_embed_type='person_joint'; :298-323
 *   init_ref_method='sample_space'; norm2absolute multi_view_pose_transformer.py:575-580).
 *   joint_emb (J,2C), inst_emb (Q,2C) fp32; lin (grid_n) fp32 = torch.linspace(0,1,grid_n),
 *   grid_n = ceil(sqrt(Q)); tpose (J,3) float64 mm; space_size / space_center: 3 HOST floats.
 *   Outputs: query_pos, tgt (B,Q*J,C) fp32 (first / second half of the summed embeddings),
 *   ref (B,Q*J,3) fp32 mm.
 *
 * mvg_assemble_predictions: poses (B,Q*J,3), prob (B,Q,2) = last layer's class output ->
 *   pred (B,Q,J,5) = [x, y, z, (score > thr) - 1, score], score = sigmoid(inverse_sigmoid(prob[...,1]))
 *   (dq_transformer.py:568, util/misc.py:608-612, lib/core/function.py:386-392), plus the score
 *   filter of run/validate_3d.py:229: valid_ids (B,Q) int32 = query ids with score > thr in
 *   ascending order (first valid_count[b] entries valid), valid_count (B) int32.
 *
 * mvg_nearby_joints_nms: lib/core/nms.py:210-284 (combined_input=True, max_dets=-1) on the valid
 *   poses of every frame.  workspace: B * Q * ceil(Q/32) uint32 (close-instance bit matrix).
 *   keep_compact (B,Q) int32 = kept indices INTO THE FILTERED ARRAY in the order the reference
 *   appends them (its return value), keep_query (B,Q) = the same as query ids, keep_count (B).
 *   Q <= 4096, J <= 32.  Equal scores: the reference's argsort is not stable (order
 *   unspecified); here the larger index goes first.
 */
MVG_API int mvg_init_queries(const float* joint_emb, const float* inst_emb, const float* lin,
                     const double* tpose, const float* space_size, const float* space_center,
                     int batch, int queries, int joints, int channels, int grid_n,
                     float* query_pos, float* tgt, float* ref, void* stream);
MVG_API int mvg_assemble_predictions(const float* poses, const float* prob, int batch, int queries,
                             int joints, float threshold, float* pred, int32_t* valid_ids,
                             int32_t* valid_count, void* stream);
MVG_API int mvg_nearby_joints_nms(const float* pred, const int32_t* valid_ids, const int32_t* valid_count,
                          int batch, int queries, int joints, float dist_thr,
                          int num_nearby_joints_thr, uint32_t* workspace, int32_t* keep_compact,
                          int32_t* keep_query, int32_t* keep_count, void* stream);

/* ---------------------------------------------------------------------------------------
 * Whole-path drivers (no torch, no Python): the launch sequence of one DQDecoderLayer.forward
 * (eval, indices=None; lib/models/dq_decoder.py:850-1045) and of DQDecoder.forward with
 * return_intermediate=True (:1107-1172) on a caller-provided workspace.  Only the shipped
 * configuration is built (feature_update_method='MLP', open_forward_ffn, 3-layer offset_net,
 * d_model 256, 8 heads x 8 points, module n_levels 1).
 */
typedef struct {
  int batch, views, queries, joints, layers;      /* B, V, Q, J, L */
  int num_levels;                                 /* Lv */
  int level_h[MVG_MAX_LEVELS], level_w[MVG_MAX_LEVELS];
  float img_w, img_h;                             /* NETWORK.IMAGE_SIZE */
  float threshold;                                /* query filter, dq_decoder.py:596-612 */
  int filter_query;                               /* 1: prob[...,1] > threshold; 0: all queries */
  int local_min_one;                              /* 1: apply the "always one query" rule (:620-623) here;
                                                     0: query-sharded rank (rule applied after the global count) */
  int d_ffn;                                      /* multiple of 256 */
} MvgDecoderConfig;

/* Device pointers to one layer's operands as mvgformer_b200 packs them: GEMM weights bf16 row-major
 * (nn.Linear layout), biases / LayerNorm parameters / class head fp32. */
typedef struct {
  const void* w_q;  const float* b_q;             /* [sampling_offsets; attention_weights] (192,256), (192) */
  const void* w_o;  const float* b_o;             /* proj_attn.output_proj (256,256) */
  const void* w_fu; const float* b_fu;            /* feature_update_mlp */
  const float* g2;  const float* e2;  float eps2; /* norm2 */
  const void* w1;   const float* b1;              /* linear1 (d_ffn,256) */
  const void* w2;   const float* b2;              /* linear2 (256,d_ffn) */
  const float* g3;  const float* e3;  float eps3; /* norm3 */
  const float* wc;  const float* bc;              /* class_embed (2,256), (2) fp32 */
  const void* w_m1; const float* b_m1;            /* pose_embed.MLP.layers.0 */
  const void* w_m2; const float* b_m2;            /* pose_embed.MLP.layers.1 */
  const float* w_m3; const float* b_m3;           /* pose_embed.MLP.layers.2 (3,256), (3) fp32 */
} MvgLayerWeights;

/* Bytes of `workspace` for mvg_decoder / mvg_decoder_layer (pyramid_is_nchw: the pyramid arrives as
 * NCHW levels and is transposed into the workspace first).  -1 on a bad configuration. */
MVG_API int64_t mvg_decoder_workspace_bytes(const MvgDecoderConfig* cfg, int pyramid_is_nchw);

/* One layer on pre-projected maps: value_hm / gmap point at THIS layer's 8 heads / 192 columns of the
 * mvg_value_proj_gemm outputs (ld_g, value_head_stride as in MvgSampleParams).  cams from
 * mvg_pack_cameras.  tgt, query_pos (B,N,256), ref3d (B,N,3) fp32 in; outputs = the layer's 5-tuple:
 * tgt_out (B,N,256), ref_out (B,N,3), refined_abs / projs_abs (B,V,N,2), class_prob (B,Q,2);
 * selected_count (1 int32, may be NULL) = queries selected here.  workspace: 256-byte aligned,
 * >= mvg_decoder_workspace_bytes(cfg, 0). */
MVG_API int mvg_decoder_layer(const MvgDecoderConfig* cfg, const MvgLayerWeights* w, const void* value_hm,
                      const void* gmap, int ld_g, int64_t value_head_stride, const float* cams,
                      const float* tgt, const float* query_pos, const float* ref3d, float* tgt_out,
                      float* ref_out, float* refined_abs, float* projs_abs, float* class_prob,
                      int32_t* selected_count, void* workspace, int64_t workspace_bytes, void* stream);

/* The L-layer decoder.  layers: HOST array of L MvgLayerWeights; w_vg_all (L*448,256) bf16 = per layer
 * [rayconv | sampling_offsets | attention_weights], b_vg_all (L*448) fp32 (zeros for the last 192 of
 * every layer: their bias rides in b_q).  Pyramid: EITHER pyramid_levels (HOST array of Lv device
 * pointers to (V*B, 256, H_l, W_l) maps of dtype pyramid_dtype) OR pyramid_cl ((V*B, S, 256) bf16,
 * consumed in place); the other one NULL.  Outputs (return_intermediate): hs (L,B,N,256),
 * refs (L,B,N,3), refs2d / projs2d (L,B,V,N,2), class_probs (L,B,Q,2), selected_counts (L) int32
 * or NULL. */
MVG_API int mvg_decoder(const MvgDecoderConfig* cfg, const MvgLayerWeights* layers, const void* w_vg_all,
                const float* b_vg_all, const void* const* pyramid_levels, int pyramid_dtype,
                const void* pyramid_cl, const float* cams, const float* tgt, const float* query_pos,
                const float* ref3d, float* hs, float* refs, float* refs2d, float* projs2d,
                float* class_probs, int32_t* selected_counts, void* workspace, int64_t workspace_bytes,
                void* stream);

/* Exchange step of the query-sharded mode: ONE all-gather of every rank's final
 * poses (B, Ql*J, 3), class prob (B, Ql, 2) and per-layer selected counts (L) -> poses (B, Q*J, 3),
 * prob (B, Q, 2), counts (L) fp32 = global per-layer counts, identical on every rank.  Ranks own
 * contiguous query blocks whose sizes differ by at most one (rank r: Q/world + (r < Q % world)).
 * nccl_comm: the caller's ncclComm_t (ignored for world == 1); ncclAllGather is resolved at run time from
 * the NCCL library already loaded in the process.  workspace: mvg_allgather_poses_workspace_bytes. */
MVG_API int64_t mvg_allgather_poses_workspace_bytes(int batch, int queries, int joints, int layers, int world);
MVG_API int mvg_allgather_poses(void* nccl_comm, int rank, int world, const float* poses_local,
                        const float* prob_local, const int32_t* counts_local, int batch, int queries,
                        int joints, int layers, float* poses, float* prob, float* counts,
                        void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVG_B200_H_ */
