/*
 * unibev_b200 -- C ABI of libunibev_b200.so (sm_100a).
 *
 * Drop-in boundary for the UniBEV uniform-BEV-encoder hot path.  Plain pointers
 * and sizes only: no torch types cross this boundary.  All pointers are DEVICE
 * pointers unless the parameter comment says "host".  Every call is asynchronous
 * on `stream` (a cudaStream_t) and returns 0 on success or a negative UB_E* code
 * (message: ub_last_error(), thread-local).  The caller owns every tensor and
 * workspace, and makes the tensors' device current before calling.  No entry point
 * allocates device memory or keeps per-call device state.  Host-side state is
 * limited to: the thread-local error text, two diagnostic counters
 * (ub_launch_count / ub_unsupported_count), a mutex-guarded cache of encoded TMA
 * tensor maps, and per-(kernel, device) records of the shared-memory opt-in.
 * Calls are re-entrant and may be captured into CUDA graphs.  Tensors are
 * contiguous row-major fp32 unless stated.
 *
 * Reference interfaces replaced (paths under /root/reference):
 *   [R1] mmcv MultiScaleDeformableAttnFunction.apply / ext_module.ms_deform_attn_forward|backward,
 *        called at projects/UniBEV/unibev_plugin/models/modules/spatial_cross_attention_img.py:432-435,
 *        spatial_cross_attention_pts.py:439-442, decoder.py:324-327 (ext handles loaded at
 *        spatial_cross_attention_img.py:19-20).
 *   [R2] ImgEncoder.get_reference_points + point_sampling, encoder_unibev_detr_img.py:45-187.
 *   [R3] SpatialCrossAttentionImg.forward rebatch/sample/scatter/count, spatial_cross_attention_img.py:141-212,
 *        with MSDeformableAttention3DImg.forward softmax/offset/anchor logic, :385-419.
 *   [R4] PtsEncoder.point_sampling + SpatialCrossAttentionPts/MSDeformableAttention3DPts.forward,
 *        encoder_unibev_detr_pts.py:105-127, spatial_cross_attention_pts.py:159-204,383-437; and the BEV
 *        self-attention (mmcv MultiScaleDeformableAttention, verbatim copy at decoder.py:278-330).
 *   [R5] residual add + nn.LayerNorm steps of BaseTransformerLayer ('norm' ops), encoder_unibev_detr_img.py:434-436.
 *   [R6] UniBEVTransformer.channel_feature_norm / spatial_feature_norm / multi_modal_fusion,
 *        transformer_fusion.py:316-337, 386-413, 280-314.
 *   [R7] UniBEVTransformer._pre_process_img_feats / _pre_process_pts_feats, transformer_fusion.py:231-278.
 *   [R8] UniBEV.voxelize -> self.pts_voxel_layer(res) (mmdet3d Voxelization / hard_voxelize_forward, max_num_points=10,
 *        voxel_size=[0.075, 0.075, 0.2], max_voxels=(90000, 120000)) and the HardSimpleVFE mean,
 *        projects/UniBEV/unibev_plugin/models/detectors/unibev_detector.py:151-175, :112-124;
 *        configs/unibev/unibev_nus_LC_cnw_256_modality_dropout.py:186-193.
 */
#ifndef UNIBEV_B200_H_
#define UNIBEV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ub_stream_t; /* cudaStream_t */

enum {
  UB_OK = 0,
  UB_EINVAL = -1,   /* bad argument / unsupported shape */
  UB_ECUDA = -2,    /* CUDA runtime error at launch */
  UB_EALIGN = -3,   /* pointer not 16-byte aligned */
  UB_EUNSUPPORTED = -4 /* valid arguments, but this entry point has no kernel for the shape: use the generic one */
};

/* Fusion modes of ub_cnw_fuse [R6]. */
enum { UB_FUSE_LINEAR = 0, UB_FUSE_AVG = 1, UB_FUSE_CAT = 2 };

/* ABI version (major*1000 + minor) and last error text of the calling thread. */
int ub_version(void);
const char* ub_last_error(void);
/* Number of kernels this library has launched since load / since the last reset (bench bookkeeping). */
int64_t ub_launch_count(void);
void ub_launch_count_reset(void);
/* Calls that returned UB_EUNSUPPORTED since the last ub_launch_count_reset (each made the caller take a generic entry point). */
int64_t ub_unsupported_count(void);

/* ---- [R1] generic multi-scale deformable attention -------------------------------------------------
 * value (B, Nv, H, D); spatial_shapes (L, 2) int64 (h, w); level_start_index (L) int64;
 * sampling_loc (B, Nq, H, L, P, 2) normalised (x, y); attn_weight (B, Nq, H, L, P); out (B, Nq, H*D).
 * Pixel convention x_pix = x*W - 0.5, zero padding, each corner bounds-checked (mmcv kernel semantics).
 * No im2col_step restriction (mmcv asserts B % min(B, 64) == 0; this does not). */
int ub_msda_fwd(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                const float* sampling_loc, const float* attn_weight, float* out,
                int B, int Nv, int H, int D, int Nq, int L, int P, ub_stream_t stream);
/* grad_value must be zero-filled by the caller (accumulated with atomics); grad_loc / grad_w are overwritten. */
int ub_msda_bwd(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                const float* sampling_loc, const float* attn_weight, const float* grad_out,
                float* grad_value, float* grad_loc, float* grad_w,
                int B, int Nv, int H, int D, int Nq, int L, int P, ub_stream_t stream);

/* ---- [R2] pillar reference points -> camera planes ------------------------------------------------
 * For every (b, q = h*bev_w + w, cam, anchor d): lift the BEV cell centre to z-anchor d, scale to
 * metres with pc_range, project with lidar2img[b, cam] (row-major 4x4), clamp depth at 1e-5, divide by
 * (img_w, img_h); mask = depth > 1e-5 && 0 < x < 1 && 0 < y < 1.
 * lidar2img (B, N, 16); zs_host: D normalised anchor heights (host, D <= 8); pc_range_host: 6 floats (host).
 * ref_cam out (B, Nq, N, D, 2); mask out (B, Nq, N) uint8, bit d = anchor d visible. */
int ub_project_points(const float* lidar2img, const float* zs_host, const float* pc_range_host,
                      float img_h, float img_w, float* ref_cam, uint8_t* mask,
                      int B, int N, int bev_h, int bev_w, int D, ub_stream_t stream);

/* ---- [R4] fused BEV-grid deformable sampling (BEV self-attention, LiDAR cross-attention) ----------
 * One level.  value (B, fH*fW, H*Dh) already value-projected.  qproj holds, per query row of stride
 * `ld` floats, the raw sampling offsets (H, P, 2) at column off_col and the raw attention logits (H, P)
 * at column logit_col (i.e. the un-normalised outputs of the sampling_offsets / attention_weights
 * linears).  The kernel generates the reference point ((w+.5)/bev_w, (h+.5)/bev_h) itself, adds
 * offset/(fW, fH), applies softmax over the P logits, gathers bilinearly and reduces: out (B, Nq, H*Dh). */
int ub_bev_sample_fwd(const float* value, const float* qproj, float* out,
                      int B, int bev_h, int bev_w, int fH, int fW, int H, int Dh, int P,
                      int ld, int off_col, int logit_col, ub_stream_t stream);

/* ---- [R3] fused camera cross-attention sampling ---------------------------------------------------
 * value (B, N, fH*fW, H*Dh); qproj as above (one row per BEV query, shared by all cameras);
 * ref_cam / mask from ub_project_points with D anchors; sampling point p uses anchor p % D.
 * A camera contributes to query q of batch item b iff batch item 0's mask for (q, cam) is non-zero
 * (reference quirk, spatial_cross_attention_img.py:142); the sum over cameras is divided by
 * max(1, #cameras whose mask for (b, q) is non-zero) (:209-212).  out (B, Nq, H*Dh). */
int ub_img_sample_fwd(const float* value, const float* qproj, const float* ref_cam, const uint8_t* mask,
                      float* out, int B, int N, int bev_h, int bev_w, int fH, int fW, int H, int Dh, int P,
                      int D, int ld, int off_col, int logit_col, ub_stream_t stream);

/* ---- [R4]/[R3] backward twins of the two kernels above (training step, BASELINE configs[4]) ----------
 * Replace ms_deform_attn_backward + the autograd of the softmax / offset normalisation / reference-point add
 * around it (spatial_cross_attention_img.py:390-419, spatial_cross_attention_pts.py:396-426) and, in camera mode,
 * of the per-camera rebatch / scatter / count division (spatial_cross_attention_img.py:141-212).
 * grad_out (B, Nq, H*Dh).  grad_value: same shape as value, ZERO-FILLED by the caller (red.global.add).
 * grad_qproj (B, Nq, ld): gradient with respect to the RAW offset / logit columns (softmax backward folded in);
 * only the columns [off_col, off_col + 2 H P) and [logit_col, logit_col + H P) of each row are written.
 * Dh = 4 x a power of two <= 128, P <= 16; other shapes return UB_EUNSUPPORTED. */
int ub_bev_sample_bwd(const float* value, const float* qproj, const float* grad_out, float* grad_value,
                      float* grad_qproj, int B, int bev_h, int bev_w, int fH, int fW, int H, int Dh, int P, int ld,
                      int off_col, int logit_col, ub_stream_t stream);
int ub_img_sample_bwd(const float* value, const float* qproj, const float* ref_cam, const uint8_t* mask,
                      const float* grad_out, float* grad_value, float* grad_qproj, int B, int N, int bev_h, int bev_w,
                      int fH, int fW, int H, int Dh, int P, int D, int ld, int off_col, int logit_col,
                      ub_stream_t stream);

/* ---- [R4]/[R3] window-staged fast path (fp16-staged values, fp32 accumulation) ------------------------
 * Same arithmetic as ub_bev_sample_fwd / ub_img_sample_fwd, but the value map is read from fp16 head-major
 * planes staged in shared memory by TMA, and the bilinear x attention weights are rounded to fp16 before the
 * fp32-accumulated multiply-add (error well below that of a TF32 projection GEMM).  Head dim 32
 * and 4 or 8 points only: other shapes return UB_EUNSUPPORTED and the caller uses the fp32 entry points.
 *
 * ub_value_to_half: value (G*Nv, H*Dh) fp32 token-major  ->  value16 (G, H, Nv, Dh) fp16 (saturating). */
int ub_value_to_half(const float* value, void* value16, int G, int Nv, int H, int Dh, ub_stream_t stream);
/* value16 (B, H, fH*fW, 32) fp16; qproj / out as in ub_bev_sample_fwd (ld, off_col, logit_col multiples of 4).
 * out_f16 != 0: `out` is (B, Nq, H*32) fp16 instead of fp32 -- the A operand of ub_linear_f16 (the output projection);
 * the rounding is the one a TF32 projection would apply to its operand anyway (11-bit significand). */
/* workspace: 2 ints owned by the caller, zero before the first use and left zero by every call (the persistent CTAs
 * draw their work units from it); calls that may run concurrently need distinct workspaces.
 * flags bit 0 (UB_WIN_ROUND_TF32): round the fp32 outputs to the nearest TF32 value -- for callers that feed them to
 * ub_linear_tf32, whose operand fetch truncates instead. */
enum { UB_WIN_ROUND_TF32 = 1 };
int ub_bev_sample_win_fwd(const void* value16, const float* qproj, void* out, int out_f16,
                          int B, int bev_h, int bev_w, int fH, int fW, int H, int Dh, int P,
                          int ld, int off_col, int logit_col, int* workspace, int flags, ub_stream_t stream);
/* fp32 twin of ub_bev_sample_win_fwd (the default precision class): planes32 (B, 2 H, fH*fW, 16) fp32 half-head planes
 * from ub_linear_tf32x3 (channel c of head h lives in plane 2 h + c / 16), fp32 weights, exact softmax; out (B, Nq, H*32)
 * fp32.  Same shapes covered (head dim 32, 4 or 8 points), same workspace contract. */
int ub_bev_sample_win32_fwd(const float* planes32, const float* qproj, float* out,
                            int B, int bev_h, int bev_w, int fH, int fW, int H, int Dh, int P,
                            int ld, int off_col, int logit_col, int* workspace, ub_stream_t stream);
/* mask (B, Nq, N) from ub_project_points -> the queries batch item 0 sees in camera n
 * (spatial_cross_attention_img.py:141-152), split by rank: hit_idx (N + 1, Nq) int32, row n = the hits whose lowest
 * seeing camera is n ("first", ascending, from the front) and the other hits of camera n ("later", from the back:
 * hit_idx[n][Nq - 1 - k]); row N = the queries no camera sees.  hit_cnt (2 N + 1) int32 = first counts, later
 * counts, unseen count.  inv_cnt (B, Nq) = 1 / max(1, #cameras whose mask for (b, q) is non-zero) (:209-212).
 * hit_ic (B, N, Nq) or NULL: inv_cnt in hit-list order (hit_ic[b][n][pos] = inv_cnt[b][hit_idx[n][pos]]), which the
 * sampling kernel reads coalesced. */
int ub_build_hits(const uint8_t* mask, int* hit_idx, int* hit_cnt, float* inv_cnt, float* hit_ic, int B, int N, int Nq,
                  ub_stream_t stream);
/* Hit-list-ordered inputs of ub_img_sample_win32_fwd, once per frame (hit_idx / hit_cnt / inv_cnt from ub_build_hits):
 * q_dst (Nq, N) int32 = for query q the rows n * Nq + pos (pos = its position in camera n's row of hit_idx) of every camera
 * that sees it, -1 padded -- the scatter map of ub_linear_tf32x3_scatter; hit_ref (B, N, Nq, 2 D) = ref_cam (B, Nq, N, D, 2)
 * gathered into hit-list order; hit_meta (B, N, Nq, 4) fp32 records {query index (int bits), 1 / #cameras, 0, 0}. */
int ub_hit_order(const uint8_t* mask, const float* ref_cam, const int* hit_idx, const int* hit_cnt, const float* inv_cnt,
                 int* q_dst, float* hit_ref, float* hit_meta, int B, int N, int Nq, int D, ub_stream_t stream);
/* fp32 twin of ub_img_sample_win_fwd (the default precision class).  planes32 (B*N, 2 H, fH*fW, 16) fp32 half-head planes;
 * qp_hit (B, N, Nq, ld) the offset|logit rows in hit-list order (ub_linear_tf32x3_scatter; rows that are no hit are never
 * read as valid and may be uninitialised); hit_ref / hit_meta from ub_hit_order; hit_idx / hit_cnt from ub_build_hits.
 * Every row of out (B, Nq, H*32) is written (zero rows for unseen queries, later hits accumulated with red.global.add).
 * Head dim 32, 4 or 8 points, an even number of Z-anchors. */
int ub_img_sample_win32_fwd(const float* planes32, const float* qp_hit, const float* hit_ref, const float* hit_meta,
                            const int* hit_idx, const int* hit_cnt, float* out, int B, int N, int bev_h, int bev_w,
                            int fH, int fW, int H, int Dh, int P, int D, int ld, int off_col, int logit_col,
                            ub_stream_t stream);
/* value16 (B, N, H, fH*fW, 32) fp16; hit_idx / hit_cnt / inv_cnt from ub_build_hits.  Every row of out
 * (B, Nq, H*32) is written: first hits with plain stores (zero rows for unseen queries), later hits accumulated with
 * red.global.add on top (sums over more than three cameras are order-dependent in the last bit).  out_f16 as in
 * ub_bev_sample_win_fwd (later hits are then accumulated in fp16). */
int ub_img_sample_win_fwd(const void* value16, const float* qproj, const float* ref_cam, const int* hit_idx,
                          const int* hit_cnt, const float* inv_cnt, const float* hit_ic /* or NULL */, void* out,
                          int out_f16, int B, int N, int bev_h,
                          int bev_w, int fH, int fW, int H, int Dh, int P, int D, int ld, int off_col, int logit_col,
                          ub_stream_t stream);

/* ---- [R5] dense projection on the tcgen05 tensor cores with fused epilogue -------------------------
 * acc = A (M, K) @ W (N, K)^T, TF32 inputs, fp32 accumulation (the nn.Linear layers of the encoder:
 * value_proj / sampling_offsets|attention_weights / output_proj / FFN, e.g. spatial_cross_attention_img.py:381-389,
 * :212-215, and mmcv FFN).  A and W contiguous row-major fp32.
 *   flags bit 0: ReLU.  flags bit 1: LayerNorm over the N outputs of (acc + bias + residual) with gamma / beta / eps
 *   (N <= 256) -- the 'norm' step that follows every attention / FFN, encoder_unibev_detr_img.py:434-436.
 *   residual (M, N) row stride ldr, or NULL.  out (M, N) row stride ldc.
 *   planes != NULL: instead of `out`, write fp16(acc + bias) head-major (M / Nv, N / 32, Nv, 32), the value-map
 *   layout of ub_bev_sample_win_fwd / ub_img_sample_win_fwd (no ReLU / residual / LayerNorm).
 * K % 32 == 0, N % 32 == 0 and (N <= 256 or N % 256 == 0), else UB_EUNSUPPORTED. */
enum { UB_LIN_RELU = 1, UB_LIN_LAYERNORM = 2 };
int ub_linear_tf32(const float* A, const float* W, const float* bias, const float* residual, int ldr,
                   const float* gamma, const float* beta, float eps, float* out, int ldc, void* planes, int Nv,
                   int M, int N, int K, int flags, ub_stream_t stream);

/* ub_linear_tf32 that also writes an fp16 copy `out16` (row stride ldc16) of the result rows. */
int ub_linear_tf32_dual(const float* A, const float* W, const float* bias, const float* residual, int ldr,
                        const float* gamma, const float* beta, float eps, float* out, int ldc, void* out16, int ldc16,
                        int M, int N, int K, int flags, ub_stream_t stream);
/* The same projection with fp16 operands A16 (M, K) / W16 (N, K) (the significand of TF32 in half the bytes; the
 * weight tile then stays resident in shared memory) and, optionally, an fp16 copy `out16` (row stride ldc16) of
 * the result next to / instead of the fp32 `out`: the A operand of the next projection.  K % 64 == 0. */
int ub_linear_f16(const void* A16, const void* W16, const float* bias, const float* residual, int ldr,
                  const float* gamma, const float* beta, float eps, float* out, int ldc, void* out16, int ldc16,
                  void* planes, int Nv, int M, int N, int K, int flags, ub_stream_t stream);

/* fp32-grade projection on the tensor cores ("3xTF32"): every product a*w is evaluated as a_hi*w_hi + a_lo*w_hi +
 * a_hi*w_lo with three tcgen05 kind::tf32 MMAs into the same fp32 accumulator (x_hi = upper 19 bits of x, x_lo = x - x_hi),
 * error ~2^-21 relative per product -- the arithmetic class of the reference's fp32 nn.Linear
 * (spatial_cross_attention_img.py:381-389, encoder_unibev_detr_img.py:434-436,476-479).  A (M, K) fp32 is split on chip;
 * W_hi / W_lo (N, K) come from ub_split_tf32 (once per weight).  Epilogues, flags and shape limits as ub_linear_tf32.
 *   planes32 != NULL: instead of `out`, write fp32 (acc + bias) as half-head planes (M / Nv, N / 16, Nv, 16), the value-map
 *   layout of the fp32 window-staged sampling kernels (no ReLU / residual / LayerNorm). */
int ub_linear_tf32x3(const float* A, const float* W_hi, const float* W_lo, const float* bias, const float* residual, int ldr,
                     const float* gamma, const float* beta, float eps, float* out, int ldc, float* planes32, int Nv,
                     int M, int N, int K, int flags, ub_stream_t stream);
/* ub_linear_tf32x3 (bias only) whose result rows leave in another order: row b * rows_per_item + q of the product is
 * written to the rows b * dst_rows_per_item + scatter[q * scatter_r + j] of `out` (row stride ldc), j = 0, 1, ... up to the
 * first negative entry (none: the row is dropped).  With scatter = q_dst of ub_hit_order the camera cross-attention's
 * offset|logit projection writes its rows in hit-list order (spatial_cross_attention_img.py:156-170 rebatches the queries
 * per camera BEFORE the projection; the effect is the same). */
int ub_linear_tf32x3_scatter(const float* A, const float* W_hi, const float* W_lo, const float* bias, float* out, int ldc,
                             const int* scatter, int scatter_r, int rows_per_item, int dst_rows_per_item,
                             int M, int N, int K, ub_stream_t stream);
/* "fp16 x3": the three-product scheme of ub_linear_tf32x3 on kind::f16 MMAs (x_hi = fp16(x), x_lo = fp16(x - x_hi): two 11-bit
 * significands, the same precision class, twice the tensor-core rate, half the operand bytes).  VALID ONLY where the caller can
 * bound |A|: every element of A times a_scale must stay below the fp16 range (65504) -- unibev_b200/plugin/fused.py derives the
 * bound from the LayerNorm / projection weights and picks a_scale (a power of two) from it; operands it cannot bound go through
 * ub_linear_tf32x3.  W16_hi / W16_lo (N, K) fp16 and col_scale (N) from ub_split_f16 (row-wise power-of-two scaling, undone
 * exactly in the epilogue).  A (M, K) fp32 is scaled and split on chip.  K % 64 == 0; out / residual 32-byte aligned with
 * row strides % 8 == 0.  Epilogues as ub_linear_tf32x3; scatter != NULL: rows leave as in ub_linear_tf32x3_scatter. */
int ub_linear_f16x3(const float* A, float a_scale, const void* W16_hi, const void* W16_lo, const float* col_scale,
                    const float* bias, const float* residual, int ldr, const float* gamma, const float* beta, float eps,
                    float* out, int ldc, float* planes32, int Nv, const int* scatter, int scatter_r, int rows_per_item,
                    int dst_rows_per_item, int M, int N, int K, int flags, ub_stream_t stream);
/* ub_linear_f16x3 whose activation bound is only known on the device: |A| <= *bound_dev * bound_mul + bound_add, bound_dev a
 * device float written by an earlier kernel on the stream (ub_flatten_feats_max: the largest magnitude of an input tensor),
 * bound_mul / bound_add what the caller proves about the path from there (row sums / bias magnitudes of a projection in
 * between).  The kernel derives a_scale from it; col_scale = 1 / s_n from ub_split_f16 with a_scale = 1. */
int ub_linear_f16x3_dyn(const float* A, const float* bound_dev, float bound_mul, float bound_add, const void* W16_hi,
                        const void* W16_lo, const float* col_scale, const float* bias, const float* residual, int ldr,
                        const float* gamma, const float* beta, float eps, float* out, int ldc, float* planes32, int Nv,
                        int M, int N, int K, int flags, ub_stream_t stream);
/* w (rows, cols) fp32 -> hi16 / lo16 (rows, cols) fp16 of w[n, :] * s_n (s_n: the power of two that brings the row's largest
 * magnitude into [4096, 8192)), col_scale (rows) = 1 / (s_n a_scale); col_scale == NULL: no scaling. */
int ub_split_f16(const float* w, void* hi16, void* lo16, float* col_scale, int rows, int cols, float a_scale,
                 ub_stream_t stream);
/* Generic fp32 (FFMA) projection for shapes the tensor-core entry points reject (UB_EUNSUPPORTED): any M, N, K.
 * out (M, N; row stride ldc) = [relu](A (M, K) @ W (N, K)^T + bias + residual (row stride ldr)); bias / residual may be NULL. */
int ub_linear_simt(const float* A, const float* W, const float* bias, const float* residual, int ldr, float* out, int ldc,
                   int M, int N, int K, int relu, ub_stream_t stream);
/* w (n) -> hi (n) = w with the 13 low mantissa bits cleared, lo (n) = round_to_tf32(w - hi). */
int ub_split_tf32(const float* w, float* hi, float* lo, int64_t n, ub_stream_t stream);

/* ---- [R5] y = LayerNorm(x + bias + residual) * gamma + beta over the last dim C ---------------------
 * bias (C) and residual (rows, C) may be NULL.  C % 4 == 0, C <= 1024.  out may alias x. */
int ub_add_layernorm(const float* x, const float* bias, const float* residual, const float* gamma,
                     const float* beta, float* out, int64_t rows, int C, float eps, ub_stream_t stream);
/* ---- training step: streaming reductions of the backward pass (BASELINE configs[4]) ------------------
 * ub_colsum: out (N) += column sums of x (M, N): the bias gradient of a projection (autograd of the nn.Linear layers,
 * spatial_cross_attention_img.py:59,285-289).  out is ACCUMULATED into (zero it first).  N % 4 == 0, N <= 1024.
 * ub_layernorm_bwd: backward of y = LayerNorm(x + residual) * gamma + beta over the last dim C (the 'norm' steps with the
 * `dropout(out) + identity` that precedes them, encoder_unibev_detr_img.py:434-436,476-479; residual may be NULL): dx (rows, C)
 * written -- the gradient of the sum, i.e. of x and of residual alike; dgamma (C), dbeta (C) ACCUMULATED into (zero them
 * first).  Row statistics are recomputed from the inputs.  C % 4 == 0, C <= 1024. */
int ub_colsum(const float* x, float* out, int64_t M, int N, ub_stream_t stream);
/* y = LayerNorm(dropout(x, p) + residual) * gamma + beta in one pass per direction: the `self.dropout(out) + identity` of the
 * attentions / FFN (spatial_cross_attention_img.py:215, mmcv FFN) and the 'norm' step that follows
 * (encoder_unibev_detr_img.py:434-436,476-479).  mask (rows * C / 4 bytes): keep bits of every float4, written by _fwd, read
 * by _bwd.  rng_state: two int64 on the device {seed, step}; the caller advances step once per training step and numbers the
 * call sites of a step (call_site), so that no mask repeats.  The drop probability actually applied is round(p 65536) / 65536.
 * _bwd: dx = gradient of x (through the mask), dresidual = gradient of the sum; dgamma / dbeta ACCUMULATED into. */
int ub_dropout_add_layernorm_fwd(const float* x, const float* residual, const float* gamma, const float* beta, float* out,
                                 uint8_t* mask, int64_t rows, int C, float eps, float p, const int64_t* rng_state,
                                 int call_site, ub_stream_t stream);
int ub_dropout_add_layernorm_bwd(const float* x, const float* residual, const uint8_t* mask, const float* dy,
                                 const float* gamma, float* dx, float* dresidual, float* dgamma, float* dbeta, int64_t rows,
                                 int C, float eps, float p, ub_stream_t stream);

int ub_layernorm_bwd(const float* x, const float* residual, const float* dy, const float* gamma, float* dx, float* dgamma,
                     float* dbeta, int64_t rows, int C, float eps, ub_stream_t stream);
/* The same, additionally writing an fp16 copy `out16` (rows, C) of the result (may be NULL). */
int ub_add_layernorm16(const float* x, const float* bias, const float* residual, const float* gamma,
                       const float* beta, float* out, void* out16, int64_t rows, int C, float eps, ub_stream_t stream);

/* ---- [R6] channel-normalised-weight fusion --------------------------------------------------------
 * img / pts (rows, C), either may be NULL (missing modality == zeros).  w_img / w_pts (C) CNW parameters
 * or NULL for feature_norm=None.  s_img / s_pts (rows_per_item) spatial-norm parameters or NULL.
 * modal_embed (C_out) or NULL.  out (rows, C) for LINEAR/AVG, (rows, 2C) for CAT. */
int ub_cnw_fuse(const float* img, const float* pts, const float* w_img, const float* w_pts,
                const float* s_img, const float* s_pts, const float* modal_embed, float* out,
                int64_t rows, int rows_per_item, int C, int mode, int c_flag, int l_flag, ub_stream_t stream);

/* ---- [R7] backbone feature map -> token-major value input -----------------------------------------
 * in (G, C, HW) -> out (G, HW, C) with out[g, p, c] = in[g, c, p] + embed_a[g % n_a, c] + embed_b[c];
 * embed_a / embed_b may be NULL. */
int ub_flatten_feats(const float* in, const float* embed_a, int n_a, const float* embed_b, float* out,
                     int G, int C, int HW, ub_stream_t stream);
/* ub_flatten_feats that also raises *absmax (a device float the caller zeroed) to the largest magnitude it wrote. */
int ub_flatten_feats_max(const float* in, const float* embed_a, int n_a, const float* embed_b, float* out, float* absmax,
                         int G, int C, int HW, ub_stream_t stream);
/* The same with an fp16 copy out16 (G, HW, C) next to / instead of the fp32 `out` (either may be NULL): the A operand
 * of the fp16 value projection (ub_linear_f16). */
int ub_flatten_feats16(const float* in, const float* embed_a, int n_a, const float* embed_b, float* out, void* out16,
                       int G, int C, int HW, ub_stream_t stream);
/* BEV query table (rows, C) repeated for the B samples of a batch (transformer_fusion.py:493-498): out32 (B, rows, C)
 * fp32 and / or out16 (B, rows, C) fp16 (either may be NULL).  C % 8 == 0. */
int ub_broadcast_rows(const float* src, int64_t rows, int C, int B, float* out32, void* out16, ub_stream_t stream);

/* ---- [R8] LiDAR hard voxelisation (integer index path, bit-exact with the sequential CPU algorithm) ----
 * points (N, C) fp32 with x, y, z in columns 0..2 (C >= 3).  voxel_size (3) and pc_range (6) are HOST arrays
 * (x, y, z order).  A point belongs to cell floor((p - range_min) / voxel_size) (fp32) and is dropped when the cell
 * is outside grid = round((range_max - range_min) / voxel_size).  Voxel ids follow the order of first occurrence
 * in the point list; a voxel keeps its first max_points points in point order; voxels beyond max_voxels are
 * dropped.  Outputs (sized for max_voxels): voxels (max_voxels, max_points, C), zero where empty;
 * coors (max_voxels, 3) int32 (z, y, x); num_points_per_voxel (max_voxels) int32; voxel_num (1) int32 on the
 * DEVICE (no host sync).  workspace: ub_voxelize_workspace_bytes(N) bytes, 256-byte aligned. */
int ub_voxelize_workspace_bytes(int num_points, size_t* bytes /* host */);
int ub_hard_voxelize(const float* points, int N, int C, const float* voxel_size /* host */,
                     const float* pc_range /* host */, int max_points, int max_voxels, float* voxels, int* coors,
                     int* num_points_per_voxel, int* voxel_num, void* workspace, size_t workspace_bytes,
                     ub_stream_t stream);
/* HardSimpleVFE: out (M, num_features) = sum over the points of a voxel / num_points_per_voxel. */
int ub_voxel_mean(const float* voxels, const int* num_points_per_voxel, int M, int max_points, int C,
                  int num_features, float* out, ub_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIBEV_B200_H_ */
