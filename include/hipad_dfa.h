/*
 * hipad_dfa.h — C ABI of the B200-native deformable feature aggregation (DFA).
 *
 * This is the drop-in boundary for HiP-AD's hot path.  Every entry point takes
 * plain device pointers + sizes + a CUDA stream and returns an int status
 * (0 = success, >0 = cudaError_t value, <0 = HIPAD_DFA_ERR_*).  No torch types,
 * no hidden device allocation: the caller owns every buffer,
 * including the workspaces.  All launches go to the given stream (the
 * backward additionally forks part of its work onto one helper stream per caller
 * stream, created on first use at the device's highest stream priority, joined before the call returns on every path, and kept for the life of
 * the process: the only process-wide state; a caller stream must be driven by one host thread at a time,
 * and cudaStreamPerThread callers run the backward without the helper) and are
 * CUDA-graph capturable (no host synchronisation, no legacy-stream use).
 *
 * Reference interfaces replaced (paths relative to the HiP-AD repository):
 *   projects/mmdet3d_plugin/ops/src/deformable_aggregation_cuda.cu:265-288
 *       void deformable_aggregation(float* output, const float* mc_ms_feat, ...)
 *       -> hipad_dfa_forward_f32 / _bf16
 *   projects/mmdet3d_plugin/ops/src/deformable_aggregation_cuda.cu:291-318
 *       void deformable_aggregation_grad(const float* mc_ms_feat, ..., float* grad_weights, ...)
 *       -> hipad_dfa_backward_f32 / _bf16 (+ hipad_dfa_backward_workspace_bytes)
 *   projects/mmdet3d_plugin/ops/src/deformable_aggregation.cpp:31-62, 86-124
 *       ATen glue (shape extraction, at::zeros) -> done by the Python host
 *       (hip-ad_b200/ops/deformable_aggregation.py) on top of this ABI.
 *
 * Alignment: mc_ms_feat, output, grad_output, grad_mc_ms_feat, weights and grad_weights should be 16-byte aligned
 * (the 16-byte vector kernels are used then; otherwise a scalar kernel family is); sample_location,
 * grad_sampling_location and sample_location_out MUST be 8-byte aligned, and grad_weights MUST be 16-byte aligned
 * when num_scale*num_groups is a multiple of 4 (HIPAD_DFA_ERR_BAD_ARGUMENT otherwise).
 *
 * Tensor layouts (row-major, identical to the reference, deformable_aggregation.cpp:22-28):
 *   mc_ms_feat        [bs, num_feat, num_embeds]               f32 or bf16
 *   spatial_shape     [num_cams, num_scale, 2]  (h, w)         int32, device
 *   scale_start_index [num_cams, num_scale]     absolute row   int32, device
 *   sample_location   [bs, num_anchors, num_pts, num_cams, 2]  f32 (x, y) normalised
 *   weights           [bs, num_anchors, num_pts, num_cams, num_scale, num_groups]  f32
 *   output            [bs, num_anchors, num_embeds]            f32
 *
 * Differences from the reference launchers, all deliberate:
 *   - outputs are fully written by the kernels: they need NOT be pre-zeroed
 *     (the reference accumulates with atomicAdd into at::zeros buffers);
 *   - results are deterministic (no floating-point atomics anywhere);
 *   - 64-bit offsets: bs*num_pts*num_embeds*num_anchors*num_cams*num_scale may exceed 2^31.
 */
#ifndef HIPAD_DFA_H_
#define HIPAD_DFA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HIPAD_DFA_VERSION 1

/* negative status codes (positive values are cudaError_t) */
#define HIPAD_DFA_ERR_BAD_ARGUMENT   (-1)  /* null pointer, non-positive dim, num_embeds % num_groups != 0 */
#define HIPAD_DFA_ERR_UNSUPPORTED    (-2)  /* shape outside the compiled kernel family */
#define HIPAD_DFA_ERR_WORKSPACE      (-3)  /* workspace pointer null / too small / misaligned */

int hipad_dfa_version(void);
const char *hipad_dfa_error_string(int status);

/* ---- forward: replaces deformable_aggregation() (deformable_aggregation_cuda.cu:265-288) ---- */
int hipad_dfa_forward_f32(float *output, const float *mc_ms_feat,
                          const int32_t *spatial_shape, const int32_t *scale_start_index,
                          const float *sample_location, const float *weights,
                          int batch_size, int num_cams, int num_feat, int num_embeds,
                          int num_scale, int num_anchors, int num_pts, int num_groups,
                          void *stream);

/* same, feature maps stored as bf16 (raw uint16 bits); fp32 accumulation, fp32 output */
int hipad_dfa_forward_bf16(float *output, const uint16_t *mc_ms_feat,
                           const int32_t *spatial_shape, const int32_t *scale_start_index,
                           const float *sample_location, const float *weights,
                           int batch_size, int num_cams, int num_feat, int num_embeds,
                           int num_scale, int num_anchors, int num_pts, int num_groups,
                           void *stream);

/* ---- backward: replaces deformable_aggregation_grad() (deformable_aggregation_cuda.cu:291-318) ----
 * Bytes of scratch the backward needs (compacted visible samples, sorted records, segment
 * tables).  Pure host arithmetic on the sizes; valid for any spatial_shape / scale_start_index
 * whose (cam, level) row ranges are disjoint and lie inside [0, num_feat). */
size_t hipad_dfa_backward_workspace_bytes(int batch_size, int num_cams, int num_feat,
                                          int num_embeds, int num_scale, int num_anchors,
                                          int num_pts, int num_groups);

/* Writes ALL of grad_mc_ms_feat [bs,num_feat,C], grad_sampling_location and grad_weights
 * (zeros where nothing contributes).  grad_mc_ms_feat may be NULL: the feature-gradient pass is
 * then skipped (frozen backbone) and workspace may be NULL.  workspace: device memory, 256-byte aligned, at least
 * hipad_dfa_backward_workspace_bytes(...) bytes, private to this call until it completes. */
int hipad_dfa_backward_f32(const float *mc_ms_feat,
                           const int32_t *spatial_shape, const int32_t *scale_start_index,
                           const float *sample_location, const float *weights,
                           const float *grad_output,
                           float *grad_mc_ms_feat, float *grad_sampling_location, float *grad_weights,
                           int batch_size, int num_cams, int num_feat, int num_embeds,
                           int num_scale, int num_anchors, int num_pts, int num_groups,
                           void *workspace, size_t workspace_bytes, void *stream);

/* bf16 feature maps; grad_mc_ms_feat is written as bf16 (fp32 accumulation, one rounding) */
int hipad_dfa_backward_bf16(const uint16_t *mc_ms_feat,
                            const int32_t *spatial_shape, const int32_t *scale_start_index,
                            const float *sample_location, const float *weights,
                            const float *grad_output,
                            uint16_t *grad_mc_ms_feat, float *grad_sampling_location, float *grad_weights,
                            int batch_size, int num_cams, int num_feat, int num_embeds,
                            int num_scale, int num_anchors, int num_pts, int num_groups,
                            void *workspace, size_t workspace_bytes, void *stream);

/* Shared feature-gradient buffer ("next" row f1 of the scope table: one dense g_feat per decoder layer / per step
 * instead of one per call).  Same as hipad_dfa_backward_*, except that grad_mc_ms_feat is NOT zero-filled and the
 * touched rows are read-modify-written: grad_mc_ms_feat += this call's feature gradient.  The first call of a
 * group uses hipad_dfa_backward_* (which writes every row), the others this one, all on one stream; the
 * summation order is then the call order, i.e. still deterministic.  Removes the 115 MB zero fill per call and the
 * caller's (autograd's) dense accumulation of one g_feat per call. */
int hipad_dfa_backward_accumulate_f32(const float *mc_ms_feat,
                                      const int32_t *spatial_shape, const int32_t *scale_start_index,
                                      const float *sample_location, const float *weights,
                                      const float *grad_output,
                                      float *grad_mc_ms_feat, float *grad_sampling_location, float *grad_weights,
                                      int batch_size, int num_cams, int num_feat, int num_embeds,
                                      int num_scale, int num_anchors, int num_pts, int num_groups,
                                      void *workspace, size_t workspace_bytes, void *stream);

int hipad_dfa_backward_accumulate_bf16(const uint16_t *mc_ms_feat,
                                       const int32_t *spatial_shape, const int32_t *scale_start_index,
                                       const float *sample_location, const float *weights,
                                       const float *grad_output,
                                       uint16_t *grad_mc_ms_feat, float *grad_sampling_location, float *grad_weights,
                                       int batch_size, int num_cams, int num_feat, int num_embeds,
                                       int num_scale, int num_anchors, int num_pts, int num_groups,
                                       void *workspace, size_t workspace_bytes, void *stream);

/* Measurement variant of the backward: runs only the kernels selected by stage_mask
 * (bit0 = sample-major kernel writing grad_weights/grad_sampling_location, which also zero-fills
 * grad_mc_ms_feat; bit1 = visible-sample compaction + per-(b,cam,level,band) sort into the workspace;
 * bit2 = feature-major reduce overwriting the touched rows of grad_mc_ms_feat; 7 = the full backward;
 * bit3 = keep the zero fill in its own kernel instead of folding it into the sample-major kernel;
 * bit5 = accumulate into grad_mc_ms_feat as hipad_dfa_backward_accumulate_* does).
 * bench.py uses it to put CUDA events around each stage; results are only meaningful when the stages
 * are issued in order on one stream with the same workspace. */
int hipad_dfa_backward_stages(int feat_is_bf16, int stage_mask,
                              const void *mc_ms_feat,
                              const int32_t *spatial_shape, const int32_t *scale_start_index,
                              const float *sample_location, const float *weights,
                              const float *grad_output,
                              void *grad_mc_ms_feat, float *grad_sampling_location, float *grad_weights,
                              int batch_size, int num_cams, int num_feat, int num_embeds,
                              int num_scale, int num_anchors, int num_pts, int num_groups,
                              void *workspace, size_t workspace_bytes, void *stream);

/* ---- grouped launches ("next" row f1 of the scope table) ----
 * The aggregation calls of one decoder layer (det, map, plan, ego queries; sparse_onedecoder.py:867-887 of the
 * reference issues them one after the other) read the SAME feature maps.  A group runs them as ONE forward launch and
 * ONE backward chain, with ONE feature gradient for the whole group:
 *   output / grad_output are PACKED [bs, A_total, C] (A_total = sum of the calls' num_anchors; call k owns rows
 *   [sum_{j<k} A_j, +A_k) of every batch element), grad_mc_ms_feat is the sum of the calls' feature gradients.
 * Returns HIPAD_DFA_ERR_UNSUPPORTED for layouts the grouped kernels do not cover (anything but 4 levels, <= 8
 * groups, C = 128/256 f32 or C = 256 bf16): callers then issue the calls one by one. */
#define HIPAD_DFA_MAX_GROUP_CALLS 8
typedef struct hipad_dfa_call_t {
    const float *sample_location;      /* [bs, A, P, cams, 2], 8-byte aligned */
    const float *weights;              /* [bs, A, P, cams, L, G], 16-byte aligned */
    float *grad_sampling_location;     /* backward only, 8-byte aligned */
    float *grad_weights;               /* backward only, 16-byte aligned */
    int32_t num_anchors;
    int32_t num_pts;
} hipad_dfa_call_t;

size_t hipad_dfa_group_forward_workspace_bytes(const hipad_dfa_call_t *calls, int num_calls,
                                               int batch_size, int num_cams, int num_embeds);

/* workspace (256-byte aligned) holds the partial rows of output rows that are cut into several work units; it may be
 * NULL when every call has num_pts*num_cams <= 128. */
int hipad_dfa_group_forward(int feat_is_bf16, float *output_packed, const void *mc_ms_feat,
                            const int32_t *spatial_shape, const int32_t *scale_start_index,
                            const hipad_dfa_call_t *calls, int num_calls,
                            int batch_size, int num_cams, int num_feat, int num_embeds,
                            int num_scale, int num_groups,
                            void *workspace, size_t workspace_bytes, void *stream);

size_t hipad_dfa_group_backward_workspace_bytes(const hipad_dfa_call_t *calls, int num_calls,
                                                int batch_size, int num_cams, int num_feat,
                                                int num_embeds, int num_scale, int num_groups);

/* grad_feat_flags: bit0 accumulate into grad_mc_ms_feat (no zero fill, touched rows read-modify-written);
 *                  bit1 grad_mc_ms_feat is fp32 whatever the feature type (a shared accumulation buffer stays fp32
 *                       and is narrowed once by its owner).
 * grad_mc_ms_feat may be NULL (frozen features): only the calls' location / weight gradients are written. */
int hipad_dfa_group_backward(int feat_is_bf16, int grad_feat_flags, const void *mc_ms_feat,
                             const int32_t *spatial_shape, const int32_t *scale_start_index,
                             const hipad_dfa_call_t *calls, int num_calls,
                             const float *grad_output_packed, void *grad_mc_ms_feat,
                             int batch_size, int num_cams, int num_feat, int num_embeds,
                             int num_scale, int num_groups,
                             void *workspace, size_t workspace_bytes, void *stream);

/* Measurement variant of hipad_dfa_group_backward: runs only the kernels selected by stage_mask (bit0 sample-major
 * kernel + zero fill, bit1 compaction + sort, bit2 classification + reduce, bit3 zero fill as its own kernel), like
 * hipad_dfa_backward_stages.  grad_mc_ms_feat must not be NULL. */
int hipad_dfa_group_backward_stages(int feat_is_bf16, int grad_feat_flags, int stage_mask, const void *mc_ms_feat,
                                    const int32_t *spatial_shape, const int32_t *scale_start_index,
                                    const hipad_dfa_call_t *calls, int num_calls,
                                    const float *grad_output_packed, void *grad_mc_ms_feat,
                                    int batch_size, int num_cams, int num_feat, int num_embeds,
                                    int num_scale, int num_groups,
                                    void *workspace, size_t workspace_bytes, void *stream);

/* Debugging: byte offset of the backward's 8 work counters inside its workspace (part items, -, partial slots used,
 * tiny rows, reduce queue head, ...). */
size_t hipad_dfa_debug_counters_offset(int batch_size, int num_cams, int num_feat, int num_embeds,
                                       int num_scale, int num_anchors, int num_pts, int num_groups);

/* ---- integer sampling contract (parity instrument) ----
 * indices: int32 [bs, A, P, cams, L, 6] = {valid, h_low, w_low, level_offset, corner_mask, row0},
 * computed by the SAME device routine the kernels above use
 * (mirrors deformable_aggregation_cuda.cu:166-181 and :19-53). */
int hipad_dfa_sample_indices(int32_t *indices,
                             const int32_t *spatial_shape, const int32_t *scale_start_index,
                             const float *sample_location,
                             int batch_size, int num_cams, int num_scale, int num_anchors, int num_pts,
                             void *stream);

/* ---- fused module forward (inference): projection + group softmax + aggregation ----
 * Replaces, in one launch sequence, the chain of
 *   DeformableFeatureAggregation.project_points   (models/blocks.py:217-225)
 *   softmax over cams*L*P per group               (models/blocks.py:196-208)
 *   the two permute/contiguous copies             (models/blocks.py:138-158)
 *   deformable_aggregation()                      (deformable_aggregation_cuda.cu:265-288)
 *   key_points      [bs, A, P, 3]                 f32
 *   projection_mat  [bs, cams, 4, 4]              f32
 *   image_wh        [bs, cams, 2] or NULL         f32
 *   logits          [bs, A, cams, L, P, G]        f32, raw weights_fc output (pre-softmax)
 * Optional outputs (may be NULL): sample_location_out [bs,A,P,cams,2] as consumed by the sampler. */
int hipad_dfa_fused_forward_f32(float *output, const float *mc_ms_feat,
                                const int32_t *spatial_shape, const int32_t *scale_start_index,
                                const float *key_points, const float *projection_mat,
                                const float *image_wh, const float *logits,
                                float *sample_location_out,
                                int batch_size, int num_cams, int num_feat, int num_embeds,
                                int num_scale, int num_anchors, int num_pts, int num_groups,
                                void *stream);

int hipad_dfa_fused_forward_bf16(float *output, const uint16_t *mc_ms_feat,
                                 const int32_t *spatial_shape, const int32_t *scale_start_index,
                                 const float *key_points, const float *projection_mat,
                                 const float *image_wh, const float *logits,
                                 float *sample_location_out,
                                 int batch_size, int num_cams, int num_feat, int num_embeds,
                                 int num_scale, int num_anchors, int num_pts, int num_groups,
                                 void *stream);

/* ---- weights producer of the aggregation module, one pass each way ("next" row f2, training half) ----
 * Replaces the chain softmax over cams*L*P per group -> attn-drop mask (drawn on the CPU and copied over per call in the
 * reference) -> permute(0,1,4,2,3,5).contiguous()  (models/blocks.py:196-212, 147-158) and its mirror image in the backward.
 *   logits        [rows, cams, L, P, G]  raw weights_fc output, rows = bs*num_anchors
 *   weights       [rows, P, cams, L, G]  what hipad_dfa_forward_* / hipad_dfa_group_forward consume
 *   stats         [rows, G, 2]           (max, 1/sum) per group, saved for the backward (may be NULL in forward)
 *   keep_mask     [rows, cams, P] or NULL: explicit keep mask (1 = keep); NULL = drawn in the kernel from
 *                 (seed, row, cam, pt): keep iff uniform > drop_p, kept weights scaled by 1/(1-drop_p); drop_p = 0: none
 * num_groups must divide 256. */
int hipad_dfa_weights_forward(const float *logits, float *weights, float *stats, const float *keep_mask,
                              unsigned long long seed, float drop_p, long long rows,
                              int num_cams, int num_scale, int num_pts, int num_groups, void *stream);
int hipad_dfa_weights_backward(const float *logits, const float *stats, const float *grad_weights, float *grad_logits,
                               const float *keep_mask, unsigned long long seed, float drop_p, long long rows,
                               int num_cams, int num_scale, int num_pts, int num_groups, void *stream);

/* ---- feature_maps_format as one transposing pass ----
 * Replaces the cat + permute + flatten chain of feature_maps_format
 * (projects/mmdet3d_plugin/ops/__init__.py:74-103; two full copies of every feature map, the second a strided
 * transpose) by one read and one write:
 *   level_ptrs  HOST array of num_scale DEVICE pointers, level l = [bs, cams, C, H_l, W_l] contiguous
 *   level_hw    HOST array [num_scale][2] = (H_l, W_l)
 *   col_feats   device [bs, cams * sum_l(H_l*W_l), C]  (camera-major, level, row-major pixels, channels last)
 *   *_dtype     0 = f32, 1 = bf16 (raw uint16 bits); f32 levels may be narrowed to a bf16 col_feats on the way
 *   inverse     0: levels -> col_feats;  1: col_feats -> levels (the autograd of the format step, and the dense
 *               form of the inverse=True branch, ops/__init__.py:34-65) */
int hipad_dfa_format_features(int level_dtype, int col_dtype, int inverse,
                              void *const *level_ptrs, const int32_t *level_hw, void *col_feats,
                              int batch_size, int num_cams, int num_embeds, int num_scale, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HIPAD_DFA_H_ */
