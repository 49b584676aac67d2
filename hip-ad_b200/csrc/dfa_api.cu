// dfa_api.cu — the extern "C" boundary declared in include/hipad_dfa.h.
#include "../../include/hipad_dfa.h"
#include "dfa_launch.h"

using namespace hipad;

namespace {
inline bool bad_dims(int bs, int cams, int num_feat, int C, int L, int A, int P, int G) {
    return bs <= 0 || cams <= 0 || num_feat <= 0 || C <= 0 || L <= 0 || A <= 0 || P <= 0 || G <= 0 || (C % G) != 0;
}
inline Dims mk(int bs, int cams, int num_feat, int C, int L, int A, int P, int G) {
    Dims d; d.bs = bs; d.cams = cams; d.num_feat = num_feat; d.C = C; d.L = L; d.A = A; d.P = P; d.G = G;
    return d;
}

int forward_common(ElemType t, float* output, const void* feat, const int32_t* shapes, const int32_t* starts,
                   const float* loc, const float* weights, int bs, int cams, int num_feat, int C, int L, int A,
                   int P, int G, void* stream) {
    if (!output || !feat || !shapes || !starts || !loc || !weights || bad_dims(bs, cams, num_feat, C, L, A, P, G))
        return HIPAD_DFA_ERR_BAD_ARGUMENT;
    FwdArgs a = {};
    a.type = t; a.out = output; a.feat = feat; a.shapes = shapes; a.starts = starts;
    a.loc = loc; a.weights = weights; a.fused = false;
    a.d = mk(bs, cams, num_feat, C, L, A, P, G);
    a.stream = reinterpret_cast<cudaStream_t>(stream);
    return launch_forward(a);
}

int fused_common(ElemType t, float* output, const void* feat, const int32_t* shapes, const int32_t* starts,
                 const float* key_points, const float* proj, const float* image_wh, const float* logits,
                 float* loc_out, int bs, int cams, int num_feat, int C, int L, int A, int P, int G, void* stream) {
    if (!output || !feat || !shapes || !starts || !key_points || !proj || !logits ||
        bad_dims(bs, cams, num_feat, C, L, A, P, G))
        return HIPAD_DFA_ERR_BAD_ARGUMENT;
    FwdArgs a = {};
    a.type = t; a.out = output; a.feat = feat; a.shapes = shapes; a.starts = starts;
    a.weights = logits; a.key_points = key_points; a.proj = proj; a.image_wh = image_wh; a.loc_out = loc_out;
    a.fused = true;
    a.d = mk(bs, cams, num_feat, C, L, A, P, G);
    a.stream = reinterpret_cast<cudaStream_t>(stream);
    return launch_forward(a);
}

int backward_common(ElemType t, const void* feat, const int32_t* shapes, const int32_t* starts, const float* loc,
                    const float* weights, const float* grad_output, void* g_feat, float* g_loc, float* g_w, int bs,
                    int cams, int num_feat, int C, int L, int A, int P, int G, void* workspace, size_t workspace_bytes,
                    void* stream, int stage_mask = 7) {
    if (!feat || !shapes || !starts || !loc || !weights || !grad_output || !g_loc || !g_w ||
        bad_dims(bs, cams, num_feat, C, L, A, P, G))
        return HIPAD_DFA_ERR_BAD_ARGUMENT;
    BwdArgs a = {};
    a.type = t; a.feat = feat; a.shapes = shapes; a.starts = starts; a.loc = loc; a.weights = weights;
    a.grad_out = grad_output; a.g_feat = g_feat; a.g_loc = g_loc; a.g_w = g_w;
    a.d = mk(bs, cams, num_feat, C, L, A, P, G);
    a.workspace = workspace; a.workspace_bytes = workspace_bytes;
    a.stream = reinterpret_cast<cudaStream_t>(stream);
    a.stage_mask = stage_mask & 7;
    a.separate_zero_fill = (stage_mask & 8) != 0;
    a.classify_only = (stage_mask & 16) != 0;
    a.accumulate = (stage_mask & 32) != 0;
    return launch_backward(a);
}
}  // namespace

extern "C" {

int hipad_dfa_version(void) { return HIPAD_DFA_VERSION; }

const char* hipad_dfa_error_string(int status) {
    if (status == 0) return "success";
    if (status == HIPAD_DFA_ERR_BAD_ARGUMENT) return "hipad_dfa: bad argument (null pointer, non-positive size, or num_embeds % num_groups != 0)";
    if (status == HIPAD_DFA_ERR_UNSUPPORTED) return "hipad_dfa: shape outside the compiled kernel family";
    if (status == HIPAD_DFA_ERR_WORKSPACE) return "hipad_dfa: workspace null, misaligned or too small";
    if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
    return "hipad_dfa: unknown status";
}

int hipad_dfa_forward_f32(float* output, const float* mc_ms_feat, const int32_t* spatial_shape,
                          const int32_t* scale_start_index, const float* sample_location, const float* weights,
                          int batch_size, int num_cams, int num_feat, int num_embeds, int num_scale, int num_anchors,
                          int num_pts, int num_groups, void* stream) {
    return forward_common(kF32, output, mc_ms_feat, spatial_shape, scale_start_index, sample_location, weights,
                          batch_size, num_cams, num_feat, num_embeds, num_scale, num_anchors, num_pts, num_groups, stream);
}

int hipad_dfa_forward_bf16(float* output, const uint16_t* mc_ms_feat, const int32_t* spatial_shape,
                           const int32_t* scale_start_index, const float* sample_location, const float* weights,
                           int batch_size, int num_cams, int num_feat, int num_embeds, int num_scale, int num_anchors,
                           int num_pts, int num_groups, void* stream) {
    return forward_common(kBF16, output, mc_ms_feat, spatial_shape, scale_start_index, sample_location, weights,
                          batch_size, num_cams, num_feat, num_embeds, num_scale, num_anchors, num_pts, num_groups, stream);
}

size_t hipad_dfa_debug_counters_offset(int bs, int cams, int num_feat, int C, int L, int A, int P, int G) {
    if (bad_dims(bs, cams, num_feat, C, L, A, P, G)) return 0;
    return backward_counters_offset(mk(bs, cams, num_feat, C, L, A, P, G));
}

size_t hipad_dfa_backward_workspace_bytes(int batch_size, int num_cams, int num_feat, int num_embeds, int num_scale,
                                          int num_anchors, int num_pts, int num_groups) {
    if (bad_dims(batch_size, num_cams, num_feat, num_embeds, num_scale, num_anchors, num_pts, num_groups)) return 0;
    return backward_workspace_bytes(mk(batch_size, num_cams, num_feat, num_embeds, num_scale, num_anchors, num_pts, num_groups));
}

int hipad_dfa_backward_f32(const float* mc_ms_feat, const int32_t* spatial_shape, const int32_t* scale_start_index,
                           const float* sample_location, const float* weights, const float* grad_output,
                           float* grad_mc_ms_feat, float* grad_sampling_location, float* grad_weights, int batch_size,
                           int num_cams, int num_feat, int num_embeds, int num_scale, int num_anchors, int num_pts,
                           int num_groups, void* workspace, size_t workspace_bytes, void* stream) {
    return backward_common(kF32, mc_ms_feat, spatial_shape, scale_start_index, sample_location, weights, grad_output,
                           grad_mc_ms_feat, grad_sampling_location, grad_weights, batch_size, num_cams, num_feat,
                           num_embeds, num_scale, num_anchors, num_pts, num_groups, workspace, workspace_bytes, stream);
}

int hipad_dfa_backward_bf16(const uint16_t* mc_ms_feat, const int32_t* spatial_shape,
                            const int32_t* scale_start_index, const float* sample_location, const float* weights,
                            const float* grad_output, uint16_t* grad_mc_ms_feat, float* grad_sampling_location,
                            float* grad_weights, int batch_size, int num_cams, int num_feat, int num_embeds,
                            int num_scale, int num_anchors, int num_pts, int num_groups, void* workspace,
                            size_t workspace_bytes, void* stream) {
    return backward_common(kBF16, mc_ms_feat, spatial_shape, scale_start_index, sample_location, weights, grad_output,
                           grad_mc_ms_feat, grad_sampling_location, grad_weights, batch_size, num_cams, num_feat,
                           num_embeds, num_scale, num_anchors, num_pts, num_groups, workspace, workspace_bytes, stream);
}

int hipad_dfa_backward_accumulate_f32(const float* mc_ms_feat, const int32_t* spatial_shape,
                                      const int32_t* scale_start_index, const float* sample_location,
                                      const float* weights, const float* grad_output, float* grad_mc_ms_feat,
                                      float* grad_sampling_location, float* grad_weights, int batch_size, int num_cams,
                                      int num_feat, int num_embeds, int num_scale, int num_anchors, int num_pts,
                                      int num_groups, void* workspace, size_t workspace_bytes, void* stream) {
    if (!grad_mc_ms_feat) return HIPAD_DFA_ERR_BAD_ARGUMENT;
    return backward_common(kF32, mc_ms_feat, spatial_shape, scale_start_index, sample_location, weights, grad_output,
                           grad_mc_ms_feat, grad_sampling_location, grad_weights, batch_size, num_cams, num_feat,
                           num_embeds, num_scale, num_anchors, num_pts, num_groups, workspace, workspace_bytes, stream,
                           7 | 32);
}

int hipad_dfa_backward_accumulate_bf16(const uint16_t* mc_ms_feat, const int32_t* spatial_shape,
                                       const int32_t* scale_start_index, const float* sample_location,
                                       const float* weights, const float* grad_output, uint16_t* grad_mc_ms_feat,
                                       float* grad_sampling_location, float* grad_weights, int batch_size, int num_cams,
                                       int num_feat, int num_embeds, int num_scale, int num_anchors, int num_pts,
                                       int num_groups, void* workspace, size_t workspace_bytes, void* stream) {
    if (!grad_mc_ms_feat) return HIPAD_DFA_ERR_BAD_ARGUMENT;
    return backward_common(kBF16, mc_ms_feat, spatial_shape, scale_start_index, sample_location, weights, grad_output,
                           grad_mc_ms_feat, grad_sampling_location, grad_weights, batch_size, num_cams, num_feat,
                           num_embeds, num_scale, num_anchors, num_pts, num_groups, workspace, workspace_bytes, stream,
                           7 | 32);
}

int hipad_dfa_backward_stages(int feat_is_bf16, int stage_mask, const void* mc_ms_feat, const int32_t* spatial_shape,
                              const int32_t* scale_start_index, const float* sample_location, const float* weights,
                              const float* grad_output, void* grad_mc_ms_feat, float* grad_sampling_location,
                              float* grad_weights, int batch_size, int num_cams, int num_feat, int num_embeds,
                              int num_scale, int num_anchors, int num_pts, int num_groups, void* workspace,
                              size_t workspace_bytes, void* stream) {
    return backward_common(feat_is_bf16 ? kBF16 : kF32, mc_ms_feat, spatial_shape, scale_start_index, sample_location,
                           weights, grad_output, grad_mc_ms_feat, grad_sampling_location, grad_weights, batch_size,
                           num_cams, num_feat, num_embeds, num_scale, num_anchors, num_pts, num_groups, workspace,
                           workspace_bytes, stream, stage_mask & 63);
}

static int fill_calls(CallDesc* dst, const hipad_dfa_call_t* calls, int num_calls, bool backward) {
    if (!calls || num_calls < 1 || num_calls > HIPAD_DFA_MAX_GROUP_CALLS) return HIPAD_DFA_ERR_BAD_ARGUMENT;
    for (int k = 0; k < num_calls; ++k) {
        const hipad_dfa_call_t& c = calls[k];
        if (!c.sample_location || !c.weights || c.num_anchors <= 0 || c.num_pts <= 0) return HIPAD_DFA_ERR_BAD_ARGUMENT;
        if (backward && (!c.grad_sampling_location || !c.grad_weights)) return HIPAD_DFA_ERR_BAD_ARGUMENT;
        dst[k].loc = c.sample_location; dst[k].weights = c.weights;
        dst[k].g_loc = c.grad_sampling_location; dst[k].g_w = c.grad_weights;
        dst[k].A = c.num_anchors; dst[k].P = c.num_pts;
    }
    return 0;
}

size_t hipad_dfa_group_forward_workspace_bytes(const hipad_dfa_call_t* calls, int num_calls, int batch_size,
                                               int num_cams, int num_embeds) {
    CallDesc cd[kMaxCalls] = {};
    if (batch_size <= 0 || num_cams <= 0 || num_embeds <= 0) return 0;
    if (!calls || num_calls < 1 || num_calls > HIPAD_DFA_MAX_GROUP_CALLS) return 0;
    for (int k = 0; k < num_calls; ++k) { cd[k].A = calls[k].num_anchors; cd[k].P = calls[k].num_pts; }
    return group_forward_workspace_bytes(cd, num_calls, batch_size, num_cams, num_embeds);
}

int hipad_dfa_group_forward(int feat_is_bf16, float* output_packed, const void* mc_ms_feat,
                            const int32_t* spatial_shape, const int32_t* scale_start_index,
                            const hipad_dfa_call_t* calls, int num_calls, int batch_size, int num_cams, int num_feat,
                            int num_embeds, int num_scale, int num_groups, void* workspace, size_t workspace_bytes,
                            void* stream) {
    if (!output_packed || !mc_ms_feat || !spatial_shape || !scale_start_index ||
        bad_dims(batch_size, num_cams, num_feat, num_embeds, num_scale, 1, 1, num_groups))
        return HIPAD_DFA_ERR_BAD_ARGUMENT;
    GroupFwdArgs a = {};
    if (int rc = fill_calls(a.calls, calls, num_calls, false)) return rc;
    a.type = feat_is_bf16 ? kBF16 : kF32; a.out = output_packed; a.feat = mc_ms_feat;
    a.shapes = spatial_shape; a.starts = scale_start_index; a.ncalls = num_calls;
    a.bs = batch_size; a.cams = num_cams; a.num_feat = num_feat; a.C = num_embeds; a.L = num_scale; a.G = num_groups;
    a.workspace = workspace; a.workspace_bytes = workspace_bytes;
    a.stream = reinterpret_cast<cudaStream_t>(stream);
    return launch_group_forward(a);
}

size_t hipad_dfa_group_backward_workspace_bytes(const hipad_dfa_call_t* calls, int num_calls, int batch_size,
                                                int num_cams, int num_feat, int num_embeds, int num_scale,
                                                int num_groups) {
    CallDesc cd[kMaxCalls] = {};
    if (bad_dims(batch_size, num_cams, num_feat, num_embeds, num_scale, 1, 1, num_groups)) return 0;
    if (!calls || num_calls < 1 || num_calls > HIPAD_DFA_MAX_GROUP_CALLS) return 0;
    for (int k = 0; k < num_calls; ++k) {
        if (calls[k].num_anchors <= 0 || calls[k].num_pts <= 0) return 0;
        cd[k].A = calls[k].num_anchors; cd[k].P = calls[k].num_pts;
    }
    return group_backward_workspace_bytes(cd, num_calls, batch_size, num_cams, num_feat, num_embeds, num_scale, num_groups);
}

int hipad_dfa_group_backward(int feat_is_bf16, int grad_feat_flags, const void* mc_ms_feat,
                             const int32_t* spatial_shape, const int32_t* scale_start_index,
                             const hipad_dfa_call_t* calls, int num_calls, const float* grad_output_packed,
                             void* grad_mc_ms_feat, int batch_size, int num_cams, int num_feat, int num_embeds,
                             int num_scale, int num_groups, void* workspace, size_t workspace_bytes, void* stream) {
    if (!mc_ms_feat || !spatial_shape || !scale_start_index || !grad_output_packed ||
        bad_dims(batch_size, num_cams, num_feat, num_embeds, num_scale, 1, 1, num_groups))
        return HIPAD_DFA_ERR_BAD_ARGUMENT;
    if ((grad_feat_flags & 1) && !grad_mc_ms_feat) return HIPAD_DFA_ERR_BAD_ARGUMENT;
    GroupBwdArgs a = {};
    if (int rc = fill_calls(a.calls, calls, num_calls, true)) return rc;
    a.type = feat_is_bf16 ? kBF16 : kF32; a.feat = mc_ms_feat; a.shapes = spatial_shape; a.starts = scale_start_index;
    a.ncalls = num_calls; a.grad_out = grad_output_packed; a.g_feat = grad_mc_ms_feat;
    a.accumulate = (grad_feat_flags & 1) != 0;
    a.g_feat_f32 = !feat_is_bf16 || (grad_feat_flags & 2) != 0;
    a.bs = batch_size; a.cams = num_cams; a.num_feat = num_feat; a.C = num_embeds; a.L = num_scale; a.G = num_groups;
    a.workspace = workspace; a.workspace_bytes = workspace_bytes;
    a.stream = reinterpret_cast<cudaStream_t>(stream);
    a.stage_mask = 7;
    return launch_group_backward(a);
}

int hipad_dfa_group_backward_stages(int feat_is_bf16, int grad_feat_flags, int stage_mask, const void* mc_ms_feat,
                                    const int32_t* spatial_shape, const int32_t* scale_start_index,
                                    const hipad_dfa_call_t* calls, int num_calls, const float* grad_output_packed,
                                    void* grad_mc_ms_feat, int batch_size, int num_cams, int num_feat, int num_embeds,
                                    int num_scale, int num_groups, void* workspace, size_t workspace_bytes, void* stream) {
    if (!mc_ms_feat || !spatial_shape || !scale_start_index || !grad_output_packed || !grad_mc_ms_feat ||
        bad_dims(batch_size, num_cams, num_feat, num_embeds, num_scale, 1, 1, num_groups))
        return HIPAD_DFA_ERR_BAD_ARGUMENT;
    GroupBwdArgs a = {};
    if (int rc = fill_calls(a.calls, calls, num_calls, true)) return rc;
    a.type = feat_is_bf16 ? kBF16 : kF32; a.feat = mc_ms_feat; a.shapes = spatial_shape; a.starts = scale_start_index;
    a.ncalls = num_calls; a.grad_out = grad_output_packed; a.g_feat = grad_mc_ms_feat;
    a.accumulate = (grad_feat_flags & 1) != 0;
    a.g_feat_f32 = !feat_is_bf16 || (grad_feat_flags & 2) != 0;
    a.bs = batch_size; a.cams = num_cams; a.num_feat = num_feat; a.C = num_embeds; a.L = num_scale; a.G = num_groups;
    a.workspace = workspace; a.workspace_bytes = workspace_bytes;
    a.stream = reinterpret_cast<cudaStream_t>(stream);
    a.stage_mask = stage_mask & 7;
    a.separate_zero_fill = (stage_mask & 8) != 0;
    return launch_group_backward(a);
}

int hipad_dfa_sample_indices(int32_t* indices, const int32_t* spatial_shape, const int32_t* scale_start_index,
                             const float* sample_location, int batch_size, int num_cams, int num_scale,
                             int num_anchors, int num_pts, void* stream) {
    if (!indices || !spatial_shape || !scale_start_index || !sample_location || batch_size <= 0 || num_cams <= 0 ||
        num_scale <= 0 || num_anchors <= 0 || num_pts <= 0)
        return HIPAD_DFA_ERR_BAD_ARGUMENT;
    return launch_indices(indices, spatial_shape, scale_start_index, sample_location, batch_size, num_cams, num_scale,
                          num_anchors, num_pts, reinterpret_cast<cudaStream_t>(stream));
}

int hipad_dfa_fused_forward_f32(float* output, const float* mc_ms_feat, const int32_t* spatial_shape,
                                const int32_t* scale_start_index, const float* key_points,
                                const float* projection_mat, const float* image_wh, const float* logits,
                                float* sample_location_out, int batch_size, int num_cams, int num_feat,
                                int num_embeds, int num_scale, int num_anchors, int num_pts, int num_groups,
                                void* stream) {
    return fused_common(kF32, output, mc_ms_feat, spatial_shape, scale_start_index, key_points, projection_mat,
                        image_wh, logits, sample_location_out, batch_size, num_cams, num_feat, num_embeds, num_scale,
                        num_anchors, num_pts, num_groups, stream);
}

int hipad_dfa_fused_forward_bf16(float* output, const uint16_t* mc_ms_feat, const int32_t* spatial_shape,
                                 const int32_t* scale_start_index, const float* key_points,
                                 const float* projection_mat, const float* image_wh, const float* logits,
                                 float* sample_location_out, int batch_size, int num_cams, int num_feat,
                                 int num_embeds, int num_scale, int num_anchors, int num_pts, int num_groups,
                                 void* stream) {
    return fused_common(kBF16, output, mc_ms_feat, spatial_shape, scale_start_index, key_points, projection_mat,
                        image_wh, logits, sample_location_out, batch_size, num_cams, num_feat, num_embeds, num_scale,
                        num_anchors, num_pts, num_groups, stream);
}

}  // extern "C"
