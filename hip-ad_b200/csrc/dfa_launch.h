// dfa_launch.h — internal launch interface between the C ABI (dfa_api.cu) and the kernel
// translation units (dfa_forward.cu, dfa_backward.cu).  Not installed; see include/hipad_dfa.h.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "dfa_common.cuh"

namespace hipad {

enum ElemType { kF32 = 0, kBF16 = 1 };

struct KernelShape {
    bool vector;   // 16-byte vector path (V = 4 fp32 / 8 bf16) vs scalar path (V = 1)
    int nch;       // 32-lane channel chunks per row
    bool ok;
};

// picks the kernel family member for (C, G, alignment); ok=false -> HIPAD_DFA_ERR_UNSUPPORTED
KernelShape pick_shape(ElemType t, int C, int G, bool aligned16);
// point slices per output row: enough CTAs to fill the machine, and few enough pairs per slice that the
// slice's gather metadata (bytes_per_pair each) fits the shared-memory budget; 0 if max_slices cannot do it
int choose_slices(long long rows, int pairs, int target_ctas, int bytes_per_pair, int max_slices);

struct FwdArgs {
    ElemType type;
    float* out;
    const void* feat;
    const int* shapes;
    const int* starts;
    const float* loc;       // unfused
    const float* weights;   // unfused: softmaxed weights; fused: logits
    const float* key_points;
    const float* proj;
    const float* image_wh;
    float* loc_out;
    bool fused;
    Dims d;
    cudaStream_t stream;
};
int launch_forward(const FwdArgs& a);

struct BwdArgs {
    ElemType type;
    const void* feat;
    const int* shapes;
    const int* starts;
    const float* loc;
    const float* weights;
    const float* grad_out;
    void* g_feat;
    float* g_loc;
    float* g_w;
    Dims d;
    void* workspace;
    size_t workspace_bytes;
    cudaStream_t stream;
    int stage_mask;   // bit0 sample-major (g_w,g_loc) + zero fill of g_feat, bit1 compaction + band sort, bit2 reduce
    bool accumulate;           // g_feat += (shared buffer across calls): no zero fill, touched rows are read-modify-written
    bool classify_only;        // debugging: stop after the row classification (stage bit 4 without the reduce)
    bool separate_zero_fill;   // measurement: never fold the zero fill into the sample-major kernel
};
int launch_backward(const BwdArgs& a);
size_t backward_workspace_bytes(const Dims& d);
size_t backward_counters_offset(const Dims& d);   // debugging: where the 8 work counters live in the workspace

// ---- grouped launches: several aggregation calls that read the same feature maps (one decoder layer) -------------
constexpr int kMaxCalls = 8;
struct CallDesc {
    const float* loc;        // [bs, A, P, cams, 2]
    const float* weights;    // [bs, A, P, cams, L, G]
    float* g_loc;            // backward only
    float* g_w;              // backward only
    int A, P;
};
struct GroupFwdArgs {
    ElemType type;
    float* out;              // packed [bs, A_total, C]: call k owns rows [a_begin_k, a_begin_k + A_k) of every batch element
    const void* feat;
    const int* shapes;
    const int* starts;
    CallDesc calls[kMaxCalls];
    int ncalls;
    int bs, cams, num_feat, C, L, G;
    void* workspace;         // partial rows + tickets of sliced rows (group_forward_workspace_bytes)
    size_t workspace_bytes;
    cudaStream_t stream;
};
int launch_group_forward(const GroupFwdArgs& a);
size_t group_forward_workspace_bytes(const CallDesc* calls, int ncalls, int bs, int cams, int C);
// true when the grouped sample kernel (dfa_group.cuh) covers this layout; otherwise callers use the per-call kernels
bool group_kernel_supported(ElemType t, int C, int L, int G, int cams);

struct GroupBwdArgs {
    ElemType type;
    const void* feat;
    const int* shapes;
    const int* starts;
    CallDesc calls[kMaxCalls];
    int ncalls;
    const float* grad_out;   // packed [bs, A_total, C]
    void* g_feat;            // [bs, num_feat, C] or null (frozen features)
    bool g_feat_f32;         // g_feat is fp32 whatever the feature type (shared accumulation buffer)
    bool accumulate;         // g_feat += instead of zero fill + overwrite
    int bs, cams, num_feat, C, L, G;
    void* workspace;
    size_t workspace_bytes;
    cudaStream_t stream;
    int stage_mask;          // bit0 sample-major kernel (+ zero fill), bit1 compaction + sort, bit2 classify + reduce
    bool classify_only;
    bool separate_zero_fill;
};
int launch_group_backward(const GroupBwdArgs& a);
size_t group_backward_workspace_bytes(const CallDesc* calls, int ncalls, int bs, int cams, int num_feat, int C, int L, int G);
size_t group_backward_counters_offset(const CallDesc* calls, int ncalls, int bs, int cams, int num_feat, int C, int L, int G);

int launch_indices(int32_t* idx, const int* shapes, const int* starts, const float* loc, int bs, int cams, int L,
                   int A, int P, cudaStream_t stream);

}  // namespace hipad
