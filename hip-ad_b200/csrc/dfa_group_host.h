// dfa_group_host.h — host-side planning of grouped launches, shared by dfa_group.cu and dfa_backward.cu.
#pragma once
#include "dfa_launch.h"

namespace hipad {

struct GroupParams;

struct GroupPlan {
    int S[kMaxCalls], PS[kMaxCalls];
    long long unit_begin[kMaxCalls], part_begin[kMaxCalls], row_begin[kMaxCalls];
    long long units, parts, rows;
    int ps_max;
    int order[kMaxCalls];    // slot i of the kernel's call table holds call order[i] (units are dealt slot by slot)
    size_t partial_bytes, ticket_bytes;
};

// units of a grouped launch: rows are cut into slices of at most ps_max (p,cam) pairs
GroupPlan plan_group(bool bwd, const CallDesc* calls, int ncalls, int bs, int cams, int C, int forced_single_slice);
int fill_group_params(GroupParams& gp, const GroupPlan& pl, const CallDesc* calls, int ncalls, int bs, int cams,
                      int num_feat, int C, int G, long long io_bstride, float* out, const float* grad_out);
int launch_group_sample(bool bwd, ElemType t, const GroupParams& gp, long long units, cudaStream_t st);

}  // namespace hipad
