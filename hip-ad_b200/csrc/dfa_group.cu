// dfa_group.cu — host side of the grouped sample-major kernel (dfa_group.cuh): unit planning, forward launcher,
// and the sample-major half of the grouped backward (called from dfa_backward.cu).
#include <algorithm>
#include "dfa_dispatch.cuh"
#include "dfa_group.cuh"
#include "dfa_group_host.h"

namespace hipad {

namespace {
constexpr size_t kAlignG = 256;
// CTAs of a launch that carries a dense zero fill (measured, det call backward: 8 / 12 / 18 / 24 per SM -> 113 / 106 / 104 / 102 us)
constexpr long long kZeroFillMinCtas = 148 * 24;
inline size_t align_g(size_t v) { return (v + kAlignG - 1) / kAlignG * kAlignG; }

#ifndef HIPAD_GROUP_WARPS
#define HIPAD_GROUP_WARPS 4
#endif
// warps per unit.  Measured on the stage-2 layer: 8-warp CTAs 1.3x slower (r2b); 2-warp CTAs at twice the CTAs per SM, 48- to
// 96-pair units: forward 101 vs 74 us, backward sample kernel 116-122 vs 103 us (x5)
constexpr int kGW = HIPAD_GROUP_WARPS;
constexpr int kGM = 4 / kGW;                 // CTAs per SM scale with the CTA size
inline int group_warps() { return kGW; }
// (p,cam) pairs per unit.  Longer units find more key points of an anchor in the same quad (fewer gathers) but hold
// more shared memory, which is taken from the L1 cache the gather lives on
inline int group_ps_max(bool bwd) {
    const int dflt = kGroupMaxPS;
    const int v = hipad_env_int(bwd ? "HIPAD_DFA_GROUP_PS_BWD" : "HIPAD_DFA_GROUP_PS_FWD", dflt);
    return (v >= 16 && v <= kGroupMaxPS) ? v : dflt;
}

template <typename T, int V, int NCH, bool kBwd, int kW, int kDepth, int kMinCtas, int kG>
int launch_inst_g(const GroupParams& gp, int grid, size_t smem, cudaStream_t st) {
    auto kern = dfa_group_kernel<T, V, NCH, kBwd, kW, kDepth, kMinCtas, kG>;
    cudaError_t e = ensure_smem(kern, smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kW * 32, smem, st>>>(gp);
    return (int)cudaGetLastError();
}

// G = 8 (every aggregation module of the reference's configs) gets the instantiation with G folded in.  Measured on the
// stage-2 layer, bs = 1 / 4: backward sample kernel f32 119 -> 107 / 421 -> 381 us, bf16 98 -> 90 us; forward bf16
// 93 -> 84 us; the fp32 FORWARD is slower with it (78 -> 84 / 251 -> 271 us: ptxas spills at its 64-register cap) and
// keeps the generic instantiation.
template <typename T, int V, int NCH, bool kBwd, int kW, int kDepth, int kMinCtas>
int launch_inst(const GroupParams& gp, int grid, size_t smem, cudaStream_t st) {
    constexpr bool kFold = kBwd || sizeof(T) == 2;
    if (kFold && gp.G == 8 && hipad_env_int("HIPAD_DFA_GROUP_GENERIC_G", 0) == 0)
        return launch_inst_g<T, V, NCH, kBwd, kW, kDepth, kMinCtas, kFold ? 8 : 0>(gp, grid, smem, st);
    return launch_inst_g<T, V, NCH, kBwd, kW, kDepth, kMinCtas, 0>(gp, grid, smem, st);
}

// deep = 2 quads in flight per warp for fp32 rows (4 for packed bf16), 128 registers, 4 CTAs of 4 warps per SM;
// shallow (default) = 1 (2) in flight, <= 80 registers, 6 CTAs per SM: more units resident, so the per-unit
// visibility / sort phases of some overlap the gather of others (measured, stage-2 layer: forward 92 vs 119 us,
// backward 311 vs 336 us; HIPAD_DFA_GROUP_DEEP=1 selects the deep variant)
template <bool kBwd>
int dispatch(ElemType t, int variant, const GroupParams& gp, int grid, size_t smem, cudaStream_t st) {
    // variant 0: deep pipeline, 4 CTAs/SM; 1: shallow, 6 CTAs/SM (<= 80 registers); 2: shallow, 8 CTAs/SM (64 registers)
    if (t == kF32) {
        if (gp.C == 128) return launch_inst<float, 4, 1, kBwd, kGW, 2, 5 * kGM>(gp, grid, smem, st);
        if (gp.C == 256) {
            if (variant == 0) return launch_inst<float, 4, 2, kBwd, kGW, 2, 4 * kGM>(gp, grid, smem, st);
            if (variant == 2) return launch_inst<float, 4, 2, kBwd, kGW, 1, 8 * kGM>(gp, grid, smem, st);
            return launch_inst<float, 4, 2, kBwd, kGW, 1, 6 * kGM>(gp, grid, smem, st);
        }
    } else {
        if (gp.C == 256) {
            if (variant == 0) return launch_inst<__nv_bfloat16, 8, 1, kBwd, kGW, 4, 4 * kGM>(gp, grid, smem, st);
            if (variant == 2) return launch_inst<__nv_bfloat16, 8, 1, kBwd, kGW, 2, 8 * kGM>(gp, grid, smem, st);
            return launch_inst<__nv_bfloat16, 8, 1, kBwd, kGW, 2, 5 * kGM>(gp, grid, smem, st);
        }
    }
    return -2;
}
}  // namespace

bool group_kernel_supported(ElemType t, int C, int L, int G, int cams) {
    if (hipad_env_int("HIPAD_DFA_GROUP_KERNEL", 1) == 0) return false;      // A/B knob: round-1 kernels only
    if (L != kGroupL || G < 1 || G > kGroupMaxG || (G & (G - 1)) != 0 || C % G != 0 || cams * L > kMaxCamLevels) return false;
    const int V = (t == kF32) ? 4 : 8;
    const int gd = C / G;
    if (gd % V != 0) return false;
    const int lpg = gd / V;
    if (lpg < 4 || lpg > 32 || (lpg & (lpg - 1)) != 0) return false;   // the backward's transposing butterfly needs >= 4 lanes
    if (t == kF32) return C == 128 || C == 256;
    return C == 256;
}

GroupPlan plan_group(bool bwd, const CallDesc* calls, int ncalls, int bs, int cams, int C, int forced_single_slice) {
    GroupPlan pl = {};
    const int ps_cap = group_ps_max(bwd);
    long long units = 0, parts = 0, rows = 0;
    int ps_max = 16;
    for (int k = 0; k < ncalls; ++k) {
        const int NP = calls[k].P * cams;
        int S = forced_single_slice ? 1 : (NP + ps_cap - 1) / ps_cap;
        if (S < 1) S = 1;
        int PS = (NP + S - 1) / S;
        if (!forced_single_slice) {                   // whole key points per slice (all cameras of a point together)
            const int up = (PS + cams - 1) / cams * cams;
            if (up <= ps_cap || up <= cams) PS = up;
        }
        S = (NP + PS - 1) / PS;                       // no empty slice
        pl.S[k] = S;
        pl.PS[k] = PS;
        pl.order[k] = k;
        if (PS > ps_max) ps_max = PS;
    }
    // CTAs are dispatched in index order: calls with the longest units go first, so the launch drains on short units
    // (no result depends on the order: every sum is ordered inside its unit / row)
    if (hipad_env_int("HIPAD_DFA_GROUP_ORDER", 1) != 0)
        std::stable_sort(pl.order, pl.order + ncalls, [&](int x, int y) { return pl.PS[x] > pl.PS[y]; });
    for (int i = 0; i < ncalls; ++i) {
        const int k = pl.order[i];
        const int S = pl.S[k];
        pl.unit_begin[k] = units;
        units += (long long)bs * calls[k].A * S;
        if (!bwd && S > 1) {
            pl.part_begin[k] = parts;
            pl.row_begin[k] = rows;
            parts += (long long)bs * calls[k].A * S;
            rows += (long long)bs * calls[k].A;
        }
    }
    pl.units = units;
    pl.parts = parts;
    pl.rows = rows;
    pl.ps_max = ps_max;
    pl.partial_bytes = align_g((size_t)parts * C * sizeof(float));
    pl.ticket_bytes = align_g((size_t)rows * sizeof(int));
    return pl;
}

size_t group_forward_workspace_bytes(const CallDesc* calls, int ncalls, int bs, int cams, int C) {
    if (ncalls < 1 || ncalls > kMaxCalls) return 0;
    const GroupPlan pl = plan_group(false, calls, ncalls, bs, cams, C, 0);
    return pl.partial_bytes + pl.ticket_bytes + kAlignG;
}

// fills the call table shared by the forward and backward launches; returns 0 or an error
int fill_group_params(GroupParams& gp, const GroupPlan& pl, const CallDesc* calls, int ncalls, int bs, int cams,
                      int num_feat, int C, int G, long long io_bstride, float* out, const float* grad_out) {
    if (pl.units <= 0 || pl.units > 0x7fffffffLL) return -2;
    gp.ncalls = ncalls; gp.bs = bs; gp.cams = cams; gp.num_feat = num_feat; gp.C = C; gp.G = G;
    gp.ps_max = pl.ps_max;
    long long a_begin[kMaxCalls];
    long long a_sum = 0;
    for (int k = 0; k < ncalls; ++k) { a_begin[k] = a_sum; a_sum += calls[k].A; }
    for (int i = 0; i < ncalls; ++i) {
        const int k = pl.order[i];
        GroupCall& c = gp.calls[i];                  // table slots in unit order (the kernel finds its call by unit_begin)
        c.loc = calls[k].loc; c.weights = calls[k].weights; c.g_loc = calls[k].g_loc; c.g_w = calls[k].g_w;
        c.out = out ? out + a_begin[k] * C : nullptr;
        c.grad_out = grad_out ? grad_out + a_begin[k] * C : nullptr;
        c.io_bstride = io_bstride;
        c.A = calls[k].A; c.P = calls[k].P; c.S = pl.S[k]; c.PS = pl.PS[k];
        c.unit_begin = (int)pl.unit_begin[k];
        c.part_begin = (int)pl.part_begin[k];
        c.row_begin = (int)pl.row_begin[k];
        // 32-bit element offsets inside one call's tensors are not assumed anywhere; pair counts are ints
        if ((long long)calls[k].P * cams > (1 << 24)) return -2;
    }
    return 0;
}

int launch_group_sample(bool bwd, ElemType t, const GroupParams& gp, long long units, cudaStream_t st) {
    const int kw = group_warps();
    const int V = (t == kF32) ? 4 : 8;
    const int nch = gp.C / (32 * V);
    GroupParams g = gp;
    g.so = group_smem_layout(bwd, gp.ps_max, nch * 32 * V, kw);
    const size_t smem = (size_t)g.so.total;
    if (smem > kSampleSmemBudget) return -2;
    // a launch that carries the dense zero fill is padded with fill-only CTAs (the fill then runs on the whole machine)
    long long grid = units;
    const long long fill_ctas = 148LL * hipad_env_int("HIPAD_DFA_FILL_CTAS_PER_SM", (int)(kZeroFillMinCtas / 148));
    if (bwd && g.zero_n16 > 0 && grid < fill_ctas) grid = fill_ctas;
    g.units = (int)units;
    g.zero_per = (grid > 0) ? (g.zero_n16 + grid - 1) / grid : 0;
    if (g.zero_per >= (1LL << 31)) return -2;
    // measured (stage-2 layer, f32): forward 82 us at 8 CTAs/SM vs 89 at 6 (bs=4: 257 vs 296); the backward needs more
    // registers (64 spills inside its epilogue) and is faster at 6 (118 vs 129 us)
    int variant = hipad_env_int("HIPAD_DFA_GROUP_CTAS", bwd ? 6 : 8) >= 8 ? 2 : 1;
    if (hipad_env_int("HIPAD_DFA_GROUP_DEEP", 0) != 0) variant = 0;
    return bwd ? dispatch<true>(t, variant, g, (int)grid, smem, st) : dispatch<false>(t, variant, g, (int)units, smem, st);
}

int launch_group_forward(const GroupFwdArgs& a) {
    if (a.ncalls < 1 || a.ncalls > kMaxCalls) return -1;
    if (!group_kernel_supported(a.type, a.C, a.L, a.G, a.cams)) return -2;
    if ((long long)a.num_feat * a.C >= (1LL << 30)) return -2;
    if (reinterpret_cast<uintptr_t>(a.feat) % 16 != 0 || reinterpret_cast<uintptr_t>(a.out) % 16 != 0) return -2;
    for (int k = 0; k < a.ncalls; ++k)
        if (reinterpret_cast<uintptr_t>(a.calls[k].loc) % 8 != 0 || reinterpret_cast<uintptr_t>(a.calls[k].weights) % 16 != 0)
            return -2;
    const bool no_ws = a.workspace == nullptr;
    const GroupPlan pl = plan_group(false, a.calls, a.ncalls, a.bs, a.cams, a.C, no_ws ? 1 : 0);
    if (pl.ps_max > kGroupMaxPS) return -2;             // (only reachable without a workspace: rows cannot be sliced)
    GroupParams gp = {};
    gp.feat = a.feat; gp.shapes = a.shapes; gp.starts = a.starts;
    long long a_total = 0;
    for (int k = 0; k < a.ncalls; ++k) a_total += a.calls[k].A;
    int rc = fill_group_params(gp, pl, a.calls, a.ncalls, a.bs, a.cams, a.num_feat, a.C, a.G, a_total * a.C, a.out, nullptr);
    if (rc != 0) return rc;
    if (pl.parts > 0) {
        if (reinterpret_cast<uintptr_t>(a.workspace) % kAlignG != 0 ||
            a.workspace_bytes < pl.partial_bytes + pl.ticket_bytes)
            return -3;
        gp.partial = reinterpret_cast<float*>(a.workspace);
        gp.tickets = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(a.workspace) + pl.partial_bytes);
        const cudaError_t e = cudaMemsetAsync(gp.tickets, 0, (size_t)pl.rows * sizeof(int), a.stream);
        if (e != cudaSuccess) return (int)e;
    }
    return launch_group_sample(false, a.type, gp, pl.units, a.stream);
}

}  // namespace hipad

#ifdef HIPAD_DFA_TRACE
// development builds: copy (and clear) the per-CTA phase trace of the grouped sample kernels
extern "C" int hipad_dfa_trace_read_group(long long* host, long long count) {
    cudaDeviceSynchronize();
    const int rc = (int)cudaMemcpyFromSymbol(host, hipad::g_trace, sizeof(long long) * (size_t)count);
    void* sym = nullptr;
    cudaGetSymbolAddress(&sym, hipad::g_trace);
    cudaMemset(sym, 0, sizeof(hipad::g_trace));
    return rc;
}
#endif
