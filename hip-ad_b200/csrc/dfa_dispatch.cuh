// dfa_dispatch.cuh — host-side template dispatch for the sample-major kernel family.
#pragma once
#include "dfa_launch.h"
#include "dfa_sample.cuh"
#include <cstdlib>

namespace hipad {

inline int hipad_env_int(const char* name, int dflt) {   // tuning knobs for A/B measurements only
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

constexpr size_t kSampleSmemBudget = 220 * 1024;   // dynamic shared memory one sample-kernel CTA may ask for

// shared-memory bytes of gather metadata per (p,cam) pair of a slice (sample_smem_layout, dfa_sample.cuh)
inline int sample_smem_per_pair(int mode, int L) {
    return 8 + 4 + L * (16 + 16 + 4) + (mode == kBwd ? 1 + L * 32 : 0);
}

template <typename K>
inline cudaError_t ensure_smem(K kern, size_t smem) {
    if (smem <= 48 * 1024) return cudaSuccess;
    // attribute is per device and per function; setting it is idempotent and cheap
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

template <typename T, int V, int NCH, int kL, int kMode, int kLPG, bool kCluster, int kWarps>
int launch_sample_inst(const SampleParams& p, int grid, size_t smem, cudaStream_t st) {
    auto kern = dfa_sample_kernel<T, V, NCH, kL, kMode, kLPG, kCluster, kWarps>;
    cudaError_t e = ensure_smem(kern, smem);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(kWarps * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (kCluster) {
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)p.S;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    e = cudaLaunchKernelEx(&cfg, kern, p);
    return (int)e;
}

template <typename T, int kMode, bool kCluster, int kWarps>
int dispatch_sample_w(const SampleParams& p, KernelShape ks, int grid, size_t smem, cudaStream_t st) {
    constexpr int VV = 16 / (int)sizeof(T);
    const bool l4 = (p.d.L == 4);
    const int lpg = (p.d.C / p.d.G) / VV;
#define HIPAD_CASE(V_, NCH_, KL_, LPG_) \
    return launch_sample_inst<T, V_, NCH_, KL_, kMode, LPG_, kCluster, kWarps>(p, grid, smem, st)
    if (ks.vector) {
        // the shipped HiP-AD shape (C=256, G=8, L=4) gets fully compile-time reductions
        if (kMode == kBwd && l4 && ks.nch == 32 / VV * 8 / 32 && lpg == 32 / VV) HIPAD_CASE(VV, 32 / VV * 8 / 32, 4, 32 / VV);
        if (ks.nch == 1) { if (l4) HIPAD_CASE(VV, 1, 4, 0); else HIPAD_CASE(VV, 1, 0, 0); }
        if (ks.nch == 2) { if (l4) HIPAD_CASE(VV, 2, 4, 0); else HIPAD_CASE(VV, 2, 0, 0); }
        if (ks.nch == 3) { if (l4) HIPAD_CASE(VV, 3, 4, 0); else HIPAD_CASE(VV, 3, 0, 0); }
        if (ks.nch == 4) { if (l4) HIPAD_CASE(VV, 4, 4, 0); else HIPAD_CASE(VV, 4, 0, 0); }
    } else {
        if (ks.nch == 2) HIPAD_CASE(1, 2, 0, -1);
        if (ks.nch == 8) HIPAD_CASE(1, 8, 0, -1);
    }
#undef HIPAD_CASE
    return -2;
}

// CTA width: 2 warps (many small CTAs: ~8 resident per SM, so one CTA's visibility/metadata phases overlap the
// gather loops of the others and a bs=1 call fits one wave) or 8 warps (few output rows)
template <typename T, int kMode, bool kCluster>
int dispatch_sample(const SampleParams& p, KernelShape ks, int warps, int grid, size_t smem, cudaStream_t st) {
    if (warps == 2) return dispatch_sample_w<T, kMode, kCluster, 2>(p, ks, grid, smem, st);
    return dispatch_sample_w<T, kMode, kCluster, 8>(p, ks, grid, smem, st);
}

inline size_t sample_smem_for(int mode, const Dims& d, KernelShape ks, ElemType t, int ps, int warps) {
    const int V = ks.vector ? (t == kF32 ? 4 : 8) : 1;
    if (warps == 2)
        return (size_t)sample_smem_layout<2>(mode, d.cams * d.L, d.L, ks.nch * 32 * V, ps, d.G, d.cams).total;
    return (size_t)sample_smem_layout<8>(mode, d.cams * d.L, d.L, ks.nch * 32 * V, ps, d.G, d.cams).total;
}

// warps per CTA for a launch of `ctas` CTAs
inline int choose_sample_warps(long long ctas, int G) {
    const int forced = hipad_env_int("HIPAD_DFA_SAMPLE_WARPS", 0);
    if (forced == 2 || forced == 8) return (64 % G == 0 || forced == 8) ? forced : 8;
    (void)ctas;
    return 8;   // measured on B200: narrow CTAs lose the intra-CTA L1 reuse between neighbouring key points (r01h)
}

}  // namespace hipad
