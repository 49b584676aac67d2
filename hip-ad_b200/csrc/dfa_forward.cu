// dfa_forward.cu — forward launchers (plain 5-argument contract and fused projection+softmax).
#include "dfa_dispatch.cuh"

namespace hipad {

KernelShape pick_shape(ElemType t, int C, int G, bool aligned16) {
    KernelShape ks{false, 0, false};
    const int V = (t == kF32) ? 4 : 8;
    const int gd = C / G;
    const bool pow2_lanes = (gd % V == 0) && (((gd / V) & ((gd / V) - 1)) == 0) && (gd / V) <= 32;
    // vector path: a single (possibly partial) 32-lane chunk, or 2..4 FULL chunks
    if (aligned16 && C % V == 0 && pow2_lanes && (C <= 32 * V || (C % (32 * V) == 0 && C <= 4 * 32 * V))) {
        ks.vector = true;
        ks.nch = (C + 32 * V - 1) / (32 * V);
        ks.ok = true;
        return ks;
    }
    if (C <= 256) {
        ks.vector = false;
        ks.nch = (C <= 64) ? 2 : 8;
        ks.ok = true;
    }
    return ks;
}

int choose_slices(long long rows, int pairs, int target_ctas, int bytes_per_pair, int max_slices) {
    // Measured on B200 (gpurun r01n-r01p, stage-2 shapes):
    //  * forward (max_slices = 8, the slices of a row are a thread-block cluster): cluster gang scheduling costs
    //    ~25 us per launch, so rows are sliced only when there are fewer rows than CTA slots (2 per SM);
    //  * backward (no cluster, nothing to combine across slices): ~8 CTAs per SM worth of slices while a slice keeps
    //    >= 48 (p,cam) pairs (map: S=16 81.9 us vs S=8 95.1; plan: S=4 85.0 vs S=2 89.1; det stays S=1).
    const bool clustered = max_slices <= 8;
    const int min_pairs = clustered ? 64 : 48;
    int S = 1;
    if (!clustered || rows < 2 * 148)
        while (S < 64 && S * 2 <= max_slices && rows * S < target_ctas && pairs / (S * 2) >= min_pairs) S *= 2;
    const int forced = hipad_env_int("HIPAD_DFA_SLICES", 0);   // A/B knob
    if (forced > 0 && forced <= max_slices && forced <= 64) S = forced;
    // fixed part of the CTA's shared memory (tables, reduction scratch) is < 48 KB for every supported shape
    const long long budget = (long long)kSampleSmemBudget - 48 * 1024;
    while ((long long)((pairs + S - 1) / S) * bytes_per_pair > budget) {
        if (S * 2 > max_slices) return 0;
        S *= 2;
    }
    return S;
}

int launch_forward(const FwdArgs& a) {
    const Dims& d = a.d;
    const bool al = (reinterpret_cast<uintptr_t>(a.feat) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(a.out) % 16 == 0);
    const KernelShape ks = pick_shape(a.type, d.C, d.G, al);
    if (!ks.ok || d.cams * d.L > kMaxCamLevels || (long long)d.num_feat * d.C >= (1LL << 30)) return -2;
    // float2 accesses of the kernels: reject odd sub-buffers instead of faulting
    if (reinterpret_cast<uintptr_t>(a.loc) % 8 != 0 || reinterpret_cast<uintptr_t>(a.loc_out) % 8 != 0) return -1;
    const int mode = a.fused ? kFused : kFwd;
    if (a.fused && (256 % d.G != 0)) return -2;
    if (!a.fused && al && d.P * d.cams <= 96 && group_kernel_supported(a.type, d.C, d.L, d.G, d.cams)) {
        // whole rows fit one work unit: the grouped kernel (dfa_group.cuh) with a single call and no workspace
        GroupFwdArgs g = {};
        g.type = a.type; g.out = a.out; g.feat = a.feat; g.shapes = a.shapes; g.starts = a.starts;
        g.calls[0].loc = a.loc; g.calls[0].weights = a.weights; g.calls[0].A = d.A; g.calls[0].P = d.P;
        g.ncalls = 1;
        g.bs = d.bs; g.cams = d.cams; g.num_feat = d.num_feat; g.C = d.C; g.L = d.L; g.G = d.G;
        g.stream = a.stream;
        const int rc = launch_group_forward(g);
        if (rc != -2) return rc;
    }

    SampleParams p = {};
    p.feat = a.feat; p.shapes = a.shapes; p.starts = a.starts;
    p.loc = a.loc; p.weights = a.weights; p.out = a.out;
    p.key_points = a.key_points; p.proj = a.proj; p.image_wh = a.image_wh; p.loc_out = a.loc_out;
    p.d = d;
    const int NP = d.P * d.cams;
    const long long rows = (long long)d.bs * d.A;
    p.S = choose_slices(rows, NP, 4 * 148, sample_smem_per_pair(mode, d.L), /*max_slices=*/8);   // cluster size <= 8
    if (p.S == 0) return -2;
    p.PS = (NP + p.S - 1) / p.S;
    const long long grid = rows * p.S;
    if (grid > 0x7fffffffLL) return -2;
    const int warps = choose_sample_warps(grid, d.G);
    const size_t smem = sample_smem_for(mode, d, ks, a.type, p.PS, warps);
    if (smem > kSampleSmemBudget) return -2;

#define HIPAD_GO(T_, MODE_)                                                                  \
    return (p.S > 1) ? dispatch_sample<T_, MODE_, true>(p, ks, warps, (int)grid, smem, a.stream)    \
                     : dispatch_sample<T_, MODE_, false>(p, ks, warps, (int)grid, smem, a.stream)
    if (a.type == kF32) {
        if (a.fused) HIPAD_GO(float, kFused); else HIPAD_GO(float, kFwd);
    } else {
        if (a.fused) HIPAD_GO(__nv_bfloat16, kFused); else HIPAD_GO(__nv_bfloat16, kFwd);
    }
#undef HIPAD_GO
}

// ------------------------------------------------------------------ integer sampling contract
__global__ void dfa_indices_kernel(int32_t* __restrict__ idx, const int* __restrict__ shapes,
                                   const int* __restrict__ starts, const float* __restrict__ loc,
                                   long long n_sample, int cams, int L) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sample) return;
    const int cam = (int)(s % cams);
    const float2 xy = __ldg(reinterpret_cast<const float2*>(loc) + s);
    const bool vis = loc_valid(xy.x, xy.y);
    for (int l = 0; l < L; ++l) {
        const int cl = cam * L + l;
        const int h = __ldg(shapes + cl * 2), w = __ldg(shapes + cl * 2 + 1), st = __ldg(starts + cl);
        int32_t* o = idx + (s * L + l) * 6;
        o[0] = vis ? 1 : 0;
        o[3] = st;
        if (vis) {
            const Quad q = quad_setup(xy.x, xy.y, h, w);
            o[1] = q.h_low;
            o[2] = q.w_low;
            o[4] = (int)q.ok1 | ((int)q.ok2 << 1) | ((int)q.ok3 << 2) | ((int)q.ok4 << 3);
            o[5] = st + q.h_low * w + q.w_low;
        } else {
            o[1] = 0; o[2] = 0; o[4] = 0; o[5] = 0;
        }
    }
}

int launch_indices(int32_t* idx, const int* shapes, const int* starts, const float* loc, int bs, int cams, int L,
                   int A, int P, cudaStream_t stream) {
    const long long n = (long long)bs * A * P * cams;
    const long long blocks = (n + 255) / 256;
    if (blocks > 0x7fffffffLL) return -2;
    dfa_indices_kernel<<<(unsigned)blocks, 256, 0, stream>>>(idx, shapes, starts, loc, n, cams, L);
    return (int)cudaGetLastError();
}

}  // namespace hipad
