// dfa_backward.cu — backward launchers: sample-major (g_w, g_loc) + feature-major (g_feat).
#include "dfa_dispatch.cuh"
#include "dfa_gfeat.cuh"

namespace hipad {

namespace {
constexpr size_t kAlign = 256;
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct WorkspaceLayout {
    size_t rec, seg, counts, sortbuf, heavy_list, heavy_count, total;
    int seg_stride;
};

WorkspaceLayout workspace_layout(const Dims& d) {
    WorkspaceLayout w;
    const size_t n_cl = (size_t)d.cams * d.L, AP = (size_t)d.A * d.P;
    // sum over (cam,level) of (H+1)(W+1)+1 <= 2*num_feat + 3*cams*L for any shapes with H,W >= 1
    w.seg_stride = (int)(2 * (size_t)d.num_feat + 3 * n_cl);
    size_t off = 0;
    w.rec = off;     off += align_up((size_t)d.bs * n_cl * AP * sizeof(int));
    w.seg = off;     off += align_up((size_t)d.bs * w.seg_stride * sizeof(int));
    w.counts = off;  off += align_up((size_t)d.bs * n_cl * sizeof(int));
    w.sortbuf = off; off += align_up((size_t)d.bs * n_cl * 2 * AP * sizeof(unsigned long long));
    w.heavy_list = off;  off += align_up((size_t)d.bs * d.num_feat * sizeof(int2));
    w.heavy_count = off; off += align_up(sizeof(int));
    w.total = off;
    return w;
}

template <typename T>
int launch_reduce(const GfeatParams& gp, KernelShape ks, dim3 grid, cudaStream_t st) {
    constexpr int VV = 16 / (int)sizeof(T);
#define HIPAD_RED(V_, NCH_)                                                                \
    do {                                                                                   \
        dfa_gfeat_reduce_kernel<T, V_, NCH_><<<grid, kReduceWarps * 32, 0, st>>>(gp);      \
        cudaError_t e_ = cudaGetLastError();                                               \
        if (e_ != cudaSuccess) return (int)e_;                                             \
        dfa_gfeat_heavy_kernel<T, V_, NCH_><<<kHeavyCtas, kReduceWarps * 32, 0, st>>>(gp); \
        return (int)cudaGetLastError();                                                    \
    } while (0)
    if (ks.vector) {
        if (ks.nch == 1) HIPAD_RED(VV, 1);
        if (ks.nch == 2) HIPAD_RED(VV, 2);
        if (ks.nch == 3) HIPAD_RED(VV, 3);
        if (ks.nch == 4) HIPAD_RED(VV, 4);
    } else {
        if (ks.nch == 2) HIPAD_RED(1, 2);
        if (ks.nch == 8) HIPAD_RED(1, 8);
    }
#undef HIPAD_RED
    return -2;
}
}  // namespace

size_t backward_workspace_bytes(const Dims& d) { return workspace_layout(d).total; }

int launch_backward(const BwdArgs& a) {
    const Dims& d = a.d;
    const bool al = (reinterpret_cast<uintptr_t>(a.feat) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(a.g_feat) % 16 == 0) &&   /* nullptr passes */
                    (reinterpret_cast<uintptr_t>(a.grad_out) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(a.g_w) % 16 == 0);
    const KernelShape ks = pick_shape(a.type, d.C, d.G, al);
    if (!ks.ok || d.cams * d.L > kMaxCamLevels || (long long)d.num_feat * d.C >= (1LL << 30)) return -2;
    const WorkspaceLayout wl = workspace_layout(d);
    if (a.workspace == nullptr || a.workspace_bytes < wl.total ||
        reinterpret_cast<uintptr_t>(a.workspace) % kAlign != 0)
        return -3;
    if ((long long)d.A * d.P > (1LL << 30)) return -2;
    if ((long long)d.A * d.P * d.cams * d.L * d.G >= (1LL << 31) || (long long)d.A * d.C >= (1LL << 31)) return -2;

    // ---- K1: sample-major, g_w + g_loc (fully written)
    SampleParams p = {};
    p.feat = a.feat; p.shapes = a.shapes; p.starts = a.starts;
    p.loc = a.loc; p.weights = a.weights;
    p.grad_out = a.grad_out; p.g_loc = a.g_loc; p.g_w = a.g_w;
    p.d = d;
    const int NP = d.P * d.cams;
    const long long rows = (long long)d.bs * d.A;
    p.S = choose_slices(rows, NP, 4 * 148);
    p.PS = (NP + p.S - 1) / p.S;
    if (p.PS > kMaxPairsPerSlice) return -2;
    const long long grid = rows * p.S;
    if (grid > 0x7fffffffLL) return -2;
    const size_t smem = sample_smem_for(kBwd, d, ks, a.type, p.PS);
    if (a.stage_mask & 1) {
        const int rc = (a.type == kF32)
                           ? dispatch_sample<float, kBwd, false>(p, ks, (int)grid, smem, a.stream)
                           : dispatch_sample<__nv_bfloat16, kBwd, false>(p, ks, (int)grid, smem, a.stream);
        if (rc != 0) return rc;
    }
    if (a.g_feat == nullptr) return 0;   // caller does not need the feature-map gradient

    // ---- K2a: per-(b,cam,level) bucket sort of the visible samples by quad key
    unsigned char* ws = reinterpret_cast<unsigned char*>(a.workspace);
    GfeatParams gp = {};
    gp.shapes = a.shapes; gp.starts = a.starts; gp.loc = a.loc; gp.weights = a.weights;
    gp.grad_out = a.grad_out; gp.g_feat = a.g_feat;
    gp.rec = reinterpret_cast<int*>(ws + wl.rec);
    gp.seg = reinterpret_cast<int*>(ws + wl.seg);
    gp.counts = reinterpret_cast<int*>(ws + wl.counts);
    gp.sortbuf = reinterpret_cast<unsigned long long*>(ws + wl.sortbuf);
    gp.heavy_list = reinterpret_cast<int2*>(ws + wl.heavy_list);
    gp.heavy_count = reinterpret_cast<int*>(ws + wl.heavy_count);
    gp.d = d;
    gp.seg_stride = wl.seg_stride;
    const long long AP = (long long)d.A * d.P;
    int cap = 24576;
    if (AP < cap) cap = (int)((AP + 63) / 64 * 64);
    gp.smem_cap = cap;
    if (a.stage_mask & 2) {
        const size_t sort_smem = bucket_sort_smem_bytes(cap);
        cudaError_t e = ensure_smem(dfa_bucket_sort_kernel, sort_smem);
        if (e != cudaSuccess) return (int)e;
        dfa_bucket_sort_kernel<<<dim3((unsigned)(d.cams * d.L), (unsigned)d.bs), kSortThreads, sort_smem, a.stream>>>(gp);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    if (!(a.stage_mask & 4)) return 0;

    // ---- K2b: feature-major reduce, writes every row of g_feat once
    const unsigned tiles = (unsigned)((d.num_feat + kRowsPerTile - 1) / kRowsPerTile + d.cams * d.L);
    const dim3 rgrid(tiles, (unsigned)d.bs);
    return (a.type == kF32) ? launch_reduce<float>(gp, ks, rgrid, a.stream)
                            : launch_reduce<__nv_bfloat16>(gp, ks, rgrid, a.stream);
}

}  // namespace hipad
