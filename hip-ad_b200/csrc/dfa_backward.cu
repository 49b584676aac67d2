// dfa_backward.cu — backward launchers: sample-major (g_w, g_loc) + feature-major (g_feat).
#include "dfa_dispatch.cuh"
#include "dfa_gfeat.cuh"
#include <cstdio>
#include <map>
#include <mutex>
#include <utility>

namespace hipad {

namespace {
// HIPAD_DFA_DEBUG_SYNC=1: synchronise after every launch of the backward and name the kernel that failed
inline int debug_sync(const char* what, cudaStream_t st) {
    static const int on = hipad_env_int("HIPAD_DFA_DEBUG_SYNC", 0);
    if (!on) return 0;
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) fprintf(stderr, "hipad_dfa: %s failed: %s\n", what, cudaGetErrorString(e));
    return (int)e;
}
// The compaction -> band sort -> classification chain of a backward call depends only on the sampling locations, not
// on the sample-major kernel, so it runs beside that kernel on a helper stream (fork / join with events; under CUDA-graph
// capture the helper stream joins the capture and the two become parallel branches).  One helper stream + two events per
// (device, caller stream), created on first use outside of any capture and kept for the life of the process.
struct SideStream {
    cudaStream_t stream;
    cudaEvent_t fork, join;
};
SideStream* acquire_side_stream(cudaStream_t main) {
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, SideStream> pool;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    auto it = pool.find(std::make_pair(dev, main));
    if (it != pool.end()) return &it->second;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(main, &st) != cudaSuccess) {
        cudaGetLastError();          // e.g. legacy stream while another stream is capturing: stay serial
        return nullptr;
    }
    if (st != cudaStreamCaptureStatusNone) return nullptr;   // never create resources inside a capture
    SideStream s;
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return &pool.emplace(std::make_pair(dev, main), s).first->second;
}

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct WorkspaceLayout {
    size_t vis_id, vis_xy, vis_cnt, band_cnt, rec, seg, cursor, sortbuf, part_list, tiny_list, partial, unit_done, counters, total;
    size_t partial_slots;
    int seg_stride, n_chunks;
};

WorkspaceLayout workspace_layout(const Dims& d) {
    WorkspaceLayout w;
    const size_t n_cl = (size_t)d.cams * d.L, AP = (size_t)d.A * d.P;
    w.n_chunks = (int)((AP + kVisChunk - 1) / kVisChunk);
    // band tables of a bucket occupy [8*start, 8*(start + h*w)) (dfa_gfeat.cuh, seg_offset)
    w.seg_stride = (int)((size_t)kSegScale * d.num_feat);
    size_t off = 0;
    w.vis_id = off;  off += align_up((size_t)d.bs * d.cams * AP * sizeof(int));
    w.vis_xy = off;  off += align_up((size_t)d.bs * d.cams * AP * sizeof(float2));
    w.vis_cnt = off; off += align_up((size_t)d.bs * d.cams * w.n_chunks * sizeof(int));
    w.band_cnt = off; off += align_up((size_t)d.bs * d.cams * w.n_chunks * d.L * kMaxBands * sizeof(int));
    w.rec = off;     off += align_up((size_t)d.bs * n_cl * AP * sizeof(int4));
    w.seg = off;     off += align_up((size_t)d.bs * w.seg_stride * sizeof(int));
    w.cursor = off;  off += align_up((size_t)d.bs * n_cl * sizeof(int));
    w.sortbuf = off; off += align_up((size_t)d.bs * n_cl * 2 * AP * sizeof(unsigned long long));
    // part sums of rows with more than kPart contributions.  Worst case (every sample visible and piled onto few
    // rows) would need bs*AP*cams*L*8/kPart slots; half of that covers every realistic input, and a row that finds no
    // free slots is summed by a single warp instead (dfa_row_classify_kernel).
    w.partial_slots = (size_t)d.bs * (AP * n_cl / 16 + 1024);
    w.part_list = off;   off += align_up(((size_t)d.bs * d.num_feat + w.partial_slots) * kEntryInts * sizeof(int));
    w.tiny_list = off;   off += align_up((size_t)d.bs * d.num_feat * kEntryInts * sizeof(int));
    w.partial = off;     off += align_up(w.partial_slots * d.C * sizeof(float));
    w.unit_done = off;   off += align_up(w.partial_slots * sizeof(int));
    w.counters = off;    off += align_up(8 * sizeof(int));
    w.total = off;
    return w;
}

template <typename T, int V, int NCH>
int launch_reduce_nq(const GfeatParams& gp, cudaStream_t st) {
#define HIPAD_NQ(NQ_)                                                                          \
    case NQ_:                                                                                  \
        dfa_gfeat_reduce_kernel<T, V, NCH, NQ_><<<kReduceCtas, 256, 0, st>>>(gp);              \
        break
    switch (gp.tiny_ok ? gp.d.C / 32 : 0) {
        HIPAD_NQ(0); HIPAD_NQ(1); HIPAD_NQ(2); HIPAD_NQ(4); HIPAD_NQ(8);
        default: return -2;
    }
#undef HIPAD_NQ
    return (int)cudaGetLastError();
}

template <typename T>
int launch_reduce(const GfeatParams& gp, KernelShape ks, cudaStream_t st) {
    constexpr int VV = 16 / (int)sizeof(T);
    if (ks.vector && sizeof(T) == 2) {
        // grad_out and the partial sums are fp32 whatever the feature type: use the fp32 lane map (4 channels per lane,
        // 512 contiguous bytes per warp load) and store 4 bf16 per lane; the 8-channel map reads half of every sector
        const int C = gp.d.C;
        if (C <= 128) return launch_reduce_nq<T, 4, 1>(gp, st);
        if (C == 256) return launch_reduce_nq<T, 4, 2>(gp, st);
        if (C == 384) return launch_reduce_nq<T, 4, 3>(gp, st);
        if (C == 512) return launch_reduce_nq<T, 4, 4>(gp, st);
    }
    if (ks.vector) {
        if (ks.nch == 1) return launch_reduce_nq<T, VV, 1>(gp, st);
        if (ks.nch == 2) return launch_reduce_nq<T, VV, 2>(gp, st);
        if (ks.nch == 3) return launch_reduce_nq<T, VV, 3>(gp, st);
        if (ks.nch == 4) return launch_reduce_nq<T, VV, 4>(gp, st);
    } else {
        if (ks.nch == 2) return launch_reduce_nq<T, 1, 2>(gp, st);
        if (ks.nch == 8) return launch_reduce_nq<T, 1, 8>(gp, st);
    }
    return -2;
}
}  // namespace

size_t backward_workspace_bytes(const Dims& d) { return workspace_layout(d).total; }
size_t backward_counters_offset(const Dims& d) { return workspace_layout(d).counters; }

int launch_backward(const BwdArgs& a) {
    const Dims& d = a.d;
    const bool al = (reinterpret_cast<uintptr_t>(a.feat) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(a.g_feat) % 16 == 0) &&   /* nullptr passes */
                    (reinterpret_cast<uintptr_t>(a.grad_out) % 16 == 0) &&
                    (reinterpret_cast<uintptr_t>(a.g_w) % 16 == 0);
    const KernelShape ks = pick_shape(a.type, d.C, d.G, al);
    if (!ks.ok || d.cams * d.L > kMaxCamLevels || (long long)d.num_feat * d.C >= (1LL << 30)) return -2;
    const WorkspaceLayout wl = workspace_layout(d);
    if (a.g_feat != nullptr &&
        (a.workspace == nullptr || a.workspace_bytes < wl.total ||
         reinterpret_cast<uintptr_t>(a.workspace) % kAlign != 0))
        return -3;
    if ((long long)d.A * d.P > (long long)kMaxChunks * kVisChunk) return -2;
    // 32-bit BYTE offsets inside one batch element's weights / grad_out (dfa_gfeat.cuh, accumulate_row)
    if ((long long)d.A * d.P * d.cams * d.L * d.G >= (1LL << 30) || (long long)d.A * d.C >= (1LL << 30)) return -2;

    // ---- K1: sample-major, g_w + g_loc (fully written); zero-fills g_feat on the side when it has the CTAs
    SampleParams p = {};
    p.feat = a.feat; p.shapes = a.shapes; p.starts = a.starts;
    p.loc = a.loc; p.weights = a.weights;
    p.grad_out = a.grad_out; p.g_loc = a.g_loc; p.g_w = a.g_w;
    p.d = d;
    const int NP = d.P * d.cams;
    const long long rows = (long long)d.bs * d.A;
    p.S = choose_slices(rows, NP, 8 * 148, sample_smem_per_pair(kBwd, d.L), /*max_slices=*/1 << 20);
    if (p.S == 0 || d.bs > 65535 || (long long)d.bs * d.num_feat >= (1LL << 31)) return -2;
    p.PS = (NP + p.S - 1) / p.S;
    const long long grid = rows * p.S;
    if (grid > 0x7fffffffLL) return -2;
    const int warps = choose_sample_warps(grid, d.G);
    const size_t smem = sample_smem_for(kBwd, d, ks, a.type, p.PS, warps);
    if (smem > kSampleSmemBudget) return -2;
    const size_t gfeat_bytes = (size_t)d.bs * d.num_feat * d.C * (a.type == kF32 ? 4 : 2);
    // full backward with a feature gradient: the sort chain goes to the helper stream, forked BEFORE the sample kernel
    cudaStream_t chain = a.stream;
    SideStream* side = nullptr;
    if (a.g_feat != nullptr && a.stage_mask == 7 && !a.classify_only && hipad_env_int("HIPAD_DFA_OVERLAP", 1) != 0 &&
        hipad_env_int("HIPAD_DFA_DEBUG_SYNC", 0) == 0)
        side = acquire_side_stream(a.stream);
    if (side != nullptr) {
        if (cudaEventRecord(side->fork, a.stream) != cudaSuccess || cudaStreamWaitEvent(side->stream, side->fork, 0) != cudaSuccess) {
            cudaGetLastError();
            side = nullptr;
        } else {
            chain = side->stream;
        }
    }
    if (a.stage_mask & 1) {
        if (a.g_feat != nullptr && !a.accumulate) {
            const bool vec_ok = (gfeat_bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(a.g_feat) % 16 == 0);
            if (vec_ok && grid >= 2 * 148 && !a.separate_zero_fill) {
                p.zero_ptr = reinterpret_cast<uint4*>(a.g_feat);
                p.zero_n16 = (long long)(gfeat_bytes / 16);
            } else if (vec_ok) {
                dfa_zero_kernel<<<148 * 16, 256, 0, a.stream>>>(reinterpret_cast<uint4*>(a.g_feat),
                                                              (long long)(gfeat_bytes / 16));
                const cudaError_t e = cudaGetLastError();
                if (e != cudaSuccess) return (int)e;
            } else {
                const cudaError_t e = cudaMemsetAsync(a.g_feat, 0, gfeat_bytes, a.stream);
                if (e != cudaSuccess) return (int)e;
            }
        }
        const int rc = (a.type == kF32)
                           ? dispatch_sample<float, kBwd, false>(p, ks, warps, (int)grid, smem, a.stream)
                           : dispatch_sample<__nv_bfloat16, kBwd, false>(p, ks, warps, (int)grid, smem, a.stream);
        if (rc != 0) return rc;
        if (int e = debug_sync("sample-major backward kernel", a.stream)) return e;
    }
    if (a.g_feat == nullptr) return 0;   // caller does not need the feature-map gradient

    // ---- K2: visible-sample compaction, then per-(b,cam,level,band) sort by quad key
    unsigned char* ws = reinterpret_cast<unsigned char*>(a.workspace);
    GfeatParams gp = {};
    gp.shapes = a.shapes; gp.starts = a.starts; gp.loc = a.loc; gp.weights = a.weights;
    gp.grad_out = a.grad_out; gp.g_feat = a.g_feat;
    gp.vis_id = reinterpret_cast<int*>(ws + wl.vis_id);
    gp.vis_xy = reinterpret_cast<float2*>(ws + wl.vis_xy);
    gp.vis_cnt = reinterpret_cast<int*>(ws + wl.vis_cnt);
    gp.band_cnt = reinterpret_cast<int*>(ws + wl.band_cnt);
    gp.rec = reinterpret_cast<int4*>(ws + wl.rec);
    gp.seg = reinterpret_cast<int*>(ws + wl.seg);
    gp.cursor = reinterpret_cast<int*>(ws + wl.cursor);
    gp.sortbuf = reinterpret_cast<unsigned long long*>(ws + wl.sortbuf);
    gp.part_list = reinterpret_cast<int4*>(ws + wl.part_list);
    gp.partial = reinterpret_cast<float*>(ws + wl.partial);
    gp.unit_done = reinterpret_cast<int*>(ws + wl.unit_done);
    gp.partial_cap = (int)wl.partial_slots;
    {   // test knob: shrink the partial-slot pool to exercise the single-warp fallback of rows that find no slots
        const int cap = hipad_env_int("HIPAD_DFA_PARTIAL_CAP", -1);
        if (cap >= 0 && cap < gp.partial_cap) gp.partial_cap = cap;
    }
    gp.tiny_list = reinterpret_cast<int4*>(ws + wl.tiny_list);
    gp.counters = reinterpret_cast<int*>(ws + wl.counters);
    gp.d = d;
    gp.seg_stride = wl.seg_stride;
    gp.n_chunks = wl.n_chunks;
    // enough bands for ~2 sort CTAs per SM, whatever the batch size; twice that for long sample lists, where a band
    // CTA's cost is the pass over its camera's visible samples (measured, stage-2 map / plan: 24 bands 36 / 40 us
    // against 42 / 46 us with 12-16; the det call, A*P = 11 700, is 2 us faster with 12)
    const int buckets = d.cams * d.L * d.bs;
    int nb = (2 * 148) / buckets;
    if ((long long)d.A * d.P >= 24000) nb *= 2;
    gp.NB = nb < 1 ? 1 : (nb > kMaxBands ? kMaxBands : nb);
    {   // A/B knobs
        const int fb = hipad_env_int("HIPAD_DFA_BANDS", 0);
        if (fb >= 1 && fb <= kMaxBands) gp.NB = fb;
        const int tm = hipad_env_int("HIPAD_DFA_TINY_MAX", kTinyRow);
        gp.tiny_max = tm < 1 ? 1 : (tm > kTinyRow ? kTinyRow : tm);
    }
    gp.accumulate = a.accumulate ? 1 : 0;
    gp.tiny_ok = (ks.vector && (d.C == 32 || d.C == 64 || d.C == 128 || d.C == 256) && (d.C / d.G) % 32 == 0 &&
                  hipad_env_int("HIPAD_DFA_TINY", 1) != 0) ? 1 : 0;
    if (a.stage_mask & 2) {
        dfa_vis_compact_kernel<<<dim3((unsigned)wl.n_chunks, (unsigned)d.cams, (unsigned)d.bs), kVisThreads, 0,
                                 chain>>>(gp);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        const size_t sort_smem = band_sort_smem_bytes(wl.n_chunks);
        e = ensure_smem(dfa_band_sort_kernel, sort_smem);
        if (e != cudaSuccess) return (int)e;
        dfa_band_sort_kernel<<<dim3((unsigned)gp.NB, (unsigned)(d.cams * d.L), (unsigned)d.bs), kSortThreads, sort_smem,
                               chain>>>(gp);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        if (int e2 = debug_sync("compaction / band sort", chain)) return e2;
    }
    if (!(a.stage_mask & 4)) return 0;

    // ---- K3: feature-major reduce, overwrites every touched row of the zero-filled g_feat once
    dfa_row_classify_kernel<<<dim3((unsigned)((d.num_feat + kClassifyThreads - 1) / kClassifyThreads), (unsigned)d.bs),
                              kClassifyThreads, 0, chain>>>(gp);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (int e2 = debug_sync("row classification", chain)) return e2;
    if (a.classify_only) return 0;
    if (side != nullptr) {   // join: the reduce needs the work lists (helper stream) and the zero-filled g_feat (caller's)
        cudaError_t ej = cudaEventRecord(side->join, side->stream);
        if (ej == cudaSuccess) ej = cudaStreamWaitEvent(a.stream, side->join, 0);
        if (ej != cudaSuccess) return (int)ej;
    }
    const int rc = (a.type == kF32) ? launch_reduce<float>(gp, ks, a.stream) : launch_reduce<__nv_bfloat16>(gp, ks, a.stream);
    if (rc != 0) return rc;
    return debug_sync("feature-gradient reduce", a.stream);
}

}  // namespace hipad
