// dfa_backward.cu — backward launchers: sample-major (g_w, g_loc) + feature-major (g_feat).
#include "dfa_dispatch.cuh"
#include "dfa_gfeat.cuh"
#include "dfa_group.cuh"
#include "dfa_group_host.h"
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
    // cudaStreamPerThread names a different stream in every host thread: one pool entry (one event pair) would be
    // shared by all of them, so such callers stay serial.  A given caller stream must be driven by one thread at a time.
    if (main == cudaStreamPerThread) return nullptr;
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
    // The chain is a few hundred latency-bound CTAs; the sample kernel beside it has thousands.  At the highest stream
    // priority the chain's CTAs take the SM slots that free up first instead of queueing behind the sample kernel's
    // (measured, stage-2 layer backward, bs = 1 / 4: 264 -> 235 us, 875 -> 848 us; HIPAD_DFA_SIDE_PRIORITY=0 turns it off).
    int prio_lo = 0, prio_hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) { cudaGetLastError(); prio_hi = 0; }
    if (hipad_env_int("HIPAD_DFA_SIDE_PRIORITY", 1) == 0) prio_hi = 0;
    if (cudaStreamCreateWithPriority(&s.stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return &pool.emplace(std::make_pair(dev, main), s).first->second;
}

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

struct WorkspaceLayout {
    size_t vis_key, chunk_cum, ghist, ghist_bytes, rec, recw, seg, cursor, sortbuf, part_list, tiny_list, partial, unit_done, counters, total;
    size_t partial_slots;
    int seg_stride, n_chunks;
};

// sample id space of a group: call k owns [id_begin[k], id_begin[k] + A*P), id_begin a multiple of kVisChunk
struct IdSpace {
    int id_begin[kMaxCalls], anchor_begin[kMaxCalls];
    long long n_ids, a_total;
};
IdSpace id_space(const CallDesc* calls, int ncalls) {
    IdSpace s = {};
    for (int k = 0; k < ncalls; ++k) {
        s.id_begin[k] = (int)s.n_ids;
        s.anchor_begin[k] = (int)s.a_total;
        const long long ap = (long long)calls[k].A * calls[k].P;
        s.n_ids += (ap + kVisChunk - 1) / kVisChunk * kVisChunk;
        s.a_total += calls[k].A;
    }
    return s;
}

WorkspaceLayout workspace_layout(int bs, int cams, int num_feat, int C, int L, int G, long long n_ids) {
    WorkspaceLayout w;
    const size_t n_cl = (size_t)cams * L, AP = (size_t)n_ids;
    w.n_chunks = (int)(AP / kVisChunk);
    // the segment table of a bucket occupies [kSegScale*start, kSegScale*(start + h*w)) (dfa_gfeat.cuh, seg_offset)
    w.seg_stride = (int)((size_t)kSegScale * num_feat);
    size_t off = 0;
    w.vis_key = off; off += align_up((size_t)bs * n_cl * AP * sizeof(unsigned long long));
    w.chunk_cum = off; off += align_up((size_t)bs * cams * w.n_chunks * L * kCum * sizeof(int));
    w.ghist_bytes = (size_t)bs * n_cl * kFine * sizeof(int);
    w.ghist = off;   off += align_up(w.ghist_bytes);
    w.rec = off;     off += align_up((size_t)bs * n_cl * AP * sizeof(int4));
    w.recw = off;    off += align_up((size_t)bs * n_cl * AP * G * sizeof(float));
    w.seg = off;     off += align_up((size_t)bs * w.seg_stride * sizeof(int));
    w.cursor = off;  off += align_up((size_t)bs * n_cl * sizeof(int));
    w.sortbuf = off; off += align_up((size_t)bs * n_cl * 2 * AP * sizeof(unsigned long long));
    // part sums of rows with more than kPart contributions.  Worst case (every sample visible and piled onto few
    // rows) would need bs*AP*cams*L*8/kPart slots; half of that covers every realistic input, and a row that finds no
    // free slots is summed by a single warp instead (dfa_row_classify_kernel).
    w.partial_slots = (size_t)bs * (AP * n_cl / 16 + 1024);
    w.part_list = off;   off += align_up(((size_t)bs * num_feat + w.partial_slots) * kEntryInts * sizeof(int));
    w.tiny_list = off;   off += align_up((size_t)bs * num_feat * kEntryInts * sizeof(int));
    w.partial = off;     off += align_up(w.partial_slots * C * sizeof(float));
    w.unit_done = off;   off += align_up(w.partial_slots * sizeof(int));
    w.counters = off;    off += align_up(8 * sizeof(int));
    w.total = off;
    return w;
}

template <typename T, int V, int NCH>
int launch_reduce_nq(const GfeatParams& gp, cudaStream_t st) {
    // light variant: 4 contributions in flight per warp, <= 64 registers, 4 CTAs per SM (HIPAD_DFA_REDUCE_LIGHT)
    const bool light = hipad_env_int("HIPAD_DFA_REDUCE_LIGHT", 0) != 0;
#define HIPAD_NQ(NQ_)                                                                                      \
    case NQ_:                                                                                              \
        if (light) dfa_gfeat_reduce_kernel<T, V, NCH, NQ_, 3><<<148 * 3, 256, 0, st>>>(gp);              \
        else dfa_gfeat_reduce_kernel<T, V, NCH, NQ_, 2><<<kReduceCtas, 256, 0, st>>>(gp);                  \
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

size_t group_backward_workspace_bytes(const CallDesc* calls, int ncalls, int bs, int cams, int num_feat, int C, int L, int G) {
    if (ncalls < 1 || ncalls > kMaxCalls) return 0;
    return workspace_layout(bs, cams, num_feat, C, L, G, id_space(calls, ncalls).n_ids).total;
}
size_t group_backward_counters_offset(const CallDesc* calls, int ncalls, int bs, int cams, int num_feat, int C, int L, int G) {
    if (ncalls < 1 || ncalls > kMaxCalls) return 0;
    return workspace_layout(bs, cams, num_feat, C, L, G, id_space(calls, ncalls).n_ids).counters;
}
size_t backward_workspace_bytes(const Dims& d) {
    CallDesc c = {}; c.A = d.A; c.P = d.P;
    return group_backward_workspace_bytes(&c, 1, d.bs, d.cams, d.num_feat, d.C, d.L, d.G);
}
size_t backward_counters_offset(const Dims& d) {
    CallDesc c = {}; c.A = d.A; c.P = d.P;
    return group_backward_counters_offset(&c, 1, d.bs, d.cams, d.num_feat, d.C, d.L, d.G);
}

int launch_backward(const BwdArgs& b) {
    GroupBwdArgs a = {};
    a.type = b.type; a.feat = b.feat; a.shapes = b.shapes; a.starts = b.starts;
    a.calls[0].loc = b.loc; a.calls[0].weights = b.weights; a.calls[0].g_loc = b.g_loc; a.calls[0].g_w = b.g_w;
    a.calls[0].A = b.d.A; a.calls[0].P = b.d.P;
    a.ncalls = 1;
    a.grad_out = b.grad_out; a.g_feat = b.g_feat; a.g_feat_f32 = (b.type == kF32); a.accumulate = b.accumulate;
    a.bs = b.d.bs; a.cams = b.d.cams; a.num_feat = b.d.num_feat; a.C = b.d.C; a.L = b.d.L; a.G = b.d.G;
    a.workspace = b.workspace; a.workspace_bytes = b.workspace_bytes; a.stream = b.stream;
    a.stage_mask = b.stage_mask; a.classify_only = b.classify_only; a.separate_zero_fill = b.separate_zero_fill;
    return launch_group_backward(a);
}

int launch_group_backward(const GroupBwdArgs& a) {
    if (a.ncalls < 1 || a.ncalls > kMaxCalls) return -1;
    Dims d = {};
    d.bs = a.bs; d.cams = a.cams; d.num_feat = a.num_feat; d.C = a.C; d.L = a.L; d.G = a.G;
    d.A = a.calls[0].A; d.P = a.calls[0].P;       // per-call kernels only (single call / bs == 1 fallback)
    bool al = (reinterpret_cast<uintptr_t>(a.feat) % 16 == 0) &&
              (reinterpret_cast<uintptr_t>(a.g_feat) % 16 == 0) &&   /* nullptr passes */
              (reinterpret_cast<uintptr_t>(a.grad_out) % 16 == 0);
    for (int k = 0; k < a.ncalls; ++k) {
        const CallDesc& c = a.calls[k];
        if (!c.loc || !c.weights || !c.g_loc || !c.g_w || c.A <= 0 || c.P <= 0) return -1;
        // float2 / float4 accesses of the kernels (reject instead of faulting on odd sub-buffers)
        if (reinterpret_cast<uintptr_t>(c.loc) % 8 != 0 || reinterpret_cast<uintptr_t>(c.g_loc) % 8 != 0) return -1;
        // the kernels zero / write weight-gradient rows with 16-byte stores when a row (L*G floats) allows it
        if ((d.L * d.G) % 4 == 0 && reinterpret_cast<uintptr_t>(c.g_w) % 16 != 0) return -1;
        al = al && (reinterpret_cast<uintptr_t>(c.g_w) % 16 == 0) && (reinterpret_cast<uintptr_t>(c.weights) % 16 == 0);
        if ((long long)c.A * c.P > (long long)kMaxChunks * kVisChunk) return -2;
        if ((long long)c.A * c.P * d.cams * d.L * d.G >= (1LL << 30)) return -2;
    }
    const KernelShape ks = pick_shape(a.type, d.C, d.G, al);
    const KernelShape ks_g = a.g_feat_f32 ? pick_shape(kF32, d.C, d.G, al) : ks;     // lane map of the reduce
    if (!ks.ok || !ks_g.ok || d.cams * d.L > kMaxCamLevels || (long long)d.num_feat * d.C >= (1LL << 30)) return -2;
    const IdSpace ids = id_space(a.calls, a.ncalls);
    if (ids.n_ids > (long long)kMaxChunks * kVisChunk || ids.a_total * d.C >= (1LL << 30)) return -2;
    const WorkspaceLayout wl = workspace_layout(d.bs, d.cams, d.num_feat, d.C, d.L, d.G, ids.n_ids);
    if (a.g_feat != nullptr &&
        (a.workspace == nullptr || a.workspace_bytes < wl.total ||
         reinterpret_cast<uintptr_t>(a.workspace) % kAlign != 0))
        return -3;
    if (d.bs > 65535 || (long long)d.bs * d.num_feat >= (1LL << 31)) return -2;
    const bool grouped = al && group_kernel_supported(a.type, d.C, d.L, d.G, d.cams);
    if (!grouped && a.ncalls > 1 && d.bs > 1) return -2;     // per-call kernels need contiguous [bs, A, C] gradients

    const size_t gfeat_elem = (a.g_feat_f32 || a.type == kF32) ? 4 : 2;
    const size_t gfeat_bytes = (size_t)d.bs * d.num_feat * d.C * gfeat_elem;
    // full backward with a feature gradient: the sort chain goes to the helper stream, forked BEFORE the sample kernel
    // (every fallible check is above this point, so a fork is always followed by its join)
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
    // joins the helper stream back into the caller's on every exit path after a successful fork
    auto finish = [&](int rc) {
        if (side != nullptr) {
            cudaError_t ej = cudaEventRecord(side->join, side->stream);
            if (ej == cudaSuccess) ej = cudaStreamWaitEvent(a.stream, side->join, 0);
            side = nullptr;
            if (rc == 0 && ej != cudaSuccess) rc = (int)ej;
        }
        return rc;
    };

    // ---- K1: sample-major, g_w + g_loc (fully written); zero-fills g_feat on the side when it has the CTAs
    auto run_k1 = [&]() -> int {
        if (a.stage_mask & 1) {
            const bool want_zero = a.g_feat != nullptr && !a.accumulate;
            const bool vec_ok = (gfeat_bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(a.g_feat) % 16 == 0);
            bool zero_done = !want_zero;
            auto separate_zero = [&]() -> int {
                if (vec_ok) {
                    dfa_zero_kernel<<<148 * 16, 256, 0, a.stream>>>(reinterpret_cast<uint4*>(a.g_feat), (long long)(gfeat_bytes / 16));
                    return (int)cudaGetLastError();
                }
                return (int)cudaMemsetAsync(a.g_feat, 0, gfeat_bytes, a.stream);
            };
            if (grouped) {
                const GroupPlan pl = plan_group(true, a.calls, a.ncalls, d.bs, d.cams, d.C, 0);
                GroupParams gpk = {};
                gpk.feat = a.feat; gpk.shapes = a.shapes; gpk.starts = a.starts;
                int rc = fill_group_params(gpk, pl, a.calls, a.ncalls, d.bs, d.cams, d.num_feat, d.C, d.G,
                                           ids.a_total * d.C, nullptr, a.grad_out);
                if (rc != 0) return rc;
                if (want_zero) {
                    // (measured and not kept: giving a launch too small to fold the fill in -- the ego call -- the helper
                    // stream for its sample kernel as well: 45 vs 41 us; in a group the ego call costs nothing anyway)
                    // a launch of >= 2 CTAs per SM carries the fill (padded with fill-only CTAs up to 12 per SM: det call
                    // backward 115 -> 104 us); below that (ego: 1 unit) a dedicated fill kernel is faster (37 vs 43 us)
                    if (vec_ok && pl.units >= 2 * 148 && !a.separate_zero_fill) {
                        gpk.zero_ptr = reinterpret_cast<uint4*>(a.g_feat);
                        gpk.zero_n16 = (long long)(gfeat_bytes / 16);
                    } else if (int e = separate_zero()) {
                        return e;
                    }
                    zero_done = true;
                }
                rc = launch_group_sample(true, a.type, gpk, pl.units, a.stream);
                if (rc != 0) return rc;
            } else {
                for (int k = 0; k < a.ncalls; ++k) {
                    Dims dk = d;
                    dk.A = a.calls[k].A; dk.P = a.calls[k].P;
                    SampleParams p = {};
                    p.feat = a.feat; p.shapes = a.shapes; p.starts = a.starts;
                    p.loc = a.calls[k].loc; p.weights = a.calls[k].weights;
                    p.grad_out = a.grad_out + (size_t)ids.anchor_begin[k] * d.C;     // ncalls == 1 or bs == 1: contiguous
                    p.g_loc = a.calls[k].g_loc; p.g_w = a.calls[k].g_w;
                    p.d = dk;
                    const int NP = dk.P * dk.cams;
                    const long long rows = (long long)dk.bs * dk.A;
                    p.S = choose_slices(rows, NP, 8 * 148, sample_smem_per_pair(kBwd, dk.L), /*max_slices=*/1 << 20);
                    if (p.S == 0) return -2;
                    p.PS = (NP + p.S - 1) / p.S;
                    const long long grid = rows * p.S;
                    if (grid > 0x7fffffffLL) return -2;
                    const int warps = choose_sample_warps(grid, dk.G);
                    const size_t smem = sample_smem_for(kBwd, dk, ks, a.type, p.PS, warps);
                    if (smem > kSampleSmemBudget) return -2;
                    if (!zero_done) {
                        if (vec_ok && grid >= 2 * 148 && !a.separate_zero_fill) {
                            p.zero_ptr = reinterpret_cast<uint4*>(a.g_feat);
                            p.zero_n16 = (long long)(gfeat_bytes / 16);
                        } else if (int e = separate_zero()) {
                            return e;
                        }
                        zero_done = true;
                    }
                    const int rc = (a.type == kF32)
                                       ? dispatch_sample<float, kBwd, false>(p, ks, warps, (int)grid, smem, a.stream)
                                       : dispatch_sample<__nv_bfloat16, kBwd, false>(p, ks, warps, (int)grid, smem, a.stream);
                    if (rc != 0) return rc;
                }
            }
            if (int e = debug_sync("sample-major backward kernel", a.stream)) return e;
        }
        return 0;
    };
    // HIPAD_DFA_CHAIN_FIRST=1 issues the chain's (short) launches before the sample kernel.  Measured with the chain at high
    // priority, stage-2 layer backward, bs = 1 / 4: sample kernel first 235 / 848 us, chain first 238 / 872 us (default: off).
    const bool chain_first = side != nullptr && hipad_env_int("HIPAD_DFA_CHAIN_FIRST", 0) != 0;
    if (!chain_first)
        if (int rc1 = run_k1()) return finish(rc1);
    if (a.g_feat == nullptr) return finish(0);   // caller does not need the feature-map gradient

    // ---- K2: visible-sample compaction, then per-(b,cam,level,band) sort by quad key
    unsigned char* ws = reinterpret_cast<unsigned char*>(a.workspace);
    GfeatParams gp = {};
    gp.shapes = a.shapes; gp.starts = a.starts;
    gp.ncalls = a.ncalls;
    for (int k = 0; k < a.ncalls; ++k) {
        gp.calls[k].loc = a.calls[k].loc; gp.calls[k].weights = a.calls[k].weights;
        gp.calls[k].A = a.calls[k].A; gp.calls[k].P = a.calls[k].P;
        gp.calls[k].id_begin = ids.id_begin[k]; gp.calls[k].anchor_begin = ids.anchor_begin[k];
    }
    gp.n_ids = (int)ids.n_ids;
    gp.A_total = (int)ids.a_total;
    gp.grad_out = a.grad_out; gp.g_feat = a.g_feat;
    gp.vis_key = reinterpret_cast<unsigned long long*>(ws + wl.vis_key);
    gp.chunk_cum = reinterpret_cast<int*>(ws + wl.chunk_cum);
    gp.ghist = reinterpret_cast<int*>(ws + wl.ghist);
    gp.rec = reinterpret_cast<int4*>(ws + wl.rec);
    gp.recw = reinterpret_cast<float*>(ws + wl.recw);
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
    // Band CTAs launched per (cam, level) bucket; a bucket uses ceil(samples / band_target) of them (the rest exit
    // after reading the bucket's histogram), so the launch covers the busiest camera and quiet cameras cost little
    gp.band_target = 768;
    {
        const long long most = (ids.n_ids + gp.band_target - 1) / gp.band_target;     // every sample seen by one camera
        gp.NB = (int)(most < 1 ? 1 : (most > kMaxBands ? kMaxBands : most));
    }
    {   // A/B knobs
        const int fb = hipad_env_int("HIPAD_DFA_BANDS", 0);
        if (fb >= 1 && fb <= kMaxBands) gp.NB = fb;
        const int bt = hipad_env_int("HIPAD_DFA_BAND_TARGET", 0);
        if (bt >= 64) {
            gp.band_target = bt;
            const long long most = (ids.n_ids + bt - 1) / bt;
            if (fb < 1) gp.NB = (int)(most < 1 ? 1 : (most > kMaxBands ? kMaxBands : most));
        }
        const int tm = hipad_env_int("HIPAD_DFA_TINY_MAX", kTinyRow);
        gp.tiny_max = tm < 1 ? 1 : (tm > kTinyRow ? kTinyRow : tm);
    }
    gp.accumulate = a.accumulate ? 1 : 0;
    gp.tiny_ok = (ks_g.vector && (d.C == 32 || d.C == 64 || d.C == 128 || d.C == 256) && (d.C / d.G) % 32 == 0 &&
                  hipad_env_int("HIPAD_DFA_TINY", 1) != 0) ? 1 : 0;
    if (a.stage_mask & 2) {
        cudaError_t e = cudaMemsetAsync(gp.ghist, 0, wl.ghist_bytes, chain);     // the compaction kernel adds to it
        if (e != cudaSuccess) return finish((int)e);
        dfa_vis_compact_kernel<<<dim3((unsigned)wl.n_chunks, (unsigned)d.cams, (unsigned)d.bs), kVisThreads, 0,
                                 chain>>>(gp);
        e = cudaGetLastError();
        if (e != cudaSuccess) return finish((int)e);
        // 256-thread sort CTAs (4 per SM) of ~768 samples; HIPAD_DFA_SORT_THREADS=512 selects 512-thread CTAs with twice
        // the shared memory.  Every phase of a band CTA is a chain of memory round trips, so what pays is CTAs per SM:
        // stage-2 layer, compaction + sort, bs = 1 / 4: 256 threads x 768 samples 60 / 124 us, 512 x 1536 62 / 140 us
        const bool wide = hipad_env_int("HIPAD_DFA_SORT_THREADS", 256) >= 512;
        const dim3 sgrid((unsigned)gp.NB, (unsigned)(d.cams * d.L), (unsigned)d.bs);
        if (wide) {
            const size_t sort_smem = band_sort_smem_bytes(wl.n_chunks, 512, 8192);
            e = ensure_smem(dfa_band_sort_kernel<512, 8192>, sort_smem);
            if (e != cudaSuccess) return finish((int)e);
            dfa_band_sort_kernel<512, 8192><<<sgrid, 512, sort_smem, chain>>>(gp);
        } else {
            const size_t sort_smem = band_sort_smem_bytes(wl.n_chunks, 256, 4096);
            e = ensure_smem(dfa_band_sort_kernel<256, 4096>, sort_smem);
            if (e != cudaSuccess) return finish((int)e);
            dfa_band_sort_kernel<256, 4096><<<sgrid, 256, sort_smem, chain>>>(gp);
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return finish((int)e);
        if (int e2 = debug_sync("compaction / band sort", chain)) return finish(e2);
    }
    if (!(a.stage_mask & 4)) return finish(0);

    // ---- K3: feature-major reduce, overwrites every touched row of the zero-filled g_feat once
    dfa_row_classify_kernel<<<dim3((unsigned)((d.num_feat + kClassifyThreads - 1) / kClassifyThreads), (unsigned)d.bs),
                              kClassifyThreads, 0, chain>>>(gp);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return finish((int)e);
    if (int e2 = debug_sync("row classification", chain)) return finish(e2);
    if (a.classify_only) return finish(0);
    if (chain_first)
        if (int rc1 = run_k1()) return finish(rc1);
    // join: the reduce needs the work lists (helper stream) and the zero-filled g_feat (caller's)
    if (int ej = finish(0)) return ej;
    const int rc = (a.type == kF32 || a.g_feat_f32) ? launch_reduce<float>(gp, ks_g, a.stream)
                                                    : launch_reduce<__nv_bfloat16>(gp, ks_g, a.stream);
    if (rc != 0) return rc;
    return debug_sync("feature-gradient reduce", a.stream);
}

}  // namespace hipad

#ifdef HIPAD_DFA_TRACE
// development builds: copy (and clear) the per-CTA phase trace of this translation unit's kernels
extern "C" int hipad_dfa_trace_read_gfeat(long long* host, long long count) {
    cudaDeviceSynchronize();
    const int rc = (int)cudaMemcpyFromSymbol(host, hipad::g_trace, sizeof(long long) * (size_t)count);
    void* sym = nullptr;
    cudaGetSymbolAddress(&sym, hipad::g_trace);
    cudaMemset(sym, 0, sizeof(hipad::g_trace));
    return rc;
}
#endif
