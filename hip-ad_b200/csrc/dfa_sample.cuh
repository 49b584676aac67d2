// dfa_sample.cuh — the sample-major kernel family (one CTA, or one cluster of CTAs, per output row).
//
// One template serves three entry points:
//   kFwd    out[b,a,:]  = sum_{valid (p,cam)} sum_l w * bilinear(feat)     (replaces cu:129-187)
//   kBwd    g_w, g_loc  = the sample-major half of the backward            (replaces cu:62-126 minus
//                          the feature scatter, which dfa_backward.cu does feature-major)
//   kFused  kFwd with the key-point projection (blocks.py:217-225) and the group softmax over
//           cams*L*P (blocks.py:196-208) folded in: sampling locations and weights never
//           round-trip through HBM.
//
// Work decomposition (B200: 148 SMs, 64 warps/SM):
//   * grid = bs*A*S CTAs of kWarps warps; S = 1,2,4,8 "point slices" per output row so that
//     map/plan shapes (A=100/480, P*cams=1800/540) still fill the machine.  For kFwd/kFused
//     the S CTAs of a row form a thread-block CLUSTER and combine their partial rows through
//     distributed shared memory in rank order (deterministic, no atomics, no second launch).
//   * phase 1: the CTA scans its (p,cam) pairs, tests visibility and ballot-compacts the
//     visible ones (10-20 % in practice) into shared memory in pair order.
//   * phase 2: warps stride over the compacted list; a lane owns V consecutive channels in
//     each of NCH 32-lane chunks (fp32: V=4 -> LDG.128 fully coalesced 512 B per warp
//     instruction; bf16: V=8).  Four corner rows x NCH vector loads per level are issued
//     back to back (all levels unrolled when kL>0) so each lane keeps >=8 16-byte loads in flight.
//   * phase 3: fixed-order cross-warp (+ cross-CTA) reduction, one coalesced store.
#pragma once
#include <cooperative_groups.h>

#include "dfa_common.cuh"

namespace hipad {
namespace cg = cooperative_groups;

enum SampleMode { kFwd = 0, kBwd = 1, kFused = 2 };

struct SampleParams {
    const void* feat;
    const int* shapes;
    const int* starts;
    const float* loc;      // kFwd/kBwd: [bs,A,P,cams,2]
    const float* weights;  // kFwd/kBwd: [bs,A,P,cams,L,G];  kFused: logits [bs,A,cams,L,P,G]
    float* out;            // kFwd/kFused
    const float* grad_out; // kBwd
    float* g_loc;          // kBwd
    float* g_w;            // kBwd
    const float* key_points;  // kFused [bs,A,P,3]
    const float* proj;        // kFused [bs,cams,4,4]
    const float* image_wh;    // kFused [bs,cams,2] or null
    float* loc_out;           // kFused optional [bs,A,P,cams,2]
    Dims d;
    int S;    // CTAs per output row
    int PS;   // (p,cam) pairs per slice = ceil(P*cams / S)
};

template <int kWarps>
__device__ __forceinline__ int block_compact_offset(bool flag, int* s_wcnt, int& total) {
    // ordered compaction across the CTA; returns this thread's slot (only meaningful if flag)
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const int c = s_wcnt[w];
        if (w < warp) before += c;
        all += c;
    }
    __syncthreads();
    total = all;
    return before + __popc(bal & ((1u << lane) - 1u));
}

// shared-memory carve-up (host mirrors this in sample_smem_bytes())
template <int kWarps>
__host__ __device__ inline size_t sample_smem_bytes(int mode, int n_cl, int cpad, int ps, int G, int cams) {
    size_t b = 0;
    b += (size_t)n_cl * 3 * sizeof(int);            // level table
    b += 16 * sizeof(int);                          // warp counters
    b = (b + 15) & ~(size_t)15;
    b += (size_t)kWarps * cpad * sizeof(float);     // cross-warp reduction / group scratch
    b += (size_t)cpad * sizeof(float);              // CTA partial row (cluster exchange)
    b += (size_t)ps * (sizeof(float2) + sizeof(int));  // compacted list
    if (mode == kBwd) b += ((size_t)ps + 15) & ~(size_t)15;  // validity bytes
    if (mode == kFused) b += (size_t)(cams * 14 + 4 * G + 2 * kWarps * 32) * sizeof(float);
    return (b + 15) & ~(size_t)15;
}

template <typename T, int V, int NCH, int kL, int kMode, bool kShfl, bool kCluster, int kWarps>
__global__ void __launch_bounds__(kWarps * 32) dfa_sample_kernel(const SampleParams p) {
    constexpr int kThreads = kWarps * 32;
    constexpr int CPAD = NCH * 32 * V;
    const Dims d = p.d;
    const int L = (kL > 0) ? kL : d.L;
    const int NP = d.P * d.cams;
    const int n_cl = d.cams * L;
    const int gd = d.C / d.G;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ba = blockIdx.x / p.S;
    const int slice = blockIdx.x - ba * p.S;
    const int b = ba / d.A;
    const int pair_lo = slice * p.PS;
    const int pair_hi = min(NP, pair_lo + p.PS);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* sp = smem_raw;
    int* tab = reinterpret_cast<int*>(sp);            sp += (size_t)n_cl * 3 * sizeof(int);
    int* s_wcnt = reinterpret_cast<int*>(sp);         sp += 16 * sizeof(int);
    sp = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sp) + 15) & ~(uintptr_t)15);
    float* red = reinterpret_cast<float*>(sp);        sp += (size_t)kWarps * CPAD * sizeof(float);
    float* part = reinterpret_cast<float*>(sp);       sp += (size_t)CPAD * sizeof(float);
    float2* l_xy = reinterpret_cast<float2*>(sp);     sp += (size_t)p.PS * sizeof(float2);
    int* l_pair = reinterpret_cast<int*>(sp);         sp += (size_t)p.PS * sizeof(int);
    unsigned char* s_valid = sp;
    if (kMode == kBwd) sp += ((size_t)p.PS + 15) & ~(size_t)15;
    float* s_proj = reinterpret_cast<float*>(sp);     // kFused: cams*12 matrix rows 0..2, cams*2 wh
    float* s_stat = s_proj + d.cams * 14;             // kFused: m[G], inv_s[G], scratch 2*G
    float* s_red2 = s_stat + 4 * d.G;                 // kFused: per-thread (m,s)

    load_level_table(tab, p.shapes, p.starts, n_cl);
    if (kMode == kFused) {
        for (int i = tid; i < d.cams * 12; i += kThreads)
            s_proj[i] = __ldg(p.proj + ((size_t)b * d.cams + i / 12) * 16 + (i % 12));
        for (int i = tid; i < d.cams * 2; i += kThreads)
            s_proj[d.cams * 12 + i] = p.image_wh ? __ldg(p.image_wh + (size_t)b * d.cams * 2 + i) : 1.f;
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 1: visible pairs
    int n_list = 0;
    for (int base = pair_lo; base < pair_hi; base += kThreads) {
        const int pair = base + tid;
        bool vis = false;
        float2 xy = make_float2(0.f, 0.f);
        if (pair < pair_hi) {
            if (kMode == kFused) {
                const int pt = pair / d.cams, cam = pair - pt * d.cams;
                const float* kp = p.key_points + ((size_t)ba * d.P + pt) * 3;
                const float X = __ldg(kp), Y = __ldg(kp + 1), Z = __ldg(kp + 2);
                const float* m = s_proj + cam * 12;
                // row . [X,Y,Z,1] accumulated left to right like a 4-term dot product
                const float px = __fmaf_rn(m[2], Z, __fmaf_rn(m[1], Y, m[0] * X)) + m[3];
                const float py = __fmaf_rn(m[6], Z, __fmaf_rn(m[5], Y, m[4] * X)) + m[7];
                const float pz = __fmaf_rn(m[10], Z, __fmaf_rn(m[9], Y, m[8] * X)) + m[11];
                const float zc = fmaxf(pz, 1e-5f);
                xy.x = __fdiv_rn(px, zc);
                xy.y = __fdiv_rn(py, zc);
                if (p.image_wh) {
                    xy.x = __fdiv_rn(xy.x, s_proj[d.cams * 12 + cam * 2]);
                    xy.y = __fdiv_rn(xy.y, s_proj[d.cams * 12 + cam * 2 + 1]);
                }
                if (p.loc_out) reinterpret_cast<float2*>(p.loc_out)[(size_t)ba * NP + pair] = xy;
            } else {
                xy = __ldg(reinterpret_cast<const float2*>(p.loc) + (size_t)ba * NP + pair);
            }
            vis = loc_valid(xy.x, xy.y);
            if (kMode == kBwd) {
                s_valid[pair - pair_lo] = vis ? 1 : 0;
                if (!vis) reinterpret_cast<float2*>(p.g_loc)[(size_t)ba * NP + pair] = make_float2(0.f, 0.f);
            }
        }
        int total;
        const int slot = n_list + block_compact_offset<kWarps>(vis, s_wcnt, total);
        if (vis) {
            l_xy[slot] = xy;
            l_pair[slot] = pair;
        }
        n_list += total;
    }
    __syncthreads();

    if (kMode == kBwd) {
        // zero the weight-gradient rows of invisible pairs (each row = L*G floats, contiguous)
        const int lg = L * d.G;
        float* gw_base = p.g_w + ((size_t)ba * NP + pair_lo) * lg;
        const int n_pairs = pair_hi - pair_lo;
        if ((lg & 3) == 0) {
            const int q_per = lg >> 2;
            for (int q = tid; q < n_pairs * q_per; q += kThreads)
                if (!s_valid[q / q_per]) reinterpret_cast<float4*>(gw_base)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (int q = tid; q < n_pairs * lg; q += kThreads)
                if (!s_valid[q / lg]) gw_base[q] = 0.f;
        }
    }

    // ------------------------------------------------------------------ fused: softmax statistics
    float sm_m[NCH], sm_inv[NCH];
    if (kMode == kFused) {
        // logits of this output row: [cams, L, P, G], softmax jointly over cams*L*P per group.
        const int n_log = d.cams * L * d.P * d.G;
        const float* lg_row = p.weights + (size_t)ba * n_log;
        // each CTA of the cluster scans 1/S of the logits; statistics are exchanged via DSMEM
        int chunk = (n_log + p.S - 1) / p.S;
        chunk = (chunk + kThreads - 1) / kThreads * kThreads;   // keeps (idx % G) fixed per thread
        const int lo = min(n_log, slice * chunk), hi = min(n_log, lo + chunk);
        float m_run = -INFINITY, s_run = 0.f;
        for (int i = lo + tid; i < hi; i += kThreads) {
            const float x = __ldg(lg_row + i);
            if (x > m_run) {
                s_run = s_run * expf(m_run - x) + 1.f;
                m_run = x;
            } else {
                s_run += expf(x - m_run);
            }
        }
        s_red2[tid * 2] = m_run;
        s_red2[tid * 2 + 1] = s_run;
        __syncthreads();
        if (tid < d.G) {   // kThreads % G == 0 (checked on the host): thread t saw group (lo+t) % G
            float m = -INFINITY, s = 0.f;
            for (int t = 0; t < kThreads; ++t) {
                if ((lo + t) % d.G != tid) continue;
                const float mt = s_red2[t * 2], st = s_red2[t * 2 + 1];
                if (st == 0.f) continue;
                const float mn = fmaxf(m, mt);
                s = s * expf(m - mn) + st * expf(mt - mn);
                m = mn;
            }
            s_stat[2 * d.G + tid * 2] = m;
            s_stat[2 * d.G + tid * 2 + 1] = s;
        }
        if (kCluster) {
            cg::cluster_group cl = cg::this_cluster();
            cl.sync();
            if (tid < d.G) {
                float m = -INFINITY, s = 0.f;
                for (int r = 0; r < p.S; ++r) {
                    const float* rs = cl.map_shared_rank(s_stat, r);
                    const float mt = rs[2 * d.G + tid * 2], st = rs[2 * d.G + tid * 2 + 1];
                    if (st == 0.f) continue;
                    const float mn = fmaxf(m, mt);
                    s = s * expf(m - mn) + st * expf(mt - mn);
                    m = mn;
                }
                s_stat[tid] = m;
                s_stat[d.G + tid] = 1.f / s;
            }
        } else {
            __syncthreads();
            if (tid < d.G) {
                s_stat[tid] = s_stat[2 * d.G + tid * 2];
                s_stat[d.G + tid] = 1.f / s_stat[2 * d.G + tid * 2 + 1];
            }
        }
        __syncthreads();
    }

    // ------------------------------------------------------------------ phase 2: gather
    int ch[NCH], grp[NCH];
    bool act[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        ch[j] = (j * 32 + lane) * V;
        act[j] = ch[j] < d.C;
        grp[j] = act[j] ? ch[j] / gd : 0;
        if (kMode == kFused) {
            sm_m[j] = s_stat[grp[j]];
            sm_inv[j] = s_stat[d.G + grp[j]];
        }
    }
    const int lpg = (gd / V) > 0 ? (gd / V) : 1;   // lanes per group (kShfl: power of two <= 32)

    float acc[NCH][V];
    float go[NCH][V];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
        for (int e = 0; e < V; ++e) { acc[j][e] = 0.f; go[j][e] = 0.f; }
    if (kMode == kBwd) {
#pragma unroll
        for (int j = 0; j < NCH; ++j)
            if (act[j]) VecIO<float, V>::load(p.grad_out + (size_t)ba * d.C + ch[j], go[j]);
    }

    const T* feat = reinterpret_cast<const T*>(p.feat);
    const size_t feat_b = (size_t)b * d.num_feat;

    for (int i = warp; i < n_list; i += kWarps) {
        const float2 xy = l_xy[i];
        const int pair = l_pair[i];
        const int pt = pair / d.cams, cam = pair - pt * d.cams;
        const int* t = tab + cam * L * 3;
        const float* wrow = (kMode == kFused)
                                ? p.weights + ((size_t)ba * d.cams + cam) * L * d.P * d.G + (size_t)pt * d.G
                                : p.weights + ((size_t)ba * NP + pair) * L * d.G;
        float gx = 0.f, gy = 0.f;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const int h = t[l * 3], w = t[l * 3 + 1];
            const Quad q = quad_setup(xy.x, xy.y, h, w);
            const T* r1 = feat + (feat_b + t[l * 3 + 2] + (ptrdiff_t)q.h_low * w + q.w_low) * d.C;
            const T* r3 = r1 + (size_t)w * d.C;
            const float c1 = q.hh * q.hw, c2 = q.hh * q.lw, c3 = q.lh * q.hw, c4 = q.lh * q.lw;
            float wv[NCH];
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                if (kMode == kFused) {
                    const float x = act[j] ? __ldg(wrow + (size_t)l * d.P * d.G + grp[j]) : 0.f;
                    wv[j] = expf(x - sm_m[j]) * sm_inv[j];
                } else {
                    wv[j] = act[j] ? __ldg(wrow + l * d.G + grp[j]) : 0.f;
                }
            }
            float v1[NCH][V], v2[NCH][V], v3[NCH][V], v4[NCH][V];
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
#pragma unroll
                for (int e = 0; e < V; ++e) { v1[j][e] = 0.f; v2[j][e] = 0.f; v3[j][e] = 0.f; v4[j][e] = 0.f; }
                if (act[j]) {
                    if (q.ok1) VecIO<T, V>::load(r1 + ch[j], v1[j]);
                    if (q.ok2) VecIO<T, V>::load(r1 + d.C + ch[j], v2[j]);
                    if (q.ok3) VecIO<T, V>::load(r3 + ch[j], v3[j]);
                    if (q.ok4) VecIO<T, V>::load(r3 + d.C + ch[j], v4[j]);
                }
            }
            if (kMode != kBwd) {
#pragma unroll
                for (int j = 0; j < NCH; ++j)
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const float val = c1 * v1[j][e] + c2 * v2[j][e] + c3 * v3[j][e] + c4 * v4[j][e];
                        acc[j][e] = __fmaf_rn(val, wv[j], acc[j][e]);
                    }
            } else {
                float gxl = 0.f, gyl = 0.f;
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    float gw = 0.f, dxs = 0.f, dys = 0.f;
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const float val = c1 * v1[j][e] + c2 * v2[j][e] + c3 * v3[j][e] + c4 * v4[j][e];
                        // d(val)/d(w_im) and d(val)/d(h_im)  (cu:86-121)
                        const float dw = q.hh * (v2[j][e] - v1[j][e]) + q.lh * (v4[j][e] - v3[j][e]);
                        const float dh = q.hw * (v3[j][e] - v1[j][e]) + q.lw * (v4[j][e] - v2[j][e]);
                        gw = __fmaf_rn(go[j][e], val, gw);
                        dxs = __fmaf_rn(go[j][e], dw, dxs);
                        dys = __fmaf_rn(go[j][e], dh, dys);
                    }
                    gxl = __fmaf_rn(dxs, wv[j], gxl);
                    gyl = __fmaf_rn(dys, wv[j], gyl);
                    // weight gradient: reduce over the channels of the group
                    float* gw_dst = p.g_w + (((size_t)ba * NP + pair) * L + l) * d.G;
                    if (kShfl) {
                        for (int o = lpg >> 1; o > 0; o >>= 1) gw += __shfl_xor_sync(0xffffffffu, gw, o);
                        if (act[j] && (lane & (lpg - 1)) == 0) gw_dst[grp[j]] = gw;
                    } else {
                        // generic group sizes: stage per-channel terms, then one lane per group sums
                        float* sc = red + warp * CPAD;
                        sc[ch[j]] = act[j] ? gw : 0.f;   // V == 1 on this path
                        __syncwarp();
                        if (j == NCH - 1) {
                            for (int g = lane; g < d.G; g += 32) {
                                float s = 0.f;
                                for (int c = g * gd; c < (g + 1) * gd; ++c) s += sc[c];
                                gw_dst[g] = s;
                            }
                            __syncwarp();
                        }
                    }
                }
                gx = __fmaf_rn((float)w, gxl, gx);
                gy = __fmaf_rn((float)h, gyl, gy);
            }
        }
        if (kMode == kBwd) {
            gx = warp_sum(gx);
            gy = warp_sum(gy);
            if (lane == 0) reinterpret_cast<float2*>(p.g_loc)[(size_t)ba * NP + pair] = make_float2(gx, gy);
        }
    }

    // ------------------------------------------------------------------ phase 3: reduce + store
    if (kMode != kBwd) {
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int e = 0; e < V; ++e) red[warp * CPAD + (j * 32 + lane) * V + e] = acc[j][e];
        __syncthreads();
        float* out_row = p.out + (size_t)ba * d.C;
        for (int c = tid; c < d.C; c += kThreads) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += red[w * CPAD + c];
            if (kCluster) part[c] = s; else out_row[c] = s;
        }
        if (kCluster) {
            cg::cluster_group cl = cg::this_cluster();
            cl.sync();
            if (slice == 0) {
                for (int c = tid; c < d.C; c += kThreads) {
                    float s = part[c];
                    for (int r = 1; r < p.S; ++r) s += cl.map_shared_rank(part, r)[c];
                    out_row[c] = s;
                }
            }
            cl.sync();
        }
    }
}

}  // namespace hipad
