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

// shared-memory carve-up, as byte OFFSETS from the dynamic-smem base (all multiples of 16) so the
// compiler keeps every access in the shared address space (LDS/STS, 32-bit addressing).
struct SampleSmem {
    int tab, wcnt, red, part, lxy, lpair, mrows, mcoef, valid, gxy, proj, stat, red2, total;
};
template <int kWarps>
__host__ __device__ inline SampleSmem sample_smem_layout(int mode, int n_cl, int L, int cpad, int ps, int G, int cams) {
    SampleSmem o;
    int b = 0;
    auto take = [&](int bytes) { const int at = b; b += (bytes + 15) & ~15; return at; };
    o.tab = take(n_cl * 3 * 4);
    o.wcnt = take(16 * 4);
    o.red = take(kWarps * cpad * 4);       // cross-warp reduction / group scratch
    o.part = take(cpad * 4);               // CTA partial row (cluster exchange)
    o.lxy = take(ps * 8);                  // compacted list: locations
    o.lpair = take(ps * 4);                //                 pair ids
    o.mrows = take(ps * L * 16);           // per (visible pair, level): 4 corner rows
    o.mcoef = take(ps * L * 16);           //                            4 bilinear terms
    o.valid = take(mode == kBwd ? ps : 0);
    o.gxy = take(mode == kBwd ? ps * L * 8 : 0);   // per (pair, level) location-gradient partials
    o.proj = take(mode == kFused ? cams * 14 * 4 : 0);
    o.stat = take(mode == kFused ? 4 * G * 4 : 0);
    o.red2 = take(mode == kFused ? 2 * kWarps * 32 * 4 : 0);
    o.total = b;
    return o;
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
    // slice s owns pairs s, s+S, s+2S, ... (interleaved, so the S CTAs of a row see the same mix
    // of cameras / key points and finish together)
    const int n_mine = (NP - slice + p.S - 1) / p.S;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SampleSmem so = sample_smem_layout<kWarps>(kMode, n_cl, L, CPAD, p.PS, d.G, d.cams);
    int* tab = reinterpret_cast<int*>(smem_raw + so.tab);
    int* s_wcnt = reinterpret_cast<int*>(smem_raw + so.wcnt);
    float* red = reinterpret_cast<float*>(smem_raw + so.red);
    float* part = reinterpret_cast<float*>(smem_raw + so.part);
    float2* l_xy = reinterpret_cast<float2*>(smem_raw + so.lxy);
    int* l_pair = reinterpret_cast<int*>(smem_raw + so.lpair);
    int4* m_rows = reinterpret_cast<int4*>(smem_raw + so.mrows);     // element rows of the 4 corners (clamped)
    float4* m_coef = reinterpret_cast<float4*>(smem_raw + so.mcoef); // fwd: c1..c4   bwd: lh, lw, ok, (h<<16|w)
    unsigned char* s_valid = smem_raw + so.valid;
    float2* s_gxy = reinterpret_cast<float2*>(smem_raw + so.gxy);
    float* s_proj = reinterpret_cast<float*>(smem_raw + so.proj);    // kFused: cams*12 matrix rows 0..2, cams*2 wh
    float* s_stat = reinterpret_cast<float*>(smem_raw + so.stat);    // kFused: m[G], inv_s[G], scratch 2*G
    float* s_red2 = reinterpret_cast<float*>(smem_raw + so.red2);    // kFused: per-thread (m,s)

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
    for (int base = 0; base < n_mine; base += kThreads) {
        const int k_mine = base + tid;
        const int pair = slice + k_mine * p.S;
        bool vis = false;
        float2 xy = make_float2(0.f, 0.f);
        if (k_mine < n_mine) {
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
                s_valid[k_mine] = vis ? 1 : 0;
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
        float* gw_base = p.g_w + (size_t)ba * NP * lg;
        if ((lg & 3) == 0) {
            const int q_per = lg >> 2;
            for (int q = tid; q < n_mine * q_per; q += kThreads) {
                const int k_mine = q / q_per;
                if (!s_valid[k_mine])
                    reinterpret_cast<float4*>(gw_base)[(size_t)(slice + k_mine * p.S) * q_per + (q - k_mine * q_per)] =
                        make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            for (int q = tid; q < n_mine * lg; q += kThreads) {
                const int k_mine = q / lg;
                if (!s_valid[k_mine]) gw_base[(size_t)(slice + k_mine * p.S) * lg + (q - k_mine * lg)] = 0.f;
            }
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

    // ------------------------------------------------------------------ phase 1b: gather metadata
    // One thread per (visible pair, level) turns the location into corner rows and bilinear terms
    // ONCE, instead of all 32 lanes of the consuming warp redoing the same scalar math.  Rows of
    // out-of-bounds corners are redirected to an in-bounds corner of the SAME quad and their
    // coefficient is zeroed, so phase 2 can issue its loads unconditionally (a valid sample always
    // has at least one in-bounds corner; reading a corner the quad reads anyway keeps NaN/Inf
    // behaviour identical to skipping the load).
    for (int it = tid; it < n_list * L; it += kThreads) {
        const int i = it / L, l = it - i * L;
        const float2 xy = l_xy[i];
        const int pair = l_pair[i];
        const int cam = pair - (pair / d.cams) * d.cams;
        const int* t = tab + (cam * L + l) * 3;
        const int h = t[0], w = t[1];
        const Quad q = quad_setup(xy.x, xy.y, h, w);
        const int r1 = t[2] + q.h_low * w + q.w_low, r2 = r1 + 1, r3 = r1 + w, r4 = r3 + 1;
        const int safe = q.ok1 ? r1 : (q.ok2 ? r2 : (q.ok3 ? r3 : r4));
        m_rows[it] = make_int4(q.ok1 ? r1 : safe, q.ok2 ? r2 : safe, q.ok3 ? r3 : safe, q.ok4 ? r4 : safe);
        if (kMode != kBwd) {
            m_coef[it] = make_float4(q.ok1 ? q.hh * q.hw : 0.f, q.ok2 ? q.hh * q.lw : 0.f,
                                     q.ok3 ? q.lh * q.hw : 0.f, q.ok4 ? q.lh * q.lw : 0.f);
        } else {
            const int ok = (int)q.ok1 | ((int)q.ok2 << 1) | ((int)q.ok3 << 2) | ((int)q.ok4 << 3);
            m_coef[it] = make_float4(q.lh, q.lw, __int_as_float(ok), __int_as_float((h << 16) | w));
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 2: gather
    // Work item = one (visible pair, level): 4 corner rows x NCH vector loads per lane.  Items are
    // dealt round-robin to the warps of the CTA (fine-grained: an anchor's serial chain per warp is
    // n_items / kWarps) and software-pipelined two deep: the loads of item n+1 are in flight while
    // item n is consumed, i.e. 2 x 4 x NCH 16-byte loads outstanding per lane.
    // Lanes whose channels fall beyond C (C < NCH*32*V) are clamped onto the last valid vector:
    // they load real data and compute values nobody reads, which keeps the loop branch-free.
    int ch[NCH], grp[NCH];
    bool act[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int c_raw = (j * 32 + lane) * V;
        act[j] = c_raw < d.C;
        ch[j] = act[j] ? c_raw : d.C - V;
        grp[j] = ch[j] / gd;
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
            if (act[j]) VecIO<float, V>::load(p.grad_out + (size_t)ba * d.C + ch[j], go[j]);   // inactive: stays 0
    }

    // element offsets inside one batch element fit 32 bits (num_feat*C < 2^31, checked on the host)
    const T* featb = reinterpret_cast<const T*>(p.feat) + (size_t)b * d.num_feat * d.C;
    const unsigned Cu = (unsigned)d.C;
    const int n_items = n_list * L;

    struct Item {
        float v[4][NCH][V];
        float wv[NCH];
        float4 cf;
        int pair, l;
    };

    auto issue = [&](int it, Item& r) {
        const int i = it / L;
        r.l = it - i * L;
        r.pair = l_pair[i];
        const int4 rw = m_rows[it];
        r.cf = m_coef[it];
        const int pt = r.pair / d.cams, cam = r.pair - pt * d.cams;
        const float* wrow = (kMode == kFused)
                                ? p.weights + (((size_t)ba * d.cams + cam) * L + r.l) * d.P * d.G + (size_t)pt * d.G
                                : p.weights + (((size_t)ba * NP + r.pair) * L + r.l) * d.G;
#pragma unroll
        for (int j = 0; j < NCH; ++j) r.wv[j] = __ldg(wrow + grp[j]);
        // one 64-bit row pointer per corner; the NCH chunks are constant offsets from it
        const T* q1 = featb + (size_t)((unsigned)rw.x * Cu);
        const T* q2 = featb + (size_t)((unsigned)rw.y * Cu);
        const T* q3 = featb + (size_t)((unsigned)rw.z * Cu);
        const T* q4 = featb + (size_t)((unsigned)rw.w * Cu);
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            VecIO<T, V>::load(q1 + ch[j], r.v[0][j]);
            VecIO<T, V>::load(q2 + ch[j], r.v[1][j]);
            VecIO<T, V>::load(q3 + ch[j], r.v[2][j]);
            VecIO<T, V>::load(q4 + ch[j], r.v[3][j]);
        }
    };

    auto consume = [&](int it, const Item& r) {
        if (kMode != kBwd) {
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                float wj = r.wv[j];
                if (kMode == kFused) wj = expf(wj - sm_m[j]) * sm_inv[j];
                // weight folded into the four bilinear coefficients: 4 FMAs per channel
                const float k1 = r.cf.x * wj, k2 = r.cf.y * wj, k3 = r.cf.z * wj, k4 = r.cf.w * wj;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    float a_ = acc[j][e];
                    a_ = __fmaf_rn(k1, r.v[0][j][e], a_);
                    a_ = __fmaf_rn(k2, r.v[1][j][e], a_);
                    a_ = __fmaf_rn(k3, r.v[2][j][e], a_);
                    a_ = __fmaf_rn(k4, r.v[3][j][e], a_);
                    acc[j][e] = a_;
                }
            }
        } else {
            const float lh = r.cf.x, lw = r.cf.y, hh = 1.f - lh, hw = 1.f - lw;
            const int ok = __float_as_int(r.cf.z), hwp = __float_as_int(r.cf.w);
            const bool o1 = ok & 1, o2 = ok & 2, o3 = ok & 4, o4 = ok & 8;
            // value, d/dw and d/dh as linear forms of the four corner values (cu:86-121)
            const float c1 = o1 ? hh * hw : 0.f, c2 = o2 ? hh * lw : 0.f, c3 = o3 ? lh * hw : 0.f, c4 = o4 ? lh * lw : 0.f;
            const float a1 = o1 ? -hh : 0.f, a2 = o2 ? hh : 0.f, a3 = o3 ? -lh : 0.f, a4 = o4 ? lh : 0.f;
            const float b1 = o1 ? -hw : 0.f, b2 = o2 ? -lw : 0.f, b3 = o3 ? hw : 0.f, b4 = o4 ? lw : 0.f;
            float gxl = 0.f, gyl = 0.f;
            float* gw_dst = p.g_w + (((size_t)ba * NP + r.pair) * L + r.l) * d.G;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                // s_k = <grad_out, corner_k> over this lane's channels, then three 4-term forms
                float s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    s1 = __fmaf_rn(go[j][e], r.v[0][j][e], s1);
                    s2 = __fmaf_rn(go[j][e], r.v[1][j][e], s2);
                    s3 = __fmaf_rn(go[j][e], r.v[2][j][e], s3);
                    s4 = __fmaf_rn(go[j][e], r.v[3][j][e], s4);
                }
                float gw = c1 * s1 + c2 * s2 + c3 * s3 + c4 * s4;
                const float dxs = a1 * s1 + a2 * s2 + a3 * s3 + a4 * s4;
                const float dys = b1 * s1 + b2 * s2 + b3 * s3 + b4 * s4;
                gxl = __fmaf_rn(dxs, r.wv[j], gxl);
                gyl = __fmaf_rn(dys, r.wv[j], gyl);
                // weight gradient: reduce over the channels of the group
                if (kShfl) {
                    for (int o = lpg >> 1; o > 0; o >>= 1) gw += __shfl_xor_sync(0xffffffffu, gw, o);
                    if (act[j] && (lane & (lpg - 1)) == 0) gw_dst[grp[j]] = gw;
                } else {
                    // generic group sizes: stage per-channel terms, then one lane per group sums
                    float* sc = red + warp * CPAD;
                    sc[j * 32 + lane] = act[j] ? gw : 0.f;   // V == 1 on this path
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
            // location gradient of this level; the L levels of a pair are summed in phase 3
            const float gx = warp_sum((float)(hwp & 0xffff) * gxl);
            const float gy = warp_sum((float)(hwp >> 16) * gyl);
            if (lane == 0) s_gxy[it] = make_float2(gx, gy);
        }
    };

    if constexpr (NCH * V <= 8) {
        Item A, B;
        if (warp < n_items) issue(warp, A);
        for (int it = warp; it < n_items; it += 2 * kWarps) {
            const int itB = it + kWarps;
            if (itB < n_items) issue(itB, B);
            consume(it, A);
            const int itA = itB + kWarps;
            if (itA < n_items) issue(itA, A);
            if (itB < n_items) consume(itB, B);
        }
    } else {   // wide rows: one item's loads already fill the register budget
        Item A;
        for (int it = warp; it < n_items; it += kWarps) {
            issue(it, A);
            consume(it, A);
        }
    }

    if (kMode == kBwd) {
        // g_loc[pair] = sum over levels, fixed order
        __syncthreads();
        for (int i = tid; i < n_list; i += kThreads) {
            float gx = 0.f, gy = 0.f;
            for (int l = 0; l < L; ++l) {
                const float2 g = s_gxy[i * L + l];
                gx += g.x;
                gy += g.y;
            }
            reinterpret_cast<float2*>(p.g_loc)[(size_t)ba * NP + l_pair[i]] = make_float2(gx, gy);
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
