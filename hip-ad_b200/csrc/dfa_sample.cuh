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
    uint4* zero_ptr;          // kBwd optional: dense buffer (grad_mc_ms_feat) this launch zero-fills on the side
    long long zero_n16;       //               its size in 16-byte units
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
    int tab, wcnt, red, part, lxy, lpair, mrows, mcoef, mwoff, valid, mdx, mdy, proj, stat, red2, total;
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
    o.mwoff = take(ps * L * 4);            //                            weight offset
    o.valid = take(mode == kBwd ? ps : 0);
    o.mdx = take(mode == kBwd ? ps * L * 16 : 0);  // bwd: d/dx coefficient vectors (scaled by W)
    o.mdy = take(mode == kBwd ? ps * L * 16 : 0);  //      d/dy coefficient vectors (scaled by H)
    o.proj = take(mode == kFused ? cams * 14 * 4 : 0);
    o.stat = take(mode == kFused ? 4 * G * 4 : 0);
    o.red2 = take(mode == kFused ? 2 * kWarps * 32 * 4 : 0);
    o.total = b;
    return o;
}

// kLPG: lanes per channel group. >0 compile-time (power of two), 0 run-time (power of two), -1 generic groups
template <typename T, int V, int NCH, int kL, int kMode, int kLPG, bool kCluster, int kWarps>
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
    float4* m_coef = reinterpret_cast<float4*>(smem_raw + so.mcoef); // c1..c4 (0 for out-of-bounds corners)
    int* m_woff = reinterpret_cast<int*>(smem_raw + so.mwoff);
    unsigned char* s_valid = smem_raw + so.valid;
    float4* m_dx = reinterpret_cast<float4*>(smem_raw + so.mdx);
    float4* m_dy = reinterpret_cast<float4*>(smem_raw + so.mdy);
    float* s_proj = reinterpret_cast<float*>(smem_raw + so.proj);    // kFused: cams*12 matrix rows 0..2, cams*2 wh
    float* s_stat = reinterpret_cast<float*>(smem_raw + so.stat);    // kFused: m[G], inv_s[G], scratch 2*G
    float* s_red2 = reinterpret_cast<float*>(smem_raw + so.red2);    // kFused: per-thread (m,s)

    if (kMode == kBwd && p.zero_n16 > 0) {
        // Dense zero fill of the feature gradient, 1/gridDim of it per CTA.  The stores are fire-and-forget:
        // they drain to HBM while this CTA waits on its gather loads (the write path is otherwise idle here).
        const long long per = (p.zero_n16 + gridDim.x - 1) / gridDim.x;
        const long long z0 = (long long)blockIdx.x * per;
        const long long z1 = (z0 + per < p.zero_n16) ? z0 + per : p.zero_n16;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (long long i = z0 + tid; i < z1; i += kThreads) p.zero_ptr[i] = z;
    }
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
        for (int i0 = lo + tid; i0 < hi; i0 += kThreads * 8) {
            // eight independent loads in flight per thread, then one max and one rescale for the batch
            float x[8];
            float mx = m_run;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + u * kThreads;
                x[u] = (i < hi) ? __ldg(lg_row + i) : -INFINITY;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) mx = fmaxf(mx, x[u]);
            float acc = (s_run == 0.f) ? 0.f : s_run * expf(m_run - mx);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc += expf(x[u] - mx);     // exp(-inf) = 0 for the padding
            m_run = mx;
            s_run = acc;
        }
        // combine the per-thread (max, sum) pairs of each group: thread t saw group (lo + t) % G = t % G (lo is a
        // multiple of kThreads, kThreads % G == 0).  When G divides 32 the lanes of a warp that share a group are
        // merged with shuffles first, so the final per-group loop runs over kWarps entries instead of kThreads.
        const bool g_pow2 = (d.G & (d.G - 1)) == 0 && d.G <= 32;
        if (g_pow2) {
            for (int o = 16; o >= d.G; o >>= 1) {
                const float mo = __shfl_xor_sync(0xffffffffu, m_run, o), so_ = __shfl_xor_sync(0xffffffffu, s_run, o);
                const float mn = fmaxf(m_run, mo);
                const float a_ = (s_run == 0.f) ? 0.f : s_run * expf(m_run - mn);
                const float b_ = (so_ == 0.f) ? 0.f : so_ * expf(mo - mn);
                m_run = mn;
                s_run = a_ + b_;
            }
            if (lane < d.G) {
                s_red2[(warp * d.G + lane) * 2] = m_run;
                s_red2[(warp * d.G + lane) * 2 + 1] = s_run;
            }
        } else {
            s_red2[tid * 2] = m_run;
            s_red2[tid * 2 + 1] = s_run;
        }
        __syncthreads();
        if (tid < d.G) {
            float m = -INFINITY, s = 0.f;
            const int n_ent = g_pow2 ? kWarps : kThreads;
            for (int t = 0; t < n_ent; ++t) {
                int at = t;
                if (g_pow2) at = t * d.G + tid;
                else if (t % d.G != tid) continue;
                const float mt = s_red2[at * 2], st = s_red2[at * 2 + 1];
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
        const int pt = pair / d.cams, cam = pair - pt * d.cams;
        const int* t = tab + (cam * L + l) * 3;
        const int h = t[0], w = t[1];
        const Quad q = quad_setup(xy.x, xy.y, h, w);
        const int r1 = t[2] + q.h_low * w + q.w_low, r2 = r1 + 1, r3 = r1 + w, r4 = r3 + 1;
        const int safe = q.ok1 ? r1 : (q.ok2 ? r2 : (q.ok3 ? r3 : r4));
        // BYTE offsets of the corner rows inside this batch element (< 2^32, checked on the host)
        const unsigned row_bytes = (unsigned)d.C * (unsigned)sizeof(T);
        m_rows[it] = make_int4((int)((unsigned)(q.ok1 ? r1 : safe) * row_bytes), (int)((unsigned)(q.ok2 ? r2 : safe) * row_bytes),
                               (int)((unsigned)(q.ok3 ? r3 : safe) * row_bytes), (int)((unsigned)(q.ok4 ? r4 : safe) * row_bytes));
        // element offset of this item's G weights inside the output row's weight block
        m_woff[it] = (kMode == kFused) ? ((cam * L + l) * d.P + pt) * d.G : (pair * L + l) * d.G;
        m_coef[it] = make_float4(q.ok1 ? q.hh * q.hw : 0.f, q.ok2 ? q.hh * q.lw : 0.f,
                                 q.ok3 ? q.lh * q.hw : 0.f, q.ok4 ? q.lh * q.lw : 0.f);
        if (kMode == kBwd) {
            // d(val)/d(loc_x) = W * (-hh v1 + hh v2 - lh v3 + lh v4), d/d(loc_y) = H * (-hw v1 - lw v2 + hw v3 + lw v4)
            // (cu:86-125), as coefficient vectors with out-of-bounds corners zeroed
            const float W_ = (float)w, H_ = (float)h;
            m_dx[it] = make_float4(q.ok1 ? -q.hh * W_ : 0.f, q.ok2 ? q.hh * W_ : 0.f,
                                   q.ok3 ? -q.lh * W_ : 0.f, q.ok4 ? q.lh * W_ : 0.f);
            m_dy[it] = make_float4(q.ok1 ? -q.hw * H_ : 0.f, q.ok2 ? -q.lw * H_ : 0.f,
                                   q.ok3 ? q.hw * H_ : 0.f, q.ok4 ? q.lw * H_ : 0.f);
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
    const int lpg = (kLPG > 0) ? kLPG : ((gd / V) > 0 ? (gd / V) : 1);   // lanes per group (power of two <= 32 unless kLPG < 0)

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

    // Addressing: one 64-bit base per tensor (uniform), everything else 32-bit.  On the vector path
    // with NCH > 1 every lane is active (C == NCH*32*V), so chunk j is a compile-time constant
    // offset from chunk 0 and folds into the load's immediate field.
    const char* featb = reinterpret_cast<const char*>(p.feat) + (size_t)b * d.num_feat * d.C * sizeof(T);
    const unsigned lane_byte = (unsigned)ch[0] * (unsigned)sizeof(T);
    const float* wblock = (kMode == kFused) ? p.weights + (size_t)ba * d.cams * L * d.P * d.G
                                            : p.weights + (size_t)ba * NP * L * d.G;
    const int n_items = n_list * L;
    constexpr bool kConstChunks = (V > 1);   // vector path: full chunks (host guarantees it when NCH > 1)

    // bf16 x8 rows at one chunk per lane stay PACKED (4 registers per corner) while in flight, which lets the gather
    // run 4 items deep in the registers the fp32 path needs for 2; every other instantiation is unchanged
    constexpr bool kPacked = (sizeof(T) == 2 && V == 8 && NCH == 1);
    struct Item {
        float v[4][kPacked ? 1 : NCH][kPacked ? 1 : V];
        uint4 raw[kPacked ? 4 : 1];
        float wv[NCH];
        float4 cf;
        int woff;
    };

    // Warp `warp` owns visible pairs warp, warp+kWarps, ... and walks their L levels back to back
    // (sequence index q -> pair warp + kWarps*(q / L), level q % L), so location gradients
    // accumulate in-lane over the levels and need one warp reduction per PAIR.
    const int my_pairs = (n_list > warp) ? (n_list - warp + kWarps - 1) / kWarps : 0;
    const int my_items = my_pairs * L;
    auto item_of = [&](int q) { const int k = q / L; return (warp + k * kWarps) * L + (q - k * L); };

    auto issue = [&](int it, Item& r) {
        const int4 rw = m_rows[it];
        r.cf = m_coef[it];
        r.woff = m_woff[it];
        const float* wp = wblock + (unsigned)(r.woff + grp[0]);
        const char* q1 = featb + ((unsigned)rw.x + lane_byte);
        const char* q2 = featb + ((unsigned)rw.y + lane_byte);
        const char* q3 = featb + ((unsigned)rw.z + lane_byte);
        const char* q4 = featb + ((unsigned)rw.w + lane_byte);
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int dj = kConstChunks ? j * 32 * V : ch[j] - ch[0];          // channels
            r.wv[j] = __ldg(wp + (grp[j] - grp[0]));
            if constexpr (kPacked) {
                r.raw[0] = ldg_nc_v4_pinned(reinterpret_cast<const T*>(q1) + dj);
                r.raw[1] = ldg_nc_v4_pinned(reinterpret_cast<const T*>(q2) + dj);
                r.raw[2] = ldg_nc_v4_pinned(reinterpret_cast<const T*>(q3) + dj);
                r.raw[3] = ldg_nc_v4_pinned(reinterpret_cast<const T*>(q4) + dj);
            } else {
                VecIO<T, V>::load(reinterpret_cast<const T*>(q1) + dj, r.v[0][j]);
                VecIO<T, V>::load(reinterpret_cast<const T*>(q2) + dj, r.v[1][j]);
                VecIO<T, V>::load(reinterpret_cast<const T*>(q3) + dj, r.v[2][j]);
                VecIO<T, V>::load(reinterpret_cast<const T*>(q4) + dj, r.v[3][j]);
            }
        }
    };

    float gx = 0.f, gy = 0.f;   // kBwd: lane partials of the current pair's location gradient
    float* const gw_block = (kMode == kBwd) ? p.g_w + (size_t)ba * NP * L * d.G : nullptr;
    float* const gloc_row = (kMode == kBwd) ? p.g_loc + (size_t)ba * NP * 2 : nullptr;

    auto consume = [&](int q, int it, const Item& r) {
        if (kMode != kBwd) {
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                float vv[4][V];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if constexpr (kPacked) {
                        unpack_bf16x8(r.raw[k], vv[k]);
                    } else {
#pragma unroll
                        for (int e = 0; e < V; ++e) vv[k][e] = r.v[k][j][e];
                    }
                }
                float wj = r.wv[j];
                if (kMode == kFused) wj = expf(wj - sm_m[j]) * sm_inv[j];
                // weight folded into the four bilinear coefficients: 4 FMAs per channel,
                // issued as packed fp32x2 FMAs (FFMA2) where the vector width allows
                const float kk[4] = {r.cf.x * wj, r.cf.y * wj, r.cf.z * wj, r.cf.w * wj};
                if constexpr (V % 2 == 0) {
#pragma unroll
                    for (int e = 0; e < V; e += 2) {
                        float2 a2 = make_float2(acc[j][e], acc[j][e + 1]);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            a2 = __ffma2_rn(make_float2(kk[k], kk[k]), make_float2(vv[k][e], vv[k][e + 1]), a2);
                        acc[j][e] = a2.x;
                        acc[j][e + 1] = a2.y;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < V; ++e)
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[j][e] = __fmaf_rn(kk[k], vv[k][e], acc[j][e]);
                }
            }
        } else {
            const float4 cx4 = m_dx[it], cy4 = m_dy[it];
            const float cc[4] = {r.cf.x, r.cf.y, r.cf.z, r.cf.w};
            const float ax[4] = {cx4.x, cx4.y, cx4.z, cx4.w};
            const float by[4] = {cy4.x, cy4.y, cy4.z, cy4.w};
            float* gw_dst = gw_block + (unsigned)r.woff;
            float gwv[NCH];
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                // s_k = <grad_out, corner_k> over this lane's channels, then three 4-term forms
                float vv[4][V];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if constexpr (kPacked) {
                        unpack_bf16x8(r.raw[k], vv[k]);
                    } else {
#pragma unroll
                        for (int e = 0; e < V; ++e) vv[k][e] = r.v[k][j][e];
                    }
                }
                float sk[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if constexpr (V % 2 == 0) {
                        float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
                        for (int e = 0; e < V; e += 2)
                            s2 = __ffma2_rn(make_float2(go[j][e], go[j][e + 1]),
                                            make_float2(vv[k][e], vv[k][e + 1]), s2);
                        sk[k] = s2.x + s2.y;
                    } else {
                        float s1 = 0.f;
#pragma unroll
                        for (int e = 0; e < V; ++e) s1 = __fmaf_rn(go[j][e], vv[k][e], s1);
                        sk[k] = s1;
                    }
                }
                gwv[j] = cc[0] * sk[0] + cc[1] * sk[1] + cc[2] * sk[2] + cc[3] * sk[3];
                const float dxs = ax[0] * sk[0] + ax[1] * sk[1] + ax[2] * sk[2] + ax[3] * sk[3];
                const float dys = by[0] * sk[0] + by[1] * sk[1] + by[2] * sk[2] + by[3] * sk[3];
                gx = __fmaf_rn(dxs, r.wv[j], gx);
                gy = __fmaf_rn(dys, r.wv[j], gy);
            }
            // weight gradient: reduce over the lanes of each channel group, then one store per group
            if (kLPG >= 0) {
                if constexpr (NCH == 2) {
                    // two group sums per lane: the first butterfly step also splits them between the
                    // halves of the lane group (3 shuffles instead of 6 for 8-lane groups)
                    if (lpg >= 2) {
                        const int half = lpg >> 1;
                        const bool up = (lane & half) != 0;
                        const float send = up ? gwv[0] : gwv[1];
                        float keep = up ? gwv[1] : gwv[0];
                        keep += __shfl_xor_sync(0xffffffffu, send, half);
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1)
                            if (o < half) keep += __shfl_xor_sync(0xffffffffu, keep, o);
                        const int sub = lane & (lpg - 1);
                        if (sub == 0) gw_dst[grp[0]] = keep;
                        if (sub == half) gw_dst[grp[1]] = keep;
                    } else {
                        gw_dst[grp[0]] = gwv[0];
                        gw_dst[grp[1]] = gwv[1];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        float gw = gwv[j];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1)
                            if (o < lpg) gw += __shfl_xor_sync(0xffffffffu, gw, o);
                        if (act[j] && (lane & (lpg - 1)) == 0) gw_dst[grp[j]] = gw;
                    }
                }
            } else {
                // generic group sizes (V == 1): stage per-channel terms, one lane per group sums
                float* sc = red + warp * CPAD;
#pragma unroll
                for (int j = 0; j < NCH; ++j) sc[j * 32 + lane] = act[j] ? gwv[j] : 0.f;
                __syncwarp();
                for (int g = lane; g < d.G; g += 32) {
                    float s_ = 0.f;
                    for (int c = g * gd; c < (g + 1) * gd; ++c) s_ += sc[c];
                    gw_dst[g] = s_;
                }
                __syncwarp();
            }
            if (q % L == L - 1) {
                // last level of the pair: gx and gy reduced together (lower half-warp ends with gx,
                // upper with gy: 5 shuffles instead of 10)
                const bool up = (lane & 16) != 0;
                const float send = up ? gx : gy;
                float keep = up ? gy : gx;
                keep += __shfl_xor_sync(0xffffffffu, send, 16);
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
                const int pair = l_pair[it / L];
                if ((lane & 15) == 0) gloc_row[pair * 2 + (up ? 1 : 0)] = keep;
                gx = 0.f;
                gy = 0.f;
            }
        }
    };

    if constexpr (kPacked) {   // 4 items deep: 4 x 16 registers of packed rows in flight per lane
        Item A, B, C_, D_;
        if (my_items > 0) issue(item_of(0), A);
        if (my_items > 1) issue(item_of(1), B);
        if (my_items > 2) issue(item_of(2), C_);
        for (int q = 0; q < my_items; q += 4) {
            if (q + 3 < my_items) issue(item_of(q + 3), D_);
            consume(q, item_of(q), A);
            if (q + 4 < my_items) issue(item_of(q + 4), A);
            if (q + 1 < my_items) consume(q + 1, item_of(q + 1), B);
            if (q + 5 < my_items) issue(item_of(q + 5), B);
            if (q + 2 < my_items) consume(q + 2, item_of(q + 2), C_);
            if (q + 6 < my_items) issue(item_of(q + 6), C_);
            if (q + 3 < my_items) consume(q + 3, item_of(q + 3), D_);
        }
    } else if constexpr (NCH * V <= 8) {
        Item A, B;
        if (my_items > 0) issue(item_of(0), A);
        for (int q = 0; q < my_items; q += 2) {
            const int itA = item_of(q);
            const bool hasB = q + 1 < my_items;
            const int itB = hasB ? item_of(q + 1) : 0;
            if (hasB) issue(itB, B);
            consume(q, itA, A);
            if (q + 2 < my_items) issue(item_of(q + 2), A);
            if (hasB) consume(q + 1, itB, B);
        }
    } else {   // wide rows: one item's loads already fill the register budget
        Item A;
        for (int q = 0; q < my_items; ++q) {
            const int it = item_of(q);
            issue(it, A);
            consume(q, it, A);
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
