// dfa_group.cuh — grouped sample-major kernel: ONE launch for several aggregation calls that read the same
// feature maps (the det / map / plan / ego calls of a decoder layer, sparse_onedecoder.py:867-887), also used
// with a single call.  Specialised for the shipped HiP-AD layout: 16-byte vector rows (C = NCH*32*V), 4 levels,
// <= 8 groups; everything else stays on dfa_sample.cuh.
//
// Compared with dfa_sample_kernel (round 1):
//   * work unit = (call, output row, CONTIGUOUS range of <= 128 (p,cam) pairs), one CTA of kW warps each; a launch
//     has thousands of similar units (no cluster, no per-call tail), the slices of a row are combined through
//     partial rows + a ticket (the last CTA to arrive adds the partials in slice order: deterministic);
//   * the weights of an item (8 floats) are staged in shared memory by the metadata pass, so the gather loop issues
//     nothing but the 4 corner rows; forward items are dealt to warps one by one (balanced), not pair by pair;
//   * address arithmetic per corner is one 64-bit add of a 32-bit byte offset to a lane-resolved base.
#pragma once
#include "dfa_common.cuh"

namespace hipad {

constexpr int kMaxGroupCalls = 8;
constexpr int kGroupL = 4;          // levels (compile time)
constexpr int kGroupMaxG = 8;       // weights of one item held as 8 floats in shared memory
constexpr int kGroupMaxPS = 128;    // (p,cam) pairs per unit (one visibility pass of a 4-warp CTA)

struct GroupCall {
    const float* loc;        // [bs, A, P, cams, 2]
    const float* weights;    // [bs, A, P, cams, L, G]
    float* out;              // forward:  row (b, a) at out + b*io_bstride + a*C
    const float* grad_out;   // backward: same addressing
    float* g_loc;            // backward [bs, A, P, cams, 2]
    float* g_w;              // backward [bs, A, P, cams, L, G]
    long long io_bstride;    // elements between batch elements of out / grad_out (A*C, or A_total*C when packed)
    int A, P, S, PS;         // S slices per row of PS pairs each (the last one may be shorter)
    int unit_begin;          // first CTA of this call
    int part_begin;          // forward, S > 1: first partial-row slot of this call
    int row_begin;           // forward, S > 1: first ticket of this call
    int pad_;
};

struct GroupSmem {
    int tab, wcnt, flag, vis, lxy, lpair, mrows, mcoef, mw, mdx, mdy, red, total;
};
__host__ __device__ inline GroupSmem group_smem_layout(bool bwd, int ps_max, int cpad, int warps) {
    GroupSmem o;
    int b = 0;
    auto take = [&](int bytes) { const int at = b; b += (bytes + 15) & ~15; return at; };
    const int items = ps_max * kGroupL;
    o.tab = take(kMaxCamLevels * 3 * 4);
    o.wcnt = take(16 * 4);
    o.flag = take(16);
    o.vis = take(bwd ? ps_max : 0);
    o.lxy = take(ps_max * 8);
    o.lpair = take(ps_max * 4);
    o.mrows = take(items * 16);
    o.mcoef = take(items * 16);
    o.mw = take(items * kGroupMaxG * 4);
    o.mdx = take(bwd ? items * 16 : 0);
    o.mdy = take(bwd ? items * 16 : 0);
    o.red = take(bwd ? 0 : warps * cpad * 4);
    o.total = b;
    return o;
}

struct GroupParams {
    const void* feat;
    const int* shapes;
    const int* starts;
    float* partial;          // forward workspace: [sum over calls with S>1 of bs*A*S][C]
    int* tickets;            // forward workspace: [sum over calls with S>1 of bs*A], zeroed before the launch
    uint4* zero_ptr;         // backward: dense buffer this launch zero-fills on the side (or null)
    long long zero_n16;
    int ncalls, bs, cams, num_feat, C, G, ps_max, pad_;
    GroupSmem so;            // shared-memory carve-up (filled by the host: the kernel reads the offsets as constants)
    GroupCall calls[kMaxGroupCalls];
};

template <typename T, int V, int NCH, bool kBwd, int kW>
__global__ void __launch_bounds__(kW * 32) dfa_group_kernel(const GroupParams p) {
    constexpr int kThreads = kW * 32;
    constexpr int L = kGroupL;
    constexpr int CPAD = NCH * 32 * V;                       // == C (checked on the host)
    constexpr bool kPacked = (sizeof(T) == 2 && V == 8 && NCH == 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ------------------------------------------------------------------ which unit am I
    int c = 0;
#pragma unroll
    for (int k = 1; k < kMaxGroupCalls; ++k)
        if (k < p.ncalls && (int)blockIdx.x >= p.calls[k].unit_begin) c = k;
    const GroupCall& gc = p.calls[c];
    const int A = gc.A, S = gc.S, PS = gc.PS;
    const int NP = gc.P * p.cams;
    const int r = (int)blockIdx.x - gc.unit_begin;
    const int ba = r / S, slice = r - ba * S;
    const int b = ba / A, a = ba - b * A;
    const int p0 = slice * PS;
    const int n_mine = min(PS, NP - p0);
    const int G = p.G, gd = p.C / G;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GroupSmem& so = p.so;
    int* tab = reinterpret_cast<int*>(smem_raw + so.tab);
    int* s_wcnt = reinterpret_cast<int*>(smem_raw + so.wcnt);
    int* s_flag = reinterpret_cast<int*>(smem_raw + so.flag);
    unsigned char* s_vis = smem_raw + so.vis;
    float2* l_xy = reinterpret_cast<float2*>(smem_raw + so.lxy);
    int* l_pair = reinterpret_cast<int*>(smem_raw + so.lpair);
    uint4* m_rows = reinterpret_cast<uint4*>(smem_raw + so.mrows);
    float4* m_coef = reinterpret_cast<float4*>(smem_raw + so.mcoef);
    float* m_w = reinterpret_cast<float*>(smem_raw + so.mw);
    float4* m_dx = reinterpret_cast<float4*>(smem_raw + so.mdx);
    float4* m_dy = reinterpret_cast<float4*>(smem_raw + so.mdy);
    float* red = reinterpret_cast<float*>(smem_raw + so.red);

    if (kBwd && p.zero_n16 > 0) {
        // dense zero fill of the feature gradient, 1/gridDim of it per CTA: fire-and-forget stores that drain
        // while this CTA waits on its gather loads
        const long long per = (p.zero_n16 + gridDim.x - 1) / gridDim.x;
        const long long z0 = (long long)blockIdx.x * per;
        const long long z1 = (z0 + per < p.zero_n16) ? z0 + per : p.zero_n16;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (long long i = z0 + tid; i < z1; i += kThreads) p.zero_ptr[i] = z;
    }
    load_level_table(tab, p.shapes, p.starts, p.cams * L);

    // ------------------------------------------------------------------ phase 1: visible pairs (ordered compaction)
    const float2* loc2 = reinterpret_cast<const float2*>(gc.loc) + (size_t)ba * NP + p0;
    int n_list = 0;
    for (int base = 0; base < n_mine; base += kThreads) {       // one trip for PS <= kThreads
        const int k = base + tid;
        float2 xy = make_float2(-1.f, -1.f);
        if (k < n_mine) xy = __ldg(loc2 + k);
        const bool vis = (k < n_mine) && loc_valid(xy.x, xy.y);
        if (kBwd && k < n_mine) {
            s_vis[k] = vis ? 1 : 0;
            // gradients of an invisible pair are exactly zero (its weight-gradient rows are zeroed below)
            if (!vis) reinterpret_cast<float2*>(gc.g_loc)[(size_t)ba * NP + p0 + k] = make_float2(0.f, 0.f);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, vis);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < kW; ++w) {
            const int cnt = s_wcnt[w];
            if (w < warp) before += cnt;
            all += cnt;
        }
        if (vis) {
            const int slot = n_list + before + __popc(bal & ((1u << lane) - 1u));
            l_xy[slot] = xy;
            l_pair[slot] = k;
        }
        n_list += all;
        __syncthreads();
    }

    if (kBwd) {
        // zero the weight-gradient rows of invisible pairs (L*G contiguous floats each), all threads, coalesced
        const int lg = L * G;
        float* gw_base = gc.g_w + ((size_t)ba * NP + p0) * lg;
        if ((lg & 3) == 0) {
            const int q_per = lg >> 2;
            for (int q = tid; q < n_mine * q_per; q += kThreads)
                if (!s_vis[q / q_per]) reinterpret_cast<float4*>(gw_base)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (int q = tid; q < n_mine * lg; q += kThreads)
                if (!s_vis[q / lg]) gw_base[q] = 0.f;
        }
    }

    // ------------------------------------------------------------------ phase 2: gather metadata, one thread per item
    const int n_items = n_list * L;
    const unsigned row_bytes = (unsigned)p.C * (unsigned)sizeof(T);
    for (int it = tid; it < n_items; it += kThreads) {
        const int i = it >> 2, l = it & 3;
        const float2 xy = l_xy[i];
        const int pair = p0 + l_pair[i];
        const int pt = pair / p.cams, cam = pair - pt * p.cams;
        const int* t = tab + (cam * L + l) * 3;
        const int h = t[0], w = t[1];
        const Quad q = quad_setup(xy.x, xy.y, h, w);
        const int r1 = t[2] + q.h_low * w + q.w_low, r2 = r1 + 1, r3 = r1 + w, r4 = r3 + 1;
        // out-of-bounds corners are redirected to an in-bounds corner of the same quad with coefficient 0
        const int safe = q.ok1 ? r1 : (q.ok2 ? r2 : (q.ok3 ? r3 : r4));
        m_rows[it] = make_uint4((unsigned)(q.ok1 ? r1 : safe) * row_bytes, (unsigned)(q.ok2 ? r2 : safe) * row_bytes,
                                (unsigned)(q.ok3 ? r3 : safe) * row_bytes, (unsigned)(q.ok4 ? r4 : safe) * row_bytes);
        m_coef[it] = make_float4(q.ok1 ? q.hh * q.hw : 0.f, q.ok2 ? q.hh * q.lw : 0.f,
                                 q.ok3 ? q.lh * q.hw : 0.f, q.ok4 ? q.lh * q.lw : 0.f);
        const float* wsrc = gc.weights + (((size_t)ba * NP + pair) * L + l) * G;
        float* wdst = m_w + it * kGroupMaxG;
        if (G == 8) {
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(wsrc));
            const float4 w1 = __ldg(reinterpret_cast<const float4*>(wsrc) + 1);
            reinterpret_cast<float4*>(wdst)[0] = w0;
            reinterpret_cast<float4*>(wdst)[1] = w1;
        } else {
            for (int g = 0; g < G; ++g) wdst[g] = __ldg(wsrc + g);
        }
        if (kBwd) {
            // d(val)/d(loc_x) = W * (-hh v1 + hh v2 - lh v3 + lh v4), d/d(loc_y) = H * (-hw v1 - lw v2 + hw v3 + lw v4)
            const float W_ = (float)w, H_ = (float)h;
            m_dx[it] = make_float4(q.ok1 ? -q.hh * W_ : 0.f, q.ok2 ? q.hh * W_ : 0.f,
                                   q.ok3 ? -q.lh * W_ : 0.f, q.ok4 ? q.lh * W_ : 0.f);
            m_dy[it] = make_float4(q.ok1 ? -q.hw * H_ : 0.f, q.ok2 ? -q.lw * H_ : 0.f,
                                   q.ok3 ? q.hw * H_ : 0.f, q.ok4 ? q.lw * H_ : 0.f);
        }
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 3: gather
    // lane owns V consecutive channels in each of NCH chunks of 32*V channels
    int grp[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) grp[j] = ((j * 32 + lane) * V) / gd;
    const int lpg = (gd / V) > 0 ? (gd / V) : 1;          // lanes per channel group (power of two, checked on the host)
    const char* fl = reinterpret_cast<const char*>(p.feat) + ((size_t)b * p.num_feat * p.C + (size_t)lane * V) * sizeof(T);
    asm volatile("" : "+l"(fl));      // keep the lane's base in registers (nvcc otherwise re-derives it per item)

    struct Item {
        float v[4][kPacked ? 1 : NCH][kPacked ? 1 : V];
        uint4 raw[kPacked ? 4 : 1];
    };
    auto issue = [&](int it, Item& r_) {
        const uint4 rw = m_rows[it];
        const char* q1 = fl + rw.x;
        const char* q2 = fl + rw.y;
        const char* q3 = fl + rw.z;
        const char* q4 = fl + rw.w;
        if constexpr (kPacked) {
            r_.raw[0] = ldg_nc_v4_pinned(q1);
            r_.raw[1] = ldg_nc_v4_pinned(q2);
            r_.raw[2] = ldg_nc_v4_pinned(q3);
            r_.raw[3] = ldg_nc_v4_pinned(q4);
        } else {
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                VecIO<T, V>::load(reinterpret_cast<const T*>(q1) + j * 32 * V, r_.v[0][j]);
                VecIO<T, V>::load(reinterpret_cast<const T*>(q2) + j * 32 * V, r_.v[1][j]);
                VecIO<T, V>::load(reinterpret_cast<const T*>(q3) + j * 32 * V, r_.v[2][j]);
                VecIO<T, V>::load(reinterpret_cast<const T*>(q4) + j * 32 * V, r_.v[3][j]);
            }
        }
    };
    auto corner = [&](const Item& r_, int k, int j, float (&vv)[V]) {
        if constexpr (kPacked) {
            unpack_bf16x8(r_.raw[k], vv);
        } else {
#pragma unroll
            for (int e = 0; e < V; ++e) vv[e] = r_.v[k][j][e];
        }
    };

    if constexpr (!kBwd) {
        float acc[NCH][V];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int e = 0; e < V; ++e) acc[j][e] = 0.f;

        auto consume = [&](int it, const Item& r_) {
            const float4 cf = m_coef[it];
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                const float wj = m_w[it * kGroupMaxG + grp[j]];
                const float kk[4] = {cf.x * wj, cf.y * wj, cf.z * wj, cf.w * wj};
                float vv[4][V];
#pragma unroll
                for (int k = 0; k < 4; ++k) corner(r_, k, j, vv[k]);
#pragma unroll
                for (int e = 0; e < V; e += 2) {
                    float2 a2 = make_float2(acc[j][e], acc[j][e + 1]);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        a2 = __ffma2_rn(make_float2(kk[k], kk[k]), make_float2(vv[k][e], vv[k][e + 1]), a2);
                    acc[j][e] = a2.x;
                    acc[j][e + 1] = a2.y;
                }
            }
        };
        // items are dealt to the warps one by one; kDepth items in flight per warp
        if constexpr (kPacked) {
            Item I0, I1, I2, I3;
            int it = warp;
            if (it < n_items) issue(it, I0);
            if (it + kW < n_items) issue(it + kW, I1);
            if (it + 2 * kW < n_items) issue(it + 2 * kW, I2);
            for (; it < n_items; it += 4 * kW) {
                if (it + 3 * kW < n_items) issue(it + 3 * kW, I3);
                consume(it, I0);
                if (it + 4 * kW < n_items) issue(it + 4 * kW, I0);
                if (it + kW < n_items) consume(it + kW, I1);
                if (it + 5 * kW < n_items) issue(it + 5 * kW, I1);
                if (it + 2 * kW < n_items) consume(it + 2 * kW, I2);
                if (it + 6 * kW < n_items) issue(it + 6 * kW, I2);
                if (it + 3 * kW < n_items) consume(it + 3 * kW, I3);
            }
        } else {
            Item I0, I1;
            int it = warp;
            if (it < n_items) issue(it, I0);
            for (; it < n_items; it += 2 * kW) {
                const bool has1 = it + kW < n_items;
                if (has1) issue(it + kW, I1);
                consume(it, I0);
                if (it + 2 * kW < n_items) issue(it + 2 * kW, I0);
                if (has1) consume(it + kW, I1);
            }
        }

        // -------------------------------------------------------------- phase 4: cross-warp sum in warp order
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int e = 0; e < V; ++e) red[warp * CPAD + (j * 32 + lane) * V + e] = acc[j][e];
        __syncthreads();
        float* out_row = gc.out + (size_t)b * gc.io_bstride + (size_t)a * p.C;
        if (S == 1) {
            for (int ch = tid; ch < CPAD; ch += kThreads) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < kW; ++w) s += red[w * CPAD + ch];
                out_row[ch] = s;
            }
        } else {
            // partial row of this slice; the last slice of the row to arrive adds all S partials in slice order
            float* mine = p.partial + ((size_t)gc.part_begin + (size_t)ba * S + slice) * CPAD;
            for (int ch = tid; ch < CPAD; ch += kThreads) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < kW; ++w) s += red[w * CPAD + ch];
                mine[ch] = s;
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) s_flag[0] = (atomicAdd(p.tickets + gc.row_begin + ba, 1) == S - 1) ? 1 : 0;
            __syncthreads();
            if (s_flag[0]) {
                __threadfence();
                const float* parts = p.partial + ((size_t)gc.part_begin + (size_t)ba * S) * CPAD;
                for (int ch = tid; ch < CPAD; ch += kThreads) {
                    float s = 0.f;
                    for (int q = 0; q < S; ++q) s += __ldcg(parts + (size_t)q * CPAD + ch);
                    out_row[ch] = s;
                }
            }
        }
    } else {
        // -------------------------------------------------------------- backward: g_w per item, g_loc per pair
        float go[NCH][V];
        {
            const float* go_row = gc.grad_out + (size_t)b * gc.io_bstride + (size_t)a * p.C;
#pragma unroll
            for (int j = 0; j < NCH; ++j) VecIO<float, V>::load(go_row + (j * 32 + lane) * V, go[j]);
        }
        float* const gw_block = gc.g_w + (size_t)ba * NP * (L * G);
        float2* const gloc_row = reinterpret_cast<float2*>(gc.g_loc) + (size_t)ba * NP;
        float gx = 0.f, gy = 0.f;

        auto consume = [&](int it, const Item& r_) {
            const float4 cf = m_coef[it], cx4 = m_dx[it], cy4 = m_dy[it];
            const float cc[4] = {cf.x, cf.y, cf.z, cf.w};
            const float ax[4] = {cx4.x, cx4.y, cx4.z, cx4.w};
            const float by[4] = {cy4.x, cy4.y, cy4.z, cy4.w};
            const int pair = p0 + l_pair[it >> 2];
            float* gw_dst = gw_block + ((size_t)pair * L + (it & 3)) * G;
            float gwv[NCH];
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                float sk[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float vv[V];
                    corner(r_, k, j, vv);
                    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int e = 0; e < V; e += 2)
                        s2 = __ffma2_rn(make_float2(go[j][e], go[j][e + 1]), make_float2(vv[e], vv[e + 1]), s2);
                    sk[k] = s2.x + s2.y;
                }
                const float wj = m_w[it * kGroupMaxG + grp[j]];
                gwv[j] = cc[0] * sk[0] + cc[1] * sk[1] + cc[2] * sk[2] + cc[3] * sk[3];
                const float dxs = ax[0] * sk[0] + ax[1] * sk[1] + ax[2] * sk[2] + ax[3] * sk[3];
                const float dys = by[0] * sk[0] + by[1] * sk[1] + by[2] * sk[2] + by[3] * sk[3];
                gx = __fmaf_rn(dxs, wj, gx);
                gy = __fmaf_rn(dys, wj, gy);
            }
            // weight gradient: sum over the lpg lanes of each channel group, one store per group
            if constexpr (NCH == 2) {
                if (lpg >= 2) {
                    // two group sums per lane: the first butterfly step also splits them between the halves
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
                    if ((lane & (lpg - 1)) == 0) gw_dst[grp[j]] = gw;
                }
            }
            if ((it & 3) == L - 1) {
                // last level of the pair: gx and gy reduced together (lower half-warp ends with gx, upper with gy)
                const bool up = (lane & 16) != 0;
                const float send = up ? gx : gy;
                float keep = up ? gy : gx;
                keep += __shfl_xor_sync(0xffffffffu, send, 16);
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
                if ((lane & 15) == 0) reinterpret_cast<float*>(gloc_row + pair)[up ? 1 : 0] = keep;
                gx = 0.f;
                gy = 0.f;
            }
        };
        // warp `warp` owns visible pairs warp, warp+kW, ... and walks their 4 levels back to back
        const int my_pairs = (n_list > warp) ? (n_list - warp + kW - 1) / kW : 0;
        const int my_items = my_pairs * L;
        auto item_of = [&](int q) { return ((warp + (q >> 2) * kW) << 2) + (q & 3); };
        if constexpr (kPacked) {
            Item I0, I1, I2, I3;
            if (my_items > 0) issue(item_of(0), I0);
            if (my_items > 1) issue(item_of(1), I1);
            if (my_items > 2) issue(item_of(2), I2);
            for (int q = 0; q < my_items; q += 4) {      // my_items is a multiple of 4
                issue(item_of(q + 3), I3);
                consume(item_of(q), I0);
                if (q + 4 < my_items) issue(item_of(q + 4), I0);
                consume(item_of(q + 1), I1);
                if (q + 5 < my_items) issue(item_of(q + 5), I1);
                consume(item_of(q + 2), I2);
                if (q + 6 < my_items) issue(item_of(q + 6), I2);
                consume(item_of(q + 3), I3);
            }
        } else {
            Item I0, I1;
            if (my_items > 0) issue(item_of(0), I0);
            for (int q = 0; q < my_items; q += 2) {      // my_items is a multiple of 4
                issue(item_of(q + 1), I1);
                consume(item_of(q), I0);
                if (q + 2 < my_items) issue(item_of(q + 2), I0);
                consume(item_of(q + 1), I1);
            }
        }
    }
}

}  // namespace hipad
