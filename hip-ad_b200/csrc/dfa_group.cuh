// dfa_group.cuh — grouped sample-major kernel: ONE launch for several aggregation calls that read the same
// feature maps (the det / map / plan / ego calls of a decoder layer, sparse_onedecoder.py:867-887), also used
// with a single call.  Specialised for the shipped HiP-AD layout: 16-byte vector rows (C = NCH*32*V), 4 levels,
// <= 8 groups, >= 4 lanes per channel group; everything else stays on dfa_sample.cuh.
//
// Work unit = (call, output row, CONTIGUOUS range of <= 96 (p,cam) pairs), one CTA of kW warps; a launch has
// thousands of similar units (no cluster, no per-call tail).  Inside a unit:
//   1. visibility test + ordered compaction of the unit's pairs (10-20 % are visible);
//   2. every (visible pair, level) ITEM gets a key (camera-level, quad row, quad column); a rank sort in shared
//      memory brings items that read the SAME quad of the same map together — the key points of one anchor fall
//      into a handful of quads on the coarse levels, so a unit has 1.5-3x fewer distinct quads than items;
//   3. forward: per quad the bilinear coefficients x group weights of its items are summed into one
//      [group][corner] table (weights read once, coalesced); the gather loop then reads each distinct quad ONCE
//      (4 corner rows, LDG.128, two quads in flight per warp) and applies the merged table;
//      backward: the gather loop reduces <grad_out, corner_k> per (quad, group) once (transposing butterfly, 4
//      shuffles per 128 channels), and a thread-per-item epilogue turns those 32 numbers into the item's weight
//      gradients and location-gradient terms;
//   4. slices of a row are combined through partial rows + a ticket: the last CTA to arrive adds the partials in
//      slice order (deterministic, no floating-point atomics anywhere).
// Every floating-point sum has a fixed order: items of a quad in (pair, level) order, quads of a warp in list order,
// warps in index order, slices in slice order.
#pragma once
#include "dfa_common.cuh"

namespace hipad {

constexpr int kMaxGroupCalls = 8;
constexpr int kGroupL = 4;          // levels (compile time)
constexpr int kGroupMaxG = 8;       // groups held per quad table
constexpr int kGroupMaxPS = 96;     // (p,cam) pairs per unit
constexpr int kQuadChunk = 64;      // distinct quads whose tables are resident at a time

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
    int tab, wcnt, flag, vis, lxy, lpair, key, mask, lhlw, woff, sorted, qof, qstart, gxy, qrows, qtab, total;
};
__host__ __device__ inline GroupSmem group_smem_layout(bool bwd, int ps_max, int cpad, int warps) {
    GroupSmem o;
    int b = 0;
    auto take = [&](int bytes) { const int at = b; b += (bytes + 15) & ~15; return at; };
    const int items = ((ps_max + 3) & ~3) * kGroupL;       // level-major, every level padded to 4 entries
    o.tab = take(kMaxCamLevels * 3 * 4);
    o.wcnt = take(16 * 4);
    o.flag = take(16);
    o.vis = take(bwd ? ps_max : 0);
    o.lxy = take(ps_max * 8);
    o.lpair = take(ps_max * 2);
    o.key = take(items * 4);
    o.mask = take(items);
    o.lhlw = take(items * (bwd ? 8 : 16));   // backward: (lh, lw) per item; forward: its four masked bilinear terms
    o.woff = take(items * 2);
    o.sorted = take(items * 2);
    o.qof = take(items * 2);
    o.qstart = take((items + 1) * 2);
    o.gxy = take(bwd ? items * 8 : 0);
    o.qrows = take(kQuadChunk * 16);
    // quad tables [quad][group] float4; the forward's cross-warp reduction scratch reuses the same bytes
    int qtab = kQuadChunk * kGroupMaxG * 16;
    if (!bwd && warps * cpad * 4 > qtab) qtab = warps * cpad * 4;
    o.qtab = take(qtab);
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
    long long zero_per;      // 16-byte units per CTA: ceil(zero_n16 / grid), < 2^31 (filled by launch_group_sample)
    int ncalls, bs, cams, num_feat, C, G, ps_max;
    int units;               // CTAs [units, gridDim.x) exist only for the zero fill (launches with few units)
    GroupSmem so;            // shared-memory carve-up (filled by the host: the kernel reads the offsets as constants)
    GroupCall calls[kMaxGroupCalls];
};

struct ItemGeo {
    float lh, lw, hh, hw;
};

// kDepth: distinct quads whose rows are in flight per warp (registers: kDepth * 32 for fp32 C = 256);
// kMinCtas: CTAs per SM the register allocation must allow
// kG: the number of channel groups when known at compile time (0 = read it from the parameters): with G fixed, the lanes
// per group, the butterfly's shuffle distances and the weight-row length fold into constants (the backward kernel spends
// a quarter of its instructions on that butterfly)
template <typename T, int V, int NCH, bool kBwd, int kW, int kDepth, int kMinCtas, int kG = 0>
__global__ void __launch_bounds__(kW * 32, kMinCtas) dfa_group_kernel(const GroupParams p) {
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
    const int cams = p.cams;
    const int NP = gc.P * cams;
    const int r_unit = (int)blockIdx.x - gc.unit_begin;
    const int ba = r_unit / S, slice = r_unit - ba * S;
    const int b = ba / A, a = ba - b * A;
    const int p0 = slice * PS;
    const int n_mine = min(PS, NP - p0);
    const int G = (kG > 0) ? kG : p.G;
    const int gd = (NCH * 32 * V) / G;                     // C == NCH*32*V (checked on the host)
    DFA_TRACE_BEGIN((int)blockIdx.x);
    DFA_TRACE((int)blockIdx.x, 2);
    DFA_TRACE_DECL(tr5);
    DFA_TRACE_DECL(tr6);
    DFA_TRACE_DECL(tr7);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GroupSmem& so = p.so;
    int* tab = reinterpret_cast<int*>(smem_raw + so.tab);
    int* s_wcnt = reinterpret_cast<int*>(smem_raw + so.wcnt);
    int* s_flag = reinterpret_cast<int*>(smem_raw + so.flag);
    unsigned char* s_vis = smem_raw + so.vis;
    float2* l_xy = reinterpret_cast<float2*>(smem_raw + so.lxy);
    unsigned short* l_pair = reinterpret_cast<unsigned short*>(smem_raw + so.lpair);
    unsigned* s_key = reinterpret_cast<unsigned*>(smem_raw + so.key);
    unsigned char* s_mask = smem_raw + so.mask;
    float2* s_lhlw = reinterpret_cast<float2*>(smem_raw + so.lhlw);
    float4* s_c4 = reinterpret_cast<float4*>(smem_raw + so.lhlw);
    unsigned short* s_woff = reinterpret_cast<unsigned short*>(smem_raw + so.woff);
    unsigned short* s_sorted = reinterpret_cast<unsigned short*>(smem_raw + so.sorted);
    unsigned short* s_qof = reinterpret_cast<unsigned short*>(smem_raw + so.qof);
    unsigned short* s_qstart = reinterpret_cast<unsigned short*>(smem_raw + so.qstart);
    float2* s_gxy = reinterpret_cast<float2*>(smem_raw + so.gxy);
    uint4* q_rows = reinterpret_cast<uint4*>(smem_raw + so.qrows);
    float4* q_tab = reinterpret_cast<float4*>(smem_raw + so.qtab);
    float* red = reinterpret_cast<float*>(smem_raw + so.qtab);

    if (kBwd && p.zero_n16 > 0) {
        // dense zero fill of the feature gradient, 1/gridDim of it per CTA: fire-and-forget stores that drain
        // while this CTA waits on its gather loads
        const long long z0 = (long long)blockIdx.x * p.zero_per;          // zero_per = ceil(zero_n16 / gridDim.x), host
        const long long left = p.zero_n16 - z0;
        const int cnt = (int)(left < p.zero_per ? (left > 0 ? left : 0) : p.zero_per);
        uint4* zp = p.zero_ptr + z0;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        int i = tid;
        for (; i + 3 * kThreads < cnt; i += 4 * kThreads) {
            zp[i] = z;
            zp[i + kThreads] = z;
            zp[i + 2 * kThreads] = z;
            zp[i + 3 * kThreads] = z;
        }
        for (; i < cnt; i += kThreads) zp[i] = z;
        if ((int)blockIdx.x >= p.units) return;      // a CTA added to the launch for the fill alone (block-uniform)
    }
    if (tid == 0) {
        // The weights of the unit's pairs (L*G floats each, contiguous) are needed a few microseconds from now (phase 5
        // forward, phase 7 backward) and come from DRAM: one bulk prefetch into L2 now takes that latency off the unit's
        // critical path (cp.async.bulk.prefetch: no shared memory, no barrier to wait on).
        const float* w0 = gc.weights + ((size_t)ba * NP + p0) * (L * G);
        const unsigned bytes = (unsigned)n_mine * (unsigned)(L * G) * 4u;
        if ((bytes & 15u) == 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(w0), "r"(bytes) : "memory");
    }
    load_level_table(tab, p.shapes, p.starts, cams * L);

    // ------------------------------------------------------------------ phase 1: visible pairs (ordered compaction)
    const float2* loc2 = reinterpret_cast<const float2*>(gc.loc) + (size_t)ba * NP + p0;
    int n_list = 0;
    for (int base = 0; base < n_mine; base += kThreads) {       // one trip (PS <= 96 <= kThreads)
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
            l_pair[slot] = (unsigned short)k;
        }
        n_list += all;
        __syncthreads();
    }

    DFA_TRACE((int)blockIdx.x, 3);
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

    // ------------------------------------------------------------------ phase 2: item keys
    // An ITEM is (visible pair i, level l), stored LEVEL-MAJOR at position l*n_pad + i (n_pad = n_list rounded up to 4,
    // the padding holds keys that sort last): items of different levels never share a quad, so the sort below only
    // looks at one level's segment.  key = camera (6 bits) | quad row + 1 (13) | quad column + 1 (13): level maps up to
    // 8190 x 8190, at most 64 cameras.
    const int n_pad = (n_list + 3) & ~3;
    const int n_items = n_list * L;
    for (int idx = tid; idx < n_pad * L; idx += kThreads) {
        const int i = idx >> 2, l = idx & 3;
        unsigned key = 0xffffffffu;
        if (i < n_list) {
            const float2 xy = l_xy[i];
            const int pair = p0 + l_pair[i];
            const int cam = pair % cams;
            const int* t = tab + (cam * L + l) * 3;
            const Quad q = quad_setup(xy.x, xy.y, t[0], t[1]);
            key = ((unsigned)cam << 26) | ((unsigned)(q.h_low + 1) << 13) | (unsigned)(q.w_low + 1);
            s_mask[l * n_pad + i] = (unsigned char)((int)q.ok1 | ((int)q.ok2 << 1) | ((int)q.ok3 << 2) | ((int)q.ok4 << 3));
            if constexpr (kBwd) {
                s_lhlw[l * n_pad + i] = make_float2(q.lh, q.lw);
            } else {
                s_c4[l * n_pad + i] = make_float4(q.ok1 ? q.hh * q.hw : 0.f, q.ok2 ? q.hh * q.lw : 0.f,
                                                  q.ok3 ? q.lh * q.hw : 0.f, q.ok4 ? q.lh * q.lw : 0.f);
            }
            s_woff[l * n_pad + i] = (unsigned short)(((int)l_pair[i] * L + l) * G);     // < 96 * 4 * 8
        }
        s_key[l * n_pad + i] = key;
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 3: rank sort inside each level
    // (stable: ties by position).  Rank r of level l goes to slot l*n_list + r of the sorted list.
    for (int idx = tid; idx < n_items; idx += kThreads) {
        const int i = idx >> 2, l = idx & 3;
        const unsigned* seg = s_key + l * n_pad;
        const unsigned k = seg[i];
        int rank = 0;
        for (int j0 = 0; j0 < n_pad; j0 += 4) {
            const uint4 kk = *reinterpret_cast<const uint4*>(seg + j0);
            rank += (kk.x < k || (kk.x == k && j0 + 0 < i)) ? 1 : 0;
            rank += (kk.y < k || (kk.y == k && j0 + 1 < i)) ? 1 : 0;
            rank += (kk.z < k || (kk.z == k && j0 + 2 < i)) ? 1 : 0;
            rank += (kk.w < k || (kk.w == k && j0 + 3 < i)) ? 1 : 0;
        }
        s_sorted[l * n_list + rank] = (unsigned short)(l * n_pad + i);
    }
    __syncthreads();

    // ------------------------------------------------------------------ phase 4: distinct quads
    // a quad starts wherever the sorted key changes; quad index of an item = number of starts up to its rank - 1
    int n_quads = 0;
    for (int base = 0; base < n_items; base += kThreads) {
        const int r = base + tid;
        bool head = false;
        int it = 0;
        if (r < n_items) {
            it = s_sorted[r];
            // a quad starts where the key changes or a new level's segment begins (n_list ranks per level)
            const int lv = (r >= n_list) + (r >= 2 * n_list) + (r >= 3 * n_list);
            head = (r == lv * n_list) || (s_key[it] != s_key[s_sorted[r - 1]]);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, head);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < kW; ++w) {
            const int cnt = s_wcnt[w];
            if (w < warp) before += cnt;
            all += cnt;
        }
        if (r < n_items) {
            const int q = n_quads + before + __popc(bal & ((2u << lane) - 1u)) - 1;     // inclusive count - 1
            s_qof[it] = (unsigned short)q;
            if (head) s_qstart[q] = (unsigned short)r;
        }
        n_quads += all;
        __syncthreads();
    }
    if (tid == 0) s_qstart[n_quads] = (unsigned short)n_items;
    __syncthreads();
    DFA_TRACE((int)blockIdx.x, 4);

    // position of an item in the level-major arrays -> its level
    auto level_of = [&](int pos) { return (pos >= n_pad) + (pos >= 2 * n_pad) + (pos >= 3 * n_pad); };

    // lane owns V consecutive channels in each of NCH chunks of 32*V channels
    int grp[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) grp[j] = ((j * 32 + lane) * V) / gd;
    const int lpg = gd / V;                               // lanes per channel group (power of two >= 4, host-checked)
    const char* fl = reinterpret_cast<const char*>(p.feat) + ((size_t)b * p.num_feat * p.C + (size_t)lane * V) * sizeof(T);
    asm volatile("" : "+l"(fl));      // keep the lane's base in registers (nvcc otherwise re-derives it per quad)
    const unsigned row_bytes = (unsigned)p.C * (unsigned)sizeof(T);
    const float* w_unit = gc.weights + ((size_t)ba * NP + p0) * (L * G);     // weights of the unit's first pair

    struct Rows {
        float v[4][kPacked ? 1 : NCH][kPacked ? 1 : V];
        uint4 raw[kPacked ? 4 : 1];
    };
    auto issue = [&](int qi, Rows& r_) {
        const uint4 rw = q_rows[qi];
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
    auto corner = [&](const Rows& r_, int k, int j, float (&vv)[V]) {
        if constexpr (kPacked) {
            unpack_bf16x8(r_.raw[k], vv);
        } else {
#pragma unroll
            for (int e = 0; e < V; ++e) vv[e] = r_.v[k][j][e];
        }
    };

    float acc[NCH][V];      // forward accumulators
    float go[NCH][V];       // backward: grad_out of this row at the lane's channels
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
        for (int e = 0; e < V; ++e) { acc[j][e] = 0.f; go[j][e] = 0.f; }
    if (kBwd) {
        const float* go_row = gc.grad_out + (size_t)b * gc.io_bstride + (size_t)a * p.C;
#pragma unroll
        for (int j = 0; j < NCH; ++j) VecIO<float, V>::load(go_row + (j * 32 + lane) * V, go[j]);
    }

    for (int qc0 = 0; qc0 < n_quads; qc0 += kQuadChunk) {
        const int nq = min(kQuadChunk, n_quads - qc0);
        DFA_TRACE_NOW(t5);

        // -------------------------------------------------------------- phase 5: per-quad rows (+ forward tables)
        for (int qi = tid; qi < nq; qi += kThreads) {
            const int it = s_sorted[s_qstart[qc0 + qi]];         // first member: all members share the quad
            const unsigned key = s_key[it];
            const int cl = (int)(key >> 26) * L + level_of(it);
            const int hl = (int)((key >> 13) & 8191u) - 1, wl = (int)(key & 8191u) - 1;
            const int* t = tab + cl * 3;
            const int w = t[1];
            const int m = s_mask[it];
            const int r1 = t[2] + hl * w + wl, r2 = r1 + 1, r3 = r1 + w, r4 = r3 + 1;
            // out-of-bounds corners are redirected to an in-bounds corner of the same quad (their terms are 0)
            const int safe = (m & 1) ? r1 : ((m & 2) ? r2 : ((m & 4) ? r3 : r4));
            q_rows[qi] = make_uint4((unsigned)((m & 1) ? r1 : safe) * row_bytes, (unsigned)((m & 2) ? r2 : safe) * row_bytes,
                                    (unsigned)((m & 4) ? r3 : safe) * row_bytes, (unsigned)((m & 8) ? r4 : safe) * row_bytes);
        }
        if constexpr (!kBwd) {
            // table[quad][group] = sum over the quad's items of (bilinear term of corner k) * weight[group], k = 1..4
            const int lg2 = 31 - __clz(G);                        // G is a power of two (host-checked)
            for (int task = tid; task < (nq << lg2); task += kThreads) {
                const int qi = task >> lg2, g = task & (G - 1);
                const int r0 = s_qstart[qc0 + qi], r1 = s_qstart[qc0 + qi + 1];
                float4 tsum = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int r = r0; r < r1; ++r) {
                    const int it = s_sorted[r];
                    const float4 c4 = s_c4[it];
                    const float wv = __ldg(w_unit + (s_woff[it] + g));
                    tsum.x = __fmaf_rn(c4.x, wv, tsum.x);
                    tsum.y = __fmaf_rn(c4.y, wv, tsum.y);
                    tsum.z = __fmaf_rn(c4.z, wv, tsum.z);
                    tsum.w = __fmaf_rn(c4.w, wv, tsum.w);
                }
                q_tab[qi * kGroupMaxG + g] = tsum;
            }
        }
        __syncthreads();
        DFA_TRACE_ACC(tr5, t5);
        DFA_TRACE_NOW(t6);

        // -------------------------------------------------------------- phase 6: gather, one distinct quad at a time
        if constexpr (!kBwd) {
            auto consume = [&](int qi, const Rows& r_) {
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const float4 kk4 = q_tab[qi * kGroupMaxG + grp[j]];
                    const float kk[4] = {kk4.x, kk4.y, kk4.z, kk4.w};
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
            {
                // quads are dealt to the warps one by one; kDepth quads in flight per warp (rotating register sets)
                Rows I[kDepth];
#pragma unroll
                for (int d = 0; d + 1 < kDepth; ++d)
                    if (warp + d * kW < nq) issue(warp + d * kW, I[d]);
                for (int qi = warp; qi < nq; qi += kDepth * kW) {
#pragma unroll
                    for (int d = 0; d < kDepth; ++d) {
                        const int qn = qi + (d + kDepth - 1) * kW;
                        if (qn < nq) issue(qn, I[(d + kDepth - 1) % kDepth]);
                        if (qi + d * kW < nq) consume(qi + d * kW, I[d]);
                    }
                }
            }
        } else {
            // S[quad][group][k] = <grad_out, corner_k> over the channels of the group: per lane 4 partial dot products
            // per chunk, reduced over the lpg lanes of the group by a transposing butterfly (each step halves the
            // values a lane carries), the surviving lanes write one float each
            auto consume = [&](int qi, const Rows& r_) {
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
                    const int hA = lpg >> 1, hB = lpg >> 2;
                    const bool upA = (lane & hA) != 0, upB = (lane & hB) != 0;
                    // step A: lanes with bit hA clear keep corners {0,1}, the others {2,3}
                    float a0 = upA ? sk[2] : sk[0], a1 = upA ? sk[3] : sk[1];
                    const float s0 = upA ? sk[0] : sk[2], s1 = upA ? sk[1] : sk[3];
                    a0 += __shfl_xor_sync(0xffffffffu, s0, hA);
                    a1 += __shfl_xor_sync(0xffffffffu, s1, hA);
                    // step B: bit hB clear keeps the first of the pair
                    float v = upB ? a1 : a0;
                    const float sb = upB ? a0 : a1;
                    v += __shfl_xor_sync(0xffffffffu, sb, hB);
                    for (int o = hB >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if ((lane & (hB - 1)) == 0) {
                        const int k = (upA ? 2 : 0) + (upB ? 1 : 0);
                        reinterpret_cast<float*>(q_tab + qi * kGroupMaxG + grp[j])[k] = v;
                    }
                }
            };
            {
                // quads are dealt to the warps one by one; kDepth quads in flight per warp (rotating register sets)
                Rows I[kDepth];
#pragma unroll
                for (int d = 0; d + 1 < kDepth; ++d)
                    if (warp + d * kW < nq) issue(warp + d * kW, I[d]);
                for (int qi = warp; qi < nq; qi += kDepth * kW) {
#pragma unroll
                    for (int d = 0; d < kDepth; ++d) {
                        const int qn = qi + (d + kDepth - 1) * kW;
                        if (qn < nq) issue(qn, I[(d + kDepth - 1) % kDepth]);
                        if (qi + d * kW < nq) consume(qi + d * kW, I[d]);
                    }
                }
            }
            __syncthreads();
            DFA_TRACE_ACC(tr6, t6);
            DFA_TRACE_NOW(t7);

            // ---------------------------------------------------------- phase 7 (backward): one thread per item
            // g_w[item][g] = sum_k c_k S[g][k];  location-gradient terms of the item = sum_g w[g] * (dx . S[g], dy . S[g])
            // (cu:86-125: d/dx = W (-hh v1 + hh v2 - lh v3 + lh v4), d/dy = H (-hw v1 - lw v2 + hw v3 + lw v4))
            const int ra = s_qstart[qc0], rb = s_qstart[qc0 + nq];
            for (int r = ra + tid; r < rb; r += kThreads) {
                const int it = s_sorted[r];
                const int l = level_of(it);
                const int qi = (int)s_qof[it] - qc0;
                const int* t = tab + ((int)(s_key[it] >> 26) * L + l) * 3;
                const float2 f = s_lhlw[it];
                ItemGeo ig;
                ig.lh = f.x; ig.lw = f.y; ig.hh = 1.f - f.x; ig.hw = 1.f - f.y;   // same expressions as quad_setup()
                const int m = s_mask[it];
                const float W_ = (float)t[1], H_ = (float)t[0];
                const float c1 = (m & 1) ? ig.hh * ig.hw : 0.f, c2 = (m & 2) ? ig.hh * ig.lw : 0.f;
                const float c3 = (m & 4) ? ig.lh * ig.hw : 0.f, c4 = (m & 8) ? ig.lh * ig.lw : 0.f;
                const float x1 = (m & 1) ? -ig.hh * W_ : 0.f, x2 = (m & 2) ? ig.hh * W_ : 0.f;
                const float x3 = (m & 4) ? -ig.lh * W_ : 0.f, x4 = (m & 8) ? ig.lh * W_ : 0.f;
                const float y1 = (m & 1) ? -ig.hw * H_ : 0.f, y2 = (m & 2) ? -ig.lw * H_ : 0.f;
                const float y3 = (m & 4) ? ig.hw * H_ : 0.f, y4 = (m & 8) ? ig.lw * H_ : 0.f;
                const size_t woff = s_woff[it];
                const float* wsrc = w_unit + woff;
                float* gdst = gc.g_w + ((size_t)ba * NP + p0) * (L * G) + woff;
                float gx = 0.f, gy = 0.f;
                for (int g = 0; g < G; ++g) {
                    const float4 s4 = q_tab[qi * kGroupMaxG + g];
                    const float wv = __ldg(wsrc + g);
                    gdst[g] = c1 * s4.x + c2 * s4.y + c3 * s4.z + c4 * s4.w;
                    gx = __fmaf_rn(x1 * s4.x + x2 * s4.y + x3 * s4.z + x4 * s4.w, wv, gx);
                    gy = __fmaf_rn(y1 * s4.x + y2 * s4.y + y3 * s4.z + y4 * s4.w, wv, gy);
                }
                s_gxy[it] = make_float2(gx, gy);
            }
            DFA_TRACE_ACC(tr7, t7);
        }
        __syncthreads();      // the quad tables are rewritten by the next chunk / reused as reduction scratch
        if constexpr (!kBwd) DFA_TRACE_ACC(tr6, t6);
    }

    DFA_TRACE_V((int)blockIdx.x, 5, tr5);
    DFA_TRACE_V((int)blockIdx.x, 6, tr6);
    DFA_TRACE_V((int)blockIdx.x, 7, tr7);
    DFA_TRACE_V((int)blockIdx.x, 8, n_items);
    DFA_TRACE_V((int)blockIdx.x, 9, n_quads);
    DFA_TRACE((int)blockIdx.x, 10);
    if constexpr (!kBwd) {
        // -------------------------------------------------------------- phase 8: cross-warp sum in warp order
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
            __syncthreads();
            if (tid == 0) s_flag[0] = (atomic_add_release_gpu(p.tickets + gc.row_begin + ba, 1) == S - 1) ? 1 : 0;
            __syncthreads();
            if (s_flag[0]) {
                __threadfence();
                const float* parts = p.partial + ((size_t)gc.part_begin + (size_t)ba * S) * CPAD;
                // The S partial rows are contiguous: they are staged in this CTA's (now dead) shared-memory tables by bulk
                // asynchronous copies (cp.async.bulk + mbarrier: one thread issues up to ~18 KB, one round trip per chunk
                // instead of one per few slices) and added from there in slice order.
                constexpr int kRowBytes = CPAD * 4;
                constexpr int kChPer = (CPAD + kThreads - 1) / kThreads;
                float* stage = reinterpret_cast<float*>(smem_raw + so.lxy);
                const int cap_rows = (so.total - so.lxy) / kRowBytes;
                if (cap_rows >= 2 && (reinterpret_cast<uintptr_t>(parts) & 15) == 0) {
                    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(s_wcnt);       // 8-byte aligned, dead by now
                    const unsigned stage_a = (unsigned)__cvta_generic_to_shared(stage);
                    if (tid == 0) {
                        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1) : "memory");
                        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                    }
                    float acc_s[kChPer];
#pragma unroll
                    for (int c2 = 0; c2 < kChPer; ++c2) acc_s[c2] = 0.f;
                    unsigned phase = 0;
                    for (int q0 = 0; q0 < S; q0 += cap_rows) {
                        const int nr = min(cap_rows, S - q0);
                        __syncthreads();         // barrier initialised / previous chunk consumed; the tables are dead
                        if (tid == 0) {
                            const unsigned bytes = (unsigned)nr * (unsigned)kRowBytes;
                            // generic-proxy accesses (this CTA's earlier use of the tables; the other slices' partial rows,
                            // acquired through the ticket) before the async-proxy copy
                            asm volatile("fence.proxy.async;" ::: "memory");
                            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         ::"r"(stage_a), "l"(parts + (size_t)q0 * CPAD), "r"(bytes), "r"(bar_a) : "memory");
                        }
                        unsigned ok = 0;
                        do {
                            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                         : "=r"(ok) : "r"(bar_a), "r"(phase) : "memory");
                        } while (!ok);
                        phase ^= 1u;
#pragma unroll
                        for (int c2 = 0; c2 < kChPer; ++c2) {
                            const int ch = tid + c2 * kThreads;
                            if (ch < CPAD)
                                for (int q = 0; q < nr; ++q) acc_s[c2] += stage[q * CPAD + ch];
                        }
                    }
#pragma unroll
                    for (int c2 = 0; c2 < kChPer; ++c2) {
                        const int ch = tid + c2 * kThreads;
                        if (ch < CPAD) out_row[ch] = acc_s[c2];
                    }
                } else {
                    for (int ch = tid; ch < CPAD; ch += kThreads) {
                        float s = 0.f;
                        for (int q = 0; q < S; ++q) s += __ldcg(parts + (size_t)q * CPAD + ch);
                        out_row[ch] = s;
                    }
                }
            }
        }
    } else {
        // -------------------------------------------------------------- phase 8 (backward): location gradient per pair
        float2* const gloc = reinterpret_cast<float2*>(gc.g_loc) + (size_t)ba * NP + p0;
        for (int i = tid; i < n_list; i += kThreads) {
            const float2 t0 = s_gxy[i], t1 = s_gxy[n_pad + i], t2 = s_gxy[2 * n_pad + i], t3 = s_gxy[3 * n_pad + i];
            gloc[l_pair[i]] = make_float2(((t0.x + t1.x) + t2.x) + t3.x, ((t0.y + t1.y) + t2.y) + t3.y);
        }
    }
    DFA_TRACE((int)blockIdx.x, 11);
    DFA_TRACE_END((int)blockIdx.x, 12);
}

}  // namespace hipad
