// dfa_gfeat.cuh — deterministic feature-map gradient (the scatter half of cu:62-126).
//
// The reference scatters 4 fp32 atomicAdds per (sample, level, channel) into grad_mc_ms_feat.
// Here the scatter is turned into a gather, so every feature row is written exactly once,
// in a fixed summation order, with zeros for untouched rows (no memset, no atomics):
//
//   dfa_bucket_sort_kernel   one CTA per (b, cam, level) "bucket".  Scans the bucket's samples
//       in canonical order, compacts the visible ones, computes the PADDED quad key
//       (h_low+1)*(W+1) + (w_low+1) of each, and radix-sorts (key, sample) words inside shared
//       memory (stable LSD passes built on __match_any_sync; global-memory ping-pong only if a
//       bucket exceeds the 227 KB CTA budget).  Emits the sorted sample ids and a dense
//       segment table seg[key] = first sorted position with key' >= key.
//   dfa_gfeat_reduce_kernel  one warp per feature row.  A row (y,x) is corner 1/2/3/4 of the
//       quads keyed (y,x), (y,x-1), (y-1,x), (y-1,x-1): four segments of the table.  The warp
//       walks them in order, re-derives the bilinear coefficient from the sample location with
//       the same quad_setup() as the forward, and accumulates coef * w[g] * grad_out[b,a,:]
//       in registers (lane = V channels x NCH chunks, LDG.128 coalesced).  Rows with many
//       contributions (coarse levels) are split across the CTA's warps and combined in fixed
//       order through shared memory.  Coarse levels are scheduled first (longest work first).
#pragma once
#include "dfa_common.cuh"

namespace hipad {

constexpr int kSortThreads = 1024;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kReduceWarps = 8;
constexpr int kRowsPerTile = 32;         // rows of one (cam, level) per CTA
constexpr int kHeavyRow = 64;
constexpr int kHeavyCtas = 148 * 6;      // persistent CTAs of the heavy-row kernel   // contributions above which a row is split across the CTA

struct GfeatParams {
    const int* shapes;
    const int* starts;
    const float* loc;
    const float* weights;
    const float* grad_out;
    void* g_feat;
    int* rec;        // [bs][cams*L][A*P]   sorted sample ids (a*P + p)
    int* seg;        // [bs][seg_stride]    per bucket: (H+1)(W+1)+1 entries
    int* counts;     // [bs][cams*L]        visible samples per bucket
    unsigned long long* sortbuf;  // [bs][cams*L][2][A*P] global ping-pong (large buckets only)
    int2* heavy_list;  // [bs*num_feat] rows with > kHeavyRow contributions: (b*cams*L + cl, row in level)
    int* heavy_count;  // zeroed by the bucket-sort kernel
    Dims d;
    int seg_stride;  // ints per batch element in seg
    int smem_cap;    // words of packed records that fit in shared memory (per ping-pong half)
};

__host__ __device__ inline int bits_for(unsigned v) {   // number of bits to represent values < v
    int b = 0;
    while ((1ull << b) < (unsigned long long)v) ++b;
    return b;
}

// Stable LSD radix sort of n packed words on bits [lo_bit, lo_bit + nbits) by one CTA.
// a/b: ping-pong arrays (shared or global).  Returns the array holding the result.
template <typename W>
__device__ W* block_radix_sort(W* a, W* b, int n, int lo_bit, int nbits, unsigned* hist /*[kSortWarps][kRadix]*/,
                               unsigned* tot /*[kRadix]*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int chunk = (n + kSortWarps - 1) / kSortWarps;
    chunk = (chunk + 31) & ~31;
    const int beg = min(n, warp * chunk), end = min(n, beg + chunk);
    for (int shift = lo_bit; shift < lo_bit + nbits; shift += kRadixBits) {
        unsigned* my = hist + warp * kRadix;
        for (int i = lane; i < kRadix; i += 32) my[i] = 0;
        __syncwarp();
        for (int i0 = beg; i0 < end; i0 += 32) {
            const int i = i0 + lane;
            const bool has = i < end;
            const unsigned dgt = has ? (unsigned)((a[i] >> shift) & (kRadix - 1)) : kRadix;
            const unsigned peers = __match_any_sync(0xffffffffu, dgt);
            if (has && lane == __ffs(peers) - 1) my[dgt] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        if (tid < kRadix) {
            unsigned run = 0;
            for (int w = 0; w < kSortWarps; ++w) {
                const unsigned c = hist[w * kRadix + tid];
                hist[w * kRadix + tid] = run;
                run += c;
            }
            tot[tid] = run;
        }
        __syncthreads();
        if (tid < kRadix) {   // exclusive scan of the 256 digit totals: warp scan + 8 warp carries
            const unsigned mine = tot[tid];
            unsigned inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) tot[kRadix + warp] = inc;     // tot has kRadix + 32 entries
            __syncwarp();
            asm volatile("bar.sync 1, 256;");              // only the first 8 warps take part
            unsigned base = 0;
            for (int ww = 0; ww < warp; ++ww) base += tot[kRadix + ww];
            base += inc - mine;
            for (int w = 0; w < kSortWarps; ++w) hist[w * kRadix + tid] += base;
        }
        __syncthreads();
        for (int i0 = beg; i0 < end; i0 += 32) {
            const int i = i0 + lane;
            const bool has = i < end;
            W word = 0;
            unsigned dgt = kRadix;
            if (has) {
                word = a[i];
                dgt = (unsigned)((word >> shift) & (kRadix - 1));
            }
            const unsigned peers = __match_any_sync(0xffffffffu, dgt);
            if (has) {
                const unsigned pos = my[dgt] + __popc(peers & ((1u << lane) - 1u));
                b[pos] = word;
            }
            __syncwarp();
            if (has && lane == __ffs(peers) - 1) my[dgt] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        W* t = a; a = b; b = t;
    }
    return a;
}

// exclusive scan of kSortWarps (=32) per-warp counts by warp 0; returns total via smem
__device__ __forceinline__ void scan_warp_counts(int* s_wcnt /*[32] in: counts, out: exclusive offsets*/, int* s_total) {
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int c = s_wcnt[lane];
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        s_wcnt[lane] = inc - c;
        if (lane == 31) *s_total = inc;
    }
}

template <typename W>
__device__ void bucket_sort_body(const GfeatParams& p, W* a, W* b, unsigned* hist, unsigned* tot, const int* s_woff,
                                 int n, int b_idx, int cam, int cl, int h, int w, int segoff) {
    const Dims d = p.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int AP = d.A * d.P;
    const int vb = bits_for((unsigned)AP);
    const int K = (h + 1) * (w + 1);
    const int kb = bits_for((unsigned)K);
    const float2* loc2 = reinterpret_cast<const float2*>(p.loc) + (size_t)b_idx * AP * d.cams + cam;

    // ordered compaction: warp `warp` owns the contiguous sample range [beg, end) and writes its
    // visible samples at s_woff[warp] + running rank -> canonical (sample-index) order overall.
    int chunk = (AP + kSortWarps - 1) / kSortWarps;
    chunk = (chunk + 31) & ~31;
    const int beg = min(AP, warp * chunk), end = min(AP, beg + chunk);
    int pos = s_woff[warp];
    for (int s0 = beg; s0 < end; s0 += 32) {
        const int s = s0 + lane;
        bool vis = false;
        W word = 0;
        if (s < end) {
            const float2 xy = __ldg(loc2 + (size_t)s * d.cams);
            vis = loc_valid(xy.x, xy.y);
            if (vis) {
                const Quad q = quad_setup(xy.x, xy.y, h, w);
                const unsigned key = (unsigned)((q.h_low + 1) * (w + 1) + (q.w_low + 1));
                word = ((W)key << vb) | (W)s;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, vis);
        if (vis) a[pos + __popc(bal & ((1u << lane) - 1u))] = word;
        pos += __popc(bal);
    }
    __syncthreads();

    W* sorted = block_radix_sort<W>(a, b, n, vb, kb, hist, tot);

    int* rec = p.rec + ((size_t)b_idx * d.cams * d.L + cl) * AP;
    const W vmask = ((W)1 << vb) - 1;
    for (int i = tid; i < n; i += kSortThreads) rec[i] = (int)(sorted[i] & vmask);

    // seg[k] = first sorted position whose key >= k, for k = 0..K:
    //   fill with n, mark segment starts, then a suffix-min scan closes the gaps of empty keys.
    int* seg = p.seg + (size_t)b_idx * p.seg_stride + segoff;
    for (int k = tid; k <= K; k += kSortThreads) seg[k] = n;
    __syncthreads();
    for (int i = tid; i < n; i += kSortThreads) {
        const unsigned key = (unsigned)(sorted[i] >> vb);
        if (i == 0 || (unsigned)(sorted[i - 1] >> vb) != key) seg[key] = i;
    }
    __syncthreads();
    {
        const int per = (K + 1 + kSortThreads - 1) / kSortThreads;
        const int lo = min(K + 1, tid * per), hi = min(K + 1, lo + per);
        int run = 0x7fffffff;
        for (int k = hi - 1; k >= lo; --k) run = min(run, seg[k]);
        // exclusive suffix-min across threads (thread t needs min over threads > t)
        int v = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_down_sync(0xffffffffu, v, o);
            if (lane + o < 32) v = min(v, t);
        }
        int* s_wmin = reinterpret_cast<int*>(tot);          // 32 ints of scratch
        __syncthreads();
        if (lane == 0) s_wmin[warp] = v;
        __syncthreads();
        int after = 0x7fffffff;                             // min over later warps
        for (int ww = warp + 1; ww < kSortWarps; ++ww) after = min(after, s_wmin[ww]);
        int excl = __shfl_down_sync(0xffffffffu, v, 1);     // inclusive suffix-min of lanes > lane
        if (lane == 31) excl = 0x7fffffff;
        int carry = min(excl, after);
        for (int k = hi - 1; k >= lo; --k) {
            carry = min(carry, seg[k]);
            seg[k] = carry;
        }
    }
    if (tid == 0) p.counts[(size_t)b_idx * d.cams * d.L + cl] = n;
}

// grid (cams*L, bs), block kSortThreads, dynamic smem = hist + tot + counters + 2*smem_cap words
__global__ void __launch_bounds__(kSortThreads) dfa_bucket_sort_kernel(const GfeatParams p) {
    const Dims d = p.d;
    const int cl = blockIdx.x, b_idx = blockIdx.y;
    const int cam = cl / d.L;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* hist = reinterpret_cast<unsigned*>(smem_raw);
    unsigned* tot = hist + kSortWarps * kRadix;            // kRadix totals + 32 scratch
    int* s_wcnt = reinterpret_cast<int*>(tot + kRadix + 32);
    int* s_misc = s_wcnt + kSortWarps;           // [0] segoff  [1] visible count of (b,cam)
    unsigned char* data = reinterpret_cast<unsigned char*>(s_misc + 16);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int AP = d.A * d.P;
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) *p.heavy_count = 0;
    if (warp == 0) {
        int off = 0;
        for (int i = lane; i < cl; i += 32)
            off += (__ldg(p.shapes + i * 2) + 1) * (__ldg(p.shapes + i * 2 + 1) + 1) + 1;
        off = __reduce_add_sync(0xffffffffu, off);
        if (lane == 0) s_misc[0] = off;
    }
    // per-warp visible counts over the same contiguous ranges the compaction pass will use
    {
        const float2* loc2 = reinterpret_cast<const float2*>(p.loc) + (size_t)b_idx * AP * d.cams + cam;
        int chunk = (AP + kSortWarps - 1) / kSortWarps;
        chunk = (chunk + 31) & ~31;
        const int beg = min(AP, warp * chunk), end = min(AP, beg + chunk);
        int c = 0;
#pragma unroll 8
        for (int s = beg + lane; s < end; s += 32) {
            const float2 xy = __ldg(loc2 + (size_t)s * d.cams);
            c += loc_valid(xy.x, xy.y) ? 1 : 0;
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) s_wcnt[warp] = c;
    }
    __syncthreads();
    scan_warp_counts(s_wcnt, &s_misc[1]);
    __syncthreads();
    const int n_vis = s_misc[1];
    const int segoff = s_misc[0];
    const int h = __ldg(p.shapes + cl * 2), w = __ldg(p.shapes + cl * 2 + 1);
    const int bits = bits_for((unsigned)AP) + bits_for((unsigned)((h + 1) * (w + 1)));
    unsigned long long* gbuf = p.sortbuf + ((size_t)b_idx * d.cams * d.L + cl) * 2 * AP;
    if (bits <= 32) {
        if (n_vis <= p.smem_cap) {
            unsigned* a = reinterpret_cast<unsigned*>(data);
            bucket_sort_body<unsigned>(p, a, a + p.smem_cap, hist, tot, s_wcnt, n_vis, b_idx, cam, cl, h, w, segoff);
        } else {
            unsigned* a = reinterpret_cast<unsigned*>(gbuf);
            bucket_sort_body<unsigned>(p, a, a + AP, hist, tot, s_wcnt, n_vis, b_idx, cam, cl, h, w, segoff);
        }
    } else {
        if (n_vis <= p.smem_cap / 2) {
            unsigned long long* a = reinterpret_cast<unsigned long long*>(data);
            bucket_sort_body<unsigned long long>(p, a, a + p.smem_cap / 2, hist, tot, s_wcnt, n_vis, b_idx, cam, cl, h, w,
                                                 segoff);
        } else {
            bucket_sort_body<unsigned long long>(p, gbuf, gbuf + AP, hist, tot, s_wcnt, n_vis, b_idx, cam, cl, h, w,
                                                 segoff);
        }
    }
}

inline size_t bucket_sort_smem_bytes(int smem_cap) {
    return (size_t)(kSortWarps * kRadix + kRadix + 32 + kSortWarps + 16) * 4 + (size_t)smem_cap * 2 * 4;
}

// ---------------------------------------------------------------------------------------------
// Feature-major reduce.  Two kernels:
//   dfa_gfeat_reduce_kernel  grid (tiles_upper_bound, bs): tile = kRowsPerTile consecutive rows of one
//       (cam, level), coarsest level first.  One warp per row; rows with more than kHeavyRow
//       contributions are only APPENDED to a device work list (which CTA ends up processing such a row
//       is irrelevant for determinism: the summation order inside a row is fixed).
//   dfa_gfeat_heavy_kernel   persistent CTAs walk that list; the 8 warps of a CTA split one row's
//       contributions into contiguous ranges and combine the partials in warp order.
template <int V, int NCH>
struct RowCtx {
    const int* rec;        // sorted sample ids of the bucket
    const float2* loc2;    // locations of (b, :, :, cam)
    const float* wts;      // weights of (b, :, :, cam, l, :)
    const float* gout;     // grad_out of batch element b
    int h, w, cams, P, C, w_stride;
    int beg[4];            // begin of corner-1/2/3/4 segments in rec[]
    int e1, e2, e3;        // cumulative ends of the first three segments
    int ch[NCH], grp[NCH];
};

// accumulate contributions [c_lo, c_hi) of one row, indices in the row's concatenated
// (corner1, corner2, corner3, corner4) order -- the fixed summation order of this library.
template <int V, int NCH>
__device__ __forceinline__ void accumulate_row(const RowCtx<V, NCH>& cx, int c_lo, int c_hi, float (&acc)[NCH][V]) {
    const int lane = threadIdx.x & 31;
    for (int v0 = c_lo; v0 < c_hi; v0 += 32) {
        // lane-parallel metadata for up to 32 contributions
        const int v = v0 + lane;
        int go_off = 0, w_off = 0;
        float coef = 0.f;
        if (v < c_hi) {
            const int k = (v >= cx.e1) + (v >= cx.e2) + (v >= cx.e3);
            const int first = (k == 0) ? 0 : (k == 1) ? cx.e1 : (k == 2) ? cx.e2 : cx.e3;
            const int bk = (k == 0) ? cx.beg[0] : (k == 1) ? cx.beg[1] : (k == 2) ? cx.beg[2] : cx.beg[3];
            const int s_id = __ldg(cx.rec + bk + (v - first));
            const float2 xy = __ldg(cx.loc2 + (size_t)s_id * cx.cams);
            const Quad q = quad_setup(xy.x, xy.y, cx.h, cx.w);
            coef = (k == 0) ? q.hh * q.hw : (k == 1) ? q.hh * q.lw : (k == 2) ? q.lh * q.hw : q.lh * q.lw;
            go_off = (s_id / cx.P) * cx.C;
            w_off = s_id * cx.w_stride;     // < 2^31 checked on the host
        }
        const int cnt = min(32, c_hi - v0);
#pragma unroll 4
        for (int m = 0; m < cnt; ++m) {
            const int go_m = __shfl_sync(0xffffffffu, go_off, m);
            const int w_m = __shfl_sync(0xffffffffu, w_off, m);
            const float cf = __shfl_sync(0xffffffffu, coef, m);
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                float g[V];
                VecIO<float, V>::load(cx.gout + go_m + cx.ch[j], g);
                const float wg = __ldg(cx.wts + w_m + cx.grp[j]) * cf;
#pragma unroll
                for (int e = 0; e < V; ++e) acc[j][e] = __fmaf_rn(wg, g[e], acc[j][e]);
            }
        }
    }
}

// row (y,x) of a (cam,level) of size h x w: the four quad-key segments that contribute to it
__device__ __forceinline__ int row_segments(const int* __restrict__ seg, int row, int w, int (&beg)[4], int& e1, int& e2,
                                            int& e3) {
    const int y = row / w, x = row - y * w;
    const int k4 = y * (w + 1) + x;            // quad (y-1,x-1): this row is its corner 4
    const int k1 = k4 + (w + 1) + 1;           // quad (y,x):     corner 1
    const int sc = __ldg(seg + k1 - 1), sa = __ldg(seg + k1), sb = __ldg(seg + k1 + 1);
    const int sd = __ldg(seg + k4), se = __ldg(seg + k4 + 1), sf = __ldg(seg + k4 + 2);
    // accumulation order: corner 1 [sa,sb), corner 2 [sc,sa), corner 3 [se,sf), corner 4 [sd,se)
    beg[0] = sa; beg[1] = sc; beg[2] = se; beg[3] = sd;
    e1 = sb - sa;
    e2 = e1 + (sa - sc);
    e3 = e2 + (sf - se);
    return e3 + (se - sd);
}

template <typename T, int V, int NCH>
__global__ void __launch_bounds__(kReduceWarps * 32) dfa_gfeat_reduce_kernel(const GfeatParams p) {
    constexpr int kRowsPerWarp = kRowsPerTile / kReduceWarps;
    const Dims d = p.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b_idx = blockIdx.y;
    const int n_cl = d.cams * d.L;
    const int AP = d.A * d.P;
    const int gd = d.C / d.G;

    __shared__ int s_tile[8];                  // cl, row0, h, w, start, segoff
    __shared__ int s_beg[kRowsPerTile][4];
    __shared__ int s_end[kRowsPerTile][4];     // e1, e2, e3, n

    if (warp == 0) {
        // tile decode, lane-parallel over the (level, cam) list ordered coarsest level first
        int t = blockIdx.x, found = -1, row0 = 0, segoff = 0;
        for (int base = 0; base < n_cl && found < 0; base += 32) {
            const int i = base + lane;
            int cl_i = -1, nt = 0;
            if (i < n_cl) {
                const int l = d.L - 1 - i / d.cams, cam = i - (i / d.cams) * d.cams;
                cl_i = cam * d.L + l;
                nt = (__ldg(p.shapes + cl_i * 2) * __ldg(p.shapes + cl_i * 2 + 1) + kRowsPerTile - 1) / kRowsPerTile;
            }
            int inc = nt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            const unsigned hit = __ballot_sync(0xffffffffu, i < n_cl && t < inc);
            if (hit) {
                const int src = __ffs(hit) - 1;
                found = __shfl_sync(0xffffffffu, cl_i, src);
                row0 = (t - (__shfl_sync(0xffffffffu, inc, src) - __shfl_sync(0xffffffffu, nt, src))) * kRowsPerTile;
            } else {
                t -= __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        if (found >= 0) {
            for (int i = lane; i < found; i += 32)
                segoff += (__ldg(p.shapes + i * 2) + 1) * (__ldg(p.shapes + i * 2 + 1) + 1) + 1;
            segoff = __reduce_add_sync(0xffffffffu, segoff);
        }
        if (lane == 0) {
            s_tile[0] = found;
            s_tile[1] = row0;
            if (found >= 0) {
                s_tile[2] = __ldg(p.shapes + found * 2);
                s_tile[3] = __ldg(p.shapes + found * 2 + 1);
                s_tile[4] = __ldg(p.starts + found);
                s_tile[5] = segoff;
            }
        }
    }
    __syncthreads();
    const int cl = s_tile[0];
    if (cl < 0) return;
    const int row0 = s_tile[1], h = s_tile[2], w = s_tile[3], start = s_tile[4];
    const int cam = cl / d.L, l = cl - cam * d.L;
    const int* seg = p.seg + (size_t)b_idx * p.seg_stride + s_tile[5];
    T* gfeat = reinterpret_cast<T*>(p.g_feat) + ((size_t)b_idx * d.num_feat + start) * d.C;

    // segment bounds of every row of the tile, one thread per row
    if (tid < kRowsPerTile) {
        const int row = row0 + tid;
        int beg[4], e1 = 0, e2 = 0, e3 = 0, n = -1;
        beg[0] = beg[1] = beg[2] = beg[3] = 0;
        if (row < h * w) {
            n = row_segments(seg, row, w, beg, e1, e2, e3);
            if (n > kHeavyRow) {
                const int slot = atomicAdd(p.heavy_count, 1);
                p.heavy_list[slot] = make_int2(b_idx * n_cl + cl, row);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) s_beg[tid][k] = beg[k];
        s_end[tid][0] = e1; s_end[tid][1] = e2; s_end[tid][2] = e3; s_end[tid][3] = n;
    }
    __syncthreads();

    RowCtx<V, NCH> cx;
    cx.rec = p.rec + ((size_t)b_idx * n_cl + cl) * AP;
    cx.loc2 = reinterpret_cast<const float2*>(p.loc) + (size_t)b_idx * AP * d.cams + cam;
    cx.wts = p.weights + ((size_t)b_idx * AP * d.cams + cam) * d.L * d.G + (size_t)l * d.G;
    cx.gout = p.grad_out + (size_t)b_idx * d.A * d.C;
    cx.h = h; cx.w = w; cx.cams = d.cams; cx.P = d.P; cx.C = d.C; cx.w_stride = d.cams * d.L * d.G;
    bool act[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int c_raw = (j * 32 + lane) * V;
        act[j] = c_raw < d.C;
        cx.ch[j] = act[j] ? c_raw : d.C - V;    // clamped lanes compute values nobody stores
        cx.grp[j] = cx.ch[j] / gd;
    }

#pragma unroll 1
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int r = rr * kReduceWarps + warp;
        const int n = s_end[r][3];
        if (n < 0 || n > kHeavyRow) continue;     // beyond the level / handled by the heavy kernel
        float acc[NCH][V];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int e = 0; e < V; ++e) acc[j][e] = 0.f;
        if (n > 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) cx.beg[k] = s_beg[r][k];
            cx.e1 = s_end[r][0]; cx.e2 = s_end[r][1]; cx.e3 = s_end[r][2];
            accumulate_row<V, NCH>(cx, 0, n, acc);
        }
#pragma unroll
        for (int j = 0; j < NCH; ++j)
            if (act[j]) VecIO<T, V>::store(gfeat + (size_t)(row0 + r) * d.C + cx.ch[j], acc[j]);
    }
}

// grid (kHeavyCtas), block kReduceWarps*32: persistent walk over the heavy-row list
template <typename T, int V, int NCH>
__global__ void __launch_bounds__(kReduceWarps * 32) dfa_gfeat_heavy_kernel(const GfeatParams p) {
    constexpr int CPAD = NCH * 32 * V;
    const Dims d = p.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_cl = d.cams * d.L;
    const int AP = d.A * d.P;
    const int gd = d.C / d.G;
    __shared__ __align__(16) float red[kReduceWarps * CPAD];

    const int n_heavy = *p.heavy_count;
    RowCtx<V, NCH> cx;
    bool act[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int c_raw = (j * 32 + lane) * V;
        act[j] = c_raw < d.C;
        cx.ch[j] = act[j] ? c_raw : d.C - V;
        cx.grp[j] = cx.ch[j] / gd;
    }
    cx.cams = d.cams; cx.P = d.P; cx.C = d.C; cx.w_stride = d.cams * d.L * d.G;

    for (int e = blockIdx.x; e < n_heavy; e += gridDim.x) {
        const int2 ent = p.heavy_list[e];
        const int b_idx = ent.x / n_cl, cl = ent.x - b_idx * n_cl, row = ent.y;
        const int cam = cl / d.L, l = cl - cam * d.L;
        const int h = __ldg(p.shapes + cl * 2), w = __ldg(p.shapes + cl * 2 + 1), start = __ldg(p.starts + cl);
        int segoff = 0;
        for (int i = 0; i < cl; ++i) segoff += (__ldg(p.shapes + i * 2) + 1) * (__ldg(p.shapes + i * 2 + 1) + 1) + 1;
        const int* seg = p.seg + (size_t)b_idx * p.seg_stride + segoff;
        const int n = row_segments(seg, row, w, cx.beg, cx.e1, cx.e2, cx.e3);
        cx.rec = p.rec + ((size_t)b_idx * n_cl + cl) * AP;
        cx.loc2 = reinterpret_cast<const float2*>(p.loc) + (size_t)b_idx * AP * d.cams + cam;
        cx.wts = p.weights + ((size_t)b_idx * AP * d.cams + cam) * d.L * d.G + (size_t)l * d.G;
        cx.gout = p.grad_out + (size_t)b_idx * d.A * d.C;
        cx.h = h; cx.w = w;

        // contiguous ranges in multiples of 32 so every warp runs full metadata batches
        int per = (n + kReduceWarps - 1) / kReduceWarps;
        per = (per + 31) & ~31;
        float acc[NCH][V];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int q = 0; q < V; ++q) acc[j][q] = 0.f;
        accumulate_row<V, NCH>(cx, min(n, warp * per), min(n, (warp + 1) * per), acc);
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int q = 0; q < V; ++q) red[warp * CPAD + (j * 32 + lane) * V + q] = acc[j][q];
        __syncthreads();
        if (warp == 0) {
            T* dst = reinterpret_cast<T*>(p.g_feat) + ((size_t)b_idx * d.num_feat + start + row) * d.C;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                float sum[V];
#pragma unroll
                for (int q = 0; q < V; ++q) {
                    float s = 0.f;
#pragma unroll
                    for (int ww = 0; ww < kReduceWarps; ++ww) s += red[ww * CPAD + (j * 32 + lane) * V + q];
                    sum[q] = s;
                }
                if (act[j]) VecIO<T, V>::store(dst + cx.ch[j], sum);
            }
        }
        __syncthreads();
    }
}

}  // namespace hipad
