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
constexpr int kRowsPerTile = kReduceWarps;
constexpr int kHeavyRow = 48;   // contributions above which a row is split across the CTA

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
        if (tid < kRadix) {
            unsigned base = 0;
            for (int dd = 0; dd < tid; ++dd) base += tot[dd];
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

template <typename W>
__device__ void bucket_sort_body(const GfeatParams& p, W* a, W* b, unsigned* hist, unsigned* tot, int* s_wcnt,
                                 int b_idx, int cam, int cl, int h, int w, int segoff) {
    const Dims d = p.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int AP = d.A * d.P;
    const int vb = bits_for((unsigned)AP);
    const int K = (h + 1) * (w + 1);
    const int kb = bits_for((unsigned)K);
    const float2* loc2 = reinterpret_cast<const float2*>(p.loc) + (size_t)b_idx * AP * d.cams + cam;

    // ordered compaction of the visible samples of (b, cam): word = key << vb | sample
    int n = 0;
    for (int base = 0; base < AP; base += kSortThreads) {
        const int s = base + tid;
        bool vis = false;
        W word = 0;
        if (s < AP) {
            const float2 xy = __ldg(loc2 + (size_t)s * d.cams);
            vis = loc_valid(xy.x, xy.y);
            if (vis) {
                const Quad q = quad_setup(xy.x, xy.y, h, w);
                const unsigned key = (unsigned)((q.h_low + 1) * (w + 1) + (q.w_low + 1));
                word = ((W)key << vb) | (W)s;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, vis);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        int before = 0, all = 0;
        for (int ww = 0; ww < kSortWarps; ++ww) {
            const int c = s_wcnt[ww];
            if (ww < warp) before += c;
            all += c;
        }
        __syncthreads();
        if (vis) a[n + before + __popc(bal & ((1u << lane) - 1u))] = word;
        n += all;
    }
    __syncthreads();

    W* sorted = block_radix_sort<W>(a, b, n, vb, kb, hist, tot);

    int* rec = p.rec + ((size_t)b_idx * d.cams * d.L + cl) * AP;
    const W vmask = ((W)1 << vb) - 1;
    for (int i = tid; i < n; i += kSortThreads) rec[i] = (int)(sorted[i] & vmask);
    // seg[k] = first position whose key >= k, k = 0..K
    int* seg = p.seg + (size_t)b_idx * p.seg_stride + segoff;
    for (int k = tid; k <= K; k += kSortThreads) {
        int lo = 0, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((unsigned)(sorted[mid] >> vb) < (unsigned)k) lo = mid + 1; else hi = mid;
        }
        seg[k] = lo;
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
    unsigned* tot = hist + kSortWarps * kRadix;
    int* s_wcnt = reinterpret_cast<int*>(tot + kRadix);
    int* s_misc = s_wcnt + kSortWarps;           // [0] segoff  [1] visible count of (b,cam)
    unsigned char* data = reinterpret_cast<unsigned char*>(s_misc + 16);

    const int tid = threadIdx.x;
    const int AP = d.A * d.P;
    if (tid == 0) {
        int off = 0;
        for (int i = 0; i < cl; ++i) off += (__ldg(p.shapes + i * 2) + 1) * (__ldg(p.shapes + i * 2 + 1) + 1) + 1;
        s_misc[0] = off;
        s_misc[1] = 0;
    }
    __syncthreads();
    // count visible samples first: decides shared vs global staging for this bucket
    {
        const float2* loc2 = reinterpret_cast<const float2*>(p.loc) + (size_t)b_idx * AP * d.cams + cam;
        int c = 0;
        for (int s = tid; s < AP; s += kSortThreads) {
            const float2 xy = __ldg(loc2 + (size_t)s * d.cams);
            c += loc_valid(xy.x, xy.y) ? 1 : 0;
        }
        c = (int)warp_sum((float)c);   // exact: per-warp counts < 2^24
        if ((tid & 31) == 0 && c) atomicAdd(&s_misc[1], c);
    }
    __syncthreads();
    const int n_vis = s_misc[1];
    const int segoff = s_misc[0];
    const int h = __ldg(p.shapes + cl * 2), w = __ldg(p.shapes + cl * 2 + 1);
    const int bits = bits_for((unsigned)AP) + bits_for((unsigned)((h + 1) * (w + 1)));
    unsigned long long* gbuf = p.sortbuf + ((size_t)b_idx * d.cams * d.L + cl) * 2 * AP;
    if (bits <= 32) {
        if (n_vis <= p.smem_cap) {
            unsigned* a = reinterpret_cast<unsigned*>(data);
            bucket_sort_body<unsigned>(p, a, a + p.smem_cap, hist, tot, s_wcnt, b_idx, cam, cl, h, w, segoff);
        } else {
            unsigned* a = reinterpret_cast<unsigned*>(gbuf);
            bucket_sort_body<unsigned>(p, a, a + AP, hist, tot, s_wcnt, b_idx, cam, cl, h, w, segoff);
        }
    } else {
        if (n_vis <= p.smem_cap / 2) {
            unsigned long long* a = reinterpret_cast<unsigned long long*>(data);
            bucket_sort_body<unsigned long long>(p, a, a + p.smem_cap / 2, hist, tot, s_wcnt, b_idx, cam, cl, h, w, segoff);
        } else {
            bucket_sort_body<unsigned long long>(p, gbuf, gbuf + AP, hist, tot, s_wcnt, b_idx, cam, cl, h, w, segoff);
        }
    }
}

inline size_t bucket_sort_smem_bytes(int smem_cap) {
    return (size_t)(kSortWarps * kRadix + kRadix + kSortWarps + 16) * 4 + (size_t)smem_cap * 2 * 4;
}

// ---------------------------------------------------------------------------------------------
// grid (tiles_upper_bound, bs), block kReduceWarps*32.  Tile = kRowsPerTile consecutive rows of one
// (cam, level); tiles are numbered coarsest level first.
template <typename T, int V, int NCH>
__global__ void __launch_bounds__(kReduceWarps * 32) dfa_gfeat_reduce_kernel(const GfeatParams p) {
    constexpr int CPAD = NCH * 32 * V;
    const Dims d = p.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b_idx = blockIdx.y;
    const int n_cl = d.cams * d.L;
    const int AP = d.A * d.P;
    const int gd = d.C / d.G;

    __shared__ int tab[kMaxCamLevels * 3];
    __shared__ int s_tile[4];                 // cl, row0
    __shared__ int s_seg[kRowsPerTile][8];    // 4 x (begin,end) per row
    __shared__ int s_nrow[kRowsPerTile];
    __shared__ __align__(16) float red[kReduceWarps * CPAD];

    load_level_table(tab, p.shapes, p.starts, n_cl);
    __syncthreads();
    if (tid == 0) {
        int t = blockIdx.x, found = -1, row0 = 0;
        for (int l = d.L - 1; l >= 0 && found < 0; --l)
            for (int cam = 0; cam < d.cams; ++cam) {
                const int cl = cam * d.L + l;
                const int nt = (tab[cl * 3] * tab[cl * 3 + 1] + kRowsPerTile - 1) / kRowsPerTile;
                if (t < nt) { found = cl; row0 = t * kRowsPerTile; break; }
                t -= nt;
            }
        s_tile[0] = found;
        s_tile[1] = row0;
    }
    __syncthreads();
    const int cl = s_tile[0];
    if (cl < 0) return;
    const int cam = cl / d.L, l = cl - cam * d.L;
    const int h = tab[cl * 3], w = tab[cl * 3 + 1], start = tab[cl * 3 + 2];
    int segoff = 0;
    for (int i = 0; i < cl; ++i) segoff += (tab[i * 3] + 1) * (tab[i * 3 + 1] + 1) + 1;
    const int* seg = p.seg + (size_t)b_idx * p.seg_stride + segoff;
    const int* rec = p.rec + ((size_t)b_idx * n_cl + cl) * AP;
    const float2* loc2 = reinterpret_cast<const float2*>(p.loc) + (size_t)b_idx * AP * d.cams + cam;
    const float* wts = p.weights + ((size_t)b_idx * AP * d.cams + cam) * d.L * d.G + (size_t)l * d.G;
    const float* gout = p.grad_out + (size_t)b_idx * d.A * d.C;
    T* gfeat = reinterpret_cast<T*>(p.g_feat) + ((size_t)b_idx * d.num_feat + start) * d.C;

    int ch[NCH], grp[NCH];
    bool act[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        ch[j] = (j * 32 + lane) * V;
        act[j] = ch[j] < d.C;
        grp[j] = act[j] ? ch[j] / gd : 0;
    }

    // segments of this warp's row
    const int my_row = s_tile[1] + warp;
    if (lane == 0) {
        int n = 0;
        if (my_row < h * w) {
            const int y = my_row / w, x = my_row - y * w;
            const int k4 = y * (w + 1) + x;            // quad (y-1,x-1): this row is its corner 4
            const int k1 = k4 + (w + 1) + 1;           // quad (y,x):     corner 1
            const int sa = __ldg(seg + k1), sb = __ldg(seg + k1 + 1);       // corner 1
            const int sc = __ldg(seg + k1 - 1);                             // corner 2: [sc, sa)
            const int sd = __ldg(seg + k4), se = __ldg(seg + k4 + 1), sf = __ldg(seg + k4 + 2);
            s_seg[warp][0] = sa; s_seg[warp][1] = sb;   // corner 1
            s_seg[warp][2] = sc; s_seg[warp][3] = sa;   // corner 2  key k1-1
            s_seg[warp][4] = se; s_seg[warp][5] = sf;   // corner 3  key k4+1 : quad (y-1,x)
            s_seg[warp][6] = sd; s_seg[warp][7] = se;   // corner 4  key k4
            n = (sb - sa) + (sa - sc) + (sf - se) + (se - sd);
        }
        s_nrow[warp] = (my_row < h * w) ? n : -1;
    }
    __syncthreads();

    // accumulate contributions [c_lo, c_hi) of row r (indices in the row's concatenated segment space)
    auto accumulate = [&](int r, int c_lo, int c_hi, float (&acc)[NCH][V]) {
        int pos = 0;
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
            const int sb = s_seg[r][2 * k], se = s_seg[r][2 * k + 1];
            const int lo = max(c_lo - pos, 0) + sb, hi = min(c_hi - pos, se - sb) + sb;
            pos += se - sb;
            for (int i0 = lo; i0 < hi; i0 += 32) {
                // lane-parallel metadata for up to 32 contributions
                const int i = i0 + lane;
                int s_id = 0;
                float coef = 0.f;
                if (i < hi) {
                    s_id = __ldg(rec + i);
                    const float2 xy = __ldg(loc2 + (size_t)s_id * d.cams);
                    const Quad q = quad_setup(xy.x, xy.y, h, w);
                    coef = (k == 0) ? q.hh * q.hw : (k == 1) ? q.hh * q.lw : (k == 2) ? q.lh * q.hw : q.lh * q.lw;
                }
                const int cnt = min(32, hi - i0);
#pragma unroll 4
                for (int m = 0; m < cnt; ++m) {
                    const int sm = __shfl_sync(0xffffffffu, s_id, m);
                    const float cf = __shfl_sync(0xffffffffu, coef, m);
                    const int a_idx = sm / d.P;
                    const float* wrow = wts + (size_t)sm * d.cams * d.L * d.G;
                    const float* grow = gout + (size_t)a_idx * d.C;
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        if (!act[j]) continue;
                        float g[V];
                        VecIO<float, V>::load(grow + ch[j], g);
                        const float wg = __ldg(wrow + grp[j]);
#pragma unroll
                        for (int e = 0; e < V; ++e) acc[j][e] = __fmaf_rn(cf, g[e] * wg, acc[j][e]);
                    }
                }
            }
        }
    };

    // light rows: one warp each
    {
        const int n = s_nrow[warp];
        if (n >= 0 && n <= kHeavyRow) {
            float acc[NCH][V];
#pragma unroll
            for (int j = 0; j < NCH; ++j)
#pragma unroll
                for (int e = 0; e < V; ++e) acc[j][e] = 0.f;
            if (n > 0) accumulate(warp, 0, n, acc);
#pragma unroll
            for (int j = 0; j < NCH; ++j)
                if (act[j]) VecIO<T, V>::store(gfeat + (size_t)my_row * d.C + ch[j], acc[j]);
        }
    }
    // heavy rows: all warps of the CTA share one row, fixed-order combine
    for (int r = 0; r < kRowsPerTile; ++r) {
        const int n = s_nrow[r];
        if (n <= kHeavyRow) continue;   // uniform across the CTA
        const int per = (n + kReduceWarps - 1) / kReduceWarps;
        float acc[NCH][V];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int e = 0; e < V; ++e) acc[j][e] = 0.f;
        accumulate(r, min(n, warp * per), min(n, (warp + 1) * per), acc);
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int e = 0; e < V; ++e) red[warp * CPAD + (j * 32 + lane) * V + e] = acc[j][e];
        __syncthreads();
        if (warp == 0) {
            float sum[NCH][V];
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    float s = 0.f;
#pragma unroll
                    for (int ww = 0; ww < kReduceWarps; ++ww) s += red[ww * CPAD + (j * 32 + lane) * V + e];
                    sum[j][e] = s;
                }
                if (act[j]) VecIO<T, V>::store(gfeat + (size_t)(s_tile[1] + r) * d.C + ch[j], sum[j]);
            }
        }
        __syncthreads();
    }
}

}  // namespace hipad
