// dfa_gfeat.cuh — deterministic feature-map gradient (the scatter half of cu:62-126).
//
// The reference scatters 4 fp32 atomicAdds per (sample, level, channel) into a pre-zeroed
// grad_mc_ms_feat.  Here the scatter is turned into a gather, every touched feature row is summed in a
// fixed order and written once, and no floating-point atomic exists:
//
//   dfa_zero_kernel         dense zero fill of grad_mc_ms_feat (streaming 16-byte stores).  When the
//       sample-major backward kernel has enough CTAs the fill is folded into that kernel instead (its
//       stores ride under the gather latency of that kernel, which leaves the HBM write path idle).
//   dfa_vis_compact_kernel  per (b, cam, chunk of 1024 samples): ordered compaction of the visible
//       samples; per level it writes the packed word (quad key << id bits | sample id) of every visible sample,
//       the chunk's histogram over kFine equal key ranges of the (cam, level) bucket as an exclusive prefix, and
//       adds the histogram to the bucket's global one (integer atomics).  The only pass that touches every location.
//   dfa_band_sort_kernel    per (b, cam, level, BAND): the bucket's key space is cut into bands of roughly EQUAL
//       SAMPLE COUNT from the global histogram (every band CTA derives the same cuts; a band's records start at the
//       prefix of the counts before it -- no cursor, no atomics).  The CTA picks the words of its key range out of
//       the bucket's list (ordered: chunk order, then order inside the chunk), radix-sorts them by key in shared
//       memory (stable LSD, ballot ranks), and emits 16-byte records {sample, anchor, lh, lw} + the
//       record's G weights in sorted order plus its slice of the bucket's segment table seg[key] = first record with
//       key' >= key.  The order INSIDE a key segment is the stable sample order, which is all the summation order
//       depends on.  (Bands used to be equal ranges of quad ROWS: on a stage-2 layer one band of 24 held 11 700 of a
//       camera's 14 400 samples on the coarsest level and ran 55 us while the average band CTA ran 17.)
//   dfa_row_classify_kernel one thread per feature row: 4 segment lookups (row (y,x) is corner 1/2/3/4 of
//       the quads keyed (y,x), (y,x-1), (y-1,x), (y-1,x-1)); touched rows are appended (warp-aggregated
//       integer atomics) to work lists together with their segment bounds: rows with <= kTinyRow
//       contributions to the tiny list, the others as one PART item per kPart contributions.
//       Untouched rows (most of them) cost a few loads.
//   dfa_gfeat_reduce_kernel persistent warps pull items from one dynamic queue: a part item is summed by a
//       warp (lane = V channels x NCH chunks, LDG.128, loads of 8 contributions in flight), a tiny item
//       is four rows summed by the four quarters of a warp.  Rows of several parts are combined in part
//       order by whichever warp finishes last (see the kernel).
#pragma once
#include "dfa_common.cuh"

namespace hipad {

constexpr int kVisChunk = 1024;          // samples per compaction chunk
constexpr int kVisThreads = 256;
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kMaxBands = 32;            // band CTAs launched per (cam, level) bucket; a bucket uses as many as its samples need
constexpr int kFine = 128;               // equal key ranges per bucket the histograms count (band cuts fall between them)
constexpr int kCum = kFine + 4;          // ints per (chunk, level) prefix row: kFine + 1 used, padded to 16 bytes
constexpr int kSegScale = 5;             // ints of segment table per feature row (upper bound, see seg_offset)
constexpr int kClassifyThreads = 256;    // rows classified per CTA (one thread each)
constexpr int kReduceCtas = 148 * 2;     // persistent CTAs (8 warps, 2 per SM; 4 per SM for the light variant) of the reduce kernel
constexpr int kTinyRow = 8;              // contributions up to which a row is summed by a quarter warp
constexpr int kPart = 64;                // contributions per part item (longer rows are split into parts); measured on the
                                         // stage-2 det / map / plan calls: 32 -> 46 / 76 / 81 us, 64 -> 45 / 68 / 65 us, 128 -> 49 / 73 / 66 us
constexpr int kMaxChunks = 4096;         // A*P <= 4 Mi samples per batch element

// One aggregation call of a group (all calls of a group read the same feature maps and share ONE feature gradient).
// Samples of the group live in one id space: call k owns ids [id_begin, id_begin + A*P), id_begin a multiple of
// kVisChunk, so a compaction chunk never straddles two calls.
constexpr int kMaxGfeatCalls = 8;
struct GfeatCall {
    const float* loc;        // [bs, A, P, cams, 2]
    const float* weights;    // [bs, A, P, cams, L, G]
    int A, P;
    int id_begin;            // first sample id of this call in the group's id space
    int anchor_begin;        // first row of this call in the packed grad_out [bs, A_total, C]
};

struct GfeatParams {
    const int* shapes;
    const int* starts;
    GfeatCall calls[kMaxGfeatCalls];
    int ncalls;
    int n_ids;         // size of the group's sample id space (sum of the calls' A*P rounded up to kVisChunk)
    int A_total;       // rows per batch element of the packed grad_out
    const float* grad_out;   // [bs, A_total, C]
    void* g_feat;
    unsigned long long* vis_key;  // [bs][cams*L][n_ids] slots of 8 bytes: packed words (key << id bits | id) of the visible
                                  // samples, chunk c at [c*kVisChunk, ...), grouped by fine bin inside the chunk; 32-bit
                                  // words where key and id fit (list_is_wide)
    int* chunk_cum;    // [bs][cams][n_chunks][L][kCum]  per chunk and level: samples whose fine bin is < f, f = 0..kFine
    int* ghist;        // [bs][cams*L][kFine]  samples per fine bin of the bucket (zeroed before the compaction kernel)
    int4* rec;         // [bs][cams*L][n_ids]  sorted records {sample id, packed anchor row, lh, lw}
    float* recw;       // [bs][cams*L][n_ids][G] the G weights of each record's (sample, cam, level), in record order
    int* seg;          // [bs][seg_stride]     segment tables: seg[key] = first record of the bucket with key' >= key
    int* cursor;       // [bs][cams*L]         records of each bucket (written by the bucket's band 0)
    unsigned long long* sortbuf;  // [bs][cams*L][2][n_ids] global ping-pong (bands that exceed shared memory)
    int4* part_list;   // [bs*num_feat + partial_cap][4]  work-list entries (kEntryInts ints), one per part item
    int4* tiny_list;   // [bs*num_feat][4]                work-list entries, one per tiny row (tiny_ok only)
    float* partial;    // [partial_cap][C]     part sums of rows with more than one part
    int* unit_done;    // [partial_cap]        parts finished, indexed by a row's first slot
    int* counters;     // [0] part items  [2] partial slots  [3] tiny rows  [4] reduce queue head  (zeroed by the compaction kernel)
    int partial_cap;
    Dims d;
    int seg_stride;    // ints per batch element in seg
    int n_chunks;
    int NB;            // band CTAs launched per bucket (<= kMaxBands)
    int band_target;   // samples per band the cuts aim for
    int accumulate;    // add to the rows already in g_feat instead of overwriting them (shared buffer across calls)
    int tiny_max;      // rows with at most this many contributions go to the quarter-warp path (<= kTinyRow)
    int tiny_ok;       // shape supported by the quarter-warp kernel (C % 32 == 0, C <= 256, (C/G) % 32 == 0)
};

// call that owns sample id `sid` (ids of call k start at calls[k].id_begin; at most kMaxGfeatCalls entries)
__device__ __forceinline__ int call_of_id(const GfeatParams& p, int sid) {
    int k = 0;
#pragma unroll
    for (int i = 1; i < kMaxGfeatCalls; ++i)
        if (i < p.ncalls && sid >= p.calls[i].id_begin) k = i;
    return k;
}

__host__ __device__ inline int bits_for(unsigned v) {   // number of bits to represent values < v
#ifdef __CUDA_ARCH__
    return (v <= 1u) ? 0 : 32 - __clz(v - 1u);
#else
    int b = 0;
    while ((1ull << b) < (unsigned long long)v) ++b;
    return b;
#endif
}

// Key space of a (cam, level) bucket of h x w pixels.  Quads are keyed in PADDED coordinates (h_low+1, w_low+1) in
// [0,h] x [0,w], row-major key = row * (w+1) + column; fine bin f owns the keys [f*KF, (f+1)*KF).
struct KeySpace {
    int keys, row_keys, KF;      // (h+1)*(w+1), keys per quad row (w+1), keys per fine bin
};
__device__ __forceinline__ KeySpace key_space(int h, int w) {
    KeySpace g;
    g.row_keys = w + 1;
    g.keys = (h + 1) * g.row_keys;
    g.KF = (g.keys + kFine - 1) / kFine;
    return g;
}
// the key of a sample at a level: the same FMA + floor as quad_setup()
__device__ __forceinline__ int quad_key(float x, float y, int h, int w) {
    const int qr = __float2int_rd(__fmaf_rn(y, (float)h, -0.5f)) + 1;
    const int qc = __float2int_rd(__fmaf_rn(x, (float)w, -0.5f)) + 1;
    return qr * (w + 1) + qc;
}
// words of the bucket's list are 64-bit when (key, sample id) does not fit 32 bits (both kernels evaluate this)
__device__ __forceinline__ bool list_is_wide(int keys, int id_bits) { return bits_for((unsigned)keys) + id_bits > 32; }
// The segment table of bucket `cl` (keys + 1 entries) starts at kSegScale * start[cl]: (h+1)(w+1) + 1 <= 5*h*w for
// every h, w >= 1, so buckets whose row ranges are disjoint never overlap, whatever their order.
__device__ __forceinline__ size_t seg_offset(int start) { return (size_t)kSegScale * (size_t)start; }

// ------------------------------------------------------------------------------------------ zero fill
__global__ void __launch_bounds__(256) dfa_zero_kernel(uint4* __restrict__ dst, long long n16) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {     // four independent 16-byte stores per thread and trip
        dst[i] = z;
        dst[i + stride] = z;
        dst[i + 2 * stride] = z;
        dst[i + 3 * stride] = z;
    }
    for (; i < n16; i += stride) dst[i] = z;
}

// lanes of the warp that hold the same `db`-bit digit (inactive lanes pass a digit nobody else has: bit `db` set).
// One ballot per digit bit; __match_any_sync costs a multiple of this when the warp holds many distinct digits.
__device__ __forceinline__ unsigned digit_peers(unsigned dgt, int db) {
    unsigned peers = 0xffffffffu;
    for (int bit = 0; bit <= db; ++bit) {
        const unsigned vote = __ballot_sync(0xffffffffu, (dgt >> bit) & 1u);
        peers &= ((dgt >> bit) & 1u) ? vote : ~vote;
    }
    return peers;
}

// ------------------------------------------------------------------------------------------ compaction
// grid (n_chunks, cams, bs), block kVisThreads.  Per level the visible samples of the chunk
// are written to the bucket's list GROUPED BY FINE BIN (stable: warp order, then order inside the warp -- the sample
// order), so that a band CTA finds the words of its key range as one contiguous run per chunk and never looks at the
// others (walking the whole list in every band CTA was the largest phase of the sort kernel: ~20 instructions per
// word, bands x words times).
__global__ void __launch_bounds__(kVisThreads) dfa_vis_compact_kernel(const GfeatParams p) {
    constexpr int kWarps = kVisThreads / 32;
    constexpr int kPerWarp = kVisChunk / kWarps;     // 128
    constexpr int kIter = kPerWarp / 32;             // 4
    const Dims d = p.d;
    const int c = blockIdx.x, cam = blockIdx.y, b_idx = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kc = call_of_id(p, c * kVisChunk);          // a chunk belongs to exactly one call
    const GfeatCall& gc = p.calls[kc];
    const int AP = gc.A * gc.P;                           // samples of that call (per batch element)
    __shared__ __align__(16) int s_wh[kWarps * kFine];    // per warp and fine bin: samples, then first slot inside the bin
    __shared__ __align__(16) int s_cum[kFine];            // exclusive prefix over the bins of the level at hand

    if (c == 0 && cam == 0 && b_idx == 0 && tid < 8) p.counters[tid] = 0;
    const float2* loc2 = reinterpret_cast<const float2*>(gc.loc) + (size_t)b_idx * AP * d.cams + cam;
    const int s_base = c * kVisChunk + warp * kPerWarp;   // group sample id
    const int s_local0 = s_base - gc.id_begin;            // sample index inside the call
    float2 xy[kIter];
    unsigned vis = 0;                                     // bit `it`: this lane's sample of iteration `it` is visible
    bool any = false;
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
        const int s = s_local0 + it * 32 + lane;
        xy[it] = make_float2(-1.f, -1.f);
        if (s < AP) xy[it] = __ldg(loc2 + (size_t)s * d.cams);
        if (loc_valid(xy[it].x, xy[it].y)) vis |= 1u << it;
    }
    any = __syncthreads_or(vis != 0);
    const int id_bits = bits_for((unsigned)p.n_ids);
    int* cum = p.chunk_cum + (((size_t)b_idx * d.cams + cam) * p.n_chunks + c) * d.L * kCum;
    if (!any) {      // nothing of this chunk is visible to the camera (most chunks of the side and rear cameras)
        for (int i = tid; i < d.L * kCum; i += kVisThreads) cum[i] = 0;
        return;
    }
    const unsigned lt = (1u << lane) - 1u;
    for (int l = 0; l < d.L; ++l) {
        const int cl = cam * d.L + l;
        const int h = __ldg(p.shapes + cl * 2), w = __ldg(p.shapes + cl * 2 + 1);
        const KeySpace g = key_space(h, w);
        for (int i = tid; i < kWarps * kFine; i += kVisThreads) s_wh[i] = 0;
        __syncthreads();
        // rank of every visible sample among the samples of its warp that fall into the same bin
        int key[kIter], rank[kIter], fbin[kIter];
        int* my = s_wh + warp * kFine;
#pragma unroll
        for (int it = 0; it < kIter; ++it) {
            const bool v = (vis >> it) & 1u;
            key[it] = v ? quad_key(xy[it].x, xy[it].y, h, w) : 0;
            const unsigned bin = v ? (unsigned)(key[it] / g.KF) : (unsigned)kFine;
            fbin[it] = (int)bin;
            const unsigned peers = digit_peers(bin, 7);
            rank[it] = v ? my[bin] + __popc(peers & lt) : 0;
            __syncwarp();
            if (v && lane == __ffs(peers) - 1) my[bin] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        int count = 0;
        if (tid < kFine) {     // bin totals; s_wh becomes the first slot of (warp, bin) inside the bin
#pragma unroll
            for (int wq = 0; wq < kWarps; ++wq) {
                const int n = s_wh[wq * kFine + tid];
                s_wh[wq * kFine + tid] = count;
                count += n;
            }
            s_cum[tid] = count;
        }
        __syncthreads();
        if (warp == 0) {       // exclusive prefix over the bins (4 per lane), the chunk's prefix row, the bucket's histogram
            const int4 v = *reinterpret_cast<const int4*>(s_cum + lane * 4);
            const int mine = v.x + v.y + v.z + v.w;
            int inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            const int ex = inc - mine;
            const int4 e4 = make_int4(ex, ex + v.x, ex + v.x + v.y, ex + v.x + v.y + v.z);
            *reinterpret_cast<int4*>(s_cum + lane * 4) = e4;
            *reinterpret_cast<int4*>(cum + l * kCum + lane * 4) = e4;
            if (lane == 31) cum[l * kCum + kFine] = inc;
            if (mine > 0) {
                int* gh = p.ghist + ((size_t)b_idx * d.cams * d.L + cl) * kFine + lane * 4;
                if (v.x) atomicAdd(gh + 0, v.x);
                if (v.y) atomicAdd(gh + 1, v.y);
                if (v.z) atomicAdd(gh + 2, v.z);
                if (v.w) atomicAdd(gh + 3, v.w);
            }
        }
        __syncthreads();
        const bool wide = list_is_wide(g.keys, id_bits);
        unsigned long long* slots = p.vis_key + ((size_t)b_idx * d.cams * d.L + cl) * p.n_ids;
        unsigned long long* l64 = slots + (size_t)c * kVisChunk;
        unsigned* l32 = reinterpret_cast<unsigned*>(slots) + (size_t)c * kVisChunk;
        // rank[] was taken before s_wh turned into first slots: slot = bin start + the warp's first slot in the bin + rank
#pragma unroll
        for (int it = 0; it < kIter; ++it) {
            if ((vis >> it) & 1u) {
                const int bin = fbin[it];
                const int at = s_cum[bin] + my[bin] + rank[it];
                const unsigned id = (unsigned)(s_base + it * 32 + lane);
                if (wide) l64[at] = ((unsigned long long)key[it] << id_bits) | id;
                else l32[at] = ((unsigned)key[it] << id_bits) | id;
            }
        }
        __syncthreads();       // s_wh / s_cum are rewritten by the next level
    }
}

// ------------------------------------------------------------------------------------------ band sort
// Stable LSD radix sort of n packed words on bits [lo_bit, lo_bit + nbits) by one CTA of kT threads.
// a/b: ping-pong arrays (shared or global).  Returns the array holding the result.
// The nbits are split evenly over ceil(nbits / 8) passes (10 bits -> 5 + 5, not 8 + 2): what a pass costs besides the
// two walks over the words is proportional to its digit count (per-warp histograms: zero, column scan, base add).
template <typename W, int kT>
__device__ W* block_radix_sort(W* a, W* b, int n, int lo_bit, int nbits, unsigned* hist /*[kSortWarps][kRadix]*/,
                               unsigned* tot /*[kRadix + 32]*/) {
    constexpr int kSortWarps = kT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int chunk = (n + kSortWarps - 1) / kSortWarps;
    chunk = (chunk + 31) & ~31;
    const int beg = min(n, warp * chunk), end = min(n, beg + chunk);
    const int passes = (nbits + kRadixBits - 1) / kRadixBits;
    int shift = lo_bit;
    for (int pass = 0; pass < passes; ++pass) {
        const int left = lo_bit + nbits - shift;
        const int db = (left + (passes - pass) - 1) / (passes - pass);     // digit bits of this pass (<= kRadixBits)
        const int R = 1 << db;
        const unsigned dmask = (unsigned)R - 1u;
        unsigned* my = hist + warp * R;
        for (int i = lane; i < R; i += 32) my[i] = 0;
        __syncwarp();
        for (int i0 = beg; i0 < end; i0 += 32) {
            const int i = i0 + lane;
            const bool has = i < end;
            const unsigned dgt = has ? ((unsigned)(a[i] >> shift) & dmask) : (unsigned)R;
            const unsigned peers = digit_peers(dgt, db);
            if (has && lane == __ffs(peers) - 1) my[dgt] += __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        if (tid < R) {                      // digit totals over the warps
            unsigned run = 0;
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) run += hist[w * R + tid];
            tot[tid] = run;
        }
        __syncthreads();
        if (warp == 0) {                    // exclusive scan of the R totals: lane owns R/32 consecutive digits
            const int per = (R + 31) >> 5;
            unsigned mine = 0;
            for (int j = 0; j < per; ++j) {
                const int dg = lane * per + j;
                if (dg < R) mine += tot[dg];
            }
            unsigned inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            unsigned run = inc - mine;
            for (int j = 0; j < per; ++j) {
                const int dg = lane * per + j;
                if (dg < R) {
                    const unsigned c = tot[dg];
                    tot[dg] = run;
                    run += c;
                }
            }
        }
        __syncthreads();
        if (tid < R) {                      // first slot of (warp, digit): digit base + the warps before
            unsigned run = tot[tid];
#pragma unroll
            for (int w = 0; w < kSortWarps; ++w) {
                const unsigned c = hist[w * R + tid];
                hist[w * R + tid] = run;
                run += c;
            }
        }
        __syncthreads();
        for (int i0 = beg; i0 < end; i0 += 32) {
            const int i = i0 + lane;
            const bool has = i < end;
            W word = 0;
            unsigned dgt = (unsigned)R;
            if (has) {
                word = a[i];
                dgt = (unsigned)(word >> shift) & dmask;
            }
            const unsigned peers = digit_peers(dgt, db);
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
        shift += db;
    }
    return a;
}

struct BandCtx {
    int b_idx, cam, cl, h, w, q0, q1, K, n, base, vb, kb, keys;     // [q0, q1): the band's key range
    const int* s_coff;    // per chunk: offset of its first in-band sample inside the band's list ([n_chunks] = n)
    const int* s_from;    // per chunk: where the band's run starts inside the chunk's words
    int tr;               // trace slot of this CTA (development builds)
};

// gather (chunk order, then order inside the chunk) + sort + record/segment emission of one band.
// LW: word type of the bucket's list, W: word type of the band's sort (local key << id bits | id)
template <typename W, typename LW, int kT>
__device__ void band_sort_body(const GfeatParams& p, const BandCtx& bc, W* a, W* b, unsigned* hist, unsigned* tot,
                               int* seg_bucket) {
    constexpr int kSortThreads = kT, kSortWarps = kT / 32;
    const Dims d = p.d;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int AP = p.n_ids;
    const LW* list = reinterpret_cast<const LW*>(p.vis_key + ((size_t)bc.b_idx * d.cams * d.L + bc.cl) * AP);

    // the band's words are one contiguous run per chunk (the compaction kernel groups a chunk's words by fine bin):
    // warp w copies the runs of chunks w, w + kSortWarps, ..., four chunks' loads in flight at a time, and turns the
    // bucket-wide key into the band's local one
    {
        const LW sub = (LW)bc.q0 << bc.vb;
        for (int c0 = warp; c0 < p.n_chunks; c0 += 4 * kSortWarps) {
            LW wd[4];
            int cnt[4], dst[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = c0 + u * kSortWarps;
                cnt[u] = 0;
                dst[u] = 0;
                wd[u] = 0;
                if (c < p.n_chunks) {
                    dst[u] = bc.s_coff[c];
                    cnt[u] = bc.s_coff[c + 1] - dst[u];
                    if (lane < cnt[u]) wd[u] = __ldg(list + (size_t)c * kVisChunk + bc.s_from[c] + lane);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (lane < cnt[u]) a[dst[u] + lane] = (W)(wd[u] - sub);
                if (cnt[u] > 32) {     // warp-uniform; a chunk rarely holds more than 32 words of one band
                    const int c = c0 + u * kSortWarps;
                    const LW* src = list + (size_t)c * kVisChunk + bc.s_from[c];
                    for (int i = 32 + lane; i < cnt[u]; i += 32) a[dst[u] + i] = (W)(__ldg(src + i) - sub);
                }
            }
        }
    }
    __syncthreads();
    DFA_TRACE(bc.tr, 4);

    W* sorted = block_radix_sort<W, kT>(a, b, bc.n, bc.vb, bc.kb, hist, tot);
    DFA_TRACE(bc.tr, 5);

    // sorted records: everything the reduce needs per contribution, so it never re-derives the quad
    int4* rec = p.rec + ((size_t)bc.b_idx * d.cams * d.L + bc.cl) * AP + bc.base;
    float* recw = p.recw + (((size_t)bc.b_idx * d.cams * d.L + bc.cl) * AP + bc.base) * d.G;
    const int lvl = bc.cl - bc.cam * d.L;
    const W vmask = ((W)1 << bc.vb) - 1;
    constexpr int kEmit = 4;      // records per thread whose loads are in flight together
    for (int i0 = 0; i0 < bc.n; i0 += kSortThreads * kEmit) {
        int sid[kEmit], arow[kEmit];
        float2 xy[kEmit];
#pragma unroll
        for (int u = 0; u < kEmit; ++u) {
            const int i = i0 + u * kSortThreads + tid;
            sid[u] = (i < bc.n) ? (int)(sorted[i] & vmask) : p.calls[0].id_begin;
            const GfeatCall& gc = p.calls[call_of_id(p, sid[u])];
            const int s_local = sid[u] - gc.id_begin;
            const size_t s_abs = (size_t)bc.b_idx * gc.A * gc.P + s_local;      // sample inside the call's tensors
            xy[u] = __ldg(reinterpret_cast<const float2*>(gc.loc) + s_abs * d.cams + bc.cam);
            arow[u] = gc.anchor_begin + s_local / gc.P;
        }
#pragma unroll
        for (int u = 0; u < kEmit; ++u) {
            const int i = i0 + u * kSortThreads + tid;
            if (i < bc.n) {
                const Quad q = quad_setup(xy[u].x, xy[u].y, bc.h, bc.w);
                rec[i] = make_int4(sid[u], arow[u], __float_as_int(q.lh), __float_as_int(q.lw));
            }
        }
    }
    // the record's G weights travel with it, so the reduce reads them sequentially and never has to find the call a
    // contribution came from: one thread per 16 bytes, stores in record order (coalesced)
    if ((d.G & 3) == 0) {
        const int q_per = d.G >> 2, total = bc.n * q_per;
        float4* dst = reinterpret_cast<float4*>(recw);
        for (int e0 = 0; e0 < total; e0 += kSortThreads * kEmit) {
            float4 v[kEmit];
#pragma unroll
            for (int u = 0; u < kEmit; ++u) {
                const int e = min(e0 + u * kSortThreads + tid, total - 1);
                const int i = e / q_per, part = e - i * q_per;
                const int sid = (int)(sorted[i] & vmask);
                const GfeatCall& gc = p.calls[call_of_id(p, sid)];
                const size_t s_abs = (size_t)bc.b_idx * gc.A * gc.P + (sid - gc.id_begin);
                v[u] = __ldg(reinterpret_cast<const float4*>(gc.weights + ((s_abs * d.cams + bc.cam) * d.L + lvl) * d.G) + part);
            }
#pragma unroll
            for (int u = 0; u < kEmit; ++u) {
                const int e = e0 + u * kSortThreads + tid;
                if (e < total) dst[e] = v[u];
            }
        }
    } else {
        for (int e = tid; e < bc.n * d.G; e += kSortThreads) {
            const int i = e / d.G, g = e - i * d.G;
            const int sid = (int)(sorted[i] & vmask);
            const GfeatCall& gc = p.calls[call_of_id(p, sid)];
            const size_t s_abs = (size_t)bc.b_idx * gc.A * gc.P + (sid - gc.id_begin);
            recw[e] = __ldg(gc.weights + ((s_abs * d.cams + bc.cam) * d.L + lvl) * d.G + g);
        }
    }
    DFA_TRACE(bc.tr, 6);

    // seg[q0 + k] = first record (position inside the bucket) whose key >= q0 + k, k = 0..K-1: a lower-bound search in
    // the sorted words per key -- no fill / mark / scan phases, no barriers.  The band that owns the bucket's last key
    // also writes the end marker seg[keys].
    // record i is the first of key k_i: the keys (k_{i-1}, k_i] start at i.  Keys up to the first record's start at 0,
    // keys after the last record's at n (the band that owns the bucket's last key also writes the end marker seg[keys]).
    const int k_end = bc.K + ((bc.q1 == bc.keys) ? 1 : 0);
    const int k_first = (int)(sorted[0] >> bc.vb), k_last = (int)(sorted[bc.n - 1] >> bc.vb);   // n > 0 (checked by the kernel)
    int* seg_band = seg_bucket + bc.q0;
    for (int k = tid; k <= k_first; k += kSortThreads) seg_band[k] = bc.base;
    for (int k = k_last + 1 + tid; k < k_end; k += kSortThreads) seg_band[k] = bc.base + bc.n;
    for (int i0 = warp * 32; i0 < bc.n; i0 += kSortThreads) {
        const int i = i0 + lane;
        int key = 0, gap = 0;
        if (i > 0 && i < bc.n) {
            key = (int)(sorted[i] >> bc.vb);
            gap = key - (int)(sorted[i - 1] >> bc.vb);
        }
        // short gaps (the usual case: 0 = same key, 1 = next key) by the record's own thread, long ones by the warp
        for (int t = 0; t < min(gap, 4); ++t) seg_band[key - t] = bc.base + i;
        unsigned big = __ballot_sync(0xffffffffu, gap > 4);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const int kk = __shfl_sync(0xffffffffu, key, src), gg = __shfl_sync(0xffffffffu, gap, src);
            for (int t = 4 + lane; t < gg; t += 32) seg_band[kk - t] = bc.base + i0 + src;
        }
    }
    DFA_TRACE(bc.tr, 7);
    DFA_TRACE_V(bc.tr, 8, bc.n);
    DFA_TRACE_END(bc.tr, 9);
}

// dynamic smem: hist[warps*kRadix] + tot[kRadix+32] + misc[16] + fine[kFine+4] + coff[n_chunks] + from[n_chunks] + 2*cap words
inline size_t band_sort_smem_bytes(int n_chunks, int threads, int cap) {
    return (size_t)((threads / 32) * kRadix + kRadix + 32 + 16 + (kFine + 4) + 2 * ((n_chunks + 4) & ~3)) * 4 + (size_t)cap * 2 * 4;
}

// grid (NB, cams*L, bs), block kT threads (>= 256: the digit scan uses the first 8 warps); kCap packed words per
// ping-pong half are kept in shared memory, longer bands sort in global memory
template <int kT, int kCap>
__global__ void __launch_bounds__(kT, (kT >= 512) ? 2 : 4) dfa_band_sort_kernel(const GfeatParams p) {
    constexpr int kSortThreads = kT, kSortWarps = kT / 32, kBandCap = kCap;
    const Dims d = p.d;
    const int band = blockIdx.x, cl = blockIdx.y, b_idx = blockIdx.z;
    const int cam = cl / d.L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int AP = p.n_ids;
    BandCtx bc;
    bc.tr = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    DFA_TRACE_BEGIN(bc.tr);
    DFA_TRACE(bc.tr, 2);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned* hist = reinterpret_cast<unsigned*>(smem_raw);
    unsigned* tot = hist + kSortWarps * kRadix;
    int* s_misc = reinterpret_cast<int*>(tot + kRadix + 32);     // [0] f0  [1] f1  [2] last bin with samples  [4..7] warp totals of the fine scan
    int* s_fine = s_misc + 16;                                   // [kFine + 1] exclusive prefix of the bucket's histogram
    int* s_coff = s_fine + kFine + 4;      // the two per-chunk arrays are (n_chunks + 4) & ~3 ints apart
    int* s_from = s_coff + ((p.n_chunks + 4) & ~3);
    unsigned char* data = reinterpret_cast<unsigned char*>(s_from + ((p.n_chunks + 4) & ~3));

    // the bucket's histogram -> exclusive prefix E[0..kFine] (E[kFine] = samples of the bucket)
    {
        int mine = 0, inc = 0;
        if (tid < kFine) mine = __ldg(p.ghist + ((size_t)b_idx * d.cams * d.L + cl) * kFine + tid);
        // A bucket without samples (rear cameras of the planning query, every camera of the ego query) stays empty:
        // cursor == 0 is what the row classification checks before it reads any segment table
        if (!__syncthreads_or(mine > 0)) {
            if (band == 0 && tid == 0) p.cursor[(size_t)b_idx * d.cams * d.L + cl] = 0;
            return;
        }
        if (tid < kFine) {
            inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) s_misc[4 + warp] = inc;
        }
        __syncthreads();
        if (tid < kFine) {
            int before = 0;
#pragma unroll
            for (int w = 0; w < kFine / 32; ++w)
                if (w < warp) before += s_misc[4 + w];
            s_fine[tid] = before + inc - mine;
            if (tid == kFine - 1) s_fine[kFine] = before + inc;
        }
        __syncthreads();
    }
    const int N = s_fine[kFine];
    if (band == 0 && tid == 0) p.cursor[(size_t)b_idx * d.cams * d.L + cl] = N;
    // bands of ~band_target samples, cut between fine bins: band j starts at the first bin whose prefix reaches
    // ceil(j*N/nb) -- every CTA of the bucket derives the same cuts
    const int nb = max(1, min(p.NB, (N + p.band_target - 1) / p.band_target));
    if (band >= nb) return;
    {
        const long long t0 = ((long long)band * N + nb - 1) / nb, t1 = ((long long)(band + 1) * N + nb - 1) / nb;
        if (tid <= kFine) {
            const int e = s_fine[tid], prev = (tid == 0) ? -1 : s_fine[tid - 1];
            if (e >= t0 && prev < t0) s_misc[0] = tid;
            if (e >= t1 && prev < t1) s_misc[1] = tid;
        }
        __syncthreads();
    }
    const int f0 = s_misc[0], f1 = (band == nb - 1) ? kFine : s_misc[1];
    // last fine bin of the band that holds a sample: the sort only has to tell the keys up to there apart
    if (tid == 0) s_misc[2] = f0;
    __syncthreads();
    if (tid >= f0 && tid < f1 && s_fine[tid + 1] > s_fine[tid]) atomicMax(&s_misc[2], tid);
    __syncthreads();
    const int h = __ldg(p.shapes + cl * 2), w = __ldg(p.shapes + cl * 2 + 1);
    const KeySpace g = key_space(h, w);
    bc.b_idx = b_idx; bc.cam = cam; bc.cl = cl; bc.h = h; bc.w = w; bc.keys = g.keys;
    bc.q0 = min(g.keys, f0 * g.KF);
    bc.q1 = min(g.keys, f1 * g.KF);
    if (bc.q0 >= bc.q1) return;        // no key maps to this band (then it holds no sample either)
    bc.K = bc.q1 - bc.q0;
    bc.vb = bits_for((unsigned)AP);
    bc.kb = bits_for((unsigned)min(bc.K, (s_misc[2] + 1 - f0) * g.KF));
    bc.base = s_fine[f0];
    bc.n = s_fine[f1] - s_fine[f0];
    int* seg_bucket = p.seg + (size_t)b_idx * p.seg_stride + seg_offset(__ldg(p.starts + cl));
    if (bc.n == 0) {                   // keys without samples (sky, road ahead of nothing): empty segments at `base`
        for (int k = tid; k < bc.K + ((bc.q1 == bc.keys) ? 1 : 0); k += kSortThreads) seg_bucket[bc.q0 + k] = bc.base;
        return;
    }

    // in-band samples per chunk (two prefix entries of the chunk's histogram), then their exclusive scan
    {
        const int l = cl - cam * d.L;
        const int* cum = p.chunk_cum + (((size_t)b_idx * d.cams + cam) * p.n_chunks * d.L + l) * kCum;
        for (int c = tid; c < p.n_chunks; c += kSortThreads) {
            const int* row = cum + (size_t)c * d.L * kCum;
            const int from = __ldg(row + f0);
            s_coff[c] = __ldg(row + f1) - from;
            s_from[c] = from;
        }
    }
    __syncthreads();
    if (warp == 0) {    // exclusive scan of the in-band counts over the chunks
        int run = 0;
        for (int c0 = 0; c0 < p.n_chunks; c0 += 32) {
            const int c = c0 + lane;
            const int m = (c < p.n_chunks) ? s_coff[c] : 0;
            int inc = m;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (c < p.n_chunks) s_coff[c] = run + inc - m;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) s_coff[p.n_chunks] = run;
    }
    __syncthreads();
    bc.s_coff = s_coff;
    bc.s_from = s_from;
    DFA_TRACE(bc.tr, 3);

    // bands that overflow shared memory sort in their slice [base, base+n) of the bucket's global buffers
    unsigned long long* gbuf = p.sortbuf + ((size_t)b_idx * d.cams * d.L + cl) * 2 * AP;
    const bool wide_list = list_is_wide(g.keys, bc.vb);
    if (bc.vb + bc.kb <= 32) {
        unsigned* a = reinterpret_cast<unsigned*>(data);
        unsigned* b2 = a + kBandCap;
        if (bc.n > kBandCap) {
            a = reinterpret_cast<unsigned*>(gbuf) + bc.base;
            b2 = a + AP;
        }
        if (wide_list) band_sort_body<unsigned, unsigned long long, kT>(p, bc, a, b2, hist, tot, seg_bucket);
        else band_sort_body<unsigned, unsigned, kT>(p, bc, a, b2, hist, tot, seg_bucket);
    } else {
        unsigned long long* a = reinterpret_cast<unsigned long long*>(data);
        unsigned long long* b2 = a + kBandCap / 2;
        if (bc.n > kBandCap / 2) {
            a = gbuf + bc.base;
            b2 = gbuf + AP + bc.base;
        }
        band_sort_body<unsigned long long, unsigned long long, kT>(p, bc, a, b2, hist, tot, seg_bucket);
    }
}

// ------------------------------------------------------------------------------------------ reduce
constexpr int kEntryInts = 16;   // work-list entry, shared by the part and the tiny list:
//   [0] b*num_feat + row   [1] n (contributions)   [2] b   [3] cam | level << 8 | bucket << 16
//   [4..7] beg[4]          [8..10] e1, e2, e3      [12] part | parts << 16   [13] first partial slot
//   [14], [15] contribution range [c_lo, c_hi) of the part

template <int V, int NCH>
struct RowCtx {
    const int4* rec;              // sorted records of the bucket
    const char* wts_lane[NCH];    // record weights of the bucket at the group of this lane's chunk j, as bytes
    const char* gout_lane[NCH];   // grad_out of batch element b at this lane's channels of chunk j, as bytes
    unsigned c_bytes, w_stride_bytes;
    int beg[4];                   // begin of the corner-1/2/3/4 segments in rec[]
    int e1, e2, e3;               // cumulative ends of the first three segments
};

// per-lane channel ownership, fixed for the whole kernel
template <int V, int NCH>
struct LaneMap {
    int ch[NCH], grp[NCH];
    bool act[NCH];
    __device__ __forceinline__ void init(int C, int G) {
        const int lane = threadIdx.x & 31, gd = C / G;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int c_raw = (j * 32 + lane) * V;
            act[j] = c_raw < C;
            ch[j] = act[j] ? c_raw : C - V;    // clamped lanes compute values nobody stores
            grp[j] = ch[j] / gd;
        }
    }
};

template <int V, int NCH>
__device__ __forceinline__ void row_ctx_from_entry(RowCtx<V, NCH>& cx, const LaneMap<V, NCH>& lm, const GfeatParams& p,
                                                   int b_idx, int packed) {
    const Dims& d = p.d;
    const int rcl = packed >> 16;
    const size_t AP = (size_t)p.n_ids;
    cx.rec = p.rec + ((size_t)b_idx * d.cams * d.L + rcl) * AP;
    const float* wts = p.recw + ((size_t)b_idx * d.cams * d.L + rcl) * AP * d.G;
    const float* gout = p.grad_out + (size_t)b_idx * p.A_total * d.C;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        cx.wts_lane[j] = reinterpret_cast<const char*>(wts + lm.grp[j]);
        cx.gout_lane[j] = reinterpret_cast<const char*>(gout + lm.ch[j]);
    }
}

// accumulate contributions [c_lo, c_hi) of one row, indices in the row's concatenated
// (corner1, corner2, corner3, corner4) order -- the fixed summation order of this library.
// Contributions are consumed in batches whose loads are all issued before the first FMA; a short last
// batch is padded with copies of the range's last contribution at coefficient 0, so a typical row (a
// handful of contributions) costs one memory round trip instead of one per contribution.  Addresses are
// one 64-bit base per tensor plus 32-bit byte offsets (the loop is issue-bound otherwise).
template <int V, int NCH, int kBatchCap>
__device__ __forceinline__ void accumulate_row(const RowCtx<V, NCH>& cx, int c_lo, int c_hi, float (&acc)[NCH][V]) {
    // contributions whose loads are in flight together: 64 data registers per lane (kBatchCap = 4: 32, for the
    // light variant of the kernel that keeps 4 CTAs per SM resident)
    constexpr int kRegsPer = NCH * (V + 1);
    constexpr int kBatch0 = (kRegsPer <= 10) ? 8 : (kRegsPer <= 20) ? 4 : (kRegsPer <= 40) ? 2 : 1;
    constexpr int kReduceBatch = kBatch0 < kBatchCap ? kBatch0 : kBatchCap;
    constexpr bool kConstChunks = (V > 1);   // vector path: chunk j sits j*32*V channels after chunk 0
    const int lane = threadIdx.x & 31;
    for (int v0 = c_lo; v0 < c_hi; v0 += 32) {
        // lane-parallel metadata for up to 32 contributions (lanes past the end clamp onto the last one)
        const int v = min(v0 + lane, c_hi - 1);
        const int k = (v >= cx.e1) + (v >= cx.e2) + (v >= cx.e3);
        const int first = (k == 0) ? 0 : (k == 1) ? cx.e1 : (k == 2) ? cx.e2 : cx.e3;
        const int bk = (k == 0) ? cx.beg[0] : (k == 1) ? cx.beg[1] : (k == 2) ? cx.beg[2] : cx.beg[3];
        const int pos = bk + (v - first);                             // record index inside the bucket
        const int4 e = __ldg(cx.rec + pos);
        const float lh = __int_as_float(e.z), lw = __int_as_float(e.w);
        const float hh = 1.f - lh, hw = 1.f - lw;    // same expressions as quad_setup()
        float coef = (k == 0) ? hh * hw : (k == 1) ? hh * lw : (k == 2) ? lh * hw : lh * lw;
        if (v0 + lane >= c_hi) coef = 0.f;
        const unsigned go_off = (unsigned)e.y * cx.c_bytes;           // < 2^32 checked on the host
        const unsigned w_off = (unsigned)pos * cx.w_stride_bytes;
        const int cnt = min(32, c_hi - v0);
        for (int m0 = 0; m0 < cnt; m0 += kReduceBatch) {
            float g[kReduceBatch][NCH][V];
            float wg[kReduceBatch][NCH];
#pragma unroll
            for (int u = 0; u < kReduceBatch; ++u) {
                const int m = min(m0 + u, 31);
                const unsigned go_m = __shfl_sync(0xffffffffu, go_off, m);
                const unsigned w_m = __shfl_sync(0xffffffffu, w_off, m);
                const float cf = __shfl_sync(0xffffffffu, coef, m);
                const char* g0 = cx.gout_lane[0] + go_m;
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const char* gp = kConstChunks ? g0 + j * 32 * V * 4 : cx.gout_lane[j] + go_m;
                    VecIO<float, V>::load(reinterpret_cast<const float*>(gp), g[u][j]);
                    wg[u][j] = __ldg(reinterpret_cast<const float*>(cx.wts_lane[j] + w_m)) * cf;
                }
            }
#pragma unroll
            for (int u = 0; u < kReduceBatch; ++u)
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    if constexpr (V % 2 == 0) {
#pragma unroll
                        for (int q = 0; q < V; q += 2) {
                            const float2 r = __ffma2_rn(make_float2(wg[u][j], wg[u][j]),
                                                        make_float2(g[u][j][q], g[u][j][q + 1]),
                                                        make_float2(acc[j][q], acc[j][q + 1]));
                            acc[j][q] = r.x;
                            acc[j][q + 1] = r.y;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < V; ++q) acc[j][q] = __fmaf_rn(wg[u][j], g[u][j][q], acc[j][q]);
                    }
                }
        }
    }
}

// row (y,x) of a bucket: the four quad-key segments that contribute to it; a key's segment is [seg[k], seg[k+1])
__device__ __forceinline__ int row_segments(const int* __restrict__ seg_bucket, int row_keys, int y, int x,
                                            int (&beg)[4], int& e1, int& e2, int& e3) {
    // quad (y,x) -> padded (y+1, x+1): this row is its corner 1; (y, x-1) -> (y+1, x): corner 2;
    // (y-1, x) -> (y, x+1): corner 3; (y-1, x-1) -> (y, x): corner 4.  Keys k1-1, k1 and k3-1, k3 are neighbours.
    const int k1 = (y + 1) * row_keys + x + 1, k3 = y * row_keys + x + 1;
    const int a0 = __ldg(seg_bucket + k1 - 1), a1 = __ldg(seg_bucket + k1), a2 = __ldg(seg_bucket + k1 + 1);
    const int c0 = __ldg(seg_bucket + k3 - 1), c1 = __ldg(seg_bucket + k3), c2 = __ldg(seg_bucket + k3 + 1);
    beg[0] = a1; beg[1] = a0; beg[2] = c1; beg[3] = c0;
    // accumulation order: corner 1, corner 2, corner 3, corner 4
    e1 = a2 - a1;
    e2 = e1 + (a1 - a0);
    e3 = e2 + (c2 - c1);
    return e3 + (c1 - c0);
}

// grid (ceil(num_feat / kClassifyThreads), bs), block kClassifyThreads
__global__ void __launch_bounds__(kClassifyThreads) dfa_row_classify_kernel(const GfeatParams p) {
    const Dims d = p.d;
    const int tid = threadIdx.x, lane = tid & 31;
    const int b_idx = blockIdx.y;
    const int n_cl = d.cams * d.L;
    __shared__ int s_tab[kMaxCamLevels * 3];
    load_level_table(s_tab, p.shapes, p.starts, n_cl);
    __syncthreads();

    const int row = blockIdx.x * kClassifyThreads + tid;
    int n = 0, cl = -1;
    int beg[4] = {0, 0, 0, 0}, e1 = 0, e2 = 0, e3 = 0;
    if (row < d.num_feat) {
        for (int i = 0; i < n_cl; ++i) {
            const int st = s_tab[i * 3 + 2];
            if (row >= st && row < st + s_tab[i * 3] * s_tab[i * 3 + 1]) cl = i;
        }
        if (cl >= 0) {
            const int w = s_tab[cl * 3 + 1], st = s_tab[cl * 3 + 2];
            const int r = row - st, y = r / w, x = r - y * w;
            if (__ldg(p.cursor + (size_t)b_idx * n_cl + cl) > 0)     // empty bucket: its table was never written
                n = row_segments(p.seg + (size_t)b_idx * p.seg_stride + seg_offset(st), w + 1, y, x, beg, e1, e2, e3);
        }
    }
    const bool tiny = p.tiny_ok && n > 0 && n <= p.tiny_max;
    const bool single = n > 0 && !tiny && n <= kPart, multi = n > kPart;
    const int cam = (cl >= 0) ? cl / d.L : 0;
    const int4 q0 = make_int4(b_idx * d.num_feat + row, n, b_idx, cam | ((cl - cam * d.L) << 8) | (cl << 16));
    const int4 q1 = make_int4(beg[0], beg[1], beg[2], beg[3]);
    const int4 q2 = make_int4(e1, e2, e3, 0);
    const unsigned lt = (1u << lane) - 1u;
    // tiny rows and single-part rows: warp-aggregated appends, one integer atomic per warp and list
    const unsigned bt = __ballot_sync(0xffffffffu, tiny), bs1 = __ballot_sync(0xffffffffu, single);
    int base_t = 0, base_s = 0;
    if (lane == 0) {
        if (bt) base_t = atomicAdd(p.counters + 3, __popc(bt));
        if (bs1) base_s = atomicAdd(p.counters + 0, __popc(bs1));
    }
    base_t = __shfl_sync(0xffffffffu, base_t, 0);
    base_s = __shfl_sync(0xffffffffu, base_s, 0);
    if (tiny) {
        int4* e = p.tiny_list + (size_t)(base_t + __popc(bt & lt)) * (kEntryInts / 4);
        e[0] = q0; e[1] = q1; e[2] = q2;
    }
    if (single) {
        int4* e = p.part_list + (size_t)(base_s + __popc(bs1 & lt)) * (kEntryInts / 4);
        e[0] = q0; e[1] = q1; e[2] = q2;
        e[3] = make_int4(0 | (1 << 16), -1, 0, n);
    }
    // rows with more than kPart contributions: one entry per part, partial-sum slots for the parts
    if (multi) {
        int parts = (n + kPart - 1) / kPart;
        int slot = atomicAdd(p.counters + 2, parts);
        if (slot + parts > p.partial_cap) {   // out of partial slots (pathological inputs): one warp sums the whole row
            parts = 1;
            slot = -1;
        } else {
            p.unit_done[slot] = 0;
        }
        const int at = atomicAdd(p.counters + 0, parts);
        for (int u = 0; u < parts; ++u) {
            int4* e = p.part_list + (size_t)(at + u) * (kEntryInts / 4);
            e[0] = q0; e[1] = q1; e[2] = q2;
            e[3] = make_int4(u | (parts << 16), slot, u * kPart, (parts == 1) ? n : min(n, (u + 1) * kPart));
        }
    }
}

// One reduce kernel, persistent warps, ONE dynamic queue over all work items (integer atomic; which warp takes
// which item does not influence any result):
//   part items  a range of <= kPart contributions of one row, summed by a warp in the row's fixed order.  Rows
//               with a single part are written directly; parts of longer rows go to partial-sum slots and the
//               warp that finishes a row's last part adds the slots in part order and writes the row.
//   tiny items  four rows with <= kTinyRow contributions each, one per QUARTER warp (lane s of a quarter owns
//               channels (i*8+s)*4.. of every 32-channel chunk i < NQ = C/32; (C/G) % 32 == 0).  Such rows are
//               pure latency (entry -> records -> grad_out rows); the lever is rows in flight per SM.
// The next item's index and its list entry are fetched while the current item is processed.
template <typename T, int V, int NCH, int NQ, int kMinCtas>
__global__ void __launch_bounds__(256, kMinCtas) dfa_gfeat_reduce_kernel(const GfeatParams p) {
    constexpr int kBatchCap = (kMinCtas >= 3) ? 4 : 8;
    const Dims d = p.d;
    const int lane = threadIdx.x & 31, sub = lane & 7, quarter = lane >> 3;
    const int n_parts = p.counters[0];
    const int n_tiny = (NQ > 0) ? p.counters[3] : 0;
    const int total = n_parts + (n_tiny + 3) / 4;
    const int n_cl = d.cams * d.L;
    const size_t AP = (size_t)p.n_ids;

    LaneMap<V, NCH> lm;
    lm.init(d.C, d.G);
    RowCtx<V, NCH> cx;
    cx.c_bytes = (unsigned)d.C * 4u;
    cx.w_stride_bytes = (unsigned)d.G * 4u;       // one record's weights
    const int* plist = reinterpret_cast<const int*>(p.part_list);
    const int* tlist = reinterpret_cast<const int*>(p.tiny_list);

    // entry words of item `idx`: part items -> lane j holds word j (j < 16); tiny items -> the lanes of quarter
    // q hold words sub and sub + 8 of row 4*t + q
    auto fetch = [&](int idx, int& fa, int& fb) {
        fa = 0; fb = 0;
        if (idx < n_parts) {
            if (lane < kEntryInts) fa = __ldg(plist + (size_t)idx * kEntryInts + lane);
        } else if (idx < total) {
            const int r = (idx - n_parts) * 4 + quarter;
            if (r < n_tiny) {
                fa = __ldg(tlist + (size_t)r * kEntryInts + sub);
                fb = __ldg(tlist + (size_t)r * kEntryInts + 8 + sub);
            }
        }
    };

    // One integer atomic per item on the queue head; the next item's index is requested and its list entry fetched
    // while the current item is processed.  (Measured alternatives, both slower on the stage-2 det call: grabs of
    // 4 items -- coarser balance, 54 vs 46 us; three atomics in flight per warp -- the single-address atomic
    // throughput, ~2.4 G/s, becomes the bottleneck, 50 us.)
    int cur = 0, nxt = 0;
    if (lane == 0) {
        cur = atomicAdd(p.counters + 4, 1);
        nxt = atomicAdd(p.counters + 4, 1);
    }
    cur = __shfl_sync(0xffffffffu, cur, 0);
    int fa, fb;
    fetch(cur, fa, fb);
#pragma unroll 1
    while (cur < total) {
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        int na, nb2;
        fetch(nxt, na, nb2);                                   // in flight while this item is processed
        int after = 0;
        if (lane == 0) after = atomicAdd(p.counters + 4, 1);   // likewise

        if (cur < n_parts) {
            // ------------------------------------------------------------------ part item
            const int grow = __shfl_sync(0xffffffffu, fa, 0);
            row_ctx_from_entry<V, NCH>(cx, lm, p, __shfl_sync(0xffffffffu, fa, 2), __shfl_sync(0xffffffffu, fa, 3));
#pragma unroll
            for (int k = 0; k < 4; ++k) cx.beg[k] = __shfl_sync(0xffffffffu, fa, 4 + k);
            cx.e1 = __shfl_sync(0xffffffffu, fa, 8);
            cx.e2 = __shfl_sync(0xffffffffu, fa, 9);
            cx.e3 = __shfl_sync(0xffffffffu, fa, 10);
            const int up = __shfl_sync(0xffffffffu, fa, 12), slot = __shfl_sync(0xffffffffu, fa, 13);
            const int c_lo = __shfl_sync(0xffffffffu, fa, 14), c_hi = __shfl_sync(0xffffffffu, fa, 15);
            const int u = up & 0xffff, parts = up >> 16;
            float acc[NCH][V];
#pragma unroll
            for (int j = 0; j < NCH; ++j)
#pragma unroll
                for (int e = 0; e < V; ++e) acc[j][e] = 0.f;
            accumulate_row<V, NCH, kBatchCap>(cx, c_lo, c_hi, acc);
            bool write_row = parts == 1;
            if (parts > 1) {
                float* mine = p.partial + (size_t)(slot + u) * d.C;
#pragma unroll
                for (int j = 0; j < NCH; ++j)
                    if (lm.act[j]) VecIO<float, V>::store(mine + lm.ch[j], acc[j]);
                __syncwarp();
                int old = 0;
                if (lane == 0) old = atomic_add_release_gpu(p.unit_done + slot, 1);
                old = __shfl_sync(0xffffffffu, old, 0);
                if (old == parts - 1) {        // every part of this row is in memory: add them in part order
                    __threadfence();
                    write_row = true;
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
#pragma unroll
                        for (int e = 0; e < V; ++e) acc[j][e] = 0.f;
                    // eight part rows in flight at a time (vector loads that bypass L1), added in part order
                    constexpr int kQ = (V % 4 == 0) ? V / 4 : 1;      // float4 pieces of this lane's V channels
                    for (int u0 = 0; u0 < parts; u0 += 8) {
                        float4 t[8][NCH][kQ];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const float* part = p.partial + (size_t)(slot + min(u0 + k, parts - 1)) * d.C;
#pragma unroll
                            for (int j = 0; j < NCH; ++j)
#pragma unroll
                                for (int q = 0; q < kQ; ++q) {
                                    if constexpr (V % 4 == 0) {
                                        t[k][j][q] = __ldcg(reinterpret_cast<const float4*>(part + lm.ch[j]) + q);
                                    } else {
                                        t[k][j][q] = make_float4(0.f, 0.f, 0.f, 0.f);
                                    }
                                }
                        }
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            if (u0 + k >= parts) break;      // warp-uniform
                            const float* part = p.partial + (size_t)(slot + u0 + k) * d.C;
#pragma unroll
                            for (int j = 0; j < NCH; ++j) {
                                if (!lm.act[j]) continue;
                                if constexpr (V % 4 == 0) {
#pragma unroll
                                    for (int q = 0; q < kQ; ++q) {
                                        acc[j][4 * q + 0] += t[k][j][q].x; acc[j][4 * q + 1] += t[k][j][q].y;
                                        acc[j][4 * q + 2] += t[k][j][q].z; acc[j][4 * q + 3] += t[k][j][q].w;
                                    }
                                } else {
#pragma unroll
                                    for (int e = 0; e < V; ++e) acc[j][e] += __ldcg(part + lm.ch[j] + e);
                                }
                            }
                        }
                    }
                }
            }
            if (write_row) {
                T* dst = reinterpret_cast<T*>(p.g_feat) + (size_t)grow * d.C;
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    if (!lm.act[j]) continue;
                    if (p.accumulate) {     // shared g_feat buffer: calls are serialised by the stream, so the order is fixed
                        float old[V];
                        VecIO<T, V>::load(dst + lm.ch[j], old);
#pragma unroll
                        for (int e = 0; e < V; ++e) acc[j][e] += old[e];
                    }
                    VecIO<T, V>::store(dst + lm.ch[j], acc[j]);
                }
            }
        } else if constexpr (NQ > 0) {
            // ------------------------------------------------------------------ tiny item (4 rows)
            const bool live = (cur - n_parts) * 4 + quarter < n_tiny;    // a dead quarter walks sample 0 at coefficient 0
            const int grow = __shfl_sync(0xffffffffu, fa, 0, 8);
            const int n_ent = __shfl_sync(0xffffffffu, fa, 1, 8);   // every lane takes part in the shuffle
            const int n = live ? n_ent : 0;
            const int b_idx = __shfl_sync(0xffffffffu, fa, 2, 8), packed = __shfl_sync(0xffffffffu, fa, 3, 8);
            const int b1 = __shfl_sync(0xffffffffu, fa, 4, 8), b2 = __shfl_sync(0xffffffffu, fa, 5, 8);
            const int b3 = __shfl_sync(0xffffffffu, fa, 6, 8), b4 = __shfl_sync(0xffffffffu, fa, 7, 8);
            const int e1 = __shfl_sync(0xffffffffu, fb, 0, 8), e2 = __shfl_sync(0xffffffffu, fb, 1, 8);
            const int e3 = __shfl_sync(0xffffffffu, fb, 2, 8);
            const int rcl = packed >> 16;
            const int4* rec = p.rec + ((size_t)b_idx * n_cl + rcl) * AP;
            const char* wts = reinterpret_cast<const char*>(p.recw + ((size_t)b_idx * n_cl + rcl) * AP * d.G);
            const char* gout = reinterpret_cast<const char*>(p.grad_out + (size_t)b_idx * p.A_total * d.C + sub * 4);
            const int gd = d.C / d.G;

            // lane `sub` holds contribution `sub` of its quarter's row (clamped onto the last one, coefficient 0)
            unsigned go_off = 0, w_off = 0;
            float coef = 0.f;
            if (n > 0) {
                const int v = min(sub, n - 1);
                const int k = (v >= e1) + (v >= e2) + (v >= e3);
                const int first = (k == 0) ? 0 : (k == 1) ? e1 : (k == 2) ? e2 : e3;
                const int bk = (k == 0) ? b1 : (k == 1) ? b2 : (k == 2) ? b3 : b4;
                const int pos = bk + (v - first);
                const int4 e = __ldg(rec + pos);
                const float lh = __int_as_float(e.z), lw = __int_as_float(e.w);
                const float hh = 1.f - lh, hw = 1.f - lw;    // same expressions as quad_setup()
                coef = (k == 0) ? hh * hw : (k == 1) ? hh * lw : (k == 2) ? lh * hw : lh * lw;
                if (sub >= n) coef = 0.f;
                go_off = (unsigned)e.y * cx.c_bytes;
                w_off = (unsigned)pos * cx.w_stride_bytes;
            }
            const int n_max = __reduce_max_sync(0xffffffffu, n);
            constexpr int NQ1 = (NQ > 0) ? NQ : 1;
            float acc[NQ1][4];
#pragma unroll
            for (int i = 0; i < NQ1; ++i)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
            for (int m0 = 0; m0 < n_max; m0 += 2) {
                float4 g[2][NQ1];
                float wg[2][NQ1];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int m = min(m0 + u, 7);
                    const unsigned go_m = __shfl_sync(0xffffffffu, go_off, m, 8);
                    const unsigned w_m = __shfl_sync(0xffffffffu, w_off, m, 8);
                    const float cf = __shfl_sync(0xffffffffu, coef, m, 8);
                    const char* g0 = gout + go_m;
                    const float* w0 = reinterpret_cast<const float*>(wts + w_m);
#pragma unroll
                    for (int i = 0; i < NQ1; ++i) {
                        g[u][i] = __ldg(reinterpret_cast<const float4*>(g0 + i * 128));
                        wg[u][i] = __ldg(w0 + (i * 32) / gd) * cf;
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int i = 0; i < NQ1; ++i) {
                        const float2 w2 = make_float2(wg[u][i], wg[u][i]);
                        const float2 lo =
                            __ffma2_rn(w2, make_float2(g[u][i].x, g[u][i].y), make_float2(acc[i][0], acc[i][1]));
                        const float2 hi =
                            __ffma2_rn(w2, make_float2(g[u][i].z, g[u][i].w), make_float2(acc[i][2], acc[i][3]));
                        acc[i][0] = lo.x; acc[i][1] = lo.y; acc[i][2] = hi.x; acc[i][3] = hi.y;
                    }
            }
            if (live) {
                T* dst = reinterpret_cast<T*>(p.g_feat) + (size_t)grow * d.C + sub * 4;
#pragma unroll
                for (int i = 0; i < NQ1; ++i) {
                    if (p.accumulate) {
                        float old[4];
                        VecIO<T, 4>::load(dst + i * 32, old);
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[i][q] += old[q];
                    }
                    VecIO<T, 4>::store(dst + i * 32, acc[i]);
                }
            }
        }
        cur = nxt;
        fa = na;
        fb = nb2;
        nxt = after;
    }
}

}  // namespace hipad
