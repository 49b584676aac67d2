// dfa_weights.cu — the weights producer of the aggregation module as ONE pass each way ("next" row f2, training half).
//
// Reference chain (models/blocks.py:196-212 and :147-158), every step its own ATen kernel and its own round trip of the
// 9 / 23 / 33 MB (det / map / plan, bs = 1) weight tensor:
//     logits [bs, A, cams, L*P*G] -> reshape [bs, A, cams*L*P, G] -> softmax(dim=-2)           (read + write)
//     training: mask = torch.rand(bs, A, cams, 1, P, 1) ON THE CPU, .to(device)                (H2D copy per call)
//               weights = (mask > attn_drop) * weights / (1 - attn_drop)                       (2 x read + write)
//     permute(0, 1, 4, 2, 3, 5).contiguous() -> [bs, A, P, cams, L, G]                         (read + write)
// and the mirror image in the backward (permute copy, mask, softmax backward).
// Here: dfa_weights_forward reads the logits once (online max / sum per group, logits of a row re-read from L2 for the
// write pass) and writes the op's weight layout directly, with the drop mask drawn in the kernel from (seed, b, a, cam, p);
// dfa_weights_backward turns the op's grad_weights into grad_logits in one pass over both.  One CTA per output row (b, a).
#include "../../include/hipad_dfa.h"
#include "dfa_launch.h"

namespace hipad {
namespace {

constexpr int kWThreads = 256;

// counter-based uniform in [0, 1): two rounds of a 32-bit integer mixer over (seed, index)
__device__ __forceinline__ float drop_uniform(unsigned long long seed, unsigned long long idx) {
    unsigned x = (unsigned)(idx ^ (idx >> 32)) * 0x9E3779B1u ^ (unsigned)seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    x += (unsigned)(seed >> 32);
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return (float)(x >> 8) * (1.0f / 16777216.0f);
}

struct WParams {
    const float* logits;   // [bs*A, cams, L, P, G]
    float* weights;        // [bs*A, P, cams, L, G]
    float* stats;          // [bs*A, G, 2] (max, 1/sum) saved for the backward
    const float* mask;     // optional explicit keep mask [bs*A, cams, P] (1 = keep), overrides the in-kernel draw
    const float* g_w;      // backward: [bs*A, P, cams, L, G]
    float* g_logits;       // backward: [bs*A, cams, L, P, G]
    unsigned long long seed;
    float drop_p;          // 0 = no dropping
    int cams, L, P, G;
};

__device__ __forceinline__ float keep_scale(const WParams& p, long long row, int cam, int pt) {
    if (p.drop_p <= 0.f) return 1.f;
    bool keep;
    if (p.mask) keep = p.mask[(row * p.cams + cam) * p.P + pt] > 0.5f;
    else keep = drop_uniform(p.seed, (unsigned long long)((row * p.cams + cam) * p.P + pt)) > p.drop_p;
    return keep ? 1.f / (1.f - p.drop_p) : 0.f;
}

// per-group (max, sum) of one row; thread t only ever sees group t % G (kWThreads % G == 0)
__device__ __forceinline__ void row_stats(const float* __restrict__ x, int n, int G, float* s_red /*[2*kWThreads]*/,
                                          float* s_stat /*[2*G]*/) {
    const int tid = threadIdx.x;
    float m = -INFINITY, s = 0.f;
    for (int i0 = tid; i0 < n; i0 += kWThreads * 4) {
        float v[4];
        float mx = m;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kWThreads;
            v[u] = (i < n) ? __ldg(x + i) : -INFINITY;
            mx = fmaxf(mx, v[u]);
        }
        float acc = (s == 0.f) ? 0.f : s * expf(m - mx);
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += expf(v[u] - mx);
        m = mx;
        s = acc;
    }
    s_red[tid * 2] = m;
    s_red[tid * 2 + 1] = s;
    __syncthreads();
    if (tid < G) {
        float mm = -INFINITY, ss = 0.f;
        for (int t = tid; t < kWThreads; t += G) {
            const float mt = s_red[t * 2], st = s_red[t * 2 + 1];
            if (st == 0.f) continue;
            const float mn = fmaxf(mm, mt);
            ss = ss * expf(mm - mn) + st * expf(mt - mn);
            mm = mn;
        }
        s_stat[tid * 2] = mm;
        s_stat[tid * 2 + 1] = 1.f / ss;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kWThreads) dfa_weights_fwd_kernel(const WParams p) {
    __shared__ float s_red[2 * kWThreads];
    __shared__ float s_stat[2 * 32];
    const long long row = blockIdx.x;
    const int n = p.cams * p.L * p.P * p.G;
    const float* x = p.logits + row * n;
    row_stats(x, n, p.G, s_red, s_stat);
    const int tid = threadIdx.x;
    if (p.stats && tid < 2 * p.G) p.stats[row * 2 * p.G + tid] = s_stat[tid];
    const int g = tid % p.G;
    const float m = s_stat[g * 2], inv = s_stat[g * 2 + 1];
    float* out = p.weights + row * n;
    const int PG = p.P * p.G, LPG = p.L * PG;
    for (int i = tid; i < n; i += kWThreads) {
        const int cam = i / LPG, r1 = i - cam * LPG;
        const int l = r1 / PG, r2 = r1 - l * PG;
        const int pt = r2 / p.G;                                   // r2 - pt*G == g
        const float w = expf(__ldg(x + i) - m) * inv * keep_scale(p, row, cam, pt);
        out[((pt * p.cams + cam) * p.L + l) * p.G + g] = w;
    }
}

// g_logits_i = soft_i * (gs_i - sum_j gs_j soft_j) per group, with gs = g_w * keep_scale (softmax + mask backward)
__global__ void __launch_bounds__(kWThreads) dfa_weights_bwd_kernel(const WParams p) {
    __shared__ float s_red[kWThreads];
    __shared__ float s_dot[32];
    const long long row = blockIdx.x;
    const int n = p.cams * p.L * p.P * p.G;
    const float* x = p.logits + row * n;
    const float* gw = p.g_w + row * n;
    const int tid = threadIdx.x, g = tid % p.G;
    const float m = p.stats[row * 2 * p.G + g * 2], inv = p.stats[row * 2 * p.G + g * 2 + 1];
    const int PG = p.P * p.G, LPG = p.L * PG;
    float dot = 0.f;
    for (int i = tid; i < n; i += kWThreads) {
        const int cam = i / LPG, r1 = i - cam * LPG;
        const int l = r1 / PG, r2 = r1 - l * PG;
        const int pt = r2 / p.G;
        const float soft = expf(__ldg(x + i) - m) * inv;
        const float gs = __ldg(gw + ((pt * p.cams + cam) * p.L + l) * p.G + g) * keep_scale(p, row, cam, pt);
        dot = fmaf(gs, soft, dot);
    }
    s_red[tid] = dot;
    __syncthreads();
    if (tid < p.G) {
        float d = 0.f;
        for (int t = tid; t < kWThreads; t += p.G) d += s_red[t];   // fixed order
        s_dot[tid] = d;
    }
    __syncthreads();
    const float dg = s_dot[g];
    float* out = p.g_logits + row * n;
    for (int i = tid; i < n; i += kWThreads) {
        const int cam = i / LPG, r1 = i - cam * LPG;
        const int l = r1 / PG, r2 = r1 - l * PG;
        const int pt = r2 / p.G;
        const float soft = expf(__ldg(x + i) - m) * inv;
        const float gs = __ldg(gw + ((pt * p.cams + cam) * p.L + l) * p.G + g) * keep_scale(p, row, cam, pt);
        out[i] = soft * (gs - dg);
    }
}

inline bool bad(long long rows, int cams, int L, int P, int G, float drop_p) {
    return rows <= 0 || rows > 0x7fffffffLL || cams <= 0 || L <= 0 || P <= 0 || G <= 0 || drop_p < 0.f || drop_p >= 1.f;
}
}  // namespace
}  // namespace hipad

using namespace hipad;

extern "C" {

int hipad_dfa_weights_forward(const float* logits, float* weights, float* stats, const float* keep_mask,
                              unsigned long long seed, float drop_p, long long rows, int num_cams, int num_scale,
                              int num_pts, int num_groups, void* stream) {
    if (!logits || !weights || bad(rows, num_cams, num_scale, num_pts, num_groups, drop_p)) return HIPAD_DFA_ERR_BAD_ARGUMENT;
    if (kWThreads % num_groups != 0 || num_groups > 32 ||
        (long long)num_cams * num_scale * num_pts * num_groups >= (1LL << 30))
        return HIPAD_DFA_ERR_UNSUPPORTED;
    WParams p = {};
    p.logits = logits; p.weights = weights; p.stats = stats; p.mask = keep_mask; p.seed = seed; p.drop_p = drop_p;
    p.cams = num_cams; p.L = num_scale; p.P = num_pts; p.G = num_groups;
    dfa_weights_fwd_kernel<<<(unsigned)rows, kWThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return (int)cudaGetLastError();
}

int hipad_dfa_weights_backward(const float* logits, const float* stats, const float* grad_weights, float* grad_logits,
                               const float* keep_mask, unsigned long long seed, float drop_p, long long rows,
                               int num_cams, int num_scale, int num_pts, int num_groups, void* stream) {
    if (!logits || !stats || !grad_weights || !grad_logits || bad(rows, num_cams, num_scale, num_pts, num_groups, drop_p))
        return HIPAD_DFA_ERR_BAD_ARGUMENT;
    if (kWThreads % num_groups != 0 || num_groups > 32 ||
        (long long)num_cams * num_scale * num_pts * num_groups >= (1LL << 30))
        return HIPAD_DFA_ERR_UNSUPPORTED;
    WParams p = {};
    p.logits = logits; p.stats = const_cast<float*>(stats); p.g_w = grad_weights; p.g_logits = grad_logits;
    p.mask = keep_mask; p.seed = seed; p.drop_p = drop_p;
    p.cams = num_cams; p.L = num_scale; p.P = num_pts; p.G = num_groups;
    dfa_weights_bwd_kernel<<<(unsigned)rows, kWThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    return (int)cudaGetLastError();
}

}  // extern "C"
