// dfa_format.cu — feature_maps_format as ONE transposing pass (SURVEY §8 a7 / "next" row f3).
//
// The reference (projects/mmdet3d_plugin/ops/__init__.py:74-103) builds the op's input with
//   cat over levels of [bs,cams,C,H_l*W_l]  ->  permute(0,1,3,2)  ->  flatten(1,2)
// i.e. two full copies of every feature map (115 MB each way per sample at 352x640), the second one a
// strided transpose.  Here every (b, cam, level) plane [C][H_l*W_l] is transposed through shared memory
// straight into its rows of col_feats [bs, cams*sum(H_l*W_l), C]: one read and one write of the data, both
// sides in full 128-byte lines.  The same kernel run backwards (kInverse) scatters a col_feats-shaped
// gradient back into the per-level NCHW tensors (autograd of the format step; also the dense form of
// ops/__init__.py:34-65, whose forward direction the Python side serves with zero-copy views).
#include "../../include/hipad_dfa.h"
#include "dfa_common.cuh"

namespace hipad {
namespace {

constexpr int kMaxLevels = 8;
constexpr int kTile = 64;            // tile edge (pixels x channels)
constexpr int kFormatThreads = 256;

struct FormatParams {
    void* level[kMaxLevels];         // per level: [bs, cams, C, H_l*W_l] contiguous
    int hw[kMaxLevels];              // H_l * W_l
    int row0[kMaxLevels];            // first row of the level inside one camera's block of col_feats
    int tile0[kMaxLevels + 1];       // first pixel-tile index of the level (prefix sum of ceil(hw/kTile))
    void* col;                       // [bs, cams * rows_per_cam, C]
    int C, L, rows_per_cam;
};

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// grid (pixel tiles of all levels, channel tiles, bs*cams), block 256.
// TL: element type of the per-level tensors, TC: element type of col_feats (f32 -> bf16 narrows on the way).
template <typename TL, typename TC, bool kInverse>
__global__ void __launch_bounds__(kFormatThreads) dfa_format_kernel(const FormatParams p) {
    __shared__ float tile[kTile][kTile + 1];
    int l = 0;
    while (l + 1 < p.L && (int)blockIdx.x >= p.tile0[l + 1]) ++l;
    const int px0 = ((int)blockIdx.x - p.tile0[l]) * kTile;      // first pixel of the tile inside the level
    const int c0 = (int)blockIdx.y * kTile;
    const int plane = (int)blockIdx.z;                            // b * cams + cam
    const int hw = p.hw[l];
    TL* lev = reinterpret_cast<TL*>(p.level[l]) + (size_t)plane * p.C * hw;
    TC* col = reinterpret_cast<TC*>(p.col) + ((size_t)plane * p.rows_per_cam + p.row0[l]) * p.C;
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;       // 64 x 4

    if (!kInverse) {
        // read [c][pixel] rows (pixel fastest), write [pixel][c] rows (channel fastest)
#pragma unroll 4
        for (int r = ty; r < kTile; r += 4) {
            const int c = c0 + r, px = px0 + tx;
            if (c < p.C && px < hw) tile[r][tx] = to_f<TL>(lev[(size_t)c * hw + px]);
        }
        __syncthreads();
#pragma unroll 4
        for (int r = ty; r < kTile; r += 4) {
            const int px = px0 + r, c = c0 + tx;
            if (px < hw && c < p.C) col[(size_t)px * p.C + c] = from_f<TC>(tile[tx][r]);
        }
    } else {
#pragma unroll 4
        for (int r = ty; r < kTile; r += 4) {
            const int px = px0 + r, c = c0 + tx;
            if (px < hw && c < p.C) tile[tx][r] = to_f<TC>(col[(size_t)px * p.C + c]);
        }
        __syncthreads();
#pragma unroll 4
        for (int r = ty; r < kTile; r += 4) {
            const int c = c0 + r, px = px0 + tx;
            if (c < p.C && px < hw) lev[(size_t)c * hw + px] = from_f<TL>(tile[r][tx]);
        }
    }
}

template <typename TL, typename TC>
int launch_format(const FormatParams& p, bool inverse, dim3 grid, cudaStream_t st) {
    if (inverse)
        dfa_format_kernel<TL, TC, true><<<grid, kFormatThreads, 0, st>>>(p);
    else
        dfa_format_kernel<TL, TC, false><<<grid, kFormatThreads, 0, st>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace
}  // namespace hipad

extern "C" int hipad_dfa_format_features(int level_dtype, int col_dtype, int inverse, void* const* level_ptrs,
                                         const int32_t* level_hw, void* col_feats, int batch_size, int num_cams,
                                         int num_embeds, int num_scale, void* stream) {
    using namespace hipad;
    if (!level_ptrs || !level_hw || !col_feats || batch_size <= 0 || num_cams <= 0 || num_embeds <= 0 || num_scale <= 0)
        return HIPAD_DFA_ERR_BAD_ARGUMENT;
    if (num_scale > kMaxLevels || (long long)batch_size * num_cams > 65535) return HIPAD_DFA_ERR_UNSUPPORTED;
    if ((level_dtype != 0 && level_dtype != 1) || (col_dtype != 0 && col_dtype != 1)) return HIPAD_DFA_ERR_BAD_ARGUMENT;
    FormatParams p = {};
    int rows = 0, tiles = 0;
    for (int l = 0; l < num_scale; ++l) {
        const long long hw = (long long)level_hw[2 * l] * level_hw[2 * l + 1];
        if (!level_ptrs[l] || hw <= 0 || hw > (1 << 28)) return HIPAD_DFA_ERR_BAD_ARGUMENT;
        p.level[l] = level_ptrs[l];
        p.hw[l] = (int)hw;
        p.row0[l] = rows;
        p.tile0[l] = tiles;
        rows += (int)hw;
        tiles += (int)((hw + kTile - 1) / kTile);
    }
    p.tile0[num_scale] = tiles;
    p.col = col_feats;
    p.C = num_embeds;
    p.L = num_scale;
    p.rows_per_cam = rows;
    const dim3 grid((unsigned)tiles, (unsigned)((num_embeds + kTile - 1) / kTile), (unsigned)(batch_size * num_cams));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool inv = inverse != 0;
    if (level_dtype == 0 && col_dtype == 0) return launch_format<float, float>(p, inv, grid, st);
    if (level_dtype == 0 && col_dtype == 1) return launch_format<float, __nv_bfloat16>(p, inv, grid, st);
    if (level_dtype == 1 && col_dtype == 1) return launch_format<__nv_bfloat16, __nv_bfloat16>(p, inv, grid, st);
    return launch_format<__nv_bfloat16, float>(p, inv, grid, st);
}
