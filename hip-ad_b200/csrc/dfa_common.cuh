// dfa_common.cuh — shared device helpers of the B200 deformable-aggregation kernels.
//
// Sampling convention restated from the reference op
// (projects/mmdet3d_plugin/ops/src/deformable_aggregation_cuda.cu:13-59, 166-181):
//   sample valid  <=>  0 < x < 1 && 0 < y < 1          (strict, whole sample dropped otherwise)
//   h_im = y*H - 0.5, w_im = x*W - 0.5                 (ONE fp32 FMA in the compiled reference)
//   corners (h_low,w_low) (h_low,w_low+1) (h_low+1,w_low) (h_low+1,w_low+1), zero padding
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

namespace hipad {

constexpr int kMaxCamLevels = 64;  // cams * levels held in shared-memory tables

struct Dims {
    int bs, cams, num_feat, C, L, A, P, G;
};

struct Quad {
    int h_low, w_low;
    float lh, lw, hh, hw;
    bool ok1, ok2, ok3, ok4;  // (low,low) (low,high) (high,low) (high,high) in-bounds
};

__device__ __forceinline__ bool loc_valid(float x, float y) {
    return x > 0.f && x < 1.f && y > 0.f && y < 1.f;
}

// The one place where sample coordinates become integer indices.  Every kernel in
// this library (forward, both backward kernels, the index-export kernel) calls it.
__device__ __forceinline__ Quad quad_setup(float x, float y, int h, int w) {
    Quad q;
    const float h_im = __fmaf_rn(y, (float)h, -0.5f);
    const float w_im = __fmaf_rn(x, (float)w, -0.5f);
    q.h_low = __float2int_rd(h_im);
    q.w_low = __float2int_rd(w_im);
    q.lh = h_im - (float)q.h_low;
    q.lw = w_im - (float)q.w_low;
    q.hh = 1.f - q.lh;
    q.hw = 1.f - q.lw;
    const bool hl = q.h_low >= 0, wl = q.w_low >= 0;
    const bool hh_ = q.h_low + 1 <= h - 1, wh_ = q.w_low + 1 <= w - 1;
    q.ok1 = hl && wl;
    q.ok2 = hl && wh_;
    q.ok3 = hh_ && wl;
    q.ok4 = hh_ && wh_;
    return q;
}

// ---- vector I/O: V consecutive channels of type T <-> V floats --------------------------
template <typename T, int V>
struct VecIO;

template <>
struct VecIO<float, 4> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};

template <>
struct VecIO<float, 8> {   // fp32 side-band data (grad_out) on the bf16 vector path
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};

template <>
struct VecIO<float, 1> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
    static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { *p = v[0]; }
};

__device__ __forceinline__ float bf16_bits_to_float(uint32_t bits16) {
    return __uint_as_float(bits16 << 16);
}
__device__ __forceinline__ uint32_t float_to_bf16_bits(float f) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(f));
}

template <>
struct VecIO<__nv_bfloat16, 8> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(u[i] << 16);
            v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint32_t u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            u[i] = float_to_bf16_bits(v[2 * i]) | (float_to_bf16_bits(v[2 * i + 1]) << 16);
        *reinterpret_cast<uint4*>(p) = make_uint4(u[0], u[1], u[2], u[3]);
    }
};

template <>
struct VecIO<__nv_bfloat16, 4> {   // 4 channels of a bf16 row (quarter-warp reduce)
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
        *reinterpret_cast<uint2*>(p) = make_uint2(float_to_bf16_bits(v[0]) | (float_to_bf16_bits(v[1]) << 16),
                                                  float_to_bf16_bits(v[2]) | (float_to_bf16_bits(v[3]) << 16));
    }
};

template <>
struct VecIO<__nv_bfloat16, 1> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[1]) {
        v[0] = bf16_bits_to_float(__ldg(reinterpret_cast<const unsigned short*>(p)));
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[1]) {
        *p = __float2bfloat16_rn(v[0]);
    }
};

// 16-byte read-only load whose position in the instruction stream is pinned (asm volatile): the software pipeline
// of the packed-bf16 gather depends on the loads of item q+3 being ISSUED before item q is consumed, and nvcc
// otherwise sinks plain __ldg loads next to their first use.
__device__ __forceinline__ uint4 ldg_nc_v4_pinned(const void* p) {
    uint4 t;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "l"(p));
    return t;
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& t, float (&v)[8]) {
    const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(u[i] << 16);
        v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
    }
}

// Ticket increment with RELEASE semantics at gpu scope: orders the writes this thread has observed (its own and, through a
// preceding __syncwarp / __syncthreads, those of the threads it synchronised with) before the increment.  Unlike
// __threadfence() (MEMBAR.SC + CCTL.IVALL in every thread) it costs one MEMBAR in one thread and does NOT invalidate the
// SM's L1, which the gather kernels live on.  Readers of the published data use ld.global.cg (L2) after an acquire fence.
__device__ __forceinline__ int atomic_add_release_gpu(int* p, int v) {
    int old;
    asm volatile("atom.add.release.gpu.global.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Per-CTA phase trace of one kernel (development builds only: HIPAD_DFA_NVCC_EXTRA=-DHIPAD_DFA_TRACE python build.py).
// One trace array per translation unit (read back with hipad_dfa_trace_read(unit, ...)).
// slot 0 = globaltimer at entry, 1 = SM id, 2.. = clock64 at the marks, read back with hipad_dfa_trace_read().
#ifdef HIPAD_DFA_TRACE
constexpr int kTraceSlots = 14, kTraceCtas = 1 << 14;
static __device__ long long g_trace[kTraceCtas * kTraceSlots];
__device__ __forceinline__ void trace_mark(int cta, int slot, long long v) {
    if (threadIdx.x == 0 && cta < kTraceCtas) g_trace[cta * kTraceSlots + slot] = v;
}
__device__ __forceinline__ long long trace_globaltimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int trace_smid() {
    int s;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    return s;
}
#define DFA_TRACE(cta, slot) trace_mark((cta), (slot), clock64())
#define DFA_TRACE_V(cta, slot, v) trace_mark((cta), (slot), (long long)(v))
#define DFA_TRACE_END(cta, slot) trace_mark((cta), (slot), trace_globaltimer())
#define DFA_TRACE_DECL(name) long long name = 0
#define DFA_TRACE_ACC(name, t0) name += clock64() - (t0)
#define DFA_TRACE_NOW(name) const long long name = clock64()
#define DFA_TRACE_BEGIN(cta) do { trace_mark((cta), 0, trace_globaltimer()); trace_mark((cta), 1, trace_smid()); } while (0)
#else
#define DFA_TRACE(cta, slot) ((void)0)
#define DFA_TRACE_V(cta, slot, v) ((void)0)
#define DFA_TRACE_BEGIN(cta) ((void)0)
#define DFA_TRACE_END(cta, slot) ((void)0)
#define DFA_TRACE_DECL(name) ((void)0)
#define DFA_TRACE_ACC(name, t0) ((void)0)
#define DFA_TRACE_NOW(name) ((void)0)
#endif

// (cam,level) table in shared memory: {h, w, start} per entry
__device__ __forceinline__ void load_level_table(int* tab, const int* __restrict__ shapes,
                                                 const int* __restrict__ starts, int n_cl) {
    for (int i = threadIdx.x; i < n_cl; i += blockDim.x) {
        tab[i * 3 + 0] = __ldg(shapes + i * 2);
        tab[i * 3 + 1] = __ldg(shapes + i * 2 + 1);
        tab[i * 3 + 2] = __ldg(starts + i);
    }
}

}  // namespace hipad
