// Shared helpers for the invertavatar_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/invertavatar_b200.h"

namespace ia {

void set_error(const char* fmt, ...);
void count_launch();
// Optional per-launch CUDA-event bracketing (ia_profile_begin/ia_profile_report): prof_begin records the start event on
// the launching stream, IA_LAUNCH_CHECK closes the bracket.  Both are no-ops unless profiling is enabled.
void prof_begin(const char* name, cudaStream_t stream);
void prof_end();
const char* prof_detail_name(const char* base, int ntaps, int gh, int gw, int cin, int cout);

#define IA_CHECK(cond, ...)                                   \
    do {                                                      \
        if (!(cond)) {                                        \
            ia::set_error(__VA_ARGS__);                       \
            return 1;                                         \
        }                                                     \
    } while (0)

#define IA_LAUNCH_CHECK(name)                                                        \
    do {                                                                             \
        ia::count_launch();                                                          \
        ia::prof_end();                                                              \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            ia::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
            return 2;                                                                \
        }                                                                            \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// bf16 split: hi = rn(v), lo = rn(v - hi)
__device__ __forceinline__ void split_bf16(float v, uint16_t& hi, uint16_t& lo) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    float r = v - __bfloat162float(h);
    __nv_bfloat16 l = __float2bfloat16_rn(r);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(l);
}
// Two values at once (packed cvt.rn.bf16x2.f32: same rounding as above, half the instructions): hi / lo hold a in their
// low and b in their high 16 bits -- the in-memory order of two consecutive channels.
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    uint32_t h;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
    const float ra = a - __uint_as_float(h << 16);
    const float rb = b - __uint_as_float(h & 0xffff0000u);
    uint32_t l;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(rb), "f"(ra));
    hi = h; lo = l;
}
__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }

// IA_OPFMT_F16X1 operands: round to nearest fp16, saturating at the largest finite value (an activation beyond 65504 must not
// become inf inside the accumulation).  Two values -> one 32-bit word in memory order (a low, b high).
__device__ __forceinline__ uint32_t pack_f16x2_sat(float a, float b) {
    a = fminf(fmaxf(a, -65504.f), 65504.f);
    b = fminf(fmaxf(b, -65504.f), 65504.f);
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float f16_bits_to_float(uint16_t b) { return __half2float(__ushort_as_half(b)); }
// Four consecutive channels of one operand pixel in format `fmt`: bf16 hi/lo pair, or fp16 in `hi` alone.
__device__ __forceinline__ void store_operand4(int fmt, uint16_t* hi, uint16_t* lo, float a, float b, float c, float d) {
    if (fmt == IA_OPFMT_F16X1) {
        *reinterpret_cast<uint2*>(hi) = make_uint2(pack_f16x2_sat(a, b), pack_f16x2_sat(c, d));
    } else {
        uint2 hv, lv;
        split_bf16x2(a, b, hv.x, lv.x);
        split_bf16x2(c, d, hv.y, lv.y);
        *reinterpret_cast<uint2*>(hi) = hv;
        *reinterpret_cast<uint2*>(lo) = lv;
    }
}

// activation + gain + clamp shared by every epilogue (reference bias_act.cu:27-151 forward path; only the
// activations the generator uses get the fast path, the rest are exact expressions).
__device__ __forceinline__ float apply_act(float x, int act, float alpha) {
    switch (act) {
        case IA_ACT_LINEAR: return x;
        case IA_ACT_RELU: return x > 0.f ? x : 0.f;
        case IA_ACT_LRELU: return x > 0.f ? x : x * alpha;
        case IA_ACT_TANH: return tanhf(x);
        case IA_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
        case IA_ACT_ELU: return x > 0.f ? x : expm1f(x);
        case IA_ACT_SELU: return x > 0.f ? 1.0507009873554805f * x : 1.0507009873554805f * 1.6732632423543772f * expm1f(x);
        case IA_ACT_SOFTPLUS: return x > 20.f ? x : log1pf(expf(x));
        case IA_ACT_SWISH: return x / (1.f + expf(-x));
        case IA_ACT_PRELU: return x > 0.f ? x : x * alpha;      // alpha = this channel's slope (convolution epilogues only)
    }
    return x;
}
__device__ __forceinline__ float act_gain_clamp(float x, int act, float alpha, float gain, float clamp) {
    x = apply_act(x, act, alpha) * gain;
    if (clamp >= 0.f) x = fminf(fmaxf(x, -clamp), clamp);
    return x;
}

// Emit one group of 4 consecutive channels (co..co+3) of pixel `pix` of image b.
__device__ __forceinline__ void emit4(const ia_emit& e, int b, int64_t pix, int co, int C, const float v[4]) {
    if (e.out32) {
        float* o = e.out32 + pix * e.out32_ld + co;
        if (co + 3 < C && ((e.out32_ld & 3) == 0)) {
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            for (int k = 0; k < 4; ++k) if (co + k < C) o[k] = v[k];
        }
    }
    if (e.hi1) {
        float m[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = (co + k < C) ? v[k] * (e.s1 ? e.s1[(int64_t)b * C + co + k] : 1.f) : 0.f;
        store_operand4(e.fmt1, e.hi1 + pix * e.c1_pad + co, e.lo1 + pix * e.c1_pad + co, m[0], m[1], m[2], m[3]);
    }
    if (e.hi2) {
        float m[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = (co + k < C) ? v[k] * (e.s2 ? e.s2[(int64_t)b * C + co + k] : 1.f) : 0.f;
        store_operand4(e.fmt2, e.hi2 + pix * e.c2_pad + co, e.lo2 + pix * e.c2_pad + co, m[0], m[1], m[2], m[3]);
    }
}


// Warp-cooperative dot products of one weight row with up to NB activation rows: acc[i] += sum_k f(x[i * xs + k]) * w[k], f = square or
// identity, partial sums per lane (the caller reduces across the warp).  The tiny GEMVs of the mapping network / style affines /
// demodulation coefficients are latency-bound: with K a multiple of 128 and 16-byte aligned rows every lane issues its float4 loads of
// four 128-element chunks back to back (one exposed memory latency per 512 elements instead of one per 32).
template <int NB, bool SQUARE>
__device__ __forceinline__ void warp_dot_rows(const float* __restrict__ w, const float* __restrict__ x, int64_t xs, int K, int nb, int lane,
                                              float (&acc)[NB]) {
    const bool vec = (K & 127) == 0 && ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(x)) & 15) == 0 && (xs & 3) == 0;
    if (vec) {
#pragma unroll 4
        for (int k0 = 0; k0 < K; k0 += 128) {
            const int k = k0 + 4 * lane;
            const float4 wv = __ldg(reinterpret_cast<const float4*>(w + k));
#pragma unroll
            for (int i = 0; i < NB; ++i)
                if (i < nb) {
                    float4 xv = *reinterpret_cast<const float4*>(x + (int64_t)i * xs + k);
                    if (SQUARE) { xv.x *= xv.x; xv.y *= xv.y; xv.z *= xv.z; xv.w *= xv.w; }
                    acc[i] = fmaf(xv.w, wv.w, fmaf(xv.z, wv.z, fmaf(xv.y, wv.y, fmaf(xv.x, wv.x, acc[i]))));
                }
        }
        return;
    }
    for (int k = lane; k < K; k += 32) {
        const float wv = w[k];
#pragma unroll
        for (int i = 0; i < NB; ++i)
            if (i < nb) {
                float xv = x[(int64_t)i * xs + k];
                if (SQUARE) xv *= xv;
                acc[i] = fmaf(xv, wv, acc[i]);
            }
    }
}

}  // namespace ia
