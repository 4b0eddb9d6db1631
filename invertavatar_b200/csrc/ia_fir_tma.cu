// FIR epilogue of the up=2 layers as a TMA-fed streaming kernel.
//
// The register kernels of ia_modconv.cu keep at most two raw rows per thread in flight (~39 KB of unique reads per SM) and
// are latency-bound at 3.3-3.8 TB/s (profiles/r1_fir_full_v3.txt, ncu stall sampling: 42 % long-scoreboard).  Here a
// producer warp streams the raw (2H+1)x(2W+1) transposed-convolution output through a 4-stage shared-memory ring with
// cp.async.bulk.tensor (boxes of 4 rows x 35 columns x 32 channels fp32, zero-filled outside the image, which implements
// the [1,1,1,1] padding), three CTAs per SM: ~215 KB of reads in flight per SM.  The 256 consumer threads own 4 channels of
// one output column each, walk the rows of a 64-row strip with the separable 4-tap filter held as running accumulators
// (same arithmetic, same order as fir_epilogue_kernel -> bit-identical results) and apply demod / noise / bias /
// activation / clamp and the operand emission of the consumers.
//
// Replaces upfirdn2d(up=1, pad, gain=up^2) + bias_act after the transposed convolution of reference
// torch_utils/ops/conv2d_resample.py:127-128 and networks_stylegan2_new.py:74-90.
#include <cuda.h>
#include <stdlib.h>

#include "ia_common.cuh"

using namespace ia;

namespace {

constexpr int FT_XT = 32;        // output columns per CTA
constexpr int FT_CC = 32;        // channels per CTA (128 bytes: the innermost box dimension)
constexpr int FT_R = 4;          // raw rows per stage
constexpr int FT_STAGES = 4;
constexpr int FT_YT = 64;        // output rows per CTA
constexpr int FT_COLS = FT_XT + 3;
constexpr int FT_STAGE_FLOATS = FT_R * FT_COLS * FT_CC;
constexpr uint32_t FT_STAGE_BYTES = FT_STAGE_FLOATS * 4u;
constexpr int FT_THREADS = 256 + 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("ia_fir_tma: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <int ACT, bool NPF>     // NPF: noise loaded one emission ahead (see fir_epilogue_x2_kernel)
__global__ void __launch_bounds__(FT_THREADS, 3) fir_tma_kernel(const __grid_constant__ CUtensorMap tm_raw, const ia_fir_params p, int cchunks) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    const float* ring = reinterpret_cast<const float*>(smem_raw + (base - smem_u32(smem_raw)));
    const uint32_t bar_base = base + FT_STAGES * FT_STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(FT_STAGES + s); };

    const int ox0 = blockIdx.x * FT_XT;
    const int oy0 = blockIdx.y * FT_YT;
    const int oy1 = min(oy0 + FT_YT, p.OH);
    const int b = blockIdx.z / cchunks;
    const int cb = (blockIdx.z % cchunks) * FT_CC;
    // raw rows oy0-1 .. oy1+1 feed the output rows [oy0, oy1)
    const int nrows = oy1 - oy0 + 3;
    const int nchunks = (nrows + FT_R - 1) / FT_R;

    if (threadIdx.x == 0) {
        for (int s = 0; s < FT_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (threadIdx.x >= 256) {
        // ===================== TMA producer (one lane) =====================
        if (threadIdx.x == 256) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_raw) : "memory");
            for (int i = 0; i < nchunks; ++i) {
                const int s = i % FT_STAGES;
                mbar_wait(empty_bar(s), (((uint32_t)(i / FT_STAGES)) & 1u) ^ 1u);
                mbar_expect_tx(full_bar(s), FT_STAGE_BYTES);
                tma_load_4d(base + (uint32_t)s * FT_STAGE_BYTES, &tm_raw, full_bar(s), cb, ox0 - 1, oy0 - 1 + i * FT_R, b);
            }
        }
        return;
    }

    // ===================== consumers =====================
    const int c4 = threadIdx.x & 7;
    const int xl = threadIdx.x >> 3;
    const int lane = threadIdx.x & 31;
    const int ox = ox0 + xl;
    const int c0 = cb + c4 * 4;
    const bool active = ox < p.OW && c0 < p.C;
    // out[oy][ox] = sum_{ty,tx} F[3-ty][3-tx] * raw[oy+ty-1][ox+tx-1]  (pad [1,1,1,1], gain folded in); F = fy (x) fx
    float fy[4], fx[4];
    {
        float tot = 0.f, rs[4] = {0.f, 0.f, 0.f, 0.f}, cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float f = p.fir[i * 4 + j]; rs[i] += f; cs[j] += f; tot += f; }
#pragma unroll
        for (int t = 0; t < 4; ++t) { fy[t] = rs[3 - t]; fx[t] = cs[3 - t] / tot; }
    }
    const int grp = p.groups > 1 ? b / p.imgs_per_group : 0;
    float dc[4], bs[4], s1[4], s2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool cv = c0 + k < p.C;
        dc[k] = (p.dcoef && cv) ? p.dcoef[(int64_t)b * p.C + c0 + k] : 1.f;
        bs[k] = (p.bias && cv) ? p.bias[(int64_t)grp * p.C + c0 + k] : 0.f;
        s1[k] = (p.emit.hi1 && p.emit.s1 && cv) ? p.emit.s1[(int64_t)b * p.C + c0 + k] : 1.f;
        s2[k] = (p.emit.hi2 && p.emit.s2 && cv) ? p.emit.s2[(int64_t)b * p.C + c0 + k] : 1.f;
    }
    const float nstr = p.noise ? p.noise_strength[grp] : 0.f;
    const float* nptr = (p.noise && active) ? p.noise + (int64_t)grp * p.noise_gstride + (int64_t)b * p.noise_bstride + (int64_t)oy0 * p.OW + ox : nullptr;
    const float gain = p.gain, alpha = p.alpha, clampv = p.clamp;
    // running 32-bit element offsets (the launcher checks that every emitted tensor has < 2^31 elements)
    const int64_t opix = ((int64_t)b * p.OH + oy0) * p.OW + (active ? ox : 0);
    const uint32_t o32ld = (uint32_t)p.emit.out32_ld, c1p = (uint32_t)p.emit.c1_pad, c2p = (uint32_t)p.emit.c2_pad;
    uint32_t o32 = (uint32_t)(opix * o32ld + c0), o1 = (uint32_t)(opix * c1p + c0), o2 = (uint32_t)(opix * c2p + c0);
    const uint32_t o32row = (uint32_t)p.OW * o32ld, r1row = (uint32_t)p.OW * c1p, r2row = (uint32_t)p.OW * c2p;
    const bool has32 = p.emit.out32 != nullptr, has1 = p.emit.hi1 != nullptr, has2 = p.emit.hi2 != nullptr;

    float nq = 0.f;                                       // NPF: noise of the next output row to be emitted
    int nrows_left = oy1 - oy0;
    if (NPF && nptr) { nq = nptr[0]; nptr += p.OW; --nrows_left; }
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc0 = z4, acc1 = z4, acc2 = z4;               // output rows ry-2, ry-1, ry
    const float* my = ring + xl * FT_CC + c4 * 4;         // column ox-1 of the box = raw column ox0-1+xl
    int ry = oy0 - 1;
    for (int i = 0; i < nchunks; ++i) {
        const int s = i % FT_STAGES;
        mbar_wait(full_bar(s), ((uint32_t)(i / FT_STAGES)) & 1u);
        const float* st = my + s * FT_STAGE_FLOATS;
#pragma unroll
        for (int rr = 0; rr < FT_R; ++rr, ++ry) {
            const float* q = st + rr * (FT_COLS * FT_CC);
            const float4 r0 = *reinterpret_cast<const float4*>(q);
            const float4 r1 = *reinterpret_cast<const float4*>(q + FT_CC);
            const float4 r2 = *reinterpret_cast<const float4*>(q + 2 * FT_CC);
            const float4 r3 = *reinterpret_cast<const float4*>(q + 3 * FT_CC);
            float4 h;
            h.x = fmaf(fx[3], r3.x, fmaf(fx[2], r2.x, fmaf(fx[1], r1.x, fx[0] * r0.x)));
            h.y = fmaf(fx[3], r3.y, fmaf(fx[2], r2.y, fmaf(fx[1], r1.y, fx[0] * r0.y)));
            h.z = fmaf(fx[3], r3.z, fmaf(fx[2], r2.z, fmaf(fx[1], r1.z, fx[0] * r0.z)));
            h.w = fmaf(fx[3], r3.w, fmaf(fx[2], r2.w, fmaf(fx[1], r1.w, fx[0] * r0.w)));
            // acc_j accumulates output row (ry - 2 + j): raw row ry is its tap ty = 3 - j
            acc0.x = fmaf(fy[3], h.x, acc0.x); acc0.y = fmaf(fy[3], h.y, acc0.y); acc0.z = fmaf(fy[3], h.z, acc0.z); acc0.w = fmaf(fy[3], h.w, acc0.w);
            acc1.x = fmaf(fy[2], h.x, acc1.x); acc1.y = fmaf(fy[2], h.y, acc1.y); acc1.z = fmaf(fy[2], h.z, acc1.z); acc1.w = fmaf(fy[2], h.w, acc1.w);
            acc2.x = fmaf(fy[1], h.x, acc2.x); acc2.y = fmaf(fy[1], h.y, acc2.y); acc2.z = fmaf(fy[1], h.z, acc2.z); acc2.w = fmaf(fy[1], h.w, acc2.w);
            const float4 acc3 = make_float4(fy[0] * h.x, fy[0] * h.y, fy[0] * h.z, fy[0] * h.w);
            if (ry - 2 >= oy0 && ry - 2 < oy1) {          // output row ry-2 is complete
                if (active) {
                    float nz = 0.f;
                    if (nptr) {
                        if (NPF) {
                            nz = nq * nstr;
                            if (nrows_left > 0) { nq = nptr[0]; nptr += p.OW; --nrows_left; }
                        } else {
                            nz = nptr[0] * nstr; nptr += p.OW;
                        }
                    }
                    const float a4[4] = {acc0.x, acc0.y, acc0.z, acc0.w};
                    float v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float t = p.dcoef ? fmaf(a4[k], dc[k], nz) : a4[k] + nz;     // fma(x, dcoef, noise), networks_stylegan2_new.py:74
                        t += bs[k];
                        if (ACT == IA_ACT_LRELU) t = (t > 0.f ? t : t * alpha) * gain;
                        else if (ACT == IA_ACT_LINEAR) t = t * gain;
                        else t = apply_act(t, p.act, alpha) * gain;
                        if (clampv >= 0.f) t = fminf(fmaxf(t, -clampv), clampv);
                        v[k] = t;
                    }
                    if (has32) {
                        float* o = p.emit.out32 + o32;
                        if ((o32ld & 3) == 0) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                        else { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3]; }
                    }
                    if (has1) {
                        uint2 hv, lv;
                        split_bf16x2(v[0] * s1[0], v[1] * s1[1], hv.x, lv.x);
                        split_bf16x2(v[2] * s1[2], v[3] * s1[3], hv.y, lv.y);
                        *reinterpret_cast<uint2*>(p.emit.hi1 + o1) = hv;
                        *reinterpret_cast<uint2*>(p.emit.lo1 + o1) = lv;
                    }
                    if (has2) {
                        uint2 hv, lv;
                        split_bf16x2(v[0] * s2[0], v[1] * s2[1], hv.x, lv.x);
                        split_bf16x2(v[2] * s2[2], v[3] * s2[3], hv.y, lv.y);
                        *reinterpret_cast<uint2*>(p.emit.hi2 + o2) = hv;
                        *reinterpret_cast<uint2*>(p.emit.lo2 + o2) = lv;
                    }
                }
                o32 += o32row; o1 += r1row; o2 += r2row;
            }
            acc0 = acc1; acc1 = acc2; acc2 = acc3;
        }
        // this warp is done reading stage s
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
    return fn;
}

}  // namespace

namespace ia {

// 1 when the TMA kernel can take this launch (ia_fir_epilogue falls back to the register kernels otherwise).
bool fir_tma_eligible(const ia_fir_params* p) {
    // Measured on B200 (512^2 x 128 channels, batch 8): 592 us against 569 us of the two-column register kernel -- both sit at
    // ~3.8 TB/s with DRAM only 38 % busy (ncu), i.e. the bytes in flight were not what bounds this pass -- so the variant is
    // off by default; IA_FIR_TMA=1 enables it (read per call; tests/test_gpu_regress.py checks it bit for bit).
    const char* e = getenv("IA_FIR_TMA");
    if (!e || atoi(e) == 0) return false;
    if ((p->emit.hi1 && p->emit.fmt1 != IA_OPFMT_BF16X3) || (p->emit.hi2 && p->emit.fmt2 != IA_OPFMT_BF16X3)) return false;   // bf16 hi/lo emission only
    if (p->C % FT_CC != 0 || p->OW % FT_XT != 0 || p->OH < 8) return false;
    if ((reinterpret_cast<uintptr_t>(p->raw) & 15) != 0) return false;
    const int64_t out_pix = (int64_t)p->B * p->OH * p->OW;
    if (out_pix * (p->emit.out32 ? p->emit.out32_ld : 0) >= (1ll << 31) || out_pix * (p->emit.hi1 ? p->emit.c1_pad : 0) >= (1ll << 31) ||
        out_pix * (p->emit.hi2 ? p->emit.c2_pad : 0) >= (1ll << 31))
        return false;
    return (int64_t)p->B * (p->C / FT_CC) <= 65535;
}

int fir_tma_launch(const ia_fir_params* p, void* stream) {
    EncodeTiledFn enc = get_encode_fn();
    IA_CHECK(enc, "ia_fir_epilogue: cuTensorMapEncodeTiled unavailable");
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)p->C, (cuuint64_t)p->RW, (cuuint64_t)p->RH, (cuuint64_t)p->B};
    cuuint64_t strides[3] = {(cuuint64_t)p->C * 4, (cuuint64_t)p->RW * p->C * 4, (cuuint64_t)p->RH * p->RW * p->C * 4};
    cuuint32_t box[4] = {(cuuint32_t)FT_CC, (cuuint32_t)FT_COLS, (cuuint32_t)FT_R, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p->raw), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IA_CHECK(r == CUDA_SUCCESS, "ia_fir_epilogue: raw tensor map encode failed (CUresult %d; B=%d RH=%d RW=%d C=%d)", (int)r, p->B, p->RH,
             p->RW, p->C);
    const int cchunks = p->C / FT_CC;
    const size_t smem = (size_t)FT_STAGES * FT_STAGE_BYTES + 128 + 8 * 2 * FT_STAGES;
    dim3 grid((unsigned)(p->OW / FT_XT), (unsigned)cdiv(p->OH, FT_YT), (unsigned)(p->B * cchunks));
    cudaError_t e;
    bool npf = false;
    { const char* ev = getenv("IA_FIR_NOISE_PREFETCH"); if (ev && atoi(ev) != 0) npf = true; }
#define IA_FIR_LAUNCH(A, N)                                                                                                  \
    e = cudaFuncSetAttribute(fir_tma_kernel<A, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                  \
    IA_CHECK(e == cudaSuccess, "ia_fir_epilogue: cudaFuncSetAttribute: %s", cudaGetErrorString(e));                          \
    ia::prof_begin("ia_fir_epilogue", as_stream(stream));                                                                    \
    fir_tma_kernel<A, N><<<grid, FT_THREADS, smem, as_stream(stream)>>>(tm, *p, cchunks)
    if (npf) {
        if (p->act == IA_ACT_LRELU) { IA_FIR_LAUNCH(IA_ACT_LRELU, true); }
        else if (p->act == IA_ACT_LINEAR) { IA_FIR_LAUNCH(IA_ACT_LINEAR, true); }
        else { IA_FIR_LAUNCH(-1, true); }
    } else {
        if (p->act == IA_ACT_LRELU) { IA_FIR_LAUNCH(IA_ACT_LRELU, false); }
        else if (p->act == IA_ACT_LINEAR) { IA_FIR_LAUNCH(IA_ACT_LINEAR, false); }
        else { IA_FIR_LAUNCH(-1, false); }
    }
#undef IA_FIR_LAUNCH
    IA_LAUNCH_CHECK("ia_fir_epilogue");
    return 0;
}

}  // namespace ia
