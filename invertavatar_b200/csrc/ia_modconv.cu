// Modulated-convolution support kernels: style affines + demodulation coefficients, operand preparation
// (modulate + bf16 hi/lo split), weight packing, the up=2 FIR epilogue, the ToRGB tail and the CUDA-core
// implicit-GEMM convolution used for bring-up / cross-checking the tcgen05 kernel.
// Arithmetic follows reference training_avatar_texture/networks_stylegan2_new.py:34-91 (unfused branch :70-79),
// :311-357, torch_utils/ops/conv2d_resample.py:114-131 and torch_utils/ops/upfirdn2d.py:315-350.
#include "ia_common.cuh"

using namespace ia;

// ------------------------------------------------------------------------------------------------
// styles + demodulation coefficients for every layer of a network in two launches
// ------------------------------------------------------------------------------------------------
__global__ void styles_kernel(const ia_style_layer* __restrict__ layers, const float* __restrict__ ws, int B, int num_ws) {
    const ia_style_layer L = layers[blockIdx.y];
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= L.Cin) return;
    const float* wr = L.affine_w + (int64_t)warp * L.w_dim;
    for (int b0 = 0; b0 < B; b0 += 8) {
        int nb = min(8, B - b0);
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        warp_dot_rows<8, false>(wr, ws + ((int64_t)b0 * num_ws + L.w_index) * L.w_dim, (int64_t)num_ws * L.w_dim, L.w_dim, nb, lane, acc);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = acc[i];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && i < nb)
                L.styles[(int64_t)(b0 + i) * L.Cin + warp] = (v * L.affine_gain + L.affine_b[warp]) * L.style_gain;
        }
    }
}

__global__ void demod_kernel(const ia_style_layer* __restrict__ layers, int B) {
    const ia_style_layer L = layers[blockIdx.y];
    if (L.wsq == nullptr || L.dcoef == nullptr) return;
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= L.Cout) return;
    const float* wr = L.wsq + (int64_t)warp * L.Cin;
    for (int b0 = 0; b0 < B; b0 += 8) {
        int nb = min(8, B - b0);
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
        warp_dot_rows<8, true>(wr, L.styles + (int64_t)b0 * L.Cin, (int64_t)L.Cin, L.Cin, nb, lane, acc);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = acc[i];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && i < nb) L.dcoef[(int64_t)(b0 + i) * L.Cout + warp] = rsqrtf(v + 1e-8f);
        }
    }
}

extern "C" int ia_styles(const ia_style_layer* layers_dev, const ia_style_layer* layers_host, int32_t n_layers,
                         const float* ws, int32_t B, int32_t num_ws, void* stream) {
    IA_CHECK(layers_dev && layers_host && ws, "ia_styles: null argument");
    if (n_layers == 0 || B == 0) return 0;
    int max_cin = 0, max_cout = 0;
    for (int i = 0; i < n_layers; ++i) {
        IA_CHECK(layers_host[i].w_index >= 0 && layers_host[i].w_index < num_ws, "ia_styles: layer %d w_index out of range", i);
        max_cin = layers_host[i].Cin > max_cin ? layers_host[i].Cin : max_cin;
        if (layers_host[i].wsq) max_cout = layers_host[i].Cout > max_cout ? layers_host[i].Cout : max_cout;
    }
    dim3 g1((unsigned)cdiv((int64_t)max_cin * 32, 256), n_layers);
    ia::prof_begin("ia_styles(styles)", as_stream(stream));
    styles_kernel<<<g1, 256, 0, as_stream(stream)>>>(layers_dev, ws, B, num_ws);
    IA_LAUNCH_CHECK("ia_styles(styles)");
    if (max_cout > 0) {
        dim3 g2((unsigned)cdiv((int64_t)max_cout * 32, 256), n_layers);
        ia::prof_begin("ia_styles(demod)", as_stream(stream));
        demod_kernel<<<g2, 256, 0, as_stream(stream)>>>(layers_dev, B);
        IA_LAUNCH_CHECK("ia_styles(demod)");
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// modulate (+ optional cond blend) and split into bf16 hi/lo, zero-padded to C_pad
// ------------------------------------------------------------------------------------------------
__global__ void modsplit_kernel(ia_modsplit_params p) {
    const int groups = p.C_pad >> 2;  // 4 channels per thread
    int64_t total = (int64_t)p.B * p.HW * groups;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int g = i % groups;
    int64_t pix = i / groups;  // b*HW + hw
    int b = (int)(pix / p.HW);
    int c0 = g * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < p.C) {
        const float* xp = p.x + pix * p.x_ld + c0;
        if (c0 + 3 < p.C && (p.x_ld & 3) == 0) {
            float4 t = *reinterpret_cast<const float4*>(xp);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
            for (int k = 0; k < 4; ++k) if (c0 + k < p.C) v[k] = xp[k];
        }
        if (p.cond) {
            float a = p.cond_alpha[pix];
            const float* cp = p.cond + pix * p.cond_ld + c0;
            for (int k = 0; k < 4; ++k)
                if (c0 + k < p.C) v[k] = cp[k] * a + v[k] * (1.f - a);
        }
        if (p.styles) {
            const float* sp = p.styles + (int64_t)b * p.C + c0;
            for (int k = 0; k < 4; ++k)
                if (c0 + k < p.C) v[k] *= sp[k];
        }
    }
    const int64_t opix = p.out_img_pix ? (int64_t)b * p.out_img_pix + (pix - (int64_t)b * p.HW) : pix;   // padded image stride
    store_operand4(p.fmt, p.hi + opix * p.C_pad + c0, p.lo + opix * p.C_pad + c0, v[0], v[1], v[2], v[3]);
}

extern "C" int ia_modsplit(const ia_modsplit_params* p, void* stream) {
    IA_CHECK(p && p->x && p->hi && (p->lo || p->fmt == IA_OPFMT_F16X1), "ia_modsplit: null tensor");
    IA_CHECK(p->fmt == IA_OPFMT_BF16X3 || p->fmt == IA_OPFMT_F16X1, "ia_modsplit: unknown operand format %d", p->fmt);
    IA_CHECK(p->C > 0 && p->C_pad >= p->C && (p->C_pad & 3) == 0, "ia_modsplit: C_pad must be >= C and a multiple of 4");
    IA_CHECK(p->cond == nullptr || p->cond_alpha != nullptr, "ia_modsplit: cond needs cond_alpha");
    int64_t total = (int64_t)p->B * p->HW * (p->C_pad >> 2);
    if (total == 0) return 0;
    ia::prof_begin("ia_modsplit", as_stream(stream));
    modsplit_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(*p);
    IA_LAUNCH_CHECK("ia_modsplit");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> [tap][Cout_pad][Cin_pad] bf16 hi/lo, wsq[o][i] = sum_taps w^2
// ------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int Cout_pad, int Cin_pad,
                                   uint16_t* __restrict__ w_hi, uint16_t* __restrict__ w_lo, int fmt) {
    int64_t total = (int64_t)taps * Cout_pad * Cin_pad;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int ci = i % Cin_pad; int64_t t = i / Cin_pad;
    int co = t % Cout_pad; int tap = (int)(t / Cout_pad);
    float v = (ci < Cin && co < Cout) ? w[((int64_t)co * Cin + ci) * taps + tap] : 0.f;
    if (fmt == IA_OPFMT_F16X1) {
        w_hi[i] = __half_as_ushort(__float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f)));
        return;
    }
    uint16_t h, l;
    split_bf16(v, h, l);
    w_hi[i] = h;
    w_lo[i] = l;
}
__global__ void wsq_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, float* __restrict__ wsq) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)Cout * Cin) return;
    float s = 0.f;
    for (int t = 0; t < taps; ++t) { float v = w[i * taps + t]; s = fmaf(v, v, s); }
    wsq[i] = s;
}

extern "C" int ia_pack_conv_weight(const float* w, int32_t Cout, int32_t Cin, int32_t kh, int32_t kw, int32_t Cout_pad,
                                   int32_t Cin_pad, uint16_t* w_hi, uint16_t* w_lo, float* wsq, int32_t fmt, void* stream) {
    IA_CHECK(fmt == IA_OPFMT_BF16X3 || fmt == IA_OPFMT_F16X1, "ia_pack_conv_weight: unknown operand format %d", fmt);
    IA_CHECK(w && w_hi && (w_lo || fmt == IA_OPFMT_F16X1), "ia_pack_conv_weight: null tensor");
    IA_CHECK(Cout_pad >= Cout && Cin_pad >= Cin, "ia_pack_conv_weight: bad padding");
    int taps = kh * kw;
    int64_t total = (int64_t)taps * Cout_pad * Cin_pad;
    ia::prof_begin("ia_pack_conv_weight", as_stream(stream));
    pack_weight_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(w, Cout, Cin, taps, Cout_pad, Cin_pad, w_hi, w_lo, fmt);
    IA_LAUNCH_CHECK("ia_pack_conv_weight");
    if (wsq) {
        ia::prof_begin("ia_pack_conv_weight(wsq)", as_stream(stream));
        wsq_kernel<<<(unsigned)cdiv((int64_t)Cout * Cin, 256), 256, 0, as_stream(stream)>>>(w, Cout, Cin, taps, wsq);
        IA_LAUNCH_CHECK("ia_pack_conv_weight(wsq)");
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// shared epilogue math
// ------------------------------------------------------------------------------------------------
struct EpiArgs {
    const float* dcoef; const float* noise; const float* noise_strength; const float* bias;
    int act; float alpha, gain, clamp;
};
__device__ __forceinline__ float epi_value(float acc, const EpiArgs& e, int b, int co, int C, float noise_term) {
    float v = acc;
    if (e.dcoef) v = fmaf(v, e.dcoef[(int64_t)b * C + co], noise_term);  // fma(x, dcoef, noise), :74
    else v += noise_term;
    if (e.bias) v += e.bias[co];
    return act_gain_clamp(v, e.act, e.alpha, e.gain, e.clamp);
}

// ------------------------------------------------------------------------------------------------
// FIR epilogue for up=2 layers
// ------------------------------------------------------------------------------------------------
// Strip-mined: a thread owns 4 channels of one output column and walks FIR_YT consecutive output rows.  Each raw row it
// reads (4 neighbouring float4, of which the next-column thread of the same block re-reads 3 -> L1 hits) is filtered
// horizontally once and then contributes to the 4 output rows whose window covers it, held as 4 running accumulators;
// every raw element comes from DRAM once (plus a 3/FIR_YT row halo) -- measured traffic = algorithmic bytes
// (profiles/r1_fir_full.txt).  The 4x4 filter must be separable (rank 1), which the resample filter
// outer([1,3,3,1]) of every synthesis layer is: its row/column factors are recovered from the row and column sums.
constexpr int FIR_YT = 32;

template <int ACT>
__global__ void __launch_bounds__(256, 3) fir_epilogue_kernel(ia_fir_params p, int cg, int xt, int cchunks) {
    const int c4 = threadIdx.x % cg;
    const int xl = threadIdx.x / cg;
    const int ox = blockIdx.x * xt + xl;
    const int oy0 = blockIdx.y * FIR_YT;
    const int b = blockIdx.z / cchunks;
    const int c0 = ((blockIdx.z % cchunks) * cg + c4) * 4;
    if (ox >= p.OW || c0 >= p.C) return;
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
    const int oy1 = min(oy0 + FIR_YT, p.OH);
    float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0, acc2 = acc0, acc3 = acc0;   // output rows ry-2 .. ry+1
    const int grp = p.groups > 1 ? b / p.imgs_per_group : 0;
    float dc[4], bs[4], s1[4], s2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        dc[k] = p.dcoef ? p.dcoef[(int64_t)b * p.C + c0 + k] : 1.f;
        bs[k] = p.bias ? p.bias[(int64_t)grp * p.C + c0 + k] : 0.f;
        s1[k] = (p.emit.hi1 && p.emit.s1) ? p.emit.s1[(int64_t)b * p.C + c0 + k] : 1.f;
        s2[k] = (p.emit.hi2 && p.emit.s2) ? p.emit.s2[(int64_t)b * p.C + c0 + k] : 1.f;
    }
    const float nstr = p.noise ? p.noise_strength[grp] : 0.f;
    const float* nptr = p.noise ? p.noise + (int64_t)grp * p.noise_gstride + (int64_t)b * p.noise_bstride + (int64_t)oy0 * p.OW + ox : nullptr;
    const int64_t row_f = (int64_t)p.RW * p.C;
    const float* rp = p.raw + (int64_t)b * p.RH * row_f + (int64_t)(oy0 - 1) * row_f + (int64_t)(ox - 1) * p.C + c0;
    const bool v0 = ox - 1 >= 0, v3 = ox + 2 < p.RW;      // ox, ox+1 always valid (ox < OW = RW-1)
    const int C = p.C;
    int64_t opix = ((int64_t)b * p.OH + oy0) * p.OW + ox;
    const float gain = p.gain, alpha = p.alpha, clampv = p.clamp;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // raw row ry feeds output rows oy = ry+1-ty, ty = 0..3.  Walk ry from oy0-1 to oy1+1.
    float4 n0 = z4, n1 = z4, n2 = z4, n3 = z4;      // next raw row, in flight while the current one is consumed
    auto load_row = [&](int ry, const float* q) {
        n0 = z4; n1 = z4; n2 = z4; n3 = z4;
        if (ry >= 0 && ry < p.RH) {
            if (v0) n0 = __ldg(reinterpret_cast<const float4*>(q));
            n1 = __ldg(reinterpret_cast<const float4*>(q + C));
            n2 = __ldg(reinterpret_cast<const float4*>(q + 2 * C));
            if (v3) n3 = __ldg(reinterpret_cast<const float4*>(q + 3 * C));
        }
    };
    load_row(oy0 - 1, rp);
    for (int ry = oy0 - 1; ry <= oy1 + 1; ++ry, rp += row_f) {
        const float4 r0 = n0, r1 = n1, r2 = n2, r3 = n3;
        if (ry < oy1 + 1) load_row(ry + 1, rp + row_f);
        float4 h;
        h.x = fmaf(fx[3], r3.x, fmaf(fx[2], r2.x, fmaf(fx[1], r1.x, fx[0] * r0.x)));
        h.y = fmaf(fx[3], r3.y, fmaf(fx[2], r2.y, fmaf(fx[1], r1.y, fx[0] * r0.y)));
        h.z = fmaf(fx[3], r3.z, fmaf(fx[2], r2.z, fmaf(fx[1], r1.z, fx[0] * r0.z)));
        h.w = fmaf(fx[3], r3.w, fmaf(fx[2], r2.w, fmaf(fx[1], r1.w, fx[0] * r0.w)));
        // acc_j accumulates output row (ry - 2 + j): raw row ry is its tap ty = 3 - j
        acc0.x = fmaf(fy[3], h.x, acc0.x); acc0.y = fmaf(fy[3], h.y, acc0.y); acc0.z = fmaf(fy[3], h.z, acc0.z); acc0.w = fmaf(fy[3], h.w, acc0.w);
        acc1.x = fmaf(fy[2], h.x, acc1.x); acc1.y = fmaf(fy[2], h.y, acc1.y); acc1.z = fmaf(fy[2], h.z, acc1.z); acc1.w = fmaf(fy[2], h.w, acc1.w);
        acc2.x = fmaf(fy[1], h.x, acc2.x); acc2.y = fmaf(fy[1], h.y, acc2.y); acc2.z = fmaf(fy[1], h.z, acc2.z); acc2.w = fmaf(fy[1], h.w, acc2.w);
        acc3.x = fy[0] * h.x; acc3.y = fy[0] * h.y; acc3.z = fy[0] * h.z; acc3.w = fy[0] * h.w;
        if (ry - 2 >= oy0) {      // output row ry-2 (< oy1 by the loop bound) is complete
            float nz = 0.f;
            if (nptr) { nz = nptr[0] * nstr; nptr += p.OW; }
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
            if (p.emit.out32) {
                float* o = p.emit.out32 + opix * p.emit.out32_ld + c0;
                if ((p.emit.out32_ld & 3) == 0) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                else { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3]; }
            }
            if (p.emit.hi1)
                store_operand4(p.emit.fmt1, p.emit.hi1 + opix * p.emit.c1_pad + c0, p.emit.lo1 + opix * p.emit.c1_pad + c0,
                               v[0] * s1[0], v[1] * s1[1], v[2] * s1[2], v[3] * s1[3]);
            if (p.emit.hi2)
                store_operand4(p.emit.fmt2, p.emit.hi2 + opix * p.emit.c2_pad + c0, p.emit.lo2 + opix * p.emit.c2_pad + c0,
                               v[0] * s2[0], v[1] * s2[1], v[2] * s2[2], v[3] * s2[3]);
            opix += p.OW;
        }
        acc0 = acc1; acc1 = acc2; acc2 = acc3;
    }
}

// Two output columns per thread and two raw rows in flight.  The one-column kernel above keeps a single row of 4 x 16 B per
// thread in flight, of which only 16 B are unique DRAM bytes (the neighbours re-read the rest through L1): ~12 KB of unique
// reads in flight per SM, about a third of what the HBM latency-bandwidth product asks for -- it ran at 3.3 TB/s, issue- and
// latency-bound rather than bandwidth-bound (profiles/r1_fir_full_v3.txt).  Here a thread reads 5 neighbouring float4 per
// raw row for 2 outputs (2.5 loads per output instead of 4) and prefetches two rows ahead (64 B unique per thread in flight).
// NPF (default; IA_FIR_NOISE_PREFETCH=0 disables): the noise value of an output row is loaded one emission ahead instead of in the
// iteration that consumes it (ncu: 34 % of this kernel's stall samples sat on that consumer, profiles/r1_fir_x2_full_v6.txt); same
// values, same arithmetic -- bit-identical on hardware (tests/test_gpu_regress.py), 1.62 -> 1.54 ms per step.
template <int ACT, bool NPF>
__global__ void __launch_bounds__(256, 2) fir_epilogue_x2_kernel(ia_fir_params p, int cg, int xt, int cchunks) {
    const int c4 = threadIdx.x % cg;
    const int xl = threadIdx.x / cg;
    const int ox = (blockIdx.x * xt + xl) * 2;            // even output column; this thread owns ox and ox + 1 (OW is even)
    const int oy0 = blockIdx.y * FIR_YT;
    const int b = blockIdx.z / cchunks;
    const int c0 = ((blockIdx.z % cchunks) * cg + c4) * 4;
    if (ox >= p.OW || c0 >= p.C) return;
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
    const int oy1 = min(oy0 + FIR_YT, p.OH);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc[2][3];                                     // [column][output rows ry-2, ry-1, ry]
#pragma unroll
    for (int j = 0; j < 2; ++j) { acc[j][0] = z4; acc[j][1] = z4; acc[j][2] = z4; }
    const int grp = p.groups > 1 ? b / p.imgs_per_group : 0;
    float dc[4], bs[4], s1[4], s2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        dc[k] = p.dcoef ? p.dcoef[(int64_t)b * p.C + c0 + k] : 1.f;
        bs[k] = p.bias ? p.bias[(int64_t)grp * p.C + c0 + k] : 0.f;
        s1[k] = (p.emit.hi1 && p.emit.s1) ? p.emit.s1[(int64_t)b * p.C + c0 + k] : 1.f;
        s2[k] = (p.emit.hi2 && p.emit.s2) ? p.emit.s2[(int64_t)b * p.C + c0 + k] : 1.f;
    }
    const float nstr = p.noise ? p.noise_strength[grp] : 0.f;
    const float* nptr = p.noise ? p.noise + (int64_t)grp * p.noise_gstride + (int64_t)b * p.noise_bstride + (int64_t)oy0 * p.OW + ox : nullptr;
    const int64_t row_f = (int64_t)p.RW * p.C;
    const float* rp = p.raw + (int64_t)b * p.RH * row_f + (int64_t)(oy0 - 1) * row_f + (int64_t)(ox - 1) * p.C + c0;
    const bool v0 = ox - 1 >= 0, v4 = ox + 3 < p.RW;      // raw columns ox, ox+1, ox+2 always exist (ox + 1 < OW = RW - 1)
    const int C = p.C;
    // running 32-bit element offsets (the launcher checks that every emitted tensor has < 2^31 elements) instead of 64-bit
    // index products per pixel
    const int64_t opix = ((int64_t)b * p.OH + oy0) * p.OW + ox;
    const uint32_t o32ld = (uint32_t)p.emit.out32_ld, c1p = (uint32_t)p.emit.c1_pad, c2p = (uint32_t)p.emit.c2_pad;
    uint32_t o32 = (uint32_t)(opix * o32ld + c0), o1 = (uint32_t)(opix * c1p + c0), o2 = (uint32_t)(opix * c2p + c0);
    const uint32_t o32row = (uint32_t)p.OW * o32ld, r1row = (uint32_t)p.OW * c1p, r2row = (uint32_t)p.OW * c2p;
    const bool has32 = p.emit.out32 != nullptr, has1 = p.emit.hi1 != nullptr, has2 = p.emit.hi2 != nullptr;
    const float gain = p.gain, alpha = p.alpha, clampv = p.clamp;
    float nq0 = 0.f, nq1 = 0.f;                           // NPF: noise of the next output row to be emitted
    int nrows_left = oy1 - oy0;                           // NPF: output rows whose noise has not been requested yet
    if (NPF && nptr) { nq0 = nptr[0]; nq1 = nptr[1]; nptr += p.OW; --nrows_left; }
    float4 ra[5], rb[5];                                  // the next two raw rows, in flight while the current one is consumed
    auto load_row = [&](float4 (&r)[5], int ry, const float* q) {
#pragma unroll
        for (int k = 0; k < 5; ++k) r[k] = z4;
        if (ry >= 0 && ry < p.RH) {
            if (v0) r[0] = __ldg(reinterpret_cast<const float4*>(q));
            r[1] = __ldg(reinterpret_cast<const float4*>(q + C));
            r[2] = __ldg(reinterpret_cast<const float4*>(q + 2 * C));
            r[3] = __ldg(reinterpret_cast<const float4*>(q + 3 * C));
            if (v4) r[4] = __ldg(reinterpret_cast<const float4*>(q + 4 * C));
        }
    };
    load_row(ra, oy0 - 1, rp);
    load_row(rb, oy0, rp + row_f);
#pragma unroll 3
    for (int ry = oy0 - 1; ry <= oy1 + 1; ++ry, rp += row_f) {
        float4 r[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) { r[k] = ra[k]; ra[k] = rb[k]; }
        if (ry + 2 <= oy1 + 1) load_row(rb, ry + 2, rp + 2 * row_f);
        float nz0 = 0.f, nz1 = 0.f;
        const bool emit_row = ry - 2 >= oy0;              // output row ry-2 (< oy1 by the loop bound) completes with this raw row
        if (emit_row && nptr) {
            if (NPF) {
                nz0 = nq0 * nstr; nz1 = nq1 * nstr;
                if (nrows_left > 0) { nq0 = nptr[0]; nq1 = nptr[1]; nptr += p.OW; --nrows_left; }
            } else {
                nz0 = nptr[0] * nstr; nz1 = nptr[1] * nstr; nptr += p.OW;
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float4 h;
            h.x = fmaf(fx[3], r[j + 3].x, fmaf(fx[2], r[j + 2].x, fmaf(fx[1], r[j + 1].x, fx[0] * r[j].x)));
            h.y = fmaf(fx[3], r[j + 3].y, fmaf(fx[2], r[j + 2].y, fmaf(fx[1], r[j + 1].y, fx[0] * r[j].y)));
            h.z = fmaf(fx[3], r[j + 3].z, fmaf(fx[2], r[j + 2].z, fmaf(fx[1], r[j + 1].z, fx[0] * r[j].z)));
            h.w = fmaf(fx[3], r[j + 3].w, fmaf(fx[2], r[j + 2].w, fmaf(fx[1], r[j + 1].w, fx[0] * r[j].w)));
            float4& a0 = acc[j][0]; float4& a1 = acc[j][1]; float4& a2 = acc[j][2];
            a0.x = fmaf(fy[3], h.x, a0.x); a0.y = fmaf(fy[3], h.y, a0.y); a0.z = fmaf(fy[3], h.z, a0.z); a0.w = fmaf(fy[3], h.w, a0.w);
            a1.x = fmaf(fy[2], h.x, a1.x); a1.y = fmaf(fy[2], h.y, a1.y); a1.z = fmaf(fy[2], h.z, a1.z); a1.w = fmaf(fy[2], h.w, a1.w);
            a2.x = fmaf(fy[1], h.x, a2.x); a2.y = fmaf(fy[1], h.y, a2.y); a2.z = fmaf(fy[1], h.z, a2.z); a2.w = fmaf(fy[1], h.w, a2.w);
            const float4 a3 = make_float4(fy[0] * h.x, fy[0] * h.y, fy[0] * h.z, fy[0] * h.w);
            if (emit_row) {
                const float nz = j ? nz1 : nz0;
                const float a4[4] = {a0.x, a0.y, a0.z, a0.w};
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
                    float* o = p.emit.out32 + (o32 + j * o32ld);
                    if ((o32ld & 3) == 0) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                    else { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3]; }
                }
                if (has1)
                    store_operand4(p.emit.fmt1, p.emit.hi1 + (o1 + j * c1p), p.emit.lo1 + (o1 + j * c1p), v[0] * s1[0], v[1] * s1[1], v[2] * s1[2], v[3] * s1[3]);
                if (has2)
                    store_operand4(p.emit.fmt2, p.emit.hi2 + (o2 + j * c2p), p.emit.lo2 + (o2 + j * c2p), v[0] * s2[0], v[1] * s2[1], v[2] * s2[2], v[3] * s2[3]);
            }
            a0 = a1; a1 = a2; a2 = a3;
        }
        if (emit_row) { o32 += o32row; o1 += r1row; o2 += r2row; }
    }
}

// The two-column kernel restructured around what its ncu capture showed (profiles/r2_fir_full_v24.txt: 39 % of the stall
// samples on the long scoreboard -- not on the consumers of a row but on the register MOVES that rotated the prefetched rows
// (r <- ra <- rb), which wait for the newest load and so shorten the prefetch distance to one row; issue slots 60 % busy at
// 316 instructions per thread and row, 42 of them uniform compares / branches on launch-constant flags and 16 register
// zeroings).  Here the three row buffers form a ring that is addressed statically (three copies of the loop body, no moves),
// rows are zero-filled only when they really lie outside the image, the demodulation is always an FMA (coefficient 1 when
// absent: fma(a, 1, n) == a + n bit for bit) and the emission layout is a template parameter for the two layouts the hot path
// uses: EM = 1 -> operand 1 only, single-pass fp16 (backbone up-layers); EM = 2 -> operand 1 only, bf16 hi/lo (super-resolution
// up-layers); EM = 0 -> whatever ia_emit says.  Same arithmetic in the same order as fir_epilogue_kernel: bit-identical
// (tests/test_gpu_regress.py); IA_FIR_RING=0 selects the previous kernel.
template <int ACT, int EM>
__global__ void __launch_bounds__(256, 2) fir_epilogue_x2r_kernel(ia_fir_params p, int cg, int xt, int cchunks) {
    const int c4 = threadIdx.x % cg;
    const int xl = threadIdx.x / cg;
    const int ox = (blockIdx.x * xt + xl) * 2;            // even output column; this thread owns ox and ox + 1 (OW is even)
    const int oy0 = blockIdx.y * FIR_YT;
    const int b = blockIdx.z / cchunks;
    const int c0 = ((blockIdx.z % cchunks) * cg + c4) * 4;
    if (ox >= p.OW || c0 >= p.C) return;
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
    const int oy1 = min(oy0 + FIR_YT, p.OH);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc[2][3];                                     // [column][output rows ry-2, ry-1, ry]
#pragma unroll
    for (int j = 0; j < 2; ++j) { acc[j][0] = z4; acc[j][1] = z4; acc[j][2] = z4; }
    const int grp = p.groups > 1 ? b / p.imgs_per_group : 0;
    const bool has32 = EM == 0 && p.emit.out32 != nullptr;
    const bool has1 = EM != 0 || p.emit.hi1 != nullptr;
    const bool has2 = EM == 0 && p.emit.hi2 != nullptr;
    const int fmt1 = EM == 1 ? IA_OPFMT_F16X1 : (EM == 2 ? IA_OPFMT_BF16X3 : p.emit.fmt1);
    float dc[4], bs[4], s1[4], s2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        dc[k] = p.dcoef ? p.dcoef[(int64_t)b * p.C + c0 + k] : 1.f;
        bs[k] = p.bias ? p.bias[(int64_t)grp * p.C + c0 + k] : 0.f;
        s1[k] = (has1 && p.emit.s1) ? p.emit.s1[(int64_t)b * p.C + c0 + k] : 1.f;
        s2[k] = (has2 && p.emit.s2) ? p.emit.s2[(int64_t)b * p.C + c0 + k] : 1.f;
    }
    const float nstr = p.noise ? p.noise_strength[grp] : 0.f;
    const float* nptr = p.noise ? p.noise + (int64_t)grp * p.noise_gstride + (int64_t)b * p.noise_bstride + (int64_t)oy0 * p.OW + ox : nullptr;
    const int64_t row_f = (int64_t)p.RW * p.C;
    const float* rq = p.raw + (int64_t)b * p.RH * row_f + (int64_t)(oy0 - 1) * row_f + (int64_t)(ox - 1) * p.C + c0;
    const bool v0 = ox - 1 >= 0, v4 = ox + 3 < p.RW;      // raw columns ox, ox+1, ox+2 always exist (ox + 1 < OW = RW - 1)
    const int C = p.C, RH = p.RH, OW = p.OW;
    const int64_t opix = ((int64_t)b * p.OH + oy0) * p.OW + ox;
    const uint32_t o32ld = (uint32_t)p.emit.out32_ld, c1p = (uint32_t)p.emit.c1_pad, c2p = (uint32_t)p.emit.c2_pad;
    uint32_t o32 = has32 ? (uint32_t)(opix * o32ld + c0) : 0u, o1 = has1 ? (uint32_t)(opix * c1p + c0) : 0u, o2 = has2 ? (uint32_t)(opix * c2p + c0) : 0u;
    const uint32_t o32row = (uint32_t)OW * o32ld, r1row = (uint32_t)OW * c1p, r2row = (uint32_t)OW * c2p;
    const bool vec32 = (o32ld & 3) == 0;
    const float gain = p.gain, alpha = p.alpha, clampv = p.clamp;
    const bool clamped = clampv >= 0.f;
    float nq0 = 0.f, nq1 = 0.f;                           // noise of the next output row to be emitted
    int nrows_left = oy1 - oy0;                           // output rows whose noise has not been requested yet
    if (nptr) { nq0 = nptr[0]; nq1 = nptr[1]; nptr += OW; --nrows_left; }
    auto load_row = [&](float4 (&r)[5], int ry, const float* q) {
        if (ry >= 0 && ry < RH) {
            r[0] = z4; r[4] = z4;
            if (v0) r[0] = __ldg(reinterpret_cast<const float4*>(q));
            r[1] = __ldg(reinterpret_cast<const float4*>(q + C));
            r[2] = __ldg(reinterpret_cast<const float4*>(q + 2 * C));
            r[3] = __ldg(reinterpret_cast<const float4*>(q + 3 * C));
            if (v4) r[4] = __ldg(reinterpret_cast<const float4*>(q + 4 * C));
        } else {
#pragma unroll
            for (int k = 0; k < 5; ++k) r[k] = z4;
        }
    };
    // raw row ry: horizontal taps for both columns, vertical accumulation, emission of output row ry - 2 when it completes
    auto consume = [&](const float4 (&r)[5], int ry) {
        const bool emit_row = ry - 2 >= oy0;              // output row ry-2 (< oy1 by the loop bound) completes with this raw row
        float nz0 = 0.f, nz1 = 0.f;
        if (emit_row && nptr) {
            nz0 = nq0 * nstr; nz1 = nq1 * nstr;
            if (nrows_left > 0) { nq0 = nptr[0]; nq1 = nptr[1]; nptr += OW; --nrows_left; }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float4 h;
            h.x = fmaf(fx[3], r[j + 3].x, fmaf(fx[2], r[j + 2].x, fmaf(fx[1], r[j + 1].x, fx[0] * r[j].x)));
            h.y = fmaf(fx[3], r[j + 3].y, fmaf(fx[2], r[j + 2].y, fmaf(fx[1], r[j + 1].y, fx[0] * r[j].y)));
            h.z = fmaf(fx[3], r[j + 3].z, fmaf(fx[2], r[j + 2].z, fmaf(fx[1], r[j + 1].z, fx[0] * r[j].z)));
            h.w = fmaf(fx[3], r[j + 3].w, fmaf(fx[2], r[j + 2].w, fmaf(fx[1], r[j + 1].w, fx[0] * r[j].w)));
            float4& a0 = acc[j][0]; float4& a1 = acc[j][1]; float4& a2 = acc[j][2];
            a0.x = fmaf(fy[3], h.x, a0.x); a0.y = fmaf(fy[3], h.y, a0.y); a0.z = fmaf(fy[3], h.z, a0.z); a0.w = fmaf(fy[3], h.w, a0.w);
            a1.x = fmaf(fy[2], h.x, a1.x); a1.y = fmaf(fy[2], h.y, a1.y); a1.z = fmaf(fy[2], h.z, a1.z); a1.w = fmaf(fy[2], h.w, a1.w);
            a2.x = fmaf(fy[1], h.x, a2.x); a2.y = fmaf(fy[1], h.y, a2.y); a2.z = fmaf(fy[1], h.z, a2.z); a2.w = fmaf(fy[1], h.w, a2.w);
            const float4 a3 = make_float4(fy[0] * h.x, fy[0] * h.y, fy[0] * h.z, fy[0] * h.w);
            if (emit_row) {
                const float nz = j ? nz1 : nz0;
                const float a4[4] = {a0.x, a0.y, a0.z, a0.w};
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float t = fmaf(a4[k], dc[k], nz) + bs[k];                    // fma(x, dcoef, noise), networks_stylegan2_new.py:74
                    if (ACT == IA_ACT_LRELU) t = (t > 0.f ? t : t * alpha) * gain;
                    else if (ACT == IA_ACT_LINEAR) t = t * gain;
                    else t = apply_act(t, p.act, alpha) * gain;
                    v[k] = t;
                }
                if (clamped) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = fminf(fmaxf(v[k], -clampv), clampv);
                }
                if (has32) {
                    float* o = p.emit.out32 + (o32 + j * o32ld);
                    if (vec32) *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
                    else { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3]; }
                }
                if (has1)
                    store_operand4(fmt1, p.emit.hi1 + (o1 + j * c1p), p.emit.lo1 + (o1 + j * c1p), v[0] * s1[0], v[1] * s1[1], v[2] * s1[2], v[3] * s1[3]);
                if (has2)
                    store_operand4(p.emit.fmt2, p.emit.hi2 + (o2 + j * c2p), p.emit.lo2 + (o2 + j * c2p), v[0] * s2[0], v[1] * s2[1], v[2] * s2[2], v[3] * s2[3]);
            }
            a0 = a1; a1 = a2; a2 = a3;
        }
        if (emit_row) { o32 += o32row; o1 += r1row; o2 += r2row; }
    };
    float4 ra[5], rb[5], rc[5];                           // ring of raw rows: one being consumed, two in flight
    int ry = oy0 - 1;
    const int last = oy1 + 1;
    load_row(ra, ry, rq);
    load_row(rb, ry + 1, rq + row_f);
    rq += 2 * row_f;                                      // row ry + 2
#define IA_FIR_STEP(CUR, NXT)                                   \
    if (ry + 2 <= last) load_row(NXT, ry + 2, rq);              \
    rq += row_f;                                                \
    consume(CUR, ry);                                           \
    if (++ry > last) break;
    for (;;) {
        IA_FIR_STEP(ra, rc)
        IA_FIR_STEP(rb, ra)
        IA_FIR_STEP(rc, rb)
    }
#undef IA_FIR_STEP
}

namespace ia {
bool fir_tma_eligible(const ia_fir_params* p);      // ia_fir_tma.cu: TMA-fed streaming variant
int fir_tma_launch(const ia_fir_params* p, void* stream);
}

extern "C" int ia_fir_epilogue(const ia_fir_params* p, void* stream) {
    IA_CHECK(p && p->raw && p->fir, "ia_fir_epilogue: null tensor");
    IA_CHECK((p->C & 3) == 0, "ia_fir_epilogue: C must be a multiple of 4");
    IA_CHECK(p->RH == p->OH + 1 && p->RW == p->OW + 1, "ia_fir_epilogue: raw must be (OH+1) x (OW+1)");
    IA_CHECK(p->noise == nullptr || p->noise_strength != nullptr, "ia_fir_epilogue: noise needs noise_strength");
    if ((int64_t)p->B * p->OH * p->OW * p->C == 0) return 0;
    if (ia::fir_tma_eligible(p)) return ia::fir_tma_launch(p, stream);
    const int groups = p->C >> 2;
    int cg = groups < 32 ? groups : 32;                 // channel groups (of 4) per block row; 32 -> one warp = 512 contiguous bytes
    while (256 % cg) --cg;                              // block is 256 threads = cg x xt
    const int xt = 256 / cg;
    const int cchunks = (int)cdiv(groups, cg);
    int x2 = 1;                                         // IA_FIR_X2=0 selects the one-column kernel (cross-check / profiling)
    { const char* e = getenv("IA_FIR_X2"); if (e) x2 = atoi(e); }
    ia::prof_begin("ia_fir_epilogue", as_stream(stream));
    const int64_t out_pix = (int64_t)p->B * p->OH * p->OW;
    const bool small = out_pix * (p->emit.out32 ? p->emit.out32_ld : 0) < (1ll << 31) && out_pix * (p->emit.hi1 ? p->emit.c1_pad : 0) < (1ll << 31) &&
                       out_pix * (p->emit.hi2 ? p->emit.c2_pad : 0) < (1ll << 31);
    if (x2 && (p->OW & 1) == 0 && small) {
        dim3 grid((unsigned)cdiv(p->OW / 2, xt), (unsigned)cdiv(p->OH, FIR_YT), (unsigned)(p->B * cchunks));
        bool npf = true;       // noise of the next output row loaded one emission ahead (bit-identical; IA_FIR_NOISE_PREFETCH=0: in the consuming iteration)
        { const char* e = getenv("IA_FIR_NOISE_PREFETCH"); if (e && atoi(e) == 0) npf = false; }
        bool ring = npf;       // statically addressed row ring + templated emission layout (IA_FIR_RING=0: the previous two-column kernel)
        { const char* e = getenv("IA_FIR_RING"); if (e && atoi(e) == 0) ring = false; }
        if (ring) {
            const bool only1 = p->emit.hi1 && !p->emit.hi2 && !p->emit.out32;
            const int em = only1 ? (p->emit.fmt1 == IA_OPFMT_F16X1 ? 1 : 2) : 0;
#define IA_FIR_RING_LAUNCH(A)                                                                                              \
            do {                                                                                                           \
                if (em == 1) fir_epilogue_x2r_kernel<A, 1><<<grid, 256, 0, as_stream(stream)>>>(*p, cg, xt, cchunks);       \
                else if (em == 2) fir_epilogue_x2r_kernel<A, 2><<<grid, 256, 0, as_stream(stream)>>>(*p, cg, xt, cchunks);  \
                else fir_epilogue_x2r_kernel<A, 0><<<grid, 256, 0, as_stream(stream)>>>(*p, cg, xt, cchunks);               \
            } while (0)
            if (p->act == IA_ACT_LRELU) IA_FIR_RING_LAUNCH(IA_ACT_LRELU);
            else if (p->act == IA_ACT_LINEAR) IA_FIR_RING_LAUNCH(IA_ACT_LINEAR);
            else IA_FIR_RING_LAUNCH(-1);
#undef IA_FIR_RING_LAUNCH
        } else if (npf) {
            if (p->act == IA_ACT_LRELU) fir_epilogue_x2_kernel<IA_ACT_LRELU, true><<<grid, 256, 0, as_stream(stream)>>>(*p, cg, xt, cchunks);
            else if (p->act == IA_ACT_LINEAR) fir_epilogue_x2_kernel<IA_ACT_LINEAR, true><<<grid, 256, 0, as_stream(stream)>>>(*p, cg, xt, cchunks);
            else fir_epilogue_x2_kernel<-1, true><<<grid, 256, 0, as_stream(stream)>>>(*p, cg, xt, cchunks);
        } else {
            if (p->act == IA_ACT_LRELU) fir_epilogue_x2_kernel<IA_ACT_LRELU, false><<<grid, 256, 0, as_stream(stream)>>>(*p, cg, xt, cchunks);
            else if (p->act == IA_ACT_LINEAR) fir_epilogue_x2_kernel<IA_ACT_LINEAR, false><<<grid, 256, 0, as_stream(stream)>>>(*p, cg, xt, cchunks);
            else fir_epilogue_x2_kernel<-1, false><<<grid, 256, 0, as_stream(stream)>>>(*p, cg, xt, cchunks);
        }
        IA_LAUNCH_CHECK("ia_fir_epilogue");
        return 0;
    }
    dim3 grid((unsigned)cdiv(p->OW, xt), (unsigned)cdiv(p->OH, FIR_YT), (unsigned)(p->B * cchunks));
    if (p->act == IA_ACT_LRELU) fir_epilogue_kernel<IA_ACT_LRELU><<<grid, 256, 0, as_stream(stream)>>>(*p, cg, xt, cchunks);
    else if (p->act == IA_ACT_LINEAR) fir_epilogue_kernel<IA_ACT_LINEAR><<<grid, 256, 0, as_stream(stream)>>>(*p, cg, xt, cchunks);
    else fir_epilogue_kernel<-1><<<grid, 256, 0, as_stream(stream)>>>(*p, cg, xt, cchunks);
    IA_LAUNCH_CHECK("ia_fir_epilogue");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// ToRGB tail: img_out = upsample2d(img_prev) + clamp(raw + bias)
// ------------------------------------------------------------------------------------------------
// one output element: the arithmetic (and its order) every ToRGB-tail kernel below shares
__device__ __forceinline__ float torgb_value(const ia_torgb_params& p, int b, int c, int y, int x) {
    float v = p.raw[(((int64_t)b * p.H + y) * p.W + x) * p.raw_ld + c];
    if (p.bias) v += p.bias[(p.groups > 1 ? b / p.imgs_per_group : 0) * p.C + c];
    if (p.clamp >= 0.f) v = fminf(fmaxf(v, -p.clamp), p.clamp);
    if (p.img_prev) {
        // upsample2d: zero-stuff x2, pad (2,1), 4-tap [1,3,3,1]/8 * 2 per axis.  Even output 2m: .25*x[m-1] + .75*x[m];
        // odd output 2m+1: .75*x[m] + .25*x[m+1]; samples outside the image are zero.
        const int h2 = p.H >> 1, w2 = p.W >> 1;
        int my = y >> 1, mx = x >> 1;
        int y0, y1, x0, x1; float wy0, wy1, wx0, wx1;
        if (y & 1) { y0 = my; y1 = my + 1; wy0 = 0.75f; wy1 = 0.25f; } else { y0 = my - 1; y1 = my; wy0 = 0.25f; wy1 = 0.75f; }
        if (x & 1) { x0 = mx; x1 = mx + 1; wx0 = 0.75f; wx1 = 0.25f; } else { x0 = mx - 1; x1 = mx; wx0 = 0.25f; wx1 = 0.75f; }
        const float* ip = p.img_prev + (int64_t)b * h2 * w2 * p.C + c;
        float up = 0.f;
        // accumulate in the order of the reference's 4x4 correlation (row-major over taps)
        if (y0 >= 0 && y0 < h2) {
            if (x0 >= 0 && x0 < w2) up = fmaf(wy0 * wx0, ip[((int64_t)y0 * w2 + x0) * p.C], up);
            if (x1 >= 0 && x1 < w2) up = fmaf(wy0 * wx1, ip[((int64_t)y0 * w2 + x1) * p.C], up);
        }
        if (y1 >= 0 && y1 < h2) {
            if (x0 >= 0 && x0 < w2) up = fmaf(wy1 * wx0, ip[((int64_t)y1 * w2 + x0) * p.C], up);
            if (x1 >= 0 && x1 < w2) up = fmaf(wy1 * wx1, ip[((int64_t)y1 * w2 + x1) * p.C], up);
        }
        v = up + v;
    }
    return v;
}

__global__ void torgb_finish_kernel(ia_torgb_params p) {
    int64_t total = (int64_t)p.B * p.H * p.W * p.C;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c, x, y, b;
    if (p.out_nchw) {
        x = i % p.W; int64_t t = i / p.W; y = t % p.H; t /= p.H; c = t % p.C; b = (int)(t / p.C);
    } else {
        c = i % p.C; int64_t t = i / p.C; x = t % p.W; t /= p.W; y = t % p.H; b = (int)(t / p.H);
    }
    const float v = torgb_value(p, b, c, y, x);
    if (p.out_nchw) {
        const int64_t o = (((int64_t)b * p.C + c) * p.H + y) * p.W + x;
        p.img_out[o] = v;
        // fused gather of the final frames: the same value goes to this rank's slot of every rank's gathered buffer, through
        // the NVSwitch multicast mapping when there is one (one store, replicated by the switch), else over the peer mappings
        if (p.mc_out) {
            asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p.mc_out + p.peer_offset + o), "f"(v) : "memory");
        } else {
            for (int k = 0; k < p.n_peers; ++k) p.peer_out[k][p.peer_offset + o] = v;
        }
    } else {
        p.img_out[(((int64_t)b * p.H + y) * p.W + x) * p.C + c] = v;
    }
}

// Planar output, 4 adjacent pixels of every channel per thread (W % 4 == 0, 16-byte aligned outputs): the raw NHWC pixels of
// a warp are one contiguous run, every plane row is written as 16-byte stores, and the fused gather issues ONE
// multimem.st.v4.f32 (or one 16-byte store per peer) per 4 values instead of four scalar ones.  Values are torgb_value's.
__global__ void __launch_bounds__(256) torgb_finish_nchw4_kernel(ia_torgb_params p) {
    const int w4 = p.W >> 2;
    const unsigned total = (unsigned)p.B * p.H * w4;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % w4) * 4;
    unsigned t = i / w4;
    const int y = (int)(t % p.H); const int b = (int)(t / p.H);
    for (int c = 0; c < p.C; ++c) {
        const float4 v = make_float4(torgb_value(p, b, c, y, x), torgb_value(p, b, c, y, x + 1), torgb_value(p, b, c, y, x + 2), torgb_value(p, b, c, y, x + 3));
        const int64_t o = (((int64_t)b * p.C + c) * p.H + y) * p.W + x;
        *reinterpret_cast<float4*>(p.img_out + o) = v;
        if (p.mc_out) {
            asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                         ::"l"(p.mc_out + p.peer_offset + o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        } else {
            for (int k = 0; k < p.n_peers; ++k) *reinterpret_cast<float4*>(p.peer_out[k] + p.peer_offset + o) = v;
        }
    }
}

// 4 channels per thread, 32-bit index arithmetic (NHWC output, C % 4 == 0): the element-per-thread kernel above spends its
// time in 64-bit divisions, not in memory.
__global__ void __launch_bounds__(256) torgb_finish_vec4_kernel(ia_torgb_params p) {
    const int groups = p.C >> 2;
    const unsigned total = (unsigned)p.B * p.H * p.W * groups;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % groups) * 4;
    unsigned t = i / groups;
    const int x = (int)(t % p.W); t /= p.W;
    const int y = (int)(t % p.H); const int b = (int)(t / p.H);
    const int64_t pix = ((int64_t)b * p.H + y) * p.W + x;
    float4 v;
    if ((p.raw_ld & 3) == 0) v = __ldg(reinterpret_cast<const float4*>(p.raw + pix * p.raw_ld + c));
    else { const float* q = p.raw + pix * p.raw_ld + c; v = make_float4(q[0], q[1], q[2], q[3]); }
    if (p.bias) { const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + (p.groups > 1 ? b / p.imgs_per_group : 0) * p.C + c)); v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w; }
    if (p.clamp >= 0.f) {
        v.x = fminf(fmaxf(v.x, -p.clamp), p.clamp); v.y = fminf(fmaxf(v.y, -p.clamp), p.clamp);
        v.z = fminf(fmaxf(v.z, -p.clamp), p.clamp); v.w = fminf(fmaxf(v.w, -p.clamp), p.clamp);
    }
    if (p.img_prev) {
        const int h2 = p.H >> 1, w2 = p.W >> 1;
        const int my = y >> 1, mx = x >> 1;
        int y0, y1, x0, x1; float wy0, wy1, wx0, wx1;
        if (y & 1) { y0 = my; y1 = my + 1; wy0 = 0.75f; wy1 = 0.25f; } else { y0 = my - 1; y1 = my; wy0 = 0.25f; wy1 = 0.75f; }
        if (x & 1) { x0 = mx; x1 = mx + 1; wx0 = 0.75f; wx1 = 0.25f; } else { x0 = mx - 1; x1 = mx; wx0 = 0.25f; wx1 = 0.75f; }
        const float* ip = p.img_prev + (int64_t)b * h2 * w2 * p.C + c;
        float4 up = make_float4(0.f, 0.f, 0.f, 0.f);
        auto tap = [&](int yy, int xx, float w) {
            if (yy < 0 || yy >= h2 || xx < 0 || xx >= w2) return;
            const float4 q = __ldg(reinterpret_cast<const float4*>(ip + ((int64_t)yy * w2 + xx) * p.C));
            up.x = fmaf(w, q.x, up.x); up.y = fmaf(w, q.y, up.y); up.z = fmaf(w, q.z, up.z); up.w = fmaf(w, q.w, up.w);
        };
        tap(y0, x0, wy0 * wx0); tap(y0, x1, wy0 * wx1); tap(y1, x0, wy1 * wx0); tap(y1, x1, wy1 * wx1);
        v.x += up.x; v.y += up.y; v.z += up.z; v.w += up.w;
    }
    *reinterpret_cast<float4*>(p.img_out + pix * p.C + c) = v;
}

// the planar 4-pixel kernel needs whole 16-byte groups in every buffer it stores to (IA_TORGB_NCHW4=0: element-per-thread kernel)
static bool torgb_nchw4_ok(const ia_torgb_params* p) {
    { const char* e = getenv("IA_TORGB_NCHW4"); if (e && atoi(e) == 0) return false; }
    if ((p->W & 3) != 0 || (int64_t)p->B * p->H * (p->W >> 2) >= (int64_t)0x7fffffff) return false;
    if ((reinterpret_cast<uintptr_t>(p->img_out) & 15) != 0 || (p->peer_offset & 3) != 0) return false;
    if (p->mc_out && (reinterpret_cast<uintptr_t>(p->mc_out) & 15) != 0) return false;
    for (int k = 0; k < p->n_peers; ++k)
        if ((reinterpret_cast<uintptr_t>(p->peer_out[k]) & 15) != 0) return false;
    return true;
}

extern "C" int ia_torgb_finish(const ia_torgb_params* p, void* stream) {
    IA_CHECK(p && p->raw && p->img_out, "ia_torgb_finish: null tensor");
    IA_CHECK(p->img_prev == nullptr || ((p->H & 1) == 0 && (p->W & 1) == 0), "ia_torgb_finish: odd size with skip image");
    IA_CHECK(p->n_peers >= 0 && p->n_peers <= 8, "ia_torgb_finish: at most 8 peers");
    IA_CHECK((p->n_peers == 0 && p->mc_out == nullptr) || p->out_nchw, "ia_torgb_finish: the fused gather needs the planar (out_nchw) output");
    int64_t total = (int64_t)p->B * p->H * p->W * p->C;
    if (total == 0) return 0;
    ia::prof_begin("ia_torgb_finish", as_stream(stream));
    if (!p->out_nchw && (p->C & 3) == 0 && total / 4 < (int64_t)0x7fffffff &&
        (reinterpret_cast<uintptr_t>(p->img_out) & 15) == 0 && (p->bias == nullptr || (reinterpret_cast<uintptr_t>(p->bias) & 15) == 0) &&
        (p->img_prev == nullptr || (reinterpret_cast<uintptr_t>(p->img_prev) & 15) == 0) && (reinterpret_cast<uintptr_t>(p->raw) & 15) == 0)
        torgb_finish_vec4_kernel<<<(unsigned)cdiv(total / 4, 256), 256, 0, as_stream(stream)>>>(*p);
    else if (p->out_nchw && torgb_nchw4_ok(p))
        torgb_finish_nchw4_kernel<<<(unsigned)cdiv((int64_t)p->B * p->H * (p->W >> 2), 256), 256, 0, as_stream(stream)>>>(*p);
    else
        torgb_finish_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(*p);
    IA_LAUNCH_CHECK("ia_torgb_finish");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// CUDA-core implicit GEMM (64 rows x 64 couts per CTA, 4x4 per thread, fp32 over hi+lo)
// ------------------------------------------------------------------------------------------------
#define SIMT_TM 64
#define SIMT_TN 64
#define SIMT_TK 16
__global__ void __launch_bounds__(256) conv_simt_kernel(ia_conv_params p) {
    __shared__ float As[SIMT_TK][SIMT_TM + 4];
    __shared__ float Bs[SIMT_TK][SIMT_TN + 4];
    const int tid = threadIdx.x;
    const int64_t rows_total = (int64_t)p.B * p.GH * p.GW;
    const int64_t row0 = (int64_t)blockIdx.x * SIMT_TM;
    const int col0 = blockIdx.y * SIMT_TN;
    // loader mapping: thread -> (row lr, 4 consecutive k)
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    int64_t r = row0 + lr;
    bool rvalid = r < rows_total;
    int gx = 0, gy = 0, img = 0;
    if (rvalid) { gx = r % p.GW; int64_t t = r / p.GW; gy = t % p.GH; img = (int)(t / p.GH); }
    const int ty = tid >> 4, tx = tid & 15;  // compute mapping: rows ty*4.., cols tx*4..
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int t = 0; t < p.ntaps; ++t) {
        int iy = gy + p.dy[t], ix = gx + p.dx[t];
        bool avalid = rvalid && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
        const int64_t abase = avalid ? ((((int64_t)img * p.H + iy) * p.W + ix) * p.Cin_pad) : 0;
        const int64_t wrow = ((int64_t)p.wtap[t] * p.Cout_pad + col0 + lr) * p.Cin_pad;
        const bool wvalid = (col0 + lr) < p.Cout_pad;
        for (int k0 = 0; k0 < p.Cin_pad; k0 += SIMT_TK) {
            float av[4] = {0.f, 0.f, 0.f, 0.f}, wv[4] = {0.f, 0.f, 0.f, 0.f};
            if (avalid && p.op_fmt == IA_OPFMT_F16X1) {
                uint2 h = *reinterpret_cast<const uint2*>(p.a_hi + abase + k0 + lk);
                av[0] = f16_bits_to_float(h.x & 0xffff); av[1] = f16_bits_to_float(h.x >> 16);
                av[2] = f16_bits_to_float(h.y & 0xffff); av[3] = f16_bits_to_float(h.y >> 16);
            } else if (avalid) {
                uint2 h = *reinterpret_cast<const uint2*>(p.a_hi + abase + k0 + lk);
                uint2 l = *reinterpret_cast<const uint2*>(p.a_lo + abase + k0 + lk);
                av[0] = bf16_bits_to_float(h.x & 0xffff) + bf16_bits_to_float(l.x & 0xffff);
                av[1] = bf16_bits_to_float(h.x >> 16) + bf16_bits_to_float(l.x >> 16);
                av[2] = bf16_bits_to_float(h.y & 0xffff) + bf16_bits_to_float(l.y & 0xffff);
                av[3] = bf16_bits_to_float(h.y >> 16) + bf16_bits_to_float(l.y >> 16);
            }
            if (wvalid && p.op_fmt == IA_OPFMT_F16X1) {
                uint2 h = *reinterpret_cast<const uint2*>(p.w_hi + wrow + k0 + lk);
                wv[0] = f16_bits_to_float(h.x & 0xffff); wv[1] = f16_bits_to_float(h.x >> 16);
                wv[2] = f16_bits_to_float(h.y & 0xffff); wv[3] = f16_bits_to_float(h.y >> 16);
            } else if (wvalid) {
                uint2 h = *reinterpret_cast<const uint2*>(p.w_hi + wrow + k0 + lk);
                uint2 l = *reinterpret_cast<const uint2*>(p.w_lo + wrow + k0 + lk);
                wv[0] = bf16_bits_to_float(h.x & 0xffff) + bf16_bits_to_float(l.x & 0xffff);
                wv[1] = bf16_bits_to_float(h.x >> 16) + bf16_bits_to_float(l.x >> 16);
                wv[2] = bf16_bits_to_float(h.y & 0xffff) + bf16_bits_to_float(l.y & 0xffff);
                wv[3] = bf16_bits_to_float(h.y >> 16) + bf16_bits_to_float(l.y >> 16);
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 4; ++k) { As[lk + k][lr] = av[k]; Bs[lk + k][lr] = wv[k]; }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < SIMT_TK; ++k) {
                float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
        }
    }
    EpiArgs e{p.dcoef, p.noise, p.noise_strength, p.bias, p.act, p.alpha, p.gain, p.clamp};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t rr = row0 + ty * 4 + i;
        if (rr >= rows_total) continue;
        int ox = rr % p.GW; int64_t t = rr / p.GW; int oy = t % p.GH; int b = (int)(t / p.GH);
        oy = oy * p.sy + p.py; ox = ox * p.sx + p.px;
        if (oy >= p.OH || ox >= p.OW) continue;
        int co = col0 + tx * 4;
        if (co >= p.Cout) continue;
        float v[4];
        if (p.mode == 1) {
            float nz = p.noise ? p.noise[(int64_t)b * p.noise_bstride + (int64_t)oy * p.OW + ox] * p.noise_strength[0] : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (co + j < p.Cout) ? epi_value(acc[i][j], e, b, co + j, p.Cout, nz) : 0.f;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = acc[i][j];
        }
        emit4(p.emit, b, ((int64_t)b * p.OH + oy) * p.OW + ox, co, p.Cout, v);
    }
}

int ia_conv_validate(const ia_conv_params* p, const char* who) {
    IA_CHECK(p && (p->op_fmt == IA_OPFMT_BF16X3 || p->op_fmt == IA_OPFMT_F16X1), "%s: unknown operand format", who);
    IA_CHECK(p->a_hi && p->w_hi && (p->op_fmt == IA_OPFMT_F16X1 || (p->a_lo && p->w_lo)), "%s: null operand", who);
    IA_CHECK(p->Cin_pad > 0 && (p->Cin_pad % 64) == 0, "%s: Cin_pad must be a multiple of 64 (got %d)", who, p->Cin_pad);
    IA_CHECK(p->Cout > 0 && p->Cout_pad >= p->Cout && (p->Cout_pad % 32) == 0, "%s: Cout_pad must be a multiple of 32 >= Cout", who);
    IA_CHECK(p->ntaps >= 1 && p->ntaps <= 9, "%s: ntaps must be in [1,9]", who);
    for (int t = 0; t < p->ntaps; ++t)
        IA_CHECK(p->wtap[t] >= 0 && p->wtap[t] < p->n_taps_total, "%s: tap %d weight index out of range", who, t);
    IA_CHECK(p->GH > 0 && p->GW > 0 && p->B > 0 && p->H > 0 && p->W > 0, "%s: empty geometry", who);
    IA_CHECK(p->sy >= 1 && p->sx >= 1, "%s: bad output stride", who);
    IA_CHECK(p->mode == 0 || p->mode == 1 || p->mode == 2, "%s: bad epilogue mode", who);
    IA_CHECK(p->a_img_rows == 0 || p->a_img_rows == p->H || (p->a_img_rows == p->H + 1 && p->mode == 0 && p->groups <= 1),
             "%s: a_img_rows must be H, or H+1 (zero row after every image) with mode 0 and no groups", who);
    IA_CHECK(p->emit.e1_img_pix == 0 || p->emit.e1_img_pix >= (int64_t)p->OH * p->OW, "%s: emit-1 image stride smaller than the image", who);
    IA_CHECK(p->mode != 2 || (p->emit.out32 && !p->emit.hi1 && !p->emit.hi2 && !p->emit.rgb_out && p->sy == 1 && p->sx == 1 && (p->Cout & 3) == 0 &&
                              (!p->img_prev || ((p->OH & 1) == 0 && (p->OW & 1) == 0))),
             "%s: mode 2 (ToRGB tail) writes out32 only, stride 1, Cout %% 4 == 0, even output size with img_prev", who);
    IA_CHECK(p->noise == nullptr || p->noise_strength != nullptr, "%s: noise needs noise_strength", who);
    IA_CHECK(p->act != IA_ACT_PRELU || (p->mode == 1 && p->slope && !p->emit.hi2 && !p->emit.rgb_out && (p->Cout & 3) == 0 && p->groups <= 1 &&
                                        (reinterpret_cast<uintptr_t>(p->slope) & 15) == 0),
             "%s: PReLU epilogue needs mode 1, 16-byte aligned slope[Cout], Cout %% 4 == 0, no emit 2 / fused ToRGB / groups", who);
    IA_CHECK(p->emit.out32 || p->emit.hi1 || p->emit.hi2 || p->emit.rgb_out, "%s: nothing to emit", who);
    IA_CHECK(!p->emit.rgb_out || (p->emit.rgb_w && p->emit.rgb_n >= 1 && p->emit.rgb_n <= 4 && p->mode == 1 && (p->Cout & 3) == 0 &&
                                  !p->emit.out32 && !p->emit.hi2),
             "%s: fused ToRGB needs rgb_w, 1..4 outputs, mode 1, Cout %% 4 == 0 and neither out32 nor emit 2", who);
    IA_CHECK(p->groups <= 1 || (p->imgs_per_group > 0 && p->B == p->groups * p->imgs_per_group), "%s: grouped launch needs B == groups * imgs_per_group", who);
    IA_CHECK(!p->emit.hi1 || ((p->emit.lo1 || p->emit.fmt1 == IA_OPFMT_F16X1) && p->emit.c1_pad >= p->Cout && (p->emit.c1_pad & 3) == 0), "%s: bad emit 1", who);
    IA_CHECK(!p->emit.hi2 || ((p->emit.lo2 || p->emit.fmt2 == IA_OPFMT_F16X1) && p->emit.c2_pad >= p->Cout && (p->emit.c2_pad & 3) == 0), "%s: bad emit 2", who);
    return 0;
}

extern "C" int ia_conv_simt(const ia_conv_params* p, void* stream) {
    if (int rc = ia_conv_validate(p, "ia_conv_simt")) return rc;
    IA_CHECK(p->groups <= 1, "ia_conv_simt: grouped launches are implemented by ia_conv_tc only");
    IA_CHECK(!p->emit.rgb_out, "ia_conv_simt: the fused ToRGB contraction is implemented by ia_conv_tc only");
    IA_CHECK(p->mode != 2, "ia_conv_simt: the fused ToRGB tail (mode 2) is implemented by ia_conv_tc only");
    IA_CHECK(p->act != IA_ACT_PRELU, "ia_conv_simt: the PReLU epilogue is implemented by ia_conv_tc only");
    IA_CHECK(p->a_img_rows <= p->H && p->emit.e1_img_pix == 0, "ia_conv_simt: padded operand layouts are implemented by ia_conv_tc only");
    int64_t rows = (int64_t)p->B * p->GH * p->GW;
    dim3 grid((unsigned)cdiv(rows, SIMT_TM), (unsigned)cdiv(p->Cout_pad, SIMT_TN));
    ia::prof_begin("ia_conv_simt", as_stream(stream));
    conv_simt_kernel<<<grid, 256, 0, as_stream(stream)>>>(*p);
    IA_LAUNCH_CHECK("ia_conv_simt");
    return 0;
}
