// Mix-Transformer pieces of the "improved one-shot" inversion encoder (reference encoder_inversion/models/mmseg/
// mix_transformer.py: OverlapPatchEmbed :159-198, Attention :56-115, Mlp/DWConv :18-53,379-390, Block :118-156,
// transformer_block :453-472).  The dense contractions (patch embedding, q / kv / proj / fc1 / fc2) run on ia_conv_tc as
// 1x1 convolutions over token maps stored [B][H][W][C] (tokens == pixels of an NHWC image, so no reshapes exist); this file
// holds what sits between them:
//   ia_enc_im2col     k x k / stride s patch gather of a (concatenated, pixel-shuffled) view -> bf16 hi/lo GEMM operand
//   ia_layer_norm     per-token LayerNorm (+ optional pre-bias) -> fp32 tokens and / or the next GEMM operand
//   ia_attention      softmax(q k^T * scale) v, flash-style (no N x N matrix in memory), fp32 on CUDA cores
//   ia_dwconv_gelu    depthwise 3x3 + bias + exact GELU of the fc1 output -> fc2 operand
#include "ia_common.cuh"

using namespace ia;

namespace {

__device__ __forceinline__ float view_at(const ia_view& v, int b, int y, int x, int c) {
    if (v.ps == 1) return v.p[(int64_t)b * v.s_img + (int64_t)y * v.s_row + (int64_t)x * v.s_pix + (int64_t)c * v.s_c];
    const int ps = v.ps;
    const int cc = c * ps * ps + (y % ps) * ps + (x % ps);
    return v.p[(int64_t)b * v.s_img + (int64_t)(y / ps) * v.s_row + (int64_t)(x / ps) * v.s_pix + (int64_t)cc * v.s_c];
}

// ---- im2col ------------------------------------------------------------------------------------------------------------
// out[b][oy][ox][(ky*k + kx)*Ctot + c] = cat(src)[b][oy*s - pad + ky][ox*s - pad + kx][c]  (0 outside the image / beyond K)
__global__ void __launch_bounds__(256) im2col_kernel(ia_enc_im2col_params p, int Ctot, int K) {
    const int groups = p.K_pad >> 2;
    const int64_t total = (int64_t)p.B * p.OH * p.OW * groups;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i % groups);
    const int64_t pix = i / groups;
    const int ox = (int)(pix % p.OW); const int64_t t = pix / p.OW;
    const int oy = (int)(t % p.OH); const int b = (int)(t / p.OH);
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int kk = g * 4 + e;
        float a = 0.f;
        if (kk < K) {
            const int tap = kk / Ctot;
            int c = kk - tap * Ctot;
            const int ky = tap / p.k, kx = tap - ky * p.k;
            const int y = oy * p.stride - p.pad + ky, x = ox * p.stride - p.pad + kx;
            if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
                int s = 0;
                while (s < p.nsrc - 1 && c >= p.src[s].C) { c -= p.src[s].C; ++s; }
                a = view_at(p.src[s], b, y, x, c);
            }
        }
        v[e] = a;
    }
    store_operand4(IA_OPFMT_BF16X3, p.hi + pix * p.K_pad + g * 4, p.lo + pix * p.K_pad + g * 4, v[0], v[1], v[2], v[3]);
}

// ---- LayerNorm: one warp per token; mean, centred variance (biased), affine ------------------------------------------------
__global__ void __launch_bounds__(256) layer_norm_kernel(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ pre_bias,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                         int64_t rows, int C, float* __restrict__ out32, int64_t out32_ld,
                                                         uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int C_pad) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* xr = x + row * x_ld;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c] + (pre_bias ? pre_bias[c] : 0.f);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] + (pre_bias ? pre_bias[c] : 0.f) - mean; q = fmaf(d, d, q); }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)C + eps);
    // 4 consecutive channels per lane and step, so operand stores are 8-byte
    for (int c0 = lane * 4; c0 < (hi ? C_pad : C); c0 += 128) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = c0 + k;
            v[k] = c < C ? fmaf((xr[c] + (pre_bias ? pre_bias[c] : 0.f) - mean) * rstd, gamma[c], beta[c]) : 0.f;
        }
        if (out32) for (int k = 0; k < 4; ++k) if (c0 + k < C) out32[row * out32_ld + c0 + k] = v[k];
        if (hi) store_operand4(IA_OPFMT_BF16X3, hi + row * C_pad + c0, lo + row * C_pad + c0, v[0], v[1], v[2], v[3]);
    }
}

// ---- attention ---------------------------------------------------------------------------------------------------------
// One CTA = 64 queries of one (image, head); keys / values stream through shared memory in tiles of 64; online softmax
// (running max / sum per query row) in the exp2 domain.  256 threads as 16 (ty) x 16 (tx): thread (ty,tx) owns query rows
// 4ty..4ty+3, score columns tx+16j (j<4) and output dims 4tx+64j..+3 (j < HD/64).
template <int HD>
struct AttnSmem {
    static constexpr int QLD = HD + 4;      // row pitch (floats): +4 keeps float4 alignment and shifts rows by 4 banks
    static constexpr int PLD = 68;
    float q[64 * QLD];
    float k[64 * QLD];
    float v[64 * HD];
    float p[64 * PLD];
};

template <int HD>
__global__ void __launch_bounds__(256, 1) attention_kernel(ia_attention_params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AttnSmem<HD>& sm = *reinterpret_cast<AttnSmem<HD>*>(smem_raw);
    constexpr int QLD = AttnSmem<HD>::QLD, PLD = AttnSmem<HD>::PLD;
    constexpr int DJ = HD / 64;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
    const int coff = h * HD;
    const float qscale = p.scale * 1.4426950408889634f;        // softmax in base 2
    const float* qg = p.q + (int64_t)b * p.Nq * p.q_ld + coff;
    const float* kg = p.k + (int64_t)b * p.Nk * p.k_ld + coff;
    const float* vg = p.v + (int64_t)b * p.Nk * p.v_ld + coff;

    // Q tile (bias added, pre-scaled); rows past Nq are zero
    for (int i = tid; i < 64 * (HD / 4); i += 256) {
        const int r = i / (HD / 4), d = (i - r * (HD / 4)) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < p.Nq) {
            a = *reinterpret_cast<const float4*>(qg + (int64_t)(q0 + r) * p.q_ld + d);
            if (p.q_bias) { const float4 bb = *reinterpret_cast<const float4*>(p.q_bias + coff + d); a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w; }
            a.x *= qscale; a.y *= qscale; a.z *= qscale; a.w *= qscale;
        }
        *reinterpret_cast<float4*>(&sm.q[r * QLD + d]) = a;
    }
    float o[4][DJ * 4];
    float m[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m[i] = -INFINITY; l[i] = 0.f;
#pragma unroll
        for (int j = 0; j < DJ * 4; ++j) o[i][j] = 0.f;
    }
    for (int k0 = 0; k0 < p.Nk; k0 += 64) {
        __syncthreads();           // previous tile's K / V / P fully consumed (and the Q tile written, first time round)
        for (int i = tid; i < 64 * (HD / 4); i += 256) {
            const int r = i / (HD / 4), d = (i - r * (HD / 4)) * 4;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
            if (k0 + r < p.Nk) {
                a = *reinterpret_cast<const float4*>(kg + (int64_t)(k0 + r) * p.k_ld + d);
                c = *reinterpret_cast<const float4*>(vg + (int64_t)(k0 + r) * p.v_ld + d);
                if (p.k_bias) { const float4 bb = *reinterpret_cast<const float4*>(p.k_bias + coff + d); a.x += bb.x; a.y += bb.y; a.z += bb.z; a.w += bb.w; }
                if (p.v_bias) { const float4 bb = *reinterpret_cast<const float4*>(p.v_bias + coff + d); c.x += bb.x; c.y += bb.y; c.z += bb.z; c.w += bb.w; }
            }
            *reinterpret_cast<float4*>(&sm.k[r * QLD + d]) = a;
            *reinterpret_cast<float4*>(&sm.v[r * HD + d]) = c;
        }
        __syncthreads();
        // scores: s[i][j] = q[4ty+i] . k[tx+16j]
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 4
        for (int d = 0; d < HD; d += 4) {
            float4 qa[4], ka[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) qa[i] = *reinterpret_cast<const float4*>(&sm.q[(ty * 4 + i) * QLD + d]);
#pragma unroll
            for (int j = 0; j < 4; ++j) ka[j] = *reinterpret_cast<const float4*>(&sm.k[(tx + 16 * j) * QLD + d]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    s[i][j] = fmaf(qa[i].x, ka[j].x, s[i][j]);
                    s[i][j] = fmaf(qa[i].y, ka[j].y, s[i][j]);
                    s[i][j] = fmaf(qa[i].z, ka[j].z, s[i][j]);
                    s[i][j] = fmaf(qa[i].w, ka[j].w, s[i][j]);
                }
        }
        // online softmax over the 64 columns of this tile (16 tx lanes x 4 columns each)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (k0 + tx + 16 * j >= p.Nk) s[i][j] = -INFINITY;
                mx = fmaxf(mx, s[i][j]);
            }
#pragma unroll
            for (int w = 8; w; w >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, w));
            const float mn = fmaxf(m[i], mx);            // finite: every tile has at least one valid column
            const float corr = exp2f(m[i] - mn);
            float rs = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float e = exp2f(s[i][j] - mn);
                rs += e;
                sm.p[(ty * 4 + i) * PLD + tx + 16 * j] = e;
            }
#pragma unroll
            for (int w = 8; w; w >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, w);
            l[i] = l[i] * corr + rs;
            m[i] = mn;
#pragma unroll
            for (int j = 0; j < DJ * 4; ++j) o[i][j] *= corr;
        }
        __syncthreads();
        // o[i][dims] += sum_c p[4ty+i][c] * v[c][dims]
#pragma unroll 2
        for (int c = 0; c < 64; c += 4) {
            float4 pa[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) pa[i] = *reinterpret_cast<const float4*>(&sm.p[(ty * 4 + i) * PLD + c]);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
                for (int j = 0; j < DJ; ++j) {
                    const float4 vv = *reinterpret_cast<const float4*>(&sm.v[(c + cc) * HD + tx * 4 + 64 * j]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float pe = cc == 0 ? pa[i].x : cc == 1 ? pa[i].y : cc == 2 ? pa[i].z : pa[i].w;
                        o[i][j * 4 + 0] = fmaf(pe, vv.x, o[i][j * 4 + 0]);
                        o[i][j * 4 + 1] = fmaf(pe, vv.y, o[i][j * 4 + 1]);
                        o[i][j * 4 + 2] = fmaf(pe, vv.z, o[i][j * 4 + 2]);
                        o[i][j * 4 + 3] = fmaf(pe, vv.w, o[i][j * 4 + 3]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = q0 + ty * 4 + i;
        if (r >= p.Nq) continue;
        const float inv = 1.f / l[i];
        const int64_t row = (int64_t)b * p.Nq + r;
#pragma unroll
        for (int j = 0; j < DJ; ++j) {
            const int c = coff + tx * 4 + 64 * j;
            const float a0 = o[i][j * 4] * inv, a1 = o[i][j * 4 + 1] * inv, a2 = o[i][j * 4 + 2] * inv, a3 = o[i][j * 4 + 3] * inv;
            if (p.out32) *reinterpret_cast<float4*>(p.out32 + row * p.out32_ld + c) = make_float4(a0, a1, a2, a3);
            if (p.hi) store_operand4(IA_OPFMT_BF16X3, p.hi + row * p.C_pad + c, p.lo + row * p.C_pad + c, a0, a1, a2, a3);
        }
    }
}

// ---- depthwise 3x3 (+ input bias, + bias) + GELU -> operand -----------------------------------------------------------
__global__ void __launch_bounds__(256) dwconv_gelu_kernel(const float* __restrict__ x, const float* __restrict__ in_bias,
                                                          const float* __restrict__ w, const float* __restrict__ bias, int B, int H, int W,
                                                          int C, float* __restrict__ out32, uint16_t* __restrict__ hi,
                                                          uint16_t* __restrict__ lo, int C_pad) {
    const int groups = C_pad >> 2;
    const int64_t total = (int64_t)B * H * W * groups;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i % groups);
    const int64_t pix = i / groups;
    const int xx = (int)(pix % W); const int64_t t = pix / W;
    const int yy = (int)(t % H); const int b = (int)(t / H);
    const int c0 = g * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (c0 < C) {      // C % 4 == 0 (checked by the host entry)
        const float4 ib = in_bias ? *reinterpret_cast<const float4*>(in_bias + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
        float acc[4];
        { const float4 bb = *reinterpret_cast<const float4*>(bias + c0); acc[0] = bb.x; acc[1] = bb.y; acc[2] = bb.z; acc[3] = bb.w; }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int y = yy + ky - 1;
            if (y < 0 || y >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xq = xx + kx - 1;
                if (xq < 0 || xq >= W) continue;
                const float4 a = *reinterpret_cast<const float4*>(x + (((int64_t)b * H + y) * W + xq) * C + c0);
                acc[0] = fmaf(a.x + ib.x, w[(c0 + 0) * 9 + ky * 3 + kx], acc[0]);
                acc[1] = fmaf(a.y + ib.y, w[(c0 + 1) * 9 + ky * 3 + kx], acc[1]);
                acc[2] = fmaf(a.z + ib.z, w[(c0 + 2) * 9 + ky * 3 + kx], acc[2]);
                acc[3] = fmaf(a.w + ib.w, w[(c0 + 3) * 9 + ky * 3 + kx], acc[3]);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = 0.5f * acc[k] * (1.f + erff(acc[k] * 0.70710678118654752f));     // exact GELU (nn.GELU default)
        if (out32) *reinterpret_cast<float4*>(out32 + pix * C + c0) = make_float4(v[0], v[1], v[2], v[3]);
    }
    if (hi) store_operand4(IA_OPFMT_BF16X3, hi + pix * C_pad + c0, lo + pix * C_pad + c0, v[0], v[1], v[2], v[3]);
}

int check_view_(const ia_view* v, const char* who) {
    IA_CHECK(v && v->p, "%s: null view", who);
    IA_CHECK(v->C > 0 && v->ps >= 1, "%s: bad view (C=%d ps=%d)", who, v ? v->C : 0, v ? v->ps : 0);
    return 0;
}

template <int HD>
int launch_attention(const ia_attention_params* p, cudaStream_t st) {
    const size_t smem = sizeof(AttnSmem<HD>);
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(attention_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        IA_CHECK(e == cudaSuccess, "ia_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set[dev] = true;
    }
    dim3 grid((unsigned)cdiv(p->Nq, 64), (unsigned)p->heads, (unsigned)p->B);
    ia::prof_begin("ia_attention", st);
    attention_kernel<HD><<<grid, 256, smem, st>>>(*p);
    IA_LAUNCH_CHECK("ia_attention");
    return 0;
}

}  // namespace

extern "C" int ia_enc_im2col(const ia_enc_im2col_params* p, void* stream) {
    IA_CHECK(p && p->nsrc >= 1 && p->nsrc <= 4, "ia_enc_im2col: 1..4 sources");
    int Ctot = 0;
    for (int s = 0; s < p->nsrc; ++s) {
        if (int rc = check_view_(&p->src[s], "ia_enc_im2col")) return rc;
        Ctot += p->src[s].C;
    }
    IA_CHECK(p->hi && p->lo, "ia_enc_im2col: null output");
    IA_CHECK(p->k >= 1 && p->stride >= 1 && p->pad >= 0 && p->B > 0 && p->H > 0 && p->W > 0, "ia_enc_im2col: bad geometry");
    IA_CHECK(p->OH == (p->H + 2 * p->pad - p->k) / p->stride + 1 && p->OW == (p->W + 2 * p->pad - p->k) / p->stride + 1,
             "ia_enc_im2col: output size %dx%d does not match floor((H + 2 pad - k) / stride) + 1", p->OH, p->OW);
    const int64_t K = (int64_t)p->k * p->k * Ctot;
    IA_CHECK(K <= p->K_pad && (p->K_pad & 3) == 0 && K < (1ll << 30), "ia_enc_im2col: K_pad (%d) must be >= k*k*C (%lld) and a multiple of 4",
             p->K_pad, (long long)K);
    const int64_t total = (int64_t)p->B * p->OH * p->OW * (p->K_pad >> 2);
    if (total == 0) return 0;
    ia::prof_begin("ia_enc_im2col", as_stream(stream));
    im2col_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(*p, Ctot, (int)K);
    IA_LAUNCH_CHECK("ia_enc_im2col");
    return 0;
}

extern "C" int ia_layer_norm(const float* x, int64_t x_ld, const float* pre_bias, const float* gamma, const float* beta, float eps,
                             int64_t rows, int32_t C, float* out32, int64_t out32_ld, uint16_t* hi, uint16_t* lo, int32_t C_pad,
                             void* stream) {
    IA_CHECK(x && gamma && beta && C > 0 && x_ld >= C, "ia_layer_norm: bad arguments");
    IA_CHECK((hi != nullptr) == (lo != nullptr) && (hi || out32), "ia_layer_norm: no output");
    IA_CHECK(!hi || (C_pad >= C && (C_pad & 3) == 0), "ia_layer_norm: C_pad (%d) must be >= C (%d) and a multiple of 4", C_pad, C);
    IA_CHECK(!out32 || out32_ld >= C, "ia_layer_norm: out32_ld < C");
    if (rows == 0) return 0;
    ia::prof_begin("ia_layer_norm", as_stream(stream));
    layer_norm_kernel<<<(unsigned)cdiv(rows, 8), 256, 0, as_stream(stream)>>>(x, x_ld, pre_bias, gamma, beta, eps, rows, C, out32, out32_ld,
                                                                            hi, lo, C_pad);
    IA_LAUNCH_CHECK("ia_layer_norm");
    return 0;
}

extern "C" int ia_attention(const ia_attention_params* p, void* stream) {
    IA_CHECK(p && p->q && p->k && p->v, "ia_attention: null input");
    IA_CHECK(p->B > 0 && p->heads > 0 && p->Nq > 0 && p->Nk > 0, "ia_attention: bad sizes");
    IA_CHECK(p->head_dim == 64 || p->head_dim == 256, "ia_attention: head_dim %d not instantiated (64: MixVisionTransformer, 256: transformer_block)",
             p->head_dim);
    const int64_t Cc = (int64_t)p->heads * p->head_dim;
    IA_CHECK(p->q_ld >= Cc && p->k_ld >= Cc && p->v_ld >= Cc && ((p->q_ld | p->k_ld | p->v_ld) & 3) == 0, "ia_attention: row pitches must be >= heads*head_dim and multiples of 4");
    IA_CHECK((((uintptr_t)p->q | (uintptr_t)p->k | (uintptr_t)p->v) & 15) == 0, "ia_attention: q / k / v must be 16-byte aligned");
    IA_CHECK((p->hi != nullptr) == (p->lo != nullptr) && (p->hi || p->out32), "ia_attention: no output");
    IA_CHECK(!p->hi || p->C_pad >= Cc, "ia_attention: C_pad < heads*head_dim");
    IA_CHECK(!p->out32 || (p->out32_ld >= Cc && (p->out32_ld & 3) == 0), "ia_attention: out32_ld");
    if (p->hi && p->C_pad > Cc) {      // padding channels of the operand are the caller's to zero; we only write [0, C)
        IA_CHECK(false, "ia_attention: C_pad (%d) must equal heads*head_dim (%lld)", p->C_pad, (long long)Cc);
    }
    return p->head_dim == 64 ? launch_attention<64>(p, as_stream(stream)) : launch_attention<256>(p, as_stream(stream));
}

extern "C" int ia_dwconv_gelu(const float* x, const float* in_bias, const float* w, const float* bias, int32_t B, int32_t H, int32_t W,
                              int32_t C, float* out32, uint16_t* hi, uint16_t* lo, int32_t C_pad, void* stream) {
    IA_CHECK(x && w && bias && B > 0 && H > 0 && W > 0 && C > 0 && (C & 3) == 0, "ia_dwconv_gelu: bad arguments (C must be a multiple of 4)");
    IA_CHECK((hi != nullptr) == (lo != nullptr) && (hi || out32), "ia_dwconv_gelu: no output");
    if (!hi) C_pad = C;
    IA_CHECK(C_pad >= C && (C_pad & 3) == 0, "ia_dwconv_gelu: C_pad");
    const int64_t total = (int64_t)B * H * W * (C_pad >> 2);
    ia::prof_begin("ia_dwconv_gelu", as_stream(stream));
    dwconv_gelu_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(x, in_bias, w, bias, B, H, W, C, out32, hi, lo, C_pad);
    IA_LAUNCH_CHECK("ia_dwconv_gelu");
    return 0;
}
