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

// ---- attention on the tensor cores (head_dim 256: transformer_block) -----------------------------------------------------
// Operands arrive as the bf16 hi/lo splits the q / kv projections emit from their epilogues ([B][N][C_pad], head h at channel
// h*256); products are the 3-term split hi*hi + hi*lo + lo*hi (fp32-grade) on mma.sync.m16n8k16 with fp32 accumulators, for
// both q k^T and p v (p in [0,1] is split in registers).  One CTA = 128 queries of one (image, head), 8 warps x 16 query rows;
// keys / values stream through two 32-key shared-memory tiles (K and V separately, so the load of the next K tile overlaps
// softmax + p v and the load of the next V tile overlaps q k^T).  Rows are 512 bytes; 16-byte chunks are XOR-swizzled with the
// row index so every ldmatrix phase touches 8 distinct bank groups.
template <int QW>               // QW = warps per CTA = 16-query row groups per CTA
struct AttnTcSmem {
    uint16_t q[2][16 * QW * 256];  // hi, lo
    uint16_t k[2][32 * 256];
    uint16_t v[2][32 * 256];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;      // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of 16-byte chunk `chunk` (0..31) of row `row` inside a [rows][256] bf16 tile
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(row * 512 + ((chunk ^ (row & 7)) << 4)); }

template <int QW>
__global__ void __launch_bounds__(32 * QW, 1) attention_tc_kernel(ia_attention_tc_params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    AttnTcSmem<QW>& sm = *reinterpret_cast<AttnTcSmem<QW>*>(smem_raw);
    constexpr int NT = 32 * QW;          // threads
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int q0 = blockIdx.x * (16 * QW), h = blockIdx.y, b = blockIdx.z;
    const float sscale = p.scale * 1.4426950408889634f;
    const uint16_t* qsrc[2] = {p.q_hi + ((int64_t)b * p.Nq) * p.q_ld + h * 256, p.q_lo + ((int64_t)b * p.Nq) * p.q_ld + h * 256};
    const uint16_t* ksrc[2] = {p.kv_hi + ((int64_t)b * p.Nk) * p.kv_ld + h * 256, p.kv_lo + ((int64_t)b * p.Nk) * p.kv_ld + h * 256};
    const int v_off = p.heads * 256;       // v follows k inside a kv row
    const uint32_t q_s[2] = {smem_u32(sm.q[0]), smem_u32(sm.q[1])};
    const uint32_t k_s[2] = {smem_u32(sm.k[0]), smem_u32(sm.k[1])};
    const uint32_t v_s[2] = {smem_u32(sm.v[0]), smem_u32(sm.v[1])};

    auto load_kv_tile = [&](const uint32_t (&dst)[2], int k0, int coff) {
        // 32 rows x 32 chunks x {hi, lo} = 2048 chunks of 16 bytes, 2048 / NT per thread
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
            for (int i = 0; i < 1024 / NT; ++i) {
                const int idx = tid + i * NT;
                const int row = idx >> 5, chunk = idx & 31;
                const bool ok = k0 + row < p.Nk;
                const uint16_t* src = ksrc[part] + (int64_t)(ok ? k0 + row : 0) * p.kv_ld + coff + chunk * 8;
                cp_async16(dst[part] + swz(row, chunk), src, ok);
            }
    };
    // group 0: Q tile + first K tile; group 1: first V tile
#pragma unroll
    for (int part = 0; part < 2; ++part)
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {       // 16 * QW rows x 32 chunks / NT threads = 16 per thread and part
            const int idx = tid + i * NT;
            const int row = idx >> 5, chunk = idx & 31;
            const bool ok = q0 + row < p.Nq;
            cp_async16(q_s[part] + swz(row, chunk), qsrc[part] + (int64_t)(ok ? q0 + row : 0) * p.q_ld + chunk * 8, ok);
        }
    load_kv_tile(k_s, 0, 0);
    cp_async_commit();
    load_kv_tile(v_s, 0, v_off);
    cp_async_commit();

    float o[32][4];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;       // rows g and g+8 of this warp's 16
    const int arow = warp * 16 + (lane & 15);       // ldmatrix row of the A (query) fragments
    const int achk = lane >> 4;                     // + 0 / 1 chunk (k 0-7 / 8-15)
    // B fragments of q k^T (K rows = keys): lanes 0-7 keys n0..7 @k0, 8-15 same keys @k0+8, 16-23 keys n0+8.. @k0, 24-31 @k0+8
    const int brow = (lane & 7) + ((lane >> 4) << 3);
    const int bchk = (lane >> 3) & 1;
    // B fragments of p v (V rows = keys, transposed load): lanes 0-7 keys 0-7 @dim n0, 8-15 keys 8-15 @n0, 16-23 keys 0-7 @n0+8, 24-31 keys 8-15 @n0+8
    const int vrow = (lane & 7) + (((lane >> 3) & 1) << 3);
    const int vchk = lane >> 4;

    for (int k0 = 0; k0 < p.Nk; k0 += 32) {
        cp_async_wait<1>();          // K tile of this iteration (and Q) landed; the V tile may still be in flight
        __syncthreads();
        float s[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll 4
        for (int ks = 0; ks < 16; ++ks) {
            uint32_t ah[4], al[4], bh[2][4], bl[2][4];
            ldsm_x4(q_s[0] + swz(arow, ks * 2 + achk), ah);
            ldsm_x4(q_s[1] + swz(arow, ks * 2 + achk), al);
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                ldsm_x4(k_s[0] + swz(jj * 16 + brow, ks * 2 + bchk), bh[jj]);
                ldsm_x4(k_s[1] + swz(jj * 16 + brow, ks * 2 + bchk), bl[jj]);
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
                mma_bf16(s[jj * 2], ah, bh[jj][0], bh[jj][1]);
                mma_bf16(s[jj * 2 + 1], ah, bh[jj][2], bh[jj][3]);
                mma_bf16(s[jj * 2], ah, bl[jj][0], bl[jj][1]);
                mma_bf16(s[jj * 2 + 1], ah, bl[jj][2], bl[jj][3]);
                mma_bf16(s[jj * 2], al, bh[jj][0], bh[jj][1]);
                mma_bf16(s[jj * 2 + 1], al, bh[jj][2], bh[jj][3]);
            }
        }
        __syncthreads();             // every warp is done with the K tile: refill it with the next one while softmax / p v run
        if (k0 + 32 < p.Nk) load_kv_tile(k_s, k0 + 32, 0);
        cp_async_commit();
        // online softmax (exp2 domain); key columns of n-tile j held by this thread: k0 + 8j + 2t, +1
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = k0 + j * 8 + 2 * t;
            s[j][0] = c < p.Nk ? s[j][0] * sscale : -INFINITY;
            s[j][1] = c + 1 < p.Nk ? s[j][1] * sscale : -INFINITY;
            s[j][2] = c < p.Nk ? s[j][2] * sscale : -INFINITY;
            s[j][3] = c + 1 < p.Nk ? s[j][3] * sscale : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);       // finite: key k0 is valid
        const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
        m0 = mn0; m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
        uint32_t ph[2][4], pl[2][4];      // A fragments of p (hi / lo) for the two 16-key k-steps
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float e0 = exp2f(s[j][0] - mn0), e1 = exp2f(s[j][1] - mn0), e2 = exp2f(s[j][2] - mn1), e3 = exp2f(s[j][3] - mn1);
            rs0 += e0 + e1; rs1 += e2 + e3;
            // n-tile j covers keys 8j..8j+7: k-step j/2, low (a0,a1) or high (a2,a3) half
            split_bf16x2(e0, e1, ph[j >> 1][(j & 1) * 2], pl[j >> 1][(j & 1) * 2]);
            split_bf16x2(e2, e3, ph[j >> 1][(j & 1) * 2 + 1], pl[j >> 1][(j & 1) * 2 + 1]);
        }
        rs0 += __shfl_xor_sync(0xffffffffu, rs0, 1); rs0 += __shfl_xor_sync(0xffffffffu, rs0, 2);
        rs1 += __shfl_xor_sync(0xffffffffu, rs1, 1); rs1 += __shfl_xor_sync(0xffffffffu, rs1, 2);
        l0 = l0 * c0 + rs0; l1 = l1 * c1 + rs1;
#pragma unroll
        for (int j = 0; j < 32; ++j) { o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1; }
        cp_async_wait<1>();          // V tile of this iteration landed (the K refill just committed may still be in flight)
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
            for (int jn = 0; jn < 16; ++jn) {        // 2 n-tiles (16 dims) per step
                uint32_t vh[4], vl[4];
                ldsm_x4_trans(v_s[0] + swz(kk * 16 + vrow, jn * 2 + vchk), vh);
                ldsm_x4_trans(v_s[1] + swz(kk * 16 + vrow, jn * 2 + vchk), vl);
                mma_bf16(o[jn * 2], ph[kk], vh[0], vh[1]);
                mma_bf16(o[jn * 2 + 1], ph[kk], vh[2], vh[3]);
                mma_bf16(o[jn * 2], ph[kk], vl[0], vl[1]);
                mma_bf16(o[jn * 2 + 1], ph[kk], vl[2], vl[3]);
                mma_bf16(o[jn * 2], pl[kk], vh[0], vh[1]);
                mma_bf16(o[jn * 2 + 1], pl[kk], vh[2], vh[3]);
            }
        }
        __syncthreads();             // V tile consumed
        if (k0 + 32 < p.Nk) load_kv_tile(v_s, k0 + 32, v_off);
        cp_async_commit();
    }
    cp_async_wait<0>();
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int c = h * 256 + j * 8 + 2 * t;
        if (r0 < p.Nq) {
            const int64_t row = (int64_t)b * p.Nq + r0;
            uint32_t hv, lv;
            split_bf16x2(o[j][0] * i0, o[j][1] * i0, hv, lv);
            *reinterpret_cast<uint32_t*>(p.hi + row * p.C_pad + c) = hv;
            *reinterpret_cast<uint32_t*>(p.lo + row * p.C_pad + c) = lv;
            if (p.out32) *reinterpret_cast<float2*>(p.out32 + row * p.out32_ld + c) = make_float2(o[j][0] * i0, o[j][1] * i0);
        }
        if (r1 < p.Nq) {
            const int64_t row = (int64_t)b * p.Nq + r1;
            uint32_t hv, lv;
            split_bf16x2(o[j][2] * i1, o[j][3] * i1, hv, lv);
            *reinterpret_cast<uint32_t*>(p.hi + row * p.C_pad + c) = hv;
            *reinterpret_cast<uint32_t*>(p.lo + row * p.C_pad + c) = lv;
            if (p.out32) *reinterpret_cast<float2*>(p.out32 + row * p.out32_ld + c) = make_float2(o[j][2] * i1, o[j][3] * i1);
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

template <int QW>
int launch_attention_tc(const ia_attention_tc_params* p, cudaStream_t st) {
    const size_t smem = sizeof(AttnTcSmem<QW>);
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<QW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        IA_CHECK(e == cudaSuccess, "ia_attention_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set[dev] = true;
    }
    dim3 grid((unsigned)cdiv(p->Nq, 16 * QW), (unsigned)p->heads, (unsigned)p->B);
    ia::prof_begin("ia_attention_tc", st);
    attention_tc_kernel<QW><<<grid, 32 * QW, smem, st>>>(*p);
    IA_LAUNCH_CHECK("ia_attention_tc");
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

extern "C" int ia_attention_tc(const ia_attention_tc_params* p, void* stream) {
    IA_CHECK(p && p->q_hi && p->q_lo && p->kv_hi && p->kv_lo && p->hi && p->lo, "ia_attention_tc: null operand");
    IA_CHECK(p->B > 0 && p->heads > 0 && p->Nq > 0 && p->Nk > 0, "ia_attention_tc: bad sizes");
    IA_CHECK(p->head_dim == 256, "ia_attention_tc: head_dim %d not instantiated (256: transformer_block); use ia_attention", p->head_dim);
    const int64_t Cc = (int64_t)p->heads * 256;
    IA_CHECK(p->q_ld >= Cc && p->kv_ld >= 2 * Cc && ((p->q_ld | p->kv_ld | p->C_pad) & 7) == 0 && p->C_pad >= Cc,
             "ia_attention_tc: row pitches must cover heads*256 (q, out) / 2*heads*256 (kv) channels and be multiples of 8");
    IA_CHECK((((uintptr_t)p->q_hi | (uintptr_t)p->q_lo | (uintptr_t)p->kv_hi | (uintptr_t)p->kv_lo) & 15) == 0, "ia_attention_tc: operands must be 16-byte aligned");
    IA_CHECK(!p->out32 || (p->out32_ld >= Cc && (p->out32_ld & 1) == 0), "ia_attention_tc: out32_ld");
    // queries per CTA: 128 (8 warps) when that already gives a CTA per SM, else 64 / 32 (4 / 2 warps): the short sequences of the
    // low-resolution UpLayers (64 ... 1024 tokens) then spread over 2 - 4x as many SMs, each warp keeping the tensor pipe of its SM
    // to fewer neighbours
    const int64_t bh = (int64_t)p->B * p->heads;
    int qw = 8;
    if (cdiv(p->Nq, 128) * bh < 120) qw = 4;
    if (qw == 4 && cdiv(p->Nq, 64) * bh < 120) qw = 2;
    { const char* ev = getenv("IA_ATTENTION_QW"); if (ev) { const int v = atoi(ev); if (v == 2 || v == 4 || v == 8) qw = v; } }
    return qw == 8 ? launch_attention_tc<8>(p, as_stream(stream)) : qw == 4 ? launch_attention_tc<4>(p, as_stream(stream)) : launch_attention_tc<2>(p, as_stream(stream));
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
