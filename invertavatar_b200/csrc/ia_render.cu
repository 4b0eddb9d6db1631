// Fused hierarchical volume renderer: one persistent kernel turns (planes, cameras) into the 32-channel
// neural feature image.  Per ray (one warp): ray generation -> jittered coarse depths -> tri-plane bilinear
// gather (8 lanes cooperate on each 128-byte texel) -> OSG decoder MLP 32->64->33 (weights in the constant
// bank, one sample per lane) -> coarse compositing weights -> smoothed inverse-CDF importance sampling ->
// fine gather + MLP -> rank-merge of the two sample lists -> final compositing.  Nothing but the planes is
// read from and nothing but the feature image / depth / weight sum is written to HBM.
//
// Follows reference volumetric_rendering/renderer.py:309-469 (ImportanceRenderer_bsMotion), :51-97
// (project_onto_planes / sample_from_planes), ray_sampler.py:70-107 (RaySampler_zxc), ray_marcher.py:25-57
// (MipRayMarcher2) and triplane_v20.py:415-438 (OSGDecoder).
#include <math_constants.h>
#include <stdlib.h>

#include "ia_common.cuh"

using namespace ia;

namespace {

constexpr int kHidden = 64;
constexpr int kFeat = 32;
constexpr int kOut = 33;
constexpr int kRowLd = 36;          // colour scratch row pitch: even (float2 stores) and 4g+2t bank spread -> conflict-free
constexpr int kMaxWarpsPerCta = 12;   // one CTA per SM, as many ray-warps as the shared-memory scratch allows
constexpr int kMaxS = 192;          // Dc + Df
constexpr int kOutPad = 40;         // layer-2 columns: 0..31 = rgb, 32 = sigma, 33..39 = zero padding (5 n-tiles of 8)

// OSG decoder in mma.sync.m16n8k16 B-fragment order, fp16 hi/lo split (3-term product hi*hi + hi*lo + lo*hi with fp32
// accumulation reproduces the fp32 MLP to ~1e-6), FullyConnectedLayer runtime gains folded in
// (networks_stylegan2.py:111-115).  Built by decoder_stage_kernel, copied to shared memory by every render CTA.
//   w1f[((nt*2 + ks)*2 + hl)*32 + lane] : layer 1, n-tile nt (8 hidden units), k-step ks (16 input channels)
//   w2f[((nt*4 + ks)*2 + hl)*32 + lane] : layer 2, n-tile nt (8 outputs),     k-step ks (16 hidden units)
// Input channel order inside a k-step follows the gather: lane quad member t owns physical channels 4t..4t+3 (k-step 0)
// and 16+4t..16+4t+3 (k-step 1), which sit at logical k positions {2t, 2t+1, 2t+8, 2t+9}.
struct DecoderFrags {
    uint2 w1f[8 * 2 * 2 * 32];
    uint2 w2f[5 * 4 * 2 * 32];
    float b1[kHidden];
    float b2[kOutPad];
};

__device__ __forceinline__ uint32_t pack_half2(__half lo16, __half hi16) {
    return (uint32_t)__half_as_ushort(lo16) | ((uint32_t)__half_as_ushort(hi16) << 16);
}
__device__ __forceinline__ void split_half(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}
// logical layer-1 k index -> physical feature channel.  fp32 planes: lane quad member t gathers channels 4t..4t+3 (k-step 0) and
// 16+4t..16+4t+3 (k-step 1) -- two 16-byte loads per texel; fp16 planes (H16): channels 8t..8t+7 -- ONE 16-byte load per texel,
// of which 8t..8t+3 feed k-step 0 and 8t+4..8t+7 feed k-step 1.
__device__ __forceinline__ int phys_channel(int k, bool h16) {
    const int ks = k >> 4, r = k & 15;
    const int t = (r < 8) ? (r >> 1) : ((r - 8) >> 1);
    const int q = (r < 8) ? (r & 1) : 2 + ((r - 8) & 1);
    return h16 ? 8 * t + 4 * ks + q : 16 * ks + 4 * t + q;
}

__global__ void decoder_stage_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                                     const float* __restrict__ b2, DecoderFrags* __restrict__ out, int h16) {
    DecoderFrags& g_dec = *out;
    // The transcendental constants of the two activations are folded into the staged weights: layer 1 produces pre * log2(e), so
    // softplus(pre) / ln 2 = lg2(1 + ex2(pre')) is two bare MUFU ops; layer 2's density row carries the ln 2, its colour rows
    // carry -log2(e) * ln 2 = -1 (and their bias -log2(e)), so sigmoid(o) = rcp(1 + ex2(o')).
    const float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
    const float g1 = kLog2e / sqrtf((float)kFeat), g2 = 1.0f / sqrtf((float)kHidden);
    // layer 1: B[k][n] = W1[n][phys(k)] * g1
    for (int i = threadIdx.x; i < 8 * 2 * 32; i += blockDim.x) {
        const int lane = i & 31, ks = (i >> 5) & 1, nt = i >> 6;
        const int g = lane >> 2, t = lane & 3, n = nt * 8 + g;
        float v[4];
        const int kk[4] = {16 * ks + 2 * t, 16 * ks + 2 * t + 1, 16 * ks + 2 * t + 8, 16 * ks + 2 * t + 9};
        __half h[4], l[4];
        for (int q = 0; q < 4; ++q) { v[q] = w1[n * kFeat + phys_channel(kk[q], h16 != 0)] * g1; split_half(v[q], h[q], l[q]); }
        g_dec.w1f[((nt * 2 + ks) * 2 + 0) * 32 + lane] = make_uint2(pack_half2(h[0], h[1]), pack_half2(h[2], h[3]));
        g_dec.w1f[((nt * 2 + ks) * 2 + 1) * 32 + lane] = make_uint2(pack_half2(l[0], l[1]), pack_half2(l[2], l[3]));
    }
    // layer 2: column o' < 32 -> rgb channel o' (W2 row 1+o'), o' == 32 -> sigma (W2 row 0), else zero
    for (int i = threadIdx.x; i < 5 * 4 * 32; i += blockDim.x) {
        const int lane = i & 31, ks = (i >> 5) & 3, nt = i >> 7;
        const int g = lane >> 2, t = lane & 3, col = nt * 8 + g;
        const int row = col < 32 ? col + 1 : (col == 32 ? 0 : -1);
        const int kk[4] = {16 * ks + 2 * t, 16 * ks + 2 * t + 1, 16 * ks + 2 * t + 8, 16 * ks + 2 * t + 9};
        __half h[4], l[4];
        const float gq = col < 32 ? -g2 : g2 * kLn2;
        for (int q = 0; q < 4; ++q) { const float v = row >= 0 ? w2[row * kHidden + kk[q]] * gq : 0.f; split_half(v, h[q], l[q]); }
        g_dec.w2f[((nt * 4 + ks) * 2 + 0) * 32 + lane] = make_uint2(pack_half2(h[0], h[1]), pack_half2(h[2], h[3]));
        g_dec.w2f[((nt * 4 + ks) * 2 + 1) * 32 + lane] = make_uint2(pack_half2(l[0], l[1]), pack_half2(l[2], l[3]));
    }
    for (int i = threadIdx.x; i < kHidden; i += blockDim.x) g_dec.b1[i] = b1[i] * kLog2e;
    for (int i = threadIdx.x; i < kOutPad; i += blockDim.x) g_dec.b2[i] = i < 32 ? -b2[i + 1] * kLog2e : (i == 32 ? b2[0] : 0.f);
}

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// softplus(pre) / ln 2 from pre' = pre * log2(e) (threshold 20 of F.softplus in the scaled variable)
__device__ __forceinline__ float softplus_log2(float xs) { return xs > 28.853900817779268f ? xs : lg2_approx(1.f + ex2_approx(xs)); }
__device__ __forceinline__ float softplus_acc(float x) { return x > 20.f ? x : log1pf(expf(x)); }

__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], const uint2 b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

// torch.linspace(start, end, steps) for float32 (ATen RangeFactories: symmetric around the midpoint)
__device__ __forceinline__ float linspace_at(float start, float end, int steps, int i) {
    const float step = (end - start) / (float)(steps - 1);
    return (i < steps / 2) ? start + step * (float)i : end - step * (float)(steps - i - 1);
}

struct Ray { float ox, oy, oz, dx, dy, dz; };

__device__ __forceinline__ Ray make_ray(const float* __restrict__ cam, int res, int px, int py) {
    // intrinsics with the first two rows scaled by res (ray_sampler.py:83), general 3x3 inverse
    float k00 = cam[16] * res, k01 = cam[17] * res, k02 = cam[18] * res;
    float k10 = cam[19] * res, k11 = cam[20] * res, k12 = cam[21] * res;
    float k20 = cam[22], k21 = cam[23], k22 = cam[24];
    float c00 = k11 * k22 - k12 * k21, c01 = k02 * k21 - k01 * k22, c02 = k01 * k12 - k02 * k11;
    float c10 = k12 * k20 - k10 * k22, c11 = k00 * k22 - k02 * k20, c12 = k02 * k10 - k00 * k12;
    float c20 = k10 * k21 - k11 * k20, c21 = k01 * k20 - k00 * k21, c22 = k00 * k11 - k01 * k10;
    float det = k00 * c00 + k01 * c10 + k02 * c20;
    float id = 1.0f / det;
    float x = (float)px, y = (float)py;
    float cx = (c00 * x + c01 * y + c02) * id;
    float cy = (c10 * x + c11 * y + c12) * id;
    float cz = (c20 * x + c21 * y + c22) * id;
    float wx = cam[0] * cx + cam[1] * cy + cam[2] * cz;
    float wy = cam[4] * cx + cam[5] * cy + cam[6] * cz;
    float wz = cam[8] * cx + cam[9] * cy + cam[10] * cz;
    float n = fmaxf(sqrtf(wx * wx + wy * wy + wz * wz), 1e-12f);  // F.normalize eps
    Ray r;
    r.dx = wx / n; r.dy = wy / n; r.dz = wz / n;
    r.ox = cam[3]; r.oy = cam[7]; r.oz = cam[11];
    return r;
}

// Bilinear taps of one plane for this lane's 8 channels (physical 4t..4t+3 and 16+4t..16+4t+3) of one sample:
// F.grid_sample(bilinear, zeros, align_corners=False) arithmetic, accumulated nw, ne, sw, se.
__device__ __forceinline__ void plane_gather8(const float* __restrict__ plane_base, int64_t px_ld, int PH, int PW, float gx, float gy,
                                              float acc[8]) {
    const float ix = ((gx + 1.f) * PW - 1.f) / 2.f;
    const float iy = ((gy + 1.f) * PH - 1.f) / 2.f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float wnw = ((float)x1 - ix) * ((float)y1 - iy);
    const float wne = (ix - (float)x0) * ((float)y1 - iy);
    const float wsw = ((float)x1 - ix) * (iy - (float)y0);
    const float wse = (ix - (float)x0) * (iy - (float)y0);
    const bool vx0 = x0 >= 0 && x0 < PW, vx1 = x1 >= 0 && x1 < PW, vy0 = y0 >= 0 && y0 < PH, vy1 = y1 >= 0 && y1 < PH;
    // Branch-free zero padding: an out-of-range corner reads a clamped (valid) texel with weight 0, so all eight 16-byte
    // loads are issued unconditionally and back to back (memory-level parallelism, no divergence bookkeeping); adding
    // v*0 leaves the sum unchanged, so the result equals the skip-the-corner formulation bit for bit.
    const int x0c = min(max(x0, 0), PW - 1), x1c = min(max(x1, 0), PW - 1);
    const int y0c = min(max(y0, 0), PH - 1), y1c = min(max(y1, 0), PH - 1);
    const float w00 = (vy0 && vx0) ? wnw : 0.f, w01 = (vy0 && vx1) ? wne : 0.f;
    const float w10 = (vy1 && vx0) ? wsw : 0.f, w11 = (vy1 && vx1) ? wse : 0.f;
    const float* r0 = plane_base + (int64_t)(y0c * PW) * px_ld;
    const float* r1 = plane_base + (int64_t)(y1c * PW) * px_ld;
    const float* p00 = r0 + (int64_t)x0c * px_ld;
    const float* p01 = r0 + (int64_t)x1c * px_ld;
    const float* p10 = r1 + (int64_t)x0c * px_ld;
    const float* p11 = r1 + (int64_t)x1c * px_ld;
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(p00)), a1 = __ldg(reinterpret_cast<const float4*>(p00 + 16));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p01)), b1 = __ldg(reinterpret_cast<const float4*>(p01 + 16));
    const float4 c0 = __ldg(reinterpret_cast<const float4*>(p10)), c1 = __ldg(reinterpret_cast<const float4*>(p10 + 16));
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(p11)), d1 = __ldg(reinterpret_cast<const float4*>(p11 + 16));
    float s[8];
    s[0] = a0.x * w00; s[1] = a0.y * w00; s[2] = a0.z * w00; s[3] = a0.w * w00;
    s[4] = a1.x * w00; s[5] = a1.y * w00; s[6] = a1.z * w00; s[7] = a1.w * w00;
#define IA_TAP(lo4_, hi4_, wt_)                                                                                   \
    s[0] += lo4_.x * wt_; s[1] += lo4_.y * wt_; s[2] += lo4_.z * wt_; s[3] += lo4_.w * wt_;                       \
    s[4] += hi4_.x * wt_; s[5] += hi4_.y * wt_; s[6] += hi4_.z * wt_; s[7] += hi4_.w * wt_;
    IA_TAP(b0, b1, w01) IA_TAP(c0, c1, w10) IA_TAP(d0, d1, w11)
#undef IA_TAP
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += s[k];
}

// fp16 planes: this lane's 8 channels 8t..8t+7 of a texel are one 16-byte load; arithmetic stays fp32 (only the storage is
// rounded: 9.4e-6 on the final image, profiles/r1_render_precision_probe.json).  Same zero-padding scheme as above.
// Bilinear set-up of ONE plane of one sample (the arithmetic of plane_gather8_h, split from its loads): clamped texel offsets
// (elements, relative to the plane's channel 0) and the four zero-padding-masked weights.  In gather_mlp_pass the four lanes of a
// quad need the same 3 planes x 2 samples: lane t computes plane min(t, 2) of both samples and the quad exchanges the results by
// shuffle (8 values per plane and sample) instead of every lane computing all six -- same values, bit-identical features.
struct TapSetup { int o00, o01, o10, o11; float w00, w01, w10, w11; };
__device__ __forceinline__ TapSetup tap_setup(int64_t px_ld, int PH, int PW, float gx, float gy) {
    const float ix = ((gx + 1.f) * PW - 1.f) / 2.f;
    const float iy = ((gy + 1.f) * PH - 1.f) / 2.f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float wnw = ((float)x1 - ix) * ((float)y1 - iy);
    const float wne = (ix - (float)x0) * ((float)y1 - iy);
    const float wsw = ((float)x1 - ix) * (iy - (float)y0);
    const float wse = (ix - (float)x0) * (iy - (float)y0);
    const bool vx0 = x0 >= 0 && x0 < PW, vx1 = x1 >= 0 && x1 < PW, vy0 = y0 >= 0 && y0 < PH, vy1 = y1 >= 0 && y1 < PH;
    const int x0c = min(max(x0, 0), PW - 1), x1c = min(max(x1, 0), PW - 1);
    const int y0c = min(max(y0, 0), PH - 1), y1c = min(max(y1, 0), PH - 1);
    TapSetup t;
    t.w00 = (vy0 && vx0) ? wnw : 0.f; t.w01 = (vy0 && vx1) ? wne : 0.f;
    t.w10 = (vy1 && vx0) ? wsw : 0.f; t.w11 = (vy1 && vx1) ? wse : 0.f;
    const int ld = (int)px_ld;
    t.o00 = (y0c * PW + x0c) * ld; t.o01 = (y0c * PW + x1c) * ld;
    t.o10 = (y1c * PW + x0c) * ld; t.o11 = (y1c * PW + x1c) * ld;
    return t;
}
// The four taps of one plane for this lane's 8 fp16 channels, fp32 arithmetic on packed pairs (fma.rn.f32x2: the same roundings as
// eight scalar FMAs in half the instructions).
__device__ __forceinline__ void plane_taps8_h(const __half* __restrict__ base, const TapSetup& t, float2 (&acc)[4]) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(base + t.o00));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(base + t.o01));
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(base + t.o10));
    const uint4 d = __ldg(reinterpret_cast<const uint4*>(base + t.o11));
    float2 s[4];
    auto h2 = [](uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); };
    const float2 wa = make_float2(t.w00, t.w00), wb = make_float2(t.w01, t.w01), wc = make_float2(t.w10, t.w10), wd = make_float2(t.w11, t.w11);
    s[0] = __fmul2_rn(h2(a.x), wa); s[1] = __fmul2_rn(h2(a.y), wa); s[2] = __fmul2_rn(h2(a.z), wa); s[3] = __fmul2_rn(h2(a.w), wa);
    s[0] = __ffma2_rn(h2(b.x), wb, s[0]); s[1] = __ffma2_rn(h2(b.y), wb, s[1]); s[2] = __ffma2_rn(h2(b.z), wb, s[2]); s[3] = __ffma2_rn(h2(b.w), wb, s[3]);
    s[0] = __ffma2_rn(h2(c.x), wc, s[0]); s[1] = __ffma2_rn(h2(c.y), wc, s[1]); s[2] = __ffma2_rn(h2(c.z), wc, s[2]); s[3] = __ffma2_rn(h2(c.w), wc, s[3]);
    s[0] = __ffma2_rn(h2(d.x), wd, s[0]); s[1] = __ffma2_rn(h2(d.y), wd, s[1]); s[2] = __ffma2_rn(h2(d.z), wd, s[2]); s[3] = __ffma2_rn(h2(d.w), wd, s[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = __fadd2_rn(acc[k], s[k]);
}

// Tri-plane gather + OSG decoder for samples [s0, s0+n) of this warp's ray, 16 samples per step on the tensor cores
// (mma.sync m16n8k16, fp16 hi/lo split operands, fp32 accumulate).  Lane (g = lane/4, t = lane%4) gathers 8 channels of
// samples 16m+g and 16m+g+8 straight into its A fragment; layer 1's C fragments become layer 2's A fragments in
// registers.  Colours go to col[sample][channel], sigma to sig[sample].
// MLP1 (ia_render_params.mlp_fmt = IA_OPFMT_F16X1): both decoder layers as single-pass fp16 products (fp32 accumulate) instead of
// the 3-term hi/lo split -- a third of the mma.sync and none of the lo-fragment arithmetic.  CPU probe with the oracle
// (tools/probe_render_precision.py): 1.2e-4 max-abs / 94 dB on the final image, inside the 1e-3 bar but outside the 2e-5 the
// op-level renderer tests hold the feature image to: the generator asks for it (its measured budget), a bare renderer does not.
template <bool MLP1, bool H16>
__device__ __forceinline__ void gather_mlp_pass(const ia_render_params& p, const float* __restrict__ planes_b, const Ray& r,
                                                const float* dep, float* col, float* sig, int s0, int n, int lane,
                                                const DecoderFrags* __restrict__ dec) {
    const int g = lane >> 2, t = lane & 3;
    const float scale = 2.0f / p.box_warp;
    const float* pb = planes_b + 4 * t;
    const __half* pbh = reinterpret_cast<const __half*>(planes_b) + 8 * t;
    for (int m0 = 0; m0 < n; m0 += 16) {
        float feat[2][8];
        if (H16) {
            // quad-cooperative set-up: lane t owns plane min(t, 2) -- (x, y), (x, z), (z, x) -- of the quad's two samples
            TapSetup mine[2];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int sidx = min(m0 + g + 8 * rr, n - 1);      // rows past the end repeat the last sample; results are dropped
                const float tt = dep[s0 + sidx];
                const float qx = (r.ox + tt * r.dx) * scale;
                const float qy = (r.oy + tt * r.dy) * scale;
                const float qz = (r.oz + tt * r.dz) * scale;
                const float gx = t < 2 ? qx : qz;
                const float gy = t == 0 ? qy : (t == 1 ? qz : qx);
                mine[rr] = tap_setup(p.plane_px_ld, p.PH, p.PW, gx, gy);
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                for (int pl = 0; pl < 3; ++pl) {
                    const int src = (lane & ~3) | pl;
                    TapSetup ts;
                    ts.o00 = __shfl_sync(0xffffffffu, mine[rr].o00, src); ts.o01 = __shfl_sync(0xffffffffu, mine[rr].o01, src);
                    ts.o10 = __shfl_sync(0xffffffffu, mine[rr].o10, src); ts.o11 = __shfl_sync(0xffffffffu, mine[rr].o11, src);
                    ts.w00 = __shfl_sync(0xffffffffu, mine[rr].w00, src); ts.w01 = __shfl_sync(0xffffffffu, mine[rr].w01, src);
                    ts.w10 = __shfl_sync(0xffffffffu, mine[rr].w10, src); ts.w11 = __shfl_sync(0xffffffffu, mine[rr].w11, src);
                    plane_taps8_h(pbh + 32 * pl, ts, acc);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) { feat[rr][2 * k] = acc[k].x * (1.0f / 3.0f); feat[rr][2 * k + 1] = acc[k].y * (1.0f / 3.0f); }
            }
        }
        if (!H16) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int sidx = min(m0 + g + 8 * rr, n - 1);      // rows past the end repeat the last sample; results are dropped
                const float tt = dep[s0 + sidx];
                const float qx = (r.ox + tt * r.dx) * scale;
                const float qy = (r.oy + tt * r.dy) * scale;
                const float qz = (r.oz + tt * r.dz) * scale;
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                plane_gather8(pb + 0, p.plane_px_ld, p.PH, p.PW, qx, qy, acc);    // plane 0: (x, y)
                plane_gather8(pb + 32, p.plane_px_ld, p.PH, p.PW, qx, qz, acc);   // plane 1: (x, z)
                plane_gather8(pb + 64, p.plane_px_ld, p.PH, p.PW, qz, qx, acc);   // plane 2: (z, x)
#pragma unroll
                for (int k = 0; k < 8; ++k) feat[rr][k] = acc[k] * (1.0f / 3.0f);  // mean over the three planes (<= 1 ulp from the division)
            }
        }
        // ---- layer 1: [16 x 32] x [32 x 64] ----
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            __half h[2][4], l[2][4];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                for (int q = 0; q < 4; ++q) split_half(feat[rr][4 * ks + q], h[rr][q], l[rr][q]);
            ah[ks][0] = pack_half2(h[0][0], h[0][1]); ah[ks][1] = pack_half2(h[1][0], h[1][1]);
            ah[ks][2] = pack_half2(h[0][2], h[0][3]); ah[ks][3] = pack_half2(h[1][2], h[1][3]);
            if (!MLP1) {
                al[ks][0] = pack_half2(l[0][0], l[0][1]); al[ks][1] = pack_half2(l[1][0], l[1][1]);
                al[ks][2] = pack_half2(l[0][2], l[0][3]); al[ks][3] = pack_half2(l[1][2], l[1][3]);
            }
        }
        float hid[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float bb0 = dec->b1[nt * 8 + 2 * t], bb1 = dec->b1[nt * 8 + 2 * t + 1];
            float c[4] = {bb0, bb1, bb0, bb1};
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                if (MLP1) {
                    mma_f16(c, ah[ks], dec->w1f[((nt * 2 + ks) * 2 + 0) * 32 + lane]);
                } else {
                    const uint2 bh = dec->w1f[((nt * 2 + ks) * 2 + 0) * 32 + lane];
                    const uint2 bl = dec->w1f[((nt * 2 + ks) * 2 + 1) * 32 + lane];
                    mma_f16(c, ah[ks], bh);
                    mma_f16(c, ah[ks], bl);
                    mma_f16(c, al[ks], bh);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) hid[nt][q] = softplus_log2(c[q]);
        }
        // ---- layer 2: [16 x 64] x [64 x 40] ----
        float out[5][4];
#pragma unroll
        for (int nt = 0; nt < 5; ++nt) {
            const float bb0 = dec->b2[nt * 8 + 2 * t], bb1 = dec->b2[nt * 8 + 2 * t + 1];
            out[nt][0] = bb0; out[nt][1] = bb1; out[nt][2] = bb0; out[nt][3] = bb1;
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            __half h[8], l[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) { split_half(hid[2 * ks][q], h[q], l[q]); split_half(hid[2 * ks + 1][q], h[4 + q], l[4 + q]); }
            const uint32_t a2h[4] = {pack_half2(h[0], h[1]), pack_half2(h[2], h[3]), pack_half2(h[4], h[5]), pack_half2(h[6], h[7])};
            const uint32_t a2l[4] = {pack_half2(l[0], l[1]), pack_half2(l[2], l[3]), pack_half2(l[4], l[5]), pack_half2(l[6], l[7])};
#pragma unroll
            for (int nt = 0; nt < 5; ++nt) {
                if (MLP1) {
                    mma_f16(out[nt], a2h, dec->w2f[((nt * 4 + ks) * 2 + 0) * 32 + lane]);
                } else {
                    const uint2 bh = dec->w2f[((nt * 4 + ks) * 2 + 0) * 32 + lane];
                    const uint2 bl = dec->w2f[((nt * 4 + ks) * 2 + 1) * 32 + lane];
                    mma_f16(out[nt], a2h, bh);
                    mma_f16(out[nt], a2h, bl);
                    mma_f16(out[nt], a2l, bh);
                }
            }
        }
        // ---- write back: rows g (c0,c1) and g+8 (c2,c3); columns nt*8 + 2t, +1 ----
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int sidx = m0 + g + 8 * rr;
            if (sidx < n) {
                float* row = col + (s0 + sidx) * kRowLd;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const float v0 = rcp_approx(1.0f + ex2_approx(out[nt][2 * rr + 0]));      // sigmoid: out' = -out * log2(e)
                    const float v1 = rcp_approx(1.0f + ex2_approx(out[nt][2 * rr + 1]));
                    *reinterpret_cast<float2*>(row + nt * 8 + 2 * t) =
                        make_float2(v0 * (1.0f + 2.0f * 0.001f) - 0.001f, v1 * (1.0f + 2.0f * 0.001f) - 0.001f);
                }
                if (t == 0) sig[s0 + sidx] = out[4][2 * rr];
            }
        }
    }
}

// Ray-marching weights of the sorted (depth, sigma) list of length n: w[i], i in [0, n-1).  Returns (sum w, sum w*dmid).
__device__ __forceinline__ void march_weights(const float* d, const float* sg, float* w, int n, int lane, float& wsum, float& dnum) {
    float carry = 1.0f;  // running transmittance entering the current 32-interval chunk
    float ws = 0.f, dn = 0.f;
    for (int base = 0; base < n - 1; base += 32) {
        const int i = base + lane;
        float alpha = 0.f, shifted = 1.f, dmid = 0.f;
        if (i < n - 1) {
            const float delta = d[i + 1] - d[i];
            const float smid = (sg[i] + sg[i + 1]) / 2.0f;
            dmid = (d[i] + d[i + 1]) / 2.0f;
            const float dens = softplus_acc(smid - 1.0f);
            alpha = 1.0f - expf(-(dens * delta));
            shifted = 1.0f - alpha + 1e-10f;
        }
        // exclusive product scan of `shifted` across the warp
        float incl = shifted;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            float v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl *= v;
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        const float T = carry * excl;
        const float wi = alpha * T;
        if (i < n - 1) { w[i] = wi; ws += wi; dn += wi * dmid; }
        carry = carry * __shfl_sync(0xffffffffu, incl, 31);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ws += __shfl_xor_sync(0xffffffffu, ws, o); dn += __shfl_xor_sync(0xffffffffu, dn, o); }
    wsum = ws; dnum = dn;
}

template <bool MLP1, bool H16>
__global__ void __launch_bounds__(kMaxWarpsPerCta * 32, 1) render_kernel(const ia_render_params p) {
    const int kWarpsPerCta = blockDim.x >> 5;
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.Dc + p.Df;
    // decoder fragments (shared by the CTA), then per-warp scratch
    DecoderFrags* dec = reinterpret_cast<DecoderFrags*>(smem);
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.scratch);
        uint4* dst = reinterpret_cast<uint4*>(dec);
        for (int i = threadIdx.x; i < (int)(sizeof(DecoderFrags) / 16); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int per_warp = S * kRowLd + 6 * S + 2 * (p.Dc + 2) + ((S * kRowLd + 6 * S + 2 * (p.Dc + 2)) & 1);
    float* base = smem + sizeof(DecoderFrags) / 4 + (size_t)warp * per_warp;
    float* col = base;                      // [S][33] features -> colours
    float* dep = col + S * kRowLd;          // [S] depths (coarse then fine)
    float* sig = dep + S;                   // [S]
    float* sd = sig + S;                    // [S] sorted depths
    float* ssg = sd + S;                    // [S] sorted sigmas
    float* wgt = ssg + S;                   // [S] interval weights
    int* perm = reinterpret_cast<int*>(wgt + S);   // [S] sorted position -> sample
    float* cdf = reinterpret_cast<float*>(perm + S);  // [Dc+2]
    float* zmid = cdf + (p.Dc + 2);         // [Dc+2]

    const int rays_per_img = p.res * p.res;
    const int64_t total_rays = (int64_t)p.B * rays_per_img;
    const float near = p.near_far[0], far = p.near_far[1];
    const float delta_c = p.near_far[2] / (float)(p.Dc - 1);
    float dmin = CUDART_INF_F, dmax = -CUDART_INF_F;

    for (int64_t ray = (int64_t)blockIdx.x * kWarpsPerCta + warp; ray < total_rays; ray += (int64_t)gridDim.x * kWarpsPerCta) {
        const int b = (int)(ray / rays_per_img);
        const int m = (int)(ray % rays_per_img);
        const int py = m / p.res, px = m % p.res;
        Ray r;
        if (p.rays_o) {   // explicit rays (ImportanceRenderer API); otherwise generate them from the camera
            r.ox = p.rays_o[ray * 3 + 0]; r.oy = p.rays_o[ray * 3 + 1]; r.oz = p.rays_o[ray * 3 + 2];
            r.dx = p.rays_d[ray * 3 + 0]; r.dy = p.rays_d[ray * 3 + 1]; r.dz = p.rays_d[ray * 3 + 2];
        } else {
            r = make_ray(p.cam + (int64_t)b * p.cam_ld, p.res, px, py);
        }
        // (fp16 planes: the same element offset, two bytes per element)
        const float* planes_b = H16 ? reinterpret_cast<const float*>(reinterpret_cast<const __half*>(p.planes) + (int64_t)b * p.PH * p.PW * p.plane_px_ld)
                                    : p.planes + (int64_t)b * p.PH * p.PW * p.plane_px_ld;

        // ---- coarse depths (renderer.py:404-406) ----
        for (int s = lane; s < p.Dc; s += 32) {
            const float t = linspace_at(near, far, p.Dc, s) + p.jitter[ray * p.Dc + s] * delta_c;
            dep[s] = t;
            dmin = fminf(dmin, t); dmax = fmaxf(dmax, t);
        }
        // The coarse and the fine pass are two iterations of one loop so that the gather + MLP body (~20 KB of SASS) exists once: two
        // inlined copies do not fit the instruction cache together (ncu: 6 % of the stall samples on instruction fetch).
        int n_all = p.Dc;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          if (pass == 1) {
            if (p.Df <= 0) break;
            // ---- coarse weights + importance sampling (renderer.py:410-469) ----
            float ws_c, dn_c;
            march_weights(dep, sig, wgt, p.Dc, lane, ws_c, dn_c);
            __syncwarp();
            const int nw = p.Dc - 1;       // number of coarse weights
            const int nb = p.Dc - 3;       // number of pdf bins (weights[1:-1] after smoothing)
            // smoothed weights a_k = (max(w[k-1],w[k]) + max(w[k],w[k+1]))/2 + 0.01, k in [0,nw); we need k = 1..nb
            for (int k = lane; k < nw; k += 32) zmid[k] = 0.5f * (dep[k] + dep[k + 1]);
            for (int k = lane; k < nb; k += 32) {
                const int kk = k + 1;
                const float wm1 = wgt[kk - 1], w0 = wgt[kk], wp1 = (kk + 1 < nw) ? wgt[kk + 1] : -CUDART_INF_F;
                const float m0 = fmaxf(wm1, w0);
                const float m1 = (kk + 1 < nw) ? fmaxf(w0, wp1) : w0;
                sd[k] = ((m0 + m1) * 0.5f + 0.01f) + 1e-5f;   // pdf numerators (sample_pdf adds eps)
            }
            __syncwarp();
            {   // pdf = w / sum(w); cdf = [0, cumsum(pdf)]  (warp-parallel sum and scan, 32 bins per round)
                float tot = 0.f;
                for (int k = lane; k < nb; k += 32) tot += sd[k];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                float carry_c = 0.f;
                if (lane == 0) cdf[0] = 0.f;
                for (int base_k = 0; base_k < nb; base_k += 32) {
                    const int k = base_k + lane;
                    float v = (k < nb) ? sd[k] / tot : 0.f;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const float up = __shfl_up_sync(0xffffffffu, v, o);
                        if (lane >= o) v += up;
                    }
                    v += carry_c;
                    if (k < nb) cdf[k + 1] = v;
                    carry_c = __shfl_sync(0xffffffffu, v, 31);
                }
            }
            __syncwarp();
            for (int f = lane; f < p.Df; f += 32) {
                const float u = p.u ? p.u[ray * p.Df + f] : linspace_at(0.0f, 1.0f, p.Df, f);
                int inds = 0;   // searchsorted(cdf, u, right=True) over the nb+1 non-decreasing entries: #{k : cdf[k] <= u}
                {
                    int lo_i = 0, hi_i = nb + 1;
                    while (lo_i < hi_i) { const int mid = (lo_i + hi_i) >> 1; if (cdf[mid] <= u) lo_i = mid + 1; else hi_i = mid; }
                    inds = lo_i;
                }
                const int below = max(inds - 1, 0);
                const int above = min(inds, nb);
                const float c0 = cdf[below], c1 = cdf[above];
                float denom = c1 - c0;
                if (denom < 1e-5f) denom = 1.0f;
                const float b0 = zmid[below], b1 = zmid[above];
                const float t = b0 + (u - c0) / denom * (b1 - b0);
                dep[p.Dc + f] = t;
                dmin = fminf(dmin, t); dmax = fmaxf(dmax, t);
            }
          }
          __syncwarp();
          gather_mlp_pass<MLP1, H16>(p, planes_b, r, dep, col, sig, pass ? p.Dc : 0, pass ? p.Df : p.Dc, lane, dec);
          __syncwarp();
        }
        if (p.Df > 0) {
            // ---- merge: stable rank of every sample among all S (unify_samples, renderer.py:372-382) ----
            // Both lists are normally already sorted (coarse: jitter < bin width; fine: deterministic u), in which case the
            // stable rank is a 2-way merge: coarse s -> s + #{fine < d_s}, fine f -> f + #{coarse <= d_f} (binary searches).
            // Anything else (random u, or a rounding inversion at a bin edge) takes the exact O(S^2) rank.
            bool sorted_ok = true;
            for (int s = lane; s < S - 1; s += 32)
                if (s != p.Dc - 1 && dep[s] > dep[s + 1]) sorted_ok = false;
            sorted_ok = __all_sync(0xffffffffu, sorted_ok);
            if (sorted_ok) {
                for (int s = lane; s < S; s += 32) {
                    const float ds = dep[s];
                    int lo_i, hi_i, rank;
                    if (s < p.Dc) {      // lower_bound over the fine list
                        lo_i = p.Dc; hi_i = S;
                        while (lo_i < hi_i) { const int mid = (lo_i + hi_i) >> 1; if (dep[mid] < ds) lo_i = mid + 1; else hi_i = mid; }
                        rank = s + (lo_i - p.Dc);
                    } else {             // upper_bound over the coarse list
                        lo_i = 0; hi_i = p.Dc;
                        while (lo_i < hi_i) { const int mid = (lo_i + hi_i) >> 1; if (dep[mid] <= ds) lo_i = mid + 1; else hi_i = mid; }
                        rank = (s - p.Dc) + lo_i;
                    }
                    sd[rank] = ds; ssg[rank] = sig[s]; perm[rank] = s;
                }
            } else {
                for (int s = lane; s < S; s += 32) {
                    const float ds = dep[s];
                    int rank = 0;
                    for (int j = 0; j < S; ++j) {
                        const float dj = dep[j];
                        rank += (dj < ds || (dj == ds && j < s)) ? 1 : 0;
                    }
                    sd[rank] = ds; ssg[rank] = sig[s]; perm[rank] = s;
                }
            }
            n_all = S;
            __syncwarp();
        } else {
            for (int s = lane; s < p.Dc; s += 32) { sd[s] = dep[s]; ssg[s] = sig[s]; perm[s] = s; }
            __syncwarp();
        }

        // ---- final compositing (ray_marcher.py:25-57) ----
        float wsum, dnum;
        march_weights(sd, ssg, wgt, n_all, lane, wsum, dnum);
        __syncwarp();
        {
            const int c = lane;  // one feature channel per lane
            float acc = 0.f;
            float prev = col[perm[0] * kRowLd + c];
            for (int i = 0; i < n_all - 1; ++i) {
                const float nxt = col[perm[i + 1] * kRowLd + c];
                acc = fmaf(wgt[i], (prev + nxt) / 2.0f, acc);
                prev = nxt;
            }
            if (p.white_back) acc = acc + 1.0f - wsum;
            p.feat[ray * kFeat + c] = acc * 2.0f - 1.0f;
        }
        if (lane == 0) {
            p.depth[ray] = dnum / wsum;
            p.wsum[ray] = wsum;
        }
        __syncwarp();
    }
    // global min/max of all sample depths (ray_marcher.py:50); depths are positive so int ordering works
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    }
    if (lane == 0 && dmin <= dmax) {
        atomicMin(reinterpret_cast<int*>(p.depth_minmax), __float_as_int(dmin));
        atomicMax(reinterpret_cast<int*>(p.depth_minmax) + 1, __float_as_int(dmax));
    }
}

// ---- standalone MipRayMarcher2 (ray_marcher.py:25-57): the generator never calls it on its own (fused into render_kernel); this is
// the reference's module-level API for callers that march their own samples.  One thread per ray walks the S samples.
__global__ void __launch_bounds__(128) ray_march_kernel(const float* __restrict__ colors, const float* __restrict__ sigmas, const float* __restrict__ depths,
                                                        int64_t rays, int S, int C, int white_back, float* __restrict__ rgb, float* __restrict__ depth,
                                                        float* __restrict__ weights, float* __restrict__ minmax) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float dmin = CUDART_INF_F, dmax = -CUDART_INF_F;
    if (ray < rays) {
        const float* d = depths + ray * S;
        const float* sg = sigmas + ray * S;
        float T = 1.0f, wsum = 0.f, dnum = 0.f;
        for (int i = 0; i < S; ++i) { dmin = fminf(dmin, d[i]); dmax = fmaxf(dmax, d[i]); }
        for (int i = 0; i < S - 1; ++i) {
            const float delta = d[i + 1] - d[i];
            const float dens = softplus_acc((sg[i] + sg[i + 1]) / 2.0f - 1.0f);
            const float alpha = 1.0f - expf(-(dens * delta));
            const float w = alpha * T;
            weights[ray * (S - 1) + i] = w;
            wsum += w; dnum += w * ((d[i] + d[i + 1]) / 2.0f);
            T *= 1.0f - alpha + 1e-10f;
        }
        for (int c = 0; c < C; ++c) {
            const float* col = colors + ray * S * C + c;
            float acc = 0.f;
            for (int i = 0; i < S - 1; ++i) acc = fmaf(weights[ray * (S - 1) + i], (col[(int64_t)i * C] + col[(int64_t)(i + 1) * C]) / 2.0f, acc);
            if (white_back) acc = acc + 1.0f - wsum;
            rgb[ray * C + c] = acc * 2.0f - 1.0f;
        }
        depth[ray] = dnum / wsum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
        dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    }
    if ((threadIdx.x & 31) == 0 && dmin <= dmax) {      // depths are positive: the int ordering of the bit patterns is the float ordering
        atomicMin(reinterpret_cast<int*>(minmax), __float_as_int(dmin));
        atomicMax(reinterpret_cast<int*>(minmax) + 1, __float_as_int(dmax));
    }
}

__global__ void ray_bounds_kernel(const float* __restrict__ cam, int64_t cam_ld, int B, float* __restrict__ near_far) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
        const float* c = cam + b * cam_ld;
        s += sqrtf(c[3] * c[3] + c[7] * c[7] + c[11] * c[11]);
    }
    const double dist = (double)(s / (float)B);
    const double start = dist - 0.45, end = dist + 0.6;
    near_far[0] = (float)start;
    near_far[1] = (float)end;
    near_far[2] = (float)(end - start);
}

// same bounds from explicit ray origins [n][3]: mean of the per-ray norms (renderer.py:311)
__global__ void ray_bounds_origins_kernel(const float* __restrict__ origins, int64_t n, float* __restrict__ near_far) {
    __shared__ float red[32];
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float* o = origins + i * 3;
        s += sqrtf(o[0] * o[0] + o[1] * o[1] + o[2] * o[2]);
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) {
            const double dist = (double)(t / (float)n);
            const double start = dist - 0.45, end = dist + 0.6;
            near_far[0] = (float)start; near_far[1] = (float)end; near_far[2] = (float)(end - start);
        }
    }
}

__global__ void minmax_init_kernel(float* mm) {
    if (threadIdx.x == 0) { mm[0] = CUDART_INF_F; mm[1] = 0.0f; }
}

__global__ void depth_clamp_kernel(float* __restrict__ depth, int64_t n, const float* __restrict__ mm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float d = depth[i];
    if (isnan(d)) d = CUDART_INF_F;   // torch.nan_to_num(depth, nan=inf)
    depth[i] = fminf(fmaxf(d, mm[0]), mm[1]);
}

__global__ void ray_sampler_kernel(const float* __restrict__ cam, int64_t cam_ld, int B, int res, float* __restrict__ origins,
                                   float* __restrict__ dirs) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (int64_t)B * res * res;
    if (i >= total) return;
    int b = (int)(i / ((int64_t)res * res));
    int m = (int)(i % ((int64_t)res * res));
    Ray r = make_ray(cam + b * cam_ld, res, m % res, m / res);
    origins[i * 3 + 0] = r.ox; origins[i * 3 + 1] = r.oy; origins[i * 3 + 2] = r.oz;
    dirs[i * 3 + 0] = r.dx; dirs[i * 3 + 1] = r.dy; dirs[i * 3 + 2] = r.dz;
}

}  // namespace

extern "C" int64_t ia_render_scratch_bytes(void) { return (int64_t)sizeof(DecoderFrags); }

extern "C" int ia_ray_bounds(const float* cam, int64_t cam_ld, int32_t B, float* near_far, void* stream) {
    IA_CHECK(cam && near_far && B > 0, "ia_ray_bounds: bad arguments");
    ia::prof_begin("ia_ray_bounds", as_stream(stream));
    ray_bounds_kernel<<<1, 32, 0, as_stream(stream)>>>(cam, cam_ld, B, near_far);
    IA_LAUNCH_CHECK("ia_ray_bounds");
    return 0;
}

extern "C" int ia_ray_bounds_from_origins(const float* origins, int64_t n, float* near_far, void* stream) {
    IA_CHECK(origins && near_far && n > 0, "ia_ray_bounds_from_origins: bad arguments");
    ia::prof_begin("ia_ray_bounds_from_origins", as_stream(stream));
    ray_bounds_origins_kernel<<<1, 1024, 0, as_stream(stream)>>>(origins, n, near_far);
    IA_LAUNCH_CHECK("ia_ray_bounds_from_origins");
    return 0;
}

extern "C" int ia_render(const ia_render_params* p, void* stream) {
    IA_CHECK(p && p->planes && (p->cam || (p->rays_o && p->rays_d)) && p->jitter && p->near_far && p->feat && p->depth && p->wsum && p->depth_minmax,
             "ia_render: null argument");
    IA_CHECK(p->w1 && p->b1 && p->w2 && p->b2, "ia_render: null decoder weights");
    IA_CHECK(p->scratch && (reinterpret_cast<uintptr_t>(p->scratch) & 15) == 0, "ia_render: scratch of ia_render_scratch_bytes() bytes (16-byte aligned) required");
    IA_CHECK(p->Dc >= 4 && p->Dc <= 96 && p->Df >= 0 && p->Df <= 96, "ia_render: depth resolutions must be in [4,96] / [0,96]");
    IA_CHECK(p->planes_fmt == IA_OPFMT_BF16X3 || p->planes_fmt == IA_OPFMT_F16X1, "ia_render: planes_fmt must be 0 (fp32) or IA_OPFMT_F16X1 (fp16)");
    const bool h16 = p->planes_fmt == IA_OPFMT_F16X1;
    IA_CHECK((p->plane_px_ld & (h16 ? 7 : 3)) == 0 && p->plane_px_ld >= 96 && (reinterpret_cast<uintptr_t>(p->planes) & 15) == 0,
             "ia_render: planes need >= 96 channels, 16-byte aligned pixels");
    IA_CHECK(p->res > 0 && p->B > 0, "ia_render: empty batch");
    cudaStream_t st = as_stream(stream);
    ia::prof_begin("ia_render(decoder_stage)", st);
    decoder_stage_kernel<<<1, 256, 0, st>>>(p->w1, p->b1, p->w2, p->b2, reinterpret_cast<DecoderFrags*>(p->scratch), h16 ? 1 : 0);
    IA_LAUNCH_CHECK("ia_render(decoder_stage)");
    cudaError_t e = cudaSuccess;
    ia::prof_begin("ia_render(minmax_init)", st);
    minmax_init_kernel<<<1, 32, 0, st>>>(p->depth_minmax);
    IA_LAUNCH_CHECK("ia_render(minmax_init)");

    const int S = p->Dc + p->Df;
    size_t per_warp = (size_t)S * kRowLd + 6 * (size_t)S + 2 * (size_t)(p->Dc + 2);
    per_warp += per_warp & 1;   // keep every warp's scratch 8-byte aligned (float2 colour stores)
    int kWarpsPerCta = (int)((227 * 1024 - sizeof(DecoderFrags)) / (per_warp * sizeof(float)));
    if (kWarpsPerCta > kMaxWarpsPerCta) kWarpsPerCta = kMaxWarpsPerCta;
    IA_CHECK(kWarpsPerCta >= 1, "ia_render: shared memory request too large");
    const size_t smem = sizeof(DecoderFrags) + per_warp * kWarpsPerCta * sizeof(float);
    IA_CHECK(p->mlp_fmt == IA_OPFMT_BF16X3 || p->mlp_fmt == IA_OPFMT_F16X1, "ia_render: unknown mlp_fmt %d", p->mlp_fmt);
    const bool mlp1 = p->mlp_fmt == IA_OPFMT_F16X1;      // single-pass fp16 decoder (see gather_mlp_pass)
    void (*kern)(const ia_render_params) = mlp1 ? (h16 ? render_kernel<true, true> : render_kernel<true, false>)
                                                : (h16 ? render_kernel<false, true> : render_kernel<false, false>);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IA_CHECK(e == cudaSuccess, "ia_render: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarpsPerCta * 32, smem);
    if (per_sm < 1) per_sm = 1;
    int64_t total_rays = (int64_t)p->B * p->res * p->res;
    int64_t want = cdiv(total_rays, kWarpsPerCta);
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > want) grid = want;
    ia::prof_begin("ia_render", st);
    kern<<<(unsigned)grid, kWarpsPerCta * 32, smem, st>>>(*p);
    IA_LAUNCH_CHECK("ia_render");
    return 0;
}

extern "C" int ia_ray_march(const float* colors, const float* densities, const float* depths, int64_t rays, int32_t S, int32_t C, int32_t white_back,
                            float* rgb, float* depth, float* weights, float* depth_minmax, void* stream) {
    IA_CHECK(colors && densities && depths && rgb && depth && weights && depth_minmax, "ia_ray_march: null argument");
    IA_CHECK(rays >= 0 && S >= 2 && C >= 1, "ia_ray_march: need at least two samples per ray and one channel");
    if (rays == 0) return 0;
    cudaStream_t st = as_stream(stream);
    ia::prof_begin("ia_ray_march(minmax_init)", st);
    minmax_init_kernel<<<1, 32, 0, st>>>(depth_minmax);
    IA_LAUNCH_CHECK("ia_ray_march(minmax_init)");
    ia::prof_begin("ia_ray_march", st);
    ray_march_kernel<<<(unsigned)cdiv(rays, 128), 128, 0, st>>>(colors, densities, depths, rays, S, C, white_back, rgb, depth, weights, depth_minmax);
    IA_LAUNCH_CHECK("ia_ray_march");
    return 0;
}

extern "C" int ia_depth_clamp(float* depth, int64_t n, const float* depth_minmax, void* stream) {
    IA_CHECK(depth && depth_minmax, "ia_depth_clamp: null argument");
    if (n == 0) return 0;
    ia::prof_begin("ia_depth_clamp", as_stream(stream));
    depth_clamp_kernel<<<(unsigned)cdiv(n, 256), 256, 0, as_stream(stream)>>>(depth, n, depth_minmax);
    IA_LAUNCH_CHECK("ia_depth_clamp");
    return 0;
}

extern "C" int ia_ray_sampler(const float* cam, int64_t cam_ld, int32_t B, int32_t res, float* origins, float* dirs, void* stream) {
    IA_CHECK(cam && origins && dirs, "ia_ray_sampler: null argument");
    int64_t total = (int64_t)B * res * res;
    if (total == 0) return 0;
    ia::prof_begin("ia_ray_sampler", as_stream(stream));
    ray_sampler_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(cam, cam_ld, B, res, origins, dirs);
    IA_LAUNCH_CHECK("ia_ray_sampler");
    return 0;
}
