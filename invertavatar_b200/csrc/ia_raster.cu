// UV rasterize / stitch kernels: GPU flood fill replacing cv2.floodFill, bilinear grid_sample, antialiased
// bilinear resize.  Semantics follow reference training_avatar_texture/volumetric_rendering/renderer.py:716-741
// (fill_mouth), triplane_v20.py:317-339 (rasterize) and the ATen ops they call.
#include <stdlib.h>

#include "ia_common.cuh"

using namespace ia;

// ------------------------------------------------------------------------------------------------
// fill_mouth: one CTA per image, bit-packed masks in shared memory.
//   passable(p) := 0 <= 255*alpha(p) - 255*alpha(0,0) <= 254   (FLOODFILL_FIXED_RANGE, lo=0, up=254, float image)
//   reached     := 4-connected component of (0,0) inside passable
// Propagation alternates a bit-parallel fill along rows (carry trick, one thread per row) with sequential
// sweeps along columns (32 columns per thread) until nothing changes.
// ------------------------------------------------------------------------------------------------
#define FM_MAXW 8  // words per row (W <= 256)
__global__ void __launch_bounds__(256) fill_mouth_kernel(const float* __restrict__ alpha, int64_t a_stride, int64_t a_batch,
                                                         int H, int W, int upper_row0, float* __restrict__ full_alpha,
                                                         float* __restrict__ mouth, float* __restrict__ upper_alpha) {
    __shared__ uint32_t pass[256][FM_MAXW];
    __shared__ uint32_t reach[256][FM_MAXW];
    __shared__ int changed;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const float* ab = alpha + b * a_batch;
    const int words = (W + 31) >> 5;
    const float seed = ab[0] * 255.0f;
    // build the passable mask: one thread per (row, word) pair, strided
    for (int i = tid; i < H * FM_MAXW; i += blockDim.x) {
        int y = i / FM_MAXW, k = i % FM_MAXW;
        uint32_t m = 0;
        if (k < words) {
            for (int bit = 0; bit < 32; ++bit) {
                int x = k * 32 + bit;
                if (x < W) {
                    float d = ab[((int64_t)y * W + x) * a_stride] * 255.0f - seed;
                    if (d >= 0.0f && d <= 254.0f) m |= 1u << bit;
                }
            }
        }
        pass[y][k] = m;
        reach[y][k] = 0;
    }
    __syncthreads();
    if (tid == 0) { reach[0][0] = 1u; pass[0][0] |= 1u; }
    __syncthreads();
    for (int iter = 0; iter < 4096; ++iter) {
        if (tid == 0) changed = 0;
        __syncthreads();
        // ---- rows: fill every passable run that contains a reached bit
        if (tid < H) {
            int y = tid;
            bool ch = false;
            uint32_t carry = 0;
            for (int k = 0; k < words; ++k) {  // towards increasing x
                uint32_t P = pass[y][k], R = reach[y][k];
                uint32_t x = R & P;
                if (carry && (P & 1u)) x |= 1u;
                uint64_t sum = (uint64_t)P + (uint64_t)x;
                uint32_t f = ((((uint32_t)sum) ^ P) & P) | x;
                carry = (uint32_t)(sum >> 32) & 1u;
                if (f & ~R) { reach[y][k] = R | f; ch = true; }
            }
            carry = 0;
            for (int k = words - 1; k >= 0; --k) {  // towards decreasing x (bit-reversed words)
                uint32_t P = __brev(pass[y][k]), R = __brev(reach[y][k]);
                uint32_t x = R & P;
                if (carry && (P & 1u)) x |= 1u;
                uint64_t sum = (uint64_t)P + (uint64_t)x;
                uint32_t f = ((((uint32_t)sum) ^ P) & P) | x;
                carry = (uint32_t)(sum >> 32) & 1u;
                if (f & ~R) { reach[y][k] = __brev(R | f); ch = true; }
            }
            if (ch) changed = 1;
        }
        __syncthreads();
        // ---- columns: sequential sweeps, 32 columns per thread
        if (tid < words) {
            int k = tid;
            bool ch = false;
            uint32_t prev = reach[0][k];
            for (int y = 1; y < H; ++y) {
                uint32_t R = reach[y][k];
                uint32_t n = R | (prev & pass[y][k]);
                if (n != R) { reach[y][k] = n; ch = true; }
                prev = n;
            }
            for (int y = H - 2; y >= 0; --y) {
                uint32_t R = reach[y][k];
                uint32_t n = R | (prev & pass[y][k]);
                if (n != R) { reach[y][k] = n; ch = true; }
                prev = n;
            }
            if (ch) changed = 1;
        }
        __syncthreads();
        if (!changed) break;
        __syncthreads();
    }
    // ---- outputs
    for (int i = tid; i < H * W; i += blockDim.x) {
        int y = i / W, x = i % W;
        float a = ab[(int64_t)i * a_stride];
        bool filled = (reach[y][x >> 5] >> (x & 31)) & 1u;
        float m = filled ? 0.0f : (255.0f - a * 255.0f) / 255.0f;
        int64_t o = (int64_t)b * H * W + i;
        if (mouth) mouth[o] = m;
        if (full_alpha) full_alpha[o] = fminf(fmaxf(a + m, 0.0f), 1.0f);
        if (upper_alpha) upper_alpha[o] = fminf(fmaxf(a + (y >= upper_row0 ? m : 0.0f), 0.0f), 1.0f);
    }
}

extern "C" int ia_fill_mouth(const float* alpha, int64_t a_stride, int64_t a_batch_stride, int32_t B, int32_t H, int32_t W,
                             int32_t upper_row0, float* full_alpha, float* mouth, float* upper_alpha, void* stream) {
    IA_CHECK(alpha, "ia_fill_mouth: null alpha");
    IA_CHECK(H >= 1 && W >= 1 && H <= 256 && W <= 256, "ia_fill_mouth: image must be at most 256x256 (got %dx%d)", H, W);
    if (B == 0) return 0;
    ia::prof_begin("ia_fill_mouth", as_stream(stream));
    fill_mouth_kernel<<<B, 256, 0, as_stream(stream)>>>(alpha, a_stride, a_batch_stride, H, W, upper_row0, full_alpha, mouth, upper_alpha);
    IA_LAUNCH_CHECK("ia_fill_mouth");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// grid_sample: bilinear, zeros padding, align_corners=False, NHWC
// ------------------------------------------------------------------------------------------------
__global__ void grid_sample_kernel(const float* __restrict__ in, int B, int Hi, int Wi, int C, int64_t in_ld,
                                   const float* __restrict__ grid, int64_t g_ld, int Ho, int Wo, float* __restrict__ out,
                                   int64_t out_ld) {
    const int groups = (C + 3) >> 2;
    int64_t total = (int64_t)B * Ho * Wo * groups;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int g = i % groups; int64_t pix = i / groups;
    int b = (int)(pix / ((int64_t)Ho * Wo));
    const float gx = grid[pix * g_ld + 0], gy = grid[pix * g_ld + 1];
    const float ix = ((gx + 1.f) * Wi - 1.f) / 2.f;
    const float iy = ((gy + 1.f) * Hi - 1.f) / 2.f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    // ATen weights: nw = (x1-ix)(y1-iy), ne = (ix-x0)(y1-iy), sw = (x1-ix)(iy-y0), se = (ix-x0)(iy-y0)
    const float wnw = ((float)x1 - ix) * ((float)y1 - iy);
    const float wne = (ix - (float)x0) * ((float)y1 - iy);
    const float wsw = ((float)x1 - ix) * (iy - (float)y0);
    const float wse = (ix - (float)x0) * (iy - (float)y0);
    const bool vx0 = x0 >= 0 && x0 < Wi, vx1 = x1 >= 0 && x1 < Wi, vy0 = y0 >= 0 && y0 < Hi, vy1 = y1 >= 0 && y1 < Hi;
    const int c0 = g * 4;
    const int nc = min(4, C - c0);
    const bool vec = (nc == 4) && ((in_ld & 3) == 0);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float* base = in + (int64_t)b * Hi * Wi * in_ld + c0;
    auto corner = [&](int yy, int xx, float w) {
        const float* p = base + ((int64_t)yy * Wi + xx) * in_ld;
        if (vec) {
            float4 t = *reinterpret_cast<const float4*>(p);
            acc[0] += t.x * w; acc[1] += t.y * w; acc[2] += t.z * w; acc[3] += t.w * w;
        } else {
            for (int k = 0; k < nc; ++k) acc[k] += p[k] * w;
        }
    };
    if (vy0 && vx0) corner(y0, x0, wnw);
    if (vy0 && vx1) corner(y0, x1, wne);
    if (vy1 && vx0) corner(y1, x0, wsw);
    if (vy1 && vx1) corner(y1, x1, wse);
    float* o = out + pix * out_ld + c0;
    if (vec && (out_ld & 3) == 0) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else for (int k = 0; k < nc; ++k) o[k] = acc[k];
}

extern "C" int ia_grid_sample(const float* in, int32_t B, int32_t Hi, int32_t Wi, int32_t C, int64_t in_ld, const float* grid,
                              int64_t g_ld, int32_t Ho, int32_t Wo, float* out, int64_t out_ld, void* stream) {
    IA_CHECK(in && grid && out, "ia_grid_sample: null tensor");
    IA_CHECK(C > 0 && in_ld >= C && out_ld >= C && g_ld >= 2, "ia_grid_sample: bad strides");
    int64_t total = (int64_t)B * Ho * Wo * ((C + 3) >> 2);
    if (total == 0) return 0;
    ia::prof_begin("ia_grid_sample", as_stream(stream));
    grid_sample_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(in, B, Hi, Wi, C, in_ld, grid, g_ld, Ho, Wo, out, out_ld);
    IA_LAUNCH_CHECK("ia_grid_sample");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// antialiased bilinear resize (separable taps; horizontal pass nested inside the vertical one, matching the
// width-then-height order of ATen's separable implementation)
// ------------------------------------------------------------------------------------------------
__global__ void resize_aa_kernel(ia_resize_params p) {
    const int groups = (p.C + 3) >> 2;
    int64_t total = (int64_t)p.B * p.oh * p.ow * groups;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int g = i % groups; int64_t t = i / groups;
    int ox = t % p.ow; t /= p.ow;
    int oy = t % p.oh; int b = (int)(t / p.oh);
    const int c0 = g * 4;
    const int nc = min(4, p.C - c0);
    const bool vec = (nc == 4) && ((p.in_ld & 3) == 0);
    const int ys = p.y_start[oy], yn = p.y_count[oy];
    const int xs = p.x_start[ox], xn = p.x_count[ox];
    const float* wy = p.y_w + (int64_t)oy * p.y_max_taps;
    const float* wx = p.x_w + (int64_t)ox * p.x_max_taps;
    const float* base = p.in + (int64_t)b * p.in_H * p.in_W * p.in_ld + c0;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ty = 0; ty < yn; ++ty) {
        const float* row = base + ((int64_t)(p.in_y0 + ys + ty) * p.in_W + p.in_x0 + xs) * p.in_ld;
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        for (int tx = 0; tx < xn; ++tx) {
            const float w = wx[tx];
            const float* q = row + (int64_t)tx * p.in_ld;
            if (vec) {
                float4 v = *reinterpret_cast<const float4*>(q);
                r[0] += v.x * w; r[1] += v.y * w; r[2] += v.z * w; r[3] += v.w * w;
            } else {
                for (int k = 0; k < nc; ++k) r[k] += q[k] * w;
            }
        }
        const float w = wy[ty];
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += r[k] * w;
    }
    float* o = p.out + (((int64_t)b * p.out_H + p.out_y0 + oy) * p.out_W + p.out_x0 + ox) * p.out_ld + c0;
    if (vec && (p.out_ld & 3) == 0) *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    else for (int k = 0; k < nc; ++k) o[k] = acc[k];
}

extern "C" int ia_resize_aa(const ia_resize_params* p, void* stream) {
    IA_CHECK(p && p->in && p->out, "ia_resize_aa: null tensor");
    IA_CHECK(p->y_start && p->y_count && p->y_w && p->x_start && p->x_count && p->x_w, "ia_resize_aa: null tap table");
    IA_CHECK(p->C > 0 && p->in_ld >= p->C && p->out_ld >= p->C, "ia_resize_aa: bad strides");
    int64_t total = (int64_t)p->B * p->oh * p->ow * ((p->C + 3) >> 2);
    if (total == 0) return 0;
    ia::prof_begin("ia_resize_aa", as_stream(stream));
    resize_aa_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(*p);
    IA_LAUNCH_CHECK("ia_resize_aa");
    return 0;
}


// ------------------------------------------------------------------------------------------------
// fused rasterize level: grid_sample (+) horizontal antialias taps, then vertical taps + static resize + alpha blend
// ------------------------------------------------------------------------------------------------
namespace {

// pass 1: thread = (NX adjacent output pixels (b, y, x'..x'+NX-1), lane); lane owns channel groups c4 = lane + LPP*k, k < KC
// (LPP lanes per pixel), so every LDG.128 of a warp covers contiguous 16*LPP bytes of one texel and the coordinate / weight
// arithmetic of a tap is done once per KC*4 loads.  The antialias (triangle) filter of a down-scaling level makes every
// 256^2 sample a tap of two neighbouring outputs: a thread that owns NX neighbours gathers each sample of their union once
// (NX = 4: 40 instead of 64 gathers at scale 8) and adds it to each output whose window holds it, in the same ascending-tap
// order as before -- results are bit-identical to NX = 1.
// COOP (lpp == 32: the 32 lanes of a warp share one pixel group): the bilinear set-up of a sample -- uv load, coordinates, the four
// zero-padding-masked weights and clamped texel offsets -- is computed ONCE, by lane (x - x_lo) % 32, and broadcast by shuffle,
// instead of by all 32 lanes; out-of-range corners read a clamped texel with weight 0 (branch-free, v * 0 adds nothing).  The
// accumulations use packed fp32 FMAs (fma.rn.f32x2: the roundings of the scalar FMAs).  Results are bit-identical to the plain path.
__device__ __forceinline__ void fma4(float4& a, const float4& v, float w) {
    const float2 ww = make_float2(w, w);
    const float2 lo = __ffma2_rn(make_float2(v.x, v.y), ww, make_float2(a.x, a.y));
    const float2 hi = __ffma2_rn(make_float2(v.z, v.w), ww, make_float2(a.z, a.w));
    a = make_float4(lo.x, lo.y, hi.x, hi.y);
}

template <int KC, int NX, bool COOP>
__global__ void __launch_bounds__(256) raster_hpass_kernel(ia_raster_level_params p, int lpp) {
    const int rg = (p.r + NX - 1) / NX;
    const int64_t total = (int64_t)p.B * p.UH * rg * lpp;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int lane = (int)(i % lpp);
    int64_t t = i / lpp;
    const int ox0 = (int)(t % rg) * NX; t /= rg;
    const int y = (int)(t % p.UH); const int b = (int)(t / p.UH);
    int xs[NX], xn[NX];
    int x_lo = 1 << 30, x_hi = -(1 << 30);
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        const bool on = ox0 + j < p.r;
        xs[j] = on ? p.ux_start[ox0 + j] : 0;
        xn[j] = on ? p.ux_count[ox0 + j] : 0;
        if (on) { x_lo = min(x_lo, xs[j]); x_hi = max(x_hi, xs[j] + xn[j]); }
    }
    const float* uvrow = p.uv + (int64_t)(b * p.UH + y) * p.UW * p.uv_ld;
    const float* tbase = p.tex + (int64_t)b * p.Ht * p.Wt * p.C + lane * 4;
    float4 acc[NX][KC];
#pragma unroll
    for (int j = 0; j < NX; ++j)
#pragma unroll
        for (int k = 0; k < KC; ++k) acc[j][k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int Wi = p.Wt, Hi = p.Ht;
    const int kstride = lpp * 4;
    // grid_sample(bilinear, zeros, align_corners=False), same arithmetic as grid_sample_kernel
    auto setup = [&](int x, float& wnw, float& wne, float& wsw, float& wse, int& x0, int& y0, bool& vx0, bool& vx1, bool& vy0, bool& vy1) {
        const float gx = uvrow[(int64_t)x * p.uv_ld + 0], gy = uvrow[(int64_t)x * p.uv_ld + 1];
        const float ix = ((gx + 1.f) * Wi - 1.f) / 2.f;
        const float iy = ((gy + 1.f) * Hi - 1.f) / 2.f;
        const float fx = floorf(ix), fy = floorf(iy);
        x0 = (int)fx; y0 = (int)fy;
        const int x1 = x0 + 1, y1 = y0 + 1;
        wnw = ((float)x1 - ix) * ((float)y1 - iy);
        wne = (ix - (float)x0) * ((float)y1 - iy);
        wsw = ((float)x1 - ix) * (iy - (float)y0);
        wse = (ix - (float)x0) * (iy - (float)y0);
        vx0 = x0 >= 0 && x0 < Wi; vx1 = x1 >= 0 && x1 < Wi; vy0 = y0 >= 0 && y0 < Hi; vy1 = y1 >= 0 && y1 < Hi;
    };
    auto accumulate = [&](int x, const float4 (&s)[KC]) {
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            const int tx = x - xs[j];
            if (tx >= 0 && tx < xn[j]) {
                const float w = p.ux_w[(int64_t)(ox0 + j) * p.ux_max_taps + tx];
#pragma unroll
                for (int k = 0; k < KC; ++k) fma4(acc[j][k], s[k], w);
            }
        }
    };
    if (COOP) {
        for (int xb = x_lo; xb < x_hi; xb += 32) {
            float mw00 = 0.f, mw01 = 0.f, mw10 = 0.f, mw11 = 0.f;
            int mo00 = 0, mo01 = 0, mo10 = 0, mo11 = 0;
            if (xb + lane < x_hi) {
                float wnw, wne, wsw, wse; int x0, y0; bool vx0, vx1, vy0, vy1;
                setup(xb + lane, wnw, wne, wsw, wse, x0, y0, vx0, vx1, vy0, vy1);
                const int x0c = min(max(x0, 0), Wi - 1), x1c = min(max(x0 + 1, 0), Wi - 1);
                const int y0c = min(max(y0, 0), Hi - 1), y1c = min(max(y0 + 1, 0), Hi - 1);
                mw00 = (vy0 && vx0) ? wnw : 0.f; mw01 = (vy0 && vx1) ? wne : 0.f;
                mw10 = (vy1 && vx0) ? wsw : 0.f; mw11 = (vy1 && vx1) ? wse : 0.f;
                mo00 = (y0c * Wi + x0c) * p.C; mo01 = (y0c * Wi + x1c) * p.C;
                mo10 = (y1c * Wi + x0c) * p.C; mo11 = (y1c * Wi + x1c) * p.C;
            }
            const int cnt = min(32, x_hi - xb);
            for (int q = 0; q < cnt; ++q) {
                const float w00 = __shfl_sync(0xffffffffu, mw00, q), w01 = __shfl_sync(0xffffffffu, mw01, q);
                const float w10 = __shfl_sync(0xffffffffu, mw10, q), w11 = __shfl_sync(0xffffffffu, mw11, q);
                const int o00 = __shfl_sync(0xffffffffu, mo00, q), o01 = __shfl_sync(0xffffffffu, mo01, q);
                const int o10 = __shfl_sync(0xffffffffu, mo10, q), o11 = __shfl_sync(0xffffffffu, mo11, q);
                float4 s[KC];
                float4 va[KC], vb[KC], vc[KC], vd[KC];
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    va[k] = __ldg(reinterpret_cast<const float4*>(tbase + o00 + k * kstride));
                    vb[k] = __ldg(reinterpret_cast<const float4*>(tbase + o01 + k * kstride));
                    vc[k] = __ldg(reinterpret_cast<const float4*>(tbase + o10 + k * kstride));
                    vd[k] = __ldg(reinterpret_cast<const float4*>(tbase + o11 + k * kstride));
                }
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    s[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    fma4(s[k], va[k], w00); fma4(s[k], vb[k], w01); fma4(s[k], vc[k], w10); fma4(s[k], vd[k], w11);
                }
                accumulate(xb + q, s);
            }
        }
    } else {
        for (int x = x_lo; x < x_hi; ++x) {
            float wnw, wne, wsw, wse; int x0, y0; bool vx0, vx1, vy0, vy1;
            setup(x, wnw, wne, wsw, wse, x0, y0, vx0, vx1, vy0, vy1);
            const int x1 = x0 + 1, y1 = y0 + 1;
            float4 s[KC];
#pragma unroll
            for (int k = 0; k < KC; ++k) s[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            auto corner = [&](int yy, int xx, float cw) {
                const float* q = tbase + ((int64_t)yy * Wi + xx) * p.C;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(q + k * kstride));
                    fma4(s[k], v, cw);
                }
            };
            if (vy0 && vx0) corner(y0, x0, wnw);
            if (vy0 && vx1) corner(y0, x1, wne);
            if (vy1 && vx0) corner(y1, x0, wsw);
            if (vy1 && vx1) corner(y1, x1, wse);
            accumulate(x, s);
        }
    }
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        if (ox0 + j >= p.r) continue;
        float* o = p.tmp + ((int64_t)(b * p.UH + y) * p.r + ox0 + j) * p.C + lane * 4;
#pragma unroll
        for (int k = 0; k < KC; ++k) *reinterpret_cast<float4*>(o + k * kstride) = acc[j][k];
    }
}

// Cell-merged horizontal pass (the 32 lanes of a warp share one pixel group, lpp == 32).  ncu of the kernel above
// (profiles/r2_misc_full_v24.txt): L1 data path 60-75 % busy -- every 256^2 sample gathers its four texels although, on a 32^2 ...
// 128^2 texture, 8 ... 2 consecutive samples of a row fall into the SAME texel cell.  The pass is linear in the texels:
//     out_j = sum_x a_j(x) * sum_corner w_corner(x) * T[cell(x)][corner]  =  sum_cells sum_corner T[cell][corner] * (sum_{x in cell} a_j(x) * w_corner(x)),
// so while consecutive samples stay in one cell only their 4 x NX scalar weight products are accumulated, and the cell's four
// texels are gathered ONCE, when the cell changes (warp-uniform decision: the cell of a sample is broadcast to all lanes).  A
// different summation order than the per-sample kernel (fp32 reassociation, ~1e-7 relative; IA_RASTER_MERGE=0 selects the
// per-sample kernel, which is the bit-exact restatement of grid_sample followed by the resize).
template <int KC, int NX>
__global__ void __launch_bounds__(256) raster_hpass_merge_kernel(ia_raster_level_params p) {
    const int rg = (p.r + NX - 1) / NX;
    const int64_t total = (int64_t)p.B * p.UH * rg * 32;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int lane = (int)(i & 31);
    int64_t t = i >> 5;
    const int ox0 = (int)(t % rg) * NX; t /= rg;
    const int y = (int)(t % p.UH); const int b = (int)(t / p.UH);
    int xs[NX], xn[NX];
    int x_lo = 1 << 30, x_hi = -(1 << 30);
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        const bool on = ox0 + j < p.r;
        xs[j] = on ? p.ux_start[ox0 + j] : 0;
        xn[j] = on ? p.ux_count[ox0 + j] : 0;
        if (on) { x_lo = min(x_lo, xs[j]); x_hi = max(x_hi, xs[j] + xn[j]); }
    }
    const float* uvrow = p.uv + (int64_t)(b * p.UH + y) * p.UW * p.uv_ld;
    const float* tbase = p.tex + (int64_t)b * p.Ht * p.Wt * p.C + lane * 4;
    float4 acc[NX][KC];
    float cw[NX][4];
#pragma unroll
    for (int j = 0; j < NX; ++j) {
#pragma unroll
        for (int k = 0; k < KC; ++k) acc[j][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        cw[j][0] = cw[j][1] = cw[j][2] = cw[j][3] = 0.f;
    }
    const int Wi = p.Wt, Hi = p.Ht;
    constexpr int kstride = 32 * 4;
    int c00 = -1, c01 = -1, c10 = -1, c11 = -1;           // texel offsets of the open cell (-1: none yet)
    auto flush = [&]() {
        float4 va[KC], vb[KC], vc[KC], vd[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            va[k] = __ldg(reinterpret_cast<const float4*>(tbase + c00 + k * kstride));
            vb[k] = __ldg(reinterpret_cast<const float4*>(tbase + c01 + k * kstride));
            vc[k] = __ldg(reinterpret_cast<const float4*>(tbase + c10 + k * kstride));
            vd[k] = __ldg(reinterpret_cast<const float4*>(tbase + c11 + k * kstride));
        }
#pragma unroll
        for (int j = 0; j < NX; ++j) {
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                fma4(acc[j][k], va[k], cw[j][0]); fma4(acc[j][k], vb[k], cw[j][1]);
                fma4(acc[j][k], vc[k], cw[j][2]); fma4(acc[j][k], vd[k], cw[j][3]);
            }
        }
    };
    // 32 samples per round, ONE PER LANE: the lane computes its sample's bilinear set-up and the 4 x NX products
    // a_j(x) * w_corner(x); runs of lanes with equal cells are found with one ballot and summed with a segmented warp scan
    // (5 shuffle steps); only the per-run work -- 4 offsets + 4 x NX sums broadcast, gathers, FMAs -- is serial.
    for (int xb = x_lo; xb < x_hi; xb += 32) {
        const int cnt = min(32, x_hi - xb);
        const int x = xb + lane;
        int mo00 = -2 - lane, mo01 = 0, mo10 = 0, mo11 = 0;            // lanes past the end: cells of their own, never equal to a neighbour's
        float pr[NX][4];
#pragma unroll
        for (int j = 0; j < NX; ++j) pr[j][0] = pr[j][1] = pr[j][2] = pr[j][3] = 0.f;
        if (lane < cnt) {
            // grid_sample(bilinear, zeros, align_corners=False), same arithmetic as raster_hpass_kernel's set-up
            const float gx = uvrow[(int64_t)x * p.uv_ld + 0], gy = uvrow[(int64_t)x * p.uv_ld + 1];
            const float ix = ((gx + 1.f) * Wi - 1.f) / 2.f;
            const float iy = ((gy + 1.f) * Hi - 1.f) / 2.f;
            const float fx = floorf(ix), fy = floorf(iy);
            const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
            const float wnw = ((float)x1 - ix) * ((float)y1 - iy);
            const float wne = (ix - (float)x0) * ((float)y1 - iy);
            const float wsw = ((float)x1 - ix) * (iy - (float)y0);
            const float wse = (ix - (float)x0) * (iy - (float)y0);
            const bool vx0 = x0 >= 0 && x0 < Wi, vx1 = x1 >= 0 && x1 < Wi, vy0 = y0 >= 0 && y0 < Hi, vy1 = y1 >= 0 && y1 < Hi;
            const int x0c = min(max(x0, 0), Wi - 1), x1c = min(max(x1, 0), Wi - 1);
            const int y0c = min(max(y0, 0), Hi - 1), y1c = min(max(y1, 0), Hi - 1);
            const float mw00 = (vy0 && vx0) ? wnw : 0.f, mw01 = (vy0 && vx1) ? wne : 0.f;
            const float mw10 = (vy1 && vx0) ? wsw : 0.f, mw11 = (vy1 && vx1) ? wse : 0.f;
            mo00 = (y0c * Wi + x0c) * p.C; mo01 = (y0c * Wi + x1c) * p.C;
            mo10 = (y1c * Wi + x0c) * p.C; mo11 = (y1c * Wi + x1c) * p.C;
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                const int tx = x - xs[j];
                if (tx >= 0 && tx < xn[j]) {
                    const float a = p.ux_w[(ox0 + j) * p.ux_max_taps + tx];
                    pr[j][0] = a * mw00; pr[j][1] = a * mw01; pr[j][2] = a * mw10; pr[j][3] = a * mw11;
                }
            }
        }
        // run heads: lane 0, and every lane whose cell differs from its left neighbour's
        const int l00 = __shfl_up_sync(0xffffffffu, mo00, 1), l01 = __shfl_up_sync(0xffffffffu, mo01, 1);
        const int l10 = __shfl_up_sync(0xffffffffu, mo10, 1), l11 = __shfl_up_sync(0xffffffffu, mo11, 1);
        const bool head = lane == 0 || l00 != mo00 || l01 != mo01 || l10 != mo10 || l11 != mo11;
        const unsigned heads = __ballot_sync(0xffffffffu, head);
        const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));      // first lane of this lane's run
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const bool take = lane - d >= start;
#pragma unroll
            for (int j = 0; j < NX; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float up = __shfl_up_sync(0xffffffffu, pr[j][c], d);
                    if (take) pr[j][c] += up;
                }
        }
        // the runs of this round, left to right (warp-uniform loop); a run that continues the open cell adds to its sums
        unsigned m = heads & (cnt == 32 ? 0xffffffffu : ((1u << cnt) - 1u));
        while (m) {
            const int h = __ffs(m) - 1;
            m &= m - 1;
            const int tail = (m ? __ffs(m) - 1 : cnt) - 1;
            const int o00 = __shfl_sync(0xffffffffu, mo00, h), o01 = __shfl_sync(0xffffffffu, mo01, h);
            const int o10 = __shfl_sync(0xffffffffu, mo10, h), o11 = __shfl_sync(0xffffffffu, mo11, h);
            const bool same = o00 == c00 && o01 == c01 && o10 == c10 && o11 == c11;
            if (!same && c00 >= 0) flush();
#pragma unroll
            for (int j = 0; j < NX; ++j)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float sum = __shfl_sync(0xffffffffu, pr[j][c], tail);
                    cw[j][c] = same ? cw[j][c] + sum : sum;
                }
            c00 = o00; c01 = o01; c10 = o10; c11 = o11;
        }
    }
    if (c00 >= 0) flush();
#pragma unroll
    for (int j = 0; j < NX; ++j) {
        if (ox0 + j >= p.r) continue;
        float* o = p.tmp + ((int64_t)(b * p.UH + y) * p.r + ox0 + j) * p.C + lane * 4;
#pragma unroll
        for (int k = 0; k < KC; ++k) *reinterpret_cast<float4*>(o + k * kstride) = acc[j][k];
    }
}

__global__ void __launch_bounds__(256) raster_vpass_kernel(ia_raster_level_params p) {
    const int groups = p.C >> 2;
    const int64_t total = (int64_t)p.B * p.r * p.r * groups;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c0 = (int)(i % groups) * 4;
    int64_t t = i / groups;
    const int ox = (int)(t % p.r); t /= p.r;
    const int oy = (int)(t % p.r); const int b = (int)(t / p.r);
    // vertical antialias taps over the horizontally filtered samples
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        const int ys = p.uy_start[oy], yn = p.uy_count[oy];
        const float* wy = p.uy_w + (int64_t)oy * p.uy_max_taps;
        const float* q = p.tmp + ((int64_t)(b * p.UH + ys) * p.r + ox) * p.C + c0;
        for (int ty = 0; ty < yn; ++ty, q += (int64_t)p.r * p.C) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(q));
            const float w = wy[ty];
            f.x += v.x * w; f.y += v.y * w; f.z += v.z * w; f.w += v.w * w;
        }
    }
    // static crop resized to r x r (rows of horizontal taps, then the vertical weight: ATen's pass order)
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        const int ys = p.sy_start[oy], yn = p.sy_count[oy], xs = p.sx_start[ox], xn = p.sx_count[ox];
        const float* wy = p.sy_w + (int64_t)oy * p.sy_max_taps;
        const float* wx = p.sx_w + (int64_t)ox * p.sx_max_taps;
        const bool vec = (p.stat_ld & 3) == 0;
        for (int ty = 0; ty < yn; ++ty) {
            const float* row = p.stat + (((int64_t)b * p.SH + p.sy0 + ys + ty) * p.SW + p.sx0 + xs) * p.stat_ld + c0;
            float4 rr = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int tx = 0; tx < xn; ++tx) {
                const float* q = row + (int64_t)tx * p.stat_ld;
                float4 v;
                if (vec) v = __ldg(reinterpret_cast<const float4*>(q));
                else v = make_float4(q[0], q[1], q[2], q[3]);
                const float w = wx[tx];
                rr.x += v.x * w; rr.y += v.y * w; rr.z += v.z * w; rr.w += v.w * w;
            }
            const float w = wy[ty];
            s.x += rr.x * w; s.y += rr.y * w; s.z += rr.z * w; s.w += rr.w * w;
        }
    }
    const float al = p.alpha[((int64_t)b * p.r + oy) * p.r + ox];
    const float bl = 1.f - al;
    float* o = p.out + (((int64_t)b * p.r + oy) * p.r + ox) * p.out_ld + c0;
    const float4 res = make_float4(f.x * al + s.x * bl, f.y * al + s.y * bl, f.z * al + s.z * bl, f.w * al + s.w * bl);
    if ((p.out_ld & 3) == 0) *reinterpret_cast<float4*>(o) = res;
    else { o[0] = res.x; o[1] = res.y; o[2] = res.z; o[3] = res.w; }
}

// One launch per level: the 2-D form of the cell merge.  The antialias filter is separable, so an output pixel is
//     out(y', x') = sum_{y, x} ay(y', y) * ax(x', x) * bilinear(tex, uv[y][x])
// over its (2 * scale)^2 window of 256^2 samples -- 256 samples for a 32^2 level, 64 for 64^2, 16 for 128^2 -- which on a smooth UV
// map touch only a handful of texel cells.  One warp per output pixel: the LANES take the samples of the window (set-up, weight
// product ay * ax * w_corner), runs of equal cells are summed with a segmented warp scan as in raster_hpass_merge_kernel, and the
// run sums are collected in a table of up to 32 cells held one per lane (key = the cell's four clamped texel offsets; look-up =
// one ballot).  When the window is done the lanes switch to channels: each table entry's four texels are gathered ONCE and applied
// with its four summed weights.  (A table keyed per TEXEL gathers fewer texels -- a 3 x 3 block of cells shares 16 -- but needs four
// look-ups per run: measured slower, 123 / 225 us against 91 / 198 us for the 32^2 / 64^2 levels; the serial per-run work bounds this kernel.)  Then the vertical-pass tail (static-crop resize, alpha blend) runs in the same thread.  Neither
// the [B][256][r][C] intermediate nor the second launch exists; texel gathers drop from 4 per sample and output column to 4 per
// touched cell and output pixel.  Deterministic (fixed sample, run and table order); against the two-pass kernels the sums are
// reassociated (~1e-6 relative), IA_RASTER_FUSED=0 selects them.  Three CTAs per SM (80 registers: the gathers of a cell go out KB
// channel groups at a time) measured faster than two with all gathers of a cell in flight (91 -> 84 us, 199 -> 178 us, 421 -> 336 us per
// level) and than four (64 registers, spills: 143 / 285 / 324 us).
__device__ __forceinline__ void prefetch_l1(const void* q) { asm volatile("prefetch.global.L1 [%0];" ::"l"(q)); }

template <int KC>
__global__ void __launch_bounds__(256, 3) raster_fused_kernel(ia_raster_level_params p) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= (int64_t)p.B * p.r * p.r) return;
    const int ox = (int)(warp % p.r);
    const int64_t tq = warp / p.r;
    const int oy = (int)(tq % p.r), b = (int)(tq / p.r);
    const int ys = p.uy_start[oy], yn = p.uy_count[oy], xs = p.ux_start[ox], xn = p.ux_count[ox];
    const int n = yn * xn;
    const float* wyp = p.uy_w + (int64_t)oy * p.uy_max_taps;
    const float* wxp = p.ux_w + (int64_t)ox * p.ux_max_taps;
    const float* uvb = p.uv + (int64_t)b * p.UH * p.UW * p.uv_ld;
    const float* tbase = p.tex + (int64_t)b * p.Ht * p.Wt * p.C + lane * 4;
    const int Wi = p.Wt, Hi = p.Ht;
    constexpr int kstride = 32 * 4;
    constexpr int KB = KC == 4 ? 2 : KC;                 // channel groups processed together (see flush_all)
    const unsigned full = 0xffffffffu;
    // the static-crop taps the tail of this thread reads are known now: pull their lines towards L1 while the window is processed
    {
        const int sys_ = p.sy_start[oy], syn_ = p.sy_count[oy], sxs_ = p.sx_start[ox], sxn_ = p.sx_count[ox];
        for (int ty = 0; ty < syn_; ++ty)
            for (int tx = 0; tx < sxn_; ++tx) {
                const float* q = p.stat + (((int64_t)b * p.SH + p.sy0 + sys_ + ty) * p.SW + p.sx0 + sxs_ + tx) * p.stat_ld + lane * 4;
#pragma unroll
                for (int k = 0; k < KC; ++k) prefetch_l1(q + k * kstride);
            }
    }
    float4 acc[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    int k00 = -1, k01 = -1, k10 = -1, k11 = -1;          // this lane's table entry: cell key ...
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;        // ... and its four summed corner weights
    int nent = 0;                                        // entries in use (warp-uniform)
    auto flush_all = [&]() {
        for (int e = 0; e < nent; ++e) {
            const int o00 = __shfl_sync(full, k00, e), o01 = __shfl_sync(full, k01, e);
            const int o10 = __shfl_sync(full, k10, e), o11 = __shfl_sync(full, k11, e);
            const float w0 = __shfl_sync(full, t0, e), w1 = __shfl_sync(full, t1, e);
            const float w2 = __shfl_sync(full, t2, e), w3 = __shfl_sync(full, t3, e);
            // (KB channel groups at a time: 4 x KB gathers in flight -- with all four groups of a 512-channel level the kernel needs
            // 128 registers and two CTAs per SM; with two it fits three, and the occupancy is worth more than the deeper batch)
#pragma unroll
            for (int k0 = 0; k0 < KC; k0 += KB) {
                float4 va[KB], vb[KB], vc[KB], vd[KB];
#pragma unroll
                for (int k = 0; k < KB; ++k) {
                    va[k] = __ldg(reinterpret_cast<const float4*>(tbase + o00 + (k0 + k) * kstride));
                    vb[k] = __ldg(reinterpret_cast<const float4*>(tbase + o01 + (k0 + k) * kstride));
                    vc[k] = __ldg(reinterpret_cast<const float4*>(tbase + o10 + (k0 + k) * kstride));
                    vd[k] = __ldg(reinterpret_cast<const float4*>(tbase + o11 + (k0 + k) * kstride));
                }
#pragma unroll
                for (int k = 0; k < KB; ++k) {
                    fma4(acc[k0 + k], va[k], w0); fma4(acc[k0 + k], vb[k], w1); fma4(acc[k0 + k], vc[k], w2); fma4(acc[k0 + k], vd[k], w3);
                }
            }
        }
        nent = 0;
    };
    for (int base = 0; base < n; base += 32) {
        const int cnt = min(32, n - base);
        const int idx = base + lane;
        int mo00 = -2 - lane, mo01 = 0, mo10 = 0, mo11 = 0;            // lanes past the end: cells of their own
        float pr0 = 0.f, pr1 = 0.f, pr2 = 0.f, pr3 = 0.f;
        if (lane < cnt) {
            const int sy = idx / xn, sx = idx - sy * xn;
            const float a = wyp[sy] * wxp[sx];
            const float* uvp = uvb + ((int64_t)(ys + sy) * p.UW + xs + sx) * p.uv_ld;
            // grid_sample(bilinear, zeros, align_corners=False), same arithmetic as raster_hpass_kernel's set-up
            const float gx = uvp[0], gy = uvp[1];
            const float ix = ((gx + 1.f) * Wi - 1.f) / 2.f;
            const float iy = ((gy + 1.f) * Hi - 1.f) / 2.f;
            const float fx = floorf(ix), fy = floorf(iy);
            const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
            const float wnw = ((float)x1 - ix) * ((float)y1 - iy);
            const float wne = (ix - (float)x0) * ((float)y1 - iy);
            const float wsw = ((float)x1 - ix) * (iy - (float)y0);
            const float wse = (ix - (float)x0) * (iy - (float)y0);
            const bool vx0 = x0 >= 0 && x0 < Wi, vx1 = x1 >= 0 && x1 < Wi, vy0 = y0 >= 0 && y0 < Hi, vy1 = y1 >= 0 && y1 < Hi;
            const int x0c = min(max(x0, 0), Wi - 1), x1c = min(max(x1, 0), Wi - 1);
            const int y0c = min(max(y0, 0), Hi - 1), y1c = min(max(y1, 0), Hi - 1);
            pr0 = (vy0 && vx0) ? a * wnw : 0.f; pr1 = (vy0 && vx1) ? a * wne : 0.f;
            pr2 = (vy1 && vx0) ? a * wsw : 0.f; pr3 = (vy1 && vx1) ? a * wse : 0.f;
            mo00 = (y0c * Wi + x0c) * p.C; mo01 = (y0c * Wi + x1c) * p.C;
            mo10 = (y1c * Wi + x0c) * p.C; mo11 = (y1c * Wi + x1c) * p.C;
        }
        const int l00 = __shfl_up_sync(full, mo00, 1), l01 = __shfl_up_sync(full, mo01, 1);
        const int l10 = __shfl_up_sync(full, mo10, 1), l11 = __shfl_up_sync(full, mo11, 1);
        const bool head = lane == 0 || l00 != mo00 || l01 != mo01 || l10 != mo10 || l11 != mo11;
        const unsigned heads = __ballot_sync(full, head);
        const int start = 31 - __clz(heads & (full >> (31 - lane)));      // first lane of this lane's run
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const bool take = lane - d >= start;
            const float u0 = __shfl_up_sync(full, pr0, d), u1 = __shfl_up_sync(full, pr1, d);
            const float u2 = __shfl_up_sync(full, pr2, d), u3 = __shfl_up_sync(full, pr3, d);
            if (take) { pr0 += u0; pr1 += u1; pr2 += u2; pr3 += u3; }
        }
        unsigned m = heads & (cnt == 32 ? full : ((1u << cnt) - 1u));
        while (m) {
            const int h = __ffs(m) - 1;
            m &= m - 1;
            const int tail = (m ? __ffs(m) - 1 : cnt) - 1;
            const int o00 = __shfl_sync(full, mo00, h), o01 = __shfl_sync(full, mo01, h);
            const int o10 = __shfl_sync(full, mo10, h), o11 = __shfl_sync(full, mo11, h);
            const float r0 = __shfl_sync(full, pr0, tail), r1 = __shfl_sync(full, pr1, tail);
            const float r2 = __shfl_sync(full, pr2, tail), r3 = __shfl_sync(full, pr3, tail);
            const unsigned hit = __ballot_sync(full, lane < nent && k00 == o00 && k01 == o01 && k10 == o10 && k11 == o11);
            if (hit) {
                if (lane == __ffs(hit) - 1) { t0 += r0; t1 += r1; t2 += r2; t3 += r3; }
            } else {
                if (nent == 32) flush_all();
                if (lane == nent) { k00 = o00; k01 = o01; k10 = o10; k11 = o11; t0 = r0; t1 = r1; t2 = r2; t3 = r3; }
                ++nent;
                // a new cell: start fetching this lane's share of its four texels (the gathers of flush_all then hit L1)
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    prefetch_l1(tbase + o00 + k * kstride); prefetch_l1(tbase + o01 + k * kstride);
                    prefetch_l1(tbase + o10 + k * kstride); prefetch_l1(tbase + o11 + k * kstride);
                }
            }
        }
    }
    flush_all();
    // tail of the two-pass path's second kernel: static crop resized to r x r (rows of horizontal taps, then the vertical weight:
    // ATen's pass order), alpha blend, store -- this lane's channel groups lane*4 + 128*k
    const int sys = p.sy_start[oy], syn = p.sy_count[oy], sxs = p.sx_start[ox], sxn = p.sx_count[ox];
    const float* swy = p.sy_w + (int64_t)oy * p.sy_max_taps;
    const float* swx = p.sx_w + (int64_t)ox * p.sx_max_taps;
    const bool vec = (p.stat_ld & 3) == 0;
    const float al = p.alpha[((int64_t)b * p.r + oy) * p.r + ox];
    const float bl = 1.f - al;
    // (taps outside, channel groups inside: the KC loads of a tap are independent and go out back to back; per channel the
    // sums and their order are raster_vpass_kernel's)
#pragma unroll
    for (int k0 = 0; k0 < KC; k0 += KB) {
        float4 sacc[KB];
#pragma unroll
        for (int k = 0; k < KB; ++k) sacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int ty = 0; ty < syn; ++ty) {
            const float* row = p.stat + (((int64_t)b * p.SH + p.sy0 + sys + ty) * p.SW + p.sx0 + sxs) * p.stat_ld + lane * 4 + k0 * kstride;
            float4 rr[KB];
#pragma unroll
            for (int k = 0; k < KB; ++k) rr[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int tx = 0; tx < sxn; ++tx) {
                float4 v[KB];
#pragma unroll
                for (int k = 0; k < KB; ++k) {
                    const float* q = row + (int64_t)tx * p.stat_ld + k * kstride;
                    if (vec) v[k] = __ldg(reinterpret_cast<const float4*>(q));
                    else v[k] = make_float4(q[0], q[1], q[2], q[3]);
                }
                const float w = swx[tx];
#pragma unroll
                for (int k = 0; k < KB; ++k) { rr[k].x += v[k].x * w; rr[k].y += v[k].y * w; rr[k].z += v[k].z * w; rr[k].w += v[k].w * w; }
            }
            const float w = swy[ty];
#pragma unroll
            for (int k = 0; k < KB; ++k) { sacc[k].x += rr[k].x * w; sacc[k].y += rr[k].y * w; sacc[k].z += rr[k].z * w; sacc[k].w += rr[k].w * w; }
        }
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            const float4 f = acc[k0 + k];
            float* o = p.out + (((int64_t)b * p.r + oy) * p.r + ox) * p.out_ld + lane * 4 + (k0 + k) * kstride;
            const float4 res = make_float4(f.x * al + sacc[k].x * bl, f.y * al + sacc[k].y * bl, f.z * al + sacc[k].z * bl, f.w * al + sacc[k].w * bl);
            if ((p.out_ld & 3) == 0) *reinterpret_cast<float4*>(o) = res;
            else { o[0] = res.x; o[1] = res.y; o[2] = res.z; o[3] = res.w; }
        }
    }
}

}  // namespace

extern "C" int ia_raster_level(const ia_raster_level_params* p, void* stream) {
    IA_CHECK(p && p->tex && p->uv && p->tmp && p->stat && p->alpha && p->out, "ia_raster_level: null tensor");
    IA_CHECK(p->C > 0 && (p->C & 3) == 0 && p->uv_ld >= 2 && p->stat_ld >= p->C && p->out_ld >= p->C, "ia_raster_level: bad channel count / strides");
    IA_CHECK(p->ux_start && p->uy_start && p->sx_start && p->sy_start, "ia_raster_level: null tap table");
    if ((int64_t)p->B * p->r * p->r == 0) return 0;
    const int groups = p->C >> 2;
    int kc = 1;
    if (groups % 128 == 0) kc = 4; else if (groups % 64 == 0) kc = 2;
    while (groups / kc > 32 && kc < 4) kc *= 2;
    { const char* e = getenv("IA_RASTER_KC"); if (e) { const int v = atoi(e); if ((v == 1 || v == 2 || v == 4) && groups % v == 0 && groups / v <= 64) kc = v; } }
    IA_CHECK(groups % kc == 0 && groups / kc <= 64, "ia_raster_level: unsupported channel count %d", p->C);
    const int lpp = groups / kc;
    // outputs per thread: 4 when the level shrinks the 256^2 samples by >= 4 (neighbouring windows overlap by half), 2 at
    // scale 2, 1 when every output has its own sample(s); IA_RASTER_NX overrides (1 = the one-output kernel)
    int nx = p->UW >= 4 * p->r ? 4 : (p->UW >= 2 * p->r ? 2 : 1);
    { const char* e = getenv("IA_RASTER_NX"); if (e) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) nx = v; } }
    if (kc == 4 && nx == 4) nx = 2;      // 4 x 4 float4 accumulators + the sample would not fit the register budget
    // one-launch 2-D cell merge when a warp's lanes cover the channels (C = 128 * kc) and the level shrinks the samples
    {
        // (measured per level, tools/prof_raster.py: 135 -> 84 us @32^2, 292 -> 178 us @64^2, 411 -> 336 us @128^2 against the two-pass kernels)
        int fused_min_scale = 2;
        { const char* e = getenv("IA_RASTER_FUSED_SCALE"); if (e && atoi(e) > 0) fused_min_scale = atoi(e); }
        bool fused = lpp == 32 && groups == 32 * kc && p->UW >= fused_min_scale * p->r && p->UH >= fused_min_scale * p->r &&
                     (reinterpret_cast<uintptr_t>(p->tex) & 15) == 0;
        { const char* e = getenv("IA_RASTER_FUSED"); if (e && atoi(e) == 0) fused = false; }
        { const char* e = getenv("IA_RASTER_MERGE"); if (e && atoi(e) == 0) fused = false; }
        if (fused) {
            const int64_t warps = (int64_t)p->B * p->r * p->r;
            const unsigned gridf = (unsigned)cdiv(warps * 32, 256);
            ia::prof_begin("ia_raster_level(fused)", as_stream(stream));
            if (kc == 4) raster_fused_kernel<4><<<gridf, 256, 0, as_stream(stream)>>>(*p);
            else if (kc == 2) raster_fused_kernel<2><<<gridf, 256, 0, as_stream(stream)>>>(*p);
            else raster_fused_kernel<1><<<gridf, 256, 0, as_stream(stream)>>>(*p);
            IA_LAUNCH_CHECK("ia_raster_level(fused)");
            return 0;
        }
    }
    const int64_t total1 = (int64_t)p->B * p->UH * cdiv(p->r, nx) * lpp;
    const unsigned grid1 = (unsigned)cdiv(total1, 256);
    ia::prof_begin("ia_raster_level(hpass)", as_stream(stream));
    static int coop_env = -1;      // IA_RASTER_COOP=0: every lane computes every sample's set-up (the round-1 kernel)
    if (coop_env < 0) { const char* e = getenv("IA_RASTER_COOP"); coop_env = e ? atoi(e) : 1; }
    const bool coop = coop_env != 0 && lpp == 32;
    // cell-merged gathers (IA_RASTER_MERGE=0: one gather set per sample, read per call so that tests can compare the two)
    int merge_min_scale = 2;                      // merge when the level shrinks the samples by at least this factor (IA_RASTER_MERGE_SCALE)
    { const char* e = getenv("IA_RASTER_MERGE_SCALE"); if (e && atoi(e) > 0) merge_min_scale = atoi(e); }
    bool merge = coop && p->UW >= merge_min_scale * p->r;
    { const char* e = getenv("IA_RASTER_MERGE"); if (e && atoi(e) == 0) merge = false; }
#define IA_HPASS(K, N) do { if (merge) raster_hpass_merge_kernel<K, N><<<grid1, 256, 0, as_stream(stream)>>>(*p); else if (coop) raster_hpass_kernel<K, N, true><<<grid1, 256, 0, as_stream(stream)>>>(*p, lpp); else raster_hpass_kernel<K, N, false><<<grid1, 256, 0, as_stream(stream)>>>(*p, lpp); } while (0)
    if (kc == 4) { if (nx == 2) IA_HPASS(4, 2); else IA_HPASS(4, 1); }
    else if (kc == 2) { if (nx == 4) IA_HPASS(2, 4); else if (nx == 2) IA_HPASS(2, 2); else IA_HPASS(2, 1); }
    else { if (nx == 4) IA_HPASS(1, 4); else if (nx == 2) IA_HPASS(1, 2); else IA_HPASS(1, 1); }
#undef IA_HPASS
    IA_LAUNCH_CHECK("ia_raster_level(hpass)");
    const int64_t total2 = (int64_t)p->B * p->r * p->r * groups;
    ia::prof_begin("ia_raster_level(vpass)", as_stream(stream));
    raster_vpass_kernel<<<(unsigned)cdiv(total2, 256), 256, 0, as_stream(stream)>>>(*p);
    IA_LAUNCH_CHECK("ia_raster_level(vpass)");
    return 0;
}


// ------------------------------------------------------------------------------------------------
// plane stitch (triplane_v20.py:119-128): copy of the static planes with the face-backbone output blended into plane 0
// inside the face window, written once in the renderer's storage format
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stitch_planes_kernel(ia_stitch_params p) {
    const int groups = p.C >> 2;
    const int64_t total = (int64_t)p.B * p.H * p.W * groups;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % groups) * 4;
    int64_t t = i / groups;
    const int x = (int)(t % p.W); t /= p.W;
    const int y = (int)(t % p.H); const int b = (int)(t / p.H);
    const int64_t pix = ((int64_t)b * p.H + y) * p.W + x;
    float4 v = __ldg(reinterpret_cast<const float4*>(p.planes + pix * p.planes_ld + c));
    const int wy = y - p.y0, wx = x - p.x0;
    if (c < 32 && wy >= 0 && wy < p.wh && wx >= 0 && wx < p.ww) {
        const int64_t wp = ((int64_t)b * p.wh + wy) * p.ww + wx;
        const float a = p.alpha[wp];
        const float4 s = __ldg(reinterpret_cast<const float4*>(p.stitch + wp * 32 + c));
        // a*alpha + b*(1-alpha), the arithmetic of ia_lerp_alpha
        v.x = s.x * a + v.x * (1.f - a); v.y = s.y * a + v.y * (1.f - a);
        v.z = s.z * a + v.z * (1.f - a); v.w = s.w * a + v.w * (1.f - a);
    }
    if (p.out_fmt == IA_OPFMT_F16X1) {
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        uint2 o;
        o.x = *reinterpret_cast<const uint32_t*>(&h0); o.y = *reinterpret_cast<const uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + pix * p.C + c) = o;
    } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + pix * p.C + c) = v;
    }
}

extern "C" int ia_stitch_planes(const ia_stitch_params* p, void* stream) {
    IA_CHECK(p && p->planes && p->stitch && p->alpha && p->out, "ia_stitch_planes: null tensor");
    IA_CHECK(p->C >= 32 && (p->C & 7) == 0 && (p->planes_ld & 3) == 0, "ia_stitch_planes: C must be a multiple of 8 (>= 32), fp32 pixel stride a multiple of 4");
    IA_CHECK(p->y0 >= 0 && p->x0 >= 0 && p->y0 + p->wh <= p->H && p->x0 + p->ww <= p->W, "ia_stitch_planes: window outside the planes");
    IA_CHECK(p->out_fmt == IA_OPFMT_BF16X3 || p->out_fmt == IA_OPFMT_F16X1, "ia_stitch_planes: out_fmt must be 0 (fp32) or IA_OPFMT_F16X1");
    IA_CHECK((reinterpret_cast<uintptr_t>(p->planes) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->stitch) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->out) & 15) == 0,
             "ia_stitch_planes: tensors must be 16-byte aligned");
    const int64_t total = (int64_t)p->B * p->H * p->W * (p->C >> 2);
    if (total == 0) return 0;
    ia::prof_begin("ia_stitch_planes", as_stream(stream));
    stitch_planes_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(*p);
    IA_LAUNCH_CHECK("ia_stitch_planes");
    return 0;
}
