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

#include "ia_common.cuh"

using namespace ia;

namespace {

constexpr int kHidden = 64;
constexpr int kFeat = 32;
constexpr int kOut = 33;
constexpr int kRowLd = 33;          // padded row of the per-sample feature/colour scratch
constexpr int kWarpsPerCta = 4;
constexpr int kMaxS = 192;          // Dc + Df

// decoder weights with the FullyConnectedLayer runtime gains folded in (networks_stylegan2.py:111-115)
__constant__ float c_w1[kHidden * kFeat];   // [j][i]
__constant__ float c_b1[kHidden];
__constant__ float c_w2t[kHidden * kOut];   // [j][k]  (transposed second layer)
__constant__ float c_b2[kOut];
__device__ float g_dec_stage[kHidden * kFeat + kHidden + kHidden * kOut + kOut];

__global__ void decoder_stage_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                                     const float* __restrict__ b2) {
    const float g1 = 1.0f / sqrtf((float)kFeat), g2 = 1.0f / sqrtf((float)kHidden);
    float* s_w1 = g_dec_stage;
    float* s_b1 = s_w1 + kHidden * kFeat;
    float* s_w2t = s_b1 + kHidden;
    float* s_b2 = s_w2t + kHidden * kOut;
    for (int i = threadIdx.x; i < kHidden * kFeat; i += blockDim.x) s_w1[i] = w1[i] * g1;
    for (int i = threadIdx.x; i < kHidden; i += blockDim.x) s_b1[i] = b1[i];
    for (int i = threadIdx.x; i < kHidden * kOut; i += blockDim.x) {
        int j = i / kOut, k = i % kOut;
        s_w2t[i] = w2[k * kHidden + j] * g2;
    }
    for (int i = threadIdx.x; i < kOut; i += blockDim.x) s_b2[i] = b2[i];
}

__device__ __forceinline__ float softplus_fast(float x) { return x > 20.f ? x : __logf(1.f + __expf(x)); }
__device__ __forceinline__ float softplus_acc(float x) { return x > 20.f ? x : log1pf(expf(x)); }

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

// One bilinear tap set of a plane for the 8-lane group: accumulate 4 channels (sub*4..sub*4+3).
__device__ __forceinline__ void plane_gather(const float* __restrict__ plane_base, int64_t px_ld, int PH, int PW, float gx, float gy,
                                             int sub, float acc[4]) {
    const float ix = ((gx + 1.f) * PW - 1.f) / 2.f;
    const float iy = ((gy + 1.f) * PH - 1.f) / 2.f;
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    const float wnw = ((float)x1 - ix) * ((float)y1 - iy);
    const float wne = (ix - (float)x0) * ((float)y1 - iy);
    const float wsw = ((float)x1 - ix) * (iy - (float)y0);
    const float wse = (ix - (float)x0) * (iy - (float)y0);
    const bool vx0 = x0 >= 0 && x0 < PW, vx1 = x1 >= 0 && x1 < PW, vy0 = y0 >= 0 && y0 < PH, vy1 = y1 >= 0 && y1 < PH;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const float* b = plane_base + sub * 4;
    if (vy0 && vx0) { float4 t = __ldg(reinterpret_cast<const float4*>(b + ((int64_t)y0 * PW + x0) * px_ld)); s[0] += t.x * wnw; s[1] += t.y * wnw; s[2] += t.z * wnw; s[3] += t.w * wnw; }
    if (vy0 && vx1) { float4 t = __ldg(reinterpret_cast<const float4*>(b + ((int64_t)y0 * PW + x1) * px_ld)); s[0] += t.x * wne; s[1] += t.y * wne; s[2] += t.z * wne; s[3] += t.w * wne; }
    if (vy1 && vx0) { float4 t = __ldg(reinterpret_cast<const float4*>(b + ((int64_t)y1 * PW + x0) * px_ld)); s[0] += t.x * wsw; s[1] += t.y * wsw; s[2] += t.z * wsw; s[3] += t.w * wsw; }
    if (vy1 && vx1) { float4 t = __ldg(reinterpret_cast<const float4*>(b + ((int64_t)y1 * PW + x1) * px_ld)); s[0] += t.x * wse; s[1] += t.y * wse; s[2] += t.z * wse; s[3] += t.w * wse; }
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += s[k];
}

// Gather the mean tri-plane feature of samples [s0, s0+n) of this warp's ray into rows of `col`.
__device__ __forceinline__ void gather_pass(const ia_render_params& p, const float* __restrict__ planes_b, const Ray& r,
                                            const float* dep, float* col, int s0, int n, int lane) {
    const int grp = lane >> 3, sub = lane & 7;
    const float scale = 2.0f / p.box_warp;
    for (int s = grp; s < n; s += 4) {
        const float t = dep[s0 + s];
        const float qx = (r.ox + t * r.dx) * scale;
        const float qy = (r.oy + t * r.dy) * scale;
        const float qz = (r.oz + t * r.dz) * scale;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        plane_gather(planes_b + 0, p.plane_px_ld, p.PH, p.PW, qx, qy, sub, acc);    // plane 0: (x, y)
        plane_gather(planes_b + 32, p.plane_px_ld, p.PH, p.PW, qx, qz, sub, acc);   // plane 1: (x, z)
        plane_gather(planes_b + 64, p.plane_px_ld, p.PH, p.PW, qz, qx, sub, acc);   // plane 2: (z, x)
        float* o = col + (s0 + s) * kRowLd + sub * 4;
        o[0] = acc[0] / 3.0f; o[1] = acc[1] / 3.0f; o[2] = acc[2] / 3.0f; o[3] = acc[3] / 3.0f;
    }
}

// OSG decoder on samples [s0, s0+n): one sample per lane; features in, colours out (same rows), sigma to sig[].
__device__ __forceinline__ void mlp_pass(float* col, float* sig, int s0, int n, int lane) {
    for (int s = lane; s < n; s += 32) {
        float* row = col + (s0 + s) * kRowLd;
        float x[kFeat];
#pragma unroll
        for (int i = 0; i < kFeat; ++i) x[i] = row[i];
        float out[kOut];
#pragma unroll
        for (int k = 0; k < kOut; ++k) out[k] = c_b2[k];
#pragma unroll
        for (int j = 0; j < kHidden; ++j) {
            float h = c_b1[j];
#pragma unroll
            for (int i = 0; i < kFeat; ++i) h = fmaf(x[i], c_w1[j * kFeat + i], h);
            const float a = softplus_fast(h);
#pragma unroll
            for (int k = 0; k < kOut; ++k) out[k] = fmaf(a, c_w2t[j * kOut + k], out[k]);
        }
        sig[s0 + s] = out[0];
#pragma unroll
        for (int k = 0; k < kFeat; ++k) {
            const float sg = 1.0f / (1.0f + __expf(-out[1 + k]));
            row[k] = sg * (1.0f + 2.0f * 0.001f) - 0.001f;
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

__global__ void __launch_bounds__(kWarpsPerCta * 32) render_kernel(const ia_render_params p) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.Dc + p.Df;
    // per-warp scratch
    const int per_warp = S * kRowLd + 6 * S + 2 * (p.Dc + 2);
    float* base = smem + (size_t)warp * per_warp;
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
        const float* planes_b = p.planes + (int64_t)b * p.PH * p.PW * p.plane_px_ld;

        // ---- coarse depths (renderer.py:404-406) ----
        for (int s = lane; s < p.Dc; s += 32) {
            const float t = linspace_at(near, far, p.Dc, s) + p.jitter[ray * p.Dc + s] * delta_c;
            dep[s] = t;
            dmin = fminf(dmin, t); dmax = fmaxf(dmax, t);
        }
        __syncwarp();
        gather_pass(p, planes_b, r, dep, col, 0, p.Dc, lane);
        __syncwarp();
        mlp_pass(col, sig, 0, p.Dc, lane);
        __syncwarp();

        int n_all = p.Dc;
        if (p.Df > 0) {
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
            if (lane == 0) {
                float tot = 0.f;
                for (int k = 0; k < nb; ++k) tot += sd[k];
                float c = 0.f;
                cdf[0] = 0.f;
                for (int k = 0; k < nb; ++k) { c += sd[k] / tot; cdf[k + 1] = c; }
            }
            __syncwarp();
            for (int f = lane; f < p.Df; f += 32) {
                const float u = p.u ? p.u[ray * p.Df + f] : linspace_at(0.0f, 1.0f, p.Df, f);
                int inds = 0;   // searchsorted(cdf, u, right=True) over nb+1 entries
                for (int k = 0; k <= nb; ++k) inds += (cdf[k] <= u) ? 1 : 0;
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
            __syncwarp();
            gather_pass(p, planes_b, r, dep, col, p.Dc, p.Df, lane);
            __syncwarp();
            mlp_pass(col, sig, p.Dc, p.Df, lane);
            __syncwarp();
            // ---- merge: stable rank of every sample among all S (unify_samples, renderer.py:372-382) ----
            for (int s = lane; s < S; s += 32) {
                const float ds = dep[s];
                int rank = 0;
                for (int j = 0; j < S; ++j) {
                    const float dj = dep[j];
                    rank += (dj < ds || (dj == ds && j < s)) ? 1 : 0;
                }
                sd[rank] = ds; ssg[rank] = sig[s]; perm[rank] = s;
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
    IA_CHECK(p->Dc >= 4 && p->Dc <= 96 && p->Df >= 0 && p->Df <= 96, "ia_render: depth resolutions must be in [4,96] / [0,96]");
    IA_CHECK((p->plane_px_ld & 3) == 0 && p->plane_px_ld >= 96, "ia_render: planes need >= 96 channels, pixel stride multiple of 4");
    IA_CHECK(p->res > 0 && p->B > 0, "ia_render: empty batch");
    cudaStream_t st = as_stream(stream);
    ia::prof_begin("ia_render(decoder_stage)", st);
    decoder_stage_kernel<<<1, 256, 0, st>>>(p->w1, p->b1, p->w2, p->b2);
    IA_LAUNCH_CHECK("ia_render(decoder_stage)");
    void* stage = nullptr;
    cudaError_t e = cudaGetSymbolAddress(&stage, g_dec_stage);
    IA_CHECK(e == cudaSuccess, "ia_render: cudaGetSymbolAddress: %s", cudaGetErrorString(e));
    const float* sp = static_cast<const float*>(stage);
    e = cudaMemcpyToSymbolAsync(c_w1, sp, sizeof(float) * kHidden * kFeat, 0, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_b1, sp + kHidden * kFeat, sizeof(float) * kHidden, 0, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_w2t, sp + kHidden * kFeat + kHidden, sizeof(float) * kHidden * kOut, 0, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_b2, sp + kHidden * kFeat + kHidden + kHidden * kOut, sizeof(float) * kOut, 0, cudaMemcpyDeviceToDevice, st);
    IA_CHECK(e == cudaSuccess, "ia_render: constant upload: %s", cudaGetErrorString(e));
    ia::prof_begin("ia_render(minmax_init)", st);
    minmax_init_kernel<<<1, 32, 0, st>>>(p->depth_minmax);
    IA_LAUNCH_CHECK("ia_render(minmax_init)");

    const int S = p->Dc + p->Df;
    const size_t per_warp = (size_t)S * kRowLd + 6 * (size_t)S + 2 * (size_t)(p->Dc + 2);
    const size_t smem = per_warp * kWarpsPerCta * sizeof(float);
    IA_CHECK(smem <= 227 * 1024, "ia_render: shared memory request too large");
    e = cudaFuncSetAttribute(render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    IA_CHECK(e == cudaSuccess, "ia_render: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_kernel, kWarpsPerCta * 32, smem);
    if (per_sm < 1) per_sm = 1;
    int64_t total_rays = (int64_t)p->B * p->res * p->res;
    int64_t want = cdiv(total_rays, kWarpsPerCta);
    int64_t grid = (int64_t)sms * per_sm;
    if (grid > want) grid = want;
    ia::prof_begin("ia_render", st);
    render_kernel<<<(unsigned)grid, kWarpsPerCta * 32, smem, st>>>(*p);
    IA_LAUNCH_CHECK("ia_render");
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
