// Layout / normalisation / gating kernels of the inversion encoder (reference encoder_inversion/models/
// {helpers,e4e,unet_encoders,uvnet}.py).  The dense contractions of the encoder (3x3 / 1x1 convolutions of the IR-SE50
// trunks, DoubleConv, ConvGRU, SFT heads, GradualStyleBlocks) run on ia_conv_tc; everything between two convolutions is
// one of the HBM-bound passes below.  All of them read through ia_view (include/invertavatar_b200.h), so channel
// concatenation, PixelShuffle, stride-2 subsampling, NCHW inputs and batch broadcast never materialise a tensor.
#include "ia_common.cuh"

using namespace ia;

namespace {

__device__ __forceinline__ float view_at(const ia_view& v, int b, int y, int x, int c) {
    if (v.ps == 1) return v.p[(int64_t)b * v.s_img + (int64_t)y * v.s_row + (int64_t)x * v.s_pix + (int64_t)c * v.s_c];
    const int ps = v.ps;
    const int cc = c * ps * ps + (y % ps) * ps + (x % ps);
    return v.p[(int64_t)b * v.s_img + (int64_t)(y / ps) * v.s_row + (int64_t)(x / ps) * v.s_pix + (int64_t)cc * v.s_c];
}

// ---- per-channel sums (float64) --------------------------------------------------------------------------------
// block = (32 channels) x (8 pixel lanes); each block walks a contiguous chunk of pixels.
__global__ void __launch_bounds__(256) chan_stats_kernel(ia_view v, int B, int H, int W, int64_t pix_per_block, double* __restrict__ sums) {
    const int c = blockIdx.y * 32 + (threadIdx.x & 31);
    const int lane_p = threadIdx.x >> 5;
    const int64_t npix = (int64_t)B * H * W;
    const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
    const int64_t p1 = p0 + pix_per_block < npix ? p0 + pix_per_block : npix;
    double s = 0.0, q = 0.0;
    if (c < v.C) {
        for (int64_t p = p0 + lane_p; p < p1; p += 8) {
            const int x = (int)(p % W); const int64_t t = p / W;
            const int y = (int)(t % H); const int b = (int)(t / H);
            const double a = (double)view_at(v, b, y, x, c);
            s += a; q += a * a;
        }
    }
    __shared__ double sh[2][8][32];
    sh[0][lane_p][threadIdx.x & 31] = s;
    sh[1][lane_p][threadIdx.x & 31] = q;
    __syncthreads();
    if (lane_p == 0 && c < v.C) {
        for (int k = 1; k < 8; ++k) { s += sh[0][k][threadIdx.x & 31]; q += sh[1][k][threadIdx.x & 31]; }
        atomicAdd(&sums[c], s);
        atomicAdd(&sums[v.C + c], q);
    }
}

__global__ void bn_fold_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma, const float* __restrict__ beta,
                               float* __restrict__ rmean, float* __restrict__ rvar, int training, float momentum, float eps, int C,
                               float* __restrict__ scale, float* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float mean, var;
    if (training) {
        const double m = sums[c] / count;
        double vv = sums[C + c] / count - m * m;
        if (vv < 0.0) vv = 0.0;
        mean = (float)m; var = (float)vv;
        if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
        if (rvar) {
            const double unb = count > 1.0 ? vv * count / (count - 1.0) : vv;
            rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unb;
        }
    } else {
        mean = rmean[c]; var = rvar[c];
    }
    const float inv = 1.f / sqrtf(var + eps);
    const float g = gamma ? gamma[c] : 1.f;
    const float sc = g * inv;
    scale[c] = sc;
    shift[c] = (beta ? beta[c] : 0.f) - mean * sc;
}

// ---- train-mode BatchNorm in ONE launch: per-channel sums (as chan_stats_kernel) + the fold of the last CTA to arrive --------------
// `sums` [2C] float64 and `counter` are zero on entry and left zero; the last CTA (threadfence + ticket) computes mean / biased
// variance, updates the running statistics as torch does and writes (scale, shift).
__global__ void __launch_bounds__(256) bn_stats_fold_kernel(ia_view v, int B, int H, int W, int64_t pix_per_block, double* __restrict__ sums,
                                                            int* __restrict__ counter, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                                                            float momentum, float eps, float* __restrict__ scale, float* __restrict__ shift) {
    const int C = v.C;
    const int c = blockIdx.y * 32 + (threadIdx.x & 31);
    const int lane_p = threadIdx.x >> 5;
    const int64_t npix = (int64_t)B * H * W;
    const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
    const int64_t p1 = p0 + pix_per_block < npix ? p0 + pix_per_block : npix;
    double s = 0.0, q = 0.0;
    if (c < C) {
        for (int64_t p = p0 + lane_p; p < p1; p += 8) {
            const int x = (int)(p % W); const int64_t t = p / W;
            const int y = (int)(t % H); const int b = (int)(t / H);
            const double a = (double)view_at(v, b, y, x, c);
            s += a; q += a * a;
        }
    }
    __shared__ double sh[2][8][32];
    __shared__ int s_last;
    sh[0][lane_p][threadIdx.x & 31] = s;
    sh[1][lane_p][threadIdx.x & 31] = q;
    __syncthreads();
    if (lane_p == 0 && c < C) {
        for (int k = 1; k < 8; ++k) { s += sh[0][k][threadIdx.x & 31]; q += sh[1][k][threadIdx.x & 31]; }
        atomicAdd(&sums[c], s);
        atomicAdd(&sums[C + c], q);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(counter, 1) == (int)(gridDim.x * gridDim.y) - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const double count = (double)npix;
    for (int i = threadIdx.x; i < C; i += 256) {
        const double m = __ldcg(&sums[i]) / count;
        double vv = __ldcg(&sums[C + i]) / count - m * m;
        if (vv < 0.0) vv = 0.0;
        sums[i] = 0.0; sums[C + i] = 0.0;
        const float mean = (float)m, var = (float)vv;
        if (rmean) rmean[i] = (1.f - momentum) * rmean[i] + momentum * mean;
        if (rvar) {
            const double unb = count > 1.0 ? vv * count / (count - 1.0) : vv;
            rvar[i] = (1.f - momentum) * rvar[i] + momentum * (float)unb;
        }
        const float inv = 1.f / sqrtf(var + eps);
        const float sc = (gamma ? gamma[i] : 1.f) * inv;
        scale[i] = sc;
        shift[i] = (beta ? beta[i] : 0.f) - mean * sc;
    }
    if (threadIdx.x == 0) *counter = 0;
}

// ---- operand builder -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) enc_prep_kernel(ia_enc_prep_params p, int Ctot) {
    const int groups = (p.hi ? p.C_pad : Ctot + 3) >> 2;     // 4 channels per thread
    const int64_t total = (int64_t)p.B * p.H * p.W * groups;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i % groups);
    const int64_t pix = i / groups;
    const int x = (int)(pix % p.W); const int64_t t = pix / p.W;
    const int y = (int)(t % p.H); const int b = (int)(t / p.H);
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = g * 4 + k;
        float a = 0.f;
        if (c < Ctot) {
            int cc = c, s = 0;
            while (s < p.nsrc - 1 && cc >= p.src[s].C) { cc -= p.src[s].C; ++s; }
            a = view_at(p.src[s], b, y, x, cc);
            if (p.scale) a = fmaf(a, p.scale[c], p.shift ? p.shift[c] : 0.f);
            else if (p.shift) a += p.shift[c];
            if (p.slope) a = a >= 0.f ? a : a * p.slope[c];
            else if (p.lrelu != 1.f) a = a >= 0.f ? a : a * p.lrelu;
        }
        v[k] = a;
    }
    if (p.out32) {
        float* o = p.out32 + pix * Ctot + g * 4;
        for (int k = 0; k < 4; ++k) if (g * 4 + k < Ctot) o[k] = v[k];
    }
    if (p.hi && g * 4 < p.C_pad) {
        uint16_t h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) split_bf16(v[k], h[k], l[k]);
        *reinterpret_cast<uint2*>(p.hi + pix * p.C_pad + g * 4) = make_uint2(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16));
        *reinterpret_cast<uint2*>(p.lo + pix * p.C_pad + g * 4) = make_uint2(l[0] | ((uint32_t)l[1] << 16), l[2] | ((uint32_t)l[3] << 16));
    }
}

// ---- affine + activation + gate + residual ----------------------------------------------------------------------
__global__ void __launch_bounds__(256) enc_affine_kernel(ia_enc_affine_params p) {
    const int64_t total = (int64_t)p.B * p.H * p.W * p.C;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % p.C);
    const int64_t pix = i / p.C;
    const int x = (int)(pix % p.W); const int64_t t = pix / p.W;
    const int y = (int)(t % p.H); const int b = (int)(t / p.H);
    float a = view_at(p.x, b, y, x, c);
    if (p.scale) a = fmaf(a, p.scale[c], p.shift ? p.shift[c] : 0.f);
    else if (p.shift) a += p.shift[c];
    if (p.slope1) a = a >= 0.f ? a : a * p.slope1[c];
    else a = apply_act(a, p.act, p.alpha);
    if (p.slope2) a = a >= 0.f ? a : a * p.slope2[c];
    if (p.gate) a *= p.gate[(int64_t)b * p.C + c];
    if (p.res.p) {
        float r = view_at(p.res, b, y, x, c);
        if (p.res_scale) r = fmaf(r, p.res_scale[c], p.res_shift ? p.res_shift[c] : 0.f);
        else if (p.res_shift) r += p.res_shift[c];
        a += r;
    }
    p.y[pix * p.y_ld + c] = a;
}

// Same arithmetic, 4 channels per thread, plus the operand of the NEXT convolution: split(y * e_scale[c] + e_shift[c]) as
// [B][H][W][e_C_pad] bf16 hi/lo (zero padded) -- the BatchNorm affine that opens the next IR-SE unit, folded into the pass that
// closes this one (no separate ia_enc_prep launch, no re-read of y).
__global__ void __launch_bounds__(256) enc_affine_emit_kernel(ia_enc_affine_params p) {
    const int groups = p.e_C_pad >> 2;
    const int64_t total = (int64_t)p.B * p.H * p.W * groups;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i % groups);
    const int64_t pix = i / groups;
    const int x = (int)(pix % p.W); const int64_t t = pix / p.W;
    const int y = (int)(t % p.H); const int b = (int)(t / p.H);
    float e[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = g * 4 + k;
        e[k] = 0.f;
        if (c >= p.C) continue;
        float a = view_at(p.x, b, y, x, c);
        if (p.scale) a = fmaf(a, p.scale[c], p.shift ? p.shift[c] : 0.f);
        else if (p.shift) a += p.shift[c];
        if (p.slope1) a = a >= 0.f ? a : a * p.slope1[c];
        else a = apply_act(a, p.act, p.alpha);
        if (p.slope2) a = a >= 0.f ? a : a * p.slope2[c];
        if (p.gate) a *= p.gate[(int64_t)b * p.C + c];
        if (p.res.p) {
            float r = view_at(p.res, b, y, x, c);
            if (p.res_scale) r = fmaf(r, p.res_scale[c], p.res_shift ? p.res_shift[c] : 0.f);
            else if (p.res_shift) r += p.res_shift[c];
            a += r;
        }
        p.y[pix * p.y_ld + c] = a;
        e[k] = p.e_scale ? fmaf(a, p.e_scale[c], p.e_shift ? p.e_shift[c] : 0.f) : (p.e_shift ? a + p.e_shift[c] : a);
    }
    store_operand4(IA_OPFMT_BF16X3, p.e_hi + pix * p.e_C_pad + g * 4, p.e_lo + pix * p.e_C_pad + g * 4, e[0], e[1], e[2], e[3]);
}

// ---- global average pool of an affine'd view: grid (B, ceil(C/32), pixel chunks), block 32 x 8; partial sums are added
// atomically into `pooled` (zeroed by the host wrapper), a second tiny kernel turns the sums into scale*mean + shift ------
__global__ void __launch_bounds__(256) global_pool_kernel(ia_view v, int H, int W, int pix_per_block, float* __restrict__ pooled) {
    const int b = blockIdx.x;
    const int c = blockIdx.y * 32 + (threadIdx.x & 31);
    const int lane_p = threadIdx.x >> 5;
    const int npix = H * W;
    const int p0 = blockIdx.z * pix_per_block;
    const int p1 = min(p0 + pix_per_block, npix);
    float s = 0.f;
    if (c < v.C) {
        for (int p = p0 + lane_p; p < p1; p += 8) s += view_at(v, b, p / W, p % W, c);
    }
    __shared__ float sh[8][32];
    sh[lane_p][threadIdx.x & 31] = s;
    __syncthreads();
    if (lane_p == 0 && c < v.C) {
        for (int k = 1; k < 8; ++k) s += sh[k][threadIdx.x & 31];
        atomicAdd(&pooled[(int64_t)b * v.C + c], s);
    }
}

__global__ void global_pool_finish_kernel(float* __restrict__ pooled, const float* __restrict__ scale, const float* __restrict__ shift, int B, int C,
                                          float inv_npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    const int c = i % C;
    float m = pooled[i] * inv_npix;
    if (scale) m = fmaf(m, scale[c], shift ? shift[c] : 0.f);
    pooled[i] = m;
}

// ---- squeeze-excite gate in ONE launch (helpers.py:62-80): gate[b][c] = sigmoid(fc2(relu(fc1(mean_hw(x*scale + shift))))) ---------
// grid (B, ceil(C/32), pixel chunks) as global_pool_kernel: partial channel sums are added atomically into `sums` [B][C]; the last
// CTA of an image to arrive on the image's ticket (threadfence + atomic counter) turns the sums into the pooled vector, evaluates
// the two tiny bias-free layers from shared memory and leaves `sums` / the counter zeroed for the next launch on the same stream.
__global__ void __launch_bounds__(256) se_gate_kernel(ia_view v, const float* __restrict__ scale, const float* __restrict__ shift, int H, int W,
                                                      int pix_per_block, float* __restrict__ sums, int* __restrict__ counters,
                                                      const float* __restrict__ w1, const float* __restrict__ w2, int Cr,
                                                      float* __restrict__ gate) {
    const int b = blockIdx.x;
    const int C = v.C;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + lane;
    const int npix = H * W;
    const int p0 = blockIdx.z * pix_per_block;
    const int p1 = min(p0 + pix_per_block, npix);
    float s = 0.f;
    if (c < C) {
        for (int p = p0 + wrp; p < p1; p += 8) s += view_at(v, b, p / W, p % W, c);
    }
    __shared__ float sh[8][32];
    __shared__ int s_last;
    sh[wrp][lane] = s;
    __syncthreads();
    if (wrp == 0 && c < C) {
        for (int k = 1; k < 8; ++k) s += sh[k][lane];
        atomicAdd(&sums[(int64_t)b * C + c], s);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&counters[b], 1) == (int)(gridDim.y * gridDim.z) - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    extern __shared__ float dyn[];          // pooled [C] | hidden [Cr]
    float* pooled = dyn;
    float* hid = dyn + C;
    const float inv = 1.f / (float)npix;
    for (int i = threadIdx.x; i < C; i += 256) {
        float m = __ldcg(&sums[(int64_t)b * C + i]) * inv;
        if (scale) m = fmaf(m, scale[i], shift ? shift[i] : 0.f);
        pooled[i] = m;
        sums[(int64_t)b * C + i] = 0.f;
    }
    if (threadIdx.x == 0) counters[b] = 0;
    __syncthreads();
    for (int o = wrp; o < Cr; o += 8) {     // fc1 + ReLU: one warp per output
        float a = 0.f;
        for (int i = lane; i < C; i += 32) a = fmaf(pooled[i], w1[(int64_t)o * C + i], a);
#pragma unroll
        for (int k = 16; k; k >>= 1) a += __shfl_xor_sync(0xffffffffu, a, k);
        if (lane == 0) hid[o] = a > 0.f ? a : 0.f;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < C; o += 256) {     // fc2 + sigmoid
        float a = 0.f;
        for (int i = 0; i < Cr; ++i) a = fmaf(hid[i], w2[(int64_t)o * Cr + i], a);
        gate[(int64_t)b * C + o] = 1.f / (1.f + expf(-a));
    }
}

__global__ void __launch_bounds__(256) avgpool_kernel(ia_view v, int B, int OH, int OW, int k, float* __restrict__ y) {
    const int64_t total = (int64_t)B * OH * OW * v.C;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % v.C);
    const int64_t pix = i / v.C;
    const int x = (int)(pix % OW); const int64_t t = pix / OW;
    const int yy = (int)(t % OH); const int b = (int)(t / OH);
    float s = 0.f;
    for (int dy = 0; dy < k; ++dy)
        for (int dx = 0; dx < k; ++dx) s += view_at(v, b, yy * k + dy, x * k + dx, c);
    y[i] = s / (float)(k * k);
}

__global__ void __launch_bounds__(256) upsample_add_kernel(const float* __restrict__ x, int B, int h, int w, int C,
                                                           const float* __restrict__ lat, int H, int W, float* __restrict__ y) {
    const int64_t total = (int64_t)B * H * W * C;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    const int64_t pix = i / C;
    const int ox = (int)(pix % W); const int64_t t = pix / W;
    const int oy = (int)(t % H); const int b = (int)(t / H);
    // align_corners=True: src = dst * (in-1)/(out-1)   (ATen area_pixel_compute_source_index)
    const float ry = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
    const float rx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const float fy = ry * oy, fx = rx * ox;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0;
    const float* xb = x + (int64_t)b * h * w * C + c;
    const float v00 = xb[((int64_t)y0 * w + x0) * C], v01 = xb[((int64_t)y0 * w + x1) * C];
    const float v10 = xb[((int64_t)y1 * w + x0) * C], v11 = xb[((int64_t)y1 * w + x1) * C];
    const float up = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
    y[i] = up + lat[i];
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256) gru_gate0_kernel(const float* __restrict__ raw, const float* __restrict__ bias, const float* __restrict__ h,
                                                        float* __restrict__ rh, float* __restrict__ z, int64_t n, int C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * C) return;
    const int c = (int)(i % C);
    const int64_t pix = i / C;
    const float r = sigmoidf_(raw[pix * 2 * C + c] + bias[c]);
    const float zz = sigmoidf_(raw[pix * 2 * C + C + c] + bias[C + c]);
    rh[i] = h ? r * h[i] : 0.f;
    z[i] = zz;
}

__global__ void __launch_bounds__(256) gru_gate1_kernel(const float* __restrict__ raw, const float* __restrict__ bias, const float* __restrict__ h,
                                                        const float* __restrict__ z, float* __restrict__ h_out, int64_t n, int C) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * C) return;
    const int c = (int)(i % C);
    const float cand = tanhf(raw[i] + bias[c]);
    const float zz = z[i];
    const float hv = h ? h[i] : 0.f;
    h_out[i] = (1.f - zz) * hv + zz * cand;
}

__global__ void __launch_bounds__(256) sft_half_kernel(float* __restrict__ x, int64_t x_ld, ia_view sc, ia_view sh, int B, int H, int W, int C) {
    const int half = C >> 1;
    const int64_t total = (int64_t)B * H * W * half;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % half);
    const int64_t pix = i / half;
    const int xx = (int)(pix % W); const int64_t t = pix / W;
    const int y = (int)(t % H); const int b = (int)(t / H);
    float* px = x + pix * x_ld + half + c;
    *px = fmaf(*px, view_at(sc, b, y, xx, c), view_at(sh, b, y, xx, c));
}

int check_view(const ia_view* v, const char* who) {
    IA_CHECK(v && v->p, "%s: null view", who);
    IA_CHECK(v->C > 0 && v->ps >= 1, "%s: bad view (C=%d ps=%d)", who, v ? v->C : 0, v ? v->ps : 0);
    return 0;
}

}  // namespace

extern "C" int ia_enc_chan_stats(const ia_view* x, int32_t B, int32_t H, int32_t W, double* sums, void* stream) {
    if (int rc = check_view(x, "ia_enc_chan_stats")) return rc;
    IA_CHECK(sums && B > 0 && H > 0 && W > 0, "ia_enc_chan_stats: bad arguments");
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * x->C, as_stream(stream));
    IA_CHECK(e == cudaSuccess, "ia_enc_chan_stats: memset: %s", cudaGetErrorString(e));
    const int64_t npix = (int64_t)B * H * W;
    const int cblocks = (int)cdiv(x->C, 32);
    int64_t want_blocks = cdiv(148 * 4, cblocks);
    int64_t ppb = cdiv(npix, want_blocks);
    if (ppb < 64) ppb = 64;
    dim3 grid((unsigned)cdiv(npix, ppb), (unsigned)cblocks);
    ia::prof_begin("ia_enc_chan_stats", as_stream(stream));
    chan_stats_kernel<<<grid, 256, 0, as_stream(stream)>>>(*x, B, H, W, ppb, sums);
    IA_LAUNCH_CHECK("ia_enc_chan_stats");
    return 0;
}

extern "C" int ia_enc_bn_fold(const double* sums, int64_t count, const float* gamma, const float* beta, float* running_mean,
                              float* running_var, int32_t training, float momentum, float eps, int32_t C, float* scale, float* shift,
                              void* stream) {
    IA_CHECK(scale && shift && C > 0, "ia_enc_bn_fold: null output");
    IA_CHECK(training ? (sums != nullptr && count > 0) : (running_mean && running_var), "ia_enc_bn_fold: %s",
             training ? "training mode needs sums and count" : "eval mode needs running statistics");
    ia::prof_begin("ia_enc_bn_fold", as_stream(stream));
    bn_fold_kernel<<<(unsigned)cdiv(C, 128), 128, 0, as_stream(stream)>>>(sums, (double)count, gamma, beta, running_mean, running_var,
                                                                          training, momentum, eps, C, scale, shift);
    IA_LAUNCH_CHECK("ia_enc_bn_fold");
    return 0;
}

extern "C" int ia_enc_bn_stats_fold(const ia_view* x, int32_t B, int32_t H, int32_t W, double* sums, int32_t* counter, const float* gamma,
                                    const float* beta, float* running_mean, float* running_var, float momentum, float eps, float* scale,
                                    float* shift, void* stream) {
    if (int rc = check_view(x, "ia_enc_bn_stats_fold")) return rc;
    IA_CHECK(sums && counter && scale && shift && B > 0 && H > 0 && W > 0, "ia_enc_bn_stats_fold: bad arguments");
    const int64_t npix = (int64_t)B * H * W;
    const int cblocks = (int)cdiv(x->C, 32);
    int64_t want_blocks = cdiv(148 * 4, cblocks);
    int64_t ppb = cdiv(npix, want_blocks);
    if (ppb < 64) ppb = 64;
    dim3 grid((unsigned)cdiv(npix, ppb), (unsigned)cblocks);
    ia::prof_begin("ia_enc_bn_stats_fold", as_stream(stream));
    bn_stats_fold_kernel<<<grid, 256, 0, as_stream(stream)>>>(*x, B, H, W, ppb, sums, counter, gamma, beta, running_mean, running_var, momentum, eps,
                                                             scale, shift);
    IA_LAUNCH_CHECK("ia_enc_bn_stats_fold");
    return 0;
}

extern "C" int ia_enc_prep(const ia_enc_prep_params* p, void* stream) {
    IA_CHECK(p && p->nsrc >= 1 && p->nsrc <= 4, "ia_enc_prep: 1..4 sources");
    int Ctot = 0;
    for (int s = 0; s < p->nsrc; ++s) {
        if (int rc = check_view(&p->src[s], "ia_enc_prep")) return rc;
        Ctot += p->src[s].C;
    }
    IA_CHECK((p->hi != nullptr) == (p->lo != nullptr) && (p->hi || p->out32), "ia_enc_prep: no output");
    IA_CHECK(!p->hi || (p->C_pad >= Ctot && (p->C_pad & 3) == 0), "ia_enc_prep: C_pad (%d) must be >= sum of source channels (%d) and a multiple of 4",
             p->C_pad, Ctot);
    const int groups = (p->hi ? p->C_pad : Ctot + 3) >> 2;
    const int64_t total = (int64_t)p->B * p->H * p->W * groups;
    if (total == 0) return 0;
    ia::prof_begin("ia_enc_prep", as_stream(stream));
    enc_prep_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(*p, Ctot);
    IA_LAUNCH_CHECK("ia_enc_prep");
    return 0;
}

extern "C" int ia_enc_affine_act(const ia_enc_affine_params* p, void* stream) {
    IA_CHECK(p && p->y, "ia_enc_affine_act: null output");
    if (int rc = check_view(&p->x, "ia_enc_affine_act")) return rc;
    IA_CHECK(p->x.C >= p->C && p->y_ld >= p->C, "ia_enc_affine_act: channel mismatch");
    IA_CHECK(p->res.p == nullptr || (p->res.C >= p->C && p->res.ps >= 1), "ia_enc_affine_act: bad residual view");
    const int64_t total = (int64_t)p->B * p->H * p->W * p->C;
    if (total == 0) return 0;
    if (p->e_hi) {
        IA_CHECK(p->e_lo && p->e_C_pad >= p->C && (p->e_C_pad & 3) == 0, "ia_enc_affine_act: emitted operand needs e_lo and e_C_pad >= C, multiple of 4");
        const int64_t tot4 = (int64_t)p->B * p->H * p->W * (p->e_C_pad >> 2);
        ia::prof_begin("ia_enc_affine_act", as_stream(stream));
        enc_affine_emit_kernel<<<(unsigned)cdiv(tot4, 256), 256, 0, as_stream(stream)>>>(*p);
        IA_LAUNCH_CHECK("ia_enc_affine_act");
        return 0;
    }
    ia::prof_begin("ia_enc_affine_act", as_stream(stream));
    enc_affine_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(*p);
    IA_LAUNCH_CHECK("ia_enc_affine_act");
    return 0;
}

extern "C" int ia_enc_global_pool(const ia_view* x, const float* scale, const float* shift, int32_t B, int32_t H, int32_t W,
                                  float* pooled, void* stream) {
    if (int rc = check_view(x, "ia_enc_global_pool")) return rc;
    IA_CHECK(pooled && B > 0 && H > 0 && W > 0, "ia_enc_global_pool: bad arguments");
    cudaError_t e = cudaMemsetAsync(pooled, 0, sizeof(float) * (size_t)B * x->C, as_stream(stream));
    IA_CHECK(e == cudaSuccess, "ia_enc_global_pool: memset: %s", cudaGetErrorString(e));
    const int npix = H * W;
    const int cblocks = (int)cdiv(x->C, 32);
    int chunks = (int)cdiv(148 * 4, (int64_t)B * cblocks);
    int ppb = (int)cdiv(npix, chunks);
    if (ppb < 64) ppb = 64;
    chunks = (int)cdiv(npix, ppb);
    dim3 grid((unsigned)B, (unsigned)cblocks, (unsigned)chunks);
    ia::prof_begin("ia_enc_global_pool", as_stream(stream));
    global_pool_kernel<<<grid, 256, 0, as_stream(stream)>>>(*x, H, W, ppb, pooled);
    IA_LAUNCH_CHECK("ia_enc_global_pool");
    ia::prof_begin("ia_enc_global_pool", as_stream(stream));
    global_pool_finish_kernel<<<(unsigned)cdiv((int64_t)B * x->C, 128), 128, 0, as_stream(stream)>>>(pooled, scale, shift, B, x->C, 1.0f / (float)npix);
    IA_LAUNCH_CHECK("ia_enc_global_pool");
    return 0;
}

extern "C" int ia_enc_se_gate(const ia_view* x, const float* scale, const float* shift, int32_t B, int32_t H, int32_t W, const float* w1,
                              const float* w2, int32_t Cr, float* sums, int32_t* counters, float* gate, void* stream) {
    if (int rc = check_view(x, "ia_enc_se_gate")) return rc;
    IA_CHECK(w1 && w2 && sums && counters && gate && B > 0 && H > 0 && W > 0 && Cr > 0, "ia_enc_se_gate: bad arguments");
    IA_CHECK((size_t)(x->C + Cr) * sizeof(float) <= 40 * 1024, "ia_enc_se_gate: C + Cr too large for the shared-memory vectors");
    const int npix = H * W;
    const int cblocks = (int)cdiv(x->C, 32);
    int chunks = (int)cdiv(148 * 4, (int64_t)B * cblocks);
    int ppb = (int)cdiv(npix, chunks);
    if (ppb < 64) ppb = 64;
    chunks = (int)cdiv(npix, ppb);
    dim3 grid((unsigned)B, (unsigned)cblocks, (unsigned)chunks);
    ia::prof_begin("ia_enc_se_gate", as_stream(stream));
    se_gate_kernel<<<grid, 256, (size_t)(x->C + Cr) * sizeof(float), as_stream(stream)>>>(*x, scale, shift, H, W, ppb, sums, counters, w1, w2, Cr, gate);
    IA_LAUNCH_CHECK("ia_enc_se_gate");
    return 0;
}

extern "C" int ia_enc_avgpool(const ia_view* x, int32_t B, int32_t H, int32_t W, int32_t k, float* y, void* stream) {
    if (int rc = check_view(x, "ia_enc_avgpool")) return rc;
    IA_CHECK(y && k >= 1 && H % k == 0 && W % k == 0, "ia_enc_avgpool: H, W must be multiples of k");
    const int64_t total = (int64_t)B * (H / k) * (W / k) * x->C;
    if (total == 0) return 0;
    ia::prof_begin("ia_enc_avgpool", as_stream(stream));
    avgpool_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(*x, B, H / k, W / k, k, y);
    IA_LAUNCH_CHECK("ia_enc_avgpool");
    return 0;
}

extern "C" int ia_enc_upsample_add(const float* x, int32_t B, int32_t h, int32_t w, int32_t C, const float* lateral, int32_t H,
                                   int32_t W, float* y, void* stream) {
    IA_CHECK(x && lateral && y && h > 0 && w > 0 && H > 0 && W > 0, "ia_enc_upsample_add: bad arguments");
    const int64_t total = (int64_t)B * H * W * C;
    if (total == 0) return 0;
    ia::prof_begin("ia_enc_upsample_add", as_stream(stream));
    upsample_add_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(x, B, h, w, C, lateral, H, W, y);
    IA_LAUNCH_CHECK("ia_enc_upsample_add");
    return 0;
}

extern "C" int ia_enc_gru_gate(int32_t stage, const float* raw, const float* bias, const float* h, float* rh, float* z,
                               float* h_out, int64_t n, int32_t C, void* stream) {
    IA_CHECK(raw && bias && z && n >= 0 && C > 0, "ia_enc_gru_gate: bad arguments");
    if (n == 0) return 0;
    const unsigned blocks = (unsigned)cdiv(n * C, 256);
    if (stage == 0) {
        IA_CHECK(rh, "ia_enc_gru_gate: stage 0 needs rh");
        ia::prof_begin("ia_enc_gru_gate", as_stream(stream));
        gru_gate0_kernel<<<blocks, 256, 0, as_stream(stream)>>>(raw, bias, h, rh, z, n, C);
    } else {
        IA_CHECK(h_out, "ia_enc_gru_gate: stage 1 needs h_out");
        ia::prof_begin("ia_enc_gru_gate", as_stream(stream));
        gru_gate1_kernel<<<blocks, 256, 0, as_stream(stream)>>>(raw, bias, h, z, h_out, n, C);
    }
    IA_LAUNCH_CHECK("ia_enc_gru_gate");
    return 0;
}

extern "C" int ia_sft_half(float* x, int64_t x_ld, const ia_view* scale, const ia_view* shift, int32_t B, int32_t H, int32_t W,
                           int32_t C, void* stream) {
    IA_CHECK(x && (C & 1) == 0 && x_ld >= C, "ia_sft_half: bad arguments");
    if (int rc = check_view(scale, "ia_sft_half")) return rc;
    if (int rc = check_view(shift, "ia_sft_half")) return rc;
    IA_CHECK(scale->C >= C / 2 && shift->C >= C / 2, "ia_sft_half: condition has too few channels");
    const int64_t total = (int64_t)B * H * W * (C / 2);
    if (total == 0) return 0;
    ia::prof_begin("ia_sft_half", as_stream(stream));
    sft_half_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(x, x_ld, *scale, *shift, B, H, W, C);
    IA_LAUNCH_CHECK("ia_sft_half");
    return 0;
}
