/*
 * invertavatar_b200 -- C-ABI of the B200-native generator-forward hot path.
 *
 * This is the drop-in boundary: a plain-C shared library (libinvertavatar_b200.so) with raw device
 * pointers, sizes and a CUDA stream handle.  No torch types appear here.  Every entry point cites the
 * reference interface it replaces (paths relative to the XChenZ/invertAvatar tree).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host
 *   - activations are fp32, channels-last ("NHWC": [B][H][W][C]) unless stated otherwise
 *   - "split" activations are a pair of bf16 tensors (hi, lo) with hi = rn_bf16(v), lo = rn_bf16(v - hi);
 *     the tensor-core convolution computes hi*hi + hi*lo + lo*hi with fp32 accumulation (3-term split)
 *   - stream is a cudaStream_t passed as void*; work is enqueued, never synchronised
 *   - return value: 0 on success, non-zero on error; ia_last_error() gives the message
 *     (the reference raises RuntimeError through TORCH_CHECK, e.g. torch_utils/ops/bias_act.cpp:39-55;
 *      the Python host layer turns a non-zero return into RuntimeError as well)
 *   - inputs are borrowed and never mutated; outputs are caller-allocated
 */
#ifndef INVERTAVATAR_B200_H_
#define INVERTAVATAR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IA_ABI_VERSION 3

/* ---- library management ------------------------------------------------------------------------ */
int ia_abi_version(void);
const char* ia_last_error(void);
/* Select the device for the calling host thread (the library carries its own CUDA runtime instance). */
int ia_set_device(int device);
/* Number of kernels this library has launched since load / since the last reset (bench.py gpu_launches). */
int64_t ia_launch_count(void);
void ia_reset_launch_count(void);
/* Per-launch timing for bench.py's roofline leg: between ia_profile_begin() and ia_profile_report() every kernel launch
 * of this library is bracketed by CUDA events on its launching stream.  ia_profile_report synchronises the device, stops
 * profiling and writes a JSON object {"<entry point>": {"ms": total, "launches": n}, ...} into buf (NUL-terminated,
 * truncated to buflen); returns the untruncated length, or -1 on error. */
int ia_profile_begin(void);
int64_t ia_profile_report(char* buf, int64_t buflen);

/* ---- torch_utils/ops plugin equivalents --------------------------------------------------------- */

/* Activation ids follow the reference's cuda_idx (torch_utils/ops/bias_act.py:23-33). */
enum { IA_ACT_LINEAR = 1, IA_ACT_RELU = 2, IA_ACT_LRELU = 3, IA_ACT_TANH = 4, IA_ACT_SIGMOID = 5,
       IA_ACT_ELU = 6, IA_ACT_SELU = 7, IA_ACT_SOFTPLUS = 8, IA_ACT_SWISH = 9,
       /* not a reference id: torch.nn.PReLU with per-channel slopes inside a convolution epilogue (ia_conv_params.slope) */
       IA_ACT_PRELU = 10 };

/* Element types of the three plugin-level entry points.  The reference dispatches bias_act / upfirdn2d over double, float
 * and half (AT_DISPATCH_FLOATING_TYPES_AND_HALF, bias_act.cpp:81, upfirdn2d.cpp:67) and filtered_lrelu over float / half
 * (filtered_lrelu.cpp:33); arithmetic is float for half/float and double for double (bias_act.cu:18-21).  Everything else
 * in this header is fp32. */
enum { IA_DTYPE_F32 = 0, IA_DTYPE_F16 = 1, IA_DTYPE_F64 = 2 };

/* y = clamp(act(x + b[c]) * gain).  Replaces bias_act_plugin.bias_act(x,b,xref,yref,dy,grad=0,dim,act,alpha,
 * gain,clamp) (torch_utils/ops/bias_act.cpp:36-94, kernel bias_act.cu:27-151), forward only.
 * x is a dense tensor viewed as [outer][C][inner]; channel of element i is (i / inner) % C.
 * x, b, y have element type `dtype` (bias_act.cpp:41: b.dtype == x.dtype).  b may be NULL (no bias).  clamp < 0 disables
 * clamping. */
int ia_bias_act(const void* x, const void* b, void* y, int64_t numel, int64_t C, int64_t inner,
                int act, float alpha, float gain, float clamp, int dtype, void* stream);

/* Generic upfirdn2d: zero-insert x(up), pad/crop, 2-D FIR, decimate.  Replaces
 * upfirdn2d_plugin.upfirdn2d(x,f,upx,upy,downx,downy,padx0,padx1,pady0,pady1,flip,gain)
 * (torch_utils/ops/upfirdn2d.cpp:20-102, kernels upfirdn2d.cu:33-204).  Arbitrary element strides for x and y
 * (covers contiguous and channels-last, like upfirdn2d.cpp:56-63); f is a dense fp32 [fh][fw] filter. */
typedef struct {
    const void* x; const float* f; void* y;      /* x, y: element type `dtype`; f: fp32 (upfirdn2d.cpp:26) */
    int32_t N, C, inH, inW, outH, outW, fh, fw;
    int32_t upx, upy, downx, downy, padx0, pady0;
    int32_t flip;              /* reference semantics: flip=0 -> true convolution (filter is flipped) */
    float gain;
    int64_t xs_n, xs_c, xs_h, xs_w;   /* element strides of x */
    int64_t ys_n, ys_c, ys_h, ys_w;   /* element strides of y */
    int32_t dtype;                    /* IA_DTYPE_* of x and y */
} ia_upfirdn2d_params;
int ia_upfirdn2d(const ia_upfirdn2d_params* p, void* stream);

/* filtered_lrelu: y = downFIR(clamp(lrelu(upFIR(x + b) * up^2) * gain)).  Replaces
 * filtered_lrelu_plugin.filtered_lrelu(x,fu,fd,b,si,up,down,px0,px1,py0,py1,sx,sy,gain,slope,clamp,flip_filters,writeSigns)
 * -> (y, so, rc) (torch_utils/ops/filtered_lrelu.cpp:20-214), forward only, as two kernels through a caller-provided fp32
 * workspace of ia_filtered_lrelu_workspace(p) bytes.  fu / fd: fp32 2-D [fh][fw], or separable 1-D [fw] with fh = 0
 * (filtered_lrelu.cpp:52-53), or NULL (identity).  x / y: element strides, element type `dtype` (F32 or F16); b [C] of the
 * same type or NULL.  outH / outW must equal the reference's output size (filtered_lrelu.cpp:77-86).
 * Return value: 0 = done; -1 = "no specialised kernel" exactly like the reference's rc (filtered_lrelu.cpp:56-60): given when
 * sign tensors are requested (si / so / write_signs: they only serve the backward pass) -- the caller then takes the
 * bias_act / upfirdn2d composition as filtered_lrelu.py:225-231 does; > 0 = invalid arguments (ia_last_error). */
typedef struct {
    const void* x; void* y; const void* b;
    const float* fu; const float* fd;
    int32_t N, C, inH, inW, outH, outW;
    int32_t fuw, fuh, fdw, fdh;
    int32_t up, down, px0, px1, py0, py1;
    float gain, slope, clamp;         /* clamp < 0: none */
    int32_t flip;                     /* flip_filters */
    int64_t xs_n, xs_c, xs_h, xs_w;
    int64_t ys_n, ys_c, ys_h, ys_w;
    int32_t dtype;
    void* workspace; int64_t workspace_bytes;
    const uint8_t* si; int32_t sx, sy; int32_t write_signs; uint8_t* so;
} ia_filtered_lrelu_params;
int64_t ia_filtered_lrelu_workspace(const ia_filtered_lrelu_params* p);   /* bytes, or -1 (ia_last_error) */
int ia_filtered_lrelu(const ia_filtered_lrelu_params* p, void* stream);
/* In place x = clamp(lrelu(x) * gain).  Replaces filtered_lrelu_plugin.filtered_lrelu_act_(x,si,sx,sy,gain,slope,clamp,
 * writeSigns) (filtered_lrelu.cpp:217-296), the one entry point that mutates its input.  Sign tensors: -1 as above. */
int ia_filtered_lrelu_act(void* x, int64_t numel, int dtype, const uint8_t* si, int32_t sx, int32_t sy, float gain, float slope,
                          float clamp, int32_t write_signs, uint8_t* so, void* stream);

/* ---- small dense layers (MappingNetwork, style affines; networks_stylegan2_new.py:96-127,233-268) ------- */

/* y[b][o] = act((sum_i x[b][i] * w[o][i]) * w_gain + bias[o] * b_gain) * act_gain ; w is [Out][In] row-major. */
int ia_fully_connected(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t In,
                       int32_t Out, float w_gain, float b_gain, int act, float alpha, float act_gain,
                       int64_t x_stride, int64_t y_stride, void* stream);
/* y = x * rsqrt(mean(x^2, dim=1) + eps)  (normalize_2nd_moment, networks_stylegan2_new.py:28-29) */
int ia_normalize_2nd_moment(const float* x, float* y, int32_t B, int32_t D, float eps, int64_t x_stride,
                            int64_t y_stride, void* stream);
/* ws[b][k][:] = k < cutoff ? lerp(w_avg, w[b], psi) : w[b]   (broadcast + truncation, :256-267) */
int ia_broadcast_truncate(const float* w, const float* w_avg, float* ws, int32_t B, int32_t num_ws, int32_t D,
                          float psi, int32_t cutoff, void* stream);

/* ---- modulated convolution stack (networks_stylegan2_new.py:34-91,311-357; conv2d_resample.py:48-143) --- */

/* One entry per SynthesisLayer / ToRGBLayer of a network; the table lives in device memory. */
typedef struct {
    const float* affine_w;   /* [Cin][w_dim] */
    const float* affine_b;   /* [Cin] */
    const float* wsq;        /* [Cout][Cin]  sum over taps of weight^2, or NULL (no demodulation: ToRGB) */
    float* styles;           /* out [B][Cin] */
    float* dcoef;            /* out [B][Cout] or NULL */
    int32_t Cin, Cout, w_index, w_dim;
    float affine_gain;       /* 1/sqrt(w_dim) */
    float style_gain;        /* 1 for conv layers, 1/sqrt(Cin) for ToRGB (:354) */
} ia_style_layer;
/* styles[l] = (affine_w[l] * ws[:, w_index[l]] * affine_gain + affine_b[l]) * style_gain ;
 * dcoef[l][b][o] = rsqrt(sum_i styles[b][i]^2 * wsq[o][i] + 1e-8)   (:63-65, reassociated) */
int ia_styles(const ia_style_layer* layers_dev, const ia_style_layer* layers_host, int32_t n_layers,
              const float* ws, int32_t B, int32_t num_ws, void* stream);

/* Operand formats of the tensor-core convolution (both operands of a launch share one).
 *   IA_OPFMT_BF16X3: a pair of bf16 tensors (hi = rn(v), lo = rn(v - hi)); hi*hi + hi*lo + lo*hi, fp32 accumulate: fp32-grade
 *                    products (2^-16 relative), three MMAs per k-step.
 *   IA_OPFMT_F16X1 : ONE fp16 tensor (rn(v), saturated to +-65504); a single MMA per k-step, 2^-11 relative per operand.  Chosen
 *                    per layer by the host from a measured error budget (profiles/r2_conv_precision_probe.json): the three
 *                    backbones tolerate it (final image 1e-4..3e-4 max-abs vs fp32), the super-resolution blocks do not.
 * The `lo` pointers of an F16X1 operand are ignored (may be NULL). */
enum { IA_OPFMT_BF16X3 = 0, IA_OPFMT_F16X1 = 1 };

/* Prepare the tensor-core A operand: v = x * styles[b][c] (optionally after x = cond*a + x*(1-a), the
 * cond_list blend of networks_stylegan2_new.py:538-540), then split into bf16 hi/lo, zero-padded to C_pad. */
typedef struct {
    const float* x; int64_t x_ld;            /* [B][HW][C], pixel stride x_ld */
    const float* styles;                      /* [B][C] or NULL (=1) */
    const float* cond; int64_t cond_ld;       /* [B][HW][C] or NULL */
    const float* cond_alpha;                  /* [B][HW] */
    uint16_t* hi; uint16_t* lo;               /* bf16 [B][HW][C_pad] */
    int32_t B, HW, C, C_pad;
    int64_t out_img_pix;                      /* pixel stride between images of hi / lo (0: HW, dense) */
    int32_t fmt;                              /* IA_OPFMT_* of the produced operand */
} ia_modsplit_params;
int ia_modsplit(const ia_modsplit_params* p, void* stream);

/* Pack an OIHW fp32 weight into the GEMM layout [tap][Cout_pad][Cin_pad] in operand format `fmt` (bf16 hi/lo, or fp16 in
 * w_hi alone) (+ wsq[Cout][Cin], always from the fp32 weight). */
int ia_pack_conv_weight(const float* w, int32_t Cout, int32_t Cin, int32_t kh, int32_t kw, int32_t Cout_pad,
                        int32_t Cin_pad, uint16_t* w_hi, uint16_t* w_lo, float* wsq, int32_t fmt, void* stream);

typedef struct {
    /* what to emit from the fp32 result v[b][oy][ox][co] (any subset) */
    float* out32; int64_t out32_ld;                       /* fp32 NHWC */
    uint16_t* hi1; uint16_t* lo1; const float* s1; int32_t c1_pad;   /* split of v*s1[b][co] (next conv) */
    uint16_t* hi2; uint16_t* lo2; const float* s2; int32_t c2_pad;   /* split of v*s2[b][co] (ToRGB) */
    /* Fused ToRGB contraction (networks_stylegan2_new.py:354-356: 1x1 modulated convolution without demodulation, few output
     * channels): rgb_out[b][pixel][j] += sum_co (v*rgb_s[b][co]) * rgb_w[j][co], j < rgb_n <= 4, accumulated in fp32 with
     * atomic adds (one per N tile of the launch) into a buffer the caller zero-filled.  ia_conv_tc only, persistent kernel,
     * mode 1, Cout % 4 == 0; NULL: not used. */
    float* rgb_out; const float* rgb_w; const float* rgb_s; int32_t rgb_n;
    /* Pixel stride between consecutive images of the emit-1 tensors (0: OH*OW, dense).  A consumer that is a stride-2 transposed
     * convolution takes its operand as [B][H+1][W][C] with a zero row after every image (ia_conv_params.a_img_rows). */
    int64_t e1_img_pix;
    int32_t fmt1, fmt2;     /* IA_OPFMT_* of the emit-1 / emit-2 operands (the format their consuming convolutions run in) */
} ia_emit;

typedef struct {
    /* A operand: split activations [B][H][W][Cin_pad]; B operand: packed weights */
    const uint16_t* a_hi; const uint16_t* a_lo; int32_t B, H, W, Cin_pad;
    const uint16_t* w_hi; const uint16_t* w_lo; int32_t Cout, Cout_pad, n_taps_total;
    /* tile grid and taps: output grid position (gy,gx) accumulates sum_t A[gy+dy_t][gx+dx_t] * W[wtap_t] */
    int32_t GH, GW, ntaps; int32_t dy[9]; int32_t dx[9]; int32_t wtap[9];
    /* output pixel = (gy*sy+py, gx*sx+px) of an [B][OH][OW] image */
    int32_t OH, OW, sy, sx, py, px;
    /* epilogue: mode 0 = raw accumulator; mode 1 = v = acc*dcoef + noise*strength; v = act(v+bias)*gain, clamp;
     * mode 2 = ToRGB tail (see img_prev below) */
    int32_t mode; const float* dcoef; const float* noise; const float* noise_strength; const float* bias;
    int64_t noise_bstride;   /* 0: one [OH][OW] noise image shared by the batch ('const'); OH*OW: per-sample ('random') */
    int32_t act; float alpha; float gain; float clamp;
    ia_emit emit;
    /* Grouped launch (several networks with identical layer shapes evaluated as one batch, e.g. the low-resolution blocks of
     * the three backbones): image b belongs to group g = b / imgs_per_group and uses weight taps [g*n_taps_total, ...),
     * bias[g*Cout + co], noise_strength[g] and noise + g*noise_gstride.  groups <= 1: a single set (fields may be 0). */
    int32_t groups; int32_t imgs_per_group; int64_t noise_gstride;
    /* mode 2 (ia_conv_tc, persistent kernel, emit.out32 only): ToRGB tail fused into the 1x1 convolution --
     * out = upsample2d(img_prev) + clamp(acc + bias)  (networks_stylegan2_new.py:354-363,456-463; the arithmetic of
     * ia_torgb_finish, bit for bit).  img_prev: [B][OH/2][OW/2][Cout] fp32 NHWC or NULL. */
    const float* img_prev;
    /* Rows per image of the A tensors (0: H, dense).  H+1 = every image is followed by one all-zero row: ia_conv_tc_phases then
     * tiles the [B*(H+1)] x W concatenation of the images instead of every (H+1)-row phase grid on its own (a tile may span two
     * images; the zero row is the padding between them) -- fewer, fuller tiles.  mode 0, groups <= 1 only. */
    int32_t a_img_rows;
    /* Split-K scratch (ia_conv_tc / ia_conv_tc_phases, persistent kernel; all NULL / 0: never split).  A launch with fewer
     * output tiles than SMs splits the input channels over up to 16 CTAs per tile: partial accumulators go through
     * splitk_ws (fp32, any contents), the last CTA to arrive on a tile's ticket in splitk_counters sums them in split order
     * (deterministic) and runs the epilogue.  splitk_counters: splitk_n_counters int32, ZERO before the first launch that
     * uses them -- every launch leaves them zero again.  Both buffers belong to ONE stream at a time (launches on the same
     * stream may share them; concurrent streams need their own).  The split factor is clipped to what the buffers hold
     * (per split: tiles x 256 x N-tile x 4 bytes; counters: 8 per tile). */
    float* splitk_ws; int64_t splitk_ws_bytes; int32_t* splitk_counters; int32_t splitk_n_counters;
    int32_t op_fmt;          /* IA_OPFMT_* of a_* and w_* (a_lo / w_lo unused for F16X1) */
    /* act == IA_ACT_PRELU (mode 1): v = prelu(v + bias; slope[co]) -- the PReLU that follows a convolution of the inversion encoder
     * (helpers.py:111, unet_encoders.py:58-62), so that the layer emits the next convolution's operand (emit.hi1/lo1) without an fp32
     * round trip.  slope: [Cout]; emit 2 must be unused. */
    const float* slope;
} ia_conv_params;
/* Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA-fed). */
int ia_conv_tc(const ia_conv_params* p, void* stream);
/* n (<= 4) launches that share operands, epilogue and emitted tensors and differ only in taps / grid / output parity -- the four
 * phases of the stride-2 transposed convolution of conv2d_resample.py:114-127 -- executed as ONE persistent launch (the phases
 * share the ramp/tail and, tiles being interleaved phase-minor, the activation tile in L2).  Falls back to n ia_conv_tc calls
 * when a sub-problem is too small for the persistent kernel. */
int ia_conv_tc_phases(const ia_conv_params* p, int32_t n, void* stream);
/* Same contract on CUDA cores (fp32 FMA over hi+lo); cross-check and bring-up path. */
int ia_conv_simt(const ia_conv_params* p, void* stream);

/* Post-processing of an up=2 layer: 4x4 FIR (gain 4, pad 1) over the raw (2H+1)x(2W+1) transposed-conv
 * output, then demod/noise/bias/activation exactly as mode 1 above (conv2d_resample.py:127-128 + bias_act). */
typedef struct {
    const float* raw; int32_t B, RH, RW, C;     /* raw [B][RH][RW][C], RH = 2H+1 */
    const float* fir;                            /* [4][4] filter already multiplied by the gain */
    int32_t OH, OW;
    const float* dcoef; const float* noise; const float* noise_strength; const float* bias;
    int64_t noise_bstride;
    int32_t act; float alpha; float gain; float clamp;
    ia_emit emit;
    int32_t groups; int32_t imgs_per_group; int64_t noise_gstride;   /* as in ia_conv_params */
} ia_fir_params;
int ia_fir_epilogue(const ia_fir_params* p, void* stream);

/* ToRGB tail: img_out = upsample2d(img_prev) + clamp(raw + bias)  (networks_stylegan2_new.py:456-463,
 * upfirdn2d.upsample2d :315-350 with the [1,3,3,1] filter).  img_prev may be NULL.  out_nchw=1 writes planar. */
typedef struct {
    const float* raw; int64_t raw_ld;           /* [B][H][W][>=C] */
    const float* bias; float clamp;
    const float* img_prev;                       /* [B][H/2][W/2][C] or NULL */
    float* img_out; int32_t B, H, W, C; int32_t out_nchw;
    int32_t groups; int32_t imgs_per_group;      /* bias[g*C + c] with g = b / imgs_per_group (groups <= 1: one bias) */
    /* Fused image gather (the one exchange step of the path).  With out_nchw=1 every value is also stored at
     * element offset peer_offset + (its index in img_out) of the gathered [world*B][C][H][W] buffer of every rank: either one
     * multimem.st to mc_out (NVSwitch multicast address of the symmetric buffer) or n_peers plain stores to the peer-mapped
     * pointers peer_out[0..n_peers).  All NULL / 0: no gather. */
    float* peer_out[8]; int32_t n_peers; int64_t peer_offset; float* mc_out;
} ia_torgb_params;
int ia_torgb_finish(const ia_torgb_params* p, void* stream);

/* ---- UV rasterize / stitch (triplane_v20.py:119-128,317-339; renderer.py:716-741) ------------------------ */

/* GPU replacement of cv2.floodFill (seed (0,0), FIXED_RANGE lo 0 / up 254, 4-connectivity) + mask algebra:
 * mouth = (255 - filled(alpha*255))/255 ; full_alpha = clip(alpha+mouth,0,1) ;
 * upper_alpha = clip(alpha + mouth[rows >= upper_row0], 0, 1).  alpha is read with element stride a_stride
 * (the mask channel of uvcoords_image [B][H][W][3] has stride 3).  H,W <= 256. */
int ia_fill_mouth(const float* alpha, int64_t a_stride, int64_t a_batch_stride, int32_t B, int32_t H, int32_t W,
                  int32_t upper_row0, float* full_alpha, float* mouth, float* upper_alpha, void* stream);

/* F.grid_sample(bilinear, zeros, align_corners=False) on NHWC input; grid [B][Ho][Wo][>=2] with pixel stride g_ld. */
int ia_grid_sample(const float* in, int32_t B, int32_t Hi, int32_t Wi, int32_t C, int64_t in_ld,
                   const float* grid, int64_t g_ld, int32_t Ho, int32_t Wo, float* out, int64_t out_ld, void* stream);

/* F.interpolate(bilinear, antialias=True) (ATen _upsample_bilinear2d_aa) on a window of an NHWC tensor,
 * written into a window of another NHWC tensor.  Tap tables are built by the host (one per axis):
 * for output index i, taps j in [start[i], start[i]+count[i]) with weights w[i*max_taps + k]. */
typedef struct {
    const float* in; int64_t in_ld; int32_t in_H, in_W;      /* full input image dims (per batch) */
    int32_t in_y0, in_x0;                                      /* crop origin */
    float* out; int64_t out_ld; int32_t out_H, out_W;         /* full output image dims (per batch) */
    int32_t out_y0, out_x0;                                    /* paste origin */
    int32_t B, C, oh, ow;                                      /* resized window size */
    const int32_t* y_start; const int32_t* y_count; const float* y_w; int32_t y_max_taps;
    const int32_t* x_start; const int32_t* x_count; const float* x_w; int32_t x_max_taps;
} ia_resize_params;
int ia_resize_aa(const ia_resize_params* p, void* stream);

/* out = a*alpha + b*(1-alpha) with per-pixel alpha; windows given by pixel strides and base pointers. */
typedef struct {
    const float* a; int64_t a_ld; int64_t a_row; int64_t a_batch;
    const float* b; int64_t b_ld; int64_t b_row; int64_t b_batch;
    const float* alpha; int64_t al_ld; int64_t al_row; int64_t al_batch;
    float* out; int64_t o_ld; int64_t o_row; int64_t o_batch;
    int32_t B, H, W, C;
} ia_lerp_params;
int ia_lerp_alpha(const ia_lerp_params* p, void* stream);

/* One pyramid level of TriPlaneGenerator.rasterize (triplane_v20.py:328-338) without the [B][256][256][C] intermediate:
 *   out = aa_resize(grid_sample(tex, uv))*alpha + aa_resize(static[crop])*(1 - alpha)
 * Levels that shrink the samples (UW >= 2r) with C = 128, 256 or 512 run as ONE launch: a warp per output pixel sums the weights of
 * the samples of its window per texel cell, gathers each cell's four texels once and applies the static-crop resize and the alpha
 * blend (tmp is not touched; fp32 reassociation against the two-pass form, IA_RASTER_FUSED=0 / IA_RASTER_MERGE=0 select it).
 * Otherwise two passes: (1) grid_sample fused with the horizontal antialias taps -> tmp [B][UH][r][C]; (2) vertical taps + the
 * (up-sampling) resize of the static crop + the alpha blend.  Tap tables as in ia_resize_aa: (ux_*) UW -> r, (uy_*) UH -> r,
 * (sx_*) sw -> r, (sy_*) sh -> r.  tex is contiguous NHWC; C must be a multiple of 4. */
typedef struct {
    const float* tex; int32_t Ht, Wt, C;
    const float* uv; int64_t uv_ld; int32_t UH, UW;
    float* tmp;
    const float* stat; int64_t stat_ld; int32_t SH, SW, sy0, sx0;
    const float* alpha;
    float* out; int64_t out_ld;
    int32_t B, r;
    const int32_t* ux_start; const int32_t* ux_count; const float* ux_w; int32_t ux_max_taps;
    const int32_t* uy_start; const int32_t* uy_count; const float* uy_w; int32_t uy_max_taps;
    const int32_t* sx_start; const int32_t* sx_count; const float* sx_w; int32_t sx_max_taps;
    const int32_t* sy_start; const int32_t* sy_count; const float* sy_w; int32_t sy_max_taps;
} ia_raster_level_params;
int ia_raster_level(const ia_raster_level_params* p, void* stream);

/* Plane stitch (triplane_v20.py:119-128) in one pass: out = planes, except channels [0,32) (plane 0) inside the window
 * rows [y0,y0+wh) x cols [x0,x0+ww), which become stitch*alpha + planes*(1-alpha); out is fp32 or fp16 (out_fmt as
 * ia_render_params.planes_fmt).  planes [B][H][W][C] fp32 NHWC (pixel stride planes_ld), stitch [B][wh][ww][32], alpha [B][wh][ww]. */
typedef struct {
    const float* planes; int64_t planes_ld; int32_t B, H, W, C;
    const float* stitch; const float* alpha; int32_t y0, x0, wh, ww;
    void* out; int32_t out_fmt;
} ia_stitch_params;
int ia_stitch_planes(const ia_stitch_params* p, void* stream);

/* ---- volume renderer (renderer.py:309-469, ray_sampler.py:70-107, ray_marcher.py:25-57, triplane_v20.py:415-438) */
typedef struct {
    const float* planes; int64_t plane_px_ld;    /* [B][PH][PW][>=96]: plane p = channels [32p, 32p+32) */
    int32_t B, PH, PW;
    const float* cam;  int64_t cam_ld;           /* [B][>=25] c2w(16) | K(9); may be NULL when rays are given */
    const float* rays_o; const float* rays_d;    /* optional explicit rays [B][rays][3] (ImportanceRenderer API) */
    int32_t res;                                  /* neural rendering resolution N; rays = N*N */
    int32_t Dc, Df;                               /* coarse / importance samples (Dc<=96, Df<=96, Dc+Df<=192) */
    const float* jitter;                          /* [B][rays][Dc] U[0,1) (replaces rand_like, renderer.py:406) */
    const float* u;                               /* [B*rays][Df] or NULL -> linspace(0,1,Df) (evaluation) */
    float box_warp; int32_t white_back;
    const float* near_far;                        /* device [2] from ia_ray_bounds */
    const float* w1; const float* b1; const float* w2; const float* b2;  /* decoder [64][32],[64],[33][64],[33] raw */
    float* feat;                                  /* out [B][res][res][32] (NHWC feature image) */
    float* depth;                                 /* out [B][res][res] unclamped composite depth */
    float* wsum;                                  /* out [B][res][res] */
    float* depth_minmax;                          /* device [2] global min/max of all sample depths (init by callee) */
    /* Device scratch of ia_render_scratch_bytes() bytes, 16-byte aligned, owned by the caller for the duration of the call's
     * stream work: the decoder weights are staged into it in tensor-core fragment order by a small kernel in front of the
     * render kernel.  Per call, so that the library keeps no mutable device state: concurrent renders on different streams
     * (two generators, a graph replay next to an eager call) cannot race. */
    void* scratch;
    /* Decoder MLP arithmetic: IA_OPFMT_BF16X3-style 3-term split of fp16 hi/lo operands (0: reproduces the fp32 MLP to ~1e-6) or
     * single-pass fp16 products (IA_OPFMT_F16X1 = 1: a third of the mma.sync work; moves the final image by ~1.2e-4,
     * profiles/r1_render_precision_probe.json, r2_conv_precision_probe_mix_mlp.json).  fp32 accumulation in both. */
    int32_t mlp_fmt;
    /* Storage type of `planes`: 0 = fp32 (as declared), IA_OPFMT_F16X1 = fp16 (the pointer is then a __half*; plane_px_ld stays
     * in elements, a multiple of 8).  Halves the gather traffic -- the kernel's bound: 1536 B of L2 reads per sample in fp32 --
     * at 9.4e-6 on the final image; interpolation arithmetic is fp32 either way.  ia_stitch_planes produces such planes. */
    int32_t planes_fmt;
} ia_render_params;
int64_t ia_render_scratch_bytes(void);
/* near/far = mean_b ||c2w_b[:3,3]|| - 0.45 / + 0.6 (renderer.py:311-313), computed on device (no host sync). */
int ia_ray_bounds(const float* cam, int64_t cam_ld, int32_t B, float* near_far, void* stream);
int ia_ray_bounds_from_origins(const float* origins, int64_t n, float* near_far, void* stream);
int ia_render(const ia_render_params* p, void* stream);
/* MipRayMarcher2.run_forward (ray_marcher.py:25-57) on caller-provided samples: colors [rays][S][C], densities / depths [rays][S]
 * (sorted along S) -> rgb [rays][C] (scaled to (-1,1), + 1 - sum w with white_back), depth [rays] = sum w d_mid / sum w (unclamped;
 * follow with ia_depth_clamp on depth_minmax, which receives the global min / max of `depths`), weights [rays][S-1]. */
int ia_ray_march(const float* colors, const float* densities, const float* depths, int64_t rays, int32_t S, int32_t C, int32_t white_back,
                 float* rgb, float* depth, float* weights, float* depth_minmax, void* stream);
/* depth = clamp(nan_to_num(depth, inf), min, max) (ray_marcher.py:49-50) */
int ia_depth_clamp(float* depth, int64_t n, const float* depth_minmax, void* stream);
/* Ray generation only (RaySampler_zxc API): origins/dirs [B][rays][3]. */
int ia_ray_sampler(const float* cam, int64_t cam_ld, int32_t B, int32_t res, float* origins, float* dirs, void* stream);

/* ---- inversion encoder (encoder_inversion/models/{helpers,e4e,unet_encoders,uvnet}.py) -------------------------
 * The IR-SE50 trunks, FPN/GradualStyleBlocks, DoubleConv/ConvGRU decoders and SFT heads are built from ia_conv_tc
 * (mode 0: raw fp32 accumulators) plus the layout / normalisation / gating kernels below.  A view addresses element
 * (b, y, x, c) of a logical [B][H][W][C] tensor as
 *     p[b*s_img + (y/ps)*s_row + (x/ps)*s_pix + (c*ps*ps + (y%ps)*ps + (x%ps))*s_c]
 * so NCHW inputs, channel slices, stride-2 subsampling (MaxPool2d(1,2), the output of a stride-2 convolution computed
 * at full resolution), batch broadcast (s_img = 0, the `expand(T)` of unet_encoders.py:219) and
 * torch.nn.PixelShuffle(ps) (unet_encoders.py:74,89,282) are all plain views -- nothing is materialised. */
typedef struct {
    const float* p; int32_t C; int32_t ps;
    int64_t s_c, s_pix, s_row, s_img;
} ia_view;

/* Per-channel sum and sum of squares over [B][H][W] in float64 (sums[0..C) = sum, sums[C..2C) = sum of squares); the
 * callee zeroes `sums` first.  Feeds train-mode BatchNorm (eval_seq.py:92 leaves e4e and the UNet decoders in train
 * mode) and the squeeze of SEModule. */
int ia_enc_chan_stats(const ia_view* x, int32_t B, int32_t H, int32_t W, double* sums, void* stream);
/* torch.nn.BatchNorm2d folded to y = x*scale[c] + shift[c].  training != 0: batch statistics from `sums` over `count`
 * elements (biased variance), and running_mean / running_var (may be NULL) updated with `momentum` (unbiased variance) as
 * torch does; training == 0: running statistics.  gamma/beta may be NULL (1 / 0). */
int ia_enc_bn_fold(const double* sums, int64_t count, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, int32_t training, float momentum, float eps, int32_t C, float* scale, float* shift,
                   void* stream);
/* ia_enc_chan_stats + ia_enc_bn_fold (training) of ONE view in one launch: the last CTA to finish its partial sums folds.  sums
 * [2C] float64 and counter [1] int32 are scratch, ZERO before the first launch that uses them and left zero by every launch; they
 * belong to one stream at a time. */
int ia_enc_bn_stats_fold(const ia_view* x, int32_t B, int32_t H, int32_t W, double* sums, int32_t* counter, const float* gamma,
                         const float* beta, float* running_mean, float* running_var, float momentum, float eps, float* scale,
                         float* shift, void* stream);
/* Operand builder of an encoder convolution: concatenates up to 4 views along channels (torch.cat of
 * unet_encoders.py:78,96, uvnet.py:121,183), applies the per-channel affine of a preceding BatchNorm (scale/shift over
 * the concatenated channel index, may be NULL), an optional bias-free activation -- PReLU with per-channel `slope`, or
 * leaky ReLU with the scalar `lrelu` when slope is NULL and lrelu != 1 -- and writes the bf16 hi/lo split
 * [B][H][W][C_pad] (zero padded) that ia_conv_tc consumes, and/or the fp32 NHWC tensor out32 [B][H][W][sum C]. */
typedef struct {
    ia_view src[4]; int32_t nsrc;
    const float* scale; const float* shift; const float* slope; float lrelu;
    uint16_t* hi; uint16_t* lo; float* out32;
    int32_t B, H, W, C_pad;
} ia_enc_prep_params;
int ia_enc_prep(const ia_enc_prep_params* p, void* stream);
/* y = gate[b][c] * act2(act1(x*scale[c] + shift[c])) + (res*res_scale[c] + res_shift[c]).
 * act1: PReLU with per-channel slope1 when given, else IA_ACT_* `act` with `alpha`; act2: optional second PReLU (slope2;
 * DoubleConv ends in two PReLUs, unet_encoders.py:62-63).  Every pointer except x.p and y may be NULL.  Covers
 * conv+bias, BatchNorm+PReLU (helpers.py:113-118), the SE scale + shortcut add of bottleneck_IR_SE (helpers.py:120-124),
 * LeakyReLU heads, feature offsets (uvnet.py:187) and image differences (uvnet.py:181).  y is [B][H][W][C] with pixel
 * stride y_ld (>= C). */
typedef struct {
    ia_view x; const float* scale; const float* shift; const float* slope1; const float* slope2;
    int32_t act; float alpha;
    const float* gate;
    ia_view res; const float* res_scale; const float* res_shift;
    float* y; int64_t y_ld;
    int32_t B, H, W, C;
    /* optional second output (all NULL / 0: none): the bf16 hi/lo operand [B][H][W][e_C_pad] of the next convolution,
     * split(y * e_scale[c] + e_shift[c]) -- e.g. the BatchNorm that opens the next IR-SE unit (helpers.py:109) in eval mode */
    const float* e_scale; const float* e_shift; uint16_t* e_hi; uint16_t* e_lo; int32_t e_C_pad;
} ia_enc_affine_params;
int ia_enc_affine_act(const ia_enc_affine_params* p, void* stream);
/* pooled[b][c] = mean over H,W of x*scale[c] + shift[c] (AdaptiveAvgPool2d(1) of SEModule, helpers.py:65,74). */
int ia_enc_global_pool(const ia_view* x, const float* scale, const float* shift, int32_t B, int32_t H, int32_t W,
                       float* pooled, void* stream);
/* The whole SEModule gate in one launch (helpers.py:62-80): gate[b][c] = sigmoid(w2 . relu(w1 . pooled[b])) with pooled as in
 * ia_enc_global_pool; w1 [Cr][C], w2 [C][Cr] (the bias-free 1x1 convolutions fc1 / fc2).  sums [B][C] fp32 and counters [B] int32 are
 * scratch that must be ZERO before the first launch that uses them; every launch leaves them zero again.  They belong to one stream
 * at a time (launches on one stream may share them). */
int ia_enc_se_gate(const ia_view* x, const float* scale, const float* shift, int32_t B, int32_t H, int32_t W, const float* w1,
                   const float* w2, int32_t Cr, float* sums, int32_t* counters, float* gate, void* stream);
/* k x k box average, NHWC out [B][H/k][W/k][C] (AdaptiveAvgPool2d((256,256)) on 512^2 inputs, uvnet.py:108-109,
 * unet_encoders.py:199-200; only integer ratios occur). */
int ia_enc_avgpool(const ia_view* x, int32_t B, int32_t H, int32_t W, int32_t k, float* y, void* stream);
/* y = bilinear_upsample(x -> [H][W], align_corners=True) + lateral  (FPN _upsample_add, e4e.py:48-65).
 * x [B][h][w][C] and lateral/y [B][H][W][C] contiguous NHWC. */
int ia_enc_upsample_add(const float* x, int32_t B, int32_t h, int32_t w, int32_t C, const float* lateral, int32_t H,
                        int32_t W, float* y, void* stream);
/* ConvGRU gating (unet_encoders.py:27-32).  stage 0: raw [n][2C] = ih conv accumulators; writes rh = sigmoid(raw[:C] +
 * bias[:C]) * h and z = sigmoid(raw[C:] + bias[C:]).  stage 1: raw [n][C] = hh conv accumulators;
 * h_out = (1 - z)*h + z*tanh(raw + bias).  n = B*H*W pixels, all tensors contiguous NHWC; h may be NULL (zero state). */
int ia_enc_gru_gate(int32_t stage, const float* raw, const float* bias, const float* h, float* rh, float* z,
                    float* h_out, int64_t n, int32_t C, void* stream);
/* CS-SFT of the static backbone (networks_stylegan2_new.py:448-452): x[..., C/2:] = x[..., C/2:]*scale + shift, in
 * place on the fp32 NHWC activation x [B][HW][C]; scale/shift are views of [1|B][HW][C/2]. */
int ia_sft_half(float* x, int64_t x_ld, const ia_view* scale, const ia_view* shift, int32_t B, int32_t H, int32_t W,
                int32_t C, void* stream);

/* ---- "improved one-shot" encoder: Mix-Transformer pieces ------------------------------------------------------
 * encoder_inversion/models/mmseg/mix_transformer.py (MixVisionTransformer :201, Block :118, Attention :56, Mlp :18, DWConv :379,
 * OverlapPatchEmbed :159, transformer_block :453) and the UpLayer decoders of unet_transformer.py:255,340,523.  Tokens are
 * stored as [B][H][W][C] fp32 -- an NHWC image whose pixels are the tokens -- so every nn.Linear is a 1x1 ia_conv_tc and the
 * reference's flatten / transpose / reshape / permute calls do not exist here. */

/* Patch gather for a k x k, stride s, zero-pad convolution evaluated as ONE GEMM (OverlapPatchEmbed.proj: 7x7 s2 / s4, 3x3 s2;
 * Attention.sr: k = stride = sr_ratio): out[b][oy][ox][(ky*k + kx)*C + c] = cat(src)[b][oy*s - pad + ky][ox*s - pad + kx][c] as
 * the bf16 hi/lo operand [B][OH][OW][K_pad] (zero beyond k*k*C).  The matching weight is the OIHW filter permuted to
 * [O][kh][kw][I] and packed as a 1x1 convolution with k*k*C input channels. */
typedef struct {
    ia_view src[4]; int32_t nsrc;
    int32_t B, H, W, k, stride, pad, OH, OW;
    uint16_t* hi; uint16_t* lo; int32_t K_pad;
} ia_enc_im2col_params;
int ia_enc_im2col(const ia_enc_im2col_params* p, void* stream);
/* torch.nn.LayerNorm over the last dimension of x [rows][C] (row pitch x_ld) after adding pre_bias[c] (may be NULL: the bias of
 * the convolution whose raw accumulators x holds): y = (x - mean) * rsqrt(var + eps) * gamma + beta, biased variance.  Writes
 * fp32 out32 [rows][out32_ld] and / or the bf16 hi/lo operand [rows][C_pad] of the linear layer that follows. */
int ia_layer_norm(const float* x, int64_t x_ld, const float* pre_bias, const float* gamma, const float* beta, float eps,
                  int64_t rows, int32_t C, float* out32, int64_t out32_ld, uint16_t* hi, uint16_t* lo, int32_t C_pad,
                  void* stream);
/* Multi-head attention (mix_transformer.py:97-112): out[b][n][h*hd + d] = sum_m softmax_m((q[b][n][h] + q_bias) . (k[b][m][h] +
 * k_bias) * scale) (v[b][m][h][d] + v_bias).  q [B][Nq][q_ld], k / v [B][Nk][k_ld / v_ld] fp32 (raw accumulators of the q / kv
 * projections; k and v may point into one kv tensor), head h at channel offset h*head_dim; biases [heads*head_dim] or NULL
 * (qkv_bias=False).  head_dim 64 or 256.  The N x N score matrix is never stored (online softmax over 64-key tiles). */
typedef struct {
    const float* q; const float* k; const float* v; int64_t q_ld, k_ld, v_ld;
    const float* q_bias; const float* k_bias; const float* v_bias;
    int32_t B, heads, head_dim, Nq, Nk; float scale;
    float* out32; int64_t out32_ld;
    uint16_t* hi; uint16_t* lo; int32_t C_pad;      /* operand of the proj layer [B*Nq][C_pad], C_pad == heads*head_dim */
} ia_attention_params;
int ia_attention(const ia_attention_params* p, void* stream);
/* The same attention on the tensor cores for head_dim 256 without q/kv bias (transformer_block: 4 heads x 256, qkv_bias=False).
 * Operands are the bf16 hi/lo splits that the q and kv projections emit from their epilogues (ia_emit.hi1/lo1): q [B][Nq][q_ld],
 * kv [B][Nk][kv_ld] with k in channels [0, heads*256) and v in [heads*256, 2*heads*256) of a row; both q k^T and p v are 3-term
 * split products (hi*hi + hi*lo + lo*hi, fp32 accumulate) on mma.sync.m16n8k16, softmax in fp32.  Writes the proj operand
 * [B*Nq][C_pad] (+ optional fp32 copy). */
typedef struct {
    const uint16_t* q_hi; const uint16_t* q_lo; int64_t q_ld;
    const uint16_t* kv_hi; const uint16_t* kv_lo; int64_t kv_ld;
    int32_t B, heads, head_dim, Nq, Nk; float scale;
    float* out32; int64_t out32_ld;
    uint16_t* hi; uint16_t* lo; int32_t C_pad;
} ia_attention_tc_params;
int ia_attention_tc(const ia_attention_tc_params* p, void* stream);
/* Mix-FFN middle (mix_transformer.py:46-49): gelu(depthwise3x3(x + in_bias) + bias), exact (erf) GELU, zero padding.
 * x [B][H][W][C] fp32 contiguous (raw fc1 accumulators), w [C][3][3]; writes fp32 out32 [B][H][W][C] and / or the fc2 operand. */
int ia_dwconv_gelu(const float* x, const float* in_bias, const float* w, const float* bias, int32_t B, int32_t H, int32_t W,
                   int32_t C, float* out32, uint16_t* hi, uint16_t* lo, int32_t C_pad, void* stream);

/* ---- mesh-condition producer (the step in front of the path) -------------------------------------------------
 * Faceverse_manager.make_driven_rendering (data_preprocess/FaceVerse/renderer.py:45-84): driving coefficients -> 3DMM vertices
 * -> orthographic rasterisation of per-vertex (u, v, mask) -> the [256][256][3] uvcoords_image of TriPlaneGenerator.synthesis.
 * Replaces FaceVerseModel_v3.get_vs / compute_eye_rotation_matrix (FaceVerseModel_v3.py:237-244,303-325) and the pytorch3d
 * MeshRasterizer + render_after_rasterize (ortho_renderer.py:52-100, volumetric_rendering/renderer.py:556-571). */

/* Per driving frame b: exp_out[b] = expression coefficients coeff[b][exp_off .. +exp_dims) with entries -4 / -2 clamped to
 * [-0.75, 0.6] / [-0.75, 0.75] (renderer.py:48-49) and, when the bases are given, retargeted (e - base_drive_exp) +
 * base_avatar_exp (:50-53); eye_rot[b][2][9] = Ry(eye[1]) Rx(eye[0]) of the left / right eye-ball from coeff[b][eye_off .. +4). */
int ia_mesh_coeffs(const float* coeff, int64_t coeff_ld, int32_t B, int32_t exp_off, int32_t exp_dims, int32_t eye_off,
                   const float* base_drive_exp, const float* base_avatar_exp, float* exp_out, float* eye_rot, void* stream);
/* centres[2][3]: mean of the identity's neutral eye-ball vertices [eye0,eye1) / [eye1,eye2) with z + 0.005 (FaceVerseModel_v3.py:252-264). */
int ia_mesh_eye_centres(const float* neutral, int32_t eye0, int32_t eye1, int32_t eye2, float* centres, void* stream);
typedef struct {
    const float* neutral;        /* [NV][3] identity shape: meanshape + idBase * id (loader conventions applied) */
    const float* exp_basis_t;    /* [exp_dims][NV*3] expression basis, transposed (coalesced over vertices) */
    const float* exp;            /* [B][exp_dims] from ia_mesh_coeffs */
    int32_t B, NV, exp_dims;
    int32_t eye0, eye1, eye2;    /* eye-ball vertex ranges */
    const float* eye_rot;        /* [B][2][9] */
    const float* eye_centre;     /* [2][3] */
    float M[12];                 /* last step: out = M[:, :3] v + M[:, 3]  (fv2fl transform, orthographic shift / scale, z flip) */
    float* verts;                /* out [B][NV][3] */
} ia_blendshape_params;
int ia_blendshape(const ia_blendshape_params* p, void* stream);
typedef struct {
    const float* verts; int32_t B, NV;        /* [B][NV][3]; camera: x_ndc = -x, y_ndc = -y, view z = z + cam_z */
    const int32_t* tri; int32_t F;             /* [F][3] */
    const float* attr;                         /* [NV][3] per-vertex (u, v, face mask) */
    int32_t size; float cam_z; float blur_radius;
    int32_t crop_x, crop_y, crop_w, crop_h;    /* window of the size x size raster that is written out */
    unsigned long long* zbuf;                  /* scratch, ia_ortho_raster_scratch_bytes(B, size) */
    float* out;                                /* [B][crop_h][crop_w][3]: (u, v, mask) * vis * mask, mask binarised at 0.5 */
    int32_t* pix_to_face;                      /* optional [B][crop_h][crop_w] face index or -1 (diagnostics), or NULL */
} ia_ortho_raster_params;
int64_t ia_ortho_raster_scratch_bytes(int32_t B, int32_t size);
int ia_ortho_raster(const ia_ortho_raster_params* p, void* stream);

/* ---- output stage (the step after the path) ------------------------------------------------------------
 * layout_grid (reenact_avatar_next3d.py:117-131) with float_to_uint8 and chw_to_hwc: frame g = gy*grid_w + gx of the
 * [B][C][H][W]-indexed image (element strides s_b, s_c, s_h, s_w -- the generator's channels-last output or a planar
 * tensor) goes to out[(gy*H + y)][(gx*W + x)][c] = (uint8) clamp(img*127.5 + 128, 0, 255) (truncation, as .to(torch.uint8)). */
int ia_layout_grid_u8(const float* img, int64_t s_b, int64_t s_c, int64_t s_h, int64_t s_w, int32_t grid_h, int32_t grid_w,
                      int32_t C, int32_t H, int32_t W, uint8_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* INVERTAVATAR_B200_H_ */
