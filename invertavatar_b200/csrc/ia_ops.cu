// Library management + the torch_utils/ops plugin equivalents + small dense layers.
// Behaviour follows reference torch_utils/ops/{bias_act,upfirdn2d}.{py,cu} and
// training_avatar_texture/networks_stylegan2_new.py:96-127,233-268 (written from scratch).
#include <stdarg.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <stdlib.h>
#include "ia_common.cuh"

namespace ia {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- per-launch event profiler (bench.py's roofline leg) -------------------------------------------
struct ProfRec { const char* name; cudaEvent_t t0, t1; };
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static thread_local int g_prof_open = -1;
static thread_local cudaStream_t g_prof_stream = nullptr;
void prof_begin(const char* name, cudaStream_t stream) {
    if (!g_prof_on) return;
    ProfRec r{name, nullptr, nullptr};
    if (cudaEventCreate(&r.t0) != cudaSuccess || cudaEventCreate(&r.t1) != cudaSuccess) return;
    cudaEventRecord(r.t0, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
    g_prof_open = (int)g_prof.size() - 1;
    g_prof_stream = stream;
}
// IA_PROF_DETAIL=1: per-shape names for the convolution launches (tools/prof_layers.py); interned, never freed.
const char* prof_detail_name(const char* base, int ntaps, int gh, int gw, int cin, int cout) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("IA_PROF_DETAIL"); on = (e && atoi(e)) ? 1 : 0; }
    if (!on || !g_prof_on) return base;
    static std::map<std::string, std::string*> names;
    char tmp[160];
    snprintf(tmp, sizeof(tmp), "%s[t%d %dx%d %d->%d]", base, ntaps, gh, gw, cin, cout);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    auto it = names.find(tmp);
    if (it == names.end()) it = names.emplace(tmp, new std::string(tmp)).first;
    return it->second->c_str();
}
void prof_end() {
    if (g_prof_open < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_prof_open < (int)g_prof.size()) cudaEventRecord(g_prof[g_prof_open].t1, g_prof_stream);
    g_prof_open = -1;
}
}  // namespace ia

extern "C" int ia_profile_begin(void) {
    std::lock_guard<std::mutex> lk(ia::g_prof_mu);
    for (auto& r : ia::g_prof) { cudaEventDestroy(r.t0); cudaEventDestroy(r.t1); }
    ia::g_prof.clear();
    ia::g_prof_on = true;
    return 0;
}

extern "C" int64_t ia_profile_report(char* buf, int64_t buflen) {
    std::lock_guard<std::mutex> lk(ia::g_prof_mu);
    ia::g_prof_on = false;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { ia::set_error("ia_profile_report: %s", cudaGetErrorString(e)); return -1; }
    std::map<std::string, std::pair<double, int64_t>> agg;
    for (auto& r : ia::g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.t0, r.t1) == cudaSuccess) { auto& a = agg[r.name]; a.first += ms; a.second += 1; }
        cudaEventDestroy(r.t0); cudaEventDestroy(r.t1);
    }
    ia::g_prof.clear();
    std::string out = "{";
    bool first = true;
    for (auto& kv : agg) {
        char tmp[256];
        snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"ms\": %.6f, \"launches\": %lld}", first ? "" : ", ", kv.first.c_str(), kv.second.first,
                 (long long)kv.second.second);
        out += tmp;
        first = false;
    }
    out += "}";
    if (buf && buflen > 0) {
        size_t n = out.size() < (size_t)buflen - 1 ? out.size() : (size_t)buflen - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return (int64_t)out.size();
}

using namespace ia;

extern "C" int ia_abi_version(void) { return IA_ABI_VERSION; }
extern "C" const char* ia_last_error(void) { return ia::g_err; }
extern "C" int64_t ia_launch_count(void) { return ia::g_launches.load(); }
extern "C" void ia_reset_launch_count(void) { ia::g_launches.store(0); }
extern "C" int ia_set_device(int device) {
    cudaError_t e = cudaSetDevice(device);
    IA_CHECK(e == cudaSuccess, "ia_set_device(%d): %s", device, cudaGetErrorString(e));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// element types of the plugin-level entry points: the reference dispatches bias_act / upfirdn2d over double, float and half
// (AT_DISPATCH_FLOATING_TYPES_AND_HALF, bias_act.cpp:81, upfirdn2d.cpp:67) with float arithmetic for half/float and double
// arithmetic for double (InternalType<T>, bias_act.cu:18-21)
// ------------------------------------------------------------------------------------------------
#include <cuda_fp16.h>
template <typename T> struct Internal { typedef float type; };
template <> struct Internal<double> { typedef double type; };
template <typename T, typename S> __device__ __forceinline__ S ld_as(const T* p) { return (S)(*p); }
template <> __device__ __forceinline__ float ld_as<__half, float>(const __half* p) { return __half2float(*p); }
template <typename T, typename S> __device__ __forceinline__ void st_as(T* p, S v) { *p = (T)v; }
template <> __device__ __forceinline__ void st_as<__half, float>(__half* p, float v) { *p = __float2half_rn(v); }

__device__ __forceinline__ double apply_act_f64(double x, int act, double alpha) {
    switch (act) {
        case IA_ACT_LINEAR: return x;
        case IA_ACT_RELU: return x > 0. ? x : 0.;
        case IA_ACT_LRELU: return x > 0. ? x : x * alpha;
        case IA_ACT_TANH: return tanh(x);
        case IA_ACT_SIGMOID: return 1. / (1. + exp(-x));
        case IA_ACT_ELU: return x > 0. ? x : expm1(x);
        case IA_ACT_SELU: return x > 0. ? 1.0507009873554805 * x : 1.0507009873554805 * 1.6732632423543772 * expm1(x);
        case IA_ACT_SOFTPLUS: return x > 20. ? x : log1p(exp(x));
        case IA_ACT_SWISH: return x / (1. + exp(-x));
    }
    return x;
}
__device__ __forceinline__ float act_gc(float x, int act, float alpha, float gain, float clamp) { return act_gain_clamp(x, act, alpha, gain, clamp); }
__device__ __forceinline__ double act_gc(double x, int act, float alpha, float gain, float clamp) {
    x = apply_act_f64(x, act, (double)alpha) * (double)gain;
    if (clamp >= 0.f) x = fmin(fmax(x, -(double)clamp), (double)clamp);
    return x;
}

// ------------------------------------------------------------------------------------------------
// bias_act
// ------------------------------------------------------------------------------------------------
__global__ void bias_act_kernel(const float* __restrict__ x, const float* __restrict__ b, float* __restrict__ y,
                                int64_t numel, int64_t C, int64_t inner, int act, float alpha, float gain, float clamp) {
    int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= numel) return;
    if (i0 + 3 < numel && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
        float4 v = *reinterpret_cast<const float4*>(x + i0);
        float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float t = r[k];
            if (b) t += b[((i0 + k) / inner) % C];
            r[k] = act_gain_clamp(t, act, alpha, gain, clamp);
        }
        *reinterpret_cast<float4*>(y + i0) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
        for (int k = 0; k < 4 && i0 + k < numel; ++k) {
            float t = x[i0 + k];
            if (b) t += b[((i0 + k) / inner) % C];
            y[i0 + k] = act_gain_clamp(t, act, alpha, gain, clamp);
        }
    }
}

// half / double I/O (the fp32 kernel above is the vectorised hot variant)
template <typename T>
__global__ void bias_act_any_kernel(const T* __restrict__ x, const T* __restrict__ b, T* __restrict__ y, int64_t numel, int64_t C,
                                    int64_t inner, int act, float alpha, float gain, float clamp) {
    typedef typename Internal<T>::type S;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numel) return;
    S t = ld_as<T, S>(x + i);
    if (b) t += ld_as<T, S>(b + (i / inner) % C);
    st_as<T, S>(y + i, act_gc(t, act, alpha, gain, clamp));
}

extern "C" int ia_bias_act(const void* x, const void* b, void* y, int64_t numel, int64_t C, int64_t inner,
                           int act, float alpha, float gain, float clamp, int dtype, void* stream) {
    IA_CHECK(x && y, "ia_bias_act: null tensor");
    IA_CHECK(act >= IA_ACT_LINEAR && act <= IA_ACT_SWISH, "ia_bias_act: bad activation id %d", act);
    IA_CHECK(dtype == IA_DTYPE_F32 || dtype == IA_DTYPE_F16 || dtype == IA_DTYPE_F64, "ia_bias_act: x must be float16, float32 or float64 (dtype id %d)", dtype);
    IA_CHECK(numel >= 0 && numel <= 0x7fffffffLL * 4, "ia_bias_act: numel too large");
    IA_CHECK(b == nullptr || (C > 0 && inner > 0), "ia_bias_act: bias needs C>0 and inner>0");
    if (numel == 0) return 0;
    int threads = 256;
    const int64_t Cc = C > 0 ? C : 1, in = inner > 0 ? inner : 1;
    ia::prof_begin("ia_bias_act", as_stream(stream));
    if (dtype == IA_DTYPE_F32) {
        int64_t blocks = cdiv(cdiv(numel, 4), threads);
        bias_act_kernel<<<(unsigned)blocks, threads, 0, as_stream(stream)>>>((const float*)x, (const float*)b, (float*)y, numel, Cc, in, act, alpha, gain, clamp);
    } else if (dtype == IA_DTYPE_F16) {
        bias_act_any_kernel<__half><<<(unsigned)cdiv(numel, threads), threads, 0, as_stream(stream)>>>((const __half*)x, (const __half*)b, (__half*)y, numel, Cc, in, act, alpha, gain, clamp);
    } else {
        bias_act_any_kernel<double><<<(unsigned)cdiv(numel, threads), threads, 0, as_stream(stream)>>>((const double*)x, (const double*)b, (double*)y, numel, Cc, in, act, alpha, gain, clamp);
    }
    IA_LAUNCH_CHECK("ia_bias_act");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// upfirdn2d (generic, strided)
// ------------------------------------------------------------------------------------------------
// One output element: sum over the filter window of the zero-stuffed, padded input.  BIAS/ACT (filtered_lrelu's first stage):
// `bias` is added to every real input sample before filtering and leaky-ReLU * act_gain, clamp are applied to the result.
// fh == 0 marks a separable filter: f holds fw taps applied along both axes (weight f[ky]*f[kx]).
template <typename T, typename S, bool FUSED>
__device__ __forceinline__ S upfirdn2d_point(const T* __restrict__ xb, int64_t xs_h, int64_t xs_w, int inH, int inW, const float* __restrict__ f,
                                             int fh, int fw, int upx, int upy, int uy0, int ux0, int flip, S bias) {
    const int fhh = fh > 0 ? fh : fw;
    S acc = (S)0;
    for (int fy = 0; fy < fhh; ++fy) {
        const int uy = uy0 + fy;
        if (uy < 0 || uy % upy != 0) continue;
        const int iy = uy / upy;
        if (iy >= inH) continue;
        const int ky = flip ? fy : (fhh - 1 - fy);   // reference: correlate with the flipped filter unless flip_filter is set
        for (int fx = 0; fx < fw; ++fx) {
            const int ux = ux0 + fx;
            if (ux < 0 || ux % upx != 0) continue;
            const int ix = ux / upx;
            if (ix >= inW) continue;
            const int kx = flip ? fx : (fw - 1 - fx);
            const S w = fh > 0 ? (S)f[ky * fw + kx] : (S)f[ky] * (S)f[kx];
            S v = ld_as<T, S>(xb + iy * xs_h + ix * xs_w);
            if (FUSED) v += bias;
            acc += w * v;
        }
    }
    return acc;
}

template <typename T>
__global__ void upfirdn2d_kernel(ia_upfirdn2d_params p) {
    typedef typename Internal<T>::type S;
    int64_t total = (int64_t)p.N * p.C * p.outH * p.outW;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    // decode with the fastest index following the smallest output stride (NCHW vs NHWC)
    int n, c, oy, ox;
    if (p.ys_c == 1) {  // channels-last: c fastest
        c = idx % p.C; int64_t t = idx / p.C;
        ox = t % p.outW; t /= p.outW;
        oy = t % p.outH; n = (int)(t / p.outH);
    } else {
        ox = idx % p.outW; int64_t t = idx / p.outW;
        oy = t % p.outH; t /= p.outH;
        c = t % p.C; n = (int)(t / p.C);
    }
    // position in the zero-stuffed, padded signal: top-left of the filter window in upsampled coordinates
    const T* xb = reinterpret_cast<const T*>(p.x) + n * p.xs_n + c * p.xs_c;
    const S acc = upfirdn2d_point<T, S, false>(xb, p.xs_h, p.xs_w, p.inH, p.inW, p.f, p.fh, p.fw, p.upx, p.upy, oy * p.downy - p.pady0,
                                               ox * p.downx - p.padx0, p.flip, (S)0);
    st_as<T, S>(reinterpret_cast<T*>(p.y) + n * p.ys_n + c * p.ys_c + oy * p.ys_h + ox * p.ys_w, acc * (S)p.gain);
}

extern "C" int ia_upfirdn2d(const ia_upfirdn2d_params* p, void* stream) {
    IA_CHECK(p && p->x && p->y && p->f, "ia_upfirdn2d: null tensor");
    IA_CHECK(p->upx >= 1 && p->upy >= 1 && p->downx >= 1 && p->downy >= 1, "ia_upfirdn2d: bad up/down factors");
    IA_CHECK(p->fh >= 1 && p->fw >= 1 && p->fh <= 64 && p->fw <= 64, "ia_upfirdn2d: bad filter size");
    IA_CHECK(p->dtype == IA_DTYPE_F32 || p->dtype == IA_DTYPE_F16 || p->dtype == IA_DTYPE_F64, "ia_upfirdn2d: x must be float16, float32 or float64 (dtype id %d)", p->dtype);
    int64_t total = (int64_t)p->N * p->C * p->outH * p->outW;
    IA_CHECK(total >= 0 && total < (1LL << 40), "ia_upfirdn2d: output too large");
    if (total == 0) return 0;
    ia::prof_begin("ia_upfirdn2d", as_stream(stream));
    const unsigned blocks = (unsigned)cdiv(total, 256);
    if (p->dtype == IA_DTYPE_F32) upfirdn2d_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(*p);
    else if (p->dtype == IA_DTYPE_F16) upfirdn2d_kernel<__half><<<blocks, 256, 0, as_stream(stream)>>>(*p);
    else upfirdn2d_kernel<double><<<blocks, 256, 0, as_stream(stream)>>>(*p);
    IA_LAUNCH_CHECK("ia_upfirdn2d");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// filtered_lrelu (third plugin of the reference's boundary, filtered_lrelu.cpp:20,217): bias -> up-FIR -> leaky ReLU * gain,
// clamp -> down-FIR as two kernels through an fp32 workspace.  Forward only: the sign tensors that the reference's backward
// pass reads/writes are not produced -- a call that asks for them answers -1, the reference's own "no specialised kernel"
// code, on which its Python falls back to the bias_act / upfirdn2d composition (filtered_lrelu.py:225-231).
// ------------------------------------------------------------------------------------------------
struct FluGeom { int cw, ch; };
static int flu_geometry(const ia_filtered_lrelu_params* p, FluGeom* g) {
    const int fuw = p->fu ? p->fuw : 1, fuh = p->fu ? (p->fuh > 0 ? p->fuh : p->fuw) : 1;
    const int fdw = p->fd ? p->fdw : 1, fdh = p->fd ? (p->fdh > 0 ? p->fdh : p->fdw) : 1;
    const int64_t cw = (int64_t)p->inW * p->up + (p->px0 + p->px1) - (fuw - 1);
    const int64_t ch = (int64_t)p->inH * p->up + (p->py0 + p->py1) - (fuh - 1);
    IA_CHECK(cw > fdw - 1 && ch > fdh - 1, "ia_filtered_lrelu: upsampled buffer must be at least the size of downsampling filter");
    IA_CHECK(cw <= 0x7fffffff && ch <= 0x7fffffff, "ia_filtered_lrelu: upsampled buffer is too large");
    const int64_t yw = (cw - (fdw - 1) + (p->down - 1)) / p->down, yh = (ch - (fdh - 1) + (p->down - 1)) / p->down;
    IA_CHECK(yw > 0 && yh > 0, "ia_filtered_lrelu: output must be at least 1x1");
    IA_CHECK(yw == p->outW && yh == p->outH, "ia_filtered_lrelu: output is %lld x %lld for these arguments, caller allocated %d x %d",
             (long long)yh, (long long)yw, p->outH, p->outW);
    g->cw = (int)cw; g->ch = (int)ch;
    return 0;
}
static int flu_validate(const ia_filtered_lrelu_params* p) {
    IA_CHECK(p && p->x && p->y, "ia_filtered_lrelu: null tensor");
    IA_CHECK(p->dtype == IA_DTYPE_F32 || p->dtype == IA_DTYPE_F16, "ia_filtered_lrelu: x and b must be float16 or float32");
    IA_CHECK(p->N > 0 && p->C > 0 && p->inH > 0 && p->inW > 0, "ia_filtered_lrelu: x is empty");
    IA_CHECK(p->up >= 1 && p->down >= 1, "ia_filtered_lrelu: up and down must be at least 1");
    IA_CHECK(!p->fu || (p->fuw >= 1 && p->fuw <= 64 && p->fuh >= 0 && p->fuh <= 64), "ia_filtered_lrelu: bad fu shape");
    IA_CHECK(!p->fd || (p->fdw >= 1 && p->fdw <= 64 && p->fdh >= 0 && p->fdh <= 64), "ia_filtered_lrelu: bad fd shape");
    return 0;
}

template <typename T>
__global__ void flu_up_kernel(ia_filtered_lrelu_params p, int cw, int ch, const float* __restrict__ one) {
    const int64_t total = (int64_t)p.N * p.C * ch * cw;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ux = (int)(idx % cw); int64_t t = idx / cw;
    const int uy = (int)(t % ch); t /= ch;
    const int c = (int)(t % p.C), n = (int)(t / p.C);
    const T* xb = reinterpret_cast<const T*>(p.x) + n * p.xs_n + c * p.xs_c;
    const float bias = p.b ? ld_as<T, float>(reinterpret_cast<const T*>(p.b) + c) : 0.f;
    const float* f = p.fu ? p.fu : one;
    const int fw = p.fu ? p.fuw : 1, fh = p.fu ? p.fuh : 1;
    float v = upfirdn2d_point<T, float, true>(xb, p.xs_h, p.xs_w, p.inH, p.inW, f, fh, fw, p.up, p.up, uy - p.py0, ux - p.px0, p.flip, bias);
    v *= (float)(p.up * p.up);
    v = (v > 0.f ? v : v * p.slope) * p.gain;
    if (p.clamp >= 0.f) v = fminf(fmaxf(v, -p.clamp), p.clamp);
    reinterpret_cast<float*>(p.workspace)[idx] = v;
}

template <typename T>
__global__ void flu_down_kernel(ia_filtered_lrelu_params p, int cw, int ch, const float* __restrict__ one) {
    const int64_t total = (int64_t)p.N * p.C * p.outH * p.outW;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int ox = (int)(idx % p.outW); int64_t t = idx / p.outW;
    const int oy = (int)(t % p.outH); t /= p.outH;
    const int c = (int)(t % p.C), n = (int)(t / p.C);
    const float* wb = reinterpret_cast<const float*>(p.workspace) + ((int64_t)n * p.C + c) * ch * cw;
    const float* f = p.fd ? p.fd : one;
    const int fw = p.fd ? p.fdw : 1, fh = p.fd ? p.fdh : 1;
    const float v = upfirdn2d_point<float, float, false>(wb, cw, 1, ch, cw, f, fh, fw, 1, 1, oy * p.down, ox * p.down, p.flip, 0.f);
    st_as<T, float>(reinterpret_cast<T*>(p.y) + n * p.ys_n + c * p.ys_c + oy * p.ys_h + ox * p.ys_w, v);
}

__device__ float g_flu_one = 1.0f;     // identity filter (read-only)

extern "C" int64_t ia_filtered_lrelu_workspace(const ia_filtered_lrelu_params* p) {
    if (flu_validate(p)) return -1;
    FluGeom g;
    if (flu_geometry(p, &g)) return -1;
    return (int64_t)p->N * p->C * g.ch * g.cw * (int64_t)sizeof(float);
}

extern "C" int ia_filtered_lrelu(const ia_filtered_lrelu_params* p, void* stream) {
    if (int rc = flu_validate(p)) return rc;
    if (p->si || p->so || p->write_signs) {
        ia::set_error("ia_filtered_lrelu: sign tensors (backward pass) are not produced by this forward-only library; -1 = no specialised kernel");
        return -1;
    }
    FluGeom g;
    if (int rc = flu_geometry(p, &g)) return rc;
    const int64_t need = (int64_t)p->N * p->C * g.ch * g.cw * (int64_t)sizeof(float);
    IA_CHECK(p->workspace && p->workspace_bytes >= need, "ia_filtered_lrelu: workspace of %lld bytes needed (ia_filtered_lrelu_workspace)", (long long)need);
    float* one = nullptr;
    cudaError_t e = cudaGetSymbolAddress((void**)&one, g_flu_one);
    IA_CHECK(e == cudaSuccess, "ia_filtered_lrelu: %s", cudaGetErrorString(e));
    const int64_t tot_up = (int64_t)p->N * p->C * g.ch * g.cw, tot_dn = (int64_t)p->N * p->C * p->outH * p->outW;
    ia::prof_begin("ia_filtered_lrelu(up)", as_stream(stream));
    if (p->dtype == IA_DTYPE_F32) flu_up_kernel<float><<<(unsigned)cdiv(tot_up, 256), 256, 0, as_stream(stream)>>>(*p, g.cw, g.ch, one);
    else flu_up_kernel<__half><<<(unsigned)cdiv(tot_up, 256), 256, 0, as_stream(stream)>>>(*p, g.cw, g.ch, one);
    IA_LAUNCH_CHECK("ia_filtered_lrelu(up)");
    ia::prof_begin("ia_filtered_lrelu(down)", as_stream(stream));
    if (p->dtype == IA_DTYPE_F32) flu_down_kernel<float><<<(unsigned)cdiv(tot_dn, 256), 256, 0, as_stream(stream)>>>(*p, g.cw, g.ch, one);
    else flu_down_kernel<__half><<<(unsigned)cdiv(tot_dn, 256), 256, 0, as_stream(stream)>>>(*p, g.cw, g.ch, one);
    IA_LAUNCH_CHECK("ia_filtered_lrelu(down)");
    return 0;
}

template <typename T>
__global__ void flu_act_kernel(T* __restrict__ x, int64_t numel, float gain, float slope, float clamp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numel) return;
    float v = ld_as<T, float>(x + i);
    v = (v > 0.f ? v : v * slope) * gain;
    if (clamp >= 0.f) v = fminf(fmaxf(v, -clamp), clamp);
    st_as<T, float>(x + i, v);
}

extern "C" int ia_filtered_lrelu_act(void* x, int64_t numel, int dtype, const uint8_t* si, int32_t sx, int32_t sy, float gain, float slope,
                                     float clamp, int32_t write_signs, uint8_t* so, void* stream) {
    (void)sx; (void)sy;
    IA_CHECK(x, "ia_filtered_lrelu_act: null tensor");
    IA_CHECK(dtype == IA_DTYPE_F32 || dtype == IA_DTYPE_F16 || dtype == IA_DTYPE_F64, "ia_filtered_lrelu_act: x must be float16, float32 or float64");
    if (si || so || write_signs) {
        ia::set_error("ia_filtered_lrelu_act: sign tensors (backward pass) are not supported by this forward-only library");
        return -1;
    }
    if (numel == 0) return 0;
    ia::prof_begin("ia_filtered_lrelu_act", as_stream(stream));
    const unsigned blocks = (unsigned)cdiv(numel, 256);
    if (dtype == IA_DTYPE_F32) flu_act_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>((float*)x, numel, gain, slope, clamp);
    else if (dtype == IA_DTYPE_F16) flu_act_kernel<__half><<<blocks, 256, 0, as_stream(stream)>>>((__half*)x, numel, gain, slope, clamp);
    else flu_act_kernel<double><<<blocks, 256, 0, as_stream(stream)>>>((double*)x, numel, gain, slope, clamp);
    IA_LAUNCH_CHECK("ia_filtered_lrelu_act");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// fully connected: one warp per output feature, all batch rows per warp (weights read once)
// ------------------------------------------------------------------------------------------------
template <int MAXB>
__global__ void fc_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                          float* __restrict__ y, int B, int In, int Out, float w_gain, float b_gain, int act,
                          float alpha, float act_gain, int64_t xs, int64_t ys, int b0) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= Out) return;
    int nb = min(MAXB, B - b0);
    float acc[MAXB];
#pragma unroll
    for (int i = 0; i < MAXB; ++i) acc[i] = 0.f;
    const float* wr = w + (int64_t)warp * In;
    ia::warp_dot_rows<MAXB, false>(wr, x + (int64_t)b0 * xs, xs, In, nb, lane, acc);
#pragma unroll
    for (int i = 0; i < MAXB; ++i) {
        float v = acc[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && i < nb) {
            float t = v * w_gain + (bias ? bias[warp] * b_gain : 0.f);
            y[(int64_t)(b0 + i) * ys + warp] = apply_act(t, act, alpha) * act_gain;
        }
    }
}

extern "C" int ia_fully_connected(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t In,
                                  int32_t Out, float w_gain, float b_gain, int act, float alpha, float act_gain,
                                  int64_t x_stride, int64_t y_stride, void* stream) {
    IA_CHECK(x && w && y, "ia_fully_connected: null tensor");
    IA_CHECK(B >= 0 && In > 0 && Out > 0, "ia_fully_connected: bad shape");
    if (B == 0) return 0;
    int threads = 256;
    int blocks = (int)cdiv((int64_t)Out * 32, threads);
    for (int b0 = 0; b0 < B; b0 += 8) {
        ia::prof_begin("ia_fully_connected", as_stream(stream));
        fc_kernel<8><<<blocks, threads, 0, as_stream(stream)>>>(x, w, bias, y, B, In, Out, w_gain, b_gain, act, alpha,
                                                                  act_gain, x_stride, y_stride, b0);
        IA_LAUNCH_CHECK("ia_fully_connected");
    }
    return 0;
}

__global__ void normalize_2nd_moment_kernel(const float* __restrict__ x, float* __restrict__ y, int D, float eps,
                                            int64_t xs, int64_t ys) {
    int b = blockIdx.x;
    __shared__ float red[32];
    float s = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        float v = x[b * xs + i];
        s += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) red[0] = rsqrtf(t / (float)D + eps);
    }
    __syncthreads();
    float r = red[0];
    for (int i = threadIdx.x; i < D; i += blockDim.x) y[b * ys + i] = x[b * xs + i] * r;
}

extern "C" int ia_normalize_2nd_moment(const float* x, float* y, int32_t B, int32_t D, float eps, int64_t x_stride,
                                       int64_t y_stride, void* stream) {
    IA_CHECK(x && y && D > 0, "ia_normalize_2nd_moment: bad arguments");
    if (B == 0) return 0;
    ia::prof_begin("ia_normalize_2nd_moment", as_stream(stream));
    normalize_2nd_moment_kernel<<<B, 256, 0, as_stream(stream)>>>(x, y, D, eps, x_stride, y_stride);
    IA_LAUNCH_CHECK("ia_normalize_2nd_moment");
    return 0;
}

__global__ void broadcast_truncate_kernel(const float* __restrict__ w, const float* __restrict__ w_avg,
                                          float* __restrict__ ws, int B, int num_ws, int D, float psi, int cutoff) {
    int64_t total = (int64_t)B * num_ws * D;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int d = i % D;
    int k = (i / D) % num_ws;
    int b = (int)(i / ((int64_t)D * num_ws));
    float v = w[(int64_t)b * D + d];
    if (psi != 1.f && k < cutoff) {
        // torch.lerp(start=w_avg, end=v, weight=psi): weight < 0.5 ? a + t*(b-a) : b - (b-a)*(1-t)
        float a = w_avg[d];
        v = (psi < 0.5f) ? a + psi * (v - a) : v - (v - a) * (1.f - psi);
    }
    ws[i] = v;
}

extern "C" int ia_broadcast_truncate(const float* w, const float* w_avg, float* ws, int32_t B, int32_t num_ws,
                                     int32_t D, float psi, int32_t cutoff, void* stream) {
    IA_CHECK(w && ws, "ia_broadcast_truncate: null tensor");
    IA_CHECK(psi == 1.f || w_avg, "ia_broadcast_truncate: truncation needs w_avg");
    int64_t total = (int64_t)B * num_ws * D;
    if (total == 0) return 0;
    ia::prof_begin("ia_broadcast_truncate", as_stream(stream));
    broadcast_truncate_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(w, w_avg, ws, B, num_ws, D,
                                                                                        psi, cutoff);
    IA_LAUNCH_CHECK("ia_broadcast_truncate");
    return 0;
}

// ------------------------------------------------------------------------------------------------
// per-pixel alpha blend
// ------------------------------------------------------------------------------------------------
__global__ void lerp_alpha_kernel(ia_lerp_params p) {
    int64_t total = (int64_t)p.B * p.H * p.W * p.C;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c = i % p.C; int64_t t = i / p.C;
    int x = t % p.W; t /= p.W;
    int y = t % p.H; int b = (int)(t / p.H);
    float al = p.alpha[b * p.al_batch + y * p.al_row + x * p.al_ld];
    float av = p.a[b * p.a_batch + y * p.a_row + x * p.a_ld + c];
    float bv = p.b[b * p.b_batch + y * p.b_row + x * p.b_ld + c];
    p.out[b * p.o_batch + y * p.o_row + x * p.o_ld + c] = av * al + bv * (1.f - al);
}

extern "C" int ia_lerp_alpha(const ia_lerp_params* p, void* stream) {
    IA_CHECK(p && p->a && p->b && p->alpha && p->out, "ia_lerp_alpha: null tensor");
    int64_t total = (int64_t)p->B * p->H * p->W * p->C;
    if (total == 0) return 0;
    ia::prof_begin("ia_lerp_alpha", as_stream(stream));
    lerp_alpha_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(*p);
    IA_LAUNCH_CHECK("ia_lerp_alpha");
    return 0;
}


// ------------------------------------------------------------------------------------------------
// output stage: layout_grid with uint8 quantisation, HWC
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layout_grid_u8_kernel(const float* __restrict__ img, int64_t s_b, int64_t s_c, int64_t s_h, int64_t s_w,
                                                             int grid_h, int grid_w, int C, int H, int W, uint8_t* __restrict__ out) {
    const int64_t total = (int64_t)grid_h * H * grid_w * W;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int X = (int)(i % ((int64_t)grid_w * W));
    const int Y = (int)(i / ((int64_t)grid_w * W));
    const int gx = X / W, x = X - gx * W, gy = Y / H, y = Y - gy * H;
    const float* p = img + (int64_t)(gy * grid_w + gx) * s_b + (int64_t)y * s_h + (int64_t)x * s_w;
    uint8_t* o = out + i * C;
    for (int c = 0; c < C; ++c) {
        float v = p[(int64_t)c * s_c] * 127.5f + 128.0f;
        v = fminf(fmaxf(v, 0.0f), 255.0f);
        o[c] = (uint8_t)v;      // truncation toward zero, like tensor.to(torch.uint8)
    }
}

extern "C" int ia_layout_grid_u8(const float* img, int64_t s_b, int64_t s_c, int64_t s_h, int64_t s_w, int32_t grid_h, int32_t grid_w,
                                 int32_t C, int32_t H, int32_t W, uint8_t* out, void* stream) {
    IA_CHECK(img && out && grid_h > 0 && grid_w > 0 && C > 0 && H > 0 && W > 0, "ia_layout_grid_u8: bad arguments");
    const int64_t total = (int64_t)grid_h * H * grid_w * W;
    ia::prof_begin("ia_layout_grid_u8", as_stream(stream));
    layout_grid_u8_kernel<<<(unsigned)cdiv(total, 256), 256, 0, as_stream(stream)>>>(img, s_b, s_c, s_h, s_w, grid_h, grid_w, C, H, W, out);
    IA_LAUNCH_CHECK("ia_layout_grid_u8");
    return 0;
}
