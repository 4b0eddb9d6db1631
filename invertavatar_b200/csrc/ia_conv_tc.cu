// Implicit-GEMM modulated convolution on the Blackwell 5th-generation tensor cores.
//
//   D[pixel, cout] = sum_{tap} sum_{cin} A[pixel + shift(tap), cin] * W[tap][cout, cin]
//
// A = modulated activations, NHWC bf16 hi/lo pair (3-term split: hi*hi + hi*lo + lo*hi, fp32 accumulate in
// TMEM), W = packed weights [tap][Cout_pad][Cin_pad] bf16 hi/lo.  Both operands are K-major, staged by TMA
// (cp.async.bulk.tensor, 128-byte swizzle) straight from their natural layouts: the activation tile of a tap is
// a 4-D box {64 ch, tw, th, nb} whose origin is shifted by the tap offset -- out-of-bounds rows are zero-filled
// by the TMA unit, which implements the convolution padding and the ragged edges of transposed-conv phases.
//
// One CTA computes a 128-pixel x n_tile-cout tile.  Warp roles (192 threads): warp 0 = TMA producer,
// warp 1 = tcgen05.mma issuer (one elected lane), warps 2..5 = epilogue (tcgen05.ld TMEM -> registers ->
// demod / noise / bias / leaky-ReLU / clamp -> fp32 NHWC and/or modulated bf16 hi/lo for the next layer).
//
// Replaces the cuDNN grouped conv2d / conv_transpose2d reached through reference
// torch_utils/ops/conv2d_gradfix.py:127-129 <- conv2d_resample.py:31-43 <- modulated_conv2d
// (training_avatar_texture/networks_stylegan2_new.py:34-91) together with the bias_act that follows it.
#include <cuda.h>
#include <stdlib.h>

#include "ia_common.cuh"

using namespace ia;

int ia_conv_validate(const ia_conv_params* p, const char* who);

namespace {

constexpr int kTileM = 128;
constexpr int kBlockK = 64;           // bf16 elements per k-block = 128 bytes = one swizzle row
constexpr int kMaxStages = 6;
constexpr int kThreads = 192;
constexpr uint32_t kABytes = kTileM * kBlockK * 2;  // 16 KB per A tile

struct TcParams {
    // tile geometry
    int B, GH, GW, th, tw, nb, tiles_x, tiles_y;
    int Cin_blocks;       // Cin_pad / 64
    int Cout, Cout_pad, n_tile, stages, tmem_cols;
    int ntaps; int dy[9]; int dx[9]; int wtap[9];
    int OH, OW, sy, sx, py, px;
    int mode; const float* dcoef; const float* noise; const float* noise_strength; const float* bias;
    long long noise_bstride;
    int act; float alpha, gain, clamp; const float* slope;   // slope: per-channel PReLU slopes (IA_ACT_PRELU)
    ia_emit emit;
    int groups, ipg, n_taps_total; long long noise_gstride;   // grouped launch (see ia_conv_params)
    int nops;             // operand tensors per side: 2 (bf16 hi/lo, 3 MMAs per k-step) or 1 (fp16, 1 MMA)
    uint32_t idesc_fmt;   // A/B format bits of the instruction descriptor
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
            printf("ia_conv_tc: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row atoms of 1024 B (SBO), descriptor version 1.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);        // start address, 16-byte units
    d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset between 8-row atoms
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// CTA-pair variant (tcgen05 cta_group::2): the two CTAs of a cluster (one TPC) execute ONE M=256 MMA per instruction -- 128 pixel
// rows from each CTA's shared memory, the weight tile split between them (each CTA stages and reads half of its rows), 128 TMEM
// lanes of accumulator in each.  Only the leader (cluster rank 0) issues MMAs; both CTAs' TMA loads signal the leader's "full"
// barriers (.cta_group::2 form with the barrier address mapped into the leader), the leader's commits release the slots /
// publish the accumulators in both CTAs (multicast arrive), and both CTAs' epilogue warps release an accumulator buffer on the
// leader's barrier.
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {      // arrives on `bar` (same offset) in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
// One lane of a converged warp (the issuer of the asynchronous tensor-core / TMA instructions).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// MMA from the low words of the two shared-memory descriptors (address >> 4 | LBO) and their common high word.
template <bool PAIR>
__device__ __forceinline__ void umma_issue(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
    if (PAIR) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "setp.ne.b32 p, %5, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
            ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
            "setp.ne.b32 p, %5, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
            ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate) : "memory");
    }
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- the kernel ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
               const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
               const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B-swizzled tiles
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t b_bytes = (uint32_t)p.n_tile * 128u;
    const uint32_t nops = (uint32_t)p.nops;
    const uint32_t stage_bytes = nops * (kABytes + b_bytes);      // [A hi][A lo]?[B hi][B lo]?
    const uint32_t b_off = nops * kABytes;
    const uint32_t bar_base = smem_base + (uint32_t)p.stages * stage_bytes;
    // barriers: full[stages], empty[stages], tmem_full ; then the TMEM base-address slot
    auto full_bar = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (uint32_t)(kMaxStages + s); };
    const uint32_t tmem_full_bar = bar_base + 8u * (2 * kMaxStages);
    const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // tile coordinates
    int tile = blockIdx.x;
    const int txi = tile % p.tiles_x; tile /= p.tiles_x;
    const int tyi = tile % p.tiles_y; tile /= p.tiles_y;
    const int tni = tile;
    const int x0 = txi * p.tw, y0 = tyi * p.th, n0 = tni * p.nb;
    const int col0 = blockIdx.y * p.n_tile;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a_hi); prefetch_tmap(&tm_a_lo); prefetch_tmap(&tm_w_hi); prefetch_tmap(&tm_w_lo);
    }
    if (warp == 2) {  // TMEM allocation (whole warp); the same warp frees it at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int num_kb = p.ntaps * p.Cin_blocks;

    if (warp == 0) {
        // ===================== TMA producer (whole warp runs the loop, one elected lane issues: see conv_tc2_kernel) =====================
        {
            const bool el = elect_one();
            int stage = 0; uint32_t phase = 0;
            for (int t = 0; t < p.ntaps; ++t) {
                const int wrow = ((n0 / p.ipg) * p.n_taps_total + p.wtap[t]) * p.Cout_pad + col0;   // a tile never spans two groups
                for (int kc = 0; kc < p.Cin_blocks; ++kc) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
                    if (el) {
                        mbar_expect_tx(full_bar(stage), stage_bytes);
                        tma_load_4d(sa, &tm_a_hi, full_bar(stage), kc * kBlockK, x0 + p.dx[t], y0 + p.dy[t], n0);
                        if (nops == 2u) tma_load_4d(sa + kABytes, &tm_a_lo, full_bar(stage), kc * kBlockK, x0 + p.dx[t], y0 + p.dy[t], n0);
                        tma_load_2d(sa + b_off, &tm_w_hi, full_bar(stage), kc * kBlockK, wrow);
                        if (nops == 2u) tma_load_2d(sa + b_off + b_bytes, &tm_w_lo, full_bar(stage), kc * kBlockK, wrow);
                    }
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        {
            const bool el = elect_one();
            // instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n_tile
            const uint32_t idesc = (1u << 4) | p.idesc_fmt | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
            const uint32_t desc_hi = (uint32_t)(make_sw128_desc(0u) >> 32);
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
                const uint32_t a_hi = ((sa & 0x3FFFFu) >> 4) | (1u << 16);
                const uint32_t a_lo = a_hi + (kABytes >> 4), b_hi = a_hi + (b_off >> 4), b_lo = b_hi + (b_bytes >> 4);
                if (el) {
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        const uint32_t ko = (uint32_t)(k * 2);  // +32 bytes per 16-element K step
                        umma_issue<false>(tmem_base, a_hi + ko, b_hi + ko, desc_hi, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        if (nops == 2u) {
                            umma_issue<false>(tmem_base, a_hi + ko, b_lo + ko, desc_hi, idesc, 1u);
                            umma_issue<false>(tmem_base, a_lo + ko, b_hi + ko, desc_hi, idesc, 1u);
                        }
                    }
                    umma_commit(empty_bar(stage));   // frees the smem slot once these MMAs have read it
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            if (el) umma_commit(tmem_full_bar);          // accumulator complete
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int lg = warp & 3;                 // TMEM lane group this warp may access
        const int row = lg * 32 + lane;          // tile row == TMEM lane
        const int w_l = row % p.tw;
        const int h_l = (row / p.tw) % p.th;
        const int n_l = row / (p.tw * p.th);
        const int gy = y0 + h_l, gx = x0 + w_l, img = n0 + n_l;
        const int oy = gy * p.sy + p.py, ox = gx * p.sx + p.px;
        const bool valid = gy < p.GH && gx < p.GW && img < p.B && oy < p.OH && ox < p.OW;
        const int64_t pix = ((int64_t)img * p.OH + oy) * p.OW + ox;
        const int grp = img / p.ipg;
        float nz = 0.f;
        if (valid && p.mode == 1 && p.noise)
            nz = p.noise[(int64_t)grp * p.noise_gstride + (int64_t)img * p.noise_bstride + (int64_t)oy * p.OW + ox] * p.noise_strength[grp];
        const float* bias_g = p.bias ? p.bias + (int64_t)grp * p.Cout : nullptr;

        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        for (int c = 0; c < p.n_tile; c += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c, r);
            if (!valid) continue;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int co = col0 + c + q * 4;
                if (co >= p.Cout) break;
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float a = __uint_as_float(r[q * 4 + k]);
                    if (p.mode == 1 && co + k < p.Cout) {
                        if (p.dcoef) a = fmaf(a, p.dcoef[(int64_t)img * p.Cout + co + k], nz); else a += nz;
                        if (bias_g) a += bias_g[co + k];
                        a = act_gain_clamp(a, p.act, p.act == IA_ACT_PRELU ? p.slope[co + k] : p.alpha, p.gain, p.clamp);
                    }
                    v[k] = a;
                }
                emit4(p.emit, img, pix, co, p.Cout, v);
            }
        }
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------------
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

int make_act_map(CUtensorMap* m, const void* ptr, int B, int H, int W, int C_pad, int nb, int th, int tw, int img_rows = 0) {
    EncodeTiledFn enc = get_encode_fn();
    IA_CHECK(enc, "ia_conv_tc: cuTensorMapEncodeTiled unavailable");
    if (img_rows < H) img_rows = H;
    cuuint64_t dims[4] = {(cuuint64_t)C_pad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C_pad * 2, (cuuint64_t)W * C_pad * 2, (cuuint64_t)img_rows * W * C_pad * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)nb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IA_CHECK(r == CUDA_SUCCESS, "ia_conv_tc: activation tensor map encode failed (CUresult %d; B=%d H=%d W=%d C=%d box %d,%d,%d)",
             (int)r, B, H, W, C_pad, nb, th, tw);
    return 0;
}

int make_weight_map(CUtensorMap* m, const void* ptr, int rows, int Cin_pad, int n_tile) {
    EncodeTiledFn enc = get_encode_fn();
    IA_CHECK(enc, "ia_conv_tc: cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)Cin_pad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Cin_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)n_tile};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IA_CHECK(r == CUDA_SUCCESS, "ia_conv_tc: weight tensor map encode failed (CUresult %d; rows=%d Cin=%d n_tile=%d)", (int)r,
             rows, Cin_pad, n_tile);
    return 0;
}

// Choose the 128-row patch {nb, th, tw} (powers of two) that covers the [B][GH][GW] grid with the fewest tiles.
// nb_cap > 0 (grouped launch): a tile may not span two groups, so nb must divide nb_cap (= images per group).
void choose_patch(int B, int GH, int GW, int& nb, int& th, int& tw, int nb_cap = 0) {
    int64_t best = -1;
    int wmax = 1, hmax = 1;
    while (wmax < GW && wmax < 128) wmax <<= 1;
    while (hmax < GH && hmax < 128) hmax <<= 1;
    if (nb_cap > 0) {
        // the patch may have to be larger than the image so that it holds few enough images (out-of-range rows are zero-filled
        // by TMA and masked in the epilogue): let the search run over every power-of-two patch shape
        wmax = 128; hmax = 128;
    }
    for (int w = 1; w <= wmax; w <<= 1) {
        for (int h = 1; h <= hmax && h * w <= 128; h <<= 1) {
            int n = 128 / (w * h);
            if (nb_cap > 0 && (n > nb_cap || nb_cap % n != 0)) continue;
            int64_t tiles = cdiv(GW, w) * cdiv(GH, h) * cdiv(B, n);
            // prefer fewer tiles, then wider rows (longer contiguous runs per TMA box row)
            if (best < 0 || tiles < best || (tiles == best && w > tw)) { best = tiles; nb = n; th = h; tw = w; }
        }
    }
}


// =============================================================================================================
// v2: persistent, 256-pixel tiles, double-buffered TMEM accumulators, row-halo reuse of the activation tile
// =============================================================================================================
// One CTA per SM loops over (pixel tile, cout tile) pairs.  A pixel tile is TH x tw pixels of one image (256 pixels =
// two M=128 halves sharing every weight tile).  Per 32/64-channel k-block and per horizontal tap offset dx the activation
// tile is loaded ONCE with its vertical halo ((TH+halo) x tw pixels); the taps (dy, dx) of that column group read it
// through UMMA descriptors whose start address is shifted by (dy - dy_min) * tw rows -- a multiple of the 8-row swizzle
// atom because tw is a multiple of 8.  Weights stream through their own ring, one (tap, k-block) tile per slot.
// The accumulators of tile i are drained by the epilogue warps while the MMA warp already works on tile i+1.
template <int BK> struct SwzTraits;
template <> struct SwzTraits<64> { static constexpr uint64_t kLayout = 2; static constexpr uint32_t kSbo = 1024; static constexpr CUtensorMapSwizzle kTma = CU_TENSOR_MAP_SWIZZLE_128B; };
template <> struct SwzTraits<32> { static constexpr uint64_t kLayout = 4; static constexpr uint32_t kSbo = 512; static constexpr CUtensorMapSwizzle kTma = CU_TENSOR_MAP_SWIZZLE_64B; };

template <int BK>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(SwzTraits<BK>::kSbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= SwzTraits<BK>::kLayout << 61;
    return d;
}

constexpr int kV2MaxASlots = 6, kV2MaxBSlots = 8;

// One sub-problem of a launch: its taps (grouped by horizontal offset), its grid and where its outputs go.  A plain launch has
// one; a stride-2 transposed convolution runs its four output-parity phases as four sub-problems of ONE persistent launch, so the
// phases share the launch, the ramp/tail and -- tiles being interleaved phase-minor -- the activation tile in L2.
struct Tc2Phase {
    int GH, GW, py, px, tiles_x, tiles_y, ngroups;
    int g_dx[4], g_first[5];
    int t_dyoff[9], t_wtap[9];
};

struct Tc2Params {
    int nph, S_tx, S_ty;           // sub-problems; common tile grid = the largest sub-problem grid
    Tc2Phase ph[4];
    int B, GH, GW, TH, tw, tiles_x, tiles_y, m_tiles, n_tiles, total_tiles;
    int Cin_blocks, Cout, Cout_pad, n_tile, acc_stride;
    int ntaps, ngroups;
    int g_dx[4], g_first[5];
    int t_dyoff[9], t_wtap[9];
    int dy_min;
    uint32_t a_bytes, b_bytes;     // slot stride of ONE (hi or lo) tile, 1024-byte aligned
    uint32_t a_tx, b_tx;           // bytes one TMA box actually delivers
    int a_slots, b_slots;
    int OH, OW, sy, sx, py, px;
    int mode; const float* dcoef; const float* noise; const float* noise_strength; const float* bias;
    long long noise_bstride;
    int act; float alpha, gain, clamp; const float* slope;   // slope: per-channel PReLU slopes (IA_ACT_PRELU)
    ia_emit emit;
    int groups, ipg, n_taps_total; long long noise_gstride;   // grouped launch (see ia_conv_params)
    // CTA-pair variant: the schedule of the plain variant with every (chunk, sub-problem, N tile) list of ic * T tiles padded to an
    // even length, so that the two CTAs of a pair always hold two tiles of the same sub-problem and N tile
    int pp_cum[5];                 // pp_cum[q] = sum_{k<q} ceil(ic * T_k / 2)
    int chunk_pairs, total_pairs;  // n_tiles * pp_cum[nph]; (B / ic) * chunk_pairs
    uint32_t b_half_tx;            // bytes of the half weight tile one CTA of a pair stages
    int dbg_skip_epi;              // IA_DBG_SKIP_EPI=1 (experiments only): drain TMEM but skip the epilogue arithmetic and stores
    // balanced schedule (plain variant): tiles ordered (image chunk, sub-problem, N tile, image, tile) with the sub-problems by
    // descending tap count, only real tiles enumerated -- see decode()
    int ic, chunk_tiles;           // images per chunk; schedule entries of one chunk = ic * n_tiles * ph_cum[nph]
    int ph_cum[5];                 // ph_cum[q] = sum_{k<q} tiles_x[k] * tiles_y[k]
    int epi_vec4;                  // 1: Cout % 4 == 0 and every output pointer / pitch is 16-byte friendly -> epilogue_chunk_v4
    const float* img_prev;         // mode 2: previous-resolution image [B][OH/2][OW/2][Cout] (or null)
    int cat_rows, cat_B;           // > 0: the grid rows are the concatenation of cat_B images of cat_rows rows each (see launch_v2)
    // split-K (plain variant): every tile is computed by `ksplit` CTAs, each over kc_per k-blocks of the input channels; the
    // partial accumulators go through `ws`, the last CTA to arrive (ticket in `cnt`) sums them in split order and runs the epilogue
    int ksplit, kc_per;
    float4* ws; int* cnt;
    int nops;             // operand tensors per side: 2 (bf16 hi/lo, 3 MMAs per k-step) or 1 (fp16, 1 MMA)
    uint32_t idesc_fmt;   // A/B format bits of the instruction descriptor
};


// Inner epilogue loop of one [32 rows][32 channels] chunk: lane = channel.  ACT: 0 = raw accumulator (mode 0),
// IA_ACT_LINEAR / IA_ACT_LRELU specialised, -1 = any activation through the generic switch.
template <int ACT>
__device__ __forceinline__ void epilogue_chunk(const Tc2Params& p, const float* tsm, int lane, uint32_t vmask, int my_pix, float my_nz,
                                               int64_t img_pix0, int64_t img_pix1, int co, bool cvalid, float dc, float bs, float s1v, float s2v) {
    float* o32 = p.emit.out32 ? p.emit.out32 + img_pix0 * p.emit.out32_ld + co : nullptr;
    uint16_t* h1 = p.emit.hi1 ? p.emit.hi1 + img_pix1 * p.emit.c1_pad + co : nullptr;
    uint16_t* l1 = p.emit.hi1 ? p.emit.lo1 + img_pix1 * p.emit.c1_pad + co : nullptr;
    uint16_t* h2 = p.emit.hi2 ? p.emit.hi2 + img_pix0 * p.emit.c2_pad + co : nullptr;
    uint16_t* l2 = p.emit.hi2 ? p.emit.lo2 + img_pix0 * p.emit.c2_pad + co : nullptr;
    const bool has_dc = p.dcoef != nullptr;
    const float gain = p.gain, alpha = p.alpha, clampv = p.clamp;
    const bool do_clamp = clampv >= 0.f;
#pragma unroll 4
    for (int rr = 0; rr < 32; ++rr) {
        if (!((vmask >> rr) & 1u)) continue;
        float a = tsm[rr * 33 + lane];
        const int pix_r = __shfl_sync(0xffffffffu, my_pix, rr);
        if (ACT != 0) {
            const float nz = __shfl_sync(0xffffffffu, my_nz, rr);
            a = has_dc ? fmaf(a, dc, nz) : a + nz;
            a += bs;
            if (ACT == IA_ACT_LRELU) a = (a > 0.f ? a : a * alpha) * gain;
            else if (ACT == IA_ACT_LINEAR) a = a * gain;
            else a = apply_act(a, p.act, p.act == IA_ACT_PRELU ? s2v : alpha) * gain;      // PReLU: the slope travels in the s2 slot
            if (do_clamp) a = fminf(fmaxf(a, -clampv), clampv);
        }
        if (!cvalid) continue;
        if (o32) o32[(int64_t)pix_r * p.emit.out32_ld] = a;
        if (h1) {
            const int64_t o = (int64_t)pix_r * p.emit.c1_pad;
            if (p.emit.fmt1 == IA_OPFMT_F16X1) {
                h1[o] = __half_as_ushort(__float2half_rn(fminf(fmaxf(a * s1v, -65504.f), 65504.f)));
            } else {
                uint16_t h, l;
                split_bf16(a * s1v, h, l);
                h1[o] = h; l1[o] = l;
            }
        }
        if (h2) {
            const int64_t o = (int64_t)pix_r * p.emit.c2_pad;
            if (p.emit.fmt2 == IA_OPFMT_F16X1) {
                h2[o] = __half_as_ushort(__float2half_rn(fminf(fmaxf(a * s2v, -65504.f), 65504.f)));
            } else {
                uint16_t h, l;
                split_bf16(a * s2v, h, l);
                h2[o] = h; l2[o] = l;
            }
        }
    }
}

// Vectorised variant (Cout % 4 == 0, 16-byte aligned outputs): lane = (row_sub = lane/8, c4 = lane%8) owns 4 consecutive
// channels of row 4*i + row_sub, so one pass over a [32 rows][32 channels] chunk is 8 iterations of {LDS.128, 2 shuffles,
// 4-wide epilogue math, 16-byte fp32 / 8-byte bf16 stores} instead of 32 iterations of scalar work -- the scalar loop
// costs ~80 instructions per row and bounds every few-tap launch (profiles/r1_conv_up_phase.txt).  tsm is [32][36] floats.
constexpr int kTsmLd = 36;
// Per-chunk operands of the fused ToRGB contraction: styles and weight rows of this lane's 4 channels, running sums of its 8 rows.
struct RgbLane { float4 s; float4 w[4]; };
template <int ACT, bool O32, bool E1, bool E2, bool RGB = false>
__device__ __forceinline__ void epilogue_chunk_v4(const Tc2Params& p, const float* tsm, int lane, uint32_t vmask, int my_pix, float my_nz,
                                                  float* o32, uint16_t* h1, uint16_t* l1, uint16_t* h2, uint16_t* l2,
                                                  const float4 dc, const float4 bs, const float4 s1, const float4 s2,
                                                  const RgbLane& rg, float (&racc)[8][4]) {
    const int rs = lane >> 3, c4 = lane & 7;
    const float gain = p.gain, alpha = p.alpha, clampv = p.clamp;
    const bool do_clamp = clampv >= 0.f;
#pragma unroll(RGB ? 8 : 2)
    for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + rs;
        const int pix_r = __shfl_sync(0xffffffffu, my_pix, rr);
        float nz = 0.f;
        if (ACT != 0) nz = __shfl_sync(0xffffffffu, my_nz, rr);
        if (!((vmask >> rr) & 1u)) continue;
        float4 a = *reinterpret_cast<const float4*>(tsm + rr * kTsmLd + c4 * 4);
        if (ACT != 0) {
            a.x = fmaf(a.x, dc.x, nz) + bs.x; a.y = fmaf(a.y, dc.y, nz) + bs.y;
            a.z = fmaf(a.z, dc.z, nz) + bs.z; a.w = fmaf(a.w, dc.w, nz) + bs.w;
            if (ACT == IA_ACT_LRELU) {
                a.x = (a.x > 0.f ? a.x : a.x * alpha) * gain; a.y = (a.y > 0.f ? a.y : a.y * alpha) * gain;
                a.z = (a.z > 0.f ? a.z : a.z * alpha) * gain; a.w = (a.w > 0.f ? a.w : a.w * alpha) * gain;
            } else if (ACT == IA_ACT_LINEAR) {
                a.x *= gain; a.y *= gain; a.z *= gain; a.w *= gain;
            } else if (p.act == IA_ACT_PRELU) {    // per-channel slopes travel in the s2 slot (emit 2 is not available with PReLU)
                a.x = (a.x > 0.f ? a.x : a.x * s2.x) * gain; a.y = (a.y > 0.f ? a.y : a.y * s2.y) * gain;
                a.z = (a.z > 0.f ? a.z : a.z * s2.z) * gain; a.w = (a.w > 0.f ? a.w : a.w * s2.w) * gain;
            } else {
                a.x = apply_act(a.x, p.act, alpha) * gain; a.y = apply_act(a.y, p.act, alpha) * gain;
                a.z = apply_act(a.z, p.act, alpha) * gain; a.w = apply_act(a.w, p.act, alpha) * gain;
            }
            if (do_clamp) {
                a.x = fminf(fmaxf(a.x, -clampv), clampv); a.y = fminf(fmaxf(a.y, -clampv), clampv);
                a.z = fminf(fmaxf(a.z, -clampv), clampv); a.w = fminf(fmaxf(a.w, -clampv), clampv);
            }
        }
        if (O32) *reinterpret_cast<float4*>(o32 + (int64_t)pix_r * p.emit.out32_ld) = a;
        if (RGB) {
            // (v * style) * weight, channel by channel, as the modulated 1x1 convolution does (networks_stylegan2_new.py:70-79)
            const float mx = a.x * rg.s.x, my = a.y * rg.s.y, mz = a.z * rg.s.z, mw = a.w * rg.s.w;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                racc[i][j] = fmaf(mw, rg.w[j].w, fmaf(mz, rg.w[j].z, fmaf(my, rg.w[j].y, fmaf(mx, rg.w[j].x, racc[i][j]))));
        }
        if (E1) {
            const int64_t o = (int64_t)pix_r * p.emit.c1_pad;
            store_operand4(p.emit.fmt1, h1 + o, l1 + o, a.x * s1.x, a.y * s1.y, a.z * s1.z, a.w * s1.w);
        }
        if (E2) {
            const int64_t o = (int64_t)pix_r * p.emit.c2_pad;
            store_operand4(p.emit.fmt2, h2 + o, l2 + o, a.x * s2.x, a.y * s2.y, a.z * s2.z, a.w * s2.w);
        }
    }
}

// mode 2: ToRGB tail on a [32 rows][32 channels] chunk -- out = upsample2d(img_prev) + clamp(acc + bias), the arithmetic (and
// the order of operations) of torgb_finish_vec4_kernel in ia_modconv.cu.  `prev` points at this image's img_prev, channel co0.
// Per-row set-up of the ToRGB tail, computed once per tile by the lane that owns the row: pixel offset of the (y0, x0) tap inside
// the half-resolution previous image (coordinates clamped into the image) and a code holding the steps to x1 / y1 (0 when clamped
// away) and the four separable weights as indices into {0, 0.25, 0.75} (0 = tap outside the image: fmaf(0, q, up) == up, so the
// sum equals the skip-the-tap formulation of torgb_finish_vec4_kernel bit for bit).
__device__ __forceinline__ void torgb_row_setup(int y, int x, int h2, int w2, int& o00, int& code) {
    const int my = y >> 1, mx = x >> 1;
    int y0, y1, x0, x1, iy0, iy1, ix0, ix1;      // weight indices: 1 = 0.25, 2 = 0.75
    if (y & 1) { y0 = my; y1 = my + 1; iy0 = 2; iy1 = 1; } else { y0 = my - 1; y1 = my; iy0 = 1; iy1 = 2; }
    if (x & 1) { x0 = mx; x1 = mx + 1; ix0 = 2; ix1 = 1; } else { x0 = mx - 1; x1 = mx; ix0 = 1; ix1 = 2; }
    if (y0 < 0) { y0 = 0; iy0 = 0; }
    if (y1 >= h2) { y1 = h2 - 1; iy1 = 0; }
    if (x0 < 0) { x0 = 0; ix0 = 0; }
    if (x1 >= w2) { x1 = w2 - 1; ix1 = 0; }
    o00 = y0 * w2 + x0;
    code = (x1 - x0) | ((y1 - y0) << 1) | (iy0 << 2) | (iy1 << 4) | (ix0 << 6) | (ix1 << 8);
}

// mode 2: ToRGB tail on a [32 rows][32 channels] chunk -- out = upsample2d(img_prev) + clamp(acc + bias), the arithmetic (and
// the order of operations) of torgb_finish_vec4_kernel in ia_modconv.cu.  `prev` points at this image's img_prev, channel co0.
// The four taps of a row are loaded unconditionally (clamped addresses, zero weights) and back to back: one exposed L2 latency per
// iteration instead of four.
__device__ __forceinline__ void epilogue_chunk_v4_torgb(const Tc2Params& p, const float* tsm, int lane, uint32_t vmask, int my_pix,
                                                        int my_o00, int my_code, float* o32, const float4 bs, const float* prev) {
    const int rs = lane >> 3, c4 = lane & 7;
    const float clampv = p.clamp;
    const int w2 = p.OW >> 1;
#pragma unroll 2
    for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + rs;
        const int pix_r = __shfl_sync(0xffffffffu, my_pix, rr);
        const int o00 = __shfl_sync(0xffffffffu, my_o00, rr);
        const int code = __shfl_sync(0xffffffffu, my_code, rr);
        const bool ok = (vmask >> rr) & 1u;
        float4 q00, q01, q10, q11;
        if (prev) {      // rows that are not valid carry o00 = 0, code = 0: in-bounds loads, never stored
            const float* b00 = prev + (int64_t)o00 * p.Cout;
            const int dx = (code & 1) * p.Cout, dy = ((code >> 1) & 1) * w2 * p.Cout;
            q00 = __ldg(reinterpret_cast<const float4*>(b00));
            q01 = __ldg(reinterpret_cast<const float4*>(b00 + dx));
            q10 = __ldg(reinterpret_cast<const float4*>(b00 + dy));
            q11 = __ldg(reinterpret_cast<const float4*>(b00 + dy + dx));
        }
        float4 v = *reinterpret_cast<const float4*>(tsm + rr * kTsmLd + c4 * 4);
        v.x += bs.x; v.y += bs.y; v.z += bs.z; v.w += bs.w;
        if (clampv >= 0.f) {
            v.x = fminf(fmaxf(v.x, -clampv), clampv); v.y = fminf(fmaxf(v.y, -clampv), clampv);
            v.z = fminf(fmaxf(v.z, -clampv), clampv); v.w = fminf(fmaxf(v.w, -clampv), clampv);
        }
        if (prev) {
            auto wt = [](int idx) { return 0.25f * (float)(idx + (idx >> 1)); };      // 0, 0.25, 0.75
            const float wy0 = wt((code >> 2) & 3), wy1 = wt((code >> 4) & 3), wx0 = wt((code >> 6) & 3), wx1 = wt((code >> 8) & 3);
            float4 up = make_float4(0.f, 0.f, 0.f, 0.f);
            auto tap = [&](const float4& q, float w) {
                up.x = fmaf(w, q.x, up.x); up.y = fmaf(w, q.y, up.y); up.z = fmaf(w, q.z, up.z); up.w = fmaf(w, q.w, up.w);
            };
            tap(q00, wy0 * wx0); tap(q01, wy0 * wx1); tap(q10, wy1 * wx0); tap(q11, wy1 * wx1);
            v.x += up.x; v.y += up.y; v.z += up.z; v.w += up.w;
        }
        if (ok) *reinterpret_cast<float4*>(o32 + (int64_t)pix_r * p.emit.out32_ld) = v;
    }
}

template <int ACT>
__device__ __forceinline__ void epilogue_chunk_v4_dispatch(const Tc2Params& p, const float* tsm, int lane, uint32_t vmask, int my_pix, float my_nz,
                                                           float* o32, uint16_t* h1, uint16_t* l1, uint16_t* h2, uint16_t* l2,
                                                           const float4 dc, const float4 bs, const float4 s1, const float4 s2,
                                                           const RgbLane& rg, float (&racc)[8][4], bool rgb) {
    if (ACT != 0 && rgb) {      // fused ToRGB: only with (optionally) the next convolution's operand -- checked by the launcher
        if (h1) epilogue_chunk_v4<ACT, false, true, false, true>(p, tsm, lane, vmask, my_pix, my_nz, o32, h1, l1, h2, l2, dc, bs, s1, s2, rg, racc);
        else epilogue_chunk_v4<ACT, false, false, false, true>(p, tsm, lane, vmask, my_pix, my_nz, o32, h1, l1, h2, l2, dc, bs, s1, s2, rg, racc);
        return;
    }
    const int sel = (o32 ? 1 : 0) | (h1 ? 2 : 0) | (h2 ? 4 : 0);
#define IA_EPI(O, A, B) epilogue_chunk_v4<ACT, O, A, B>(p, tsm, lane, vmask, my_pix, my_nz, o32, h1, l1, h2, l2, dc, bs, s1, s2, rg, racc)
    switch (sel) {
        case 1: IA_EPI(true, false, false); break;
        case 2: IA_EPI(false, true, false); break;
        case 3: IA_EPI(true, true, false); break;
        case 4: IA_EPI(false, false, true); break;
        case 5: IA_EPI(true, false, true); break;
        case 6: IA_EPI(false, true, true); break;
        case 7: IA_EPI(true, true, true); break;
        default: break;
    }
#undef IA_EPI
}

constexpr int kEpiWarps2 = 8;                       // epilogue warps of the v2 kernel: 2 per TMEM lane quarter (one per half tile)
constexpr int kThreads2 = (2 + kEpiWarps2) * 32;

template <int BK, bool CL>
__global__ void __launch_bounds__(kThreads2, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                const Tc2Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t nops = (uint32_t)p.nops;
    const uint32_t a_slot_bytes = nops * p.a_bytes, b_slot_bytes = nops * p.b_bytes;
    const uint32_t a_base = smem_base;
    const uint32_t b_base = a_base + (uint32_t)p.a_slots * a_slot_bytes;
    const uint32_t bar_base = b_base + (uint32_t)p.b_slots * b_slot_bytes;
    auto a_full = [&](int s) { return bar_base + 8u * (uint32_t)s; };
    auto a_empty = [&](int s) { return bar_base + 8u * (uint32_t)(kV2MaxASlots + s); };
    auto b_full = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kV2MaxASlots + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kV2MaxASlots + kV2MaxBSlots + s); };
    auto t_full = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kV2MaxASlots + 2 * kV2MaxBSlots + s); };
    auto t_empty = [&](int s) { return bar_base + 8u * (uint32_t)(2 * kV2MaxASlots + 2 * kV2MaxBSlots + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (uint32_t)(2 * kV2MaxASlots + 2 * kV2MaxBSlots + 4);
    const uint32_t epi_base = bar_base + 512u;   // 4 x [32][33] fp32 transpose tiles of the epilogue warps

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.a_slots; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < p.b_slots; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        // pair: the leader's t_empty collects the epilogue warps of both CTAs
        for (int s = 0; s < 2; ++s) { mbar_init(t_full(s), 1); mbar_init(t_empty(s), CL ? 2 * kEpiWarps2 : kEpiWarps2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a_hi); prefetch_tmap(&tm_a_lo); prefetch_tmap(&tm_w_hi); prefetch_tmap(&tm_w_lo);
    }
    if (warp == 2) {
        if (CL) {       // the same warp of both CTAs allocates (and frees) the pair's tensor memory
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL) cluster_sync_all();      // the peer's barriers are initialised before anything arrives on them
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const int th_half = p.TH >> 1;
    // tile schedule.  Plain: CTA b takes tiles b, b+grid, ...  Pair: the CTAs (2q, 2q+1) take tile pairs q, q+grid/2, ...; both
    // tiles of a pair belong to the same sub-problem and N tile (every such list is padded to an even length: a padding tile
    // recomputes the list's last tile and stores nothing), so one M=256 MMA serves both.
    const uint32_t crank = CL ? cluster_ctarank() : 0u;
    const int it_first = CL ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int it_step = CL ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int it_end = CL ? p.total_pairs : p.total_tiles * p.ksplit;
    // -> N tile, image, tile origin, sub-problem; `null_tile` (cluster padding) computes but stores nothing, `skip` (a tile
    // position outside this sub-problem's grid) is not executed at all -- every role takes the same decision.
    auto decode = [&](int it, int& n_idx, int& img, int& txi, int& tyi, int& pi, bool& null_tile, bool& skip) {
        if (CL) {
            skip = false;
            const int ch = it / p.chunk_pairs;
            int r = it - ch * p.chunk_pairs;
            pi = 0;
            while (pi + 1 < p.nph && r >= p.n_tiles * p.pp_cum[pi + 1]) ++pi;
            r -= p.n_tiles * p.pp_cum[pi];
            const int Lp = p.pp_cum[pi + 1] - p.pp_cum[pi];
            n_idx = r / Lp; r -= n_idx * Lp;
            const int T = p.ph_cum[pi + 1] - p.ph_cum[pi];
            int lin = 2 * r + (int)crank;
            null_tile = lin >= p.ic * T;
            if (null_tile) lin = p.ic * T - 1;
            const int il = lin / T; lin -= il * T;
            img = ch * p.ic + il;
            tyi = lin / p.ph[pi].tiles_x; txi = lin - tyi * p.ph[pi].tiles_x;
            return;
        }
        // Static round-robin over a cost-sorted tile list: inside a chunk of p.ic images all tiles of the 4-tap sub-problem come
        // first, then the 2-tap ones, then the 1-tap one, so consecutive schedule entries -- which go to consecutive CTAs -- cost
        // the same and every CTA draws (within one tile) the same number of tiles of each cost.  (An interleaved order gives a
        // CTA a random mix: max/mean load 1.2-1.5 on the (H+1)^2 phase grids of the 32^2..128^2 transposed convolutions.)
        null_tile = false; skip = false;
        const int tile_it = p.ksplit > 1 ? it / p.ksplit : it;       // schedule entry -> tile (the splits of a tile are adjacent entries)
        const int ch = tile_it / p.chunk_tiles;
        int r = tile_it - ch * p.chunk_tiles;
        const int per_ph = p.ic * p.n_tiles;
        pi = 0;
        while (pi + 1 < p.nph && r >= per_ph * p.ph_cum[pi + 1]) ++pi;
        r -= per_ph * p.ph_cum[pi];
        const int T = p.ph_cum[pi + 1] - p.ph_cum[pi];
        const int per_n = p.ic * T;
        n_idx = r / per_n; r -= n_idx * per_n;
        const int il = r / T; r -= il * T;
        img = ch * p.ic + il;
        tyi = r / p.ph[pi].tiles_x; txi = r - tyi * p.ph[pi].tiles_x;
    };
    // k-blocks [kc0, kc1) of schedule entry `it` (split-K: a contiguous slice of the input channels; else all of them)
    auto kc_range = [&](int it, int& kc0, int& kc1) {
        if (CL || p.ksplit <= 1) { kc0 = 0; kc1 = p.Cin_blocks; return; }
        const int sp = it % p.ksplit;
        kc0 = sp * p.kc_per;
        kc1 = min(kc0 + p.kc_per, p.Cin_blocks);
    };

    // The producer and the MMA issuer run their loops with the WHOLE warp (every value is warp-uniform and lives in uniform
    // registers) and let one elected lane issue the asynchronous instructions.  Running the loop inside `if (lane == 0)` makes
    // the control flow divergent: every UTCHMMA / UTMALDG operand then sits in a vector register and is moved to the uniform
    // datapath through an R2UR + ELECT + BRA.U.ANY waterfall -- ~430 SASS instructions per weight slot, which bound the MMA
    // warp's issue rate well below the tensor pipe (ncu source view, profiles/r2_conv_mma_issue.txt).
    if (warp == 0) {
        // ===================== TMA producer =====================
        const bool el = elect_one();
        uint32_t as = 0, a_par = 1, bs = 0, b_par = 1;      // ring slot + the parity to wait for on its "empty" barrier
        const uint32_t a_tx = (CL ? 2u : 1u) * nops * p.a_tx, b_tx = CL ? 2u * nops * p.b_half_tx : nops * p.b_tx;
        for (int it = it_first; it < it_end; it += it_step) {
            int n_idx, img, txi, tyi, pi; bool null_tile, skip;
            decode(it, n_idx, img, txi, tyi, pi, null_tile, skip);
            if (skip) continue;
            const Tc2Phase& ph = p.ph[pi];
            const int x0 = txi * p.tw, y0 = tyi * p.TH, col0 = n_idx * p.n_tile;
            const int wrow_base = (img / p.ipg) * p.n_taps_total;
            int kc0, kc1;
            kc_range(it, kc0, kc1);
            for (int kc = kc0; kc < kc1; ++kc) {
                for (int g = 0; g < ph.ngroups; ++g) {
                    mbar_wait(a_empty(as), a_par);
                    const uint32_t sa = a_base + as * a_slot_bytes;
                    if (el) {
                        if (CL) {     // the leader arms its barrier for the bytes of both CTAs; both load their own pixel tile
                            if (crank == 0) mbar_expect_tx(a_full(as), a_tx);
                            const uint32_t lbar = mapa_rank(a_full(as), 0u);
                            tma_load_4d_pair(sa, &tm_a_hi, lbar, kc * BK, x0 + ph.g_dx[g], y0 + p.dy_min, img);
                            if (nops == 2u) tma_load_4d_pair(sa + p.a_bytes, &tm_a_lo, lbar, kc * BK, x0 + ph.g_dx[g], y0 + p.dy_min, img);
                        } else {
                            mbar_expect_tx(a_full(as), a_tx);
                            tma_load_4d(sa, &tm_a_hi, a_full(as), kc * BK, x0 + ph.g_dx[g], y0 + p.dy_min, img);
                            if (nops == 2u) tma_load_4d(sa + p.a_bytes, &tm_a_lo, a_full(as), kc * BK, x0 + ph.g_dx[g], y0 + p.dy_min, img);
                        }
                    }
                    if (++as == (uint32_t)p.a_slots) { as = 0; a_par ^= 1u; }
                    for (int t = ph.g_first[g]; t < ph.g_first[g + 1]; ++t) {
                        mbar_wait(b_empty(bs), b_par);
                        const uint32_t sb = b_base + bs * b_slot_bytes;
                        const int wrow = (wrow_base + ph.t_wtap[t]) * p.Cout_pad + col0;
                        if (el) {
                            if (CL) {     // every CTA stages its half of the weight rows (rank r: rows [r, r + 1) * n_tile / 2)
                                if (crank == 0) mbar_expect_tx(b_full(bs), b_tx);
                                const uint32_t lbar = mapa_rank(b_full(bs), 0u);
                                const int hrow = wrow + (int)crank * (p.n_tile >> 1);
                                tma_load_2d_pair(sb, &tm_w_hi, lbar, kc * BK, hrow);
                                if (nops == 2u) tma_load_2d_pair(sb + p.b_bytes, &tm_w_lo, lbar, kc * BK, hrow);
                            } else {
                                mbar_expect_tx(b_full(bs), b_tx);
                                tma_load_2d(sb, &tm_w_hi, b_full(bs), kc * BK, wrow);
                                if (nops == 2u) tma_load_2d(sb + p.b_bytes, &tm_w_lo, b_full(bs), kc * BK, wrow);
                            }
                        }
                        if (++bs == (uint32_t)p.b_slots) { bs = 0; b_par ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (pair: the leader CTA only) =====================
        if (crank == 0u) {
            const bool el = elect_one();
            // instruction descriptor: D = f32, A/B format, both K-major, N = n_tile, M = 128 (256 across the pair)
            const uint32_t idesc = (1u << 4) | p.idesc_fmt | ((uint32_t)(p.n_tile >> 3) << 17) | ((uint32_t)((CL ? 2 * kTileM : kTileM) >> 4) << 24);
            // shared-memory descriptors: the high word is constant, the low word is (address >> 4) | LBO
            const uint32_t desc_hi = (uint32_t)(make_kmajor_desc<BK>(0u) >> 32);
            auto dlo = [](uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); };
            uint32_t as = 0, a_par = 0, bs = 0, b_par = 0, j = 0;
            const uint32_t row_bytes = (uint32_t)BK * 2u;
            const uint32_t a_lo_off = p.a_bytes >> 4, b_lo_off = p.b_bytes >> 4;       // hi -> lo tile, in descriptor units
            const uint32_t half_off = ((uint32_t)(th_half * p.tw) * row_bytes) >> 4;  // first -> second 128-pixel half
            for (int it = it_first; it < it_end; it += it_step) {
                int mma_pi;
                {
                    int n_idx, img, txi, tyi; bool null_tile, skip;
                    decode(it, n_idx, img, txi, tyi, mma_pi, null_tile, skip);
                    if (skip) continue;
                }
                const Tc2Phase& ph = p.ph[mma_pi];
                const uint32_t acc = j & 1u;
                mbar_wait(t_empty(acc), ((j >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d0 = tmem_base + (acc * 2u + 0u) * (uint32_t)p.acc_stride;
                const uint32_t d1 = tmem_base + (acc * 2u + 1u) * (uint32_t)p.acc_stride;
                uint32_t accum = 0u;
                int kc0, kc1;
                kc_range(it, kc0, kc1);
                for (int kc = kc0; kc < kc1; ++kc) {
                    for (int g = 0; g < ph.ngroups; ++g) {
                        mbar_wait(a_full(as), a_par);
                        const uint32_t a_lo0 = dlo(a_base + as * a_slot_bytes);
                        for (int t = ph.g_first[g]; t < ph.g_first[g + 1]; ++t) {
                            mbar_wait(b_full(bs), b_par);
                            tc_fence_after();
                            const uint32_t b_hi = dlo(b_base + bs * b_slot_bytes);
                            const uint32_t a0_hi = a_lo0 + (((uint32_t)(ph.t_dyoff[t] * p.tw) * row_bytes) >> 4);
                            const uint32_t a1_hi = a0_hi + half_off;
                            if (el) {
#pragma unroll
                                for (int k = 0; k < BK / 16; ++k) {
                                    const uint32_t ko = (uint32_t)(k * 2);          // +32 bytes per 16-element K step
                                    if (nops == 2u) {
                                        umma_issue<CL>(d0, a0_hi + ko, b_hi + ko, desc_hi, idesc, accum);
                                        umma_issue<CL>(d0, a0_hi + ko, b_hi + b_lo_off + ko, desc_hi, idesc, 1u);
                                        umma_issue<CL>(d0, a0_hi + a_lo_off + ko, b_hi + ko, desc_hi, idesc, 1u);
                                        umma_issue<CL>(d1, a1_hi + ko, b_hi + ko, desc_hi, idesc, accum);
                                        umma_issue<CL>(d1, a1_hi + ko, b_hi + b_lo_off + ko, desc_hi, idesc, 1u);
                                        umma_issue<CL>(d1, a1_hi + a_lo_off + ko, b_hi + ko, desc_hi, idesc, 1u);
                                    } else {
                                        umma_issue<CL>(d0, a0_hi + ko, b_hi + ko, desc_hi, idesc, accum);
                                        umma_issue<CL>(d1, a1_hi + ko, b_hi + ko, desc_hi, idesc, accum);
                                    }
                                    accum = 1u;
                                }
                                if (CL) umma_commit_pair(b_empty(bs)); else umma_commit(b_empty(bs));
                            }
                            accum = 1u;
                            if (++bs == (uint32_t)p.b_slots) { bs = 0; b_par ^= 1u; }
                        }
                        if (el) { if (CL) umma_commit_pair(a_empty(as)); else umma_commit(a_empty(as)); }
                        if (++as == (uint32_t)p.a_slots) { as = 0; a_par ^= 1u; }
                    }
                }
                if (el) { if (CL) umma_commit_pair(t_full(acc)); else umma_commit(t_full(acc)); }
                ++j;
            }
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        // TMEM -> registers (thread = pixel row) -> per-warp shared-memory transpose -> lane = output channel, so that the
        // per-channel epilogue operands (demod coefficient, bias, next-layer styles) live in registers and every global
        // store instruction writes one contiguous run (128 B of fp32 or 64 B of bf16) of a single pixel.
        const int lg = warp & 3;
        const int half = (warp - 2) >> 2;           // which 128-pixel half of the tile this warp drains
        float* tsm = reinterpret_cast<float*>(smem_raw + (epi_base - smem_u32(smem_raw))) + (warp - 2) * (32 * kTsmLd);
        const int tw_shift = p.tw == 8 ? 3 : (p.tw == 16 ? 4 : 5);
        uint32_t j = 0;
        for (int it = it_first; it < it_end; it += it_step) {
            int n_idx, img, txi, tyi, pi; bool null_tile, skip;
            decode(it, n_idx, img, txi, tyi, pi, null_tile, skip);
            if (skip) continue;
            const Tc2Phase& ph = p.ph[pi];
            const uint32_t acc = j & 1u;
            const int x0 = txi * p.tw, y0 = tyi * p.TH, col0 = n_idx * p.n_tile;
            mbar_wait(t_full(acc), (j >> 1) & 1u);
            tc_fence_after();
            bool released_any = false;      // split-K releases the TMEM buffer as soon as the partials are parked
            {
                const int row = lg * 32 + lane;
                const int w_l = row & (p.tw - 1);
                const int h_l = (row >> tw_shift) + half * th_half;
                int gy = y0 + h_l;
                const int gx = x0 + w_l;
                // concatenated-rows launch: grid row -> (image, row inside the image); the pixel index then carries the image
                // offset (mode 0 only: no per-image epilogue operand is read) and `img` stays 0
                int cat_img = 0;
                if (p.cat_rows > 0) { cat_img = gy / p.cat_rows; gy -= cat_img * p.cat_rows; }
                const int oy = gy * p.sy + ph.py, ox = gx * p.sx + ph.px;
                const bool valid = !null_tile && cat_img < p.cat_B && gy < ph.GH && gx < ph.GW && oy < p.OH && ox < p.OW;
                const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
                const int my_pix = (cat_img * p.OH + oy) * p.OW + ox;  // pixel index inside the image (fits in int)
                float my_nz = 0.f;
                int my_o00 = 0, my_code = 0;        // ToRGB tail (mode 2): per-row tap set-up, see torgb_row_setup
                if (p.mode == 2 && p.img_prev && valid) torgb_row_setup(oy, ox, p.OH >> 1, p.OW >> 1, my_o00, my_code);
                const int grp = img / p.ipg;
                if (valid && p.mode == 1 && p.noise)
                    my_nz = p.noise[(int64_t)grp * p.noise_gstride + (int64_t)img * p.noise_bstride + my_pix] * p.noise_strength[grp];
                const float* bias_g = p.bias ? p.bias + (int64_t)grp * p.Cout : nullptr;
                const int64_t img_pix0 = (int64_t)img * p.OH * p.OW;
                const int64_t img_pix1 = p.emit.e1_img_pix ? (int64_t)img * p.emit.e1_img_pix : img_pix0;   // emit 1 may be row-padded
                const uint32_t tcol = (acc * 2u + (uint32_t)half) * (uint32_t)p.acc_stride;
                const bool rgb = p.emit.rgb_out != nullptr;
                float racc[8][4];
#pragma unroll
                for (int i = 0; i < 8; ++i) { racc[i][0] = 0.f; racc[i][1] = 0.f; racc[i][2] = 0.f; racc[i][3] = 0.f; }
                bool vmask_done = false;     // split-K: another CTA of this tile runs the epilogue
                // split-K: park this CTA's partial accumulators in the workspace ([tile][split][warp][chunk][8][32 lanes] float4: every
                // store / load instruction moves one contiguous 512-byte run), release the TMEM buffer, take a ticket; the last of the
                // tile's `ksplit` CTAs to arrive sums the partials in split order 0..ksplit-1 (the same order whoever arrives last:
                // the result is deterministic) and runs the epilogue.  Rows are warp-private, so the ticket is per (tile, epilogue
                // warp) and no CTA-wide synchronisation is needed.
                const bool split = !CL && p.ksplit > 1;
                const float4* ws_src = nullptr;
                int64_t ws_sstride = 0;
                if (split) {
                    const int tile_it = it / p.ksplit, sp = it - tile_it * p.ksplit;
                    const int nchunks = p.n_tile >> 5;
                    const int wslot = warp - 2;
                    float4* wsp = p.ws + (((int64_t)tile_it * p.ksplit + sp) * kEpiWarps2 + wslot) * (int64_t)(nchunks * 256);
#pragma unroll 1
                    for (int c = 0; c < p.n_tile; c += 32) {
                        uint32_t r[32];
                        tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + tcol + (uint32_t)c, r);
                        if (vmask == 0u) continue;
                        float4* dst = wsp + (c >> 5) * 256 + lane;
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            dst[q * 32] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(t_empty(acc)) : "memory");     // (split-K: never a pair launch)
                    released_any = true;
                    vmask_done = true;
                    if (vmask != 0u) {      // (vmask depends on the tile only: all splits of a tile take the same branch)
                        __threadfence();
                        __syncwarp();
                        int ticket = 0;
                        int* cnt = p.cnt + tile_it * kEpiWarps2 + wslot;
                        if (lane == 0) ticket = atomicAdd(cnt, 1);
                        ticket = __shfl_sync(0xffffffffu, ticket, 0);
                        if (ticket == p.ksplit - 1) {
                            if (lane == 0) *cnt = 0;       // every split has arrived: leave the counter ready for the next launch
                            __threadfence();
                            vmask_done = false;
                            ws_src = p.ws + (((int64_t)tile_it * p.ksplit) * kEpiWarps2 + wslot) * (int64_t)(nchunks * 256);
                            ws_sstride = (int64_t)kEpiWarps2 * nchunks * 256;
                        }
                    }
                }
                if (!vmask_done) {
#pragma unroll 1
                    for (int c = 0; c < p.n_tile; c += 32) {
                        uint32_t r[32];
                        if (!split) {
                            tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + tcol + (uint32_t)c, r);
                        } else {
                            // split-major: the eight 16-byte loads of one split are in flight together (one exposed L2 latency per
                            // split instead of one per load); every element is still summed in split order 0, 1, ...
                            const float4* src = ws_src + (c >> 5) * 256 + lane;
                            float4 a[8];
#pragma unroll
                            for (int q = 0; q < 8; ++q) a[q] = __ldcg(src + q * 32);
#pragma unroll 2
                            for (int s2 = 1; s2 < p.ksplit; ++s2) {
                                float4 b[8];
#pragma unroll
                                for (int q = 0; q < 8; ++q) b[q] = __ldcg(src + (int64_t)s2 * ws_sstride + q * 32);
#pragma unroll
                                for (int q = 0; q < 8; ++q) { a[q].x += b[q].x; a[q].y += b[q].y; a[q].z += b[q].z; a[q].w += b[q].w; }
                            }
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                r[4 * q] = __float_as_uint(a[q].x); r[4 * q + 1] = __float_as_uint(a[q].y);
                                r[4 * q + 2] = __float_as_uint(a[q].z); r[4 * q + 3] = __float_as_uint(a[q].w);
                            }
                        }
                        if (vmask == 0u || p.dbg_skip_epi) continue;
                        __syncwarp();
                        if (p.epi_vec4) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                *reinterpret_cast<float4*>(tsm + lane * kTsmLd + 4 * q) =
                                    make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
                            __syncwarp();
                            const int co0 = col0 + c + 4 * (lane & 7);
                            const bool cval = co0 < p.Cout;
                            float4 dc4 = make_float4(1.f, 1.f, 1.f, 1.f), bs4 = make_float4(0.f, 0.f, 0.f, 0.f), s14 = dc4, s24 = dc4;
                            if (cval) {
                                if (p.mode == 1) {
                                    if (p.dcoef) dc4 = *reinterpret_cast<const float4*>(p.dcoef + (int64_t)img * p.Cout + co0);
                                    if (bias_g) bs4 = *reinterpret_cast<const float4*>(bias_g + co0);
                                } else if (p.mode == 2) {
                                    if (bias_g) bs4 = *reinterpret_cast<const float4*>(bias_g + co0);
                                }
                                if (p.emit.hi1 && p.emit.s1) s14 = *reinterpret_cast<const float4*>(p.emit.s1 + (int64_t)img * p.Cout + co0);
                                if (p.emit.hi2 && p.emit.s2) s24 = *reinterpret_cast<const float4*>(p.emit.s2 + (int64_t)img * p.Cout + co0);
                                if (p.mode == 1 && p.act == IA_ACT_PRELU) s24 = *reinterpret_cast<const float4*>(p.slope + co0);
                            }
                            // rows of lanes whose channel group is past Cout are masked out, the shuffles inside stay warp-wide
                            const uint32_t vm = cval ? vmask : 0u;
                            float* o32 = p.emit.out32 ? p.emit.out32 + img_pix0 * p.emit.out32_ld + co0 : nullptr;
                            uint16_t* h1 = p.emit.hi1 ? p.emit.hi1 + img_pix1 * p.emit.c1_pad + co0 : nullptr;
                            uint16_t* l1 = p.emit.hi1 ? p.emit.lo1 + img_pix1 * p.emit.c1_pad + co0 : nullptr;
                            uint16_t* h2 = p.emit.hi2 ? p.emit.hi2 + img_pix0 * p.emit.c2_pad + co0 : nullptr;
                            uint16_t* l2 = p.emit.hi2 ? p.emit.lo2 + img_pix0 * p.emit.c2_pad + co0 : nullptr;
                            RgbLane rg;
                            rg.s = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
                            for (int j = 0; j < 4; ++j) rg.w[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (rgb && cval) {
                                if (p.emit.rgb_s) rg.s = *reinterpret_cast<const float4*>(p.emit.rgb_s + (int64_t)img * p.Cout + co0);
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (j < p.emit.rgb_n) rg.w[j] = *reinterpret_cast<const float4*>(p.emit.rgb_w + (int64_t)j * p.Cout + co0);
                            }
                            if (p.mode == 2) {
                                const float* prev = p.img_prev ? p.img_prev + (int64_t)img * (p.OH >> 1) * (p.OW >> 1) * p.Cout + co0 : nullptr;
                                epilogue_chunk_v4_torgb(p, tsm, lane, vm, my_pix, my_o00, my_code, o32, bs4, prev);
                            }
                            else if (p.mode == 0) epilogue_chunk_v4_dispatch<0>(p, tsm, lane, vm, my_pix, my_nz, o32, h1, l1, h2, l2, dc4, bs4, s14, s24, rg, racc, false);
                            else if (p.act == IA_ACT_LRELU) epilogue_chunk_v4_dispatch<IA_ACT_LRELU>(p, tsm, lane, vm, my_pix, my_nz, o32, h1, l1, h2, l2, dc4, bs4, s14, s24, rg, racc, rgb);
                            else if (p.act == IA_ACT_LINEAR && !rgb) epilogue_chunk_v4_dispatch<IA_ACT_LINEAR>(p, tsm, lane, vm, my_pix, my_nz, o32, h1, l1, h2, l2, dc4, bs4, s14, s24, rg, racc, false);
                            else epilogue_chunk_v4_dispatch<-1>(p, tsm, lane, vm, my_pix, my_nz, o32, h1, l1, h2, l2, dc4, bs4, s14, s24, rg, racc, rgb);
                            continue;
                        }
#pragma unroll
                        for (int q = 0; q < 32; ++q) tsm[lane * 33 + q] = __uint_as_float(r[q]);
                        __syncwarp();
                        const int co = col0 + c + lane;
                        const bool cvalid = co < p.Cout;
                        float dc = 1.f, bs = 0.f, s1v = 1.f, s2v = 1.f;
                        if (cvalid) {
                            if (p.mode == 1) { if (p.dcoef) dc = p.dcoef[(int64_t)img * p.Cout + co]; if (bias_g) bs = bias_g[co]; }
                            if (p.emit.hi1 && p.emit.s1) s1v = p.emit.s1[(int64_t)img * p.Cout + co];
                            if (p.emit.hi2 && p.emit.s2) s2v = p.emit.s2[(int64_t)img * p.Cout + co];
                            if (p.mode == 1 && p.act == IA_ACT_PRELU) s2v = p.slope[co];
                        }
                        if (p.mode == 0) epilogue_chunk<0>(p, tsm, lane, vmask, my_pix, my_nz, img_pix0, img_pix1, co, cvalid, dc, bs, s1v, s2v);
                        else if (p.act == IA_ACT_LRELU) epilogue_chunk<IA_ACT_LRELU>(p, tsm, lane, vmask, my_pix, my_nz, img_pix0, img_pix1, co, cvalid, dc, bs, s1v, s2v);
                        else if (p.act == IA_ACT_LINEAR) epilogue_chunk<IA_ACT_LINEAR>(p, tsm, lane, vmask, my_pix, my_nz, img_pix0, img_pix1, co, cvalid, dc, bs, s1v, s2v);
                        else epilogue_chunk<-1>(p, tsm, lane, vmask, my_pix, my_nz, img_pix0, img_pix1, co, cvalid, dc, bs, s1v, s2v);
                    }
                }
                if (rgb && vmask != 0u && !vmask_done) {
                    // sum the partial contractions of the 8 lanes that share a row (their 4-channel groups), then lane c4 == 0 adds
                    // this N tile's share into the zero-filled output (one add per N tile: two tiles commute, the result is exact)
                    const int rs = lane >> 3;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float v = racc[i][j];
                            v += __shfl_xor_sync(0xffffffffu, v, 1);
                            v += __shfl_xor_sync(0xffffffffu, v, 2);
                            v += __shfl_xor_sync(0xffffffffu, v, 4);
                            racc[i][j] = v;
                        }
                        const int rr = 4 * i + rs;
                        const int pix_r = __shfl_sync(0xffffffffu, my_pix, rr);
                        if ((lane & 7) == 0 && ((vmask >> rr) & 1u)) {
                            float* o = p.emit.rgb_out + (img_pix0 + pix_r) * p.emit.rgb_n;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (j < p.emit.rgb_n) atomicAdd(o + j, racc[i][j]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (!released_any && lane == 0) {
                if (CL) mbar_arrive_cluster(mapa_rank(t_empty(acc), 0u));      // the leader's MMA warp waits for both CTAs' epilogues
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(t_empty(acc)) : "memory");
            }
            ++j;
        }
    }

    __syncthreads();
    if (CL) cluster_sync_all();      // no CTA leaves while its peer may still read its tiles or arrive on its barriers
    if (warp == 2) {
        tc_fence_after();
        if (CL) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <int BK>
int make_act_map2(CUtensorMap* m, const void* ptr, int B, int H, int W, int C_pad, int rows, int tw, int img_rows = 0) {
    EncodeTiledFn enc = get_encode_fn();
    IA_CHECK(enc, "ia_conv_tc: cuTensorMapEncodeTiled unavailable");
    if (img_rows < H) img_rows = H;      // rows between the starts of consecutive images (row-padded operand layouts)
    cuuint64_t dims[4] = {(cuuint64_t)C_pad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C_pad * 2, (cuuint64_t)W * C_pad * 2, (cuuint64_t)img_rows * W * C_pad * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)tw, (cuuint32_t)rows, 1u};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, SwzTraits<BK>::kTma, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IA_CHECK(r == CUDA_SUCCESS, "ia_conv_tc(v2): activation tensor map encode failed (CUresult %d; B=%d H=%d W=%d C=%d box %d x %d)",
             (int)r, B, H, W, C_pad, rows, tw);
    return 0;
}

template <int BK>
int make_weight_map2(CUtensorMap* m, const void* ptr, int rows, int Cin_pad, int n_tile) {
    EncodeTiledFn enc = get_encode_fn();
    IA_CHECK(enc, "ia_conv_tc: cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[2] = {(cuuint64_t)Cin_pad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Cin_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)n_tile};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, SwzTraits<BK>::kTma, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IA_CHECK(r == CUDA_SUCCESS, "ia_conv_tc(v2): weight tensor map encode failed (CUresult %d)", (int)r);
    return 0;
}

int g_sm_count = 0;

template <int BK>
int launch_v2(const ia_conv_params* const* ps, int nph, void* stream) {
    const ia_conv_params* p = ps[0];       // operands, epilogue and emitted tensors are shared by all sub-problems
    Tc2Params t;
    memset(&t, 0, sizeof(t));
    t.B = p->B;
    int GHm = 0, GWm = 0;
    for (int i = 0; i < nph; ++i) { GHm = ps[i]->GH > GHm ? ps[i]->GH : GHm; GWm = ps[i]->GW > GWm ? ps[i]->GW : GWm; }
    t.GH = GHm; t.GW = GWm;
    // Concatenated rows (ia_conv_params.a_img_rows == H + 1): the operand is [B][H+1][W][C] with a zero row after every image,
    // so the B phase grids of (H+1) rows each form one tall grid of B*(H+1) rows whose taps never see a neighbouring image's
    // data (row -1 / row H of an image is a zero row, or outside the tensor).  Tiles then run across image boundaries:
    // ceil(B*(H+1)/TH) tile rows instead of B*ceil((H+1)/TH).
    const bool cat = nph > 1 && p->a_img_rows == p->H + 1 && p->mode == 0 && p->groups <= 1 && GHm == p->H + 1;
    const int rowsB = cat ? p->B * (p->H + 1) : 0;           // rows of the concatenated grid
    t.cat_rows = cat ? p->H + 1 : 0; t.cat_B = cat ? p->B : 1;
    if (cat) t.B = 1;
    // tile geometry: 256 pixels as TH x tw with tw in {8,16,32}; fewest tiles wins, ties -> taller tiles (smaller halo share)
    {
        int64_t best = -1;
        const int cand[3][2] = {{32, 8}, {16, 16}, {8, 32}};
        for (int i = 0; i < 3; ++i) {
            int64_t tiles = 0;
            for (int q = 0; q < nph; ++q) tiles += cdiv(cat ? rowsB : ps[q]->GH, cand[i][0]) * cdiv(ps[q]->GW, cand[i][1]);
            if (best < 0 || tiles < best) { best = tiles; t.TH = cand[i][0]; t.tw = cand[i][1]; }
        }
    }
    t.nph = nph;
    t.S_tx = (int)cdiv(GWm, t.tw); t.S_ty = (int)cdiv(cat ? rowsB : GHm, t.TH);
    t.tiles_x = t.S_tx; t.tiles_y = t.S_ty;
    t.m_tiles = t.S_tx * t.S_ty * nph * t.B;           // schedule entries (positions outside a sub-problem's grid are skipped)
    int64_t real_tiles = 0;
    for (int q = 0; q < nph; ++q) real_tiles += cdiv(cat ? rowsB : ps[q]->GH, t.TH) * cdiv(ps[q]->GW, t.tw) * t.B;
    t.Cin_blocks = p->Cin_pad / BK;
    t.Cout = p->Cout; t.Cout_pad = p->Cout_pad;
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    int n_tile = 32;
    // Split-K (needs the caller's workspace + ticket counters): a launch with fewer tiles than SMs keeps the WIDEST N tile --
    // every activation tile is then amortised over 128 output channels instead of being re-read by 4 narrow-tile CTAs -- and is
    // spread over the idle SMs by splitting the input channels instead.  Measured motivation (round 1): 1024->512 @16^2 x 4
    // images ran as 64 CTAs of N=32 with a serial chain of 288 (tap, k-block) stages each: 203 us, 48 TF/s.
    int ksplit = 1;
    static int splitk_max = -1;    // IA_CONV_SPLITK: 0 = off, n = upper bound on the split factor (default 16)
    if (splitk_max < 0) { const char* ev = getenv("IA_CONV_SPLITK"); splitk_max = ev ? atoi(ev) : 16; }
    const bool can_split = splitk_max > 1 && p->splitk_ws && p->splitk_counters && !p->emit.rgb_out;
    for (int cand = 128; cand >= 32; cand -= 32) {
        if (p->Cout_pad % cand) continue;
        n_tile = cand;
        // widest N tile that still gives at least `min_tiles` CTAs: a narrow tile amortises every activation tile over few
        // output channels and leaves too little MMA work per pipeline stage to cover the TMA latency (measured, 512->512 at
        // 16x16 x 24 images: 96 CTAs of N=128 take 126 us, 384 CTAs of N=32 take 213 us); IA_CONV_MIN_TILES overrides
        static int min_tiles = -1;
        if (min_tiles < 0) { const char* ev = getenv("IA_CONV_MIN_TILES"); min_tiles = ev ? atoi(ev) : 48; }
        if (can_split) break;
        if (real_tiles * (p->Cout_pad / cand) >= min_tiles) break;
        // fused ToRGB: every N tile adds its share into the same pixels; with at most two tiles the two adds commute and the
        // result does not depend on their order -> take the widest tile whatever the grid size
        if (p->emit.rgb_out) break;
    }
    if (can_split) {
        const int64_t tiles = real_tiles * (p->Cout_pad / n_tile);
        const int cin_blocks = p->Cin_pad / BK;
        static int min_kc = -1;        // IA_CONV_SPLITK_MIN_KC: fewest k-blocks a split may get (default 2 = 64 channels x all taps)
        if (min_kc < 0) { const char* ev = getenv("IA_CONV_SPLITK_MIN_KC"); min_kc = ev ? atoi(ev) : 2; if (min_kc < 1) min_kc = 1; }
        if (tiles * 4 < (int64_t)g_sm_count * 3 && cin_blocks >= 2 * min_kc) {
            int want = (int)(g_sm_count / tiles);                         // splits that fill the machine once
            if (want > splitk_max) want = splitk_max;
            if (want > cin_blocks / min_kc) want = cin_blocks / min_kc;
            // workspace / counter capacity
            const int64_t per_split = tiles * 256 * (int64_t)n_tile * 4;
            if (per_split > 0 && want > p->splitk_ws_bytes / per_split) want = (int)(p->splitk_ws_bytes / per_split);
            if (tiles * kEpiWarps2 > p->splitk_n_counters) want = 1;
            if (want > 1) {
                const int kc_per = (cin_blocks + want - 1) / want;
                ksplit = (cin_blocks + kc_per - 1) / kc_per;                 // every split non-empty
                t.kc_per = kc_per;
            }
        }
        if (ksplit <= 1) {
            // not split after all: fall back to the narrow-tile rule
            for (int cand = 128; cand >= 32; cand -= 32) {
                if (p->Cout_pad % cand) continue;
                n_tile = cand;
                static int min_tiles2 = -1;
                if (min_tiles2 < 0) { const char* ev = getenv("IA_CONV_MIN_TILES"); min_tiles2 = ev ? atoi(ev) : 48; }
                if (real_tiles * (p->Cout_pad / cand) >= min_tiles2) break;
            }
        }
    }
    t.ksplit = ksplit;
    if (ksplit <= 1) { t.ksplit = 1; t.kc_per = p->Cin_pad / BK; }
    // CTA-pair launch (tcgen05 cta_group::2; IA_CONV_PAIR=0 disables): M = 256 MMAs over two CTAs' pixel tiles with the weight tile
    // split between them -- half the weight bytes staged and read per CTA.  Needs at least one tile pair per SM pair, one weight
    // group, no split-K.
    // (read per launch, not cached: the bit-identity test toggles them inside one process)
    int pair_env = 1, pair_min_tiles = 4;      // IA_CONV_PAIR_MIN_TILES: fewest (M tile, N tile) entries of a pair launch, in units of SMs x 1/4
    { const char* ev = getenv("IA_CONV_PAIR"); if (ev) pair_env = atoi(ev); }
    { const char* ev = getenv("IA_CONV_PAIR_MIN_TILES"); if (ev) pair_min_tiles = atoi(ev); }
    bool use_pair = pair_env != 0 && p->groups <= 1 && t.ksplit <= 1 && real_tiles * (p->Cout_pad / n_tile) * 4 >= (int64_t)g_sm_count * pair_min_tiles;
    t.ws = reinterpret_cast<float4*>(p->splitk_ws); t.cnt = p->splitk_counters;
    t.n_tile = n_tile; t.acc_stride = 128;
    t.n_tiles = p->Cout_pad / n_tile;
    t.total_tiles = t.m_tiles * t.n_tiles;
    // taps of every sub-problem grouped by dx; one activation box (common dy range) serves all of them
    int dy_min = 1 << 30, dy_max = -(1 << 30);
    for (int q = 0; q < nph; ++q)
        for (int i = 0; i < ps[q]->ntaps; ++i) {
            dy_min = ps[q]->dy[i] < dy_min ? ps[q]->dy[i] : dy_min;
            dy_max = ps[q]->dy[i] > dy_max ? ps[q]->dy[i] : dy_max;
        }
    t.dy_min = dy_min;
    const int halo = dy_max - dy_min;
    int taps_sum = 0, groups_sum = 0;
    for (int q = 0; q < nph; ++q) {
        const ia_conv_params* pq = ps[q];
        Tc2Phase& ph = t.ph[q];
        ph.GH = pq->GH; ph.GW = pq->GW; ph.py = pq->py; ph.px = pq->px;
        ph.tiles_x = (int)cdiv(pq->GW, t.tw); ph.tiles_y = (int)cdiv(cat ? rowsB : pq->GH, t.TH);
        ph.ngroups = 0;
        int nt = 0;
        bool used[9] = {false, false, false, false, false, false, false, false, false};
        for (int i = 0; i < pq->ntaps; ++i) {
            if (used[i]) continue;
            IA_CHECK(ph.ngroups < 4, "ia_conv_tc(v2): more than 4 distinct horizontal tap offsets");
            ph.g_dx[ph.ngroups] = pq->dx[i];
            ph.g_first[ph.ngroups] = nt;
            for (int k = i; k < pq->ntaps; ++k)
                if (!used[k] && pq->dx[k] == pq->dx[i]) { used[k] = true; ph.t_dyoff[nt] = pq->dy[k] - dy_min; ph.t_wtap[nt] = pq->wtap[k]; ++nt; }
            ++ph.ngroups;
        }
        ph.g_first[ph.ngroups] = nt;
        taps_sum += pq->ntaps; groups_sum += ph.ngroups;
    }
    t.ntaps = taps_sum; t.ngroups = groups_sum;      // (ring-depth balance below; the kernel reads the per-phase tables)
    const int a_rows = (t.TH + halo) * t.tw;
    t.a_tx = (uint32_t)a_rows * BK * 2u;
    t.a_bytes = (t.a_tx + 1023u) & ~1023u;
    t.b_tx = (uint32_t)n_tile * BK * 2u;
    t.b_half_tx = t.b_tx >> 1;
    if (use_pair) t.b_tx = t.b_half_tx;      // a pair CTA's weight slot holds its half of the tile
    t.b_bytes = (t.b_tx + 1023u) & ~1023u;
    const uint32_t kEpiBytes = (uint32_t)kEpiWarps2 * 32u * (uint32_t)kTsmLd * 4u;
    const uint32_t budget = 227u * 1024u - 1024u - 512u - kEpiBytes;
    const uint32_t nops = p->op_fmt == IA_OPFMT_F16X1 ? 1u : 2u;      // operand tensors per side (a slot holds hi [+ lo])
    t.nops = (int)nops;
    t.idesc_fmt = p->op_fmt == IA_OPFMT_F16X1 ? 0u : ((1u << 7) | (1u << 10));     // A/B format: F16 = 0, BF16 = 1
    // at least 2 + 2 slots; then spend the rest alternately (weights first: they turn over once per tap)
    t.a_slots = 2; t.b_slots = 2;
    // a wide k-block with a tall halo tile may not leave room for 2 + 2 slots at the widest N tile: narrow the N tile
    while (nops * (2u * t.a_bytes + 2u * t.b_bytes) > budget && n_tile > 32) {
        int next = n_tile - 32;
        while (next > 32 && p->Cout_pad % next) next -= 32;
        n_tile = next;
        t.ksplit = 1; t.kc_per = p->Cin_pad / BK;      // (the split plan was sized for the wider tile)
        t.n_tile = n_tile; t.n_tiles = p->Cout_pad / n_tile; t.total_tiles = t.m_tiles * t.n_tiles;
        t.b_tx = (uint32_t)n_tile * BK * 2u;
        t.b_half_tx = t.b_tx >> 1;
        if (use_pair) t.b_tx = t.b_half_tx;
        t.b_bytes = (t.b_tx + 1023u) & ~1023u;
    }
    IA_CHECK(nops * (2u * t.a_bytes + 2u * t.b_bytes) <= budget, "ia_conv_tc(v2): tile does not fit shared memory");
    // Spend the rest so that both rings hold about the same number of k-blocks of work: a k-block consumes `ngroups`
    // activation slots and `ntaps` weight slots.  Few-tap launches (transposed-conv phases, 1x1) therefore get a deep
    // activation ring -- their activation tiles stream from DRAM and two slots in flight bound them by latency
    // (measured: 3.85 TB/s aggregate with 2 slots) -- while 3x3 layers keep the deep weight ring.
    for (;;) {
        const uint32_t used_b = nops * ((uint32_t)t.a_slots * t.a_bytes + (uint32_t)t.b_slots * t.b_bytes);
        const bool a_fits = t.a_slots < kV2MaxASlots && used_b + nops * t.a_bytes <= budget;
        const bool b_fits = t.b_slots < kV2MaxBSlots && used_b + nops * t.b_bytes <= budget;
        if (!a_fits && !b_fits) break;
        // depth_a = a_slots / ngroups, depth_b = b_slots / ntaps; grow the shallower ring (ties -> activations)
        const bool want_a = (int64_t)t.a_slots * t.ntaps <= (int64_t)t.b_slots * t.ngroups;
        if ((want_a && a_fits) || !b_fits) ++t.a_slots; else ++t.b_slots;
    }
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    // Balanced schedule of the plain variant (decode()).  Images per chunk: a chunk's activations should stay in L2 while its
    // sub-problems are swept one after the other (each sweep re-reads them), so only chunk sizes up to ~72 MB of operand bytes
    // are considered (always at least one image); among those the one with the smallest maximum per-CTA load (in tap units,
    // evaluated exactly for the round-robin assignment) wins, ties go to the smaller chunk.
    {
        t.ph_cum[0] = 0;
        for (int q = 0; q < nph; ++q) t.ph_cum[q + 1] = t.ph_cum[q] + t.ph[q].tiles_x * t.ph[q].tiles_y;
        const int64_t total = (int64_t)t.ph_cum[nph] * t.B * t.n_tiles;
        const int g = total < g_sm_count ? (int)total : g_sm_count;
        const double img_bytes = (double)p->H * p->W * p->Cin_pad * 4.0;
        int best_ic = 1; int64_t best_max = -1;
        for (int ic = 1; ic <= t.B && nph > 1; ++ic) {
            if (t.B % ic) continue;
            if (ic > 1 && ic * img_bytes > 72e6) break;
            int64_t load[1024];
            const int gg = g < 1024 ? g : 1024;
            for (int i = 0; i < gg; ++i) load[i] = 0;
            int64_t pos = 0;
            for (int ch = 0; ch < t.B / ic; ++ch)
                for (int q = 0; q < nph; ++q) {
                    const int64_t L = (int64_t)ic * t.n_tiles * (t.ph_cum[q + 1] - t.ph_cum[q]);
                    const int64_t full = L / gg, rem = L % gg;
                    const int64_t w = ps[q]->ntaps;
                    for (int i = 0; i < gg; ++i) load[i] += full * w;
                    for (int64_t i = 0; i < rem; ++i) load[(pos + i) % gg] += w;
                    pos += L;
                }
            int64_t mx = 0;
            for (int i = 0; i < gg; ++i) mx = load[i] > mx ? load[i] : mx;
            if (best_max < 0 || mx < best_max) { best_max = mx; best_ic = ic; }
        }
        if (nph == 1) best_ic = use_pair ? t.B : 1;      // (pair: one list per N tile, so at most one padding tile each)
        t.ic = best_ic;
        t.chunk_tiles = t.ic * t.n_tiles * t.ph_cum[nph];
        t.total_tiles = (int)total;
        t.pp_cum[0] = 0;
        for (int q = 0; q < nph; ++q) t.pp_cum[q + 1] = t.pp_cum[q] + (t.ic * (t.ph_cum[q + 1] - t.ph_cum[q]) + 1) / 2;
        t.chunk_pairs = t.n_tiles * t.pp_cum[nph];
        t.total_pairs = (t.B / t.ic) * t.chunk_pairs;
    }
    t.OH = p->OH; t.OW = p->OW; t.sy = p->sy; t.sx = p->sx; t.py = p->py; t.px = p->px;
    t.mode = p->mode; t.dcoef = p->dcoef; t.noise = p->noise; t.noise_strength = p->noise_strength; t.bias = p->bias;
    t.noise_bstride = p->noise_bstride;
    t.act = p->act; t.alpha = p->alpha; t.gain = p->gain; t.clamp = p->clamp; t.slope = p->slope;
    { static int dbg = -1; if (dbg < 0) { const char* ev = getenv("IA_DBG_SKIP_EPI"); dbg = ev ? atoi(ev) : 0; } t.dbg_skip_epi = dbg; }
    t.emit = p->emit;
    t.groups = p->groups > 1 ? p->groups : 1; t.ipg = p->groups > 1 ? p->imgs_per_group : (p->B > 0 ? p->B : 1);
    t.n_taps_total = p->n_taps_total; t.noise_gstride = p->groups > 1 ? p->noise_gstride : 0;
    {
        auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
        const ia_emit& e = p->emit;
        t.epi_vec4 = (p->Cout % 4 == 0) && (!e.out32 || (al16(e.out32) && e.out32_ld % 4 == 0)) &&
                     (!e.hi1 || (al16(e.hi1) && al16(e.lo1) && e.c1_pad % 8 == 0)) && (!e.hi2 || (al16(e.hi2) && al16(e.lo2) && e.c2_pad % 8 == 0)) &&
                     (!p->dcoef || al16(p->dcoef)) && (!p->bias || al16(p->bias)) && (!e.s1 || al16(e.s1)) && (!e.s2 || al16(e.s2));
        IA_CHECK(p->mode != 2 || (t.epi_vec4 && (!p->img_prev || al16(p->img_prev))),
                 "ia_conv_tc: the fused ToRGB tail (mode 2) needs the vectorised epilogue (aligned operands, Cout %% 4 == 0)");
        t.img_prev = p->img_prev;
        IA_CHECK(!e.rgb_out || t.n_tiles <= 2, "ia_conv_tc: the fused ToRGB contraction allows at most two N tiles (Cout_pad %d, N tile %d)", p->Cout_pad, t.n_tile);
        IA_CHECK(!e.rgb_out || (t.epi_vec4 && (!e.rgb_s || al16(e.rgb_s)) && al16(e.rgb_w)),
                 "ia_conv_tc: the fused ToRGB contraction needs the vectorised epilogue (aligned operands, Cout %% 4 == 0)");
        static int force_scalar = -1;
        if (force_scalar < 0) { const char* ev = getenv("IA_CONV_EPI_SCALAR"); force_scalar = (ev && atoi(ev)) ? 1 : 0; }
        if (force_scalar && !e.rgb_out && p->mode != 2) t.epi_vec4 = 0;
    }

    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    // concatenated rows: one image of B*(H+1) rows (the trailing zero row of the last image included); else B images whose
    // starts are a_img_rows rows apart
    const int mapB = cat ? 1 : p->B, mapH = cat ? rowsB : p->H, mapR = cat ? rowsB : p->a_img_rows;
    if (int rc = make_act_map2<BK>(&ma_hi, p->a_hi, mapB, mapH, p->W, p->Cin_pad, t.TH + halo, t.tw, mapR)) return rc;
    if (int rc = make_act_map2<BK>(&ma_lo, nops == 2u ? p->a_lo : p->a_hi, mapB, mapH, p->W, p->Cin_pad, t.TH + halo, t.tw, mapR)) return rc;
    const int wrows = t.groups * p->n_taps_total * p->Cout_pad;
    const int w_box = use_pair ? n_tile / 2 : n_tile;
    if (int rc = make_weight_map2<BK>(&mw_hi, p->w_hi, wrows, p->Cin_pad, w_box)) return rc;
    if (int rc = make_weight_map2<BK>(&mw_lo, nops == 2u ? p->w_lo : p->w_hi, wrows, p->Cin_pad, w_box)) return rc;

    const size_t smem = (size_t)nops * ((size_t)t.a_slots * t.a_bytes + (size_t)t.b_slots * t.b_bytes) + 1024 + 512 + kEpiBytes;
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    IA_CHECK(e == cudaSuccess, "ia_conv_tc(v2): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    e = cudaFuncSetAttribute(conv_tc2_kernel<BK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    IA_CHECK(e == cudaSuccess, "ia_conv_tc(v2): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    if (use_pair) {
        const int sm_even = g_sm_count & ~1;
        const int grid = 2 * t.total_pairs < sm_even ? 2 * t.total_pairs : sm_even;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid, 1, 1);
        cfg.blockDim = dim3(kThreads2, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = as_stream(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        ia::prof_begin(ia::prof_detail_name("ia_conv_tc", taps_sum, GHm, GWm, p->Cin_pad, p->Cout), as_stream(stream));
        cudaError_t le = cudaLaunchKernelEx(&cfg, conv_tc2_kernel<BK, true>, ma_hi, ma_lo, mw_hi, mw_lo, t);
        IA_CHECK(le == cudaSuccess, "ia_conv_tc(v2): pair launch failed: %s", cudaGetErrorString(le));
        IA_LAUNCH_CHECK("ia_conv_tc");
        return 0;
    }
    const int64_t real_total = real_tiles * t.n_tiles * t.ksplit;
    const int grid = real_total < g_sm_count ? (int)real_total : g_sm_count;
    ia::prof_begin(ia::prof_detail_name("ia_conv_tc", taps_sum, GHm, GWm, p->Cin_pad, p->Cout), as_stream(stream));
    conv_tc2_kernel<BK, false><<<grid, kThreads2, smem, as_stream(stream)>>>(ma_hi, ma_lo, mw_hi, mw_lo, t);
    IA_LAUNCH_CHECK("ia_conv_tc");
    return 0;
}

}  // namespace

extern "C" int ia_conv_tc(const ia_conv_params* p, void* stream) {
    if (int rc = ia_conv_validate(p, "ia_conv_tc")) return rc;
    IA_CHECK((reinterpret_cast<uintptr_t>(p->a_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->a_lo) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(p->w_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->w_lo) & 15) == 0,
             "ia_conv_tc: operands must be 16-byte aligned");     // (a NULL lo pointer of an F16X1 operand passes)
    // v2 (persistent, 256-pixel tiles, halo reuse) needs images of at least one 128-pixel half tile; the low-resolution
    // layers (4x4, 8x8 and their transposed-conv phase grids) pack several images into a tile with the v1 kernel below.
    {
        static int mode = -1;   // IA_CONV_TC: 0 = v1 only, 32 / 64 = v2 with that k-block (default 32)
        static int mode_few = -1;   // IA_CONV_TC_FEW: k-block for launches with <= 4 taps (transposed-conv phases, 1x1): these
                                    // issue few MMAs per activation tile, so the TMA row rate (one request per 64/128-byte
                                    // row) bounds them and 128-byte rows (k-block 64) halve the request count
        if (mode < 0) { const char* e = getenv("IA_CONV_TC"); mode = e ? atoi(e) : 32; }
        if (mode_few < 0) { const char* e = getenv("IA_CONV_TC_FEW"); mode_few = e ? atoi(e) : 32; }
        static int mode_f16 = -1;   // IA_CONV_TC_F16: k-block of single-pass fp16 launches.  One operand tensor per side: a 64-channel
                                    // k-block costs the shared memory of a 32-channel bf16 hi/lo one, and its 128-byte rows halve the
                                    // number of TMA row requests per byte
        if (mode_f16 < 0) { const char* e = getenv("IA_CONV_TC_F16"); mode_f16 = e ? atoi(e) : 64; }
        if (mode != 0 && p->GH * p->GW >= 128 && p->GW >= 8) {
            int bk = (p->ntaps <= 4) ? mode_few : mode;
            if (p->op_fmt == IA_OPFMT_F16X1) bk = mode_f16;
            return bk == 64 ? launch_v2<64>(&p, 1, stream) : launch_v2<32>(&p, 1, stream);
        }
    }
    IA_CHECK(!p->emit.rgb_out, "ia_conv_tc: the fused ToRGB contraction needs the persistent kernel (images of >= 128 pixels, width >= 8)");
    IA_CHECK(p->mode != 2, "ia_conv_tc: the fused ToRGB tail (mode 2) needs the persistent kernel (images of >= 128 pixels, width >= 8)");
    IA_CHECK(p->emit.e1_img_pix == 0, "ia_conv_tc: a row-padded emit-1 layout needs the persistent kernel (images of >= 128 pixels, width >= 8)");
    TcParams t;
    memset(&t, 0, sizeof(t));
    t.B = p->B; t.GH = p->GH; t.GW = p->GW;
    t.groups = p->groups > 1 ? p->groups : 1; t.ipg = p->groups > 1 ? p->imgs_per_group : (p->B > 0 ? p->B : 1);
    t.n_taps_total = p->n_taps_total; t.noise_gstride = p->groups > 1 ? p->noise_gstride : 0;
    choose_patch(p->B, p->GH, p->GW, t.nb, t.th, t.tw, t.groups > 1 ? t.ipg : 0);
    t.tiles_x = (int)cdiv(p->GW, t.tw); t.tiles_y = (int)cdiv(p->GH, t.th);
    const int tiles_n = (int)cdiv(p->B, t.nb);
    t.Cin_blocks = p->Cin_pad / kBlockK;
    t.Cout = p->Cout; t.Cout_pad = p->Cout_pad;
    // N tile: the largest multiple of 32 (<= 256) dividing Cout_pad that still yields a grid of at least ~one wave;
    // low-resolution layers have few M tiles and are bound by the serial MMA chain of a CTA, so they take narrow tiles.
    const int64_t m_tiles = (int64_t)t.tiles_x * t.tiles_y * tiles_n;
    int n_tile = 32;
    for (int cand = 256; cand >= 32; cand -= 32) {
        if (p->Cout_pad % cand) continue;
        n_tile = cand;
        static int min_tiles1 = -1;
        if (min_tiles1 < 0) { const char* ev = getenv("IA_CONV_MIN_TILES_V1"); min_tiles1 = ev ? atoi(ev) : 128; }
        if (m_tiles * (p->Cout_pad / cand) >= min_tiles1) break;
    }
    t.n_tile = n_tile;
    t.tmem_cols = 32; while (t.tmem_cols < n_tile) t.tmem_cols <<= 1;
    const uint32_t nops = p->op_fmt == IA_OPFMT_F16X1 ? 1u : 2u;
    t.nops = (int)nops;
    t.idesc_fmt = p->op_fmt == IA_OPFMT_F16X1 ? 0u : ((1u << 7) | (1u << 10));
    const uint32_t stage_bytes = nops * (kABytes + (uint32_t)n_tile * 128u);
    const uint32_t budget = 227u * 1024u - 1024u /*alignment*/ - 256u /*barriers*/;
    int stages = (int)(budget / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    IA_CHECK(stages >= 2, "ia_conv_tc: tile does not fit shared memory");
    t.stages = stages;
    t.ntaps = p->ntaps;
    for (int i = 0; i < p->ntaps; ++i) { t.dy[i] = p->dy[i]; t.dx[i] = p->dx[i]; t.wtap[i] = p->wtap[i]; }
    t.OH = p->OH; t.OW = p->OW; t.sy = p->sy; t.sx = p->sx; t.py = p->py; t.px = p->px;
    t.mode = p->mode; t.dcoef = p->dcoef; t.noise = p->noise; t.noise_strength = p->noise_strength; t.bias = p->bias;
    t.noise_bstride = p->noise_bstride;
    t.act = p->act; t.alpha = p->alpha; t.gain = p->gain; t.clamp = p->clamp; t.slope = p->slope;
    t.emit = p->emit;

    CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
    if (int rc = make_act_map(&ma_hi, p->a_hi, p->B, p->H, p->W, p->Cin_pad, t.nb, t.th, t.tw, p->a_img_rows)) return rc;
    if (int rc = make_act_map(&ma_lo, nops == 2u ? p->a_lo : p->a_hi, p->B, p->H, p->W, p->Cin_pad, t.nb, t.th, t.tw, p->a_img_rows)) return rc;
    const int wrows = t.groups * p->n_taps_total * p->Cout_pad;
    if (int rc = make_weight_map(&mw_hi, p->w_hi, wrows, p->Cin_pad, n_tile)) return rc;
    if (int rc = make_weight_map(&mw_lo, nops == 2u ? p->w_lo : p->w_hi, wrows, p->Cin_pad, n_tile)) return rc;

    const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
    {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        IA_CHECK(e == cudaSuccess, "ia_conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    dim3 grid((unsigned)(t.tiles_x * t.tiles_y * tiles_n), (unsigned)(p->Cout_pad / n_tile));
    ia::prof_begin(ia::prof_detail_name("ia_conv_tc", p->ntaps, p->GH, p->GW, p->Cin_pad, p->Cout), as_stream(stream));
    conv_tc_kernel<<<grid, kThreads, smem, as_stream(stream)>>>(ma_hi, ma_lo, mw_hi, mw_lo, t);
    IA_LAUNCH_CHECK("ia_conv_tc");
    return 0;
}


// n launches that share operands, epilogue and emitted tensors and differ only in taps / grid / output parity (the four phases
// of a stride-2 transposed convolution, conv2d_resample.py:114-127) executed as ONE persistent launch.
extern "C" int ia_conv_tc_phases(const ia_conv_params* p, int32_t n, void* stream) {
    IA_CHECK(p && n >= 1 && n <= 4, "ia_conv_tc_phases: 1..4 sub-problems");
    bool merged = n > 1;
    {
        static int on = -1;
        if (on < 0) { const char* e = getenv("IA_CONV_MERGE_PHASES"); on = e ? atoi(e) : 1; }
        if (!on) merged = false;
        static int mode = -1;
        if (mode < 0) { const char* e = getenv("IA_CONV_TC"); mode = e ? atoi(e) : 32; }
        if (mode == 0) merged = false;
    }
    for (int i = 0; i < n && merged; ++i) {
        if (int rc = ia_conv_validate(&p[i], "ia_conv_tc_phases")) return rc;
        const ia_conv_params& a = p[0]; const ia_conv_params& b = p[i];
        // every sub-problem must be eligible for the persistent kernel and share everything but its geometry
        if (!(b.GH * b.GW >= 128 && b.GW >= 8)) merged = false;
        if (a.a_hi != b.a_hi || a.a_lo != b.a_lo || a.w_hi != b.w_hi || a.w_lo != b.w_lo || a.B != b.B || a.H != b.H || a.W != b.W ||
            a.Cin_pad != b.Cin_pad || a.Cout != b.Cout || a.Cout_pad != b.Cout_pad || a.n_taps_total != b.n_taps_total || a.OH != b.OH ||
            a.OW != b.OW || a.sy != b.sy || a.sx != b.sx || a.mode != b.mode || a.dcoef != b.dcoef || a.noise != b.noise || a.bias != b.bias ||
            a.act != b.act || a.gain != b.gain || a.clamp != b.clamp || a.emit.out32 != b.emit.out32 || a.emit.hi1 != b.emit.hi1 ||
            a.emit.hi2 != b.emit.hi2 || a.groups != b.groups || a.imgs_per_group != b.imgs_per_group || a.a_img_rows != b.a_img_rows || a.op_fmt != b.op_fmt)
            merged = false;
    }
    if (!merged) {
        for (int i = 0; i < n; ++i)
            if (int rc = ia_conv_tc(&p[i], stream)) return rc;
        return 0;
    }
    IA_CHECK((reinterpret_cast<uintptr_t>(p->a_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->a_lo) & 15) == 0 &&
             (reinterpret_cast<uintptr_t>(p->w_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->w_lo) & 15) == 0,
             "ia_conv_tc_phases: operands must be 16-byte aligned");
    const ia_conv_params* ps[4];
    for (int i = 0; i < n; ++i) ps[i] = &p[i];
    static int ph_f16 = -1;     // IA_CONV_TC_F16_PHASES: k-block of single-pass fp16 merged-phase launches
    if (ph_f16 < 0) { const char* e = getenv("IA_CONV_TC_F16_PHASES"); ph_f16 = e ? atoi(e) : 64; }
    if (p->op_fmt == IA_OPFMT_F16X1 && ph_f16 == 64) return launch_v2<64>(ps, n, stream);
    return launch_v2<32>(ps, n, stream);
}
