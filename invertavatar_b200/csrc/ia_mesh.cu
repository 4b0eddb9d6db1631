// Mesh-condition producer: 3DMM blend shapes + eye-ball rotation + rigid / orthographic transform, and an
// orthographic z-buffer rasteriser that emits the (u, v, mask) UV-coordinate image the generator's rasterize stage consumes.
// Replaces reference data_preprocess/FaceVerse/renderer.py:45-84 (Faceverse_manager.make_driven_rendering), FaceVerseModel_v3.py
// :237-244 (get_vs), :303-325 (compute_eye_rotation_matrix) and the pytorch3d MeshRasterizer reached through
// training_avatar_texture/volumetric_rendering/ortho_renderer.py:52-100 + renderer.py:556-571 (render_after_rasterize).
// Rasterisation rule = pytorch3d's naive kernel (pixel centres at NDC 1 - (2i+1)/S, +X left / +Y up, strictly positive
// barycentrics, nearest non-negative view z, ties to the lower face index); the arithmetic is written with explicit
// round-to-nearest intrinsics (no FMA contraction) so that it reproduces the float32 oracle bit for bit.
#include "ia_common.cuh"

using namespace ia;

namespace {

// drive coefficients -> clamped / retargeted expression vector and the two eye-ball rotation matrices (renderer.py:47-57)
__global__ void mesh_coeff_kernel(const float* __restrict__ coeff, int64_t coeff_ld, int B, int exp_off, int exp_dims, int eye_off,
                                  const float* __restrict__ base_drive_exp, const float* __restrict__ base_avatar_exp,
                                  float* __restrict__ exp_out, float* __restrict__ eye_rot) {
    const int b = blockIdx.x;
    const float* c = coeff + (int64_t)b * coeff_ld;
    for (int j = threadIdx.x; j < exp_dims; j += blockDim.x) {
        float e = c[exp_off + j];
        if (j == exp_dims - 4) e = fmaxf(fminf(e, 0.6f), -0.75f);
        if (j == exp_dims - 2) e = fmaxf(fminf(e, 0.75f), -0.75f);
        if (base_drive_exp) e = (e - base_drive_exp[j]) + base_avatar_exp[j];
        exp_out[(int64_t)b * exp_dims + j] = e;
    }
    if (threadIdx.x < 2) {
        // R = Ry(eye[1]) @ Rx(eye[0]); vertices are multiplied from the left: (v - c) @ R + c   (FaceVerseModel_v3.py:303-325)
        const float ex = c[eye_off + 2 * threadIdx.x], ey = c[eye_off + 2 * threadIdx.x + 1];
        const float sx = sinf(ex), cx = cosf(ex), sy = sinf(ey), cy = cosf(ey);
        float* R = eye_rot + ((int64_t)b * 2 + threadIdx.x) * 9;
        // Ry = [[cy,0,sy],[0,1,0],[-sy,0,cy]], Rx = [[1,0,0],[0,cx,-sx],[0,sx,cx]]
        R[0] = cy;  R[1] = sy * sx;  R[2] = sy * cx;
        R[3] = 0.f; R[4] = cx;       R[5] = -sx;
        R[6] = -sy; R[7] = cy * sx;  R[8] = cy * cx;
    }
}

// centre of an eye-ball of the identity: mean over its vertex range of the neutral shape with z + 0.005 (FaceVerseModel_v3.py:252-264)
__global__ void mesh_eye_centre_kernel(const float* __restrict__ neutral, int v0, int v1, float* __restrict__ centre) {
    __shared__ float sh[3][32];
    float s[3] = {0.f, 0.f, 0.f};
    for (int v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        s[0] += neutral[v * 3]; s[1] += neutral[v * 3 + 1]; s[2] += neutral[v * 3 + 2] + 0.005f;
    }
    for (int k = 0; k < 3; ++k) {
        float t = s[k];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = t;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[threadIdx.x][w];
        centre[threadIdx.x] = t / (float)(v1 - v0);
    }
}

__global__ void __launch_bounds__(256) blendshape_kernel(ia_blendshape_params p) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (v >= p.NV) return;
    float x = p.neutral[v * 3], y = p.neutral[v * 3 + 1], z = p.neutral[v * 3 + 2];
    const float* e = p.exp + (int64_t)b * p.exp_dims;
    const float* bt = p.exp_basis_t + (int64_t)v * 3;
    const int64_t ld = (int64_t)p.NV * 3;
    for (int j = 0; j < p.exp_dims; ++j) {
        const float w = e[j];
        x = fmaf(bt[0], w, x); y = fmaf(bt[1], w, y); z = fmaf(bt[2], w, z);
        bt += ld;
    }
    int eye = -1;
    if (v >= p.eye0 && v < p.eye1) eye = 0; else if (v >= p.eye1 && v < p.eye2) eye = 1;
    if (eye >= 0) {
        const float* c = p.eye_centre + eye * 3;
        const float* R = p.eye_rot + ((int64_t)b * 2 + eye) * 9;
        const float dx = x - c[0], dy = y - c[1], dz = z - c[2];
        x = dx * R[0] + dy * R[3] + dz * R[6] + c[0];
        y = dx * R[1] + dy * R[4] + dz * R[7] + c[1];
        z = dx * R[2] + dy * R[5] + dz * R[8] + c[2];
    }
    float* o = p.verts + ((int64_t)b * p.NV + v) * 3;
    o[0] = p.M[0] * x + p.M[1] * y + p.M[2] * z + p.M[3];
    o[1] = p.M[4] * x + p.M[5] * y + p.M[6] * z + p.M[7];
    o[2] = p.M[8] * x + p.M[9] * y + p.M[10] * z + p.M[11];
}

// ---- rasteriser ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pix_centre(int i, float S) {      // 1 - (2i + 1)/S, every operation rounded on its own
    return __fsub_rn(1.0f, __fdiv_rn(__fadd_rn(__fmul_rn(2.0f, (float)i), 1.0f), S));
}
__device__ __forceinline__ float edge_fn(float ax, float ay, float bx, float by, float px, float py) {
    return __fsub_rn(__fmul_rn(__fsub_rn(px, ax), __fsub_rn(by, ay)), __fmul_rn(__fsub_rn(py, ay), __fsub_rn(bx, ax)));
}
struct Face { float x0, y0, z0, x1, y1, z1, x2, y2, z2, area; };
__device__ __forceinline__ Face load_face(const ia_ortho_raster_params& p, int b, int f, int (&vi)[3]) {
    const int* t = p.tri + (int64_t)f * 3;
    vi[0] = t[0]; vi[1] = t[1]; vi[2] = t[2];
    const float* vb = p.verts + (int64_t)b * p.NV * 3;
    Face F;
    F.x0 = -vb[vi[0] * 3]; F.y0 = -vb[vi[0] * 3 + 1]; F.z0 = __fadd_rn(vb[vi[0] * 3 + 2], p.cam_z);
    F.x1 = -vb[vi[1] * 3]; F.y1 = -vb[vi[1] * 3 + 1]; F.z1 = __fadd_rn(vb[vi[1] * 3 + 2], p.cam_z);
    F.x2 = -vb[vi[2] * 3]; F.y2 = -vb[vi[2] * 3 + 1]; F.z2 = __fadd_rn(vb[vi[2] * 3 + 2], p.cam_z);
    F.area = edge_fn(F.x0, F.y0, F.x1, F.y1, F.x2, F.y2);
    return F;
}
__device__ __forceinline__ void barycentric(const Face& F, float px, float py, float& w0, float& w1, float& w2) {
    w0 = __fdiv_rn(edge_fn(F.x1, F.y1, F.x2, F.y2, px, py), F.area);
    w1 = __fdiv_rn(edge_fn(F.x2, F.y2, F.x0, F.y0, px, py), F.area);
    w2 = __fdiv_rn(edge_fn(F.x0, F.y0, F.x1, F.y1, px, py), F.area);
}

__global__ void raster_clear_kernel(unsigned long long* z, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) z[i] = ~0ull;
}

__global__ void __launch_bounds__(128) raster_faces_kernel(ia_ortho_raster_params p, float rad) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (f >= p.F) return;
    int vi[3];
    const Face F = load_face(p, b, f, vi);
    if (fabsf(F.area) <= 1e-8f) return;
    const float S = (float)p.size;
    const float xmin = __fsub_rn(fminf(F.x0, fminf(F.x1, F.x2)), rad), xmax = __fadd_rn(fmaxf(F.x0, fmaxf(F.x1, F.x2)), rad);
    const float ymin = __fsub_rn(fminf(F.y0, fminf(F.y1, F.y2)), rad), ymax = __fadd_rn(fmaxf(F.y0, fmaxf(F.y1, F.y2)), rad);
    // pixel index of an NDC coordinate c: i = ((1 - c) S - 1) / 2; scan one pixel beyond the analytic range and apply the exact predicate
    int c_lo = (int)floorf(((1.f - xmax) * S - 1.f) * 0.5f) - 1, c_hi = (int)ceilf(((1.f - xmin) * S - 1.f) * 0.5f) + 1;
    int r_lo = (int)floorf(((1.f - ymax) * S - 1.f) * 0.5f) - 1, r_hi = (int)ceilf(((1.f - ymin) * S - 1.f) * 0.5f) + 1;
    c_lo = max(c_lo, 0); r_lo = max(r_lo, 0); c_hi = min(c_hi, p.size - 1); r_hi = min(r_hi, p.size - 1);
    unsigned long long* zb = p.zbuf + (int64_t)b * p.size * p.size;
    for (int r = r_lo; r <= r_hi; ++r) {
        const float py = pix_centre(r, S);
        if (!(py >= ymin && py <= ymax)) continue;
        for (int c = c_lo; c <= c_hi; ++c) {
            const float px = pix_centre(c, S);
            if (!(px >= xmin && px <= xmax)) continue;
            float w0, w1, w2;
            barycentric(F, px, py, w0, w1, w2);
            if (!(w0 > 0.f && w1 > 0.f && w2 > 0.f)) continue;
            const float pz = __fadd_rn(__fadd_rn(__fmul_rn(w0, F.z0), __fmul_rn(w1, F.z1)), __fmul_rn(w2, F.z2));
            if (!(pz >= 0.f)) continue;
            const unsigned long long key = ((unsigned long long)__float_as_uint(pz) << 32) | (unsigned)f;
            atomicMin(zb + (int64_t)r * p.size + c, key);
        }
    }
}

__global__ void __launch_bounds__(256) raster_resolve_kernel(ia_ortho_raster_params p) {
    const int64_t total = (int64_t)p.B * p.crop_h * p.crop_w;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = (int)(i % p.crop_w); const int64_t t = i / p.crop_w;
    const int y = (int)(t % p.crop_h); const int b = (int)(t / p.crop_h);
    const int r = y + p.crop_y, c = x + p.crop_x;
    float u = 0.f, v = 0.f, m = 0.f;
    int face = -1;
    if (r >= 0 && r < p.size && c >= 0 && c < p.size) {
        const unsigned long long key = p.zbuf[((int64_t)b * p.size + r) * p.size + c];
        if (key != ~0ull) {
            face = (int)(key & 0xffffffffull);
            int vi[3];
            const Face F = load_face(p, b, face, vi);
            const float S = (float)p.size;
            float w0, w1, w2;
            barycentric(F, pix_centre(c, S), pix_centre(r, S), w0, w1, w2);
            const float* a0 = p.attr + (int64_t)vi[0] * 3; const float* a1 = p.attr + (int64_t)vi[1] * 3; const float* a2 = p.attr + (int64_t)vi[2] * 3;
            float val[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) val[k] = __fadd_rn(__fadd_rn(__fmul_rn(w0, a0[k]), __fmul_rn(w1, a1[k])), __fmul_rn(w2, a2[k]));
            // rendering *= vismask * face_mask (renderer.py:72-74); the mask channel is then binarised at 0.5 (:83)
            const float rm = val[2];
            u = __fmul_rn(val[0], rm); v = __fmul_rn(val[1], rm);
            m = __fmul_rn(val[2], rm) >= 0.5f ? 1.f : 0.f;
        }
    }
    float* o = p.out + i * 3;
    o[0] = u; o[1] = v; o[2] = m;
    if (p.pix_to_face && x < p.crop_w) p.pix_to_face[i] = face;
}

}  // namespace

extern "C" int ia_mesh_coeffs(const float* coeff, int64_t coeff_ld, int32_t B, int32_t exp_off, int32_t exp_dims, int32_t eye_off,
                              const float* base_drive_exp, const float* base_avatar_exp, float* exp_out, float* eye_rot, void* stream) {
    IA_CHECK(coeff && exp_out && eye_rot && B > 0 && exp_dims >= 4, "ia_mesh_coeffs: bad arguments");
    IA_CHECK((base_drive_exp == nullptr) == (base_avatar_exp == nullptr), "ia_mesh_coeffs: retargeting needs both expression bases");
    ia::prof_begin("ia_mesh_coeffs", as_stream(stream));
    mesh_coeff_kernel<<<B, 128, 0, as_stream(stream)>>>(coeff, coeff_ld, B, exp_off, exp_dims, eye_off, base_drive_exp, base_avatar_exp, exp_out, eye_rot);
    IA_LAUNCH_CHECK("ia_mesh_coeffs");
    return 0;
}

extern "C" int ia_mesh_eye_centres(const float* neutral, int32_t eye0, int32_t eye1, int32_t eye2, float* centres, void* stream) {
    IA_CHECK(neutral && centres && eye0 >= 0 && eye1 > eye0 && eye2 > eye1, "ia_mesh_eye_centres: bad arguments");
    ia::prof_begin("ia_mesh_eye_centres", as_stream(stream));
    mesh_eye_centre_kernel<<<1, 256, 0, as_stream(stream)>>>(neutral, eye0, eye1, centres);
    IA_LAUNCH_CHECK("ia_mesh_eye_centres");
    ia::prof_begin("ia_mesh_eye_centres", as_stream(stream));
    mesh_eye_centre_kernel<<<1, 256, 0, as_stream(stream)>>>(neutral, eye1, eye2, centres + 3);
    IA_LAUNCH_CHECK("ia_mesh_eye_centres");
    return 0;
}

extern "C" int ia_blendshape(const ia_blendshape_params* p, void* stream) {
    IA_CHECK(p && p->neutral && p->exp_basis_t && p->exp && p->eye_rot && p->eye_centre && p->verts, "ia_blendshape: null tensor");
    IA_CHECK(p->B > 0 && p->NV > 0 && p->exp_dims > 0, "ia_blendshape: empty problem");
    IA_CHECK(p->eye0 <= p->eye1 && p->eye1 <= p->eye2 && p->eye2 <= p->NV, "ia_blendshape: eye vertex ranges out of order");
    dim3 grid((unsigned)cdiv(p->NV, 256), (unsigned)p->B);
    ia::prof_begin("ia_blendshape", as_stream(stream));
    blendshape_kernel<<<grid, 256, 0, as_stream(stream)>>>(*p);
    IA_LAUNCH_CHECK("ia_blendshape");
    return 0;
}

extern "C" int64_t ia_ortho_raster_scratch_bytes(int32_t B, int32_t size) { return (int64_t)B * size * size * 8; }

extern "C" int ia_ortho_raster(const ia_ortho_raster_params* p, void* stream) {
    IA_CHECK(p && p->verts && p->tri && p->attr && p->zbuf && p->out, "ia_ortho_raster: null tensor");
    IA_CHECK(p->B > 0 && p->NV > 0 && p->F > 0 && p->size > 0 && p->size <= 8192, "ia_ortho_raster: bad geometry");
    IA_CHECK(p->crop_w > 0 && p->crop_h > 0, "ia_ortho_raster: empty crop");
    IA_CHECK((reinterpret_cast<uintptr_t>(p->zbuf) & 7) == 0, "ia_ortho_raster: zbuf must be 8-byte aligned");
    const int64_t npix = (int64_t)p->B * p->size * p->size;
    cudaStream_t st = as_stream(stream);
    ia::prof_begin("ia_ortho_raster(clear)", st);
    raster_clear_kernel<<<(unsigned)cdiv(npix, 256), 256, 0, st>>>(p->zbuf, npix);
    IA_LAUNCH_CHECK("ia_ortho_raster(clear)");
    dim3 grid((unsigned)cdiv(p->F, 128), (unsigned)p->B);
    ia::prof_begin("ia_ortho_raster(faces)", st);
    raster_faces_kernel<<<grid, 128, 0, st>>>(*p, sqrtf(p->blur_radius > 0.f ? p->blur_radius : 0.f));
    IA_LAUNCH_CHECK("ia_ortho_raster(faces)");
    const int64_t total = (int64_t)p->B * p->crop_h * p->crop_w;
    ia::prof_begin("ia_ortho_raster(resolve)", st);
    raster_resolve_kernel<<<(unsigned)cdiv(total, 256), 256, 0, st>>>(*p);
    IA_LAUNCH_CHECK("ia_ortho_raster(resolve)");
    return 0;
}
