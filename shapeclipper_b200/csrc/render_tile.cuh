// render_tile.cuh — the per-tile forward program shared by the forward kernel and the backward kernel's recompute.
// See render_common.cuh for the tiling and tests/kernel_model.py for the algorithm (validated against autograd).
#pragma once
#include "render_common.cuh"
#include "sc_b200.h"

namespace scr {

struct Tile {
    // shared memory
    float *X, *Y, *Z, *U, *P, *cst, *cb, *pt, *ray;
    float* stash;           // this CTA's global scratch slot
    WeightPipe wp;
    int tid, lane, warp;
    // tile identity
    int b;                  // image
    int first;              // first ray (mode 0) / first point (mode 1) of the tile within the image
    int S;                  // samples per ray (mode 0)
    int rays_per_tile;      // M_TILE / S
    float beta;
    __device__ __forceinline__ void mark() const {}      // trace hook of the tensor-core tile type (render_ray.cuh is generic)

    __device__ __forceinline__ float* pv(int v) const { return pt + v * M_TILE; }
};

__device__ __forceinline__ float sgnf(float v) { return (float)(v > 0.f) - (float)(v < 0.f); }

// density sigma(s) and cfac = -d sigma / d s = exp(-|s|/beta) / (2 beta^2)      (model/implicit.py:70-79)
__device__ __forceinline__ void density(float s, float beta, float& sigma, float& cfac, float& e_half) {
    e_half = 0.5f * expf(-fabsf(s) / beta);
    sigma = ((s >= 0.f) ? e_half : 1.f - e_half) / beta;
    cfac = e_half / (beta * beta);
}

// d pe_k / d x~_c(k) for row k of the positional-encoding plane, from the plane itself (column p)
__device__ __forceinline__ float dpe_row(const float* __restrict__ P, int k, int p) {
    if (k < 3) return 1.f;
    const int f = (k - 3) / 6, r = (k - 3) % 6;
    const float fr = (float)(1 << f);
    return (r < 3) ? fr * P[(k + 3) * LD + p] : -fr * P[(k - 3) * LD + p];
}
__device__ __forceinline__ float d2pe_row(const float* __restrict__ P, int k, int p) {
    if (k < 3) return 0.f;
    const int f = (k - 3) / 6;
    const float fr = (float)(1 << f);
    return -fr * fr * P[k * LD + p];
}

// ---------------------------------------------------------------------------------------------------------
// Point setup: sample depths, sample points, positional encoding.      (renderer.py:13-37,84-86; implicit.py:28-38)
template <int MODE>
__device__ __forceinline__ void tile_setup(const Tile& T, const ScRenderArgs& a)
{
    for (int i = T.tid; i < 256; i += kThreads) T.cb[i] = a.cb[(size_t)T.b * 256 + i];
    if (T.tid < M_TILE) {
        const int p = T.tid;
        float x0, x1, x2, z = 0.f;
        bool valid;
        if (MODE == 0) {
            const int R = a.n_per_image, S = T.S;
            const int r = T.first + p / S, s = p % S;
            valid = r < R;
            if (valid) {
                const float c = __fmul_rn(a.cam_dist, a.scale_dist[T.b]);
                const float nr = __fsub_rn(c, a.half_range), fr = __fadd_rn(c, a.half_range);
                auto zb = [&](int i) {
                    const float t = a.t_vals[i];
                    return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.f, t)), __fmul_rn(fr, t));
                };
                z = zb(s);
                if (a.jitter != nullptr) {
                    const float up = (s < S - 1) ? __fmul_rn(0.5f, __fadd_rn(zb(s + 1), z)) : z;
                    const float lo = (s > 0) ? __fmul_rn(0.5f, __fadd_rn(z, zb(s - 1))) : z;
                    const float u = a.jitter[((size_t)T.b * R + r) * S + s];
                    z = __fadd_rn(lo, __fmul_rn(__fsub_rn(up, lo), u));
                }
                const float* d = a.ray_dirs + ((size_t)T.b * R + r) * 3;
                const float* o = a.cam_loc + (size_t)T.b * 3;
                x0 = __fadd_rn(o[0], __fmul_rn(z, d[0]));
                x1 = __fadd_rn(o[1], __fmul_rn(z, d[1]));
                x2 = __fadd_rn(o[2], __fmul_rn(z, d[2]));
            }
        } else {
            const int n = T.first + p;
            valid = n < a.n_per_image;
            if (valid) {
                const float* q = a.points + ((size_t)T.b * a.n_per_image + n) * 3;
                x0 = q[0]; x1 = q[1]; x2 = q[2];
            }
        }
        if (!valid) { x0 = 0.25f; x1 = 0.25f; x2 = 0.25f; z = 0.f; }
        T.pv(PV_Z)[p] = z;
        T.pv(PV_SGN)[p] = sgnf(x0);
        T.pv(PV_X0)[p] = x0; T.pv(PV_X1)[p] = x1; T.pv(PV_X2)[p] = x2;
        const float xt[3] = {fabsf(x0), x1, x2};
#pragma unroll
        for (int c = 0; c < 3; ++c) T.P[c * LD + p] = xt[c];
#pragma unroll
        for (int f = 0; f < 6; ++f) {
            const float fr = (float)(1 << f);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float sn, cs;
                sincosf(xt[c] * fr, &sn, &cs);
                T.P[(3 + 6 * f + c) * LD + p] = sn;
                T.P[(3 + 6 * f + 3 + c) * LD + p] = cs;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Forward program of one tile. STASH_ALL = also park what the backward sweep needs (Q, FEAT, R, GPE).
// End state (mode 0): P = posenc, U = g4, Z rows 0..39 = gpe, per-point vectors SDF, COL*, GX*, SIG, CF, UN, NS*.
template <int MODE, bool STASH_ALL>
__device__ __forceinline__ void tile_forward(Tile& T, const ScRenderArgs& a, bool want_grad, bool want_feat)
{
    const int lane = T.lane, warp = T.warp;
    float acc[4][8];
    float hv[4][8];
    const float* W;

    // ---- F.0: h0 = softplus(A0 pe + c0)
    W = T.wp.acquire(); zero(acc); gemm64(acc, T.P, NPE, W, lane, warp);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) hv[i][j] = softplus100(acc[i][j] + T.cb[0 * 64 + 8 * warp + j]);
    tile_store(T.X, hv, lane, warp); stash_store(T.stash + (ST_H + 0) * LD, hv, lane, warp);
    // ---- F.1: h1 = softplus(B1 h0 + A1 pe + c1)
    W = T.wp.acquire(); zero(acc); gemm64(acc, T.X, HID, W, lane, warp);
    W = T.wp.acquire(); gemm64(acc, T.P, NPE, W, lane, warp);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) hv[i][j] = softplus100(acc[i][j] + T.cb[1 * 64 + 8 * warp + j]);
    tile_store(T.Y, hv, lane, warp); stash_store(T.stash + (ST_H + 64) * LD, hv, lane, warp);
    // ---- F.2
    W = T.wp.acquire(); zero(acc); gemm64(acc, T.Y, HID, W, lane, warp);
    W = T.wp.acquire(); gemm64(acc, T.P, NPE, W, lane, warp);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) hv[i][j] = softplus100(acc[i][j] + T.cb[2 * 64 + 8 * warp + j]);
    tile_store(T.X, hv, lane, warp); stash_store(T.stash + (ST_H + 128) * LD, hv, lane, warp);
    // ---- F.3
    W = T.wp.acquire(); zero(acc); gemm64(acc, T.X, HID, W, lane, warp);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) hv[i][j] = softplus100(acc[i][j] + T.cst[C_B3 + 8 * warp + j]);
    tile_store(T.Y, hv, lane, warp); stash_store(T.stash + (ST_H + 192) * LD, hv, lane, warp);
    // ---- F.4 -> U (kept for the gradient pass)
    W = T.wp.acquire(); zero(acc); gemm64(acc, T.Y, HID, W, lane, warp);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) hv[i][j] = softplus100(acc[i][j] + T.cst[C_B4 + 8 * warp + j]);
    tile_store(T.U, hv, lane, warp); stash_store(T.stash + (ST_H + 256) * LD, hv, lane, warp);
    // ---- F.5: feat = W5f h4 + b5f -> Z ; sdf = w5 . h4 + b5 (per point)
    if (MODE == 0 || want_feat) {
        W = T.wp.acquire(); zero(acc); gemm64(acc, T.U, HID, W, lane, warp);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) hv[i][j] = acc[i][j] + T.cst[C_B5F + 8 * warp + j];
        if (MODE == 0) {
            tile_store(T.Z, hv, lane, warp);
            if (STASH_ALL) stash_store(T.stash + ST_FEAT * LD, hv, lane, warp);
        } else if (a.feat != nullptr) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int n = T.first + 4 * lane + i;
                if (n < a.n_per_image) {
                    float* dst = a.feat + ((size_t)T.b * a.n_per_image + n) * 64 + 8 * warp;
                    *reinterpret_cast<float4*>(dst) = make_float4(hv[i][0], hv[i][1], hv[i][2], hv[i][3]);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(hv[i][4], hv[i][5], hv[i][6], hv[i][7]);
                }
            }
        }
    }
    __syncthreads();     // U (and Z) are published for the per-point dot product below
    if (T.tid < M_TILE) {
        const int p = T.tid;
        float s = T.cst[C_B5];
#pragma unroll 16
        for (int k = 0; k < HID; ++k) s = fmaf(T.cst[C_W5 + k], T.U[k * LD + p], s);
        T.pv(PV_SDF)[p] = s;
    }

    if (MODE == 0) {
        // ---- RGB.0: r0 = relu(V0p pe + V0f feat + c_rgb) -> X
        W = T.wp.acquire(); zero(acc); gemm64(acc, T.P, NPE, W, lane, warp);
        W = T.wp.acquire(); gemm64(acc, T.Z, HID, W, lane, warp);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) hv[i][j] = fmaxf(acc[i][j] + T.cb[3 * 64 + 8 * warp + j], 0.f);
        tile_store(T.X, hv, lane, warp);
        if (STASH_ALL) stash_store(T.stash + (ST_R + 0) * LD, hv, lane, warp);
        // ---- RGB.1 -> Y
        W = T.wp.acquire(); zero(acc); gemm64(acc, T.X, HID, W, lane, warp);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) hv[i][j] = fmaxf(acc[i][j] + T.cst[C_C1R + 8 * warp + j], 0.f);
        tile_store(T.Y, hv, lane, warp);
        if (STASH_ALL) stash_store(T.stash + (ST_R + 64) * LD, hv, lane, warp);
        // ---- RGB.2 -> X
        W = T.wp.acquire(); zero(acc); gemm64(acc, T.Y, HID, W, lane, warp);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) hv[i][j] = fmaxf(acc[i][j] + T.cst[C_C2R + 8 * warp + j], 0.f);
        tile_store(T.X, hv, lane, warp);
        if (STASH_ALL) stash_store(T.stash + (ST_R + 128) * LD, hv, lane, warp);
        __syncthreads();
        // ---- RGB.3: colour = sigmoid(V3 r2 + c3) (per point)
        if (T.tid < M_TILE) {
            const int p = T.tid;
            float o[3] = {T.cst[C_C3R + 0], T.cst[C_C3R + 1], T.cst[C_C3R + 2]};
#pragma unroll 8
            for (int k = 0; k < HID; ++k) {
                const float r = T.X[k * LD + p];
                o[0] = fmaf(T.cst[C_V3 + k], r, o[0]);
                o[1] = fmaf(T.cst[C_V3 + 64 + k], r, o[1]);
                o[2] = fmaf(T.cst[C_V3 + 128 + k], r, o[2]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) T.pv(PV_COL0 + c)[p] = 1.f / (1.f + expf(-o[c]));
        }
    }

    if (MODE == 1 && !want_grad) return;

    // ---- gradient pass (reverse mode for d sdf / d x)
    // G.4: g4 = w5 * softplus'(h4), in place in U
    __syncthreads();
    tile_load(T.U, hv, lane, warp);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) hv[i][j] = T.cst[C_W5 + 8 * warp + j] * sp_slope(hv[i][j]);
    tile_store(T.U, hv, lane, warp);
    float gpe[4][5];
    zero5(gpe);
    // G.3: q3 = W4^T g4 ; g3 = q3 * s3 -> X
    W = T.wp.acquire(); zero(acc); gemm64(acc, T.U, HID, W, lane, warp);
    if (STASH_ALL) stash_store(T.stash + (ST_Q + 192) * LD, acc, lane, warp);
    stash_load(T.stash + (ST_H + 192) * LD, hv, lane, warp);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) hv[i][j] = acc[i][j] * sp_slope(hv[i][j]);
    tile_store(T.X, hv, lane, warp);
    // G.2: q2 = W3^T g3 ; g2 -> Y ; gpe += A2^T g2
    W = T.wp.acquire(); zero(acc); gemm64(acc, T.X, HID, W, lane, warp);
    if (STASH_ALL) stash_store(T.stash + (ST_Q + 128) * LD, acc, lane, warp);
    stash_load(T.stash + (ST_H + 128) * LD, hv, lane, warp);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) hv[i][j] = acc[i][j] * sp_slope(hv[i][j]);
    tile_store(T.Y, hv, lane, warp);
    W = T.wp.acquire(); gemm40(gpe, T.Y, HID, W, lane, warp);
    // G.1: q1 = B2^T g2 ; g1 -> X ; gpe += A1^T g1
    W = T.wp.acquire(); zero(acc); gemm64(acc, T.Y, HID, W, lane, warp);
    if (STASH_ALL) stash_store(T.stash + (ST_Q + 64) * LD, acc, lane, warp);
    stash_load(T.stash + (ST_H + 64) * LD, hv, lane, warp);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) hv[i][j] = acc[i][j] * sp_slope(hv[i][j]);
    tile_store(T.X, hv, lane, warp);
    W = T.wp.acquire(); gemm40(gpe, T.X, HID, W, lane, warp);
    // G.0: q0 = B1^T g1 ; g0 -> Y ; gpe += A0^T g0
    W = T.wp.acquire(); zero(acc); gemm64(acc, T.X, HID, W, lane, warp);
    if (STASH_ALL) stash_store(T.stash + (ST_Q + 0) * LD, acc, lane, warp);
    stash_load(T.stash + (ST_H + 0) * LD, hv, lane, warp);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) hv[i][j] = acc[i][j] * sp_slope(hv[i][j]);
    tile_store(T.Y, hv, lane, warp);
    W = T.wp.acquire(); gemm40(gpe, T.Y, HID, W, lane, warp);
    // gpe -> Z rows 0..39 (row 39 = 0 because the padded weight column is 0)
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const float4 v = make_float4(gpe[0][j], gpe[1][j], gpe[2][j], gpe[3][j]);
        *reinterpret_cast<float4*>(T.Z + (5 * warp + j) * LD + 4 * lane) = v;
        if (STASH_ALL) __stcg(reinterpret_cast<float4*>(T.stash + (ST_GPE + 5 * warp + j) * LD + 4 * lane), v);
    }
    __syncthreads();
    // ---- per point: gx = S J^T gpe ; density ; per-sample unit normal
    if (T.tid < M_TILE) {
        const int p = T.tid;
        float g[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < NPE; ++k) g[k % 3] = fmaf(T.Z[k * LD + p], dpe_row(T.P, k, p), g[k % 3]);
        g[0] *= T.pv(PV_SGN)[p];
        T.pv(PV_GX0)[p] = g[0]; T.pv(PV_GX1)[p] = g[1]; T.pv(PV_GX2)[p] = g[2];
        if (MODE == 0) {
            float sigma, cf, eh;
            density(T.pv(PV_SDF)[p], T.beta, sigma, cf, eh);
            const float u0 = cf * g[0], u1 = cf * g[1], u2 = cf * g[2];
            const float un = sqrtf(u0 * u0 + u1 * u1 + u2 * u2);
            const float inv = 1.f / fmaxf(un, 1e-12f);
            T.pv(PV_SIG)[p] = sigma; T.pv(PV_CF)[p] = cf; T.pv(PV_UN)[p] = un;
            T.pv(PV_NS0)[p] = u0 * inv; T.pv(PV_NS1)[p] = u1 * inv; T.pv(PV_NS2)[p] = u2 * inv;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Compositing weights of the tile's rays (renderer.py:187-209): thread p < M_TILE owns sample p.
// Writes PV_W; returns this sample's (delta, E, T, exp(-E)).
__device__ __forceinline__ void tile_weights(const Tile& T, float& delta, float& E, float& Tr, float& ea, float& w)
{
    const int p = T.tid, S = T.S;
    const int s = p % S;
    const float* z = T.pv(PV_Z);
    delta = (s < S - 1) ? z[p + 1] - z[p] : 0.f;
    E = delta * T.pv(PV_SIG)[p];
    // exclusive prefix sum of E along the ray (segments of S consecutive threads)
    const int seg = S < 32 ? S : 32;
    float incl = E;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, incl, o, seg);
        if ((T.lane & (seg - 1)) >= o) incl += t;
    }
    float before = incl - E;
    if (S > 32) {
        if (T.lane == 31) T.ray[T.warp] = incl;          // warp totals
        asm volatile("bar.sync 1, 128;");                 // the 4 point warps only
        const int w0 = (p / S) * (S / 32);
        for (int ww = w0; ww < T.warp; ++ww) before += T.ray[ww];
        asm volatile("bar.sync 1, 128;");
    }
    Tr = expf(-before);
    ea = expf(-E);
    w = (1.f - ea) * Tr;
    T.pv(PV_W)[p] = w;
}

// sum over the S samples of each ray; result valid in the thread with s == 0 ... (all lanes of a <=32 segment get
// the segment sum; for S > 32 the caller combines warps through shared atomics)
__device__ __forceinline__ float seg_sum(float v, int seg) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        if (o < seg) v += t;
    }
    return v;
}

}  // namespace scr
