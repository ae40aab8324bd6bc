// render_bwd.cu — BACKWARD of the fused volume renderer / SDF query (sm_100a), and the gradient finalisation.
//
// Replaces what autograd does for model/renderer.py:57-209 + model/implicit.py:138-239 in the reference, including
// the double backward through the SDF MLP (normals and eikonal terms are themselves spatial gradients).
// Algorithm = tests/kernel_model.py (validated against autograd in fp64): per tile, recompute the forward while
// parking H/Q/FEAT/R/GPE in the CTA's L2-resident stash, run the compositing adjoint per ray, then sweep
// RGB backward -> second-order (gradient-pass) backward -> first-order backward. Every weight-gradient GEMM
// (contraction over the tile's 128 points) accumulates into a per-CTA partial that sc_render_grad_finalize reduces.
#include <cuda_runtime.h>
#include <stdint.h>

#include "render_tile.cuh"

namespace scr {

__device__ __forceinline__ void build_seq_bwd(int8_t* seq, int& len, int mode, bool second_order)
{
    int n = 0;
    const int8_t base[] = {A0T, B1T, A1T, B2T, A2T, W3T, W4T};
    for (int i = 0; i < 7; ++i) seq[n++] = base[i];
    if (mode == 0) { seq[n++] = W5FT; seq[n++] = V0PT; seq[n++] = V0FT; seq[n++] = V1T; seq[n++] = V2T; }
    if (second_order) {
        const int8_t g[] = {W4N, W3N, A2N40, B2N, A1N40, B1N, A0N40};
        for (int i = 0; i < 7; ++i) seq[n++] = g[i];
    }
    if (mode == 0) { seq[n++] = V2N; seq[n++] = V1N; seq[n++] = V0FN; seq[n++] = V0PN40; }
    if (second_order) {
        const int8_t s2[] = {A0T, A1T, B1T, A2T, B2T, W3T, W4T};
        for (int i = 0; i < 7; ++i) seq[n++] = s2[i];
    }
    if (mode == 0) seq[n++] = W5FN;
    const int8_t f1[] = {W4N, W3N, A2N40, B2N, A1N40, B1N, A0N40};
    for (int i = 0; i < 7; ++i) seq[n++] = f1[i];
    len = n;
}

// dst[8 warp + j] += sum over the tile's points of v[.][j]   (single owner: lane 0 of the warp)
__device__ __forceinline__ void rowsum_add(float* __restrict__ dst, const float (&v)[4][8], int lane, int warp) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float s = warp_sum(v[0][j] + v[1][j] + v[2][j] + v[3][j]);
        if (lane == 0) dst[8 * warp + j] += s;
    }
}
__device__ __forceinline__ void rowsum_atomic(float* __restrict__ dst, const float (&v)[4][8], int lane, int warp) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float s = warp_sum(v[0][j] + v[1][j] + v[2][j] + v[3][j]);
        if (lane == 0) atomicAdd(dst + 8 * warp + j, s);
    }
}
// pe_bar contribution (8 warps x 5 rows) folded straight into x~_bar: XTB[k % 3][p] += pe_bar_k * d pe_k / d x~
__device__ __forceinline__ void fold_pe(const Tile& T, const float (&pacc)[4][5]) {
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const int k = 5 * T.warp + j;
        if (k < NPE) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int p = 4 * T.lane + i;
                atomicAdd(T.pv(PV_XTB0 + k % 3) + p, pacc[i][j] * dpe_row(T.P, k, p));
            }
        }
    }
}
// inclusive prefix sum along the ray for the 128 point threads (segments of S consecutive threads)
__device__ __forceinline__ float ray_scan(const Tile& T, float v, float& total) {
    const int S = T.S, p = T.tid;
    const int seg = S < 32 ? S : 32;
    float incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, incl, o, seg);
        if ((T.lane & (seg - 1)) >= o) incl += t;
    }
    total = __shfl_sync(0xffffffffu, incl, seg - 1, seg);
    if (S > 32) {
        asm volatile("bar.sync 1, 128;");
        if (T.lane == 31) T.ray[T.warp] = incl;
        asm volatile("bar.sync 1, 128;");
        const int w0 = (p / S) * (S / 32), w1 = w0 + S / 32;
        float before = 0.f, tot = 0.f;
        for (int ww = w0; ww < w1; ++ww) { const float t = T.ray[ww]; tot += t; if (ww < T.warp) before += t; }
        incl += before; total = tot;
    }
    return incl;
}

constexpr int RAY_ACC = 32;     // [ray][8] forward sums
constexpr int RAY_BAR = 288;    // [ray][8] upstream adjoints: rgb(3) mask depth normal(3)
constexpr int RAY_NB = 544;     // [ray][4] Nsum_bar(3), spare

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) render_bwd_kernel(const ScRenderArgs a, float* stash_base)
{
    extern __shared__ __align__(128) float sm[];
    __shared__ int8_t seq[64];
    __shared__ int seq_len;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_FLOATS);

    Tile T;
    T.X = sm + SM_X; T.Y = sm + SM_Y; T.Z = sm + SM_Z; T.U = sm + SM_U; T.P = sm + SM_P;
    T.cst = sm + SM_CONST; T.cb = sm + SM_CB; T.pt = sm + SM_PT; T.ray = sm + SM_RAY;
    T.tid = threadIdx.x; T.lane = threadIdx.x & 31; T.warp = threadIdx.x >> 5;
    T.stash = stash_base + (size_t)blockIdx.x * ST_BWD_ROWS * LD;
    T.S = (MODE == 0) ? a.n_samples : 1;
    T.rays_per_tile = (MODE == 0) ? M_TILE / a.n_samples : M_TILE;
    T.beta = (MODE == 0) ? fabsf(*a.beta_param) + a.beta_min : 1.f;
    const int lane = T.lane, warp = T.warp, tid = T.tid;

    const bool second = (MODE == 0) || (a.want_grad && a.grad_bar != nullptr);
    float* part = a.grad_partial + (size_t)blockIdx.x * kGradFloats;
    for (int i = tid; i < kGradFloats; i += kThreads) part[i] = 0.f;
    if (tid == 0) {
        mbar_init(bars, 1); mbar_init(bars + 1, 1); mbar_fence_init();
        int len; build_seq_bwd(seq, len, MODE, second); seq_len = len;
    }
    for (int i = tid; i < kConstFloats; i += kThreads) T.cst[i] = a.blob[kConstOffset + i];
    for (int i = tid; i < (P_ROWS - NPE) * LD; i += kThreads) T.P[NPE * LD + i] = 0.f;
    __syncthreads();
    T.wp.blob = a.blob; T.wp.slots = sm + SM_W; T.wp.bars = bars; T.wp.seq = seq; T.wp.seq_len = seq_len;

    const int per_tile = (MODE == 0) ? T.rays_per_tile : M_TILE;
    const int tiles_per_image = (a.n_per_image + per_tile - 1) / per_tile;
    const int total = a.batch * tiles_per_image;
    if ((int)blockIdx.x >= total) return;
    T.wp.prologue();

    float acc[4][8], hv[4][8], t1[4][8], t2[4][8];
    float pacc[4][5];
    const float* W;
    const int cb_rows = a.detach_latent ? CB_C0D : CB_C0;

    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        T.b = tile / tiles_per_image;
        T.first = (tile % tiles_per_image) * per_tile;
        float* cbb = a.cb_bar + (size_t)T.b * kCbRows * 64;
        __syncthreads();
        tile_setup<MODE>(T, a);
        tile_forward<MODE, true>(T, a, second, MODE == 0);
        __syncthreads();

        // ======================================================================================= upstream + ray phase
        if (tid < M_TILE) { T.pv(PV_XTB0)[tid] = 0.f; T.pv(PV_XTB1)[tid] = 0.f; T.pv(PV_XTB2)[tid] = 0.f; }
        if (MODE == 1) {
            if (tid < M_TILE) {
                const int n = T.first + tid;
                const bool valid = n < a.n_per_image;
                const size_t g = (size_t)T.b * a.n_per_image + n;
                T.pv(PV_SDFB)[tid] = (valid && a.sdf_bar) ? a.sdf_bar[g] : 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) T.pv(PV_GXB0 + c)[tid] = (valid && second) ? a.grad_bar[g * 3 + c] : 0.f;
            }
        } else {
            for (int i = tid; i < 32 * 8; i += kThreads) T.ray[RAY_ACC + i] = 0.f;
            if (tid < T.rays_per_tile) {
                const int r = T.first + tid;
                const bool valid = r < a.n_per_image;
                const size_t g = (size_t)T.b * a.n_per_image + r;
                float* ub = T.ray + RAY_BAR + tid * 8;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    ub[c] = (valid && a.rgb_bar) ? a.rgb_bar[g * 3 + c] : 0.f;
                    ub[5 + c] = (valid && a.normal_bar) ? a.normal_bar[g * 3 + c] : 0.f;
                }
                ub[3] = (valid && a.mask_bar) ? a.mask_bar[g] : 0.f;
                ub[4] = (valid && a.depth_bar) ? a.depth_bar[g] : 0.f;
            }
            __syncthreads();
            float delta = 0.f, E = 0.f, Tr = 0.f, ea = 0.f, w = 0.f, wp = 0.f, z = 0.f, fac = 0.f;
            int rl = 0;
            if (tid < M_TILE) {
                const int p = tid, S = T.S;
                rl = p / S;
                tile_weights(T, delta, E, Tr, ea, w);
                z = T.pv(PV_Z)[p];
                wp = (a.normal_pow == 1.f) ? w : powf(w, a.normal_pow);
                float v[4] = {w * z, wp * T.pv(PV_NS0)[p], wp * T.pv(PV_NS1)[p], wp * T.pv(PV_NS2)[p]};
                const int seg = S < 32 ? S : 32;
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = seg_sum(v[q], seg);
                if ((lane & (seg - 1)) == 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) atomicAdd(&T.ray[RAY_ACC + rl * 8 + q], v[q]);
                }
            }
            __syncthreads();
            if (tid < T.rays_per_tile) {
                const int r = T.first + tid;
                const float* ac = T.ray + RAY_ACC + tid * 8;
                const float* ub = T.ray + RAY_BAR + tid * 8;
                const float nn = sqrtf(ac[1] * ac[1] + ac[2] * ac[2] + ac[3] * ac[3]);
                float nb[3];
                if (nn > 1e-12f) {
                    const float inv = 1.f / nn;
                    const float n0 = ac[1] * inv, n1 = ac[2] * inv, n2 = ac[3] * inv;
                    const float dt = n0 * ub[5] + n1 * ub[6] + n2 * ub[7];
                    nb[0] = (ub[5] - n0 * dt) * inv; nb[1] = (ub[6] - n1 * dt) * inv; nb[2] = (ub[7] - n2 * dt) * inv;
                } else {
                    nb[0] = ub[5] * 1e12f; nb[1] = ub[6] * 1e12f; nb[2] = ub[7] * 1e12f;
                }
                T.ray[RAY_NB + tid * 4 + 0] = nb[0]; T.ray[RAY_NB + tid * 4 + 1] = nb[1]; T.ray[RAY_NB + tid * 4 + 2] = nb[2];
                if (r < a.n_per_image) a.depth_fac_bar[(size_t)T.b * a.n_per_image + r] = ub[4] * ac[0];
            }
            __syncthreads();
            if (tid < M_TILE) {
                const int p = tid, S = T.S, s = p % S;
                const float* ub = T.ray + RAY_BAR + rl * 8;
                const float* nb = T.ray + RAY_NB + rl * 4;
                const int r = T.first + rl;
                fac = (r < a.n_per_image) ? a.depth_fac[(size_t)T.b * a.n_per_image + r] : 0.f;
                const float c0 = T.pv(PV_COL0)[p], c1 = T.pv(PV_COL1)[p], c2 = T.pv(PV_COL2)[p];
                const float ns0 = T.pv(PV_NS0)[p], ns1 = T.pv(PV_NS1)[p], ns2 = T.pv(PV_NS2)[p];
                const float ndot = nb[0] * ns0 + nb[1] * ns1 + nb[2] * ns2;
                float w_bar = ub[0] * (c0 - a.bg_color) + ub[1] * (c1 - a.bg_color) + ub[2] * (c2 - a.bg_color)
                            + ub[3] + ub[4] * z * fac;
                w_bar += (a.normal_pow == 1.f) ? ndot : a.normal_pow * powf(w, a.normal_pow - 1.f) * ndot;
                T.pv(PV_CB0)[p] = w * ub[0]; T.pv(PV_CB1)[p] = w * ub[1]; T.pv(PV_CB2)[p] = w * ub[2];
                float z_bar = ub[4] * w * fac;
                // per-sample normal: n_s = u / max(|u|, eps), u = cf * gx
                const float nsb0 = wp * nb[0], nsb1 = wp * nb[1], nsb2 = wp * nb[2];
                const float un = T.pv(PV_UN)[p], cf = T.pv(PV_CF)[p];
                float ub0, ub1, ub2;
                if (un > 1e-12f) {
                    const float inv = 1.f / un, dt = ns0 * nsb0 + ns1 * nsb1 + ns2 * nsb2;
                    ub0 = (nsb0 - ns0 * dt) * inv; ub1 = (nsb1 - ns1 * dt) * inv; ub2 = (nsb2 - ns2 * dt) * inv;
                } else { ub0 = nsb0 * 1e12f; ub1 = nsb1 * 1e12f; ub2 = nsb2 * 1e12f; }
                T.pv(PV_GXB0)[p] = cf * ub0; T.pv(PV_GXB1)[p] = cf * ub1; T.pv(PV_GXB2)[p] = cf * ub2;
                const float c_bar = ub0 * T.pv(PV_GX0)[p] + ub1 * T.pv(PV_GX1)[p] + ub2 * T.pv(PV_GX2)[p];
                // weights: w = (1 - ea) * Tr,  Tr = exp(-sum_{j<i} E_j)
                const float alpha_bar = w_bar * Tr;
                const float C_bar = -(w_bar * (1.f - ea)) * Tr;
                float tot;
                const float incl = ray_scan(T, C_bar, tot);
                const float E_bar = (tot - incl) + alpha_bar * ea;
                const float sigma = T.pv(PV_SIG)[p];
                const float sigma_bar = E_bar * delta;
                const float delta_bar = (s < S - 1) ? E_bar * sigma : 0.f;
                T.pv(PV_TMP)[p] = delta_bar;
                asm volatile("bar.sync 1, 128;");
                z_bar -= delta_bar;
                if (s > 0) z_bar += T.pv(PV_TMP)[p - 1];
                T.pv(PV_ZB)[p] = z_bar;
                // density: sigma(s, beta), cf(s, beta)
                const float sd = T.pv(PV_SDF)[p], beta = T.beta;
                const float sg = (sd >= 0.f) ? 1.f : -1.f;
                const float eh = cf * beta * beta;                         // 0.5 exp(-|s|/beta)
                const float dsig_dbeta = -sigma / beta + eh * sd / (beta * beta * beta);
                const float dc_ds = -sg / beta * cf;
                const float dc_dbeta = cf * (-2.f / beta + fabsf(sd) / (beta * beta));
                T.pv(PV_SDFB)[p] = sigma_bar * (-cf) + c_bar * dc_ds;
                const float bb = warp_sum(sigma_bar * dsig_dbeta + c_bar * dc_dbeta);
                if (lane == 0) atomicAdd(part + G_BETA, bb);
            }
            __syncthreads();

            // =================================================================================== RGB backward
            // B1: o3_bar = colour_bar * col (1 - col)
            if (tid < M_TILE) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float col = T.pv(PV_COL0 + c)[tid];
                    const float v = T.pv(PV_CB0 + c)[tid] * col * (1.f - col);
                    T.pv(PV_CB0 + c)[tid] = v;
                    const float sres = warp_sum(v);
                    if (lane == 0) atomicAdd(part + G_C3R + c, sres);
                }
            }
            plane_copy(T.X, T.stash + (ST_R + 128) * LD, 64);       // r2
            __syncthreads();
            if (tid < 192) {                                         // dV3[c][k] += sum_p o3_bar[c][p] r2[k][p]
                const int c = tid >> 6, k = tid & 63;
                const float* ob = T.pv(PV_CB0 + c);
                const float* rr = T.X + k * LD;
                float s = 0.f;
#pragma unroll 8
                for (int p = 0; p < M_TILE; ++p) s = fmaf(ob[p], rr[p], s);
                part[G_V3 + c * 64 + k] += s;
            }
            tile_load(T.X, hv, lane, warp);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int p = 4 * lane + i;
                const float o0 = T.pv(PV_CB0)[p], o1 = T.pv(PV_CB1)[p], o2 = T.pv(PV_CB2)[p];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int k = 8 * warp + j;
                    const float v = T.cst[C_V3 + k] * o0 + T.cst[C_V3 + 64 + k] * o1 + T.cst[C_V3 + 128 + k] * o2;
                    t1[i][j] = hv[i][j] > 0.f ? v : 0.f;
                }
            }
            tile_store(T.Y, t1, lane, warp);                         // o2_bar
            rowsum_add(part + G_C2R, t1, lane, warp);
            plane_copy(T.Z, T.stash + (ST_R + 64) * LD, 64);         // r1
            __syncthreads();
            wgrad<4>(T.Y, T.Z, part + G_V2, 64, 64, M_TILE);
            W = T.wp.acquire(); zero(acc); gemm64(acc, T.Y, HID, W, lane, warp);       // V2N
            tile_load(T.Z, hv, lane, warp);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) t1[i][j] = hv[i][j] > 0.f ? acc[i][j] : 0.f;
            tile_store(T.X, t1, lane, warp);                         // o1_bar
            rowsum_add(part + G_C1R, t1, lane, warp);
            __syncthreads();
            plane_copy(T.Y, T.stash + (ST_R + 0) * LD, 64);          // r0
            __syncthreads();
            wgrad<4>(T.X, T.Y, part + G_V1, 64, 64, M_TILE);
            W = T.wp.acquire(); zero(acc); gemm64(acc, T.X, HID, W, lane, warp);       // V1N
            tile_load(T.Y, hv, lane, warp);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) t1[i][j] = hv[i][j] > 0.f ? acc[i][j] : 0.f;
            tile_store(T.Z, t1, lane, warp);                         // o0_bar
            rowsum_atomic(cbb + CB_RGB * 64, t1, lane, warp);
            __syncthreads();
            plane_copy(T.X, T.stash + ST_FEAT * LD, 64);             // feat
            __syncthreads();
            wgrad<4>(T.Z, T.X, part + G_V0F, 64, 64, M_TILE);
            wgrad<3>(T.Z, T.P, part + G_V0P, NPE, NPE, M_TILE);
            W = T.wp.acquire(); zero(acc); gemm64(acc, T.Z, HID, W, lane, warp);       // V0FN -> feat_bar
            stash_store(T.stash + ST_FB * LD, acc, lane, warp);
            rowsum_add(part + G_B5F, acc, lane, warp);
            W = T.wp.acquire(); zero5(pacc); gemm40(pacc, T.Z, HID, W, lane, warp);    // V0PN40 -> pe_bar (rgb)
            fold_pe(T, pacc);
        }
        __syncthreads();

        // ======================================================================================= second-order sweep
        if (second) {
            // C1: gpe_bar = J (S gx_bar) -> X rows 0..39 ; x~_bar += S gx_bar * sum_k d2pe_k gpe_k
            if (tid < M_TILE) {
                const int p = tid;
                float gb[3] = {T.pv(PV_GXB0)[p] * T.pv(PV_SGN)[p], T.pv(PV_GXB1)[p], T.pv(PV_GXB2)[p]};
                float curv[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < NPE; ++k) {
                    T.X[k * LD + p] = gb[k % 3] * dpe_row(T.P, k, p);
                    curv[k % 3] = fmaf(d2pe_row(T.P, k, p), __ldcg(T.stash + (ST_GPE + k) * LD + p), curv[k % 3]);
                }
                T.X[NPE * LD + p] = 0.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) T.pv(PV_XTB0 + c)[p] += gb[c] * curv[c];
            }
            // C2: layer 0
            W = T.wp.acquire(); zero(acc); gemm64(acc, T.X, NPE, W, lane, warp);       // A0T
            stash_load(T.stash + (ST_H + 0) * LD, hv, lane, warp);
            stash_load(T.stash + (ST_Q + 0) * LD, t1, lane, warp);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float s = sp_slope(hv[i][j]);
                    t2[i][j] = acc[i][j] * t1[i][j];         // SB0 = g0_bar * q0
                    hv[i][j] = t1[i][j] * s;                 // g0
                    t1[i][j] = acc[i][j] * s;                // q0_bar
                }
            stash_store(T.stash + (ST_SB + 0) * LD, t2, lane, warp);
            tile_store(T.Y, t1, lane, warp); tile_store(T.Z, hv, lane, warp);
            __syncthreads();
            wgrad<3>(T.Z, T.X, part + G_A0, NPE, NPE, M_TILE);
            // C3: layer 1
            W = T.wp.acquire(); zero(acc); gemm64(acc, T.X, NPE, W, lane, warp);       // A1T
            W = T.wp.acquire(); gemm64(acc, T.Y, HID, W, lane, warp);                  // B1T
            stash_load(T.stash + (ST_H + 64) * LD, hv, lane, warp);
            stash_load(T.stash + (ST_Q + 64) * LD, t1, lane, warp);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float s = sp_slope(hv[i][j]);
                    t2[i][j] = acc[i][j] * t1[i][j];
                    hv[i][j] = t1[i][j] * s;
                    t1[i][j] = acc[i][j] * s;
                }
            stash_store(T.stash + (ST_SB + 64) * LD, t2, lane, warp);
            tile_store(T.U, t1, lane, warp); tile_store(T.Z, hv, lane, warp);          // q1_bar -> U, g1 -> Z
            __syncthreads();
            wgrad<3>(T.Z, T.X, part + G_A1, NPE, NPE, M_TILE);
            wgrad<4>(T.Z, T.Y, part + G_B1, 64, 64, M_TILE);
            // C4: layer 2
            W = T.wp.acquire(); zero(acc); gemm64(acc, T.X, NPE, W, lane, warp);       // A2T
            W = T.wp.acquire(); gemm64(acc, T.U, HID, W, lane, warp);                  // B2T
            stash_load(T.stash + (ST_H + 128) * LD, hv, lane, warp);
            stash_load(T.stash + (ST_Q + 128) * LD, t1, lane, warp);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float s = sp_slope(hv[i][j]);
                    t2[i][j] = acc[i][j] * t1[i][j];
                    hv[i][j] = t1[i][j] * s;
                    t1[i][j] = acc[i][j] * s;
                }
            stash_store(T.stash + (ST_SB + 128) * LD, t2, lane, warp);
            tile_store(T.Y, t1, lane, warp); tile_store(T.Z, hv, lane, warp);          // q2_bar -> Y, g2 -> Z
            __syncthreads();
            wgrad<3>(T.Z, T.X, part + G_A2, NPE, NPE, M_TILE);
            wgrad<4>(T.Z, T.U, part + G_B2, 64, 64, M_TILE);
            // C5: layer 3
            W = T.wp.acquire(); zero(acc); gemm64(acc, T.Y, HID, W, lane, warp);       // W3T
            stash_load(T.stash + (ST_H + 192) * LD, hv, lane, warp);
            stash_load(T.stash + (ST_Q + 192) * LD, t1, lane, warp);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float s = sp_slope(hv[i][j]);
                    t2[i][j] = acc[i][j] * t1[i][j];
                    hv[i][j] = t1[i][j] * s;
                    t1[i][j] = acc[i][j] * s;
                }
            stash_store(T.stash + (ST_SB + 192) * LD, t2, lane, warp);
            tile_store(T.U, t1, lane, warp); tile_store(T.Z, hv, lane, warp);          // q3_bar -> U, g3 -> Z
            __syncthreads();
            wgrad<4>(T.Z, T.Y, part + G_W3, 64, 64, M_TILE);
            // C6: layer 4 (q4 = w5)
            W = T.wp.acquire(); zero(acc); gemm64(acc, T.U, HID, W, lane, warp);       // W4T
            stash_load(T.stash + (ST_H + 256) * LD, hv, lane, warp);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float s = sp_slope(hv[i][j]);
                    const float w5 = T.cst[C_W5 + 8 * warp + j];
                    t2[i][j] = acc[i][j] * w5;               // SB4
                    t1[i][j] = acc[i][j] * s;                // -> dw5
                    hv[i][j] = w5 * s;                       // g4
                }
            stash_store(T.stash + (ST_SB + 256) * LD, t2, lane, warp);
            rowsum_add(part + G_W5, t1, lane, warp);
            tile_store(T.Z, hv, lane, warp);
            __syncthreads();
            wgrad<4>(T.Z, T.U, part + G_W4, 64, 64, M_TILE);
        }

        // ======================================================================================= first-order sweep
        // D1: h4_bar = w5 sdf_bar (+ W5f^T feat_bar) ; a4_bar = h4_bar s4 + SB4 t4 -> Y
        zero(acc);
        if (MODE == 0) {
            __syncthreads();
            plane_copy(T.X, T.stash + ST_FB * LD, 64);
            W = T.wp.acquire(); gemm64(acc, T.X, HID, W, lane, warp);                  // W5FN
        } else {
            __syncthreads();
        }
        stash_load(T.stash + (ST_H + 256) * LD, hv, lane, warp);
        if (second) stash_load(T.stash + (ST_SB + 256) * LD, t2, lane, warp); else zero(t2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float sb = T.pv(PV_SDFB)[4 * lane + i];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float s, t;
                sp_slope_curv(hv[i][j], s, t);
                const float hb = acc[i][j] + T.cst[C_W5 + 8 * warp + j] * sb;
                t1[i][j] = hb * s + t2[i][j] * t;            // a4_bar
                t2[i][j] = sb * hv[i][j];                    // -> dw5
            }
        }
        tile_store(T.Y, t1, lane, warp);
        rowsum_add(part + G_B4, t1, lane, warp);
        rowsum_add(part + G_W5, t2, lane, warp);
        if (tid < M_TILE) {
            const float sres = warp_sum(T.pv(PV_SDFB)[tid]);
            if (lane == 0) atomicAdd(part + G_B5, sres);
        }
        if (MODE == 0) {
            plane_copy(T.Z, T.stash + (ST_H + 256) * LD, 64);        // h4
            __syncthreads();
            wgrad<4>(T.X, T.Z, part + G_W5F, 64, 64, M_TILE);
        }
        // D2
        plane_copy(T.U, T.stash + (ST_H + 192) * LD, 64);            // h3
        __syncthreads();
        wgrad<4>(T.Y, T.U, part + G_W4, 64, 64, M_TILE);
        W = T.wp.acquire(); zero(acc); gemm64(acc, T.Y, HID, W, lane, warp);           // W4N
        tile_load(T.U, hv, lane, warp);
        if (second) stash_load(T.stash + (ST_SB + 192) * LD, t2, lane, warp); else zero(t2);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float s, t;
                sp_slope_curv(hv[i][j], s, t);
                t1[i][j] = acc[i][j] * s + t2[i][j] * t;     // a3_bar
            }
        tile_store(T.X, t1, lane, warp);
        rowsum_add(part + G_B3, t1, lane, warp);
        // D3
        plane_copy(T.Z, T.stash + (ST_H + 128) * LD, 64);            // h2
        __syncthreads();
        wgrad<4>(T.X, T.Z, part + G_W3, 64, 64, M_TILE);
        W = T.wp.acquire(); zero(acc); gemm64(acc, T.X, HID, W, lane, warp);           // W3N
        tile_load(T.Z, hv, lane, warp);
        if (second) stash_load(T.stash + (ST_SB + 128) * LD, t2, lane, warp); else zero(t2);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float s, t;
                sp_slope_curv(hv[i][j], s, t);
                t1[i][j] = acc[i][j] * s + t2[i][j] * t;     // a2_bar
            }
        tile_store(T.Y, t1, lane, warp);
        rowsum_atomic(cbb + (cb_rows + 2) * 64, t1, lane, warp);
        // D4
        plane_copy(T.U, T.stash + (ST_H + 64) * LD, 64);             // h1
        __syncthreads();
        wgrad<4>(T.Y, T.U, part + G_B2, 64, 64, M_TILE);
        wgrad<3>(T.Y, T.P, part + G_A2, NPE, NPE, M_TILE);
        W = T.wp.acquire(); zero5(pacc); gemm40(pacc, T.Y, HID, W, lane, warp);        // A2N40
        fold_pe(T, pacc);
        W = T.wp.acquire(); zero(acc); gemm64(acc, T.Y, HID, W, lane, warp);           // B2N
        tile_load(T.U, hv, lane, warp);
        if (second) stash_load(T.stash + (ST_SB + 64) * LD, t2, lane, warp); else zero(t2);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float s, t;
                sp_slope_curv(hv[i][j], s, t);
                t1[i][j] = acc[i][j] * s + t2[i][j] * t;     // a1_bar
            }
        tile_store(T.X, t1, lane, warp);
        rowsum_atomic(cbb + (cb_rows + 1) * 64, t1, lane, warp);
        // D5
        plane_copy(T.Z, T.stash + (ST_H + 0) * LD, 64);              // h0
        __syncthreads();
        wgrad<4>(T.X, T.Z, part + G_B1, 64, 64, M_TILE);
        wgrad<3>(T.X, T.P, part + G_A1, NPE, NPE, M_TILE);
        W = T.wp.acquire(); zero5(pacc); gemm40(pacc, T.X, HID, W, lane, warp);        // A1N40
        fold_pe(T, pacc);
        W = T.wp.acquire(); zero(acc); gemm64(acc, T.X, HID, W, lane, warp);           // B1N
        tile_load(T.Z, hv, lane, warp);
        if (second) stash_load(T.stash + (ST_SB + 0) * LD, t2, lane, warp); else zero(t2);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float s, t;
                sp_slope_curv(hv[i][j], s, t);
                t1[i][j] = acc[i][j] * s + t2[i][j] * t;     // a0_bar
            }
        tile_store(T.Y, t1, lane, warp);
        rowsum_atomic(cbb + (cb_rows + 0) * 64, t1, lane, warp);
        // D6
        __syncthreads();
        wgrad<3>(T.Y, T.P, part + G_A0, NPE, NPE, M_TILE);
        W = T.wp.acquire(); zero5(pacc); gemm40(pacc, T.Y, HID, W, lane, warp);        // A0N40
        fold_pe(T, pacc);
        __syncthreads();

        // D7: x_bar = S x~_bar -> points_bar (mode 1) or ray geometry (mode 0)
        if (MODE == 1) {
            if (tid < M_TILE) {
                const int n = T.first + tid;
                if (n < a.n_per_image && a.points_bar != nullptr) {
                    const size_t g = ((size_t)T.b * a.n_per_image + n) * 3;
                    a.points_bar[g + 0] = T.pv(PV_XTB0)[tid] * T.pv(PV_SGN)[tid];
                    a.points_bar[g + 1] = T.pv(PV_XTB1)[tid];
                    a.points_bar[g + 2] = T.pv(PV_XTB2)[tid];
                }
            }
        } else {
            if (tid < M_TILE) {
                const int p = tid, S = T.S, rl = p / S, r = T.first + rl;
                const bool valid = r < a.n_per_image;
                const float xb0 = T.pv(PV_XTB0)[p] * T.pv(PV_SGN)[p], xb1 = T.pv(PV_XTB1)[p], xb2 = T.pv(PV_XTB2)[p];
                const float z = T.pv(PV_Z)[p];
                float d0 = 0.f, d1 = 0.f, d2 = 0.f;
                if (valid) {
                    const float* d = a.ray_dirs + ((size_t)T.b * a.n_per_image + r) * 3;
                    d0 = d[0]; d1 = d[1]; d2 = d[2];
                }
                const float zb = T.pv(PV_ZB)[p] + d0 * xb0 + d1 * xb1 + d2 * xb2;
                float v[7] = {xb0, xb1, xb2, z * xb0, z * xb1, z * xb2, zb};
                const int seg = S < 32 ? S : 32;
#pragma unroll
                for (int q = 0; q < 7; ++q) v[q] = seg_sum(v[q], seg);
                if ((lane & (seg - 1)) == 0 && valid) {
                    float* db = a.ray_dirs_bar + ((size_t)T.b * a.n_per_image + r) * 3;
                    if (S <= 32) { db[0] = v[3]; db[1] = v[4]; db[2] = v[5]; }
                    else { atomicAdd(db + 0, v[3]); atomicAdd(db + 1, v[4]); atomicAdd(db + 2, v[5]); }
                    atomicAdd(a.cam_loc_bar + T.b * 3 + 0, v[0]);
                    atomicAdd(a.cam_loc_bar + T.b * 3 + 1, v[1]);
                    atomicAdd(a.cam_loc_bar + T.b * 3 + 2, v[2]);
                    atomicAdd(a.scale_dist_bar + T.b, a.cam_dist * v[6]);
                }
            }
        }
    }
    T.wp.drain();
}

// ---------------------------------------------------------------------------------------------------------
// Finalisation: per-CTA partials + per-image bias adjoints -> nn.Linear-layout gradients, latent and beta grads.
struct FinalizeOut { float* w[10]; float* b[10]; float* z_sdf_bar; float* z_rgb_bar; float* beta_bar;     int accumulate;
};

__device__ __forceinline__ float psum(const float* __restrict__ partial, int n, int idx) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += partial[(size_t)i * kGradFloats + idx];
    return s;
}
__device__ __forceinline__ float cb_outer(const float* __restrict__ cbb, const float* __restrict__ z, int B, int row,
                                          int row_d, int o, int k) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
        float c = cbb[((size_t)b * kCbRows + row) * 64 + o];
        if (row_d >= 0) c += cbb[((size_t)b * kCbRows + row_d) * 64 + o];
        s = fmaf(c, z[(size_t)b * 64 + k], s);
    }
    return s;
}
__device__ __forceinline__ float cb_sum(const float* __restrict__ cbb, int B, int row, int row_d, int o) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) {
        s += cbb[((size_t)b * kCbRows + row) * 64 + o];
        if (row_d >= 0) s += cbb[((size_t)b * kCbRows + row_d) * 64 + o];
    }
    return s;
}

constexpr int F_W0 = 0, F_W1 = F_W0 + 64 * 103, F_W2 = F_W1 + 64 * 167, F_W3 = F_W2 + 64 * 167, F_W4 = F_W3 + 4096,
              F_W5 = F_W4 + 4096, F_V0 = F_W5 + 65 * 64, F_V1 = F_V0 + 64 * 167, F_V2 = F_V1 + 4096, F_V3 = F_V2 + 4096,
              F_BIAS = F_V3 + 192, F_BIAS_END = F_BIAS + 64 * 5 + 65 + 64 * 3 + 3, F_END = F_BIAS_END;

__global__ void finalize_kernel(const float* __restrict__ partial, int n, const float* __restrict__ cbb,
                                const float* __restrict__ z_sdf, const float* __restrict__ z_rgb,
                                const float* __restrict__ blob, int B, FinalizeOut out)
{
    const float r2 = 0.70710678118654752440f;
    // parameter gradients: overwrite, or add to what is there (fused gradient accumulation into p.grad)
    auto put = [&](float* p, float v) { *p = out.accumulate ? *p + v : v; };
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < F_W1) {                                               // sdf lin0.weight [64][103]
        if (!out.w[0]) return;
        const int o = idx / 103, c = idx % 103;
        put(&out.w[0][idx], c < 39 ? psum(partial, n, G_A0 + o * 39 + c) : cb_outer(cbb, z_sdf, B, CB_C0, CB_C0D, o, c - 39));
    } else if (idx < F_W3) {                                        // sdf lin1 / lin2 .weight [64][167]
        const int l = idx < F_W2 ? 1 : 2;
        if (!out.w[l]) return;
        const int e = idx - (l == 1 ? F_W1 : F_W2), o = e / 167, c = e % 167;
        const int gB = l == 1 ? G_B1 : G_B2, gA = l == 1 ? G_A1 : G_A2;
        float v;
        if (c < 64) v = psum(partial, n, gB + o * 64 + c);
        else if (c < 103) v = psum(partial, n, gA + o * 39 + (c - 64));
        else v = cb_outer(cbb, z_sdf, B, CB_C0 + l, CB_C0D + l, o, c - 103);
        put(&out.w[l][e], r2 * v);
    } else if (idx < F_W5) {                                        // lin3 / lin4
        const int l = idx < F_W4 ? 3 : 4;
        if (!out.w[l]) return;
        const int e = idx - (l == 3 ? F_W3 : F_W4);
        put(&out.w[l][e], psum(partial, n, (l == 3 ? G_W3 : G_W4) + e));
    } else if (idx < F_V0) {                                        // lin5.weight [65][64]
        if (!out.w[5]) return;
        const int e = idx - F_W5, o = e / 64, k = e % 64;
        put(&out.w[5][e], o == 0 ? psum(partial, n, G_W5 + k) : psum(partial, n, G_W5F + (o - 1) * 64 + k));
    } else if (idx < F_V1) {                                        // rgb lin0.weight [64][167]
        if (!out.w[6]) return;
        const int e = idx - F_V0, o = e / 167, c = e % 167;
        float v;
        if (c < 39) v = psum(partial, n, G_V0P + o * 39 + c);
        else if (c < 103) v = cb_outer(cbb, z_rgb, B, CB_RGB, -1, o, c - 39);
        else v = psum(partial, n, G_V0F + o * 64 + (c - 103));
        put(&out.w[6][e], v);
    } else if (idx < F_V3) {
        const int l = idx < F_V2 ? 7 : 8;
        if (!out.w[l]) return;
        const int e = idx - (l == 7 ? F_V1 : F_V2);
        put(&out.w[l][e], psum(partial, n, (l == 7 ? G_V1 : G_V2) + e));
    } else if (idx < F_BIAS) {
        if (!out.w[9]) return;
        const int e = idx - F_V3;
        put(&out.w[9][e], psum(partial, n, G_V3 + e));
    } else if (idx < F_BIAS_END) {
        int e = idx - F_BIAS;
        if (e < 192) {                                              // sdf lin0..2 bias
            const int l = e / 64, o = e % 64;
            if (out.b[l]) put(&out.b[l][o], cb_sum(cbb, B, CB_C0 + l, CB_C0D + l, o));
            return;
        }
        e -= 192;
        if (e < 128) { const int l = 3 + e / 64, o = e % 64; if (out.b[l]) put(&out.b[l][o], psum(partial, n, (l == 3 ? G_B3 : G_B4) + o)); return; }
        e -= 128;
        if (e < 65) { if (out.b[5]) put(&out.b[5][e], e == 0 ? psum(partial, n, G_B5) : psum(partial, n, G_B5F + e - 1)); return; }
        e -= 65;
        if (e < 64) { if (out.b[6]) put(&out.b[6][e], cb_sum(cbb, B, CB_RGB, -1, e)); return; }
        e -= 64;
        if (e < 128) { const int l = 7 + e / 64, o = e % 64; if (out.b[l]) put(&out.b[l][o], psum(partial, n, (l == 7 ? G_C1R : G_C2R) + o)); return; }
        e -= 128;
        if (out.b[9]) put(&out.b[9][e], psum(partial, n, G_C3R + e));
    } else {
        idx -= F_END;
        const float* lat = blob + kLatentOffset;
        if (idx < B * 64) {                                         // z_sdf_bar[b][i] = sum_l Z_l^T c_l_bar[b]
            if (!out.z_sdf_bar) return;
            const int b = idx / 64, i = idx % 64;
            float s = 0.f;
            for (int l = 0; l < 3; ++l)
                for (int o = 0; o < 64; ++o)
                    s = fmaf(lat[l * 4096 + o * 64 + i], cbb[((size_t)b * kCbRows + CB_C0 + l) * 64 + o], s);
            out.z_sdf_bar[idx] = s;
        } else if (idx < 2 * B * 64) {
            if (!out.z_rgb_bar) return;
            idx -= B * 64;
            const int b = idx / 64, i = idx % 64;
            float s = 0.f;
            for (int o = 0; o < 64; ++o) s = fmaf(lat[L_V0Z + o * 64 + i], cbb[((size_t)b * kCbRows + CB_RGB) * 64 + o], s);
            out.z_rgb_bar[idx] = s;
        } else if (idx == 2 * B * 64) {
            if (out.beta_bar) out.beta_bar[0] = psum(partial, n, G_BETA);
        }
    }
}

int num_sms_bwd() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
}

}  // namespace scr

using namespace scr;

extern "C" int sc_render_backward(const ScRenderArgs* a, cudaStream_t stream)
{
    if (a == nullptr || a->blob == nullptr || a->cb == nullptr || a->scratch == nullptr || a->grad_partial == nullptr ||
        a->cb_bar == nullptr)
        return (int)cudaErrorInvalidValue;
    if (a->mode == 0) {
        const int S = a->n_samples;
        if (S < 4 || S > M_TILE || (S % 4) != 0 || (M_TILE % S) != 0 || a->beta_param == nullptr) return (int)cudaErrorInvalidValue;
        if (!a->ray_dirs_bar || !a->depth_fac_bar || !a->cam_loc_bar || !a->scale_dist_bar) return (int)cudaErrorInvalidValue;
    } else if (a->mode != 1) return (int)cudaErrorInvalidValue;
    const int grid = num_sms_bwd();     // every CTA zeroes its gradient partial, tiles or not
    cudaError_t err;
    if (a->mode == 0) {
        err = cudaFuncSetAttribute(render_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (err != cudaSuccess) return (int)err;
        render_bwd_kernel<0><<<grid, kThreads, kSmemBytes, stream>>>(*a, (float*)a->scratch);
    } else {
        err = cudaFuncSetAttribute(render_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (err != cudaSuccess) return (int)err;
        render_bwd_kernel<1><<<grid, kThreads, kSmemBytes, stream>>>(*a, (float*)a->scratch);
    }
    return (int)cudaGetLastError();
}

static int finalize_impl(const float* grad_partial, int n_ctas, const float* cb_bar, const float* z_sdf,
                         const float* z_rgb, const float* blob, int batch, float* const* out_w,
                         float* const* out_b, float* z_sdf_bar, float* z_rgb_bar, float* beta_bar, int accumulate,
                         cudaStream_t stream)
{
    FinalizeOut o;
    o.accumulate = accumulate;
    for (int i = 0; i < 10; ++i) { o.w[i] = out_w[i]; o.b[i] = out_b[i]; }
    o.z_sdf_bar = z_sdf_bar; o.z_rgb_bar = (z_rgb != nullptr) ? z_rgb_bar : nullptr; o.beta_bar = beta_bar;
    const int total = F_END + 2 * batch * 64 + 1;
    finalize_kernel<<<(total + 127) / 128, 128, 0, stream>>>(grad_partial, n_ctas, cb_bar, z_sdf, z_rgb, blob, batch, o);
    return (int)cudaGetLastError();
}

extern "C" int sc_render_grad_finalize(const float* grad_partial, int n_ctas, const float* cb_bar, const float* z_sdf,
                                       const float* z_rgb, const float* blob, int batch, float* const* out_w,
                                       float* const* out_b, float* z_sdf_bar, float* z_rgb_bar, float* beta_bar,
                                       cudaStream_t stream)
{
    return finalize_impl(grad_partial, n_ctas, cb_bar, z_sdf, z_rgb, blob, batch, out_w, out_b, z_sdf_bar, z_rgb_bar, beta_bar, 0, stream);
}

extern "C" int sc_render_grad_finalize_accumulate(const float* grad_partial, int n_ctas, const float* cb_bar, const float* z_sdf,
                                                  const float* z_rgb, const float* blob, int batch, float* const* out_w,
                                                  float* const* out_b, float* z_sdf_bar, float* z_rgb_bar, float* beta_bar,
                                                  cudaStream_t stream)
{
    return finalize_impl(grad_partial, n_ctas, cb_bar, z_sdf, z_rgb, blob, batch, out_w, out_b, z_sdf_bar, z_rgb_bar, beta_bar, 1, stream);
}
