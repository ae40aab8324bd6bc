// render_ray.cuh — compositing forward + adjoint of one tile's rays (renderer.py:115-152,187-209 and its autograd),
// generic over the tile type (FFMA Tile / tensor-core TileTC: both expose pv(), ray, tid, lane, warp, S, first, b, beta,
// rays_per_tile, sync() = barrier over the executing threads, scan_sync() = barrier over the threads that own a sample). Inputs: per-point vectors Z, SIG, CF, UN, NS*, COL*, GX*, SDF. Outputs: SDFB, GXB*, CB* (colour adjoint),
// ZB (partial), depth_fac_bar (global), beta adjoint (atomicAdd into `beta_acc`). Executed by all threads of the CTA; only
// the first 128 own a sample. Algorithm: tests/kernel_model.py::composite_backward.
#pragma once
#include "render_tile.cuh"

namespace scr {

constexpr int RAYX_ACC = 32;     // [ray][8] forward sums
constexpr int RAYX_BAR = 288;    // [ray][8] upstream adjoints: rgb(3) mask depth normal(3)
constexpr int RAYX_NB = 544;     // [ray][4] Nsum_bar(3)

template <class TT>
__device__ __forceinline__ float ray_scan_t(const TT& T, float v, float& total) {
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
        T.scan_sync();
        if (T.lane == 31) T.ray[T.warp] = incl;
        T.scan_sync();
        const int w0 = (p / S) * (S / 32), w1 = w0 + S / 32;
        float before = 0.f, tot = 0.f;
        for (int ww = w0; ww < w1; ++ww) { const float t = T.ray[ww]; tot += t; if (ww < T.warp) before += t; }
        incl += before; total = tot;
    }
    return incl;
}

template <class TT>
__device__ __forceinline__ void tile_weights_t(const TT& T, float& delta, float& E, float& Tr, float& ea, float& w)
{
    const int p = T.tid, S = T.S;
    const int s = p % S;
    const float* z = T.pv(PV_Z);
    delta = (s < S - 1) ? z[p + 1] - z[p] : 0.f;
    E = delta * T.pv(PV_SIG)[p];
    float tot;
    const float incl = ray_scan_t(T, E, tot);
    Tr = expf(-(incl - E));
    ea = expf(-E);
    w = (1.f - ea) * Tr;
    T.pv(PV_W)[p] = w;
}

// upstream adjoints of ray r (8 floats: rgb(3) mask depth normal(3)); NULL pointers read as zero
__device__ __forceinline__ void load_upstream(const ScRenderArgs& a, int b, int r, float (&ub)[8]) {
    const bool valid = r < a.n_per_image;
    const size_t g = (size_t)b * a.n_per_image + r;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        ub[c] = (valid && a.rgb_bar) ? a.rgb_bar[g * 3 + c] : 0.f;
        ub[5 + c] = (valid && a.normal_bar) ? a.normal_bar[g * 3 + c] : 0.f;
    }
    ub[3] = (valid && a.mask_bar) ? a.mask_bar[g] : 0.f;
    ub[4] = (valid && a.depth_bar) ? a.depth_bar[g] : 0.f;
}

// ub_pre: this thread's ray's upstream adjoints if the caller prefetched them (threads tid < rays_per_tile), else nullptr
template <class TT>
__device__ __forceinline__ void ray_phase_backward(TT& T, const ScRenderArgs& a, float* beta_acc, int nthreads,
                                                   const float* ub_pre = nullptr)
{
    const int tid = T.tid, lane = T.lane;
    float* part_beta = beta_acc;
                for (int i = tid; i < 32 * 8; i += nthreads) T.ray[RAYX_ACC + i] = 0.f;
                if (tid < T.rays_per_tile) {
                    float* ub = T.ray + RAYX_BAR + tid * 8;
                    float u8[8];
                    if (ub_pre != nullptr) {
    #pragma unroll
                        for (int c = 0; c < 8; ++c) u8[c] = ub_pre[c];
                    } else {
                        load_upstream(a, T.b, T.first + tid, u8);
                    }
    #pragma unroll
                    for (int c = 0; c < 8; ++c) ub[c] = u8[c];
                }
                T.sync();
                T.mark();
                float delta = 0.f, E = 0.f, Tr = 0.f, ea = 0.f, w = 0.f, wp = 0.f, z = 0.f, fac = 0.f;
                int rl = 0;
                if (tid < M_TILE) {
                    const int p = tid, S = T.S;
                    rl = p / S;
                    tile_weights_t(T, delta, E, Tr, ea, w);
                    z = T.pv(PV_Z)[p];
                    wp = (a.normal_pow == 1.f) ? w : powf(w, a.normal_pow);
                    float v[4] = {w * z, wp * T.pv(PV_NS0)[p], wp * T.pv(PV_NS1)[p], wp * T.pv(PV_NS2)[p]};
                    const int seg = S < 32 ? S : 32;
    #pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = seg_sum(v[q], seg);
                    if ((lane & (seg - 1)) == 0) {
    #pragma unroll
                        for (int q = 0; q < 4; ++q) atomicAdd(&T.ray[RAYX_ACC + rl * 8 + q], v[q]);
                    }
                }
                T.sync();
                T.mark();
                if (tid < T.rays_per_tile) {
                    const int r = T.first + tid;
                    const float* ac = T.ray + RAYX_ACC + tid * 8;
                    const float* ub = T.ray + RAYX_BAR + tid * 8;
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
                    T.ray[RAYX_NB + tid * 4 + 0] = nb[0]; T.ray[RAYX_NB + tid * 4 + 1] = nb[1]; T.ray[RAYX_NB + tid * 4 + 2] = nb[2];
                    if (r < a.n_per_image) a.depth_fac_bar[(size_t)T.b * a.n_per_image + r] = ub[4] * ac[0];
                }
                T.sync();
                T.mark();
                if (tid < M_TILE) {
                    const int p = tid, S = T.S, s = p % S;
                    const float* ub = T.ray + RAYX_BAR + rl * 8;
                    const float* nb = T.ray + RAYX_NB + rl * 4;
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
                    const float incl = ray_scan_t(T, C_bar, tot);
                    const float E_bar = (tot - incl) + alpha_bar * ea;
                    const float sigma = T.pv(PV_SIG)[p];
                    const float sigma_bar = E_bar * delta;
                    const float delta_bar = (s < S - 1) ? E_bar * sigma : 0.f;
                    T.pv(PV_TMP)[p] = delta_bar;
                    T.scan_sync();
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
                    if (lane == 0) atomicAdd(part_beta, bb);
                }
                T.sync();
                T.mark();
}

}  // namespace scr
