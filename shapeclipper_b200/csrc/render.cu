// render.cu — weight packing, latent biases and the FORWARD volume-render / SDF-query kernel (sm_100a).
// Reference: model/renderer.py:57-209, model/implicit.py:138-239, utils/eval_3D.py:21-38.
#include <cuda_runtime.h>
#include <stdint.h>

#include "render_tile.cuh"

namespace scr {

// ---------------------------------------------------------------------------------------------------------
// weight packing: nn.Linear tensors -> blob (render_common.cuh). One thread per blob float.
struct PackSrc { const float* w[10]; const float* b[10]; };

__device__ __forceinline__ float pack_elem(const PackSrc& s, int idx)
{
    const float r2 = 0.70710678118654752440f;
    if (idx < kSegTotal) {
        int seg = 0, off = 0;
        while (off + seg_floats(seg) <= idx) { off += seg_floats(seg); ++seg; }
        const int e = idx - off;
        // source description: tensor t with row length `in`, sub-block column offset c0, scale
        int t, in_dim, c0; float sc = 1.f; int kind;  // kind 0: T [in][out=64] ; 1: N [out][in=64] ; 2: N40 [out][40]
        switch (seg) {
            case A0T: t = 0; in_dim = 103; c0 = 0; kind = 0; break;
            case B1T: t = 1; in_dim = 167; c0 = 0; kind = 0; sc = r2; break;
            case A1T: t = 1; in_dim = 167; c0 = 64; kind = 0; sc = r2; break;
            case B2T: t = 2; in_dim = 167; c0 = 0; kind = 0; sc = r2; break;
            case A2T: t = 2; in_dim = 167; c0 = 64; kind = 0; sc = r2; break;
            case W3T: t = 3; in_dim = 64; c0 = 0; kind = 0; break;
            case W4T: t = 4; in_dim = 64; c0 = 0; kind = 0; break;
            case W5FT: t = 5; in_dim = 64; c0 = 0; kind = 3; break;            // rows 1..64 of lin5
            case V0PT: t = 6; in_dim = 167; c0 = 0; kind = 0; break;
            case V0FT: t = 6; in_dim = 167; c0 = 103; kind = 0; break;
            case V1T: t = 7; in_dim = 64; c0 = 0; kind = 0; break;
            case V2T: t = 8; in_dim = 64; c0 = 0; kind = 0; break;
            case W4N: t = 4; in_dim = 64; c0 = 0; kind = 1; break;
            case W3N: t = 3; in_dim = 64; c0 = 0; kind = 1; break;
            case B2N: t = 2; in_dim = 167; c0 = 0; kind = 1; sc = r2; break;
            case B1N: t = 1; in_dim = 167; c0 = 0; kind = 1; sc = r2; break;
            case W5FN: t = 5; in_dim = 64; c0 = 0; kind = 4; break;            // rows 1..64 of lin5, natural
            case V2N: t = 8; in_dim = 64; c0 = 0; kind = 1; break;
            case V1N: t = 7; in_dim = 64; c0 = 0; kind = 1; break;
            case V0FN: t = 6; in_dim = 167; c0 = 103; kind = 1; break;
            case A0N40: t = 0; in_dim = 103; c0 = 0; kind = 2; break;
            case A1N40: t = 1; in_dim = 167; c0 = 64; kind = 2; sc = r2; break;
            case A2N40: t = 2; in_dim = 167; c0 = 64; kind = 2; sc = r2; break;
            default /*V0PN40*/: t = 6; in_dim = 167; c0 = 0; kind = 2; break;
        }
        if (kind == 0) { const int k = e / 64, o = e % 64; return sc * s.w[t][o * in_dim + c0 + k]; }
        if (kind == 1) { const int o = e / 64, k = e % 64; return sc * s.w[t][o * in_dim + c0 + k]; }
        if (kind == 2) { const int o = e / 40, k = e % 40; return k < NPE ? sc * s.w[t][o * in_dim + c0 + k] : 0.f; }
        if (kind == 3) { const int k = e / 64, o = e % 64; return s.w[5][(o + 1) * 64 + k]; }
        { const int o = e / 64, k = e % 64; return s.w[5][(o + 1) * 64 + k]; }
    }
    idx -= kConstOffset;
    if (idx < kConstFloats) {
        if (idx < C_B3) return s.w[5][idx];                       // w5 = row 0 of lin5
        if (idx < C_B4) return s.b[3][idx - C_B3];
        if (idx < C_B5F) return s.b[4][idx - C_B4];
        if (idx < C_C1R) return s.b[5][1 + idx - C_B5F];
        if (idx < C_C2R) return s.b[7][idx - C_C1R];
        if (idx < C_V3) return s.b[8][idx - C_C2R];
        if (idx < C_C3R) return s.w[9][idx - C_V3];               // rgb lin3 [3][64]
        if (idx < C_B5) return (idx - C_C3R) < 3 ? s.b[9][idx - C_C3R] : 0.f;
        return idx == C_B5 ? s.b[5][0] : 0.f;
    }
    idx -= kConstFloats;
    if (idx < L_Z1) { const int o = idx / 64, k = idx % 64; return s.w[0][o * 103 + 39 + k]; }
    if (idx < L_Z2) { idx -= L_Z1; const int o = idx / 64, k = idx % 64; return r2 * s.w[1][o * 167 + 103 + k]; }
    if (idx < L_V0Z) { idx -= L_Z2; const int o = idx / 64, k = idx % 64; return r2 * s.w[2][o * 167 + 103 + k]; }
    if (idx < L_B0) { idx -= L_V0Z; const int o = idx / 64, k = idx % 64; return s.w[6][o * 167 + 39 + k]; }
    if (idx < L_B1) return s.b[0][idx - L_B0];
    if (idx < L_B2) return s.b[1][idx - L_B1];
    if (idx < L_C0R) return s.b[2][idx - L_B2];
    return s.b[6][idx - L_C0R];
}

__global__ void pack_kernel(PackSrc s, float* __restrict__ blob) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < kBlobFloats) blob[idx] = pack_elem(s, idx);
}

// cb[b][l][o] = sum_i Zl[o][i] z[b][i] + bias_l[o]   (l = 0,1,2: sdf latent; l = 3: rgb latent)
__global__ void latent_bias_kernel(const float* __restrict__ blob, const float* __restrict__ z_sdf,
                                   const float* __restrict__ z_rgb, int batch, float* __restrict__ cb)
{
    const int b = blockIdx.x, l = threadIdx.x >> 6, o = threadIdx.x & 63;
    const float* lat = blob + kLatentOffset;
    const float* z = (l < 3) ? z_sdf : z_rgb;
    float v = lat[L_B0 + l * 64 + o];
    if (z != nullptr) {
        const float* M = lat + l * 4096 + o * 64;
        const float* zz = z + (size_t)b * 64;
#pragma unroll 8
        for (int i = 0; i < 64; ++i) v = fmaf(M[i], zz[i], v);
    }
    cb[(size_t)b * 256 + threadIdx.x] = v;
}

// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void build_seq(int8_t* seq, int& len, int mode, bool want_grad, bool want_feat)
{
    int n = 0;
    const int8_t base[] = {A0T, B1T, A1T, B2T, A2T, W3T, W4T};
    for (int i = 0; i < 7; ++i) seq[n++] = base[i];
    if (mode == 0 || want_feat) seq[n++] = W5FT;
    if (mode == 0) { seq[n++] = V0PT; seq[n++] = V0FT; seq[n++] = V1T; seq[n++] = V2T; }
    if (mode == 0 || want_grad) {
        const int8_t g[] = {W4N, W3N, A2N40, B2N, A1N40, B1N, A0N40};
        for (int i = 0; i < 7; ++i) seq[n++] = g[i];
    }
    len = n;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) render_fwd_kernel(const ScRenderArgs a, float* stash_base, int stash_rows)
{
    extern __shared__ __align__(128) float sm[];
    __shared__ int8_t seq[64];
    __shared__ int seq_len;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_FLOATS);

    Tile T;
    T.X = sm + SM_X; T.Y = sm + SM_Y; T.Z = sm + SM_Z; T.U = sm + SM_U; T.P = sm + SM_P;
    T.cst = sm + SM_CONST; T.cb = sm + SM_CB; T.pt = sm + SM_PT; T.ray = sm + SM_RAY;
    T.tid = threadIdx.x; T.lane = threadIdx.x & 31; T.warp = threadIdx.x >> 5;
    T.stash = stash_base + (size_t)blockIdx.x * stash_rows * LD;
    T.S = (MODE == 0) ? a.n_samples : 1;
    T.rays_per_tile = (MODE == 0) ? M_TILE / a.n_samples : M_TILE;
    T.beta = (MODE == 0) ? fabsf(*a.beta_param) + a.beta_min : 1.f;

    const bool want_grad = (MODE == 0) || a.want_grad;
    const bool want_feat = (MODE == 0) || a.want_feat;
    if (T.tid == 0) {
        mbar_init(bars, 1); mbar_init(bars + 1, 1); mbar_fence_init();
        int len; build_seq(seq, len, MODE, want_grad, want_feat); seq_len = len;
    }
    for (int i = T.tid; i < kConstFloats; i += kThreads) T.cst[i] = a.blob[kConstOffset + i];
    for (int i = T.tid; i < (P_ROWS - NPE) * LD; i += kThreads) T.P[NPE * LD + i] = 0.f;
    __syncthreads();
    T.wp.blob = a.blob; T.wp.slots = sm + SM_W; T.wp.bars = bars; T.wp.seq = seq; T.wp.seq_len = seq_len;

    const int per_tile = (MODE == 0) ? T.rays_per_tile : M_TILE;
    const int tiles_per_image = (a.n_per_image + per_tile - 1) / per_tile;
    const int total = a.batch * tiles_per_image;
    if ((int)blockIdx.x >= total) return;
    T.wp.prologue();

    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        T.b = tile / tiles_per_image;
        T.first = (tile % tiles_per_image) * per_tile;
        __syncthreads();                       // previous tile's per-point readers are done
        tile_setup<MODE>(T, a);
        tile_forward<MODE, false>(T, a, want_grad, want_feat);
        __syncthreads();

        if (MODE == 1) {
            if (T.tid < M_TILE) {
                const int n = T.first + T.tid;
                if (n < a.n_per_image) {
                    const size_t g = (size_t)T.b * a.n_per_image + n;
                    a.sdf[g] = T.pv(PV_SDF)[T.tid];
                    if (a.grad != nullptr && want_grad) {
                        a.grad[g * 3 + 0] = T.pv(PV_GX0)[T.tid];
                        a.grad[g * 3 + 1] = T.pv(PV_GX1)[T.tid];
                        a.grad[g * 3 + 2] = T.pv(PV_GX2)[T.tid];
                    }
                }
            }
        } else {
            // ---- compositing (renderer.py:115-152,187-209)
            if (T.tid < 32 * 8) T.ray[32 + T.tid] = 0.f;          // per-ray accumulators [ray][8] after the warp totals
            __syncthreads();
            if (T.tid < M_TILE) {
                const int p = T.tid, S = T.S, rl = p / S;
                float delta, E, Tr, ea, w;
                tile_weights(T, delta, E, Tr, ea, w);
                const float z = T.pv(PV_Z)[p];
                const float wp = (a.normal_pow == 1.f) ? w : powf(w, a.normal_pow);
                float v[8] = {w, w * T.pv(PV_COL0)[p], w * T.pv(PV_COL1)[p], w * T.pv(PV_COL2)[p], w * z,
                              wp * T.pv(PV_NS0)[p], wp * T.pv(PV_NS1)[p], wp * T.pv(PV_NS2)[p]};
                const int seg = S < 32 ? S : 32;
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = seg_sum(v[q], seg);
                if ((T.lane & (seg - 1)) == 0) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) atomicAdd(&T.ray[32 + rl * 8 + q], v[q]);
                }
            }
            __syncthreads();
            if (T.tid < T.rays_per_tile) {
                const int r = T.first + T.tid;
                if (r < a.n_per_image) {
                    const float* acc = T.ray + 32 + T.tid * 8;
                    const size_t g = (size_t)T.b * a.n_per_image + r;
                    const float m = acc[0];
                    a.mask[g] = m;
                    a.mask_hard[g] = (m > 0.5f) ? 1.f : 0.f;
                    const float bgc = (1.f - m) * a.bg_color;
                    a.rgb[g * 3 + 0] = acc[1] + bgc; a.rgb[g * 3 + 1] = acc[2] + bgc; a.rgb[g * 3 + 2] = acc[3] + bgc;
                    a.depth[g] = acc[4] * a.depth_fac[g];
                    const float nn = sqrtf(acc[5] * acc[5] + acc[6] * acc[6] + acc[7] * acc[7]);
                    const float inv = 1.f / fmaxf(nn, 1e-12f);
                    a.normal[g * 3 + 0] = acc[5] * inv; a.normal[g * 3 + 1] = acc[6] * inv; a.normal[g * 3 + 2] = acc[7] * inv;
                }
            }
        }
    }
    T.wp.drain();
}

int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        g_num_sms = n;
    }
    return g_num_sms;
}

int check_args(const ScRenderArgs* a) {
    if (a == nullptr || a->blob == nullptr || a->cb == nullptr || a->scratch == nullptr) return (int)cudaErrorInvalidValue;
    if (a->mode == 0) {
        const int S = a->n_samples;
        if (S < 4 || S > M_TILE || (S % 4) != 0 || (M_TILE % S) != 0) return (int)cudaErrorInvalidValue;
        if (a->beta_param == nullptr) return (int)cudaErrorInvalidValue;
    } else if (a->mode != 1) return (int)cudaErrorInvalidValue;
    return 0;
}

}  // namespace scr

using namespace scr;

extern "C" size_t sc_render_blob_floats(void) { return (size_t)kBlobFloats; }
extern "C" size_t sc_render_grad_floats(void) { return (size_t)kGradFloats; }
extern "C" int sc_render_num_ctas(void) { return num_sms(); }
extern "C" size_t sc_render_scratch_bytes(int backward) {
    return (size_t)num_sms() * (size_t)(backward ? ST_BWD_ROWS : ST_FWD_ROWS) * LD * sizeof(float);
}

extern "C" int sc_render_pack_weights(const float* const* w, const float* const* b, float* blob, cudaStream_t stream)
{
    PackSrc s;
    for (int i = 0; i < 10; ++i) { s.w[i] = w[i]; s.b[i] = b[i]; }
    pack_kernel<<<(kBlobFloats + 255) / 256, 256, 0, stream>>>(s, blob);
    return (int)cudaGetLastError();
}

extern "C" int sc_render_latent_bias(const float* blob, const float* z_sdf, const float* z_rgb, int batch, float* cb,
                                     cudaStream_t stream)
{
    if (batch <= 0) return 0;
    latent_bias_kernel<<<batch, 256, 0, stream>>>(blob, z_sdf, z_rgb, batch, cb);
    return (int)cudaGetLastError();
}

extern "C" int sc_render_forward(const ScRenderArgs* a, cudaStream_t stream)
{
    int rc = check_args(a);
    if (rc) return rc;
    if (a->batch <= 0 || a->n_per_image <= 0) return 0;
    const int per_tile = a->mode == 0 ? M_TILE / a->n_samples : M_TILE;
    const long total = (long)a->batch * ((a->n_per_image + per_tile - 1) / per_tile);
    int grid = num_sms();
    if (total < grid) grid = (int)total;
    cudaError_t err;
    if (a->mode == 0) {
        err = cudaFuncSetAttribute(render_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (err != cudaSuccess) return (int)err;
        render_fwd_kernel<0><<<grid, kThreads, kSmemBytes, stream>>>(*a, (float*)a->scratch, ST_FWD_ROWS);
    } else {
        err = cudaFuncSetAttribute(render_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (err != cudaSuccess) return (int)err;
        render_fwd_kernel<1><<<grid, kThreads, kSmemBytes, stream>>>(*a, (float*)a->scratch, ST_FWD_ROWS);
    }
    return (int)cudaGetLastError();
}
