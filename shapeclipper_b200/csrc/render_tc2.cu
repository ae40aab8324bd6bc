// render_tc2.cu — two-chain tensor-core FORWARD kernel (see render_tc2.cuh). Weight blob: sc_render_tc_pack_weights.
#include <cuda_runtime.h>
#include <stdint.h>

#include "render_tc2_tile.cuh"

namespace sct2 {

__device__ __forceinline__ void build_seq_fwd(int8_t* seq, int& len, int mode, bool want_grad, bool want_feat)
{
    int n = 0;
    using namespace sct;      // SegTC ids
    const int8_t base[] = {A0N, B1N, A1N, B2N, A2N, W3N, W4N};
    for (int i = 0; i < 7; ++i) seq[n++] = base[i];
    if (mode == 0 || want_feat) seq[n++] = W5FN;
    if (mode == 0) { seq[n++] = V0PN; seq[n++] = V0FN; seq[n++] = V1N; seq[n++] = V2N; }
    if (mode == 0 || want_grad) {
        const int8_t g[] = {W4T, W3T, B2T, B1T, A2T, A1T, A0T};
        for (int i = 0; i < 7; ++i) seq[n++] = g[i];
    }
    len = n;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) render_tc2_fwd_kernel(const ScRenderArgs a, float* stash_base)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    int8_t* seq = reinterpret_cast<int8_t*>(reinterpret_cast<float*>(smem + SMB_F32) + SF_MISC);
    int& seq_len = *reinterpret_cast<int*>(reinterpret_cast<float*>(smem + SMB_F32) + SF_MISC + 16);
    const uint8_t* blob = reinterpret_cast<const uint8_t*>(a.blob);

    const bool want_grad = (MODE == 0) || a.want_grad;
    const bool want_feat = (MODE == 0) || a.want_feat;
    if (threadIdx.x == 0) { int len; build_seq_fwd(seq, len, MODE, want_grad, want_feat); seq_len = len; }
    const uint32_t tmem = tc2_prologue(smem, blob);

    TileTC2 T;
    tc2_init_tile(T, smem, blob, seq, seq_len, 4, 2);
    T.tmem = tmem;
    T.stash = stash_base + ((size_t)blockIdx.x * kGroups + T.g) * TS_PLANES_FWD * kStashPlane;
    T.S = (MODE == 0) ? a.n_samples : 1;
    T.rays_per_tile = (MODE == 0) ? MT / a.n_samples : MT;
    T.beta = (MODE == 0) ? fabsf(*a.beta_param) + a.beta_min : 1.f;

    const int per_tile = (MODE == 0) ? T.rays_per_tile : MT;
    const int tiles_per_image = (a.n_per_image + per_tile - 1) / per_tile;
    const int total = a.batch * tiles_per_image;
    const int first_tile = kGroups * blockIdx.x + T.g;
    if (first_tile < total) {
        T.wr.prologue(T.issuer);
        for (int tile = first_tile; tile < total; tile += kGroups * gridDim.x) {
            T.b = tile / tiles_per_image;
            T.first = (tile % tiles_per_image) * per_tile;
            T.sync();
            tc2_tile_setup<MODE>(T, a);
            tc2_tile_forward<MODE, false>(T, a, want_grad, want_feat);

            if (MODE == 1) {
                if (T.tg < MT) {
                    const int n = T.first + T.tg;
                    if (n < a.n_per_image) {
                        const size_t g = (size_t)T.b * a.n_per_image + n;
                        a.sdf[g] = T.pv(scr::PV_SDF)[T.tg];
                        if (a.grad != nullptr && want_grad) {
                            a.grad[g * 3 + 0] = T.pv(scr::PV_GX0)[T.tg];
                            a.grad[g * 3 + 1] = T.pv(scr::PV_GX1)[T.tg];
                            a.grad[g * 3 + 2] = T.pv(scr::PV_GX2)[T.tg];
                        }
                    }
                }
            } else {
                using namespace scr;
                if (T.tg < 16 * 8) T.ray[RAY2_ACC + T.tg] = 0.f;
                T.sync();
                if (T.tg < MT) {
                    const int p = T.tg, S = T.S, rl = p / S;
                    // compositing weights (renderer.py:187-209)
                    const int s = p % S;
                    const float* zv = T.pv(PV_Z);
                    const float delta = (s < S - 1) ? zv[p + 1] - zv[p] : 0.f;
                    const float E = delta * T.pv(PV_SIG)[p];
                    const int seg = S < 32 ? S : 32;
                    float incl = E;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const float t = __shfl_up_sync(0xffffffffu, incl, o, seg);
                        if ((T.lane & (seg - 1)) >= o) incl += t;
                    }
                    float before = incl - E;
                    if (S > 32) {                                   // S == 64: the ray spans the group's warps 0 and 1
                        if (T.lane == 31) T.ray[T.wg] = incl;
                        rays_sync(T.g);
                        if (T.wg == 1) before += T.ray[0];
                        rays_sync(T.g);
                    }
                    const float w = (1.f - expf(-E)) * expf(-before);
                    const float z = zv[p];
                    const float wp = (a.normal_pow == 1.f) ? w : powf(w, a.normal_pow);
                    float vv[8] = {w, w * T.pv(PV_COL0)[p], w * T.pv(PV_COL1)[p], w * T.pv(PV_COL2)[p], w * z,
                                   wp * T.pv(PV_NS0)[p], wp * T.pv(PV_NS1)[p], wp * T.pv(PV_NS2)[p]};
#pragma unroll
                    for (int q = 0; q < 8; ++q) vv[q] = seg_sum(vv[q], seg);
                    if ((T.lane & (seg - 1)) == 0) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) atomicAdd(&T.ray[RAY2_ACC + rl * 8 + q], vv[q]);
                    }
                }
                T.sync();
                if (T.tg < T.rays_per_tile) {
                    const int r = T.first + T.tg;
                    if (r < a.n_per_image) {
                        const float* acc = T.ray + RAY2_ACC + T.tg * 8;
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
        T.wr.drain(T.issuer);
    }
    sctc::tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) sctc::tmem_dealloc<512>(tmem);
}

}  // namespace sct2

static int tc2_grid(const ScRenderArgs* a) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_tile = a->mode == 0 ? sct2::MT / a->n_samples : sct2::MT;
    const long total = (long)a->batch * ((a->n_per_image + per_tile - 1) / per_tile);
    const long want = (total + sct2::kGroups - 1) / sct2::kGroups;
    return (int)(want < sms ? want : sms);
}

// 64-point tiles: a ray must fit one tile (4 <= S <= 64, S | 64)
extern "C" int sc_render_tc2_supported(int mode, int n_samples) {
    if (mode == 1) return 1;
    return (mode == 0 && n_samples >= 4 && n_samples <= sct2::MT && (n_samples % 4) == 0 && (sct2::MT % n_samples) == 0) ? 1 : 0;
}

extern "C" int sc_render_tc2_forward(const ScRenderArgs* a, cudaStream_t stream)
{
    if (a == nullptr || a->blob == nullptr || a->cb == nullptr || a->scratch == nullptr) return (int)cudaErrorInvalidValue;
    if (a->mode != 0 && a->mode != 1) return (int)cudaErrorInvalidValue;
    if (!sc_render_tc2_supported(a->mode, a->n_samples) || (a->mode == 0 && a->beta_param == nullptr)) return (int)cudaErrorInvalidValue;
    if (a->batch <= 0 || a->n_per_image <= 0) return 0;
    const int grid = tc2_grid(a);
    cudaError_t err;
    if (a->mode == 0) {
        err = cudaFuncSetAttribute(sct2::render_tc2_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, sct2::kSmemBytes);
        if (err != cudaSuccess) return (int)err;
        sct2::render_tc2_fwd_kernel<0><<<grid, sct2::kThreads, sct2::kSmemBytes, stream>>>(*a, (float*)a->scratch);
    } else {
        err = cudaFuncSetAttribute(sct2::render_tc2_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sct2::kSmemBytes);
        if (err != cudaSuccess) return (int)err;
        sct2::render_tc2_fwd_kernel<1><<<grid, sct2::kThreads, sct2::kSmemBytes, stream>>>(*a, (float*)a->scratch);
    }
    return (int)cudaGetLastError();
}
