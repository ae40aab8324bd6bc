// losses.cu — the render-consuming losses of one render, values AND input gradients, in two launches (+ one sort).
// Replaces model/loss.py:19-97 as used by model/graph.py:220-265 — MSE_loss(rgb), mask_loss = soft-IoU (+ mask_mse * MSE),
// the trimmed ("robust") normal_loss and the eikonal MSE (mask_mse enters the gradient only: callers use the torch form when it is non-zero) — which are ~130 small torch launches per render forward+backward.
//   pass 1 (per pixel): rgb MSE partial sums + its gradient; per-image IoU sums; normal-loss sort key
//                       key = 1 - <n, n_t> where both masks > 0.5, +inf elsewhere; per-pixel value l1 * |n - n_t|_1 + key
//   (caller)          : order = stable argsort(key)            — torch.sort, the only part left to a library
//   pass 2            : n_keep = floor(n_valid * (1 - tol)) (double, as the reference); normal loss = mean of the per-pixel
//                       values of order[0 .. n_keep); IoU loss and gradients; eikonal MSE
// Gradients are written for unit upstream weight of each loss ("unit gradients"); the autograd Function scales them.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "sc_b200.h"

namespace scloss {

// workspace layout (floats): [0] sum (rgb - t)^2, [1] n_valid, [2] normal-loss sum, [3] eikonal sum, [4 + 2 b] I_b, [5 + 2 b] U_b
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void losses_pass1(const float* __restrict__ rgb, const float* __restrict__ rgb_t, const float* __restrict__ mask,
                             const float* __restrict__ mask_t, const float* __restrict__ normal, const float* __restrict__ normal_t,
                             const float* __restrict__ eik, int n_eik, int B, int R, float normal_l1,
                             float* __restrict__ ws, float* __restrict__ key, float* __restrict__ per_px,
                             float* __restrict__ rgb_unit, float* __restrict__ normal_unit, float* __restrict__ normal_t_unit,
                             float* __restrict__ eik_unit)
{
    const int b = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
    float se = 0.f, I = 0.f, U = 0.f, nv = 0.f, ee = 0.f;
    if (r < R) {
        const size_t g = (size_t)b * R + r;
        const float inv_n = 2.f / ((float)B * (float)R * 3.f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float d = rgb[g * 3 + c] - rgb_t[g * 3 + c];
            se += d * d;
            rgb_unit[g * 3 + c] = d * inv_n;
        }
        const float a = mask[g], t = mask_t[g];
        I = a * t; U = a + t - a * t + 1.e-8f;
        if (normal != nullptr) {
            const bool valid = (t > 0.5f) && (a > 0.5f);
            float dot = 0.f, l1 = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float n = normal[g * 3 + c], nt = normal_t[g * 3 + c];
                dot += n * nt; l1 += fabsf(n - nt);
                normal_unit[g * 3 + c] = 0.f; normal_t_unit[g * 3 + c] = 0.f;
            }
            const float ang = 1.f - dot;
            key[g] = valid ? ang : INFINITY;
            per_px[g] = normal_l1 * l1 + ang;
            nv = valid ? 1.f : 0.f;
        }
    }
    // eikonal MSE to 1 over n_eik values, spread over the whole grid
    if (eik != nullptr) {
        const int stride = gridDim.x * gridDim.y * blockDim.x;
        for (int i = (blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x; i < n_eik; i += stride) {
            const float d = eik[i] - 1.f;
            ee += d * d;
            eik_unit[i] = 2.f * d / (float)n_eik;
        }
    }
    __shared__ float red[8][5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    se = warp_sum(se); I = warp_sum(I); U = warp_sum(U); nv = warp_sum(nv); ee = warp_sum(ee);
    if (lane == 0) { red[warp][0] = se; red[warp][1] = I; red[warp][2] = U; red[warp][3] = nv; red[warp][4] = ee; }
    __syncthreads();
    if (threadIdx.x < 5) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
        float* dst = threadIdx.x == 0 ? ws + 0 : (threadIdx.x == 1 ? ws + 4 + 2 * b : (threadIdx.x == 2 ? ws + 5 + 2 * b :
                     (threadIdx.x == 3 ? ws + 1 : ws + 3)));
        atomicAdd(dst, s);
    }
}

__global__ void losses_pass2(const float* __restrict__ mask, const float* __restrict__ mask_t, const float* __restrict__ normal,
                             const float* __restrict__ normal_t, const int64_t* __restrict__ order, const float* __restrict__ per_px,
                             int B, int R, float normal_l1, double normal_tol, float mask_mse, float* __restrict__ ws,
                             float* __restrict__ mask_unit, float* __restrict__ normal_unit, float* __restrict__ normal_t_unit)
{
    const int b = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)B * R;
    const long i = (long)b * R + r;
    float ns = 0.f;
    if (r < R) {
        // IoU gradient: loss = mean_b (1 - I_b / U_b)
        const float I = ws[4 + 2 * b], U = ws[5 + 2 * b];
        const float a = mask[i], t = mask_t[i];
        float gm = -(t * U - I * (1.f - t)) / (U * U * (float)B);
        if (mask_mse != 0.f) gm += mask_mse * 2.f * (a - t) / (float)total;
        mask_unit[i] = gm;
        if (order != nullptr) {
            const long n_keep = (long)floor((double)ws[1] * (1.0 - normal_tol));
            if (i < n_keep) {                                     // the i-th smallest angular error among the valid pixels
                const long j = order[i];
                ns = per_px[j];
                const float inv = 1.f / (float)n_keep;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float n = normal[j * 3 + c], nt = normal_t[j * 3 + c];
                    const float sg = (n > nt) ? 1.f : ((n < nt) ? -1.f : 0.f);
                    normal_unit[j * 3 + c] = (normal_l1 * sg - nt) * inv;
                    normal_t_unit[j * 3 + c] = (-normal_l1 * sg - n) * inv;
                }
            }
        }
    }
    ns = warp_sum(ns);
    if ((threadIdx.x & 31) == 0 && ns != 0.f) atomicAdd(ws + 2, ns);
}

// losses [4] = render MSE, mask loss, normal loss, eikonal MSE
__global__ void losses_final(const float* __restrict__ ws, int B, int R, int n_eik, double normal_tol, int has_normal,
                             float* __restrict__ losses)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    losses[0] = ws[0] / ((float)B * (float)R * 3.f);
    float m = 0.f;
    for (int b = 0; b < B; ++b) m += 1.f - ws[4 + 2 * b] / ws[5 + 2 * b];
    losses[1] = m / (float)B;
    if (has_normal) {
        const long n_keep = (long)floor((double)ws[1] * (1.0 - normal_tol));
        losses[2] = ws[2] / (float)n_keep;
    } else losses[2] = 0.f;
    losses[3] = n_eik > 0 ? ws[3] / (float)n_eik : 0.f;
}

}  // namespace scloss

extern "C" size_t sc_render_losses_workspace_floats(int batch) { return 4 + 2 * (size_t)(batch > 0 ? batch : 0); }

extern "C" int sc_render_losses_pass1(const float* rgb, const float* rgb_t, const float* mask, const float* mask_t,
                                      const float* normal, const float* normal_t, const float* eik, int n_eik, int batch,
                                      int n_rays, float normal_l1, float* workspace, float* key, float* per_px,
                                      float* rgb_unit, float* normal_unit, float* normal_t_unit, float* eik_unit,
                                      cudaStream_t stream)
{
    if (batch <= 0 || n_rays <= 0) return 0;
    if (!rgb || !rgb_t || !mask || !mask_t || !workspace || !rgb_unit) return (int)cudaErrorInvalidValue;
    if (normal && (!normal_t || !key || !per_px || !normal_unit || !normal_t_unit)) return (int)cudaErrorInvalidValue;
    if (eik && !eik_unit) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(workspace, 0, sc_render_losses_workspace_floats(batch) * sizeof(float), stream);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((n_rays + 255) / 256, batch);
    scloss::losses_pass1<<<grid, 256, 0, stream>>>(rgb, rgb_t, mask, mask_t, normal, normal_t, eik, n_eik, batch, n_rays,
                                                   normal_l1, workspace, key, per_px, rgb_unit, normal_unit, normal_t_unit, eik_unit);
    return (int)cudaGetLastError();
}

extern "C" int sc_render_losses_pass2(const float* mask, const float* mask_t, const float* normal, const float* normal_t,
                                      const int64_t* order, const float* per_px, int n_eik, int batch, int n_rays, float normal_l1,
                                      double normal_tol, float mask_mse, float* workspace, float* mask_unit, float* normal_unit,
                                      float* normal_t_unit, float* losses, cudaStream_t stream)
{
    if (batch <= 0 || n_rays <= 0) return 0;
    if (!mask || !mask_t || !workspace || !mask_unit || !losses) return (int)cudaErrorInvalidValue;
    if (order && (!normal || !normal_t || !per_px || !normal_unit || !normal_t_unit)) return (int)cudaErrorInvalidValue;
    dim3 grid((n_rays + 255) / 256, batch);
    scloss::losses_pass2<<<grid, 256, 0, stream>>>(mask, mask_t, normal, normal_t, order, per_px, batch, n_rays, normal_l1,
                                                   normal_tol, mask_mse, workspace, mask_unit, normal_unit, normal_t_unit);
    scloss::losses_final<<<1, 32, 0, stream>>>(workspace, batch, n_rays, n_eik, normal_tol, order != nullptr, losses);
    return (int)cudaGetLastError();
}
