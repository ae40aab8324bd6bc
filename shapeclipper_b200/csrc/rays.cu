// rays.cu — ray geometry of the selected pixels, forward and adjoint.
// Replaces utils/camera.py:157-196 (get_camera_grid + get_center_and_ray, perspective model) and model/renderer.py:59-68:
//   pix = (x + .5, y + .5, 1);  cam = K^-1 pix;  raw = R^T cam (= world point - camera centre);  centre = -R^T t
//   ray_dirs = raw / max(|raw|, 1e-12);  depth_fac = |ray_dirs| / |raw|
// for pose = [R | t] (world -> camera), only for the R rays listed in ray_idx (the reference builds all H*W rays and
// gathers). In torch this is ~35 launches forward and ~70 backward per render; here one launch each way (+ a 1-block finish).
#include <cuda_runtime.h>
#include <stdint.h>

#include "sc_b200.h"

namespace scrays {

struct Cam { float kinv[9]; float r[9]; };

// closed-form inverse (adjugate / determinant), same arithmetic as camera.inv3x3
__device__ __forceinline__ void inv3x3(const float* m, float* o) {
    const float c0x = m[4] * m[8] - m[5] * m[7], c0y = m[5] * m[6] - m[3] * m[8], c0z = m[3] * m[7] - m[4] * m[6];   // r1 x r2
    const float c1x = m[7] * m[2] - m[8] * m[1], c1y = m[8] * m[0] - m[6] * m[2], c1z = m[6] * m[1] - m[7] * m[0];   // r2 x r0
    const float c2x = m[1] * m[5] - m[2] * m[4], c2y = m[2] * m[3] - m[0] * m[5], c2z = m[0] * m[4] - m[1] * m[3];   // r0 x r1
    const float det = m[0] * c0x + m[1] * c0y + m[2] * c0z;
    o[0] = c0x / det; o[1] = c1x / det; o[2] = c2x / det;
    o[3] = c0y / det; o[4] = c1y / det; o[5] = c2y / det;
    o[6] = c0z / det; o[7] = c1z / det; o[8] = c2z / det;
}
__device__ __forceinline__ void load_cam(const float* pose, const float* intr, int b, Cam& c) {
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) k[i] = intr[b * 9 + i];
    inv3x3(k, c.kinv);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c.r[i * 3 + j] = pose[b * 12 + i * 4 + j];
}
__device__ __forceinline__ void pixel_of(const int64_t* ray_idx, int b, int r, int R, int W, float* pix) {
    const int64_t id = ray_idx ? ray_idx[(size_t)b * R + r] : (int64_t)r;
    pix[0] = (float)(id % W) + 0.5f; pix[1] = (float)(id / W) + 0.5f; pix[2] = 1.f;
}

__global__ void rays_fwd_kernel(const float* __restrict__ pose, const float* __restrict__ intr, const int64_t* __restrict__ ray_idx,
                                int R, int W, float* __restrict__ cam_loc, float* __restrict__ dirs, float* __restrict__ fac)
{
    const int b = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
    Cam c;
    load_cam(pose, intr, b, c);
    if (blockIdx.x == 0 && threadIdx.x < 3) {                     // centre_j = -sum_i t_i R_ij
        const int j = threadIdx.x;
        cam_loc[b * 3 + j] = -(pose[b * 12 + 3] * c.r[j] + pose[b * 12 + 7] * c.r[3 + j] + pose[b * 12 + 11] * c.r[6 + j]);
    }
    if (r >= R) return;
    float pix[3], cam[3], raw[3];
    pixel_of(ray_idx, b, r, R, W, pix);
#pragma unroll
    for (int i = 0; i < 3; ++i) cam[i] = c.kinv[i * 3] * pix[0] + c.kinv[i * 3 + 1] * pix[1] + c.kinv[i * 3 + 2] * pix[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) raw[j] = cam[0] * c.r[j] + cam[1] * c.r[3 + j] + cam[2] * c.r[6 + j];
    const float n = sqrtf(raw[0] * raw[0] + raw[1] * raw[1] + raw[2] * raw[2]);
    const float inv = 1.f / fmaxf(n, 1e-12f);
    const float d0 = raw[0] * inv, d1 = raw[1] * inv, d2 = raw[2] * inv;
    const size_t g = (size_t)b * R + r;
    dirs[g * 3] = d0; dirs[g * 3 + 1] = d1; dirs[g * 3 + 2] = d2;
    fac[g] = sqrtf(d0 * d0 + d1 * d1 + d2 * d2) / n;
}

// acc [B][18]: R_bar (9) then Kinv_bar (9), zeroed by the launcher
__global__ void rays_bwd_kernel(const float* __restrict__ pose, const float* __restrict__ intr, const int64_t* __restrict__ ray_idx,
                                int R, int W, const float* __restrict__ g_dirs, const float* __restrict__ g_fac,
                                float* __restrict__ acc)
{
    const int b = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
    Cam c;
    load_cam(pose, intr, b, c);
    float v[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) v[i] = 0.f;
    if (r < R) {
        float pix[3], cam[3], raw[3];
        pixel_of(ray_idx, b, r, R, W, pix);
#pragma unroll
        for (int i = 0; i < 3; ++i) cam[i] = c.kinv[i * 3] * pix[0] + c.kinv[i * 3 + 1] * pix[1] + c.kinv[i * 3 + 2] * pix[2];
#pragma unroll
        for (int j = 0; j < 3; ++j) raw[j] = cam[0] * c.r[j] + cam[1] * c.r[3 + j] + cam[2] * c.r[6 + j];
        const float n = fmaxf(sqrtf(raw[0] * raw[0] + raw[1] * raw[1] + raw[2] * raw[2]), 1e-12f), inv = 1.f / n;
        const float d[3] = {raw[0] * inv, raw[1] * inv, raw[2] * inv};
        const size_t g = (size_t)b * R + r;
        const float gd[3] = {g_dirs ? g_dirs[g * 3] : 0.f, g_dirs ? g_dirs[g * 3 + 1] : 0.f, g_dirs ? g_dirs[g * 3 + 2] : 0.f};
        const float gf = g_fac ? g_fac[g] : 0.f;
        const float dot = d[0] * gd[0] + d[1] * gd[1] + d[2] * gd[2];
        float rb[3], cb[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) rb[j] = (gd[j] - d[j] * dot) * inv - gf * d[j] * inv * inv;    // d dirs + d(1/|raw|)
#pragma unroll
        for (int i = 0; i < 3; ++i) cb[i] = c.r[i * 3] * rb[0] + c.r[i * 3 + 1] * rb[1] + c.r[i * 3 + 2] * rb[2];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) { v[i * 3 + j] = cam[i] * rb[j]; v[9 + i * 3 + j] = cb[i] * pix[j]; }
    }
    __shared__ float red[8][18];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 18; ++i) {
        float s = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 18) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
        atomicAdd(acc + b * 18 + threadIdx.x, s);
    }
}

// pose_bar [B,3,4], intr_bar [B,3,3] from the accumulators and the centre adjoint
__global__ void rays_bwd_finish_kernel(const float* __restrict__ pose, const float* __restrict__ intr, const float* __restrict__ acc,
                                       const float* __restrict__ g_center, int B, float* __restrict__ pose_bar,
                                       float* __restrict__ intr_bar)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    Cam c;
    load_cam(pose, intr, b, c);
    const float* a = acc + b * 18;
    const float gc[3] = {g_center ? g_center[b * 3] : 0.f, g_center ? g_center[b * 3 + 1] : 0.f, g_center ? g_center[b * 3 + 2] : 0.f};
    const float t[3] = {pose[b * 12 + 3], pose[b * 12 + 7], pose[b * 12 + 11]};
    if (pose_bar) {
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) pose_bar[b * 12 + i * 4 + j] = a[i * 3 + j] - t[i] * gc[j];       // centre_j = -sum_i t_i R_ij
            pose_bar[b * 12 + i * 4 + 3] = -(c.r[i * 3] * gc[0] + c.r[i * 3 + 1] * gc[1] + c.r[i * 3 + 2] * gc[2]);
        }
    }
    if (intr_bar) {                                               // K_bar = -Kinv^T Kinv_bar Kinv^T
        float m[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                m[i * 3 + j] = c.kinv[0 * 3 + i] * a[9 + 0 * 3 + j] + c.kinv[1 * 3 + i] * a[9 + 1 * 3 + j] + c.kinv[2 * 3 + i] * a[9 + 2 * 3 + j];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                intr_bar[b * 9 + i * 3 + j] = -(m[i * 3] * c.kinv[j * 3] + m[i * 3 + 1] * c.kinv[j * 3 + 1] + m[i * 3 + 2] * c.kinv[j * 3 + 2]);
    }
}

}  // namespace scrays

extern "C" int sc_pixel_rays_forward(const float* pose, const float* intr, const int64_t* ray_idx, int batch, int n_rays,
                                     int width, float* cam_loc, float* ray_dirs, float* depth_fac, cudaStream_t stream)
{
    if (batch <= 0 || n_rays <= 0) return 0;
    if (!pose || !intr || !cam_loc || !ray_dirs || !depth_fac || width <= 0) return (int)cudaErrorInvalidValue;
    dim3 grid((n_rays + 255) / 256, batch);
    scrays::rays_fwd_kernel<<<grid, 256, 0, stream>>>(pose, intr, ray_idx, n_rays, width, cam_loc, ray_dirs, depth_fac);
    return (int)cudaGetLastError();
}

extern "C" int sc_pixel_rays_backward(const float* pose, const float* intr, const int64_t* ray_idx, int batch, int n_rays,
                                      int width, const float* cam_loc_bar, const float* ray_dirs_bar, const float* depth_fac_bar,
                                      float* workspace, float* pose_bar, float* intr_bar, cudaStream_t stream)
{
    if (batch <= 0 || n_rays <= 0) return 0;
    if (!pose || !intr || !workspace || width <= 0) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)batch * 18 * sizeof(float), stream);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((n_rays + 255) / 256, batch);
    scrays::rays_bwd_kernel<<<grid, 256, 0, stream>>>(pose, intr, ray_idx, n_rays, width, ray_dirs_bar, depth_fac_bar, workspace);
    scrays::rays_bwd_finish_kernel<<<(batch + 63) / 64, 64, 0, stream>>>(pose, intr, workspace, cam_loc_bar, batch, pose_bar, intr_bar);
    return (int)cudaGetLastError();
}

// ---- eikonal sample points (model/renderer.py:154-170 + UniformSampler.get_z_vals's z_eik, model/renderer.py:13-37) --------
// pts [B, 2R, 3] = cat( uniform points uni [B,R,3],  cam_loc + z_eik * ray_dirs )  with z_eik = the depth of the randomly picked
// sample eik_idx of each ray (stratified with jitter u when u != NULL). In torch this is ~45 launches forward and ~40 backward.
namespace scrays {

__device__ __forceinline__ float eik_depth(float sd, float cam_dist, float half_range, const float* __restrict__ t_vals, int S,
                                           int idx, const float* __restrict__ u_row)
{
    // the same un-contracted fp32 operations as the render kernel's depth set-up (and torch's elementwise ops)
    const float c = __fmul_rn(cam_dist, sd), nr = __fsub_rn(c, half_range), fr = __fadd_rn(c, half_range);
    auto zb = [&](int i) {
        const float t = t_vals[min(max(i, 0), S - 1)];
        return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.f, t)), __fmul_rn(fr, t));
    };
    const float z = zb(idx);
    if (u_row == nullptr) return z;
    const float up = (idx < S - 1) ? __fmul_rn(0.5f, __fadd_rn(zb(idx + 1), z)) : z;
    const float lo = (idx > 0) ? __fmul_rn(0.5f, __fadd_rn(z, zb(idx - 1))) : z;
    return __fadd_rn(lo, __fmul_rn(__fsub_rn(up, lo), u_row[idx]));
}

__global__ void eik_points_fwd_kernel(const float* __restrict__ cam_loc, const float* __restrict__ dirs, const float* __restrict__ sd,
                                      const float* __restrict__ t_vals, const float* __restrict__ u, const int64_t* __restrict__ eik_idx,
                                      const float* __restrict__ uni, int R, int S, float cam_dist, float half_range,
                                      float* __restrict__ pts, float* __restrict__ z_out)
{
    const int b = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const size_t g = (size_t)b * R + r;
    const int idx = (int)eik_idx[g];
    const float z = eik_depth(sd[b], cam_dist, half_range, t_vals, S, idx, u ? u + g * S : nullptr);
    z_out[g] = z;
    float* pu = pts + ((size_t)b * 2 * R + r) * 3;
    float* pn = pts + ((size_t)b * 2 * R + R + r) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) { pu[c] = uni[g * 3 + c]; pn[c] = __fadd_rn(cam_loc[b * 3 + c], __fmul_rn(z, dirs[g * 3 + c])); }
}

// pts_bar [B,2R,3] -> dirs_bar [B,R,3] (written), cam_loc_bar [B,3], sd_bar [B] (atomics; zeroed by the launcher)
__global__ void eik_points_bwd_kernel(const float* __restrict__ dirs, const float* __restrict__ z_eik, const float* __restrict__ pts_bar,
                                      int R, float cam_dist, float* __restrict__ dirs_bar, float* __restrict__ acc /*[B][4]*/)
{
    const int b = blockIdx.y, r = blockIdx.x * blockDim.x + threadIdx.x;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < R) {
        const size_t g = (size_t)b * R + r;
        const float* pb = pts_bar + ((size_t)b * 2 * R + R + r) * 3;
        const float z = z_eik[g];
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) { v[c] = pb[c]; dirs_bar[g * 3 + c] = z * pb[c]; dot += dirs[g * 3 + c] * pb[c]; }
        v[3] = cam_dist * dot;                       // d z_eik / d scale_dist = cam_dist for every sample (convex combination of bin edges)
    }
    __shared__ float red[8][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float s = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
        atomicAdd(acc + b * 4 + threadIdx.x, s);
    }
}

}  // namespace scrays

extern "C" int sc_eikonal_points_forward(const float* cam_loc, const float* ray_dirs, const float* scale_dist, const float* t_vals,
                                         const float* jitter, const int64_t* eik_idx, const float* uniform_pts, int batch, int n_rays,
                                         int n_samples, float cam_dist, float half_range, float* points, float* z_eik,
                                         cudaStream_t stream)
{
    if (batch <= 0 || n_rays <= 0) return 0;
    if (!cam_loc || !ray_dirs || !scale_dist || !t_vals || !eik_idx || !uniform_pts || !points || !z_eik || n_samples <= 0)
        return (int)cudaErrorInvalidValue;
    dim3 grid((n_rays + 255) / 256, batch);
    scrays::eik_points_fwd_kernel<<<grid, 256, 0, stream>>>(cam_loc, ray_dirs, scale_dist, t_vals, jitter, eik_idx, uniform_pts, n_rays,
                                                            n_samples, cam_dist, half_range, points, z_eik);
    return (int)cudaGetLastError();
}

// acc: batch * 4 floats (cam_loc_bar xyz, scale_dist_bar), zeroed here
extern "C" int sc_eikonal_points_backward(const float* ray_dirs, const float* z_eik, const float* points_bar, int batch, int n_rays,
                                          float cam_dist, float* ray_dirs_bar, float* acc, cudaStream_t stream)
{
    if (batch <= 0 || n_rays <= 0) return 0;
    if (!ray_dirs || !z_eik || !points_bar || !ray_dirs_bar || !acc) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(acc, 0, (size_t)batch * 4 * sizeof(float), stream);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((n_rays + 255) / 256, batch);
    scrays::eik_points_bwd_kernel<<<grid, 256, 0, stream>>>(ray_dirs, z_eik, points_bar, n_rays, cam_dist, ray_dirs_bar, acc);
    return (int)cudaGetLastError();
}
