// chamfer.cu — brute-force bidirectional nearest neighbour (squared distance + int32 argmin) for sm_100a.
//
// Replaces /root/reference/external/chamfer3D/chamfer3D.cu:12-154 (NmDistanceKernel x2) and :155-195
// (NmDistanceGradKernel x2). Result definition (bit-exact): with (x,y,z) = candidate - query,
//     d = fmaf(z, z, fmaf(x, x, y*y)),   winner = min d, lowest candidate index among exact ties.
//
// Design (not a port of the reference's 512-thread / 16-block scheme, which fills <=16 SMs at batch 1):
//  * every CTA owns 256*Q queries (Q per thread, in registers) and one slice of the candidates, so the grid
//    fills all 148 SMs whatever the batch size; slices merge through one 64-bit atomicMin on
//    (float_bits(d) << 32 | index) — for d >= 0 the bit pattern is monotone, so the u64 minimum IS
//    "min d, then lowest index" and the merge order is irrelevant;
//  * candidates are staged in shared memory as SoA planes and read back with broadcast LDS.128; two candidates
//    are evaluated per instruction with the packed FP32x2 ops of sm_100 (FADD2/FMUL2/FFMA2: same IEEE rounding
//    per lane as the scalar chain), which halves the issue slots per pair;
//  * the running minimum is value-only (one FMNMX per candidate pair); the argmin is recovered by remembering
//    the first 32-candidate group that strictly improved the minimum and re-scanning just that group.
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "sc_b200.h"

namespace {

constexpr int kThreads = 256;
constexpr int kQ = 4;             // queries per thread
constexpr int kStage = 2048;      // candidates staged per shared-memory round (3 planes x 8 KB)
constexpr int kGroup = 32;        // argmin bookkeeping granularity

__device__ __forceinline__ float pair_dist(float cx, float cy, float cz, float qx, float qy, float qz) {
    const float x = cx - qx, y = cy - qy, z = cz - qz;
    return __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
}

// grid: (query tiles, candidate slices, batch*2); blockIdx.z & 1 selects the direction.
__global__ void __launch_bounds__(kThreads)
chamfer_nn_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int n1, int n2,
                  int slice1, int slice2, unsigned long long* __restrict__ packed1,
                  unsigned long long* __restrict__ packed2)
{
    __shared__ __align__(16) float sx[kStage];
    __shared__ __align__(16) float sy[kStage];
    __shared__ __align__(16) float sz[kStage];

    const int dir = blockIdx.z & 1;
    const int b = blockIdx.z >> 1;
    const int n = dir ? n2 : n1;                  // queries
    const int m = dir ? n1 : n2;                  // candidates
    const int slice = dir ? slice2 : slice1;      // candidates per slice (multiple of kStage)
    const float* __restrict__ q = (dir ? xyz2 : xyz1) + (size_t)b * n * 3;
    const float* __restrict__ c = (dir ? xyz1 : xyz2) + (size_t)b * m * 3;
    unsigned long long* __restrict__ out = (dir ? packed2 : packed1) + (size_t)b * n;

    const int q0 = blockIdx.x * (kThreads * kQ);
    const int c_lo = blockIdx.y * slice;
    if (q0 >= n || c_lo >= m) return;
    const int c_hi = min(m, c_lo + slice);

    float2 nqx[kQ], nqy[kQ], nqz[kQ];             // (-q, -q) so that c - q is one packed add
    float best[kQ], seen[kQ];
    int grp[kQ];
#pragma unroll
    for (int i = 0; i < kQ; ++i) {
        const int j = q0 + i * kThreads + threadIdx.x;
        float x = 0.f, y = 0.f, z = 0.f;
        if (j < n) { x = q[j * 3 + 0]; y = q[j * 3 + 1]; z = q[j * 3 + 2]; }
        nqx[i] = make_float2(-x, -x); nqy[i] = make_float2(-y, -y); nqz[i] = make_float2(-z, -z);
        best[i] = INFINITY; seen[i] = INFINITY; grp[i] = c_lo / kGroup;
    }

    for (int base = c_lo; base < c_hi; base += kStage) {
        const int cnt = min(kStage, c_hi - base);
        __syncthreads();
        for (int k = threadIdx.x; k < kStage; k += kThreads) {
            float x = INFINITY, y = INFINITY, z = INFINITY;   // padding never wins: d = +inf
            if (k < cnt) { const float* p = c + (size_t)(base + k) * 3; x = p[0]; y = p[1]; z = p[2]; }
            sx[k] = x; sy[k] = y; sz[k] = z;
        }
        __syncthreads();
        const int groups = (cnt + kGroup - 1) / kGroup;
        for (int g = 0; g < groups; ++g) {
#pragma unroll
            for (int k = 0; k < kGroup; k += 4) {
                const float4 X = *reinterpret_cast<const float4*>(&sx[g * kGroup + k]);
                const float4 Y = *reinterpret_cast<const float4*>(&sy[g * kGroup + k]);
                const float4 Z = *reinterpret_cast<const float4*>(&sz[g * kGroup + k]);
#pragma unroll
                for (int i = 0; i < kQ; ++i) {
                    const float2 xa = __fadd2_rn(make_float2(X.x, X.y), nqx[i]);
                    const float2 ya = __fadd2_rn(make_float2(Y.x, Y.y), nqy[i]);
                    const float2 za = __fadd2_rn(make_float2(Z.x, Z.y), nqz[i]);
                    const float2 xb = __fadd2_rn(make_float2(X.z, X.w), nqx[i]);
                    const float2 yb = __fadd2_rn(make_float2(Y.z, Y.w), nqy[i]);
                    const float2 zb = __fadd2_rn(make_float2(Z.z, Z.w), nqz[i]);
                    const float2 da = __ffma2_rn(za, za, __ffma2_rn(xa, xa, __fmul2_rn(ya, ya)));
                    const float2 db = __ffma2_rn(zb, zb, __ffma2_rn(xb, xb, __fmul2_rn(yb, yb)));
                    best[i] = fminf(best[i], fminf(da.x, da.y));
                    best[i] = fminf(best[i], fminf(db.x, db.y));
                }
            }
            const int gid = base / kGroup + g;
#pragma unroll
            for (int i = 0; i < kQ; ++i) {
                if (best[i] < seen[i]) { seen[i] = best[i]; grp[i] = gid; }
            }
        }
    }

    // argmin recovery: first candidate of the remembered group whose distance equals the minimum.
#pragma unroll
    for (int i = 0; i < kQ; ++i) {
        const int j = q0 + i * kThreads + threadIdx.x;
        if (j >= n) continue;
        const float qx = -nqx[i].x, qy = -nqy[i].x, qz = -nqz[i].x;
        const int k0 = grp[i] * kGroup;
        const int k1 = min(c_hi, k0 + kGroup);
        int win = k0;
        float dwin = pair_dist(c[(size_t)k0 * 3], c[(size_t)k0 * 3 + 1], c[(size_t)k0 * 3 + 2], qx, qy, qz);
        if (!(dwin == best[i])) {
            for (int k = k0 + 1; k < k1; ++k) {
                const float d = pair_dist(c[(size_t)k * 3], c[(size_t)k * 3 + 1], c[(size_t)k * 3 + 2], qx, qy, qz);
                if (d == best[i]) { win = k; dwin = d; break; }
            }
        }
        const unsigned long long key =
            ((unsigned long long)__float_as_uint(dwin) << 32) | (unsigned long long)(unsigned int)win;
        atomicMin(out + j, key);
    }
}

__global__ void __launch_bounds__(256)
chamfer_unpack_kernel(const unsigned long long* __restrict__ packed, float* __restrict__ dist,
                      int32_t* __restrict__ idx, size_t count)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const unsigned long long v = packed[i];
    dist[i] = __uint_as_float((unsigned int)(v >> 32));
    idx[i] = (int32_t)(unsigned int)(v & 0xffffffffull);
}

// reference: chamfer3D.cu:155-174. One thread per (batch, point); scatter with float atomics.
__global__ void __launch_bounds__(256)
chamfer_grad_kernel(int n, int m, const float* __restrict__ p1, const float* __restrict__ p2,
                    const float* __restrict__ gdist, const int32_t* __restrict__ idx,
                    float* __restrict__ g1, float* __restrict__ g2, size_t total)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const size_t b = t / n;
    const size_t a = t * 3;
    const size_t o = (b * m + (size_t)idx[t]) * 3;
    const float g = gdist[t] * 2.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = g * (p1[a + c] - p2[o + c]);
        atomicAdd(g1 + a + c, v);
        atomicAdd(g2 + o + c, -v);
    }
}

int pick_slice(int n_queries_tiles_total, int m) {
    // enough CTAs for >= 2 waves of 148 SMs x 2 resident CTAs, slices a multiple of the staging depth
    const int want = (4 * 148 + n_queries_tiles_total - 1) / n_queries_tiles_total;
    const int stages = (m + kStage - 1) / kStage;
    int n_slices = want < 1 ? 1 : want;
    if (n_slices > stages) n_slices = stages;
    if (n_slices < 1) n_slices = 1;
    const int stages_per = (stages + n_slices - 1) / n_slices;
    return stages_per * kStage;
}

}  // namespace

extern "C" size_t sc_chamfer_workspace_bytes(int batch, int n, int m) {
    return (size_t)batch * ((size_t)n + (size_t)m) * sizeof(unsigned long long);
}

extern "C" int sc_chamfer_forward(const float* xyz1, const float* xyz2, int batch, int n, int m,
                                  float* dist1, float* dist2, int32_t* idx1, int32_t* idx2,
                                  void* workspace, size_t workspace_bytes, cudaStream_t stream)
{
    if (batch <= 0 || n <= 0 || m <= 0) return (int)cudaSuccess;   // nothing to write (reference: zero-trip loops)
    if (workspace == nullptr || workspace_bytes < sc_chamfer_workspace_bytes(batch, n, m))
        return (int)cudaErrorInvalidValue;
    unsigned long long* packed1 = reinterpret_cast<unsigned long long*>(workspace);
    unsigned long long* packed2 = packed1 + (size_t)batch * n;
    cudaError_t err = cudaMemsetAsync(workspace, 0xff, sc_chamfer_workspace_bytes(batch, n, m), stream);
    if (err != cudaSuccess) return (int)err;

    const int per_cta = kThreads * kQ;
    const int tiles1 = (n + per_cta - 1) / per_cta, tiles2 = (m + per_cta - 1) / per_cta;
    const int slice1 = pick_slice(batch * (tiles1 + tiles2), m);   // direction 0: candidates = xyz2
    const int slice2 = pick_slice(batch * (tiles1 + tiles2), n);   // direction 1: candidates = xyz1
    const int ns1 = (m + slice1 - 1) / slice1, ns2 = (n + slice2 - 1) / slice2;
    dim3 grid(tiles1 > tiles2 ? tiles1 : tiles2, ns1 > ns2 ? ns1 : ns2, batch * 2);
    chamfer_nn_kernel<<<grid, kThreads, 0, stream>>>(xyz1, xyz2, n, m, slice1, slice2, packed1, packed2);
    const size_t c1 = (size_t)batch * n, c2 = (size_t)batch * m;
    chamfer_unpack_kernel<<<(unsigned)((c1 + 255) / 256), 256, 0, stream>>>(packed1, dist1, idx1, c1);
    chamfer_unpack_kernel<<<(unsigned)((c2 + 255) / 256), 256, 0, stream>>>(packed2, dist2, idx2, c2);
    return (int)cudaGetLastError();
}

extern "C" int sc_chamfer_backward(const float* xyz1, const float* xyz2, int batch, int n, int m,
                                   const float* graddist1, const float* graddist2,
                                   const int32_t* idx1, const int32_t* idx2,
                                   float* gradxyz1, float* gradxyz2, cudaStream_t stream)
{
    if (batch <= 0 || n <= 0 || m <= 0) return (int)cudaSuccess;
    const size_t c1 = (size_t)batch * n, c2 = (size_t)batch * m;
    chamfer_grad_kernel<<<(unsigned)((c1 + 255) / 256), 256, 0, stream>>>(n, m, xyz1, xyz2, graddist1, idx1,
                                                                        gradxyz1, gradxyz2, c1);
    chamfer_grad_kernel<<<(unsigned)((c2 + 255) / 256), 256, 0, stream>>>(m, n, xyz2, xyz1, graddist2, idx2,
                                                                        gradxyz2, gradxyz1, c2);
    return (int)cudaGetLastError();
}
