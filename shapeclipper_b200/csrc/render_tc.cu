// render_tc.cu — tensor-core edition: weight packing + FORWARD kernel (see render_tc.cuh).
#include <cuda_runtime.h>
#include <stdint.h>

#include "render_tc_tile.cuh"

namespace sct {

struct PackSrcTc { const float* w[10]; const float* b[10]; };

// element (row n, col k) of segment `seg` (64 x 64, zero padded) from the nn.Linear tensors
__device__ __forceinline__ float seg_elem(const PackSrcTc& s, int seg, int n, int k)
{
    const float r2 = 0.70710678118654752440f;
    // natural: value = W[out = n][in = k] ; transposed: value = W[out = k][in = n]
    bool tr = seg >= W4T;
    int t, in_dim, c0, n_in = 64, row0 = 0; float sc = 1.f;
    switch (seg) {
        case A0N: case A0T: t = 0; in_dim = 103; c0 = 0; n_in = 39; break;
        case B1N: case B1T: t = 1; in_dim = 167; c0 = 0; sc = r2; break;
        case A1N: case A1T: t = 1; in_dim = 167; c0 = 64; n_in = 39; sc = r2; break;
        case B2N: case B2T: t = 2; in_dim = 167; c0 = 0; sc = r2; break;
        case A2N: case A2T: t = 2; in_dim = 167; c0 = 64; n_in = 39; sc = r2; break;
        case W3N: case W3T: t = 3; in_dim = 64; c0 = 0; break;
        case W4N: case W4T: t = 4; in_dim = 64; c0 = 0; break;
        case W5FN: case W5FT: t = 5; in_dim = 64; c0 = 0; row0 = 1; break;
        case V0PN: case V0PT: t = 6; in_dim = 167; c0 = 0; n_in = 39; break;
        case V0FN: case V0FT: t = 6; in_dim = 167; c0 = 103; break;
        case V1N: case V1T: t = 7; in_dim = 64; c0 = 0; break;
        default /*V2*/: t = 8; in_dim = 64; c0 = 0; break;
    }
    const int o = tr ? k : n, i = tr ? n : k;
    if (i >= n_in) return 0.f;
    return sc * s.w[t][(size_t)(o + row0) * in_dim + c0 + i];
}

__global__ void pack_tc_kernel(PackSrcTc s, uint8_t* __restrict__ blob, const float* __restrict__ fp32_tail)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < NSEG_TC * 64 * 64) {
        const int seg = idx / 4096, e = idx % 4096, n = e / 64, k = e % 64;
        const float v = seg_elem(s, seg, n, k);
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const int off = n * 128 + ((((k >> 3) ^ (n & 7))) << 4) + ((k & 7) << 1);
        uint8_t* base = blob + (size_t)seg * kWSegBytes;
        *reinterpret_cast<__nv_bfloat16*>(base + off) = hi;
        *reinterpret_cast<__nv_bfloat16*>(base + kWPlaneBytes + off) = lo;
    } else {
        const int j = idx - NSEG_TC * 64 * 64;                  // fp32 consts + latent matrices, copied from the FFMA blob
        if (j < scr::kConstFloats + scr::kLatentFloats)
            reinterpret_cast<float*>(blob + kTcConstOffsetBytes)[j] = fp32_tail[j];
    }
}

__device__ __forceinline__ void build_seq_tc(int8_t* seq, int& len, int mode, bool want_grad, bool want_feat)
{
    int n = 0;
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

__device__ __forceinline__ void tc_init_tile_struct(TileTC& T, uint8_t* smem, const uint8_t* blob, const int8_t* seq, int seq_len,
                                                    int n_act, int n_slots)
{
    for (int i = 0; i < kNumAct; ++i) T.act[i] = smem + SMB_ACT + (i < n_act ? i : n_act - 1) * kActBytes;
    float* f = reinterpret_cast<float*>(smem + SMB_F32);
    T.cst = f + SF_CONST; T.cb = f + SF_CB; T.pt = f + SF_PT; T.ray = f + SF_RAY; T.bias = f + SF_BIAS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMB_BAR);
    T.wr.blob = blob; T.wr.slots = smem + SMB_ACT + n_act * kActBytes; T.wr.wfull = bars; T.wr.wfree = bars + 4;
    T.wr.seq = seq; T.wr.seq_len = seq_len; T.wr.NS = n_slots;
    T.mma_done = bars + 8; T.mma_phase = 0;
    T.tid = threadIdx.x; T.lane = threadIdx.x & 31; T.warp = threadIdx.x >> 5;
    T.w0 = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0) == 0;          // warp-uniform for the compiler too
    T.wr.w0 = T.w0;
    T.wide = true; T.single = false;
    T.row = 32 * (T.warp & 3) + T.lane; T.ch = T.warp >> 2;
}

// SAVE (mode 0 only): a.saved receives every tile's activation planes + per-point vectors for sc_render_tc_backward
// PREC (ScRenderArgs::precision) is a TEMPLATE parameter: the single-MMA mode as a run-time branch in the issue path made the
// default (split) kernels 6 % slower (instruction-cache footprint of the issuing warp).
template <int MODE, bool SAVE, int PREC>
__global__ void __launch_bounds__(kThreads, 1) render_tc_fwd_kernel(const ScRenderArgs a, float* stash_base)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    int8_t* seq = reinterpret_cast<int8_t*>(reinterpret_cast<float*>(smem + SMB_F32) + SF_MISC);
    int& seq_len = *reinterpret_cast<int*>(reinterpret_cast<float*>(smem + SMB_F32) + SF_MISC + 16);
    const uint8_t* blob = reinterpret_cast<const uint8_t*>(a.blob);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMB_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const bool want_grad = (MODE == 0) || a.want_grad;
    const bool want_feat = (MODE == 0) || a.want_feat;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 9; ++i) scr::mbar_init(bars + i, 1);
        scr::mbar_fence_init();
        int len; build_seq_tc(seq, len, MODE, want_grad, want_feat); seq_len = len;
    }
    if ((threadIdx.x >> 5) == 0) sctc::tmem_alloc<512>(tmem_slot);
    {
        float* f = reinterpret_cast<float*>(smem + SMB_F32);
        const float* src = reinterpret_cast<const float*>(blob + kTcConstOffsetBytes);
        for (int i = threadIdx.x; i < scr::kConstFloats; i += kThreads) f[SF_CONST + i] = src[i];
        for (int i = threadIdx.x; i < 64; i += kThreads) {       // constant rows of the bias tables: b3 b4 | c1r c2r
            f[SF_BIAS + 3 * 64 + i] = src[scr::C_B3 + i]; f[SF_BIAS + 4 * 64 + i] = src[scr::C_B4 + i];
            f[SF_BIAS + 6 * 64 + i] = src[scr::C_C1R + i]; f[SF_BIAS + 7 * 64 + i] = src[scr::C_C2R + i];
        }
    }
    sctc::tc_fence_before();
    __syncthreads();
    sctc::tc_fence_after();

    TileTC T;
    tc_init_tile_struct(T, smem, blob, seq, seq_len, 4, kWSlots);
    if (PREC == 1) { T.single = true; T.wide = false; }              // 64-column accumulators: nothing to add in the epilogue
    T.tmem = *tmem_slot;
    T.stash = stash_base + (size_t)blockIdx.x * TS_PLANES_FWD * kStashPlane;
    T.S = (MODE == 0) ? a.n_samples : 1;
    T.rays_per_tile = (MODE == 0) ? M_TILE / a.n_samples : M_TILE;
    T.beta = (MODE == 0) ? fabsf(*a.beta_param) + a.beta_min : 1.f;
#ifdef SC_TC_TRACE
    T.trace = reinterpret_cast<long long*>(a.points_bar); T.trace_n = 0;
#endif

    const int per_tile = (MODE == 0) ? T.rays_per_tile : M_TILE;
    const int tiles_per_image = (a.n_per_image + per_tile - 1) / per_tile;
    const int total = a.batch * tiles_per_image;
    if ((int)blockIdx.x < total) {
        T.wr.prologue();
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            T.b = tile / tiles_per_image;
            T.first = (tile % tiles_per_image) * per_tile;
            __syncthreads();
            T.mark();
            if (SAVE) T.stash = reinterpret_cast<float*>(a.saved) + (size_t)tile * TS_SAVED_PLANES * kStashPlane;
            tc_tile_setup<MODE>(T, a);
            T.mark();
            tc_tile_forward<MODE, SAVE>(T, a, want_grad, want_feat);
            if (SAVE) saved_vectors<true>(T, T.stash + TS_SAVED_PV * kStashPlane);
            T.mark();

            if (MODE == 1) {
                if (T.tid < M_TILE) {
                    const int n = T.first + T.tid;
                    if (n < a.n_per_image) {
                        const size_t g = (size_t)T.b * a.n_per_image + n;
                        a.sdf[g] = T.pv(scr::PV_SDF)[T.tid];
                        if (a.grad != nullptr && want_grad) {
                            a.grad[g * 3 + 0] = T.pv(scr::PV_GX0)[T.tid];
                            a.grad[g * 3 + 1] = T.pv(scr::PV_GX1)[T.tid];
                            a.grad[g * 3 + 2] = T.pv(scr::PV_GX2)[T.tid];
                        }
                    }
                }
            } else {
                using namespace scr;
                T.ray[32 + T.tid] = 0.f;
                __syncthreads();
                if (T.tid < M_TILE) {
                    const int p = T.tid, S = T.S, rl = p / S;
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
                    if (S > 32) {
                        if (T.lane == 31) T.ray[T.warp] = incl;
                        asm volatile("bar.sync 1, 128;");
                        const int w0 = (p / S) * (S / 32);
                        for (int ww = w0; ww < T.warp; ++ww) before += T.ray[ww];
                        asm volatile("bar.sync 1, 128;");
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
                        for (int q = 0; q < 8; ++q) atomicAdd(&T.ray[32 + rl * 8 + q], vv[q]);
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
        T.wr.drain();
    }
    sctc::tc_fence_before();
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) sctc::tmem_dealloc<512>(T.tmem);
}

}  // namespace sct

using namespace sct;

extern "C" size_t sc_render_tc_blob_bytes(void) { return kTcBlobBytes; }
extern "C" size_t sc_render_tc_saved_bytes(int batch, int n_per_image, int n_samples) {
    if (batch <= 0 || n_per_image <= 0 || n_samples < 4 || n_samples > M_TILE || (M_TILE % n_samples) != 0) return 0;
    const int per_tile = M_TILE / n_samples;
    const size_t tiles = (size_t)batch * ((n_per_image + per_tile - 1) / per_tile);
    return tiles * TS_SAVED_PLANES * kStashPlane * sizeof(float);
}
extern "C" size_t sc_render_tc_scratch_bytes(int backward) {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return (size_t)n * (size_t)(backward ? TS_PLANES_BWD : TS_PLANES_FWD) * kStashPlane * sizeof(float);
}

// ffma_blob: the blob of sc_render_pack_weights for the same tensors (its fp32 const/latent tail is reused)
extern "C" int sc_render_tc_pack_weights(const float* const* w, const float* const* b, const float* ffma_blob, void* tc_blob,
                                         cudaStream_t stream)
{
    PackSrcTc s;
    for (int i = 0; i < 10; ++i) { s.w[i] = w[i]; s.b[i] = b[i]; }
    const int total = NSEG_TC * 64 * 64 + scr::kConstFloats + scr::kLatentFloats;
    pack_tc_kernel<<<(total + 255) / 256, 256, 0, stream>>>(s, reinterpret_cast<uint8_t*>(tc_blob), ffma_blob + scr::kConstOffset);
    return (int)cudaGetLastError();
}

extern "C" int sc_render_tc_forward(const ScRenderArgs* a, cudaStream_t stream)
{
    if (a == nullptr || a->blob == nullptr || a->cb == nullptr || a->scratch == nullptr) return (int)cudaErrorInvalidValue;
    if (a->mode == 0) {
        const int S = a->n_samples;
        if (S < 4 || S > M_TILE || (S % 4) != 0 || (M_TILE % S) != 0 || a->beta_param == nullptr) return (int)cudaErrorInvalidValue;
    } else if (a->mode != 1) return (int)cudaErrorInvalidValue;
    if (a->batch <= 0 || a->n_per_image <= 0) return 0;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_tile = a->mode == 0 ? M_TILE / a->n_samples : M_TILE;
    const long total = (long)a->batch * ((a->n_per_image + per_tile - 1) / per_tile);
    const int grid = total < sms ? (int)total : sms;
    cudaError_t err;
#define SC_LAUNCH_FWD(MODE_, SAVE_, PREC_) do { \
        err = cudaFuncSetAttribute(render_tc_fwd_kernel<MODE_, SAVE_, PREC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesTc); \
        if (err != cudaSuccess) return (int)err; \
        render_tc_fwd_kernel<MODE_, SAVE_, PREC_><<<grid, kThreads, kSmemBytesTc, stream>>>(*a, (float*)a->scratch); } while (0)
    const bool single = a->precision == 1;
    if (a->mode == 0 && a->saved != nullptr) { if (single) SC_LAUNCH_FWD(0, true, 1); else SC_LAUNCH_FWD(0, true, 0); }
    else if (a->mode == 0) { if (single) SC_LAUNCH_FWD(0, false, 1); else SC_LAUNCH_FWD(0, false, 0); }
    else { if (single) SC_LAUNCH_FWD(1, false, 1); else SC_LAUNCH_FWD(1, false, 0); }
#undef SC_LAUNCH_FWD
    return (int)cudaGetLastError();
}
