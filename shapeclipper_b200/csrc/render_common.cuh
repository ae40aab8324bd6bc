// render_common.cuh — shared definitions of the fused SDF/RGB-MLP volume-render kernels (sm_100a).
//
// Reference path replaced: model/renderer.py:57-209 (Renderer.forward / volume_rendering),
// model/implicit.py:7-239 (posenc, SDFNetwork, RGBNetwork, LaplaceDensity).
//
// One persistent CTA (256 threads) per SM walks tiles of M_TILE = 128 sample points (whole rays). Per tile the
// MLPs run as a chain of register-tiled FP32 GEMMs [128 points x 64 outputs x K]: activations live in shared
// memory as k-major planes, the layer's weight matrix is TMA-bulk-copied (cp.async.bulk + mbarrier) into a
// double-buffered 16 KB slot while the previous layer computes. Per-point activations never go to a tensor in
// HBM: what the gradient pass / backward sweep needs is parked in a per-CTA scratch slot that is rewritten
// every tile and therefore stays L2-resident.
//
// FP32 FFMA is deliberate for this (parity) path: BASELINE.md measures 2e-4..5e-4 max-rel error with tf32
// operands and 2e-3..6e-3 with bf16, above the 1e-4 target.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scr {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kPtsPerLane = 4;
constexpr int M_TILE = 32 * kPtsPerLane;   // 128 points per tile
constexpr int LD = M_TILE + 4;             // row stride of the k-major activation planes (floats)
constexpr int NPE = 39;                    // positional-encoding width (3 + 3*2*6)
constexpr int NPE_PAD = 40;
constexpr int P_ROWS = 48;                 // rows allocated for the 39-row planes (wgrad reads 3 x 16 rows)
constexpr int HID = 64;

// ---------------------------------------------------------------------------------------------------------
// Packed weight blob (floats). *_T = [in][out] (forward products), *_N = [out][in] (transposed products),
// *_N40 = [out=64][in padded to 40]. 1/sqrt(2) of the skip layers is folded into B1/A1/Z1, B2/A2/Z2.
enum Seg : int {
    A0T = 0, B1T, A1T, B2T, A2T, W3T, W4T, W5FT, V0PT, V0FT, V1T, V2T,
    W4N, W3N, B2N, B1N, W5FN, V2N, V1N, V0FN,
    A0N40, A1N40, A2N40, V0PN40,
    NSEG
};
constexpr int kSeg39 = NPE * HID;          // 2496 floats
constexpr int kSeg64 = HID * HID;          // 4096
constexpr int kSegN40 = HID * NPE_PAD;     // 2560

__host__ __device__ constexpr int seg_floats(int s) {
    return (s == A0T || s == A1T || s == A2T || s == V0PT) ? kSeg39
         : (s == A0N40 || s == A1N40 || s == A2N40 || s == V0PN40) ? kSegN40 : kSeg64;
}
__host__ __device__ constexpr int seg_offset(int s) {
    int o = 0;
    for (int i = 0; i < s; ++i) o += seg_floats(i);
    return o;
}
constexpr int kSegTotal = seg_offset(NSEG);

// small constants, copied once per CTA into shared memory
constexpr int C_W5 = 0, C_B3 = 64, C_B4 = 128, C_B5F = 192, C_C1R = 256, C_C2R = 320, C_V3 = 384 /*[3][64]*/,
              C_C3R = 576 /*3 (+1 pad)*/, C_B5 = 580 /*1 (+3 pad)*/, kConstFloats = 584;
constexpr int kConstOffset = kSegTotal;
// latent matrices [out][in] + the biases they combine with (used by the latent-bias kernels only)
constexpr int L_Z0 = 0, L_Z1 = 4096, L_Z2 = 8192, L_V0Z = 12288, L_B0 = 16384, L_B1 = 16448, L_B2 = 16512, L_C0R = 16576,
              kLatentFloats = 16640;
constexpr int kLatentOffset = kConstOffset + kConstFloats;
constexpr int kBlobFloats = kLatentOffset + kLatentFloats;

// ---------------------------------------------------------------------------------------------------------
// Folded gradient layout (per-CTA partial and reduced), natural [out][in]:
constexpr int G_A0 = 0;                       // [64][39]
constexpr int G_B1 = G_A0 + 64 * 39;          // [64][64]
constexpr int G_A1 = G_B1 + 4096;
constexpr int G_B2 = G_A1 + 64 * 39;
constexpr int G_A2 = G_B2 + 4096;
constexpr int G_W3 = G_A2 + 64 * 39;
constexpr int G_W4 = G_W3 + 4096;
constexpr int G_W5F = G_W4 + 4096;
constexpr int G_V0P = G_W5F + 4096;           // [64][39]
constexpr int G_V0F = G_V0P + 64 * 39;
constexpr int G_V1 = G_V0F + 4096;
constexpr int G_V2 = G_V1 + 4096;
constexpr int G_V3 = G_V2 + 4096;             // [3][64]
constexpr int G_W5 = G_V3 + 192;              // [64]
constexpr int G_B3 = G_W5 + 64;
constexpr int G_B4 = G_B3 + 64;
constexpr int G_B5F = G_B4 + 64;
constexpr int G_C1R = G_B5F + 64;
constexpr int G_C2R = G_C1R + 64;
constexpr int G_C3R = G_C2R + 64;             // 3 (+1)
constexpr int G_B5 = G_C3R + 4;               // 1 (+3): sdf bias
constexpr int G_BETA = G_B5 + 4;              // 1 (+3): d/d(effective beta)
constexpr int kGradFloats = G_BETA + 4;

// per-image bias adjoints (global atomics): [B][kCbRows][64]
constexpr int CB_C0 = 0, CB_C1 = 1, CB_C2 = 2, CB_RGB = 3, CB_C0D = 4, CB_C1D = 5, CB_C2D = 6, kCbRows = 7;

// ---------------------------------------------------------------------------------------------------------
// Per-CTA scratch ("stash") rows, each LD floats. Forward needs H only.
constexpr int ST_H = 0;             // H0..H4       5*64
constexpr int ST_FWD_ROWS = 320;
constexpr int ST_Q = 320;           // Q0..Q3       4*64
constexpr int ST_FEAT = 576;        // 64
constexpr int ST_R = 640;           // R0..R2       3*64
constexpr int ST_GPE = 832;         // 40
constexpr int ST_FB = 872;          // feat_bar 64
constexpr int ST_SB = 936;          // SB0..SB4     5*64
constexpr int ST_BWD_ROWS = 1256;

// ---------------------------------------------------------------------------------------------------------
// Shared-memory map (floats)
constexpr int SM_X = 0;
constexpr int SM_Y = SM_X + HID * LD;
constexpr int SM_Z = SM_Y + HID * LD;
constexpr int SM_U = SM_Z + HID * LD;
constexpr int SM_P = SM_U + HID * LD;              // positional encoding plane, P_ROWS rows
constexpr int SM_W = SM_P + P_ROWS * LD;           // 2 weight slots of 4096 floats
constexpr int SM_CONST = SM_W + 2 * 4096;          // kConstFloats
constexpr int SM_CB = SM_CONST + kConstFloats;     // per-image biases of the current tile [4][64]
constexpr int SM_PT = SM_CB + 256;                 // per-point vectors, kPtVecs x M_TILE
enum PtVec : int {
    PV_Z = 0, PV_SGN, PV_SDF, PV_SIG, PV_CF, PV_UN, PV_GX0, PV_GX1, PV_GX2, PV_NS0, PV_NS1, PV_NS2,
    PV_COL0, PV_COL1, PV_COL2, PV_X0, PV_X1, PV_X2,
    PV_SDFB, PV_GXB0, PV_GXB1, PV_GXB2, PV_CB0, PV_CB1, PV_CB2, PV_ZB, PV_XTB0, PV_XTB1, PV_XTB2, PV_W, PV_TMP,
    kPtVecs
};
constexpr int SM_RAY = SM_PT + kPtVecs * M_TILE;   // per-ray scratch (<= 32 rays x 32 floats)
constexpr int SM_FLOATS = SM_RAY + 32 * 32;
constexpr int SM_BAR_BYTES = 32;                   // 2 mbarriers (+pad), placed after the float area
constexpr size_t kSmemBytes = (size_t)SM_FLOATS * 4 + SM_BAR_BYTES;

// ---------------------------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Weight pipeline: a static per-tile sequence of blob segments; slot (n & 1) holds the n-th weight matrix.
struct WeightPipe {
    const float* blob;
    float* slots;          // shared: 2 x 4096 floats
    uint64_t* bars;        // shared: 2 mbarriers
    const int8_t* seq;     // segment ids, length seq_len (repeats every tile)
    int seq_len;
    uint32_t n;            // matrices consumed so far by this CTA
    int pos;               // position in seq of matrix n

    __device__ __forceinline__ void issue(uint32_t idx, int seg) {
        uint64_t* bar = bars + (idx & 1);
        const uint32_t bytes = (uint32_t)seg_floats(seg) * 4u;
        mbar_expect_tx(bar, bytes);
        tma_bulk_g2s(slots + (idx & 1) * 4096, blob + seg_offset_rt(seg), bytes, bar);
    }
    __device__ static int seg_offset_rt(int s) {
        int o = 0;
#pragma unroll 1
        for (int i = 0; i < s; ++i) o += seg_floats(i);
        return o;
    }
    __device__ __forceinline__ void prologue() {
        n = 0; pos = 0;
        if (threadIdx.x == 0) issue(0, seq[0]);
    }
    // Called by all threads. Starts with a CTA barrier (retires every read of the other slot and publishes the
    // previous epilogue's stores), starts the copy of the NEXT matrix and returns the current one.
    __device__ __forceinline__ const float* acquire() {
        __syncthreads();
        const uint32_t cur = n;
        const int nxt = (pos + 1 == seq_len) ? 0 : pos + 1;
        if (threadIdx.x == 0) issue(cur + 1, seq[nxt]);
        mbar_wait(bars + (cur & 1), (cur >> 1) & 1);
        n = cur + 1;
        pos = nxt;
        return slots + (cur & 1) * 4096;
    }
    // One copy is always in flight: wait for it before the CTA exits.
    __device__ __forceinline__ void drain() {
        mbar_wait(bars + (n & 1), (n >> 1) & 1);
        __syncthreads();
    }
};

// ---------------------------------------------------------------------------------------------------------
// Register-tiled GEMM pieces. Lane l owns points 4l..4l+3 of the tile, warp w owns outputs 8w..8w+7.
// acc[i][j] += sum_k A[k][4l+i] * W[k][8w+j]          (A: shared k-major plane, W: shared [K][64])
__device__ __forceinline__ void gemm64_step(float (&acc)[4][8], const float* __restrict__ a, const float* __restrict__ w, int k)
{
    const float4 p = *reinterpret_cast<const float4*>(a + k * LD);
    const float4 w0 = *reinterpret_cast<const float4*>(w + k * HID);
    const float4 w1 = *reinterpret_cast<const float4*>(w + k * HID + 4);
    const float av[4] = {p.x, p.y, p.z, p.w};
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
}
// The k loop is kept a REAL loop (4 steps per trip): fully unrolled, the ~50 GEMM call sites of the backward kernel
// make >0.5 MB of straight-line SASS and the SM stalls on instruction fetch (ncu: stall_no_inst 36 % of samples).
__device__ __forceinline__ void gemm64(float (&acc)[4][8], const float* __restrict__ A, int K,
                                       const float* __restrict__ W, int lane, int warp)
{
    const float* a = A + 4 * lane;
    const float* w = W + 8 * warp;
    const int K4 = K & ~3;
#pragma unroll 1
    for (int k = 0; k < K4; k += 4) {
        gemm64_step(acc, a, w, k); gemm64_step(acc, a, w, k + 1); gemm64_step(acc, a, w, k + 2); gemm64_step(acc, a, w, k + 3);
    }
#pragma unroll 1
    for (int k = K4; k < K; ++k) gemm64_step(acc, a, w, k);
}
// 40-wide output (39 used): warp w owns outputs 5w..5w+4.  W: shared [K=64][40]
__device__ __forceinline__ void gemm40_step(float (&acc)[4][5], const float* __restrict__ a, const float* __restrict__ w, int k)
{
    const float4 p = *reinterpret_cast<const float4*>(a + k * LD);
    const float av[4] = {p.x, p.y, p.z, p.w};
    float wv[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) wv[j] = w[k * NPE_PAD + j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
}
__device__ __forceinline__ void gemm40(float (&acc)[4][5], const float* __restrict__ A, int K,
                                       const float* __restrict__ W, int lane, int warp)
{
    const float* a = A + 4 * lane;
    const float* w = W + 5 * warp;
#pragma unroll 1
    for (int k = 0; k < K; k += 4) {
        gemm40_step(acc, a, w, k); gemm40_step(acc, a, w, k + 1); gemm40_step(acc, a, w, k + 2); gemm40_step(acc, a, w, k + 3);
    }
}

__device__ __forceinline__ void zero(float (&acc)[4][8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}
__device__ __forceinline__ void zero5(float (&acc)[4][5]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[i][j] = 0.f;
}

// thread-tile <-> k-major plane (shared or global stash; both have row stride LD)
__device__ __forceinline__ void tile_store(float* __restrict__ plane, const float (&v)[4][8], int lane, int warp) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(plane + (8 * warp + j) * LD + 4 * lane) = make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
}
__device__ __forceinline__ void tile_load(const float* __restrict__ plane, float (&v)[4][8], int lane, int warp) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 t = *reinterpret_cast<const float4*>(plane + (8 * warp + j) * LD + 4 * lane);
        v[0][j] = t.x; v[1][j] = t.y; v[2][j] = t.z; v[3][j] = t.w;
    }
}
// same, for the global per-CTA stash: L2-only accesses (never the non-coherent read-only path)
__device__ __forceinline__ void stash_store(float* plane, const float (&v)[4][8], int lane, int warp) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
        __stcg(reinterpret_cast<float4*>(plane + (8 * warp + j) * LD + 4 * lane), make_float4(v[0][j], v[1][j], v[2][j], v[3][j]));
}
__device__ __forceinline__ void stash_load(const float* plane, float (&v)[4][8], int lane, int warp) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(plane + (8 * warp + j) * LD + 4 * lane));
        v[0][j] = t.x; v[1][j] = t.y; v[2][j] = t.z; v[3][j] = t.w;
    }
}
// cooperative copy of `rows` rows (LD floats each, 16B aligned) global -> shared
__device__ __forceinline__ void plane_copy(float* __restrict__ dst, const float* __restrict__ src, int rows) {
    const int n4 = rows * (LD / 4);
    const float4* s = reinterpret_cast<const float4*>(src);
    float4* d = reinterpret_cast<float4*>(dst);
    for (int i = threadIdx.x; i < n4; i += kThreads) d[i] = __ldcg(s + i);
}

// softplus(beta=100, threshold=20) and its derivatives, from the activation value h alone:
//   E = exp(-100 h), s = softplus' = 1 - E, t = softplus'' = 100 s E
__device__ __forceinline__ float softplus100(float a) {
    const float t = 100.f * a;
    return (t > 20.f) ? a : 0.01f * __logf(1.f + __expf(t));
}
__device__ __forceinline__ float sp_slope(float h) { return 1.f - __expf(-100.f * h); }
__device__ __forceinline__ void sp_slope_curv(float h, float& s, float& t) {
    const float E = __expf(-100.f * h);
    s = 1.f - E;
    t = 100.f * s * E;
}

// sum of v over the 32 lanes (all lanes get the result)
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// wgrad: D[o][i] += sum_p L[o][p] * R[i][p] over the tile's points (L, R shared k-major planes of 64 / NR rows).
// Thread (to = t/16, ti = t%16) owns D rows {to + 16 a}, columns {ti + 16 b}; the result is added into the per-CTA
// partial gradient `dst` (row stride ld_dst, columns < ncols kept). NRB = number of 16-column groups (4 or 3).
template <int NRB>
__device__ __forceinline__ void wgrad(const float* __restrict__ L, const float* __restrict__ R,
                                      float* __restrict__ dst, int ld_dst, int ncols, int npts)
{
    const int to = threadIdx.x >> 4, ti = threadIdx.x & 15;
    float d[4][NRB];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < NRB; ++b) d[a][b] = 0.f;
#pragma unroll 1
    for (int p = 0; p < npts; p += 4) {
        float4 l[4], r[NRB];
#pragma unroll
        for (int a = 0; a < 4; ++a) l[a] = *reinterpret_cast<const float4*>(L + (to + 16 * a) * LD + p);
#pragma unroll
        for (int b = 0; b < NRB; ++b) r[b] = *reinterpret_cast<const float4*>(R + (ti + 16 * b) * LD + p);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < NRB; ++b) {
                d[a][b] = fmaf(l[a].x, r[b].x, d[a][b]);
                d[a][b] = fmaf(l[a].y, r[b].y, d[a][b]);
                d[a][b] = fmaf(l[a].z, r[b].z, d[a][b]);
                d[a][b] = fmaf(l[a].w, r[b].w, d[a][b]);
            }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < NRB; ++b) {
            const int col = ti + 16 * b;
            if (col < ncols) dst[(to + 16 * a) * ld_dst + col] += d[a][b];
        }
}

}  // namespace scr
