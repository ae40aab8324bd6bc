// clip_ops.cu — the non-GEMM pieces of the CLIP ViT image tower and the encoder orchestration (sm_100a).
// Architecture restated from openai/CLIP's published VisionTransformer (the package is an un-vendored, unpinned pip
// dependency of the reference — README.md:14, call sites CLIP_anno.py:16,166): conv1 patch embedding (no bias) ->
// [class token; patches] + positional embedding -> ln_pre -> L x { x += MHA(ln_1 x); x += W2 QuickGELU(W1 ln_2 x) } ->
// ln_post(token 0) -> @ proj. LayerNorm eps 1e-5; QuickGELU(x) = x sigmoid(1.702 x).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sc_b200.h"

extern "C" int sc_gemm_bf16_tc(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, int M, int N, int K,
                               const float* bias, const float* residual, int act, float scale, float* out_f32, void* out_hi,
                               void* out_lo, cudaStream_t stream);

namespace scclip {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void split_store(bf16* hi, bf16* lo, size_t i, float v) {
    const bf16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// images [B,3,S,S] fp32 -> patches [B*G*G, 3*P*P] (hi/lo bf16); column = c*P*P + py*P + px (conv1.weight flattening)
// Kp = 3*P*P rounded up to a multiple of 64 (zero padding; the packed conv1 weight is padded the same way)
__global__ void im2col_split_kernel(const float* __restrict__ img, int B, int S, int P, int Kp, bf16* __restrict__ hi,
                                    bf16* __restrict__ lo)
{
    const int G = S / P, Kc = 3 * P * P;
    const size_t total = (size_t)B * G * G * Kp;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int col = (int)(i % Kp);
        const size_t row = i / Kp;
        float v = 0.f;
        if (col < Kc) {
            const int px = col % P, py = (col / P) % P, c = col / (P * P);
            const int gx = (int)(row % G), gy = (int)((row / G) % G), b = (int)(row / ((size_t)G * G));
            v = img[(((size_t)b * 3 + c) * S + gy * P + py) * S + gx * P + px];
        }
        split_store(hi, lo, i, v);
    }
}

// one warp per row. mode 0: y = LN(x) -> hi/lo planes. mode 1: token assembly (+pos) then LN -> fp32 x (ln_pre).
__global__ void layernorm_kernel(const float* __restrict__ x, size_t row_stride, int rows, int W, const float* __restrict__ g,
                                 const float* __restrict__ bta, bf16* __restrict__ hi, bf16* __restrict__ lo,
                                 float* __restrict__ out_f32, const float* __restrict__ patch, const float* __restrict__ cls,
                                 const float* __restrict__ pos, int T)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float v[32];                                   // W <= 1024
    const int per = W / 32;
    float s = 0.f;
    for (int i = 0; i < per; ++i) {
        const int c = lane + 32 * i;
        float t;
        if (patch != nullptr) {                    // assemble token `row` of the sequence
            const int b = row / T, tk = row % T;
            t = (tk == 0 ? cls[c] : patch[((size_t)b * (T - 1) + tk - 1) * W + c]) + pos[(size_t)tk * W + c];
        } else {
            t = x[(size_t)row * row_stride + c];
        }
        v[i] = t; s += t;
    }
    const float mean = warp_sum(s) / W;
    float q = 0.f;
    for (int i = 0; i < per; ++i) { const float d = v[i] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / W + 1e-5f);
    for (int i = 0; i < per; ++i) {
        const int c = lane + 32 * i;
        const float y = (v[i] - mean) * rstd * g[c] + bta[c];
        if (out_f32) out_f32[(size_t)row * W + c] = y;
        if (hi) split_store(hi, lo, (size_t)row * W + c, y);
    }
}

// softmax(q k^T / sqrt(64)) v per (image, head); qkv fp32 [B*T, 3W] (q | k | v), head h = columns h*64..h*64+63.
// K and V of the head sit in shared memory with 16-byte-aligned rows (stride 68 floats: conflict-free LDS.128 for 8 lanes on
// consecutive keys); a warp takes TWO query rows per pass so that every K / V vector load feeds two accumulators.
constexpr int kAttLd = 68;
__global__ void attention_kernel(const float* __restrict__ qkv, int T, int W, bf16* __restrict__ hi, bf16* __restrict__ lo)
{
    extern __shared__ __align__(16) float sm[];
    float* Ks = sm;                          // [T][68]
    float* Vs = sm + (size_t)T * kAttLd;
    float* Ps = Vs + (size_t)T * kAttLd;     // per warp: P0[Tp] P1[Tp] Q0[64] Q1[64]
    const int Tp = (T + 3) & ~3;
    const int b = blockIdx.x, h = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float* base = qkv + (size_t)b * T * 3 * W + h * 64;
    for (int i = threadIdx.x; i < T * 16; i += blockDim.x) {                 // float4 granularity: 16 per row
        const int t = i >> 4, d4 = (i & 15) << 2;
        *reinterpret_cast<float4*>(Ks + t * kAttLd + d4) = *reinterpret_cast<const float4*>(base + (size_t)t * 3 * W + W + d4);
        *reinterpret_cast<float4*>(Vs + t * kAttLd + d4) = *reinterpret_cast<const float4*>(base + (size_t)t * 3 * W + 2 * W + d4);
    }
    __syncthreads();
    float* P0 = Ps + (size_t)warp * (2 * Tp + 128);
    float* P1 = P0 + Tp;
    float* Q0 = P1 + Tp;
    float* Q1 = Q0 + 64;
    for (int tq = 2 * warp; tq < T; tq += 2 * nw) {
        const bool two = tq + 1 < T;
        const float* q0 = base + (size_t)tq * 3 * W;
        const float* q1 = base + (size_t)(two ? tq + 1 : tq) * 3 * W;
        Q0[lane] = q0[lane] * 0.125f; Q0[32 + lane] = q0[32 + lane] * 0.125f;          // pre-scaled by 1/sqrt(64)
        Q1[lane] = q1[lane] * 0.125f; Q1[32 + lane] = q1[32 + lane] * 0.125f;
        __syncwarp();
        float mx0 = -INFINITY, mx1 = -INFINITY;
        for (int tk = lane; tk < T; tk += 32) {
            const float4* kr = reinterpret_cast<const float4*>(Ks + tk * kAttLd);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
            for (int d = 0; d < 16; ++d) {
                const float4 k = kr[d];
                const float4 x = reinterpret_cast<const float4*>(Q0)[d];
                const float4 y = reinterpret_cast<const float4*>(Q1)[d];
                a0 = fmaf(x.x, k.x, a0); a1 = fmaf(x.y, k.y, a1); a2 = fmaf(x.z, k.z, a2); a3 = fmaf(x.w, k.w, a3);
                c0 = fmaf(y.x, k.x, c0); c1 = fmaf(y.y, k.y, c1); c2 = fmaf(y.z, k.z, c2); c3 = fmaf(y.w, k.w, c3);
            }
            const float s0 = (a0 + a1) + (a2 + a3), s1 = (c0 + c1) + (c2 + c3);
            P0[tk] = s0; P1[tk] = s1;
            mx0 = fmaxf(mx0, s0); mx1 = fmaxf(mx1, s1);
        }
        mx0 = warp_max(mx0); mx1 = warp_max(mx1);
        float sum0 = 0.f, sum1 = 0.f;
        for (int tk = lane; tk < T; tk += 32) {
            const float e0 = __expf(P0[tk] - mx0), e1 = __expf(P1[tk] - mx1);
            P0[tk] = e0; P1[tk] = e1; sum0 += e0; sum1 += e1;
        }
        sum0 = warp_sum(sum0); sum1 = warp_sum(sum1);
        __syncwarp();
        // out[d] for d = 2 lane, 2 lane + 1 (one 8-byte V load per key feeds both queries)
        float o00 = 0.f, o01 = 0.f, o10 = 0.f, o11 = 0.f;
#pragma unroll 4
        for (int tk = 0; tk < T; ++tk) {
            const float2 v = *reinterpret_cast<const float2*>(Vs + tk * kAttLd + 2 * lane);
            const float p0 = P0[tk], p1 = P1[tk];
            o00 = fmaf(p0, v.x, o00); o01 = fmaf(p0, v.y, o01);
            o10 = fmaf(p1, v.x, o10); o11 = fmaf(p1, v.y, o11);
        }
        const float i0 = 1.f / sum0, i1 = 1.f / sum1;
        const size_t o = ((size_t)b * T + tq) * W + h * 64 + 2 * lane;
        split_store(hi, lo, o, o00 * i0); split_store(hi, lo, o + 1, o01 * i0);
        if (two) { split_store(hi, lo, o + W, o10 * i1); split_store(hi, lo, o + W + 1, o11 * i1); }
        __syncwarp();
    }
}

// rows of [rows, D]: y = x / max(|x|, 1e-12); also hi/lo planes for the similarity GEMM
__global__ void l2norm_kernel(const float* __restrict__ x, int rows, int D, float* __restrict__ y, bf16* __restrict__ hi,
                              bf16* __restrict__ lo)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    float s = 0.f;
    for (int c = lane; c < D; c += 32) { const float v = x[(size_t)row * D + c]; s += v * v; }
    const float inv = 1.f / fmaxf(sqrtf(warp_sum(s)), 1e-12f);
    for (int c = lane; c < D; c += 32) {
        const float v = x[(size_t)row * D + c] * inv;
        if (y) y[(size_t)row * D + c] = v;
        if (hi) split_store(hi, lo, (size_t)row * D + c, v);
    }
}

// top-k (largest, k <= 16) of each row of sim [Q, N]; ties -> lowest index. One 256-thread CTA per row: k rounds of a
// block-wide argmax over the entries strictly after the previous winner in (value desc, index asc) order.
__global__ void topk_kernel(const float* __restrict__ sim, int Q, int ld, int N, int k, float* __restrict__ val,
                            int32_t* __restrict__ idx)
{
    const int row = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (row >= Q) return;
    __shared__ float sv[8];
    __shared__ int si[8];
    __shared__ float wv;
    __shared__ int wi;
    const float* s = sim + (size_t)row * ld;
    float prev_v = INFINITY; int prev_i = -1;
    for (int j = 0; j < k; ++j) {
        float bv = -INFINITY; int bi = 0x7fffffff;
        for (int c = threadIdx.x; c < N; c += blockDim.x) {
            const float v = s[c];
            const bool after_prev = (v < prev_v) || (v == prev_v && c > prev_i);     // strictly after the previous winner
            if (after_prev && (v > bv || (v == bv && c < bi))) { bv = v; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float fv = sv[0]; int fi = si[0];
            for (int w = 1; w < nw; ++w) if (sv[w] > fv || (sv[w] == fv && si[w] < fi)) { fv = sv[w]; fi = si[w]; }
            wv = fv; wi = fi;
            val[(size_t)row * k + j] = fv; idx[(size_t)row * k + j] = fi;
        }
        __syncthreads();
        prev_v = wv; prev_i = wi;
    }
}

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

struct Workspace {
    bf16 *patch_hi, *patch_lo, *ln_hi, *ln_lo, *attn_hi, *attn_lo, *h_hi, *h_lo, *pool_hi, *pool_lo;
    float *patch_out, *x, *qkv, *emb_raw;
    size_t bytes;
};

Workspace carve(const ScClipConfig& c, int B, uint8_t* base) {
    const int G = c.image_size / c.patch, T = G * G + 1, W = c.width;
    const size_t Mp = (size_t)B * G * G, M = (size_t)B * T, Kc = ((size_t)3 * c.patch * c.patch + 63) / 64 * 64;
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += align_up(bytes); return p; };
    Workspace w;
    w.patch_hi = (bf16*)take(Mp * Kc * 2); w.patch_lo = (bf16*)take(Mp * Kc * 2);
    w.patch_out = (float*)take(Mp * W * 4);
    w.x = (float*)take(M * W * 4);
    w.ln_hi = (bf16*)take(M * W * 2); w.ln_lo = (bf16*)take(M * W * 2);
    w.qkv = (float*)take(M * 3 * W * 4);
    w.attn_hi = (bf16*)take(M * W * 2); w.attn_lo = (bf16*)take(M * W * 2);
    w.h_hi = (bf16*)take(M * 4 * W * 2); w.h_lo = (bf16*)take(M * 4 * W * 2);
    w.pool_hi = (bf16*)take((size_t)B * W * 2); w.pool_lo = (bf16*)take((size_t)B * W * 2);
    w.emb_raw = (float*)take((size_t)B * c.out_dim * 4);
    w.bytes = off;
    return w;
}

}  // namespace scclip

using namespace scclip;

extern "C" size_t sc_clip_workspace_bytes(const ScClipConfig* cfg, int batch) {
    if (cfg == nullptr || batch <= 0) return 0;
    return carve(*cfg, batch, nullptr).bytes;
}

#define SC_TRY(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

extern "C" int sc_clip_encode(const ScClipConfig* cfg, const ScClipWeights* wts, const float* images, int batch,
                              float* emb, float* emb_unnormalised, void* emb_hi, void* emb_lo, void* workspace,
                              size_t workspace_bytes, cudaStream_t stream)
{
    if (cfg == nullptr || wts == nullptr || images == nullptr || workspace == nullptr) return (int)cudaErrorInvalidValue;
    if (batch <= 0) return 0;
    const ScClipConfig& c = *cfg;
    const int G = c.image_size / c.patch, T = G * G + 1, W = c.width, H = c.heads;
    if (W % 64 != 0 || W > 1024 || W / H != 64 || c.out_dim % 64 != 0)
        return (int)cudaErrorInvalidValue;
    Workspace w = carve(c, batch, (uint8_t*)workspace);
    if (workspace_bytes < w.bytes) return (int)cudaErrorInvalidValue;
    const int Mp = batch * G * G, M = batch * T, Kc = (3 * c.patch * c.patch + 63) / 64 * 64;
    const bool split = c.split != 0;
    auto lo = [&](bf16* p) { return split ? p : (bf16*)nullptr; };
    auto wlo = [&](const void* p) { return split ? p : (const void*)nullptr; };

    im2col_split_kernel<<<296, 256, 0, stream>>>(images, batch, c.image_size, c.patch, Kc, w.patch_hi, lo(w.patch_lo));
    SC_TRY(sc_gemm_bf16_tc(w.patch_hi, lo(w.patch_lo), wts->conv_w_hi, wlo(wts->conv_w_lo), Mp, W, Kc, nullptr, nullptr, 0, 1.f,
                           w.patch_out, nullptr, nullptr, stream));
    const int ln_warps = 8;
    layernorm_kernel<<<(M + ln_warps - 1) / ln_warps, ln_warps * 32, 0, stream>>>(
        nullptr, 0, M, W, wts->lnpre_w, wts->lnpre_b, nullptr, nullptr, w.x, w.patch_out, wts->class_emb, wts->pos_emb, T);
    for (int l = 0; l < c.layers; ++l) {
        const ScClipLayer& L = wts->layers[l];
        layernorm_kernel<<<(M + ln_warps - 1) / ln_warps, ln_warps * 32, 0, stream>>>(
            w.x, (size_t)W, M, W, L.ln1_w, L.ln1_b, w.ln_hi, lo(w.ln_lo), nullptr, nullptr, nullptr, nullptr, T);
        SC_TRY(sc_gemm_bf16_tc(w.ln_hi, lo(w.ln_lo), L.qkv_w_hi, wlo(L.qkv_w_lo), M, 3 * W, W, L.qkv_b, nullptr, 0, 1.f,
                               w.qkv, nullptr, nullptr, stream));
        const int att_threads = 512;      // 16 warps share one (image, head)'s K/V tile: ~3 query rows per warp at T = 50
        const size_t att_smem = ((size_t)2 * T * scclip::kAttLd + (size_t)(att_threads / 32) * (2 * ((T + 3) & ~3) + 128)) * sizeof(float);
        if (att_smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem);
            if (e != cudaSuccess) return (int)e;
        }
        attention_kernel<<<dim3(batch, H), att_threads, att_smem, stream>>>(w.qkv, T, W, w.attn_hi, lo(w.attn_lo));
        SC_TRY(sc_gemm_bf16_tc(w.attn_hi, lo(w.attn_lo), L.out_w_hi, wlo(L.out_w_lo), M, W, W, L.out_b, w.x, 0, 1.f, w.x,
                               nullptr, nullptr, stream));
        layernorm_kernel<<<(M + ln_warps - 1) / ln_warps, ln_warps * 32, 0, stream>>>(
            w.x, (size_t)W, M, W, L.ln2_w, L.ln2_b, w.ln_hi, lo(w.ln_lo), nullptr, nullptr, nullptr, nullptr, T);
        SC_TRY(sc_gemm_bf16_tc(w.ln_hi, lo(w.ln_lo), L.fc1_w_hi, wlo(L.fc1_w_lo), M, 4 * W, W, L.fc1_b, nullptr, 1, 1.f, nullptr,
                               w.h_hi, lo(w.h_lo), stream));
        SC_TRY(sc_gemm_bf16_tc(w.h_hi, lo(w.h_lo), L.fc2_w_hi, wlo(L.fc2_w_lo), M, W, 4 * W, L.fc2_b, w.x, 0, 1.f, w.x, nullptr,
                               nullptr, stream));
    }
    // ln_post on the class token of every image (row stride T*W), projection, L2 normalisation (CLIP_anno.py:167)
    layernorm_kernel<<<(batch + ln_warps - 1) / ln_warps, ln_warps * 32, 0, stream>>>(
        w.x, (size_t)T * W, batch, W, wts->lnpost_w, wts->lnpost_b, w.pool_hi, lo(w.pool_lo), nullptr, nullptr, nullptr, nullptr, T);
    float* raw = emb_unnormalised ? emb_unnormalised : w.emb_raw;
    SC_TRY(sc_gemm_bf16_tc(w.pool_hi, lo(w.pool_lo), wts->proj_w_hi, wlo(wts->proj_w_lo), batch, c.out_dim, W, nullptr, nullptr, 0,
                           1.f, raw, nullptr, nullptr, stream));
    l2norm_kernel<<<(batch + 7) / 8, 256, 0, stream>>>(raw, batch, c.out_dim, emb, (bf16*)emb_hi, (bf16*)emb_lo);
    return (int)cudaGetLastError();
}

// cosine top-k of unit-norm queries against a unit-norm bank (both as hi/lo bf16 planes): sim = Q . Bank^T, then top-k
extern "C" int sc_cosine_topk(const void* q_hi, const void* q_lo, const void* bank_hi, const void* bank_lo, int n_query,
                              int n_bank, int n_bank_valid, int dim, int k, float* sim_workspace, float* values,
                              int32_t* indices, cudaStream_t stream)
{
    if (n_query <= 0) return 0;
    if (k <= 0 || n_bank_valid > n_bank || k > n_bank_valid || sim_workspace == nullptr) return (int)cudaErrorInvalidValue;
    SC_TRY(sc_gemm_bf16_tc(q_hi, q_lo, bank_hi, bank_lo, n_query, n_bank, dim, nullptr, nullptr, 0, 1.f, sim_workspace, nullptr,
                           nullptr, stream));
    topk_kernel<<<n_query, 256, 0, stream>>>(sim_workspace, n_query, n_bank, n_bank_valid, k, values, indices);
    return (int)cudaGetLastError();
}
