// clip_tower.cu — the CLIP ViT image tower as ONE persistent, warp-specialised sm_100a kernel.
//
// Replaces clip_model.encode_image + F.normalize (CLIP_anno.py:166-167; openai/CLIP is an un-vendored pip dependency, the
// architecture is restated from its published VisionTransformer, see clip_ops.cu) for the whole encoder: patch embedding,
// token assembly + ln_pre, L x {LN1+QKV, attention, out-proj+residual, LN2+fc1+QuickGELU, fc2+residual}, ln_post, projection
// and L2 normalisation run as a list of PHASES inside one cooperative launch, separated by grid barriers (3 launches' worth of
// work per layer in round 1 cost 92 launches per encode; this is 1).
//
// GEMM phases (all of them the same code):  C[M,N] = A[M,K] . W[N,K]^T
//   warp 0        TMA producer: A (128 x 64) and W (BN x 64) tiles, SWIZZLE_128B, mbarrier ring over ALL tiles of the phase
//   warp 1        tcgen05.mma issuer (UMMA 128 x BN x 16, kind::f16), accumulator DOUBLE-BUFFERED in TMEM (2 x 256 columns):
//                 the epilogue of tile i overlaps the main loop of tile i+1
//   warps 4..7    epilogue: tcgen05.ld -> bias / LayerNorm correction / QuickGELU / residual -> global
//   Tiles are walked persistently (tile = blockIdx.x + i * gridDim.x, m fastest so concurrent CTAs share a W tile in L2).
// LayerNorm is FOLDED into the GEMM that consumes it: with W'[n,k] = g[k] W[n,k], s[n] = sum_k W'[n,k], c[n] = sum_k b[k] W[n,k] + bias[n]
//   LN(x) W^T + bias = rstd (x W'^T - mean s) + c
// so the A operand is the raw residual stream in 16-bit form, which the PREVIOUS residual epilogue writes next to the fp32
// stream, together with per-row (mean, M2) partials of its BN columns; the consumer's epilogue merges the partials (Chan) into
// mean / rstd. No LayerNorm kernel, no normalised activation tensor.
// Attention runs on mma.sync m16n8k16 from ldmatrix-fed shared-memory tiles (64 queries x 64 keys per 4-warp group, online
// softmax over key tiles), three (image, head, query-tile) items per CTA at a time.
//
// Two arithmetic modes (template SPLIT): 1 = fp16 operands, one MMA per product — the reference's own precision (clip.load
// returns an fp16 model on CUDA, CLIP_anno.py:16) with an fp32 residual stream and fp32 accumulation; 3 = hi/lo bf16 operand
// pairs, three MMAs per product (fp32-class, the 1e-4 parity mode) — also inside attention.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "gemm_tc.cuh"
#include "sc_b200.h"

namespace sctw {
using namespace sctc;

enum { PH_GEMM = 0, PH_ATTN = 1, PH_IM2COL = 2, PH_TOKENS = 3, PH_HEAD = 4, PH_L2NORM = 5 };
enum { EPI_F32 = 0, EPI_LN16 = 1, EPI_RESID = 2 };

struct alignas(16) Phase {
    int type, M, N, K;
    int BN, epi, act, map_a;              // tensor-map indices of the hi planes (lo plane = index + 1)
    int map_w, parts_in, cnt_in, T;       // LN partials per row to merge (EPI_LN16) and the columns each covers; tokens per image
    int B, H, W, P;                       // batch, heads, width, patch
    int S, Kp, D, pad0;                   // image size, padded im2col K, out dim
    const float* v0;                      // EPI_LN16: s[N]      | TOKENS: class_emb | LNPOST: gain  | PROJ: proj^T [D, W]
    const float* v1;                      // EPI_LN16: c[N]; else bias[N] or null | TOKENS: pos_emb | LNPOST: bias
    const float* v2;                      // TOKENS: ln_pre gain | HEAD: proj^T [D, W]
    const float* v3;                      // TOKENS: ln_pre bias
    float* f0;                            // GEMM: fp32 out / residual stream x | TOKENS: x | LNPOST: y [B, W] | PROJ: raw emb | L2NORM: raw emb
    const float* f1;                      // TOKENS: patch_out | HEAD: x | IM2COL: images (null = kernel argument)
    void* o_hi; void* o_lo;               // 16-bit output planes
    const void* i_hi; const void* i_lo;   // ATTN: qkv planes
    const float2* st_in; float2* st_out;  // LN partial statistics [M, parts]
};

constexpr int kThreads = 384;             // warp 0 TMA, 1 MMA issue, 2 TMEM owner, 3 spare, 4..11 epilogue; 3 x 4-warp attention groups
constexpr int kTileBytes = 196 * 1024;    // operand ring / attention staging
constexpr int kMaxStages = 8;
constexpr int kSmemBytes = kTileBytes + 1024 + 512;
constexpr int kAttPitch = 72;             // 16-bit elements per staged row (64 + 8: conflict-free ldmatrix)

// ------------------------------------------------------------------------------------------------ small device helpers
template <int SPLIT> struct E16;
template <> struct E16<1> {
    typedef __half t;
    static __device__ __forceinline__ uint32_t pack(float a, float b) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
    static __device__ __forceinline__ t one(float a) { return __float2half_rn(a); }
    static __device__ __forceinline__ float lo_of(float v) { return 0.f; }
    static constexpr uint32_t kFmt = 0;   // UMMA a/b format F16
};
template <> struct E16<3> {
    typedef __nv_bfloat16 t;
    static __device__ __forceinline__ uint32_t pack(float a, float b) { __nv_bfloat162 h = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
    static __device__ __forceinline__ t one(float a) { return __float2bfloat16_rn(a); }
    static __device__ __forceinline__ float lo_of(float v) { return v - __bfloat162float(__float2bfloat16_rn(v)); }
    static constexpr uint32_t kFmt = 1;   // BF16
};

__device__ __forceinline__ uint32_t make_idesc(int N, uint32_t fmt) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void bar_named(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// 16 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld_16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// grid-wide barrier between phases: every CTA is resident (cooperative launch). `target` counts arrivals so far.
__device__ __forceinline__ void grid_sync(unsigned* counter, unsigned& target) {
    fence_proxy_async_all();                       // generic-proxy global writes of this phase vs. TMA reads of the next
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_u32(counter) < target) { }
        __threadfence();
    }
    __syncthreads();
    fence_proxy_async_all();
}

struct Smem {
    uint8_t* tiles;
    uint64_t *full, *empty, *tfull, *tempty;
    uint32_t* tmem_slot;
};

// ------------------------------------------------------------------------------------------------ GEMM phase
// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): one FULL 32-byte sector per thread and instruction. The epilogue's accesses
// are row-strided (thread = accumulator row), so every lane touches its own sector; with 128-bit accesses each sector was requested
// twice and the L2 request rate, not bytes, bounded the residual epilogues (measured: 17 of 37 us in out-proj).
// QuickGELU x sigmoid(1.702 x). Parity mode: exp + reciprocal (two SFU ops per element). fp16 mode: sigmoid(z) = 0.5 tanh(z / 2) + 0.5
// with tanh.approx (ONE SFU op, max relative error 2^-11 = the fp16 rounding of the result): the 128 x 192 fc1 tiles were SFU-paced.
template <int SPLIT>
__device__ __forceinline__ float quick_gelu(float x) {
    if (SPLIT == 3) return __fdividef(x, 1.f + __expf(-1.702f * x));
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}
struct F8 { float v[8]; };
__device__ __forceinline__ F8 ldcg256(const float* p) {
    F8 r;
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void st256u(void* p, const uint32_t* h) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"l"(p), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
}

// Epilogue of one 16-column chunk of this thread's row. Every per-phase parameter arrives in registers (the Phase lives in shared
// memory and the tcgen05 asm blocks clobber memory: reading it inside the loop costs a dependent LDS + LDG chain per element).
// `rs` = this chunk's residual values (EPI_RESID), loaded by the caller one chunk ahead.
template <int SPLIT, int EPI>
__device__ __forceinline__ void epi_chunk(uint32_t taddr, bool row_ok, size_t off, int col0, int act, int dbg, float mean, float rstd,
                                          const float* __restrict__ v0, const float* __restrict__ v1, float* __restrict__ f0,
                                          typename E16<SPLIT>::t* __restrict__ o_hi, typename E16<SPLIT>::t* __restrict__ o_lo,
                                          const F8& rs0, const F8& rs1, float& r_n, float& r_mean, float& r_m2)
{
    typedef E16<SPLIT> E;
    float4 p0[4], p1[4];
    if (EPI == EPI_LN16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { p0[i] = __ldg(reinterpret_cast<const float4*>(v0 + col0) + i); p1[i] = __ldg(reinterpret_cast<const float4*>(v1 + col0) + i); }
    } else if (v1 != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) p1[i] = __ldg(reinterpret_cast<const float4*>(v1 + col0) + i);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) p1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float v[16];
    tmem_ld_16(taddr, v);
    if (EPI == EPI_LN16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[4 * i + 0] = rstd * (v[4 * i + 0] - mean * p0[i].x) + p1[i].x;
            v[4 * i + 1] = rstd * (v[4 * i + 1] - mean * p0[i].y) + p1[i].y;
            v[4 * i + 2] = rstd * (v[4 * i + 2] - mean * p0[i].z) + p1[i].z;
            v[4 * i + 3] = rstd * (v[4 * i + 3] - mean * p0[i].w) + p1[i].w;
        }
        if (act == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = quick_gelu<SPLIT>(v[i]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[4 * i] += p1[i].x; v[4 * i + 1] += p1[i].y; v[4 * i + 2] += p1[i].z; v[4 * i + 3] += p1[i].w; }
        if (EPI == EPI_RESID) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i] += rs0.v[i]; v[8 + i] += rs1.v[i]; }
        }
    }
    if (!row_ok) return;
    if (EPI != EPI_LN16 && !(dbg & 2)) { st256(f0 + off, v); st256(f0 + off + 8, v + 8); }
    if (EPI != EPI_F32 && !(dbg & 4)) {
        uint32_t h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) h[i] = E::pack(v[2 * i], v[2 * i + 1]);
        st256u(o_hi + off, h);
        if (SPLIT == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = E::pack(E::lo_of(v[2 * i]), E::lo_of(v[2 * i + 1]));
            st256u(o_lo + off, h);
        }
    }
    if (EPI == EPI_RESID) {                                          // (mean, M2) of these 16 columns, merged into the running pair (Chan)
        float mc = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) mc += v[i];
        mc *= (1.f / 16.f);
        float qc = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float d = v[i] - mc; qc += d * d; }
        const float nn = r_n + 16.f, d = mc - r_mean;
        r_m2 += qc + d * d * r_n * 16.f / nn;
        r_mean += d * 16.f / nn;
        r_n = nn;
    }
}

template <int SPLIT>
__device__ __forceinline__ void gemm_phase(const Phase& ph_smem, const CUtensorMap* maps, const Smem& sm, uint32_t tmem_base,
                                           uint32_t& ring_par, uint32_t& acc_it)
{
    typedef E16<SPLIT> E;
    constexpr int PL = (SPLIT == 3) ? 2 : 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per-phase parameters into registers, once
    const int BN = ph_smem.BN, M = ph_smem.M, N = ph_smem.N, K = ph_smem.K, epi = ph_smem.epi, act = ph_smem.act;
    const int a_bytes = 128 * 64 * 2, w_bytes = BN * 64 * 2;
    const int stage_bytes = PL * (a_bytes + w_bytes);
    int ns = kTileBytes / stage_bytes;
    ns = ns > kMaxStages ? kMaxStages : ns;
    const int MT = (M + 127) / 128, NT = N / BN, num_kb = K / 64;
    const int tiles = MT * NT;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (lane == 0) {
            const CUtensorMap* mAh = maps + ph_smem.map_a;
            const CUtensorMap* mWh = maps + ph_smem.map_w;
            tma_prefetch_desc(mAh); tma_prefetch_desc(mWh);
            if (SPLIT == 3) { tma_prefetch_desc(mAh + 1); tma_prefetch_desc(mWh + 1); }
            int s = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int m0 = (t % MT) * 128, n0 = (t / MT) * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(sm.empty + s, ((ring_par >> s) & 1u) ^ 1u);
                    uint8_t* st = sm.tiles + (size_t)s * stage_bytes;
                    mbar_expect_tx(sm.full + s, stage_bytes);
                    tma_load_2d(st, mAh, sm.full + s, kb * 64, m0);
                    tma_load_2d(st + a_bytes, mWh, sm.full + s, kb * 64, n0);
                    if (SPLIT == 3) {
                        tma_load_2d(st + a_bytes + w_bytes, mAh + 1, sm.full + s, kb * 64, m0);
                        tma_load_2d(st + 2 * a_bytes + w_bytes, mWh + 1, sm.full + s, kb * 64, n0);
                    }
                    ring_par ^= 1u << s;
                    s = (s + 1 == ns) ? 0 : s + 1;
                }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BN, E::kFmt);
            int s = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
                const uint32_t buf = acc_it & 1u;
                mbar_wait(sm.tempty + buf, ((acc_it >> 1) & 1u) ^ 1u);       // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem_base + buf * 256u;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(sm.full + s, (ring_par >> s) & 1u);
                    tc_fence_after();
                    uint8_t* st = sm.tiles + (size_t)s * stage_bytes;
                    const uint64_t dAh = make_smem_desc_k128(st);
                    const uint64_t dWh = make_smem_desc_k128(st + a_bytes);
                    const uint64_t dAl = make_smem_desc_k128(st + a_bytes + w_bytes);
                    const uint64_t dWl = make_smem_desc_k128(st + 2 * a_bytes + w_bytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = (uint64_t)(k * 32 >> 4);
                        umma_bf16(tacc, dAh + adv, dWh + adv, idesc, (kb | k) != 0);
                        if (SPLIT == 3) {
                            umma_bf16(tacc, dAh + adv, dWl + adv, idesc, 1);
                            umma_bf16(tacc, dAl + adv, dWh + adv, idesc, 1);
                        }
                    }
                    umma_commit(sm.empty + s);
                    ring_par ^= 1u << s;
                    s = (s + 1 == ns) ? 0 : s + 1;
                }
                umma_commit(sm.tfull + buf);
                ++acc_it;
            }
        }
    } else if (warp >= 4) {
        // ---------------------------------------------------------------- epilogue: 8 warps; TMEM lane quarter = warp & 3, column half = (warp - 4) >> 2
        const int q = warp & 3, half = (warp - 4) >> 2;
        const int hw = BN >> 1, cb = half * hw;
        typename E::t* o_hi = reinterpret_cast<typename E::t*>(ph_smem.o_hi);
        typename E::t* o_lo = reinterpret_cast<typename E::t*>(ph_smem.o_lo);
        const float* v0 = ph_smem.v0; const float* v1 = ph_smem.v1; float* f0 = ph_smem.f0;
        const float2* st_in = ph_smem.st_in; float2* st_out = ph_smem.st_out;
        const int parts_in = ph_smem.parts_in, cnt_in = ph_smem.cnt_in, dbg = ph_smem.pad0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int mt = t % MT, nt = t / MT;
            const int m0 = mt * 128, n0 = nt * BN;
            const uint32_t buf = acc_it & 1u;
            const int row = m0 + q * 32 + lane;
            const bool row_ok = row < M;
            // LayerNorm statistics of this row (Chan merge of the partials the producer of x left behind)
            float mean = 0.f, rstd = 1.f;
            if (epi == EPI_LN16 && row_ok) {
                const float2* sp = st_in + (size_t)row * parts_in;
                float mu = 0.f, m2 = 0.f;
                for (int j = 0; j < parts_in; ++j) mu += __ldcg(&sp[j]).x;
                mu /= (float)parts_in;
                for (int j = 0; j < parts_in; ++j) { const float2 pj = __ldcg(&sp[j]); const float d = pj.x - mu; m2 += pj.y + (float)cnt_in * d * d; }
                mean = mu;
                rstd = rsqrtf(m2 / (float)(parts_in * cnt_in) + 1e-5f);
            }
            mbar_wait(sm.tfull + buf, (acc_it >> 1) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 256u;
            float r_n = 0.f, r_mean = 0.f, r_m2 = 0.f;                       // running statistics of the new x row (EPI_RESID)
            const size_t row_off = (size_t)row * N + n0;
            F8 z8;
#pragma unroll
            for (int i = 0; i < 8; ++i) z8.v[i] = 0.f;
            if (epi == EPI_LN16) {
#pragma unroll 2
                for (int c = cb; c < cb + hw; c += 16)
                    epi_chunk<SPLIT, EPI_LN16>(trow + (uint32_t)c, row_ok, row_off + c, n0 + c, act, dbg, mean, rstd, v0, v1, f0, o_hi, o_lo, z8, z8, r_n, r_mean, r_m2);
            } else if (epi == EPI_RESID) {
                // the residual does not depend on the MMA: its loads run one chunk ahead of the accumulator reads
                const bool ld_ok = row_ok && !(dbg & 1);
                F8 a0 = z8, a1 = z8;
                if (ld_ok) { a0 = ldcg256(f0 + row_off + cb); a1 = ldcg256(f0 + row_off + cb + 8); }
#pragma unroll 1
                for (int c = cb; c < cb + hw; c += 16) {
                    F8 b0 = z8, b1 = z8;
                    if (ld_ok && c + 16 < cb + hw) { b0 = ldcg256(f0 + row_off + c + 16); b1 = ldcg256(f0 + row_off + c + 24); }
                    epi_chunk<SPLIT, EPI_RESID>(trow + (uint32_t)c, row_ok, row_off + c, n0 + c, act, dbg, mean, rstd, v0, v1, f0, o_hi, o_lo, a0, a1, r_n, r_mean, r_m2);
                    a0 = b0; a1 = b1;
                }
            } else {
#pragma unroll 2
                for (int c = cb; c < cb + hw; c += 16)
                    epi_chunk<SPLIT, EPI_F32>(trow + (uint32_t)c, row_ok, row_off + c, n0 + c, act, dbg, mean, rstd, v0, v1, f0, o_hi, o_lo, z8, z8, r_n, r_mean, r_m2);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.tempty + buf);                     // accumulator buffer free for tile i+2 (8 arrivals)
            if (epi == EPI_RESID && row_ok && st_out != nullptr)
                st_out[(size_t)row * (2 * NT) + 2 * nt + half] = make_float2(r_mean, r_m2);
            ++acc_it;
        }
    }
}

// ------------------------------------------------------------------------------------------------ attention phase
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <int SPLIT>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (SPLIT == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// softmax(q k^T) v for one (image, head, 64-query tile) by a 4-warp group; q arrives pre-scaled by 1/sqrt(64) (folded into W_q).
// qkv planes [M, 3W] (q | k | v; head h = 64 columns); output planes [M, W].
template <int SPLIT>
__device__ __forceinline__ void attention_phase(const Phase& ph, uint8_t* tiles)
{
    typedef E16<SPLIT> E;
    typedef typename E::t T16;
    constexpr int PL = (SPLIT == 3) ? 2 : 1;
    const int grp = threadIdx.x >> 7, tg = threadIdx.x & 127, warp = tg >> 5, lane = threadIdx.x & 31;
    const int T = ph.T, W = ph.W, H = ph.H;
    const int QT = (T + 63) / 64;
    const int items = ph.B * H * QT;
    const int plane_elems = 64 * kAttPitch;
    T16* base = reinterpret_cast<T16*>(tiles) + (size_t)grp * (3 * PL * plane_elems);
    T16* Qs = base;                               // [PL][64][72]
    T16* Ks = base + PL * plane_elems;
    T16* Vs = base + 2 * PL * plane_elems;
    const T16* in_pl[2] = {reinterpret_cast<const T16*>(ph.i_hi), reinterpret_cast<const T16*>(ph.i_lo)};
    T16* out_pl[2] = {reinterpret_cast<T16*>(ph.o_hi), reinterpret_cast<T16*>(ph.o_lo)};
    const int g = lane >> 2, tq = lane & 3;
    const float kLog2e = 1.4426950408889634f;

    for (int it = 3 * blockIdx.x + grp; it < items; it += 3 * gridDim.x) {
        const int qt = it % QT, h = (it / QT) % H, b = it / (QT * H);
        const size_t row_base = (size_t)b * T;
        const int q0 = qt * 64;
        // ---- stage the query tile (rows beyond T zero-filled)
        for (int i = tg; i < PL * 64 * 8; i += 128) {
            const int pl = i / 512, r = (i >> 3) & 63, c8 = (i & 7) * 8;
            uint4 val = make_uint4(0, 0, 0, 0);
            if (q0 + r < T) val = __ldcg(reinterpret_cast<const uint4*>(in_pl[pl] + (row_base + q0 + r) * (size_t)(3 * W) + h * 64 + c8));
            *reinterpret_cast<uint4*>(Qs + pl * plane_elems + r * kAttPitch + c8) = val;
        }
        float o[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
        float mx0 = -1e30f, mx1 = -1e30f, l0 = 0.f, l1 = 0.f;

        for (int k0 = 0; k0 < T; k0 += 64) {
            bar_named(1 + grp, 128);                                      // previous K/V tile fully consumed (and Q staged)
            for (int i = tg; i < PL * 64 * 8; i += 128) {
                const int pl = i / 512, r = (i >> 3) & 63, c8 = (i & 7) * 8;
                uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
                if (k0 + r < T) {
                    const T16* src = in_pl[pl] + (row_base + k0 + r) * (size_t)(3 * W) + h * 64 + c8;
                    kv = __ldcg(reinterpret_cast<const uint4*>(src + W));
                    vv = __ldcg(reinterpret_cast<const uint4*>(src + 2 * W));
                }
                *reinterpret_cast<uint4*>(Ks + pl * plane_elems + r * kAttPitch + c8) = kv;
                *reinterpret_cast<uint4*>(Vs + pl * plane_elems + r * kAttPitch + c8) = vv;
            }
            bar_named(1 + grp, 128);
            // ---- S = Q K^T for this warp's 16 query rows x 64 keys
            float s[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {                               // 16 head dims per step
                uint32_t a[PL][4];
#pragma unroll
                for (int pl = 0; pl < PL; ++pl)
                    ldsm_x4(a[pl], smem_u32(Qs + pl * plane_elems + (warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kAttPitch + kk * 16 + 8 * (lane >> 4)));
#pragma unroll
                for (int jp = 0; jp < 4; ++jp) {                           // two 8-key tiles per ldmatrix.x4
                    uint32_t bk[PL][4];
#pragma unroll
                    for (int pl = 0; pl < PL; ++pl)
                        ldsm_x4(bk[pl], smem_u32(Ks + pl * plane_elems + (jp * 16 + (lane & 7) + 8 * (lane >> 4)) * kAttPitch + kk * 16 + 8 * ((lane >> 3) & 1)));
                    mma16816<SPLIT>(s[2 * jp], a[0], bk[0][0], bk[0][1]);
                    mma16816<SPLIT>(s[2 * jp + 1], a[0], bk[0][2], bk[0][3]);
                    if (SPLIT == 3) {
                        mma16816<SPLIT>(s[2 * jp], a[0], bk[PL - 1][0], bk[PL - 1][1]);
                        mma16816<SPLIT>(s[2 * jp + 1], a[0], bk[PL - 1][2], bk[PL - 1][3]);
                        mma16816<SPLIT>(s[2 * jp], a[PL - 1], bk[0][0], bk[0][1]);
                        mma16816<SPLIT>(s[2 * jp + 1], a[PL - 1], bk[0][2], bk[0][3]);
                    }
                }
            }
            // ---- mask, online softmax (rows g and g + 8 of this warp's 16)
            float tm0 = -1e30f, tm1 = -1e30f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int key = k0 + j * 8 + 2 * tq;
                if (key >= T) { s[j][0] = -1e30f; s[j][2] = -1e30f; }
                if (key + 1 >= T) { s[j][1] = -1e30f; s[j][3] = -1e30f; }
                tm0 = fmaxf(tm0, fmaxf(s[j][0], s[j][1]));
                tm1 = fmaxf(tm1, fmaxf(s[j][2], s[j][3]));
            }
            tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 1)); tm0 = fmaxf(tm0, __shfl_xor_sync(0xffffffffu, tm0, 2));
            tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 1)); tm1 = fmaxf(tm1, __shfl_xor_sync(0xffffffffu, tm1, 2));
            const float nm0 = fmaxf(mx0, tm0), nm1 = fmaxf(mx1, tm1);
            const float sc0 = exp2f((mx0 - nm0) * kLog2e), sc1 = exp2f((mx1 - nm1) * kLog2e);
            mx0 = nm0; mx1 = nm1;
            float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j][0] = exp2f((s[j][0] - nm0) * kLog2e); s[j][1] = exp2f((s[j][1] - nm0) * kLog2e);
                s[j][2] = exp2f((s[j][2] - nm1) * kLog2e); s[j][3] = exp2f((s[j][3] - nm1) * kLog2e);
                ps0 += s[j][0] + s[j][1]; ps1 += s[j][2] + s[j][3];
                o[j][0] *= sc0; o[j][1] *= sc0; o[j][2] *= sc1; o[j][3] *= sc1;
            }
            l0 = l0 * sc0 + ps0; l1 = l1 * sc1 + ps1;
            // ---- O += P V
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {                               // 16 keys per step
                uint32_t pa[PL][4];
                pa[0][0] = E::pack(s[2 * kk][0], s[2 * kk][1]);     pa[0][1] = E::pack(s[2 * kk][2], s[2 * kk][3]);
                pa[0][2] = E::pack(s[2 * kk + 1][0], s[2 * kk + 1][1]); pa[0][3] = E::pack(s[2 * kk + 1][2], s[2 * kk + 1][3]);
                if (SPLIT == 3) {
                    pa[PL - 1][0] = E::pack(E::lo_of(s[2 * kk][0]), E::lo_of(s[2 * kk][1]));
                    pa[PL - 1][1] = E::pack(E::lo_of(s[2 * kk][2]), E::lo_of(s[2 * kk][3]));
                    pa[PL - 1][2] = E::pack(E::lo_of(s[2 * kk + 1][0]), E::lo_of(s[2 * kk + 1][1]));
                    pa[PL - 1][3] = E::pack(E::lo_of(s[2 * kk + 1][2]), E::lo_of(s[2 * kk + 1][3]));
                }
#pragma unroll
                for (int jp = 0; jp < 4; ++jp) {                           // two 8-dim output tiles per ldmatrix.x4.trans
                    uint32_t bv[PL][4];
#pragma unroll
                    for (int pl = 0; pl < PL; ++pl)
                        ldsm_x4_t(bv[pl], smem_u32(Vs + pl * plane_elems + (kk * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * kAttPitch + jp * 16 + 8 * (lane >> 4)));
                    mma16816<SPLIT>(o[2 * jp], pa[0], bv[0][0], bv[0][1]);
                    mma16816<SPLIT>(o[2 * jp + 1], pa[0], bv[0][2], bv[0][3]);
                    if (SPLIT == 3) {
                        mma16816<SPLIT>(o[2 * jp], pa[0], bv[PL - 1][0], bv[PL - 1][1]);
                        mma16816<SPLIT>(o[2 * jp + 1], pa[0], bv[PL - 1][2], bv[PL - 1][3]);
                        mma16816<SPLIT>(o[2 * jp], pa[PL - 1], bv[0][0], bv[0][1]);
                        mma16816<SPLIT>(o[2 * jp + 1], pa[PL - 1], bv[0][2], bv[0][3]);
                    }
                }
            }
        }
        // ---- normalise and store: rows q0 + warp*16 + g (+8), columns h*64 + 8 j + 2 tq (+1)
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.f / l0, i1 = 1.f / l1;
        const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = h * 64 + j * 8 + 2 * tq;
            const float a0 = o[j][0] * i0, a1 = o[j][1] * i0, b0 = o[j][2] * i1, b1 = o[j][3] * i1;
            if (r0 < T) {
                *reinterpret_cast<uint32_t*>(out_pl[0] + (row_base + r0) * (size_t)W + col) = E::pack(a0, a1);
                if (SPLIT == 3) *reinterpret_cast<uint32_t*>(out_pl[1] + (row_base + r0) * (size_t)W + col) = E::pack(E::lo_of(a0), E::lo_of(a1));
            }
            if (r1 < T) {
                *reinterpret_cast<uint32_t*>(out_pl[0] + (row_base + r1) * (size_t)W + col) = E::pack(b0, b1);
                if (SPLIT == 3) *reinterpret_cast<uint32_t*>(out_pl[1] + (row_base + r1) * (size_t)W + col) = E::pack(E::lo_of(b0), E::lo_of(b1));
            }
        }
        bar_named(1 + grp, 128);                                          // Q tile free for the next item
    }
}

// ------------------------------------------------------------------------------------------------ element-wise phases
// images [B,3,S,S] fp32 -> patches [B*G*G, Kp] 16-bit planes; column = c*P*P + py*P + px (conv1.weight flattening). The padding
// columns [3*P*P, Kp) are never written: the workspace is zero-filled when the plan is made.
// Unit of work = one pixel row of one patch channel (P contiguous floats in, P contiguous 16-bit values out), handled by 8 lanes
// with vector loads; four units per group in flight (one CTA per SM: instruction-level parallelism is the only latency hiding),
// 32-bit index arithmetic only (the first version spent its time in 64-bit divisions).
template <int SPLIT, int VEC>
__device__ __forceinline__ void im2col_rows(const Phase& ph, const float* images)
{
    typedef E16<SPLIT> E;
    typename E::t* hi = reinterpret_cast<typename E::t*>(ph.o_hi);
    typename E::t* lo = reinterpret_cast<typename E::t*>(ph.o_lo);
    const int S = ph.S, P = ph.P, G = S / P, Kp = ph.Kp, PP = P * P;
    const int units = ph.B * G * G * 3 * P;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, l8 = gtid & 7;
    const int ngroups = (gridDim.x * blockDim.x) >> 3;
    const int px = l8 * VEC;
    constexpr int U = 4;
    for (int u0 = gtid >> 3; u0 < units; u0 += ngroups * U) {
        float v[U][VEC];
        int dst[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int u = u0 + k * ngroups;
            dst[k] = -1;
            if (u < units && px < P) {
                const int py = u % P, t = u / P, c = t % 3, row = t / 3;
                const int gx = row % G, t2 = row / G, gy = t2 % G, bb = t2 / G;
                const float* src = images + (((size_t)bb * 3 + c) * S + gy * P + py) * S + gx * P + px;
                if (VEC == 4) { const float4 q = __ldg(reinterpret_cast<const float4*>(src)); v[k][0] = q.x; v[k][1] = q.y; v[k][VEC - 2] = q.z; v[k][VEC - 1] = q.w; }
                else { const float2 q = __ldg(reinterpret_cast<const float2*>(src)); v[k][0] = q.x; v[k][1] = q.y; }
                dst[k] = row * Kp + c * PP + py * P + px;            // < 2^31 for every supported shape (checked by the host)
            }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            if (dst[k] < 0) continue;
            if (VEC == 4) {
                *reinterpret_cast<uint2*>(hi + dst[k]) = make_uint2(E::pack(v[k][0], v[k][1]), E::pack(v[k][VEC - 2], v[k][VEC - 1]));
                if (SPLIT == 3) *reinterpret_cast<uint2*>(lo + dst[k]) = make_uint2(E::pack(E::lo_of(v[k][0]), E::lo_of(v[k][1])), E::pack(E::lo_of(v[k][VEC - 2]), E::lo_of(v[k][VEC - 1])));
            } else {
                *reinterpret_cast<uint32_t*>(hi + dst[k]) = E::pack(v[k][0], v[k][1]);
                if (SPLIT == 3) *reinterpret_cast<uint32_t*>(lo + dst[k]) = E::pack(E::lo_of(v[k][0]), E::lo_of(v[k][1]));
            }
        }
    }
}
template <int SPLIT>
__device__ __forceinline__ void im2col_phase(const Phase& ph, const float* images)
{
    if (ph.P == 32) im2col_rows<SPLIT, 4>(ph, images);      // 8 lanes x 4 pixels
    else im2col_rows<SPLIT, 2>(ph, images);                 // P <= 16, even (ViT-L/14: 7 of 8 lanes active)
}

// one warp per token row: x = LN_pre((class | patch) + pos) -> fp32 stream, 16-bit planes, (mean, M2) of the row
template <int SPLIT, int PER>
__device__ __forceinline__ void tokens_rows(const Phase& ph)
{
    typedef E16<SPLIT> E;
    typename E::t* hi = reinterpret_cast<typename E::t*>(ph.o_hi);
    typename E::t* lo = reinterpret_cast<typename E::t*>(ph.o_lo);
    const int W = ph.W, T = ph.T, M = ph.M, lane = threadIdx.x & 31;
    const float* cls = ph.v0; const float* pos = ph.v1; const float* gain = ph.v2; const float* bias = ph.v3; const float* patch = ph.f1;
    float* x = ph.f0; float2* st = ph.st_out;
    const int wpb = blockDim.x >> 5;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < M; row += gridDim.x * wpb) {
        const int b = row / T, tk = row % T;
        float v[PER];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int c = i * 32 + lane;
            v[i] = (tk == 0 ? __ldg(cls + c) : __ldcg(patch + ((size_t)b * (T - 1) + tk - 1) * W + c)) + __ldg(pos + (size_t)tk * W + c);
            s += v[i];
        }
        const float mean = warp_sum(s) / W;
        float qv = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) { const float d = v[i] - mean; qv += d * d; }
        const float rstd = rsqrtf(warp_sum(qv) / W + 1e-5f);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int c = i * 32 + lane;
            v[i] = (v[i] - mean) * rstd * __ldg(gain + c) + __ldg(bias + c);
            const size_t o = (size_t)row * W + c;
            x[o] = v[i];
            hi[o] = E::one(v[i]);
            if (SPLIT == 3) lo[o] = E::one(E::lo_of(v[i]));
        }
        // statistics of the NORMALISED row as ONE (mean, M2) part over all W columns (the first QKV GEMM merges parts_in = 1)
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) ss += v[i];
        const float mp = warp_sum(ss) / W;
        float qq = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) { const float d = v[i] - mp; qq += d * d; }
        qq = warp_sum(qq);
        if (lane == 0) st[row] = make_float2(mp, qq);
    }
}
template <int SPLIT>
__device__ __forceinline__ void tokens_phase(const Phase& ph)
{
    switch (ph.W / 32) {
        case 24: tokens_rows<SPLIT, 24>(ph); break;      // ViT-B
        case 32: tokens_rows<SPLIT, 32>(ph); break;      // ViT-L
        case 16: tokens_rows<SPLIT, 16>(ph); break;
        case 8:  tokens_rows<SPLIT, 8>(ph); break;
        default: tokens_rows<SPLIT, 4>(ph); break;       // W = 128 (test configuration); the host rejects other widths
    }
}

// head, part 1: raw[b, d0 .. d0+127] = LN_post(x[b, class token]) @ proj for one (image, 128-output block) per CTA and pass:
// LayerNorm by warp 0 into shared memory (recomputed per block: 768 values), the dot products spread over the CTA's warps two
// at a time (fp32 FMA, proj^T rows from L2).
constexpr int kHeadBlock = 128;
__device__ __forceinline__ void head_phase(const Phase& ph, float* smem_f)
{
    const int W = ph.W, D = ph.D, T = ph.T, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float* gain = ph.v0; const float* bias = ph.v1; const float* projT = ph.v2; const float* xs = ph.f1;
    float* raw = ph.f0;
    float* y = smem_f;                 // [W]
    const int nblk = (D + kHeadBlock - 1) / kHeadBlock, items = ph.B * nblk;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int b = it / nblk, d_lo = (it % nblk) * kHeadBlock, d_hi = min(D, d_lo + kHeadBlock);
        __syncthreads();
        if (warp == 0) {
            const float* x = xs + (size_t)b * T * W;
            float s = 0.f;
            for (int c = lane; c < W; c += 32) { const float t = __ldcg(x + c); y[c] = t; s += t; }
            const float mean = warp_sum(s) / W;
            float qv = 0.f;
            for (int c = lane; c < W; c += 32) { const float d = y[c] - mean; qv += d * d; }
            const float rstd = rsqrtf(warp_sum(qv) / W + 1e-5f);
            for (int c = lane; c < W; c += 32) y[c] = (y[c] - mean) * rstd * __ldg(gain + c) + __ldg(bias + c);
        }
        __syncthreads();
        for (int d0 = d_lo + warp * 2; d0 < d_hi; d0 += nw * 2) {           // D is even: (d0, d0 + 1) stay inside the block
            const float* w0 = projT + (size_t)d0 * W;
            const float* w1 = w0 + W;
            float a0 = 0.f, a1 = 0.f;
#pragma unroll 2
            for (int k = lane * 4; k < W; k += 128) {
                const float4 yy = *reinterpret_cast<const float4*>(y + k);
                const float4 c0 = __ldg(reinterpret_cast<const float4*>(w0 + k));
                const float4 c1 = __ldg(reinterpret_cast<const float4*>(w1 + k));
                a0 = fmaf(yy.x, c0.x, a0); a0 = fmaf(yy.y, c0.y, a0); a0 = fmaf(yy.z, c0.z, a0); a0 = fmaf(yy.w, c0.w, a0);
                a1 = fmaf(yy.x, c1.x, a1); a1 = fmaf(yy.y, c1.y, a1); a1 = fmaf(yy.z, c1.z, a1); a1 = fmaf(yy.w, c1.w, a1);
            }
            a0 = warp_sum(a0); a1 = warp_sum(a1);
            if (lane == 0) { raw[(size_t)b * D + d0] = a0; raw[(size_t)b * D + d0 + 1] = a1; }
        }
    }
}

// head, part 2: emb = raw / max(|raw|, 1e-12) (+ optional copy of raw, + hi/lo bf16 planes of emb for sc_cosine_topk); one warp per image
__device__ __forceinline__ void l2norm_phase(const Phase& ph, float* emb, float* raw_out, __nv_bfloat16* e_hi, __nv_bfloat16* e_lo)
{
    const int D = ph.D, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int b = blockIdx.x * wpb + (threadIdx.x >> 5); b < ph.B; b += gridDim.x * wpb) {
        const float* x = ph.f0 + (size_t)b * D;
        float s = 0.f;
        for (int c = lane; c < D; c += 32) { const float v = __ldcg(x + c); s += v * v; }
        const float inv = 1.f / fmaxf(sqrtf(warp_sum(s)), 1e-12f);
        for (int c = lane; c < D; c += 32) {
            const float r = __ldcg(x + c), v = r * inv;
            const size_t o = (size_t)b * D + c;
            if (emb) emb[o] = v;
            if (raw_out) raw_out[o] = r;
            if (e_hi) { const __nv_bfloat16 hh = __float2bfloat16_rn(v); e_hi[o] = hh; if (e_lo) e_lo[o] = __float2bfloat16_rn(v - __bfloat162float(hh)); }
        }
    }
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int SPLIT>
__global__ void __launch_bounds__(kThreads, 1)
clip_tower_kernel(const Phase* __restrict__ phases, const CUtensorMap* __restrict__ maps, int ph_begin, int ph_end,
                  unsigned* barrier_counter, const float* images, float* emb, float* raw_out, __nv_bfloat16* e_hi, __nv_bfloat16* e_lo)
{
    extern __shared__ uint8_t smem_raw[];
    Smem sm;
    sm.tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    sm.full = reinterpret_cast<uint64_t*>(sm.tiles + kTileBytes);
    sm.empty = sm.full + kMaxStages;
    sm.tfull = sm.empty + kMaxStages;
    sm.tempty = sm.tfull + 2;
    sm.tmem_slot = reinterpret_cast<uint32_t*>(sm.tempty + 2);
    __shared__ Phase ph_s;

    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kMaxStages; ++s) { mbar_init(sm.full + s, 1); mbar_init(sm.empty + s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(sm.tfull + b, 1); mbar_init(sm.tempty + b, 8); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(sm.tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *sm.tmem_slot;

    uint32_t ring_par = 0, acc_it = 0;      // per-role pipeline state; identical sequences in producer / issuer / epilogue
    unsigned target = 0;
    for (int p = ph_begin; p < ph_end; ++p) {
        if (threadIdx.x < (int)(sizeof(Phase) / 4)) reinterpret_cast<uint32_t*>(&ph_s)[threadIdx.x] = reinterpret_cast<const uint32_t*>(phases + p)[threadIdx.x];
        __syncthreads();
        const Phase& ph = ph_s;
        switch (ph.type) {
            case PH_GEMM:   gemm_phase<SPLIT>(ph, maps, sm, tmem_base, ring_par, acc_it); break;
            case PH_ATTN:   attention_phase<SPLIT>(ph, sm.tiles); break;
            case PH_IM2COL: im2col_phase<SPLIT>(ph, ph.f1 != nullptr ? ph.f1 : images); break;
            case PH_TOKENS: tokens_phase<SPLIT>(ph); break;
            case PH_HEAD:   head_phase(ph, reinterpret_cast<float*>(sm.tiles)); break;
            default:        l2norm_phase(ph, emb, raw_out, e_hi, e_lo); break;
        }
        if (p + 1 < ph_end) grid_sync(barrier_counter, target);
        else __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ host: plan
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// row-major 16-bit [rows, K] -> box {64, box_rows}, 128-B swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, const void* base, int rows, int K, int box_rows, bool fp16) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return (int)cudaErrorNotSupported;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

static inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct Workspace {
    void *patch[2], *x16[2], *qkv[2], *attn[2], *h[2];
    float *patch_out, *x, *y, *raw;
    float2 *st_a, *st_b;
    unsigned* counter;
    size_t bytes;
};
static int stat_parts_max(int W) { return W / 32; }
static Workspace carve(const ScClipConfig& c, int B, uint8_t* base) {
    const int G = c.image_size / c.patch, T = G * G + 1, W = c.width;
    const size_t Mp = (size_t)B * G * G, M = (size_t)B * T, Kp = ((size_t)3 * c.patch * c.patch + 63) / 64 * 64;
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += align_up(bytes); return (void*)p; };
    Workspace w;
    w.counter = (unsigned*)take(256);
    for (int p = 0; p < 2; ++p) {
        w.patch[p] = take(Mp * Kp * 2); w.x16[p] = take(M * W * 2); w.qkv[p] = take(M * 3 * W * 2);
        w.attn[p] = take(M * W * 2); w.h[p] = take(M * 4 * W * 2);
    }
    w.patch_out = (float*)take(Mp * W * 4);
    w.x = (float*)take(M * W * 4);
    w.y = (float*)take((size_t)B * W * 4);
    w.raw = (float*)take((size_t)B * c.out_dim * 4);
    w.st_a = (float2*)take(M * stat_parts_max(W) * 8);
    w.st_b = (float2*)take(M * stat_parts_max(W) * 8);
    w.bytes = off;
    return w;
}

// tile width for a GEMM of M x N on `ctas` persistent CTAs: the multiple of 32 dividing N (<= 256, >= 64) with the best
// wave efficiency tiles / (ceil(tiles / ctas) * ctas), discounted when the MMA becomes shared-memory-read-bound (small N)
static int choose_bn(int M, int N, int ctas, int max_bn, int step = 32) {
    const int MT = (M + 127) / 128;
    int best = 0;
    double best_e = -1.0;
    for (int bn = max_bn; bn >= 64; bn -= step) {
        if (N % bn) continue;
        const int tiles = MT * (N / bn);
        const int waves = (tiles + ctas - 1) / ctas;
        double e = (double)tiles / ((double)waves * ctas);
        const double smem_clk = (4096.0 + 32.0 * bn) / 128.0, mma_clk = bn / 2.0;      // operand bytes at 128 B/clk vs MMA cycles
        if (smem_clk > mma_clk) e *= mma_clk / smem_clk;
        if (e > best_e + 0.02) { best_e = e; best = bn; }
    }
    return best;
}

}  // namespace sctw

using namespace sctw;

extern "C" size_t sc_clip_tower_workspace_bytes(const ScClipConfig* cfg, int batch) {
    if (cfg == nullptr || batch <= 0) return 0;
    return carve(*cfg, batch, nullptr).bytes;
}
extern "C" size_t sc_clip_tower_plan_bytes(const ScClipConfig* cfg) {
    if (cfg == nullptr) return 0;
    const size_t n_phase = 3 + 5 * (size_t)cfg->layers + 2, n_maps = 8 + 2 * (4 * (size_t)cfg->layers + 1);
    return align_up(n_phase * sizeof(Phase), 128) + n_maps * sizeof(CUtensorMap) + 128;
}

// Diagnostics: byte offsets of the activation buffers inside the workspace, in the order
// {x fp32, x16 hi, x16 lo, qkv hi, qkv lo, attn hi, attn lo, h hi, h lo, patch hi, patch lo, patch_out fp32, y fp32, raw fp32, stats a, stats b}.
extern "C" int sc_clip_tower_workspace_layout(const ScClipConfig* cfg, int batch, size_t* offsets16) {
    if (cfg == nullptr || batch <= 0 || offsets16 == nullptr) return (int)cudaErrorInvalidValue;
    uint8_t* base = reinterpret_cast<uint8_t*>(4096);
    Workspace w = carve(*cfg, batch, base);
    const void* p[16] = {w.x, w.x16[0], w.x16[1], w.qkv[0], w.qkv[1], w.attn[0], w.attn[1], w.h[0], w.h[1], w.patch[0], w.patch[1],
                         w.patch_out, w.y, w.raw, w.st_a, w.st_b};
    for (int i = 0; i < 16; ++i) offsets16[i] = (size_t)(reinterpret_cast<const uint8_t*>(p[i]) - base);
    return 0;
}

// Builds the phase table and every TMA descriptor ONCE for (weights, batch, workspace) and uploads them to `plan` (device).
// Not capturable (host -> device copy of host-built descriptors): call it outside CUDA-graph capture, then launch
// sc_clip_tower_encode any number of times.
extern "C" int sc_clip_tower_plan(const ScClipConfig* cfg, const ScClipTowerWeights* wts, int batch, void* workspace,
                                  size_t workspace_bytes, void* plan, size_t plan_bytes, int* n_phases_out, cudaStream_t stream)
{
    if (cfg == nullptr || wts == nullptr || workspace == nullptr || plan == nullptr || batch <= 0) return (int)cudaErrorInvalidValue;
    const ScClipConfig& c = *cfg;
    const int G = c.image_size / c.patch, T = G * G + 1, W = c.width, H = c.heads, D = c.out_dim;
    const int per = W / 32;
    if (W % 128 != 0 || !(per == 4 || per == 8 || per == 16 || per == 24 || per == 32) || W / H != 64 || D % 2 != 0 ||
        (3 * c.patch * c.patch) % 2 != 0 || c.patch % 2 != 0) return (int)cudaErrorInvalidValue;
    if ((double)batch * G * G * ((3.0 * c.patch * c.patch + 63) / 64 * 64) > 2.0e9) return (int)cudaErrorInvalidValue;      // im2col indexes with 32 bits
    Workspace w = carve(c, batch, (uint8_t*)workspace);
    if (workspace_bytes < w.bytes || plan_bytes < sc_clip_tower_plan_bytes(cfg)) return (int)cudaErrorInvalidValue;
    const int Mp = batch * G * G, M = batch * T, Kp = (3 * c.patch * c.patch + 63) / 64 * 64;
    const bool split = c.split != 0, fp16 = !split;
    const int PL = split ? 2 : 1;
    int dev = 0, ctas = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&ctas, cudaDevAttrMultiProcessorCount, dev);
    const int max_bn = 256;
    // residual GEMMs (N = W) share ONE tile width: their (mean, M2) partials must line up with what the LN-folded consumers merge
    const int bn_res = choose_bn(M, W, ctas, max_bn, 64);        // halves of 32-column multiples: the token phase's statistics segments
    const int bn_qkv = choose_bn(M, 3 * W, ctas, max_bn), bn_fc1 = choose_bn(M, 4 * W, ctas, max_bn);
    const int bn_patch = choose_bn(Mp, W, ctas, max_bn);
    if (!bn_res || !bn_qkv || !bn_fc1 || !bn_patch) return (int)cudaErrorInvalidValue;

    std::vector<Phase> ph;
    std::vector<CUtensorMap> maps;
    auto add_map = [&](const void* hi, const void* lo, int rows, int K, int box_rows) -> int {
        const int idx = (int)maps.size();
        CUtensorMap m;
        if (make_map(&m, hi, rows, K, box_rows, fp16)) return -1;
        maps.push_back(m);
        if (make_map(&m, split ? lo : hi, rows, K, box_rows, fp16)) return -1;
        maps.push_back(m);
        return idx;
    };
    const char* dbg_env = getenv("SC_TOWER_DEBUG");
    const int dbg_flags = dbg_env ? atoi(dbg_env) : 0;        // diagnostics: 1 = skip residual read, 2 = skip fp32 store, 4 = skip 16-bit store (WRONG RESULTS)
    auto zero_phase = [&]() { Phase p; memset(&p, 0, sizeof(p)); p.pad0 = dbg_flags; p.B = batch; p.H = H; p.W = W; p.P = c.patch; p.S = c.image_size; p.Kp = Kp; p.D = D; p.T = T; return p; };
    const int m_patch = add_map(w.patch[0], w.patch[1], Mp, Kp, 128);
    const int m_x = add_map(w.x16[0], w.x16[1], M, W, 128);
    const int m_attn = add_map(w.attn[0], w.attn[1], M, W, 128);
    const int m_h = add_map(w.h[0], w.h[1], M, 4 * W, 128);
    const int m_conv = add_map(wts->conv_w_hi, wts->conv_w_lo, W, Kp, bn_patch);
    if (m_patch < 0 || m_x < 0 || m_attn < 0 || m_h < 0 || m_conv < 0) return (int)cudaErrorInvalidValue;
    (void)PL;
    {   // patch embedding
        Phase p = zero_phase(); p.type = PH_IM2COL; p.o_hi = w.patch[0]; p.o_lo = w.patch[1]; ph.push_back(p);
        p = zero_phase(); p.type = PH_GEMM; p.M = Mp; p.N = W; p.K = Kp; p.BN = bn_patch; p.epi = EPI_F32; p.map_a = m_patch; p.map_w = m_conv;
        p.f0 = w.patch_out; ph.push_back(p);
        p = zero_phase(); p.type = PH_TOKENS; p.M = M; p.v0 = wts->class_emb; p.v1 = wts->pos_emb; p.v2 = wts->lnpre_w; p.v3 = wts->lnpre_b;
        p.f0 = w.x; p.f1 = w.patch_out; p.o_hi = w.x16[0]; p.o_lo = w.x16[1]; p.st_out = w.st_a; ph.push_back(p);
    }
    const int parts = 2 * (W / bn_res), cnt = bn_res / 2;       // every residual tile leaves TWO (mean, M2) partials per row (its two epilogue column halves)
    for (int l = 0; l < c.layers; ++l) {
        const ScClipTowerLayer& L = wts->layers[l];
        const int m_qkv = add_map(L.qkv_w_hi, L.qkv_w_lo, 3 * W, W, bn_qkv);
        const int m_out = add_map(L.out_w_hi, L.out_w_lo, W, W, bn_res);
        const int m_fc1 = add_map(L.fc1_w_hi, L.fc1_w_lo, 4 * W, W, bn_fc1);
        const int m_fc2 = add_map(L.fc2_w_hi, L.fc2_w_lo, W, 4 * W, bn_res);
        if (m_qkv < 0 || m_out < 0 || m_fc1 < 0 || m_fc2 < 0) return (int)cudaErrorInvalidValue;
        Phase p = zero_phase();            // qkv = LN1(x) Wqkv^T + b   (q columns pre-scaled by 1/8)
        p.type = PH_GEMM; p.M = M; p.N = 3 * W; p.K = W; p.BN = bn_qkv; p.epi = EPI_LN16; p.map_a = m_x; p.map_w = m_qkv;
        p.v0 = L.qkv_s; p.v1 = L.qkv_c; p.o_hi = w.qkv[0]; p.o_lo = w.qkv[1]; p.st_in = w.st_a; p.parts_in = (l == 0) ? 1 : parts; p.cnt_in = (l == 0) ? W : cnt; ph.push_back(p);
        p = zero_phase();                  // attention
        p.type = PH_ATTN; p.i_hi = w.qkv[0]; p.i_lo = w.qkv[1]; p.o_hi = w.attn[0]; p.o_lo = w.attn[1]; ph.push_back(p);
        p = zero_phase();                  // x += attn Wo^T + bo ; statistics for LN2
        p.type = PH_GEMM; p.M = M; p.N = W; p.K = W; p.BN = bn_res; p.epi = EPI_RESID; p.map_a = m_attn; p.map_w = m_out;
        p.v1 = L.out_b; p.f0 = w.x; p.o_hi = w.x16[0]; p.o_lo = w.x16[1]; p.st_out = w.st_b; ph.push_back(p);
        p = zero_phase();                  // h = QuickGELU(LN2(x) W1^T + b1)
        p.type = PH_GEMM; p.M = M; p.N = 4 * W; p.K = W; p.BN = bn_fc1; p.epi = EPI_LN16; p.act = 1; p.map_a = m_x; p.map_w = m_fc1;
        p.v0 = L.fc1_s; p.v1 = L.fc1_c; p.o_hi = w.h[0]; p.o_lo = w.h[1]; p.st_in = w.st_b; p.parts_in = parts; p.cnt_in = cnt; ph.push_back(p);
        p = zero_phase();                  // x += h W2^T + b2 ; statistics for the next layer's LN1
        p.type = PH_GEMM; p.M = M; p.N = W; p.K = 4 * W; p.BN = bn_res; p.epi = EPI_RESID; p.map_a = m_h; p.map_w = m_fc2;
        p.v1 = L.fc2_b; p.f0 = w.x; p.o_hi = w.x16[0]; p.o_lo = w.x16[1]; p.st_out = w.st_a; ph.push_back(p);
    }
    {
        Phase p = zero_phase(); p.type = PH_HEAD; p.v0 = wts->lnpost_w; p.v1 = wts->lnpost_b; p.v2 = wts->proj_t; p.f0 = w.raw; p.f1 = w.x; ph.push_back(p);
        p = zero_phase(); p.type = PH_L2NORM; p.f0 = w.raw; ph.push_back(p);
    }
    const size_t ph_bytes = align_up(ph.size() * sizeof(Phase), 128);
    std::vector<uint8_t> host(ph_bytes + maps.size() * sizeof(CUtensorMap));
    memcpy(host.data(), ph.data(), ph.size() * sizeof(Phase));
    memcpy(host.data() + ph_bytes, maps.data(), maps.size() * sizeof(CUtensorMap));
    if (host.size() > plan_bytes) return (int)cudaErrorInvalidValue;
    cudaError_t e = cudaMemcpyAsync(plan, host.data(), host.size(), cudaMemcpyHostToDevice, stream);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return (int)e;
    if (n_phases_out) *n_phases_out = (int)ph.size();
    return 0;
}

template <int SPLIT>
static int launch_tower(const Phase* phases, const CUtensorMap* maps, int p0, int p1, unsigned* counter, const float* images,
                        float* emb, float* raw, void* e_hi, void* e_lo, bool cooperative, cudaStream_t stream, int group = 0)
{
    cudaError_t e = cudaFuncSetAttribute(clip_tower_kernel<SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    __nv_bfloat16* eh = reinterpret_cast<__nv_bfloat16*>(e_hi);
    __nv_bfloat16* el = reinterpret_cast<__nv_bfloat16*>(e_lo);
    if (cooperative) {
        // group = 0: the whole tower in one launch. group = g > 0: one cooperative launch per g phases — pieces of ~0.1 ms that a
        // stream of other kernels (the training step's render launches) can interleave with instead of one 1-3 ms block
        const int step = group > 0 ? group : (p1 - p0);
        for (int q0 = p0; q0 < p1; q0 += step) {
            int q1 = q0 + step < p1 ? q0 + step : p1;
            if (group > 0 && q0 == p0) q1 = (p0 + 3 < p1) ? p0 + 3 : p1;      // the embedding phases first, then `group` phases at a time
            e = cudaMemsetAsync(counter, 0, sizeof(unsigned), stream);
            if (e != cudaSuccess) return (int)e;
            void* args[] = {(void*)&phases, (void*)&maps, (void*)&q0, (void*)&q1, (void*)&counter, (void*)&images, (void*)&emb, (void*)&raw, (void*)&eh, (void*)&el};
            e = cudaLaunchCooperativeKernel((const void*)clip_tower_kernel<SPLIT>, dim3(sms), dim3(kThreads), args, kSmemBytes, stream);
            if (e != cudaSuccess) return (int)e;
            if (group > 0 && q0 == p0) q0 = q1 - step;
        }
        return 0;
    }
    for (int p = p0; p < p1; ++p) {                      // one ordinary launch per phase: the profiling / debugging form
        clip_tower_kernel<SPLIT><<<sms, kThreads, kSmemBytes, stream>>>(phases, maps, p, p + 1, counter, images, emb, raw, eh, el);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

// images [B,3,S,S] fp32 (CLIP-normalised) -> emb [B,D] L2-normalised (+ optional unnormalised copy, + hi/lo bf16 planes).
// mode 0: ONE cooperative launch for the whole tower; mode 1: one launch per phase (same device code; what ncu attributes
// per GEMM); mode k >= 2: per-phase launches of the first k - 1 phases only (diagnostics); mode -g: cooperative launches of g phases
// each (the three embedding phases first). plan / workspace: as prepared by sc_clip_tower_plan for the same cfg and batch.
extern "C" int sc_clip_tower_encode(const ScClipConfig* cfg, const void* plan, int n_phases, void* workspace, const float* images,
                                    float* emb, float* emb_unnormalised, void* emb_hi, void* emb_lo, int mode, cudaStream_t stream)
{
    if (cfg == nullptr || plan == nullptr || workspace == nullptr || images == nullptr || n_phases <= 0) return (int)cudaErrorInvalidValue;
    const size_t n_phase_cap = 3 + 5 * (size_t)cfg->layers + 2;
    const Phase* phases = reinterpret_cast<const Phase*>(plan);
    const CUtensorMap* maps = reinterpret_cast<const CUtensorMap*>(reinterpret_cast<const uint8_t*>(plan) + align_up(n_phase_cap * sizeof(Phase), 128));
    unsigned* counter = reinterpret_cast<unsigned*>(workspace);
    if (mode >= 2) n_phases = (mode - 1 < n_phases) ? mode - 1 : n_phases;      // diagnostics: stop after the first (mode - 1) phases
    const int group = mode < 0 ? -mode : 0;                                     // mode -g: cooperative launches of g phases each
    if (cfg->split) return launch_tower<3>(phases, maps, 0, n_phases, counter, images, emb, emb_unnormalised, emb_hi, emb_lo, mode <= 0, stream, group);
    return launch_tower<1>(phases, maps, 0, n_phases, counter, images, emb, emb_unnormalised, emb_hi, emb_lo, mode <= 0, stream, group);
}
