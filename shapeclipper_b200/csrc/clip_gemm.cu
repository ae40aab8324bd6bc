// clip_gemm.cu — tcgen05 GEMM used by the CLIP ViT image tower.
//   C[M,N] = A[M,K] . W[N,K]^T (+bias) (QuickGELU) (+residual)  ->  fp32 and/or hi/lo bf16 planes
// Replaces the cuBLAS fp16 GEMMs behind clip_model.encode_image (CLIP_anno.py:166) and the per-query GEMV loop of
// NN_annotator.calc_matches (CLIP_anno.py:29-57). See gemm_tc.cuh for the tcgen05/TMA primitives.
//
// Kernel anatomy (192 threads, one 128 x BN output tile per CTA):
//   warp 0      TMA producer: per 64-wide k-block loads A_hi[,A_lo] (128 x 64) and W_hi[,W_lo] (BN x 64), SWIZZLE_128B
//   warp 1      TMEM allocator + MMA issuer: 4 k-steps x {1|3} tcgen05.mma (128 x BN x 16) per k-block, tcgen05.commit
//   warps 2..5  epilogue: tcgen05.ld 32x32b.x32 -> bias / QuickGELU / residual -> global (fp32 and/or split bf16)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_tc.cuh"
#include "sc_b200.h"

namespace sctc {

struct GemmEpilogue {
    const float* bias;        // [N] or null
    const float* residual;    // fp32 [M, ld] or null (may alias out_f32)
    float* out_f32;           // fp32 [M, ld] or null
    __nv_bfloat16* out_hi;    // bf16 [M, ld] or null
    __nv_bfloat16* out_lo;    // bf16 [M, ld] or null (residual v - hi)
    int ld;                   // leading dimension of the outputs (= N)
    int act;                  // 0 none, 1 QuickGELU x*sigmoid(1.702x)
    float scale;              // applied to the accumulator before the bias (1 = none)
};

constexpr int BM = 128, BK = 64, kGemmThreads = 192;

template <int BN, int SPLIT>
struct GemmCfg {
    static constexpr int kABytes = BM * BK * 2, kWBytes = BN * BK * 2;
    static constexpr int kStageBytes = (SPLIT == 3 ? 2 : 1) * (kABytes + kWBytes);
    static constexpr int kStages = (200 * 1024) / kStageBytes > 6 ? 6 : (200 * 1024) / kStageBytes;
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int kTmemCols = BN < 32 ? 32 : BN;
};

template <int BN, int SPLIT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapAh, const __grid_constant__ CUtensorMap mapAl,
               const __grid_constant__ CUtensorMap mapWh, const __grid_constant__ CUtensorMap mapWl,
               GemmEpilogue epi, int M, int N, int K)
{
    using Cfg = GemmCfg<BN, SPLIT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* empty = full + Cfg::kStages;
    uint64_t* tmem_full = empty + Cfg::kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    const int num_kb = K / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapAh); tma_prefetch_desc(&mapWh);
        if (SPLIT == 3) { tma_prefetch_desc(&mapAl); tma_prefetch_desc(&mapWl); }
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % Cfg::kStages;
                const uint32_t ph = (kb / Cfg::kStages) & 1;
                mbar_wait(empty + s, ph ^ 1);
                uint8_t* st = smem + s * Cfg::kStageBytes;
                mbar_expect_tx(full + s, Cfg::kStageBytes);
                tma_load_2d(st, &mapAh, full + s, kb * BK, m0);
                tma_load_2d(st + Cfg::kABytes, &mapWh, full + s, kb * BK, n0);
                if (SPLIT == 3) {
                    tma_load_2d(st + Cfg::kABytes + Cfg::kWBytes, &mapAl, full + s, kb * BK, m0);
                    tma_load_2d(st + 2 * Cfg::kABytes + Cfg::kWBytes, &mapWl, full + s, kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % Cfg::kStages;
                const uint32_t ph = (kb / Cfg::kStages) & 1;
                mbar_wait(full + s, ph);
                tc_fence_after();
                uint8_t* st = smem + s * Cfg::kStageBytes;
                const uint64_t dAh = make_smem_desc_k128(st);
                const uint64_t dWh = make_smem_desc_k128(st + Cfg::kABytes);
                const uint64_t dAl = make_smem_desc_k128(st + Cfg::kABytes + Cfg::kWBytes);
                const uint64_t dWl = make_smem_desc_k128(st + 2 * Cfg::kABytes + Cfg::kWBytes);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t adv = (uint64_t)(k * 32 >> 4);          // 16 bf16 = 32 B along K inside the 128-B row
                    umma_bf16(tmem_base, dAh + adv, dWh + adv, idesc, (kb | k) != 0);
                    if (SPLIT == 3) {
                        umma_bf16(tmem_base, dAh + adv, dWl + adv, idesc, 1);
                        umma_bf16(tmem_base, dAl + adv, dWh + adv, idesc, 1);
                    }
                }
                umma_commit(empty + s);            // ring slot reusable once these MMAs have read it
            }
            umma_commit(tmem_full);                // accumulator complete
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5 -> TMEM lane quarters)
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int row = m0 + q * 32 + lane;
        const bool row_ok = row < M;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            float v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
            const int col0 = n0 + c * 32;
            if (row_ok) {
                const size_t off = (size_t)row * epi.ld + col0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float x = v[i] * epi.scale;
                    if (epi.bias) x += __ldg(epi.bias + col0 + i);
                    if (epi.act == 1) x = x / (1.f + __expf(-1.702f * x));
                    v[i] = x;
                }
                if (epi.residual) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 r = *reinterpret_cast<const float4*>(epi.residual + off + i);
                        v[i] += r.x; v[i + 1] += r.y; v[i + 2] += r.z; v[i + 3] += r.w;
                    }
                }
                if (epi.out_f32) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(epi.out_f32 + off + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
                if (epi.out_hi) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        __align__(16) __nv_bfloat16 h[8];
                        __align__(16) __nv_bfloat16 l[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            h[j] = __float2bfloat16_rn(v[i + j]);
                            l[j] = __float2bfloat16_rn(v[i + j] - __bfloat162float(h[j]));
                        }
                        *reinterpret_cast<uint4*>(epi.out_hi + off + i) = *reinterpret_cast<const uint4*>(h);
                        if (epi.out_lo) *reinterpret_cast<uint4*>(epi.out_lo + off + i) = *reinterpret_cast<const uint4*>(l);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major bf16 [rows, K] -> 2-D map with box {64, box_rows}, 128-B swizzle, zero fill out of bounds
int make_map(CUtensorMap* map, const void* base, int rows, int K, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return (int)cudaErrorNotSupported;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

template <int BN, int SPLIT>
int launch(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& wh, const CUtensorMap& wl,
           const GemmEpilogue& epi, int M, int N, int K, cudaStream_t stream)
{
    using Cfg = GemmCfg<BN, SPLIT>;
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(N / BN, (M + BM - 1) / BM);
    gemm_tc_kernel<BN, SPLIT><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ah, al, wh, wl, epi, M, N, K);
    return (int)cudaGetLastError();
}

}  // namespace sctc

using namespace sctc;

// C ABI: see include/sc_b200.h
extern "C" int sc_gemm_bf16_tc(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, int M, int N, int K,
                               const float* bias, const float* residual, int act, float scale, float* out_f32, void* out_hi,
                               void* out_lo, cudaStream_t stream)
{
    if (M <= 0) return 0;
    if (a_hi == nullptr || w_hi == nullptr || (K % BK) != 0 || (N % 64) != 0 || K <= 0 || N <= 0) return (int)cudaErrorInvalidValue;
    const bool split = (a_lo != nullptr && w_lo != nullptr);
    const int bn = (N % 128 == 0 && ((long)(N / 128) * ((M + BM - 1) / BM) >= 148)) ? 128 : 64;
    CUtensorMap ah, al, wh, wl;
    int rc = make_map(&ah, a_hi, M, K, BM);
    if (!rc) rc = make_map(&wh, w_hi, N, K, bn);
    if (!rc) rc = make_map(&al, split ? a_lo : a_hi, M, K, BM);
    if (!rc) rc = make_map(&wl, split ? w_lo : w_hi, N, K, bn);
    if (rc) return rc;
    GemmEpilogue epi;
    epi.bias = bias; epi.residual = residual; epi.out_f32 = out_f32;
    epi.out_hi = reinterpret_cast<__nv_bfloat16*>(out_hi); epi.out_lo = reinterpret_cast<__nv_bfloat16*>(out_lo);
    epi.ld = N; epi.act = act; epi.scale = scale;
    if (bn == 128) return split ? launch<128, 3>(ah, al, wh, wl, epi, M, N, K, stream) : launch<128, 1>(ah, al, wh, wl, epi, M, N, K, stream);
    return split ? launch<64, 3>(ah, al, wh, wl, epi, M, N, K, stream) : launch<64, 1>(ah, al, wh, wl, epi, M, N, K, stream);
}
