// render_tc.cuh — tensor-core (tcgen05 / TMEM) edition of the fused render kernels: shared definitions.
//
// Same tile program as render_tile.cuh (128 sample points per tile, one persistent CTA per SM), but every
// [128 x 64 x 64] layer GEMM is a burst of tcgen05.mma (UMMA 128x64x16) with the accumulator in TMEM, and every
// weight-gradient GEMM (contraction over the tile's 128 points) is a UMMA 64x64x16 whose accumulator STAYS in TMEM
// for the whole kernel (12 matrices interleaved in 384 columns) — no per-tile flush.
// FP32-class accuracy on bf16 tensor cores: every operand is a hi/lo bf16 pair, every product 3 MMAs
//      (Ah + Al)(Wh + Wl) ~= Ah Wh + Ah Wl + Al Wh          (2^-17 relative operand error; tf32 would be 2^-11)
//
// Shared-memory operand format: one "plane" = [128 rows (points)] x [64 bf16 = 128 B], 16-byte chunks XOR-swizzled with
// (row & 7) (the SWIZZLE_128B image TMA produces). The SAME bytes serve as
//   * K-major A operand of a layer GEMM      (M = point, K = feature), and
//   * MN-major A/B operand of a weight-grad  (M/N = feature, K = point).
// Thread mapping of the 512 threads: TMEM lane / point row r = 32*(warp & 3) + lane, column group ch = warp >> 2
// (columns 16 ch .. 16 ch + 15) — what tcgen05.ld.32x32b.x16 delivers. 16 warps hide the MMA / L2 latencies of the
// strictly serial phase chain far better than 8 (ncu: 8-warp version 15 % issue-active, 37 % long-scoreboard stalls).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_tc.cuh"
#include "render_common.cuh"
#include "sc_b200.h"

namespace sct {

using scr::HID; using scr::NPE; using scr::M_TILE;

constexpr int kThreads = 512;                     // 16 warps: TMEM lane quarter = warp & 3, column group = warp >> 2
constexpr int NC = 16;                            // accumulator columns per thread
constexpr int kPlaneBytes = 128 * 128;            // one bf16 plane of a [128 x 64] tile
constexpr int kActBytes = 2 * kPlaneBytes;        // hi + lo
constexpr int kWPlaneBytes = 64 * 128;            // weight tile [64 x 64] bf16
constexpr int kWSegBytes = 2 * kWPlaneBytes;      // hi + lo = 16 KB
constexpr int kNumAct = 5;                        // P X Y Z U (the forward kernel uses 4)
constexpr int kWSlots = 4;                        // weight ring depth of the forward kernel (prefetch distance 3)

// ---- packed weight blob (bytes): 24 swizzled bf16 hi/lo segments, then the fp32 const + latent regions of the FFMA blob
enum SegTC : int {
    A0N = 0, B1N, A1N, B2N, A2N, W3N, W4N, W5FN, V0PN, V0FN, V1N, V2N,       // rows = out, cols = in  (y = W x)
    W4T, W3T, B2T, B1T, A2T, A1T, A0T, W5FT, V2T, V1T, V0FT, V0PT,          // rows = in, cols = out  (x_bar = W^T y_bar)
    NSEG_TC
};
constexpr size_t kSegRegionBytes = (size_t)NSEG_TC * kWSegBytes;             // 393 216
constexpr size_t kTcConstOffsetBytes = kSegRegionBytes;                      // fp32 consts (scr::kConstFloats)
constexpr size_t kTcLatentOffsetBytes = kTcConstOffsetBytes + (size_t)scr::kConstFloats * 4;
constexpr size_t kTcBlobBytes = kTcLatentOffsetBytes + (size_t)scr::kLatentFloats * 4;

// ---- TMEM columns
constexpr uint32_t TM_ACC0 = 0, TM_ACC1 = 64, TM_WGRAD = 128;               // 12 weight-grad matrices, 2 per 64 columns
enum WG : int { WG_A0 = 0, WG_A1, WG_A2, WG_B1, WG_B2, WG_W3, WG_W4, WG_W5F, WG_V0P, WG_V0F, WG_V1, WG_V2, NWG };
__device__ __forceinline__ uint32_t wg_taddr(uint32_t tmem_base, int m) {     // matrix m: columns 128 + 64 (m/2), lanes +16 (m&1)
    return tmem_base + TM_WGRAD + 64u * (uint32_t)(m >> 1) + ((uint32_t)((m & 1) * 16) << 16);
}

// ---- shared-memory map (bytes; the dynamic buffer is aligned to 1024 in the kernel)
// forward kernel: 4 activation buffers + 4 weight slots; backward kernel: 5 + 2 (both 192 KB)
constexpr int SMB_ACT = 0;
constexpr int SMB_W_FWD = SMB_ACT + 4 * kActBytes;                    // forward: weight slots start here
constexpr int SMB_W_BWD = SMB_ACT + 5 * kActBytes;                    // backward
constexpr int SMB_F32 = SMB_ACT + 6 * kActBytes;                      // float area (after 192 KB of operands)
constexpr int SF_CONST = 0;
constexpr int SF_CB = SF_CONST + scr::kConstFloats;
constexpr int SF_BIAS = SF_CB + 256;                                   // [8][64] per-tile bias tables
constexpr int SF_PT = SF_BIAS + 512;
constexpr int kPtVecsTc = scr::kPtVecs + 16;                          // + 4 scratch vectors x 4 column groups
constexpr int SF_RAY = SF_PT + kPtVecsTc * M_TILE;                    // per-ray scratch (<= 32 rays): 704 floats
constexpr int SF_VACC = SF_RAY + 704;                                 // backward: vector-gradient accumulators (588 floats)
constexpr int SF_MISC = SF_VACC + 592;                                // seq[64] bytes, seq_len, wg_mask
constexpr int SF_END = SF_MISC + 32;
constexpr int SMB_BAR = SMB_F32 + SF_END * 4;                         // mbarriers: wfull[4] wfree[4] mma_done . tmem-slot ready[4] . ray_full[2] ray_free[2]
constexpr int BAR_WFULL = 0, BAR_WFREE = 4, BAR_MMA_DONE = 8, BAR_TMEM_SLOT = 10, BAR_READY = 11, BAR_RAY_FULL = 16, BAR_RAY_FREE = 18;
constexpr int SMB_CMD = SMB_BAR + 192;                                // role-split backward: command ring, 4 batches x 8 words
constexpr int kCmdBatches = 4, kCmdWords = 8;
constexpr int kSmemBytesTc = SMB_CMD + kCmdBatches * kCmdWords * 4;   // all-dynamic, __align__(1024): no static smem, no slack
static_assert(kSmemBytesTc <= 232448, "exceeds the 227 KB shared memory of one CTA");
enum PtVecTc : int { PX_A = scr::kPtVecs, PX_B = PX_A + 4, PX_C = PX_B + 4, PX_D = PX_C + 4 };   // 4 scratch vectors x [4 column groups]

// ---------------------------------------------------------------------------------------------------------
// hi = bf16_rn(v), lo = bf16_rn(v - hi): hi + lo reproduces v to ~2^-17 relative. Two values per F2FP.
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    const float la = a - __uint_as_float(hi << 16), lb = b - __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 t = __floats2bfloat162_rn(la, lb);
    lo = *reinterpret_cast<const uint32_t*>(&t);
}

// write NC consecutive columns (group ch) of row r into a hi/lo plane pair
__device__ __forceinline__ void row_store(uint8_t* act, int r, int ch, const float (&v)[NC]) {
    uint8_t* hi = act + r * 128;
    uint8_t* lo = hi + kPlaneBytes;
#pragma unroll
    for (int q = 0; q < NC / 8; ++q) {
        const int pos = ((ch * (NC / 8) + q) ^ (r & 7)) << 4;
        uint4 h, l;
        split_pair(v[q * 8 + 0], v[q * 8 + 1], h.x, l.x);
        split_pair(v[q * 8 + 2], v[q * 8 + 3], h.y, l.y);
        split_pair(v[q * 8 + 4], v[q * 8 + 5], h.z, l.z);
        split_pair(v[q * 8 + 6], v[q * 8 + 7], h.w, l.w);
        *reinterpret_cast<uint4*>(hi + pos) = h;
        *reinterpret_cast<uint4*>(lo + pos) = l;
    }
}
// single element (row r, column c) of a plane pair
__device__ __forceinline__ float act_elem(const uint8_t* act, int r, int c) {
    const int off = r * 128 + ((((c >> 3) ^ (r & 7))) << 4) + ((c & 7) << 1);
    return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(act + off)) +
           __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(act + kPlaneBytes + off));
}

// global stash planes (128 x 64 fp32): only ever re-read by the thread that wrote them, so the layout is chosen for
// coalescing: float4 index (ch * NC/4 + q) * 128 + row — consecutive lanes (rows) are consecutive 16-byte words.
__device__ __forceinline__ void st_store(float* plane, int r, int ch, const float (&v)[NC]) {
    float4* p = reinterpret_cast<float4*>(plane) + (ch * (NC / 4)) * 128 + r;
#pragma unroll
    for (int q = 0; q < NC / 4; ++q) __stcg(p + q * 128, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
}
__device__ __forceinline__ void st_load(const float* plane, int r, int ch, float (&v)[NC]) {
    const float4* p = reinterpret_cast<const float4*>(plane) + (ch * (NC / 4)) * 128 + r;
#pragma unroll
    for (int q = 0; q < NC / 4; ++q) { const float4 t = __ldcg(p + q * 128); v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[NC]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
constexpr int kStashPlane = 128 * 64;     // floats per stash plane

// ---- UMMA descriptors for this layout
// K-major (layer GEMMs): rows of 128 B, 8-row groups 1024 B apart — identical to gemm_tc.cuh::make_smem_desc_k128
// MN-major (weight-grad GEMMs): MN = the 64 features of one 128-B row, K = rows; 8-row groups 1024 B apart (SBO)
__device__ __forceinline__ uint64_t make_smem_desc_mn128(const void* tile) {
    uint64_t d = 0;
    d |= (uint64_t)((sctc::smem_u32(tile) >> 4) & 0x3FFF);
    d |= (uint64_t)(1024 >> 4) << 16;                          // LBO: next 64-wide MN group (unused: MN = 64)
    d |= (uint64_t)(1024 >> 4) << 32;                          // SBO: next group of 8 K rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                                    // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {     // both operands MN-major
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- MMA issue. Executed by ALL 32 lanes of the issuing warp (warp 0) with warp-uniform operands; one elected lane
// executes the tcgen05.mma itself. (Issued from divergent `if (tid == 0)` code the compiler cannot keep the descriptors in
// uniform registers and wraps EVERY MMA in an ELECT / 5x R2UR.BROADCAST / BRA.U.ANY loop: ~100 cycles per MMA against 32 of
// tensor-pipe time — measured 1.3 k cycles of issue per 12-MMA layer GEMM on the critical path of every phase.)
__device__ __forceinline__ uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(sctc::smem_u32(bar)) : "memory");
}
// low / high 32-bit words of the K-major / MN-major shared-memory descriptors of a (warp-uniform) byte address
// (start address field + LBO | SBO, version, swizzle)
constexpr uint32_t kDescHi = 0x40004040u;
__device__ __forceinline__ uint32_t desc_lo_k128(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_lo_mn128(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | ((1024u >> 4) << 16); }

#if defined(SC_TC_NOINLINE_ISSUE) && !defined(SC_TC_ROLE_SPLIT)   // (was: the backward kernel before it got a dedicated issuing warp)
#define SC_TC_ISSUE_FN static __device__ __noinline__
#else
#define SC_TC_ISSUE_FN __device__ __forceinline__
#endif
// D[128 x 64] (+)= ACT[128 x 64] . W^T   (3 MMAs per 16-wide k-step). Called by the whole issuing warp.
// NOT inlined in the backward kernel (here and issue_wgrad): it has ~60 issue sites of 12-24 MMAs each; inlined they were 5 k
// SASS instructions (25 % of the kernel) that the issuing warp walked once per tile, every line an instruction-cache miss
// on the critical path of the phase (ncu: 29 % of the warp samples stalled on instruction fetch).
SC_TC_ISSUE_FN void issue_layer_gemm(uint32_t tmem_d, const uint8_t* act, const uint8_t* w, bool accumulate) {
    // ONE asm block for the 12 MMAs: descriptor low words are base + constant (the high word 0x40004040 = SBO 1024 B, version 1,
    // SWIZZLE_128B never changes), the instruction descriptor and the elected-lane predicate are set up once. Issued as 12
    // separate asm statements every MMA re-materialised its idesc and 64-bit descriptors in the (slow, scalar) uniform datapath:
    // ~70 cycles per MMA against 32 cycles of tensor-pipe time (clock64 trace: 840 cycles of issue per layer GEMM).
    constexpr uint32_t idesc = sctc::make_idesc_bf16(128, 64);
    const uint32_t d = uniform_u32(tmem_d), acc = uniform_u32(accumulate ? 1u : 0u);
    const uint32_t a0 = uniform_u32(sctc::smem_u32(act)), w0 = uniform_u32(sctc::smem_u32(w));
    const uint32_t ah = desc_lo_k128(a0), al = desc_lo_k128(a0 + kPlaneBytes), wh = desc_lo_k128(w0), wl = desc_lo_k128(w0 + kWPlaneBytes);
    asm volatile(
        "{\n\t"
        ".reg .pred p, q, t;\n\t"
        ".reg .b32 xa, xb, ya, yb;\n\t"
        ".reg .b64 da, db, ea, eb;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.u32 t, %5, %5;\n\t"
        "add.u32 xa, %1, 0; add.u32 xb, %2, 0; add.u32 ya, %3, 0; add.u32 yb, %4, 0;\n\t"
        "mov.b64 da, {xa, %6}; mov.b64 db, {xb, %6}; mov.b64 ea, {ya, %6}; mov.b64 eb, {yb, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %7, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, eb, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "add.u32 xa, %1, 2; add.u32 xb, %2, 2; add.u32 ya, %3, 2; add.u32 yb, %4, 2;\n\t"
        "mov.b64 da, {xa, %6}; mov.b64 db, {xb, %6}; mov.b64 ea, {ya, %6}; mov.b64 eb, {yb, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, eb, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "add.u32 xa, %1, 4; add.u32 xb, %2, 4; add.u32 ya, %3, 4; add.u32 yb, %4, 4;\n\t"
        "mov.b64 da, {xa, %6}; mov.b64 db, {xb, %6}; mov.b64 ea, {ya, %6}; mov.b64 eb, {yb, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, eb, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "add.u32 xa, %1, 6; add.u32 xb, %2, 6; add.u32 ya, %3, 6; add.u32 yb, %4, 6;\n\t"
        "mov.b64 da, {xa, %6}; mov.b64 db, {xb, %6}; mov.b64 ea, {ya, %6}; mov.b64 eb, {yb, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, eb, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "}"
        ::"r"(d), "r"(ah), "r"(al), "r"(wh), "r"(wl), "r"(acc), "r"(kDescHi), "r"(idesc) : "memory");
}
// Wide form (forward kernel, where TMEM has room): D[128 x 128] at tmem_d. The weight segment's hi and lo planes are contiguous,
// so ONE N = 128 MMA computes Ah.Wh^T into columns 0..63 and Ah.Wl^T into columns 64..127, and an N = 64 MMA adds Al.Wh^T to
// columns 0..63; the epilogue adds the two halves. 8 MMAs instead of 12 per GEMM, and the activation tile is read from shared
// memory twice instead of three times per k-step: at N = 64 a 128x64x16 MMA is bound by its operand reads (4 KB of A + 2 KB of
// B at 128 B/clk against 32 cycles of math; measured ~70 cycles per MMA).
__device__ __forceinline__ void issue_layer_gemm_wide(uint32_t tmem_d, const uint8_t* act, const uint8_t* w, bool accumulate) {
    constexpr uint32_t idesc128 = sctc::make_idesc_bf16(128, 128), idesc64 = sctc::make_idesc_bf16(128, 64);
    const uint32_t d = uniform_u32(tmem_d), acc = uniform_u32(accumulate ? 1u : 0u);
    const uint32_t a0 = uniform_u32(sctc::smem_u32(act)), w0 = uniform_u32(sctc::smem_u32(w));
    const uint32_t ah = desc_lo_k128(a0), al = desc_lo_k128(a0 + kPlaneBytes), wh = desc_lo_k128(w0);
    asm volatile(
        "{\n\t"
        ".reg .pred p, q, t;\n\t"
        ".reg .b32 xa, xb, ya;\n\t"
        ".reg .b64 da, db, ea;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.u32 t, %4, %4;\n\t"
        "add.u32 xa, %1, 0; add.u32 xb, %2, 0; add.u32 ya, %3, 0;\n\t"
        "mov.b64 da, {xa, %5}; mov.b64 db, {xb, %5}; mov.b64 ea, {ya, %5};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %6, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "add.u32 xa, %1, 2; add.u32 xb, %2, 2; add.u32 ya, %3, 2;\n\t"
        "mov.b64 da, {xa, %5}; mov.b64 db, {xb, %5}; mov.b64 ea, {ya, %5};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %6, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "add.u32 xa, %1, 4; add.u32 xb, %2, 4; add.u32 ya, %3, 4;\n\t"
        "mov.b64 da, {xa, %5}; mov.b64 db, {xb, %5}; mov.b64 ea, {ya, %5};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %6, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "add.u32 xa, %1, 6; add.u32 xb, %2, 6; add.u32 ya, %3, 6;\n\t"
        "mov.b64 da, {xa, %5}; mov.b64 db, {xb, %5}; mov.b64 ea, {ya, %5};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %6, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "}"
        ::"r"(d), "r"(ah), "r"(al), "r"(wh), "r"(acc), "r"(kDescHi), "r"(idesc128), "r"(idesc64) : "memory");
}
// D[64 x 64] += L^T . R over the tile's 128 points (L, R = plane pairs). Called by the whole issuing warp. 24 MMAs, one asm block.
SC_TC_ISSUE_FN void issue_wgrad(uint32_t tmem_d, const uint8_t* L, const uint8_t* R, bool accumulate) {
    constexpr uint32_t idesc = make_idesc_bf16_mn(64, 64);
    const uint32_t d = uniform_u32(tmem_d), acc = uniform_u32(accumulate ? 1u : 0u);
    const uint32_t l0 = uniform_u32(sctc::smem_u32(L)), r0 = uniform_u32(sctc::smem_u32(R));
    const uint32_t lh = desc_lo_mn128(l0), ll = desc_lo_mn128(l0 + kPlaneBytes), rh = desc_lo_mn128(r0), rl = desc_lo_mn128(r0 + kPlaneBytes);
    asm volatile(
        "{\n\t"
        ".reg .pred p, q, t;\n\t"
        ".reg .b32 xa, xb, ya, yb;\n\t"
        ".reg .b64 da, db, ea, eb;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.u32 t, %5, %5;\n\t"
        "add.u32 xa, %1, 0; add.u32 xb, %2, 0; add.u32 ya, %3, 0; add.u32 yb, %4, 0;\n\t"
        "mov.b64 da, {xa, %6}; mov.b64 db, {xb, %6}; mov.b64 ea, {ya, %6}; mov.b64 eb, {yb, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %7, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, eb, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "add.u32 xa, %1, 128; add.u32 xb, %2, 128; add.u32 ya, %3, 128; add.u32 yb, %4, 128;\n\t"
        "mov.b64 da, {xa, %6}; mov.b64 db, {xb, %6}; mov.b64 ea, {ya, %6}; mov.b64 eb, {yb, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, eb, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "add.u32 xa, %1, 256; add.u32 xb, %2, 256; add.u32 ya, %3, 256; add.u32 yb, %4, 256;\n\t"
        "mov.b64 da, {xa, %6}; mov.b64 db, {xb, %6}; mov.b64 ea, {ya, %6}; mov.b64 eb, {yb, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, eb, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "add.u32 xa, %1, 384; add.u32 xb, %2, 384; add.u32 ya, %3, 384; add.u32 yb, %4, 384;\n\t"
        "mov.b64 da, {xa, %6}; mov.b64 db, {xb, %6}; mov.b64 ea, {ya, %6}; mov.b64 eb, {yb, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, eb, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "add.u32 xa, %1, 512; add.u32 xb, %2, 512; add.u32 ya, %3, 512; add.u32 yb, %4, 512;\n\t"
        "mov.b64 da, {xa, %6}; mov.b64 db, {xb, %6}; mov.b64 ea, {ya, %6}; mov.b64 eb, {yb, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, eb, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "add.u32 xa, %1, 640; add.u32 xb, %2, 640; add.u32 ya, %3, 640; add.u32 yb, %4, 640;\n\t"
        "mov.b64 da, {xa, %6}; mov.b64 db, {xb, %6}; mov.b64 ea, {ya, %6}; mov.b64 eb, {yb, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, eb, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "add.u32 xa, %1, 768; add.u32 xb, %2, 768; add.u32 ya, %3, 768; add.u32 yb, %4, 768;\n\t"
        "mov.b64 da, {xa, %6}; mov.b64 db, {xb, %6}; mov.b64 ea, {ya, %6}; mov.b64 eb, {yb, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, eb, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "add.u32 xa, %1, 896; add.u32 xb, %2, 896; add.u32 ya, %3, 896; add.u32 yb, %4, 896;\n\t"
        "mov.b64 da, {xa, %6}; mov.b64 db, {xb, %6}; mov.b64 ea, {ya, %6}; mov.b64 eb, {yb, %6};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, eb, %7, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], db, ea, %7, t;\n\t"
        "}"
        ::"r"(d), "r"(lh), "r"(ll), "r"(rh), "r"(rl), "r"(acc), "r"(kDescHi), "r"(idesc) : "memory");
}

// Single-MMA forms (ScRenderArgs::precision = 1, plain bf16 operands): the hi planes alone, 4 / 8 MMAs instead of 12 / 24.
SC_TC_ISSUE_FN void issue_layer_gemm_single(uint32_t tmem_d, const uint8_t* act, const uint8_t* w, bool accumulate) {
    constexpr uint32_t idesc = sctc::make_idesc_bf16(128, 64);
    const uint32_t d = uniform_u32(tmem_d), acc = uniform_u32(accumulate ? 1u : 0u);
    const uint32_t ah = desc_lo_k128(uniform_u32(sctc::smem_u32(act))), wh = desc_lo_k128(uniform_u32(sctc::smem_u32(w)));
    asm volatile(
        "{\n\t"
        ".reg .pred p, q, t;\n\t"
        ".reg .b32 xa, ya;\n\t"
        ".reg .b64 da, ea;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %3, 0;\n\t"
        "setp.eq.u32 t, %3, %3;\n\t"
        "add.u32 xa, %1, 0; add.u32 ya, %2, 0; mov.b64 da, {xa, %4}; mov.b64 ea, {ya, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %5, p;\n\t"
        "add.u32 xa, %1, 2; add.u32 ya, %2, 2; mov.b64 da, {xa, %4}; mov.b64 ea, {ya, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %5, t;\n\t"
        "add.u32 xa, %1, 4; add.u32 ya, %2, 4; mov.b64 da, {xa, %4}; mov.b64 ea, {ya, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %5, t;\n\t"
        "add.u32 xa, %1, 6; add.u32 ya, %2, 6; mov.b64 da, {xa, %4}; mov.b64 ea, {ya, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %5, t;\n\t"
        "}"
        ::"r"(d), "r"(ah), "r"(wh), "r"(acc), "r"(kDescHi), "r"(idesc) : "memory");
}
SC_TC_ISSUE_FN void issue_wgrad_single(uint32_t tmem_d, const uint8_t* L, const uint8_t* R, bool accumulate) {
    constexpr uint32_t idesc = make_idesc_bf16_mn(64, 64);
    const uint32_t d = uniform_u32(tmem_d), acc = uniform_u32(accumulate ? 1u : 0u);
    const uint32_t lh = desc_lo_mn128(uniform_u32(sctc::smem_u32(L))), rh = desc_lo_mn128(uniform_u32(sctc::smem_u32(R)));
    asm volatile(
        "{\n\t"
        ".reg .pred p, q, t;\n\t"
        ".reg .b32 xa, ya;\n\t"
        ".reg .b64 da, ea;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %3, 0;\n\t"
        "setp.eq.u32 t, %3, %3;\n\t"
        "add.u32 xa, %1, 0; add.u32 ya, %2, 0; mov.b64 da, {xa, %4}; mov.b64 ea, {ya, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %5, p;\n\t"
        "add.u32 xa, %1, 128; add.u32 ya, %2, 128; mov.b64 da, {xa, %4}; mov.b64 ea, {ya, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %5, t;\n\t"
        "add.u32 xa, %1, 256; add.u32 ya, %2, 256; mov.b64 da, {xa, %4}; mov.b64 ea, {ya, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %5, t;\n\t"
        "add.u32 xa, %1, 384; add.u32 ya, %2, 384; mov.b64 da, {xa, %4}; mov.b64 ea, {ya, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %5, t;\n\t"
        "add.u32 xa, %1, 512; add.u32 ya, %2, 512; mov.b64 da, {xa, %4}; mov.b64 ea, {ya, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %5, t;\n\t"
        "add.u32 xa, %1, 640; add.u32 ya, %2, 640; mov.b64 da, {xa, %4}; mov.b64 ea, {ya, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %5, t;\n\t"
        "add.u32 xa, %1, 768; add.u32 ya, %2, 768; mov.b64 da, {xa, %4}; mov.b64 ea, {ya, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %5, t;\n\t"
        "add.u32 xa, %1, 896; add.u32 ya, %2, 896; mov.b64 da, {xa, %4}; mov.b64 ea, {ya, %4};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, ea, %5, t;\n\t"
        "}"
        ::"r"(d), "r"(lh), "r"(rh), "r"(acc), "r"(kDescHi), "r"(idesc) : "memory");
}

// ---- weight ring: NS slots, TMA-filled (wfull), released by tcgen05.commit (wfree). Prefetch distance NS - 1:
// a 16 KB bulk copy from L2 takes ~1-2 k cycles, longer than one tensor-core layer phase.
struct WeightRing {
    const uint8_t* blob;
    uint8_t* slots;
    uint64_t *wfull, *wfree;      // [NS] each
    const int8_t* seq;
    int seq_len;
    int NS;
    uint32_t n;                   // matrices consumed
    int pos_fetch;                // position in seq of the next matrix to fetch
    bool w0;                      // this thread belongs to the issuing warp (warp 0); its 32 lanes keep n / pos_fetch in step

    // issuing warp, all lanes (state); one elected lane starts the copy
    __device__ __forceinline__ void issue(uint32_t idx) {
        const uint32_t s = idx % (uint32_t)NS;
        if (elect_one()) {
            uint64_t* bar = wfull + s;
            scr::mbar_expect_tx(bar, kWSegBytes);
            scr::tma_bulk_g2s(slots + s * kWSegBytes, blob + (size_t)seq[pos_fetch] * kWSegBytes, kWSegBytes, bar);
        }
        __syncwarp();
        pos_fetch = (pos_fetch + 1 == seq_len) ? 0 : pos_fetch + 1;
    }
    __device__ __forceinline__ void prologue() {
        n = 0; pos_fetch = 0;
        if (w0) for (int i = 0; i < NS - 1; ++i) issue((uint32_t)i);
    }
    // all threads: publishes the operand stores of the previous epilogue to the async proxy, syncs the CTA and returns the
    // slot of matrix n. Only the issuing warp waits for the weights to land.
    __device__ __forceinline__ const uint8_t* acquire() {
        sctc::fence_proxy_async();
        sctc::tc_fence_before();
        __syncthreads();
        sctc::tc_fence_after();
        const uint32_t cur = n;
        if (w0) scr::mbar_wait(wfull + (cur % NS), (cur / NS) & 1);
        n = cur + 1;
        return slots + (cur % NS) * kWSegBytes;
    }
    // role-split kernels: the dedicated issuing warp owns the ring; the operand hand-over is the command queue's business
    __device__ __forceinline__ const uint8_t* wait_weights() {
        const uint32_t cur = n;
        scr::mbar_wait(wfull + (cur % NS), (cur / NS) & 1);
        n = cur + 1;
        return slots + (cur % NS) * kWSegBytes;
    }
    __device__ __forceinline__ void drain_issuer() {
        for (uint32_t m = n; m < n + (uint32_t)NS - 1; ++m) scr::mbar_wait(wfull + (m % NS), (m / NS) & 1);
    }
    // issuing warp, AFTER issuing the MMAs that read the slot returned by the last acquire(): release that slot when they
    // complete, then prefetch matrix n + NS - 2 into the slot of the matrix before it (waiting for ITS MMAs, which are
    // ahead of ours in the tensor pipe, so the wait overlaps useful work instead of delaying the issue).
    __device__ __forceinline__ void release() {
        const uint32_t cur = n - 1;
        umma_commit_elect(wfree + (cur % NS));
        const uint32_t nx = cur + (uint32_t)NS - 1;
        if (cur >= 1) scr::mbar_wait(wfree + (nx % NS), ((nx / NS) - 1) & 1);
        issue(nx);
    }
    // NS - 1 copies are always in flight: wait for them before the CTA exits
    __device__ __forceinline__ void drain() {
        if (w0)
            for (uint32_t m = n; m < n + (uint32_t)NS - 1; ++m) scr::mbar_wait(wfull + (m % NS), (m / NS) & 1);
        __syncthreads();
    }
};

}  // namespace sct
