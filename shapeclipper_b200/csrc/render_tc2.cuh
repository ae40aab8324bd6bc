// render_tc2.cuh — second generation of the tensor-core render kernels: TWO independent 64-point tile chains per CTA.
//
// render_tc.cuh runs ONE 128-point tile per CTA as a strictly serial chain of ~45 GEMM phases (issue -> MMA -> tcgen05.ld ->
// epilogue -> barrier): ncu shows the tensor pipe 10 % active and the issue slots 19 % — latency-bound, with shared memory
// and TMEM both full, so a second 128-point tile cannot be in flight. Here the CTA's 16 warps form two GROUPS of 8 warps;
// each group walks its own sequence of 64-point tiles (UMMA M = 64) with its own operand buffers (half the rows, so the
// same shared memory in total), accumulators (the M = 64 accumulator layout uses 16 lanes per TMEM sub-partition: group g
// takes lanes 16 g .. 16 g + 15 of the same columns), mbarriers, named barriers and weight ring. While one group sits in
// an MMA / L2 / instruction-fetch latency the other one computes. Only the 12 weight-gradient accumulators are shared:
// both groups' UMMA 64x64x16 accumulate into the same TMEM-resident matrices (pre-zeroed; addition commutes).
//
// Thread mapping inside a group (tg = tid & 255, wg = tg >> 5): TMEM sub-partition q = wg & 3, column half = wg >> 2;
// tcgen05.ld.16x32bx2.x16 hands lanes 0-15 of a warp the 16 rows of the sub-partition at columns [c, c+16) and lanes 16-31
// the same rows at [c+16, c+32) (measured: scripts/micro/tmem_ld_16x32bx2.cu), i.e.
//     row = 16 q + (lane & 15),   column group ch = 2 (wg >> 2) + (lane >> 4),   16 consecutive columns per thread
// — the same per-thread shape as generation 1, so the epilogue arithmetic is unchanged.
#pragma once
#include "render_tc.cuh"

namespace sct2 {

using scr::HID; using scr::NPE;
using sct::NC; using sct::kWPlaneBytes; using sct::kWSegBytes; using sct::split_pair;
using sct::TM_ACC0; using sct::TM_ACC1; using sct::TM_WGRAD; using sct::NWG;

constexpr int kThreads = 512;
constexpr int kGroups = 2;
constexpr int kGT = 256;                          // threads per group
constexpr int MT = 64;                            // points per tile
constexpr int kPlaneBytes = MT * 128;             // one bf16 plane of a [64 x 64] tile
constexpr int kActBytes = 2 * kPlaneBytes;        // hi + lo = 16 KB
constexpr int kNumAct = 5;

// ---- shared-memory map (bytes, buffer aligned to 1024). 12 x 16 KB of operands: forward 2 x (4 act + 2 weight slots),
// backward 2 x (5 act + 1 weight slot). Then the float area.
constexpr int SMB_OPERANDS = 12 * kActBytes;                          // 196 608
constexpr int kPtVecs2 = scr::kPtVecs + 16;                           // + 4 scratch vectors x 4 column groups
constexpr int kRayFloats = 328;                                       // per-group ray scratch (<= 16 rays per tile)
constexpr int RAY2_ACC = 8, RAY2_BAR = 136, RAY2_NB = 264;            // [0..1] warp partials; [16][8], [16][8], [16][4]
constexpr int GF_BIAS = 0;                                            // [8][64] per-tile bias tables
constexpr int GF_PT = GF_BIAS + 512;
constexpr int GF_RAY = GF_PT + kPtVecs2 * MT;
constexpr int GF_FLOATS = GF_RAY + kRayFloats;                        // per-group floats (3904)
constexpr int SF_CONST = 0;
constexpr int SF_GROUP = SF_CONST + scr::kConstFloats;                // 2 x GF_FLOATS
constexpr int SF_VACC = SF_GROUP + kGroups * GF_FLOATS;               // backward: vector-gradient accumulators (588 floats)
constexpr int SF_MISC = SF_VACC + 592;                                // seq[64] bytes, seq_len
constexpr int SF_END = SF_MISC + 32;
constexpr int SMB_F32 = SMB_OPERANDS;
constexpr int SMB_BAR = SMB_F32 + SF_END * 4;                         // mbarriers: per group wfull[2] wfree[2] mma_done; tmem slot
constexpr int kSmemBytes = SMB_BAR + 128;
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB shared memory of one CTA");
enum PtVec2 : int { PX_A = scr::kPtVecs, PX_B = PX_A + 4, PX_C = PX_B + 4, PX_D = PX_C + 4 };

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); }
__device__ __forceinline__ void rays_sync(int g) { asm volatile("bar.sync %0, 64;" ::"r"(3 + g) : "memory"); }   // the 64 point-owner threads

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

// global stash planes (64 x 64 fp32), only ever re-read by the thread that wrote them: float4 index (ch * 4 + q) * 64 + row
constexpr int kStashPlane = MT * 64;
__device__ __forceinline__ void st_store(float* plane, int r, int ch, const float (&v)[NC]) {
    float4* p = reinterpret_cast<float4*>(plane) + (ch * (NC / 4)) * MT + r;
#pragma unroll
    for (int q = 0; q < NC / 4; ++q) __stcg(p + q * MT, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
}
__device__ __forceinline__ void st_load(const float* plane, int r, int ch, float (&v)[NC]) {
    const float4* p = reinterpret_cast<const float4*>(plane) + (ch * (NC / 4)) * MT + r;
#pragma unroll
    for (int q = 0; q < NC / 4; ++q) { const float4 t = __ldcg(p + q * MT); v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w; }
}
// 16 rows x 32 columns per warp: lanes 0-15 columns [c, c+16), lanes 16-31 columns [c+16, c+32) of TMEM lanes base .. base+15
__device__ __forceinline__ void tmem_ld_16x32(uint32_t taddr, float (&v)[NC]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x32bx2.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16], 16;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

#ifdef SC_TC_NOINLINE_ISSUE
#define SC_TC2_ISSUE_FN static __device__ __noinline__
#else
#define SC_TC2_ISSUE_FN __device__ __forceinline__
#endif
// D[64 x 64] (+)= ACT[64 x 64] . W^T   (3 MMAs per 16-wide k-step). Issued by ONE thread of the group.
SC_TC2_ISSUE_FN void issue_layer_gemm(uint32_t tmem_d, const uint8_t* act, const uint8_t* w, bool accumulate) {
    constexpr uint32_t idesc = sctc::make_idesc_bf16(64, 64);
    const uint64_t ah = sctc::make_smem_desc_k128(act), al = sctc::make_smem_desc_k128(act + kPlaneBytes);
    const uint64_t wh = sctc::make_smem_desc_k128(w), wl = sctc::make_smem_desc_k128(w + kWPlaneBytes);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint64_t adv = (uint64_t)(2 * k);
        sctc::umma_bf16(tmem_d, ah + adv, wh + adv, idesc, (accumulate || k > 0) ? 1u : 0u);
        sctc::umma_bf16(tmem_d, ah + adv, wl + adv, idesc, 1u);
        sctc::umma_bf16(tmem_d, al + adv, wh + adv, idesc, 1u);
    }
}
// D[64 x 64] += L^T . R over the tile's 64 points (L, R = plane pairs read MN-major). Always accumulates: the weight-gradient
// accumulators are zeroed at kernel start and shared by both groups.
SC_TC2_ISSUE_FN void issue_wgrad(uint32_t tmem_d, const uint8_t* L, const uint8_t* R) {
    constexpr uint32_t idesc = sct::make_idesc_bf16_mn(64, 64);
    const uint64_t lh = sct::make_smem_desc_mn128(L), ll = sct::make_smem_desc_mn128(L + kPlaneBytes);
    const uint64_t rh = sct::make_smem_desc_mn128(R), rl = sct::make_smem_desc_mn128(R + kPlaneBytes);
#pragma unroll
    for (int k = 0; k < MT / 16; ++k) {
        const uint64_t adv = (uint64_t)(k * (16 * 128 >> 4));          // 16 points = 16 rows of 128 B
        sctc::umma_bf16(tmem_d, lh + adv, rh + adv, idesc, 1u);
        sctc::umma_bf16(tmem_d, lh + adv, rl + adv, idesc, 1u);
        sctc::umma_bf16(tmem_d, ll + adv, rh + adv, idesc, 1u);
    }
}

// ---- per-group weight ring. NS = 2 (forward): TMA-filled slots (wfull), released by tcgen05.commit (wfree), prefetch
// distance 1. NS = 1 (backward, no room for more): the next matrix is requested as soon as the phase's MMAs have completed
// (prefetch(), called by the issuing thread right after the mma_done wait), i.e. it lands under the epilogue; the second
// GEMM of a two-GEMM phase waits for the first one's MMAs to release the slot (the other group fills that gap).
struct WeightRing2 {
    const uint8_t* blob;
    uint8_t* slots;
    uint64_t *wfull, *wfree;      // [2] each
    const int8_t* seq;
    int seq_len;
    int NS;
    uint32_t n;                   // matrices consumed
    uint32_t fetched;             // matrices requested            (issuer thread only)
    uint32_t frees;               // commits on wfree[0] so far (NS == 1; issuer thread only)
    int pos_fetch;

    __device__ __forceinline__ void issue_next() {
        const uint32_t s = fetched % (uint32_t)NS;
        uint64_t* bar = wfull + s;
        scr::mbar_expect_tx(bar, kWSegBytes);
        scr::tma_bulk_g2s(slots + s * kWSegBytes, blob + (size_t)seq[pos_fetch] * kWSegBytes, kWSegBytes, bar);
        pos_fetch = (pos_fetch + 1 == seq_len) ? 0 : pos_fetch + 1;
        ++fetched;
    }
    __device__ __forceinline__ void prologue(bool issuer) {
        n = 0; pos_fetch = 0; fetched = 0; frees = 0;
        if (issuer) issue_next();
    }
    // issuer thread, after the group barrier: returns the slot of matrix n once its bytes have landed
    __device__ __forceinline__ const uint8_t* acquire_issuer() {
        const uint32_t cur = n;
        if (NS == 1 && fetched == cur) {                       // not prefetched: the slot is still being read by the previous GEMM
            if (frees > 0) scr::mbar_wait(wfree, (frees - 1) & 1);
            issue_next();
        }
        scr::mbar_wait(wfull + (cur % NS), (cur / NS) & 1);
        n = cur + 1;
        return slots + (cur % NS) * kWSegBytes;
    }
    // issuer thread, AFTER issuing the MMAs that read the slot of matrix n - 1
    __device__ __forceinline__ void release_issuer() {
        const uint32_t cur = n - 1;
        if (NS == 1) { sctc::umma_commit(wfree); ++frees; return; }
        sctc::umma_commit(wfree + (cur % NS));
        if (cur >= 1) scr::mbar_wait(wfree + ((cur + 1) % NS), (((cur + 1) / NS) - 1) & 1);
        issue_next();                                          // matrix cur + 1 into the slot matrix cur - 1 used
    }
    // issuer thread, NS == 1, right after a wait on mma_done (every MMA of the group has completed: the slot is free)
    __device__ __forceinline__ void prefetch() { if (NS == 1 && fetched == n) issue_next(); }
    __device__ __forceinline__ void drain(bool issuer) {       // outstanding copies must land before the CTA exits
        if (issuer) for (uint32_t m = n; m < fetched; ++m) scr::mbar_wait(wfull + (m % NS), (m / NS) & 1);
    }
};

}  // namespace sct2
