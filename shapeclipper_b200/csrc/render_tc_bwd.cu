// render_tc_bwd.cu — tensor-core edition of the BACKWARD render / SDF-query kernel (see render_tc.cuh, render_bwd.cu).
//
// Per tile: recompute the forward (stashing H, Q, FEAT, R, GPE in the CTA's L2-resident scratch), compositing adjoint,
// RGB backward, second-order sweep, first-order sweep. Every layer GEMM is a tcgen05.mma burst into TMEM accumulator 0/1;
// every weight-gradient GEMM is a UMMA 64x64x16 over the tile's 128 points (both operands = activation planes read
// MN-major) that accumulates into one of 12 TMEM-resident matrices for the WHOLE kernel and is never waited for,
// except before its operand buffers are overwritten. Vector gradients accumulate in shared memory. One flush at the end.
//
// Roles (608 threads = 19 warps: five per scheduler is the most that leaves 96 registers a thread): warps 0-15 = epilogue warps
// (TMEM lane quarter = warp & 3, column group = warp >> 2); warps 16-17 = ray group (saved-activation mode: loads the NEXT tile's saved
// per-point vectors and upstream adjoints and runs its compositing adjoint, 64 points at a time, while the epilogue warps sweep the
// current tile - latency-bound work of 128 threads that used to sit in front of every tile); warp 18 = MMA issuer (interprets the command ring: weights, tcgen05.mma, commits, weight prefetch). With the issue code on warp 0 — also an epilogue warp — every phase waited for warp 0 to get
// through 12-60 MMA issues (blocking on the tensor pipe's queue) AND its own epilogue before the next CTA barrier: clock64 trace,
// 2-4 k cycles of a ~7 k-cycle phase. Now no epilogue warp ever waits for another one inside the sweeps.
#include <cuda_runtime.h>
#include <stdint.h>

#define SC_TC_ROLE_SPLIT 1        // epilogue warps describe GEMMs as commands, a dedicated warp issues them (render_tc_tile.cuh)
#include "render_ray.cuh"
#include "render_tc_tile.cuh"

namespace sct {

using namespace scr;

// recompute = false (mode 0 with saved activations): the tile program starts at the RGB backward
__device__ __forceinline__ void build_seq_tc_bwd(int8_t* seq, int& len, int mode, bool second, bool recompute)
{
    int n = 0;
    if (recompute) {
        const int8_t base[] = {A0N, B1N, A1N, B2N, A2N, W3N, W4N};
        for (int i = 0; i < 7; ++i) seq[n++] = base[i];
        if (mode == 0) { seq[n++] = W5FN; seq[n++] = V0PN; seq[n++] = V0FN; seq[n++] = V1N; seq[n++] = V2N; }
        if (second) { const int8_t g[] = {W4T, W3T, B2T, B1T, A2T, A1T, A0T}; for (int i = 0; i < 7; ++i) seq[n++] = g[i]; }
    }
    if (mode == 0) { seq[n++] = V2T; seq[n++] = V1T; seq[n++] = V0FT; seq[n++] = V0PT; }
    if (second) { const int8_t s2[] = {A0N, A1N, B1N, A2N, B2N, W3N, W4N}; for (int i = 0; i < 7; ++i) seq[n++] = s2[i]; }
    if (mode == 0) seq[n++] = W5FT;
    const int8_t f1[] = {W4T, W3T, A2T, B2T, A1T, B1T, A0T};
    for (int i = 0; i < 7; ++i) seq[n++] = f1[i];
    len = n;
}

// shared-memory vector-gradient accumulators (floats), same relative order as the G_* tail of the folded gradient layout
constexpr int VA_V3 = 0, VA_W5 = 192, VA_B3 = 256, VA_B4 = 320, VA_B5F = 384, VA_C1R = 448, VA_C2R = 512, VA_C3R = 576,
              VA_B5 = 580, VA_BETA = 584, VA_FLOATS = 588;
static_assert(G_W5 - G_V3 == VA_W5 - VA_V3 && G_BETA - G_V3 == VA_BETA - VA_V3, "vector accumulators mirror the G_* layout");

// column sums over the 128 rows of the tile, added to a 64-entry accumulator (16 columns per thread).
// Butterfly transpose-reduce: after the exchange with lane ^ 16 a lane keeps 8 of its 16 columns, then 4, 2, 1 - 16
// shuffles per warp instead of 80, and ONE atomic instruction (16 lanes, 16 addresses) instead of 16.
__device__ __forceinline__ float colsum16(const float (&v)[NC], int lane) {
    float a[8], b[4], c[2];
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = h16 ? v[i] : v[i + 8], keep = h16 ? v[i + 8] : v[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h8 ? a[i] : a[i + 4], keep = h8 ? a[i + 4] : a[i];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h4 ? b[i] : b[i + 2], keep = h4 ? b[i + 2] : b[i];
        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const float send = h2 ? c[0] : c[1], keep = h2 ? c[1] : c[0];
    float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;                      // column (lane >> 1) & 15 of this thread's group, summed over the warp's 32 rows
}
__device__ __forceinline__ void colsum_shared(float* acc64, const float (&v)[NC], int ch, int lane) {
    const float s = colsum16(v, lane);
    if ((lane & 1) == 0) atomicAdd(acc64 + NC * ch + (lane >> 1), s);
}
__device__ __forceinline__ void colsum_global(float* acc64, const float (&v)[NC], int ch, int lane) {
    const float s = colsum16(v, lane);
    if ((lane & 1) == 0) atomicAdd(acc64 + NC * ch + (lane >> 1), s);
}
// stash plane -> operand buffer (same thread mapping)
__device__ __forceinline__ void plane_to_act(const TileTC& T, const float* plane, uint8_t* dst, float (&v)[NC]) {
    st_load(plane, T.row, T.ch, v);
    row_store(dst, T.row, T.ch, v);
}
// pe_bar (this thread's 16 columns of row r) folded into x~_bar: XTB[k % 3][r] += pe_bar_k * d pe_k / d x~
__device__ __forceinline__ void fold_pe_tc(const TileTC& T, const float (&v)[NC]) {
    if (NC * T.ch < NPE) {
        float g[3] = {0.f, 0.f, 0.f};
        pe_fold(T.P(), T.row, T.ch, v, g);
        atomicAdd(T.pv(PV_XTB0) + T.row, g[0]); atomicAdd(T.pv(PV_XTB1) + T.row, g[1]); atomicAdd(T.pv(PV_XTB2) + T.row, g[2]);
    }
}

constexpr int kBwdThreads = 608;         // 16 epilogue warps + 2 ray warps + the issuer
constexpr int kRayWarp0 = 16, kIssuerWarp = 18, kRayThreads = 64;

// What ray_phase_backward / saved_vectors need of a tile, for the ray group: its 64 threads take the tile as two half tiles of 64
// points (whole rays: S <= 64), `half` selects the one in flight.
// Its input vectors (Z', SIG, CF, UN, GX*, NS*, COL*, SDF, W, TMP) are private to the group in saved-activation mode; the eight
// output vectors go to slot `out_shift` (see TileTC::pv). Z' is the group's own copy of the sample depths (the epilogue warps'
// PV_Z belongs to the tile they are sweeping).
struct RayTile {
    float *pt, *ray;
    int tid, lane, warp, S, first, b, rays_per_tile, out_shift, half;
    float beta;
    __device__ __forceinline__ float* pv(int v) const {
        const int u = (v == PV_Z) ? PX_A + 8 : v + ((v >= PV_SDFB && v <= PV_ZB) ? out_shift : 0);
        return pt + u * M_TILE + kRayThreads * half;
    }
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync 3, 64;" ::: "memory"); }
    __device__ __forceinline__ void scan_sync() const { asm volatile("bar.sync 3, 64;" ::: "memory"); }
    __device__ __forceinline__ void mark() const {}
};

// The issuing warp (all 32 lanes; one elected lane executes the tcgen05 instructions): interprets the command batches.
template <int PREC>
__device__ __forceinline__ void issuer_loop(WeightRing& wr, const uint32_t* cmd, uint64_t* mma_done, uint32_t tmem,
                                            uint8_t* act0, uint32_t& wg_mask_out)
{
    uint32_t wg_init = 0;                 // which weight-gradient accumulators already hold data
    wr.prologue();
    for (uint32_t b = 0;; ++b) {
        if (b & 1u) asm volatile("bar.sync 5, 544;" ::: "memory"); else asm volatile("bar.sync 4, 544;" ::: "memory");
        sctc::tc_fence_after();
        const uint32_t* c = cmd + (b & 3u) * kCmdWords;
        const uint32_t n = c[kCmdWords - 1];
        bool end = false;
#pragma unroll 1
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t w = uniform_u32(c[i]);
            const uint32_t op = w & 3u, x = (w >> 2) & 15u;
            const uint8_t* A = act0 + ((w >> 7) & 7u) * kActBytes;
            if (op == TileTC::OP_GEMM) {
                const uint8_t* ws = wr.wait_weights();
                if (PREC == 1) issue_layer_gemm_single(tmem + 16u * x, A, ws, (w >> 6) & 1u);
                else issue_layer_gemm(tmem + 16u * x, A, ws, (w >> 6) & 1u);
                wr.release();
            } else if (op == TileTC::OP_WGRAD) {
                const uint8_t* R = act0 + ((w >> 10) & 7u) * kActBytes;
                if (PREC == 1) issue_wgrad_single(wg_taddr(tmem, (int)x), A, R, (wg_init >> x) & 1u);
                else issue_wgrad(wg_taddr(tmem, (int)x), A, R, (wg_init >> x) & 1u);
                wg_init |= 1u << x;
            } else if (op == TileTC::OP_COMMIT) {
                umma_commit_elect(mma_done);
            } else {
                end = true;
            }
        }
        if (end) break;
    }
    wr.drain_issuer();
    wg_mask_out = wg_init;
}

// PREC = ScRenderArgs::precision, SAVED = mode 0 with the activations saved by sc_render_tc_forward (no recompute): compile-time
template <int MODE, int PREC, bool SAVED>
__global__ void __launch_bounds__(kBwdThreads, 1) render_tc_bwd_kernel(const ScRenderArgs a, float* stash_base)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    int8_t* seq = reinterpret_cast<int8_t*>(reinterpret_cast<float*>(smem + SMB_F32) + SF_MISC);
    int& seq_len = *reinterpret_cast<int*>(reinterpret_cast<float*>(smem + SMB_F32) + SF_MISC + 16);
    float* vacc = reinterpret_cast<float*>(smem + SMB_F32) + SF_VACC;
    uint32_t& wg_mask = *reinterpret_cast<uint32_t*>(reinterpret_cast<float*>(smem + SMB_F32) + SF_MISC + 17);
    const uint8_t* blob = reinterpret_cast<const uint8_t*>(a.blob);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMB_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_TMEM_SLOT);
    uint32_t* cmd = reinterpret_cast<uint32_t*>(smem + SMB_CMD);

    const bool second = (MODE == 0) || (a.want_grad && a.grad_bar != nullptr);
    constexpr bool use_saved = SAVED;
    float* part = a.grad_partial + (size_t)blockIdx.x * kGradFloats;
    for (int i = threadIdx.x; i < kGradFloats; i += kBwdThreads) part[i] = 0.f;
    for (int i = threadIdx.x; i < VA_FLOATS; i += kBwdThreads) vacc[i] = 0.f;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 9; ++i) mbar_init(bars + i, 1);
        mbar_fence_init();
        int len; build_seq_tc_bwd(seq, len, MODE, second, !use_saved); seq_len = len;
    }
    if ((threadIdx.x >> 5) == 0) sctc::tmem_alloc<512>(tmem_slot);
    {
        float* f = reinterpret_cast<float*>(smem + SMB_F32);
        const float* src = reinterpret_cast<const float*>(blob + kTcConstOffsetBytes);
        for (int i = threadIdx.x; i < kConstFloats; i += kBwdThreads) f[SF_CONST + i] = src[i];
        for (int i = threadIdx.x; i < 64; i += kBwdThreads) {
            f[SF_BIAS + 3 * 64 + i] = src[C_B3 + i]; f[SF_BIAS + 4 * 64 + i] = src[C_B4 + i];
            f[SF_BIAS + 6 * 64 + i] = src[C_C1R + i]; f[SF_BIAS + 7 * 64 + i] = src[C_C2R + i];
        }
    }
    sctc::tc_fence_before();
    __syncthreads();
    sctc::tc_fence_after();

    const int per_tile_ = (MODE == 0) ? M_TILE / a.n_samples : M_TILE;
    const int total_ = a.batch * ((a.n_per_image + per_tile_ - 1) / per_tile_);
    const bool active = (int)blockIdx.x < total_;
    const uint32_t tmem_base = *tmem_slot;
    const int warp_id = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);      // warp-uniform for the compiler too

    if (warp_id == kIssuerWarp) {
        // ---------------------------------------------------------------------------------------- the issuer (keeps its 96 registers)
        if (active) {
            WeightRing wr;
            wr.blob = blob; wr.slots = smem + SMB_W_BWD; wr.wfull = bars + BAR_WFULL; wr.wfree = bars + BAR_WFREE;
            wr.seq = seq; wr.seq_len = seq_len; wr.NS = 2; wr.w0 = true;
            uint32_t mask = 0;
            issuer_loop<PREC>(wr, cmd, bars + BAR_MMA_DONE, tmem_base, smem + SMB_ACT, mask);
            if ((threadIdx.x & 31) == 0) wg_mask = mask;
        }
        __syncthreads();                  // (A) the issuer is done: wg_mask is published
        __syncthreads();                  // (B) the epilogue warps have flushed TMEM
        return;
    }
    if (warp_id >= kRayWarp0) {
        // ---------------------------------------------------------------------------------------- the ray group
        // (No setmaxnreg: 19 warps x 96 registers fit and the epilogue warps need no more. Learnt on the way: setmaxnreg.inc can only
        // take what a .dec has RELEASED into the CTA's pool - the SM's unallocated registers are not part of it, asking for more
        // blocks forever; and the register file is per scheduler: a sixth warp on one of them, 21 warps, means 80 registers.)
        if (SAVED && active) {
            RayTile R;
            float* f = reinterpret_cast<float*>(smem + SMB_F32);
            R.pt = f + SF_PT; R.ray = f + SF_RAY;
            R.tid = threadIdx.x - 32 * kRayWarp0; R.lane = threadIdx.x & 31; R.warp = R.tid >> 5;
            R.S = a.n_samples; R.rays_per_tile = kRayThreads / a.n_samples;
            R.beta = fabsf(*a.beta_param) + a.beta_min;
            const int tiles_per_image_ = (a.n_per_image + per_tile_ - 1) / per_tile_;
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < total_; tile += gridDim.x, ++it) {
                const uint32_t slot = it & 1u;
                // hand-over through named barriers (6 + slot: slot free, 8 + slot: slot full; 512 + 64 threads): hardware barrier
                // ordering without the MEMBAR.ALL.CTA of a releasing mbarrier arrive, which would make every epilogue thread wait for
                // its global atomics at the end of each tile
                if (it >= 2) { if (slot) asm volatile("bar.sync 7, 576;" ::: "memory"); else asm volatile("bar.sync 6, 576;" ::: "memory"); }
                R.b = tile / tiles_per_image_;
                R.out_shift = slot ? (PX_A - PV_SDFB) : 0;
                const float* plane = reinterpret_cast<const float*>(a.saved) + ((size_t)tile * TS_SAVED_PLANES + TS_SAVED_PV) * kStashPlane;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    R.half = half;
                    R.first = (tile % tiles_per_image_) * per_tile_ + half * R.rays_per_tile;
                    const int p = R.tid, rr = R.first + p / R.S;
                    R.pv(PV_Z)[p] = (rr < a.n_per_image) ? tc_sample_depth(a, R.b, rr, p % R.S, R.S) : 0.f;
                    saved_vectors<false>(R, const_cast<float*>(plane) + kRayThreads * half);
                    R.sync();
                    ray_phase_backward(R, a, vacc + VA_BETA, kRayThreads);
#pragma unroll
                    for (int c = 0; c < 3; ++c) {                                     // o3_bar = colour_bar * col (1 - col)
                        const float col = R.pv(PV_COL0 + c)[p];
                        const float vv = R.pv(PV_CB0 + c)[p] * col * (1.f - col);
                        R.pv(PV_CB0 + c)[p] = vv;
                        const float sres = warp_sum(vv);
                        if (R.lane == 0) atomicAdd(vacc + VA_C3R + c, sres);
                    }
                }
                if (slot) asm volatile("bar.arrive 9, 576;" ::: "memory"); else asm volatile("bar.arrive 8, 576;" ::: "memory");   // slot full
            }
        }
        __syncthreads();                  // (A)
        __syncthreads();                  // (B)
        return;
    }
    TileTC T;
    {   // same carve-up as the forward kernel, but 5 operand buffers + 2 weight slots
        for (int i = 0; i < kNumAct; ++i) T.act[i] = smem + SMB_ACT + i * kActBytes;
        float* f = reinterpret_cast<float*>(smem + SMB_F32);
        T.cst = f + SF_CONST; T.cb = f + SF_CB; T.pt = f + SF_PT; T.ray = f + SF_RAY; T.bias = f + SF_BIAS;
        T.cmd = cmd; T.batch = 0; T.ncmd = 0; T.out_shift = 0;
        T.mma_done = bars + BAR_MMA_DONE; T.mma_phase = 0;
        T.tid = threadIdx.x; T.lane = threadIdx.x & 31; T.warp = threadIdx.x >> 5;
        T.w0 = false;
        T.wide = false;            // TMEM is full here (12 weight-gradient accumulators)
        T.single = (PREC == 1);
        T.row = 32 * (T.warp & 3) + T.lane; T.ch = T.warp >> 2;
#ifdef SC_TC_TRACE
        T.trace = (MODE == 0) ? reinterpret_cast<long long*>(a.points_bar) : nullptr; T.trace_n = 0;
#endif
    }
    T.tmem = tmem_base;
    T.stash = stash_base + (size_t)blockIdx.x * TS_PLANES_BWD * kStashPlane;
    T.S = (MODE == 0) ? a.n_samples : 1;
    T.rays_per_tile = (MODE == 0) ? M_TILE / a.n_samples : M_TILE;
    T.beta = (MODE == 0) ? fabsf(*a.beta_param) + a.beta_min : 1.f;
    const int tid = T.tid, lane = T.lane, r = T.row, ch = T.ch, c0 = NC * T.ch;
    float* const sc = T.stash;            // per-CTA scratch planes (FB, SB: written and re-read inside one tile)
    float* st = T.stash;                  // activation planes: the same scratch (recompute) or the tile's saved block
    auto wgrad = [&](int m, const uint8_t* L, const uint8_t* R) { T.wgrad(m, L, R); };
    // wait until every MMA issued so far (layer GEMMs and weight gradients) has completed
    auto drain_mma = [&]() { T.commit(); T.wait_mma(); };

    const int per_tile = (MODE == 0) ? T.rays_per_tile : M_TILE;
    const int tiles_per_image = (a.n_per_image + per_tile - 1) / per_tile;
    const int total = a.batch * tiles_per_image;
    const int cb_rows = a.detach_latent ? CB_C0D : CB_C0;
    float v[NC], h[NC], w1[NC], w2[NC];

    if (active) {
        uint32_t tile_it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tile_it) {
            T.b = tile / tiles_per_image;
            T.first = (tile % tiles_per_image) * per_tile;
            float* cbb = a.cb_bar + (size_t)T.b * kCbRows * 64;
            T.sync();
            T.mark();                                                            // [trace] tile start
            tc_tile_setup<MODE>(T, a);
            T.mark();                                                            // [trace] setup done
            if (use_saved) {
                st = reinterpret_cast<float*>(a.saved) + (size_t)tile * TS_SAVED_PLANES * kStashPlane;
                st_load(st + (TS_R + 2) * kStashPlane, r, ch, h);               // r2, r1: one phase ahead
                st_load(st + (TS_R + 1) * kStashPlane, r, ch, w1);
            } else {
                tc_tile_forward<MODE, true>(T, a, second, MODE == 0);
            }

            // ================================================================================ upstream + ray phase
            if (tid < M_TILE) { T.pv(PV_XTB0)[tid] = 0.f; T.pv(PV_XTB1)[tid] = 0.f; T.pv(PV_XTB2)[tid] = 0.f; }
            if (MODE == 1) {
                if (tid < M_TILE) {
                    const int n = T.first + tid;
                    const bool valid = n < a.n_per_image;
                    const size_t g = (size_t)T.b * a.n_per_image + n;
                    T.pv(PV_SDFB)[tid] = (valid && a.sdf_bar) ? a.sdf_bar[g] : 0.f;
#pragma unroll
                    for (int c = 0; c < 3; ++c) T.pv(PV_GXB0 + c)[tid] = (valid && second) ? a.grad_bar[g * 3 + c] : 0.f;
                }
                T.sync();
            } else {
                if (use_saved) {
                    // the ray group prepared this tile's SDFB, GXB*, CB* (already x col (1 - col)), ZB while the previous tile was swept
                    const uint32_t slot = tile_it & 1u;
                    T.out_shift = slot ? (PX_A - PV_SDFB) : 0;
                    if (slot) asm volatile("bar.sync 9, 576;" ::: "memory"); else asm volatile("bar.sync 8, 576;" ::: "memory");
                    T.mark();                                                    // [trace] ray vectors ready
                } else {
                    ray_phase_backward(T, a, vacc + VA_BETA, 512);
                    T.mark();                                                    // [trace] ray phase done

                    // ======================================================================== RGB backward
                    if (tid < M_TILE) {                              // o3_bar = colour_bar * col (1 - col)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const float col = T.pv(PV_COL0 + c)[tid];
                            const float vv = T.pv(PV_CB0 + c)[tid] * col * (1.f - col);
                            T.pv(PV_CB0 + c)[tid] = vv;
                            const float sres = warp_sum(vv);
                            if (lane == 0) atomicAdd(vacc + VA_C3R + c, sres);
                        }
                    }
                    T.sync();
                    T.mark();                                                    // [trace] o3_bar done
                }
                // o2_bar = (V3^T o3_bar) * [r2 > 0] -> Y ; dV3 += o3_bar (x) r2
                // Every activation plane is loaded ONE PHASE AHEAD of its use (into w1, between the submit and the wait of the phase
                // before): a load issued right where the plane is needed costs a full L2/HBM round trip per phase, and nothing hides
                // it — all 16 warps run the same phase.
                if (!use_saved) { st_load(st + (TS_R + 2) * kStashPlane, r, ch, h); st_load(st + (TS_R + 1) * kStashPlane, r, ch, w1); }
                {
                    const float o0 = T.pv(PV_CB0)[r], o1 = T.pv(PV_CB1)[r], o2 = T.pv(PV_CB2)[r];
#pragma unroll
                    for (int i = 0; i < NC; ++i) {
                        const int k = c0 + i;
                        const float t = T.cst[C_V3 + k] * o0 + T.cst[C_V3 + 64 + k] * o1 + T.cst[C_V3 + 128 + k] * o2;
                        v[i] = h[i] > 0.f ? t : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < NC; ++i) w2[i] = o0 * h[i];
                    colsum_shared(vacc + VA_V3, w2, ch, lane);
#pragma unroll
                    for (int i = 0; i < NC; ++i) w2[i] = o1 * h[i];
                    colsum_shared(vacc + VA_V3 + 64, w2, ch, lane);
#pragma unroll
                    for (int i = 0; i < NC; ++i) w2[i] = o2 * h[i];
                    colsum_shared(vacc + VA_V3 + 128, w2, ch, lane);
                }
                row_store(T.Y(), r, ch, v);
                colsum_shared(vacc + VA_C2R, v, ch, lane);
#pragma unroll
                for (int i = 0; i < NC; ++i) h[i] = w1[i];
                row_store(T.Z(), r, ch, h);                                      // r1 -> Z (h keeps this thread's r1 values)
                T.gemm(TM_ACC0, T.Y(), false);                                   // V2T : r1_bar = V2^T o2_bar
                T.commit();                                                      // the epilogue overlaps the weight-gradient MMAs
                wgrad(WG_V2, T.Y(), T.Z());
                T.submit();
                st_load(st + (TS_R + 0) * kStashPlane, r, ch, w1);              // r0, one phase ahead
                T.wait_and_load(TM_ACC0, v);
#pragma unroll
                for (int i = 0; i < NC; ++i) v[i] = h[i] > 0.f ? v[i] : 0.f;
                row_store(T.X(), r, ch, v);                                      // o1_bar -> X
                colsum_shared(vacc + VA_C1R, v, ch, lane);
#pragma unroll
                for (int i = 0; i < NC; ++i) h[i] = w1[i];
                row_store(T.U(), r, ch, h);                                      // r0 -> U
                T.gemm(TM_ACC0, T.X(), false);                                   // V1T
                T.commit();
                wgrad(WG_V1, T.X(), T.U());
                T.submit();
                st_load(st + TS_FEAT * kStashPlane, r, ch, w1);                 // feat, one phase ahead
                T.wait_and_load(TM_ACC0, v);
#pragma unroll
                for (int i = 0; i < NC; ++i) v[i] = h[i] > 0.f ? v[i] : 0.f;
                row_store(T.Z(), r, ch, v);                                      // o0_bar -> Z  (Z: wgrad V2 completed with the V1T phase)
                colsum_global(cbb + CB_RGB * 64, v, ch, lane);
                row_store(T.Y(), r, ch, w1);                                     // feat -> Y
                T.gemm(TM_ACC0, T.Z(), false);                                   // V0FT -> feat_bar
                T.gemm(TM_ACC1, T.Z(), false);                                   // V0PT -> pe_bar (rgb)
                T.commit();
                wgrad(WG_V0F, T.Z(), T.Y()); wgrad(WG_V0P, T.Z(), T.P());
                T.submit();
                if (second) st_load(st + TS_GPE * kStashPlane, r, ch, h);       // gpe, for the second-order prologue
                T.wait_and_load(TM_ACC0, v);
                st_store(sc + TS_FB * kStashPlane, r, ch, v);
                colsum_shared(vacc + VA_B5F, v, ch, lane);
                tmem_ld_32x16(T.tmem + TM_ACC1 + ((uint32_t)(32 * (T.warp & 3)) << 16) + (uint32_t)c0, v);
                fold_pe_tc(T, v);
                T.sync();
            }

            // ================================================================================ second-order sweep
            if (second) {
                // gpe_bar = J (S gx_bar) -> X ; x~_bar += S gx_bar * sum_k d2pe_k gpe_k
                if (MODE == 1) st_load(st + TS_GPE * kStashPlane, r, ch, h);
                {
                    const float gb[3] = {T.pv(PV_GXB0)[r] * T.pv(PV_SGN)[r], T.pv(PV_GXB1)[r], T.pv(PV_GXB2)[r]};
                    float curv[3] = {0.f, 0.f, 0.f};
                    pe_second(T.P(), r, ch, gb, h, v, curv);
                    if (c0 < NPE) {
                        atomicAdd(T.pv(PV_XTB0) + r, gb[0] * curv[0]); atomicAdd(T.pv(PV_XTB1) + r, gb[1] * curv[1]);
                        atomicAdd(T.pv(PV_XTB2) + r, gb[2] * curv[2]);
                    }
                }
                row_store(T.X(), r, ch, v);                                      // X = gpe_bar for the whole sweep
                // layers 0..3: g_l_bar = A_l gpe_bar (+ B_l q_{l-1}_bar) ; q_l_bar = g_l_bar s_l ; SB_l = g_l_bar q_l ; g_l = q_l s_l
                //   buffers: q_bar alternates Y, U, Y, U ; g_l always -> Z
                // Software-pipelined: the GEMMs of layer l + 1 are queued and submitted at the END of layer l's epilogue, then the
                // SB_l store and the loads of H_{l+1}, Q_{l+1} are issued, then the loop waits.
                T.gemm(TM_ACC0, T.X(), false);                                   // A0N
                T.finish();
                st_load(st + (TS_H + 0) * kStashPlane, r, ch, h);
                st_load(st + (TS_Q + 0) * kStashPlane, r, ch, w1);
#pragma unroll 1
                for (int l = 0; l < 4; ++l) {
                    uint8_t* qprev = (l & 1) ? T.Y() : T.U();                    // q_{l-1}_bar (l >= 1)
                    uint8_t* qcur = (l & 1) ? T.U() : T.Y();
                    T.wait_and_load(TM_ACC0, v);
#pragma unroll
                    for (int i = 0; i < NC; ++i) {
                        const float s = sp_slope(h[i]);
                        w2[i] = v[i] * w1[i];            // SB_l
                        h[i] = w1[i] * s;                // g_l
                        v[i] = v[i] * s;                 // q_l_bar
                    }
                    row_store(qcur, r, ch, v); row_store(T.Z(), r, ch, h);
                    if (l < 3) wgrad(l == 0 ? WG_A0 : (l == 1 ? WG_A1 : WG_A2), T.Z(), T.X());
                    if (l == 1) wgrad(WG_B1, T.Z(), qprev);
                    if (l == 2) wgrad(WG_B2, T.Z(), qprev);
                    if (l == 3) wgrad(WG_W3, T.Z(), qprev);
                    // Z (g_l) and qprev are re-written in the next epilogue, i.e. after its wait_and_load, whose commit covers
                    // these weight-gradient MMAs.
                    if (l < 2) { T.gemm(TM_ACC0, T.X(), false); T.gemm(TM_ACC0, qcur, true); }      // A1N, B1N | A2N, B2N
                    else T.gemm(TM_ACC0, qcur, false);                                              // W3N | W4N (q3_bar is in U)
                    T.finish();
                    st_store(sc + (TS_SB + l) * kStashPlane, r, ch, w2);
                    st_load(st + (TS_H + l + 1) * kStashPlane, r, ch, h);
                    if (l < 3) st_load(st + (TS_Q + l + 1) * kStashPlane, r, ch, w1);
                }
                // layer 4 (q4 = w5)
                if (MODE == 0) st_load(sc + TS_FB * kStashPlane, r, ch, w1);    // feat_bar, one phase ahead
                T.wait_and_load(TM_ACC0, v);
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    const float s = sp_slope(h[i]);
                    const float w5 = T.cst[C_W5 + c0 + i];
                    w2[i] = v[i] * w5;                   // SB4 (stays in w2 for the first-order sweep)
                    v[i] = v[i] * s;                     // -> dw5
                    h[i] = w5 * s;                       // g4
                }
                colsum_shared(vacc + VA_W5, v, ch, lane);
                row_store(T.Z(), r, ch, h);
                wgrad(WG_W4, T.Z(), T.U());
            } else if (MODE == 0) {
                st_load(sc + TS_FB * kStashPlane, r, ch, w1);
            }

            // ================================================================================ first-order sweep
            // a4_bar = (w5 sdf_bar + W5f^T feat_bar) s4 + SB4 t4 -> Y
            if (MODE == 0) {
                row_store(T.X(), r, ch, w1);                                     // feat_bar -> X (X = gpe_bar: its readers are done
                T.gemm(TM_ACC0, T.X(), false);                                   //   once the W4N phase above completed)  W5FT
                T.finish();
                st_load(st + (TS_H + 4) * kStashPlane, r, ch, h);
                st_load(st + (TS_H + 3) * kStashPlane, r, ch, w1);              // h3, one phase ahead
                T.wait_and_load(TM_ACC0, v);
            } else {
                T.commit(); T.submit();                                          // mode 1: no GEMM here; retire the weight gradients
                st_load(st + (TS_H + 4) * kStashPlane, r, ch, h);
                st_load(st + (TS_H + 3) * kStashPlane, r, ch, w1);
                T.wait_mma();
#pragma unroll
                for (int i = 0; i < NC; ++i) v[i] = 0.f;
            }
            {
                const float sb = T.pv(PV_SDFB)[r];
                float dw5[NC];
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    float s, t;
                    sp_slope_curv(h[i], s, t);
                    const float hb = v[i] + T.cst[C_W5 + c0 + i] * sb;
                    v[i] = hb * s + (second ? w2[i] * t : 0.f);      // a4_bar
                    dw5[i] = sb * h[i];                               // -> dw5
                }
                if (ch == 0) { const float sres = warp_sum(sb); if (lane == 0) atomicAdd(vacc + VA_B5, sres); }
                colsum_shared(vacc + VA_W5, dw5, ch, lane);
            }
            row_store(T.Y(), r, ch, v);
            colsum_shared(vacc + VA_B4, v, ch, lane);
            if (MODE == 0) {
                row_store(T.Z(), r, ch, h);                                      // h4 -> Z  (Z = g4: weight gradient W4 retired above)
                wgrad(WG_W5F, T.X(), T.Z());
            }
            // layers 3..0: dW_{l+1} += a_{l+1}_bar (x) h_l ; a_l_bar = (W_{l+1}^T a_{l+1}_bar) s_l + SB_l t_l
            //   a_bar alternates Y -> X -> Y -> X -> Y ; h_l is loaded into U / Z alternately
#pragma unroll 1
            for (int l = 3; l >= 0; --l) {
                uint8_t* acur = (l & 1) ? T.Y() : T.X();                         // a_{l+1}_bar
                uint8_t* anew = (l & 1) ? T.X() : T.Y();                         // a_l_bar
                uint8_t* hbuf = (l & 1) ? T.U() : T.Z();
#pragma unroll
                for (int i = 0; i < NC; ++i) h[i] = w1[i];
                row_store(hbuf, r, ch, h);                                       // h_l (h keeps this thread's values)
                if (l < 2) { T.gemm(TM_ACC1, acur, false); }                     // A2T | A1T  -> pe_bar
                T.gemm(TM_ACC0, acur, false);                                    // W4T | W3T | B2T | B1T
                T.commit();
                wgrad(l == 3 ? WG_W4 : (l == 2 ? WG_W3 : (l == 1 ? WG_B2 : WG_B1)), acur, hbuf);
                if (l < 2) wgrad(l == 1 ? WG_A2 : WG_A1, acur, T.P());
                T.submit();
                if (second) st_load(sc + (TS_SB + l) * kStashPlane, r, ch, w2);
                if (l > 0) st_load(st + (TS_H + l - 1) * kStashPlane, r, ch, w1);       // h_{l-1}, one phase ahead
                T.wait_and_load(TM_ACC0, v);
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    float s, t;
                    sp_slope_curv(h[i], s, t);
                    v[i] = v[i] * s + (second ? w2[i] * t : 0.f);
                }
                row_store(anew, r, ch, v);
                if (l == 3) colsum_shared(vacc + VA_B3, v, ch, lane);
                else colsum_global(cbb + (cb_rows + l) * 64, v, ch, lane);
                if (l < 2) {
                    tmem_ld_32x16(T.tmem + TM_ACC1 + ((uint32_t)(32 * (T.warp & 3)) << 16) + (uint32_t)c0, w2);
                    fold_pe_tc(T, w2);
                }
            }
            // a0_bar is in Y: dA0 += a0_bar (x) pe ; pe_bar += A0^T a0_bar
            T.gemm(TM_ACC1, T.Y(), false);                                       // A0T
            wgrad(WG_A0, T.Y(), T.P());
            T.finish_and_load(TM_ACC1, w1);
            fold_pe_tc(T, w1);
            T.sync();
            T.mark();                                                            // [trace] sweeps done

            // ================================================================================ x_bar -> outputs
            if (MODE == 1) {
                if (tid < M_TILE) {
                    const int n = T.first + tid;
                    if (n < a.n_per_image && a.points_bar != nullptr) {
                        const size_t g = ((size_t)T.b * a.n_per_image + n) * 3;
                        a.points_bar[g + 0] = T.pv(PV_XTB0)[tid] * T.pv(PV_SGN)[tid];
                        a.points_bar[g + 1] = T.pv(PV_XTB1)[tid];
                        a.points_bar[g + 2] = T.pv(PV_XTB2)[tid];
                    }
                }
            } else if (tid < M_TILE) {
                const int p = tid, S = T.S, rl = p / S, rr = T.first + rl;
                const bool valid = rr < a.n_per_image;
                const float xb0 = T.pv(PV_XTB0)[p] * T.pv(PV_SGN)[p], xb1 = T.pv(PV_XTB1)[p], xb2 = T.pv(PV_XTB2)[p];
                const float z = T.pv(PV_Z)[p];
                float d0 = 0.f, d1 = 0.f, d2 = 0.f;
                if (valid) {
                    const float* d = a.ray_dirs + ((size_t)T.b * a.n_per_image + rr) * 3;
                    d0 = d[0]; d1 = d[1]; d2 = d[2];
                }
                const float zb = T.pv(PV_ZB)[p] + d0 * xb0 + d1 * xb1 + d2 * xb2;
                float vv[7] = {xb0, xb1, xb2, z * xb0, z * xb1, z * xb2, zb};
                const int seg = S < 32 ? S : 32;
#pragma unroll
                for (int q = 0; q < 7; ++q) vv[q] = seg_sum(vv[q], seg);
                if ((lane & (seg - 1)) == 0 && valid) {
                    float* db = a.ray_dirs_bar + ((size_t)T.b * a.n_per_image + rr) * 3;
                    if (S <= 32) { db[0] = vv[3]; db[1] = vv[4]; db[2] = vv[5]; }
                    else { atomicAdd(db + 0, vv[3]); atomicAdd(db + 1, vv[4]); atomicAdd(db + 2, vv[5]); }
                    atomicAdd(a.cam_loc_bar + T.b * 3 + 0, vv[0]);
                    atomicAdd(a.cam_loc_bar + T.b * 3 + 1, vv[1]);
                    atomicAdd(a.cam_loc_bar + T.b * 3 + 2, vv[2]);
                    atomicAdd(a.scale_dist_bar + T.b, a.cam_dist * vv[6]);
                }
            }
            if (use_saved && tile + 2 * (int)gridDim.x < total) {                        // this slot's vectors may be overwritten
                if (tile_it & 1u) asm volatile("bar.arrive 7, 576;" ::: "memory"); else asm volatile("bar.arrive 6, 576;" ::: "memory");
            }
        }
        T.commit(); T.end();                 // retire every MMA, then let the issuer drain its weight ring and publish wg_mask
        T.wait_mma();
    }
    __syncthreads();                         // (A) all threads
    if (active) {
        // ---- flush: 12 TMEM-resident weight gradients + the shared vector accumulators -> this CTA's partial
        const int goff[NWG] = {G_A0, G_A1, G_A2, G_B1, G_B2, G_W3, G_W4, G_W5F, G_V0P, G_V0F, G_V1, G_V2};
#pragma unroll 1
        for (int pr = 0; pr < NWG / 2; ++pr) {
            tmem_ld_32x16(T.tmem + TM_WGRAD + 64u * (uint32_t)pr + ((uint32_t)(32 * (T.warp & 3)) << 16) + (uint32_t)c0, v);
            const int m = 2 * pr + (lane >> 4);
            if ((wg_mask >> m) & 1u) {
                const int o = 16 * (T.warp & 3) + (lane & 15);
                const bool narrow = (m == WG_A0 || m == WG_A1 || m == WG_A2 || m == WG_V0P);
                const int ld = narrow ? NPE : 64;
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    const int col = c0 + i;
                    if (col < ld) part[goff[m] + o * ld + col] = v[i];
                }
            }
        }
        T.sync();
        for (int i = tid; i < VA_FLOATS; i += 512) part[G_V3 + i] = vacc[i];
    }
    sctc::tc_fence_before();
    __syncthreads();                         // (B)
    if ((threadIdx.x >> 5) == 0) sctc::tmem_dealloc<512>(T.tmem);
}

}  // namespace sct

using namespace sct;

extern "C" int sc_render_tc_backward(const ScRenderArgs* a, cudaStream_t stream)
{
    if (a == nullptr || a->blob == nullptr || a->cb == nullptr || a->scratch == nullptr || a->grad_partial == nullptr ||
        a->cb_bar == nullptr)
        return (int)cudaErrorInvalidValue;
    if (a->mode == 0) {
        const int S = a->n_samples;
        if (S < 4 || S > M_TILE || (S % 4) != 0 || (M_TILE % S) != 0 || a->beta_param == nullptr) return (int)cudaErrorInvalidValue;
        if (!a->ray_dirs_bar || !a->depth_fac_bar || !a->cam_loc_bar || !a->scale_dist_bar) return (int)cudaErrorInvalidValue;
    } else if (a->mode != 1) return (int)cudaErrorInvalidValue;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t err;
#define SC_LAUNCH_BWD(MODE_, PREC_, SAVED_) do { \
        err = cudaFuncSetAttribute(render_tc_bwd_kernel<MODE_, PREC_, SAVED_>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesTc); \
        if (err != cudaSuccess) return (int)err; \
        render_tc_bwd_kernel<MODE_, PREC_, SAVED_><<<sms, sct::kBwdThreads, kSmemBytesTc, stream>>>(*a, (float*)a->scratch); } while (0)
    const bool single = a->precision == 1, saved = a->mode == 0 && a->saved != nullptr && a->n_samples <= 64;      // ray group: whole rays per half tile
    if (a->mode == 0) {
        if (saved) { if (single) SC_LAUNCH_BWD(0, 1, true); else SC_LAUNCH_BWD(0, 0, true); }
        else { if (single) SC_LAUNCH_BWD(0, 1, false); else SC_LAUNCH_BWD(0, 0, false); }
    } else { if (single) SC_LAUNCH_BWD(1, 1, false); else SC_LAUNCH_BWD(1, 0, false); }
#undef SC_LAUNCH_BWD
    return (int)cudaGetLastError();
}
