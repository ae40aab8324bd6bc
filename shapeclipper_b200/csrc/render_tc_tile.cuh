// render_tc_tile.cuh — per-tile FORWARD program of the tensor-core render kernels (shared by forward and backward).
// Algorithm identical to render_tile.cuh / tests/kernel_model.py; see render_tc.cuh for the execution model.
#pragma once
#include "render_tc.cuh"
#include "render_tile.cuh"      // sgnf, density, seg_sum

namespace sct {

struct TileTC {
    uint8_t* act[kNumAct];      // P X Y Z U plane pairs
    float *cst, *cb, *pt, *ray, *bias;      // bias: [8][64] = sdf layers 0..4 (cb0 cb1 cb2 b3 b4), rgb layers 0..2 (cb3 c1r c2r)
    float* stash;
    WeightRing wr;
    uint64_t* mma_done;
    uint32_t mma_phase;
    uint32_t tmem;
    int tid, lane, warp, row, ch;
    bool w0;                    // member of the issuing warp (warp 0, warp-uniform)
    bool wide;                  // forward kernel: 128-column accumulators (issue_layer_gemm_wide); accumulator a lives at columns 2 a
    bool single;                // ScRenderArgs::precision = 1: one MMA per product (hi planes only); never together with `wide`
    int b, first, S, rays_per_tile;
    float beta;
#ifdef SC_TC_TRACE
    long long* trace; int trace_n;
    __device__ __forceinline__ void mark() { if (trace && blockIdx.x == 0 && (tid == 0 || tid == 300) && trace_n < 500) trace[(tid ? 512 : 0) + trace_n++] = clock64(); }
#else
    __device__ __forceinline__ void mark() {}
#endif
#ifdef SC_TC_ROLE_SPLIT
    // the eight vectors the ray group hands over (SDFB, GXB*, CB*, ZB) are double-buffered: slot 1 lives in the PX_* scratch
    int out_shift;              // 0 or PX_A - PV_SDFB
    __device__ __forceinline__ float* pv(int v) const {
        return pt + (v + ((v >= scr::PV_SDFB && v <= scr::PV_ZB) ? out_shift : 0)) * M_TILE;
    }
#else
    __device__ __forceinline__ float* pv(int v) const { return pt + v * M_TILE; }
#endif
    __device__ __forceinline__ void scan_sync() const { asm volatile("bar.sync 1, 128;" ::: "memory"); }
    __device__ __forceinline__ uint8_t* P() const { return act[0]; }
    __device__ __forceinline__ uint8_t* X() const { return act[1]; }
    __device__ __forceinline__ uint8_t* Y() const { return act[2]; }
    __device__ __forceinline__ uint8_t* Z() const { return act[3]; }
    __device__ __forceinline__ uint8_t* U() const { return act[4]; }

#ifndef SC_TC_ROLE_SPLIT
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    // thread 0: commit the phase's layer GEMMs. MMAs issued AFTER this (weight gradients) are covered by the NEXT phase's
    // commit: their operand buffers must stay untouched until the next wait_and_load() has returned.
    __device__ __forceinline__ void commit() { if (w0) umma_commit_elect(mma_done); }
    // one layer GEMM: acquire weights, (thread 0) issue + release
    __device__ __forceinline__ void gemm(uint32_t acc_col, const uint8_t* a, bool accumulate) {
        mark();
        const uint8_t* w = wr.acquire();
        mark();
        if (w0) {
            if (wide) issue_layer_gemm_wide(tmem + 2 * acc_col, a, w, accumulate);
            else if (single) issue_layer_gemm_single(tmem + acc_col, a, w, accumulate);
            else issue_layer_gemm(tmem + acc_col, a, w, accumulate);
            wr.release();
        }
    }
    __device__ __forceinline__ void submit() {}
#else
    // ---- role-split kernels (render_tc_bwd.cu): the 16 epilogue warps never issue an MMA. They describe the work as COMMANDS
    // (thread 0 writes them into a 4-batch ring in shared memory), publish their operand stores (fence.proxy.async + a NON-BLOCKING
    // bar.arrive on named barrier 4 + (batch & 1), which the issuing warp bar.sync's on: hardware barrier ordering without a
    // CTA-wide wait and without the MEMBAR.ALL.CTA a releasing mbarrier.arrive costs — that one waits for every outstanding global
    // load/store of the thread, i.e. for the prefetched saved planes) and go on to wait for the accumulator; a dedicated warp
    // (warp 16) interprets the commands: waits for the weights, issues, commits, prefetches.
    // Command word: bits 0-1 op, 2-5 accumulator column / 16 (GEMM) or weight-gradient matrix (WGRAD), 6 accumulate,
    // 7-9 operand buffer A (act[] index), 10-12 operand buffer B (WGRAD). Word 7 of a batch = its length.
    // EVERY batch ends in work covered by a COMMIT the epilogue warps wait for before they submit the next one, so they are never
    // more than one batch ahead of the issuer: two barrier ids and a ring of four command slots cannot wrap.
    uint32_t* cmd;              // [kCmdBatches][kCmdWords]
    uint32_t batch, ncmd;       // uniform over the epilogue warps
    enum : uint32_t { OP_GEMM = 0, OP_COMMIT = 1, OP_WGRAD = 2, OP_END = 3 };
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync 2, 512;" ::: "memory"); }
    __device__ __forceinline__ uint32_t buf_index(const uint8_t* a) const { return (uint32_t)(a - act[0]) / (uint32_t)kActBytes; }
    __device__ __forceinline__ void push(uint32_t c) { if (tid == 0) cmd[(batch & 3u) * kCmdWords + ncmd] = c; ++ncmd; }
    __device__ __forceinline__ void gemm(uint32_t acc_col, const uint8_t* a, bool accumulate) {
        push(OP_GEMM | ((acc_col >> 4) << 2) | ((accumulate ? 1u : 0u) << 6) | (buf_index(a) << 7));
    }
    __device__ __forceinline__ void commit() { push(OP_COMMIT); }
    __device__ __forceinline__ void wgrad(int m, const uint8_t* L, const uint8_t* R) {
        push(OP_WGRAD | ((uint32_t)m << 2) | (buf_index(L) << 7) | (buf_index(R) << 10));
    }
    __device__ __forceinline__ void end() { push(OP_END); }
    __device__ __forceinline__ void submit() {
        if (ncmd == 0) return;
        if (tid == 0) cmd[(batch & 3u) * kCmdWords + (kCmdWords - 1)] = ncmd;
        sctc::fence_proxy_async();                 // this thread's operand stores -> visible to the tensor core's (async-proxy) reads
        sctc::tc_fence_before();                   // its tcgen05.ld of the accumulator the next MMA may overwrite
        if (batch & 1u) asm volatile("bar.arrive 5, 544;" ::: "memory"); else asm volatile("bar.arrive 4, 544;" ::: "memory");
        ++batch; ncmd = 0;
    }
#endif
    __device__ __forceinline__ void finish() { commit(); submit(); }       // then: prefetch loads, wait_and_load()
    __device__ __forceinline__ void finish_and_load(uint32_t acc_col, float (&v)[NC]) { commit(); wait_and_load(acc_col, v); }
    // everyone: wait for the accumulator, then read this thread's NC columns
    __device__ __forceinline__ void wait_mma() {
        submit();
        mark();
        scr::mbar_wait(mma_done, mma_phase);
        mark();
        mma_phase ^= 1;
        sctc::tc_fence_after();
    }
    __device__ __forceinline__ void wait_and_load(uint32_t acc_col, float (&v)[NC]) {
        wait_mma();
        if (wide) {
            float u[NC];
            const uint32_t base = tmem + 2 * acc_col + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(NC * ch);
            tmem_ld_32x16(base, v);
            tmem_ld_32x16(base + 64, u);
#pragma unroll
            for (int i = 0; i < NC; ++i) v[i] += u[i];
        } else {
            tmem_ld_32x16(tmem + acc_col + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(NC * ch), v);
        }
        mark();
    }
};

// stash planes (floats, row-major [128][64])
constexpr int TS_H = 0, TS_Q = 5, TS_FEAT = 9, TS_R = 10, TS_GPE = 13, TS_FB = 14, TS_SB = 15, TS_PLANES_FWD = 5, TS_PLANES_BWD = 20;
// saved-activation buffer (ScRenderArgs::saved), per tile: planes H0..4, Q0..3, FEAT, R0..2, GPE as above, then one plane of
// per-point vectors [kSavedVecs][128]
constexpr int TS_SAVED_PV = 14, TS_SAVED_PLANES = 15;
constexpr int kSavedVecs = 13;
// forward (store = true): per-point vectors the backward needs -> plane TS_SAVED_PV of the tile's saved block; backward: back.
template <bool STORE, class TT>
__device__ __forceinline__ void saved_vectors(const TT& T, float* plane) {
    constexpr int vecs[kSavedVecs] = {scr::PV_SDF, scr::PV_COL0, scr::PV_COL1, scr::PV_COL2, scr::PV_GX0, scr::PV_GX1, scr::PV_GX2,
                                      scr::PV_SIG, scr::PV_CF, scr::PV_UN, scr::PV_NS0, scr::PV_NS1, scr::PV_NS2};
    if (T.tid < M_TILE) {
#pragma unroll
        for (int i = 0; i < kSavedVecs; ++i) {
            if (STORE) __stcg(plane + i * M_TILE + T.tid, T.pv(vecs[i])[T.tid]);
            else T.pv(vecs[i])[T.tid] = __ldcg(plane + i * M_TILE + T.tid);
        }
    }
}
// ---- posenc derivatives of one row for the 16 columns of column group CH, with compile-time column roles.
// d pe_k / d x~ = +-2^f pe_partner(k) (sin <-> cos, 3 columns away), d2 pe_k / d x~2 = -4^f pe_k. The row's posenc values are read
// back from the operand plane with four 16-byte loads per plane (columns 16 CH - 8 .. 16 CH + 23, conflict-free) instead of
// one 2-byte gather per element with run-time (k - 3) / 6 arithmetic (ncu: 30 % of the backward kernel's instructions).
template <int CH>
__device__ __forceinline__ void pe_row_values(const uint8_t* P, int r, float (&val)[32]) {
    const uint8_t* hi = P + r * 128;
    const uint8_t* lo = hi + kPlaneBytes;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int chunk = 2 * CH - 1 + c;                          // 8 columns each
        if (chunk < 0 || chunk > 4) {                              // columns >= 40 are zero padding
#pragma unroll
            for (int j = 0; j < 8; ++j) val[8 * c + j] = 0.f;
            continue;
        }
        const int pos = (chunk ^ (r & 7)) << 4;
        const uint4 h = *reinterpret_cast<const uint4*>(hi + pos);
        const uint4 l = *reinterpret_cast<const uint4*>(lo + pos);
        const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
        const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&l);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 a = __bfloat1622float2(hp[e]), b = __bfloat1622float2(lp[e]);
            val[8 * c + 2 * e] = a.x + b.x; val[8 * c + 2 * e + 1] = a.y + b.y;
        }
    }
}
// d1[i] = d pe_k / d x~, d2[i] = d2 pe_k / d x~2 for k = 16 CH + i (0 beyond the 39 posenc columns)
template <int CH>
__device__ __forceinline__ void pe_derivs(const uint8_t* P, int r, float (&d1)[NC], float (&d2)[NC]) {
    float val[32];
    pe_row_values<CH>(P, r, val);
    constexpr int base = 8 * (2 * CH - 1);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int k = NC * CH + i;
        if (k >= NPE) { d1[i] = 0.f; d2[i] = 0.f; }
        else if (k < 3) { d1[i] = 1.f; d2[i] = 0.f; }
        else {
            const int f = (k - 3) / 6, rr = (k - 3) % 6;
            const float fr = (float)(1 << f);
            d1[i] = (rr < 3) ? fr * val[k + 3 - base] : -fr * val[k - 3 - base];
            d2[i] = -fr * fr * val[k - base];
        }
    }
}
// g[c] += sum over this thread's columns k with k % 3 == c of v[i] * d pe_k / d x~
template <int CH>
__device__ __forceinline__ void pe_fold_ch(const uint8_t* P, int r, const float (&v)[NC], float (&g)[3]) {
    float d1[NC], d2[NC];
    pe_derivs<CH>(P, r, d1, d2);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int k = NC * CH + i;
        if (k < NPE) g[k % 3] = fmaf(v[i], d1[i], g[k % 3]);
    }
}
__device__ __forceinline__ void pe_fold(const uint8_t* P, int r, int ch, const float (&v)[NC], float (&g)[3]) {
    switch (ch) {                                    // ch = warp >> 2: warp-uniform
        case 0: pe_fold_ch<0>(P, r, v, g); break;
        case 1: pe_fold_ch<1>(P, r, v, g); break;
        case 2: pe_fold_ch<2>(P, r, v, g); break;
        default: break;                              // columns 48..63: no posenc feature
    }
}
// second-order prologue: v[i] = gb[k % 3] * d pe_k ; curv[c] += d2 pe_k * h[i]
template <int CH>
__device__ __forceinline__ void pe_second_ch(const uint8_t* P, int r, const float (&gb)[3], const float (&h)[NC], float (&v)[NC], float (&curv)[3]) {
    float d1[NC], d2[NC];
    pe_derivs<CH>(P, r, d1, d2);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int k = NC * CH + i;
        if (k < NPE) { v[i] = gb[k % 3] * d1[i]; curv[k % 3] = fmaf(d2[i], h[i], curv[k % 3]); }
        else v[i] = 0.f;
    }
}
__device__ __forceinline__ void pe_second(const uint8_t* P, int r, int ch, const float (&gb)[3], const float (&h)[NC], float (&v)[NC], float (&curv)[3]) {
    switch (ch) {
        case 0: pe_second_ch<0>(P, r, gb, h, v, curv); break;
        case 1: pe_second_ch<1>(P, r, gb, h, v, curv); break;
        case 2: pe_second_ch<2>(P, r, gb, h, v, curv); break;
        default:
#pragma unroll
            for (int i = 0; i < NC; ++i) v[i] = 0.f;
            break;
    }
}

// posenc columns 16 CH .. 16 CH + 15 of one point: [x, sin(2^f x), cos(2^f x)]_{f = 0..5}, 39 features then zero padding.
// Accurate sincosf at 2^0 and 2^3, angle doubling (sin 2a = 2 s c, cos 2a = 1 - 2 s^2) for the two octaves after each anchor:
// error growth 4x at most (~2.5e-7), well below the hi/lo-bf16 operand rounding of this path. Octaves a column group does
// not hold are never evaluated.
__host__ __device__ constexpr bool pe_oct_needed(int f, int c, int k0, int k1) {      // does [k0, k1) hold sin or cos of (octave f, coord c)?
    return (3 + 6 * f + c >= k0 && 3 + 6 * f + c < k1) || (6 + 6 * f + c >= k0 && 6 + 6 * f + c < k1);
}
__host__ __device__ constexpr int pe_last_needed(int anchor, int c, int k0, int k1) {  // last needed octave of {anchor .. anchor + 2}, or -1
    return pe_oct_needed(anchor + 2, c, k0, k1) ? anchor + 2 : (pe_oct_needed(anchor + 1, c, k0, k1) ? anchor + 1 :
           (pe_oct_needed(anchor, c, k0, k1) ? anchor : -1));
}
template <int CH>
__device__ __forceinline__ void posenc_cols(const float (&xt)[3], float (&v)[NC]) {
    constexpr int k0 = NC * CH, k1 = k0 + NC;
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = (k0 + i < 3) ? xt[(k0 + i) % 3] : 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float sn = 0.f, cs = 1.f;
#pragma unroll
        for (int f = 0; f < 6; ++f) {
            const int anchor = (f < 3) ? 0 : 3;
            if (f > pe_last_needed(anchor, c, k0, k1)) continue;       // compile-time after unrolling
            if (f == 0) sincosf(xt[c], &sn, &cs);
            else if (f == 3) sincosf(xt[c] * 8.f, &sn, &cs);
            else { const float s2 = 2.f * sn * cs, c2 = 1.f - 2.f * sn * sn; sn = s2; cs = c2; }
            const int ks = 3 + 6 * f + c, kc = ks + 3;
            if (ks >= k0 && ks < k1) v[ks - k0] = sn;
            if (kc >= k0 && kc < k1) v[kc - k0] = cs;
        }
    }
}

// depth of sample s of ray r (image b), exactly as the reference builds it (model/renderer.py:80-96): stratified bins, jitter inside
__device__ __forceinline__ float tc_sample_depth(const ScRenderArgs& a, int b, int r, int s, int S) {
    const float c = __fmul_rn(a.cam_dist, a.scale_dist[b]);
    const float nr = __fsub_rn(c, a.half_range), fr = __fadd_rn(c, a.half_range);
    auto zb = [&](int i) {
        const float t = a.t_vals[i];
        return __fadd_rn(__fmul_rn(nr, __fsub_rn(1.f, t)), __fmul_rn(fr, t));
    };
    float z = zb(s);
    if (a.jitter != nullptr) {
        const float up = (s < S - 1) ? __fmul_rn(0.5f, __fadd_rn(zb(s + 1), z)) : z;
        const float lo = (s > 0) ? __fmul_rn(0.5f, __fadd_rn(z, zb(s - 1))) : z;
        const float u = a.jitter[((size_t)b * a.n_per_image + r) * S + s];
        z = __fadd_rn(lo, __fmul_rn(__fsub_rn(up, lo), u));
    }
    return z;
}

template <int MODE>
__device__ __forceinline__ void tc_tile_setup(const TileTC& T, const ScRenderArgs& a)
{
    for (int i = T.tid; i < 256; i += kThreads) T.cb[i] = a.cb[(size_t)T.b * 256 + i];
    T.sync();
    // every thread (row p, column group ch) derives the sample point of its row; group 0 publishes the per-point vectors
    const int p = T.row;
    float x0, x1, x2, z = 0.f;
    bool valid;
    if (MODE == 0) {
        const int R = a.n_per_image, S = T.S;
        const int r = T.first + p / S, s = p % S;
        valid = r < R;
        if (valid) {
            z = tc_sample_depth(a, T.b, r, s, S);
            const float* d = a.ray_dirs + ((size_t)T.b * R + r) * 3;
            const float* o = a.cam_loc + (size_t)T.b * 3;
            x0 = __fadd_rn(o[0], __fmul_rn(z, d[0]));
            x1 = __fadd_rn(o[1], __fmul_rn(z, d[1]));
            x2 = __fadd_rn(o[2], __fmul_rn(z, d[2]));
        }
    } else {
        const int n = T.first + p;
        valid = n < a.n_per_image;
        if (valid) {
            const float* q = a.points + ((size_t)T.b * a.n_per_image + n) * 3;
            x0 = q[0]; x1 = q[1]; x2 = q[2];
        }
    }
    if (!valid) { x0 = 0.25f; x1 = 0.25f; x2 = 0.25f; z = 0.f; }
    if (T.ch == 0) {
        T.pv(scr::PV_Z)[p] = z;
        T.pv(scr::PV_SGN)[p] = scr::sgnf(x0);
        T.pv(scr::PV_X0)[p] = x0; T.pv(scr::PV_X1)[p] = x1; T.pv(scr::PV_X2)[p] = x2;
    }
    const float xt[3] = {fabsf(x0), x1, x2};
    float v[NC];
    switch (T.ch) {                                   // warp-uniform: column roles are compile-time inside each case
        case 0: posenc_cols<0>(xt, v); break;
        case 1: posenc_cols<1>(xt, v); break;
        case 2: posenc_cols<2>(xt, v); break;
        default:
#pragma unroll
            for (int i = 0; i < NC; ++i) v[i] = 0.f;  // columns 48..63: zero padding
            break;
    }
    row_store(T.P(), p, T.ch, v);
    // per-tile bias tables: sdf layers 0..4 and rgb layers 0..2
    for (int i = T.tid; i < 64; i += kThreads) {
        T.bias[0 * 64 + i] = T.cb[0 * 64 + i]; T.bias[1 * 64 + i] = T.cb[1 * 64 + i]; T.bias[2 * 64 + i] = T.cb[2 * 64 + i];
        T.bias[5 * 64 + i] = T.cb[3 * 64 + i];
    }
}

// Forward program (4 operand buffers: P, X, Y, Z). End state (mode 0): P = posenc, Y = g2, X = g1, Z = g0, per-point vectors SDF, COL*, GX*, SIG, CF, UN, NS*; (STASH_ALL) stash
// planes H0..4, Q0..3, FEAT, R0..2, GPE.
template <int MODE, bool STASH_ALL>
__device__ __forceinline__ void tc_tile_forward(TileTC& T, const ScRenderArgs& a, bool want_grad, bool want_feat)
{
    using namespace scr;
    const int r = T.row, ch = T.ch, c0 = NC * T.ch;
    float v[NC], h[NC];
    float* st = T.stash;

    // ---- F.0 .. F.4: h_l = softplus(W_l [h_{l-1}; pe] + bias_l)      (a REAL loop: one copy of the epilogue code)
#pragma unroll 1
    for (int l = 0; l < 5; ++l) {
        const uint8_t* src = (l == 0) ? T.P() : ((l & 1) ? T.X() : T.Y());
        uint8_t* dst = (l & 1) ? T.Y() : T.X();
        T.gemm(TM_ACC0, src, false);                                                     // A0N | B1N | B2N | W3N | W4N
        if (l == 1 || l == 2) T.gemm(TM_ACC0, T.P(), true);                              // A1N | A2N
        T.finish_and_load(TM_ACC0, v);
        const float* bl = T.bias + l * 64 + c0;
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] = softplus100(v[i] + bl[i]);
        if (l == 4) {
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < NC; ++i) dot = fmaf(T.cst[C_W5 + c0 + i], v[i], dot);
            T.pv(PX_A + ch)[r] = dot;                                                    // sdf partial (4 column groups)
        }
        row_store(dst, r, ch, v); st_store(st + (TS_H + l) * kStashPlane, r, ch, v);     // h4 ends in X
    }
    // ---- F.5 feat
    if (MODE == 0 || want_feat) {
        T.gemm(TM_ACC0, T.X(), false);                                                  // W5FN
        T.finish_and_load(TM_ACC0, v);
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] += T.cst[C_B5F + c0 + i];
        if (MODE == 0) {
            row_store(T.Z(), r, ch, v);
            if (STASH_ALL) st_store(st + TS_FEAT * kStashPlane, r, ch, v);
        } else if (a.feat != nullptr) {
            const int n = T.first + r;
            if (n < a.n_per_image) {
                float4* dstg = reinterpret_cast<float4*>(a.feat + ((size_t)T.b * a.n_per_image + n) * 64 + c0);
#pragma unroll
                for (int q = 0; q < NC / 4; ++q) dstg[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
        }
    }
    if (MODE == 0) {
        // ---- RGB.0 .. RGB.2 (+ the 3-wide output layer as per-thread partial dots)
#pragma unroll 1
        for (int l = 0; l < 3; ++l) {
            const uint8_t* src = (l == 0) ? T.P() : ((l == 1) ? T.Y() : T.X());
            uint8_t* dst = (l == 1) ? T.X() : T.Y();
            T.gemm(TM_ACC0, src, false);                                                 // V0PN | V1N | V2N
            if (l == 0) T.gemm(TM_ACC0, T.Z(), true);                                    // V0FN
            T.finish_and_load(TM_ACC0, v);
            const float* bl = T.bias + (5 + l) * 64 + c0;
#pragma unroll
            for (int i = 0; i < NC; ++i) v[i] = fmaxf(v[i] + bl[i], 0.f);
            if (l == 2) {
                float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    o0 = fmaf(T.cst[C_V3 + c0 + i], v[i], o0);
                    o1 = fmaf(T.cst[C_V3 + 64 + c0 + i], v[i], o1);
                    o2 = fmaf(T.cst[C_V3 + 128 + c0 + i], v[i], o2);
                }
                T.pv(PX_B + ch)[r] = o0; T.pv(PX_C + ch)[r] = o1; T.pv(PX_D + ch)[r] = o2;
            }
            if (l < 2 || STASH_ALL) row_store(dst, r, ch, v);                            // r0 -> Y, r1 -> X, r2 -> Y (backward only)
            if (STASH_ALL) st_store(st + (TS_R + l) * kStashPlane, r, ch, v);
        }
    }
    T.sync();
    if (T.tid < M_TILE) {
        const int p = T.tid;
        auto sum4 = [&](int base) { return T.pv(base)[p] + T.pv(base + 1)[p] + T.pv(base + 2)[p] + T.pv(base + 3)[p]; };
        T.pv(PV_SDF)[p] = T.cst[C_B5] + sum4(PX_A);
        if (MODE == 0) {
            T.pv(PV_COL0)[p] = 1.f / (1.f + expf(-(T.cst[C_C3R + 0] + sum4(PX_B))));
            T.pv(PV_COL1)[p] = 1.f / (1.f + expf(-(T.cst[C_C3R + 1] + sum4(PX_C))));
            T.pv(PV_COL2)[p] = 1.f / (1.f + expf(-(T.cst[C_C3R + 2] + sum4(PX_D))));
        }
    }
    if (MODE == 1 && !want_grad) { T.sync(); return; }

    // ---- gradient pass. Stash reads are issued BEFORE waiting for the MMA so that their L2 latency overlaps it.
    st_load(st + (TS_H + 4) * kStashPlane, r, ch, h);
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = T.cst[C_W5 + c0 + i] * sp_slope(h[i]);
    row_store(T.X(), r, ch, v);                                                          // g4 -> X
#pragma unroll 1
    for (int j = 3; j >= 0; --j) {                                                       // q_j = W_{j+1}^T g_{j+1}; g_j = q_j * s_j
        const uint8_t* src = (j == 3) ? T.X() : ((j == 2) ? T.Z() : ((j == 1) ? T.Y() : T.X()));
        uint8_t* dst = (j == 3) ? T.Z() : ((j == 2) ? T.Y() : ((j == 1) ? T.X() : T.Z()));   // g3->Z g2->Y g1->X g0->Z
        T.gemm(TM_ACC0, src, false);                                                     // W4T | W3T | B2T | B1T
        T.finish();
        st_load(st + (TS_H + j) * kStashPlane, r, ch, h);
        T.wait_and_load(TM_ACC0, v);
        if (STASH_ALL) st_store(st + (TS_Q + j) * kStashPlane, r, ch, v);
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] *= sp_slope(h[i]);
        row_store(dst, r, ch, v);
    }
    // gpe = A2^T g2 + A1^T g1 + A0^T g0  (accumulator 1)
    T.gemm(TM_ACC1, T.Y(), false); T.gemm(TM_ACC1, T.X(), true); T.gemm(TM_ACC1, T.Z(), true);   // A2T, A1T, A0T
    T.finish_and_load(TM_ACC1, v);
    if (STASH_ALL) st_store(st + TS_GPE * kStashPlane, r, ch, v);
    {   // gx~_c = sum_k gpe_k dpe_k : per-thread partial over this thread's columns
        float g[3] = {0.f, 0.f, 0.f};
        pe_fold(T.P(), r, ch, v, g);
        T.pv(PX_A + ch)[r] = g[0]; T.pv(PX_B + ch)[r] = g[1]; T.pv(PX_C + ch)[r] = g[2];
    }
    T.sync();
    if (T.tid < M_TILE) {
        const int p = T.tid;
        auto sum4 = [&](int base) { return T.pv(base)[p] + T.pv(base + 1)[p] + T.pv(base + 2)[p] + T.pv(base + 3)[p]; };
        float g[3] = {sum4(PX_A) * T.pv(PV_SGN)[p], sum4(PX_B), sum4(PX_C)};
        T.pv(PV_GX0)[p] = g[0]; T.pv(PV_GX1)[p] = g[1]; T.pv(PV_GX2)[p] = g[2];
        if (MODE == 0) {
            float sigma, cf, eh;
            density(T.pv(PV_SDF)[p], T.beta, sigma, cf, eh);
            const float u0 = cf * g[0], u1 = cf * g[1], u2 = cf * g[2];
            const float un = sqrtf(u0 * u0 + u1 * u1 + u2 * u2);
            const float inv = 1.f / fmaxf(un, 1e-12f);
            T.pv(PV_SIG)[p] = sigma; T.pv(PV_CF)[p] = cf; T.pv(PV_UN)[p] = un;
            T.pv(PV_NS0)[p] = u0 * inv; T.pv(PV_NS1)[p] = u1 * inv; T.pv(PV_NS2)[p] = u2 * inv;
        }
    }
    T.sync();
}

}  // namespace sct
