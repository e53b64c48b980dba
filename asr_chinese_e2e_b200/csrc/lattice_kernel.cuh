// k2: log-space alpha/beta recursion over the blank-extended label lattice (2U+1 states x T_b frames).
// (replaces aten::_ctc_loss's log-alpha kernel and aten::_ctc_loss_backward's log-beta kernel)
//
// Two warps per utterance.  One runs alpha forward from t=0, the other runs beta backward from
// t=T_b-1, SIMULTANEOUSLY ("meet in the middle"): in phase 1 each warp stores its rows for its half
// of the frames to a global scratch (L2-resident); at the midpoint the two halves are combined once
// to get the log-likelihood; in phase 2 each warp keeps going through the OTHER half, where the
// opposite quantity is already stored, and emits the posterior state occupancies
//     gamma_t(s) = alpha_t(s) * beta_t(s) / (y_t(l'_s) * P)
// on the fly.  Serial depth is T_b steps instead of 2*T_b and no [T,S] array is ever resident.
//
// Lane i owns the NS consecutive states s = NS*i .. NS*i+NS-1 (even = blank, odd = label), so the
// s-1 / s-2 neighbours are registers of the same lane except at the lane boundary, which is one
// warp shuffle (two for beta).  Everything is in log2 units so a logsumexp is
// max + lg2(sum ex2(. - max)) with no multiplies (MUFU ex2/lg2), and log(0) is the FINITE sentinel
// kNeg = -1e30 so that no step needs an inf/NaN guard (see lse2/lse3 below).
//
// The recursion is a single in-order warp per direction, i.e. bound by instruction latency, so the
// step is kept to ~50 instructions: the direction is a template parameter, the 8 frames of a stage
// are unrolled, and the cross-lane sum of the blank occupancies is deferred to once per stage.
// A CTA holds TWO utterances (4 warps) because warp w of a CTA always lands on SM sub-partition
// w % 4: this gives every recursion warp its own issue port and MUFU unit.
//
// Precision: alpha/beta are kept RELATIVE to a running per-direction offset (a double, warp-uniform) that
// absorbs the warp maximum once per stage, so the fp32 state stays within a few hundred log2 units of 0
// instead of growing like T*log2(V) (|alpha| ~ 5000 at T=400): the stored halves, the occupancies and the
// log-likelihood then carry ~30x less rounding error than a plain fp32 log-space recursion (torch's).
// The offset of each stored stage is kept in tile_off[b][stage].
//
// The frames of lp_lab (and, in phase 2, the stored rows of the other direction) are staged in
// shared memory by 1-D bulk TMA copies, TT frames per stage, a private ring of stages per warp,
// completion on mbarriers.
#pragma once
#include "layout.h"
#include "ptx.cuh"
#include "stream_kernels.cuh"

namespace ctcb200 {

constexpr int kLatTT = 8;                 // frames per TMA stage

template <int NS, bool GRAD>
struct LatCfg {
    static constexpr int NL = NS / 2;
    static constexpr int Lp = 4 + 32 * NL;
    static constexpr int Sp = 32 * NS;
    static constexpr int NSTG = NS == 4 ? 3 : 2;           // ring depth per warp (24 / 16 frames ahead)
    static constexpr uint32_t LP_ROW = Lp * 4;
    static constexpr uint32_t AB_ROW = Sp * 4;
    static constexpr uint32_t STAGE = kLatTT * (LP_ROW + (GRAD ? AB_ROW : 0));
    static constexpr uint32_t RING = NSTG * STAGE;
    // xch (per utterance, 32 B): {double ll2, float ll2_rel, double beta offset}; bx (loss only): Sp floats.
    // The shared-memory map of the kernel is K2Smem in lattice_lin.cuh.
};

// log2(2^a + 2^b): 2 MUFU.  With the finite sentinel a-b is never inf-inf.
__device__ __forceinline__ float lse2(float a, float b) {
    return fmaxf(a, b) + lg2f(1.f + ex2f(-fabsf(a - b)));
}
// log2(2^a + 2^b + 2^c): the max term contributes exactly 1, so the sum is >= 1.
__device__ __forceinline__ float lse3(float a, float b, float c) {
    const float m = fmaxf(fmaxf(a, b), c);
    return m + lg2f(ex2f(a - m) + ex2f(b - m) + ex2f(c - m));
}

__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float2 lds_v2(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
template <int N>
__device__ __forceinline__ void lds_vec(float (&d)[N], uint32_t a) {
    if (N == 2) {
        const float2 x = lds_v2(a);
        d[0] = x.x; d[1] = x.y;
    } else {
#pragma unroll
        for (int k = 0; k < N / 4; ++k) {
            const float4 x = lds_v4(a + 16 * k);
            d[4 * k] = x.x; d[4 * k + 1] = x.y; d[4 * k + 2] = x.z; d[4 * k + 3] = x.w;
        }
    }
}
// stores of the lattice's own scratch (stored halves, occupancies): keep them in L2 (evict_last), they are
// read back within microseconds while a [B,T,V] sweep of another chunk may be streaming through the cache
template <int N>
__device__ __forceinline__ void stg_vec(float *p, const float (&d)[N]) {
    if (N == 2) {
        stg_v2_hint((float2 *)p, make_float2(d[0], d[1]), kScratch);
    } else {
#pragma unroll
        for (int k = 0; k < N / 4; ++k)
            stg_v4_hint((float4 *)p + k, make_float4(d[4 * k], d[4 * k + 1], d[4 * k + 2], d[4 * k + 3]), kScratch);
    }
}

// Per-utterance, per-direction state shared by the phases.
template <int NS>
struct LatLane {
    float st[NS];          // alpha_t(.) (DIR 0) or beta_t(.) (DIR 1) of this lane, log2 units
    uint32_t skip;         // bit jj: the s-2 (DIR 0) / s+2 (DIR 1) transition of label jj is allowed
    int lane, Tb, Ub;
};

// One recursion step at frame t; the frame's lp row is at shared address `fa`.
template <int NS, int DIR>
__device__ __forceinline__ void lat_step(LatLane<NS> &L, uint32_t fa, int t, float &lpb, float (&lpl)[NS / 2]) {
    constexpr int NL = NS / 2;
    lpb = lds_f32(fa);
    lds_vec<NL>(lpl, fa + 16 + 4 * NL * L.lane);
    // self-describing frame values (layout.h): v > 0 is a probability, otherwise already a log2-probability
    lpb = lpb > 0.f ? lg2f(lpb) : lpb;
#pragma unroll
    for (int jj = 0; jj < NL; ++jj) lpl[jj] = lpl[jj] > 0.f ? lg2f(lpl[jj]) : lpl[jj];
    float nw[NS];
    if (DIR == 0) {
        if (t == 0) {
#pragma unroll
            for (int j = 0; j < NS; ++j) nw[j] = kNeg;
            if (L.lane == 0) { nw[0] = lpb; nw[1] = lpl[0]; }
        } else {
            float prev = __shfl_up_sync(0xffffffffu, L.st[NS - 1], 1);
            if (L.lane == 0) prev = kNeg;
            nw[0] = lpb + lse2(L.st[0], prev);
#pragma unroll
            for (int jj = 0; jj < NL; ++jj) {
                float s2 = jj == 0 ? prev : L.st[(2 * jj + NS - 1) % NS];
                if (!((L.skip >> jj) & 1)) s2 = kNeg;
                nw[2 * jj + 1] = lpl[jj] + lse3(L.st[2 * jj + 1], L.st[2 * jj], s2);
                if (2 * jj + 2 < NS) nw[(2 * jj + 2) % NS] = lpb + lse2(L.st[(2 * jj + 2) % NS], L.st[2 * jj + 1]);
            }
        }
    } else {
        if (t == L.Tb - 1) {
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const int s = NS * L.lane + j;
                nw[j] = kNeg;
                if (s == 2 * L.Ub) nw[j] = lpb;
                if ((j & 1) && s == 2 * L.Ub - 1) nw[j] = lpl[j >> 1];
            }
        } else {
            float n0 = __shfl_down_sync(0xffffffffu, L.st[0], 1);
            float n1 = __shfl_down_sync(0xffffffffu, L.st[1], 1);
            if (L.lane == 31) { n0 = kNeg; n1 = kNeg; }
#pragma unroll
            for (int jj = 0; jj < NL; ++jj) {
                nw[2 * jj] = lpb + lse2(L.st[2 * jj], L.st[2 * jj + 1]);
                const float s1 = (2 * jj + 2 < NS) ? L.st[(2 * jj + 2) % NS] : n0;
                float s2 = (2 * jj + 3 < NS) ? L.st[(2 * jj + 3) % NS] : n1;
                if (!((L.skip >> jj) & 1)) s2 = kNeg;
                nw[2 * jj + 1] = lpl[jj] + lse3(L.st[2 * jj + 1], s1, s2);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) L.st[j] = nw[j];
}

template <int NS, bool GRAD, int DIR>
__device__ __forceinline__ void lattice_dir(uint32_t ring, uint32_t bar0, uint32_t xch, uint32_t bx, int bar_id,
                                            int lane, int b, int Tb, int Ub,
                                            const int64_t *__restrict__ targets, int64_t tnumel, int64_t toff,
                                            int *__restrict__ flags, const float *__restrict__ lp_lab,
                                            float *__restrict__ gam, float *__restrict__ ab_ws,
                                            float *__restrict__ nll, int T, int zero_inf,
                                            double *__restrict__ tile_off) {
    using C = LatCfg<NS, GRAD>;
    constexpr int NL = C::NL, Lp = C::Lp, Sp = C::Sp, TT = kLatTT, NSTG = C::NSTG;
    // ring / bar0: this warp's private stage ring and its NSTG mbarriers; xch (32 B) / bx (one row): shared by
    // the two warps of the utterance; bar_id: their named barrier

    if (lane == 0) {
        for (int s = 0; s < NSTG; ++s) mbar_init(bar0 + 8 * s, 1);
        fence_mbar_init();
    }
    LatLane<NS> L;
    L.lane = lane; L.Tb = Tb; L.Ub = Ub; L.skip = 0;
    {   // label structure: which skip transitions exist
        int lab[NL];
#pragma unroll
        for (int jj = 0; jj < NL; ++jj) {
            const int li = NL * lane + jj;
            lab[jj] = -1 - li;                                   // unique negative: never equals a neighbour
            if (li < Ub) { const int64_t idx = toff + li; lab[jj] = idx < tnumel ? (int)targets[idx] : 0; }
        }
        int pl = __shfl_up_sync(0xffffffffu, lab[NL - 1], 1);
        int nl = __shfl_down_sync(0xffffffffu, lab[0], 1);
        if (lane == 0) pl = -1000000;
        if (lane == 31) nl = -1000001;
#pragma unroll
        for (int jj = 0; jj < NL; ++jj) {
            const int li = NL * lane + jj;
            if (DIR == 0) {
                const int prev = jj == 0 ? pl : lab[(jj + NL - 1) % NL];
                if (li < Ub && li >= 1 && lab[jj] != prev) L.skip |= 1u << jj;
            } else {
                const int next = jj == NL - 1 ? nl : lab[(jj + 1) % NL];
                if (li + 1 < Ub && lab[jj] != next) L.skip |= 1u << jj;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) L.st[j] = kNeg;
    double off = 0.0;                                            // true value = L.st[j] + off
    // Once per stage the state is shifted by (roughly) its warp maximum and the shift is folded into `off`.
    // Any shift is exact bookkeeping, so the maximum is taken from a snapshot at the START of the stage: its
    // shuffle butterfly is independent of the recursion and is scheduled into the stage's idle issue slots.
    auto snapshot_max = [&]() -> float {
        float m = L.st[0];
#pragma unroll
        for (int j = 1; j < NS; ++j) m = fmaxf(m, L.st[j]);
        return warp_max(m);
    };
    auto renorm = [&](float m) {
        if (m > kNegTest) {
#pragma unroll
            for (int j = 0; j < NS; ++j) L.st[j] -= m;           // the -1e30 sentinel absorbs this
            off += (double)m;
        }
    };

    // ---- tiling of time ----
    const int Qtot = (Tb + TT - 1) / TT;
    int Tm = ((Tb / 2 + TT / 2) / TT) * TT;
    if (Tm >= Tb) Tm = ((Tb - 1) / TT) * TT;
    const int Qm = Tm / TT;                                      // alpha phase 1: tiles [0,Qm); beta phase 1: [Qm,Qtot)
    const int n1 = DIR ? (Qtot - Qm) : Qm;                       // phase-1 jobs of this warp
    const int ntot = GRAD ? Qtot : (DIR ? n1 : Qm + 1);
    const float *lp_base = lp_lab + (size_t)b * T * Lp;
    float *ab_base = ab_ws;                                      // this utterance's part of the scratch
    float *gam_base = gam + (size_t)b * T * Lp;
    double *toff_base = tile_off + (size_t)b * ((T + TT - 1) / TT);

    auto issue = [&](int n) {                                    // lane 0 only
        const int q = DIR ? (Qtot - 1 - n) : n;
        const int t0 = q * TT;
        const int rows = (Tb - t0) < TT ? (Tb - t0) : TT;
        const bool ph2 = GRAD && (DIR ? (q < Qm) : (q >= Qm));
        const int stg = n % NSTG;
        const uint32_t dst = ring + stg * C::STAGE, bar = bar0 + 8 * stg;
        mbar_expect_tx(bar, rows * C::LP_ROW + (ph2 ? rows * C::AB_ROW : 0));
        tma_load_1d_hint(dst, lp_base + (size_t)t0 * Lp, rows * C::LP_ROW, bar, kScratch);
        if (ph2) tma_load_1d_hint(dst + TT * C::LP_ROW, ab_base + (size_t)t0 * Sp, rows * C::AB_ROW, bar, kScratch);
    };
    __syncwarp();

    // ================= phase 1: recursion + store =================
    int n_issue = 0;
    for (; n_issue < NSTG && n_issue < n1; ++n_issue) if (lane == 0) issue(n_issue);
    for (int n = 0; n < n1; ++n) {
        const int stg = n % NSTG;
        mbar_wait(bar0 + 8 * stg, (n / NSTG) & 1);
        const int q = DIR ? (Qtot - 1 - n) : n;
        const int t0 = q * TT;
        const int rows = (Tb - t0) < TT ? (Tb - t0) : TT;
        const uint32_t tile = ring + stg * C::STAGE;
        const float m_snap = snapshot_max();
        if (rows == TT) {                                        // full stage: straight-line, no per-step branch
#pragma unroll
            for (int r = 0; r < TT; ++r) {
                const int rr = DIR ? (TT - 1 - r) : r;
                float lpb, lpl[NL];
                lat_step<NS, DIR>(L, tile + rr * C::LP_ROW, t0 + rr, lpb, lpl);
                if (GRAD) stg_vec<NS>(ab_base + (size_t)(t0 + rr) * Sp + NS * lane, L.st);
            }
        } else {                                                 // the (single) ragged stage at the end of the utterance
#pragma unroll 1
            for (int r = 0; r < rows; ++r) {
                const int rr = DIR ? (rows - 1 - r) : r;
                float lpb, lpl[NL];
                lat_step<NS, DIR>(L, tile + rr * C::LP_ROW, t0 + rr, lpb, lpl);
                if (GRAD) stg_vec<NS>(ab_base + (size_t)(t0 + rr) * Sp + NS * lane, L.st);
            }
        }
        __syncwarp();
        if (n_issue < n1) { if (lane == 0) issue(n_issue); ++n_issue; }
        if (GRAD && lane == 0) toff_base[q] = off;               // offset of the rows just stored
        renorm(m_snap);
    }
    // ================= midpoint =================
    if (GRAD) fence_proxy_async_global();                        // our stored rows -> the other warp's TMA reads
    else if (DIR == 1) {
#pragma unroll
        for (int j = 0; j < NS; ++j)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(bx + 4 * (NS * lane + j)), "f"(L.st[j]) : "memory");
        if (lane == 0) asm volatile("st.shared.f64 [%0], %1;" ::"r"(xch + 16), "d"(off) : "memory");
    }
    named_bar_sync(bar_id, 64);

    // ================= phase 2: recursion + occupancies =================
    for (; n_issue < n1 + NSTG && n_issue < ntot; ++n_issue) if (lane == 0) issue(n_issue);
    double ll2d = 0.0;                                           // log2-likelihood (true value)
    float K = 0.f;                                               // off + other offset - ll2d for the current stage
    bool infeasible = false;
    if (DIR == 1) {                                              // alpha publishes ll2 at its first phase-2 step
        named_bar_sync(bar_id, 64);
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(ll2d) : "r"(xch));
        infeasible = lds_f32(xch + 8) < kNegTest;
    }
    auto other_off = [&](int n) -> double {                      // offset of the other direction's stored stage of job n
        if (!GRAD) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(xch + 16)); return v; }
        return toff_base[DIR ? (Qtot - 1 - n) : n];
    };
    double other = (n1 < ntot && !infeasible) ? other_off(n1) : 0.0;
    int n_waited = n1;                                           // jobs [n_waited, n_issue) are still in flight
    for (int n = n1; n < ntot && !infeasible; ++n) {
        const int stg = n % NSTG;
        mbar_wait(bar0 + 8 * stg, (n / NSTG) & 1);
        n_waited = n + 1;
        const int q = DIR ? (Qtot - 1 - n) : n;
        const int t0 = q * TT;
        const int rows = (Tb - t0) < TT ? (Tb - t0) : TT;
        const uint32_t tile = ring + stg * C::STAGE;
        const double other_cur = other;
        if (n + 1 < ntot) other = other_off(n + 1);              // prefetch the next stage's offset
        const float m_snap = snapshot_max();
        K = (float)(off + other_cur - ll2d);
        // one phase-2 step: recursion, then gamma = 2^(alpha+beta-lp-ll2); returns this lane's blank part
        auto step2 = [&](int rr, bool first) -> float {
            const uint32_t fa = tile + rr * C::LP_ROW;
            float lpb, lpl[NL];
            lat_step<NS, DIR>(L, fa, t0 + rr, lpb, lpl);
            float ot[NS];                                        // the other direction's stored row at t
            lds_vec<NS>(ot, (GRAD ? tile + TT * C::LP_ROW + rr * C::AB_ROW : bx) + 4 * NS * lane);
            if (DIR == 0 && first) {                             // midpoint: log-likelihood (warp-uniform branch)
                float e[NS], m = kNeg * 4.f;
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    e[j] = (L.st[j] + ot[j]) - ((j & 1) ? lpl[j >> 1] : lpb);
                    m = fmaxf(m, e[j]);
                }
                m = warp_max(m);
                float sm = 0.f;
#pragma unroll
                for (int j = 0; j < NS; ++j) sm += ex2f(e[j] - m);
                sm = warp_sum(sm);
                const float ll2_rel = m + lg2f(sm);              // relative to the two offsets
                infeasible = ll2_rel < kNegTest;
                ll2d = off + other_cur + (double)ll2_rel;
                K = -ll2_rel;
                if (lane == 0) {
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(xch), "d"(ll2d) : "memory");
                    asm volatile("st.shared.f32 [%0], %1;" ::"r"(xch + 8), "f"(ll2_rel) : "memory");
                    nll[b] = infeasible ? (zero_inf ? 0.f : __int_as_float(0x7f800000))
                                        : (float)(-ll2d * 0.6931471805599453);
                    flags[b] = infeasible ? 1 : 0;
                }
                named_bar_sync(bar_id, 64);
            }
            float gb = 0.f;
            if (GRAD && !infeasible) {
                const float cb = lpb - K;
                float gl[NL];
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const float c = (j & 1) ? (lpl[j >> 1] - K) : cb;
                    const float g = ex2f((L.st[j] + ot[j]) - c);
                    if (j & 1) gl[j >> 1] = g; else gb += g;
                }
                if (NL * lane < Ub) stg_vec<NL>(gam_base + (size_t)(t0 + rr) * Lp + 4 + NL * lane, gl);
            }
            return gb;
        };
        const bool first_job = (DIR == 0 && n == n1);
        if (rows == TT && !first_job) {
            // full stage: straight-line; the cross-lane sums of the blank occupancies are deferred to
            // TT independent, interleaved butterflies
            float gbl[TT];
#pragma unroll
            for (int r = 0; r < TT; ++r) gbl[r] = step2(DIR ? (TT - 1 - r) : r, false);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int r = 0; r < TT; ++r) gbl[r] += __shfl_xor_sync(0xffffffffu, gbl[r], o);
            }
            float mine = 0.f;
#pragma unroll
            for (int r = 0; r < TT; ++r) if (lane == r) mine = gbl[r];
            if (lane < TT) {                                     // lane r writes the header of the r-th processed row
                const int rr = DIR ? (TT - 1 - lane) : lane;
                stg_v2_hint((float2 *)(gam_base + (size_t)(t0 + rr) * Lp), make_float2(mine, lds_f32(tile + rr * C::LP_ROW + 4)), kScratch);
            }
        } else {
            // ragged stage, or the stage that holds the midpoint
#pragma unroll 1
            for (int r = 0; r < rows; ++r) {
                const int rr = DIR ? (rows - 1 - r) : r;
                float gb = step2(rr, first_job && r == 0);
                if (!GRAD || infeasible) break;
                gb = warp_sum(gb);
                if (lane == 0)
                    stg_v2_hint((float2 *)(gam_base + (size_t)(t0 + rr) * Lp), make_float2(gb, lds_f32(tile + rr * C::LP_ROW + 4)), kScratch);
            }
        }
        if (!GRAD || infeasible) break;
        __syncwarp();
        if (n_issue < ntot) { if (lane == 0) issue(n_issue); ++n_issue; }
        renorm(m_snap);
    }
    // never leave the CTA with bulk copies still landing in its shared memory
    for (int n = n_waited; n < n_issue; ++n) mbar_wait(bar0 + 8 * (n % NSTG), (n / NSTG) & 1);
}

}  // namespace ctcb200
