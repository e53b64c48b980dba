// k2: log-space alpha/beta recursion over the blank-extended label lattice (2U+1 states x T_b frames).
// (replaces aten::_ctc_loss's log-alpha kernel and aten::_ctc_loss_backward's log-beta kernel)
//
// One CTA (2 warps) per utterance.  Warp 0 runs alpha forward from t=0, warp 1 runs beta backward
// from t=T_b-1, SIMULTANEOUSLY ("meet in the middle"): in phase 1 each warp stores its rows for its
// half of the frames to a global scratch (L2-resident); at the midpoint the two halves are combined
// once to get the log-likelihood; in phase 2 each warp keeps going through the OTHER half, where the
// opposite quantity is already stored, and emits the posterior state occupancies
//     gamma_t(s) = alpha_t(s) * beta_t(s) / (y_t(l'_s) * P)
// on the fly.  Serial depth is T_b steps instead of 2*T_b and no [T,S] array is ever resident.
//
// Lane i owns the NS consecutive states s = NS*i .. NS*i+NS-1 (even = blank, odd = label), so the
// s-1 / s-2 neighbours are registers of the same lane except at the lane boundary, which is one
// warp shuffle (two for beta).  Everything is in log2 units so the 3-way logsumexp is
// max + lg2(ex2+ex2+ex2) with no multiplies (MUFU ex2/lg2).
//
// The frames of lp_lab (and, in phase 2, the stored rows of the other direction) are staged in
// shared memory by 1-D bulk TMA copies, TT frames per stage, a private ring of NSTG stages per
// warp, completion on mbarriers.
#pragma once
#include "layout.h"
#include "ptx.cuh"

namespace ctcb200 {

constexpr int kLatTT = 8;     // frames per TMA stage
constexpr int kLatStages = 4; // ring depth per warp

template <int NS, bool GRAD>
struct LatCfg {
    static constexpr int NL = NS / 2;
    static constexpr int Lp = 4 + 32 * NL;
    static constexpr int Sp = 32 * NS;
    static constexpr uint32_t LP_ROW = Lp * 4;
    static constexpr uint32_t AB_ROW = Sp * 4;
    static constexpr uint32_t STAGE = kLatTT * (LP_ROW + (GRAD ? AB_ROW : 0));
    static constexpr uint32_t RING = kLatStages * STAGE;
    // [2 rings][2*NSTG mbarriers][xch: ll2 + pad][bx: Sp floats (loss-only exchange)]
    static constexpr uint32_t OFF_BARS = 2 * RING;
    static constexpr uint32_t OFF_XCH = OFF_BARS + 2 * kLatStages * 8;
    static constexpr uint32_t OFF_BX = OFF_XCH + 16;
    static constexpr uint32_t SMEM = OFF_BX + (GRAD ? 0 : AB_ROW);
};

template <int NS>
__device__ __forceinline__ void load_states(float (&d)[NS], const float *p) {
#pragma unroll
    for (int k = 0; k < NS / 4; ++k) {
        const float4 x = ((const float4 *)p)[k];
        d[4 * k] = x.x; d[4 * k + 1] = x.y; d[4 * k + 2] = x.z; d[4 * k + 3] = x.w;
    }
}
template <int NS>
__device__ __forceinline__ void store_states(float *p, const float (&d)[NS]) {
#pragma unroll
    for (int k = 0; k < NS / 4; ++k)
        ((float4 *)p)[k] = make_float4(d[4 * k], d[4 * k + 1], d[4 * k + 2], d[4 * k + 3]);
}
template <int NL>
__device__ __forceinline__ void load_labels_lp(float (&d)[NL], const float *frame, int lane, uint32_t vmask) {
    const float *p = frame + 4 + NL * lane;
    if (NL == 2) {
        const float2 x = *(const float2 *)p;
        d[0] = x.x; d[1] = x.y;
    } else {
#pragma unroll
        for (int k = 0; k < NL / 4; ++k) {
            const float4 x = ((const float4 *)p)[k];
            d[4 * k] = x.x; d[4 * k + 1] = x.y; d[4 * k + 2] = x.z; d[4 * k + 3] = x.w;
        }
    }
#pragma unroll
    for (int jj = 0; jj < NL; ++jj) if (!((vmask >> jj) & 1)) d[jj] = CTC_NEG_INF;
}

template <int NS, bool GRAD>
__global__ void __launch_bounds__(64)
k2_lattice(const int64_t *__restrict__ targets, int64_t tnumel, const int *__restrict__ Tb_arr,
           const int *__restrict__ Ub_arr, const int64_t *__restrict__ toff_arr, int *__restrict__ flags,
           const float *__restrict__ lp_lab, float *__restrict__ gam, float *__restrict__ ab_ws,
           float *__restrict__ nll, float *__restrict__ loss_sums, unsigned *__restrict__ ticket, int B,
           int T, int zero_inf) {
    using C = LatCfg<NS, GRAD>;
    constexpr int NL = C::NL, Lp = C::Lp, Sp = C::Sp, TT = kLatTT, NSTG = kLatStages;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, dir = tid >> 5;   // dir 0 = alpha, 1 = beta
    const int b = blockIdx.x;
    const int Tb = Tb_arr[b], Ub = Ub_arr[b];
    float *xch = (float *)(smem + C::OFF_XCH);

    if (Tb > 0) {
        const uint32_t ring = smem_u32(smem) + dir * C::RING;
        const unsigned char *ring_p = smem + dir * C::RING;
        const uint32_t bar0 = smem_u32(smem + C::OFF_BARS) + dir * NSTG * 8;
        if (lane == 0) {
            for (int s = 0; s < NSTG; ++s) mbar_init(bar0 + 8 * s, 1);
            fence_mbar_init();
        }
        // ---- per-lane label structure ----
        const int64_t toff = toff_arr[b];
        int lab[NL];
        uint32_t vmask = 0, fskip = 0, bskip = 0;
#pragma unroll
        for (int jj = 0; jj < NL; ++jj) {
            const int li = NL * lane + jj;
            lab[jj] = -1 - li;                               // unique negative: never equal to a neighbour
            if (li < Ub) {
                const int64_t idx = toff + li;
                lab[jj] = idx < tnumel ? (int)targets[idx] : 0;
                vmask |= 1u << jj;
            }
        }
        {
            int pl = __shfl_up_sync(0xffffffffu, lab[NL - 1], 1);
            int nl = __shfl_down_sync(0xffffffffu, lab[0], 1);
            if (lane == 0) pl = -1000000;
            if (lane == 31) nl = -1000001;
#pragma unroll
            for (int jj = 0; jj < NL; ++jj) {
                const int li = NL * lane + jj;
                const int prev = jj == 0 ? pl : lab[jj - 1];
                const int next = jj == NL - 1 ? nl : lab[jj + 1];
                if (li < Ub && li >= 1 && lab[jj] != prev) fskip |= 1u << jj;
                if (li + 1 < Ub && lab[jj] != next) bskip |= 1u << jj;
            }
        }
        // ---- tiling of time ----
        const int Qtot = (Tb + TT - 1) / TT;
        int Tm = ((Tb / 2 + TT / 2) / TT) * TT;
        if (Tm >= Tb) Tm = ((Tb - 1) / TT) * TT;
        const int Qm = Tm / TT;                              // alpha phase 1: tiles [0,Qm); beta phase 1: [Qm,Qtot)
        const int n1 = dir ? (Qtot - Qm) : Qm;               // phase-1 jobs of this warp
        const int ntot = (GRAD ? Qtot : (dir ? n1 : Qm + 1));
        const float *lp_base = lp_lab + (size_t)b * T * Lp;
        float *ab_base = ab_ws + (size_t)b * T * Sp;
        float *gam_base = gam + (size_t)b * T * Lp;

        auto issue = [&](int n) {   // lane 0 only
            const int q = dir ? (Qtot - 1 - n) : n;
            const int t0 = q * TT;
            const int rows = (Tb - t0) < TT ? (Tb - t0) : TT;
            const bool ph2 = GRAD && (dir ? (q < Qm) : (q >= Qm));
            const int stg = n % NSTG;
            const uint32_t dst = ring + stg * C::STAGE, bar = bar0 + 8 * stg;
            mbar_expect_tx(bar, rows * C::LP_ROW + (ph2 ? rows * C::AB_ROW : 0));
            tma_load_1d(dst, lp_base + (size_t)t0 * Lp, rows * C::LP_ROW, bar);
            if (ph2) tma_load_1d(dst + TT * C::LP_ROW, ab_base + (size_t)t0 * Sp, rows * C::AB_ROW, bar);
        };
        __syncwarp();

        float st[NS];                                        // alpha_t(.) or beta_t(.) of this lane, log2 units
#pragma unroll
        for (int j = 0; j < NS; ++j) st[j] = CTC_NEG_INF;

        // one recursion step at frame t given that frame's lp frame in shared memory
        auto step = [&](const float *frame, int t, float &lpb, float (&lpl)[NL]) {
            lpb = frame[0];
            load_labels_lp<NL>(lpl, frame, lane, vmask);
            float nw[NS];
            if (dir == 0) {
                if (t == 0) {
#pragma unroll
                    for (int j = 0; j < NS; ++j) nw[j] = CTC_NEG_INF;
                    if (lane == 0) { nw[0] = lpb; nw[1] = lpl[0]; }
                } else {
                    float prev = __shfl_up_sync(0xffffffffu, st[NS - 1], 1);
                    if (lane == 0) prev = CTC_NEG_INF;
                    nw[0] = lpb + lse2_2(st[0], prev);
#pragma unroll
                    for (int jj = 0; jj < NL; ++jj) {
                        const float s1 = st[2 * jj];
                        float s2 = jj == 0 ? prev : st[2 * jj - 1];
                        if (!((fskip >> jj) & 1)) s2 = CTC_NEG_INF;
                        nw[2 * jj + 1] = lpl[jj] + lse3_2(st[2 * jj + 1], s1, s2);
                        if (2 * jj + 2 < NS) nw[2 * jj + 2] = lpb + lse2_2(st[2 * jj + 2], st[2 * jj + 1]);
                    }
                }
            } else {
                if (t == Tb - 1) {
#pragma unroll
                    for (int j = 0; j < NS; ++j) {
                        const int s = NS * lane + j;
                        nw[j] = CTC_NEG_INF;
                        if (s == 2 * Ub) nw[j] = lpb;
                        if ((j & 1) && s == 2 * Ub - 1) nw[j] = lpl[j >> 1];
                    }
                } else {
                    float n0 = __shfl_down_sync(0xffffffffu, st[0], 1);
                    float n1v = __shfl_down_sync(0xffffffffu, st[1], 1);
                    if (lane == 31) { n0 = CTC_NEG_INF; n1v = CTC_NEG_INF; }
#pragma unroll
                    for (int jj = 0; jj < NL; ++jj) {
                        // blank state 2jj: successors 2jj, 2jj+1
                        nw[2 * jj] = lpb + lse2_2(st[2 * jj], st[2 * jj + 1]);
                        // label state 2jj+1: successors 2jj+1, 2jj+2, (2jj+3)
                        const float s1 = (2 * jj + 2 < NS) ? st[(2 * jj + 2) % NS] : n0;
                        float s2 = (2 * jj + 3 < NS) ? st[(2 * jj + 3) % NS] : n1v;
                        if (!((bskip >> jj) & 1)) s2 = CTC_NEG_INF;
                        nw[2 * jj + 1] = lpl[jj] + lse3_2(st[2 * jj + 1], s1, s2);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < NS; ++j) st[j] = nw[j];
        };

        // ================= phase 1 =================
        int n_issue = 0;
        for (; n_issue < NSTG && n_issue < n1; ++n_issue) if (lane == 0) issue(n_issue);
        for (int n = 0; n < n1; ++n) {
            const int stg = n % NSTG;
            mbar_wait(bar0 + 8 * stg, (n / NSTG) & 1);
            const int q = dir ? (Qtot - 1 - n) : n;
            const int t0 = q * TT;
            const int rows = (Tb - t0) < TT ? (Tb - t0) : TT;
            const unsigned char *tile = ring_p + stg * C::STAGE;
            for (int r = 0; r < rows; ++r) {
                const int rr = dir ? (rows - 1 - r) : r;
                const int t = t0 + rr;
                float lpb, lpl[NL];
                step((const float *)(tile + rr * C::LP_ROW), t, lpb, lpl);
                if (GRAD) store_states<NS>(ab_base + (size_t)t * Sp + NS * lane, st);
            }
            __syncwarp();
            if (n_issue < n1) { if (lane == 0) issue(n_issue); ++n_issue; }
        }
        // ================= midpoint =================
        if (GRAD) fence_proxy_async_global();                // our stored rows -> the other warp's TMA reads
        else if (dir == 1) store_states<NS>((float *)(smem + C::OFF_BX) + NS * lane, st);
        named_bar_sync(1, 64);

        // ================= phase 2 =================
        for (; n_issue < n1 + NSTG && n_issue < ntot; ++n_issue) if (lane == 0) issue(n_issue);
        float ll2 = 0.f;
        if (dir == 1) { named_bar_sync(1, 64); ll2 = xch[0]; }   // alpha publishes ll2 at its first step
        bool infeasible = (dir == 1) && (ll2 == CTC_NEG_INF);
        int n_waited = n1;                                   // jobs [n_waited, n_issue) are still in flight
        for (int n = n1; n < ntot && !infeasible; ++n) {
            const int stg = n % NSTG;
            mbar_wait(bar0 + 8 * stg, (n / NSTG) & 1);
            n_waited = n + 1;
            const int q = dir ? (Qtot - 1 - n) : n;
            const int t0 = q * TT;
            const int rows = (Tb - t0) < TT ? (Tb - t0) : TT;
            const unsigned char *tile = ring_p + stg * C::STAGE;
            for (int r = 0; r < rows; ++r) {
                const int rr = dir ? (rows - 1 - r) : r;
                const int t = t0 + rr;
                const float *frame = (const float *)(tile + rr * C::LP_ROW);
                float lpb, lpl[NL];
                step(frame, t, lpb, lpl);
                float ot[NS];                                // the other direction's stored row at t
                if (GRAD) load_states<NS>(ot, (const float *)(tile + TT * C::LP_ROW + rr * C::AB_ROW) + NS * lane);
                else load_states<NS>(ot, (const float *)(smem + C::OFF_BX) + NS * lane);
                float e[NS];                                 // log2 of alpha*beta/y, unnormalised
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    float lp = lpb;
                    if (j & 1) lp = ((vmask >> (j >> 1)) & 1) ? lpl[j >> 1] : 0.f;
                    e[j] = (st[j] + ot[j]) - lp;
                }
                if (dir == 0 && n == n1 && r == 0) {         // midpoint: log-likelihood
                    float m = e[0];
#pragma unroll
                    for (int j = 1; j < NS; ++j) m = fmaxf(m, e[j]);
                    m = warp_max(m);
                    const float ms = (m == CTC_NEG_INF) ? 0.f : m;
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < NS; ++j) s += ex2f(e[j] - ms);
                    s = warp_sum(s);
                    ll2 = ms + lg2f(s);
                    infeasible = (ll2 == CTC_NEG_INF);
                    if (lane == 0) {
                        xch[0] = ll2;
                        nll[b] = infeasible ? (zero_inf ? 0.f : __int_as_float(0x7f800000)) : -ll2 * kLn2;
                        flags[b] = infeasible ? 1 : 0;
                    }
                    named_bar_sync(1, 64);
                    if (!GRAD || infeasible) break;
                }
                if (GRAD) {
                    float gb = 0.f, gl[NL];
#pragma unroll
                    for (int j = 0; j < NS; ++j) {
                        const float gmm = ex2f(e[j] - ll2);
                        if (j & 1) gl[j >> 1] = gmm; else gb += gmm;
                    }
                    gb = warp_sum(gb);
                    float *gf = gam_base + (size_t)t * Lp;
                    if (lane == 0) *(float2 *)gf = make_float2(gb, frame[1]);
                    if (NL * lane < Ub) {
                        if (NL == 2) *(float2 *)(gf + 4 + NL * lane) = make_float2(gl[0], gl[1]);
                        else {
#pragma unroll
                            for (int k = 0; k < NL / 4; ++k)
                                ((float4 *)(gf + 4 + NL * lane))[k] =
                                    make_float4(gl[4 * k], gl[4 * k + 1], gl[4 * k + 2], gl[4 * k + 3]);
                        }
                    }
                }
            }
            if (!GRAD || infeasible) break;
            __syncwarp();
            if (n_issue < ntot) { if (lane == 0) issue(n_issue); ++n_issue; }
        }
        // never leave the CTA with bulk copies still landing in its shared memory
        for (int n = n_waited; n < n_issue; ++n) mbar_wait(bar0 + 8 * (n % NSTG), (n / NSTG) & 1);
    } else if (tid == 0) {
        // no frames: empty target -> probability 1, anything else is infeasible (torch: inf, zero grad)
        nll[b] = (Ub == 0) ? 0.f : (zero_inf ? 0.f : __int_as_float(0x7f800000));
        flags[b] = (Ub != 0);
    }

    // ---- deterministic batch reduction by the last CTA to finish ----
    if (loss_sums != nullptr && dir == 0) {
        unsigned tk = 0;
        if (lane == 0) { __threadfence(); tk = atomicAdd(ticket, 1u); }
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk == (unsigned)B - 1) {
            __threadfence();
            float s_norm = 0.f, s_sum = 0.f;
            for (int i = lane; i < B; i += 32) {
                const float v = __ldcg(nll + i);
                const int u = Ub_arr[i];
                s_sum += v;
                s_norm += v / (float)(u > 1 ? u : 1);
            }
            s_norm = warp_sum(s_norm);
            s_sum = warp_sum(s_sum);
            if (lane == 0) { loss_sums[0] = s_norm; loss_sums[1] = s_sum; loss_sums[2] = (float)B; *ticket = 0; }
        }
    }
}

}  // namespace ctcb200
