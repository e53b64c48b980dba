// k2, fast path: the alpha/beta recursion in the LINEAR domain (probabilities, not logs), in float64 with a
// per-lane block exponent.  Same decomposition as lattice_kernel.cuh (two warps per utterance meeting in the
// middle, lane i owns NS consecutive states, TMA-staged frames, occupancies emitted on the fly in phase 2),
// but a recursion step is 2 adds + 1 multiply per state on the FP64 pipe instead of a log-sum-exp
// (4 MUFU per state): the step of the single in-order warp drops from ~370 to ~60 cycles (NS = 4).
//
// Representation: the true alpha_t(s) (or beta_t(s)) of state j of lane l is  m[j] * 2^E_l  with m a double
// and E_l an int that is constant within a stage of TT frames.  At every stage boundary each lane shifts its
// own maximum to [1,2) (an exact power-of-two scaling) and folds the shift into E_l; lanes that are still
// all-zero adopt the exponent of the nearest live lane on the side the mass comes from.  The value crossing
// a lane boundary is rescaled by f = 2^(E_neighbour - E_l) (one DMUL).  A single warp-wide scale would not
// do: while a model still predicts mostly blank, alpha spreads over hundreds of binary orders of magnitude
// ACROSS the states of one frame (2^-30 per label), far more than any float format holds; within one lane's
// NS states plus TT frames of decay the spread stays inside the range of a double as long as no gathered
// log-probability is below `lin_thr` (about -75..-90 log2 units).  The sweep kernel checks exactly that
// while it gathers (slow[b] = 1 otherwise) and such utterances -- and any whose likelihood still underflows
// to 0, which includes the truly infeasible ones -- run the log-space recursion of lattice_kernel.cuh
// instead, so the result is defined for every input.
//
// The emission probabilities come from the sweep kernel, which stores a gathered value as the float probability
// 2^lp itself whenever it is in range (layout.h; representable because the linear path only runs when all of
// them are >= 2^-90); a lane converts its NS/2+1 values of a frame to double (exact), off the dependent chain.
// Relative error 2^-22 per factor, i.e. ~1e-9 relative on the log-likelihood -- far inside what an fp32
// log-space recursion gives.
//
// What is stored for the other direction's phase 2 is the PRE-emission sum  a^_t(s) = sum of predecessors
// (alpha_t = a^_t * y_t), so the occupancy is  gamma_t(s) = alpha_t(s) * b^_t(s) / P  with no division by
// y_t.  A stored stage is one contiguous block {32 lane exponents, rows of doubles for the lanes that hold
// real states}, fetched by one bulk copy.
#pragma once
#include "lattice_kernel.cuh"

namespace ctcb200 {

template <int NS, bool GRAD>
struct LinCfg {
    static constexpr int NL = NS / 2;
    static constexpr int Lp = 4 + 32 * NL;
    static constexpr int TT = NS == 16 ? 4 : 8;             // frames per stage = renormalisation interval
    static constexpr int NSTG = NS == 4 ? 3 : 2;
    static constexpr uint32_t LP_ROW = Lp * 4;
    static constexpr uint32_t AB_ROW = 32 * NS * 8;         // widest stored row
    static constexpr uint32_t EXPS = 128;                   // 32 lane exponents
    static constexpr uint32_t STAGE = TT * LP_ROW + (GRAD ? EXPS + TT * AB_ROW : 0);
    static constexpr uint32_t RING = NSTG * STAGE;
    // xch (per utterance, 32 B): {double 1/Lm, int EL, int status}; bx (loss only): {exps, one row}
};

// Shared-memory map of k2_lattice.  Every warp owns one ring (wide enough for either recursion) and two sets of
// mbarriers (one per recursion: an utterance that falls back to log space must not re-initialise live barriers,
// and the other utterance of the CTA may still be in the linear recursion, so nothing is shared across pairs).
template <int NS, bool GRAD>
struct K2Smem {
    using Lin = LinCfg<NS, GRAD>;
    using Log = LatCfg<NS, GRAD>;
    static constexpr int NSTG = Lin::NSTG;
    static_assert(Lin::NSTG == Log::NSTG, "ring depth");
    static constexpr uint32_t RING = Lin::RING > Log::RING ? Lin::RING : Log::RING;
    static constexpr uint32_t OFF_BARS_LIN = 4 * RING;
    static constexpr uint32_t OFF_BARS_LOG = OFF_BARS_LIN + 4 * NSTG * 8;
    static constexpr uint32_t OFF_XCH = OFF_BARS_LOG + 4 * NSTG * 8;
    static constexpr uint32_t BX = Lin::EXPS + Lin::AB_ROW;
    static constexpr uint32_t OFF_BX = OFF_XCH + 64;
    static constexpr uint32_t SMEM = OFF_BX + (GRAD ? 0 : 2 * BX);
};

__device__ __forceinline__ double pow2i(int d) {            // 2^d, d in [-1022, 1023]
    return __hiloint2double((1023 + d) << 20, 0);
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int N>
__device__ __forceinline__ void lds_vec_d(double (&d)[N], uint32_t a) {
#pragma unroll
    for (int k = 0; k < N / 2; ++k)
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(d[2 * k]), "=d"(d[2 * k + 1]) : "r"(a + 16 * k));
}

template <int NS>
struct LinLane {
    double m[NS];          // mantissas of alpha_t(.) / beta_t(.) of this lane
    double sum[NS];        // the pre-emission sums of the last step
    double sk[NS / 2];     // 1.0 where the s-2 (DIR 0) / s+2 (DIR 1) transition of label jj exists, else 0.0
    double f;              // 2^(E_neighbour - E): scale of the value that crosses the lane boundary
    int E;                 // block exponent of this lane
    int lane, Tb, Ub;
};

// The emission probabilities of one frame for this lane's states.  Loading and converting them does not
// depend on the recursion, so a whole stage is converted up front (lin_load_p x TT) and the compiler is free to
// slot that work into the latency gaps of the dependent chain (lin_chain).
template <int NS>
struct LinP {
    double pb;             // blank
    double pl[NS / 2];     // this lane's labels
};
template <int NS>
__device__ __forceinline__ void lin_load_p(LinP<NS> &P, uint32_t fa, int lane) {
    constexpr int NL = NS / 2;
    float pl[NL];
    const float pb = lds_f32(fa);
    lds_vec<NL>(pl, fa + 16 + 4 * NL * lane);
    // exact conversions: the sweep stored 2^lp as a float (>= 2^-90 on this path); the only non-positive values
    // here are the sentinels of unused label slots -> probability 0
    P.pb = (double)fmaxf(pb, 0.f);
#pragma unroll
    for (int jj = 0; jj < NL; ++jj) P.pl[jj] = (double)fmaxf(pl[jj], 0.f);
}

// One recursion step.  INIT: the first step of the direction (t = 0 for alpha, t = T_b - 1 for beta).
template <int NS, int DIR, bool INIT>
__device__ __forceinline__ void lin_chain(LinLane<NS> &L, const LinP<NS> &P) {
    constexpr int NL = NS / 2;
    double (&m)[NS] = L.m;
    double (&s)[NS] = L.sum;
    if (INIT) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const int st = NS * L.lane + j;
            if (DIR == 0) s[j] = st < 2 ? 1.0 : 0.0;
            else s[j] = (st == 2 * L.Ub || ((j & 1) && st == 2 * L.Ub - 1)) ? 1.0 : 0.0;
        }
    } else if (DIR == 0) {
        const double prev = __shfl_up_sync(0xffffffffu, m[NS - 1], 1) * L.f;
        s[0] = m[0] + prev;
#pragma unroll
        for (int jj = 0; jj < NL; ++jj) {
            const double s2 = jj == 0 ? prev : m[(2 * jj + NS - 1) % NS];
            s[2 * jj + 1] = fma(L.sk[jj], s2, m[2 * jj + 1] + m[2 * jj]);
            if (2 * jj + 2 < NS) s[(2 * jj + 2) % NS] = m[(2 * jj + 2) % NS] + m[2 * jj + 1];
        }
    } else {
        const double n0 = __shfl_down_sync(0xffffffffu, m[0], 1) * L.f;
        const double n1 = __shfl_down_sync(0xffffffffu, m[1], 1) * L.f;
#pragma unroll
        for (int jj = 0; jj < NL; ++jj) {
            s[2 * jj] = m[2 * jj] + m[2 * jj + 1];
            const double s1 = (2 * jj + 2 < NS) ? m[(2 * jj + 2) % NS] : n0;
            const double s2 = (2 * jj + 3 < NS) ? m[(2 * jj + 3) % NS] : n1;
            s[2 * jj + 1] = fma(L.sk[jj], s2, m[2 * jj + 1] + s1);
        }
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) m[j] = s[j] * ((j & 1) ? P.pl[j >> 1] : P.pb);
}

// predicated stores (no branch in the recursion loop)
__device__ __forceinline__ void stg_v2f64_hint_if(int pred, double *p, double a, double b, uint64_t policy) {
    asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %4, 0;\n @q st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;\n}"
                 ::"l"(p), "d"(a), "d"(b), "l"(policy), "r"(pred) : "memory");
}
template <int N>
__device__ __forceinline__ void stg_vec_if(int pred, float *p, const float (&d)[N]) {
    if (N == 2) {
        asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %4, 0;\n @q st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;\n}"
                     ::"l"(p), "f"(d[0]), "f"(d[1]), "l"(kScratch), "r"(pred) : "memory");
    } else {
#pragma unroll
        for (int k = 0; k < N / 4; ++k)
            asm volatile("{\n .reg .pred q;\n setp.ne.s32 q, %6, 0;\n @q st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;\n}"
                         ::"l"(p + 4 * k), "f"(d[4 * k]), "f"(d[4 * k + 1]), "f"(d[4 * k + 2]), "f"(d[4 * k + 3]),
                         "l"(kScratch), "r"(pred) : "memory");
    }
}

// Returns true when the utterance is finished (nll / flags / occupancies written); false when the likelihood
// underflowed to zero and the caller must redo the utterance in log space.  Both warps of the pair return the
// same value.
template <int NS, bool GRAD, int DIR>
__device__ __forceinline__ bool lattice_lin_dir(uint32_t ring, uint32_t bar0, uint32_t xch, uint32_t bx, int bar_id,
                                                int lane, int b, int Tb, int Ub,
                                                const int64_t *__restrict__ targets, int64_t tnumel, int64_t toff,
                                                int *__restrict__ flags, const float *__restrict__ lp_lab,
                                                float *__restrict__ gam, unsigned char *__restrict__ ab_utt,
                                                float *__restrict__ nll, int T) {
    using C = LinCfg<NS, GRAD>;
    constexpr int NL = C::NL, Lp = C::Lp, TT = C::TT, NSTG = C::NSTG;
    if (lane == 0) {
        for (int s = 0; s < NSTG; ++s) mbar_init(bar0 + 8 * s, 1);
        fence_mbar_init();
    }
    LinLane<NS> L;
    L.lane = lane; L.Tb = Tb; L.Ub = Ub; L.E = 0; L.f = 0.0;
    {   // label structure: which skip transitions exist
        int lab[NL];
#pragma unroll
        for (int jj = 0; jj < NL; ++jj) {
            const int li = NL * lane + jj;
            lab[jj] = -1 - li;
            if (li < Ub) { const int64_t idx = toff + li; lab[jj] = idx < tnumel ? (int)targets[idx] : 0; }
        }
        int pl = __shfl_up_sync(0xffffffffu, lab[NL - 1], 1);
        int nl = __shfl_down_sync(0xffffffffu, lab[0], 1);
        if (lane == 0) pl = -1000000;
        if (lane == 31) nl = -1000001;
#pragma unroll
        for (int jj = 0; jj < NL; ++jj) {
            const int li = NL * lane + jj;
            bool ok;
            if (DIR == 0) {
                const int prev = jj == 0 ? pl : lab[(jj + NL - 1) % NL];
                ok = li < Ub && li >= 1 && lab[jj] != prev;
            } else {
                const int next = jj == NL - 1 ? nl : lab[(jj + 1) % NL];
                ok = li + 1 < Ub && lab[jj] != next;
            }
            L.sk[jj] = ok ? 1.0 : 0.0;
        }
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) { L.m[j] = 0.0; L.sum[j] = 0.0; }

    // Per-lane renormalisation at a stage boundary (see the header comment).
    auto renorm = [&]() {
        int hmax = 0;
#pragma unroll
        for (int j = 0; j < NS; ++j) hmax = max(hmax, __double2hiint(L.m[j]));   // m >= 0: integer order = value order
        const bool zero = hmax < (1 << 20);                      // all zero (or denormal: flushed)
        if (zero) {
#pragma unroll
            for (int j = 0; j < NS; ++j) L.m[j] = 0.0;
        } else {
            int e = (hmax >> 20) - 1023;
            e = e > 1000 ? 1000 : e;
            const double sc = pow2i(-e);
#pragma unroll
            for (int j = 0; j < NS; ++j) L.m[j] *= sc;
            L.E += e;
        }
        const unsigned nz = __ballot_sync(0xffffffffu, !zero);
        int src = lane;
        if (DIR == 0) {
            const unsigned below = nz & (0xffffffffu >> (31 - lane));
            if (below) src = 31 - __clz(below);
        } else {
            const unsigned above = nz & (0xffffffffu << lane);
            if (above) src = __ffs(above) - 1;
        }
        const int Es = __shfl_sync(0xffffffffu, L.E, src);
        if (zero) L.E = Es;
        const int En = DIR == 0 ? __shfl_up_sync(0xffffffffu, L.E, 1) : __shfl_down_sync(0xffffffffu, L.E, 1);
        int d = En - L.E;
        d = d < -900 ? -900 : (d > 900 ? 900 : d);
        L.f = (lane == (DIR ? 31 : 0)) ? 0.0 : pow2i(d);
    };

    // ---- tiling of time (as in lattice_kernel.cuh, with this path's TT) ----
    const int Qtot = (Tb + TT - 1) / TT;
    int Tm = ((Tb / 2 + TT / 2) / TT) * TT;
    if (Tm >= Tb) Tm = ((Tb - 1) / TT) * TT;
    const int Qm = Tm / TT;
    const int n1 = DIR ? (Qtot - Qm) : Qm;
    const int ntot = GRAD ? Qtot : (DIR ? n1 : Qm + 1);
    const float *lp_base = lp_lab + (size_t)b * T * Lp;           // on this path every blank/label value is a probability
    float *gam_base = gam + (size_t)b * T * Lp;
    int nact = (2 * Ub + 1 + NS - 1) / NS;                       // lanes that hold a real state
    nact = nact > 32 ? 32 : nact;
    const uint32_t RS = (uint32_t)nact * NS * 8;                 // bytes of a stored row
    const uint32_t BLK = C::EXPS + TT * RS;                      // bytes of a stored stage

    auto issue = [&](int n) {                                    // lane 0 only
        const int q = DIR ? (Qtot - 1 - n) : n;
        const int t0 = q * TT;
        const int rows = (Tb - t0) < TT ? (Tb - t0) : TT;
        const bool ph2 = GRAD && (DIR ? (q < Qm) : (q >= Qm));
        const int stg = n % NSTG;
        const uint32_t dst = ring + stg * C::STAGE, bar = bar0 + 8 * stg;
        mbar_expect_tx(bar, rows * C::LP_ROW + (ph2 ? C::EXPS + rows * RS : 0));
        tma_load_1d_hint(dst, lp_base + (size_t)t0 * Lp, rows * C::LP_ROW, bar, kScratch);
        if (ph2) tma_load_1d_hint(dst + TT * C::LP_ROW, ab_utt + (size_t)q * BLK, C::EXPS + rows * RS, bar, kScratch);
    };
    const int act = lane < nact;
    const uint32_t RSd = RS / 8;                                 // row stride in doubles
    auto store_row = [&](double *row) {                          // row: this lane's part of the stored row
#pragma unroll
        for (int k = 0; k < NS / 2; ++k) stg_v2f64_hint_if(act, row + 2 * k, L.sum[2 * k], L.sum[2 * k + 1], kScratch);
    };

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
        renorm();
        if (GRAD) ((int *)(ab_utt + (size_t)q * BLK))[lane] = L.E;
        double *row0 = (double *)(ab_utt + (size_t)q * BLK + C::EXPS) + NS * lane;
        if (rows == TT && n > 0) {                               // full stage, not the first: straight-line
            LinP<NS> P[TT];
#pragma unroll
            for (int r = 0; r < TT; ++r) lin_load_p<NS>(P[r], tile + (DIR ? (TT - 1 - r) : r) * C::LP_ROW, lane);
            double *row = row0 + (DIR ? (TT - 1) * RSd : 0);
#pragma unroll
            for (int r = 0; r < TT; ++r) {
                lin_chain<NS, DIR, false>(L, P[r]);
                if (GRAD) { store_row(row); row = DIR ? row - RSd : row + RSd; }
            }
        } else {                                                 // first stage (holds the initial step) or ragged
#pragma unroll 1
            for (int r = 0; r < rows; ++r) {
                const int rr = DIR ? (rows - 1 - r) : r;
                LinP<NS> P;
                lin_load_p<NS>(P, tile + rr * C::LP_ROW, lane);
                if (n == 0 && r == 0) lin_chain<NS, DIR, true>(L, P);
                else lin_chain<NS, DIR, false>(L, P);
                if (GRAD) store_row(row0 + rr * RSd);
            }
        }
        __syncwarp();
        if (n_issue < n1) { if (lane == 0) issue(n_issue); ++n_issue; }
    }
    // ================= midpoint =================
    if (GRAD) fence_proxy_async_global();
    else if (DIR == 1) {                                         // loss only: hand the last row over in shared memory
        asm volatile("st.shared.s32 [%0], %1;" ::"r"(bx + 4 * lane), "r"(L.E) : "memory");
#pragma unroll
        for (int j = 0; j < NS; ++j)
            asm volatile("st.shared.f64 [%0], %1;" ::"r"(bx + C::EXPS + 8 * (NS * lane + j)), "d"(L.sum[j]) : "memory");
    }
    named_bar_sync(bar_id, 64);

    // ================= phase 2: recursion + occupancies =================
    for (; n_issue < n1 + NSTG && n_issue < ntot; ++n_issue) if (lane == 0) issue(n_issue);
    double inv_lm = 0.0;                                         // 1 / mantissa of the likelihood
    int EL = 0;                                                  // exponent of the likelihood
    bool failed = false;
    if (DIR == 1) {                                              // alpha publishes the likelihood at its first phase-2 step
        named_bar_sync(bar_id, 64);
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(inv_lm) : "r"(xch));
        int st;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(EL) : "r"(xch + 8));
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(st) : "r"(xch + 12));
        failed = st != 0;
    }
    int n_waited = n1;
    for (int n = n1; n < ntot && !failed; ++n) {
        const int stg = n % NSTG;
        mbar_wait(bar0 + 8 * stg, (n / NSTG) & 1);
        n_waited = n + 1;
        const int q = DIR ? (Qtot - 1 - n) : n;
        const int t0 = q * TT;
        const int rows = (Tb - t0) < TT ? (Tb - t0) : TT;
        const uint32_t tile = ring + stg * C::STAGE;
        const uint32_t orow = GRAD ? tile + TT * C::LP_ROW : bx;   // {exps, rows} of the other direction
        renorm();
        int Eo;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(Eo) : "r"(orow + 4 * lane));
        const int Esum = L.E + Eo;
        auto gscale = [&]() -> double {
            int d = Esum - EL;
            d = d < -1000 ? -1000 : (d > 1000 ? 1000 : d);
            return pow2i(d) * inv_lm;
        };
        double sc = gscale();
        float *gam_t = gam_base + (size_t)t0 * Lp + 4 + NL * lane;   // this lane's label slots of the stage's first frame
        const int gst = NL * lane < Ub;
        // one phase-2 step on frame t0+rr: recursion, then gamma = alpha * beta^ / P; returns this lane's blank part
        auto step2 = [&](const LinP<NS> &P, int rr, bool first, bool init) -> float {
            if (init) lin_chain<NS, DIR, true>(L, P);
            else lin_chain<NS, DIR, false>(L, P);
            double ot[NS];
            lds_vec_d<NS>(ot, orow + C::EXPS + (GRAD ? rr * RS : 0) + 8 * NS * lane);
            if (DIR == 0 && first) {                             // midpoint: the likelihood (warp-uniform branch)
                double part = 0.0;
#pragma unroll
                for (int j = 0; j < NS; ++j) part = fma(L.m[j], ot[j], part);
                if (!act) part = 0.0;                            // lanes past the stored row read stale shared memory
                const int hp = __double2hiint(part);
                const bool live = hp >= (1 << 20);
                const int X = live ? Esum + ((hp >> 20) - 1023) : -(1 << 29);
                const int Xmax = warp_max_i(X);
                failed = Xmax == -(1 << 29);
                int d = Esum - Xmax;
                d = d > 1000 ? 1000 : d;
                const double scaled = (live && d >= -1000) ? part * pow2i(d) : 0.0;
                const double Lm = warp_sum_d(scaled);            // in [1, 64)
                EL = Xmax;
                inv_lm = failed ? 0.0 : 1.0 / Lm;
                if (lane == 0) {
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(xch), "d"(inv_lm) : "memory");
                    asm volatile("st.shared.s32 [%0], %1;" ::"r"(xch + 8), "r"(EL) : "memory");
                    asm volatile("st.shared.s32 [%0], %1;" ::"r"(xch + 12), "r"((int)failed) : "memory");
                    if (!failed) {
                        nll[b] = (float)(-((double)EL + log2(Lm)) * 0.6931471805599453);
                        flags[b] = 0;
                    }
                }
                named_bar_sync(bar_id, 64);
                sc = gscale();
            }
            float gb = 0.f;
            if (GRAD && !failed) {
                float gl[NL];
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const float g = (float)((L.m[j] * ot[j]) * sc);
                    if (j & 1) gl[j >> 1] = g; else gb += g;
                }
                stg_vec_if<NL>(gst, gam_t + rr * Lp, gl);
                gb = act ? gb : 0.f;                             // (stale shared memory past the stored row)
            }
            return gb;
        };
        const bool first_job = (DIR == 0 && n == n1);
        if (rows == TT && !first_job) {
            LinP<NS> P[TT];
#pragma unroll
            for (int r = 0; r < TT; ++r) lin_load_p<NS>(P[r], tile + (DIR ? (TT - 1 - r) : r) * C::LP_ROW, lane);
            float gbl[TT];
#pragma unroll
            for (int r = 0; r < TT; ++r) gbl[r] = step2(P[r], DIR ? (TT - 1 - r) : r, false, false);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int r = 0; r < TT; ++r) gbl[r] += __shfl_xor_sync(0xffffffffu, gbl[r], o);
            }
            float mine = 0.f;
#pragma unroll
            for (int r = 0; r < TT; ++r) if (lane == r) mine = gbl[r];
            if (lane < TT) {
                const int rr = DIR ? (TT - 1 - lane) : lane;
                stg_v2_hint((float2 *)(gam_base + (size_t)(t0 + rr) * Lp), make_float2(mine, lds_f32(tile + rr * C::LP_ROW + 4)), kScratch);
            }
        } else {
#pragma unroll 1
            for (int r = 0; r < rows; ++r) {
                const int rr = DIR ? (rows - 1 - r) : r;
                LinP<NS> P;
                lin_load_p<NS>(P, tile + rr * C::LP_ROW, lane);
                float gb = step2(P, rr, first_job && r == 0, DIR == 0 && t0 + rr == 0);
                if (!GRAD || failed) break;
                gb = warp_sum(gb);
                if (lane == 0)
                    stg_v2_hint((float2 *)(gam_base + (size_t)(t0 + rr) * Lp), make_float2(gb, lds_f32(tile + rr * C::LP_ROW + 4)), kScratch);
            }
        }
        if (!GRAD || failed) break;
        __syncwarp();
        if (n_issue < ntot) { if (lane == 0) issue(n_issue); ++n_issue; }
    }
    // never leave with bulk copies still landing in shared memory
    for (int n = n_waited; n < n_issue; ++n) mbar_wait(bar0 + 8 * (n % NSTG), (n / NSTG) & 1);
    return !failed;
}

template <int NS, bool GRAD>
constexpr uint32_t k2_smem_bytes() { return K2Smem<NS, GRAD>::SMEM; }

template <int NS, bool GRAD>
__global__ void __launch_bounds__(128, 1)
k2_lattice(const int64_t *__restrict__ targets, int64_t tnumel, const int *__restrict__ Tb_arr,
           const int *__restrict__ Ub_arr, const int64_t *__restrict__ toff_arr, int *__restrict__ flags,
           const float *__restrict__ lp_lab, float *__restrict__ gam, float *__restrict__ ab_ws,
           float *__restrict__ nll, float *__restrict__ loss_sums, unsigned *__restrict__ ticket, int B,
           int T, int zero_inf, float *__restrict__ zero_grad, const int *__restrict__ rowstart, int V,
           double *__restrict__ tile_off, float mean_scale, const int *__restrict__ slow, size_t ab_utt_bytes,
           const int *__restrict__ bad_arr) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    griddep_wait();                                              // the sweep's lp_lab frames
    griddep_launch_dependents();
    const int n_lat = (B + 1) / 2;
    if ((int)blockIdx.x >= n_lat) {
        // Extra CTAs of the same launch: while the (latency-bound) lattice CTAs run, these write the zeros of
        // the padded frames of grad -- HBM work of the step that would otherwise sit in the sweep kernel.
        zero_padded_frames<128>(zero_grad, Tb_arr, rowstart, B, T, V, tid, (int)blockIdx.x - n_lat,
                                (int)gridDim.x - n_lat);
        return;
    }
    const int pair = warp >> 1, dir = warp & 1;
    const int b = 2 * blockIdx.x + pair;
    if (b >= B) return;                                          // odd batch: the last CTA has one utterance
    const int Tb = __ldcg(Tb_arr + b), Ub = __ldcg(Ub_arr + b);   // (ld.global.cg: see stream_kernels.cuh, cursor helpers)

    // Structural feasibility: an alignment exists iff T_b >= U_b + (number of adjacent equal labels).  An infeasible
    // utterance is settled here -- otherwise it would run the linear recursion up to its underflow at the midpoint AND
    // the log-space recursion after it, and the launch lasts as long as its slowest utterance (C4: 8 of 64).
    bool infeasible = false;
    if (Tb > 0) {
        const int64_t toff = __ldcg(toff_arr + b);
        int rep = 0;
        for (int j = 1 + lane; j < Ub; j += 32) {
            const int64_t i1 = toff + j;
            rep += (i1 < tnumel ? (int)targets[i1] : 0) == (i1 - 1 < tnumel ? (int)targets[i1 - 1] : 0);
        }
        rep = __reduce_add_sync(0xffffffffu, rep);
        infeasible = Tb < Ub + rep;
    }
    if (infeasible) {
        if (dir == 0 && lane == 0) {
            nll[b] = zero_inf ? 0.f : __int_as_float(0x7f800000);
            flags[b] = 1;
        }
    } else if (Tb > 0) {
        const int64_t toff = __ldcg(toff_arr + b);
        // fast path: linear-domain recursion; utterances outside its range (slow[b], set by the sweep) or whose
        // likelihood underflows (which includes the infeasible ones) run the log-space recursion
        unsigned char *ab_utt = (unsigned char *)ab_ws + (size_t)b * ab_utt_bytes;
        using M = K2Smem<NS, GRAD>;
        const int wq = pair * 2 + dir;
        const uint32_t sb = smem_u32(smem);
        const uint32_t ring = sb + wq * M::RING;
        const uint32_t xch = sb + M::OFF_XCH + pair * 32, bx = sb + M::OFF_BX + pair * M::BX;
        const uint32_t bars_lin = sb + M::OFF_BARS_LIN + wq * M::NSTG * 8, bars_log = sb + M::OFF_BARS_LOG + wq * M::NSTG * 8;
        bool done = false;
        if (!__ldcg(slow + b)) {                                 // (written by the sweep: coherent load, see k3p_patch)
            done = dir == 0 ? lattice_lin_dir<NS, GRAD, 0>(ring, bars_lin, xch, bx, 1 + pair, lane, b, Tb, Ub, targets,
                                                           tnumel, toff, flags, lp_lab, gam, ab_utt, nll, T)
                            : lattice_lin_dir<NS, GRAD, 1>(ring, bars_lin, xch, bx, 1 + pair, lane, b, Tb, Ub, targets,
                                                           tnumel, toff, flags, lp_lab, gam, ab_utt, nll, T);
            if (!done) {
                named_bar_sync(1 + pair, 64);                    // both warps are out of the rings and of xch
                if (dir == 0 && lane == 0) atomicAdd(ticket + 2, 1u);   // debug counter: underflowed / infeasible
            }
        }
        if (!done) {
            if (dir == 0 && lane == 0) atomicAdd(ticket + 1, 1u);       // debug counter: utterances run in log space
            if (dir == 0)
                lattice_dir<NS, GRAD, 0>(ring, bars_log, xch, bx, 1 + pair, lane, b, Tb, Ub, targets, tnumel, toff, flags,
                                         lp_lab, gam, (float *)ab_utt, nll, T, zero_inf, tile_off);
            else
                lattice_dir<NS, GRAD, 1>(ring, bars_log, xch, bx, 1 + pair, lane, b, Tb, Ub, targets, tnumel, toff, flags,
                                         lp_lab, gam, (float *)ab_utt, nll, T, zero_inf, tile_off);
        }
    } else if (dir == 0 && lane == 0) {
        // no frames: empty target -> probability 1, anything else is infeasible (torch: inf, zero grad)
        nll[b] = (Ub == 0) ? 0.f : (zero_inf ? 0.f : __int_as_float(0x7f800000));
        flags[b] = (Ub != 0);
    }

    // invalid lengths / labels (clamped for memory safety by k0_prep / the sweep): F.ctc_loss raises on such inputs;
    // here the utterance's nll -- and with it every reduced loss -- becomes NaN, so the error cannot go unnoticed
    if (dir == 0 && lane == 0 && __ldcg(bad_arr + b)) nll[b] = __int_as_float(0x7fc00000);

    // ---- deterministic batch reduction by the last utterance to finish ----
    if (loss_sums != nullptr && dir == 0) {
        unsigned tk = 0;
        if (lane == 0) { __threadfence(); tk = atomicAdd(ticket, 1u); }
        tk = __shfl_sync(0xffffffffu, tk, 0);
        if (tk == (unsigned)B - 1) {
            __threadfence();
            float s_norm = 0.f, s_sum = 0.f;
            for (int i = lane; i < B; i += 32) {
                const float v = __ldcg(nll + i);
                const int u = __ldcg(Ub_arr + i);
                s_sum += v;
                s_norm += v / (float)(u > 1 ? u : 1);
            }
            s_norm = warp_sum(s_norm);
            s_sum = warp_sum(s_sum);
            if (lane == 0) {
                loss_sums[0] = s_norm; loss_sums[1] = s_sum; loss_sums[2] = (float)B; loss_sums[3] = s_norm * mean_scale;
                *ticket = 0;
            }
        }
    }
}

}  // namespace ctcb200
