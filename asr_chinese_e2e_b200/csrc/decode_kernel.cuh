// k5: on-device greedy (best-path) CTC decode + edit distance.
// The sweep kernel leaves the per-frame argmax class in best[B,T]; one warp per utterance collapses
// repeats, drops blanks and runs a Levenshtein DP against the reference labels, so the per-step
// D->H sync + Python Levenshtein loop of the reference's cal_metrics
// (Predictor/Models/transformer_official.py:87-91, Predictor/Utils/score.py:4-13) is not needed for a
// CTC-branch character error rate (token-level distance).
// k6: the reference's own `cer` metric for any pair of id matrices (attention-branch argmax vs gold), on the
// device: Levenshtein distance between the SPACE-JOINED strings, exactly what calculate_cer computes.
#pragma once
#include "ptx.cuh"

namespace ctcb200 {

template <int NREF>   // reference labels per lane: U <= 32*NREF
__global__ void __launch_bounds__(128)
k5_greedy_cer(const int64_t *__restrict__ targets, int64_t tnumel, const int *__restrict__ Tb_arr,
              const int *__restrict__ Ub_arr, const int64_t *__restrict__ toff_arr, const int *__restrict__ best,
              int *__restrict__ edit, int *__restrict__ hyp_len, int64_t *__restrict__ hyp_out, int B, int T, int V,
              int blank) {
    extern __shared__ int smem_i[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = 4 * blockIdx.x + warp;
    if (b >= B) return;
    int *hyp = smem_i + warp * T;
    const int Tb = Tb_arr[b], Ub = Ub_arr[b];
    const int64_t toff = toff_arr[b];
    // ---- 1. collapse repeats, drop blanks ----
    int n = 0, carry = blank;
    for (int t0 = 0; t0 < Tb; t0 += 32) {
        const int t = t0 + lane;
        int v = t < Tb ? best[(size_t)b * T + t] : blank;
        if (v < 0 || v >= V) v = blank;                       // NaN rows leave no valid argmax
        int prev = __shfl_up_sync(0xffffffffu, v, 1);
        if (lane == 0) prev = carry;
        const bool keep = t < Tb && v != blank && v != prev;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) hyp[n + __popc(m & ((1u << lane) - 1u))] = v;
        n += __popc(m);
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
    __syncwarp();
    if (hyp_out != nullptr)
        for (int t = lane; t < T; t += 32) hyp_out[(size_t)b * T + t] = t < n ? hyp[t] : blank;
    // ---- 2. Levenshtein distance hyp[0..n) vs labels y[0..Ub) ----
    // lane owns columns c = NREF*lane + jj + 1 (column 0 is the boundary D[i][0] = i)
    int y[NREF], D[NREF];
#pragma unroll
    for (int jj = 0; jj < NREF; ++jj) {
        const int j = NREF * lane + jj;
        y[jj] = -1;
        if (j < Ub) { const int64_t idx = toff + j; y[jj] = idx < tnumel ? (int)targets[idx] : -1; }
        D[jj] = j + 1;
    }
    for (int i = 0; i < n; ++i) {
        const int h = hyp[i];
        int diag = __shfl_up_sync(0xffffffffu, D[NREF - 1], 1);
        if (lane == 0) diag = i;                             // D[i][0]
        int a[NREF], run = 0x3fffffff;
#pragma unroll
        for (int jj = 0; jj < NREF; ++jj) {
            const int up = D[jj];
            const int tmp = min(up + 1, diag + (y[jj] != h));
            diag = up;
            run = min(run, tmp - (NREF * lane + jj + 1));     // a[k] = tmp[k] - k, inclusive prefix min in the lane
            a[jj] = run;
        }
        int inc = run;                                        // inclusive prefix min over lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc = min(inc, v);
        }
        int excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0x3fffffff;
        excl = min(excl, i + 1);                              // column 0: a[0] = (i+1) - 0
#pragma unroll
        for (int jj = 0; jj < NREF; ++jj) D[jj] = (NREF * lane + jj + 1) + min(a[jj], excl);
    }
    int res = n;                                              // Ub == 0: distance = hyp length
    if (Ub > 0) {
        const int owner = (Ub - 1) / NREF, slot = (Ub - 1) % NREF;
        int v = 0;
#pragma unroll
        for (int jj = 0; jj < NREF; ++jj) if (jj == slot) v = D[jj];
        res = __shfl_sync(0xffffffffu, v, owner);
    }
    if (lane == 0) { edit[b] = res; hyp_len[b] = n; }
}

// ------------------------------------------------------------------------------------------------
// k6: edit distance between two id sequences per row (hyp[B,L], gold[B,L], int64, ids == pad dropped like
// Vocab.convert_id2str, Predictor/data_handler/vocab.py:74-78).
//   mode 0: token-level Levenshtein distance.
//   mode 1: the reference's metric (Predictor/Utils/score.py:4-13, called from transformer_official.py:89-91):
//           Lev.distance on the space-joined strings "t1 t2 ... tn", i.e. on the symbol sequences
//           t1 SP t2 SP ... tn (2n-1 symbols).  Identical to python-Levenshtein on the strings because every token
//           of the reference's vocabulary is a single character (vocab.py:4-5 tokenises by character).
// words[b] = number of words of the gold string = max(n_gold, 1)  (len("".split(" ")) == 1).
// One warp per row; lane owns NREF consecutive DP columns (gold symbols): 32*NREF >= 2L-1.
// ------------------------------------------------------------------------------------------------
template <int NREF>
__global__ void __launch_bounds__(128)
k6_seq_edit(const int64_t *__restrict__ hyp, int64_t hyp_stride, const int64_t *__restrict__ gold, int64_t gold_stride,
            int B, int L, int pad, int mode, int *__restrict__ edit, int *__restrict__ words) {
    extern __shared__ int smem_i[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = 4 * blockIdx.x + warp;
    if (b >= B) return;
    int *hs = smem_i + warp * 2 * L, *gs = hs + L;
    auto compact = [&](const int64_t *row, int *dst) -> int {
        int n = 0;
        for (int t0 = 0; t0 < L; t0 += 32) {
            const int t = t0 + lane;
            const long long v = t < L ? row[t] : (long long)pad;
            const bool keep = t < L && v != (long long)pad;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) dst[n + __popc(m & ((1u << lane) - 1u))] = (int)v;
            n += __popc(m);
        }
        __syncwarp();
        return n;
    };
    const int nh = compact(hyp + (size_t)b * hyp_stride, hs);
    const int ng = compact(gold + (size_t)b * gold_stride, gs);
    constexpr int SP = -7;                                    // the separator: no id is negative
    const int mh = mode ? (nh > 0 ? 2 * nh - 1 : 0) : nh;     // symbols of the (joined) hypothesis / gold
    const int mg = mode ? (ng > 0 ? 2 * ng - 1 : 0) : ng;
    auto sym = [&](const int *s, int i) -> int { return mode ? ((i & 1) ? SP : s[i >> 1]) : s[i]; };
    int y[NREF], D[NREF];
#pragma unroll
    for (int jj = 0; jj < NREF; ++jj) {
        const int j = NREF * lane + jj;
        y[jj] = j < mg ? sym(gs, j) : -1000 - j;
        D[jj] = j + 1;
    }
    for (int i = 0; i < mh; ++i) {
        const int h = sym(hs, i);
        int diag = __shfl_up_sync(0xffffffffu, D[NREF - 1], 1);
        if (lane == 0) diag = i;
        int a[NREF], run = 0x3fffffff;
#pragma unroll
        for (int jj = 0; jj < NREF; ++jj) {
            const int up = D[jj];
            const int tmp = min(up + 1, diag + (y[jj] != h));
            diag = up;
            run = min(run, tmp - (NREF * lane + jj + 1));
            a[jj] = run;
        }
        int inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc = min(inc, v);
        }
        int excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0x3fffffff;
        excl = min(excl, i + 1);
#pragma unroll
        for (int jj = 0; jj < NREF; ++jj) D[jj] = (NREF * lane + jj + 1) + min(a[jj], excl);
    }
    int res = mh;
    if (mg > 0) {
        const int owner = (mg - 1) / NREF, slot = (mg - 1) % NREF;
        int v = 0;
#pragma unroll
        for (int jj = 0; jj < NREF; ++jj) if (jj == slot) v = D[jj];
        res = __shfl_sync(0xffffffffu, v, owner);
    }
    if (lane == 0) { edit[b] = res; words[b] = ng > 0 ? ng : 1; }
}

}  // namespace ctcb200
