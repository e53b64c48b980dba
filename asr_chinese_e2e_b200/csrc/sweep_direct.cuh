// k1d: the fused sweep of round 2 -- aligned frame groups, direct loads, one barrier per group.
//
// Same job as k1_lse_gather (stream_kernels.cuh): ONE read of every valid logits frame -> per-frame log-sum-exp and the
// U+1 label log-probabilities the lattice needs, and (FUSED) the dense part of the gradient g_b * softmax(x) written in
// the logits' layout.  Round 1 measured what bounds that kernel (profiles/r01_bw_probe_output.txt): a read+write
// stream through a shared-memory TMA ring tops out near 6.0 TB/s on this GPU, a frame of V = 4234 floats starts 8
// bytes off every other row (16-byte hull copy + two scalar edge threads: -5 %), and a bare read-exp-write kernel with
// direct LDG.128 loads reaches 6.5-6.6 TB/s -- as long as shared memory stays small, because direct loads live on
// the L1 left over by the carve-out.  This kernel is that probe grown into the real thing:
//   * work unit = an ALIGNED GROUP of P = 4 / gcd(V, 4) consecutive frames of one utterance (P = 2 for V = 4234:
//     33 872 bytes = 2117 naturally aligned 16-byte chunks; the chunk in the middle holds the last two floats of the
//     even frame and the first two of the odd one).  Every load is an aligned LDG.128, every store an aligned
//     STG.128: no hull, no scalar edge code, nothing fetched or written twice;
//   * the group lives in registers (MAXC chunks per thread), no shared-memory staging at all: shared memory is 1 KB
//     of reduction scratch + the utterance's class table, so the L1 keeps its default size;
//   * the label logits of a frame are fetched by U+1 extra 4-byte loads issued together with the group's loads
//     (same sectors, no extra DRAM traffic) instead of being picked out of a staged copy;
//   * ONE block barrier per group (P frames): every warp reduces against its own maxima first, the partials are
//     combined after the barrier (see k1_lse_gather).
// Used when V % 4 is 0 or 2, T % P == 0 and a group fits 128 x MAXC chunks; otherwise ctcb200.cu launches
// k1_lse_gather.
#pragma once
#include "stream_kernels.cuh"

namespace ctcb200 {

struct K1dArgs {
    const float *logits; const int64_t *targets; int64_t tnumel; const int *Tb, *Ub; const int64_t *toff;
    const int *rowstart, *gstart; float *lp_lab; int *hdr; int B, T, V, Lp, blank, P;
    float *grad; int reduction; float inv_batch;       // FUSED only
    int *best; int zero_pad_here; int *slow; float lin_thr; int *bad;
};

__device__ __forceinline__ void stg_v4_cs(float4 *p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int MAXC, bool FUSED>
__global__ void __launch_bounds__(128, (MAXC <= 17 ? 4 : 2)) k1d_sweep(const K1dArgs a) {
    constexpr int NT = 128;
    __shared__ float red[2][4][4];           // [parity][warp][max0, max1, sum0, sum1]
    __shared__ int redi[2][4][2];            // [parity][warp][argmax0, argmax1]   (greedy decode only)
    __shared__ int cls_s[264];               // class id per frame slot of the current utterance
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    griddep_wait();                          // k0_prep's lengths / prefix sums
    griddep_launch_dependents();
    const int B = a.B, T = a.T, V = a.V, P = a.P, Lp = a.Lp;
    if (FUSED && a.zero_pad_here) zero_padded_frames<NT>(a.grad, a.Tb, a.rowstart, B, T, V, tid);
    int g0, ng;
    grid_share(a.gstart[B], g0, ng);         // this CTA's contiguous share of the live groups
    if (ng <= 0) return;

    // chunk geometry of a group: nch chunks of 4 floats; chunk c belongs to frame 0 below `mid`, to frame 1 above,
    // and (P == 2) chunk `mid` itself is split: .x .y = last floats of frame 0, .z .w = first floats of frame 1
    const int nch = (P * V) >> 2;
    const int mid = P == 2 ? (V >> 2) : 0x7fffffff;      // V % 4 == 2  =>  V = 4 * mid + 2

    int b, j;                                // group cursor: utterance b, group j of it (frames jP .. jP+P-1)
    {
        int lo = 0, hi = B - 1;
        while (lo < hi) {
            const int m = (lo + hi) >> 1;
            if (a.gstart[m + 1] > g0) hi = m; else lo = m + 1;
        }
        b = lo; j = g0 - a.gstart[lo];
    }
    int ngb = a.gstart[b + 1] - a.gstart[b], cur_b = -1, Tbb = 0;
    float gsc = 0.f;
    constexpr int MAXQ = (2 * 264 + NT - 1) / NT;        // label loads per thread: P frames x Lp slots over 128 threads
    for (int i = 0; i < ng; ++i) {
        if (b != cur_b) {                    // block-uniform: (re)load the utterance's class ids
            cur_b = b;
            Tbb = a.Tb[b];
            const int Ub = a.Ub[b];
            if (FUSED) gsc = a.reduction == 1 ? a.inv_batch * __frcp_rn((float)(Ub > 1 ? Ub : 1)) : 1.f;
            const int64_t toff = a.toff[b];
            __syncthreads();                 // the previous utterance's table reads are done
            for (int k = tid; k < Lp; k += NT) {
                int cls;
                if (k == 0) cls = a.blank;
                else if (k == 1) cls = -2;   // slot of lse2
                else if (k < 4) cls = -3;    // unused header slots -> 0
                else if (k - 4 < Ub) {
                    const int64_t idx = toff + (k - 4);
                    long long c = idx < a.tnumel ? a.targets[idx] : -1;
                    if (c < 0 || c >= V || c == a.blank) {
                        atomicOr(&a.hdr[0], 4);
                        if (c < 0 || c >= V) atomicOr(&a.bad[b], 4);
                        c = c < 0 ? 0 : (c >= V ? V - 1 : c);
                    }
                    cls = (int)c;
                } else cls = -1;             // beyond U_b -> sentinel
                cls_s[k] = cls;
            }
            __syncthreads();
        }
        const int t0 = j * P;
        const bool live1 = P == 2 && t0 + 1 < Tbb;       // (the odd frame of an utterance's last group may be padding)
        const float *grow = a.logits + ((size_t)b * T + t0) * V;
        const float4 *g4 = (const float4 *)grow;
        // ---- the group: global -> registers, every load in flight before the first use ----
        float4 v[MAXC];
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
            const int c = tid + k * NT;
            v[k] = make_float4(CTC_NEG_INF, CTC_NEG_INF, CTC_NEG_INF, CTC_NEG_INF);
            if (c < nch) v[k] = ldg_v4_stream(g4 + c);
        }
        float xg[MAXQ];                      // label logits: slot q of frame r  <-  item tid + kk*NT = r*Lp + q
#pragma unroll
        for (int kk = 0; kk < MAXQ; ++kk) {
            const int it = tid + kk * NT;
            xg[kk] = 0.f;
            if (it < P * Lp) {
                const int r = it >= Lp, q = it - r * Lp, c = cls_s[q];
                if (c >= 0 && (r == 0 || live1)) xg[kk] = __ldg(grow + (size_t)r * V + c);
            }
        }
        // ---- per-frame max over this warp, then 2^(x - max_w) and its sum: no other warp needed ----
        float mx0 = CTC_NEG_INF, mx1 = CTC_NEG_INF;
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
            const int c = tid + k * NT;
            const float lo2 = fmaxf(v[k].x, v[k].y), hi2 = fmaxf(v[k].z, v[k].w);
            if (c < mid) mx0 = fmaxf(mx0, fmaxf(lo2, hi2));
            else if (c > mid) mx1 = fmaxf(mx1, fmaxf(lo2, hi2));
            else { mx0 = fmaxf(mx0, lo2); mx1 = fmaxf(mx1, hi2); }
        }
        const float mw0 = warp_max(mx0), mw1 = warp_max(mx1);
        const float e0 = (mw0 == CTC_NEG_INF ? 0.f : mw0) * kLog2e, e1 = (mw1 == CTC_NEG_INF ? 0.f : mw1) * kLog2e;
        float *rd = &red[i & 1][warp][0];
        if (a.best != nullptr) {             // lowest class index attaining the warp maximum, per frame
            int c0 = 0x7fffffff, c1 = 0x7fffffff;
#pragma unroll
            for (int k = MAXC - 1; k >= 0; --k) {
                const int c = tid + k * NT, e = 4 * c;                      // float index within the group
                if (c < nch) {
                    const bool f0lo = c <= mid, f0hi = c < mid;            // does .xy / .zw belong to frame 0?
                    if (f0hi ? v[k].w == mw0 : v[k].w == mw1) { if (f0hi) c0 = e + 3; else c1 = e + 3 - V; }
                    if (f0hi ? v[k].z == mw0 : v[k].z == mw1) { if (f0hi) c0 = e + 2; else c1 = e + 2 - V; }
                    if (f0lo ? v[k].y == mw0 : v[k].y == mw1) { if (f0lo) c0 = e + 1; else c1 = e + 1 - V; }
                    if (f0lo ? v[k].x == mw0 : v[k].x == mw1) { if (f0lo) c0 = e; else c1 = e - V; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                c0 = min(c0, __shfl_xor_sync(0xffffffffu, c0, o));
                c1 = min(c1, __shfl_xor_sync(0xffffffffu, c1, o));
            }
            if (lane == 0) { redi[i & 1][warp][0] = c0; redi[i & 1][warp][1] = c1; }
        }
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
            const int c = tid + k * NT;
            const float ba = c <= mid ? e0 : e1, bb2 = c < mid ? e0 : e1;   // exponent base of .xy / .zw
            v[k].x = ex2f(fmaf(v[k].x, kLog2e, -ba)); v[k].y = ex2f(fmaf(v[k].y, kLog2e, -ba));
            v[k].z = ex2f(fmaf(v[k].z, kLog2e, -bb2)); v[k].w = ex2f(fmaf(v[k].w, kLog2e, -bb2));
            const float lo2 = v[k].x + v[k].y, hi2 = v[k].z + v[k].w;
            if (c < mid) s0 += lo2 + hi2;
            else if (c > mid) s1 += lo2 + hi2;
            else { s0 += lo2; s1 += hi2; }
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        if (lane == 0) { rd[0] = mw0; rd[1] = mw1; rd[2] = s0; rd[3] = s1; }
        __syncthreads();                     // the group's only barrier
        float M0 = CTC_NEG_INF, M1 = CTC_NEG_INF;
#pragma unroll
        for (int w = 0; w < 4; ++w) { M0 = fmaxf(M0, red[i & 1][w][0]); M1 = fmaxf(M1, red[i & 1][w][1]); }
        const float M0l = M0 * kLog2e, M1l = M1 * kLog2e;
        float S0 = 0.f, S1 = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float a0 = red[i & 1][w][0], a1 = red[i & 1][w][1];
            S0 += red[i & 1][w][2] * ex2f(fmaf(a0 == CTC_NEG_INF ? 0.f : a0, kLog2e, -M0l));
            S1 += red[i & 1][w][3] * ex2f(fmaf(a1 == CTC_NEG_INF ? 0.f : a1, kLog2e, -M1l));
        }
        const float lse0 = M0l + lg2f(S0), lse1 = M1l + lg2f(S1);
        if (a.best != nullptr && tid < P && (tid == 0 || live1)) {
            int bi = 0x7fffffff;
#pragma unroll
            for (int w = 0; w < 4; ++w) if (red[i & 1][w][tid] == (tid ? M1 : M0)) bi = min(bi, redi[i & 1][w][tid]);
            a.best[(size_t)b * T + t0 + tid] = bi;
        }
        // ---- the frames for the lattice kernel (same format as k1_lse_gather) ----
#pragma unroll
        for (int kk = 0; kk < MAXQ; ++kk) {
            const int it = tid + kk * NT;
            if (it < P * Lp) {
                const int r = it >= Lp, q = it - r * Lp, c = cls_s[q];
                if (r == 0 || live1) {
                    const float lse2 = r ? lse1 : lse0;
                    float o;
                    if (c >= 0) {
                        o = fminf(fmaxf(fmaf(xg[kk], kLog2e, -lse2), kNeg), 0.f);
                        if (o >= a.lin_thr) o = ex2f(o);
                        else a.slow[b] = 1;                                // (also NaN)
                    } else {
                        o = c == -2 ? lse2 : (c == -3 ? 0.f : kNeg);
                    }
                    stg_f32_hint(a.lp_lab + ((size_t)b * T + t0 + r) * Lp + q, o, kScratch);   // re-read by the lattice
                }
            }
        }
        if (FUSED) {
            // ---- dense gradient g * softmax = 2^(x - max_w) * 2^(max_w - M) * g / S: aligned 16-byte stores ----
            const float f0 = gsc * ex2f(e0 - M0l) * __frcp_rn(S0);
            const float f1 = live1 ? gsc * ex2f(e1 - M1l) * __frcp_rn(S1) : 0.f;   // a padded odd frame gets its zeros here
            float4 *o4 = (float4 *)(a.grad + ((size_t)b * T + t0) * V);
#pragma unroll
            for (int k = 0; k < MAXC; ++k) {
                const int c = tid + k * NT;
                if (c < nch && (c <= mid || live1)) {
                    const float fa = c <= mid ? f0 : f1, fb = c < mid ? f0 : f1;
                    const bool hi_live = c < mid || live1;                 // (upper half of the middle chunk: frame 1)
                    stg_v4_cs(o4 + c, make_float4(v[k].x * fa, v[k].y * fa, hi_live ? v[k].z * fb : 0.f,
                                                  hi_live ? v[k].w * fb : 0.f));
                }
            }
        }
        if (++j >= ngb) {                    // next group
            j = 0;
            do { ++b; } while (b < B && a.gstart[b + 1] == a.gstart[b]);
            if (b < B) ngb = a.gstart[b + 1] - a.gstart[b];
        }
    }
}

// k1p: the same aligned-group sweep fed by a bulk-TMA ring (round 2, after the knock-out bisect of k1_lse_gather).
// What the bisect showed: at 2 CTAs/SM the ring kernel is bound by its per-row dependency chain (DRAM writes are free
// there), and every launch shape with more than two frame streams per SM loses as soon as the gradient goes to DRAM.
// So: keep TWO streams per SM, and get the latency hiding from instruction-level parallelism instead of occupancy --
// a ring slot holds a whole aligned group (two adjacent frames, one 33 872-byte bulk copy, no hull, no edge threads),
// every thread carries both frames through max / exp / sum interleaved, and one barrier serves the group.
// BULKST (FUSED only): the gradient of a group does not leave through 17 STG.128 per thread but goes back into the
// group's ring slot (in place) and is written to global memory by ONE bulk-TMA store (cp.async.bulk.global.shared)
// issued by thread 0 after a second block barrier; a slot is refilled one group later, once its store has read it
// (cp.async.bulk.wait_group.read), so the ring needs three slots: one landing, one being computed, one leaving.  A
// padded odd frame is written as zeros (what zero_padded_frames writes there as well).
template <int NT, int MAXC, bool FUSED, bool BULKST = false>
__global__ void __launch_bounds__(NT, 2) k1p_sweep(const K1dArgs a, const int nst, const uint32_t slot_bytes) {
    constexpr int NW = NT / 32;
    extern __shared__ __align__(128) unsigned char smem_p[];
    uint64_t *bars = (uint64_t *)(smem_p + (size_t)nst * slot_bytes);
    float (*red)[NW][4] = (float (*)[NW][4])(bars + nst);   // [parity][warp][max0, max1, sum0, sum1]
    int (*redi)[NW][2] = (int (*)[NW][2])(&red[2][0][0]);    // [parity][warp][argmax0, argmax1]   (greedy decode only)
    int *cls_s = &redi[2][0][0];                           // [264] class id per frame slot of the current utterance
    const uint32_t slot0 = smem_u32(smem_p), bar0 = smem_u32(bars);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    griddep_wait();                          // k0_prep's lengths / prefix sums
    griddep_launch_dependents();
    const int B = a.B, T = a.T, V = a.V, P = a.P, Lp = a.Lp;
    if (FUSED && a.zero_pad_here) zero_padded_frames<NT>(a.grad, a.Tb, a.rowstart, B, T, V, tid);
    int g0, ng;
    grid_share(a.gstart[B], g0, ng);         // this CTA's contiguous share of the live groups
    if (ng <= 0) return;

    // chunk geometry of a group: nch chunks of 4 floats; chunk c belongs to frame 0 below `mid`, to frame 1 above,
    // and (P == 2) chunk `mid` itself is split: .x .y = last floats of frame 0, .z .w = first floats of frame 1
    const int nch = (P * V) >> 2;
    const int mid = P == 2 ? (V >> 2) : 0x7fffffff;      // V % 4 == 2  =>  V = 4 * mid + 2

    int b, j;                                // group cursor: utterance b, group j of it (frames jP .. jP+P-1)
    {
        int lo = 0, hi = B - 1;
        while (lo < hi) {
            const int m = (lo + hi) >> 1;
            if (a.gstart[m + 1] > g0) hi = m; else lo = m + 1;
        }
        b = lo; j = g0 - a.gstart[lo];
    }
    int ngb = a.gstart[b + 1] - a.gstart[b], cur_b = -1, Tbb = 0;
    float gsc = 0.f;
    // ---- the ring: thread 0 keeps a producer cursor nst groups ahead of the consumers ----
    if (tid == 0) {
        for (int st = 0; st < nst; ++st) mbar_init(bar0 + 8 * st, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const uint32_t group_bytes = (uint32_t)P * V * 4;            // a multiple of 16 by construction of P
    int pb = b, pj = j, pngb = ngb, issued = 0;
    auto issue = [&](int st) {                                  // thread 0 only
        const float *src = a.logits + ((size_t)pb * T + (size_t)pj * P) * V;
        mbar_expect_tx(bar0 + 8 * st, group_bytes);
        tma_load_1d_hint(slot0 + st * slot_bytes, src, group_bytes, bar0 + 8 * st, kEvictFirst);
        ++issued;
        if (++pj >= pngb) {
            pj = 0;
            do { ++pb; } while (pb < B && a.gstart[pb + 1] == a.gstart[pb]);
            if (pb < B) pngb = a.gstart[pb + 1] - a.gstart[pb];
        }
    };
    if (tid == 0) while (issued < nst && issued < ng) issue(issued);
    int stage = 0;
    uint32_t parity = 0;
    constexpr int MAXQ = (2 * 264 + NT - 1) / NT;        // label loads per thread: P frames x Lp slots over 128 threads
    for (int i = 0; i < ng; ++i) {
        if (b != cur_b) {                    // block-uniform: (re)load the utterance's class ids
            cur_b = b;
            Tbb = a.Tb[b];
            const int Ub = a.Ub[b];
            if (FUSED) gsc = a.reduction == 1 ? a.inv_batch * __frcp_rn((float)(Ub > 1 ? Ub : 1)) : 1.f;
            const int64_t toff = a.toff[b];
            __syncthreads();                 // the previous utterance's table reads are done
            for (int k = tid; k < Lp; k += NT) {
                int cls;
                if (k == 0) cls = a.blank;
                else if (k == 1) cls = -2;   // slot of lse2
                else if (k < 4) cls = -3;    // unused header slots -> 0
                else if (k - 4 < Ub) {
                    const int64_t idx = toff + (k - 4);
                    long long c = idx < a.tnumel ? a.targets[idx] : -1;
                    if (c < 0 || c >= V || c == a.blank) {
                        atomicOr(&a.hdr[0], 4);
                        if (c < 0 || c >= V) atomicOr(&a.bad[b], 4);
                        c = c < 0 ? 0 : (c >= V ? V - 1 : c);
                    }
                    cls = (int)c;
                } else cls = -1;             // beyond U_b -> sentinel
                cls_s[k] = cls;
            }
            __syncthreads();
        }
        const int t0 = j * P;
        const bool live1 = P == 2 && t0 + 1 < Tbb;       // (the odd frame of an utterance's last group may be padding)
        mbar_wait(bar0 + 8 * stage, parity);                     // the group has landed in its slot
        const float *grow = (const float *)(smem_p + (size_t)stage * slot_bytes);
        const float4 *g4 = (const float4 *)grow;
        // ---- the group: shared memory -> registers ----
        float4 v[MAXC];
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
            const int c = tid + k * NT;
            v[k] = make_float4(CTC_NEG_INF, CTC_NEG_INF, CTC_NEG_INF, CTC_NEG_INF);
            if (c < nch) v[k] = g4[c];
        }
        float xg[MAXQ];                      // label logits: slot q of frame r  <-  item tid + kk*NT = r*Lp + q
#pragma unroll
        for (int kk = 0; kk < MAXQ; ++kk) {
            const int it = tid + kk * NT;
            xg[kk] = 0.f;
            if (it < P * Lp) {
                const int r = it >= Lp, q = it - r * Lp, c = cls_s[q];
                if (c >= 0 && (r == 0 || live1)) xg[kk] = grow[r * V + c];
            }
        }
        // ---- per-frame max over this warp, then 2^(x - max_w) and its sum: no other warp needed ----
        float mx0 = CTC_NEG_INF, mx1 = CTC_NEG_INF;
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
            const int c = tid + k * NT;
            const float lo2 = fmaxf(v[k].x, v[k].y), hi2 = fmaxf(v[k].z, v[k].w);
            if (c < mid) mx0 = fmaxf(mx0, fmaxf(lo2, hi2));
            else if (c > mid) mx1 = fmaxf(mx1, fmaxf(lo2, hi2));
            else { mx0 = fmaxf(mx0, lo2); mx1 = fmaxf(mx1, hi2); }
        }
        const float mw0 = warp_max(mx0), mw1 = warp_max(mx1);
        const float e0 = (mw0 == CTC_NEG_INF ? 0.f : mw0) * kLog2e, e1 = (mw1 == CTC_NEG_INF ? 0.f : mw1) * kLog2e;
        float *rd = &red[i & 1][warp][0];
        if (a.best != nullptr) {             // lowest class index attaining the warp maximum, per frame
            int c0 = 0x7fffffff, c1 = 0x7fffffff;
#pragma unroll
            for (int k = MAXC - 1; k >= 0; --k) {
                const int c = tid + k * NT, e = 4 * c;                      // float index within the group
                if (c < nch) {
                    const bool f0lo = c <= mid, f0hi = c < mid;            // does .xy / .zw belong to frame 0?
                    if (f0hi ? v[k].w == mw0 : v[k].w == mw1) { if (f0hi) c0 = e + 3; else c1 = e + 3 - V; }
                    if (f0hi ? v[k].z == mw0 : v[k].z == mw1) { if (f0hi) c0 = e + 2; else c1 = e + 2 - V; }
                    if (f0lo ? v[k].y == mw0 : v[k].y == mw1) { if (f0lo) c0 = e + 1; else c1 = e + 1 - V; }
                    if (f0lo ? v[k].x == mw0 : v[k].x == mw1) { if (f0lo) c0 = e; else c1 = e - V; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                c0 = min(c0, __shfl_xor_sync(0xffffffffu, c0, o));
                c1 = min(c1, __shfl_xor_sync(0xffffffffu, c1, o));
            }
            if (lane == 0) { redi[i & 1][warp][0] = c0; redi[i & 1][warp][1] = c1; }
        }
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
            const int c = tid + k * NT;
            const float ba = c <= mid ? e0 : e1, bb2 = c < mid ? e0 : e1;   // exponent base of .xy / .zw
            v[k].x = ex2f(fmaf(v[k].x, kLog2e, -ba)); v[k].y = ex2f(fmaf(v[k].y, kLog2e, -ba));
            v[k].z = ex2f(fmaf(v[k].z, kLog2e, -bb2)); v[k].w = ex2f(fmaf(v[k].w, kLog2e, -bb2));
            const float lo2 = v[k].x + v[k].y, hi2 = v[k].z + v[k].w;
            if (c < mid) s0 += lo2 + hi2;
            else if (c > mid) s1 += lo2 + hi2;
            else { s0 += lo2; s1 += hi2; }
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        if (lane == 0) { rd[0] = mw0; rd[1] = mw1; rd[2] = s0; rd[3] = s1; }
        __syncthreads();                     // the group's only barrier: partials visible, slot consumed
        const int cur_stage = stage;
        if (!(BULKST && FUSED) && tid == 0 && issued < ng) issue(stage);   // refill the slot with group i + nst
        if (++stage == nst) { stage = 0; parity ^= 1; }
        float M0 = CTC_NEG_INF, M1 = CTC_NEG_INF;
#pragma unroll
        for (int w = 0; w < NW; ++w) { M0 = fmaxf(M0, red[i & 1][w][0]); M1 = fmaxf(M1, red[i & 1][w][1]); }
        const float M0l = M0 * kLog2e, M1l = M1 * kLog2e;
        float S0 = 0.f, S1 = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float a0 = red[i & 1][w][0], a1 = red[i & 1][w][1];
            S0 += red[i & 1][w][2] * ex2f(fmaf(a0 == CTC_NEG_INF ? 0.f : a0, kLog2e, -M0l));
            S1 += red[i & 1][w][3] * ex2f(fmaf(a1 == CTC_NEG_INF ? 0.f : a1, kLog2e, -M1l));
        }
        const float lse0 = M0l + lg2f(S0), lse1 = M1l + lg2f(S1);
        if (a.best != nullptr && tid < P && (tid == 0 || live1)) {
            int bi = 0x7fffffff;
#pragma unroll
            for (int w = 0; w < NW; ++w) if (red[i & 1][w][tid] == (tid ? M1 : M0)) bi = min(bi, redi[i & 1][w][tid]);
            a.best[(size_t)b * T + t0 + tid] = bi;
        }
        // ---- the frames for the lattice kernel (same format as k1_lse_gather) ----
#pragma unroll
        for (int kk = 0; kk < MAXQ; ++kk) {
            const int it = tid + kk * NT;
            if (it < P * Lp) {
                const int r = it >= Lp, q = it - r * Lp, c = cls_s[q];
                if (r == 0 || live1) {
                    const float lse2 = r ? lse1 : lse0;
                    float o;
                    if (c >= 0) {
                        o = fminf(fmaxf(fmaf(xg[kk], kLog2e, -lse2), kNeg), 0.f);
                        if (o >= a.lin_thr) o = ex2f(o);
                        else a.slow[b] = 1;                                // (also NaN)
                    } else {
                        o = c == -2 ? lse2 : (c == -3 ? 0.f : kNeg);
                    }
                    stg_f32_hint(a.lp_lab + ((size_t)b * T + t0 + r) * Lp + q, o, kScratch);   // re-read by the lattice
                }
            }
        }
        if (FUSED) {
            // ---- dense gradient g * softmax = 2^(x - max_w) * 2^(max_w - M) * g / S: aligned 16-byte stores ----
            const float f0 = gsc * ex2f(e0 - M0l) * __frcp_rn(S0);
            const float f1 = live1 ? gsc * ex2f(e1 - M1l) * __frcp_rn(S1) : 0.f;   // a padded odd frame gets its zeros here
            float4 *o4 = (float4 *)(a.grad + ((size_t)b * T + t0) * V);
            if (BULKST) {
                float4 *w4 = (float4 *)(smem_p + (size_t)cur_stage * slot_bytes);   // in place: this group's own slot
#pragma unroll
                for (int k = 0; k < MAXC; ++k) {
                    const int c = tid + k * NT;
                    if (c < nch) {
                        const float fa = c <= mid ? f0 : f1, fb = c < mid ? f0 : f1;     // (f1 = 0 for a padded odd frame)
                        w4[c] = make_float4(v[k].x * fa, v[k].y * fa, v[k].z * fb, v[k].w * fb);
                    }
                }
                fence_proxy_async_smem_cta();                    // these writes -> the bulk store's (async proxy) reads
                __syncthreads();                                 // second barrier of the group: the slot is complete
                if (tid == 0) {
                    tma_store_1d_hint(o4, slot0 + cur_stage * slot_bytes, group_bytes, c_gstore_policy);
                    bulk_commit_group();
                    if (i >= 1 && issued < ng) {                 // slot of group i-1: its store has read it by now
                        bulk_wait_group_read<1>();
                        issue(issued % nst);
                    }
                }
            } else {
#pragma unroll
            for (int k = 0; k < MAXC; ++k) {
                const int c = tid + k * NT;
                if (c < nch && (c <= mid || live1)) {
                    const float fa = c <= mid ? f0 : f1, fb = c < mid ? f0 : f1;
                    const bool hi_live = c < mid || live1;                 // (upper half of the middle chunk: frame 1)
                    stg_v4_cs(o4 + c, make_float4(v[k].x * fa, v[k].y * fa, hi_live ? v[k].z * fb : 0.f,
                                                  hi_live ? v[k].w * fb : 0.f));
                }
            }
            }
        }
        if (++j >= ngb) {                    // next group
            j = 0;
            do { ++b; } while (b < B && a.gstart[b + 1] == a.gstart[b]);
            if (b < B) ngb = a.gstart[b + 1] - a.gstart[b];
        }
    }
    if (BULKST && FUSED && tid == 0) bulk_wait_group<0>();      // shared memory must outlive the bulk stores
}

}  // namespace ctcb200
