// k1w: the sweep of round 2 -- "one warp per frame, frames staged in aligned groups".
//
// Same job as k1_lse_gather (stream_kernels.cuh): ONE read of every valid logits frame -> per-frame log-sum-exp and the
// U+1 label log-probabilities the lattice needs, and (FUSED) the dense part of the gradient g_b * softmax(x) written in
// the logits' layout.  What changed, and why (round-1 verdict: 0.86 of the measured copy bandwidth, 1.2-1.4 barrier
// stalls per issued instruction, a 16-byte hull copy + two scalar edge threads per frame because a frame of V = 4234
// floats starts 8 bytes off every other row):
//   * frames are fetched in ALIGNED GROUPS of P = 4 / gcd(V, 4) consecutive frames (P = 2 for V = 4234: 33 872 bytes,
//     a multiple of 16 that starts 16-byte aligned), one 1-D bulk TMA copy per group, no hull, nothing fetched twice;
//   * ONE WARP owns a frame from the first byte to the last: max, sum and the gradient are three passes over the
//     frame in shared memory with warp shuffles in between -- there is no block barrier anywhere in the loop, and no
//     thread ever waits for another warp's frame;
//   * the 2^(x-max) values are written back into the slot in pass 2 and rescaled in pass 3 (one MUFU per element);
//   * within a group, a frame is [aligned float4 chunks] plus one float2 (the last two floats of the even frame, the
//     first two of the odd one): every global store is a naturally aligned STG.128 / STG.64, no scalar edge code.
// The ring of group slots is self-service (mbarrier full/empty per slot): P warps share a group, one frame each; the
// first of them refills the slot with the group `nslot` ahead as soon as both have their frame in registers.  There is
// no producer warp: 8 warps = 2 per scheduler, so each may use up to 255 registers (the register file is per
// scheduler: a ninth warp would cap everyone at 168).
//
// Used when V % 4 is 0 or 2, T % P == 0 and a group fits the ring; otherwise ctcb200.cu falls back to k1_lse_gather.
#pragma once
#include "lattice_kernel.cuh"
#include "stream_kernels.cuh"

namespace ctcb200 {

constexpr int KW_MAX_CONSUMERS = 8;

struct K1wArgs {
    const float *logits; const int64_t *targets; int64_t tnumel; const int *Tb, *Ub; const int64_t *toff;
    const int *rowstart, *gstart; float *lp_lab; int *hdr; int B, T, V, Lp, blank, P;
    int nslot; uint32_t slot_bytes; int nw;             // ring geometry, consumer warps
    float *grad; int reduction; float inv_batch;       // FUSED only
    int *best; int zero_pad_here; int *slow; float lin_thr; int *bad;
};

__device__ __forceinline__ void sts_v4(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_v2(uint32_t a, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
// generic-proxy accesses to shared memory -> later async-proxy (bulk copy) writes of the same bytes
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// NCH = float4 chunks a lane keeps in registers: 32 * NCH >= aligned chunks of a frame ((V - 2) / 4 or V / 4).
template <int NCH, bool FUSED>
__global__ void __launch_bounds__(32 * KW_MAX_CONSUMERS, 1) k1w_sweep(const K1wArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x;
    griddep_wait();                                  // k0_prep's lengths / prefix sums
    griddep_launch_dependents();
    const int B = a.B, T = a.T, V = a.V, P = a.P, Lp = a.Lp, NW = a.nw;
    if (FUSED && a.zero_pad_here) {
        // padded frames (t >= T_b) of the whole batch -> zeros; every CTA takes an equal share
        if (nthreads == 256) zero_padded_frames<256>(a.grad, a.Tb, a.rowstart, B, T, V, tid);
        else if (nthreads == 192) zero_padded_frames<192>(a.grad, a.Tb, a.rowstart, B, T, V, tid);
        else if (nthreads == 128) zero_padded_frames<128>(a.grad, a.Tb, a.rowstart, B, T, V, tid);
        else zero_padded_frames<64>(a.grad, a.Tb, a.rowstart, B, T, V, tid);
    }
    int g0, ng;
    grid_share(a.gstart[B], g0, ng);                 // this CTA's contiguous share of the live groups
    if (ng <= 0) return;

    const uint32_t slot0 = smem_u32(smem);
    const uint32_t bars = slot0 + (uint32_t)a.nslot * a.slot_bytes;          // full[nslot], empty[nslot]
    int *cls_all = (int *)(smem + (size_t)a.nslot * a.slot_bytes + 16 * a.nslot);   // [NW][Lp] class id per frame slot
    if (tid == 0) {
        // a slot is handed back by the P warps that each took one frame of the group out of it
        for (int s = 0; s < a.nslot; ++s) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 8 * (a.nslot + s), P); }
        fence_mbar_init();
    }
    __syncthreads();                                 // the only block barrier of the kernel

    // group cursor: flattened live-group index <-> (utterance b, group j of it); group j holds frames jP .. jP+P-1
    auto seek = [&](int gidx, int &b, int &j) {
        int lo = 0, hi = B - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (a.gstart[mid + 1] > gidx) hi = mid; else lo = mid + 1;
        }
        b = lo; j = gidx - a.gstart[lo];
    };
    const uint32_t group_bytes = (uint32_t)P * (uint32_t)V * 4u;             // multiple of 16 by the choice of P

    // issue the bulk copy of live group `gi` (index within this CTA's share) into its slot; one lane
    auto issue_group = [&](int gi) {
        int b, j;
        seek(g0 + gi, b, j);
        const int s = gi % a.nslot;
        const float *src = a.logits + ((size_t)b * T + (size_t)j * P) * V;
        mbar_expect_tx(bars + 8 * s, group_bytes);
        tma_load_1d_hint(slot0 + s * a.slot_bytes, src, group_bytes, bars + 8 * s, kEvictFirst);
    };
    if (warp >= NW) return;
    if (warp == 0 && lane == 0)
        for (int gi = 0; gi < a.nslot && gi < ng; ++gi) issue_group(gi);     // prologue: fill the ring

    // ===================== consumers: P warps per group stream, one frame each =====================
    // stream q = warp / P takes the groups q, q + NS, ... (NS = NW / P streams); warp r = warp % P of the stream
    // owns frame r of each of them.  The frame goes from the slot into registers in ONE pass of LDS.128 and the slot
    // is handed back at once: the ring is almost entirely in flight, the arithmetic runs out of registers.
    const int NS = NW / P, q = warp / P, r = warp - q * P;
    int *cls_s = cls_all + warp * Lp;
    int cur_b = -1, Tbb = 0, Ub = 0;
    float g = 0.f;
    const int nfull = (V - (P == 2 ? 2 : 0)) >> 2;                           // aligned float4 chunks per frame
    const int c0 = (P == 2 && r == 1) ? 2 : 0;                               // first float of the aligned chunks
    const int h2 = (P == 2) ? (r == 0 ? V - 2 : 0) : -1;                     // the two floats that fill no chunk
    const float4 ninf4 = make_float4(CTC_NEG_INF, CTC_NEG_INF, CTC_NEG_INF, CTC_NEG_INF);
    for (int i = q; i < ng; i += NS) {
        int b, j;
        seek(g0 + i, b, j);
        if (b != cur_b) {                                                    // (re)load this warp's class table
            cur_b = b;
            Tbb = a.Tb[b]; Ub = a.Ub[b];
            if (FUSED) g = a.reduction == 1 ? a.inv_batch * __frcp_rn((float)(Ub > 1 ? Ub : 1)) : 1.f;
            const int64_t toff = a.toff[b];
            __syncwarp();
            for (int k = lane; k < Lp; k += 32) {
                int cls;
                if (k == 0) cls = a.blank;
                else if (k == 1) cls = -2;                                   // slot of lse2
                else if (k < 4) cls = -3;                                    // unused header slots -> 0
                else if (k - 4 < Ub) {
                    const int64_t idx = toff + (k - 4);
                    long long c = idx < a.tnumel ? a.targets[idx] : -1;
                    if (c < 0 || c >= V || c == a.blank) {
                        atomicOr(&a.hdr[0], 4);
                        if (c < 0 || c >= V) atomicOr(&a.bad[b], 4);
                        c = c < 0 ? 0 : (c >= V ? V - 1 : c);
                    }
                    cls = (int)c;
                } else cls = -1;                                             // beyond U_b -> sentinel
                cls_s[k] = cls;
            }
            __syncwarp();
        }
        const int s = i % a.nslot;
        mbar_wait_bounded(bars + 8 * s, (uint32_t)(i / a.nslot) & 1u);
        const int t = j * P + r;
        const bool live = t < Tbb;                                           // (the odd frame of the last group may be padding)
        // ---- the frame: slot -> registers (one pass), label logits gathered on the way ----
        const unsigned char *fb = smem + (size_t)s * a.slot_bytes + (size_t)r * V * 4;
        const float4 *f4 = (const float4 *)(fb + 4 * c0);
        float4 v[NCH];
        float2 xh = make_float2(CTC_NEG_INF, CTC_NEG_INF);
        constexpr int MAXG = (260 + 31) / 32;
        float xg[MAXG];
        if (live) {
#pragma unroll
            for (int k = 0; k < NCH; ++k) v[k] = lane + 32 * k < nfull ? f4[lane + 32 * k] : ninf4;
            if (h2 >= 0 && lane == 31) xh = *(const float2 *)(fb + 4 * (size_t)h2);
#pragma unroll
            for (int kk = 0; kk < MAXG; ++kk) {
                const int k = lane + 32 * kk;
                xg[kk] = 0.f;
                if (k < Lp) { const int c = cls_s[k]; if (c >= 0) xg[kk] = ((const float *)fb)[c]; }
            }
        }
        // hand the slot back: this warp's (generic-proxy) reads of it are ordered before the next bulk copy into it
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(bars + 8 * (a.nslot + s));
            if (r == 0 && i + a.nslot < ng) {                               // refill: the group one ring lap ahead
                mbar_wait_bounded(bars + 8 * (a.nslot + s), (uint32_t)(i / a.nslot) & 1u);   // all P warps are out
                issue_group(i + a.nslot);
            }
        }
        if (!live) continue;

        // ---- max (and arg max for the greedy decode) ----
        float mx = fmaxf(xh.x, xh.y);
#pragma unroll
        for (int k = 0; k < NCH; ++k) mx = fmaxf(mx, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
        const float m = warp_max(mx);
        if (a.best != nullptr) {                     // lowest class index attaining the frame maximum
            int cand = 0x7fffffff;
#pragma unroll
            for (int k = NCH - 1; k >= 0; --k) {
                const int e = c0 + 4 * (lane + 32 * k);
                if (v[k].w == m) cand = e + 3;
                if (v[k].z == m) cand = e + 2;
                if (v[k].y == m) cand = e + 1;
                if (v[k].x == m) cand = e;
            }
            if (h2 >= 0 && lane == 31) {
                if (xh.y == m) cand = min(cand, h2 + 1);
                if (xh.x == m) cand = min(cand, h2);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
            if (lane == 0) a.best[(size_t)b * T + t] = cand;
        }
        const float m2 = m * kLog2e;
        // ---- e = 2^(x - max), sum ----
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            v[k].x = ex2f(fmaf(v[k].x, kLog2e, -m2)); v[k].y = ex2f(fmaf(v[k].y, kLog2e, -m2));
            v[k].z = ex2f(fmaf(v[k].z, kLog2e, -m2)); v[k].w = ex2f(fmaf(v[k].w, kLog2e, -m2));
            sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
        }
        xh.x = ex2f(fmaf(xh.x, kLog2e, -m2)); xh.y = ex2f(fmaf(xh.y, kLog2e, -m2));
        sum += xh.x + xh.y;                          // (0 for the lanes / layouts without the float2)
        const float tot = warp_sum(sum);
        const float lse2 = m2 + lg2f(tot);
        // ---- the frame for the lattice kernel (same format as k1_lse_gather) ----
        float *frame = a.lp_lab + ((size_t)b * T + t) * Lp;
#pragma unroll
        for (int kk = 0; kk < MAXG; ++kk) {
            const int k = lane + 32 * kk;
            if (k < Lp) {
                const int c = cls_s[k];
                float o;
                if (c >= 0) {
                    o = fminf(fmaxf(fmaf(xg[kk], kLog2e, -lse2), kNeg), 0.f);
                    if (o >= a.lin_thr) o = ex2f(o);
                    else a.slow[b] = 1;                                      // (also NaN)
                } else {
                    o = c == -2 ? lse2 : (c == -3 ? 0.f : kNeg);
                }
                stg_f32_hint(frame + k, o, kEvictLast);                      // re-read by the lattice kernel
            }
        }
        if (FUSED) {
            // ---- dense gradient g * softmax = e * g / sum: naturally aligned 16-byte (8-byte) stores ----
            const float sc = g * __frcp_rn(tot);
            float *orow = a.grad + ((size_t)b * T + t) * V;
            float4 *o4 = (float4 *)(orow + c0);
#pragma unroll
            for (int k = 0; k < NCH; ++k)
                if (lane + 32 * k < nfull)
                    stg_v4_hint(o4 + lane + 32 * k, make_float4(v[k].x * sc, v[k].y * sc, v[k].z * sc, v[k].w * sc), kEvictFirst);
            if (h2 >= 0 && lane == 31) stg_v2_hint((float2 *)(orow + h2), make_float2(xh.x * sc, xh.y * sc), kEvictFirst);
        }
    }
}

}  // namespace ctcb200
