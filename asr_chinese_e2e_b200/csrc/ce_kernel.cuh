// Attention-branch loss on the same machinery: cross-entropy with optional label
// smoothing over pred[N*L, V] logits with PAD-ignored targets, exactly the reference's
// cal_loss / calculate_loss (Predictor/Utils/loss.py:26-76):
//   eps == 0:  F.cross_entropy(pred, gold, ignore_index, 'mean')
//   eps  > 0:  target = (1-eps) on the label, eps/C elsewhere;  loss = sum_rows(-sum_c target_c*logp_c) / n_word
// One sweep: every non-pad row is read once through the bulk-TMA ring, its loss is written to rowloss[]
// and (GRAD) its gradient  w/n_word * (softmax * sum(target) - target)  is stored from the registers that
// hold the row; pad rows get zeros without being read.  n_word is counted on the device (no host sync).
#pragma once
#include "stream_kernels.cuh"

namespace ctcb200 {

// kce_prep: one CTA.  Compact lists of non-pad / pad row indices, n_word.
//   hdr[0] = n_word, hdr[1] = ticket (unused), vlist[0..n_word) valid rows (ascending), plist the rest.
__global__ void __launch_bounds__(1024) kce_prep(const int64_t *__restrict__ gold, int rows, int ignore_index,
                                                 int *__restrict__ hdr, int *__restrict__ vlist,
                                                 int *__restrict__ plist) {
    __shared__ int s_part[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < rows; base += 1024) {
        const int r = base + tid;
        const int valid = (r < rows) && (gold[r] != ignore_index);
        int inc = valid;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += a;
        }
        if (lane == 31) s_part[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int p = s_part[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int a = __shfl_up_sync(0xffffffffu, p, o);
                if (lane >= o) p += a;
            }
            s_part[lane] = p;
        }
        __syncthreads();
        const int carry = s_carry;
        const int before = carry + (warp ? s_part[warp - 1] : 0) + inc - valid;   // valid rows before r
        if (r < rows) {
            if (valid) vlist[before] = r;
            else plist[r - before] = r;
        }
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_part[31];
        __syncthreads();
    }
    if (tid == 0) { hdr[0] = s_carry; hdr[1] = 0; }
}

template <int NT, int MAXC, bool EXACT, bool GRAD>
__global__ void __launch_bounds__(NT)
kce_rows(const float *__restrict__ pred, const int64_t *__restrict__ gold, int rows, int V,
         const int *__restrict__ hdr, const int *__restrict__ vlist, const int *__restrict__ plist,
         float *__restrict__ rowloss, float *__restrict__ grad, float eps, float weight, int nst,
         uint32_t slot_bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = hdr[0];
    if (GRAD) {                                      // pad rows: zero gradient, never read
        int z0, zc;
        grid_share(rows - nw, z0, zc);
        for (int i = 0; i < zc; ++i) zero_span<NT>(grad + (size_t)plist[z0 + i] * V, (size_t)V, tid);
    }
    int r0, nrows;
    grid_share(nw, r0, nrows);
    if (nrows <= 0) return;

    uint64_t *bars = (uint64_t *)(smem + (size_t)nst * slot_bytes);
    float *red = (float *)(bars + nst);              // [2 parity][3 max/sum/sumx][4 warps]
    const uint32_t slot0 = smem_u32(smem), bar0 = smem_u32(bars);
    if (tid == 0) {
        for (int s = 0; s < nst; ++s) mbar_init(bar0 + 8 * s, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const uintptr_t end16 = ((uintptr_t)pred + (size_t)rows * V * 4) & ~(uintptr_t)15;
    int issued = 0;
    if (tid == 0)
        for (; issued < nst && issued < nrows; ++issued)
            issue_row(pred + (size_t)vlist[r0 + issued] * V, V, end16, slot0 + issued * slot_bytes, bar0 + 8 * issued, 0);

    const float invC = 1.f / (float)V;
    const float tsum = 1.f - eps * invC;             // sum_c target_c  (1 when eps == 0)
    const float s = weight / (float)nw;              // gradient scale (speculative upstream gradient 1)
    int stage = 0;
    uint32_t parity = 0;
    for (int i = 0; i < nrows; ++i) {
        const int row = vlist[r0 + i];
        const int g = (int)gold[row];
        mbar_wait(bar0 + 8 * stage, parity);
        const float *grow = pred + (size_t)row * V;
        const int head = (int)(((uintptr_t)grow & 15) >> 2);
        const int nch = (head + V + 3) >> 2;
        const unsigned char *slot = smem + (size_t)stage * slot_bytes;
        const float4 *s4 = (const float4 *)slot;
        const float xg = ((const float *)slot)[head + (g < 0 ? 0 : (g >= V ? V - 1 : g))];

        float4 v[MAXC];
        float mx = CTC_NEG_INF, sx = 0.f;
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
            const int c = 1 + tid + k * NT;
            float4 x;
            if ((EXACT && k < MAXC - 1) || c <= nch - 2) {
                x = s4[c];
                sx += (x.x + x.y) + (x.z + x.w);
            } else {
                x = make_float4(CTC_NEG_INF, CTC_NEG_INF, CTC_NEG_INF, CTC_NEG_INF);
            }
            mx = fmaxf(mx, fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w)));
            v[k] = x;
        }
        float4 ve = make_float4(CTC_NEG_INF, CTC_NEG_INF, CTC_NEG_INF, CTC_NEG_INF);
        if (tid == 0 || (tid == 1 && nch > 1)) {
            const int c = tid == 0 ? 0 : nch - 1;
            ve = s4[c];
            const int e = 4 * c - head;
            if (e < 0 || e >= V) ve.x = CTC_NEG_INF; else sx += ve.x;
            if (e + 1 < 0 || e + 1 >= V) ve.y = CTC_NEG_INF; else sx += ve.y;
            if (e + 2 < 0 || e + 2 >= V) ve.z = CTC_NEG_INF; else sx += ve.z;
            if (e + 3 < 0 || e + 3 >= V) ve.w = CTC_NEG_INF; else sx += ve.w;
            mx = fmaxf(mx, fmaxf(fmaxf(ve.x, ve.y), fmaxf(ve.z, ve.w)));
        }
        float *rd = red + (i & 1) * 12;
        mx = warp_max(mx);
        if (lane == 0) rd[warp] = mx;
        __syncthreads();                                   // B1: slot fully consumed
        if (tid == 0 && issued < nrows) {
            issue_row(pred + (size_t)vlist[r0 + issued] * V, V, end16, slot0 + stage * slot_bytes, bar0 + 8 * stage, 0);
            ++issued;
        }
        const float m = NT == 128 ? fmaxf(fmaxf(rd[0], rd[1]), fmaxf(rd[2], rd[3])) : fmaxf(rd[0], rd[1]);
        const float m2 = m * kLog2e;
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
            v[k].x = ex2f(fmaf(v[k].x, kLog2e, -m2)); v[k].y = ex2f(fmaf(v[k].y, kLog2e, -m2));
            v[k].z = ex2f(fmaf(v[k].z, kLog2e, -m2)); v[k].w = ex2f(fmaf(v[k].w, kLog2e, -m2));
            sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
        }
        if (warp == 0) {
            ve.x = ex2f(fmaf(ve.x, kLog2e, -m2)); ve.y = ex2f(fmaf(ve.y, kLog2e, -m2));
            ve.z = ex2f(fmaf(ve.z, kLog2e, -m2)); ve.w = ex2f(fmaf(ve.w, kLog2e, -m2));
            sum += (ve.x + ve.y) + (ve.z + ve.w);
        }
        sum = warp_sum(sum);
        sx = warp_sum(sx);
        if (lane == 0) { rd[4 + warp] = sum; rd[8 + warp] = sx; }
        __syncthreads();                                   // B2
        const float tot = NT == 128 ? (rd[4] + rd[5]) + (rd[6] + rd[7]) : rd[4] + rd[5];
        const float totx = NT == 128 ? (rd[8] + rd[9]) + (rd[10] + rd[11]) : rd[8] + rd[9];
        if (tid == 0) {
            const float lse = (m2 + lg2f(tot)) * kLn2;
            rowloss[row] = tsum * lse - (1.f - eps) * xg - eps * invC * (totx - xg);
        }
        if (GRAD) {
            // grad_c = s * (softmax_c * tsum - target_c);  softmax_c = 2^(x_c - max) / tot
            const float A = s * tsum * __frcp_rn(tot), Bc = s * eps * invC;
            const float gfix = s * (1.f - eps) - Bc;       // extra amount subtracted at the label column
            const int gchunk = (head + g) >> 2, gsub = (head + g) & 3;
            float *orow = grad + (size_t)row * V;
            float4 *g4 = (float4 *)((uintptr_t)orow & ~(uintptr_t)15);
#pragma unroll
            for (int k = 0; k < MAXC; ++k) {
                const int c = 1 + tid + k * NT;
                if ((EXACT && k < MAXC - 1) || c <= nch - 2) {
                    float4 y = make_float4(fmaf(v[k].x, A, -Bc), fmaf(v[k].y, A, -Bc), fmaf(v[k].z, A, -Bc),
                                           fmaf(v[k].w, A, -Bc));
                    if (c == gchunk) {
                        if (gsub == 0) y.x -= gfix; else if (gsub == 1) y.y -= gfix;
                        else if (gsub == 2) y.z -= gfix; else y.w -= gfix;
                    }
                    stg_v4_hint(g4 + c, y, kEvictFirst);
                }
            }
            if (tid == 0 || (tid == 1 && nch > 1)) {
                const int c = tid == 0 ? 0 : nch - 1;
                const int e = 4 * c - head;
                float y0 = fmaf(ve.x, A, -Bc), y1 = fmaf(ve.y, A, -Bc), y2 = fmaf(ve.z, A, -Bc), y3 = fmaf(ve.w, A, -Bc);
                if (c == gchunk) {
                    if (gsub == 0) y0 -= gfix; else if (gsub == 1) y1 -= gfix;
                    else if (gsub == 2) y2 -= gfix; else y3 -= gfix;
                }
                if (e >= 0 && e < V) orow[e] = y0;
                if (e + 1 >= 0 && e + 1 < V) orow[e + 1] = y1;
                if (e + 2 >= 0 && e + 2 < V) orow[e + 2] = y2;
                if (e + 3 >= 0 && e + 3 < V) orow[e + 3] = y3;
            }
        }
        if (++stage == nst) { stage = 0; parity ^= 1; }
    }
}

// kce_finish: one CTA, deterministic sum of the row losses in row order -> out[0] = weight * mean, out[1] = n_word
__global__ void __launch_bounds__(1024) kce_finish(const int *__restrict__ hdr, const int *__restrict__ vlist,
                                                   const float *__restrict__ rowloss, float weight,
                                                   float *__restrict__ out) {
    __shared__ float s_part[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = hdr[0];
    float acc = 0.f;
    for (int i = tid; i < nw; i += 1024) acc += rowloss[vlist[i]];
    acc = warp_sum(acc);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        float p = s_part[lane];
        p = warp_sum(p);
        if (lane == 0) { out[0] = weight * p / (float)nw; out[1] = (float)nw; }
    }
}

}  // namespace ctcb200
