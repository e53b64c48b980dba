// Developer probe (GPU box): achievable HBM bandwidth of the TMA-ring row streaming pattern used by
// k1/k3, as a function of how rows are dealt to the persistent CTAs.
//   mode 0: contiguous share per CTA   mode 1: row-interleaved (row = i*G + cta)   mode 2: block-cyclic (R rows)
// op 0: read-only (sum)   op 1: read + write (scaled copy)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bw_probe tools/bw_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../asr_chinese_e2e_b200/csrc/ptx.cuh"
using namespace ctcb200;

__global__ void __launch_bounds__(128) probe(const float *__restrict__ in, float *__restrict__ out, float *sink,
                                             int rows, int V, int mode, int R, int op, int nst, uint32_t slot_bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, G = gridDim.x, bid = blockIdx.x;
    uint64_t *bars = (uint64_t *)(smem + (size_t)nst * slot_bytes);
    const uint32_t slot0 = smem_u32(smem), bar0 = smem_u32(bars);
    int n;   // rows of this CTA
    int base = rows / G, rem = rows - base * G, first = bid * base + (bid < rem ? bid : rem);
    if (mode == 0) n = base + (bid < rem);
    else if (mode == 1) n = (rows - bid + G - 1) / G;
    else { int nb = (rows + R - 1) / R; int mb = (nb - bid + G - 1) / G; n = 0; for (int j = 0; j < mb; ++j) { int r0 = (bid + j * G) * R; n += min(R, rows - r0); } }
    auto row_of = [&](int i) -> int {
        if (mode == 0) return first + i;
        if (mode == 1) return bid + i * G;
        return (bid + (i / R) * G) * R + (i % R);   // exact when all blocks are full (rows % R == 0)
    };
    if (tid == 0) { for (int s = 0; s < nst; ++s) mbar_init(bar0 + 8 * s, 1); fence_mbar_init(); }
    __syncthreads();
    const uint32_t rb = (uint32_t)V * 4;   // V*4 multiple of 16 in the probe
    int issued = 0;
    if (tid == 0) for (; issued < nst && issued < n; ++issued) {
        mbar_expect_tx(bar0 + 8 * issued, rb);
        tma_load_1d(slot0 + issued * slot_bytes, in + (size_t)row_of(issued) * V, rb, bar0 + 8 * issued);
    }
    int stage = 0; uint32_t parity = 0; float acc = 0.f;
    const int nch = V / 4;
    for (int i = 0; i < n; ++i) {
        mbar_wait(bar0 + 8 * stage, parity);
        const float4 *s4 = (const float4 *)(smem + (size_t)stage * slot_bytes);
        float4 *o4 = (float4 *)(out + (size_t)row_of(i) * V);
        if (op <= 1) {
            for (int c = tid; c < nch; c += 128) {
                float4 x = s4[c];
                if (op) { x.x *= 1.5f; x.y *= 1.5f; x.z *= 1.5f; x.w *= 1.5f; o4[c] = x; }
                else acc += x.x + x.y + x.z + x.w;
            }
            __syncthreads();
        } else {
            // k1-like: row in registers, max, (op>=3: block reduce + barrier), exp-sum (op 2,4: MUFU; op 3: FADD only)
            __shared__ float red[2][8];
            float4 v[9];
            float mx = -1e30f;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                const int c = tid + 128 * k;
                v[k] = c < nch ? s4[c] : make_float4(-1e30f, -1e30f, -1e30f, -1e30f);
                mx = fmaxf(mx, fmaxf(fmaxf(v[k].x, v[k].y), fmaxf(v[k].z, v[k].w)));
            }
            float *rd = red[i & 1];
            if (op >= 3) {
                mx = warp_max(mx);
                if ((tid & 31) == 0) rd[tid >> 5] = mx;
            }
            __syncthreads();
            if (tid == 0 && issued < n) {
                mbar_expect_tx(bar0 + 8 * stage, rb);
                tma_load_1d(slot0 + stage * slot_bytes, in + (size_t)row_of(issued) * V, rb, bar0 + 8 * stage);
                ++issued;
            }
            if (op >= 3) mx = fmaxf(fmaxf(rd[0], rd[1]), fmaxf(rd[2], rd[3]));
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (op == 3) sum += (v[k].x - mx) + (v[k].y - mx) + (v[k].z - mx) + (v[k].w - mx);
                else sum += ex2f(v[k].x - mx) + ex2f(v[k].y - mx) + ex2f(v[k].z - mx) + ex2f(v[k].w - mx);
            }
            if (op >= 3) {
                sum = warp_sum(sum);
                if ((tid & 31) == 0) rd[4 + (tid >> 5)] = sum;
                __syncthreads();
                sum = rd[4] + rd[5] + rd[6] + rd[7];
            }
            acc += sum;
            if (++stage == nst) { stage = 0; parity ^= 1; }
            continue;
        }
        if (tid == 0 && issued < n) {
            mbar_expect_tx(bar0 + 8 * stage, rb);
            tma_load_1d(slot0 + stage * slot_bytes, in + (size_t)row_of(issued) * V, rb, bar0 + 8 * stage);
            ++issued;
        }
        if (++stage == nst) { stage = 0; parity ^= 1; }
    }
    if (acc == 123.456f) sink[0] = acc;
}

__global__ void plain_copy(const float4 *__restrict__ in, float4 *__restrict__ out, size_t n4, int op, float *sink) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (; i + 3 * st < n4; i += 4 * st) {
        float4 a = in[i], b = in[i + st], c = in[i + 2 * st], d = in[i + 3 * st];
        if (op) { out[i] = a; out[i + st] = b; out[i + 2 * st] = c; out[i + 3 * st] = d; }
        else acc += a.x + b.y + c.z + d.w;
    }
    if (acc == 123.456f) sink[0] = acc;
}

int main(int argc, char **argv) {
    const int rows = 102400, V = 4236;   // 16-byte aligned rows for the probe
    const size_t n = (size_t)rows * V;
    float *in, *out, *sink;
    cudaMalloc(&in, n * 4); cudaMalloc(&out, n * 4); cudaMalloc(&sink, 4);
    cudaMemset(in, 0, n * 4); cudaMemset(out, 0, n * 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const uint32_t slot = (V * 4 + 127) / 128 * 128;
    auto time_it = [&](auto launch) { for (int i = 0; i < 3; ++i) launch(); cudaEventRecord(e0); for (int i = 0; i < 10; ++i) launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 10; };
    const char *opname[] = {"read", "copy", "read+max+ex2 (1 barrier)", "read+max+reduce+add (2 barriers, no MUFU)", "read+k1-like (2 barriers, MUFU)"};
    for (int op = 0; op < 5; ++op) {
        if (op < 2) {
            float ms = time_it([&] { plain_copy<<<sms * 8, 512>>>((const float4 *)in, (float4 *)out, n / 4, op, sink); });
            printf("plain %s: %.1f us  %.0f GB/s\n", op ? "copy" : "read", ms * 1e3, (op ? 2 : 1) * n * 4 / ms / 1e6);
        }
        for (int cps : {3, 4}) for (int nst : {3}) {
            size_t smem = (size_t)nst * slot + 8 * nst;
            if (smem * cps > 225 * 1024) continue;
            cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            for (int mode = 0; mode < 3; ++mode) {
                int R = 16;
                float ms2 = time_it([&] { probe<<<sms * cps, 128, smem>>>(in, out, sink, rows, V, mode, R, op, nst, slot); });
                printf("  tma-ring %s cps=%d nst=%d mode=%d: %.1f us  %.0f GB/s\n", opname[op], cps, nst, mode, ms2 * 1e3, (op == 1 ? 2 : 1) * n * 4 / ms2 / 1e6);
            }
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
