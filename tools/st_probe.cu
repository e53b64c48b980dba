// Micro-probe: how much do stores issued just before a shuffle delay a dependent chain through it?
// (stores and shuffles share the SM's memory-instruction queue).  Unrolled stores, immediate offsets.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int NST>
__global__ void chain(double *out, double *g, double a, int iters, long long *cyc) {
    __shared__ __align__(16) double sm[4][8 * 64];
    double x = a + threadIdx.x;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    double *p = g + (size_t)blockIdx.x * 65536 + (size_t)w * 8192 + l * 2;
    double *s = &sm[w][2 * l];
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < NST; ++k) {
            if (MODE == 0) asm volatile("st.global.v2.f64 [%0], {%1,%1};" ::"l"(p + k * 64), "d"(x) : "memory");
            if (MODE == 1) asm volatile("st.shared.v2.f64 [%0], {%1,%1};" ::"r"((unsigned)__cvta_generic_to_shared(s + k * 64)), "d"(x) : "memory");
            if (MODE == 2) asm volatile("st.global.v2.f32 [%0], {%1,%1};" ::"l"((float *)(p + k * 64)), "f"((float)x) : "memory");
        }
        double q = __shfl_up_sync(0xffffffffu, x, 1);
        x = x * 0.5 + q * 0.25;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + sm[w][l];
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out, *g; long long *cyc, h;
    cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 8); cudaMalloc(&g, (size_t)65536 * 8 * 160);
    const int it = 4000;
#define RUN(name, call) call; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-52s %8.2f cycles/iter\n", name, (double)h / it);
    RUN("chain, no stores, 1 warp", (chain<0, 0><<<1, 32>>>(out, g, 1e-9, it, cyc)));
    RUN("chain + 1 STG.128, 1 warp", (chain<0, 1><<<1, 32>>>(out, g, 1e-9, it, cyc)));
    RUN("chain + 2 STG.128, 1 warp", (chain<0, 2><<<1, 32>>>(out, g, 1e-9, it, cyc)));
    RUN("chain + 4 STG.128, 1 warp", (chain<0, 4><<<1, 32>>>(out, g, 1e-9, it, cyc)));
    RUN("chain + 8 STG.128, 1 warp", (chain<0, 8><<<1, 32>>>(out, g, 1e-9, it, cyc)));
    RUN("chain + 4 STG.128, 4 warps/CTA", (chain<0, 4><<<1, 128>>>(out, g, 1e-9, it, cyc)));
    RUN("chain + 4 STG.128, 128 CTAs x 4 warps", (chain<0, 4><<<128, 128>>>(out, g, 1e-9, it, cyc)));
    RUN("chain + 8 STG.128, 128 CTAs x 4 warps", (chain<0, 8><<<128, 128>>>(out, g, 1e-9, it, cyc)));
    RUN("chain + 4 STS.128, 1 warp", (chain<1, 4><<<1, 32>>>(out, g, 1e-9, it, cyc)));
    RUN("chain + 8 STS.128, 1 warp", (chain<1, 8><<<1, 32>>>(out, g, 1e-9, it, cyc)));
    RUN("chain + 8 STS.128, 128 CTAs x 4 warps", (chain<1, 8><<<128, 128>>>(out, g, 1e-9, it, cyc)));
    RUN("chain + 4 STG.64, 1 warp", (chain<2, 4><<<1, 32>>>(out, g, 1e-9, it, cyc)));
    cudaDeviceSynchronize();
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
