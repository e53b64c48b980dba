// Micro-probe: cost of float<->double conversions (F2F.F64.F32 / F2F.F32.F64) for one warp per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void f2d(double *out, const float *in, int iters, long long *cyc) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = in[threadIdx.x * 8 + i];
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double d;
            asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(v[i]));
            acc[i] += d;
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void f2d_int(double *out, const float *in, int iters, long long *cyc) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = in[threadIdx.x * 8 + i];
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int fb;
            asm volatile("mov.b32 %0, %1;" : "=r"(fb) : "f"(v[i]));
            acc[i] += __hiloint2double((fb >> 3) + 0x38000000, fb << 29);
        }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void d2f(float *out, const double *in, int iters, long long *cyc) {
    double v[8];
    for (int i = 0; i < 8; ++i) v[i] = in[threadIdx.x * 8 + i];
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float f;
            asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(v[i]));
            acc[i] += f;
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void dmix(double *out, double a, double b, int iters, long long *cyc) {
    // 8 independent chains of DADD -> DFMA -> DMUL (the lattice's mix)
    double x[8];
    for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { double s = x[i] + x[(i + 1) & 7]; s = fma(b, x[(i + 2) & 7], s); x[i] = s * b; }
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; ++i) s += x[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void shfl_st(double *out, double *g, double a, int iters, long long *cyc, int nst) {
    // a 64-bit shuffle whose result is needed right away, with nst STG.128 per thread issued just before it
    double x = a + threadIdx.x;
    double *p = g + (size_t)blockIdx.x * 65536 + threadIdx.x * 2;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < nst; ++k) asm volatile("st.global.v2.f64 [%0], {%1,%1};" ::"l"(p + ((it * 8 + k) & 255) * 64), "d"(x) : "memory");
        double q = __shfl_up_sync(0xffffffffu, x, 1);
        x = x * 0.5 + q * 0.25;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
__global__ void shfl_mode(double *out, double *g, double a, int iters, long long *cyc, int nst) {
    // MODE 0: st.shared.v2.f64   1: st.global.v2.f32 (8 B/lane)   2: st.global.f32 (4 B/lane)   3: st.global.v2.f64 to one fixed row
    __shared__ __align__(16) double sm[4][8][64];
    double x = a + threadIdx.x;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    double *p = g + (size_t)blockIdx.x * 65536 + threadIdx.x * 2;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int k = 0; k < nst; ++k) {
            if (MODE == 0) { sm[w][k & 7][2 * l] = x; sm[w][k & 7][2 * l + 1] = x; asm volatile("" ::: "memory"); }
            if (MODE == 1) asm volatile("st.global.v2.f32 [%0], {%1,%1};" ::"l"((float *)(p + ((it * 8 + k) & 255) * 64)), "f"((float)x) : "memory");
            if (MODE == 2) asm volatile("st.global.f32 [%0], %1;" ::"l"((float *)(g + (size_t)blockIdx.x * 65536 + ((it * 8 + k) & 255) * 64) + threadIdx.x), "f"((float)x) : "memory");
            if (MODE == 3) asm volatile("st.global.v2.f64 [%0], {%1,%1};" ::"l"(p + (k & 7) * 64), "d"(x) : "memory");
        }
        double q = __shfl_up_sync(0xffffffffu, x, 1);
        x = x * 0.5 + q * 0.25;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + sm[w][0][l];
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out, *g; long long *cyc, h; float *in;
    cudaMalloc(&out, 1 << 16); cudaMalloc(&cyc, 8); cudaMalloc(&in, 1 << 16); cudaMemset(in, 0x3c, 1 << 16);
    cudaMalloc(&g, (size_t)65536 * 8 * 160);
    const int it = 4000;
#define RUN(name, call) call; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-44s %8.2f cycles/iter\n", name, (double)h / it);
    RUN("cvt f32->f64 x8, 1 warp", (f2d<<<1, 32>>>(out, in, it, cyc)));
    RUN("cvt f32->f64 x8, 4 warps (1/SMSP)", (f2d<<<1, 128>>>(out, in, it, cyc)));
    RUN("int-unpack f32->f64 x8, 1 warp", (f2d_int<<<1, 32>>>(out, in, it, cyc)));
    RUN("cvt f64->f32 x8, 1 warp", (d2f<<<1, 32>>>((float *)out, (double *)in, it, cyc)));
    RUN("cvt f64->f32 x8, 4 warps (1/SMSP)", (d2f<<<1, 128>>>((float *)out, (double *)in, it, cyc)));
    RUN("dadd+dfma+dmul x8, 1 warp", (dmix<<<1, 32>>>(out, 1e-9, 0.999, it, cyc)));
    RUN("dadd+dfma+dmul x8, 4 warps (1/SMSP)", (dmix<<<1, 128>>>(out, 1e-9, 0.999, it, cyc)));
    RUN("dadd+dfma+dmul x8, 128 CTAs x 4 warps", (dmix<<<128, 128>>>(out, 1e-9, 0.999, it, cyc)));
    for (int nst = 0; nst <= 8; nst += 4) {
        char nm[64]; snprintf(nm, 64, "shfl64 chain + %d STG.128, 1 warp", nst);
        RUN(nm, (shfl_st<<<1, 32>>>(out, g, 1e-9, it, cyc, nst)));
        snprintf(nm, 64, "shfl64 chain + %d STG.128, 128 CTAs x 4 warps", nst);
        RUN(nm, (shfl_st<<<128, 128>>>(out, g, 1e-9, it, cyc, nst)));
    }
    RUN("shfl64 chain + 4 STS.128, 1 warp", (shfl_mode<0><<<1, 32>>>(out, g, 1e-9, it, cyc, 4)));
    RUN("shfl64 chain + 8 STS.128, 128x4 warps", (shfl_mode<0><<<128, 128>>>(out, g, 1e-9, it, cyc, 8)));
    RUN("shfl64 chain + 4 STG.64, 1 warp", (shfl_mode<1><<<1, 32>>>(out, g, 1e-9, it, cyc, 4)));
    RUN("shfl64 chain + 4 STG.32, 1 warp", (shfl_mode<2><<<1, 32>>>(out, g, 1e-9, it, cyc, 4)));
    RUN("shfl64 chain + 4 STG.128 same rows, 1 warp", (shfl_mode<3><<<1, 32>>>(out, g, 1e-9, it, cyc, 4)));
    cudaDeviceSynchronize();
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
