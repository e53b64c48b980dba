// Micro-probe for the linear-domain lattice: FP64 add/mul latency and single-warp issue rate on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void dchain(double *out, double a, double b, int iters, long long *cyc) {
    double x[CH];
    for (int i = 0; i < CH; ++i) x[i] = a + threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fma(x[i], b, a);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < CH; ++i) s += x[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
__global__ void fchain(float *out, float a, float b, int iters, long long *cyc) {
    float x[CH];
    for (int i = 0; i < CH; ++i) x[i] = a + threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fmaf(x[i], b, a);
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < CH; ++i) s += x[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void shfl_chain(double *out, double a, int iters, long long *cyc) {
    double x = a + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double p = __shfl_up_sync(0xffffffffu, x, 1);
        x = x + p;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void cvt_chain(double *out, const float *lp, int iters, long long *cyc) {
    // throughput of the lp -> double p conversion (4 independent per iteration)
    double acc = 0;
    float v[4];
    for (int i = 0; i < 4; ++i) v[i] = lp[threadIdx.x * 4 + i];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float x = v[i] - (float)it;
            const float t = x + 12582912.f;
            const float fr = x - (t - 12582912.f);
            float e;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fr));
            double d = (double)e;
            int hi = __double2hiint(d) + ((__float_as_int(t) - 0x4B400000) << 20);
            acc += __hiloint2double(hi, __double2loint(d));
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = acc;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double *out; long long *cyc, h; float *lp;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8); cudaMalloc(&lp, 4096); cudaMemset(lp, 0, 4096);
    const int it = 10000;
#define RUN(name, call) call; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-28s %8.2f cycles/iter\n", name, (double)h / it);
    for (int rep = 0; rep < 2; ++rep) {
        RUN("dfma chain x1 (latency)", (dchain<1><<<1, 32>>>(out, 1e-9, 0.999, it, cyc)));
        RUN("dfma x4 independent", (dchain<4><<<1, 32>>>(out, 1e-9, 0.999, it, cyc)));
        RUN("dfma x8 independent", (dchain<8><<<1, 32>>>(out, 1e-9, 0.999, it, cyc)));
        RUN("dfma x16 independent", (dchain<16><<<1, 32>>>(out, 1e-9, 0.999, it, cyc)));
        RUN("dfma x8, 4 warps/CTA(1/SMSP)", (dchain<8><<<1, 128>>>(out, 1e-9, 0.999, it, cyc)));
        RUN("ffma chain x1 (latency)", (fchain<1><<<1, 32>>>((float *)out, 1e-9f, 0.999f, it, cyc)));
        RUN("ffma x8 independent", (fchain<8><<<1, 32>>>((float *)out, 1e-9f, 0.999f, it, cyc)));
        RUN("shfl64 + dadd chain", (shfl_chain<<<1, 32>>>(out, 1e-9, it, cyc)));
        RUN("lp->double p x4", (cvt_chain<<<1, 32>>>(out, lp, it, cyc)));
    }
    cudaDeviceSynchronize();
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
