// Thin inline-PTX wrappers for sm_100a: mbarrier, 1-D bulk TMA (cp.async.bulk -> SASS UBLKCP),
// proxy fences, fast exp2/log2 and warp reductions.  No CUTLASS/CuTe dependency.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ctcb200 {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
#define CTC_NEG_INF (__int_as_float(0xff800000))
// log2(0) sentinel of the lattice (finite, so a-b is never inf-inf); real values are > -1e29 by a wide margin
constexpr float kNeg = -1.0e30f;
constexpr float kNegTest = -1.0e29f;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// The same wait, giving up (trap) after ~2 s: a hang on a shared GPU box is far worse than a failed launch.
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

// ---- TMA: 1-D bulk copy global -> shared, completion signalled on an mbarrier -----------------
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
// L2 eviction-priority policies (the encodings CUTLASS uses for TMA cache hints on sm_90/sm_100).
// The [B,T,V] streams are touched once per kernel -> evict_first, so that they do not flush the small
// per-frame arrays (lp_lab, gam, stored alpha/beta halves) the lattice kernel re-reads -> evict_last.
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
// L2 policy of the per-frame scratch arrays (lp_lab, gam, the stored lattice halves): written by one kernel and
// re-read by the next.  A __constant__ so that the `scratch_policy` option can switch it (ctcb200.cu); per
// translation unit, only ctcb200.cu's copy is ever changed or used.
static __constant__ uint64_t c_scratch_policy = kEvictLast;
#define kScratch c_scratch_policy
// L2 policy of k1p_sweep's bulk gradient stores (option k1p_store_policy: 0 evict_first, 1 evict_normal, 2 evict_last)
static __constant__ uint64_t c_gstore_policy = kEvictFirst;

__device__ __forceinline__ void tma_load_1d_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
// streaming 16-byte load: no L1 allocation (every byte of the [B,T,V] tensors is touched once per kernel)
__device__ __forceinline__ float4 ldg_v4_stream(const float4 *p) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_v4_hint(float4 *p, float4 v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void stg_v2_hint(float2 *p, float2 v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(policy) : "memory");
}
__device__ __forceinline__ void stg_v2f64_hint(double *p, double a, double b, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(policy) : "memory");
}
__device__ __forceinline__ void stg_f32_hint(float *p, float v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(policy) : "memory");
}

// generic-proxy writes to global -> later async-proxy (TMA) reads of the same bytes
__device__ __forceinline__ void fence_proxy_async_global() {
    asm volatile("fence.proxy.async.global;" ::: "memory");
}
// 1-D bulk TMA store shared -> global (bulk async-group completion), with an L2 policy; dst/src 16-byte aligned
__device__ __forceinline__ void tma_store_1d_hint(void *dst, uint32_t src, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(dst), "r"(src), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups have not yet READ their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// generic-proxy writes to shared memory -> later async-proxy reads of the same bytes (tcgen05.mma operands, bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem_cta() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may be scheduled while its
// predecessor on the stream is still running; it must not touch the predecessor's results before
// griddep_wait().  griddep_launch_dependents() in the predecessor allows that early scheduling, which hides
// the launch latency and the CTA set-up of the next kernel behind the tail of this one.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- math -------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2f(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// log2(2^a + 2^b), -inf safe
__device__ __forceinline__ float lse2_2(float a, float b) {
    const float m = fmaxf(a, b);
    const float ms = (m == CTC_NEG_INF) ? 0.f : m;
    return ms + lg2f(ex2f(a - ms) + ex2f(b - ms));
}
// log2(2^a + 2^b + 2^c), -inf safe
__device__ __forceinline__ float lse3_2(float a, float b, float c) {
    const float m = fmaxf(fmaxf(a, b), c);
    const float ms = (m == CTC_NEG_INF) ? 0.f : m;
    return ms + lg2f(ex2f(a - ms) + ex2f(b - ms) + ex2f(c - ms));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace ctcb200
