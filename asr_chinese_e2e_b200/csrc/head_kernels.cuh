// f1: the CTC head GEMM  logits = enc[M,K] x W[V,K]^T + bias  fused with the log-softmax
// statistics + label gather (pass 1) and with the gradient  g * (softmax - occupancy)  (pass 2, recomputing the logits
// tile), so that the [B,T,V] logits tensor is never written to or read from HBM.  Tap point of the reference:
// Predictor/Models/transformer_official.py:76 (encoder output) -> nn.Linear(d_model, vocab).
//
// Blackwell structure (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d (UTMALDG) of 128 x 32 / 256 x 32 fp32 operand boxes, SWIZZLE_128B,
//            into an mbarrier-guarded shared-memory ring
//   warp 1   MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (UTCMMA), M=128, N=256, K=8 per
//            instruction, fp32 accumulators in TMEM (2 x 256 columns, double buffered against the epilogue);
//            tcgen05.commit frees ring slots / publishes accumulators
//   warp 2   TMEM allocator (512 columns)
//   warps 4-7 epilogue: tcgen05.ld (LDTM) 32 lanes x 32 columns at a time, + bias, online max / sum-of-2^x per row,
//            the <= U+1 label logits of the row picked straight out of TMEM (tcgen05.ld .x1 at the label's column)
//
// Precision: NPASS = 3 is "3xTF32": every fp32 operand is split as x = hi + lo with hi = rna_tf32(x) (a pre-pass,
// k_split_tf32) and the product accumulates hi*hi + hi*lo + lo*hi in fp32 -- the error per product is ~2^-21, i.e.
// fp32-GEMM grade, which the 1e-5 relative parity bar on the per-utterance loss needs.  NPASS = 1 is plain TF32
// (10-bit mantissa operands; its own, looser tolerance -- tests/test_gpu_head.py).
//
// Rows are the flattened (b, t) frames; a 128-row tile may span utterance boundaries, every epilogue thread owns one
// row.  Tiles without a valid frame (t >= T_b) are skipped.
#pragma once
#include <cuda.h>

#include "layout.h"
#include "ptx.cuh"

namespace ctcb200 {

constexpr int HM = 128, HN = 256, HK = 32;                     // tile: 128 rows x 256 classes, K chunk of 32 floats = 128 B
constexpr uint32_t H_A_BYTES = HM * HK * 4, H_B_BYTES = HN * HK * 4;
constexpr int H_THREADS = 256;

template <int NPASS>
struct HeadCfg {
    static constexpr uint32_t STAGE = (NPASS == 3 ? 2 : 1) * (H_A_BYTES + H_B_BYTES);
    static constexpr int NSTAGE = NPASS == 3 ? 2 : 4;
    static constexpr uint32_t RING = NSTAGE * STAGE;          // 192 KB
    static constexpr uint32_t OFF_BARS = RING;                 // full[NSTAGE], empty[NSTAGE], tfull[2], tempty[2], conv[NSTAGE]
    static constexpr uint32_t OFF_TMEM = OFF_BARS + 8 * (3 * NSTAGE + 4);
    static constexpr uint32_t OFF_BIAS = OFF_TMEM + 16;        // float[2][HN]
    static constexpr uint32_t SMEM = OFF_BIAS + 2 * HN * 4 + 1024;   // + slack to align the ring to 1024 B
};

// ---- PTX wrappers (tcgen05 / TMA tensor) -----------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar,
                                                 uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
                 "[%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(dst), "l"((uint64_t)map), "r"(c0), "r"(c1), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 operands, fp32 accumulate.  One thread issues for the whole CTA.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                   "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return __uint_as_float(r);
}
__device__ __forceinline__ void mbar_wait_or_trap(uint32_t bar, uint32_t parity) { mbar_wait_bounded(bar, parity); }
// shared-memory matrix descriptor: K-major operand, rows of 128 B, SWIZZLE_128B atoms of 8 rows (1024 B apart)
// (cute::UMMA::SmemDescriptor: start address >> 4 in [0,14), LBO >> 4 in [16,30) -- unused for swizzled K-major, 1 like
//  CUTLASS --, SBO >> 4 in [32,46), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64))
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 at [4,6), a/b format TF32 = 2 at [7,10)/[10,13),
// a/b K-major (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- x = hi + lo with hi = round-to-nearest tf32(x): the operand split of the 3xTF32 product ----------------------
__global__ void __launch_bounds__(256) k_split_tf32(const float *__restrict__ x, float *__restrict__ hi,
                                                    float *__restrict__ lo, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = ((const float4 *)x)[i];
        float4 h, l;
        uint32_t u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.x)); h.x = __uint_as_float(u); l.x = v.x - h.x;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.y)); h.y = __uint_as_float(u); l.y = v.y - h.y;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.z)); h.z = __uint_as_float(u); l.z = v.z - h.z;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.w)); h.w = __uint_as_float(u); l.w = v.w - h.w;
        ((float4 *)hi)[i] = h;
        ((float4 *)lo)[i] = l;
    }
}

struct HeadArgs {
    const float *bias;                 // [V] or nullptr
    const int64_t *targets; int64_t tnumel;
    const int *Tb, *Ub; const int64_t *toff; const int *flags; int *slow; int *bad; int *hdr;
    float *lp_lab;                     // [B*T, Lp] frames (pass 1 writes, pass 2 reads slot 1 = lse2)
    const float *gam;                  // [B*T, Lp] occupancies (pass 2)
    float *dlogits; int64_t pitch;     // pass 2 output [B*T, pitch] floats, pitch >= V, multiple of 4
    int B, T, V, K, Lp, blank, zero_inf, reduction; float inv_batch, lin_thr, occ_skip;
};

// live tile: some row of [128*tile, 128*tile+128) is a valid frame
__device__ __forceinline__ bool head_tile_live(int tile, const int *Tb, int B, int T) {
    const long long r0 = (long long)tile * HM, r1 = r0 + HM - 1;
    int b0 = (int)(r0 / T), b1 = (int)(r1 / T);
    if (b0 >= B) return false;
    if (b1 >= B) b1 = B - 1;
    for (int b = b0; b <= b1; ++b) {
        const long long first = (long long)b * T > r0 ? (long long)b * T : r0;      // first row of utterance b in the tile
        if ((int)(first - (long long)b * T) < __ldcg(Tb + b)) return true;
    }
    return false;
}

// INRING (3xTF32 only; option head_inring, off by default -- measured slower, see ctcb200.cu): the operands are NOT pre-split in HBM.  The TMA delivers
// the raw fp32 tiles of enc / W (tmA_hi / tmB_hi then map the caller's tensors themselves); the tensor core ignores
// the 13 low mantissa bits of a tf32 operand, so the raw tile IS hi = trunc_tf32(x); warps 2-3 write
// lo = x - trunc_tf32(x) (exact) into the stage's second tile and the hi*hi third of the stage's MMAs is issued the
// moment the tile lands, under the conversion (as in k_gemm3, gemm_tf32x3.cuh).  Against the pre-split variant this
// halves the operand bytes the ring pulls through L2 (the gradient pass re-read 5.0 GB from DRAM for 0.45 GB of
// operands because the hi/lo tiles fell out of L2 under the gradient stream), and removes the k_split_tf32 pre-pass
// with its 2 x 4*B*T*K bytes of workspace.
template <int NPASS, bool GRADPASS, bool INRING = false>
__global__ void __launch_bounds__(H_THREADS, 1)
k_head(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
       const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, const HeadArgs a) {
    static_assert(!INRING || NPASS == 3, "the in-ring split is the 3xTF32 product");
    using C = HeadCfg<NPASS>;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;           // SWIZZLE_128B atoms need 1024-byte alignment
    unsigned char *sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t bars = sbase + C::OFF_BARS;
    auto full = [&](int s) { return bars + 8 * s; };
    auto empty = [&](int s) { return bars + 8 * (C::NSTAGE + s); };
    auto tfull = [&](int i) { return bars + 8 * (2 * C::NSTAGE + i); };
    auto tempty = [&](int i) { return bars + 8 * (2 * C::NSTAGE + 2 + i); };
    auto conv = [&](int s) { return bars + 8 * (2 * C::NSTAGE + 4 + s); };
    volatile uint32_t *tmem_slot = (volatile uint32_t *)(sgen + C::OFF_TMEM);
    float *sbias = (float *)(sgen + C::OFF_BIAS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = a.B * a.T;
    const int n_tiles = (M + HM - 1) / HM, n_nt = (a.V + HN - 1) / HN, n_kc = a.K / HK;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA_hi); prefetch_tmap(&tmB_hi);
        if (NPASS == 3 && !INRING) { prefetch_tmap(&tmA_lo); prefetch_tmap(&tmB_lo); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); mbar_init(conv(s), 2); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull(i), 1); mbar_init(tempty(i), 4); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(smem_u32((const void *)tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                if (!head_tile_live(tile, a.Tb, a.B, a.T)) continue;
                const int row0 = tile * HM;
                for (int nt = 0; nt < n_nt; ++nt) {
                    for (int kc = 0; kc < n_kc; ++kc) {
                        mbar_wait_or_trap(empty(s), ph ^ 1);
                        const uint32_t st = sbase + s * C::STAGE;
                        mbar_expect_tx(full(s), INRING ? H_A_BYTES + H_B_BYTES : C::STAGE);
                        tma_load_2d(st, &tmA_hi, kc * HK, row0, full(s));
                        if (INRING) {
                            tma_load_2d_hint(st + 2 * H_A_BYTES, &tmB_hi, kc * HK, nt * HN, full(s), kEvictLast);
                        } else if (NPASS == 3) {
                            tma_load_2d(st + H_A_BYTES, &tmA_lo, kc * HK, row0, full(s));
                            tma_load_2d_hint(st + 2 * H_A_BYTES, &tmB_hi, kc * HK, nt * HN, full(s), kEvictLast);
                            tma_load_2d_hint(st + 2 * H_A_BYTES + H_B_BYTES, &tmB_lo, kc * HK, nt * HN, full(s), kEvictLast);
                        } else {
                            tma_load_2d_hint(st + H_A_BYTES, &tmB_hi, kc * HK, nt * HN, full(s), kEvictLast);
                        }
                        if (++s == C::NSTAGE) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(HM, HN);
            int s = 0; uint32_t ph = 0;
            int acc = 0; uint32_t aph = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                if (!head_tile_live(tile, a.Tb, a.B, a.T)) continue;
                for (int nt = 0; nt < n_nt; ++nt) {
                    mbar_wait_or_trap(tempty(acc), aph ^ 1);               // the epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t tacc = tmem_base + (uint32_t)(acc * HN);
                    for (int kc = 0; kc < n_kc; ++kc) {
                        mbar_wait_or_trap(full(s), ph);
                        tc_fence_after();
                        const uint32_t st = sbase + s * C::STAGE;
                        const uint32_t a_hi = st, a_lo = st + H_A_BYTES;
                        const uint32_t b_hi = st + (NPASS == 3 ? 2 : 1) * H_A_BYTES, b_lo = b_hi + H_B_BYTES;
                        if (INRING) {
#pragma unroll
                            for (int k = 0; k < HK / 8; ++k)               // the raw tiles are the hi halves: no wait for warps 2-3
                                umma_tf32(tacc, umma_desc_k128(a_hi + k * 32), umma_desc_k128(b_hi + k * 32), idesc, (kc | k) != 0);
                            mbar_wait_or_trap(conv(s), ph);                // the lo halves are written
                            tc_fence_after();
#pragma unroll
                            for (int k = 0; k < HK / 8; ++k) {
                                umma_tf32(tacc, umma_desc_k128(a_lo + k * 32), umma_desc_k128(b_hi + k * 32), idesc, 1);
                                umma_tf32(tacc, umma_desc_k128(a_hi + k * 32), umma_desc_k128(b_lo + k * 32), idesc, 1);
                            }
                        } else
#pragma unroll
                        for (int k = 0; k < HK / 8; ++k) {                 // UMMA K = 8 tf32 = 32 bytes inside the 128 B row
                            const uint32_t ko = k * 32;
                            if (NPASS == 3) {
                                umma_tf32(tacc, umma_desc_k128(a_lo + ko), umma_desc_k128(b_hi + ko), idesc, (kc | k) != 0);
                                umma_tf32(tacc, umma_desc_k128(a_hi + ko), umma_desc_k128(b_lo + ko), idesc, 1);
                                umma_tf32(tacc, umma_desc_k128(a_hi + ko), umma_desc_k128(b_hi + ko), idesc, 1);
                            } else {
                                umma_tf32(tacc, umma_desc_k128(a_hi + ko), umma_desc_k128(b_hi + ko), idesc, (kc | k) != 0);
                            }
                        }
                        umma_commit(empty(s));                             // slot reusable once these MMAs have read it
                        if (++s == C::NSTAGE) { s = 0; ph ^= 1; }
                    }
                    umma_commit(tfull(acc));                               // accumulator complete
                    if (++acc == 2) { acc = 0; aph ^= 1; }
                }
            }
        }
    } else if (INRING && (warp == 2 || warp == 3)) {
        // ===================== converters: lo = x - trunc_tf32(x) of the freshly landed fp32 tiles =====================
        const int ct = tid - 64;                                          // 0..63
        int s = 0; uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            if (!head_tile_live(tile, a.Tb, a.B, a.T)) continue;
            for (int it = 0; it < n_nt * n_kc; ++it) {
                mbar_wait_or_trap(full(s), ph);
                float4 *A4 = (float4 *)(sgen + (size_t)s * C::STAGE);
                float4 *B4 = (float4 *)(sgen + (size_t)s * C::STAGE + 2 * H_A_BYTES);
                auto lo_of = [](const float4 x) {
                    float4 l;
                    l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
                    l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
                    l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
                    l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
                    return l;
                };
#pragma unroll 4
                for (int i = ct; i < (int)(H_A_BYTES / 16); i += 64) A4[H_A_BYTES / 16 + i] = lo_of(A4[i]);
#pragma unroll 4
                for (int i = ct; i < (int)(H_B_BYTES / 16); i += 64) B4[H_B_BYTES / 16 + i] = lo_of(B4[i]);
                fence_proxy_async_smem_cta();                             // generic-proxy writes -> tcgen05.mma (async proxy) reads
                __syncwarp();
                if (lane == 0) mbar_arrive(conv(s));
                if (++s == C::NSTAGE) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue: one row per thread =====================
        const int wq = warp & 3;                                          // TMEM lane quarter of this warp
        const int et = tid - 128;                                         // 0..127
        const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
        int acc = 0; uint32_t aph = 0;
        int bias_buf = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const bool live = head_tile_live(tile, a.Tb, a.B, a.T);
            const int row = tile * HM + wq * 32 + lane;
            const int b = row < M ? row / a.T : a.B - 1;
            const int t = row - b * a.T;
            const int myTb = __ldcg(a.Tb + b), myUb = __ldcg(a.Ub + b);
            const bool valid = row < M && t < myTb;
            float *drow = GRADPASS ? a.dlogits + (size_t)row * a.pitch : nullptr;
            if (!live) {
                if (GRADPASS && row < M) {                                // padded frames only: zero rows of dlogits
                    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int n = 0; n < a.pitch; n += 4) stg_v4_hint((float4 *)(drow + n), z, kEvictFirst);
                }
                continue;
            }
            float *frame = a.lp_lab + (size_t)row * a.Lp;
            // utterances present in this warp's 32 rows (warp-uniform range)
            const int rw0 = tile * HM + wq * 32;
            int wb0 = rw0 < M ? rw0 / a.T : a.B - 1, wb1 = (rw0 + 31 < M ? rw0 + 31 : M - 1) / a.T;
            if (wb1 < wb0) wb1 = wb0;
            float m2 = CTC_NEG_INF, ssum = 0.f;                           // pass 1: running max / sum of 2^(x2 - m2)
            float lse2 = 0.f, g = 0.f;                                    // pass 2
            bool nanrow = false, zerorow = false;
            if (GRADPASS) {
                if (valid) {
                    lse2 = frame[1];
                    g = a.reduction == 1 ? a.inv_batch * __frcp_rn((float)(myUb > 1 ? myUb : 1)) : 1.f;
                    const int infeasible = __ldcg(a.flags + b), isbad = __ldcg(a.bad + b);
                    nanrow = isbad || (infeasible && !a.zero_inf);
                    zerorow = infeasible && a.zero_inf && !isbad;
                } else zerorow = true;
            }
            for (int nt = 0; nt < n_nt; ++nt) {
                const int n0 = nt * HN;
                // bias of this class tile (and -inf for classes >= V) -> shared memory, double buffered
                float *sb = sbias + bias_buf * HN;
                for (int i = et; i < HN; i += 128) {
                    const int n = n0 + i;
                    sb[i] = n < a.V ? (a.bias ? __ldg(a.bias + n) : 0.f) : CTC_NEG_INF;
                }
                named_bar_sync(1, 128);
                mbar_wait_or_trap(tfull(acc), aph);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(acc * HN) + lane_sel;
#pragma unroll 1
                for (int ch = 0; ch < HN / 32; ++ch) {
                    if (n0 + ch * 32 >= a.V) break;
                    float v[32];
                    tmem_ld32(tacc + ch * 32, v);
                    if (!GRADPASS) {
                        float cm = CTC_NEG_INF;
#pragma unroll
                        for (int i = 0; i < 32; ++i) { v[i] = (v[i] + sb[ch * 32 + i]) * kLog2e; cm = fmaxf(cm, v[i]); }
                        const float mn = fmaxf(m2, cm);
                        float add = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) add += ex2f(v[i] - mn);
                        ssum = ssum * ex2f(m2 - mn) + add;
                        m2 = mn;
                    } else if (row < M) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float4 y;
                            y.x = g * ex2f(fmaf(v[i] + sb[ch * 32 + i], kLog2e, -lse2));
                            y.y = g * ex2f(fmaf(v[i + 1] + sb[ch * 32 + i + 1], kLog2e, -lse2));
                            y.z = g * ex2f(fmaf(v[i + 2] + sb[ch * 32 + i + 2], kLog2e, -lse2));
                            y.w = g * ex2f(fmaf(v[i + 3] + sb[ch * 32 + i + 3], kLog2e, -lse2));
                            if (zerorow) y = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (nanrow) { const float q = __int_as_float(0x7fc00000); y = make_float4(q, q, q, q); }
                            const int n = n0 + ch * 32 + i;
                            // streamed once: evict_first, so that the gradient rows do not push the operand tiles (re-read
                            // for every class tile) out of L2 -- round-2 ncu: 5.2 GB of DRAM reads with plain stores
                            if (n + 3 < a.pitch) stg_v4_hint((float4 *)(drow + n), y, kEvictFirst);
                            else {
                                if (n < a.pitch) drow[n] = y.x;
                                if (n + 1 < a.pitch) drow[n + 1] = y.y;
                                if (n + 2 < a.pitch) drow[n + 2] = y.z;
                            }
                        }
                    }
                }
                // the label columns of this class tile, straight out of TMEM (warp-collective single-column loads)
                for (int ub = wb0; ub <= wb1; ++ub) {
                    const int Uu = __ldcg(a.Ub + ub);
                    const int64_t toff = __ldcg(a.toff + ub);
                    const bool mine = valid && b == ub;
                    for (int j = -1; j < Uu; ++j) {
                        long long c = a.blank;
                        if (j >= 0) {
                            const int64_t idx = toff + j;
                            c = idx < a.tnumel ? __ldg(a.targets + idx) : 0;
                            if (c < 0 || c >= a.V) {
                                if (!GRADPASS && lane == 0 && nt == 0) { atomicOr(&a.hdr[0], 4); atomicOr(&a.bad[ub], 4); }
                                c = c < 0 ? 0 : a.V - 1;
                            }
                        }
                        const int col = (int)c - n0;
                        if (col < 0 || col >= HN) continue;               // warp-uniform
                        const int k = j < 0 ? 0 : 4 + j;
                        if (!GRADPASS) {
                            const float x = tmem_ld1(tacc + col) + sb[col];
                            if (mine) frame[k] = x;                       // raw logit; normalised once lse2 is known
                        } else if (mine && !zerorow && !nanrow) {
                            const float occ = __ldcg(a.gam + (size_t)row * a.Lp + k);
                            if (occ > a.occ_skip) drow[c] -= g * occ;     // same thread wrote drow[c] above: program order
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty(acc));
                if (++acc == 2) { acc = 0; aph ^= 1; }
                bias_buf ^= 1;
            }
            if (!GRADPASS && valid) {
                // frame in the lattice kernel's format (stream_kernels.cuh, k1): [0] blank, [1] lse2, [4+j] label j;
                // probabilities when in range of the linear-domain lattice, log2-probabilities otherwise
                const float l2 = m2 + lg2f(ssum);
                for (int k = 0; k < a.Lp; ++k) {
                    float o;
                    if (k == 1) o = l2;
                    else if (k == 2 || k == 3) o = 0.f;
                    else if (k >= 4 && k - 4 >= myUb) o = kNeg;
                    else {
                        o = fminf(fmaxf(fmaf(frame[k], kLog2e, -l2), kNeg), 0.f);
                        if (o >= a.lin_thr) o = ex2f(o);
                        else a.slow[b] = 1;
                    }
                    frame[k] = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace ctcb200
