// k_gemm3: C[Mc, Nc] = A x B^T on the tensor cores with fp32-grade accuracy ("3xTF32"), for the two parameter-gradient
// GEMMs of the fused CTC head:
//     d enc[B*T, K]  = dlogits[B*T, V] x W[V, K]            (reduction over the classes)
//     d W[V, K]      = dlogits[B*T, V]^T x enc[B*T, K]      (reduction over the frames)
// which torch would run as fp32 SIMT GEMMs (6.8 + 7.2 ms at the C2 shape on B200; cuBLAS has no tensor-core path for
// fp32 accuracy).  Same Blackwell machinery as k_head (head_kernels.cuh): TMA (UTMALDG) operand boxes into a
// SWIZZLE_128B shared-memory ring, tcgen05.mma kind::tf32 M128 N256 K8 issued by one thread, fp32 accumulators in TMEM,
// tcgen05.commit to recycle ring slots.  Two things are new here:
//   * operands may be "MN-major" (the reduction index is the ROW index of the matrix in memory, as for W in d enc and
//     for both operands of d W): the tile is fetched as 32 x 32 boxes and described to the MMA with the MN-major
//     canonical layout of 32-bit operands (SWIZZLE_128B_BASE32B: ((8,n),(4,k)) in 16-byte units, LBO = 4096 B between
//     32-element MN blocks, SBO = 512 B between groups of 4 reduction rows), a_major / b_major = 1 in the instruction descriptor;
//     no transposed copies of the 1.7 GB gradient are ever made;
//   * the hi/lo split of the 3xTF32 product happens IN the ring: two converter warps turn each freshly landed fp32
//     tile into hi = rna_tf32(x) (in place) and lo = x - hi (a second tile) before the MMA warp is released, so the
//     activations are never split in HBM either.
// Rows / columns / reduction tails outside the matrices read as zeros (TMA out-of-bounds fill).
#pragma once
#include "head_kernels.cuh"

namespace ctcb200 {

constexpr int G_THREADS = 384;                                   // TMA, MMA, 2 converter warps | 8 epilogue warps
constexpr uint32_t G_STAGE = 2 * (H_A_BYTES + H_B_BYTES);        // A raw->hi, A lo, B raw->hi, B lo  (96 KB)
constexpr int G_NSTAGE = 2;
constexpr uint32_t G_OFF_BARS = G_NSTAGE * G_STAGE;             // full[2], conv[2], empty[2], tfull[2], tempty[2]
constexpr uint32_t G_OFF_TMEM = G_OFF_BARS + 8 * (3 * G_NSTAGE + 4);
// The tensor core's fp32 accumulation truncates (measured on B200: over 19 200 accumulations into one TMEM tile the
// result drifts by 4e-4 relative, 26 x cuBLAS's fp32 error), so an accumulator only ever holds G_SEG reduction chunks
// (128 elements = 48 accumulations): the eight epilogue warps -- two per TMEM lane quarter, 128 of the tile's 256
// columns each -- add every finished segment to a running sum in REGISTERS (128 per thread, fp32 round-to-nearest)
// while the MMA warp fills the other accumulator, and store the tile once at the end.  (A first version kept the
// running sum in the output tile itself, a read-modify-write through L2 per segment: 9.6 ms instead of 5.4 ms for
// the two C2 GEMMs.)  setmaxnreg moves registers from the producer warpgroup to the two epilogue warpgroups.
constexpr int G_SEG = 4;
constexpr int G_REGS_PRODUCER = 72, G_REGS_EPILOGUE = 216;
constexpr uint32_t G_SMEM = G_OFF_TMEM + 16 + 1024;

struct GemmArgs {
    float *C; int64_t ldc;            // output, row-major
    int Mc, Nc, R;                    // C is Mc x Nc, reduction length R
    int ksplit; int64_t split_stride; // reduction split: slice s accumulates into C + s * split_stride (floats)
    uint64_t mn_desc;                 // upper part (LBO, SBO, version, layout type) of the MN-major operand descriptors
};

// shared-memory descriptor of an MN-major operand tile made of 32 x 32 fp32 boxes (4 KB each).  32-bit MN-major operands
// have ONE legal shared-memory layout on sm_100: SWIZZLE_128B_BASE32B (layout type 1; 32-byte chunks of a 128-byte row
// XOR-ed with the row index mod 4, i.e. Swizzle<2,5,2> on the byte address -- the tensor map's
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), canonical form ((8,n),(4,k)) in 16-byte units: LBO between 32-element MN
// blocks (= one box, 4096 B), SBO between groups of 4 reduction rows (512 B).  The plain SWIZZLE_128B description
// (round 2, first attempt) multiplies zeros.
__host__ __device__ constexpr uint64_t umma_desc_mn_hi(uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint64_t hi) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | hi;
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32_major(int M, int N, bool a_mn, bool b_mn) {
    return umma_idesc_tf32(M, N) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16);
}

// A_MN / B_MN: operand is MN-major (tensor map: inner dimension = the operand's M / N index, box 32 x 32);
// otherwise K-major (inner dimension = reduction index, box 32 x 128 for A / 32 x 256 for B).
// TRUNC: how x = hi + lo is formed.  false: hi = rna_tf32(x) written in place, lo = x - hi (|lo| <= 2^-11 |x|); every
// MMA of a stage waits for the converters.  true (default): the tensor core ignores the 13 low mantissa bits of a
// tf32 operand, so the RAW tile the TMA delivered already IS hi = trunc_tf32(x); the converters only write
// lo = x - trunc_tf32(x) (exact, |lo| < 2^-10 |x|) and the hi*hi third of the stage's MMAs is issued the moment the
// tile lands, under the conversion -- the stage's critical path loses the converter latency.
template <bool A_MN, bool B_MN, bool TRUNC>
__global__ void __launch_bounds__(G_THREADS, 1)
k_gemm3(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs a) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *sgen = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t bars = sbase + G_OFF_BARS;
    auto full = [&](int s) { return bars + 8 * s; };
    auto conv = [&](int s) { return bars + 8 * (G_NSTAGE + s); };
    auto empty = [&](int s) { return bars + 8 * (2 * G_NSTAGE + s); };
    auto tfull = [&](int i) { return bars + 8 * (3 * G_NSTAGE + i); };
    auto tempty = [&](int i) { return bars + 8 * (3 * G_NSTAGE + 2 + i); };
    volatile uint32_t *tmem_slot = (volatile uint32_t *)(sgen + G_OFF_TMEM);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_mt = (a.Mc + HM - 1) / HM, n_nt = (a.Nc + HN - 1) / HN;
    const int n_kc_all = (a.R + HK - 1) / HK;
    const int kc_per = (n_kc_all + a.ksplit - 1) / a.ksplit;
    const int n_items = n_mt * n_nt * a.ksplit;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < G_NSTAGE; ++s) { mbar_init(full(s), 1); mbar_init(conv(s), 2); mbar_init(empty(s), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tfull(i), 1); mbar_init(tempty(i), 8); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(smem_u32((const void *)tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int wg = __shfl_sync(0xffffffffu, tid >> 7, 0);             // warpgroup index, uniform as far as ptxas can tell

    // item -> (m tile, n tile, reduction slice): consecutive items share the m tile (A stays in L2)
    auto decode = [&](int it, int &mt, int &nt, int &ks) { nt = it % n_nt; ks = (it / n_nt) % a.ksplit; mt = it / (n_nt * a.ksplit); };

    if (wg == 0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(G_REGS_PRODUCER));
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                int mt, nt, ks;
                decode(it, mt, nt, ks);
                const int kc0 = ks * kc_per, kc1 = min(kc0 + kc_per, n_kc_all);
                for (int kc = kc0; kc < kc1; ++kc) {
                    mbar_wait_bounded(empty(s), ph ^ 1);
                    const uint32_t st = sbase + s * G_STAGE;
                    mbar_expect_tx(full(s), H_A_BYTES + H_B_BYTES);
                    if (A_MN) {
                        for (int j = 0; j < HM / 32; ++j) tma_load_2d(st + j * 4096, &tmA, mt * HM + 32 * j, kc * HK, full(s));
                    } else {
                        tma_load_2d(st, &tmA, kc * HK, mt * HM, full(s));
                    }
                    const uint32_t sb = st + 2 * H_A_BYTES;
                    if (B_MN) {
                        for (int j = 0; j < HN / 32; ++j) tma_load_2d(sb + j * 4096, &tmB, nt * HN + 32 * j, kc * HK, full(s));
                    } else {
                        tma_load_2d(sb, &tmB, kc * HK, nt * HN, full(s));
                    }
                    if (++s == G_NSTAGE) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32_major(HM, HN, A_MN, B_MN);
            int s = 0; uint32_t ph = 0;
            int acc = 0; uint32_t aph = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                int mt, nt, ks;
                decode(it, mt, nt, ks);
                const int kc0 = ks * kc_per, kc1 = min(kc0 + kc_per, n_kc_all);
                for (int sg0 = kc0; sg0 < kc1; sg0 += G_SEG) {
                    const int sg1 = min(sg0 + G_SEG, kc1);
                    mbar_wait_bounded(tempty(acc), aph ^ 1);               // the epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t tacc = tmem_base + (uint32_t)(acc * HN);
                    for (int kc = sg0; kc < sg1; ++kc) {
                        const uint32_t st = sbase + s * G_STAGE;
                        const uint32_t a_hi = st, a_lo = st + H_A_BYTES, b_hi = st + 2 * H_A_BYTES, b_lo = b_hi + H_B_BYTES;
                        // 8 reduction indices per instruction: 32 bytes along a K-major row, 8 rows (1024 B) of an MN-major tile
                        auto adesc = [&](uint32_t base, int k) {
                            return A_MN ? umma_desc_mn(base + k * 1024, a.mn_desc) : umma_desc_k128(base + k * 32);
                        };
                        auto bdesc = [&](uint32_t base, int k) {
                            return B_MN ? umma_desc_mn(base + k * 1024, a.mn_desc) : umma_desc_k128(base + k * 32);
                        };
                        if (TRUNC) {
                            mbar_wait_bounded(full(s), ph);                // landed: the raw tiles are the hi halves
                            tc_fence_after();
#pragma unroll
                            for (int k = 0; k < HK / 8; ++k)
                                umma_tf32(tacc, adesc(a_hi, k), bdesc(b_hi, k), idesc, (kc > sg0 || k > 0) ? 1u : 0u);
                        }
                        mbar_wait_bounded(conv(s), ph);                    // the lo halves (RNA split: hi as well) are written
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < HK / 8; ++k) {
                            umma_tf32(tacc, adesc(a_lo, k), bdesc(b_hi, k), idesc, (TRUNC || kc > sg0 || k > 0) ? 1u : 0u);
                            umma_tf32(tacc, adesc(a_hi, k), bdesc(b_lo, k), idesc, 1);
                            if (!TRUNC) umma_tf32(tacc, adesc(a_hi, k), bdesc(b_hi, k), idesc, 1);
                        }
                        umma_commit(empty(s));
                        if (++s == G_NSTAGE) { s = 0; ph ^= 1; }
                    }
                    umma_commit(tfull(acc));
                    if (++acc == 2) { acc = 0; aph ^= 1; }
                }
            }
        }
    } else {
        // ===================== converters: fp32 tile -> hi (in place) + lo =====================
        const int ct = tid - 64;                                          // 0..63
        int s = 0; uint32_t ph = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            int mt, nt, ks;
            decode(it, mt, nt, ks);
            const int kc0 = ks * kc_per, kc1 = min(kc0 + kc_per, n_kc_all);
            for (int kc = kc0; kc < kc1; ++kc) {
                mbar_wait_bounded(full(s), ph);
                float4 *A4 = (float4 *)(sgen + (size_t)s * G_STAGE);
                float4 *B4 = (float4 *)(sgen + (size_t)s * G_STAGE + 2 * H_A_BYTES);
                auto split = [](float4 *hi, float4 *lo, int i) {
                    const float4 x = hi[i];
                    float4 h, l;
                    if (TRUNC) {
                        h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
                        h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
                        h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
                        h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
                        lo[i] = l;                                        // the raw tile stays as it is
                    } else {
                        uint32_t u;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.x)); h.x = __uint_as_float(u); l.x = x.x - h.x;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.y)); h.y = __uint_as_float(u); l.y = x.y - h.y;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.z)); h.z = __uint_as_float(u); l.z = x.z - h.z;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x.w)); h.w = __uint_as_float(u); l.w = x.w - h.w;
                        hi[i] = h; lo[i] = l;
                    }
                };
#pragma unroll 4
                for (int i = ct; i < (int)(H_A_BYTES / 16); i += 64) split(A4, A4 + H_A_BYTES / 16, i);
#pragma unroll 4
                for (int i = ct; i < (int)(H_B_BYTES / 16); i += 64) split(B4, B4 + H_B_BYTES / 16, i);
                fence_proxy_async_smem_cta();                             // generic-proxy writes -> tcgen05.mma (async proxy) reads
                __syncwarp();
                if (lane == 0) mbar_arrive(conv(s));
                if (++s == G_NSTAGE) { s = 0; ph ^= 1; }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(G_REGS_EPILOGUE));
        // ===================== epilogue: TMEM -> registers -> C =====================
        const int wq = warp & 3, half = (warp - 4) >> 2;                  // TMEM lane quarter, column half of the tile
        const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
        int acc = 0; uint32_t aph = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            int mt, nt, ks;
            decode(it, mt, nt, ks);
            const int kc0 = ks * kc_per, kc1 = min(kc0 + kc_per, n_kc_all);
            const int row = mt * HM + wq * 32 + lane;
            const int n0 = nt * HN + half * (HN / 2);                     // first output column of this thread
            float sum[HN / 2];
#pragma unroll
            for (int i = 0; i < HN / 2; ++i) sum[i] = 0.f;                // (an empty reduction slice stores zeros)
            for (int sg0 = kc0; sg0 < kc1; sg0 += G_SEG) {
                mbar_wait_bounded(tfull(acc), aph);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(acc * HN + half * (HN / 2)) + lane_sel;
#pragma unroll
                for (int c = 0; c < HN / 64; ++c) {
                    float v[32];
                    tmem_ld32(tacc + c * 32, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) sum[c * 32 + i] += v[i];
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty(acc));
                if (++acc == 2) { acc = 0; aph ^= 1; }
            }
            if (row < a.Mc && n0 < a.Nc) {
                float *crow = a.C + (size_t)ks * a.split_stride + (size_t)row * a.ldc + n0;
                if (n0 + HN / 2 <= a.Nc) {
#pragma unroll
                    for (int i = 0; i < HN / 2; i += 4)
                        *(float4 *)(crow + i) = make_float4(sum[i], sum[i + 1], sum[i + 2], sum[i + 3]);
                } else {
#pragma unroll
                    for (int i = 0; i < HN / 2; ++i) if (n0 + i < a.Nc) crow[i] = sum[i];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// sum of `ksplit` partial results: out[i] = sum_s part[s * stride + i]
__global__ void __launch_bounds__(256) k_sum_partials(const float *__restrict__ part, float *__restrict__ out, size_t n4,
                                                      int ksplit, size_t stride4) {
    const size_t st = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += st) {
        float4 acc = ((const float4 *)part)[i];
        for (int s = 1; s < ksplit; ++s) {
            const float4 x = ((const float4 *)part)[(size_t)s * stride4 + i];
            acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
        }
        ((float4 *)out)[i] = acc;
    }
}

}  // namespace ctcb200
