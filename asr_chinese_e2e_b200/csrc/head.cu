// C-ABI of the fused CTC head (include/ctcb200.h, ctcb200_head_*): TMA tensor maps, workspace carve-up and the
// launch sequence  prep -> [operand split] -> k_head<pass 1> -> lattice -> k_head<pass 2>.
#include "../../include/ctcb200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "gemm_tf32x3.cuh"
#include "head_kernels.cuh"
#include "internal.h"
#include "layout.h"

using namespace ctcb200;

namespace {

constexpr size_t kSmemBudget = 226 * 1024;

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point: the library does not link libcuda, so it still
// loads (and its argument validation still runs) on a machine without a driver
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// cuTensorMapEncodeTiled needs a current driver context in the CALLING thread (CUDA_ERROR_INVALID_CONTEXT otherwise):
// torch's autograd worker threads have a current device but may not have touched the runtime yet, so bind the primary
// context of the current device here (cudaSetDevice does that since CUDA 12)
int ensure_context() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaSetDevice(dev) != cudaSuccess) return CTCB200_ERR_NO_DEVICE;
    return 0;
}

// fp32 matrix [rows, K] row-major (K contiguous) -> boxes of box_rows x 32 floats (128 B), SWIZZLE_128B; rows past the
// end read as zeros
int make_map(CUtensorMap *m, const float *base, int64_t rows, int K, int box_rows) {
    EncodeTiledFn f = encode_fn();
    if (!f) return CTCB200_ERR_NO_DEVICE;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {(cuuint32_t)HK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = f(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : CTCB200_ERR_SHAPE;
}

// fp32 matrix [rows, inner] with a row pitch (floats) -> boxes of box_rows x 32 floats, zero fill outside.  K-major
// operand tiles (mn = false) use SWIZZLE_128B, MN-major ones SWIZZLE_128B_ATOM_32B (gemm_tf32x3.cuh)
int make_map_pitched(CUtensorMap *m, const float *base, int64_t rows, int64_t inner, int64_t pitch, int box_rows, bool mn) {
    EncodeTiledFn f = encode_fn();
    if (!f) return CTCB200_ERR_NO_DEVICE;
    const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle swz = mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    if (mn && internal_g3_opt(0) > 0) swz = (CUtensorMapSwizzle)internal_g3_opt(0);
    const CUresult r = f(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS && getenv("CTCB200_DEBUG"))
        fprintf(stderr, "ctcb200: cuTensorMapEncodeTiled -> %d (base %p rows %lld inner %lld pitch %lld box_rows %d mn %d)\n",
                (int)r, (const void *)base, (long long)rows, (long long)inner, (long long)pitch, box_rows, (int)mn);
    return r == CUDA_SUCCESS ? 0 : CTCB200_ERR_SHAPE;
}

uint64_t mn_desc_bits() {
    const int lbo = internal_g3_opt(1), sbo = internal_g3_opt(2), layout = internal_g3_opt(3);
    return umma_desc_mn_hi(lbo > 0 ? lbo : 4096, sbo > 0 ? sbo : 512, layout > 0 ? layout : 1);
}

template <bool A_MN, bool B_MN, bool TRUNC>
int launch_gemm3_t(int sms, cudaStream_t s, const CUtensorMap &ma, const CUtensorMap &mb, const GemmArgs &ga) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm3<A_MN, B_MN, TRUNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
    if (e != cudaSuccess) return (int)e;
    const int items = ((ga.Mc + HM - 1) / HM) * ((ga.Nc + HN - 1) / HN) * ga.ksplit;
    k_gemm3<A_MN, B_MN, TRUNC><<<items < sms ? items : sms, G_THREADS, G_SMEM, s>>>(ma, mb, ga);
    return (int)cudaGetLastError();
}
template <bool A_MN, bool B_MN>
int launch_gemm3(int sms, cudaStream_t s, const CUtensorMap &ma, const CUtensorMap &mb, const GemmArgs &ga) {
    return internal_g3_opt(4) ? launch_gemm3_t<A_MN, B_MN, false>(sms, s, ma, mb, ga)      // option g3_rna_split
                              : launch_gemm3_t<A_MN, B_MN, true>(sms, s, ma, mb, ga);
}

struct HeadWs {
    Workspace ctc;
    size_t enc_hi, enc_lo, w_hi, w_lo, total;
};

HeadWs head_layout(int B, int T, int V, int K, const Geom &g, int precision) {
    HeadWs h;
    h.ctc = workspace_layout(B, T, g);
    size_t o = h.ctc.total;
    const size_t ne = align_up((size_t)(B > 0 ? B : 1) * T * K * 4), nw = align_up((size_t)V * K * 4);
    h.enc_hi = h.enc_lo = h.w_hi = h.w_lo = 0;
    if (precision == CTCB200_HEAD_3XTF32 && internal_g3_opt(5) == 0) {       // pre-split operands (not with option head_inring)
        h.enc_hi = o; o += ne;
        h.enc_lo = o; o += ne;
        h.w_hi = o;   o += nw;
        h.w_lo = o;   o += nw;
    }
    h.total = o;
    return h;
}

int check_head(const float *enc, const float *weight, const void *targets, const void *in_len, const void *tgt_len,
               int B, int T, int V, int K, int Umax, int blank, int precision, const void *ws, size_t ws_bytes, Geom *g,
               HeadWs *h) {
    if (B < 0 || T < 1 || V < 2 || K < HK || (K % HK) != 0 || (long long)B * T > 0x7fffff00LL) return CTCB200_ERR_SHAPE;
    if (blank < 0 || blank >= V) return CTCB200_ERR_BLANK;
    if (!geom_for(Umax, g)) return CTCB200_ERR_UMAX;
    if (precision != CTCB200_HEAD_3XTF32 && precision != CTCB200_HEAD_TF32) return CTCB200_ERR_OPTION;
    if (!enc || !weight || !targets || !in_len || !tgt_len || !ws) return CTCB200_ERR_NULL;
    if (((uintptr_t)enc & 15) || ((uintptr_t)weight & 15) || ((uintptr_t)ws & 255)) return CTCB200_ERR_ALIGN;
    *h = head_layout(B, T, V, K, *g, precision);
    if (ws_bytes < h->total) return CTCB200_ERR_WORKSPACE;
    return 0;
}

template <int NPASS, bool GRADPASS, bool INRING = false>
int launch_head(int sms, cudaStream_t s, const CUtensorMap &a_hi, const CUtensorMap &a_lo, const CUtensorMap &b_hi,
                const CUtensorMap &b_lo, const HeadArgs &args) {
    constexpr size_t smem = HeadCfg<NPASS>::SMEM;
    static_assert(smem <= kSmemBudget, "shared memory budget");
    cudaError_t e = cudaFuncSetAttribute(k_head<NPASS, GRADPASS, INRING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int M = args.B * args.T, n_tiles = (M + HM - 1) / HM;
    const int grid = n_tiles < sms ? n_tiles : sms;
    k_head<NPASS, GRADPASS, INRING><<<grid, H_THREADS, smem, s>>>(a_hi, a_lo, b_hi, b_lo, args);
    return (int)cudaGetLastError();
}

int head_impl(bool want_grad, const float *enc, const float *weight, const float *bias, const int64_t *targets,
              int64_t targets_stride, int64_t targets_numel, const int64_t *in_len, const int64_t *tgt_len, int B, int T,
              int V, int K, int Umax, int blank, int flags, int precision, int reduction, float inv_batch, float *nll,
              float *loss_sums, float *dlogits, int64_t dlogits_pitch, void *workspace, size_t workspace_bytes,
              ctcb200_stream_t stream) {
    Geom g;
    HeadWs h;
    int rc = check_head(enc, weight, targets, in_len, tgt_len, B, T, V, K, Umax, blank, precision, workspace,
                        workspace_bytes, &g, &h);
    if (rc) return rc;
    if (!nll) return CTCB200_ERR_NULL;
    if (targets_stride < 0 || targets_numel < 0) return CTCB200_ERR_SHAPE;
    if (flags & CTCB200_FLAG_DECODE) return CTCB200_ERR_OPTION;          // no per-frame argmax on this path (yet)
    if (want_grad) {
        if (!dlogits) return CTCB200_ERR_NULL;
        if ((uintptr_t)dlogits & 15) return CTCB200_ERR_ALIGN;
        if (dlogits_pitch < V || (dlogits_pitch & 3) || dlogits_pitch >= (int64_t)V + 32) return CTCB200_ERR_SHAPE;
        if (reduction < 0 || reduction > 2) return CTCB200_ERR_REDUCTION;
    }
    if (B == 0) return CTCB200_OK;
    int sms = 0;
    if ((rc = internal_sm_count(&sms)) || (rc = ensure_context())) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char *ws = (unsigned char *)workspace;
    const int64_t tnumel = targets_stride ? (int64_t)B * targets_stride : targets_numel;
    const int zero_inf = flags & 1;

    if ((rc = internal_prep(in_len, tgt_len, targets_stride, B, T, Umax, workspace, h.ctc, s))) return rc;

    // 3xTF32: round-to-nearest pre-split in HBM (default) or, option head_inring = 1, the in-ring truncation split
    const bool presplit = precision == CTCB200_HEAD_3XTF32 && internal_g3_opt(5) == 0;
    const bool inring = precision == CTCB200_HEAD_3XTF32 && !presplit;
    const float *a_hi = enc, *a_lo = enc, *b_hi = weight, *b_lo = weight;
    if (presplit) {
        float *eh = (float *)(ws + h.enc_hi), *el = (float *)(ws + h.enc_lo), *wh = (float *)(ws + h.w_hi),
              *wl = (float *)(ws + h.w_lo);
        const size_t ne4 = (size_t)B * T * K / 4, nw4 = (size_t)V * K / 4;       // K % 32 == 0
        k_split_tf32<<<sms * 8, 256, 0, s>>>(enc, eh, el, ne4);
        k_split_tf32<<<sms * 2, 256, 0, s>>>(weight, wh, wl, nw4);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        a_hi = eh; a_lo = el; b_hi = wh; b_lo = wl;
    }
    CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo;
    if ((rc = make_map(&mA_hi, a_hi, (int64_t)B * T, K, HM)) || (rc = make_map(&mA_lo, a_lo, (int64_t)B * T, K, HM)) ||
        (rc = make_map(&mB_hi, b_hi, V, K, HN)) || (rc = make_map(&mB_lo, b_lo, V, K, HN)))
        return rc;

    HeadArgs a;
    a.bias = bias; a.targets = targets; a.tnumel = tnumel;
    a.Tb = (const int *)(ws + h.ctc.Tb); a.Ub = (const int *)(ws + h.ctc.Ub); a.toff = (const int64_t *)(ws + h.ctc.toff);
    a.flags = (const int *)(ws + h.ctc.flags); a.slow = (int *)(ws + h.ctc.slow); a.bad = (int *)(ws + h.ctc.bad);
    a.hdr = (int *)(ws + h.ctc.hdr);
    a.lp_lab = (float *)(ws + h.ctc.lp_lab); a.gam = (const float *)(ws + h.ctc.gam);
    a.dlogits = dlogits; a.pitch = dlogits_pitch;
    a.B = B; a.T = T; a.V = V; a.K = K; a.Lp = g.Lp; a.blank = blank; a.zero_inf = zero_inf; a.reduction = reduction;
    a.inv_batch = inv_batch; a.lin_thr = internal_lin_thr(g, flags); a.occ_skip = internal_occ_skip();

    rc = inring     ? launch_head<3, false, true>(sms, s, mA_hi, mA_lo, mB_hi, mB_lo, a)
         : presplit ? launch_head<3, false>(sms, s, mA_hi, mA_lo, mB_hi, mB_lo, a)
                    : launch_head<1, false>(sms, s, mA_hi, mA_lo, mB_hi, mB_lo, a);
    if (rc) return rc;
    const float mean_scale = want_grad ? inv_batch : 1.f / (float)B;
    if ((rc = internal_lattice(want_grad, targets, tnumel, B, T, V, zero_inf, nll, loss_sums, mean_scale, workspace, h.ctc,
                               g, s)))
        return rc;
    if (!want_grad) return CTCB200_OK;
    return inring     ? launch_head<3, true, true>(sms, s, mA_hi, mA_lo, mB_hi, mB_lo, a)
           : presplit ? launch_head<3, true>(sms, s, mA_hi, mA_lo, mB_hi, mB_lo, a)
                      : launch_head<1, true>(sms, s, mA_hi, mA_lo, mB_hi, mB_lo, a);
}

size_t param_grads_ws(int V, int K, int *ksplit) {
    *ksplit = 2;                                            // 34 class tiles x 2 column tiles x 2 slices = 136 work items
    return align_up((size_t)(*ksplit) * V * K * 4);
}

}  // namespace

extern "C" {

int ctcb200_head_param_grads_workspace_bytes(int V, int K, size_t *out_bytes) {
    if (!out_bytes) return CTCB200_ERR_NULL;
    if (V < 2 || K < HK || (K % HK) != 0) return CTCB200_ERR_SHAPE;
    int ks;
    *out_bytes = param_grads_ws(V, K, &ks);
    return CTCB200_OK;
}

int ctcb200_head_param_grads(const float *dlogits, int64_t dlogits_pitch, const float *enc, const float *weight, int B,
                             int T, int V, int K, float *d_enc, float *d_weight, void *workspace, size_t workspace_bytes,
                             ctcb200_stream_t stream) {
    if (!dlogits || !enc || !weight || (!d_enc && !d_weight)) return CTCB200_ERR_NULL;
    if (B < 0 || T < 1 || V < 2 || K < HK || (K % HK) != 0 || (long long)B * T > 0x7fffff00LL) return CTCB200_ERR_SHAPE;
    if (dlogits_pitch < V || (dlogits_pitch & 3)) return CTCB200_ERR_SHAPE;
    if (((uintptr_t)dlogits & 15) || ((uintptr_t)enc & 15) || ((uintptr_t)weight & 15) || ((uintptr_t)d_enc & 15) ||
        ((uintptr_t)d_weight & 15) || ((uintptr_t)workspace & 255))
        return CTCB200_ERR_ALIGN;
    int ksplit;
    const size_t need = d_weight ? param_grads_ws(V, K, &ksplit) : 0;
    if (d_weight && (!workspace || workspace_bytes < need)) return CTCB200_ERR_WORKSPACE;
    if (B == 0) return CTCB200_OK;
    int sms = 0, rc;
    if ((rc = internal_sm_count(&sms)) || (rc = ensure_context())) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t M = (int64_t)B * T;
    if (d_enc) {
        // d enc[M, K] = dlogits[M, V] x W[V, K]:  A = dlogits, K-major (reduction = class index, contiguous);
        //                                         B = W as [N = K columns, reduction = V rows]: MN-major
        CUtensorMap ma, mb;
        if ((rc = make_map_pitched(&ma, dlogits, M, dlogits_pitch, dlogits_pitch, HM, false)) ||
            (rc = make_map_pitched(&mb, weight, V, K, K, 32, true)))
            return rc;
        const GemmArgs ga = {d_enc, K, (int)M, K, (int)dlogits_pitch, 1, 0, mn_desc_bits()};
        if ((rc = launch_gemm3<false, true>(sms, s, ma, mb, ga))) return rc;
    }
    if (d_weight) {
        // d W[V, K] = dlogits^T x enc:  A = dlogits as [M' = V, reduction = M rows]: MN-major;
        //                               B = enc as [N = K, reduction = M rows]: MN-major; reduction split in `ksplit` slices
        CUtensorMap ma, mb;
        if ((rc = make_map_pitched(&ma, dlogits, M, dlogits_pitch, dlogits_pitch, 32, true)) ||
            (rc = make_map_pitched(&mb, enc, M, K, K, 32, true)))
            return rc;
        int ks = ksplit;
        const int n_kc = (int)((M + HK - 1) / HK);
        if (ks > n_kc) ks = 1;
        float *part = (float *)workspace;
        const GemmArgs ga = {ks > 1 ? part : d_weight, K, V, K, (int)M, ks, (int64_t)V * K, mn_desc_bits()};
        if ((rc = launch_gemm3<true, true>(sms, s, ma, mb, ga))) return rc;
        if (ks > 1) {
            const size_t n4 = (size_t)V * K / 4;
            k_sum_partials<<<sms * 2, 256, 0, s>>>(part, d_weight, n4, ks, n4);
            if ((rc = (int)cudaGetLastError())) return rc;
        }
    }
    return CTCB200_OK;
}


int ctcb200_head_workspace_bytes(int B, int T, int V, int K, int Umax, int precision, size_t *out_bytes) {
    if (!out_bytes) return CTCB200_ERR_NULL;
    if (B < 0 || T < 1 || V < 2 || K < HK || (K % HK) != 0) return CTCB200_ERR_SHAPE;
    if (precision != CTCB200_HEAD_3XTF32 && precision != CTCB200_HEAD_TF32) return CTCB200_ERR_OPTION;
    Geom g;
    if (!geom_for(Umax, &g)) return CTCB200_ERR_UMAX;
    *out_bytes = head_layout(B, T, V, K, g, precision).total;
    return CTCB200_OK;
}

int ctcb200_head_loss(const float *enc, const float *weight, const float *bias, const int64_t *targets,
                      int64_t targets_stride, int64_t targets_numel, const int64_t *in_len, const int64_t *tgt_len, int B,
                      int T, int V, int K, int Umax, int blank, int flags, int precision, float *nll, float *loss_sums,
                      void *workspace, size_t workspace_bytes, ctcb200_stream_t stream) {
    return head_impl(false, enc, weight, bias, targets, targets_stride, targets_numel, in_len, tgt_len, B, T, V, K, Umax,
                     blank, flags, precision, 0, 0.f, nll, loss_sums, nullptr, 0, workspace, workspace_bytes, stream);
}

int ctcb200_head_loss_grad(const float *enc, const float *weight, const float *bias, const int64_t *targets,
                           int64_t targets_stride, int64_t targets_numel, const int64_t *in_len, const int64_t *tgt_len,
                           int B, int T, int V, int K, int Umax, int blank, int flags, int precision, int reduction,
                           float inv_batch, float *nll, float *loss_sums, float *dlogits, int64_t dlogits_pitch,
                           void *workspace, size_t workspace_bytes, ctcb200_stream_t stream) {
    return head_impl(true, enc, weight, bias, targets, targets_stride, targets_numel, in_len, tgt_len, B, T, V, K, Umax,
                     blank, flags, precision, reduction, inv_batch, nll, loss_sums, dlogits, dlogits_pitch, workspace,
                     workspace_bytes, stream);
}

}  // extern "C"
