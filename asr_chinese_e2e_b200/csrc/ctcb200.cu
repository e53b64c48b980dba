// C-ABI of libctcb200.so (declared in include/ctcb200.h): argument validation, workspace
// carve-up, launch configuration.  All device work is asynchronous on the caller's stream.
#include "../../include/ctcb200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ce_kernel.cuh"
#include "decode_kernel.cuh"
#include "internal.h"
#include "lattice_lin.cuh"
#include "layout.h"
#include "stream_kernels.cuh"
#include "sweep_direct.cuh"

using namespace ctcb200;

namespace {

constexpr int kMaxV = 16384;
constexpr size_t kSmemBudget = 226 * 1024;   // leave 1 KB of the 227 KB opt-in limit to the driver

// Developer tunables.  Every knob is read from the environment ONCE, when the library is loaded; afterwards
// ctcb200_set_option / ctcb200_get_option (include/ctcb200.h) are the only way to change one.  0 = "auto" for the
// launch-shape knobs (the tuned default of the kernel in question).
enum OptId {
    OPT_PDL, OPT_LATTICE_LOG, OPT_LIN_THR, OPT_K1F_NT, OPT_K1F_NST, OPT_K1F_CPS, OPT_K1_NT, OPT_K1_NST, OPT_K1_CPS,
    OPT_K3_NT, OPT_K3_NST, OPT_K3_CPS, OPT_CE_NST, OPT_CE_CPS, OPT_K3P_CPS, OPT_OCC_SKIP_BITS,
    OPT_ZERO_IN_LATTICE, OPT_ZERO_CPS, OPT_SKIP_LATTICE, OPT_LABEL_KEEP_L2, OPT_SWEEP_DIRECT, OPT_K1D_CPS,
    OPT_G3_SWZ, OPT_G3_LBO, OPT_G3_SBO, OPT_G3_LAYOUT, OPT_G3_RNA_SPLIT, OPT_HEAD_INRING, OPT_SCRATCH_POLICY, OPT_K1P_BULKST, OPT_K1P_STORE_POLICY, OPT_COUNT
};
struct Opt { const char *name, *env; int value; };
Opt g_opt[OPT_COUNT] = {
    // bit 0: sweep after prep (measured: +70 us per step in a back-to-back loop; also not safe as is: the sweep reads
    // k0_prep's arrays through const __restrict__ pointers, i.e. possibly non-coherent loads), bit 1: lattice after
    // sweep, bit 2: patch after lattice (its class tables are built while the lattice drains: -5 us), bit 3: also when
    // the patch is launched as a stage of its own behind an event record (sharded_ctc_loss splits the call there so
    // that its all-reduce can start; griddepcontrol.wait keeps it correct whatever the runtime makes of the attribute)
    {"pdl", "CTCB200_PDL", 14},
    {"lattice_log", "CTCB200_LATTICE_LOG", 0},          // 1: log-space recursion for every utterance
    {"lin_thr", "CTCB200_LIN_THR", 0},                  // range limit of the linear-domain lattice (bits; 0 = auto)
    {"k1f_nt", "CTCB200_K1F_NT", 0}, {"k1f_nst", "CTCB200_K1F_NST", 0}, {"k1f_cps", "CTCB200_K1F_CPS", 0},
    {"k1_nt", "CTCB200_K1_NT", 0}, {"k1_nst", "CTCB200_K1_NST", 0}, {"k1_cps", "CTCB200_K1_CPS", 0},
    {"k3_nt", "CTCB200_K3_NT", 0}, {"k3_nst", "CTCB200_K3_NST", 0}, {"k3_cps", "CTCB200_K3_CPS", 0},
    {"ce_nst", "CTCB200_CE_NST", 0}, {"ce_cps", "CTCB200_CE_CPS", 0},
    {"k3p_cps", "CTCB200_K3P_CPS", 32},
    {"occ_skip_bits", "CTCB200_OCC_SKIP_BITS", 40},     // the patch skips occupancies <= 2^-bits (0: exact zeros only)
    {"zero_in_lattice", "CTCB200_ZERO_IN_LATTICE", 0}, {"zero_cps", "CTCB200_ZERO_CPS", 2},
    {"skip_lattice", "CTCB200_DEBUG_SKIP_LATTICE", 0},  // profiling aid: time the sweep alone
    // fused sweep: evict_last for the gradient chunks the sparse patch revisits.  Measured on B200 (round 2): the patch
    // still misses L2 (104 MB of DRAM reads either way) and the sweep gets 4 us slower -> off
    {"label_keep_l2", "CTCB200_LABEL_KEEP_L2", 0},
    // which sweep kernel: 0 = auto (k1p_sweep when frames pair up into 16-byte-aligned groups -- even V, even T --
    // else k1_lse_gather), 1 = k1d_sweep (aligned groups, direct LDG.128 loads, 4 CTAs/SM: the same speed as
    // k1_lse_gather, 0.618 vs 0.619 ms at C2 full lengths), 2 = k1p_sweep wherever it applies (aligned groups behind a
    // bulk-TMA ring, 2 CTAs/SM, both frames of a group carried through the reductions interleaved: 0.587 vs 0.621 ms),
    // 3 = k1_lse_gather (the round-1 kernel: one frame per ring slot, hull copies)
    {"sweep_direct", "CTCB200_SWEEP_DIRECT", 0},
    {"k1d_cps", "CTCB200_K1D_CPS", 0},                  // CTAs per SM of k1d_sweep (0 = 4, or 2 for wide vocabularies)
    // k_gemm3 (gemm_tf32x3.cuh), MN-major operand tiles: tensor-map swizzle enum, descriptor LBO / SBO (bytes) and layout
    // type (0 = the kernel's defaults: SWIZZLE_128B_ATOM_32B boxes, LBO 4096, SBO 512, layout SWIZZLE_128B_BASE32B)
    {"g3_swz", "CTCB200_G3_SWZ", 0}, {"g3_lbo", "CTCB200_G3_LBO", 0}, {"g3_sbo", "CTCB200_G3_SBO", 0},
    {"g3_layout", "CTCB200_G3_LAYOUT", 0},
    // k_gemm3: 1 = round-to-nearest hi/lo split written by the converter warps before any MMA of the stage (the first
    // version); 0 = truncation split, the raw tile is the hi half and a third of the MMAs runs under the conversion
    {"g3_rna_split", "CTCB200_G3_RNA_SPLIT", 0},
    // k_head 3xTF32: 0 = operands pre-split in HBM by k_split_tf32 (round-to-nearest), 1 = in-ring truncation split as in
    // k_gemm3 (no pre-pass, no 2 x 4*B*T*K bytes of workspace).  Measured on B200 at the C2 shape: evaluation 2.05-2.12 ms
    // pre-split vs 2.32-2.42 ms in-ring (the forward pass is tensor-bound and the converters' shared-memory traffic
    // slows the MMA), training 9.25-9.37 vs 9.36-9.49 ms -> pre-split stays the default
    {"head_inring", "CTCB200_HEAD_INRING", 0},
    // L2 policy of the scratch arrays between the kernels of a step: 0 evict_last, 1 evict_normal, 2 evict_first
    {"scratch_policy", "CTCB200_SCRATCH_POLICY", 0},
    // k1p_sweep: 1 (default) = the gradient of a group goes back into its ring slot and leaves through ONE bulk-TMA
    // store (cp.async.bulk.global.shared) instead of 17 STG.128 per thread: 0.586 -> 0.572 ms (C2 full), 0.5195 -> 0.507
    {"k1p_bulkst", "CTCB200_K1P_BULKST", 1},
    {"k1p_store_policy", "CTCB200_K1P_STORE_POLICY", 0},  // L2 policy of those bulk stores: 0 evict_first, 1 normal, 2 evict_last
};
const bool g_opt_loaded = [] {
    for (Opt &o : g_opt) {
        const char *s = getenv(o.env);
        if (s && *s) o.value = atoi(s);
    }
    return true;
}();
inline int opt(OptId id) { return g_opt[id].value; }
inline int opt_or(OptId id, int dflt) { return g_opt[id].value > 0 ? g_opt[id].value : dflt; }

// push the scratch-array L2 policy to the device when the option changed (stream-ordered with the launches after it)
int apply_scratch_policy(cudaStream_t s) {
    static int applied = 0;
    const int want = opt(OPT_SCRATCH_POLICY);
    if (want == applied) return 0;
    const uint64_t pol = want == 1 ? kEvictNormal : (want == 2 ? kEvictFirst : kEvictLast);
    cudaError_t e = cudaMemcpyToSymbolAsync(c_scratch_policy, &pol, sizeof(pol), 0, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(s);                               // `pol` is a stack variable
    if (e != cudaSuccess) return (int)e;
    applied = want;
    return 0;
}
int apply_gstore_policy(cudaStream_t s) {
    static int applied = 0;
    const int want = opt(OPT_K1P_STORE_POLICY);
    if (want == applied) return 0;
    const uint64_t pol = want == 1 ? kEvictNormal : (want == 2 ? kEvictLast : kEvictFirst);
    cudaError_t e = cudaMemcpyToSymbolAsync(c_gstore_policy, &pol, sizeof(pol), 0, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return (int)e;
    applied = want;
    return 0;
}

struct DevInfo {
    int sms;
    int cc_major;
};
int device_info(DevInfo *d) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return CTCB200_ERR_NO_DEVICE;
    if (cudaDeviceGetAttribute(&d->sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&d->cc_major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
        return CTCB200_ERR_NO_DEVICE;
    if (d->cc_major != 10) return CTCB200_ERR_NO_DEVICE;   // sm_100a SASS only: no other path exists
    return 0;
}

int check_common(const void *logits, const void *targets, const void *in_len, const void *tgt_len,
                 int B, int T, int V, int Umax, int blank, const void *workspace, size_t ws_bytes,
                 Geom *g, Workspace *w) {
    if (B < 0 || T < 1 || V < 2 || V > kMaxV) return CTCB200_ERR_SHAPE;
    if (blank < 0 || blank >= V) return CTCB200_ERR_BLANK;
    if (!geom_for(Umax, g)) return CTCB200_ERR_UMAX;
    if (!logits || !targets || !in_len || !tgt_len || !workspace) return CTCB200_ERR_NULL;
    if (((uintptr_t)logits & 15) || ((uintptr_t)workspace & 255)) return CTCB200_ERR_ALIGN;
    *w = workspace_layout(B, T, *g);
    if (ws_bytes < w->total) return CTCB200_ERR_WORKSPACE;
    return 0;
}

struct StreamCfg {
    int nst;
    uint32_t slot_bytes, stage_bytes;
    size_t smem;
    int grid;
};
// Ring geometry of the two sweep kernels.  stage_extra = bytes appended to every stage, fixed_extra =
// per-CTA tail.  Defaults (env-overridable for experiments) were picked by sweeping on a B200
// (profiles/): the read-only sweep likes many small CTAs with a shallow ring, the read+write sweep
// fewer CTAs with a deeper ring; both leave >= 52 KB of shared memory per SM free so that a lattice
// CTA of another utterance chunk can be co-resident.
constexpr size_t kLatticeReserve = 52 * 1024;
int stream_cfg(int V, uint32_t stage_extra, size_t fixed_extra, int sms, int dflt_nst, int dflt_cps,
               OptId opt_nst, OptId opt_cps, StreamCfg *c) {
    c->slot_bytes = (uint32_t)align_up((size_t)V * 4 + 32, 128);
    c->stage_bytes = c->slot_bytes + stage_extra;
    int nst = opt_or(opt_nst, dflt_nst);
    if (nst < 2) nst = 2;
    if (nst > 8) nst = 8;
    while (nst > 2 && (size_t)nst * c->stage_bytes + fixed_extra + 8 * nst > kSmemBudget) --nst;
    c->nst = nst;
    c->smem = (size_t)nst * c->stage_bytes + 8 * nst + fixed_extra;
    if (c->smem > kSmemBudget) return CTCB200_ERR_SHAPE;
    int cps = (int)((kSmemBudget - kLatticeReserve) / (c->smem + 1024));   // + per-CTA reserved shared memory
    if (cps > dflt_cps) cps = dflt_cps;
    if (cps < 1) cps = 1;
    cps = opt_or(opt_cps, cps);
    if (cps < 1) cps = 1;
    c->grid = sms * cps;
    return 0;
}

// Launch with programmatic stream serialization (PDL): the kernel may be scheduled while its predecessor on
// the stream drains; every such kernel calls griddep_wait() before touching the predecessor's results.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(int edge, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                       Args... args) {
    // edge < 0: plain launch (the previous operation on the stream is not the producer kernel of this call)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = edge >= 0 ? (opt(OPT_PDL) >> edge) & 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

struct K1Args {
    const float *logits; const int64_t *targets; int64_t tnumel; const int *Tb, *Ub; const int64_t *toff;
    const int *rowstart; float *lp_lab; int *hdr; int B, T, V, Lp, blank;
    float *grad; int reduction; float inv_batch;   // fused (2-sweep) mode only
    int *best; int zero_pad_here; int *slow; float lin_thr; int *bad;
};
struct K3Args {
    const float *logits; const int64_t *targets; int64_t tnumel; const int *Tb, *Ub; const int64_t *toff;
    const int *flags, *rowstart; const float *gam, *grad_out; int64_t go_stride; int reduction; float inv_batch;
    float *grad; int B, T, V, Lp, blank, zero_inf; const int *bad;
};

template <int NT, int MAXC, bool EXACT, bool FUSED, bool DIRECT = false>
cudaError_t launch_k1x(const StreamCfg &c, cudaStream_t s, const K1Args &a) {
    cudaError_t e = cudaFuncSetAttribute(k1_lse_gather<NT, MAXC, EXACT, FUSED, DIRECT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
    if (e != cudaSuccess) return e;
    return launch_pdl(0, k1_lse_gather<NT, MAXC, EXACT, FUSED, DIRECT>, dim3(c.grid), dim3(NT), c.smem, s, a.logits, a.targets,
                      a.tnumel, a.Tb, a.Ub, a.toff, a.rowstart, a.lp_lab, a.hdr, a.B, a.T, a.V, a.Lp, a.blank, c.nst,
                      c.slot_bytes, a.grad, a.reduction, a.inv_batch, a.best, a.zero_pad_here, a.slow, a.lin_thr, a.bad,
                      (FUSED && MAXC <= 32) ? opt(OPT_LABEL_KEEP_L2) : 0);
}
template <int NT, int MAXC, bool EXACT>
cudaError_t launch_k1(const StreamCfg &c, cudaStream_t s, const K1Args &a) {
    return launch_k1x<NT, MAXC, EXACT, false>(c, s, a);
}
template <int NT, int MAXC, bool EXACT>
cudaError_t launch_k1f(const StreamCfg &c, cudaStream_t s, const K1Args &a) {
    return launch_k1x<NT, MAXC, EXACT, true>(c, s, a);
}
template <int NT, int MAXC, bool EXACT>
cudaError_t launch_k3(const StreamCfg &c, cudaStream_t s, const K3Args &a) {
    cudaError_t e = cudaFuncSetAttribute(k3_grad<NT, MAXC, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)c.smem);
    if (e != cudaSuccess) return e;
    k3_grad<NT, MAXC, EXACT><<<c.grid, NT, c.smem, s>>>(a.logits, a.targets, a.tnumel, a.Tb, a.Ub, a.toff, a.flags,
                                                        a.rowstart, a.gam, a.grad_out, a.go_stride, a.reduction,
                                                        a.inv_batch, a.grad, a.B, a.T, a.V, a.Lp, a.blank, a.zero_inf,
                                                        c.nst, c.slot_bytes, c.stage_bytes, a.bad);
    return cudaGetLastError();
}

// Instantiation for a vocabulary size: interior 16-byte chunks per row lie between imin and imax
// depending on the row's misalignment; rounds = ceil(imax / NT); EXACT when every round but the last is
// full for every row ((rounds-1)*NT <= imin) and `rounds` is one of the compiled sizes.
#define STREAM_DISPATCH(KERN, NT, rounds, exact, ...)                                                      \
    ((rounds) <= 2    ? ((exact) && (rounds) == 2 ? KERN<NT, 2, true>(__VA_ARGS__) : KERN<NT, 2, false>(__VA_ARGS__))    \
     : (rounds) <= 5  ? ((exact) && (rounds) == 5 ? KERN<NT, 5, true>(__VA_ARGS__) : KERN<NT, 5, false>(__VA_ARGS__))    \
     : (rounds) <= 9  ? ((exact) && (rounds) == 9 ? KERN<NT, 9, true>(__VA_ARGS__) : KERN<NT, 9, false>(__VA_ARGS__))    \
     : (rounds) <= 17 ? ((exact) && (rounds) == 17 ? KERN<NT, 17, true>(__VA_ARGS__) : KERN<NT, 17, false>(__VA_ARGS__)) \
                      : ((exact) && (rounds) == 33 ? KERN<NT, 33, true>(__VA_ARGS__) : KERN<NT, 33, false>(__VA_ARGS__)))

static inline void stream_pick(int V, int want_nt, int *nt, int *rounds, bool *exact) {
    const int imax = ((V + 6) >> 2) - 2, imin = ((V + 3) >> 2) - 2;
    int n = want_nt == 64 ? 64 : 128;
    int r = imax <= 0 ? 1 : (imax + n - 1) / n;
    if (n == 64 && r > 33) { n = 128; r = (imax + n - 1) / n; }   // keep the register-resident row <= 33 chunks
    *nt = n; *rounds = r;
    *exact = r >= 2 && (r - 1) * n <= imin;
}

template <int NS, bool GRAD>
cudaError_t launch_k2(cudaStream_t s, const int64_t *targets, int64_t tnumel, const int *Tb, const int *Ub,
                      const int64_t *toff, int *flags, const float *lp_lab, float *gam, float *ab, float *nll,
                      float *loss_sums, unsigned *ticket, int B, int T, int zero_inf, float *zero_grad,
                      const int *rowstart, int V, int zero_ctas, double *tile_off, float mean_scale, const int *slow,
                      size_t ab_utt, bool follows_sweep, const int *bad) {
    constexpr uint32_t smem = k2_smem_bytes<NS, GRAD>();
    cudaError_t e = cudaFuncSetAttribute(k2_lattice<NS, GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return launch_pdl(follows_sweep ? 1 : -1, k2_lattice<NS, GRAD>, dim3((B + 1) / 2 + (zero_grad ? zero_ctas : 0)), dim3(128), smem, s,
                      targets, tnumel, Tb, Ub, toff, flags, lp_lab, gam, ab, nll, loss_sums, ticket, B, T, zero_inf,
                      zero_grad, rowstart, V, tile_off, mean_scale, slow, ab_utt, bad);
}

// Every kernel of the path asks for the maximum shared-memory carveout: a launch whose carveout differs
// from the previous kernel's makes the SMs drain and reconfigure (several microseconds per launch).
template <typename K>
void prefer_max_carveout(K kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

// Frames per aligned group of the direct sweep: the smallest P with P * V * 4 a multiple of 16; 0 = not applicable
// (odd V would need groups of four frames; T must be a multiple of P so that no group spans two utterances).
int group_frames(int V, int T) {
    const int P = (V % 4 == 0) ? 1 : ((V % 2 == 0) ? 2 : 0);
    if (P == 0 || T % P != 0) return 0;
    return P;
}

template <int MAXC, bool FUSED>
cudaError_t launch_k1d(cudaStream_t s, int sms, const K1dArgs &a) {
    // default carve-out (small shared memory, large L1): direct loads are staged through the L1 data array
    cudaFuncSetAttribute(k1d_sweep<MAXC, FUSED>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutDefault);
    const int cps = opt_or(OPT_K1D_CPS, MAXC <= 17 ? 4 : 2);
    return launch_pdl(0, k1d_sweep<MAXC, FUSED>, dim3(sms * cps), dim3(128), 0, s, a);
}
// k1d_sweep launch; returns false (and launches nothing) when the shape does not fit
template <bool FUSED>
bool try_launch_k1d(cudaStream_t s, const DevInfo &dev, const K1dArgs &a, cudaError_t *err) {
    if (a.P <= 0 || opt(OPT_SWEEP_DIRECT) != 1 || a.Lp > 264) return false;
    if (((uintptr_t)a.logits & 15) || (FUSED && ((uintptr_t)a.grad & 15))) return false;
    const int nch = a.P * a.V / 4;
    if (nch <= 128 * 5) *err = launch_k1d<5, FUSED>(s, dev.sms, a);
    else if (nch <= 128 * 9) *err = launch_k1d<9, FUSED>(s, dev.sms, a);
    else if (nch <= 128 * 17) *err = launch_k1d<17, FUSED>(s, dev.sms, a);
    else if (nch <= 128 * 33) *err = launch_k1d<33, FUSED>(s, dev.sms, a);
    else return false;
    return true;
}

// k1p_sweep (sweep_direct = 2): the aligned-group sweep behind a bulk-TMA ring, two CTAs (= two frame streams) per SM
template <int NT, int MAXC, bool FUSED, bool BULKST = false>
cudaError_t launch_k1p(cudaStream_t s, int sms, const K1dArgs &a) {
    const uint32_t slot = (uint32_t)align_up((size_t)a.P * a.V * 4, 128);
    const size_t fixed = 2 * (NT / 32) * 16 + 2 * (NT / 32) * 8 + 264 * 4 + 64;
    int nst = opt_or(OPT_K1F_NST, BULKST ? 3 : 2);
    if (BULKST && nst < 3) nst = 3;
    if (nst > 8) nst = 8;
    while (nst > 2 && 2 * ((size_t)nst * slot + 8 * nst + fixed + 1024) > kSmemBudget) --nst;
    const size_t smem = (size_t)nst * slot + 8 * nst + fixed;
    if (smem > kSmemBudget) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(k1p_sweep<NT, MAXC, FUSED, BULKST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    prefer_max_carveout(k1p_sweep<NT, MAXC, FUSED, BULKST>);
    const int cps = opt_or(OPT_K1D_CPS, 2);
    return launch_pdl(0, k1p_sweep<NT, MAXC, FUSED, BULKST>, dim3(sms * cps), dim3(NT), smem, s, a, nst, slot);
}
template <bool FUSED>
bool try_launch_k1p(cudaStream_t s, const DevInfo &dev, const K1dArgs &a, cudaError_t *err) {
    // auto (0): groups of two frames only -- that is where the instruction-level parallelism comes from; 2 forces it
    const int mode = opt(OPT_SWEEP_DIRECT);
    // (auto also wants the group to fit 17 chunks per thread: two CTAs per SM with three ring slots each)
    if (a.P <= 0 || !(mode == 2 || (mode == 0 && a.P == 2 && a.P * a.V / 4 <= 128 * 17)) || a.Lp > 264) return false;
    if (((uintptr_t)a.logits & 15) || (FUSED && ((uintptr_t)a.grad & 15))) return false;
    const int nch = a.P * a.V / 4;
    if ((size_t)a.P * a.V * 4 > 100 * 1024) return false;       // two stages of a group must fit
    if (FUSED && opt(OPT_K1P_BULKST) && nch <= 128 * 17 && nch > 128 * 9 && 2 * (3 * (size_t)a.P * a.V * 4 + 4096) <= kSmemBudget) {
        *err = launch_k1p<128, 17, FUSED, true>(s, dev.sms, a);   // the gradient leaves through bulk-TMA stores
        return true;
    }
    if (nch <= 128 * 5) *err = launch_k1p<128, 5, FUSED>(s, dev.sms, a);
    else if (nch <= 128 * 9) *err = launch_k1p<128, 9, FUSED>(s, dev.sms, a);
    else if (nch <= 128 * 17) *err = launch_k1p<128, 17, FUSED>(s, dev.sms, a);
    else if (nch <= 128 * 33) *err = launch_k1p<128, 33, FUSED>(s, dev.sms, a);
    else return false;
    return true;
}

struct FusedGrad {          // non-null grad => 2-sweep mode: the sweep writes g*softmax, k3p adds -g*occupancy
    float *grad; int reduction; float inv_batch;
    int stages;             // bit 0: prep + sweep, bit 1: lattice, bit 2: sparse patch (7 = the whole call)
};

int forward_impl(bool want_grad, const FusedGrad *fg, const float *logits, const int64_t *targets, int64_t targets_stride,
                 int64_t targets_numel, const int64_t *in_len, const int64_t *tgt_len, int B, int T, int V,
                 int Umax, int blank, int zero_infinity, float *nll, float *loss_sums, void *workspace,
                 size_t workspace_bytes, ctcb200_stream_t stream, ctcb200_event_t sweep_done) {
    Geom g;
    Workspace w;
    int rc = check_common(logits, targets, in_len, tgt_len, B, T, V, Umax, blank, workspace, workspace_bytes,
                          &g, &w);
    if (rc) return rc;
    if (!nll) return CTCB200_ERR_NULL;
    if (targets_stride < 0 || targets_numel < 0) return CTCB200_ERR_SHAPE;
    if (B == 0) return CTCB200_OK;
    DevInfo dev;
    if ((rc = device_info(&dev))) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = apply_scratch_policy(s)) || (rc = apply_gstore_policy(s))) return rc;
    unsigned char *ws = (unsigned char *)workspace;
    int *hdr = (int *)(ws + w.hdr);
    int *Tb = (int *)(ws + w.Tb), *Ub = (int *)(ws + w.Ub), *flags = (int *)(ws + w.flags);
    int64_t *toff = (int64_t *)(ws + w.toff);
    int *rowstart = (int *)(ws + w.rowstart);
    int *slow = (int *)(ws + w.slow), *bad = (int *)(ws + w.bad);
    float *lp_lab = (float *)(ws + w.lp_lab), *gam = (float *)(ws + w.gam), *ab = (float *)(ws + w.ab);
    const int64_t tnumel = targets_stride ? (int64_t)B * targets_stride : targets_numel;
    // Range of the linear-domain lattice (lattice_lin.cuh): a stage of TT frames plus the NS/2 labels of one lane
    // may shrink a value by (TT + NS/2) * |lp| binary orders; keep that inside ~900 of a double's 1022.
    // CTCB200_FLAG_LATTICE_LOG (per call) or the lattice_log option forces the log-space recursion for every utterance.
    const int lattice_mode = opt(OPT_LATTICE_LOG) || (zero_infinity & CTCB200_FLAG_LATTICE_LOG);
    const int thr_env = opt(OPT_LIN_THR);
    const float lin_thr = lattice_mode ? 1.f
                                       : -(float)(thr_env > 0 ? thr_env : 900 / (lin_tile_frames(g.NS) + g.NS / 2));

    const bool fused = fg != nullptr;
    const int stages = fused ? fg->stages : 7;
    const bool want_argmax = (zero_infinity & CTCB200_FLAG_DECODE) != 0;
    zero_infinity &= 1;
    // experiment (off by default): write the zeros of the padded frames from extra CTAs of the lattice launch,
    // where the HBM is otherwise idle.  Measured on B200: the sweep gets 75 us shorter and the lattice 77 us
    // longer -- any co-resident memory traffic doubles the latency-bound lattice -- so nothing is gained.
    const bool zero_in_lattice = fused && stages == 7 && opt(OPT_ZERO_IN_LATTICE);
    cudaError_t e = cudaSuccess;
    const int P = group_frames(V, T);
    int *gstart = (int *)(ws + w.gstart);
    if (stages & (1 | 8)) {      // prep (stage bit 0 = prep + sweep; bit 3 = prep alone, bit 4 = sweep alone: profiling)
    prefer_max_carveout(k0_prep);
    k0_prep<<<1, 1024, 0, s>>>(in_len, tgt_len, targets_stride, B, T, Umax, hdr, Tb, Ub, flags, toff, rowstart, slow, bad,
                               P > 0 ? P : 1, gstart);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    }
    if (stages & (1 | 16)) {
    StreamCfg c;
    int nt1, rounds1;
    bool exact1;
    stream_pick(V, fused ? opt_or(OPT_K1F_NT, 128) : opt_or(OPT_K1_NT, 64), &nt1, &rounds1, &exact1);
    if (fused)
        rc = stream_cfg(V, 0, 96 + (size_t)g.Lp * 4, dev.sms, 4, 2, OPT_K1F_NST, OPT_K1F_CPS, &c);
    else
        rc = stream_cfg(V, 0, 96 + (size_t)g.Lp * 4, dev.sms, nt1 == 64 ? 2 : 3, nt1 == 64 ? 5 : 3,
                        OPT_K1_NST, OPT_K1_CPS, &c);
    if (rc) return rc;
    {
        const K1Args a = {logits, targets, tnumel, Tb, Ub, toff, rowstart, lp_lab, hdr, B, T, V, g.Lp, blank,
                          fused ? fg->grad : nullptr, fused ? fg->reduction : 0, fused ? fg->inv_batch : 0.f,
                          want_argmax ? (int *)(ws + w.best) : nullptr, zero_in_lattice ? 0 : 1, slow, lin_thr, bad};
        const K1dArgs da = {logits, targets, tnumel, Tb, Ub, toff, rowstart, gstart, lp_lab, hdr, B, T, V, g.Lp, blank, P,
                            a.grad, a.reduction, a.inv_batch, a.best, a.zero_pad_here, slow, lin_thr, bad};
        if (fused ? try_launch_k1d<true>(s, dev, da, &e) : try_launch_k1d<false>(s, dev, da, &e)) {
            // the direct sweep was launched (or failed to launch: e)
        } else if (fused ? try_launch_k1p<true>(s, dev, da, &e) : try_launch_k1p<false>(s, dev, da, &e)) {
            // the ring-fed aligned-group sweep was launched
        } else if (fused) {
            if (nt1 == 64) e = STREAM_DISPATCH(launch_k1f, 64, rounds1, exact1, c, s, a);
            else e = STREAM_DISPATCH(launch_k1f, 128, rounds1, exact1, c, s, a);
        } else {
            if (nt1 == 64) e = STREAM_DISPATCH(launch_k1, 64, rounds1, exact1, c, s, a);
            else e = STREAM_DISPATCH(launch_k1, 128, rounds1, exact1, c, s, a);
        }
    }
    if (e != cudaSuccess) return (int)e;
    if (sweep_done && (e = cudaEventRecord((cudaEvent_t)sweep_done, s)) != cudaSuccess) return (int)e;
    }   // stage: sweep

    if (opt(OPT_SKIP_LATTICE)) return CTCB200_OK;   // profiling aid: time the sweep alone
    if (stages & 2) {
    unsigned *ticket = (unsigned *)(hdr + 1);
#define K2_ARGS s, targets, tnumel, Tb, Ub, toff, flags, lp_lab, gam, ab, nll, loss_sums, ticket, B, T, zero_infinity, \
                (zero_in_lattice ? fg->grad : nullptr), rowstart, V, dev.sms * opt_or(OPT_ZERO_CPS, 2),   \
                (double *)(ws + w.tile_off), (fused ? fg->inv_batch : 1.f / (float)B), slow, w.ab_utt, ((stages & 1) && !sweep_done), bad
    if (want_grad) {
        if (g.NS == 4) e = launch_k2<4, true>(K2_ARGS);
        else if (g.NS == 8) e = launch_k2<8, true>(K2_ARGS);
        else e = launch_k2<16, true>(K2_ARGS);
    } else {
        if (g.NS == 4) e = launch_k2<4, false>(K2_ARGS);
        else if (g.NS == 8) e = launch_k2<8, false>(K2_ARGS);
        else e = launch_k2<16, false>(K2_ARGS);
    }
#undef K2_ARGS
    }   // stage: lattice
    if (e != cudaSuccess || !fused || !(stages & 4)) return (int)e;
    {
        const int per = opt_or(OPT_K3P_CPS, 32);
        // occupancies <= 2^-bits are not applied to the gradient: at the default 40 that is < 1e-12 of the utterance's
        // gradient scale (fp32 resolves 3e-11 at a softmax value of 1/V; the parity bar is 1e-4 absolute), and on
        // diffuse posteriors it is a third of all (frame, class) pairs, each a 32-byte DRAM read-modify-write.
        // CTCB200_OCC_SKIP_BITS=0 applies everything but exact zeros.
        const int skip_bits = opt(OPT_OCC_SKIP_BITS);
        const float occ_skip = skip_bits > 0 ? ldexpf(1.f, -skip_bits) : 0.f;
        const size_t smem = 2 * (size_t)g.Lp * 4;
        prefer_max_carveout(k3p_patch<64>);
        e = launch_pdl((stages & 2) ? 2 : 3, k3p_patch<64>, dim3(dev.sms * (per < 1 ? 1 : per)), dim3(64), smem, s, targets, tnumel, Tb, Ub, toff,
                       flags, rowstart, gam, fg->grad, fg->reduction, fg->inv_batch, B, T, V, g.Lp, blank, zero_infinity, occ_skip, bad);
    }
    return (int)e;
}


template <int NT, int MAXC, bool EXACT>
cudaError_t launch_kce(const StreamCfg &c, cudaStream_t s, bool want_grad, const float *pred, const int64_t *gold,
                       int rows, int V, const int *hdr, const int *vlist, const int *plist, float *rowloss,
                       float *grad, float eps, float weight) {
    cudaError_t e;
    if (want_grad) {
        e = cudaFuncSetAttribute(kce_rows<NT, MAXC, EXACT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
        if (e != cudaSuccess) return e;
        kce_rows<NT, MAXC, EXACT, true><<<c.grid, NT, c.smem, s>>>(pred, gold, rows, V, hdr, vlist, plist, rowloss, grad,
                                                                   eps, weight, c.nst, c.slot_bytes);
    } else {
        e = cudaFuncSetAttribute(kce_rows<NT, MAXC, EXACT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
        if (e != cudaSuccess) return e;
        kce_rows<NT, MAXC, EXACT, false><<<c.grid, NT, c.smem, s>>>(pred, gold, rows, V, hdr, vlist, plist, rowloss, grad,
                                                                    eps, weight, c.nst, c.slot_bytes);
    }
    return cudaGetLastError();
}

}  // namespace

namespace ctcb200 {

int internal_sm_count(int *sms) {
    DevInfo d;
    const int rc = device_info(&d);
    if (!rc) *sms = d.sms;
    return rc;
}

float internal_lin_thr(const Geom &g, int flags) {
    const int lattice_mode = opt(OPT_LATTICE_LOG) || (flags & CTCB200_FLAG_LATTICE_LOG);
    const int thr_env = opt(OPT_LIN_THR);
    return lattice_mode ? 1.f : -(float)(thr_env > 0 ? thr_env : 900 / (lin_tile_frames(g.NS) + g.NS / 2));
}

int internal_g3_opt(int which) { return opt((OptId)(OPT_G3_SWZ + which)); }

float internal_occ_skip() {
    const int skip_bits = opt(OPT_OCC_SKIP_BITS);
    return skip_bits > 0 ? ldexpf(1.f, -skip_bits) : 0.f;
}

int internal_prep(const int64_t *in_len, const int64_t *tgt_len, int64_t targets_stride, int B, int T, int Umax,
                  void *workspace, const Workspace &w, cudaStream_t s) {
    unsigned char *ws = (unsigned char *)workspace;
    prefer_max_carveout(k0_prep);
    k0_prep<<<1, 1024, 0, s>>>(in_len, tgt_len, targets_stride, B, T, Umax, (int *)(ws + w.hdr), (int *)(ws + w.Tb),
                               (int *)(ws + w.Ub), (int *)(ws + w.flags), (int64_t *)(ws + w.toff),
                               (int *)(ws + w.rowstart), (int *)(ws + w.slow), (int *)(ws + w.bad), 1,
                               (int *)(ws + w.gstart));
    return (int)cudaGetLastError();
}

int internal_lattice(bool want_grad, const int64_t *targets, int64_t tnumel, int B, int T, int V, int zero_infinity,
                     float *nll, float *loss_sums, float mean_scale, void *workspace, const Workspace &w, const Geom &g,
                     cudaStream_t s) {
    unsigned char *ws = (unsigned char *)workspace;
    int *hdr = (int *)(ws + w.hdr);
    cudaError_t e;
#define K2_ARGS s, targets, tnumel, (int *)(ws + w.Tb), (int *)(ws + w.Ub), (int64_t *)(ws + w.toff),                    \
                (int *)(ws + w.flags), (float *)(ws + w.lp_lab), (float *)(ws + w.gam), (float *)(ws + w.ab), nll,         \
                loss_sums, (unsigned *)(hdr + 1), B, T, zero_infinity & 1, (float *)nullptr, (int *)(ws + w.rowstart), V, \
                0, (double *)(ws + w.tile_off), mean_scale, (int *)(ws + w.slow), w.ab_utt, false, (int *)(ws + w.bad)
    if (want_grad) {
        if (g.NS == 4) e = launch_k2<4, true>(K2_ARGS);
        else if (g.NS == 8) e = launch_k2<8, true>(K2_ARGS);
        else e = launch_k2<16, true>(K2_ARGS);
    } else {
        if (g.NS == 4) e = launch_k2<4, false>(K2_ARGS);
        else if (g.NS == 8) e = launch_k2<8, false>(K2_ARGS);
        else e = launch_k2<16, false>(K2_ARGS);
    }
#undef K2_ARGS
    return (int)e;
}

}  // namespace ctcb200

extern "C" {

int ctcb200_version(void) { return CTCB200_VERSION; }

const char *ctcb200_strerror(int code) {
    switch (code) {
        case CTCB200_OK: return "ok";
        case CTCB200_ERR_NULL: return "required pointer is NULL";
        case CTCB200_ERR_SHAPE: return "bad shape (need B>=0, T>=1, 2<=V<=16384, strides>=0)";
        case CTCB200_ERR_BLANK: return "blank index outside [0,V)";
        case CTCB200_ERR_UMAX: return "Umax outside [0,255]";
        case CTCB200_ERR_ALIGN: return "logits/grad must be 16-byte aligned, workspace 256-byte aligned";
        case CTCB200_ERR_REDUCTION: return "unknown reduction code";
        case CTCB200_ERR_WORKSPACE: return "workspace too small (see ctcb200_workspace_bytes)";
        case CTCB200_ERR_NO_DEVICE: return "no CUDA device with compute capability 10.x (sm_100a) is current";
        case CTCB200_ERR_OPTION: return "unknown option name";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown ctcb200 error";
}

int ctcb200_set_option(const char *name, int value) {
    if (!name) return CTCB200_ERR_NULL;
    for (Opt &o : g_opt)
        if (!strcmp(o.name, name)) { o.value = value; return CTCB200_OK; }
    return CTCB200_ERR_OPTION;
}

int ctcb200_get_option(const char *name, int *value) {
    if (!name || !value) return CTCB200_ERR_NULL;
    for (const Opt &o : g_opt)
        if (!strcmp(o.name, name)) { *value = o.value; return CTCB200_OK; }
    return CTCB200_ERR_OPTION;
}

int ctcb200_workspace_bytes(int B, int T, int V, int Umax, size_t *out_bytes) {
    if (!out_bytes) return CTCB200_ERR_NULL;
    if (B < 0 || T < 1 || V < 2 || V > kMaxV) return CTCB200_ERR_SHAPE;
    Geom g;
    if (!geom_for(Umax, &g)) return CTCB200_ERR_UMAX;
    *out_bytes = workspace_layout(B, T, g).total;
    return CTCB200_OK;
}

int ctcb200_forward(const float *logits, const int64_t *targets, int64_t targets_stride, int64_t targets_numel,
                    const int64_t *in_len, const int64_t *tgt_len, int B, int T, int V, int Umax, int blank,
                    int zero_infinity, float *nll, float *loss_sums, void *workspace, size_t workspace_bytes,
                    ctcb200_stream_t stream, ctcb200_event_t sweep_done) {
    return forward_impl(true, nullptr, logits, targets, targets_stride, targets_numel, in_len, tgt_len, B, T, V, Umax,
                        blank, zero_infinity, nll, loss_sums, workspace, workspace_bytes, stream, sweep_done);
}

int ctcb200_loss_only(const float *logits, const int64_t *targets, int64_t targets_stride, int64_t targets_numel,
                      const int64_t *in_len, const int64_t *tgt_len, int B, int T, int V, int Umax, int blank,
                      int zero_infinity, float *nll, float *loss_sums, void *workspace, size_t workspace_bytes,
                      ctcb200_stream_t stream, ctcb200_event_t sweep_done) {
    return forward_impl(false, nullptr, logits, targets, targets_stride, targets_numel, in_len, tgt_len, B, T, V, Umax,
                        blank, zero_infinity, nll, loss_sums, workspace, workspace_bytes, stream, sweep_done);
}

int ctcb200_loss_grad(const float *logits, const int64_t *targets, int64_t targets_stride, int64_t targets_numel,
                      const int64_t *in_len, const int64_t *tgt_len, int B, int T, int V, int Umax, int blank,
                      int zero_infinity, int reduction, float inv_batch, float *nll, float *loss_sums,
                      float *grad_logits, void *workspace, size_t workspace_bytes, ctcb200_stream_t stream,
                      ctcb200_event_t sweep_done) {
    if (!grad_logits) return CTCB200_ERR_NULL;
    if ((uintptr_t)grad_logits & 15) return CTCB200_ERR_ALIGN;
    if (reduction < 0 || reduction > 2) return CTCB200_ERR_REDUCTION;
    const FusedGrad fg = {grad_logits, reduction, inv_batch, 7};
    return forward_impl(true, &fg, logits, targets, targets_stride, targets_numel, in_len, tgt_len, B, T, V, Umax,
                        blank, zero_infinity, nll, loss_sums, workspace, workspace_bytes, stream, sweep_done);
}

int ctcb200_loss_grad_stages(int stages, const float *logits, const int64_t *targets, int64_t targets_stride,
                             int64_t targets_numel, const int64_t *in_len, const int64_t *tgt_len, int B, int T, int V,
                             int Umax, int blank, int zero_infinity, int reduction, float inv_batch, float *nll,
                             float *loss_sums, float *grad_logits, void *workspace, size_t workspace_bytes,
                             ctcb200_stream_t stream) {
    if (!grad_logits) return CTCB200_ERR_NULL;
    if ((uintptr_t)grad_logits & 15) return CTCB200_ERR_ALIGN;
    if (reduction < 0 || reduction > 2) return CTCB200_ERR_REDUCTION;
    if (stages < 1 || stages > 31) return CTCB200_ERR_SHAPE;
    const FusedGrad fg = {grad_logits, reduction, inv_batch, stages};
    return forward_impl(true, &fg, logits, targets, targets_stride, targets_numel, in_len, tgt_len, B, T, V, Umax,
                        blank, zero_infinity, nll, loss_sums, workspace, workspace_bytes, stream, nullptr);
}

int ctcb200_backward(const float *logits, const int64_t *targets, int64_t targets_stride, int64_t targets_numel,
                     const float *grad_out, int64_t grad_out_stride, int reduction, float inv_batch, int B,
                     int T, int V, int Umax, int blank, int zero_infinity, float *grad_logits,
                     const void *workspace, size_t workspace_bytes, ctcb200_stream_t stream) {
    Geom g;
    Workspace w;
    // in_len/tgt_len are not re-read: the clamped copies of the forward live in the workspace
    int rc = check_common(logits, targets, workspace, workspace, B, T, V, Umax, blank, workspace, workspace_bytes,
                          &g, &w);
    if (rc) return rc;
    if (!grad_out || !grad_logits) return CTCB200_ERR_NULL;
    if ((uintptr_t)grad_logits & 15) return CTCB200_ERR_ALIGN;
    zero_infinity &= 1;
    if (reduction < 0 || reduction > 2) return CTCB200_ERR_REDUCTION;
    if (targets_stride < 0 || targets_numel < 0 || grad_out_stride < 0) return CTCB200_ERR_SHAPE;
    if (B == 0) return CTCB200_OK;
    DevInfo dev;
    if ((rc = device_info(&dev))) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned char *ws = (const unsigned char *)workspace;
    const int *Tb = (const int *)(ws + w.Tb), *Ub = (const int *)(ws + w.Ub), *flags = (const int *)(ws + w.flags);
    const int64_t *toff = (const int64_t *)(ws + w.toff);
    const int *rowstart = (const int *)(ws + w.rowstart);
    const float *gam = (const float *)(ws + w.gam);
    const int64_t tnumel = targets_stride ? (int64_t)B * targets_stride : targets_numel;

    StreamCfg c;
    const uint32_t gam_stage = (uint32_t)align_up((size_t)g.Lp * 4, 128);
    if ((rc = stream_cfg(V, gam_stage, 3 * (size_t)g.Lp * 4, dev.sms, 3, 3, OPT_K3_NST, OPT_K3_CPS, &c)))
        return rc;
    cudaError_t e;
    {
        const K3Args a = {logits, targets, tnumel, Tb, Ub, toff, flags, rowstart, gam, grad_out, grad_out_stride,
                          reduction, inv_batch, grad_logits, B, T, V, g.Lp, blank, zero_infinity,
                          (const int *)(ws + w.bad)};
        int nt, rounds;
        bool exact;
        stream_pick(V, opt_or(OPT_K3_NT, 128), &nt, &rounds, &exact);
        if (nt == 64) e = STREAM_DISPATCH(launch_k3, 64, rounds, exact, c, s, a);
        else e = STREAM_DISPATCH(launch_k3, 128, rounds, exact, c, s, a);
    }
    return (int)e;
}

int ctcb200_rescale_grad(float *grad_logits, const float *grad_out, int64_t grad_out_stride,
                         const float *applied_in, float *applied_out, int B, int T, int V,
                         ctcb200_stream_t stream) {
    if (!grad_logits || !grad_out || !applied_in || !applied_out) return CTCB200_ERR_NULL;
    if (B < 0 || T < 1 || V < 2 || grad_out_stride < 0) return CTCB200_ERR_SHAPE;
    if (B == 0) return CTCB200_OK;
    DevInfo dev;
    int rc = device_info(&dev);
    if (rc) return rc;
    int per = (2 * dev.sms + B - 1) / B;
    if (per < 1) per = 1;
    prefer_max_carveout(k4_rescale);
    k4_rescale<<<dim3(per, B), 256, 0, (cudaStream_t)stream>>>(grad_logits, grad_out, grad_out_stride, applied_in,
                                                                 applied_out, T, V);
    return (int)cudaGetLastError();
}

// ---- greedy CTC decode + edit distance --------------------------------------------
int ctcb200_greedy_decode(const int64_t *targets, int64_t targets_stride, int64_t targets_numel, int B, int T, int V,
                          int Umax, int blank, const void *workspace, size_t workspace_bytes, int *edit_out,
                          int *hyp_len_out, int64_t *hyp_out, ctcb200_stream_t stream) {
    Geom g;
    Workspace w;
    int rc = check_common(workspace, targets, workspace, workspace, B, T, V, Umax, blank, workspace, workspace_bytes,
                          &g, &w);
    if (rc == CTCB200_ERR_ALIGN && !((uintptr_t)workspace & 255)) rc = 0;
    if (rc) return rc;
    if (!edit_out || !hyp_len_out) return CTCB200_ERR_NULL;
    if (B == 0) return CTCB200_OK;
    const unsigned char *ws = (const unsigned char *)workspace;
    const int *Tb = (const int *)(ws + w.Tb), *Ub = (const int *)(ws + w.Ub), *best = (const int *)(ws + w.best);
    const int64_t *toff = (const int64_t *)(ws + w.toff);
    const int64_t tnumel = targets_stride ? (int64_t)B * targets_stride : targets_numel;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = 4 * (size_t)T * sizeof(int);
    if (smem > kSmemBudget) return CTCB200_ERR_SHAPE;
    const int grid = (B + 3) / 4;
    cudaError_t e;
#define K5(N)                                                                                                   \
    do {                                                                                                        \
        e = cudaFuncSetAttribute(k5_greedy_cer<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
        if (e == cudaSuccess) {                                                                                 \
            k5_greedy_cer<N><<<grid, 128, smem, s>>>(targets, tnumel, Tb, Ub, toff, best, edit_out, hyp_len_out, \
                                                     hyp_out, B, T, V, blank);                                  \
            e = cudaGetLastError();                                                                             \
        }                                                                                                       \
    } while (0)
    if (Umax <= 64) K5(2);
    else if (Umax <= 128) K5(4);
    else K5(8);
#undef K5
    return (int)e;
}

int ctcb200_edit_distance(const int64_t *hyp, int64_t hyp_stride, const int64_t *gold, int64_t gold_stride, int B,
                          int L, int pad, int mode, int *edit_out, int *words_out, ctcb200_stream_t stream) {
    if (!hyp || !gold || !edit_out || !words_out) return CTCB200_ERR_NULL;
    if (B < 0 || L < 1 || hyp_stride < L || gold_stride < L || mode < 0 || mode > 1) return CTCB200_ERR_SHAPE;
    const int nsym = mode ? 2 * L - 1 : L;
    if (nsym > 32 * 32) return CTCB200_ERR_UMAX;
    if (B == 0) return CTCB200_OK;
    const size_t smem = 4 * 2 * (size_t)L * sizeof(int);
    const int grid = (B + 3) / 4;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
#define K6(N)                                                                                                     \
    do {                                                                                                          \
        e = cudaFuncSetAttribute(k6_seq_edit<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
        if (e == cudaSuccess) {                                                                                   \
            k6_seq_edit<N><<<grid, 128, smem, s>>>(hyp, hyp_stride, gold, gold_stride, B, L, pad, mode, edit_out, \
                                                   words_out);                                                    \
            e = cudaGetLastError();                                                                               \
        }                                                                                                         \
    } while (0)
    if (nsym <= 128) K6(4);
    else if (nsym <= 256) K6(8);
    else if (nsym <= 512) K6(16);
    else K6(32);
#undef K6
    return (int)e;
}

// ---- attention-branch cross-entropy -------------------------------------------
static size_t ce_ws_layout(int64_t rows, size_t *o_vlist, size_t *o_plist, size_t *o_rowloss) {
    size_t o = kAlign;                      // hdr
    *o_vlist = o;   o += align_up(sizeof(int) * (size_t)(rows > 0 ? rows : 1));
    *o_plist = o;   o += align_up(sizeof(int) * (size_t)(rows > 0 ? rows : 1));
    *o_rowloss = o; o += align_up(sizeof(float) * (size_t)(rows > 0 ? rows : 1));
    return o;
}

int ctcb200_ce_workspace_bytes(int64_t rows, size_t *out_bytes) {
    if (!out_bytes) return CTCB200_ERR_NULL;
    if (rows < 0 || rows > 0x7fffffff) return CTCB200_ERR_SHAPE;
    size_t a, b, c;
    *out_bytes = ce_ws_layout(rows, &a, &b, &c);
    return CTCB200_OK;
}

int ctcb200_ce_loss_grad(const float *pred, const int64_t *gold, int64_t rows, int V, int ignore_index,
                         float smoothing, float weight, float *loss_out, float *grad, void *workspace,
                         size_t workspace_bytes, ctcb200_stream_t stream) {
    if (!pred || !gold || !loss_out || !workspace) return CTCB200_ERR_NULL;
    if (rows < 0 || rows > 0x7fffffff || V < 2 || V > kMaxV || smoothing < 0.f || smoothing >= 1.f) return CTCB200_ERR_SHAPE;
    if (((uintptr_t)pred & 15) || ((uintptr_t)grad & 15) || ((uintptr_t)workspace & 255)) return CTCB200_ERR_ALIGN;
    size_t o_v, o_p, o_r;
    if (workspace_bytes < ce_ws_layout(rows, &o_v, &o_p, &o_r)) return CTCB200_ERR_WORKSPACE;
    if (rows == 0) return CTCB200_OK;
    DevInfo dev;
    int rc = device_info(&dev);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned char *ws = (unsigned char *)workspace;
    int *hdr = (int *)ws, *vlist = (int *)(ws + o_v), *plist = (int *)(ws + o_p);
    float *rowloss = (float *)(ws + o_r);
    prefer_max_carveout(kce_prep);
    kce_prep<<<1, 1024, 0, s>>>(gold, (int)rows, ignore_index, hdr, vlist, plist);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    StreamCfg c;
    int nt, rounds;
    bool exact;
    stream_pick(V, 128, &nt, &rounds, &exact);
    if ((rc = stream_cfg(V, 0, 128, dev.sms, grad ? 4 : 3, grad ? 2 : 4, OPT_CE_NST, OPT_CE_CPS, &c))) return rc;
    e = STREAM_DISPATCH(launch_kce, 128, rounds, exact, c, s, grad != nullptr, pred, gold, (int)rows, V, hdr, vlist, plist,
                        rowloss, grad, smoothing, weight);
    if (e != cudaSuccess) return (int)e;
    prefer_max_carveout(kce_finish);
    kce_finish<<<1, 1024, 0, s>>>(hdr, vlist, rowloss, weight, loss_out);
    return (int)cudaGetLastError();
}

int ctcb200_read_status(const void *workspace, int *host_status, ctcb200_stream_t stream) {
    if (!workspace || !host_status) return CTCB200_ERR_NULL;
    cudaError_t e = cudaMemcpyAsync(host_status, workspace, sizeof(int), cudaMemcpyDeviceToHost,
                                    (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaStreamSynchronize((cudaStream_t)stream);
}

int ctcb200_read_lattice_stats(const void *workspace, int *host_stats, ctcb200_stream_t stream) {
    if (!workspace || !host_stats) return CTCB200_ERR_NULL;
    cudaError_t e = cudaMemcpyAsync(host_stats, (const int *)workspace + 2, 2 * sizeof(int), cudaMemcpyDeviceToHost,
                                    (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaStreamSynchronize((cudaStream_t)stream);
}

}  // extern "C"
