// Internal (not exported) entry points shared between the translation units of libctcb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "layout.h"

namespace ctcb200 {

// k0_prep on `s`: clamped lengths, label offsets, valid-frame prefix sums, per-utterance flags (ctcb200.cu).
__attribute__((visibility("hidden"))) int internal_prep(const int64_t *in_len, const int64_t *tgt_len,
                                                        int64_t targets_stride, int B, int T, int Umax, void *workspace,
                                                        const Workspace &w, cudaStream_t s);
// k2_lattice<NS, want_grad> on `s` over the lp_lab frames already in the workspace (ctcb200.cu).
__attribute__((visibility("hidden"))) int internal_lattice(bool want_grad, const int64_t *targets, int64_t tnumel, int B,
                                                           int T, int V, int zero_infinity, float *nll, float *loss_sums,
                                                           float mean_scale, void *workspace, const Workspace &w,
                                                           const Geom &g, cudaStream_t s);
// range limit (log2 units, <= 0) below which a gathered log-probability sends its utterance to the log-space lattice
__attribute__((visibility("hidden"))) float internal_lin_thr(const Geom &g, int flags);
__attribute__((visibility("hidden"))) float internal_occ_skip();
__attribute__((visibility("hidden"))) int internal_sm_count(int *sms);
// developer overrides of k_gemm3's MN-major operand description: 0 swizzle enum, 1 LBO, 2 SBO, 3 layout type (0 = default), 4 RNA split, 5 k_head in-ring split
__attribute__((visibility("hidden"))) int internal_g3_opt(int which);

}  // namespace ctcb200
