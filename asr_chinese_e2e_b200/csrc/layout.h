// Geometry and HBM workspace layout shared by host launchers and kernels.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace ctcb200 {

// The lattice kernel gives each lane NS consecutive blank-extended states (NS/2 labels);
// one warp covers S = 2U+1 <= 32*NS states.  NS is picked from Umax.
struct Geom {
    int NS;  // states per lane: 4, 8 or 16
    int Lp;  // floats per (b,t) frame of lp_lab / gam: 4 header + 16*NS label slots
    int Sp;  // floats per (b,t) row of the stored alpha/beta halves: 32*NS
};

static inline bool geom_for(int Umax, Geom *g) {
    int NS;
    if (Umax < 0) return false;
    if (Umax <= 63) NS = 4;
    else if (Umax <= 127) NS = 8;
    else if (Umax <= 255) NS = 16;
    else return false;
    g->NS = NS;
    g->Lp = 4 + 16 * NS;
    g->Sp = 32 * NS;
    return true;
}

// Frame layout (floats):  [0] blank, [1] lse2 of the frame (log2 units), [2..3] unused, [4+j] label slot j
// (j < U_b; the finite log(0) sentinel -1e30 beyond).  A blank/label value is SELF-DESCRIBING: v > 0 is the
// probability 2^lp itself (what the linear-domain lattice multiplies by; written whenever lp >= the path's range
// limit, i.e. p >= 2^-90), v <= 0 (or NaN) is the log2-probability lp (out-of-range values, the sentinel).  The
// log-space lattice decodes v > 0 with one lg2.  In `gam` the same slots hold the posterior
// state occupancies: [0] sum over all blank states, [4+j] label state of slot j.

// Frames per stage of the linear-domain lattice (also its renormalisation interval).
static inline int lin_tile_frames(int NS) { return NS == 16 ? 4 : 8; }
// `ab` per utterance, linear-domain lattice: one block per stage = 32 lane exponents (int) followed by
// up to TT rows of up to Sp doubles.  The log-space lattice uses the same area as float[T][Sp].
static inline size_t lin_ab_utt_bytes(int T, const Geom &g) {
    const size_t tt = (size_t)lin_tile_frames(g.NS);
    const size_t lin = ((size_t)T + tt - 1) / tt * (128 + tt * (size_t)g.Sp * 8);
    const size_t lg = (size_t)T * g.Sp * 4;
    const size_t m = lin > lg ? lin : lg;
    return (m + 127) / 128 * 128;
}

constexpr size_t kAlign = 256;
static inline size_t align_up(size_t x, size_t a = kAlign) { return (x + a - 1) / a * a; }

struct Workspace {
    // byte offsets from the workspace base
    size_t hdr;       // int status; unsigned ticket; (256 B)
    size_t Tb;        // int[B]   clamped input lengths
    size_t Ub;        // int[B]   clamped target lengths
    size_t flags;     // int[B]   1 = infeasible (no valid alignment)
    size_t slow;      // int[B]   1 = some gathered log-probability is outside the linear-domain lattice's range
    size_t bad;       // int[B]   != 0: invalid input (length out of range, label out of range or == blank): the
                      //          utterance's nll and gradient are poisoned with NaN (F.ctc_loss raises on these)
    size_t toff;      // int64[B] element offset of utterance b's labels in `targets`
    size_t rowstart;  // int[B+1] exclusive prefix sum of Tb (valid-frame numbering)
    size_t gstart;    // int[B+1] exclusive prefix sum of ceil(Tb / P): numbering of the aligned frame groups (k1d_sweep)
    size_t lp_lab;    // float[B*T*Lp]
    size_t gam;       // float[B*T*Lp]
    size_t ab;        // stored alpha/beta halves: B blocks of ab_utt bytes (see lin_ab_utt_bytes)
    size_t ab_utt;    // bytes per utterance of `ab`
    size_t best;      // int[B*T]   per-frame argmax class (greedy CTC path), written by the sweep
    size_t tile_off;  // double[B*ceil(T/8)] running offset of each stored 8-frame stage of the lattice
    size_t total;
};

static inline Workspace workspace_layout(int B, int T, const Geom &g) {
    Workspace w;
    size_t o = 0;
    const size_t b = (size_t)(B > 0 ? B : 1);
    w.hdr = o;       o += kAlign;
    w.Tb = o;        o += align_up(sizeof(int) * b);
    w.Ub = o;        o += align_up(sizeof(int) * b);
    w.flags = o;     o += align_up(sizeof(int) * b);
    w.slow = o;      o += align_up(sizeof(int) * b);
    w.bad = o;       o += align_up(sizeof(int) * b);
    w.toff = o;      o += align_up(sizeof(int64_t) * b);
    w.rowstart = o;  o += align_up(sizeof(int) * (b + 1));
    w.gstart = o;    o += align_up(sizeof(int) * (b + 1));
    w.lp_lab = o;    o += align_up(sizeof(float) * b * T * g.Lp);
    w.gam = o;       o += align_up(sizeof(float) * b * T * g.Lp);
    w.ab_utt = lin_ab_utt_bytes(T, g);
    w.ab = o;        o += align_up(w.ab_utt * b);
    w.best = o;      o += align_up(sizeof(int) * b * T);
    w.tile_off = o;  o += align_up(sizeof(double) * b * ((size_t)(T + 7) / 8));
    w.total = o;
    return w;
}

}  // namespace ctcb200
