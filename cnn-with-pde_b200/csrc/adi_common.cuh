// Definitions shared by the two implementations of the implicit ADI layers:
//   adi.cu        a lane owns a whole line of one sample pair (plain Thomas recurrences);
//   adi_split.cu  a thread owns HALF a line (twisted factorisation: both halves eliminate
//                 towards the middle at once), P sample pairs share a block and its
//                 coefficient loads.
#pragma once
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace pde {
namespace adi {

constexpr int kHeaderBytes = 4096;
constexpr float kAmpLimit = 256.0f;  // see DESIGN.md "reverse reconstruction"

struct Header {
    int mode_exact;
    float amp_bound;
    int any_clamped;   // 1 if any cell of any sweep sits outside the clamp interval
    int pad;
    float scale[PDE_MAX_SWEEPS];
    float t[PDE_MAX_SWEEPS];
    unsigned rmax_bits[PDE_MAX_SWEEPS];
    int clamped[PDE_MAX_SWEEPS];   // 1 if any cell of sweep s sits outside the clamp interval
    // sweeps with the same axis, time, time step and spacing have the same tables (Strang: the
    // closing half sweep of a step and the opening one of the next): slot[s] numbers the distinct
    // ones in order of first appearance, rep[u] is the first sweep of slot u
    int nslots;
    short slot[PDE_MAX_SWEEPS];
    short rep[PDE_MAX_SWEEPS];
};
static_assert(sizeof(Header) <= kHeaderBytes, "header too large");

__host__ __device__ inline int sweeps_per_step(const pde_adi_desc &d) { return d.lie ? 2 : 3; }
// axis 0: lines along W (alpha); axis 1: lines along H (beta)
__host__ __device__ inline int sweep_axis(int k_in_step) { return k_in_step == 1 ? 1 : 0; }

__host__ __device__ inline size_t table_elems(const pde_adi_desc &d) {
    return (size_t)d.steps * sweeps_per_step(d) * d.C * d.N * d.N;
}

struct Tables {
    const Header *hdr;
    const float *r, *inv, *e, *msk;
};

__host__ __device__ inline Tables split_tables(const void *tables, const pde_adi_desc &d) {
    Tables t;
    const char *b = static_cast<const char *>(tables);
    t.hdr = reinterpret_cast<const Header *>(b);
    const float *f = reinterpret_cast<const float *>(b + kHeaderBytes);
    const size_t T = table_elems(d);
    t.r = f;
    t.inv = f + T;
    t.e = f + 2 * T;
    t.msk = f + 3 * T;
    return t;
}

// Software prefetch (no register destination).
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct Args {
    pde_adi_desc d;
    int S, sps, G, nitems, need_gin;
    int tile_sets, tmem_cols;
    const char *tables;
    const float *u, *gout, *chan, *skipw;
    float *out, *gin;
    float *scratch, *part_maps, *part_chan, *part_skip;
    // adi_split.cu only: split-layout tables, step checkpoints (written by the forward kernel,
    // read by the backward one) and the number of partial sets per channel
    const float *stab;
    float *ckpt;
};

inline int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    if (!v || !*v) return dflt;
    return atoi(v);
}

// adi.cu: sums the per-set gradient partials (double, fixed order).  Map partials: nsets_maps sets
// [set][kind][plane] with set % C == channel; channel / skip partials: nsets_small sets.
void launch_finish(const pde_adi_desc &d, int nsets_maps, int nsets_small, const float *part_maps,
                   const float *part_chan, const float *part_skip, const float *skipw, float *g_ab,
                   float *g_atc, float *g_bb, float *g_btc, float *g_chan, float *g_skip, cudaStream_t st);

// adi_split.cu
namespace split {
bool supported(const pde_adi_desc &d);
size_t table_floats(const pde_adi_desc &d);             // 4 tables in the split layout
size_t checkpoint_bytes(const pde_adi_desc &d);
size_t workspace_bytes(const pde_adi_desc &d);          // partials (+ exact-mode scratch), no checkpoints
int prepare(const pde_adi_desc &d, const pde_adi_schedule &sch, const float *ab, const float *bb,
            const float *atc, const float *btc, char *tables, cudaStream_t st);
int forward(const pde_adi_desc &d, const char *tables, const float *u, const float *chan, const float *skipw,
            float *out, float *ckpt, cudaStream_t st);
int backward(const pde_adi_desc &d, const char *tables, const float *u, const float *gout, const float *chan,
             const float *skipw, const float *ckpt, float *gin, float *g_ab, float *g_bb, float *g_atc,
             float *g_btc, float *g_chan, float *g_skip, void *workspace, size_t workspace_bytes,
             cudaStream_t st);
}  // namespace split

}  // namespace adi
}  // namespace pde
