// sse.cuh -- device-side view of one batched SSE handle (R replicas of one lattice) and the
// lattice helpers shared by all SSE kernels.  Layout is replica-major structure-of-arrays:
// every per-slot array is [R][cap] so that a warp working on one replica streams contiguous
// 128-byte lines of its operator string.
#pragma once
#include "common.cuh"

struct SseDev {
    // lattice (replicated tables, read-only)
    uint32_t N, E, Nb, Nw;  // variables, edges, bond-index count (qmc_ising.rs:664-670), state words
    int has_h;              // |h| > f64::EPSILON
    const uint32_t *va, *vb;
    const uint2 *vab;       // [E] both variables of an edge in one 8-byte word
    // packed edge table for the block-shared copy in shared memory (sse_fast.cu): v0 | v1 << 14 | coupling code << 28,
    // couplings from the 16-entry dictionary; NULL when N > 16384, more than 16 distinct couplings, or per-replica rows
    const uint32_t *epk;
    double jdict[16];
    const double *J;
    uint64_t zone;          // rand 0.8 gen_range(0..Nb) acceptance zone: (Nb << lzcnt(Nb)) - 1
    double gamma, h;
    // batch
    uint32_t R;
    uint64_t cap;  // slots allocated per replica
    uint32_t *ops;      // [R][cap] operator words
    uint32_t *state;    // [R][Nw]  spin configuration at p = 0, one bit per variable
    uint32_t *n;        // [R] number of non-identity ops
    uint32_t *M;        // [R] cutoff
    uint64_t *cursor;   // [R] position in the replica's word stream
    uint64_t *key;      // [R] Philox key
    double *beta;       // [R]
    uint64_t *done;     // [R] sweeps completed
    unsigned long long *sum_n;   // [R] sum of n at sampled sweeps (energy estimator)
    unsigned long long *vupd;    // [R] sum of n after every sweep (vertex updates)
    uint32_t *ncl;      // [R] clusters found by the last cluster step
    uint32_t *ends;     // [R][4] first_p, last_p, first_site_p, (unused)
    int *status;        // [1] DEV_ERR_* bits
    unsigned long long *dbg;  // [16] optional event counters (NULL = off)
    // per-variable scratch
    uint32_t *vfirst, *vlast;  // [R][N] first / last leg (p << 1 | rel) on each variable, or NONE32
    uint32_t *cur;             // [R][N] FAST: current segment id per variable
    // STRICT workspace (allocated on demand)
    uint32_t *rec;       // [R][strict_rec_stride] STRICT workspace: one 32-byte record per slot (op word, 4 links, 2 cluster ids), or
                         // the world-line layout: one 16-byte entry per leg + two sentinels per variable (sse_serial.cu)
    uint32_t *ent;       // [R][cap] world-line layout: entry of leg 0 of the op in each slot
    uint32_t *frontier;  // [R][2*cap+16]
    uint32_t *interior;  // [R][4*cap+16]
    uint32_t *bits;      // [R][cap/32+2] flip bit per cluster (STRICT) / per segment (FAST)
    uint32_t *frozen;    // [R][cap/32+2] cluster holds a longitudinal op
    // FAST workspace
    uint32_t *parent;    // [R][N+cap+1] union-find parents over segment ids
    uint32_t *sid;       // [R][cap] COUNTER mode: a member of the cluster on the input side of each op (sse_counter.cu)
    // generic interactions (Qmc, qmc_runner.rs:113-156): diagonal matrix elements from tables instead of (J, Gamma, h).
    // g_w2[4 b + (s0 | s1 << 1)] for the two-variable interaction b, g_gam[v] for the constant one-variable op of
    // variable v, stored at its bond index: g_gam[E + v].  NULL = the transverse-field Ising weights of qmc_ising.rs:863-888.
    const double *g_w2, *g_gam;
    // loop updates (directed_loop.rs:103-301) need every matrix element: g_full[16 b + (o0 o1 i0 i1)] for the two-variable
    // interaction b, g_full[16 E + 4 v + (o i)] for the one-variable interaction of variable v (Interaction::at indexing,
    // qmc_runner.rs:560-600, 650-664).  NULL = no loop updates.  no_site: the model has no one-variable interactions at all
    // (Nb = E, no cluster edges, hence no cluster step: qmc_runner.rs:278-281).  loop_path: the sweep takes the all-serial
    // step on the per-slot records (loop updates on, two-variable ops that can be off-diagonal, or no_site).
    const double *g_full;
    int loop_updates, no_site, loop_path;
    // heat-bath diagonal update (heatbath.rs:10-61 BondWeights); NULL = Metropolis rule
    const double *hb_cum, *hb_maxw;  // [Nb] cumulative / per-bond maximum diagonal weight
    double hb_total;
    // per-replica Hamiltonians (tempering between unequal Hamiltonians, tempering_traits.rs:122-154;
    // same edges, couplings of the same sign: qmc_ising.rs:563-590).  NULL = every replica uses J/gamma/h.
    const uint32_t *ham;                     // [R] table row of each replica
    const double *J_tab;                     // [H][E]
    const double *gam_tab, *h_tab;           // [H]
    const double *hb_cum_tab, *hb_maxw_tab;  // [H][Nb] (heat-bath on)
    const double *hb_total_tab;              // [H]
};

// workspace of the RVB update (sse_rvb.cu), allocated by qmcb_set_run_rvb / qmcb_single_rvb_sweep.  Per replica:
// u32: var_starts[N+1] var_lengths[N] zero_vars[N] constant_ps[cap] | boundary_flips map/keys(v)/keys(p) [3][cap] |
//      boundary_noflips map/keys [2][N] | subvars[N] var_to_subvar[N] | bonds, bonds_before, bonds_after map [3][E], keys [3][E] | fill[N] |
//      world lines: start[N] len[N] cap[N] cursor index[N] cursor position[N] positions[rvb_lines_total]
// f64: key weights of boundary_flips [cap], boundary_noflips [N], the three bond sets [3][E]
// u8 : var_pos_popped[cap] var_nopos_popped[N] cluster_state[N] substate[N] mark[N]
struct RvbDev {
    const uint32_t *vb_start, *vb_list;  // classical_bonds (make_classical_bonds, qmc_ising.rs:420-432): bonds of each variable, in bond order
    uint32_t *u32;
    double *f64;
    uint8_t *u8;
    size_t stride32, stride64, stride8;
    unsigned long long *succ, *count;  // [R] total_rvb_successes, rvb_clusters_counted (qmc_ising.rs:42-43)
};
// world lines (positions of the ops on each variable, with slack for rotations): at most 2 n * 1.25 + 32 N entries
__host__ __device__ __forceinline__ size_t rvb_lines_total(const SseDev &D) { return (size_t)D.cap * 5 / 2 + 32 * (size_t)D.N + 64; }
__host__ __device__ __forceinline__ size_t rvb_stride32(const SseDev &D) {
    return 13 * (size_t)D.N + 1 + 4 * (size_t)D.cap + 6 * (size_t)D.E + rvb_lines_total(D);
}
__host__ __device__ __forceinline__ size_t rvb_stride64(const SseDev &D) { return (size_t)D.cap + D.N + 3 * (size_t)D.E; }
__host__ __device__ __forceinline__ size_t rvb_stride8(const SseDev &D) { return (size_t)D.cap + 4 * (size_t)D.N; }

// u32 words of one replica's STRICT workspace row: 8 per slot (two legs per slot at most) + the sentinels of every variable
__host__ __device__ __forceinline__ size_t strict_rec_stride(const SseDev &D) { return 8 * ((size_t)D.cap + D.N + 1); }

// kernel-selection knobs of the warp-parallel sweep, owned by the handle (qmcb_set_option); nsm is the SM count of the
// handle's device, queried once at creation
struct SseTuning {
    int minblocks = 0;   // resident blocks per SM the kernel is compiled for (register cap); 0 = choose by occupancy
    int pipe = 1;        // 0: never use two warps per replica
    int epk = 1;         // 0: never use the packed edge table
    int pad = 0;         // experiments: extra dynamic shared memory per block
    int carveout = -1;   // experiments: shared-memory carve-out preference (-1 = driver default)
    int nsm = 148;
};

// the Hamiltonian one replica is updated with
struct Ham {
    const double *J;
    const double *gw2, *ggam;  // generic interaction tables (Interaction::at for in == out), or NULL
    double gamma, h;
    const double *hb_cum, *hb_maxw;
    double hb_total;
};
template <bool MH>
__device__ __forceinline__ Ham ham_view(const SseDev &D, uint32_t r) {
    Ham m;
    m.gw2 = D.g_w2, m.ggam = D.g_gam;
    if (MH && D.ham) {
        const uint32_t hi = D.ham[r];
        m.J = D.J_tab + (size_t)hi * D.E, m.gamma = D.gam_tab[hi], m.h = D.h_tab[hi];
        m.hb_cum = D.hb_cum ? D.hb_cum_tab + (size_t)hi * D.Nb : nullptr;
        m.hb_maxw = D.hb_cum ? D.hb_maxw_tab + (size_t)hi * D.Nb : nullptr;
        m.hb_total = D.hb_cum ? D.hb_total_tab[hi] : 0.0;
    } else {
        m.J = D.J, m.gamma = D.gamma, m.h = D.h, m.hb_cum = D.hb_cum, m.hb_maxw = D.hb_maxw, m.hb_total = D.hb_total;
    }
    return m;
}

enum { KIND_BOND = 0, KIND_SITE = 1, KIND_LONG = 2 };

// bonds_fn of qmc_ising.rs:671-681
__device__ __forceinline__ int bond_kind(const SseDev &D, uint32_t b) {
    return b < D.E ? KIND_BOND : (b < D.E + D.N ? KIND_SITE : KIND_LONG);
}
__device__ __forceinline__ void bond_vars(const SseDev &D, uint32_t b, int kind, uint32_t &v0, uint32_t &v1) {
    if (kind == KIND_BOND) {
#ifdef QMCB_SPLIT_EDGES
        v0 = __ldg(D.va + b), v1 = __ldg(D.vb + b);
#else
        const uint2 e = __ldg(D.vab + b);
        v0 = e.x, v1 = e.y;
#endif
    } else {
        v0 = b - D.E - (kind == KIND_LONG ? D.N : 0u), v1 = v0;
    }
}
// diagonal matrix element <s|H_b|s>: qmc_ising.rs:863-888 with inputs == outputs
__device__ __forceinline__ double bond_weight(const Ham &Hm, uint32_t b, int kind, uint32_t s0, uint32_t s1) {
    if (Hm.gw2) return kind == KIND_BOND ? __ldg(Hm.gw2 + 4 * (size_t)b + (s0 | (s1 << 1))) : __ldg(Hm.ggam + b);
    if (kind == KIND_BOND) {
        double j = __ldg(Hm.J + b);
        return fabs(j) + (s0 == s1 ? -j : j);
    }
    if (kind == KIND_SITE) return Hm.gamma;
    return fabs(Hm.h) + (s0 ? Hm.h : -Hm.h);
}
// BondWeights::index_for_cumulative (heatbath.rs:56-60): slice::binary_search_by, insertion point
// when no element compares equal
__device__ __forceinline__ uint32_t hb_index_for_cumulative(const double *cum, uint32_t len, double val) {
    uint32_t size = len, left = 0, right = len;
    while (left < right) {
        const uint32_t mid = left + size / 2;
        const double c = __ldg(cum + mid);
        if (c < val) left = mid + 1;
        else if (c > val) right = mid;
        else return mid;
        size = right - left;
    }
    return left;
}
// the 52-bit fraction of rand 0.8's UniformFloat<f64>::sample_single: value1_2 - 1.0
__device__ __forceinline__ double unit_f64(uint64_t word) {
    return __longlong_as_double((long long)((word >> 12) | 0x3FF0000000000000ull)) - 1.0;
}
__device__ __forceinline__ uint32_t state_bit(const uint32_t *st, uint32_t v) { return (st[v >> 5] >> (v & 31)) & 1u; }
