// classical.cu -- checkerboard Metropolis sweeps of R independent classical Ising replicas.
//
// Per-site rule = the reference's (delta_e of do_spin_flip classical/graph.rs:98-115, should_flip
// :339-347); the checkerboard schedule and the fixed-width draw are builder-defined (the reference
// only has a random-site schedule, graph.rs:350-406).  Draws are BIT-SLICED: the sites of a colour
// are grouped 32 consecutive ranks per group; bit-plane k (MSB first) of group g is output word (k & 3)
// of Philox4x32-10(key, ctr = (4 g + ((k & 15) >> 2), sweep_lo, sweep_hi, tag | c)), tag 'CB'/'CC' << 16
// for planes 0..15 / 16..31, and site j of the group draws d = sum_k bit_j(plane_k) << (31 - k).  It
// flips iff d < T, T = #{d : d * 2^-32 < exp(-beta * delta_e)} (2^32 when delta_e <= 0).  The
// thresholds are tabulated on the host with libm's exp so no device transcendental can leak in.
// A 32-site word is compared against a threshold with a bit-serial ripple over the planes (3 logic
// ops per plane for all 32 sites, ties included), and the low 16 planes are only generated when a
// site of a probabilistic class still ties after the high 16 (p = 2^-16 per site).
//
// Two layouts:
//  * generic graph: one byte per spin (the reference's Vec<bool>), CSR neighbours, one thread per
//    group of 32 ranks of a colour, threshold table per (replica, site class).
//  * L x L periodic square lattice, uniform J and bias: spins bit-packed in two colour planes,
//    one thread per 32 sites: neighbour counts by bit-sliced adders; two Philox calls (eight bit planes) decide nine
//    words in ten, the words that still hold a tie are parked in a per-warp queue and get their further planes 32 at a
//    time (k_cls_square_sweeps: both colours of many sweeps in one cooperative launch).
#include <algorithm>

#include "classical.cuh"

// ------------------------------------------------------------------------------------------
// generic graph: one thread per group of 32 ranks of the colour
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void planes16(uint32_t group, uint64_t sweep, uint32_t tag_c, uint32_t k0, uint32_t k1, uint32_t *pl) {
#pragma unroll
    for (uint32_t q = 0; q < 4; q++) {
        Philox4 o = philox4x32_10(4 * group + q, (uint32_t)sweep, (uint32_t)(sweep >> 32), tag_c, k0, k1);
        pl[4 * q] = o.x, pl[4 * q + 1] = o.y, pl[4 * q + 2] = o.z, pl[4 * q + 3] = o.w;
    }
}

__global__ void __launch_bounds__(128) k_cls_generic(ClsDev D, uint32_t colour, uint32_t cstart, uint32_t ccount,
                                                     uint64_t sweep) {
    const uint32_t r = blockIdx.y;
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;  // group of 32 ranks
    if (g * 32 >= ccount) return;
    const uint64_t key = D.key[r];
    uint32_t pl[32];
    planes16(g, sweep, QMCB_TAG_CB | colour, (uint32_t)key, (uint32_t)(key >> 32), pl);
    planes16(g, sweep, QMCB_TAG_CB2 | colour, (uint32_t)key, (uint32_t)(key >> 32), pl + 16);
    uint8_t *sp = D.spins + (size_t)r * D.N;
    const unsigned long long *thr = D.thr + (size_t)r * D.thr_stride;
    for (uint32_t j = 0; j < 32; j++) {
        const uint32_t rank = g * 32 + j;
        if (rank >= ccount) break;
        uint32_t d = 0;
#pragma unroll
        for (int k = 0; k < 32; k++) d |= ((pl[k] >> j) & 1u) << (31 - k);
        const uint32_t i = __ldg(D.colour_sites + cstart + rank);
        const uint32_t s = sp[i];
        const uint32_t a0 = __ldg(D.adj_start + i), a1 = __ldg(D.adj_start + i + 1);
        uint32_t mask = 0;
        for (uint32_t e = a0; e < a1; e++) mask |= (uint32_t)(sp[__ldg(D.adj_idx + e)] == s) << (e - a0);
        const uint32_t cls = __ldg(D.site_class + i);
        const unsigned long long T = thr[__ldg(D.class_off + cls) + ((s << (a1 - a0)) | mask)];
        if ((unsigned long long)d < T) sp[i] = (uint8_t)(s ^ 1u);
    }
}

// ------------------------------------------------------------------------------------------
// square lattice, bit-packed colour planes.
// plane[c][y][w] bit j  <->  site (x, y) with x = 2 * (32 w + j) + ((y + c) & 1); colour = (x+y)&1.
// The 32 sites of a word are exactly one group of 32 consecutive ranks (rank = y * L/2 + 32 w + j).
// ------------------------------------------------------------------------------------------
// Philox4x32-10 with the ten round keys precomputed (block-uniform, in shared memory)
__device__ __forceinline__ void philox_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t *rk, uint32_t *out) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ rk[2 * i], n2 = h0 ^ c3 ^ rk[2 * i + 1];
        c0 = n0, c1 = l1, c2 = n2, c3 = l0;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

#define SQ_MAXG 10  // threshold groups (distinct thresholds strictly between 0 and 2^32)

// One thread per 32 sites.  All prob. classes are compared in ONE bit-serial ripple: per plane the
// threshold bit of every site is muxed from the block-uniform masks of its class group (G logic ops),
// then lt |= eq & ~p & t;  eq &= ~(p ^ t)  (3 logic ops) decide all 32 sites, ties included.
template <int G>
struct SqTables {
    uint32_t rk[20];
    uint32_t tk[G > 0 ? G : 1][32];               // tk[g][k]: bit (31 - k) of the threshold of group g (0/1: a multiplier, FMA pipe)
    __align__(16) uint32_t cmsk[G + 1][12];       // [set][own * 5 + cnt] class membership, 0/1 (a multiplier); set G = always-flip
};
template <int G>
__device__ __forceinline__ void sq_load_tables(const ClsDev &D, uint32_t r, SqTables<G> &S) {
    const uint64_t key = D.key[r];
    if (threadIdx.x < 10) {
        S.rk[2 * threadIdx.x] = (uint32_t)key + 0x9E3779B9u * threadIdx.x;
        S.rk[2 * threadIdx.x + 1] = (uint32_t)(key >> 32) + 0xBB67AE85u * threadIdx.x;
    }
    for (uint32_t i = threadIdx.x; i < (uint32_t)G * 32; i += blockDim.x) {
        const uint32_t T = D.sq_gT[(size_t)r * SQ_MAXG + (i >> 5)];
        S.tk[i >> 5][i & 31] = (T >> (31 - (i & 31))) & 1u;
    }
    for (uint32_t i = threadIdx.x; i < (uint32_t)(G + 1) * 10; i += blockDim.x) {
        const uint32_t set = i / 10, j = i % 10;  // j = own * 5 + cnt  ->  class bit own * 8 + cnt
        const uint32_t members = set < (uint32_t)G ? D.sq_gmem[(size_t)r * SQ_MAXG + set] : D.sq_always[r];
        S.cmsk[set][j] = (members >> ((j / 5) * 8 + (j % 5))) & 1u;
    }
}
// classes of the 32 sites of a word from its four neighbour words: always-flip sites (returned), sites of every
// probabilistic group sel[g], eq = their union
template <int G>
__device__ __forceinline__ uint32_t sq_classes(const SqTables<G> &S, uint32_t own, uint32_t same, uint32_t side, uint32_t up, uint32_t dn,
                                               uint32_t (&sel)[G > 0 ? G : 1], uint32_t &eq) {
    // bit-sliced count of anti-aligned neighbours: cnt = lo + 2 mid + 4 hi, then one-hot masks
    const uint32_t a = own ^ same, b = own ^ side, c = own ^ up, d = own ^ dn;
    const uint32_t s1 = a ^ b ^ c, c1 = (a & b) | (c & (a ^ b));
    const uint32_t lo = s1 ^ d, c2 = s1 & d;
    const uint32_t mid = c1 ^ c2, hi = c1 & c2;
    uint32_t cm[5];
    cm[0] = ~(lo | mid | hi), cm[1] = lo & ~mid, cm[2] = mid & ~lo, cm[3] = lo & mid, cm[4] = hi;
    // sites of a set of classes: the membership of class (own, cnt) is a block-uniform 0/1 in shared memory; the one-hot
    // masks are disjoint, so "or of the selected masks" is a sum of products (integer multiply-adds: the FMA pipe is idle,
    // the ALU pipe that runs the logic ops is the binding unit of this kernel)
    auto class_sites = [&](const uint32_t *mk) -> uint32_t {
        uint32_t m0 = 0, m1 = 0;
#pragma unroll
        for (int cc = 0; cc < 5; cc++) m0 = cm[cc] * mk[cc] + m0, m1 = cm[cc] * mk[5 + cc] + m1;
        return (~own & m0) | (own & m1);
    };
    eq = 0;
#pragma unroll
    for (int g = 0; g < G; g++) sel[g] = class_sites(S.cmsk[g]), eq |= sel[g];
    return class_sites(S.cmsk[G]);  // delta_e <= 0 (threshold 2^32): always
}
// the update of one word (32 sites of one colour) at sweep `sweep`, in pieces.  CG: the planes are read through L2
// (another SM wrote the neighbouring rows earlier in the same launch).
// sq_front: own word, always-flip sites (returned), sites of every probabilistic group sel[g], eq = their union
template <int G, bool CG>
__device__ __forceinline__ uint32_t sq_front(const ClsDev &D, const SqTables<G> &S, const uint32_t *mine, const uint32_t *other, uint32_t t,
                                             uint32_t colour, uint32_t &own, uint32_t (&sel)[G > 0 ? G : 1], uint32_t &eq) {
    auto ld = [](const uint32_t *p) -> uint32_t { return CG ? __ldcg(p) : *p; };
    const uint32_t WPR = D.L >> 6;
    uint32_t y, w;
    if (D.sq_wpr_shift >= 0) y = t >> D.sq_wpr_shift, w = t & (WPR - 1u);
    else y = t / WPR, w = t - y * WPR;
    const uint32_t yu = y == 0 ? D.L - 1 : y - 1, yd = y + 1 == D.L ? 0 : y + 1;
    own = ld(mine + t);
    const uint32_t same = ld(other + t);
    const uint32_t up = ld(other + yu * WPR + w), dn = ld(other + yd * WPR + w);
    uint32_t side;
    if ((y + colour) & 1u) {  // my x is odd: the second horizontal neighbour has compressed index xc + 1
        const uint32_t nxt = ld(other + (w + 1 == WPR ? t + 1 - WPR : t + 1));
        side = (same >> 1) | (nxt << 31);
    } else {  // my x is even: neighbour xc - 1
        const uint32_t prv = ld(other + (w == 0 ? t + WPR - 1 : t - 1));
        side = (same << 1) | (prv >> 31);
    }
    return sq_classes<G>(S, own, same, side, up, dn, sel, eq);
}
// four planes K0..K0+3 of the bit-serial compare: per plane the threshold bit of every site is muxed from its group
// (disjoint sel: mux by multiply-add, FMA pipe), then lt |= eq & ~p & t; eq &= ~(p ^ t)
template <int G, int K0>
__device__ __forceinline__ void sq_ripple4(const SqTables<G> &S, const uint32_t (&sel)[G > 0 ? G : 1], const uint32_t *pl, uint32_t &lt, uint32_t &eq) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t tk = 0;
#pragma unroll
        for (int g = 0; g < G; g++) tk = sel[g] * S.tk[g][K0 + i] + tk;
        const uint32_t p = pl[i];
        lt |= eq & ~p & tk;
        eq &= ~(p ^ tk);
    }
}
// planes 12..15 and 16..31: only while some site still ties with its threshold (p = 2^-12 resp. 2^-16 per site)
template <int G>
__device__ __forceinline__ void sq_low_planes(const SqTables<G> &S, const uint32_t (&sel)[G > 0 ? G : 1], uint32_t t, uint32_t colour, uint64_t sweep,
                                              uint32_t &lt, uint32_t &eq) {
    uint32_t pl[4];
    philox_rk(4 * t + 3, (uint32_t)sweep, (uint32_t)(sweep >> 32), QMCB_TAG_CB | colour, S.rk, pl);
    sq_ripple4<G, 12>(S, sel, pl, lt, eq);
    if (!eq) return;
#pragma unroll 1
    for (int q = 0; q < 4; q++) {
        philox_rk(4 * t + q, (uint32_t)sweep, (uint32_t)(sweep >> 32), QMCB_TAG_CB2 | colour, S.rk, pl);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int k = 16 + 4 * q + i;
            uint32_t tk = 0;
#pragma unroll
            for (int g = 0; g < G; g++) tk = sel[g] * S.tk[g][k] + tk;
            const uint32_t p = pl[i];
            lt |= eq & ~p & tk;
            eq &= ~(p ^ tk);
        }
    }
}
// One thread per 32 sites; planes 0..11 always, so most words need three Philox calls.
template <int G, bool CG>
__device__ __forceinline__ void sq_update_word(const ClsDev &D, const SqTables<G> &S, uint32_t *mine, const uint32_t *other, uint32_t t,
                                               uint32_t colour, uint64_t sweep) {
    uint32_t own, eq, sel[G > 0 ? G : 1];
    uint32_t flip = sq_front<G, CG>(D, S, mine, other, t, colour, own, sel, eq);
    if (G > 0) {
        uint32_t lt = 0, pl[4];
        philox_rk(4 * t + 0, (uint32_t)sweep, (uint32_t)(sweep >> 32), QMCB_TAG_CB | colour, S.rk, pl);
        sq_ripple4<G, 0>(S, sel, pl, lt, eq);
        philox_rk(4 * t + 1, (uint32_t)sweep, (uint32_t)(sweep >> 32), QMCB_TAG_CB | colour, S.rk, pl);
        sq_ripple4<G, 4>(S, sel, pl, lt, eq);
        philox_rk(4 * t + 2, (uint32_t)sweep, (uint32_t)(sweep >> 32), QMCB_TAG_CB | colour, S.rk, pl);
        sq_ripple4<G, 8>(S, sel, pl, lt, eq);
        if (eq) sq_low_planes<G>(S, sel, t, colour, sweep, lt, eq);
        flip |= lt;
    }
    if (CG) __stcg(mine + t, own ^ flip);
    else mine[t] = own ^ flip;
}
// The same update with the third Philox call DEFERRED: after planes 0..7 a word still has a tie with probability
// ~ 32 * 2^-8, so almost every warp would execute the third call for a handful of its lanes.  Words with a tie are
// parked in a per-warp queue in shared memory instead ({t, eq, lt}: everything else is recomputed -- the other colour
// does not change during the pass and the word itself is only written when it is resolved) and resolved 32 at a time
// with all lanes busy.  Same planes, same result: only WHEN a plane is generated changes.
#define SQ_QCAP 64
struct SqQueue {
    uint32_t t[SQ_QCAP], eq[SQ_QCAP], lt[SQ_QCAP];
};
template <int G, bool CG>
__device__ __forceinline__ void sq_resolve(const ClsDev &D, const SqTables<G> &S, uint32_t *mine, const uint32_t *other, uint32_t t, uint32_t eq,
                                           uint32_t lt, uint32_t colour, uint64_t sweep) {
    uint32_t own, eq0, sel[G > 0 ? G : 1], pl[4];
    const uint32_t flip = sq_front<G, CG>(D, S, mine, other, t, colour, own, sel, eq0);
    philox_rk(4 * t + 2, (uint32_t)sweep, (uint32_t)(sweep >> 32), QMCB_TAG_CB | colour, S.rk, pl);
    sq_ripple4<G, 8>(S, sel, pl, lt, eq);
    if (eq) sq_low_planes<G>(S, sel, t, colour, sweep, lt, eq);
    if (CG) __stcg(mine + t, own ^ (flip | lt));
    else mine[t] = own ^ (flip | lt);
}
// one colour pass over the words [t0, t1) of a block, a warp taking 32 consecutive words per round
template <int G, bool CG>
__device__ __forceinline__ void sq_pass_deferred(const ClsDev &D, const SqTables<G> &S, SqQueue &Q, uint32_t *mine, const uint32_t *other, uint32_t t0,
                                                 uint32_t t1, uint32_t colour, uint64_t sweep) {
    auto ld = [](const uint32_t *p) -> uint32_t { return CG ? __ldcg(p) : *p; };
    const uint32_t lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    const uint32_t WPR = D.L >> 6, nwords = D.L * WPR;
    // when a round of the block covers an even number of whole rows, a thread stays in its column and on its row parity:
    // the five loads of a word sit at fixed distances from it (except in the first and the last row of the lattice)
    const bool fixed = D.sq_wpr_shift >= 0 && (blockDim.x >> D.sq_wpr_shift) >= 2 && ((blockDim.x >> D.sq_wpr_shift) << D.sq_wpr_shift) == blockDim.x &&
                       ((blockDim.x >> D.sq_wpr_shift) & 1u) == 0;
    int dside = 0;
    bool odd = false;  // my x is odd: the second horizontal neighbour has compressed index xc + 1, else xc - 1
    if (fixed) {
        const uint32_t tt = t0 + threadIdx.x, y = tt >> D.sq_wpr_shift, w = tt & (WPR - 1u);
        odd = (y + colour) & 1u;
        dside = odd ? (w + 1 == WPR ? 1 - (int)WPR : 1) : (w == 0 ? (int)WPR - 1 : -1);
    }
    uint32_t qn = 0;
    for (uint32_t base = t0 + (threadIdx.x & ~31u); base < t1; base += blockDim.x) {
        const uint32_t t = base + lane;
        uint32_t eq = 0, lt = 0;
        if (t < t1) {
            uint32_t own, flip, sel[G > 0 ? G : 1], pl[8];
            if (fixed) {
                const uint32_t *po = other + t;
                own = ld(mine + t);
                const uint32_t same = ld(po), sidew = ld(po + dside);
                const uint32_t up = ld(t < WPR ? po + (nwords - WPR) : po - WPR), dn = ld(t + WPR >= nwords ? po - (nwords - WPR) : po + WPR);
                const uint32_t side = odd ? (same >> 1) | (sidew << 31) : (same << 1) | (sidew >> 31);
                flip = sq_classes<G>(S, own, same, side, up, dn, sel, eq);
            } else {
                flip = sq_front<G, CG>(D, S, mine, other, t, colour, own, sel, eq);
            }
            philox_rk(4 * t + 0, (uint32_t)sweep, (uint32_t)(sweep >> 32), QMCB_TAG_CB | colour, S.rk, pl);
            philox_rk(4 * t + 1, (uint32_t)sweep, (uint32_t)(sweep >> 32), QMCB_TAG_CB | colour, S.rk, pl + 4);
            sq_ripple4<G, 0>(S, sel, pl, lt, eq);
            sq_ripple4<G, 4>(S, sel, pl + 4, lt, eq);
            if (!eq) {
                if (CG) __stcg(mine + t, own ^ (flip | lt));
                else mine[t] = own ^ (flip | lt);
            }
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, eq != 0);
        if (m) {
            if (eq) {
                const uint32_t pos = qn + (uint32_t)__popc(m & lt_mask);
                Q.t[pos] = t, Q.eq[pos] = eq, Q.lt[pos] = lt;
            }
            qn += (uint32_t)__popc(m);
            __syncwarp();
            if (qn >= 32) {
                qn -= 32;
                const uint32_t qt = Q.t[qn + lane], qe = Q.eq[qn + lane], ql = Q.lt[qn + lane];
                __syncwarp();
                sq_resolve<G, CG>(D, S, mine, other, qt, qe, ql, colour, sweep);
            }
        }
    }
    if (lane < qn) sq_resolve<G, CG>(D, S, mine, other, Q.t[lane], Q.eq[lane], Q.lt[lane], colour, sweep);
    __syncwarp();
}

// one colour pass of one sweep per launch
template <int G>
__global__ void __launch_bounds__(256, 4) k_cls_square(ClsDev D, uint32_t colour, uint64_t sweep) {
    __shared__ SqTables<G> S;
    const uint32_t r = blockIdx.y;
    sq_load_tables<G>(D, r, S);
    __syncthreads();
    const uint32_t words_per_plane = D.L * (D.L >> 6);
    uint32_t *mine = D.planes + ((size_t)r * 2 + colour) * words_per_plane;
    const uint32_t *other = D.planes + ((size_t)r * 2 + (colour ^ 1u)) * words_per_plane;
    // a block walks over several chunks of 256 words so that the set-up above is paid once
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < words_per_plane; t += gridDim.x * blockDim.x)
        sq_update_word<G, false>(D, S, mine, other, t, colour, sweep);
}

// FUSED: both colours of `nsweeps` sweeps in one launch.  SQ_BPR blocks share a replica (each owns a band of rows) and
// meet at a per-replica barrier in global memory between colour passes -- only the blocks of one replica exchange rows,
// so nothing waits for the whole grid.  All blocks of the launch are co-resident (cooperative launch).
#define SQ_BPR 4
template <int G>
__global__ void __launch_bounds__(128, 8) k_cls_square_sweeps(ClsDev D, uint32_t r0, uint64_t sweep0, uint32_t nsweeps, unsigned int *bar,
                                                               unsigned int bar_base) {
    __shared__ SqTables<G> S;
    __shared__ SqQueue Q[4];  // one per warp (128 threads)
    const uint32_t r = r0 + blockIdx.y;
    sq_load_tables<G>(D, r, S);
    __syncthreads();
    const uint32_t words_per_plane = D.L * (D.L >> 6);
    const uint32_t per = (words_per_plane + SQ_BPR - 1) / SQ_BPR;
    const uint32_t t0 = blockIdx.x * per, t1 = min(t0 + per, words_per_plane);
    uint32_t *plane0 = D.planes + (size_t)r * 2 * words_per_plane;
    unsigned int arrivals = bar_base;
    for (uint32_t s = 0; s < nsweeps; s++) {
#pragma unroll 1
        for (uint32_t colour = 0; colour < 2; colour++) {
            uint32_t *mine = plane0 + colour * words_per_plane;
            const uint32_t *other = plane0 + (colour ^ 1u) * words_per_plane;
            if (G > 0) sq_pass_deferred<G, true>(D, S, Q[threadIdx.x >> 5], mine, other, t0, t1, colour, sweep0 + s);
            else
                for (uint32_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) sq_update_word<G, true>(D, S, mine, other, t, colour, sweep0 + s);
            // per-replica barrier: my rows are visible before the neighbours' next pass reads them
            __syncthreads();
            arrivals += SQ_BPR;
            if (threadIdx.x == 0) {
                __threadfence();
                atomicAdd(bar + r, 1u);
                while ((int)(*(volatile unsigned int *)(bar + r) - arrivals) < 0) __nanosleep(32);
                __threadfence();
            }
            __syncthreads();
        }
    }
}

// energy (graph.rs:430-447 restricted to uniform J / bias: integer bond and spin counts) and
// magnetisation of the bit-packed layout: per replica  unsat = #anti-aligned bonds, up = #true spins
__global__ void __launch_bounds__(256) k_cls_square_measure(ClsDev D, unsigned long long *unsat, unsigned long long *up) {
    const uint32_t WPR = D.L >> 6;
    const uint32_t words_per_plane = D.L * WPR;
    const uint32_t r = blockIdx.y;
    unsigned long long u = 0, m = 0;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < words_per_plane; t += gridDim.x * blockDim.x) {
        const uint32_t y = t / WPR, w = t - y * WPR;
        const uint32_t *p0 = D.planes + ((size_t)r * 2) * words_per_plane;
        const uint32_t *p1 = p0 + words_per_plane;
        const uint32_t own = p0[t];  // colour-0 sites own all four of their bonds exactly once
        const uint32_t yu = y == 0 ? D.L - 1 : y - 1, yd = y + 1 == D.L ? 0 : y + 1;
        const uint32_t same = p1[y * WPR + w];
        uint32_t side;
        if (y & 1u) side = (same >> 1) | (p1[y * WPR + (w + 1 == WPR ? 0 : w + 1)] << 31);
        else side = (same << 1) | (p1[y * WPR + (w == 0 ? WPR - 1 : w - 1)] >> 31);
        u += __popc(own ^ same) + __popc(own ^ side) + __popc(own ^ p1[yu * WPR + w]) + __popc(own ^ p1[yd * WPR + w]);
        m += __popc(own) + __popc(p1[t]);
    }
    for (int o = 16; o; o >>= 1) u += __shfl_down_sync(0xFFFFFFFFu, u, o), m += __shfl_down_sync(0xFFFFFFFFu, m, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&unsat[r], u), atomicAdd(&up[r], m);
}

// bytes <-> bit planes
__global__ void k_cls_square_pack(ClsDev D, const uint8_t *bytes) {
    const uint32_t WPR = D.L >> 6, words_per_plane = D.L * WPR;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y, c = blockIdx.z;
    if (t >= words_per_plane) return;
    const uint32_t y = t / WPR, w = t - y * WPR;
    const uint8_t *row = bytes + (size_t)r * D.N + (size_t)y * D.L;
    uint32_t word = 0;
    for (int j = 0; j < 32; j++) word |= (uint32_t)(row[2 * (32 * w + j) + ((y + c) & 1u)] & 1u) << j;
    D.planes[((size_t)r * 2 + c) * words_per_plane + t] = word;
}
__global__ void k_cls_square_unpack(ClsDev D, uint8_t *bytes) {
    const uint32_t WPR = D.L >> 6, words_per_plane = D.L * WPR;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y, c = blockIdx.z;
    if (t >= words_per_plane) return;
    const uint32_t y = t / WPR, w = t - y * WPR;
    uint8_t *row = bytes + (size_t)r * D.N + (size_t)y * D.L;
    const uint32_t word = D.planes[((size_t)r * 2 + c) * words_per_plane + t];
    for (int j = 0; j < 32; j++) row[2 * (32 * w + j) + ((y + c) & 1u)] = (uint8_t)((word >> j) & 1u);
}

// stream-drawn initial spins (graph.rs:57, :451-453): spin i = top bit of stream word i
__global__ void k_cls_init_bytes(ClsDev D, uint8_t *bytes) {
    const uint32_t r = blockIdx.y;
    const uint64_t key = D.key[r];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < D.N; i += gridDim.x * blockDim.x)
        bytes[(size_t)r * D.N + i] = (uint8_t)(stream_word(key, i) >> 63);
}

// generic-layout energy: graph.rs:430-447 verbatim, one thread per replica chunk, f64 sum in site order
// is order dependent, so a single thread per replica walks the sites (not on the hot path).
__global__ void k_cls_generic_energy(ClsDev D, const double *adj_j, const double *biases, double *energy, double *mag) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= D.R) return;
    const uint8_t *sp = D.spins + (size_t)r * D.N;
    double acc = 0.0;
    long long m = 0;
    for (uint32_t i = 0; i < D.N; i++) {
        uint32_t si = sp[i];
        double total_e = 0.0;
        for (uint32_t e = D.adj_start[i]; e < D.adj_start[i + 1]; e++) {
            double oc = (si == sp[D.adj_idx[e]]) ? 1.0 : -1.0;
            total_e += adj_j[e] * oc / 2.0;
        }
        double bias_e = si ? -biases[i] : biases[i];
        acc = acc + total_e + bias_e;
        m += si ? 1 : -1;
    }
    energy[r] = acc;
    mag[r] = (double)m / (double)D.N;
}

void launch_cls_generic(const ClsDev &D, uint32_t colour, uint32_t cstart, uint32_t ccount, uint64_t sweep, cudaStream_t st) {
    uint32_t groups = (ccount + 31) / 32;
    dim3 grid((groups + 127) / 128, D.R);
    k_cls_generic<<<grid, 128, 0, st>>>(D, colour, cstart, ccount, sweep);
}
void launch_cls_square(const ClsDev &D, uint32_t colour, uint64_t sweep, cudaStream_t st) {
    uint32_t words = D.L * (D.L >> 6);
    dim3 grid((words + 1023) / 1024, D.R);  // four chunks of 256 words per block
    switch (D.sq_ngroups) {
        case 0: k_cls_square<0><<<grid, 256, 0, st>>>(D, colour, sweep); break;
        case 1: k_cls_square<1><<<grid, 256, 0, st>>>(D, colour, sweep); break;
        case 2: k_cls_square<2><<<grid, 256, 0, st>>>(D, colour, sweep); break;
        case 3: k_cls_square<3><<<grid, 256, 0, st>>>(D, colour, sweep); break;
        case 4: k_cls_square<4><<<grid, 256, 0, st>>>(D, colour, sweep); break;
        case 5: k_cls_square<5><<<grid, 256, 0, st>>>(D, colour, sweep); break;
        case 6: k_cls_square<6><<<grid, 256, 0, st>>>(D, colour, sweep); break;
        default: k_cls_square<SQ_MAXG><<<grid, 256, 0, st>>>(D, colour, sweep); break;
    }
}
// returns 0, or -1 if the fused kernel cannot hold a replica's blocks co-resident on this device
template <int G>
static int launch_sq_fused(const ClsDev &D, uint64_t sweep0, uint32_t nsweeps, unsigned int *bar, unsigned int *bar_count, int nsm, cudaStream_t st) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cls_square_sweeps<G>, 128, 0) != cudaSuccess || per_sm <= 0) return -1;
    const uint32_t slots = (uint32_t)per_sm * (uint32_t)nsm;
    const uint32_t rmax = slots / SQ_BPR;  // replicas one launch can hold co-resident
    if (rmax == 0) return -1;
    for (uint32_t r0 = 0; r0 < D.R; r0 += rmax) {
        const uint32_t nr = std::min(rmax, D.R - r0);
        ClsDev Dc = D;
        uint32_t r0v = r0, ns = nsweeps;
        uint64_t sw = sweep0;
        unsigned int base = *bar_count;
        void *args[] = {&Dc, &r0v, &sw, &ns, &bar, &base};
        if (cudaLaunchCooperativeKernel((const void *)k_cls_square_sweeps<G>, dim3(SQ_BPR, nr), dim3(128), args, 0, st) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
    }
    *bar_count += 2u * SQ_BPR * nsweeps;  // every replica's counter advances by the same amount (wraps modulo 2^32)
    return (int)((D.R + rmax - 1) / rmax);
}
int launch_cls_square_fused(const ClsDev &D, uint64_t sweep0, uint32_t nsweeps, unsigned int *bar, unsigned int *bar_count, int nsm, cudaStream_t st) {
    switch (D.sq_ngroups) {
        case 0: return launch_sq_fused<0>(D, sweep0, nsweeps, bar, bar_count, nsm, st);
        case 1: return launch_sq_fused<1>(D, sweep0, nsweeps, bar, bar_count, nsm, st);
        case 2: return launch_sq_fused<2>(D, sweep0, nsweeps, bar, bar_count, nsm, st);
        case 3: return launch_sq_fused<3>(D, sweep0, nsweeps, bar, bar_count, nsm, st);
        case 4: return launch_sq_fused<4>(D, sweep0, nsweeps, bar, bar_count, nsm, st);
        case 5: return launch_sq_fused<5>(D, sweep0, nsweeps, bar, bar_count, nsm, st);
        case 6: return launch_sq_fused<6>(D, sweep0, nsweeps, bar, bar_count, nsm, st);
        default: return launch_sq_fused<SQ_MAXG>(D, sweep0, nsweeps, bar, bar_count, nsm, st);
    }
}
void launch_cls_square_measure(const ClsDev &D, unsigned long long *unsat, unsigned long long *up, cudaStream_t st) {
    uint32_t words = D.L * (D.L >> 6);
    dim3 grid(min((words + 255) / 256, 64u), D.R);
    k_cls_square_measure<<<grid, 256, 0, st>>>(D, unsat, up);
}
void launch_cls_square_pack(const ClsDev &D, const uint8_t *bytes, cudaStream_t st) {
    uint32_t words = D.L * (D.L >> 6);
    dim3 grid((words + 255) / 256, D.R, 2);
    k_cls_square_pack<<<grid, 256, 0, st>>>(D, bytes);
}
void launch_cls_square_unpack(const ClsDev &D, uint8_t *bytes, cudaStream_t st) {
    uint32_t words = D.L * (D.L >> 6);
    dim3 grid((words + 255) / 256, D.R, 2);
    k_cls_square_unpack<<<grid, 256, 0, st>>>(D, bytes);
}
void launch_cls_init_bytes(const ClsDev &D, uint8_t *bytes, cudaStream_t st) {
    dim3 grid(min((D.N + 255) / 256, 1024u), D.R);
    k_cls_init_bytes<<<grid, 256, 0, st>>>(D, bytes);
}
void launch_cls_generic_energy(const ClsDev &D, const double *adj_j, const double *biases, double *energy, double *mag, cudaStream_t st) {
    k_cls_generic_energy<<<(D.R + 63) / 64, 64, 0, st>>>(D, adj_j, biases, energy, mag);
}
