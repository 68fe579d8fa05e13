// sse_counter.cu -- warp-parallel SSE sweep, COUNTER mode (QMCB_MODE_COUNTER): the FAST cluster order with a diagonal
// update whose uniform words are a function of (key, step nonce, SLOT) instead of positions in the sequential stream
// (contract: oracle.c diagonal_update_counter / DESIGN.md 3.8).  The acceptance arithmetic is the reference's
// (diagonal.rs:142-191, same numerator / denominator, same live n); what changes is where the words come from, and with
// it the only thing that made the pass sequential apart from n: a slot no longer has to know how many words the slots
// before it consumed.  One warp per replica, 32 slots per step:
//  P1  every lane draws its slot's Philox block, proposes / weighs its op and classifies its decision against
//      conservative thresholds valid for every n the step can reach at that lane; the (rare) lane whose word falls
//      between them is settled in lane order with the exact division at the exact n.  Result: bit-identical to the
//      sequential loop.  Fused in the same pass, as in sse_fast.cu: world-line segments, their union-find, and -- new --
//      a per-slot record `sid` of a member of the cluster on the input side of every op.
//  P2  one flip bit per segment id (parent id < child id), four words per round.
//  P3  stateless apply: op word + sid -> flip bits looked up per slot; no per-variable state, no lattice tables, no
//      match; four lines in flight per iteration.
// Batches that leave most of an SM empty run with two or three warps per replica (template parameter PIPE below).
#include <algorithm>
#include <map>
#include <mutex>

#include "sse.cuh"
#include "sse_warp.cuh"

#define T_NONE 0
#define T_EMPTY 1
#define T_DIAG 2
#define T_OFFD 3

// shared memory of one replica; fixed-size tables first (compile-time offsets), lattice-sized ones after
#define CT_NUM 0      // [36] f64: beta * Nb * <s|H_b|s> per weight class (2 * coupling code + (s0 != s1); 32 site; 33 + s long)
#define CT_RNUM 288   // [36] f64: 1 / num
#define CT_FL 576     // [32] u32: variable flipped by the off-diagonal op of lane j
// staging of the operator string: QMCB_BULK_LINES = 0: one 128-byte line ahead, copied by cp.async (8 lanes x 16 B);
// L > 0: tiles of L lines copied by ONE cp.async.bulk (TMA) into a two-stage ring, completion on an mbarrier
#ifndef QMCB_BULK_LINES
#define QMCB_BULK_LINES 0
#endif
#define CT_LINE 704   // [32] u32 next line (cp.async)  |  2 mbarriers + ring [2][L][32] u32 (bulk)
#if QMCB_BULK_LINES
#define CT_VAR (704 + 16 + 2 * QMCB_BULK_LINES * 128)
#else
#define CT_VAR 832    // st, tb, cd, sb [Nw each], rep [N]
#endif
#define CLS_SITE 32u
#define CLS_LONG 33u

// (+ 72 words when two warps share a replica: [2][32] final op words of a step, 8 words exchanged between the roles)
// (three roles: + [2][3][32] words of the unions role B hands to role C)
__host__ __device__ inline size_t cnt_smem_bytes(uint32_t N, uint32_t Nw, int pipe = 0) { return (CT_VAR + ((size_t)4 * Nw + N + (pipe ? 72u : 0u) + (pipe == 3 ? 192u : 0u)) * 4 + 15) / 16 * 16; }

#ifndef QMCB_WPB
#define QMCB_WPB 4
#endif

// PIPE: two warps per replica, for batches that leave most of an SM empty (tempering ladders, big lattices: the sweep of
// a long string is one dependent chain and nothing hides it there).  Role A runs the diagonal update of step k while role B
// does the segment bookkeeping and the unions of step k - 1; the final op words of a step are handed over through a two-slot
// ring in shared memory and the two warps meet at one named barrier per step.  B then does the closure and P2; P3 is split:
// A applies the first half of the string, B the second (it knows how many site ops precede it from its own count).
// PIPE = 3: a third warp (role C) takes the union-find of step k - 2 off role B: the climbs and the CAS are dependent L2
// round trips, and with two roles B was the longer half by a factor of two (role A spent half its time at the barrier).
// B hands C one (member of set a, member of set b, variables + flags) triple per lane; C hooks the roots and caches them in
// the per-variable table with the same CAS as before, so a stale entry is never overwritten.  P3 is split in thirds.
template <bool HAS_H, int MINB, bool MH, int PK, int PIPE>
__global__ void __launch_bounds__(PIPE ? 32 * PIPE * QMCB_WPB : 32 * QMCB_WPB, PIPE ? 2 : 4 * MINB / QMCB_WPB)
    k_sse_counter(SseDev D, uint64_t target, uint32_t phases, uint64_t sample_freq, uint64_t sample_origin, uint8_t *samples,
                  uint64_t samples_per_rep, uint32_t smem_stride, uint32_t epk_off) {
    extern __shared__ __align__(16) unsigned char smem_all[];
    if (PK == 1) {  // block-shared copy of the packed edge table (low-occupancy build)
        uint32_t *dst = (uint32_t *)(smem_all + epk_off);
        for (uint32_t i = threadIdx.x; i < D.E; i += blockDim.x) dst[i] = __ldg(D.epk + i);
        __syncthreads();
    }
    const uint32_t *const epk_s = PK == 2 ? D.epk : (const uint32_t *)(smem_all + epk_off);
    const uint32_t wraw = __reduce_max_sync(FULL, threadIdx.x >> 5);  // warp-uniform by construction (uniform registers)
    const uint32_t wib = PIPE ? wraw / (uint32_t)(PIPE ? PIPE : 1) : wraw;
    const uint32_t role = PIPE ? wraw - wib * (uint32_t)PIPE : 0u;
    const bool roleA = !PIPE || role == 0, roleB = !PIPE || role == 1, roleC = PIPE == 3 && role == 2;
#define PAIR_SYNC()                                                            \
    do {                                                                       \
        if (PIPE) asm volatile("bar.sync %0, %1;" ::"r"(wib + 1u), "n"(PIPE ? 32 * PIPE : 32) : "memory"); \
    } while (0)
    unsigned char *const smem_raw = smem_all + wib * smem_stride;
    const int lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * QMCB_WPB + wib;
    if (r >= D.R) return;
    const uint32_t N = D.N, Nw = D.Nw;
    double *const numtab = (double *)(smem_raw + CT_NUM), *const rnumtab = (double *)(smem_raw + CT_RNUM);
    uint32_t *const s_fl = (uint32_t *)(smem_raw + CT_FL), *const s_line = (uint32_t *)(smem_raw + CT_LINE);
    uint32_t *const s_st = (uint32_t *)(smem_raw + CT_VAR), *const s_tb = s_st + Nw, *const s_cd = s_st + 2 * Nw, *const s_sb = s_st + 3 * Nw;
    uint32_t *const s_rep = s_st + 4 * Nw;
    uint32_t *const hand = s_rep + N;                                   // PIPE: [2][32] final op words of a step
    volatile uint32_t *const xchg = (volatile uint32_t *)(hand + 64);  // PIPE: n, cluster count, site ops before the later parts of P3
    uint32_t *const un = hand + 72;                                     // PIPE 3: [2][3][32] unions of a step, role B -> role C
    uint32_t *ops = D.ops + (size_t)r * D.cap;
    uint32_t *sid = D.sid + (size_t)r * D.cap;
    uint32_t *gstate = D.state + (size_t)r * Nw;
    uint32_t *P = D.parent + (size_t)r * (N + D.cap + 1);
    const size_t bstride = (size_t)(D.cap / 32 + 2 + N / 32);
    uint32_t *decb = D.bits + (size_t)r * bstride;
    uint32_t *frz = D.frozen + (size_t)r * bstride;
    uint32_t lt_mask;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
    const uint64_t key = D.key[r];
    const Ham Hm = ham_view<MH>(D, r);
    const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    const uint32_t E = D.E, Nb = D.Nb;
    auto evars = [&](uint32_t b, int kind, uint32_t &v0, uint32_t &v1) -> uint32_t {
        if (PK && kind == KIND_BOND) {
            const uint32_t e = PK == 2 ? __ldg(epk_s + b) : epk_s[b];
            v0 = e & 0x3FFFu, v1 = (e >> 14) & 0x3FFFu;
            return e;
        }
        bond_vars(D, b, kind, v0, v1);
        return 0u;
    };
    // weight class of (packed table word, kind, spins) and the numerator of the acceptance ratio
    auto wclass = [&](uint32_t e, int kind, uint32_t s0, uint32_t s1) -> uint32_t {
        return kind == KIND_BOND ? 2u * (e >> 28) + (s0 ^ s1) : ((!HAS_H || kind == KIND_SITE) ? CLS_SITE : CLS_LONG + s0);
    };
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    const double G_LO = 1.0 - 9.094947017729282e-13, G_HI = 1.0 + 9.094947017729282e-13;  // 1 -+ 2^-40: guard of the reciprocal bounds

    uint32_t phase_bits = 0;  // bulk staging: parity of the next completion of each stage's mbarrier
    uint64_t done = D.done[r];
    const uint64_t nsteps = (phases & 16u) ? (done + 1 == target ? 1 : 0) : ((phases & 8u) ? (target > done ? target - done : 0) : 1);
    int err = 0;

#ifdef QMCB_PHASE_TIMERS
    // cycles of lane 0 of every role per phase, summed over replicas and sweeps: dbg[32 + 8 * role + k], k = 0 set-up,
    // 1 P1 loop, 2 closure (+ frozen marks), 3 P2, 4 P3, 5 sweep tail
    long long ct_ = clock64();
#define CTICK(k)                                                                                  \
    do {                                                                                          \
        if (lane == 0 && D.dbg) {                                                                 \
            const long long t_ = clock64();                                                       \
            atomicAdd(&D.dbg[32 + 8 * (PIPE ? role : 0u) + (k)], (unsigned long long)(t_ - ct_)); \
            ct_ = t_;                                                                             \
        }                                                                                         \
    } while (0)
#else
#define CTICK(k) ((void)0)
#endif
    for (uint64_t sw = 0; sw < nsteps; sw++) {
        const uint32_t M = D.M[r];
        if (M > D.cap) {
            if (lane == 0) atomicOr(D.status, DEV_ERR_CAPACITY);
            break;
        }
        uint32_t n = D.n[r];
        uint64_t cur = D.cursor[r];
        const double bn = D.beta[r] * (double)Nb;  // diagonal.rs:168, left-to-right product
        const bool do_diag = phases & 1u, do_clus = phases & 2u;
        if (roleA)
            for (uint32_t j = lane; j < Nw; j += 32) s_st[j] = gstate[j], s_cd[j] = 0;
        if (roleB)
            for (uint32_t j = lane; j < Nw; j += 32) s_tb[j] = 0, s_sb[j] = 0;
        if (PK && roleA) {
            for (uint32_t c = lane; c < 36; c += 32) {
                double wgt = 0.0;
                if (c < 32) {
                    const double j = D.jdict[c >> 1];
                    wgt = fabs(j) + ((c & 1u) ? j : -j);
                } else if (c == CLS_SITE) wgt = Hm.gamma;
                else if (HAS_H && c < 35) wgt = fabs(Hm.h) + (c == CLS_LONG + 1 ? Hm.h : -Hm.h);
                const double num = bn * wgt;
                numtab[c] = num, rnumtab[c] = 1.0 / num;
            }
        }
        if (do_clus && roleB) {
            for (uint32_t v = lane; v < N; v += 32) s_rep[v] = v, st_cg(P + v, v);
            if (HAS_H)
                for (uint32_t j = lane; j < (uint32_t)bstride; j += 32) st_cg(frz + j, 0u);
        }
        __syncwarp();
        PAIR_SYNC();
        uint32_t nsite = 0, ks_half = 0, ks_third = 0;
        bool anylong = false, alltb = false;
        const uint64_t cdiag = cur;  // nonce of this diagonal step
        if (do_diag) cur += 1;

        CTICK(0);
        // =========================== P1: diagonal update + segments + unions ===========================
        const uint32_t nit = (M + 31) / 32;
#if QMCB_BULK_LINES
        constexpr uint32_t LPS = QMCB_BULK_LINES;
        uint64_t *const bars = (uint64_t *)(smem_raw + CT_LINE);
        uint32_t *const ring = (uint32_t *)(smem_raw + CT_LINE + 16);
        const uint32_t ntile = (nit + LPS - 1) / LPS;
        auto issue_tile = [&](uint32_t t) {  // lane 0: lines [t * LPS, (t + 1) * LPS) of the string into stage t & 1
            const uint32_t first = t * LPS, cnt = min(LPS, nit - first);
            bulk_fetch(ring + (t & 1u) * LPS * 32, ops + (size_t)first * 32, cnt * 128u, bars + (t & 1u), pol_stream);
        };
        if (lane == 0) {
            if (sw == 0) mbar_init(bars, 1), mbar_init(bars + 1, 1), mbar_fence_init();
            if (ntile > 0) issue_tile(0);
            if (ntile > 1) issue_tile(1);
        }
        __syncwarp();
#else
        if (M && roleA) fetch_line_pol(s_line, ops, lane, pol_stream);
#endif
        // the slot's words: one Philox block per slot.  Independent of the operator string, so the block of step k + 1 is
        // computed at the end of step k, between issuing the step's union-find CAS and looking at its result.
        uint64_t wA = 0, wB = 0;
        uint32_t pb = 0;  // proposal of an empty slot: bond by multiply-shift
        auto draw = [&](uint32_t slot) {
            const Philox4 o = philox4x32_10(slot, (uint32_t)cdiag, (uint32_t)(cdiag >> 32), QMCB_TAG_DIAG, k0, k1);
            wA = ((uint64_t)o.y << 32) | o.x, wB = ((uint64_t)o.w << 32) | o.z;
            pb = (uint32_t)__umul64hi(wA, (uint64_t)Nb);
        };
        if (do_diag && roleA) draw((uint32_t)lane);
        // PIPE: P3 is split at these steps (multiples of its four-line iterations): two roles [0, half) [half, end), three roles
        // [0, half) [half, third) [third, end)
        const uint32_t half_it = PIPE == 3 ? ((nit / 3 + 3) / 4) * 4 : ((nit / 2 + 3) / 4) * 4;
        const uint32_t third_it = PIPE == 3 ? min(nit, ((2 * nit / 3 + 3) / 4) * 4) : nit;
        for (uint32_t it = 0; it < nit + (PIPE ? (uint32_t)PIPE - 1u : 0u); it++) {
            const uint32_t base = it * 32, p = base + lane;
            uint32_t w = OP_EMPTY, neww = OP_EMPTY, v0 = 0, v1 = 0, fmask = 0, flipv = NONE32;  // flipv: the variable my off-diagonal op flips
            int kind = -1;
            bool changed = false;
            if (roleA && it < nit) {
            const bool valid = p < M;
#if QMCB_BULK_LINES
            const uint32_t tile = it / LPS, tl = it % LPS;
            if (tl == 0) {  // the k-th completion of a stage's mbarrier has parity k & 1; the count runs across sweeps
                mbar_wait(bars + (tile & 1u), (phase_bits >> (tile & 1u)) & 1u);
                phase_bits ^= 1u << (tile & 1u);
            }
            w = ring[((tile & 1u) * LPS + tl) * 32 + lane];
            if (!valid) w = OP_EMPTY;
            if (tl == LPS - 1 || it + 1 == nit) {  // the stage is consumed: refill it with the tile after next
                __syncwarp();
                if (lane == 0 && tile + 2 < ntile) issue_tile(tile + 2);
            }
#else
            w = take_line(s_line, lane);
            if (!valid) w = OP_EMPTY;
            if (base + 32 < M) fetch_line_pol(s_line, ops + base + 32, lane, pol_stream);
#endif
            const int type = !valid ? T_NONE : (w == OP_EMPTY ? T_EMPTY : (op_is_diag(w) ? T_DIAG : T_OFFD));
            // one decode for every lane: the op in the slot, or the op an empty slot proposes
            const uint32_t beff = w == OP_EMPTY ? pb : op_bond(w);
            kind = bkind<HAS_H>(D, beff);
            neww = w;
            const uint32_t ee = evars(beff, kind, v0, v1);
#ifdef QMCB_PREFETCH_PARENTS
            if (do_clus && w != OP_EMPTY && kind == KIND_BOND) {  // experiment: the parents the union stage will ask for, into L2
                asm volatile("prefetch.global.L2 [%0];" ::"l"(P + s_rep[v0]));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(P + s_rep[v1]));
            }
#endif
            if (do_diag) {
                // (written with selects rather than branches: the lanes of a step are a mix of empty slots, diagonal and
                // off-diagonal ops, so every branch on the slot type would be executed both ways anyway)
                const bool isE = type == T_EMPTY, isD = type == T_DIAG, offd = type == T_OFFD;
                fmask = __ballot_sync(FULL, offd);
                flipv = offd ? v0 : NONE32;
                // spins the weight looks at: the stored input bits of an existing op, the propagated state at this slot
                // for a proposal (v1 == v0 for one-variable ops, so both loads are always in range)
                const uint32_t bit0 = state_bit(s_st, v0), bit1 = kind == KIND_BOND ? state_bit(s_st, v1) : 0u;
                uint32_t s0 = isE ? bit0 : (op_in(w) & 1u), s1 = isE ? bit1 : ((op_in(w) >> 1) & 1u);
                if (fmask) {  // off-diagonal ops flip their variable for the later lanes of this step
                    s_fl[lane] = flipv;
                    if (offd) atomicOr(&s_cd[v0 >> 5], 1u << (v0 & 31));
                    __syncwarp();
                    if (isE && (fmask & lt_mask) && (state_bit(s_cd, v0) || (kind == KIND_BOND && state_bit(s_cd, v1))))
                        for (uint32_t m = fmask & lt_mask; m; m &= m - 1) {
                            const uint32_t fv = s_fl[__ffs(m) - 1];
                            s0 ^= (fv == v0), s1 ^= (kind == KIND_BOND && fv == v1);
                        }
                }
                double num, rn;
                if (PK) {
                    const uint32_t c = wclass(ee, kind, s0, s1);
                    num = numtab[c], rn = rnumtab[c];
                } else {
                    num = bn * bweight<HAS_H>(Hm, beff, kind, s0, s1);
                    rn = isD ? 1.0 / num : 0.0;
                }
                // conservative classification over the n interval this lane can see (the rules are monotone in den)
                const uint32_t emask = __ballot_sync(FULL, isE), dmask = __ballot_sync(FULL, isD);
                const double dlo = (double)(M - (n + (uint32_t)__popc(emask & lt_mask)));
                const double dhi = (double)(M - (n - (uint32_t)__popc(dmask & lt_mask)));
                // removal of a diagonal op: sure if den + 1 > num at the smallest den or the word is below the lower
                // threshold; surely kept if it is at or above the upper one; in between: the exact rule below
                const double q_lo = (dlo + 1.0) * rn * G_LO, q_hi = (dhi + 1.0) * rn * G_HI;
                const bool d_rm = (dlo + 1.0 > num) || (wA < bool_threshold(q_lo));
                const bool d_keep = q_hi < 1.0 && wA >= bool_threshold(q_hi);
                // insertion into an empty slot: sure if num > den at the largest den
                const bool e_ins = num > dhi;
                const bool needdiv = isE && !e_ins && num > 0.0;
                int dec = isE ? (e_ins ? 1 : 0) : ((isD && d_rm) ? -1 : 0);  // +1 insert, -1 remove
                bool amb = isD && !d_rm && !d_keep;
                if (__any_sync(FULL, needdiv)) {
                    // insertion with num <= den: bounds from the two reciprocals of the step's den interval
                    const double DLO = (double)(M - (n + (uint32_t)__popc(emask))), DHI = (double)(M - (n - (uint32_t)__popc(dmask)));
                    const double rA = 1.0 / DLO, rB = 1.0 / DHI;
                    if (needdiv) {
                        const double q_lo = num * rB * G_LO, q_hi = num * rA * G_HI;
                        if (wB < bool_threshold(q_lo)) dec = 1;
                        else if (!(q_hi < 1.0 && wB >= bool_threshold(q_hi))) amb = true;
                    }
                }
                uint32_t insm = __ballot_sync(FULL, dec == 1), remm = __ballot_sync(FULL, dec == -1);
                for (uint32_t ambm = __ballot_sync(FULL, amb); ambm; ambm &= ambm - 1) {
                    // the exact rule at the exact n, in lane order (earlier lanes are final)
                    const int u = __ffs(ambm) - 1;
                    if (lane == u) {
                        const uint32_t nl = n + (uint32_t)__popc(insm & lt_mask) - (uint32_t)__popc(remm & lt_mask);
                        if (type == T_EMPTY) {
                            const double den = (double)(M - nl);
                            bool acc = num > den;
                            if (!acc) {
                                const double pr = num / den;
                                if (pr == 1.0) acc = true;
                                else if (!(pr >= 0.0 && pr < 1.0)) err |= DEV_ERR_PROB;
                                else acc = wB < bool_threshold(pr);
                            }
                            dec = acc ? 1 : 0;
                        } else {
                            const double den = (double)(M - nl) + 1.0;
                            bool rem = den > num;
                            if (!rem) {
                                const double pr = den / num;
                                if (pr == 1.0) rem = true;
                                else if (!(pr >= 0.0 && pr < 1.0)) err |= DEV_ERR_PROB;
                                else rem = wA < bool_threshold(pr);
                            }
                            dec = rem ? -1 : 0;
                        }
                    }
                    const int du = __shfl_sync(FULL, dec, u);
                    if (du == 1) insm |= 1u << u;
                    else if (du == -1) remm |= 1u << u;
                }
                n += (uint32_t)__popc(insm) - (uint32_t)__popc(remm);
                const uint32_t pbits = s0 | (s1 << 1);
                neww = dec == 1 ? make_op(beff, pbits, pbits) : (dec == -1 ? OP_EMPTY : w);
                changed = neww != w;
            }
            if (neww == OP_EMPTY) kind = -1;  // no op in this slot after the diagonal update
            }  // role A
            // what is left of the diagonal step (store, state flips) and the words of the next step do not depend on the
            // union-find: they run between the CAS and the look at its result
            auto step_tail = [&]() {
                if (changed) st_cg_pol(ops + p, neww, pol_stream);
                if (fmask) {
                    if (flipv != NONE32) atomicXor(&s_st[flipv >> 5], 1u << (flipv & 31)), atomicAnd(&s_cd[flipv >> 5], ~(1u << (flipv & 31)));
                    __syncwarp();
                }
                if (do_diag && it + 1 < nit) draw(base + 32u + (uint32_t)lane);
            };
            uint32_t pc = p, itc = it;  // the step the cluster bookkeeping works on
            if (PIPE) {
                if (roleA && it < nit) {
                    hand[(it & 1u) * 32 + lane] = neww;
                    step_tail();
                }
                pc = p - 32u, itc = it - 1u;
                if (roleB) {  // role B works on the step role A finished one iteration ago
                    kind = -1;
                    if (it >= 1) {
                        const uint32_t wf = hand[((it - 1u) & 1u) * 32 + lane];
                        if (wf != OP_EMPTY) {
                            kind = bkind<HAS_H>(D, op_bond(wf));
                            evars(op_bond(wf), kind, v0, v1);
                        }
                    }
                }
            }
            if (PIPE == 3 && roleC && do_clus && it >= 2) {  // role C: the unions of step it - 2
                const uint32_t *u = un + ((it - 2u) & 1u) * 96;
                const uint32_t meta = u[64 + lane];
                if (meta >> 30) {
                    const uint32_t ra = u[lane], rb = u[32 + lane], cv0 = meta & 0x3FFFu, cv1 = (meta >> 14) & 0x3FFFu;
                    uint32_t a = ra, b = rb;
                    uint32_t pa = ld_cg_pol(P + a, pol_keep), pq = ld_cg_pol(P + b, pol_keep);
                    const uint32_t pa0 = pa, pq0 = pq;
                    while (pa != a || pq != b) {
                        a = pa, b = pq;
                        pa = ld_cg_pol(P + a, pol_keep), pq = ld_cg_pol(P + b, pol_keep);
                    }
                    if (pa0 != a) st_cg_pol(P + ra, a, pol_keep);  // a stale value is still an ancestor
                    if (pq0 != b) st_cg_pol(P + rb, b, pol_keep);
                    if (a != b) {
                        if (a > b) {
                            const uint32_t t = a;
                            a = b, b = t;
                        }
                        if (atomicCAS(P + b, b, a) != b) uf_union_pol(P, a, b, pol_keep);  // lost a race against a lane of this step: redo
                    }
                    if ((meta >> 28) & 1u) atomicCAS(&s_rep[cv0], ra, a);  // cache the root (only over the entry the segment stage read)
                    if ((meta >> 29) & 1u) atomicCAS(&s_rep[cv1], rb, a);
                }
            }
            if (do_clus && roleB && (!PIPE || (it >= 1 && it <= nit))) {
                if (PIPE && itc == half_it) ks_half = nsite;
                if (PIPE == 3 && itc == third_it) ks_third = nsite;
                // ---- segments and unions on the final ops of this step
                const uint32_t smask = __ballot_sync(FULL, kind == KIND_SITE);
                const uint32_t myid = N + nsite + (uint32_t)__popc(smask & lt_mask);
                bool collide = false;
                if (kind == KIND_SITE) {
                    st_cg_pol(P + myid, myid, pol_keep);
                    const uint32_t bit = 1u << (v0 & 31);
                    collide = atomicOr(&s_sb[v0 >> 5], bit) & bit;  // another site op of this step on the same variable
                }
                if (!alltb && kind >= 0) {
                    if (!state_bit(s_tb, v0)) atomicOr(&s_tb[v0 >> 5], 1u << (v0 & 31));
                    if (kind == KIND_BOND && !state_bit(s_tb, v1)) atomicOr(&s_tb[v1 >> 5], 1u << (v1 & 31));
                }
                // a member of the set of the segment open on my variables before my slot: the table entry, unless a
                // site op of this step cuts the variable at an earlier lane
                uint32_t ra = 0, rb = 0, oa = 0, ob = 0;
                bool fa = false, fb = false;
                if (kind >= 0) {
                    oa = ra = s_rep[v0], fa = true;
                    if (kind == KIND_BOND) ob = rb = s_rep[v1], fb = true;
                }
                __syncwarp();
                const bool hit = collide || (kind >= 0 && kind != KIND_SITE && (state_bit(s_sb, v0) || (kind == KIND_BOND && state_bit(s_sb, v1))));
                if (__any_sync(FULL, hit)) {
                    const uint32_t nokey = 0x80000000u | (uint32_t)lane;
                    const uint32_t ma = __match_any_sync(FULL, kind >= 0 ? v0 : nokey) & smask & lt_mask;
                    const uint32_t mb = __match_any_sync(FULL, kind == KIND_BOND ? v1 : (kind == KIND_SITE ? v0 : nokey)) & smask & lt_mask;
                    if (kind >= 0 && ma) ra = N + nsite + (uint32_t)__popc(smask & ((1u << (31 - __clz(ma))) - 1u)), fa = false;
                    if (kind == KIND_BOND && mb) rb = N + nsite + (uint32_t)__popc(smask & ((1u << (31 - __clz(mb))) - 1u)), fb = false;
                }
                __syncwarp();  // new ids are initialised before anyone follows them; bitmap reads are done
                if (kind == KIND_SITE) atomicAnd(&s_sb[v0 >> 5], ~(1u << (v0 & 31)));
                if (kind >= 0) st_cg_pol(sid + pc, ra, pol_stream);  // P3 looks the input-side flip up through this id
                uint32_t ua = 0, ub = 0, uold = 0;
                bool casd = false;
                if (PIPE == 3) {  // the unions of this step go to role C
                    uint32_t *u = un + (itc & 1u) * 96;
                    const bool dou = kind == KIND_BOND && ra != rb;
                    u[lane] = ra, u[32 + lane] = rb;
                    u[64 + lane] = v0 | (v1 << 14) | (fa ? 1u << 28 : 0u) | (fb ? 1u << 29 : 0u) | (dou ? 1u << 30 : 0u);
                } else if (kind == KIND_BOND && ra != rb) {
                    // lock-free min-root union (sse_warp.cuh uf_union_pol), split: climb, hook with a CAS -- and look at
                    // the CAS result only after the independent tail of the step (a lost race is redone there)
                    uint32_t a = ra, b = rb;
                    uint32_t pa = ld_cg_pol(P + a, pol_keep), pq = ld_cg_pol(P + b, pol_keep);
                    const uint32_t pa0 = pa, pq0 = pq;
                    while (pa != a || pq != b) {
                        a = pa, b = pq;
                        pa = ld_cg_pol(P + a, pol_keep), pq = ld_cg_pol(P + b, pol_keep);
                    }
                    if (pa0 != a) st_cg_pol(P + ra, a, pol_keep);  // a stale value is still an ancestor
                    if (pq0 != b) st_cg_pol(P + rb, b, pol_keep);
                    if (a != b) {
                        if (a > b) {
                            const uint32_t t = a;
                            a = b, b = t;
                        }
                        uold = atomicCAS(P + b, b, a), casd = true, ua = a, ub = b;
                    }
                    if (fa) atomicCAS(&s_rep[v0], oa, a);  // cache the root: equal roots skip the union
                    if (fb) atomicCAS(&s_rep[v1], ob, a);
                }
                if (HAS_H && kind == KIND_LONG) {
                    atomicOr(&frz[ra >> 5], 1u << (ra & 31));
                    anylong = true;
                }
                if (kind == KIND_SITE) atomicMax(&s_rep[v0], myid);  // the site op with the highest lane owns the variable from here on (roots are minima: a cached root is below every id of this step)
                if (!PIPE) step_tail();
                if (casd && uold != ub) uf_union_pol(P, ua, ub, pol_keep);  // lost a race against a lane of this step: redo
                nsite += (uint32_t)__popc(smask);
                if (!alltb && (itc & 31u) == 31u) {  // once every variable has an op the touched bits need no more updates
                    uint32_t cnt = 0;
                    for (uint32_t j = lane; j < Nw; j += 32) cnt += (uint32_t)__popc(s_tb[j]);
                    alltb = __reduce_add_sync(FULL, cnt) == N;
                }
                __syncwarp();
            } else if (!PIPE) step_tail();
#ifdef QMCB_PHASE_TIMERS
            const long long tb_ = clock64();  // cycles every role waits at the step barrier: dbg[56 + role]
            PAIR_SYNC();
            if (PIPE && lane == 0 && D.dbg) atomicAdd(&D.dbg[56 + role], (unsigned long long)(clock64() - tb_));
#else
            PAIR_SYNC();
#endif
        }
        if (do_diag && roleA && lane == 0) D.n[r] = n;
        if (PIPE) {  // role B needs n for the closure, the others the site ops before their share of P3
            if (lane == 0) {
                if (roleA) xchg[0] = n;
                if (roleB) xchg[2] = (half_it >= nit) ? nsite : ks_half, xchg[3] = (third_it >= nit) ? nsite : ks_third;
            }
            PAIR_SYNC();
            n = xchg[0];
            ks_half = xchg[2], ks_third = xchg[3];
        }

        CTICK(1);
        uint32_t ncl = 0;
        if (do_clus && n > 0) {
            const uint64_t c0 = cur;
            if (roleB) {
            // periodic closure: the segment open at the end of variable v is the one crossing p = 0
            for (uint32_t v = lane; v < N; v += 32) {
                const uint32_t rp = s_rep[v];
                if (rp != v) uf_union_cg(P, v, rp);
            }
            __syncwarp();
            __threadfence_block();
            const uint32_t nseg = N + nsite;
            const uint32_t nwords = (nseg + 31) / 32;
            bool frozen_all = false;
            if (HAS_H) {
                frozen_all = __any_sync(FULL, anylong);
                for (int32_t wd = (int32_t)nwords - 1; wd >= 0; wd--) {  // push frozen marks up to the roots: descending ids, parents are smaller
                    const uint32_t x = (uint32_t)wd * 32 + lane;
                    const uint32_t par = x < nseg ? ld_cg(P + x) : x;
                    for (;;) {
                        const uint32_t fzw = *(volatile uint32_t *)&frz[wd];
                        const bool mine = x < nseg && ((fzw >> lane) & 1u) && par != x;
                        bool changed = false;
                        if (mine) {
                            const uint32_t old = atomicOr(&frz[par >> 5], 1u << (par & 31));
                            changed = ((old >> (par & 31)) & 1u) == 0 && (par >> 5) == (uint32_t)wd;
                        }
                        if (!__any_sync(FULL, changed)) break;
                    }
                    __syncwarp();
                }
            }
            CTICK(2);
            // =========================== P2: one flip bit per segment ===========================
            // ascending ids (parent id < child id).  Four words (= one Philox block of flip bits) per round: their parents
            // are loaded together, then the decisions of parents from earlier rounds, and only the parents inside the round
            // are resolved in order -- two dependent L2 round trips per four words instead of two per word.
            uint32_t nroots = 0;
            const uint32_t nblk = (nwords + 3) >> 2;
            for (uint32_t blk = 0; blk < nblk; blk++) {
                const uint32_t wbot = blk * 4;
                uint32_t par4[4], dw4[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t x = (wbot + j) * 32 + lane;
                    par4[j] = x < nseg ? ld_cg(P + x) : x;
                    if (nsite == 0 && x < nseg) par4[j] = 0u;  // cluster.rs:98-107: no cluster edge => one cluster
                }
                const uint32_t lo_edge = wbot * 32u;  // ids < lo_edge were decided in earlier rounds
#pragma unroll
                for (int j = 0; j < 4; j++) dw4[j] = par4[j] < lo_edge ? ld_cg(decb + (par4[j] >> 5)) : 0u;
                const Philox4 rb4 = philox4x32_10(blk, (uint32_t)c0, (uint32_t)(c0 >> 32), QMCB_TAG_CLUS, k0, k1);
                uint32_t out4[4] = {0, 0, 0, 0};
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t wd = wbot + j;
                    const uint32_t x = wd * 32 + lane, par = par4[j];
                    const uint32_t rword = j == 0 ? rb4.x : (j == 1 ? rb4.y : (j == 2 ? rb4.z : rb4.w));
                    const bool root = x < nseg && par == x;
                    nroots += (uint32_t)__popc(__ballot_sync(FULL, root));
                    bool dec = false, known = root || x >= nseg;
                    if (root) {
                        dec = (rword >> lane) & 1u;
                        if (HAS_H) dec = dec && !((nsite == 0) ? frozen_all : ((ld_cg(frz + wd) >> lane) & 1u));
                    } else if (x < nseg && par < lo_edge) {
                        dec = (dw4[j] >> (par & 31)) & 1u;
                        known = true;
                    } else if (x < nseg && (par >> 5) < wd) {  // parent in an earlier word of this round
                        const uint32_t jj = (par >> 5) - wbot;
                        const uint32_t pw = jj == 0 ? out4[0] : (jj == 1 ? out4[1] : out4[2]);
                        dec = (pw >> (par & 31)) & 1u;
                        known = true;
                    }
                    // parents inside this word: resolve by rounds.  Consecutive ids are consecutive site ops of the string and
                    // min-root unions leave chains among them, so a lane whose parent is not decided yet jumps to its parent's
                    // parent (pointer doubling through a shuffle): log2(depth) rounds instead of depth
                    uint32_t pl = known ? (uint32_t)lane : par - wd * 32;
                    for (;;) {
                        const uint32_t kmask = __ballot_sync(FULL, known), dmask = __ballot_sync(FULL, dec);
                        if (kmask == FULL) {
                            out4[j] = dmask;
                            if (lane == 0 && wd < nwords) st_cg(decb + wd, dmask);
                            break;
                        }
                        const uint32_t ppl = __shfl_sync(FULL, pl, pl);
                        if (!known) {
                            if ((kmask >> pl) & 1u) dec = (dmask >> pl) & 1u, known = true;
                            else pl = ppl;
                        }
                    }
                }
                __syncwarp();
            }
            uint32_t untouched = 0;
            for (uint32_t j = lane; j < Nw; j += 32) {
                const uint32_t validm = (j == Nw - 1 && (N & 31u)) ? ((1u << (N & 31u)) - 1u) : FULL;
                untouched += (uint32_t)__popc(~s_tb[j] & validm);
            }
            untouched = __reduce_add_sync(FULL, untouched);
            ncl = nsite == 0 ? 1u : nroots - untouched;
            }  // role B
            __threadfence_block();
            __syncwarp();
            if (PIPE) {  // the flip bits are complete: both roles apply them
                if (roleB && lane == 0) xchg[1] = ncl;
                PAIR_SYNC();
                ncl = xchg[1];
            }

            CTICK(3);
            // =========================== P3: apply the flips (stateless) ===========================
            // PIPE: role A applies the slots before half_it * 32, role B the rest, starting from its own count of the site ops before them
            const uint32_t cut1 = min(M, half_it * 32u), cut2 = PIPE == 3 ? max(cut1, min(M, third_it * 32u)) : M;
            const uint32_t p3_lo = !PIPE || roleA ? 0u : (roleB ? cut1 : cut2), p3_hi = !PIPE ? M : (roleA ? cut1 : (roleB ? cut2 : M));
            uint32_t ks = !PIPE || roleA ? 0u : (roleB ? ks_half : ks_third);
            const uint32_t EN = E + N;
            // software pipeline: the op words and records of the next four lines are requested before the flip bits of the
            // current four are looked up, so one iteration waits for one round trip (the gathers), not three
            uint32_t wN[4], sN[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t p = p3_lo + 32u * j + lane;
                wN[j] = p < p3_hi ? ld_cg_pol(ops + p, pol_stream) : OP_EMPTY;
                sN[j] = p < p3_hi ? ld_cg_pol(sid + p, pol_stream) : 0u;
            }
            for (uint32_t base = p3_lo; base < p3_hi; base += 128) {
                uint32_t w4[4], s4[4], sm4[4], di4[4], do4[4];
#pragma unroll
                for (int j = 0; j < 4; j++) w4[j] = wN[j], s4[j] = sN[j];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t p = base + 128u + 32u * j + lane;
                    wN[j] = p < p3_hi ? ld_cg_pol(ops + p, pol_stream) : OP_EMPTY;
                    sN[j] = p < p3_hi ? ld_cg_pol(sid + p, pol_stream) : 0u;
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t b = op_bond(w4[j]);
                    sm4[j] = __ballot_sync(FULL, w4[j] != OP_EMPTY && b >= E && b < EN);
                }
                // every flip-bit word this iteration needs is requested before any is used (the stores below could
                // alias them as far as the compiler knows, so the order is written out by hand)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const bool site = (sm4[j] >> lane) & 1u;
                    const uint32_t id = N + ks + (uint32_t)__popc(sm4[j] & lt_mask);
                    di4[j] = w4[j] != OP_EMPTY ? ld_cg(decb + (s4[j] >> 5)) : 0u;
                    do4[j] = site ? ld_cg(decb + (id >> 5)) : 0u;
                    s4[j] = (s4[j] & 31u) | ((id & 31u) << 8) | (site ? 0x10000u : 0u);
                    ks += (uint32_t)__popc(sm4[j]);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t p = base + 32u * j + lane, w = w4[j];
                    if (w != OP_EMPTY) {
                        const uint32_t b = op_bond(w);
                        const bool din = (di4[j] >> (s4[j] & 31u)) & 1u;
                        const bool dout = (s4[j] & 0x10000u) ? ((do4[j] >> ((s4[j] >> 8) & 31u)) & 1u) : din;
                        if (din || dout) {
                            const uint32_t mask = b < E ? 3u : 1u;
                            st_cg_pol(ops + p, make_op(b, op_in(w) ^ (din ? mask : 0u), op_out(w) ^ (dout ? mask : 0u)), pol_stream);
                        }
                    }
                }
            }
            CTICK(4);
            // spins: the segment of variable v crossing p = 0 has id v
            if (roleA)
                for (uint32_t j = lane; j < Nw; j += 32) s_st[j] ^= ld_cg(decb + j) & s_tb[j];
            cur = c0 + 1;
            __syncwarp();
        }
        if (do_clus && roleA) {
            // free spins: qmc_ising.rs:780-784
            for (uint32_t base = 0; base < N; base += 32) {
                const uint32_t v = base + lane;
                const bool fr = v < N && !((s_tb[v >> 5] >> (v & 31)) & 1u);
                const uint32_t m = __ballot_sync(FULL, fr);
                bool bit = false;
                if (fr) bit = stream_word(key, cur + __popc(m & lt_mask)) < 0x8000000000000000ull;
                const uint32_t setm = __ballot_sync(FULL, bit);
                if (lane == 0 && m) s_st[base >> 5] = (s_st[base >> 5] & ~m) | setm;
                cur += __popc(m);
            }
            __syncwarp();
            if (lane == 0) D.ncl[r] = ncl;
        }
        if (roleA)
            for (uint32_t j = lane; j < Nw; j += 32) gstate[j] = s_st[j];
        if (lane == 0 && roleA) {
            D.cursor[r] = cur;
            if (phases & 4u) {
                const uint32_t grown = n + n / 2;  // qmc_ising.rs:786
                if (grown > M) D.M[r] = grown;
            }
        }
        if (phases & 8u) {
            done++;
            const uint64_t idx = done - sample_origin;
            if (lane == 0 && roleA) D.vupd[r] += n;
            if (idx % sample_freq == 0 && roleA) {
                if (lane == 0) D.sum_n[r] += n;
                if (samples) {
                    uint8_t *dst = samples + ((size_t)r * samples_per_rep + (idx / sample_freq - 1)) * N;
                    for (uint32_t v = lane; v < N; v += 32) dst[v] = (uint8_t)state_bit(s_st, v);
                }
            }
            if (lane == 0 && roleA) D.done[r] = done;
        }
        __syncwarp();
        PAIR_SYNC();  // role B reads n, the cursor and the cutoff of the next sweep after role A wrote them
        CTICK(5);
    }
    if (err) atomicOr(D.status, err);
#undef PAIR_SYNC
}

// returns the number of kernel launches, or -1 if this shape is not supported by the warp kernels
int launch_sse_counter(const SseDev &D, const SseTuning &T, uint64_t target, uint32_t phases, uint64_t sample_freq, uint64_t sample_origin,
                       uint8_t *samples, uint64_t samples_per_rep, cudaStream_t st) {
    if (D.hb_cum) return -1;  // the heat-bath rule has no COUNTER-mode contract
    const size_t smem = cnt_smem_bytes(D.N, D.Nw);
    if (smem * QMCB_WPB + 1024 > 227 * 1024) return -1;
    const uint32_t blocks = (D.R + QMCB_WPB - 1) / QMCB_WPB;
    typedef void (*Kern)(SseDev, uint64_t, uint32_t, uint64_t, uint64_t, uint8_t *, uint64_t, uint32_t, uint32_t);
    const int nsm = T.nsm > 0 ? T.nsm : 148;
    const size_t wanted = (blocks + nsm - 1) / nsm;
    int minb = T.minblocks;
    if (minb <= 0) {
        const size_t resident = std::min((size_t)(227 * 1024) / (smem * QMCB_WPB + 1024), wanted);
        minb = resident <= 4 ? 4 : 7;
    }
    if (minb != 4) minb = 7;
    const bool have_epk = D.epk && !D.ham && T.epk;
    size_t epk_bytes = 0;
    if (minb == 4 && have_epk) {  // block-shared packed edge table if it does not cost a resident block
        const size_t want = ((size_t)D.E * 4 + 15) / 16 * 16;
        const size_t without = std::min((size_t)(227 * 1024) / (smem * QMCB_WPB + 1024), wanted);
        const size_t with = std::min((size_t)(227 * 1024) / (smem * QMCB_WPB + want + 1024), wanted);
        if (with >= 1 && with >= std::min<size_t>(without, 4)) epk_bytes = want;
    }
    const int pk = !have_epk ? 0 : (epk_bytes ? 1 : 2);
    // two warps per replica (PIPE) when at most two blocks of four replicas per SM are wanted and fit
    // (T.pipe: 0 never, 1 three roles if they fit else two, 2 two roles)
    const size_t smem_p2 = cnt_smem_bytes(D.N, D.Nw, 2), smem_p3 = cnt_smem_bytes(D.N, D.Nw, 3);
    int pipe = 0;
    if (minb == 4 && T.pipe && wanted <= 2) {
        if (T.pipe != 2 && D.N <= 16384 && (size_t)(227 * 1024) / (smem_p3 * QMCB_WPB + epk_bytes + 1024) >= wanted) pipe = 3;
        else if ((size_t)(227 * 1024) / (smem_p2 * QMCB_WPB + epk_bytes + 1024) >= wanted) pipe = 2;
    }
    const size_t smem_used = pipe == 3 ? smem_p3 : (pipe == 2 ? smem_p2 : smem);
    Kern kern;
#define PICKC(MINB_, MH_, PIPE_)                                                                                                    \
    switch (pk) {                                                                                                                   \
        case 1: kern = D.has_h ? k_sse_counter<true, MINB_, MH_, 1, PIPE_> : k_sse_counter<false, MINB_, MH_, 1, PIPE_>; break;     \
        case 2: kern = D.has_h ? k_sse_counter<true, MINB_, MH_, 2, PIPE_> : k_sse_counter<false, MINB_, MH_, 2, PIPE_>; break;     \
        default: kern = D.has_h ? k_sse_counter<true, MINB_, MH_, 0, PIPE_> : k_sse_counter<false, MINB_, MH_, 0, PIPE_>; break;    \
    }
    if (D.ham) {
        if (pipe == 3) { PICKC(4, true, 3) } else if (pipe == 2) { PICKC(4, true, 2) } else if (minb == 4) { PICKC(4, true, 0) } else { PICKC(7, true, 0) }
    } else if (pipe == 3) { PICKC(4, false, 3) } else if (pipe == 2) { PICKC(4, false, 2) } else if (minb == 4) { PICKC(4, false, 0) } else { PICKC(7, false, 0) }
#undef PICKC
    {
        static std::mutex mu;
        static std::map<std::pair<const void *, int>, int> done;
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        auto it = done.find({(const void *)kern, dev});
        if (it == done.end()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (T.carveout >= 0 && (it == done.end() || it->second != T.carveout))
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, T.carveout);
        done[{(const void *)kern, dev}] = T.carveout;
    }
    size_t dyn = smem_used * QMCB_WPB + epk_bytes + (size_t)std::max(T.pad, 0);
    if (dyn > 227 * 1024) dyn = 227 * 1024;
    kern<<<blocks, (pipe ? 32 * pipe : 32) * QMCB_WPB, dyn, st>>>(D, target, phases, sample_freq, sample_origin, samples, samples_per_rep, (uint32_t)smem_used,
                                                          epk_bytes ? (uint32_t)(smem_used * QMCB_WPB) : 0u);
    return 1;
}
