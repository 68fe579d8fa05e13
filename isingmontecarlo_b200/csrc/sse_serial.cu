// sse_serial.cu -- reference-order SSE sweep, one warp per replica.
//
// This is the STRICT path: the diagonal update walks the slots in order with a live n and the
// cluster update numbers clusters in the reference's LIFO discovery order, so under the same
// injected word stream the spins and operator strings are bit-identical to the reference's own
// update path (diagonal.rs:142-191, cluster.rs:36-271, qmc_ising.rs:644-795).  The order-dependent
// parts run on lane 0 of the replica's warp; initialisation, flip-bit generation, the flip/apply
// pass, free-spin draws and sampling use all 32 lanes with coalesced accesses.
// It also carries a serial FAST-order cluster step used as the on-device cross-check of the
// warp-parallel FAST kernels in sse_fast.cu.
#include <algorithm>

#include "sse.cuh"

#define SIDE_IN 0u
#define SIDE_OUT 1u

// STRICT workspace: one 32-byte record (= one DRAM sector) per slot, so that a step of the cluster walk
// touches one sector per node:  [0] op word, [1..2] previous leg on the world line of leg 0 / 1,
// [3..4] next leg, [5] cluster of the inputs, [6] cluster of the outputs, [7] unused.
// legs are (p << 1 | rel), NONE32 at the ends (fast_ops.rs:181-190; boundaries cluster.rs:50-52)
#define REC_OP(V, p) ((V).rec[8 * (size_t)(p)])
#define REC_LINK(V, p, i) ((V).rec[8 * (size_t)(p) + 1 + (i)])
#define REC_BND(V, p, side) ((V).rec[8 * (size_t)(p) + 5 + (side)])

// phase timers of the STRICT cluster step (-DQMCB_PHASE_TIMERS + qmcb_set_option("debug_counters")): cycles of lane 0
// summed over replicas and sweeps in dbg[48..]: 0 links, 1 labelling walk, 2 flip bits, 3 apply, 4 free spins;
// dbg[56..]: 0 interior pops, 1 site arrivals, 2 interior-op arrivals that were labelled, 3 world-line wraps, 4 frontier pops
#ifdef QMCB_PHASE_TIMERS
#define ST_MARK(i)                                                                   \
    do {                                                                             \
        __syncwarp();                                                                \
        if (lane == 0 && D.dbg) {                                                    \
            const long long t_ = clock64();                                          \
            atomicAdd(&D.dbg[48 + (i)], (unsigned long long)(t_ - st_t));            \
            st_t = t_;                                                               \
        }                                                                            \
    } while (0)
#define ST_COUNT(i) (cnt_[i]++)
#define ST_COUNT_DECL unsigned long long cnt_[5] = {0, 0, 0, 0, 0}
#define ST_COUNT_FLUSH                                                               \
    do {                                                                             \
        if (D.dbg)                                                                   \
            for (int i_ = 0; i_ < 5; i_++) atomicAdd(&D.dbg[56 + i_], cnt_[i_]);     \
    } while (0)
#else
#define ST_MARK(i) ((void)0)
#define ST_COUNT(i) ((void)0)
#define ST_COUNT_DECL ((void)0)
#define ST_COUNT_FLUSH ((void)0)
#endif

struct Rep {
    uint32_t *ops, *state, *vfirst, *vlast, *cur, *rec, *frontier, *interior, *bits, *frozen, *parent;
    uint32_t *ends, *ent;
    uint4 *wl;  // world-line layout of the STRICT workspace: the same buffer as rec
    uint64_t fcap, icap;
};

__device__ __forceinline__ Rep rep_view(const SseDev &D, uint32_t r) {
    Rep v;
    v.ops = D.ops + (size_t)r * D.cap;
    v.state = D.state + (size_t)r * D.Nw;
    v.vfirst = D.vfirst + (size_t)r * D.N;
    v.vlast = D.vlast + (size_t)r * D.N;
    v.cur = D.cur + (size_t)r * D.N;
    v.rec = D.rec ? D.rec + (size_t)r * strict_rec_stride(D) : nullptr;
    v.wl = reinterpret_cast<uint4 *>(v.rec);
    v.ent = D.ent ? D.ent + (size_t)r * D.cap : nullptr;
    v.fcap = 2 * D.cap + 16, v.icap = 4 * D.cap + 16;
    v.frontier = D.frontier ? D.frontier + (size_t)r * v.fcap : nullptr;
    v.interior = D.interior ? D.interior + (size_t)r * v.icap : nullptr;
    v.bits = D.bits + (size_t)r * (D.cap / 32 + 2 + D.N / 32);
    v.frozen = D.frozen + (size_t)r * (D.cap / 32 + 2 + D.N / 32);
    v.parent = D.parent ? D.parent + (size_t)r * (D.N + D.cap + 1) : nullptr;
    v.ends = D.ends + (size_t)r * 4;
    return v;
}

// ------------------------------------------------------------------------------------------
// diagonal update (diagonal.rs:114-135, :142-191; slot loop fast_ops.rs:611-637), lane 0
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void diag_serial(const SseDev &D, uint32_t r, const Rep &V) {
    const uint32_t M = D.M[r];
    uint32_t n = D.n[r];
    uint64_t cur = D.cursor[r];
    const uint64_t key = D.key[r];
    const double bn = D.beta[r] * (double)D.Nb;  // diagonal.rs:168, left-to-right product
    const Ham Hm = ham_view<true>(D, r);
    const uint64_t range = D.Nb;
    const uint64_t zone = (range << __clzll((long long)range)) - 1ull;
    int err = 0;
    for (uint32_t p = 0; p < M; p++) {
        uint32_t w = V.ops[p];
        uint32_t b;
        if (w == OP_EMPTY) {
            uint64_t hi, lo;  // gen_range(0..Nb): widening multiply + zone rejection
            do {
                uint64_t v = stream_word(key, cur++);
                hi = __umul64hi(v, range), lo = v * range;
            } while (lo > zone);
            b = (uint32_t)hi;
        } else if (op_is_diag(w)) {
            b = op_bond(w);
        } else {  // off-diagonal: propagate the state (diagonal.rs:154-160)
            int kind = bond_kind(D, op_bond(w));
            uint32_t v0, v1;
            bond_vars(D, op_bond(w), kind, v0, v1);
            uint32_t o = op_out(w);
            V.state[v0 >> 5] = (V.state[v0 >> 5] & ~(1u << (v0 & 31))) | ((o & 1u) << (v0 & 31));
            if (kind == KIND_BOND) V.state[v1 >> 5] = (V.state[v1 >> 5] & ~(1u << (v1 & 31))) | (((o >> 1) & 1u) << (v1 & 31));
            continue;
        }
        int kind = bond_kind(D, b);
        uint32_t v0, v1;
        bond_vars(D, b, kind, v0, v1);
        uint32_t s0 = state_bit(V.state, v0), s1 = kind == KIND_BOND ? state_bit(V.state, v1) : 0u;
        double num = bn * bond_weight(Hm, b, kind, s0, s1);
        double den = (double)(M - n);
        if (w == OP_EMPTY) {
            bool accept = num > den;
            if (!accept) {
                double pr = num / den;
                if (pr == 1.0) accept = true;
                else if (!(pr >= 0.0 && pr < 1.0)) err |= DEV_ERR_PROB;
                else accept = stream_word(key, cur++) < bool_threshold(pr);
            }
            if (accept) {
                uint32_t bitsv = s0 | (s1 << 1);
                V.ops[p] = make_op(b, bitsv, bitsv);
                n++;
            }
        } else {
            den = den + 1.0;
            bool remove = den > num;
            if (!remove) {
                double pr = den / num;
                if (pr == 1.0) remove = true;
                else if (!(pr >= 0.0 && pr < 1.0)) err |= DEV_ERR_PROB;
                else remove = stream_word(key, cur++) < bool_threshold(pr);
            }
            if (remove) {
                V.ops[p] = OP_EMPTY;
                n--;
            }
        }
    }
    D.n[r] = n;
    D.cursor[r] = cur;
    if (err) atomicOr(D.status, err);
}

// ------------------------------------------------------------------------------------------
// COUNTER-mode diagonal update, literal (oracle.c diagonal_update_counter), lane 0: the fallback for shapes the warp
// kernel does not take and its on-device cross-check ("impl" = 1)
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void diag_counter_serial(const SseDev &D, uint32_t r, const Rep &V) {
    const uint32_t M = D.M[r];
    uint32_t n = D.n[r];
    const uint64_t c0 = D.cursor[r];
    const uint64_t key = D.key[r];
    const double bn = D.beta[r] * (double)D.Nb;
    const Ham Hm = ham_view<true>(D, r);
    int err = 0;
    for (uint32_t p = 0; p < M; p++) {
        const uint32_t w = V.ops[p];
        if (w != OP_EMPTY && !op_is_diag(w)) {
            const int kind = bond_kind(D, op_bond(w));
            uint32_t v0, v1;
            bond_vars(D, op_bond(w), kind, v0, v1);
            const uint32_t o = op_out(w);
            V.state[v0 >> 5] = (V.state[v0 >> 5] & ~(1u << (v0 & 31))) | ((o & 1u) << (v0 & 31));
            if (kind == KIND_BOND) V.state[v1 >> 5] = (V.state[v1 >> 5] & ~(1u << (v1 & 31))) | (((o >> 1) & 1u) << (v1 & 31));
            continue;
        }
        const Philox4 o = philox4x32_10(p, (uint32_t)c0, (uint32_t)(c0 >> 32), QMCB_TAG_DIAG, (uint32_t)key, (uint32_t)(key >> 32));
        const uint64_t wA = ((uint64_t)o.y << 32) | o.x, wB = ((uint64_t)o.w << 32) | o.z;
        const uint32_t b = w == OP_EMPTY ? (uint32_t)__umul64hi(wA, (uint64_t)D.Nb) : op_bond(w);
        const int kind = bond_kind(D, b);
        uint32_t v0, v1;
        bond_vars(D, b, kind, v0, v1);
        const uint32_t s0 = state_bit(V.state, v0), s1 = kind == KIND_BOND ? state_bit(V.state, v1) : 0u;
        const double num = bn * bond_weight(Hm, b, kind, s0, s1);
        double den = (double)(M - n);
        if (w == OP_EMPTY) {
            bool accept = num > den;
            if (!accept) {
                const double pr = num / den;
                if (pr == 1.0) accept = true;
                else if (!(pr >= 0.0 && pr < 1.0)) err |= DEV_ERR_PROB;
                else accept = wB < bool_threshold(pr);
            }
            if (accept) {
                const uint32_t bitsv = s0 | (s1 << 1);
                V.ops[p] = make_op(b, bitsv, bitsv);
                n++;
            }
        } else {
            den = den + 1.0;
            bool remove = den > num;
            if (!remove) {
                const double pr = den / num;
                if (pr == 1.0) remove = true;
                else if (!(pr >= 0.0 && pr < 1.0)) err |= DEV_ERR_PROB;
                else remove = wA < bool_threshold(pr);
            }
            if (remove) {
                V.ops[p] = OP_EMPTY;
                n--;
            }
        }
    }
    D.n[r] = n;
    D.cursor[r] = c0 + 1;
    if (err) atomicOr(D.status, err);
}

// ------------------------------------------------------------------------------------------
// heat-bath diagonal update (heatbath.rs:106-127, :149-209), lane 0
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void diag_heatbath_serial(const SseDev &D, uint32_t r, const Rep &V) {
    const uint32_t M = D.M[r];
    uint32_t n = D.n[r];
    uint64_t cur = D.cursor[r];
    const uint64_t key = D.key[r];
    const Ham Hm = ham_view<true>(D, r);
    const double total = Hm.hb_total;
    const double bt = D.beta[r] * total;  // heatbath.rs:163,194
    int err = 0;
    for (uint32_t p = 0; p < M; p++) {
        const uint32_t w = V.ops[p];
        if (w == OP_EMPTY) {
            const double den = (double)(M - n) + bt;
            const double pr = bt / den;
            bool attempt;
            if (pr == 1.0) attempt = true;  // gen_bool(1.0): no draw
            else if (!(pr >= 0.0 && pr < 1.0)) attempt = false, err |= DEV_ERR_PROB;
            else attempt = stream_word(key, cur++) < bool_threshold(pr);
            if (!attempt) continue;
            double pd, c;
            do pd = unit_f64(stream_word(key, cur++)); while (!(pd < 1.0));            // gen_range(0. ..1.0)
            do c = unit_f64(stream_word(key, cur++)) * total; while (!(c < total));     // gen_range(0. ..total)
            const uint32_t b = hb_index_for_cumulative(Hm.hb_cum, D.Nb, c);
            if (b >= D.Nb) { err |= DEV_ERR_INVARIANT; continue; }
            const int kind = bond_kind(D, b);
            uint32_t v0, v1;
            bond_vars(D, b, kind, v0, v1);
            const uint32_t s0 = state_bit(V.state, v0), s1 = kind == KIND_BOND ? state_bit(V.state, v1) : 0u;
            if (pd * __ldg(Hm.hb_maxw + b) < bond_weight(Hm, b, kind, s0, s1)) {
                const uint32_t bitsv = s0 | (s1 << 1);
                V.ops[p] = make_op(b, bitsv, bitsv);
                n++;
            }
        } else if (op_is_diag(w)) {
            const double num = (double)(M - n + 1);
            const double pr = num / (num + bt);
            bool remove;
            if (pr == 1.0) remove = true;
            else if (!(pr >= 0.0 && pr < 1.0)) remove = false, err |= DEV_ERR_PROB;
            else remove = stream_word(key, cur++) < bool_threshold(pr);
            if (remove) {
                V.ops[p] = OP_EMPTY;
                n--;
            }
        } else {
            const int kind = bond_kind(D, op_bond(w));
            uint32_t v0, v1;
            bond_vars(D, op_bond(w), kind, v0, v1);
            const uint32_t o = op_out(w);
            V.state[v0 >> 5] = (V.state[v0 >> 5] & ~(1u << (v0 & 31))) | ((o & 1u) << (v0 & 31));
            if (kind == KIND_BOND) V.state[v1 >> 5] = (V.state[v1 >> 5] & ~(1u << (v1 & 31))) | (((o >> 1) & 1u) << (v1 & 31));
        }
    }
    D.n[r] = n;
    D.cursor[r] = cur;
    if (err) atomicOr(D.status, err);
}

// ------------------------------------------------------------------------------------------
// links (what FastOps::mutate_p maintains incrementally, fast_ops.rs:337-607), lane 0
// ------------------------------------------------------------------------------------------
__device__ void links_serial(const SseDev &D, uint32_t r, const Rep &V, bool full) {
    const uint32_t M = D.M[r];
    uint32_t first_p = NONE32, last_p = NONE32, first_site = NONE32;
    for (uint32_t p = 0; p < M; p++) {
        uint32_t w = V.ops[p];
        if (w == OP_EMPTY) continue;
        uint32_t b = op_bond(w);
        int kind = bond_kind(D, b);
        uint32_t vv[2];
        bond_vars(D, b, kind, vv[0], vv[1]);
        int nv = kind == KIND_BOND ? 2 : 1;
        if (first_p == NONE32) first_p = p;
        last_p = p;
        if (kind == KIND_SITE && first_site == NONE32) first_site = p;
        if (full) {
            REC_OP(V, p) = w;
            for (int i = 0; i < 4; i++) REC_LINK(V, p, i) = NONE32;
            REC_BND(V, p, 0) = NONE32, REC_BND(V, p, 1) = NONE32;
        }
        for (int k = 0; k < nv; k++) {
            uint32_t v = vv[k], me = (p << 1) | (uint32_t)k;
            uint32_t prev = V.vlast[v];
            if (full) {
                REC_LINK(V, p, k) = prev;
                if (prev != NONE32) REC_LINK(V, prev >> 1, 2 + (prev & 1u)) = me;
            }
            if (prev == NONE32) V.vfirst[v] = me;
            V.vlast[v] = me;
        }
    }
    V.ends[0] = first_p, V.ends[1] = last_p, V.ends[2] = first_site;
}

// The same links, built by the whole warp: 32 slots per step, two half-steps of 16 slots = 32 legs, one
// lane per leg.  match.any finds the legs of the half-step that sit on the same variable (the two legs of
// one op never do), the nearest earlier one is the predecessor, otherwise the table `last` (shared
// memory, one entry per variable) holds it.
__device__ void links_parallel(const SseDev &D, uint32_t r, const Rep &V, int lane, uint32_t *last) {
    const uint32_t M = D.M[r];
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (uint32_t v = lane; v < D.N; v += 32) last[v] = NONE32, V.vfirst[v] = NONE32;
    __syncwarp();
    uint32_t first_p = NONE32, last_p = NONE32, first_site = NONE32;
    for (uint32_t base = 0; base < M; base += 32) {
        const uint32_t p = base + lane;
        const uint32_t w = p < M ? V.ops[p] : OP_EMPTY;
        int kind = -1;
        uint32_t v0 = 0, v1 = 0;
        if (w != OP_EMPTY) {
            kind = bond_kind(D, op_bond(w));
            bond_vars(D, op_bond(w), kind, v0, v1);
            uint4 a, b;
            a.x = w, a.y = NONE32, a.z = NONE32, a.w = NONE32;
            b.x = NONE32, b.y = NONE32, b.z = NONE32, b.w = 0u;
            *reinterpret_cast<uint4 *>(&V.rec[8 * (size_t)p]) = a;
            *reinterpret_cast<uint4 *>(&V.rec[8 * (size_t)p + 4]) = b;
        }
        const uint32_t hasm = __ballot_sync(0xFFFFFFFFu, kind >= 0), sitem = __ballot_sync(0xFFFFFFFFu, kind == KIND_SITE);
        if (hasm) {
            if (first_p == NONE32) first_p = base + (uint32_t)__ffs(hasm) - 1u;
            last_p = base + 31u - (uint32_t)__clz(hasm);
        }
        if (sitem && first_site == NONE32) first_site = base + (uint32_t)__ffs(sitem) - 1u;
        __syncwarp();  // records of this step exist before anyone links to them
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int s = 16 * half + (lane >> 1);
            const uint32_t rel = (uint32_t)lane & 1u;
            const int sk = __shfl_sync(0xFFFFFFFFu, kind, s);
            const uint32_t sv0 = __shfl_sync(0xFFFFFFFFu, v0, s), sv1 = __shfl_sync(0xFFFFFFFFu, v1, s);
            const bool legvalid = sk >= 0 && (rel == 0 || sk == KIND_BOND);
            const uint32_t sv = rel ? sv1 : sv0;
            const uint32_t legm = __ballot_sync(0xFFFFFFFFu, legvalid);
            const uint32_t m = __match_any_sync(0xFFFFFFFFu, legvalid ? sv : (0x80000000u | (uint32_t)lane)) & legm;
            if (legvalid) {
                const uint32_t me = ((base + (uint32_t)s) << 1) | rel;
                uint32_t prev;
                if (m & lt_mask) {
                    const uint32_t j = 31u - (uint32_t)__clz(m & lt_mask);
                    prev = ((base + 16u * half + (j >> 1)) << 1) | (j & 1u);
                } else prev = last[sv];
                REC_LINK(V, base + s, rel) = prev;
                if (prev != NONE32) REC_LINK(V, prev >> 1, 2 + (prev & 1u)) = me;
                else V.vfirst[sv] = me;
                if ((m & ~lt_mask & ~(1u << lane)) == 0) last[sv] = me;  // last leg of the half-step on this variable
            }
            __syncwarp();
        }
    }
    for (uint32_t v = lane; v < D.N; v += 32) V.vlast[v] = last[v];
    if (lane == 0) V.ends[0] = first_p, V.ends[1] = last_p, V.ends[2] = first_site;
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// STRICT cluster labelling: cluster.rs:46-108 (frontier loop) and :193-271 (expansion), lane 0
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_rec_lo(const Rep &V, uint32_t p) { return *reinterpret_cast<const uint4 *>(&V.rec[8 * (size_t)p]); }
__device__ __forceinline__ uint4 ld_rec_hi(const Rep &V, uint32_t p) { return *reinterpret_cast<const uint4 *>(&V.rec[8 * (size_t)p + 4]); }

// the walk is a chain of dependent sector loads; when a leg is pushed its link is already in registers,
// so the record it will land on is requested right away (L2 prefetch) and is close by when the leg is popped
__device__ __forceinline__ void prefetch_rec(const Rep &V, uint32_t leg) {
    if (leg != NONE32) asm volatile("prefetch.global.L2 [%0];" ::"l"(&V.rec[8 * (size_t)(leg >> 1)]));
}
__device__ __forceinline__ uint32_t rec_link(const uint4 &lo, const uint4 &hi, uint32_t li) {
    return li == 0 ? lo.y : (li == 1 ? lo.z : (li == 2 ? lo.w : hi.x));
}

// The walk itself runs on lane 0; the other lanes join for the unmapped-op scan (cluster.rs:82-88), 32 slots per ballot.
// The tops of the two LIFO stacks live in shared memory (`stk`: STK_I interior entries, then a window of STK_F frontier
// entries); what does not fit spills to the global-memory stacks, which keep the exact LIFO order.
#define STK_I 192
#define STK_F 128
#define STK_WORDS 384  // per warp: max(STK_I, STK_IW) + STK_F
__device__ uint32_t label_strict(const SseDev &D, const Rep &V, int &err, int lane, uint32_t *stk) {
    const uint32_t last_p = V.ends[1], cp = V.ends[2];
    const uint32_t E = D.E, EN = D.E + D.N;
    uint32_t *const ist = stk, *const fst = stk + STK_I;
    uint64_t flen = 0, ilen = 0, fbase = 0;  // frontier entries [fbase, flen) are in fst, [0, fbase) in V.frontier
    auto fpush = [&](uint32_t x) {
        if (flen - fbase == STK_F) {  // window full: the lower half goes to global memory
            for (uint32_t j = 0; j < STK_F / 2; j++) V.frontier[fbase + j] = fst[j];
            for (uint32_t j = 0; j < STK_F / 2; j++) fst[j] = fst[j + STK_F / 2];
            fbase += STK_F / 2;
        }
        fst[flen - fbase] = x;
        flen++;
    };
    auto fpop = [&]() -> uint32_t {
        if (flen == fbase) {  // window empty: refill from global memory
            const uint64_t take = fbase < STK_F / 2 ? fbase : STK_F / 2;
            for (uint64_t j = 0; j < take; j++) fst[j] = V.frontier[fbase - take + j];
            fbase -= take;
        }
        flen--;
        return fst[flen - fbase];
    };
    auto ipush = [&](uint32_t x) {
        if (ilen < STK_I) ist[ilen] = x;
        else V.interior[ilen] = x;
        ilen++;
    };
    auto ipop = [&]() -> uint32_t {
        ilen--;
        return ilen < STK_I ? ist[ilen] : V.interior[ilen];
    };
    if (lane == 0) {
        fpush((cp << 1) | SIDE_OUT);  // cluster.rs:57-59
        fpush((cp << 1) | SIDE_IN);
    }
    // Interior entries.  The reference stacks (p, leg, side) and, when it pops one, looks the leg's link up in p's node
    // (cluster.rs:218-241).  p's record is in registers when its legs are pushed, so the link is resolved THEN and the
    // entry holds where the leg lands: (q << 1 | leg of q) << 1 | side of q it arrives at -- one dependent record load
    // per leg instead of two.  set_boundary(p, side) at pop time (:218) is a no-op for ops reached by the walk (both of
    // their sides are set when they are reached, :252-254); it matters for the legs of a non-edge START op, whose sides
    // are set leg by leg (which is what makes the reference push duplicates): those entries carry bit 31 and set
    // boundary (p0, side ^ 1 of the entry) when popped.  Push and pop order are the reference's.
    uint32_t cnum = 0, scan = 0;
    ST_COUNT_DECL;
    auto push_leg = [&](const uint4 &lo, const uint4 &hi, uint32_t k, uint32_t side, uint32_t v0, uint32_t v1, uint32_t flag) {
        uint32_t lk = rec_link(lo, hi, 2u * side + k);
        if (lk == NONE32) {  // wrap through the ends of the world line :224-241
            const uint32_t var = k ? v1 : v0;
            lk = side == SIDE_IN ? V.vlast[var] : V.vfirst[var];
        }
        prefetch_rec(V, lk);
        ipush((lk << 1) | (side ^ 1u) | flag);
    };
    for (;;) {
        if (lane == 0) {
            while (flen) {  // :62-80
                const uint32_t e = fpop();
                ST_COUNT(4);
                const uint32_t p0 = e >> 1, side0 = e & 1u;
                const uint4 l0 = ld_rec_lo(V, p0), h0 = ld_rec_hi(V, p0);
                if (h0.y != NONE32 && h0.z != NONE32) continue;
                // ---- expand_whole_cluster(p0, (0, side0), cnum) :193-271
                {
                    const uint32_t b0 = op_bond(l0.x);
                    const int kind0 = bond_kind(D, b0);
                    uint32_t a0, a1;
                    bond_vars(D, b0, kind0, a0, a1);
                    ilen = 0;
                    if (kind0 != KIND_SITE) {  // :205-211: every leg of the start op, boundaries set as the legs are popped
                        const uint32_t nv = kind0 == KIND_BOND ? 2u : 1u;
                        for (uint32_t k = 0; k < nv; k++) push_leg(l0, h0, k, SIDE_IN, a0, a1, 0x80000000u);
                        for (uint32_t k = 0; k < nv; k++) push_leg(l0, h0, k, SIDE_OUT, a0, a1, 0x80000000u);
                    } else {  // :212-215
                        push_leg(l0, h0, 0u, side0, a0, a1, 0x80000000u);
                    }
                    while (ilen) {
                        const uint32_t it = ipop();
                        ST_COUNT(0);
                        const uint32_t sq = it & 1u, lk = (it & 0x7FFFFFFFu) >> 1;
                        if (it & 0x80000000u) {  // set_boundary(p0, side, cnum) :218, :289-306
                            const uint32_t side = sq ^ 1u;
                            const uint32_t curb = REC_BND(V, p0, side);
                            if (curb == NONE32) REC_BND(V, p0, side) = cnum;
                            else if (curb != cnum) err |= DEV_ERR_INVARIANT;  // unreachable!() in the reference
                        }
                        const uint32_t q = lk >> 1, rq = lk & 1u;
                        const uint4 ql = ld_rec_lo(V, q), qh = ld_rec_hi(V, q);  // one sector: op, links, boundaries
                        const uint32_t bq = op_bond(ql.x);
                        if (bq >= E && bq < EN) {  // cluster edge :245-248
                            ST_COUNT(1);
                            const uint32_t mine = sq ? qh.z : qh.y, other = sq ? qh.y : qh.z;
                            if (mine == NONE32) REC_BND(V, q, sq) = cnum;
                            else if (mine != cnum) err |= DEV_ERR_INVARIANT;
                            if (other == NONE32) {  // not both sides set: the other side starts a new cluster later
                                if (flen >= V.fcap) { err |= DEV_ERR_STACK; } else fpush((q << 1) | (sq ^ 1u));
                                prefetch_rec(V, rec_link(ql, qh, 2u * (sq ^ 1u)));
                            }
                        } else {  // interior op :249-268
                            const uint32_t a = qh.y, bb = qh.z;
                            const bool ok = (a == NONE32 && bb == NONE32) || (a == cnum && bb == NONE32) || (a == NONE32 && bb == cnum);
                            if (ok) {
                                ST_COUNT(2);
                                REC_BND(V, q, 0) = cnum, REC_BND(V, q, 1) = cnum;
                                const int kq = bond_kind(D, bq);
                                uint32_t c0, c1;
                                bond_vars(D, bq, kq, c0, c1);
                                const uint32_t nvq = kq == KIND_BOND ? 2u : 1u;
                                if (ilen + 4 > V.icap) { err |= DEV_ERR_STACK; break; }
                                for (uint32_t k = 0; k < nvq; k++)
                                    if (!(k == rq && sq == SIDE_IN)) push_leg(ql, qh, k, SIDE_IN, c0, c1, 0u);
                                for (uint32_t k = 0; k < nvq; k++)
                                    if (!(k == rq && sq == SIDE_OUT)) push_leg(ql, qh, k, SIDE_OUT, c0, c1, 0u);
                            }
                        }
                    }
                }
                cnum++;
            }
        }
        __syncwarp();  // the boundaries lane 0 wrote are visible to the scanning lanes
        // :82-88 the smallest unmapped op (it never decreases): 32 slots per round, whole lines of the string
        uint32_t unmapped = NONE32;
        scan = __shfl_sync(0xFFFFFFFFu, scan, 0);
        for (uint32_t base = scan & ~31u; base <= last_p && unmapped == NONE32; base += 32) {
            const uint32_t p = base + (uint32_t)lane;
            bool un = false;
            if (p >= scan && p <= last_p && V.ops[p] != OP_EMPTY) {
                const uint4 h = ld_rec_hi(V, p);
                un = h.y == NONE32 && h.z == NONE32;
            }
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, un);
            if (m) unmapped = base + (uint32_t)__ffs(m) - 1u;
            else scan = base + 32;  // everything below is mapped for good
        }
        if (unmapped == NONE32) break;
        scan = unmapped;
        if (lane == 0) {
            fpush((unmapped << 1) | SIDE_OUT);  // :89-91
            fpush((unmapped << 1) | SIDE_IN);
        }
    }
    if (lane == 0) ST_COUNT_FLUSH;
    return __shfl_sync(0xFFFFFFFFu, cnum, 0);
}

// flip_each_cluster_rng (cluster.rs:36-172), whole warp; returns n_clusters
__device__ uint32_t cluster_strict(const SseDev &D, uint32_t r, const Rep &V, int lane, uint32_t *stk, long long &st_t) {
    const uint32_t n = D.n[r];
    if (n == 0) return 0;  // :46-48
    const uint32_t last_p = V.ends[1], cp = V.ends[2];
    // (the link pass initialised the boundaries of every op to None, cluster.rs:50-52)
    uint32_t ncl = 1;
    int err = 0;
    if (cp != NONE32) {
        ncl = label_strict(D, V, err, lane, stk);
    } else {  // :98-107 the whole thing is one cluster
        for (uint32_t p = lane; p <= last_p; p += 32)
            if (V.ops[p] != OP_EMPTY) REC_BND(V, p, 0) = 0, REC_BND(V, p, 1) = 0;
    }
    if (err) atomicOr(D.status, err);
    __syncwarp();
    ST_MARK(1);
    // flips: one gen_bool per cluster in id order (:111-137).  With a longitudinal field the
    // weight product is 0.0 for a cluster holding a longitudinal op (qmc_ising.rs:759-775):
    // gen_bool(0.0) still consumes its word and returns false.
    const uint32_t nwords = (ncl + 31) / 32;
    const uint64_t key = D.key[r], c0 = D.cursor[r];
    if (D.has_h) {
        for (uint32_t j = lane; j < nwords; j += 32) V.frozen[j] = 0;
        __syncwarp();
        for (uint32_t p = lane; p <= last_p; p += 32) {
            uint32_t w = V.ops[p];
            if (w != OP_EMPTY && op_bond(w) >= D.E + D.N) {
                uint32_t c = REC_BND(V, p, 0);
                atomicOr(&V.frozen[c >> 5], 1u << (c & 31));
            }
        }
        __syncwarp();
    }
    for (uint32_t base = 0; base < ncl; base += 32) {
        uint32_t k = base + lane;
        bool f = false;
        if (k < ncl) f = stream_word(key, c0 + k) < 0x8000000000000000ull;
        uint32_t word = __ballot_sync(0xFFFFFFFFu, f);
        if (lane == 0) V.bits[base >> 5] = D.has_h ? (word & ~V.frozen[base >> 5]) : word;
    }
    __syncwarp();
    ST_MARK(2);
    // apply (:139-167)
    for (uint32_t p = lane; p <= last_p; p += 32) {
        uint32_t w = V.ops[p];
        if (w == OP_EMPTY) continue;
        uint32_t ci = REC_BND(V, p, 0), co = REC_BND(V, p, 1);
        bool fi = (V.bits[ci >> 5] >> (ci & 31)) & 1u, fo = (V.bits[co >> 5] >> (co & 31)) & 1u;
        uint32_t b = op_bond(w);
        int kind = bond_kind(D, b);
        uint32_t mask = kind == KIND_BOND ? 3u : 1u;
        uint32_t in = op_in(w) ^ (fi ? mask : 0u), out = op_out(w) ^ (fo ? mask : 0u);
        if (fi) {
            uint32_t vv[2];
            bond_vars(D, b, kind, vv[0], vv[1]);
            for (int k = 0; k < (kind == KIND_BOND ? 2 : 1); k++) {
                if (REC_LINK(V, p, k) == NONE32) {  // first op on this world line
                    uint32_t v = vv[k], bit = 1u << (v & 31);
                    if ((in >> k) & 1u) atomicOr(&V.state[v >> 5], bit);
                    else atomicAnd(&V.state[v >> 5], ~bit);
                }
            }
        }
        if (fi || fo) V.ops[p] = make_op(b, in, out);
    }
    if (lane == 0) D.cursor[r] = c0 + ncl;
    __syncwarp();
    ST_MARK(3);
    return ncl;
}

// ------------------------------------------------------------------------------------------
// STRICT cluster step on WORLD-LINE ARRAYS (round 2, second layout; `layout` = 1 of k_sse_serial)
// ------------------------------------------------------------------------------------------
// The per-slot records above make every step of the walk a jump of ~1000 slots (the next op on a variable is that
// far away in imaginary time): one DRAM sector per leg.  Here the legs of a replica are stored per VARIABLE, in p
// order, so that following a world line is index +-1 in an array (the same 128-byte line most of the time) and only
// a two-variable op jumps -- to the entry of its other leg, whose two neighbours are again adjacent.
//   entry (uint4), one per leg:  x = p << 4 | first-on-its-line << 3 | leg k << 2 | kind (0 bond, 1 site, 2 longitudinal)
//                                y = bond op: entry of the other leg << 1 | that leg is first on its line
//                                z, w = cluster of the inputs / outputs (boundaries, cluster.rs:50-52); the two
//                                       entries of a bond op carry the same pair
//   variable v owns [head sentinel][its entries][tail sentinel]; a sentinel is kind 3 with y = the entry the line
//   continues at (periodic imaginary time: the wrap of cluster.rs:224-241), so the walk needs no per-variable table.
// ent[p] = entry of leg 0 of the op in slot p.  The order of pushes and pops -- hence the cluster numbering -- is the
// reference's, exactly as in label_strict above; only where a leg's neighbour is found differs.
#define WL_KIND(x) ((x) & 3u)
#define WL_SENT 3u

__device__ void links_wl(const SseDev &D, uint32_t r, const Rep &V, int lane, uint32_t *fill) {
    const uint32_t M = D.M[r];
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (uint32_t v = lane; v < D.N; v += 32) fill[v] = 0;
    __syncwarp();
    uint32_t first_p = NONE32, last_p = NONE32, first_site = NONE32;
    // (both passes stream the operator string; the next line is requested before the current one is worked on, otherwise
    // every step of every warp waits for one DRAM round trip)
    uint32_t wnext = (uint32_t)lane < M ? V.ops[lane] : OP_EMPTY;
    for (uint32_t base = 0; base < M; base += 32) {  // pass A: legs per variable
        const uint32_t p = base + lane;
        const uint32_t w = wnext;
        wnext = p + 32 < M ? V.ops[p + 32] : OP_EMPTY;
        int kind = -1;
        if (w != OP_EMPTY) {
            uint32_t v0, v1;
            kind = bond_kind(D, op_bond(w));
            bond_vars(D, op_bond(w), kind, v0, v1);
            atomicAdd(&fill[v0], 1u);
            if (kind == KIND_BOND) atomicAdd(&fill[v1], 1u);
        }
        const uint32_t hasm = __ballot_sync(0xFFFFFFFFu, kind >= 0), sitem = __ballot_sync(0xFFFFFFFFu, kind == KIND_SITE);
        if (hasm) {
            if (first_p == NONE32) first_p = base + (uint32_t)__ffs(hasm) - 1u;
            last_p = base + 31u - (uint32_t)__clz(hasm);
        }
        if (sitem && first_site == NONE32) first_site = base + (uint32_t)__ffs(sitem) - 1u;
    }
    __syncwarp();
    uint32_t running = 0;
    for (uint32_t base = 0; base < D.N; base += 32) {  // blocks of the variables, sentinels
        const uint32_t v = base + lane;
        const uint32_t c = v < D.N ? fill[v] : 0u, tot = v < D.N ? c + 2u : 0u;
        uint32_t incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += y;
        }
        const uint32_t s0 = running + incl - tot;
        if (v < D.N) {
            V.wl[s0] = make_uint4(WL_SENT, s0 + c, NONE32, NONE32);
            V.wl[s0 + c + 1] = make_uint4(WL_SENT, s0 + 1, NONE32, NONE32);
            fill[v] = (s0 + 1) | 0x80000000u;  // next entry of v; bit 31: nothing written yet
            V.vfirst[v] = c ? 0u : NONE32;     // free_spins: does_var_have_ops
        }
        running += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    __syncwarp();
    wnext = (uint32_t)lane < M ? V.ops[lane] : OP_EMPTY;
    for (uint32_t base = 0; base < M; base += 32) {  // pass B: entries in p order
        const uint32_t p = base + lane;
        const uint32_t w = wnext;
        wnext = p + 32 < M ? V.ops[p + 32] : OP_EMPTY;
        int kind = -1;
        uint32_t v0 = 0, v1 = 0;
        if (w != OP_EMPTY) {
            kind = bond_kind(D, op_bond(w));
            bond_vars(D, op_bond(w), kind, v0, v1);
        }
        if (!__ballot_sync(0xFFFFFFFFu, kind >= 0)) continue;
#pragma unroll
        for (int half = 0; half < 2; half++) {  // 16 slots = 32 legs, one lane per leg (lane = 2 slot + leg)
            const int sl = 16 * half + (lane >> 1);
            const uint32_t rel = (uint32_t)lane & 1u;
            const int sk = __shfl_sync(0xFFFFFFFFu, kind, sl);
            const uint32_t sv0 = __shfl_sync(0xFFFFFFFFu, v0, sl), sv1 = __shfl_sync(0xFFFFFFFFu, v1, sl);
            const bool legvalid = sk >= 0 && (rel == 0 || sk == KIND_BOND);
            const uint32_t sv = rel ? sv1 : sv0;
            const uint32_t legm = __ballot_sync(0xFFFFFFFFu, legvalid);
            const uint32_t m = __match_any_sync(0xFFFFFFFFu, legvalid ? sv : (0x80000000u | (uint32_t)lane)) & legm;
            uint32_t idx = 0, first = 0;
            if (legvalid) {
                const uint32_t f = fill[sv], rank = (uint32_t)__popc(m & lt_mask);
                idx = (f & 0x7FFFFFFFu) + rank;
                first = (f >> 31) & (rank == 0 ? 1u : 0u);
            }
            __syncwarp();
            if (legvalid && (m & ~lt_mask & ~(1u << lane)) == 0) fill[sv] = idx + 1;  // last leg of the half-step on this variable
            const uint32_t xidx = __shfl_xor_sync(0xFFFFFFFFu, idx, 1), xfirst = __shfl_xor_sync(0xFFFFFFFFu, first, 1);
            if (legvalid) {
                const uint32_t ps = base + (uint32_t)sl;
                V.wl[idx] = make_uint4((ps << 4) | (first << 3) | (rel << 2) | (uint32_t)sk, sk == KIND_BOND ? ((xidx << 1) | xfirst) : 0u, NONE32, NONE32);
                if (rel == 0) V.ent[ps] = idx;
            }
            __syncwarp();
        }
    }
    if (lane == 0) V.ends[0] = first_p, V.ends[1] = last_p, V.ends[2] = first_site;
    __syncwarp();
}

__device__ __forceinline__ uint32_t *wl_bnd(const Rep &V, uint32_t idx, uint32_t side) {
    return reinterpret_cast<uint32_t *>(V.wl + idx) + 2 + side;
}

// Both LIFO stacks keep their TOP in shared memory: a ring of STK_F / STK_IW entries holds the logical positions
// [base, len); when it is full its lower half is written to the global stack (in order), when it runs empty the half below
// is read back.  (The first version kept the BOTTOM 192 interior entries in shared memory: clusters of config #3 stack
// thousands of legs, and 43 % of the pops came from global memory.)
// OPT bit 0: every interior pop requests (prefetch.global.L1) the entry the leg BELOW the popped one lands on -- the next
// pop unless this one pushes; bit 1: a bond op requests the line of its other leg in L1 as soon as it is recognised.
// Tried and dropped: next-line L2 prefetch along the walk, and a prefetch of the frontier entries two or four places below
// the top at every frontier pop (227.5 against 229 ms per sweep: frontier entries are popped while their line is still in L1).
#define STK_IW 256
template <int OPT>
__device__ uint32_t label_strict_wl(const SseDev &D, const Rep &V, int &err, int lane, uint32_t *stk) {
    const uint32_t last_p = V.ends[1], cp = V.ends[2];
    uint32_t *const ist = stk, *const fst = stk + STK_IW;
    uint32_t flen = 0, fbase = 0, ilen = 0, ibase = 0;
    const uint32_t fcap = (uint32_t)V.fcap, icap = (uint32_t)V.icap;
    // (the capacity checks live in the spill paths: a ring that never spills cannot overflow the global stacks)
    auto fpush = [&](uint32_t x) {
        if (flen - fbase == STK_F) {
            if (fbase + STK_F / 2 > fcap) err |= DEV_ERR_STACK;
            else
                for (uint32_t j = 0; j < STK_F / 2; j++) V.frontier[fbase + j] = fst[(fbase + j) & (STK_F - 1)];
            fbase += STK_F / 2;
        }
        fst[flen & (STK_F - 1)] = x;
        flen++;
    };
    auto fpop = [&]() -> uint32_t {
        if (flen == fbase) {
            const uint32_t take = fbase < STK_F / 2 ? fbase : STK_F / 2;
            for (uint32_t j = 0; j < take; j++) fst[(fbase - take + j) & (STK_F - 1)] = V.frontier[fbase - take + j];
            fbase -= take;
        }
        flen--;
        return fst[flen & (STK_F - 1)];
    };
    auto ipush = [&](uint32_t x) {
        if (ilen - ibase == STK_IW) {
            if (ibase + STK_IW / 2 > icap) err |= DEV_ERR_STACK;
            else
                for (uint32_t j = 0; j < STK_IW / 2; j++) V.interior[ibase + j] = ist[(ibase + j) & (STK_IW - 1)];
            ibase += STK_IW / 2;
        }
        ist[ilen & (STK_IW - 1)] = x;
        ilen++;
    };
    auto ipop = [&]() -> uint32_t {
        if (ilen == ibase) {
            const uint32_t take = ibase < STK_IW / 2 ? ibase : STK_IW / 2;
            for (uint32_t j = 0; j < take; j++) ist[(ibase - take + j) & (STK_IW - 1)] = V.interior[ibase - take + j];
            ibase -= take;
        }
        ilen--;
        return ist[ilen & (STK_IW - 1)];
    };
    // an interior entry is where the leg LANDS: neighbouring entry << 1 | side it arrives at
    auto push_leg = [&](uint32_t idx, uint32_t side) { ipush(((side == SIDE_IN ? idx - 1u : idx + 1u) << 1) | (side ^ 1u)); };
    if (lane == 0) {
        const uint32_t e0 = V.ent[cp];
        fpush((e0 << 1) | SIDE_OUT);  // cluster.rs:57-59
        fpush((e0 << 1) | SIDE_IN);
    }
    uint32_t cnum = 0, scan = 0;
    ST_COUNT_DECL;
    for (;;) {
        if (lane == 0) {
            while (flen) {  // :62-80
                const uint32_t fe = fpop();
                ST_COUNT(4);
                const uint32_t i0 = fe >> 1, side0 = fe & 1u;
                const uint4 E0 = V.wl[i0];
                if (E0.z != NONE32 && E0.w != NONE32) continue;
                const uint32_t kind0 = WL_KIND(E0.x), x0 = E0.y >> 1;
                ilen = 0, ibase = 0;
                // the legs of the start op (:205-215: in0, in1, out0, out1 of a non-edge op, the one leg of an edge) are popped
                // last-pushed first and everything a leg reaches is popped before the next one, so they are taken one by one
                // here instead of travelling through the stack with a flag; set_boundary(p0, side, cnum) (:218, :289-306)
                // happens when the leg is popped, as in the reference
                const uint32_t nstart = kind0 == KIND_SITE ? 1u : (kind0 == KIND_BOND ? 4u : 2u);
#pragma unroll 1
                for (uint32_t k = nstart; k-- > 0;) {
                    const uint32_t sside = kind0 == KIND_SITE ? side0 : (kind0 == KIND_BOND ? (k >> 1) : k);
                    const uint32_t sidx = (kind0 == KIND_BOND && (k & 1u)) ? x0 : i0;
                    {
                        uint32_t *b = wl_bnd(V, i0, sside);
                        const uint32_t curb = *b;
                        if (curb == NONE32) {
                            *b = cnum;
                            if (kind0 == KIND_BOND) *wl_bnd(V, x0, sside) = cnum;
                        } else if (curb != cnum) err |= DEV_ERR_INVARIANT;
                    }
                    push_leg(sidx, sside);
                    while (ilen) {
                        const uint32_t it = ipop();
                        ST_COUNT(0);
                        const uint32_t sq = it & 1u;
                        uint32_t q = it >> 1;
                        if ((OPT & 1) && ilen != ibase) asm volatile("prefetch.global.L1 [%0];" ::"l"(V.wl + (ist[(ilen - 1u) & (STK_IW - 1)] >> 1)));
                        uint4 *pe = V.wl + q;
                        uint4 e = *pe;
                        uint32_t kq = WL_KIND(e.x);
                        if (kq == WL_SENT) {  // end of the world line: continue at the other end :224-241
                            q = e.y;
                            pe = V.wl + q;
                            e = *pe;
                            kq = WL_KIND(e.x);
                            ST_COUNT(3);
                        }
                        if ((OPT & 2) && kq == KIND_BOND) asm volatile("prefetch.global.L1 [%0];" ::"l"(V.wl + (e.y >> 1)));
                        uint32_t *const bq = reinterpret_cast<uint32_t *>(pe) + 2;
                        if (kq == KIND_SITE) {  // cluster edge :245-248
                            ST_COUNT(1);
                            const uint32_t mine = sq ? e.w : e.z, other = sq ? e.z : e.w;
                            if (mine == NONE32) bq[sq] = cnum;
                            else if (mine != cnum) err |= DEV_ERR_INVARIANT;
                            if (other == NONE32) fpush((q << 1) | (sq ^ 1u));
                        } else {  // interior op :249-268
                            const uint32_t a = e.z, bb = e.w;
                            const bool ok = (a == NONE32 && bb == NONE32) || (a == cnum && bb == NONE32) || (a == NONE32 && bb == cnum);
                            if (ok) {
                                ST_COUNT(2);
                                const uint32_t xq = e.y >> 1, kme = (e.x >> 2) & 1u;
                                *reinterpret_cast<uint2 *>(bq) = make_uint2(cnum, cnum);
                                if (kq == KIND_BOND) {
                                    *reinterpret_cast<uint2 *>(wl_bnd(V, xq, 0)) = make_uint2(cnum, cnum);
                                    const uint32_t i_k0 = kme ? xq : q, i_k1 = kme ? q : xq;
                                    if (!(kme == 0 && sq == SIDE_IN)) push_leg(i_k0, SIDE_IN);
                                    if (!(kme == 1 && sq == SIDE_IN)) push_leg(i_k1, SIDE_IN);
                                    if (!(kme == 0 && sq == SIDE_OUT)) push_leg(i_k0, SIDE_OUT);
                                    if (!(kme == 1 && sq == SIDE_OUT)) push_leg(i_k1, SIDE_OUT);
                                } else {
                                    if (sq != SIDE_IN) push_leg(q, SIDE_IN);
                                    if (sq != SIDE_OUT) push_leg(q, SIDE_OUT);
                                }
                            }
                        }
                    }
                }
                cnum++;
            }
        }
        __syncwarp();
        // :82-88 the smallest unmapped op (it never decreases), 32 slots per round
        uint32_t unmapped = NONE32;
        scan = __shfl_sync(0xFFFFFFFFu, scan, 0);
        for (uint32_t base = scan & ~31u; base <= last_p && unmapped == NONE32; base += 32) {
            const uint32_t p = base + (uint32_t)lane;
            bool un = false;
            if (p >= scan && p <= last_p && V.ops[p] != OP_EMPTY) {
                const uint2 h = *reinterpret_cast<const uint2 *>(wl_bnd(V, V.ent[p], 0));
                un = h.x == NONE32 && h.y == NONE32;
            }
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, un);
            if (m) unmapped = base + (uint32_t)__ffs(m) - 1u;
            else scan = base + 32;
        }
        if (unmapped == NONE32) break;
        scan = unmapped;
        if (lane == 0) {
            const uint32_t eu = V.ent[unmapped];
            fpush((eu << 1) | SIDE_OUT);  // :89-91
            fpush((eu << 1) | SIDE_IN);
        }
    }
    if (lane == 0) ST_COUNT_FLUSH;
    return __shfl_sync(0xFFFFFFFFu, cnum, 0);
}

// flip_each_cluster_rng (cluster.rs:36-172) on the world-line layout, whole warp; returns n_clusters
__device__ uint32_t cluster_strict_wl(const SseDev &D, uint32_t r, const Rep &V, int lane, uint32_t *stk, long long &st_t, int pf) {
    const uint32_t n = D.n[r];
    if (n == 0) return 0;
    const uint32_t last_p = V.ends[1], cp = V.ends[2];
    uint32_t ncl = 1;
    int err = 0;
    if (cp != NONE32) {
        switch (pf) {  // strict_layout bits 2 and 8
            case 1: ncl = label_strict_wl<1>(D, V, err, lane, stk); break;
            case 2: ncl = label_strict_wl<2>(D, V, err, lane, stk); break;
            case 3: ncl = label_strict_wl<3>(D, V, err, lane, stk); break;
            default: ncl = label_strict_wl<0>(D, V, err, lane, stk); break;
        }
    } else {  // :98-107 the whole thing is one cluster
        for (uint32_t p = lane; p <= last_p; p += 32)
            if (V.ops[p] != OP_EMPTY) *reinterpret_cast<uint2 *>(wl_bnd(V, V.ent[p], 0)) = make_uint2(0u, 0u);
    }
    if (err) atomicOr(D.status, err);
    __syncwarp();
    ST_MARK(1);
    const uint32_t nwords = (ncl + 31) / 32;
    const uint64_t key = D.key[r], c0 = D.cursor[r];
    if (D.has_h) {
        for (uint32_t j = lane; j < nwords; j += 32) V.frozen[j] = 0;
        __syncwarp();
        for (uint32_t p = lane; p <= last_p; p += 32) {
            const uint32_t w = V.ops[p];
            if (w != OP_EMPTY && op_bond(w) >= D.E + D.N) {
                const uint32_t c = *wl_bnd(V, V.ent[p], 0);
                atomicOr(&V.frozen[c >> 5], 1u << (c & 31));
            }
        }
        __syncwarp();
    }
    for (uint32_t base = 0; base < ncl; base += 32) {  // one gen_bool per cluster in id order (:111-137)
        const uint32_t k = base + lane;
        bool f = false;
        if (k < ncl) f = stream_word(key, c0 + k) < 0x8000000000000000ull;
        const uint32_t word = __ballot_sync(0xFFFFFFFFu, f);
        if (lane == 0) V.bits[base >> 5] = D.has_h ? (word & ~V.frozen[base >> 5]) : word;
    }
    __syncwarp();
    ST_MARK(2);
    uint32_t wnext = (uint32_t)lane <= last_p ? V.ops[lane] : OP_EMPTY;
    uint32_t enext = (uint32_t)lane <= last_p ? V.ent[lane] : 0u;  // (the entry of an empty slot is never looked at)
    for (uint32_t p = lane; p <= last_p; p += 32) {  // apply (:139-167); the next line of ops and entry indices is requested first
        const uint32_t w = wnext, ei = enext;
        wnext = p + 32 <= last_p ? V.ops[p + 32] : OP_EMPTY;
        enext = p + 32 <= last_p ? V.ent[p + 32] : 0u;
        if (w == OP_EMPTY) continue;
        const uint4 e = V.wl[ei];
        const uint32_t ci = e.z, co = e.w;
        const bool fi = (V.bits[ci >> 5] >> (ci & 31)) & 1u, fo = (V.bits[co >> 5] >> (co & 31)) & 1u;
        if (!(fi || fo)) continue;
        const uint32_t b = op_bond(w), kind = WL_KIND(e.x);
        const uint32_t mask = kind == KIND_BOND ? 3u : 1u;
        const uint32_t in = op_in(w) ^ (fi ? mask : 0u), out = op_out(w) ^ (fo ? mask : 0u);
        if (fi && (((e.x >> 3) & 1u) | (e.y & 1u))) {  // a leg with no previous op on its world line: state at p = 0 changes
            uint32_t vv[2];
            bond_vars(D, b, (int)kind, vv[0], vv[1]);
            const uint32_t firsts = ((e.x >> 3) & 1u) | ((e.y & 1u) << 1);
            for (int k = 0; k < (kind == KIND_BOND ? 2 : 1); k++) {
                if ((firsts >> k) & 1u) {
                    const uint32_t v = vv[k], bit = 1u << (v & 31);
                    if ((in >> k) & 1u) atomicOr(&V.state[v >> 5], bit);
                    else atomicAnd(&V.state[v >> 5], ~bit);
                }
            }
        }
        V.ops[p] = make_op(b, in, out);
    }
    if (lane == 0) D.cursor[r] = c0 + ncl;
    __syncwarp();
    ST_MARK(3);
    return ncl;
}

// ------------------------------------------------------------------------------------------
// FAST-order cluster step, serial restatement (oracle.c cluster_update_fast), lane 0
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t uf_find(uint32_t *uf, uint32_t x) {
    while (uf[x] != x) {
        uf[x] = uf[uf[x]];
        x = uf[x];
    }
    return x;
}
__device__ __forceinline__ bool fast_flip_bit(uint64_t key, uint64_t c0, uint32_t root) {
    Philox4 o = philox4x32_10(root >> 7, (uint32_t)c0, (uint32_t)(c0 >> 32), QMCB_TAG_CLUS, (uint32_t)key, (uint32_t)(key >> 32));
    uint32_t sel = (root >> 5) & 3u;
    uint32_t x = sel == 0 ? o.x : (sel == 1 ? o.y : (sel == 2 ? o.z : o.w));
    return (x >> (root & 31)) & 1u;
}

__device__ uint32_t cluster_fast_serial(const SseDev &D, uint32_t r, const Rep &V, int lane, bool maxroot) {
    const uint32_t n = D.n[r];
    if (n == 0) return 0;
    uint32_t ncl = 0;
    if (lane == 0) {
        const uint32_t N = D.N, E = D.E, EN = D.E + D.N, last_p = V.ends[1];
        uint32_t *uf = V.parent;
        for (uint32_t v = 0; v < N; v++) uf[v] = v, V.cur[v] = v;
        uint32_t nsite = 0;
        for (uint32_t p = 0; p <= last_p; p++) {
            uint32_t w = V.ops[p];
            if (w == OP_EMPTY) continue;
            uint32_t b = op_bond(w);
            if (b >= E && b < EN) {
                uint32_t id = N + nsite++;
                uf[id] = id;
                V.cur[b - E] = id;
            } else if (b < E) {
                uint32_t x = uf_find(uf, V.cur[__ldg(D.va + b)]), y = uf_find(uf, V.cur[__ldg(D.vb + b)]);
                if (x != y) {
                    if ((x < y) != maxroot) uf[y] = x;
                    else uf[x] = y;
                }
            }
        }
        for (uint32_t v = 0; v < N; v++) {
            uint32_t x = uf_find(uf, v), y = uf_find(uf, V.cur[v]);
            if (x != y) {
                if ((x < y) != maxroot) uf[y] = x;
                else uf[x] = y;
            }
        }
        const uint32_t nseg = N + nsite, nw = (nseg + 31) / 32;
        if (nsite == 0)
            for (uint32_t x = 0; x < nseg; x++) uf[x] = maxroot ? nseg - 1 : 0;
        // frozen roots, used roots
        for (uint32_t j = 0; j < nw; j++) V.frozen[j] = 0, V.bits[j] = 0;
        for (uint32_t v = 0; v < N; v++) V.cur[v] = v;
        nsite = 0;
        for (uint32_t p = 0; p <= last_p; p++) {
            uint32_t w = V.ops[p];
            if (w == OP_EMPTY) continue;
            uint32_t b = op_bond(w);
            uint32_t ri, ro;
            if (b >= E && b < EN) {
                ri = uf_find(uf, V.cur[b - E]);
                uint32_t id = N + nsite++;
                V.cur[b - E] = id;
                ro = uf_find(uf, id);
            } else {
                uint32_t v0 = b < E ? __ldg(D.va + b) : b - EN;
                ri = ro = uf_find(uf, V.cur[v0]);
                if (b >= EN) V.frozen[ri >> 5] |= 1u << (ri & 31);
            }
            V.bits[ri >> 5] |= 1u << (ri & 31), V.bits[ro >> 5] |= 1u << (ro & 31);
        }
        for (uint32_t j = 0; j < nw; j++) ncl += __popc(V.bits[j]);
        // apply
        const uint64_t key = D.key[r], c0 = D.cursor[r];
        for (uint32_t v = 0; v < N; v++) V.cur[v] = v;
        nsite = 0;
        for (uint32_t p = 0; p <= last_p; p++) {
            uint32_t w = V.ops[p];
            if (w == OP_EMPTY) continue;
            uint32_t b = op_bond(w);
            uint32_t ri, ro, mask = b < E ? 3u : 1u;
            if (b >= E && b < EN) {
                ri = uf_find(uf, V.cur[b - E]);
                uint32_t id = N + nsite++;
                V.cur[b - E] = id;
                ro = uf_find(uf, id);
            } else {
                uint32_t v0 = b < E ? __ldg(D.va + b) : b - EN;
                ri = ro = uf_find(uf, V.cur[v0]);
            }
            bool fi = !((V.frozen[ri >> 5] >> (ri & 31)) & 1u) && fast_flip_bit(key, c0, ri);
            bool fo = ri == ro ? fi : (!((V.frozen[ro >> 5] >> (ro & 31)) & 1u) && fast_flip_bit(key, c0, ro));
            if (fi || fo) V.ops[p] = make_op(b, op_in(w) ^ (fi ? mask : 0u), op_out(w) ^ (fo ? mask : 0u));
        }
        // state: the segment of variable v crossing p = 0 has id v
        for (uint32_t v = 0; v < N; v++) {
            if (V.vfirst[v] == NONE32) continue;
            uint32_t rt = uf_find(uf, v);
            if (!((V.frozen[rt >> 5] >> (rt & 31)) & 1u) && fast_flip_bit(key, c0, rt)) V.state[v >> 5] ^= 1u << (v & 31);
        }
        D.cursor[r] = c0 + 1;
    }
    ncl = __shfl_sync(0xFFFFFFFFu, ncl, 0);
    __syncwarp();
    return ncl;
}

// free spins: qmc_ising.rs:780-784, one gen_bool(0.5) per op-less variable in ascending order
__device__ void free_spins(const SseDev &D, uint32_t r, const Rep &V, int lane) {
    const uint64_t key = D.key[r];
    uint64_t cur = D.cursor[r];
    for (uint32_t base = 0; base < D.N; base += 32) {
        uint32_t v = base + lane;
        bool fr = v < D.N && V.vfirst[v] == NONE32;
        uint32_t m = __ballot_sync(0xFFFFFFFFu, fr);
        bool bit = false;
        if (fr) bit = stream_word(key, cur + __popc(m & ((1u << lane) - 1u))) < 0x8000000000000000ull;
        uint32_t setm = __ballot_sync(0xFFFFFFFFu, bit);
        if (lane == 0 && m) V.state[base >> 5] = (V.state[base >> 5] & ~m) | setm;
        cur += __popc(m);
    }
    __syncwarp();
    if (lane == 0) D.cursor[r] = cur;
}

// phases: bit0 diagonal update, bit1 cluster update + free spins, bit2 cutoff growth,
// bit3 bookkeeping of a full timestep (done counter, estimators, sampling),
// bit4 run only the single step that takes the replica from target - 1 to target
// 7 blocks of 4 warps per SM (72 registers): 4096 replicas are resident in one wave, which is what this
// latency-bound kernel needs (at 80 registers it ran in two waves: 467 ms instead of 290 ms per sweep on config #3)
// ------------------------------------------------------------------------------------------
// Directed-loop update (directed_loop.rs:103-171 make_loop_update_with_rng with initial_n = None, as Qmc::loop_update
// qmc_runner.rs:205-220 calls it; :183-211 apply_loop_update; :214-301 loop_body), lane 0, on the per-slot records
// (links of links_serial / links_parallel).  A leg is (relative variable, side); the weight of leaving through a leg is
// the matrix element of the op with the entrance leg and that leg flipped (adjust_states, qmc_types.rs:28-37).
// ------------------------------------------------------------------------------------------
__device__ __noinline__ void loop_update_serial(const SseDev &D, uint32_t r, const Rep &V) {
    const uint32_t n = D.n[r];
    if (n == 0) return;  // :139
    const uint64_t key = D.key[r];
    uint64_t cur = D.cursor[r];
    auto gen_range = [&](uint64_t range) -> uint64_t {  // rand 0.8 UniformInt<usize>::sample_single
        const uint64_t zone = (range << __clzll((long long)range)) - 1ull;
        uint64_t hi, lo;
        do {
            const uint64_t v = stream_word(key, cur++);
            hi = __umul64hi(v, range), lo = v * range;
        } while (lo > zone);
        return hi;
    };
    const uint64_t initial_n = gen_range(n);  // :140-142
    uint32_t p0 = V.ends[0];                  // get_nth_p :76-87
    for (uint64_t seen = 0;; p0++) {
        if (V.ops[p0] == OP_EMPTY) continue;
        if (seen == initial_n) break;
        seen++;
    }
    const uint32_t w0 = REC_OP(V, p0);
    const uint32_t v0 = (uint32_t)gen_range(bond_kind(D, op_bond(w0)) == KIND_BOND ? 2u : 1u);  // :147
    const uint32_t s0 = ((int32_t)(uint32_t)(stream_word(key, cur++) >> 32) < 0) ? SIDE_IN : SIDE_OUT;  // rng.gen::<bool>() :148-152
    uint32_t sel = p0, ev = v0, es = s0;
    int err = 0;
    for (;;) {  // :195-210
        uint32_t w = REC_OP(V, sel);
        const uint32_t b = op_bond(w);
        const int kind = bond_kind(D, b);
        const uint32_t nv = kind == KIND_BOND ? 2u : 1u, nlegs = 2u * nv;
        const uint32_t in_e = op_in(w) ^ (es == SIDE_IN ? 1u << ev : 0u), out_e = op_out(w) ^ (es == SIDE_OUT ? 1u << ev : 0u);
        double wt[4], total = 0.0;
        for (uint32_t k = 0; k < nlegs; k++) {  // inputs legs, then outputs legs :231-240
            const uint32_t kv = k < nv ? k : k - nv;
            const uint32_t in = in_e ^ (k < nv ? 1u << kv : 0u), out = out_e ^ (k < nv ? 0u : 1u << kv);
            // Interaction::at: outputs more significant than inputs, first variable most significant
            const size_t idx = kind == KIND_BOND
                                   ? 16 * (size_t)b + ((((out & 1u) << 1) | (out >> 1)) << 2 | ((in & 1u) << 1) | (in >> 1))
                                   : 16 * (size_t)D.E + 4 * (size_t)(b - D.E) + ((out & 1u) << 1 | (in & 1u));
            wt[k] = __ldg(D.g_full + idx);
            total = total + wt[k];  // :242
        }
        if (!(0.0 < total) || isinf(total)) {  // gen_range(0. ..total) panics on an empty or unbounded range
            err = DEV_ERR_PROB;
            break;
        }
        double c;
        do c = unit_f64(stream_word(key, cur++)) * total + 0.0;  // gen_range(0. ..total) :243
        while (!(c < total));
        uint32_t ex = NONE32;
        for (uint32_t k = 0; k < nlegs; k++) {  // try_fold :244-253
            if (c < wt[k]) {
                ex = k;
                break;
            }
            c = c - wt[k];
        }
        if (ex == NONE32) {  // unwrap_err() on Ok: rounding left the choice beyond the last leg
            err = DEV_ERR_PROB;
            break;
        }
        const uint32_t xv = ex < nv ? ex : ex - nv, xs = ex < nv ? SIDE_IN : SIDE_OUT;
        const uint32_t in_n = in_e ^ (xs == SIDE_IN ? 1u << xv : 0u), out_n = out_e ^ (xs == SIDE_OUT ? 1u << xv : 0u);
        w = make_op(b, in_n, out_n);  // :256-259
        REC_OP(V, sel) = w, V.ops[sel] = w;
        if (sel == p0 && xv == v0 && xs == s0) break;  // :266-267
        uint32_t vv[2];
        bond_vars(D, b, kind, vv[0], vv[1]);
        const uint32_t var = vv[xv];
        uint32_t leg = REC_LINK(V, sel, (xs == SIDE_OUT ? 2u : 0u) + xv);
        if (leg == NONE32) {  // the world line closes through p = 0: the state there changes :274-289
            const uint32_t bit = ((xs == SIDE_OUT ? out_n : in_n) >> xv) & 1u;
            V.state[var >> 5] = (V.state[var >> 5] & ~(1u << (var & 31))) | (bit << (var & 31));
            leg = xs == SIDE_OUT ? V.vfirst[var] : V.vlast[var];
        }
        const uint32_t q = leg >> 1, rq = leg & 1u, ns = xs ^ 1u;  // :291
        if (q == p0 && rq == v0 && ns == s0) break;               // :293-294
        sel = q, ev = rq, es = ns;
    }
    D.cursor[r] = cur;
    if (err) atomicOr(D.status, err);
}

__global__ void __launch_bounds__(128, 7) k_sse_serial(SseDev D, int mode, uint64_t target, uint32_t phases,
                                                    uint64_t sample_freq, uint64_t sample_origin,
                                                    uint8_t *samples, uint64_t samples_per_rep, int par_links, int layout, int split) {
    // per warp: [STK_I + STK_F] stack tops of the STRICT walk, then [N] per-variable table of the link pass when par_links.
    // split = 1: build the links and stop (this launch carries the table); split = 2: the links exist, do the rest with
    // the stacks only -- 7 blocks x 16 KB of tables would otherwise take the shared memory that the walk wants as L1.
    extern __shared__ uint32_t smem_warp[];
    uint32_t *const my_smem = smem_warp + (size_t)(threadIdx.x >> 5) * (STK_WORDS + (par_links && split != 2 ? D.N : 0u));
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= D.R) return;
    const Rep V = rep_view(D, r);
    uint64_t done = D.done[r];
    const uint64_t nsteps = (phases & 16u) ? (done + 1 == target ? 1 : 0) : ((phases & 8u) ? (target > done ? target - done : 0) : 1);
    for (uint64_t s = 0; s < nsteps; s++) {
        if (D.M[r] > D.cap) {  // cannot run this sweep: host must grow the arrays first
            if (lane == 0) atomicOr(D.status, DEV_ERR_CAPACITY);
            break;
        }
        if (phases & 1u) {
            if (lane == 0) {
                if (mode == 2) diag_counter_serial(D, r, V);
                else if (D.hb_cum) diag_heatbath_serial(D, r, V);
                else diag_serial(D, r, V);
            }
            __syncwarp();
        }
        if (phases & 32u) {  // Qmc::loop_update on its own (qmc_runner.rs:205-220): links, one loop update
            for (uint32_t v = lane; v < D.N; v += 32) V.vfirst[v] = NONE32, V.vlast[v] = NONE32;
            __syncwarp();
            if (lane == 0) {
                links_serial(D, r, V, true);
                loop_update_serial(D, r, V);
            }
            __syncwarp();
            break;
        }
        if ((phases & 2u) && D.loop_path) {  // a model with loop updates: all-serial step on the per-slot records
            for (uint32_t v = lane; v < D.N; v += 32) V.vfirst[v] = NONE32, V.vlast[v] = NONE32;
            __syncwarp();
            if (lane == 0) {
                links_serial(D, r, V, true);
                if (D.loop_updates) loop_update_serial(D, r, V);  // qmc_runner.rs:366-368: between the diagonal and the cluster update
            }
            __syncwarp();
            long long st_t = 0;
            uint32_t ncl = D.no_site ? 0u : cluster_strict(D, r, V, lane, my_smem, st_t);
            if (lane == 0) D.ncl[r] = ncl;
            free_spins(D, r, V, lane);
        } else if (phases & 2u) {
            const bool wl = mode == 0 && par_links && (layout & 1);
            long long st_t = clock64();
            if (split == 2) {
            } else if (wl) {
                links_wl(D, r, V, lane, my_smem + STK_WORDS);
            } else if (mode == 0 && par_links) {
                links_parallel(D, r, V, lane, my_smem + STK_WORDS);
            } else {
                for (uint32_t v = lane; v < D.N; v += 32) V.vfirst[v] = NONE32, V.vlast[v] = NONE32;
                __syncwarp();
                if (lane == 0) links_serial(D, r, V, mode == 0);
                __syncwarp();
            }
            ST_MARK(0);
            if (split == 1) break;
            uint32_t ncl = wl ? cluster_strict_wl(D, r, V, lane, my_smem, st_t, ((layout >> 1) & 1) | ((layout >> 2) & 2))
                              : (mode == 0 ? cluster_strict(D, r, V, lane, my_smem, st_t) : cluster_fast_serial(D, r, V, lane, false));
            if (lane == 0) D.ncl[r] = ncl;
            free_spins(D, r, V, lane);
            ST_MARK(4);
        }
        if (phases & 4u) {
            if (lane == 0) {  // qmc_ising.rs:786
                uint32_t n = D.n[r], grown = n + n / 2;
                if (grown > D.M[r]) D.M[r] = grown;
            }
            __syncwarp();
        }
        if (phases & 8u) {
            done++;
            const uint64_t idx = done - sample_origin;  // 1-based sweep index inside this call
            if (lane == 0) D.vupd[r] += D.n[r];
            if (idx % sample_freq == 0) {  // qmc_stepper.rs:150-158
                if (lane == 0) D.sum_n[r] += D.n[r];
                if (samples) {
                    uint8_t *dst = samples + ((size_t)r * samples_per_rep + (idx / sample_freq - 1)) * D.N;
                    for (uint32_t v = lane; v < D.N; v += 32) dst[v] = (uint8_t)state_bit(V.state, v);
                }
            }
            if (lane == 0) D.done[r] = done;
            __syncwarp();
        }
    }
}

// OpContainer::verify (op_container.rs:137-159) + non-zero weights (qmc_ising.rs:829-860), lane 0
__global__ void k_sse_verify(SseDev D, uint32_t r, int *ok_out, uint32_t *scratch) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const Rep V = rep_view(D, r);
    const Ham Hm = ham_view<true>(D, r);
    uint32_t *roll = scratch;
    for (uint32_t j = 0; j < D.Nw; j++) roll[j] = V.state[j];
    int ok = 1;
    uint32_t n = 0;
    for (uint64_t p = 0; p < D.cap; p++) {
        uint32_t w = V.ops[p];
        if (w == OP_EMPTY) continue;
        n++;
        if (p >= D.M[r]) ok = 0;
        uint32_t b = op_bond(w);
        if (b >= D.E + 2 * D.N || (w >> 28)) { ok = 0; break; }
        int kind = bond_kind(D, b);
        if (kind == KIND_LONG && !D.has_h) ok = 0;
        uint32_t v0, v1;
        bond_vars(D, b, kind, v0, v1);
        uint32_t in = op_in(w), out = op_out(w);
        if (D.g_full) {  // generic interactions: the matrix element of the op as it stands (Interaction::at) must not vanish
            const size_t idx = kind == KIND_BOND
                                   ? 16 * (size_t)b + ((((out & 1u) << 1) | (out >> 1)) << 2 | ((in & 1u) << 1) | (in >> 1))
                                   : 16 * (size_t)D.E + 4 * (size_t)(b - D.E) + ((out & 1u) << 1 | (in & 1u));
            if (kind == KIND_LONG || !(fabs(D.g_full[idx]) > 2.220446049250313e-16)) ok = 0;
        } else if (kind != KIND_SITE) {
            if (in != out) ok = 0;  // off-diagonal two-site / longitudinal ops have zero weight
            else if (!(fabs(bond_weight(Hm, b, kind, in & 1u, (in >> 1) & 1u)) > 2.220446049250313e-16)) ok = 0;
        } else if (!(fabs(Hm.gamma) > 2.220446049250313e-16)) ok = 0;
        if (state_bit(roll, v0) != (in & 1u)) ok = 0;
        roll[v0 >> 5] = (roll[v0 >> 5] & ~(1u << (v0 & 31))) | ((out & 1u) << (v0 & 31));
        if (kind == KIND_BOND) {
            if (state_bit(roll, v1) != ((in >> 1) & 1u)) ok = 0;
            roll[v1 >> 5] = (roll[v1 >> 5] & ~(1u << (v1 & 31))) | (((out >> 1) & 1u) << (v1 & 31));
        }
    }
    for (uint32_t j = 0; j < D.Nw; j++)
        if (roll[j] != V.state[j]) ok = 0;
    if (n != D.n[r]) ok = 0;
    *ok_out = ok;
}

// imaginary_time_fold (qmc_ising.rs:815-821; OpContainer::itime_fold fast_ops.rs:1296-1315) with the magnetisation
// fold, one warp per replica: the fold sees the state before the op at p; only off-diagonal (transverse) ops change
// it, so m_p = m_0 + exclusive prefix sum of the flips.  sums[r] = {sum m, sum m^2, sum |m|} over p in [0, M).
__global__ void k_sse_itime_magnetization(SseDev D, long long *sums) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= D.R) return;
    const uint32_t *ops = D.ops + (size_t)r * D.cap, *st = D.state + (size_t)r * D.Nw;
    const uint32_t M = D.M[r];
    int up = 0;
    for (uint32_t j = lane; j < D.Nw; j += 32) up += __popc(st[j]);
    up = (int)__reduce_add_sync(0xFFFFFFFFu, (unsigned)up);
    long long m = 2ll * up - (long long)D.N, s1 = 0, s2 = 0, s3 = 0;
    for (uint32_t base = 0; base < M; base += 32) {
        const uint32_t p = base + lane;
        const uint32_t w = (p < M && p < D.cap) ? ops[p] : OP_EMPTY;
        int f = 0;  // change of m made by my op
        if (w != OP_EMPTY && !op_is_diag(w)) {
            f = 2 * ((int)(op_out(w) & 1u) - (int)(op_in(w) & 1u));
            if (bond_kind(D, op_bond(w)) == KIND_BOND) f += 2 * ((int)((op_out(w) >> 1) & 1u) - (int)((op_in(w) >> 1) & 1u));
        }
        int incl = f;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += y;
        }
        const long long mine = m + (incl - f);
        if (p < M) s1 += mine, s2 += mine * mine, s3 += mine < 0 ? -mine : mine;
        m += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    for (int d = 16; d; d >>= 1) {
        s1 += __shfl_xor_sync(0xFFFFFFFFu, s1, d), s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, d), s3 += __shfl_xor_sync(0xFFFFFFFFu, s3, d);
    }
    if (lane == 0) sums[3 * (size_t)r] = s1, sums[3 * (size_t)r + 1] = s2, sums[3 * (size_t)r + 2] = s3;
}
// propagated state of replica r before slot p_at, one warp
__global__ void k_sse_itime_state(SseDev D, uint32_t r, uint64_t p_at, uint32_t *out) {
    const int lane = threadIdx.x;
    const uint32_t *ops = D.ops + (size_t)r * D.cap, *st = D.state + (size_t)r * D.Nw;
    for (uint32_t j = lane; j < D.Nw; j += 32) out[j] = st[j];
    __syncwarp();
    const uint64_t lim = p_at < D.cap ? p_at : D.cap;
    for (uint64_t base = 0; base < lim; base += 32) {
        const uint64_t p = base + lane;
        const uint32_t w = p < lim ? ops[p] : OP_EMPTY;
        if (w != OP_EMPTY && !op_is_diag(w)) {  // transverse ops act on one variable; XOR commutes within the step
            const int kind = bond_kind(D, op_bond(w));
            uint32_t v0, v1;
            bond_vars(D, op_bond(w), kind, v0, v1);
            if ((op_in(w) ^ op_out(w)) & 1u) atomicXor(&out[v0 >> 5], 1u << (v0 & 31));
            if (kind == KIND_BOND && ((op_in(w) ^ op_out(w)) & 2u)) atomicXor(&out[v1 >> 5], 1u << (v1 & 31));
        }
    }
}

// per-bond counts of replica r (fast_ops.rs:1281-1294)
__global__ void k_sse_bond_counts(SseDev D, uint32_t r, unsigned long long *counts) {
    const uint32_t *ops = D.ops + (size_t)r * D.cap;
    for (uint64_t p = blockIdx.x * blockDim.x + threadIdx.x; p < D.M[r]; p += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t w = ops[p];
        if (w != OP_EMPTY) atomicAdd(&counts[op_bond(w)], 1ull);
    }
}

// recount n of replica r after qmcb_load_ops
__global__ void k_sse_recount(SseDev D, uint32_t r) {
    __shared__ uint32_t tot;
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    const uint32_t *ops = D.ops + (size_t)r * D.cap;
    uint32_t c = 0;
    for (uint64_t p = threadIdx.x; p < D.cap; p += blockDim.x) c += ops[p] != OP_EMPTY;
    atomicAdd(&tot, c);
    __syncthreads();
    if (threadIdx.x == 0) D.n[r] = tot;
}

// stream-drawn initial states: make_random_spin_state (classical/graph.rs:451-453)
__global__ void k_sse_init_state(SseDev D) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= D.R) return;
    uint32_t *st = D.state + (size_t)r * D.Nw;
    const uint64_t key = D.key[r];
    for (uint32_t base = 0; base < D.N; base += 32) {
        uint32_t v = base + lane;
        bool bit = v < D.N && (stream_word(key, v) >> 63);  // sign bit of next_u32 = top bit of the word
        uint32_t m = __ballot_sync(0xFFFFFFFFu, bit);
        if (lane == 0) st[base >> 5] = m;
    }
    if (lane == 0) D.cursor[r] = D.N;
}

// returns 1 when the STRICT cluster step ran on the world-line layout (layout = 1 and the per-variable table fits shared memory)
int launch_sse_serial(const SseDev &D, int mode, uint64_t target, uint32_t phases, uint64_t sample_freq,
                      uint64_t sample_origin, uint8_t *samples, uint64_t samples_per_rep, int layout, cudaStream_t st) {
    const int threads = 128;
    const uint32_t blocks = (uint32_t)(((uint64_t)D.R * 32 + threads - 1) / threads);
    const size_t stk = (size_t)(threads / 32) * STK_WORDS * sizeof(uint32_t);
    size_t smem = stk + (size_t)(threads / 32) * D.N * sizeof(uint32_t);
    const int par = mode == 0 && smem <= 160 * 1024;
    if (!par) smem = stk;
    if (D.cap >= (1ull << 27) || !D.ent) layout = 0;  // entry indices carry a side bit and a flag bit in the walk's stacks
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_sse_serial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const bool wl = par && (layout & 1) && (phases & 2u);
    if (wl && (layout & 4) && (phases & 16u)) {  // one sweep per launch: links in their own launch, the walk with the stacks only
        k_sse_serial<<<blocks, threads, smem, st>>>(D, mode, target, 2u | 16u, sample_freq, sample_origin, samples, samples_per_rep, par, layout, 1);
        k_sse_serial<<<blocks, threads, stk, st>>>(D, mode, target, phases, sample_freq, sample_origin, samples, samples_per_rep, par, layout, 2);
        return 2;
    }
    k_sse_serial<<<blocks, threads, smem, st>>>(D, mode, target, phases, sample_freq, sample_origin, samples, samples_per_rep, par, layout, 0);
    return wl;
}
void launch_sse_verify(const SseDev &D, uint32_t r, int *ok_dev, uint32_t *scratch_dev, cudaStream_t st) {
    k_sse_verify<<<1, 32, 0, st>>>(D, r, ok_dev, scratch_dev);
}
void launch_sse_bond_counts(const SseDev &D, uint32_t r, unsigned long long *counts_dev, cudaStream_t st) {
    k_sse_bond_counts<<<64, 256, 0, st>>>(D, r, counts_dev);
}
void launch_sse_itime_magnetization(const SseDev &D, long long *sums_dev, cudaStream_t st) {
    const int threads = 128;
    k_sse_itime_magnetization<<<(uint32_t)(((uint64_t)D.R * 32 + threads - 1) / threads), threads, 0, st>>>(D, sums_dev);
}
void launch_sse_itime_state(const SseDev &D, uint32_t r, uint64_t p_at, uint32_t *out_dev, cudaStream_t st) {
    k_sse_itime_state<<<1, 32, 0, st>>>(D, r, p_at, out_dev);
}
// ---- variable autocorrelation (autocorrelations.rs:48-61, :99-133) ---------------------------------------------
// pass 1: samples [R][T][N] bytes -> per (replica, variable) the time series as bits, stored twice back to back
// (2 T bits + padding) so that a circular shift is a plain window, and the number of ones
__global__ void k_ac_pack(const uint8_t *samples, uint32_t R, uint32_t N, uint32_t T, uint32_t Tw, uint32_t *bits, uint32_t *ones) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (v >= N) return;
    const uint8_t *src = samples + (size_t)r * T * N + v;
    uint32_t *dst = bits + ((size_t)r * N + v) * (2 * Tw + 1);
    for (uint32_t w = 0; w < 2 * Tw + 1; w++) dst[w] = 0;
    uint32_t cnt = 0;
    for (uint32_t t = 0; t < T; t++) {
        if (src[(size_t)t * N]) {
            cnt++;
            dst[t >> 5] |= 1u << (t & 31);
            const uint32_t u = t + T;
            dst[u >> 5] |= 1u << (u & 31);
        }
    }
    ones[(size_t)r * N + v] = cnt;
}
// pass 2: out[r][tau] = (1/N) sum_v (C_v[tau] - T m_v^2) / (T (1 - m_v^2)), C_v[tau] = sum_t s_t s_{t+tau mod T} = T - 2 mismatches.
// One block per (tau, replica); the sum over the variables is a fixed-shape tree, so the result does not depend on timing.
__global__ void __launch_bounds__(256) k_ac_corr(uint32_t N, uint32_t T, uint32_t Tw, const uint32_t *bits, const uint32_t *ones, double *out) {
    __shared__ double part[256];
    const uint32_t tau = blockIdx.x, r = blockIdx.y;
    double acc = 0.0;
    for (uint32_t v = threadIdx.x; v < N; v += blockDim.x) {
        const uint32_t *x = bits + ((size_t)r * N + v) * (2 * Tw + 1);
        uint32_t mism = 0;
        for (uint32_t w = 0; w < Tw; w++) {
            const uint32_t a = x[w];
            const uint32_t pos = 32 * w + tau, wi = pos >> 5, sh = pos & 31;
            const uint32_t b = sh ? __funnelshift_r(x[wi], x[wi + 1], sh) : x[wi];
            uint32_t d = a ^ b;
            if (w == Tw - 1 && (T & 31u)) d &= (1u << (T & 31u)) - 1u;
            mism += __popc(d);
        }
        const double Td = (double)T, m = (2.0 * (double)ones[(size_t)r * N + v] - Td) / Td;
        const double c = Td - 2.0 * (double)mism;
        acc += (c - Td * m * m) / (Td * (1.0 - m * m));  // 0/0 = NaN for a variable that never changed, as the reference
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t s = 128; s; s >>= 1) {
        if (threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[(size_t)r * T + tau] = part[0] / (double)N;
}
// series of spin products (autocorrelations.rs:53-71): derived[r][t][k] = 1 iff an even number of the spins of product k is
// true (the sign convention drops out of the normalised autocorrelation)
__global__ void k_ac_products(const uint8_t *samples, uint32_t N, uint32_t T, uint32_t K, const uint32_t *offsets, const uint32_t *vars, uint8_t *derived) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y, r = blockIdx.z;
    if (k >= K) return;
    const uint8_t *s = samples + ((size_t)r * T + t) * N;
    uint32_t par = 0;
    for (uint32_t i = offsets[k]; i < offsets[k + 1]; i++) par ^= s[vars[i]] & 1u;
    derived[((size_t)r * T + t) * K + k] = (uint8_t)(par ^ 1u);
}
void launch_spin_products(const uint8_t *samples, uint32_t R, uint32_t N, uint32_t T, uint32_t K, const uint32_t *offsets, const uint32_t *vars,
                          uint8_t *derived, cudaStream_t st) {
    for (uint32_t t0 = 0; t0 < T; t0 += 65535) {  // gridDim.y limit
        const uint32_t nt = std::min(T - t0, 65535u);
        k_ac_products<<<dim3((K + 127) / 128, nt, R), 128, 0, st>>>(samples + (size_t)t0 * N, N, T, K, offsets, vars, derived + (size_t)t0 * K);
    }
}
void launch_autocorrelation(const uint8_t *samples, uint32_t R, uint32_t N, uint32_t T, uint32_t *bits, uint32_t *ones, double *out,
                            cudaStream_t st) {
    const uint32_t Tw = (T + 31) / 32;
    k_ac_pack<<<dim3((N + 127) / 128, R), 128, 0, st>>>(samples, R, N, T, Tw, bits, ones);
    k_ac_corr<<<dim3(T, R), 256, 0, st>>>(N, T, Tw, bits, ones, out);
}
void launch_sse_recount(const SseDev &D, uint32_t r, cudaStream_t st) { k_sse_recount<<<1, 256, 0, st>>>(D, r); }
void launch_sse_init_state(const SseDev &D, cudaStream_t st) {
    const int threads = 128;
    const uint32_t blocks = (uint32_t)(((uint64_t)D.R * 32 + threads - 1) / threads);
    k_sse_init_state<<<blocks, threads, 0, st>>>(D);
}
