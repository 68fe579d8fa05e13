// sse_fast.cu -- warp-parallel SSE sweep, FAST cluster order.  One warp per replica; all state
// that is touched at random (spin bits, per-variable representatives, the RNG window) lives in
// shared memory, the operator string streams through in coalesced 128-byte lines.
//
// Per sweep and replica:
//  P1  diagonal update (diagonal.rs:142-191), 32 slots per step.  The rule is sequential in p
//      (live n, data-dependent draw count), so each step is solved as a fixed point: every lane
//      evaluates its slot for a guessed (stream cursor, n), an exclusive warp scan of the draws /
//      insertions gives the next guess, and the loop ends when no lane's input changed -- lane i is
//      exact after at most i+1 rounds, typically 2-4.  The 64 stream words a step can consume are
//      generated once per step (one Philox4x32-10 call per lane) into shared memory.  The result is
//      bit-identical to the sequential loop.
//      Fused into the same pass: world-line segments and their union-find (global memory, lock-free
//      min-root hooking), so links are never materialised.
//  P2  resolve one flip bit per segment in increasing id order (parent id < child id).
//  P3  apply the flips to the operator words (second streaming pass) and to the spins.
// Contract of the FAST order: oracle.c cluster_update_fast / DESIGN.md.
#include <algorithm>
#include <map>
#include <mutex>

#include "sse.cuh"
#include "sse_warp.cuh"

#define T_NONE 0
#define T_EMPTY 1
#define T_DIAG 2
#define T_OFFD 3
// event counters (qmcb_get_debug_counters) and phase timers are compiled in with -DQMCB_PHASE_TIMERS only: even a
// never-taken `if (D.dbg)` costs three instructions and a branch at eight places of every step
#ifdef QMCB_PHASE_TIMERS
#define DBG(i, v) do { if (D.dbg && lane == 0) atomicAdd(D.dbg + (i), (unsigned long long)(v)); } while (0)
#else
#define DBG(i, v) do { } while (0)
#endif
// phase timers, compiled in with -DQMCB_PHASE_TIMERS (tools/prof_sse.py prints them): TICK(k) adds the cycles
// since the previous TICK to dbg[32 + k]
#ifdef QMCB_PHASE_TIMERS
#define TICK(k) do { if (D.dbg) { const long long t_ = clock64(); if (lane == 0) atomicAdd(D.dbg + 32 + (k), (unsigned long long)(t_ - tick_)); tick_ = t_; } } while (0)
#else
#define TICK(k) do { } while (0)
#endif

// shared memory of one replica (= one block of one warp).  The fixed-size tables sit at constant offsets from the
// start of the block's shared memory, so that their addresses are immediates; the lattice-sized ones follow.
struct WarpSmem {
    unsigned long long *win;  // [64] stream words of this step
    uint32_t *fl;   // [32] variable flipped by the off-diagonal op of lane j (or NONE32)
    uint32_t *opw;  // [64] op word an empty slot would insert from window word y
    unsigned char *G;  // [80] cursor after an empty slot that starts reading at window position x; G[64] = 255
    unsigned char *wk, *wl;  // [32] walk: draws of diagonal ops since the previous EMPTY lane, lane of the k-th EMPTY lane
    unsigned short *wxg;     // [32] walk: start cursor | cursor after << 8 of the k-th EMPTY lane
    uint32_t *line;          // [32] next line of the operator string (asynchronous copy, see fetch_line)
    uint32_t *st;   // [Nw] spin bits at the current p
    uint32_t *tb;   // [Nw] variable has at least one op
    uint32_t *cd;   // [Nw] P1: variable is flipped inside this step; P3: flip decision of the segment open on each variable
    uint32_t *sb;   // [Nw] low-occupancy build only: variable is cut by a site op of this step (cleared at the end of the step)
    uint32_t *rep;  // [N]  P1: a member of the set of the segment currently open on each variable
};
#define SM_WIN 0
#define SM_FL 512
#define SM_OPW 640
#define SM_G 896
#define SM_WK 976
#define SM_WL 1008
#define SM_WXG 1040
#define SM_LINE 1104  // [32] the next 128-byte line of the operator string, filled by cp.async
#define SM_VAR 1232   // st, tb, cd, rep

// NB the 72-register build runs 7 blocks per SM and its shared memory (7 x (4 x 5712 + 1024) B at N = 1024) just fits
// the 164 KiB carve-out: 128 B more per block and the driver moves to 196 KiB, L1 shrinks from 92 to 60 KiB and the
// sweep takes 6 % longer (measured by padding; 228 KiB: +14 %).  Shared memory is not free here even when it fits.
__host__ __device__ inline size_t warp_smem_bytes(uint32_t N, uint32_t Nw, bool lowocc) {
    return (SM_VAR + ((size_t)(lowocc ? 4 : 3) * Nw + N + (lowocc ? 72 : 0)) * 4 + 15) / 16 * 16;  // + ring[2][32] + 8 exchange words
}

#ifndef QMCB_WPB
#define QMCB_WPB 4  // warps (= replicas) per block: one per scheduler of the SM (1 or 2 per block measured 8-12% slower)
#endif
// PK: 0 plain edge tables, 1 packed table in shared memory, 2 packed table through L1.
// PIPE: two warps per replica.  The sweep of a long operator string is one dependent chain, and when few replicas
// are resident (tempering ladders, big lattices) nothing hides it.  Role 0 runs the diagonal update of step k while
// role 1 does the segment bookkeeping and unions of step k - 1 (final op words handed over through a two-slot ring in
// shared memory, one named barrier per step); role 1 then does the closure and P2, role 0 P3 and the rest.
template <bool HAS_H, int MINB, bool HB, bool MH, int PK, bool PIPE>
__global__ void __launch_bounds__(PIPE ? 64 * QMCB_WPB : 32 * QMCB_WPB, PIPE ? 2 : 4 * MINB / QMCB_WPB) k_sse_fast(SseDev D, uint64_t target, uint32_t phases, uint64_t sample_freq,
                                                  uint64_t sample_origin, uint8_t *samples, uint64_t samples_per_rep, uint32_t smem_stride,
                                                  uint32_t epk_off) {
    extern __shared__ __align__(16) unsigned char smem_all[];
    // block-shared copy of the packed edge table (variables and coupling code of every bond): the look-ups on the
    // step's dependent chain become shared-memory loads instead of L1/L2 loads.  Compiled in (PK) only for the
    // low-occupancy builds: with 28 warps per SM the table shrinks L1 below what spills and the other tables need
    // (measured +8 % time on config #3, -20 % on L = 64 ladders).
    if (PK == 1) {
        uint32_t *dst = (uint32_t *)(smem_all + epk_off);
        for (uint32_t i = threadIdx.x; i < D.E; i += blockDim.x) dst[i] = __ldg(D.epk + i);
        __syncthreads();
    }
    const uint32_t *const epk_s = PK == 2 ? D.epk : (const uint32_t *)(smem_all + epk_off);
#if QMCB_WPB == 1
    unsigned char *const smem_raw = smem_all;
    const int lane = threadIdx.x;
    const uint32_t r = blockIdx.x;
#else
    // the warp index goes through a warp reduction so that the compiler knows it (and everything derived from it: the
    // shared-memory base, the replica index and the replica's global pointers) is warp-uniform and keeps it in the
    // uniform register file instead of rematerialising it from %tid under register pressure
    const uint32_t wraw = __reduce_max_sync(0xFFFFFFFFu, threadIdx.x >> 5);
    const uint32_t wib = PIPE ? wraw >> 1 : wraw;
    unsigned char *const smem_raw = smem_all + wib * smem_stride;
    const int lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * QMCB_WPB + wib;
    const bool roleA = !PIPE || (wraw & 1u) == 0, roleB = !PIPE || (wraw & 1u) == 1;  // diagonal update + P3 / segments + P2
#define PAIR_SYNC()                                                                        \
    do {                                                                                   \
        if (PIPE) asm volatile("bar.sync %0, 64;" ::"r"(wib + 1u) : "memory");             \
    } while (0)
#endif
    if (r >= D.R) return;
    const uint32_t N = D.N, Nw = D.Nw;
    WarpSmem S;
    S.win = (unsigned long long *)(smem_raw + SM_WIN), S.fl = (uint32_t *)(smem_raw + SM_FL), S.opw = (uint32_t *)(smem_raw + SM_OPW);
    S.G = smem_raw + SM_G, S.wk = smem_raw + SM_WK, S.wl = smem_raw + SM_WL, S.wxg = (unsigned short *)(smem_raw + SM_WXG);
    S.line = (uint32_t *)(smem_raw + SM_LINE);
    S.st = (uint32_t *)(smem_raw + SM_VAR), S.tb = S.st + Nw, S.cd = S.st + 2 * Nw, S.sb = S.st + 3 * Nw, S.rep = S.st + (MINB <= 4 ? 4 : 3) * Nw;
    if (lane < 16) S.G[64 + lane] = 255;  // positions past the window: exhausted (absorbing state of the walk)
    uint32_t *const ring = S.rep + N;                                     // PIPE: [2][32] final op words of a step
    volatile uint32_t *const xchg = (volatile uint32_t *)(ring + 64);  // PIPE: n, cursor, cluster count between the roles
    uint32_t *ops = D.ops + (size_t)r * D.cap;
    uint32_t *gstate = D.state + (size_t)r * Nw;
    uint32_t *P = D.parent + (size_t)r * (N + D.cap + 1);
    const size_t bstride = (size_t)(D.cap / 32 + 2 + N / 32);
    uint32_t *decb = D.bits + (size_t)r * bstride;
    uint32_t *frz = D.frozen + (size_t)r * bstride;
    uint32_t lt_mask;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
    const uint64_t key = D.key[r];
    const Ham Hm = ham_view<MH>(D, r);
    // variables of bond b; returns the packed table word (coupling code in the top 4 bits) when the table is in use
    auto evars = [&](uint32_t b, int kind, uint32_t &v0, uint32_t &v1) -> uint32_t {
        if (PK && kind == KIND_BOND) {
            const uint32_t e = PK == 2 ? __ldg(epk_s + b) : epk_s[b];
            v0 = e & 0x3FFFu, v1 = (e >> 14) & 0x3FFFu;
            return e;
        }
        bond_vars(D, b, kind, v0, v1);
        return 0u;
    };
    auto ewt = [&](uint32_t e, uint32_t b, int kind, uint32_t s0, uint32_t s1) -> double {
        if (PK && kind == KIND_BOND) {
            const double j = D.jdict[e >> 28];
            return fabs(j) + (s0 == s1 ? -j : j);
        }
        return bweight<HAS_H>(Hm, b, kind, s0, s1);
    };
    const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
    const uint64_t range = D.Nb;
    const uint64_t zone = D.zone;
#ifdef QMCB_PHASE_TIMERS
    long long tick_ = clock64();
#endif
    uint64_t done = D.done[r];
    // bit 4 (16): run exactly the step that takes this replica from target - 1 to target (split launches)
    const uint64_t nsteps = (phases & 16u) ? (done + 1 == target ? 1 : 0) : ((phases & 8u) ? (target > done ? target - done : 0) : 1);
    int err = 0;

    for (uint64_t sw = 0; sw < nsteps; sw++) {
        uint32_t M = D.M[r];
        if (M > D.cap) {
            if (lane == 0) atomicOr(D.status, DEV_ERR_CAPACITY);
            break;
        }
        uint32_t n = D.n[r];
        uint64_t cur = D.cursor[r];
        const double bn = D.beta[r] * (double)D.Nb;
        if (roleA)
            for (uint32_t j = lane; j < Nw; j += 32) S.st[j] = gstate[j], S.cd[j] = 0;
        if (roleB)
            for (uint32_t j = lane; j < Nw; j += 32) S.tb[j] = 0;
        if (MINB <= 4 && roleB)
            for (uint32_t j = lane; j < Nw; j += 32) S.sb[j] = 0;
        const bool do_diag = phases & 1u, do_clus = phases & 2u;
        if (do_clus && roleB) {
            for (uint32_t v = lane; v < N; v += 32) S.rep[v] = v, st_cg(P + v, v);
            if (HAS_H)
                for (uint32_t j = lane; j < (uint32_t)bstride; j += 32) st_cg(frz + j, 0u);
        }
        __syncwarp();
        PAIR_SYNC();
        uint32_t nsite = 0;
        bool anylong = false;

        // =========================== P1: diagonal update + unions ===========================
        if (M && roleA) fetch_line(S.line, ops, lane);  // software prefetch of the next 128-byte line
        const uint32_t nit = (M + 31) / 32;
        for (uint32_t it = 0; it < nit + (PIPE ? 1u : 0u); it++) {
          uint32_t base = it * 32, p = base + lane;
          bool valid = p < M;
          uint32_t neww = OP_EMPTY;
          if (roleA && it < nit) {
            uint32_t w = take_line(S.line, lane);
            if (!valid) w = OP_EMPTY;
            if (base + 32 < M) fetch_line(S.line, ops + base + 32, lane);
            DBG(0, 1);
            TICK(0);  // between steps
            int type = !valid ? T_NONE : (w == OP_EMPTY ? T_EMPTY : (op_is_diag(w) ? T_DIAG : T_OFFD));
            neww = w;
            // variables / stored bits of an existing op
            uint32_t ov0 = 0, ov1 = 0, oe = 0;
            int okind = KIND_BOND;
            if (type >= T_DIAG) {
                okind = bkind<HAS_H>(D, op_bond(w));
                oe = evars(op_bond(w), okind, ov0, ov1);
            }
            if (do_diag) {
                S.fl[lane] = type == T_OFFD ? ov0 : NONE32;
                if (type == T_OFFD) atomicOr(&S.cd[ov0 >> 5], 1u << (ov0 & 31));  // hazard bitmap of this step
                const uint32_t fmask_lt = __ballot_sync(FULL, type == T_OFFD) & lt_mask;
                const uint32_t emask = __ballot_sync(FULL, type == T_EMPTY), dmask = __ballot_sync(FULL, type == T_DIAG);
                __syncwarp();
                uint32_t rem = emask | dmask;
                if constexpr (HB) {
                    // ---- heat-bath rule (heatbath.rs:149-209).  Draws per slot: EMPTY 1 (+2 when an insertion
                    // is attempted), DIAG 1.  The attempt / removal thresholds depend on the live n only through
                    // den, monotonically, so a window word is classified once against the thresholds at both ends
                    // of the n interval the round can span; the rare word in between ends the resolved prefix
                    // (EMPTY lanes: it moves the cursor) or is settled in lane order with the exact n (DIAG lanes).
                    // A round that cannot resolve its first lane evaluates that lane literally.
                    const double total = Hm.hb_total, bt = D.beta[r] * total;
                    auto spin_here = [&](uint32_t v) -> uint32_t {  // spin of v as seen by my slot
                        uint32_t sv = state_bit(S.st, v);
                        if (state_bit(S.cd, v))
                            for (uint32_t m = fmask_lt; m; m &= m - 1) sv ^= (S.fl[__ffs(m) - 1] == v);
                        return sv;
                    };
                    auto try_insert = [&](uint64_t wp, uint64_t wc, uint32_t &nw) -> bool {  // heatbath.rs:166-187
                        const double pd = unit_f64(wp), c = unit_f64(wc) * total;
                        if (!(pd < 1.0) || !(c < total)) { err |= DEV_ERR_INVARIANT; return false; }
                        const uint32_t b = hb_index_for_cumulative(Hm.hb_cum, D.Nb, c);
                        if (b >= D.Nb) { err |= DEV_ERR_INVARIANT; return false; }
                        const int kind = bkind<HAS_H>(D, b);
                        uint32_t v0, v1;
                        const uint32_t e = evars(b, kind, v0, v1);
                        const uint32_t s0 = spin_here(v0), s1 = kind == KIND_BOND ? spin_here(v1) : 0u;
                        if (!(pd * __ldg(Hm.hb_maxw + b) < ewt(e, b, kind, s0, s1))) return false;
                        const uint32_t bitsv = s0 | (s1 << 1);
                        nw = make_op(b, bitsv, bitsv);
                        return true;
                    };
                    while (rem) {
                        const uint64_t wbase = cur & ~1ull;
                        {
                            const uint64_t blk = (wbase >> 1) + (uint64_t)lane;
                            Philox4 o = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), 0u, 0u, k0, k1);
                            S.win[2 * lane] = ((unsigned long long)o.y << 32) | o.x;
                            S.win[2 * lane + 1] = ((unsigned long long)o.w << 32) | o.z;
                        }
                        __syncwarp();
                        const bool inrem = (rem >> lane) & 1u;
                        const uint32_t remE = emask & rem, remD = dmask & rem;
                        const uint32_t head = (uint32_t)__ffs(rem) - 1u;
                        const uint32_t x0 = (uint32_t)(cur - wbase);
                        const uint32_t nlo = n - (uint32_t)__popc(remD), nhi = n + (uint32_t)__popc(remE);
                        uint32_t dc = 0, okm = 0;
                        int dn = 0;
                        if (nhi < M && bt > 0.0) {
                            // thresholds at the ends of the n interval (IEEE operations are monotone)
                            const uint64_t te_lo = bool_threshold(bt / ((double)(M - nlo) + bt)), te_hi = bool_threshold(bt / ((double)(M - nhi) + bt));
                            const double dl = (double)(M - nhi + 1), dh = (double)(M - nlo + 1);
                            const uint64_t td_lo = bool_threshold(dl / (dl + bt)), td_hi = bool_threshold(dh / (dh + bt));
                            const uint64_t w0 = S.win[lane], w1 = S.win[lane + 32];
                            const uint64_t ATT = ((uint64_t)__ballot_sync(FULL, w1 < te_lo) << 32) | __ballot_sync(FULL, w0 < te_lo);
                            const uint64_t NOA = ((uint64_t)__ballot_sync(FULL, w1 >= te_hi) << 32) | __ballot_sync(FULL, w0 >= te_hi);
                            const uint32_t predE = remE & lt_mask;
                            const int le = predE ? 31 - __clz(predE) : -1;
                            const uint32_t after_le = le < 0 ? FULL : ~((2u << le) - 1u);
                            const uint32_t dcount = (uint32_t)__popc(remD & lt_mask & after_le);  // DIAG draws since the last EMPTY lane
                            const uint32_t kidx = (uint32_t)__popc(predE);
                            if (inrem && type == T_EMPTY) S.wk[kidx] = (unsigned char)dcount, S.wl[kidx] = (unsigned char)lane;
                            __syncwarp();
                            const uint32_t nE = (uint32_t)__popc(remE);
                            uint32_t x = x0, k = 0, stop = 32u;
                            for (; k < nE; k++) {  // uniform walk over the EMPTY lanes
                                const uint32_t xs = x + S.wk[k];
                                if (xs + 2u >= 64u || !(((ATT | NOA) >> xs) & 1ull)) {
                                    stop = S.wl[k];
                                    break;
                                }
                                x = xs + 1u + 2u * (uint32_t)((ATT >> xs) & 1ull);
                                S.wxg[k] = (unsigned short)(xs | (x << 8));
                            }
                            __syncwarp();
                            const uint32_t kstop = k;
                            uint32_t myx = 0;
                            bool res = false;
                            if (inrem && type == T_EMPTY) res = kidx < kstop, myx = res ? (uint32_t)S.wxg[kidx] & 255u : 0u;
                            else if (inrem) {
                                res = kidx <= kstop && (uint32_t)lane < stop;
                                myx = (kidx == 0 ? x0 : (res ? (uint32_t)S.wxg[kidx - 1] >> 8 : 0u)) + dcount;
                            }
                            const uint32_t ovD = __ballot_sync(FULL, res && type == T_DIAG && myx >= 64u);
                            if (ovD) stop = min(stop, (uint32_t)__ffs(ovD) - 1u);
                            const uint32_t resolved = stop >= 32u ? rem : (rem & ((1u << stop) - 1u));
                            bool unres = false;
                            if ((resolved >> lane) & 1u) {
                                if (type == T_EMPTY) {
                                    dc = ((uint32_t)S.wxg[kidx] >> 8) - myx;
                                    neww = OP_EMPTY;
                                    if (dc == 3u && try_insert(S.win[myx + 1], S.win[myx + 2], neww)) dn = 1;
                                } else {
                                    const uint64_t v = S.win[myx];
                                    dc = 1;
                                    if (v < td_lo) neww = OP_EMPTY, dn = -1;
                                    else if (v >= td_hi) neww = w, dn = 0;
                                    else unres = true;
                                }
                            }
                            for (uint32_t um = __ballot_sync(FULL, unres); um; um = __ballot_sync(FULL, unres)) {
                                const uint32_t u = (uint32_t)__ffs(um) - 1u;
                                const uint32_t pre = __reduce_add_sync(FULL, (((resolved >> lane) & 1u) && (uint32_t)lane < u) ? (uint32_t)(dn + 1) : 0u);
                                if ((uint32_t)lane == u) {
                                    const uint32_t ni = n + pre - (uint32_t)__popc(resolved & lt_mask);
                                    const double num = (double)(M - ni + 1);
                                    const bool remove = S.win[myx] < bool_threshold(num / (num + bt));
                                    neww = remove ? OP_EMPTY : w, dn = remove ? -1 : 0, unres = false;
                                }
                            }
                            okm = resolved;
                        }
                        if (okm == 0) {  // literal evaluation of the first lane at the exact (cursor, n)
                            DBG(3, 1);
                            if ((uint32_t)lane == head) {
                                uint32_t x = x0;
                                if (type == T_EMPTY) {
                                    const double pr = bt / ((double)(M - n) + bt);
                                    bool attempt = false;
                                    if (pr == 1.0) attempt = true;
                                    else if (!(pr >= 0.0 && pr < 1.0)) err |= DEV_ERR_PROB;
                                    else attempt = S.win[x++] < bool_threshold(pr);
                                    neww = OP_EMPTY;
                                    if (attempt) {
                                        if (try_insert(S.win[x], S.win[x + 1], neww)) dn = 1;
                                        x += 2;
                                    }
                                } else {
                                    const double num = (double)(M - n + 1), pr = num / (num + bt);
                                    bool remove = false;
                                    if (pr == 1.0) remove = true;
                                    else if (!(pr >= 0.0 && pr < 1.0)) err |= DEV_ERR_PROB;
                                    else remove = S.win[x++] < bool_threshold(pr);
                                    neww = remove ? OP_EMPTY : w, dn = remove ? -1 : 0;
                                }
                                dc = x - x0;
                            }
                            okm = 1u << head;
                        }
                        const bool mine = (okm >> lane) & 1u;
                        const uint32_t tot = __reduce_add_sync(FULL, mine ? dc : 0u);
                        const int totn = (int)__reduce_add_sync(FULL, mine ? (uint32_t)(dn + 1) : 0u) - __popc(okm);
                        if (inrem && !mine) neww = w;
                        cur += tot;
                        n = (uint32_t)((int)n + totn);
                        rem &= ~okm;
                        __syncwarp();
                    }
                }
                double dnum = 0.0;  // num of an existing diagonal op does not depend on (cursor, n)
                if (!HB && type == T_DIAG) dnum = bn * ewt(oe, op_bond(w), okind, op_in(w) & 1u, (op_in(w) >> 1) & 1u);
                bool try_fast = true;
                TICK(1);  // load, classify, decode of existing ops, flip list
                while (!HB && rem) {
                    const uint64_t wbase = cur & ~1ull;
                    {  // stream words [wbase, wbase + 64)
                        const uint64_t blk = (wbase >> 1) + (uint64_t)lane;
                        Philox4 o = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), 0u, 0u, k0, k1);
                        S.win[2 * lane] = ((unsigned long long)o.y << 32) | o.x;
                        S.win[2 * lane + 1] = ((unsigned long long)o.w << 32) | o.z;
                    }
                    __syncwarp();
                    TICK(2);  // Philox window
                    const bool inrem = (rem >> lane) & 1u;
                    uint32_t dc = 0, okm;
                    int dn = 0;
                    if (try_fast) {
                        DBG(1, 1);
                        // ---- table walk.  Which bond a word encodes, whether rand's zone test rejects it and
                        // the weight it gets are properties of the WORD (the spins it looks at only change at
                        // the few off-diagonal ops of this step: "hazards"), so every window position is decoded
                        // once, in parallel, into: ACC (word passes gen_range's zone), EX (a second word is read),
                        // OKI (the op is inserted), HZ (needs the exact per-lane evaluation).  "Same for every
                        // admissible n" is decided on the den interval spanned by the whole step (the rules are
                        // monotone in den).  G[x] = cursor after an empty slot that starts reading at x.  The
                        // cursors of all lanes then follow from a short uniform walk x <- G[x] / x + 1.
                        const uint32_t remE = emask & rem, remD = dmask & rem;
                        const uint32_t cntE = (uint32_t)__popc(remE & lt_mask), cntD = (uint32_t)__popc(remD & lt_mask);
                        const double dlo = (double)(M - (n + cntE)), dhi = (double)(M - (n - cntD));          // this lane
                        const double dloA = (double)(M - (n + (uint32_t)__popc(remE))), dhiA = (double)(M - (n - (uint32_t)__popc(remD)));
                        // DIAG lanes: draw or not, and the thresholds at both ends of their den interval
                        bool ambdc = false, draws = false;
                        uint64_t t_lo = 0, t_hi = 0;
                        if (inrem && type == T_DIAG) {
                            const bool nd_lo = dlo + 1.0 >= dnum, nd_hi = dhi + 1.0 >= dnum;  // den >= num: removed, no draw
                            if (!nd_lo) {
                                // conservative bounds on the threshold over the den interval: one reciprocal, a
                                // guard band far above its rounding error; a word inside the band (p ~ 2^-43) is
                                // settled below with the exact division at the exact n
                                draws = true, ambdc = nd_hi;
                                const double rc = 1.0 / dnum;
                                const uint64_t a_lo = bool_threshold((dlo + 1.0) * rc), a_hi = bool_threshold((dhi + 1.0) * rc);
                                t_lo = a_lo > (1ull << 20) ? a_lo - (1ull << 20) : 0ull;
                                t_hi = a_hi < ~0ull - (1ull << 20) ? a_hi + (1ull << 20) : ~0ull;
                            }
                        }
                        const uint32_t drawD = __ballot_sync(FULL, draws), ambD = __ballot_sync(FULL, ambdc);
                        // decode window positions y = lane and lane + 32
                        uint32_t accb[2], exb[2], okb[2], hzb[2];
#pragma unroll
                        for (int hh = 0; hh < 2; hh++) {
                            const uint32_t y = (uint32_t)lane + 32u * hh;
                            const uint64_t v = S.win[y];
                            const uint64_t hi = __umul64hi(v, range), lo = v * range;
                            const bool acc = lo <= zone;
                            bool ex = false, ok = false, hz = false;
                            uint32_t opw = OP_EMPTY;
                            if (acc) {
                                const uint32_t b = (uint32_t)hi;
                                const int kind = bkind<HAS_H>(D, b);
                                uint32_t v0, v1;
                                const uint32_t eb = evars(b, kind, v0, v1);
                                const uint32_t s0 = state_bit(S.st, v0), s1 = kind == KIND_BOND ? state_bit(S.st, v1) : 0u;
                                hz = state_bit(S.cd, v0) || (kind == KIND_BOND && state_bit(S.cd, v1));
                                const double num = bn * ewt(eb, b, kind, s0, s1);
                                const uint32_t bitsv = s0 | (s1 << 1);
                                opw = make_op(b, bitsv, bitsv);
                                if (num >= dhiA) ok = true;                      // inserted without a second word
                                else if (num == 0.0) ex = true;                  // gen_bool(0.0): one word, false
                                else if (num > 0.0 && num < dloA && y + 1 < 64) {
                                    ex = true;
                                    const uint64_t v2 = S.win[y + 1];
                                    if (v2 < bool_threshold(num / dhiA)) ok = true;
                                    else if (!(v2 >= bool_threshold(num / dloA))) hz = true;
                                } else hz = true;
                            }
                            S.opw[y] = opw;
                            accb[hh] = __ballot_sync(FULL, acc), exb[hh] = __ballot_sync(FULL, ex);
                            okb[hh] = __ballot_sync(FULL, ok), hzb[hh] = __ballot_sync(FULL, hz);
                        }
                        TICK(3);  // diagonal-lane thresholds + decode of the window
                        const uint64_t ACC = ((uint64_t)accb[1] << 32) | accb[0], EX = ((uint64_t)exb[1] << 32) | exb[0];
                        const uint64_t OKI = ((uint64_t)okb[1] << 32) | okb[0], HZ = ((uint64_t)hzb[1] << 32) | hzb[0];
#pragma unroll
                        for (int hh = 0; hh < 2; hh++) {  // successor table
                            const uint32_t x = (uint32_t)lane + 32u * hh;
                            const uint64_t m = ACC >> x;
                            uint32_t g = 255u;  // window exhausted
                            if (m) {
                                const uint32_t y = x + (uint32_t)__ffsll((long long)m) - 1u;
                                g = ((HZ >> y) & 1ull) ? 254u : y + 1u + (uint32_t)((EX >> y) & 1ull);
                            }
                            S.G[x] = (unsigned char)g;
                        }
                        __syncwarp();
                        // uniform walk over the EMPTY lanes only: between two of them the cursor advances by the
                        // number of drawing diagonal ops, which every lane counts for itself with mask arithmetic
                        const uint32_t stopA = ambD ? (uint32_t)__ffs(ambD) - 1u : 32u;  // needs the exact path
                        const uint32_t predE = remE & lt_mask;
                        const int le = predE ? 31 - __clz(predE) : -1;  // last EMPTY lane before me
                        const uint32_t after_le = le < 0 ? FULL : ~((2u << le) - 1u);
                        const uint32_t dcount = (uint32_t)__popc(drawD & lt_mask & after_le);  // draws between it and me
                        const uint32_t kidx = (uint32_t)__popc(predE);  // my index among the EMPTY lanes
                        if (inrem && type == T_EMPTY) S.wk[kidx] = (unsigned char)dcount, S.wl[kidx] = (unsigned char)lane;
                        __syncwarp();
                        const uint32_t x0 = (uint32_t)(cur - wbase);
                        uint32_t x = x0, myx = 0, stop = 32u;
                        bool hz_fail = false, overridden = false;
                        const uint32_t nE = (uint32_t)__popc(stopA < 32u ? (remE & ((1u << stopA) - 1u)) : remE);
                        uint32_t k = 0;
                        for (;;) {
                            // branch-free: a hazard (254) or an exhausted window (255) is absorbing, since every later
                            // start is clamped to G[64] = 255; the first such k is found afterwards with one ballot
                            const uint32_t kbeg = k;
#pragma unroll 4
                            for (uint32_t kk = kbeg; kk < nE; kk++) {
                                const uint32_t xs = min(x + (uint32_t)S.wk[kk], 64u);
                                x = S.G[xs];
                                S.wxg[kk] = (unsigned short)(xs | (x << 8));
                            }
                            __syncwarp();
                            const uint32_t mine = ((uint32_t)lane >= kbeg && (uint32_t)lane < nE) ? (uint32_t)S.wxg[lane] : 0u;
                            const uint32_t badm = __ballot_sync(FULL, (mine >> 8) >= 254u);
                            k = badm ? (uint32_t)__ffs(badm) - 1u : nE;
                            if (!badm) break;
                            const uint32_t first = __shfl_sync(FULL, mine, (int)k);
                            if ((first >> 8) == 255u) {
                                stop = S.wl[k];
                                break;
                            }
                            const uint32_t hz_x = first & 255u;
                            // hazard: the word's spins are flipped inside this step -> evaluate for that lane
                            DBG(12, 1);
                            const uint32_t i = S.wl[k], xs = hz_x;
                            const uint32_t y = xs + (uint32_t)__ffsll((long long)(ACC >> xs)) - 1u;
                            const uint32_t b = op_bond(S.opw[y]);
                            const int kind = bkind<HAS_H>(D, b);
                            uint32_t v0, v1;
                            const uint32_t eb = evars(b, kind, v0, v1);
                            uint32_t s0 = state_bit(S.st, v0), s1 = kind == KIND_BOND ? state_bit(S.st, v1) : 0u;
                            for (uint32_t m2 = __ballot_sync(FULL, type == T_OFFD) & ((1u << i) - 1u); m2; m2 &= m2 - 1) {
                                const uint32_t fv = S.fl[__ffs(m2) - 1];
                                s0 ^= (fv == v0), s1 ^= (kind == KIND_BOND && fv == v1);
                            }
                            const double num = bn * ewt(eb, b, kind, s0, s1);
                            const uint32_t bitsv = s0 | (s1 << 1);
                            bool ok = false, ex = false, fail = false;
                            if (num >= dhiA) ok = true;
                            else if (num == 0.0) ex = true;
                            else if (num > 0.0 && num < dloA && y + 1 < 64) {
                                ex = true;
                                const uint64_t v2 = S.win[y + 1];
                                if (v2 < bool_threshold(num / dhiA)) ok = true;
                                else if (!(v2 >= bool_threshold(num / dloA))) fail = true;
                            } else fail = true;
                            if (fail) { stop = i, hz_fail = true; break; }
                            const uint32_t g = y + 1u + (ex ? 1u : 0u);
                            if ((uint32_t)lane == i) overridden = true, neww = ok ? make_op(b, bitsv, bitsv) : OP_EMPTY, dn = ok, dc = g - xs;
                            __syncwarp();
                            S.wxg[k] = (unsigned short)(xs | (g << 8));
                            x = g;
                            k++;
                        }
                        __syncwarp();
                        // every lane picks up its own cursor: EMPTY lanes their start, the others the cursor
                        // after their last EMPTY predecessor plus the diagonal draws in between
                        TICK(4);  // successor table + walk
                        const uint32_t kstop = k;  // EMPTY lanes with index >= kstop are unresolved
                        if (type == T_EMPTY) myx = kidx < kstop ? (uint32_t)S.wxg[kidx] & 255u : 0u;
                        else myx = (kidx == 0 ? x0 : (kidx <= kstop ? (uint32_t)S.wxg[kidx - 1] >> 8 : 0u)) + dcount;
                        if (stop >= 32u && kstop < nE) stop = S.wl[kstop];
                        if (stop >= 32u && stopA < 32u) stop = stopA;
                        bool need_exact = hz_fail || (stopA < 32u && stop == stopA);
                        // a drawing diagonal op beyond the window ends the resolved prefix as well
                        const uint32_t ovD = __ballot_sync(FULL, inrem && type == T_DIAG && draws && myx >= 64u && (uint32_t)lane < stop);
                        if (ovD) stop = (uint32_t)__ffs(ovD) - 1u, need_exact = false;
                        const uint32_t resolved = stop >= 32u ? rem : (rem & ((1u << stop) - 1u));
                        // decisions of the resolved lanes
                        bool unres = false;
                        if ((resolved >> lane) & 1u) {
                            if (type == T_EMPTY) {
                                if (!overridden) {
                                    const uint32_t y = myx + (uint32_t)__ffsll((long long)(ACC >> myx)) - 1u;
                                    const bool ok = (OKI >> y) & 1ull;
                                    neww = ok ? S.opw[y] : OP_EMPTY;
                                    dn = ok, dc = y + 1u + (uint32_t)((EX >> y) & 1ull) - myx;
                                }
                            } else if (!draws) neww = OP_EMPTY, dn = -1, dc = 0;
                            else {
                                const uint64_t v = S.win[myx];
                                dc = 1;
                                if (v < t_lo) neww = OP_EMPTY, dn = -1;
                                else if (v >= t_hi) neww = w, dn = 0;
                                else unres = true;  // needs the exact n: resolved below, in order
                            }
                        }
                        for (uint32_t um = __ballot_sync(FULL, unres); um; um = __ballot_sync(FULL, unres)) {
                            DBG(13, 1);
                            const uint32_t u = (uint32_t)__ffs(um) - 1u;
                            const uint32_t pre = __reduce_add_sync(FULL, (((resolved >> lane) & 1u) && (uint32_t)lane < u) ? (uint32_t)(dn + 1) : 0u);
                            if ((uint32_t)lane == u) {
                                const uint32_t ni = n + pre - (uint32_t)__popc(resolved & lt_mask);
                                const double pr = ((double)(M - ni) + 1.0) / dnum;
                                const bool remove = S.win[myx] < bool_threshold(pr);
                                neww = remove ? OP_EMPTY : w, dn = remove ? -1 : 0, unres = false;
                            }
                        }
                        okm = resolved;
                        if (okm == 0 || need_exact) {
                            if (okm == 0) {
                                if (!need_exact) err |= DEV_ERR_INVARIANT;  // a fresh window cannot be exhausted by one lane
                                DBG(2, 1);
                                try_fast = false;
                                if (inrem) neww = w;
                                __syncwarp();
                                continue;
                            }
                            try_fast = false;  // commit the prefix, then the exact path takes the next lane
                        }
                    } else {
                        // ---- exact fixed point (rare): every lane re-evaluates until no input changes
                        try_fast = true;
                        DBG(3, 1);
                        dc = inrem ? 1u : 0u;
                        bool ovf = false;
                        uint32_t last_pc = 0xFFFFFFFFu, last_pn = 0xFFFFFFFFu;
                        for (;;) {
                            const uint32_t packed = inrem ? ((dc << 8) | (uint32_t)(dn + 1)) : 0u;
                            const uint32_t ex = warp_excl_scan(packed, lane);
                            const uint32_t pc = ex >> 8, pn_b = ex & 0xFFu;
                            bool changed = false;
                            if (inrem && (pc != last_pc || pn_b != last_pn)) {
                                last_pc = pc, last_pn = pn_b;
                                const uint32_t ni = n + pn_b - (uint32_t)__popc(rem & lt_mask);
                                const uint32_t idx = (uint32_t)(cur - wbase) + pc;
                                uint32_t ndc = 0;
                                int ndn = 0;
                                bool novf = false;
                                uint32_t nw = w;
                                if (type == T_EMPTY) {
                                    uint64_t hi = 0, lo;
                                    for (;;) {  // gen_range(0..Nb)
                                        if (idx + ndc >= 64) { novf = true; break; }
                                        const uint64_t v = S.win[idx + ndc];
                                        ndc++;
                                        hi = __umul64hi(v, range), lo = v * range;
                                        if (lo <= zone) break;
                                    }
                                    if (!novf) {
                                        const uint32_t b = (uint32_t)hi;
                                        const int kind = bkind<HAS_H>(D, b);
                                        uint32_t v0, v1;
                                        const uint32_t eb = evars(b, kind, v0, v1);
                                        uint32_t s0 = state_bit(S.st, v0), s1 = kind == KIND_BOND ? state_bit(S.st, v1) : 0u;
                                        for (uint32_t m = fmask_lt; m; m &= m - 1) {  // flips by earlier lanes of this step
                                            const uint32_t fv = S.fl[__ffs(m) - 1];
                                            s0 ^= (fv == v0), s1 ^= (kind == KIND_BOND && fv == v1);
                                        }
                                        const double num = bn * ewt(eb, b, kind, s0, s1);
                                        const double den = (double)(M - ni);
                                        bool accept = num > den;
                                        if (!accept) {
                                            const double pr = num / den;
                                            if (pr == 1.0) accept = true;
                                            else if (!(pr >= 0.0 && pr < 1.0)) err |= DEV_ERR_PROB;
                                            else if (idx + ndc >= 64) novf = true;
                                            else accept = S.win[idx + ndc++] < bool_threshold(pr);
                                        }
                                        if (accept) {
                                            const uint32_t bitsv = s0 | (s1 << 1);
                                            nw = make_op(b, bitsv, bitsv);
                                            ndn = 1;
                                        } else nw = OP_EMPTY;
                                    }
                                } else {  // T_DIAG
                                    const double den = (double)(M - ni) + 1.0;
                                    bool remove = den > dnum;
                                    if (!remove) {
                                        const double pr = den / dnum;
                                        if (pr == 1.0) remove = true;
                                        else if (!(pr >= 0.0 && pr < 1.0)) err |= DEV_ERR_PROB;
                                        else if (idx >= 64) novf = true;
                                        else { remove = S.win[idx] < bool_threshold(pr); ndc = 1; }
                                    }
                                    if (remove) nw = OP_EMPTY, ndn = -1;
                                }
                                if (novf) ndc = 0, ndn = 0;
                                changed = ndc != dc || ndn != dn || novf != ovf;
                                dc = ndc, dn = ndn, ovf = novf, neww = nw;
                            }
                            if (!__any_sync(FULL, changed)) break;
                        }
                        // lanes before the first window overflow are final
                        const uint32_t ovm = __ballot_sync(FULL, inrem && ovf);
                        okm = ovm ? (rem & ((1u << (__ffs(ovm) - 1)) - 1u)) : rem;
                        if (okm == 0) {  // cannot happen: a fresh window holds >= 63 words
                            err |= DEV_ERR_INVARIANT;
                            okm = rem;
                        }
                    }
                    // advance (cursor, n) past the committed lanes
                    const bool mine = (okm >> lane) & 1u;
                    const uint32_t tot = __reduce_add_sync(FULL, mine ? dc : 0u);
                    const int totn = (int)__reduce_add_sync(FULL, mine ? (uint32_t)(dn + 1) : 0u) - __popc(okm);
                    if (inrem && !mine) neww = w;
                    cur += tot;
                    n = (uint32_t)((int)n + totn);
                    rem &= ~okm;
                    __syncwarp();
                }
                TICK(5);  // decisions + commit
                if (__any_sync(FULL, neww != w)) {
                    if (valid && neww != w) st_cg(ops + p, neww);
                }
                if (type == T_OFFD) atomicXor(&S.st[ov0 >> 5], 1u << (ov0 & 31)), atomicAnd(&S.cd[ov0 >> 5], ~(1u << (ov0 & 31)));
                __syncwarp();
            }
            TICK(6);  // store + state flips
            if (PIPE) ring[(it & 1u) * 32 + lane] = neww;
          }
            if (PIPE) {  // role 1 works on the step role 0 finished one iteration ago
                base = (it - 1u) * 32, p = base + lane, valid = it >= 1 && p < M;
                if (roleB && it >= 1) neww = ring[((it - 1u) & 1u) * 32 + lane];
            }
            if (do_clus && roleB && (!PIPE || it >= 1)) {
                // ---- segments and unions on the final ops of this step
                const uint32_t fw = neww;
                int kind = -1;
                uint32_t v0 = 0, v1 = 0;
                if (valid && fw != OP_EMPTY) {
                    kind = bkind<HAS_H>(D, op_bond(fw));
                    evars(op_bond(fw), kind, v0, v1);
                }
                const uint32_t smask = __ballot_sync(FULL, kind == KIND_SITE);
                TICK(11);  // decode of the final ops
                const uint32_t myid = N + nsite + (uint32_t)__popc(smask & lt_mask);
                if (kind == KIND_SITE) st_cg(P + myid, myid);
                if (kind >= 0) {
                    if (!state_bit(S.tb, v0)) atomicOr(&S.tb[v0 >> 5], 1u << (v0 & 31));
                    if (kind == KIND_BOND && !state_bit(S.tb, v1)) atomicOr(&S.tb[v1 >> 5], 1u << (v1 & 31));
                }
                // representative of the segment open on my variables at my slot: the table entry, unless a
                // site op of this step cuts the variable at an earlier lane (uniform loop over those site ops)
                const bool joins = kind == KIND_BOND || (HAS_H && kind == KIND_LONG);
                uint32_t ra = 0, rb = 0, oa = 0, ob = 0;
                bool fa = false, fb = false;  // value came from the table (may be refreshed with the root)
                if (joins) {
                    oa = ra = S.rep[v0], fa = true;
                    if (kind == KIND_BOND) ob = rb = S.rep[v1], fb = true;
                }
                // Does a site op of this step cut one of my variables at an earlier lane?  The lanes match their keys
                // against each other (two MATCH.ANY of ~350 cycles each on the dependent chain).  In the low-occupancy
                // build, where that latency is exposed, the site ops first mark their variable in a bitmap and the
                // joining ops look their two variables up: only about a quarter of the steps have a hit and need the
                // match (-7 % time there; the 72-register build has no shared memory to spare for the bitmap).
                constexpr bool LOWOCC = MINB <= 4;
                bool need_match = true;
                if (LOWOCC) {
                    if (kind == KIND_SITE) atomicOr(&S.sb[v0 >> 5], 1u << (v0 & 31));
                    __syncwarp();
                    const bool hit = joins && (state_bit(S.sb, v0) || (kind == KIND_BOND && state_bit(S.sb, v1)));
                    need_match = __any_sync(FULL, hit);
                }
                if (need_match) {
                    // lanes with the same key see each other: site ops publish their variable in both rounds,
                    // joining ops ask for v0 then v1; the nearest earlier site op on the variable wins
                    const uint32_t nokey = 0x80000000u | (uint32_t)lane;
                    const uint32_t ma = __match_any_sync(FULL, kind >= 0 ? v0 : nokey) & smask & lt_mask;
                    const uint32_t mb = __match_any_sync(FULL, kind == KIND_BOND ? v1 : (kind == KIND_SITE ? v0 : nokey)) & smask & lt_mask;
                    if (joins && ma) ra = N + nsite + (uint32_t)__popc(smask & ((1u << (31 - __clz(ma))) - 1u)), fa = false;
                    if (kind == KIND_BOND && mb) rb = N + nsite + (uint32_t)__popc(smask & ((1u << (31 - __clz(mb))) - 1u)), fb = false;
                }
                __syncwarp();  // new ids are initialised before anyone follows them; bitmap reads are done
                if (LOWOCC && kind == KIND_SITE) atomicAnd(&S.sb[v0 >> 5], ~(1u << (v0 & 31)));
                TICK(12);  // touched bits, representatives, same-step site ops (match)
                if (kind == KIND_BOND && ra != rb) {
                    // (an atomic-free variant -- match the roots, lowest lane stores, others retry -- measured 4 % slower)
                    const uint32_t root = uf_union_cg(P, ra, rb);
                    if (fa) atomicCAS(&S.rep[v0], oa, root);  // cache the root: equal roots skip the union
                    if (fb) atomicCAS(&S.rep[v1], ob, root);
                }
                TICK(13);  // unions
                if (HAS_H && kind == KIND_LONG) {
                    atomicOr(&frz[ra >> 5], 1u << (ra & 31));
                    anylong = true;
                }
                // the site op with the highest lane owns the variable from here on (ids grow with lane)
                if (kind == KIND_SITE) atomicMax(&S.rep[v0], myid);
                nsite += (uint32_t)__popc(smask);
                __syncwarp();
                TICK(7);  // segment ids + union-find
            }
            PAIR_SYNC();
        }
        if (do_diag && lane == 0 && roleA) D.n[r] = n;
        if (PIPE) {  // role 1 needs n and the stream position for the closure and P2
            if (roleA && lane == 0) xchg[0] = n, xchg[1] = (uint32_t)cur, xchg[2] = (uint32_t)(cur >> 32);
            PAIR_SYNC();
            if (!roleA) n = xchg[0], cur = ((uint64_t)xchg[2] << 32) | xchg[1];
        }

        uint32_t ncl = 0;
        if (do_clus && n > 0) {
          const uint64_t c0 = cur;
          if (roleB) {
            // periodic closure: the segment open at the end of variable v is the one crossing p = 0
            for (uint32_t v = lane; v < N; v += 32) {
                const uint32_t rp = S.rep[v];
                if (rp != v) uf_union_cg(P, v, rp);
            }
            __syncwarp();
            __threadfence_block();
            TICK(8);  // closure
            const uint32_t nseg = N + nsite;
            const uint32_t nwords = (nseg + 31) / 32;
            bool frozen_all = false;
            if (HAS_H) {
                frozen_all = __any_sync(FULL, anylong);
                // push frozen marks up to the roots: descending ids, parents are smaller
                for (int32_t wd = (int32_t)nwords - 1; wd >= 0; wd--) {
                    const uint32_t x = (uint32_t)wd * 32 + lane;
                    uint32_t par = x < nseg ? ld_cg(P + x) : x;
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
            // =========================== P2: one flip bit per segment ===========================
            uint32_t nroots = 0;
            Philox4 rb4 = {0, 0, 0, 0};
            for (uint32_t wd = 0; wd < nwords; wd++) {
                const uint32_t x = wd * 32 + lane;
                if ((wd & 3u) == 0) rb4 = philox4x32_10(wd >> 2, (uint32_t)c0, (uint32_t)(c0 >> 32), QMCB_TAG_CLUS, k0, k1);
                const uint32_t rword = (wd & 3u) == 0 ? rb4.x : ((wd & 3u) == 1 ? rb4.y : ((wd & 3u) == 2 ? rb4.z : rb4.w));
                uint32_t par = x < nseg ? ld_cg(P + x) : x;
                if (nsite == 0) par = x < nseg ? 0u : x;  // cluster.rs:98-107: no cluster edge => one cluster
                const bool root = x < nseg && par == x;
                nroots += (uint32_t)__popc(__ballot_sync(FULL, root));
                bool dec = false, known = root || x >= nseg;
                if (root) {
                    dec = (rword >> lane) & 1u;
                    if (HAS_H) dec = dec && !((nsite == 0) ? frozen_all : ((ld_cg(frz + wd) >> lane) & 1u));
                } else if (x < nseg && par < wd * 32) {
                    dec = (ld_cg(decb + (par >> 5)) >> (par & 31)) & 1u;
                    known = true;
                }
                // parents inside this word: resolve by rounds (parent id < child id)
                for (;;) {
                    const uint32_t kmask = __ballot_sync(FULL, known), dmask = __ballot_sync(FULL, dec);
                    if (kmask == FULL) {
                        if (lane == 0) st_cg(decb + wd, dmask);
                        break;
                    }
                    if (!known) {
                        const uint32_t pl = par - wd * 32;
                        if ((kmask >> pl) & 1u) dec = (dmask >> pl) & 1u, known = true;
                    }
                }
                __syncwarp();
            }
            uint32_t untouched = 0;
            for (uint32_t j = lane; j < Nw; j += 32) {
                const uint32_t validm = (j == Nw - 1 && (N & 31u)) ? ((1u << (N & 31u)) - 1u) : FULL;
                untouched += (uint32_t)__popc(~S.tb[j] & validm);
            }
            untouched = __reduce_add_sync(FULL, untouched);
            ncl = nsite == 0 ? 1u : nroots - untouched;
            if (PIPE && lane == 0) xchg[3] = ncl;

            TICK(9);  // P2
          }
          PAIR_SYNC();  // the flip bits are complete
          if (roleA) {
            if (PIPE) ncl = xchg[3];
            // =========================== P3: apply the flips ===========================
            for (uint32_t j = lane; j < Nw; j += 32) S.cd[j] = ld_cg(decb + j);
            __syncwarp();
            uint32_t ks = 0;
            if (M) fetch_line(S.line, ops, lane);
            // the site ops of one step have consecutive ids: their flip bits sit in two consecutive words of decb.
            // A sliding window of three words is kept in registers (loaded one step before it can be needed), so
            // that no L2 round trip sits on the step's dependent chain.
            const uint32_t dlast = (uint32_t)bstride - 1u;
            uint32_t wi = N >> 5;
            uint32_t dw0 = ld_cg(decb + min(wi, dlast)), dw1 = ld_cg(decb + min(wi + 1u, dlast)), dw2 = ld_cg(decb + min(wi + 2u, dlast));
            for (uint32_t base = 0; base < M; base += 32) {
                const uint32_t p = base + lane;
                uint32_t w = take_line(S.line, lane);
                if (p >= M) w = OP_EMPTY;
                if (base + 32 < M) fetch_line(S.line, ops + base + 32, lane);
                int kind = -1;
                uint32_t v0 = 0, v1 = 0;
                if (w != OP_EMPTY) {
                    kind = bkind<HAS_H>(D, op_bond(w));
                    evars(op_bond(w), kind, v0, v1);
                }
                const uint32_t smask = __ballot_sync(FULL, kind == KIND_SITE);
                bool outdec = false;
                if (kind == KIND_SITE) {
                    const uint32_t id = N + ks + (uint32_t)__popc(smask & lt_mask);
                    outdec = (((id >> 5) == wi ? dw0 : dw1) >> (id & 31)) & 1u;
                }
                const uint32_t odmask = __ballot_sync(FULL, outdec);
                bool din = kind >= 0 ? ((S.cd[v0 >> 5] >> (v0 & 31)) & 1u) : false;
                const uint32_t mt = __match_any_sync(FULL, kind >= 0 ? v0 : (0x80000000u | (uint32_t)lane)) & smask;
                if (mt & lt_mask) din = (odmask >> (31 - __clz(mt & lt_mask))) & 1u;  // segment opened by the nearest earlier site op
                const bool lastone = (mt & ~lt_mask & ~(1u << lane)) == 0;
                if (kind >= 0) {
                    const uint32_t mask = kind == KIND_BOND ? 3u : 1u;
                    const bool dout = kind == KIND_SITE ? outdec : din;
                    if (din || dout) st_cg(ops + p, make_op(op_bond(w), op_in(w) ^ (din ? mask : 0u), op_out(w) ^ (dout ? mask : 0u)));
                }
                __syncwarp();
                if (kind == KIND_SITE && lastone) {  // the last site op of the step on a variable sets its open decision
                    if (outdec) atomicOr(&S.cd[v0 >> 5], 1u << (v0 & 31));
                    else atomicAnd(&S.cd[v0 >> 5], ~(1u << (v0 & 31)));
                }
                ks += (uint32_t)__popc(smask);
                if (((N + ks) >> 5) != wi) {  // at most one word per step
                    wi++;
                    dw0 = dw1, dw1 = dw2, dw2 = ld_cg(decb + min(wi + 2u, dlast));
                }
                __syncwarp();
            }
            TICK(10);  // P3
            // spins: the segment of variable v crossing p = 0 has id v
            for (uint32_t j = lane; j < Nw; j += 32) S.st[j] ^= ld_cg(decb + j) & S.tb[j];
            cur = c0 + 1;
            __syncwarp();
          }
        }
        if (do_clus && roleA) {
            // free spins: qmc_ising.rs:780-784
            for (uint32_t base = 0; base < N; base += 32) {
                const uint32_t v = base + lane;
                const bool fr = v < N && !((S.tb[v >> 5] >> (v & 31)) & 1u);
                const uint32_t m = __ballot_sync(FULL, fr);
                bool bit = false;
                if (fr) bit = stream_word(key, cur + __popc(m & lt_mask)) < 0x8000000000000000ull;
                const uint32_t setm = __ballot_sync(FULL, bit);
                if (lane == 0 && m) S.st[base >> 5] = (S.st[base >> 5] & ~m) | setm;
                cur += __popc(m);
            }
            __syncwarp();
            if (lane == 0) D.ncl[r] = ncl;
        }
        if (roleA)
            for (uint32_t j = lane; j < Nw; j += 32) gstate[j] = S.st[j];
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
                    for (uint32_t v = lane; v < N; v += 32) dst[v] = (uint8_t)state_bit(S.st, v);
                }
            }
            if (lane == 0 && roleA) D.done[r] = done;
        }
        __syncwarp();
        PAIR_SYNC();  // role 1 re-initialises its tables only after role 0 is done with them
    }
    if (err) atomicOr(D.status, err);
}

// returns the number of kernel launches, or -1 if this shape is not supported by the warp kernels
int launch_sse_fast(const SseDev &D, const SseTuning &T, uint64_t target, uint32_t phases, uint64_t sample_freq, uint64_t sample_origin,
                    uint8_t *samples, uint64_t samples_per_rep, cudaStream_t st) {
    // one warp per block: up to 32 blocks are resident per SM, each with its own shared memory
    size_t smem = warp_smem_bytes(D.N, D.Nw, false);
    if (warp_smem_bytes(D.N, D.Nw, true) * QMCB_WPB + 1024 > 227 * 1024) return -1;
    const uint32_t blocks = (D.R + QMCB_WPB - 1) / QMCB_WPB;
    if (smem * QMCB_WPB + 1024 > 227 * 1024) return -1;
    typedef void (*Kern)(SseDev, uint64_t, uint32_t, uint64_t, uint64_t, uint8_t *, uint64_t, uint32_t, uint32_t);
    // register budget: the kernel is bound by the latency of each warp's dependent chain, so when shared memory or the
    // number of replicas keeps few blocks resident anyway, the 120-register build (no spills) is 10-20 % faster per
    // warp; with many replicas 72 registers keep 4096 of them (28 warps per SM) in one wave
    const int nsm = T.nsm > 0 ? T.nsm : 148;
    const size_t wanted = (blocks + nsm - 1) / nsm;
    int minb = T.minblocks;
    if (minb <= 0) {
        const size_t resident = std::min((size_t)(227 * 1024) / (smem * QMCB_WPB + 1024), wanted);
        minb = resident <= 4 ? 4 : (resident <= 6 ? 6 : 7);
    }
    if (minb <= 4) smem = warp_smem_bytes(D.N, D.Nw, true);  // the low-occupancy layout carries the site-op bitmap
    // block-shared packed edge table: low-occupancy build only, and only if it does not cost a resident block
    size_t epk_bytes = 0;
    if (minb == 4 && D.epk && !D.ham && T.epk) {
        const size_t want = ((size_t)D.E * 4 + 15) / 16 * 16;
        const size_t without = std::min((size_t)(227 * 1024) / (smem * QMCB_WPB + 1024), wanted);
        const size_t with = std::min((size_t)(227 * 1024) / (smem * QMCB_WPB + want + 1024), wanted);
        if (with >= 1 && with >= std::min<size_t>(without, 4)) epk_bytes = want;
    }
    // with many resident warps the packed table is still used, through L1: 8 KB instead of 32 KB of tables at config #3
    const bool pk_l1 = !epk_bytes && D.epk && !D.ham && T.epk && minb == 7;
    // two warps per replica (PIPE) when at most two blocks of four replicas per SM are wanted and fit
    const bool pipe = minb == 4 && T.pipe && wanted <= 2 && (size_t)(227 * 1024) / (smem * QMCB_WPB + epk_bytes + 1024) >= wanted;
    Kern kern;
#define PICK(HB_, MH_)                                                                                                  \
    switch (minb) {                                                                                                     \
        case 4:                                                                                                         \
            if (pipe) {                                                                                                 \
                if (epk_bytes) kern = D.has_h ? k_sse_fast<true, 4, HB_, MH_, 1, true> : k_sse_fast<false, 4, HB_, MH_, 1, true>; \
                else kern = D.has_h ? k_sse_fast<true, 4, HB_, MH_, 0, true> : k_sse_fast<false, 4, HB_, MH_, 0, true>;  \
            } else if (epk_bytes) kern = D.has_h ? k_sse_fast<true, 4, HB_, MH_, 1, false> : k_sse_fast<false, 4, HB_, MH_, 1, false>; \
            else kern = D.has_h ? k_sse_fast<true, 4, HB_, MH_, 0, false> : k_sse_fast<false, 4, HB_, MH_, 0, false>;    \
            break;                                                                                                      \
        case 6: kern = D.has_h ? k_sse_fast<true, 6, HB_, MH_, 0, false> : k_sse_fast<false, 6, HB_, MH_, 0, false>; break; \
        case 8: kern = (!HB_ && !MH_) ? (D.has_h ? k_sse_fast<true, 8, false, false, 0, false> : k_sse_fast<false, 8, false, false, 0, false>) \
                                      : (D.has_h ? k_sse_fast<true, 7, HB_, MH_, 0, false> : k_sse_fast<false, 7, HB_, MH_, 0, false>); break; \
        default:                                                                                                        \
            if (pk_l1 && !MH_) kern = D.has_h ? k_sse_fast<true, 7, HB_, false, 2, false> : k_sse_fast<false, 7, HB_, false, 2, false>; \
            else kern = D.has_h ? k_sse_fast<true, 7, HB_, MH_, 0, false> : k_sse_fast<false, 7, HB_, MH_, 0, false>;    \
            break;                                                                                                      \
    }
    if (D.ham) {
        if (D.hb_cum) { PICK(true, true) } else { PICK(false, true) }
    } else if (D.hb_cum) { PICK(true, false) } else { PICK(false, false) }
#undef PICK
    // function attributes are set once per kernel instance (and again only when the carve-out knob changes)
    {
        static std::mutex mu;
        static std::map<std::pair<const void *, int>, int> done;  // (kernel, device) -> carve-out last set
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        auto it = done.find({(const void *)kern, dev});
        if (it == done.end()) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (T.carveout >= 0 && (it == done.end() || it->second != T.carveout))
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, T.carveout);
        done[{(const void *)kern, dev}] = T.carveout;
    }
    size_t dyn = smem * QMCB_WPB + epk_bytes + (size_t)std::max(T.pad, 0);
    if (dyn > 227 * 1024) dyn = 227 * 1024;  // the padding knob never pushes the block past the hardware limit
    kern<<<blocks, (pipe ? 64 : 32) * QMCB_WPB, dyn, st>>>(D, target, phases, sample_freq, sample_origin, samples, samples_per_rep, (uint32_t)smem,
                                                                  epk_bytes ? (uint32_t)(smem * QMCB_WPB) : 0u);
    return 1;
}
