// sse_rvb.cu -- the resonating-valence-bond cluster move of the reference (RvbUpdater::rvb_update_with_ising_weight,
// rvb.rs:60-291, as QmcIsingGraph::timestep calls it between the diagonal and the cluster update, qmc_ising.rs:705-752;
// set_run_rvb :434-441, single_rvb_sweep :322-420).
//
// The move is a chain of dependent decisions under the replica's sequential stream (pick a constant op or an op-less
// variable, grow a cluster of world-line pieces through a weighted boundary, weigh the flip, rotate the diagonal ops on
// the cluster's border), so replicas are the parallel axis: one warp per replica, the walk on lane 0, the other lanes
// help with the workspace set-up.  The reference finds ops through per-variable links, binary heaps and "hints"
// (fast_ops.rs:639-808, 896-1172); those decide HOW ops are found, not WHICH: calculate_flip_prob (rvb.rs:649-946) and
// mutate_subsection_ops visit every op that touches a sub-variable, in p order, and get_propagated_substate_with_hint
// returns the state just before p.  This kernel makes the same visits on the flat operator words plus one sorted
// position list per variable (the ops on its world line, built once per launch with slack, patched when a rotation moves
// an op to other variables, rebuilt if a list outgrows its slack): "next op on a sub-variable's line at or after p" is a
// binary search per sub-variable, so a proposal costs what the cluster touches, not the length of the string.  What is
// kept literally because results depend on it: every draw and its order, every f64 operation (the running totals of the weighted sets are accumulated in the reference's order), and
// the key order inside BondContainer (util/bondcontainer.rs:10-159: push at the end, swap-remove), which decides what
// get_random returns.  The move is off by default in the reference (qmc_ising.rs:122) and here; it is not a tuned path.
#include "sse.cuh"

// -DQMCB_RVB_CHECK: every write with a computed index is checked against the size of its array (compute-sanitizer is not
// available on the pool this was developed on); a violation ends up in the handle's status as DEV_ERR_STACK
#ifdef QMCB_RVB_CHECK
__device__ int g_rvb_check_fail = 0;
#define RVB_CHK(cond)                                \
    do {                                             \
        if (!(cond)) atomicOr(&g_rvb_check_fail, 1); \
    } while (0)
#else
#define RVB_CHK(cond) ((void)0)
#endif

// room a world line gets beyond its length (rotations move ops between lines); -DQMCB_RVB_TIGHT_LINES builds lines with
// one spare entry so that the rebuild path runs constantly (test builds only)
#ifdef QMCB_RVB_TIGHT_LINES
#define RVB_LINE_SLACK(len) 1u
#else
#define RVB_LINE_SLACK(len) ((len) / 4 + 32)
#endif

namespace {

// BondContainer<T> (util/bondcontainer.rs): T = bond (kp == nullptr) or VarPos{v, p} (rvb.rs:957-965, index p.unwrap_or(v))
struct BC {
    uint32_t *map, *kv, *kp;
    double *kw;
    uint32_t maplen, len;
    double total;
};
__device__ __forceinline__ uint32_t bc_idx(uint32_t v, uint32_t p) { return p != NONE32 ? p : v; }
__device__ __forceinline__ bool bc_contains(const BC &c, uint32_t idx) { return idx < c.maplen && c.map[idx] != NONE32; }
__device__ __forceinline__ void bc_clear(BC &c) {  // :133-142
    for (uint32_t i = 0; i < c.len; i++) c.map[bc_idx(c.kv[i], c.kp ? c.kp[i] : NONE32)] = NONE32;
    c.len = 0, c.total = 0.0;
}
__device__ __forceinline__ void bc_insert(BC &c, uint32_t v, uint32_t p, double w) {  // :110-130
    const uint32_t t = bc_idx(v, p);
    RVB_CHK(t < c.maplen);
    const uint32_t at = c.map[t];
    if (at != NONE32) {
        const double old = c.kw[at];
        c.kw[at] = w;
        c.total = c.total + (w - old);
        if (c.total < 0.0) c.total = 0.0;  // correct_total_weight :76-87
    } else {
        RVB_CHK(c.len < c.maplen);  // the key arrays are as long as the map
        c.map[t] = c.len;
        c.kv[c.len] = v, c.kw[c.len] = w;
        if (c.kp) c.kp[c.len] = p;
        c.len++;
        c.total = c.total + w;
    }
}
__device__ __forceinline__ void bc_remove(BC &c, uint32_t v, uint32_t p) {  // :47-74: swap with the last key, pop
    const uint32_t t = bc_idx(v, p);
    if (t >= c.maplen) return;
    const uint32_t ki = c.map[t];
    if (ki == NONE32) return;
    const uint32_t last = c.len - 1;
    const uint32_t lv = c.kv[last], lp = c.kp ? c.kp[last] : NONE32;
    const double lw = c.kw[last], w = c.kw[ki];
    c.kv[ki] = lv, c.kw[ki] = lw;
    if (c.kp) c.kp[ki] = lp;
    c.map[bc_idx(lv, lp)] = ki;
    c.len = last;
    c.map[t] = NONE32;
    c.total = c.total - w;
    if (c.total < 0.0) c.total = 0.0;
}

// compiler-rt __powidf2 (what f64::powi lowers to)
__device__ __forceinline__ double powi_rt(double a, int b) {
    const bool recip = b < 0;
    double r = 1.0;
    for (;;) {
        if (b & 1) r = __dmul_rn(r, a);
        b /= 2;
        if (b == 0) break;
        a = __dmul_rn(a, a);
    }
    return recip ? __ddiv_rn(1.0, r) : r;
}

struct Op {  // a decoded operator word
    uint32_t b, nv, v[2], in, out;
    bool constant;
};

struct Ctx {
    const SseDev &D;
    const RvbDev &W;
    uint32_t *ops, *state;
    const double *J;
    uint32_t M;  // rvb.get_cutoff(): ops.len() (fast_ops.rs:1255-1257) = the cutoff
    uint64_t key, cur;
    int err;
    // find_constants, rvb.rs:1162-1188
    uint32_t *var_starts, *var_lengths, *zero_vars, *constant_ps, ncp, nzero;
    // WeightedBoundaryManager, :967-973
    BC fl, nf;
    uint8_t *pos_popped, *nopos_popped;
    uint32_t popped_idx[80], npopped_pos, npopped_nopos;
    uint32_t popped_idx2[80];
    // per update
    uint32_t cl_vars[72], cl_flips[72], ncl;
    uint32_t *subvars, *v2s, nsub;
    uint8_t *cstate, *substate, *mark;
    uint32_t toggles[150], ntog;
    BC bonds, bef, aft;
    // world lines: the positions of the ops on each variable, ascending, in lines[ln_start[v] .. + ln_len[v]) (room for ln_cap[v])
    uint32_t *ln_start, *ln_len, *ln_cap, *lines;
    size_t lines_total;
    uint32_t *c_idx, *c_pos;  // per sub-variable: cursor into its line and the position there (NONE32 past the end)

    __device__ __forceinline__ uint64_t next_u64() { return stream_word(key, cur++); }
    __device__ bool gen_bool(double p) {  // rand 0.8 Bernoulli
        if (!(p >= 0.0 && p < 1.0)) {
            if (p == 1.0) return true;
            err |= DEV_ERR_PROB;  // the reference panics
            return false;
        }
        return next_u64() < bool_threshold(p);
    }
    __device__ uint64_t gen_range(uint64_t range) {  // UniformInt<usize>::sample_single
        const uint64_t zone = (range << __clzll((long long)range)) - 1ull;
        uint64_t hi, lo;
        do {
            const uint64_t v = next_u64();
            hi = __umul64hi(v, range), lo = v * range;
        } while (lo > zone);
        return hi;
    }
    // BondContainer::get_random (:30-44): index of the chosen key, NONE32 where the reference panics
    __device__ uint32_t get_random(const BC &c) {
        if (c.len == 0 || !(0.0 < c.total) || isinf(c.total)) {
            err |= DEV_ERR_PROB;
            return NONE32;
        }
        double p;
        do p = unit_f64(next_u64()) * c.total + 0.0;  // gen_range(0. ..total_weight)
        while (!(p < c.total));
        uint32_t i = 0;
        while (i < c.len) {
            p = p - c.kw[i];
            if (p <= 0.0) break;
            i++;
        }
        if (i >= c.len) {
            err |= DEV_ERR_PROB;
            return NONE32;
        }
        return i;
    }
    __device__ __forceinline__ bool decode(uint32_t p, Op &o) const {
        const uint32_t w = ops[p];
        if (w == OP_EMPTY) return false;
        o.b = op_bond(w);
        const int kind = bond_kind(D, o.b);
        bond_vars(D, o.b, kind, o.v[0], o.v[1]);
        o.nv = kind == KIND_BOND ? 2u : 1u;
        o.in = op_in(w), o.out = op_out(w);
        o.constant = kind == KIND_SITE;
        return true;
    }
    // (re)build every world line from the operator words; slack so that rotations can move ops between lines
    __device__ void build_lines() {
        const uint32_t N = D.N;
        for (uint32_t v = 0; v < N; v++) ln_len[v] = 0;
        for (uint32_t p = 0; p < M; p++) {
            Op o;
            if (!decode(p, o)) continue;
            for (uint32_t r = 0; r < o.nv; r++) ln_len[o.v[r]]++;
        }
        size_t acc = 0;
        for (uint32_t v = 0; v < N; v++) {
            ln_start[v] = (uint32_t)acc, ln_cap[v] = ln_len[v] + RVB_LINE_SLACK(ln_len[v]);
            acc += ln_cap[v], ln_len[v] = 0;
        }
        if (acc > lines_total) {  // cannot happen: lines_total = 2.5 cap + 32 N + 64 and n <= cap
            err |= DEV_ERR_INVARIANT;
            return;
        }
        for (uint32_t p = 0; p < M; p++) {
            Op o;
            if (!decode(p, o)) continue;
            for (uint32_t r = 0; r < o.nv; r++) lines[ln_start[o.v[r]] + ln_len[o.v[r]]++] = p;
        }
    }
    // index of the first entry >= p on the line of variable v
    __device__ __forceinline__ uint32_t lower_bound(uint32_t v, uint32_t p) const {
        const uint32_t *ln = lines + ln_start[v];
        uint32_t lo = 0, hi = ln_len[v];
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (ln[mid] < p) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    }
    // point every sub-variable's cursor at its first op at or after pos
    __device__ void seek(uint32_t pos) {
        for (uint32_t s = 0; s < nsub; s++) {
            const uint32_t v = subvars[s], i = lower_bound(v, pos);
            c_idx[s] = i, c_pos[s] = i < ln_len[v] ? lines[ln_start[v] + i] : NONE32;
        }
    }
    // the first op at or after pos on a sub-variable's world line (what the reference's heaps pop next), or NONE32;
    // pos never decreases between two seeks, so the cursors only move forward
    __device__ uint32_t next_near(uint32_t pos) {
        uint32_t best = NONE32;
        for (uint32_t s = 0; s < nsub; s++) {
            uint32_t cp = c_pos[s];
            if (cp < pos) {
                const uint32_t v = subvars[s], len = ln_len[v];
                const uint32_t *ln = lines + ln_start[v];
                uint32_t i = c_idx[s];
                do cp = ++i < len ? ln[i] : NONE32;
                while (cp < pos);
                c_idx[s] = i, c_pos[s] = cp;
            }
            best = min(best, cp);
        }
        return best;
    }
    // the state of variable v just before p: the output of the last op below p on its line, else `otherwise`
    __device__ __forceinline__ uint32_t state_before(uint32_t v, uint32_t p, uint32_t otherwise) const {
        const uint32_t i = lower_bound(v, p);
        if (i == 0) return otherwise;
        Op t;
        decode(lines[ln_start[v] + i - 1], t);
        return (t.out >> (t.v[0] == v ? 0 : 1)) & 1u;
    }
    __device__ void line_remove(uint32_t v, uint32_t p) {
        uint32_t *ln = lines + ln_start[v];
        const uint32_t i = lower_bound(v, p), len = ln_len[v];
        if (i >= len || ln[i] != p) {
            err |= DEV_ERR_INVARIANT;
            return;
        }
        for (uint32_t k = i; k + 1 < len; k++) ln[k] = ln[k + 1];
        ln_len[v] = len - 1;
    }
    __device__ bool line_insert(uint32_t v, uint32_t p) {  // false: no room left, the caller rebuilds
        uint32_t *ln = lines + ln_start[v];
        const uint32_t len = ln_len[v];
        if (len >= ln_cap[v]) return false;
        const uint32_t i = lower_bound(v, p);
        RVB_CHK(ln_start[v] + len < lines_total);
        for (uint32_t k = len; k > i; k--) ln[k] = ln[k - 1];
        ln[i] = p;
        ln_len[v] = len + 1;
        return true;
    }
    __device__ __forceinline__ uint32_t other_var(uint32_t v, uint32_t b) const {  // rvb.rs:22-31
        const uint32_t a = __ldg(D.va + b), c = __ldg(D.vb + b);
        return v == a ? c : a;
    }
    // the diagonal_edge_hamiltonian closure of qmc_ising.rs:718-721 (two_site_hamiltonian :863-875 with in == out)
    __device__ __forceinline__ double edge_weight(uint32_t b, uint32_t sa, uint32_t sb) const {
        const double j = J[b];
        return fabs(j) + (sa == sb ? -j : j);
    }
    __device__ __forceinline__ double edge_weight_sub(uint32_t b) const {
        return edge_weight(b, substate[v2s[__ldg(D.va + b)]], substate[v2s[__ldg(D.vb + b)]]);
    }
    __device__ void ws_for_flip(uint32_t b, uint32_t sub_flip, double &wbef, double &waft) const {  // :665-683
        const uint32_t suba = v2s[__ldg(D.va + b)], subb = v2s[__ldg(D.vb + b)];
        uint32_t ba = substate[suba], bb = substate[subb];
        wbef = edge_weight(b, ba, bb);
        if (sub_flip == suba) ba ^= 1u;
        else bb ^= 1u;
        waft = edge_weight(b, ba, bb);
    }
    // push_adjacent, :1028-1047
    __device__ void push_adjacent(uint32_t var, uint32_t pos, bool has_w, double weight) {
        const double w = has_w ? weight : 1.0;
        BC &bd = pos != NONE32 ? fl : nf;
        const uint8_t *popped = pos != NONE32 ? pos_popped : nopos_popped;
        const uint32_t idx = bc_idx(var, pos);
        if (!popped[idx]) {
            const double curw = bd.map[idx] != NONE32 ? bd.kw[bd.map[idx]] : 0.0;
            bc_insert(bd, var, pos, curw + w);
        }
    }
    // pop_index, :1010-1026
    __device__ bool pop_index(uint32_t &v, uint32_t &p) {
        const double total = fl.total + nf.total;
        const double f_ratio = fl.total / total;
        const bool pick_flips = gen_bool(f_ratio);
        if (err) return false;
        BC &bd = pick_flips ? fl : nf;
        const uint32_t i = get_random(bd);
        if (i == NONE32) return false;
        v = bd.kv[i], p = bd.kp ? bd.kp[i] : NONE32;
        const uint32_t idx = bc_idx(v, p);
        RVB_CHK(npopped_pos < 80 && npopped_nopos < 80 && idx < (pick_flips ? fl.maplen : nf.maplen));
        if (pick_flips) pos_popped[idx] = 1, popped_idx[npopped_pos++] = idx;
        else nopos_popped[idx] = 1, popped_idx2[npopped_nopos++] = idx;
        bc_remove(bd, v, p);
        return true;
    }
    // find_overlapping_starts, :1124-1160, feeding push_adjacent
    __device__ void push_overlapping(uint32_t ov, uint32_t p_start, uint32_t p_end, double weight) {
        const uint32_t *fp = constant_ps + var_starts[ov];
        const uint32_t len = var_lengths[ov];
        const uint32_t cutoff = M;  // every position is below the cutoff (< 2^29), so the sums below fit 32 bits
        uint32_t bin_found = 0;  // binary_search(&p_start).unwrap_err()
        while (bin_found < len && fp[bin_found] < p_start) bin_found++;
        const uint32_t prev = bin_found == 0 ? len - 1 : bin_found - 1;  // (bin_found + len - 1) % len
        const uint32_t lowest = fp[prev];
        // (x + cutoff - lowest) % cutoff for x < cutoff: one conditional subtraction
        auto rel = [&](uint32_t x) { const uint32_t y = x + cutoff - lowest; return y >= cutoff ? y - cutoff : y; };
        const uint32_t off_start = rel(p_start), off_end = rel(p_end);
        uint32_t ip = prev;
        for (uint32_t k = 0; k < len; k++, ip = ip + 1 == len ? 0 : ip + 1) {  // [prev..] then [..prev]
            const uint32_t check_start = rel(fp[ip]);
            const uint32_t check_end = rel(fp[ip + 1 == len ? 0 : ip + 1]);
            const bool has_overlap_start = check_start < off_start && off_start < check_end;
            const bool has_start_within = off_start < check_start && check_start < off_end;
            const bool eq = (p_start == p_end) || (check_start == check_end);
            if (!(eq || has_overlap_start || has_start_within)) break;  // take_while
            push_adjacent(ov, ip + var_starts[ov], true, weight);
        }
    }
    // build_cluster, :1054-1122
    __device__ void build_cluster(uint32_t cluster_size, uint32_t init_var, uint32_t init_flip) {
        push_adjacent(init_var, init_flip, false, 0.0);
        while (cluster_size > 0 && !(fl.len == 0 && nf.len == 0)) {
            uint32_t v, flip;
            if (!pop_index(v, flip)) return;
            RVB_CHK(ncl < 72);
            cl_vars[ncl] = v, cl_flips[ncl] = flip, ncl++;
            const uint32_t vs = var_starts[v], vl = var_lengths[v];
            if (flip != NONE32) {
                const uint32_t rel = flip - vs;
                push_adjacent(v, (rel + vl - 1) % vl + vs, false, 0.0);
                push_adjacent(v, (rel + 1) % vl + vs, false, 0.0);
            }
            for (uint32_t k = W.vb_start[v]; k < W.vb_start[v + 1]; k++) {
                const uint32_t b = W.vb_list[k];
                const double weight = fabs(J[b]);  // bond_mag
                const uint32_t ov = other_var(v, b);
                if (var_lengths[ov] == 0) {
                    push_adjacent(ov, NONE32, true, weight);
                } else if (flip != NONE32) {
                    const uint32_t rel = flip - vs, flip_inc = (rel + 1) % vl + vs;
                    push_overlapping(ov, constant_ps[flip], constant_ps[flip_inc], weight);
                } else {
                    for (uint32_t pi = var_starts[ov]; pi < var_starts[ov] + var_lengths[ov]; pi++) push_adjacent(ov, pi, true, weight);
                }
            }
            cluster_size--;
        }
    }
    // calculate_mult, :1194-1221
    __device__ double calculate_mult(uint32_t n) const {
        const bool close = fabs(bef.total - aft.total) < 2.220446049250313e-16;
        if (n == 0 || close) return 1.0;
        return powi_rt(__ddiv_rn(aft.total, bef.total), (int)n);
    }
    // calculate_flip_prob, :649-946
    __device__ double calculate_flip_prob() {
        const double EPS = 2.220446049250313e-16;
        uint32_t cluster_size = 0, next_ci = 0, n_bonds = 0;
        double mult = 1.0;
        for (uint32_t s = 0; s < nsub; s++) cluster_size += cstate[s];
        bc_clear(bef), bc_clear(aft);
        if (cluster_size != 0) {  // set_initial_bonds, :617-646
            for (uint32_t s = 0; s < nsub; s++) {
                if (!cstate[s]) continue;
                const uint32_t v = subvars[s];
                for (uint32_t k = W.vb_start[v]; k < W.vb_start[v + 1]; k++) {
                    const uint32_t b = W.vb_list[k];
                    const uint32_t os = v2s[other_var(v, b)];
                    if (os == NONE32) {  // .unwrap()
                        err |= DEV_ERR_INVARIANT;
                        return 0.0;
                    }
                    if (!cstate[os]) {
                        double wb, wa;
                        ws_for_flip(b, s, wb, wa);
                        bc_insert(bef, b, NONE32, wb), bc_insert(aft, b, NONE32, wa);
                    }
                }
            }
        }
        uint32_t pos = 0;  // every op below pos has left the heap
        seek(0);
        for (;;) {
            Op o;
            const uint32_t q = next_near(pos);  // heap top: the next op on a sub-variable's world line
            if (q == NONE32) break;
            uint32_t p = q;
            if (cluster_size == 0) {  // :722-733: jump to the next cluster flip
                if (next_ci < ntog) p = toggles[next_ci];
                else break;
            }
            if (p > q) {  // popped < p: their outputs propagate the substate, :742-768
                for (uint32_t s = 0; s < nsub; s++) substate[s] = (uint8_t)state_before(subvars[s], p, substate[s]);
                seek(p);
            }
            pos = p + 1;
            if (p >= M || !decode(p, o)) {
                err |= DEV_ERR_INVARIANT;
                return 0.0;
            }
            const bool is_cluster_bound = next_ci < ntog && p == toggles[next_ci];
            const bool will_flip_spins = o.in != o.out;
            const bool will_change_bonds = will_flip_spins || is_cluster_bound;
            bool completely_in_cluster = true;
            for (uint32_t r = 0; r < o.nv; r++) {
                const uint32_t s = v2s[o.v[r]];
                if (s == NONE32 || !cstate[s]) completely_in_cluster = false;
            }
            if (bc_contains(bef, o.b)) {
                n_bonds++;
                continue;
            }
            if (is_cluster_bound) {  // :852-867
                const uint32_t s = v2s[o.v[0]];
                if (s == NONE32) {
                    err |= DEV_ERR_INVARIANT;
                    return 0.0;
                }
                cstate[s] ^= 1;
                if (cstate[s]) cluster_size++;
                else cluster_size--;
                next_ci++;
            }
            if (will_flip_spins)
                for (uint32_t r = 0; r < o.nv; r++)
                    if (v2s[o.v[r]] != NONE32) substate[v2s[o.v[r]]] = (uint8_t)((o.out >> r) & 1u);
            if (completely_in_cluster) {  // ising_ratio, qmc_ising.rs:722-735; rvb_update (:60-79) passes 1.0
                const bool long_bond = D.has_h && o.b >= D.E + D.N;
                mult = mult * (long_bond ? 0.0 : 1.0);
                if (mult < EPS) break;
            }
            if (will_change_bonds) {
                mult = mult * calculate_mult(n_bonds);
                n_bonds = 0;
                if (mult < EPS) break;
                for (uint32_t r = 0; r < o.nv; r++) {  // :901-934
                    const uint32_t v = o.v[r], s = v2s[v];
                    if (s == NONE32) continue;
                    for (uint32_t k = W.vb_start[v]; k < W.vb_start[v + 1]; k++) {
                        const uint32_t b = W.vb_list[k];
                        const uint32_t os = v2s[other_var(v, b)];
                        if (os == NONE32) continue;
                        if (cstate[s] == cstate[os]) {
                            if (bc_contains(bef, b)) bc_remove(bef, b, NONE32), bc_remove(aft, b, NONE32);
                        } else {
                            double wb, wa;
                            ws_for_flip(b, cstate[s] ? s : os, wb, wa);
                            bc_insert(bef, b, NONE32, wb), bc_insert(aft, b, NONE32, wa);
                        }
                    }
                }
            }
        }
        mult = mult * calculate_mult(n_bonds);
        return mult;
    }
    // the "Now update bonds" block of mutate_graph, :561-593
    __device__ void mutate_update_bonds(const Op &o) {
        for (uint32_t r = 0; r < o.nv; r++) {
            const uint32_t v = o.v[r], s = v2s[v];
            if (s == NONE32) continue;
            for (uint32_t k = W.vb_start[v]; k < W.vb_start[v + 1]; k++) {
                const uint32_t b = W.vb_list[k];
                const uint32_t os = v2s[other_var(v, b)];
                if (os == NONE32) continue;
                if (cstate[s] == cstate[os]) {
                    if (bc_contains(bonds, b)) bc_remove(bonds, b, NONE32);
                } else {
                    bc_insert(bonds, b, NONE32, edge_weight_sub(b));
                }
            }
        }
    }
    // mutate_graph, :294-616
    __device__ void mutate_graph() {
        uint32_t jump_to[152], cont_until[152];
        uint32_t nj = 0, nc = 0, count = 0;
        for (uint32_t s = 0; s < nsub; s++) count += cstate[s];
        if (count != 0) {
            jump_to[nj++] = 0;
            for (uint32_t s = 0; s < nsub; s++) substate[s] ^= cstate[s];
        }
        for (uint32_t t = 0; t < ntog; t++) {
            const uint32_t p = toggles[t];
            RVB_CHK(nj < 151 && nc < 151);
            if (count == 0) jump_to[nj++] = p;
            Op o;
            if (!decode(p, o)) {
                err |= DEV_ERR_INVARIANT;
                return;
            }
            for (uint32_t r = 0; r < o.nv; r++) {
                const uint32_t s = v2s[o.v[r]];
                if (s == NONE32) continue;
                cstate[s] ^= 1;
                if (cstate[s]) count++;
                else count--;
            }
            if (count == 0) cont_until[nc++] = p;
        }
        if (count != 0) cont_until[nc++] = M;
        if (nj != nc) {
            err |= DEV_ERR_INVARIANT;
            return;
        }
        bc_clear(bonds);
        uint32_t next_ci = 0;
        for (uint32_t s = 0; s < nsub; s++) {  // :366-381
            if (!cstate[s]) continue;
            const uint32_t v = subvars[s];
            for (uint32_t k = W.vb_start[v]; k < W.vb_start[v + 1]; k++) {
                const uint32_t b = W.vb_list[k];
                const uint32_t os = v2s[other_var(v, b)];
                if (os == NONE32) {
                    err |= DEV_ERR_INVARIANT;
                    return;
                }
                if (!cstate[os]) bc_insert(bonds, b, NONE32, edge_weight_sub(b));
            }
        }
        for (uint32_t k = 0; k < nj && !err; k++) {
            const uint32_t from = jump_to[k], until = cont_until[k];
            // get_propagated_substate_with_hint (fast_ops.rs:1027-1172): the state just before `from`
            for (uint32_t s = 0; s < nsub; s++) {
                const uint32_t v = subvars[s];
                substate[s] = (uint8_t)(state_before(v, from, state_bit(state, v)) ^ cstate[s]);  // :396-399
            }
            // mutate_subsection_ops (fast_ops.rs:639-775): ops with p in [from, until] on the sub-variables' world lines
            seek(from);
            for (uint32_t pos = from; !err;) {
                const uint32_t p = next_near(pos);
                if (p == NONE32 || p > until || p >= M) break;
                pos = p + 1;
                Op o;
                if (!decode(p, o)) {
                    err |= DEV_ERR_INVARIANT;
                    break;
                }
                const bool in_bonds = bc_contains(bonds, o.b);
                const bool at_next = next_ci < ntog && p == toggles[next_ci];
                if (in_bonds) {  // :411-432: rotate the diagonal op onto a bond drawn from the border
                    const uint32_t i = get_random(bonds);
                    if (i == NONE32) break;
                    const uint32_t nb = bonds.kv[i];
                    const uint32_t sa = v2s[__ldg(D.va + nb)], sb = v2s[__ldg(D.vb + nb)];
                    if (sa == NONE32 || sb == NONE32) {
                        err |= DEV_ERR_INVARIANT;
                        break;
                    }
                    const uint32_t st = (uint32_t)substate[sa] | ((uint32_t)substate[sb] << 1);
                    ops[p] = make_op(nb, st, st);
                    if (nb != o.b) {  // the op moves to the world lines of the new bond's variables
                        line_remove(o.v[0], p), line_remove(o.v[1], p);
                        if (!line_insert(__ldg(D.va + nb), p) || !line_insert(__ldg(D.vb + nb), p)) build_lines();
                        seek(pos);
                    }
                    continue;
                }
                uint32_t in = o.in, out = o.out;
                if (at_next) {  // :434-469
                    for (uint32_t r = 0; r < o.nv; r++) {
                        const uint32_t s = v2s[o.v[r]];
                        if (s == NONE32) {
                            err |= DEV_ERR_INVARIANT;
                            break;
                        }
                        in ^= (uint32_t)cstate[s] << r;
                        out ^= (uint32_t)(cstate[s] ^ 1) << r;
                    }
                    if (err) break;
                    for (uint32_t r = 0; r < o.nv; r++) {
                        const uint32_t s = v2s[o.v[r]];
                        cstate[s] ^= 1;
                        substate[s] = (uint8_t)((out >> r) & 1u);
                    }
                    next_ci++;
                } else {  // :470-557
                    bool any_in_cluster = false;
                    for (uint32_t r = 0; r < o.nv; r++) {
                        const uint32_t s = v2s[o.v[r]];
                        if (s != NONE32 && cstate[s]) any_in_cluster = true;
                    }
                    if (!any_in_cluster && in == out) {
                    } else if (any_in_cluster) {
                        const uint32_t mask = o.nv == 2 ? 3u : 1u;
                        for (uint32_t r = 0; r < o.nv; r++)
                            if (v2s[o.v[r]] == NONE32) err |= DEV_ERR_INVARIANT;
                        if (err) break;
                        in ^= mask, out ^= mask;
                        if (in != out)
                            for (uint32_t r = 0; r < o.nv; r++) substate[v2s[o.v[r]]] = (uint8_t)((out >> r) & 1u);
                    } else {
                        uint32_t k2 = 0;  // filter_map(var_to_subvar).zip(outputs): the k-th sub-variable takes output k
                        for (uint32_t r = 0; r < o.nv; r++)
                            if (v2s[o.v[r]] != NONE32) substate[v2s[o.v[r]]] = (uint8_t)((out >> k2++) & 1u);
                    }
                }
                if (in != o.in || out != o.out) ops[p] = make_op(o.b, in, out);
                mutate_update_bonds(o);
            }
        }
    }
};

}  // namespace

// updates < 0: the (N + 1) / 2 of the timestep (qmc_ising.rs:711).  target != 0: only replicas whose next sweep is `target`
// (the diagonal update of that sweep has run; the cluster update follows), and the counters of rvb_success_rate advance.
#ifndef QMCB_RVB_MINB
#define QMCB_RVB_MINB 7  // 72 registers: 4096 replicas are resident in one wave (measured 231 -> 181 ms per sweep of a 4096-replica L = 24 batch)
#endif
__global__ void __launch_bounds__(128, QMCB_RVB_MINB) k_sse_rvb(SseDev D, RvbDev W, uint64_t target, long long updates_arg, unsigned long long *succ_out) {
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= D.R) return;
    if (target != 0 && D.done[r] + 1 != target) return;
    const uint32_t M = D.M[r];
    if (M > D.cap) {
        if (lane == 0) atomicOr(D.status, DEV_ERR_CAPACITY);
        return;
    }
    const uint32_t N = D.N, E = D.E;
    uint32_t *u32 = W.u32 + (size_t)r * W.stride32;
    double *f64 = W.f64 + (size_t)r * W.stride64;
    uint8_t *u8 = W.u8 + (size_t)r * W.stride8;
    const size_t cap = D.cap;
    // workspace layout (see RvbDev in sse.cuh)
    uint32_t *var_starts = u32, *var_lengths = var_starts + (N + 1), *zero_vars = var_lengths + N, *constant_ps = zero_vars + N;
    uint32_t *fl_map = constant_ps + cap, *fl_kv = fl_map + cap, *fl_kp = fl_kv + cap;
    uint32_t *nf_map = fl_kp + cap, *nf_kv = nf_map + N, *subvars = nf_kv + N, *v2s = subvars + N;
    uint32_t *bd_map = v2s + N, *bd_kv = bd_map + 3 * (size_t)E, *fill = bd_kv + 3 * (size_t)E;
    uint32_t *ln_start = fill + N, *ln_len = ln_start + N, *ln_cap = ln_len + N, *c_idx = ln_cap + N, *c_pos = c_idx + N, *lines = c_pos + N;
    double *fl_kw = f64, *nf_kw = fl_kw + cap, *bd_kw = nf_kw + N;
    uint8_t *pos_popped = u8, *nopos_popped = pos_popped + cap, *cstate = nopos_popped + N, *substate = cstate + N, *mark = substate + N;
    // empty containers, nothing popped, no sub-variables (all lanes)
    for (size_t i = lane; i < cap; i += 32) fl_map[i] = NONE32, pos_popped[i] = 0;
    for (uint32_t i = lane; i < N; i += 32) nf_map[i] = NONE32, v2s[i] = NONE32, nopos_popped[i] = 0, mark[i] = 0, var_lengths[i] = 0, ln_len[i] = 0;
    for (uint32_t i = lane; i < 3 * E; i += 32) bd_map[i] = NONE32;
    __syncwarp();
    // find_constants (rvb.rs:1162-1188; constant_ops_on_var fast_ops.rs:1466-1480: the p of every constant op of a variable,
    // ascending) and the world lines, by all lanes: count, lay out, fill (order inside a 32-slot step is whatever the
    // atomics gave), then one lane per variable puts its line in order and copies out its constant ops
    const uint32_t *opw = D.ops + (size_t)r * D.cap;
    for (uint32_t base = 0; base < M; base += 32) {
        const uint32_t p = base + lane;
        const uint32_t w = p < M ? opw[p] : OP_EMPTY;
        if (w != OP_EMPTY) {
            const uint32_t b = op_bond(w);
            const int kind = bond_kind(D, b);
            uint32_t v0, v1;
            bond_vars(D, b, kind, v0, v1);
            atomicAdd(ln_len + v0, 1u);
            if (kind == KIND_BOND) atomicAdd(ln_len + v1, 1u);
            if (kind == KIND_SITE) atomicAdd(var_lengths + v0, 1u);
        }
    }
    __syncwarp();
    uint32_t ncp = 0, nzero = 0;
    if (lane == 0) {
        size_t acc = 0;
        for (uint32_t v = 0; v < N; v++) {
            ln_start[v] = (uint32_t)acc, ln_cap[v] = ln_len[v] + RVB_LINE_SLACK(ln_len[v]);
            acc += ln_cap[v], ln_len[v] = 0;
            var_starts[v] = ncp, ncp += var_lengths[v];
            if (var_lengths[v] == 0) zero_vars[nzero++] = v;
        }
        var_starts[N] = ncp;
    }
    __syncwarp();
    for (uint32_t base = 0; base < M; base += 32) {
        const uint32_t p = base + lane;
        const uint32_t w = p < M ? opw[p] : OP_EMPTY;
        if (w != OP_EMPTY) {
            const uint32_t b = op_bond(w);
            const int kind = bond_kind(D, b);
            uint32_t v0, v1;
            bond_vars(D, b, kind, v0, v1);
            const uint32_t i0 = atomicAdd(ln_len + v0, 1u);
            RVB_CHK(i0 < ln_cap[v0] && ln_start[v0] + i0 < rvb_lines_total(D));
            lines[ln_start[v0] + i0] = p;
            if (kind == KIND_BOND) {
                const uint32_t i1 = atomicAdd(ln_len + v1, 1u);
                RVB_CHK(i1 < ln_cap[v1] && ln_start[v1] + i1 < rvb_lines_total(D));
                lines[ln_start[v1] + i1] = p;
            }
        }
        __syncwarp();  // steps do not interleave: a line is out of order only inside one step's entries
    }
    for (uint32_t v = lane; v < N; v += 32) {
        uint32_t *ln = lines + ln_start[v];
        const uint32_t len = ln_len[v];
        uint32_t k = var_starts[v];
        for (uint32_t i = 0; i < len; i++) {
            const uint32_t x = ln[i];
            uint32_t j = i;
            while (j > 0 && ln[j - 1] > x) ln[j] = ln[j - 1], j--;
            ln[j] = x;
        }
        for (uint32_t i = 0; i < len; i++)
            if (bond_kind(D, op_bond(opw[ln[i]])) == KIND_SITE) {
                RVB_CHK(k < var_starts[v + 1] && k < cap);
                constant_ps[k++] = ln[i];
            }
    }
    __syncwarp();
    if (lane != 0) return;

    Ctx C{D, W};
    C.ops = D.ops + (size_t)r * D.cap, C.state = D.state + (size_t)r * D.Nw;
    C.J = ham_view<true>(D, r).J;
    C.M = M, C.key = D.key[r], C.cur = D.cursor[r], C.err = 0;
    C.var_starts = var_starts, C.var_lengths = var_lengths, C.zero_vars = zero_vars, C.constant_ps = constant_ps;
    C.fl = BC{fl_map, fl_kv, fl_kp, fl_kw, (uint32_t)cap, 0, 0.0};
    C.nf = BC{nf_map, nf_kv, nullptr, nf_kw, N, 0, 0.0};
    C.pos_popped = pos_popped, C.nopos_popped = nopos_popped, C.npopped_pos = 0, C.npopped_nopos = 0;
    C.subvars = subvars, C.v2s = v2s, C.nsub = 0, C.cstate = cstate, C.substate = substate, C.mark = mark;
    C.bonds = BC{bd_map, bd_kv, nullptr, bd_kw, E, 0, 0.0};
    C.bef = BC{bd_map + E, bd_kv + E, nullptr, bd_kw + E, E, 0, 0.0};
    C.aft = BC{bd_map + 2 * (size_t)E, bd_kv + 2 * (size_t)E, nullptr, bd_kw + 2 * (size_t)E, E, 0, 0.0};
    C.ncp = ncp, C.nzero = nzero;
    C.ln_start = ln_start, C.ln_len = ln_len, C.ln_cap = ln_cap, C.lines = lines, C.lines_total = rvb_lines_total(D);
    C.c_idx = c_idx, C.c_pos = c_pos;
    const uint64_t updates = updates_arg >= 0 ? (uint64_t)updates_arg : ((uint64_t)N + 1) / 2;
    unsigned long long num_succ = 0;
    for (uint64_t u = 0; u < updates && !C.err; u++) {
        const uint64_t choice = C.gen_range((uint64_t)ncp + nzero);  // rvb.rs:125
        uint32_t v, flip;
        if (choice < ncp) {  // :126-139: the last variable whose start is <= choice (binary search, then forward over equal starts)
            uint32_t lo = 0, hi = N;  // invariant: var_starts[lo] <= choice; first index with var_starts > choice is in (lo, hi]
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (var_starts[mid] <= choice) lo = mid;
                else hi = mid;
            }
            v = lo, flip = (uint32_t)choice;
        } else {
            v = zero_vars[choice - ncp], flip = NONE32;
        }
        const uint64_t word = C.next_u64();  // contiguous_bits, :1190-1192: trailing_ones
        const uint32_t cluster_size = (~word == 0 ? 64u : (uint32_t)(__ffsll((long long)~word) - 1)) + 1u;
        // fresh boundary manager (:152), fresh var_to_subvar (:166-167)
        bc_clear(C.fl), bc_clear(C.nf);
        for (uint32_t i = 0; i < C.npopped_pos; i++) pos_popped[C.popped_idx[i]] = 0;
        for (uint32_t i = 0; i < C.npopped_nopos; i++) nopos_popped[C.popped_idx2[i]] = 0;
        C.npopped_pos = C.npopped_nopos = 0;
        for (uint32_t s = 0; s < C.nsub; s++) v2s[subvars[s]] = NONE32;
        C.nsub = 0, C.ncl = 0;
        C.build_cluster(cluster_size, v, flip);
        if (C.err) break;
        // dissolve_into (:987-1007) and the sorted, deduplicated sub-variables (:168-180)
        auto add = [&](uint32_t w) {
            RVB_CHK(w < N && C.nsub < N + (mark[w] ? 1u : 0u));
            if (!mark[w]) mark[w] = 1, subvars[C.nsub++] = w;
        };
        for (uint32_t i = 0; i < C.ncl; i++) add(C.cl_vars[i]);
        for (uint32_t i = 0; i < C.fl.len; i++) add(C.fl.kv[i]);
        for (uint32_t i = 0; i < C.nf.len; i++) add(C.nf.kv[i]);
        for (uint32_t i = 1; i < C.nsub; i++) {  // sort_unstable + dedup (the marks kept duplicates out)
            const uint32_t x = subvars[i];
            uint32_t j = i;
            while (j > 0 && subvars[j - 1] > x) subvars[j] = subvars[j - 1], j--;
            subvars[j] = x;
        }
        for (uint32_t s = 0; s < C.nsub; s++) v2s[subvars[s]] = s, mark[subvars[s]] = 0;
        for (uint32_t s = 0; s < C.nsub; s++) cstate[s] = 0, substate[s] = (uint8_t)state_bit(C.state, subvars[s]);
        C.ntog = 0;
        for (uint32_t i = 0; i < C.ncl; i++) {  // :182-203
            const uint32_t cv = C.cl_vars[i], s = v2s[cv], fi = C.cl_flips[i];
            if (fi != NONE32) {
                const uint32_t vstart = var_starts[cv], fi_rel = fi - vstart;
                RVB_CHK(C.ntog + 2 <= 150 && s < C.nsub && fi < ncp);
                if (fi_rel + 1 >= var_lengths[cv]) {
                    cstate[s] = 1;
                    C.toggles[C.ntog++] = constant_ps[fi], C.toggles[C.ntog++] = constant_ps[vstart];
                } else {
                    C.toggles[C.ntog++] = constant_ps[fi], C.toggles[C.ntog++] = constant_ps[fi + 1];
                }
            } else {
                cstate[s] = 1;
            }
        }
        for (uint32_t i = 1; i < C.ntog; i++) {  // sort_unstable (values only: any sort gives the same array)
            const uint32_t x = C.toggles[i];
            uint32_t j = i;
            while (j > 0 && C.toggles[j - 1] > x) C.toggles[j] = C.toggles[j - 1], j--;
            C.toggles[j] = x;
        }
        {  // remove_doubles, vec_help.rs:2-23
            uint32_t ii = 0, jj = 0;
            while (jj + 1 < C.ntog) {
                if (C.toggles[jj] == C.toggles[jj + 1]) jj += 2;
                else C.toggles[ii++] = C.toggles[jj++];
            }
            if (jj < C.ntog) C.toggles[ii++] = C.toggles[jj++];
            C.ntog = ii;
        }
        const double p_to_flip = C.calculate_flip_prob();
        if (C.err) break;
        const bool should_mutate = p_to_flip >= 1.0 ? true : C.gen_bool(p_to_flip);  // :247-252
        if (C.err) break;
        if (should_mutate) {
            C.mutate_graph();
            bool starting = false;
            for (uint32_t s = 0; s < C.nsub; s++) starting |= cstate[s] != 0;
            if (starting)
                for (uint32_t s = 0; s < C.nsub; s++)
                    if (cstate[s]) C.state[subvars[s] >> 5] ^= 1u << (subvars[s] & 31);
            num_succ++;
        }
    }
    D.cursor[r] = C.cur;
    if (C.err) atomicOr(D.status, C.err);
#ifdef QMCB_RVB_CHECK
    if (g_rvb_check_fail) atomicOr(D.status, DEV_ERR_STACK);
#endif
    if (succ_out) succ_out[r] = num_succ;
    if (target != 0) W.succ[r] += num_succ, W.count[r] += updates;
}

int launch_sse_rvb(const SseDev &D, const RvbDev &W, uint64_t target, long long updates, unsigned long long *succ_out, cudaStream_t st) {
    const uint32_t wpb = 4;
    k_sse_rvb<<<(D.R + wpb - 1) / wpb, 32 * wpb, 0, st>>>(D, W, target, updates, succ_out);
    return 1;
}
