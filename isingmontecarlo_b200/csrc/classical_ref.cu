// classical_ref.cu -- the reference's OWN classical schedule, replica-parallel: GraphState::do_time_step
// (classical/graph.rs:350-406) with its three moves -- random-site spin flips (:91-119), edge flips with optional
// importance sampling (:122-153, :321-336) and worm flips (:179-318) -- under each replica's sequential stream.
//
// These moves are serial by construction (every attempt reads the spins the previous one wrote, the site is a draw),
// so one thread owns one replica and the batch is what runs in parallel; the checkerboard kernels of classical.cu are
// the throughput path.  Spins are the reference's Vec<bool> (one byte per spin).  Bit-exact with the oracle
// (oracle.c: orc_cls_spin_flips / _edge_flips / _worm_flips / _do_time_step) up to the last-place rounding of
// exp(): should_flip compares a 53-bit uniform with exp(-beta dE), and the device exp may differ from libm's in the
// last place, which changes a decision with probability ~2^-52 per draw.
#include "classical.cuh"

#define WM_NONE 0xFFFFFFFFu
#define CLS_REF_MAXDEG 12
#define CLS_REF_STACK (CLS_REF_MAXDEG * (CLS_REF_MAXDEG + 1))
#define EPS_F64 2.220446049250313e-16

namespace {

struct Rng {
    uint64_t key, cur;
    __device__ __forceinline__ uint64_t next() { return stream_word(key, cur++); }
    // rand 0.8 UniformInt<usize>::sample_single: widening multiply + approximate zone
    __device__ uint64_t range_usize(uint64_t range) {
        const uint64_t zone = (range << __clzll((long long)range)) - 1ull;
        for (;;) {
            const uint64_t v = next();
            const uint64_t hi = __umul64hi(v, range), lo = v * range;
            if (lo <= zone) return hi;
        }
    }
    // rand 0.8 UniformInt<u8>::sample_single: widened to u32 (the high half of a stream word), exact zone
    __device__ uint32_t range_u8(uint32_t range) {
        const uint32_t reject = (0xFFFFFFFFu - range + 1u) % range, zone = 0xFFFFFFFFu - reject;
        for (;;) {
            const uint32_t v = (uint32_t)(next() >> 32);
            const uint64_t m = (uint64_t)v * range;
            if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
        }
    }
    // rand 0.8 UniformFloat<f64>::sample_single(0.0, high)
    __device__ double range_f64(double high) {
        for (;;) {
            const double v01 = __longlong_as_double((long long)((next() >> 12) | 0x3FF0000000000000ull)) - 1.0;
            const double res = v01 * high + 0.0;
            if (res < high) return res;
        }
    }
    // rand 0.8 Standard f64: 53 bits * 2^-53
    __device__ __forceinline__ double f64() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

struct Worm {
    uint32_t a, b;  // Single(a) = (a, NONE), Double(a, b)
};

struct Ctx {
    const ClsRefDev &D;
    uint8_t *sp;
    Rng g;
    double beta;
    // GraphState::delta_e :155-176 (no bias term; the edge to `omit` left out)
    __device__ double delta_e(uint32_t v, uint32_t omit) const {
        const uint32_t curr = sp[v];
        double de = 0.0;
        for (uint32_t k = D.adj_start[v]; k < D.adj_start[v + 1]; k++) {
            const uint32_t o = D.adj_idx[k];
            if (o == omit) continue;
            de += -2.0 * D.adj_j[k] * (curr == sp[o] ? 1.0 : -1.0);
        }
        return de;
    }
    __device__ double worm_de(Worm m) const {  // :191-199
        if (m.b == WM_NONE) return delta_e(m.a, WM_NONE);
        const double de = delta_e(m.a, m.b);
        return de + delta_e(m.b, m.a);
    }
    __device__ bool should_flip(double de) {  // :339-347
        if (de > 0.0) {
            const double chance = exp(-beta * de);
            return g.f64() < chance;
        }
        return true;
    }
    __device__ void spin_flip() {  // :91-119
        const uint32_t i = (uint32_t)g.range_usize(D.N);
        const double de = delta_e(i, WM_NONE) + (2.0 * D.biases[i] * (sp[i] ? 1.0 : -1.0));
        if (should_flip(de)) sp[i] ^= 1u;
    }
    __device__ bool edge_flip() {  // :122-153
        uint32_t e;
        if (D.cum_w) {
            if (!(0.0 < D.total_w)) return false;
            const double p = g.range_f64(D.total_w);
            uint32_t size = D.E, left = 0, right = D.E;  // slice::binary_search_by, Ok(i) | Err(i) -> i
            e = NONE32;
            while (left < right) {
                const uint32_t mid = left + size / 2;
                const double c = D.cum_w[mid];
                if (c < p) left = mid + 1;
                else if (c > p) right = mid;
                else { e = mid; break; }
                size = right - left;
            }
            if (e == NONE32) e = left;
        } else {
            if (D.E == 0) return false;
            e = (uint32_t)g.range_usize(D.E);
        }
        if (e >= D.E) return false;
        const uint32_t va = D.ea[e], vb = D.eb[e];
        const double da = delta_e(va, vb) + (2.0 * D.biases[va] * (sp[va] ? 1.0 : -1.0));
        const double db = delta_e(vb, va) + (2.0 * D.biases[vb] * (sp[vb] ? 1.0 : -1.0));
        if (should_flip(da + db)) sp[va] ^= 1u, sp[vb] ^= 1u;
        return true;
    }
    __device__ void worm_flip(uint32_t *path, bool allow_doubles) {  // :179-318
        Worm sm[CLS_REF_STACK];
        double sde[CLS_REF_STACK];
        const uint32_t N = D.N;
        const uint32_t start = (uint32_t)g.range_usize(N);
        uint64_t plen = 0;
        auto ppush = [&](Worm m) { path[2 * plen] = m.a, path[2 * plen + 1] = m.b, plen++; };
        ppush(Worm{start, WM_NONE});
        uint32_t last_index = start;
        const double starting_e = worm_de(Worm{start, WM_NONE});
        sp[start] ^= 1u;
        bool failed = false;
        for (;;) {
            uint32_t ns = 0;
            const Worm sel{path[2 * (plen - 1)], path[2 * (plen - 1) + 1]};
            const uint32_t sel_var = sel.b == WM_NONE ? sel.a : sel.b;
            bool any_resolve = false;
            for (uint32_t k = D.adj_start[sel_var]; k < D.adj_start[sel_var + 1]; k++) {  // :214-242
                const uint32_t ov = D.adj_idx[k];
                if (ov == last_index) continue;
                const double de = delta_e(ov, WM_NONE);
                if (fabs(de) < EPS_F64) sm[ns] = Worm{ov, WM_NONE}, sde[ns++] = de;
                else if (fabs(de + starting_e) < EPS_F64) sm[ns] = Worm{ov, WM_NONE}, sde[ns++] = de, any_resolve = true;
                if (allow_doubles) {
                    sp[ov] ^= 1u;
                    for (uint32_t kk = D.adj_start[ov]; kk < D.adj_start[ov + 1]; kk++) {
                        const uint32_t oov = D.adj_idx[kk];
                        if (oov != ov && oov != sel_var) {
                            const double de2 = delta_e(oov, WM_NONE) + de;
                            if (fabs(de2) < EPS_F64) sm[ns] = Worm{ov, oov}, sde[ns++] = de2;
                            else if (fabs(de2 + starting_e) < EPS_F64) sm[ns] = Worm{ov, oov}, sde[ns++] = de2, any_resolve = true;
                        }
                    }
                    sp[ov] ^= 1u;
                }
            }
            if (any_resolve) {  // :243-245
                uint32_t w = 0;
                for (uint32_t i = 0; i < ns; i++)
                    if (fabs(sde[i] + starting_e) < EPS_F64) sm[w] = sm[i], sde[w++] = sde[i];
                ns = w;
            }
            Worm ov;
            double de;
            if (ns) {  // :247-251
                const uint32_t choice = (uint32_t)g.range_usize(ns);
                ov = sm[choice], de = sde[choice];
                ppush(ov);
            } else {  // no options: turn around, undo the last move :252-262
                ov = sel.b == WM_NONE ? sel : Worm{sel.b, sel.a};
                ppush(ov);
                de = worm_de(ov);
            }
            sp[ov.a] ^= 1u;
            if (ov.b != WM_NONE) sp[ov.b] ^= 1u;
            last_index = ov.b != WM_NONE ? ov.a : sel_var;  // :272-276
            if (fabs(de + starting_e) < EPS_F64) break;
            if (plen > N) {  // :282-285
                failed = true;
                break;
            }
        }
        // :287-298 flatten (in place: the write index never passes the read index), sort, drop pairs
        uint64_t nf = 0;
        for (uint64_t i = 0; i < plen; i++) {
            const uint32_t a = path[2 * i], b = path[2 * i + 1];
            path[nf++] = a;
            if (b != WM_NONE) path[nf++] = b;
        }
        // heap sort (sort_unstable of integers: any correct sort gives the same vector)
        auto sift = [&](uint64_t root, uint64_t end) {
            for (;;) {
                uint64_t child = 2 * root + 1;
                if (child >= end) break;
                if (child + 1 < end && path[child] < path[child + 1]) child++;
                if (path[root] >= path[child]) break;
                const uint32_t t = path[root];
                path[root] = path[child], path[child] = t;
                root = child;
            }
        };
        for (uint64_t i = nf / 2; i-- > 0;) sift(i, nf);
        for (uint64_t end = nf; end > 1;) {
            end--;
            const uint32_t t = path[0];
            path[0] = path[end], path[end] = t;
            sift(0, end);
        }
        uint64_t ii = 0, jj = 0;  // util/vec_help.rs:2-24
        while (jj + 1 < nf) {
            if (path[jj] == path[jj + 1]) jj += 2;
            else path[ii++] = path[jj++];
        }
        if (jj < nf) path[ii++] = path[jj++];
        nf = ii;
        bool undo = failed;
        if (!failed) {  // :300-311
            double total_he = 0.0;
            for (uint64_t i = 0; i < nf; i++) total_he += 2.0 * D.biases[path[i]] * (sp[path[i]] ? 1.0 : -1.0);
            undo = !should_flip(total_he);
        }
        if (undo)
            for (uint64_t i = 0; i < nf; i++) sp[path[i]] ^= 1u;
    }
};

}  // namespace

// move: 0 spin flips, 1 edge flips, 2 worm flips, 3 do_time_step (one u8 draw picks the move, :364-366)
__global__ void __launch_bounds__(32) k_cls_ref_moves(ClsRefDev D, int move, uint64_t nspin, uint64_t nedge, uint64_t nworm,
                                                      int only_basic, int allow_doubles, uint8_t *choice_out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= D.R) return;
    Ctx c{D, D.spins + (size_t)r * D.N, Rng{D.key[r], D.cursor[r]}, D.beta[r]};
    int m = move;
    if (move == 3) {
        m = (int)c.g.range_u8(only_basic ? 2u : 3u);
        allow_doubles = 1;  // :396
    }
    if (choice_out) choice_out[r] = (uint8_t)m;
    bool ok = true;
    if (m == 0) {
        for (uint64_t t = 0; t < nspin; t++) c.spin_flip();
    } else if (m == 1) {
        for (uint64_t t = 0; t < nedge && ok; t++) ok = c.edge_flip();
    } else {
        uint32_t *path = D.path + (size_t)r * 2 * ((size_t)D.N + 2);
        for (uint64_t t = 0; t < nworm; t++) c.worm_flip(path, allow_doubles != 0);
    }
    if (!ok) atomicOr(D.status, DEV_ERR_PROB);  // the reference panics (empty range / no edges)
    D.cursor[r] = c.g.cur;
}

void launch_cls_ref_moves(const ClsRefDev &D, int move, uint64_t nspin, uint64_t nedge, uint64_t nworm, int only_basic,
                          int allow_doubles, uint8_t *choice_out, cudaStream_t st) {
    k_cls_ref_moves<<<(D.R + 31) / 32, 32, 0, st>>>(D, move, nspin, nedge, nworm, only_basic, allow_doubles, choice_out);
}
