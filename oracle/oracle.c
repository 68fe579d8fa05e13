/*
 * oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h).  "parity unpinned" at the rand-0.8 boundary (no Rust toolchain here).
 *
 * All `file:line` citations are into the reference tree (Renmusxd/IsingMonteCarlo,
 * crate qmc 2.20.0).  The data-structure shape follows the reference on purpose (one
 * node per slot holding its operator and doubly linked p / per-variable links, LIFO
 * DFS for clusters) so that this file can also be timed as the CPU baseline.
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ===================================================================================
 * RNG contract: Philox4x32-10, one 64-bit word per draw
 * =================================================================================== */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        if (r > 0) {
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0, c1 = n1, c2 = n2, c3 = n3;
    }
    out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

/* W[c]: 128-bit counter = c >> 1, word = (x[2(c&1)+1] << 32) | x[2(c&1)] */
uint64_t orc_stream_word(uint64_t key, uint64_t cursor) {
    uint64_t blk = cursor >> 1;
    uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), 0u, 0u};
    uint32_t k[2] = {(uint32_t)key, (uint32_t)(key >> 32)};
    uint32_t x[4];
    orc_philox4x32_10(ctr, k, x);
    unsigned h = (unsigned)(cursor & 1u) * 2u;
    return ((uint64_t)x[h + 1] << 32) | x[h];
}

typedef struct {
    uint64_t key, cursor;
    const uint64_t *script; /* scripted words for known-answer tests (or NULL) */
    uint64_t script_len;
    int error;
    /* last Philox block (two stream words), so that a block is computed once, not twice */
    uint64_t blk_key, blk_idx, blk_w[2];
    int blk_valid;
    /* TIMING ONLY (bench.py's CPU arm): the reference's own benches seed rand's SmallRng (benches/end_to_end.rs:49),
     * which is xoshiro256++ on 64-bit targets -- two orders of magnitude cheaper per word than Philox4x32-10.  With
     * small_rng set the words come from xoshiro256++ (state seeded from the key by SplitMix64, as rand_xoshiro's
     * seed_from_u64 does); the cursor still counts words, but such a stream has no GPU counterpart and no test uses it. */
    int small_rng;
    uint64_t xs[4];
} Stream;

static uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static void small_rng_seed(Stream *s, uint64_t seed) {
    for (int i = 0; i < 4; i++) { /* SplitMix64 */
        seed += 0x9E3779B97F4A7C15ull;
        uint64_t z = seed;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        s->xs[i] = z ^ (z >> 31);
    }
    s->small_rng = 1;
}

static uint64_t next_u64(Stream *s) {
    if (s->script) {
        if (s->cursor >= s->script_len) {
            s->error = 1;
            s->cursor++;
            return 0;
        }
        return s->script[s->cursor++];
    }
    if (s->small_rng) { /* xoshiro256++ */
        uint64_t *x = s->xs;
        const uint64_t result = rotl64(x[0] + x[3], 23) + x[0], t = x[1] << 17;
        x[2] ^= x[0], x[3] ^= x[1], x[1] ^= x[2], x[0] ^= x[3], x[2] ^= t, x[3] = rotl64(x[3], 45);
        s->cursor++;
        return result;
    }
    {
        const uint64_t c = s->cursor++, blk = c >> 1;
        if (!s->blk_valid || s->blk_idx != blk || s->blk_key != s->key) {
            uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), 0u, 0u};
            uint32_t k[2] = {(uint32_t)s->key, (uint32_t)(s->key >> 32)};
            uint32_t x[4];
            orc_philox4x32_10(ctr, k, x);
            s->blk_w[0] = ((uint64_t)x[1] << 32) | x[0], s->blk_w[1] = ((uint64_t)x[3] << 32) | x[2];
            s->blk_idx = blk, s->blk_key = s->key, s->blk_valid = 1;
        }
        return s->blk_w[c & 1u];
    }
}
/* one stream word per call; a u32 is the word's high half (Appendix A.3) */
static uint32_t next_u32(Stream *s) { return (uint32_t)(next_u64(s) >> 32); }

uint64_t orc_bool_threshold(double p) { return (uint64_t)(p * 18446744073709551616.0); }

/* rand 0.8 Bernoulli (Rng::gen_bool): p==1 -> true without a draw; else v < (p*2^64) as u64 */
static int gen_bool(Stream *s, double p) {
    if (!(p >= 0.0 && p < 1.0)) {
        if (p == 1.0) return 1;
        s->error = 2; /* the reference panics here */
        return 0;
    }
    uint64_t p_int = orc_bool_threshold(p);
    uint64_t v = next_u64(s);
    return v < p_int;
}

/* rand 0.8 UniformInt<usize>::sample_single_inclusive (64-bit, widening multiply + zone) */
static uint64_t gen_range_usize(Stream *s, uint64_t range) {
    uint64_t zone = (range << __builtin_clzll(range)) - 1u;
    for (;;) {
        uint64_t v = next_u64(s);
        unsigned __int128 m = (unsigned __int128)v * range;
        uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
        if (lo <= zone) return hi;
    }
}

/* rand 0.8 UniformInt<u8>: widened to u32, exact rejection zone */
static uint32_t gen_range_u8(Stream *s, uint32_t range) {
    uint32_t ints_to_reject = (0xFFFFFFFFu - range + 1u) % range;
    uint32_t zone = 0xFFFFFFFFu - ints_to_reject;
    for (;;) {
        uint32_t v = next_u32(s);
        uint64_t m = (uint64_t)v * range;
        uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
        if (lo <= zone) return hi;
    }
}

/* rand 0.8 UniformFloat<f64>::sample_single for 0.0..1.0 (52-bit mantissa fill) */
static double gen_range_f64_01(Stream *s) {
    for (;;) {
        uint64_t v = next_u64(s);
        uint64_t bits = (v >> 12) | 0x3FF0000000000000ull;
        double value1_2;
        memcpy(&value1_2, &bits, 8);
        double value0_1 = value1_2 - 1.0;
        double res = value0_1 * 1.0 + 0.0;
        if (res < 1.0) return res;
    }
}

/* rand 0.8 UniformFloat<f64>::sample_single for low..high (finite scale): value0_1 * scale + low,
 * redrawn while the rounded result reaches `high` */
static double gen_range_f64(Stream *s, double low, double high) {
    const double scale = high - low;
    for (;;) {
        uint64_t v = next_u64(s);
        uint64_t bits = (v >> 12) | 0x3FF0000000000000ull;
        double value1_2;
        memcpy(&value1_2, &bits, 8);
        double value0_1 = value1_2 - 1.0;
        double res = value0_1 * scale + low;
        if (res < high) return res;
        if (s->error) return low; /* scripted stream ran dry */
    }
}

/* rand 0.8 Standard f64: 53 random bits * 2^-53 */
static double gen_f64(Stream *s) {
    uint64_t v = next_u64(s) >> 11;
    return (double)v * (1.0 / 9007199254740992.0);
}

/* rand 0.8 Standard bool: sign bit of a u32 */
static int gen_std_bool(Stream *s) { return (int32_t)next_u32(s) < 0; }

int orc_gen_bool(uint64_t key, uint64_t *cursor, double p) {
    Stream s = {key, *cursor, NULL, 0, 0, 0, 0, {0, 0}, 0, 0, {0, 0, 0, 0}};
    int r = gen_bool(&s, p);
    *cursor = s.cursor;
    return r;
}
uint64_t orc_gen_range_usize(uint64_t key, uint64_t *cursor, uint64_t n) {
    Stream s = {key, *cursor, NULL, 0, 0, 0, 0, {0, 0}, 0, 0, {0, 0, 0, 0}};
    uint64_t r = gen_range_usize(&s, n);
    *cursor = s.cursor;
    return r;
}
uint32_t orc_gen_range_u8(uint64_t key, uint64_t *cursor, uint32_t n) {
    Stream s = {key, *cursor, NULL, 0, 0, 0, 0, {0, 0}, 0, 0, {0, 0, 0, 0}};
    uint32_t r = gen_range_u8(&s, n);
    *cursor = s.cursor;
    return r;
}
double orc_gen_range_f64_01(uint64_t key, uint64_t *cursor) {
    Stream s = {key, *cursor, NULL, 0, 0, 0, 0, {0, 0}, 0, 0, {0, 0, 0, 0}};
    double r = gen_range_f64_01(&s);
    *cursor = s.cursor;
    return r;
}
double orc_gen_range_f64(uint64_t key, uint64_t *cursor, double low, double high) {
    Stream s = {key, *cursor, NULL, 0, 0, 0, 0, {0, 0}, 0, 0, {0, 0, 0, 0}};
    double r = gen_range_f64(&s, low, high);
    *cursor = s.cursor;
    return r;
}
double orc_gen_f64(uint64_t key, uint64_t *cursor) {
    Stream s = {key, *cursor, NULL, 0, 0, 0, 0, {0, 0}, 0, 0, {0, 0, 0, 0}};
    double r = gen_f64(&s);
    *cursor = s.cursor;
    return r;
}
int orc_gen_std_bool(uint64_t key, uint64_t *cursor) {
    Stream s = {key, *cursor, NULL, 0, 0, 0, 0, {0, 0}, 0, 0, {0, 0, 0, 0}};
    int r = gen_std_bool(&s);
    *cursor = s.cursor;
    return r;
}

/* f64::powi lowers to compiler-rt __powidf2: square-and-multiply, reciprocal at the end */
double orc_powi(double a, int b) {
    const int recip = b < 0;
    double r = 1.0;
    for (;;) {
        if (b & 1) r *= a;
        b /= 2;
        if (b == 0) break;
        a *= a;
    }
    return recip ? 1.0 / r : r;
}

/* ===================================================================================
 * SSE replica: storage (op_container.rs:224-237, fast_ops.rs:35-49,181-190)
 * =================================================================================== */

#define SIDE_IN 0
#define SIDE_OUT 1
#define NONE (-1)

typedef struct {
    uint8_t present; /* Option<Node> */
    uint8_t nv;      /* vars.len(): 1 or 2 */
    uint8_t constant;
    uint8_t in[2], out[2];
    uint32_t bond;
    uint32_t vars[2];
    int64_t prev_p, next_p;         /* previous_p / next_p */
    int64_t prev_vp[2], next_vp[2]; /* previous_for_vars / next_for_vars: PRel.p */
    int8_t prev_vr[2], next_vr[2];  /*                                     PRel.relv */
} Node;

typedef struct {
    int64_t *a;
    int8_t *b;
    int8_t *c;
    uint64_t len, cap;
} Stack; /* StackTuplizer (util/allocator.rs:73-139): LIFO of tuples */

static void stack_push(Stack *s, int64_t a, int b, int c) {
    if (s->len == s->cap) {
        s->cap = s->cap ? 2 * s->cap : 1024;
        s->a = (int64_t *)realloc(s->a, s->cap * sizeof(int64_t));
        s->b = (int8_t *)realloc(s->b, s->cap);
        s->c = (int8_t *)realloc(s->c, s->cap);
    }
    s->a[s->len] = a, s->b[s->len] = (int8_t)b, s->c[s->len] = (int8_t)c;
    s->len++;
}
static void stack_free(Stack *s) {
    free(s->a), free(s->b), free(s->c);
    memset(s, 0, sizeof(*s));
}

struct OrcSse {
    uint32_t nvars, nedges;
    uint32_t *ea, *eb;
    double *J;
    double transverse, longitudinal;
    double offset; /* total_energy_offset, qmc_ising.rs:97-99 */
    uint64_t cutoff; /* QmcIsingGraph::cutoff */
    Node *ops;       /* FastOps::ops */
    uint64_t ops_len;
    uint64_t n;
    int64_t first_p, last_p;   /* p_ends */
    int64_t *vfirst_p, *vlast_p; /* var_ends */
    int8_t *vfirst_r, *vlast_r;
    uint8_t *state;
    Stream rng;
    /* cluster scratch (boundaries StackTuplizer, cluster.rs:50-52) */
    int64_t *b_in, *b_out;
    uint64_t b_len, b_cap;
    Stack frontier, interior;
    /* fast mode scratch */
    uint32_t *uf;
    uint64_t uf_cap;
    /* heat-bath diagonal update: BondWeights (heatbath.rs:10-13), NULL = Metropolis rule */
    double *hb_maxw, *hb_cum;
    int error;
    /* generic Qmc (qmc_runner.rs:22-45): a list of interactions instead of (edges, transverse, longitudinal) */
    struct OrcInteraction *inter;
    uint32_t ninter;
    int is_qmc, has_cluster_edges, breaks_ising_symmetry, do_loop_updates;
    /* RVB update (qmc_ising.rs:39-43) */
    int run_rvb;
    uint64_t total_rvb_successes, rvb_clusters_counted;
};

/* Interaction, qmc_runner.rs:406-421: type Full(constant) | Diagonal, mat, n, vars, constant_along_diagonal */
typedef struct OrcInteraction {
    int diagonal;  /* InteractionType::Diagonal */
    int constant;  /* InteractionType::Full(true) */
    int constant_along_diagonal;
    uint32_t n, vars[2];
    double mat[16];
    uint32_t len;
} OrcInteraction;

static uint32_t num_bonds(const OrcSse *g) {
    if (g->is_qmc) return g->ninter; /* qmc_runner.rs:177 */
    /* qmc_ising.rs:664-670 */
    return g->nedges + g->nvars + (fabs(g->longitudinal) > DBL_EPSILON ? g->nvars : 0u);
}

/* bonds_fn, qmc_ising.rs:671-681 */
static void edge_fn(const OrcSse *g, uint32_t b, uint32_t vars[2], int *nv, int *constant) {
    if (g->is_qmc) { /* bonds_fn, qmc_runner.rs:178 */
        const OrcInteraction *it = &g->inter[b];
        vars[0] = it->vars[0], vars[1] = it->vars[1], *nv = (int)it->n, *constant = it->constant;
        return;
    }
    if (b < g->nedges) {
        vars[0] = g->ea[b], vars[1] = g->eb[b], *nv = 2, *constant = 0;
    } else if (b < g->nedges + g->nvars) {
        vars[0] = b - g->nedges, *nv = 1, *constant = 1;
    } else {
        vars[0] = b - g->nvars - g->nedges, *nv = 1, *constant = 0;
    }
}

/* qmc_ising.rs:863-875 */
static double two_site_hamiltonian(int i0, int i1, int o0, int o1, double bond) {
    if (i0 == o0 && i1 == o1) {
        double t;
        if (!i0 && !i1) t = -bond;
        else if (!i0 && i1) t = bond;
        else if (i0 && !i1) t = bond;
        else t = -bond;
        return fabs(bond) + t;
    }
    return 0.0;
}
/* qmc_ising.rs:877-879 */
static double transverse_hamiltonian(int in, int out, double transverse) {
    (void)in, (void)out;
    return transverse;
}
/* qmc_ising.rs:881-888 */
static double longitudinal_hamiltonian(int in, int out, double longitudinal) {
    double t;
    if (in != out) t = 0.0;
    else if (in) t = longitudinal;
    else t = -longitudinal;
    return fabs(longitudinal) + t;
}
/* QmcIsingGraph::hamiltonian, qmc_ising.rs:179-205 */
static double interaction_at(const OrcInteraction *it, const uint8_t *in, const uint8_t *out);
static double hamiltonian(const OrcSse *g, uint32_t bond, const uint8_t *in, const uint8_t *out) {
    if (g->is_qmc) return interaction_at(&g->inter[bond], in, out); /* qmc_runner.rs:172-174 */
    if (bond < g->nedges) return two_site_hamiltonian(in[0], in[1], out[0], out[1], g->J[bond]);
    if (bond < g->nedges + g->nvars) return transverse_hamiltonian(in[0], out[0], g->transverse);
    return longitudinal_hamiltonian(in[0], out[0], g->longitudinal);
}

static int node_is_diagonal(const Node *nd) {
    for (int r = 0; r < nd->nv; r++)
        if (nd->in[r] != nd->out[r]) return 0;
    return 1;
}
/* cluster.rs:284-286 */
static int is_edge(const Node *nd) { return nd->constant && nd->nv == 1; }

static void ops_resize(OrcSse *g, uint64_t len) {
    /* FastOps::set_cutoff, fast_ops.rs:1259-1263 / mutate_subsection fast_ops.rs:622-624 */
    if (len > g->ops_len) {
        g->ops = (Node *)realloc(g->ops, len * sizeof(Node));
        memset(g->ops + g->ops_len, 0, (len - g->ops_len) * sizeof(Node));
        g->ops_len = len;
    }
}

OrcSse *orc_sse_create(uint32_t nvars, uint32_t nedges, const uint32_t *ea, const uint32_t *eb,
                       const double *J, double transverse, double longitudinal, uint64_t cutoff,
                       uint64_t rng_key, const uint8_t *state_or_null) {
    /* new_with_rng_with_manager_hook, qmc_ising.rs:80-128 */
    OrcSse *g = (OrcSse *)calloc(1, sizeof(OrcSse));
    g->nvars = nvars, g->nedges = nedges;
    g->ea = (uint32_t *)malloc(sizeof(uint32_t) * (nedges ? nedges : 1));
    g->eb = (uint32_t *)malloc(sizeof(uint32_t) * (nedges ? nedges : 1));
    g->J = (double *)malloc(sizeof(double) * (nedges ? nedges : 1));
    memcpy(g->ea, ea, sizeof(uint32_t) * nedges);
    memcpy(g->eb, eb, sizeof(uint32_t) * nedges);
    memcpy(g->J, J, sizeof(double) * nedges);
    g->transverse = transverse, g->longitudinal = longitudinal;
    double edge_offset = 0.0;
    for (uint32_t e = 0; e < nedges; e++) edge_offset += fabs(J[e]);
    double field_offset = (double)nvars * (transverse + fabs(longitudinal));
    g->offset = edge_offset + field_offset;
    g->cutoff = cutoff;
    ops_resize(g, cutoff);
    g->first_p = g->last_p = NONE;
    g->vfirst_p = (int64_t *)malloc(sizeof(int64_t) * nvars);
    g->vlast_p = (int64_t *)malloc(sizeof(int64_t) * nvars);
    g->vfirst_r = (int8_t *)malloc(nvars);
    g->vlast_r = (int8_t *)malloc(nvars);
    for (uint32_t v = 0; v < nvars; v++) g->vfirst_p[v] = g->vlast_p[v] = NONE;
    g->rng.key = rng_key;
    g->state = (uint8_t *)malloc(nvars);
    if (state_or_null) {
        memcpy(g->state, state_or_null, nvars);
    } else {
        /* make_random_spin_state, classical/graph.rs:451-453 */
        for (uint32_t v = 0; v < nvars; v++) g->state[v] = (uint8_t)gen_std_bool(&g->rng);
    }
    return g;
}

void orc_sse_destroy(OrcSse *g) {
    if (!g) return;
    free(g->ea), free(g->eb), free(g->J), free(g->ops);
    free(g->vfirst_p), free(g->vlast_p), free(g->vfirst_r), free(g->vlast_r);
    free(g->state), free(g->b_in), free(g->b_out), free(g->uf);
    free(g->hb_maxw), free(g->hb_cum), free(g->inter);
    stack_free(&g->frontier), stack_free(&g->interior);
    free(g);
}

/* ===================================================================================
 * Generic Qmc: qmc_runner.rs:46-156 (construction, interactions), :363-377 (timestep), :406-680 (Interaction)
 * =================================================================================== */
/* index_from_iter over outputs then inputs, last bit least significant (qmc_runner.rs:650-664) */
static uint32_t index_from_state(const uint8_t *in, const uint8_t *out, uint32_t n, int with_outputs) {
    uint32_t acc = 0;
    if (with_outputs)
        for (uint32_t k = 0; k < n; k++) acc = (acc << 1) | (out[k] ? 1u : 0u);
    for (uint32_t k = 0; k < n; k++) acc = (acc << 1) | (in[k] ? 1u : 0u);
    return acc;
}
/* Interaction::at, qmc_runner.rs:560-600 */
static double interaction_at(const OrcInteraction *it, const uint8_t *in, const uint8_t *out) {
    if (!it->diagonal && it->constant) return it->mat[0];
    if (!it->diagonal) return it->mat[index_from_state(in, out, it->n, 1)];
    for (uint32_t k = 0; k < it->n; k++)
        if (in[k] != out[k]) return 0.0;
    return it->mat[index_from_state(in, out, it->n, 0)];
}
static int all_equal(const double *v, uint32_t len, uint32_t stride) {
    for (uint32_t k = 1; k < len; k++) /* try_fold: every item against the one before it, |old - item| < EPSILON */
        if (!(fabs(v[(k - 1) * stride] - v[k * stride]) < DBL_EPSILON)) return 0;
    return 1;
}
/* Interaction::sym_under_ising, qmc_runner.rs:626-648 */
static int sym_under_ising(const OrcInteraction *it) {
    if (!it->diagonal && it->constant) return 1;
    if (it->diagonal && it->constant_along_diagonal) return 1;
    if (!it->diagonal) {
        const uint32_t mask = (1u << (it->n << 1)) - 1u;
        for (uint32_t x = 0; x < (1u << it->n); x++)
            if (!(fabs(it->mat[x] - it->mat[(~x) & mask]) < DBL_EPSILON)) return 0;
        return 1;
    }
    const uint32_t mask = (1u << it->n) - 1u;
    for (uint32_t x = 0; x < (1u << (it->n >> 1)); x++)
        if (!(fabs(it->mat[x] - it->mat[(~x) & mask]) < DBL_EPSILON)) return 0;
    return 1;
}

/* Qmc::new_with_state (qmc_runner.rs:54-87): cutoff = nvars; state drawn from the stream when NULL (Qmc::new :48-51) */
OrcSse *orc_qmc_create(uint32_t nvars, uint64_t rng_key, const uint8_t *state_or_null) {
    uint32_t dummy = 0;
    double dj = 0.0;
    OrcSse *g = orc_sse_create(nvars, 0, &dummy, &dummy, &dj, 0.0, 0.0, nvars, rng_key, state_or_null);
    g->is_qmc = 1, g->offset = 0.0;
    return g;
}
/* make_interaction / make_interaction_and_offset / make_diagonal_interaction / make_diagonal_interaction_and_offset
 * (qmc_runner.rs:113-156) with Interaction::new / new_offset / new_diagonal / new_diagonal_offset (:424-558).
 * Returns 0, or 1 "Matrix size must be power of 2", 2 "Given x vars, expected n", 3 "Interaction contains negative
 * weights", 4 more than two variables (not restated). */
int orc_qmc_make_interaction(OrcSse *g, const double *mat, uint32_t len, const uint32_t *vars, uint32_t nvars_given, int diagonal, int and_offset) {
    OrcInteraction it;
    memset(&it, 0, sizeof it);
    uint32_t pw = 0;
    while ((1u << pw) < len) pw++;
    if ((1u << pw) != len || len > 16) return len > 16 ? 4 : 1;
    it.n = diagonal ? pw : pw >> 1;
    if (nvars_given > 2 || it.n > 2) return 4;
    memcpy(it.mat, mat, sizeof(double) * len);
    it.len = len, it.diagonal = diagonal;
    double min_diag = 0.0;
    if (and_offset) { /* new_offset :508-521 / new_diagonal_offset :424-436 */
        const uint32_t tn = 1u << it.n, stride = diagonal ? 1u : tn + 1u;
        min_diag = DBL_MAX;
        for (uint32_t k = 0; k < tn; k++) min_diag = min_diag < it.mat[k * stride] ? min_diag : it.mat[k * stride];
        for (uint32_t k = 0; k < tn; k++) it.mat[k * stride] -= min_diag;
    }
    if (!diagonal) {
        for (uint32_t k = 0; k < len; k++)
            if (it.mat[k] < 0.0) return 3;
        if (it.n != nvars_given) return 2;
        it.constant = all_equal(it.mat, len, 1);
        it.constant_along_diagonal = all_equal(it.mat, 1u << it.n, (1u << it.n) + 1u);
    } else {
        it.constant_along_diagonal = all_equal(it.mat, len, 1);
        if (it.n != nvars_given) return 2;
    }
    for (uint32_t k = 0; k < nvars_given; k++) it.vars[k] = vars[k];
    /* add_interaction :90-105 */
    if (it.constant && it.n == 1) g->has_cluster_edges = 1;
    if (!sym_under_ising(&it)) g->breaks_ising_symmetry = 1;
    g->inter = (OrcInteraction *)realloc(g->inter, sizeof(OrcInteraction) * (g->ninter + 1));
    g->inter[g->ninter++] = it;
    free(g->hb_maxw), free(g->hb_cum);
    g->hb_maxw = g->hb_cum = NULL;
    if (and_offset) g->offset -= min_diag;
    return 0;
}
int orc_qmc_flags(const OrcSse *g) { return (g->has_cluster_edges ? 1 : 0) | (g->breaks_ising_symmetry ? 2 : 0); }

void orc_sse_set_script(OrcSse *g, const uint64_t *words, uint64_t nwords) {
    g->rng.script = words, g->rng.script_len = nwords, g->rng.cursor = 0;
}
int orc_sse_error(const OrcSse *g) { return g->error | (g->rng.error << 8); }

/* Rebuild every link from the op array.  Semantically what the incremental splices in
 * FastOps::mutate_p maintain (fast_ops.rs:337-607; same construction as
 * clear_and_install_ops fast_ops.rs:89-173). */
static void links_begin(OrcSse *g) {
    for (uint32_t v = 0; v < g->nvars; v++) g->vfirst_p[v] = g->vlast_p[v] = NONE;
    g->first_p = g->last_p = NONE;
}
/* append the op in slot p (the highest occupied slot so far) to the p list and to the lists of its variables */
static inline void links_append(OrcSse *g, uint64_t p) {
    Node *nd = &g->ops[p];
    nd->prev_p = g->last_p, nd->next_p = NONE;
    if (g->last_p != NONE) g->ops[g->last_p].next_p = (int64_t)p;
    else g->first_p = (int64_t)p;
    g->last_p = (int64_t)p;
    for (int r = 0; r < nd->nv; r++) {
        uint32_t v = nd->vars[r];
        nd->next_vp[r] = NONE, nd->next_vr[r] = 0;
        if (g->vlast_p[v] != NONE) {
            Node *pn = &g->ops[g->vlast_p[v]];
            pn->next_vp[g->vlast_r[v]] = (int64_t)p, pn->next_vr[g->vlast_r[v]] = (int8_t)r;
            nd->prev_vp[r] = g->vlast_p[v], nd->prev_vr[r] = g->vlast_r[v];
        } else {
            nd->prev_vp[r] = NONE, nd->prev_vr[r] = 0;
            g->vfirst_p[v] = (int64_t)p, g->vfirst_r[v] = (int8_t)r;
        }
        g->vlast_p[v] = (int64_t)p, g->vlast_r[v] = (int8_t)r;
    }
}
static void rebuild_links(OrcSse *g) {
    links_begin(g);
    uint64_t n = 0;
    for (uint64_t p = 0; p < g->ops_len; p++) {
        if (!g->ops[p].present) continue;
        n++;
        links_append(g, p);
    }
    g->n = n;
}

/* ===================================================================================
 * Diagonal update: diagonal.rs:114-135 (driver), :142-191 (rule); loop fast_ops.rs:611-637
 * =================================================================================== */
static void diagonal_update(OrcSse *g, double beta) {
    const uint64_t cutoff = g->cutoff;
    const uint32_t nb = num_bonds(g);
    ops_resize(g, cutoff); /* fast_ops.rs:622-624 */
    uint8_t *state = g->state;
    uint64_t n = g->n; /* s.get_n(): live */
    /* the links (what mutate_p's splices maintain, fast_ops.rs:337-607) are rebuilt in the same pass: the op that ends up in
     * slot p is appended to the lists right away, so the sweep reads the string once, as the reference does */
    links_begin(g);
    for (uint64_t p = 0; p < cutoff; p++) {
        Node *nd = &g->ops[p];
        uint32_t b;
        if (!nd->present) {
            b = (uint32_t)gen_range_usize(&g->rng, nb); /* diagonal.rs:152 */
        } else if (node_is_diagonal(nd)) {
            b = nd->bond; /* :153 */
        } else {
            for (int r = 0; r < nd->nv; r++) state[nd->vars[r]] = nd->out[r]; /* :154-160 */
            links_append(g, p);
            continue;
        }
        uint32_t vars[2];
        int nv, constant;
        edge_fn(g, b, vars, &nv, &constant);
        uint8_t sub[2] = {0, 0};
        for (int r = 0; r < nv; r++) sub[r] = state[vars[r]];
        double mat_element = hamiltonian(g, b, sub, sub);
        double numerator = beta * (double)nb * mat_element; /* :168 */
        double denominator = (double)(cutoff - n);         /* :169 */
        if (!nd->present) {
            if (numerator > denominator || gen_bool(&g->rng, numerator / denominator)) { /* :173 */
                nd->present = 1, nd->nv = (uint8_t)nv, nd->constant = (uint8_t)constant, nd->bond = b;
                for (int r = 0; r < nv; r++)
                    nd->vars[r] = vars[r], nd->in[r] = sub[r], nd->out[r] = sub[r];
                n++;
            }
        } else {
            denominator = denominator + 1.0; /* :182 */
            if (denominator > numerator || gen_bool(&g->rng, denominator / numerator)) { /* :183 */
                nd->present = 0;
                n--;
            }
        }
        if (nd->present) links_append(g, p);
    }
    for (uint64_t p = cutoff; p < g->ops_len; p++) /* (slots above the cutoff hold no ops: set_cutoff refuses to orphan any) */
        if (g->ops[p].present) links_append(g, p);
    g->n = n;
}

/* ===================================================================================
 * Diagonal update, COUNTER mode (builder-defined draw source): the rule of diagonal.rs:142-191 -- the
 * same numerator / denominator arithmetic, the same live n -- but the uniform words of slot p do not come from the
 * sequential stream: they are the two 64-bit halves of ONE Philox block per slot,
 *     (x, y, z, w) = Philox4x32-10(ctr = (p, c_lo, c_hi, 'DIAG'), key),  wA = y << 32 | x,  wB = w << 32 | z,
 * c = stream cursor at the start of the step (the step then advances the cursor by 1, like the FAST cluster step).
 *   empty slot:   b = (wA * Nb) >> 64 (multiply-shift without rand's zone rejection: non-uniformity <= Nb * 2^-64);
 *                 insert iff num > den, or num / den == 1.0, or wB < (num / den * 2^64) as u64
 *   diagonal op:  remove iff den + 1 > num, or (den + 1) / num == 1.0, or wA < ((den + 1) / num * 2^64) as u64
 *   off-diagonal: state[vars] = outputs (no draw)
 * so that every slot can be decided independently once n before it is known.
 * =================================================================================== */
#define TAG_DIAG 0x44494147u

static int bernoulli_word(OrcSse *g, uint64_t word, double p) { /* gen_bool(p) fed with a given word */
    if (!(p >= 0.0 && p < 1.0)) {
        if (p == 1.0) return 1;
        g->rng.error = 2; /* the reference panics here */
        return 0;
    }
    return word < orc_bool_threshold(p);
}

static void diagonal_update_counter(OrcSse *g, double beta) {
    const uint64_t cutoff = g->cutoff;
    const uint32_t nb = num_bonds(g);
    ops_resize(g, cutoff);
    uint8_t *state = g->state;
    uint64_t n = g->n;
    const uint64_t c0 = g->rng.cursor;
    const uint32_t k[2] = {(uint32_t)g->rng.key, (uint32_t)(g->rng.key >> 32)};
    for (uint64_t p = 0; p < cutoff; p++) {
        Node *nd = &g->ops[p];
        if (nd->present && !node_is_diagonal(nd)) {
            for (int r = 0; r < nd->nv; r++) state[nd->vars[r]] = nd->out[r];
            continue;
        }
        const uint32_t ctr[4] = {(uint32_t)p, (uint32_t)c0, (uint32_t)(c0 >> 32), TAG_DIAG};
        uint32_t x[4];
        orc_philox4x32_10(ctr, k, x);
        const uint64_t wA = ((uint64_t)x[1] << 32) | x[0], wB = ((uint64_t)x[3] << 32) | x[2];
        const uint32_t b = nd->present ? nd->bond : (uint32_t)(((unsigned __int128)wA * nb) >> 64);
        uint32_t vars[2];
        int nv, constant;
        edge_fn(g, b, vars, &nv, &constant);
        uint8_t sub[2] = {0, 0};
        for (int r = 0; r < nv; r++) sub[r] = state[vars[r]];
        const double numerator = beta * (double)nb * hamiltonian(g, b, sub, sub);
        double denominator = (double)(cutoff - n);
        if (!nd->present) {
            if (numerator > denominator || bernoulli_word(g, wB, numerator / denominator)) {
                nd->present = 1, nd->nv = (uint8_t)nv, nd->constant = (uint8_t)constant, nd->bond = b;
                for (int r = 0; r < nv; r++)
                    nd->vars[r] = vars[r], nd->in[r] = sub[r], nd->out[r] = sub[r];
                n++;
            }
        } else {
            denominator = denominator + 1.0;
            if (denominator > numerator || bernoulli_word(g, wA, denominator / numerator)) {
                nd->present = 0;
                n--;
            }
        }
    }
    g->n = n;
    g->rng.cursor = c0 + 1;
    rebuild_links(g);
}

/* ===================================================================================
 * Heat-bath diagonal update: heatbath.rs:106-127 (driver), :149-209 (rule), BondWeights :10-61;
 * enabled by QmcIsingGraph::set_enable_heatbath (qmc_ising.rs:444-486)
 * =================================================================================== */
void orc_sse_set_enable_heatbath(OrcSse *g, int enable) {
    free(g->hb_maxw), free(g->hb_cum);
    g->hb_maxw = g->hb_cum = NULL;
    if (!enable) return;
    const uint32_t nb = num_bonds(g);
    g->hb_maxw = (double *)malloc(sizeof(double) * nb);
    g->hb_cum = (double *)malloc(sizeof(double) * nb);
    for (uint32_t b = 0; b < nb; b++) { /* make_bond_weights, heatbath.rs:130-146 */
        uint32_t vars[2];
        int nv, constant;
        edge_fn(g, b, vars, &nv, &constant);
        double acc = 0.0;
        for (int sub = 0; sub < (1 << nv); sub++) {
            uint8_t bits[2] = {(uint8_t)(sub & 1), (uint8_t)((sub >> 1) & 1)};
            double w = hamiltonian(g, b, bits, bits);
            if (w > acc) acc = w;
        }
        g->hb_maxw[b] = acc;
        g->hb_cum[b] = b == 0 ? acc : acc + g->hb_cum[b - 1]; /* BondWeights::new :24-30 */
    }
}
int orc_sse_get_enable_heatbath(const OrcSse *g) { return g->hb_cum != NULL; }

/* BondWeights::index_for_cumulative (heatbath.rs:56-60): slice::binary_search_by, then the insertion
 * point when there is no exact match (core::slice, the size-halving loop) */
static uint32_t hb_index_for_cumulative(const double *cum, uint32_t len, double val) {
    uint32_t size = len, left = 0, right = len;
    while (left < right) {
        uint32_t mid = left + size / 2;
        if (cum[mid] < val) left = mid + 1;
        else if (cum[mid] > val) right = mid;
        else return mid;
        size = right - left;
    }
    return left;
}

static void heatbath_diagonal_update(OrcSse *g, double beta) {
    const uint64_t cutoff = g->cutoff;
    const uint32_t nb = num_bonds(g);
    const double total = g->hb_cum[nb - 1]; /* BondWeights::total :50-54 */
    ops_resize(g, cutoff);
    uint8_t *state = g->state;
    uint64_t n = g->n;
    for (uint64_t p = 0; p < cutoff; p++) {
        Node *nd = &g->ops[p];
        if (!nd->present) { /* heatbath.rs:162-191 */
            double numerator = beta * total;
            double denominator = (double)(cutoff - n) + numerator;
            if (gen_bool(&g->rng, numerator / denominator)) {
                double pdraw = gen_range_f64(&g->rng, 0.0, 1.0);  /* :167 */
                double c = gen_range_f64(&g->rng, 0.0, total);    /* :38 */
                uint32_t b = hb_index_for_cumulative(g->hb_cum, nb, c);
                if (b >= nb) { g->error |= 4; continue; } /* reference: index out of bounds panic */
                double maxweight = g->hb_maxw[b];
                uint32_t vars[2];
                int nv, constant;
                edge_fn(g, b, vars, &nv, &constant);
                uint8_t sub[2] = {0, 0};
                for (int r = 0; r < nv; r++) sub[r] = state[vars[r]];
                double weight = hamiltonian(g, b, sub, sub);
                if (pdraw * maxweight < weight) { /* :181 */
                    nd->present = 1, nd->nv = (uint8_t)nv, nd->constant = (uint8_t)constant, nd->bond = b;
                    for (int r = 0; r < nv; r++)
                        nd->vars[r] = vars[r], nd->in[r] = sub[r], nd->out[r] = sub[r];
                    n++;
                }
            }
        } else if (node_is_diagonal(nd)) { /* :192-201 */
            double numerator = (double)(cutoff - n + 1);
            double denominator = numerator + beta * total;
            if (gen_bool(&g->rng, numerator / denominator)) {
                nd->present = 0;
                n--;
            }
        } else { /* :203-208 */
            for (int r = 0; r < nd->nv; r++) state[nd->vars[r]] = nd->out[r];
        }
    }
    g->n = n;
    rebuild_links(g);
}

/* the diagonal step of timestep / single_diagonal_step: qmc_ising.rs:250-268, :685-703 */
static void diagonal_step(OrcSse *g, double beta, int mode) {
    if (mode == ORC_MODE_COUNTER) {
        if (g->hb_cum) g->error |= 16; /* the heat-bath rule has no COUNTER-mode contract */
        else diagonal_update_counter(g, beta);
    } else if (g->hb_cum) heatbath_diagonal_update(g, beta);
    else diagonal_update(g, beta);
}

/* ===================================================================================
 * Cluster update, reference order: cluster.rs:36-172, :193-271, :289-306
 * =================================================================================== */
static void boundaries_resize(OrcSse *g, uint64_t len) {
    if (len > g->b_cap) {
        g->b_cap = len + len / 2 + 16;
        g->b_in = (int64_t *)realloc(g->b_in, g->b_cap * sizeof(int64_t));
        g->b_out = (int64_t *)realloc(g->b_out, g->b_cap * sizeof(int64_t));
    }
    g->b_len = len;
    for (uint64_t p = 0; p < len; p++) g->b_in[p] = g->b_out[p] = NONE;
}

/* set_boundary, cluster.rs:289-306.  Returns true if both sides now hold a cluster. */
static int set_boundary(OrcSse *g, int64_t p, int side, int64_t c) {
    int64_t *slot = side == SIDE_IN ? &g->b_in[p] : &g->b_out[p];
    if (*slot == NONE || *slot == c) *slot = c;
    else g->error = 3; /* unreachable!() in the reference */
    return g->b_in[p] != NONE && g->b_out[p] != NONE;
}

/* expand_whole_cluster, cluster.rs:193-271 */
static void expand_whole_cluster(OrcSse *g, int64_t p0, int relv0, int side0, int64_t c) {
    Stack *in = &g->interior;
    in->len = 0;
    const Node *op = &g->ops[p0];
    if (!is_edge(op)) {
        for (int r = 0; r < op->nv; r++) stack_push(in, p0, r, SIDE_IN); /* :205-211 */
        for (int r = 0; r < op->nv; r++) stack_push(in, p0, r, SIDE_OUT);
    } else {
        stack_push(in, p0, relv0, side0); /* :212-215 */
    }
    while (in->len) {
        in->len--;
        int64_t p = in->a[in->len];
        int relvar = in->b[in->len], side = in->c[in->len];
        set_boundary(g, p, side, c); /* :218 */
        const Node *nd = &g->ops[p];
        uint32_t var = nd->vars[relvar];
        int64_t q;
        int rq, sq;
        if (side == SIDE_IN) { /* :224-232 */
            q = nd->prev_vp[relvar], rq = nd->prev_vr[relvar];
            if (q == NONE) q = g->vlast_p[var], rq = g->vlast_r[var];
            sq = SIDE_OUT;
        } else { /* :233-241 */
            q = nd->next_vp[relvar], rq = nd->next_vr[relvar];
            if (q == NONE) q = g->vfirst_p[var], rq = g->vfirst_r[var];
            sq = SIDE_IN;
        }
        const Node *qn = &g->ops[q];
        if (is_edge(qn)) { /* :245-248 */
            if (!set_boundary(g, q, sq, c)) stack_push(&g->frontier, q, 0, sq == SIDE_IN ? SIDE_OUT : SIDE_IN);
        } else { /* :249-268 */
            int64_t a = g->b_in[q], b = g->b_out[q];
            int ok = (a == NONE && b == NONE) || (a == c && b == NONE) || (a == NONE && b == c);
            if (ok) {
                g->b_in[q] = c, g->b_out[q] = c; /* set_boundaries :308-315 */
                for (int r = 0; r < qn->nv; r++)
                    if (!(r == rq && sq == SIDE_IN)) stack_push(in, q, r, SIDE_IN);
                for (int r = 0; r < qn->nv; r++)
                    if (!(r == rq && sq == SIDE_OUT)) stack_push(in, q, r, SIDE_OUT);
            }
        }
    }
}

/* flip_each_cluster_rng, cluster.rs:36-172; `has_weights` selects the h != 0 closure of
 * qmc_ising.rs:754-776 (longitudinal ops give weight 0.0, everything else 1.0). */
static uint64_t cluster_update_strict(OrcSse *g, int has_weights) {
    if (g->n == 0) return 0; /* :46-48 */
    const int64_t last_p = g->last_p;
    boundaries_resize(g, (uint64_t)last_p + 1);
    /* find_constant_op :175-186 */
    int64_t cp = g->first_p;
    while (cp != NONE && !is_edge(&g->ops[cp])) cp = g->ops[cp].next_p;
    uint64_t n_clusters;
    if (cp != NONE) {
        Stack *fr = &g->frontier;
        fr->len = 0;
        stack_push(fr, cp, 0, SIDE_OUT); /* :57-59 */
        stack_push(fr, cp, 0, SIDE_IN);
        int64_t cluster_num = 0;
        uint64_t scan = 0; /* smallest unmapped p is monotone, so the :82-88 scan can resume */
        for (;;) {
            while (fr->len) { /* :62-80 */
                fr->len--;
                int64_t p = fr->a[fr->len];
                int side = fr->c[fr->len];
                if (g->b_in[p] != NONE && g->b_out[p] != NONE) continue;
                expand_whole_cluster(g, p, 0, side, cluster_num);
                cluster_num++;
            }
            int64_t unmapped = NONE; /* :82-88 */
            for (; scan <= (uint64_t)last_p; scan++) {
                if (g->ops[scan].present && g->b_in[scan] == NONE && g->b_out[scan] == NONE) {
                    unmapped = (int64_t)scan;
                    break;
                }
            }
            if (unmapped == NONE) break;
            stack_push(fr, unmapped, 0, SIDE_OUT); /* :89-91 */
            stack_push(fr, unmapped, 0, SIDE_IN);
        }
        n_clusters = (uint64_t)cluster_num;
    } else { /* :98-107 */
        for (int64_t p = 0; p <= last_p; p++)
            if (g->ops[p].present) g->b_in[p] = g->b_out[p] = 0;
        n_clusters = 1;
    }

    uint8_t *flips = (uint8_t *)malloc(n_clusters ? n_clusters : 1);
    if (has_weights) { /* :111-136 */
        double *w = (double *)malloc(sizeof(double) * n_clusters);
        for (uint64_t k = 0; k < n_clusters; k++) w[k] = 1.0;
        for (int64_t p = 0; p <= last_p; p++) {
            if (g->b_in[p] == NONE) continue;
            if (g->b_in[p] == g->b_out[p]) {
                double f = g->ops[p].bond >= g->nedges + g->nvars ? 0.0 : 1.0; /* qmc_ising.rs:759-775 */
                w[g->b_in[p]] *= f;
            }
        }
        for (uint64_t k = 0; k < n_clusters; k++) flips[k] = (uint8_t)gen_bool(&g->rng, w[k] * 0.5);
        free(w);
    } else { /* :137 */
        for (uint64_t k = 0; k < n_clusters; k++) flips[k] = (uint8_t)gen_bool(&g->rng, 0.5);
    }
    for (int64_t p = 0; p <= last_p; p++) { /* :139-167 */
        if (g->b_in[p] == NONE) continue;
        Node *nd = &g->ops[p];
        if (flips[g->b_in[p]]) {
            for (int r = 0; r < nd->nv; r++) nd->in[r] = !nd->in[r];
            for (int r = 0; r < nd->nv; r++)
                if (nd->prev_vp[r] == NONE) g->state[nd->vars[r]] = nd->in[r];
        }
        if (flips[g->b_out[p]])
            for (int r = 0; r < nd->nv; r++) nd->out[r] = !nd->out[r];
    }
    free(flips);
    return n_clusters;
}

/* ===================================================================================
 * Cluster update, canonical ("fast") order -- builder-defined contract, same Markov
 * kernel as cluster.rs:36-172 but order-independent so that it can be labelled with a
 * parallel union-find (DESIGN.md "fast mode"):
 *   * site ops (constant, one variable; cluster.rs:284-286) cut world lines into segments.
 *     ids: v in [0,N) = the segment of variable v crossing p = 0;  N + k = the segment that
 *     starts at the output of the k-th site op (p order).
 *   * every two-variable op joins the two segments it touches; the segment after the last
 *     site op of v is joined with segment v (periodic closure).
 *   * cluster root = smallest segment id, in FAST and COUNTER mode (a largest-id contract was tried for the GPU's
 *     cache behaviour and dropped; uf_union keeps the switch).  No site op at all => one cluster (cluster.rs:98-107).
 *   * a cluster holding a longitudinal op never flips (weight 0, qmc_ising.rs:759-775);
 *     otherwise flip = bit (root & 127) of Philox(key, ctr = (root >> 7, c_lo, c_hi, 'CLUS')),
 *     c = stream cursor at the start of the step; the step then advances the cursor by 1.
 * =================================================================================== */
#define TAG_CLUS 0x434C5553u

static uint32_t uf_find(uint32_t *uf, uint32_t x) {
    while (uf[x] != x) {
        uf[x] = uf[uf[x]];
        x = uf[x];
    }
    return x;
}
/* maxroot = 0: the root of a set is its smallest id (FAST contract); 1: its largest id (COUNTER contract) */
static void uf_union(uint32_t *uf, uint32_t a, uint32_t b, int maxroot) {
    a = uf_find(uf, a), b = uf_find(uf, b);
    if (a == b) return;
    if ((a < b) != (maxroot != 0)) uf[b] = a;
    else uf[a] = b;
}

static int fast_flip_bit(uint64_t key, uint64_t cursor, uint32_t root) {
    uint32_t ctr[4] = {root >> 7, (uint32_t)cursor, (uint32_t)(cursor >> 32), TAG_CLUS};
    uint32_t k[2] = {(uint32_t)key, (uint32_t)(key >> 32)};
    uint32_t x[4];
    orc_philox4x32_10(ctr, k, x);
    return (x[(root >> 5) & 3] >> (root & 31)) & 1u;
}

static uint64_t cluster_update_fast(OrcSse *g, int has_weights, int maxroot) {
    if (g->n == 0) return 0;
    const uint32_t N = g->nvars;
    const int64_t last_p = g->last_p;
    boundaries_resize(g, (uint64_t)last_p + 1);
    uint64_t need = (uint64_t)N + g->n + 1;
    if (need > g->uf_cap) {
        g->uf_cap = need + need / 2;
        g->uf = (uint32_t *)realloc(g->uf, g->uf_cap * sizeof(uint32_t));
    }
    uint32_t *uf = g->uf;
    uint32_t *cur = (uint32_t *)malloc(sizeof(uint32_t) * N);
    for (uint32_t v = 0; v < N; v++) uf[v] = v, cur[v] = v;
    uint32_t nsite = 0;
    /* pass 1: segment ids per leg (stored in b_in/b_out) and unions */
    for (int64_t p = 0; p <= last_p; p++) {
        const Node *nd = &g->ops[p];
        if (!nd->present) continue;
        if (is_edge(nd)) {
            uint32_t v = nd->vars[0];
            g->b_in[p] = cur[v];
            uint32_t id = N + nsite++;
            uf[id] = id;
            cur[v] = id;
            g->b_out[p] = id;
        } else {
            g->b_in[p] = g->b_out[p] = cur[nd->vars[0]];
            if (nd->nv == 2) uf_union(uf, cur[nd->vars[0]], cur[nd->vars[1]], maxroot);
        }
    }
    for (uint32_t v = 0; v < N; v++) uf_union(uf, v, cur[v], maxroot); /* periodic closure */
    const uint32_t nseg = N + nsite;
    if (nsite == 0) /* cluster.rs:98-107: no cluster edge => everything is one cluster */
        for (uint32_t x = 0; x < nseg; x++) uf[x] = maxroot ? nseg - 1 : 0;
    /* frozen clusters + cluster count (roots that own at least one leg) */
    uint8_t *frozen = (uint8_t *)calloc(nseg, 1), *used = (uint8_t *)calloc(nseg, 1);
    for (int64_t p = 0; p <= last_p; p++) {
        const Node *nd = &g->ops[p];
        if (!nd->present) continue;
        uint32_t ri = uf_find(uf, (uint32_t)g->b_in[p]), ro = uf_find(uf, (uint32_t)g->b_out[p]);
        used[ri] = used[ro] = 1;
        if (has_weights && nd->bond >= g->nedges + N) frozen[ri] = 1;
    }
    uint64_t n_clusters = 0;
    for (uint32_t x = 0; x < nseg; x++) n_clusters += used[x];
    const uint64_t c0 = g->rng.cursor;
    /* apply: same edits as cluster.rs:139-167 */
    for (int64_t p = 0; p <= last_p; p++) {
        Node *nd = &g->ops[p];
        if (!nd->present) continue;
        uint32_t ri = uf_find(uf, (uint32_t)g->b_in[p]), ro = uf_find(uf, (uint32_t)g->b_out[p]);
        g->b_in[p] = ri, g->b_out[p] = ro;
        int fi = !frozen[ri] && fast_flip_bit(g->rng.key, c0, ri);
        int fo = !frozen[ro] && fast_flip_bit(g->rng.key, c0, ro);
        if (fi) {
            for (int r = 0; r < nd->nv; r++) nd->in[r] = !nd->in[r];
            for (int r = 0; r < nd->nv; r++)
                if (nd->prev_vp[r] == NONE) g->state[nd->vars[r]] = nd->in[r];
        }
        if (fo)
            for (int r = 0; r < nd->nv; r++) nd->out[r] = !nd->out[r];
    }
    g->rng.cursor = c0 + 1;
    free(frozen), free(used), free(cur);
    return n_clusters;
}

/* single_cluster_step qmc_ising.rs:273-320 / timestep :754-784 */
static uint64_t cluster_and_free_spins(OrcSse *g, int mode) {
    int has_weights = fabs(g->longitudinal) > DBL_EPSILON;
    uint64_t ncl = mode != ORC_MODE_STRICT ? cluster_update_fast(g, has_weights, 0)
                                         : cluster_update_strict(g, has_weights);
    for (uint32_t v = 0; v < g->nvars; v++) /* qmc_ising.rs:780-784 */
        if (g->vfirst_p[v] == NONE) g->state[v] = (uint8_t)gen_bool(&g->rng, 0.5);
    return ncl;
}

/* ===================================================================================
 * Directed-loop update: directed_loop.rs:103-171 (make_loop_update_with_rng, initial_n = None as Qmc::loop_update
 * qmc_runner.rs:205-220 calls it), :183-211 (apply_loop_update), :214-301 (loop_body).  A leg is (relative variable,
 * side); the weight of leaving through a leg is the matrix element of the op with the entrance leg and that leg
 * flipped (adjust_states, qmc_types.rs:28-37; entering and leaving through the same leg is the bounce).
 * =================================================================================== */
static void flip_leg(uint8_t *in, uint8_t *out, int var, int side) {
    if (side == SIDE_IN) in[var] ^= 1;
    else out[var] ^= 1;
}
void orc_qmc_loop_update(OrcSse *g) {
    if (g->n == 0) return; /* :139; post_loop_update_hook (:174) is empty */
    const uint64_t initial_n = gen_range_usize(&g->rng, g->n); /* :140-142 */
    int64_t p = g->first_p;                                     /* get_nth_p :76-87 */
    for (uint64_t i = 0; i < initial_n; i++) p = g->ops[p].next_p;
    const int64_t p0 = p;
    const int v0 = (int)gen_range_usize(&g->rng, g->ops[p].nv);      /* :147 */
    const int s0 = gen_std_bool(&g->rng) ? SIDE_IN : SIDE_OUT;       /* :148-152 */
    int64_t sel = p0;
    int ev = v0, es = s0;
    for (;;) { /* :195-210 */
        Node *op = &g->ops[sel];
        const int nv = op->nv, nlegs = 2 * nv;
        double w[4], total = 0.0;
        for (int k = 0; k < nlegs; k++) { /* inputs legs, then outputs legs :231-240 */
            uint8_t in[2] = {op->in[0], op->in[1]}, out[2] = {op->out[0], op->out[1]};
            flip_leg(in, out, ev, es);
            flip_leg(in, out, k % nv, k < nv ? SIDE_IN : SIDE_OUT);
            w[k] = hamiltonian(g, op->bond, in, out);
            total = total + w[k]; /* :242 */
        }
        if (!(0.0 < total) || !isfinite(total)) { /* gen_range(0. ..total) panics on an empty or unbounded range */
            g->error |= 32;
            return;
        }
        double c = gen_range_f64(&g->rng, 0.0, total); /* :243 */
        int ex = -1;
        for (int k = 0; k < nlegs; k++) { /* try_fold :244-253 */
            if (c < w[k]) {
                ex = k;
                break;
            }
            c = c - w[k];
        }
        if (ex < 0) { /* unwrap_err() on Ok: rounding left the choice beyond the last leg */
            g->error |= 32;
            return;
        }
        const int xv = ex % nv, xs = ex < nv ? SIDE_IN : SIDE_OUT;
        flip_leg(op->in, op->out, ev, es); /* :256-259 */
        flip_leg(op->in, op->out, xv, xs);
        if (sel == p0 && xv == v0 && xs == s0) return; /* :266-267 */
        int64_t q;
        int rq;
        const uint32_t var = op->vars[xv];
        if (xs == SIDE_OUT) { /* :274-281 */
            q = op->next_vp[xv], rq = op->next_vr[xv];
            if (q == NONE) {
                g->state[var] = op->out[xv];
                q = g->vfirst_p[var], rq = g->vfirst_r[var];
            }
        } else { /* :282-289 */
            q = op->prev_vp[xv], rq = op->prev_vr[xv];
            if (q == NONE) {
                g->state[var] = op->in[xv];
                q = g->vlast_p[var], rq = g->vlast_r[var];
            }
        }
        const int ns = xs == SIDE_OUT ? SIDE_IN : SIDE_OUT; /* :291 */
        if (q == p0 && rq == v0 && ns == s0) return;        /* :293-294 */
        sel = q, ev = rq, es = ns;
    }
}
void orc_qmc_set_do_loop_updates(OrcSse *g, int enable) { g->do_loop_updates = enable != 0; } /* qmc_runner.rs:268-270 */

/* Qmc::timestep, qmc_runner.rs:363-377: diagonal_update (:158-201, the cutoff grows there), loop_update when
 * do_loop_updates (:366-368), cluster_update with Ising symmetry and no weights when should_do_cluster_update (:223-238,
 * :278-281), flip_free_bits (:241-256). */
void orc_qmc_timestep(OrcSse *g, double beta, int mode) {
    diagonal_step(g, beta, mode);
    uint64_t grown = g->n + g->n / 2; /* :195 */
    if (grown > g->cutoff) g->cutoff = grown;
    if (g->do_loop_updates) orc_qmc_loop_update(g);
    if (!g->breaks_ising_symmetry && g->has_cluster_edges) {
        if (mode != ORC_MODE_STRICT) cluster_update_fast(g, 0, 0);
        else cluster_update_strict(g, 0);
    }
    for (uint32_t v = 0; v < g->nvars; v++)
        if (g->vfirst_p[v] == NONE) g->state[v] = (uint8_t)gen_bool(&g->rng, 0.5);
}


/* ===================================================================================
 * RVB update: rvb.rs:60-291 (RvbUpdater::rvb_update_with_ising_weight), :294-616 (mutate_graph), :617-646
 * (set_initial_bonds), :649-946 (calculate_flip_prob), :955-1052 (VarPos, WeightedBoundaryManager), :1054-1122
 * (build_cluster), :1124-1160 (find_overlapping_starts), :1162-1188 (find_constants), :1190-1192 (contiguous_bits),
 * :1194-1221 (calculate_mult); util/bondcontainer.rs:10-159 (BondContainer); util/vec_help.rs:2-23 (remove_doubles);
 * caller qmc_ising.rs:705-752 (timestep), :322-420 (single_rvb_sweep), EdgeNav :610-636, make_classical_bonds :420-432.
 *
 * The reference walks the per-variable links with binary heaps and "hints" (fast_ops.rs:639-808, 896-1172); those decide
 * HOW ops are found, not WHICH: calculate_flip_prob and mutate_subsection_ops visit every op that touches a sub-variable,
 * in p order, and get_propagated_substate_with_hint returns the state just before p.  This restatement does the same
 * visits on the flat op array; every draw, every f64 operation and the key order inside the BondContainers (push at the
 * end, swap-remove) are the reference's.  Links are rebuilt once at the end (the reference splices them per change).
 * =================================================================================== */
typedef struct { int64_t v, p; double w; } BcKey; /* (T, f64); T = bond (p = NONE) or VarPos{v, p} (rvb.rs:957-965) */
typedef struct {
    int64_t *map; uint64_t map_len; /* Vec<Option<usize>>: index into keys, NONE */
    BcKey *keys; uint64_t len, cap;
    double total;
} BondContainer;
static uint64_t bc_index(int64_t v, int64_t p) { return (uint64_t)(p != NONE ? p : v); } /* From<VarPos> for usize, :961-965 */
static void bc_free(BondContainer *c) { free(c->map), free(c->keys); memset(c, 0, sizeof(*c)); }
static void bc_clear(BondContainer *c) { /* bondcontainer.rs:133-142 */
    for (uint64_t i = 0; i < c->len; i++) c->map[bc_index(c->keys[i].v, c->keys[i].p)] = NONE;
    c->len = 0, c->total = 0.0;
}
static int bc_contains(const BondContainer *c, int64_t v, int64_t p) { /* :90-97 */
    uint64_t t = bc_index(v, p);
    return t < c->map_len && c->map[t] != NONE;
}
static int bc_get_weight(const BondContainer *c, int64_t v, int64_t p, double *w) { /* :100-107 */
    uint64_t t = bc_index(v, p);
    if (t >= c->map_len || c->map[t] == NONE) return 0;
    *w = c->keys[c->map[t]].w;
    return 1;
}
static void bc_correct_total(BondContainer *c) { if (c->total < 0.0) c->total = 0.0; } /* :76-87 */
static void bc_insert(BondContainer *c, int64_t v, int64_t p, double w) { /* :110-130 */
    uint64_t t = bc_index(v, p);
    if (t >= c->map_len) {
        c->map = (int64_t *)realloc(c->map, (t + 1) * sizeof(int64_t));
        for (uint64_t i = c->map_len; i <= t; i++) c->map[i] = NONE;
        c->map_len = t + 1;
    }
    if (c->map[t] != NONE) {
        BcKey *k = &c->keys[c->map[t]];
        double old = k->w;
        k->w = w;
        c->total += w - old;
        bc_correct_total(c);
    } else {
        if (c->len == c->cap) c->cap = c->cap ? 2 * c->cap : 64, c->keys = (BcKey *)realloc(c->keys, c->cap * sizeof(BcKey));
        c->map[t] = (int64_t)c->len;
        c->keys[c->len].v = v, c->keys[c->len].p = p, c->keys[c->len].w = w;
        c->len++;
        c->total += w;
    }
}
static void bc_remove_index(BondContainer *c, uint64_t ki) { /* :58-74: swap with the last key, pop */
    uint64_t last = c->len - 1;
    BcKey tmp = c->keys[ki];
    c->keys[ki] = c->keys[last], c->keys[last] = tmp;
    c->map[bc_index(c->keys[ki].v, c->keys[ki].p)] = (int64_t)ki;
    BcKey out = c->keys[last];
    c->len--;
    c->map[bc_index(out.v, out.p)] = NONE;
    c->total -= out.w;
    bc_correct_total(c);
}
static void bc_remove(BondContainer *c, int64_t v, int64_t p) { /* :47-56 */
    uint64_t t = bc_index(v, p);
    if (t < c->map_len && c->map[t] != NONE) bc_remove_index(c, (uint64_t)c->map[t]);
}
/* get_random, :30-44: index of the chosen key, or NONE where the reference panics */
static int64_t bc_get_random(const BondContainer *c, OrcSse *g) {
    if (c->len == 0) { g->error |= 64; return NONE; }                  /* .unwrap() on None */
    if (!(0.0 < c->total)) { g->error |= 64; return NONE; }            /* gen_range panics on an empty range */
    double p = gen_range_f64(&g->rng, 0.0, c->total);
    uint64_t i = 0;
    while (i < c->len) {
        p -= c->keys[i].w;
        if (p <= 0.0) break;
        i++;
    }
    if (i >= c->len) { g->error |= 64; return NONE; }                  /* index out of bounds */
    return (int64_t)i;
}

typedef struct {
    OrcSse *g;
    /* EdgeNav (qmc_ising.rs:610-636): classical_bonds[v] = bonds of v in bond order (make_classical_bonds :420-432) */
    uint32_t *vb_start, *vb_list;
    /* find_constants, rvb.rs:1162-1188 */
    uint64_t *var_starts, *var_lengths, *constant_ps, ncp;
    uint32_t *zero_vars, nzero;
    /* WeightedBoundaryManager, :967-973 */
    BondContainer b_flips, b_noflips;
    uint8_t *pos_popped, *nopos_popped;
    /* per update */
    int64_t *cl_vars, *cl_flips; uint64_t ncl;
    uint32_t *subvars, nsub; int64_t *v2s;
    uint8_t *cstate, *substate;
    uint64_t *toggles, ntog;
    BondContainer bonds, bonds_before, bonds_after;
} Rvb;

static int64_t rvb_other_var(const OrcSse *g, uint32_t v, uint32_t b) { /* EdgeNavigator::other_var_for_bond, rvb.rs:22-31 */
    if (v == g->ea[b]) return g->eb[b];
    if (v == g->eb[b]) return g->ea[b];
    return NONE;
}
static double rvb_edge_weight(const OrcSse *g, uint32_t b, int sa, int sb) { /* the closure of qmc_ising.rs:718-721 */
    return two_site_hamiltonian(sa, sb, sa, sb, g->J[b]);
}
/* WeightedBoundaryManager::push_adjacent, rvb.rs:1028-1047 */
static void cbm_push_adjacent(Rvb *R, uint32_t var, int64_t pos, int has_w, double weight) {
    double w = has_w ? weight : 1.0;
    BondContainer *bd = pos != NONE ? &R->b_flips : &R->b_noflips;
    uint8_t *popped = pos != NONE ? R->pos_popped : R->nopos_popped;
    uint64_t idx = bc_index(var, pos);
    if (!popped[idx]) {
        double cur = 0.0;
        if (!bc_get_weight(bd, var, pos, &cur)) cur = 0.0;
        bc_insert(bd, var, pos, cur + w);
    }
}
/* pop_index, :1010-1026 */
static int cbm_pop_index(Rvb *R, uint32_t *v_out, int64_t *p_out) {
    OrcSse *g = R->g;
    double total = R->b_flips.total + R->b_noflips.total;
    double f_ratio = R->b_flips.total / total;
    int pick_flips = gen_bool(&g->rng, f_ratio);
    if (g->rng.error) return 0;
    BondContainer *bd = pick_flips ? &R->b_flips : &R->b_noflips;
    uint8_t *popped = pick_flips ? R->pos_popped : R->nopos_popped;
    int64_t i = bc_get_random(bd, g);
    if (i == NONE) return 0;
    BcKey k = bd->keys[i];
    popped[bc_index(k.v, k.p)] = 1;
    bc_remove(bd, k.v, k.p);
    *v_out = (uint32_t)k.v, *p_out = k.p;
    return 1;
}
/* find_overlapping_starts, :1124-1160: calls push_adjacent for every index it yields */
static void rvb_push_overlapping(Rvb *R, uint32_t ov, uint64_t p_start, uint64_t p_end, uint64_t cutoff, double weight) {
    const uint64_t *fp = R->constant_ps + R->var_starts[ov];
    const uint64_t len = R->var_lengths[ov];
    uint64_t bin_found = 0; /* binary_search(&p_start).unwrap_err(): number of entries below p_start */
    while (bin_found < len && fp[bin_found] < p_start) bin_found++;
    if (bin_found < len && fp[bin_found] == p_start) { R->g->error |= 64; return; } /* unwrap_err on Ok */
    const uint64_t prev = (bin_found + len - 1) % len;
    const uint64_t lowest = fp[prev];
    const uint64_t off_start = (p_start + cutoff - lowest) % cutoff, off_end = (p_end + cutoff - lowest) % cutoff;
    for (uint64_t k = 0; k < len; k++) {
        const uint64_t ip = (prev + k) % len; /* [prev..] then [..prev] */
        const uint64_t check_start = (fp[ip] + cutoff - lowest) % cutoff;
        const uint64_t check_end = (fp[(ip + 1) % len] + cutoff - lowest) % cutoff;
        const int has_overlap_start = check_start < off_start && off_start < check_end;
        const int has_start_within = off_start < check_start && check_start < off_end;
        const int eq = (p_start == p_end) || (check_start == check_end);
        if (!(eq || has_overlap_start || has_start_within)) break; /* take_while */
        cbm_push_adjacent(R, ov, (int64_t)(ip + R->var_starts[ov]), 1, weight);
    }
}
/* build_cluster, :1054-1122 */
static void rvb_build_cluster(Rvb *R, uint64_t cluster_size, uint32_t init_var, int64_t init_flip, uint64_t cutoff) {
    OrcSse *g = R->g;
    cbm_push_adjacent(R, init_var, init_flip, 0, 0.0);
    while (cluster_size > 0 && !(R->b_flips.len == 0 && R->b_noflips.len == 0)) {
        uint32_t v;
        int64_t flip;
        if (!cbm_pop_index(R, &v, &flip)) return;
        R->cl_vars[R->ncl] = v, R->cl_flips[R->ncl] = flip, R->ncl++;
        const uint64_t vs = R->var_starts[v], vl = R->var_lengths[v];
        if (flip != NONE) {
            uint64_t rel = (uint64_t)flip - vs;
            cbm_push_adjacent(R, v, (int64_t)((rel + vl - 1) % vl + vs), 0, 0.0);
            cbm_push_adjacent(R, v, (int64_t)((rel + 1) % vl + vs), 0, 0.0);
        }
        for (uint32_t k = R->vb_start[v]; k < R->vb_start[v + 1]; k++) {
            const uint32_t b = R->vb_list[k];
            const double weight = fabs(g->J[b]); /* bond_mag */
            const int64_t ov = rvb_other_var(g, v, b);
            if (R->var_lengths[ov] == 0) {
                cbm_push_adjacent(R, (uint32_t)ov, NONE, 1, weight);
            } else if (flip != NONE) {
                uint64_t rel = (uint64_t)flip - vs, flip_inc = (rel + 1) % vl + vs;
                rvb_push_overlapping(R, (uint32_t)ov, R->constant_ps[flip], R->constant_ps[flip_inc], cutoff, weight);
            } else {
                for (uint64_t pi = R->var_starts[ov]; pi < R->var_starts[ov] + R->var_lengths[ov]; pi++)
                    cbm_push_adjacent(R, (uint32_t)ov, (int64_t)pi, 1, weight);
            }
        }
        cluster_size--;
    }
}
static int rvb_near(const Rvb *R, const Node *nd) { /* does the op touch a sub-variable */
    for (int r = 0; r < nd->nv; r++)
        if (R->v2s[nd->vars[r]] != NONE) return 1;
    return 0;
}
/* ws_for_flip, :665-683 */
static void rvb_ws_for_flip(const Rvb *R, uint32_t b, int64_t sub_flip, double *wbef, double *waft) {
    const OrcSse *g = R->g;
    const int64_t suba = R->v2s[g->ea[b]], subb = R->v2s[g->eb[b]];
    int ba = R->substate[suba], bb = R->substate[subb];
    *wbef = rvb_edge_weight(g, b, ba, bb);
    if (sub_flip == suba) ba = !ba;
    else bb = !bb;
    *waft = rvb_edge_weight(g, b, ba, bb);
}
/* calculate_mult, :1194-1221 */
static double rvb_calculate_mult(const BondContainer *bef, const BondContainer *aft, uint64_t n) {
    const int close = fabs(bef->total - aft->total) < DBL_EPSILON;
    if (n == 0 || close) return 1.0;
    return orc_powi(aft->total / bef->total, (int)n);
}
/* calculate_flip_prob, :649-946 */
static double rvb_calculate_flip_prob(Rvb *R) {
    OrcSse *g = R->g;
    uint64_t cluster_size = 0, next_ci = 0, n_bonds = 0;
    double mult = 1.0;
    for (uint32_t s = 0; s < R->nsub; s++) cluster_size += R->cstate[s];
    bc_clear(&R->bonds_before), bc_clear(&R->bonds_after);
    if (cluster_size != 0) { /* set_initial_bonds, :617-646 */
        for (uint32_t s = 0; s < R->nsub; s++) {
            if (!R->cstate[s]) continue;
            const uint32_t v = R->subvars[s];
            for (uint32_t k = R->vb_start[v]; k < R->vb_start[v + 1]; k++) {
                const uint32_t b = R->vb_list[k];
                const int64_t os = R->v2s[rvb_other_var(g, v, b)];
                if (os == NONE) { g->error |= 64; return 0.0; } /* .unwrap() */
                if (!R->cstate[os]) {
                    double wb, wa;
                    rvb_ws_for_flip(R, b, s, &wb, &wa);
                    bc_insert(&R->bonds_before, b, NONE, wb), bc_insert(&R->bonds_after, b, NONE, wa);
                }
            }
        }
    }
    uint64_t pos = 0; /* every op below pos has been popped from the heap */
    for (;;) {
        uint64_t q = pos; /* heap top: the next op on a sub-variable's world line */
        while (q < g->ops_len && !(g->ops[q].present && rvb_near(R, &g->ops[q]))) q++;
        if (q >= g->ops_len) break;
        uint64_t p = q;
        if (cluster_size == 0) { /* :722-733: jump to the next cluster flip */
            if (next_ci < R->ntog) p = R->toggles[next_ci];
            else break;
        }
        for (uint64_t x = q; x < p; x++) { /* popped < p: their outputs propagate the substate, :742-768 */
            const Node *nd = &g->ops[x];
            if (!nd->present) continue;
            for (int r = 0; r < nd->nv; r++)
                if (R->v2s[nd->vars[r]] != NONE) R->substate[R->v2s[nd->vars[r]]] = nd->out[r];
        }
        pos = p + 1;
        const Node *op = &g->ops[p];
        if (p >= g->ops_len || !op->present) { g->error |= 64; return 0.0; }
        const int is_cluster_bound = next_ci < R->ntog && p == R->toggles[next_ci];
        const int will_flip_spins = !node_is_diagonal(op);
        const int will_change_bonds = will_flip_spins || is_cluster_bound;
        int completely_in_cluster = 1;
        for (int r = 0; r < op->nv; r++) {
            const int64_t s = R->v2s[op->vars[r]];
            if (s == NONE || !R->cstate[s]) completely_in_cluster = 0;
        }
        if (bc_contains(&R->bonds_before, op->bond, NONE)) {
            n_bonds++;
            continue;
        }
        if (is_cluster_bound) { /* :852-867 */
            const int64_t s = R->v2s[op->vars[0]];
            if (s == NONE) { g->error |= 64; return 0.0; }
            R->cstate[s] = !R->cstate[s];
            if (R->cstate[s]) cluster_size++;
            else cluster_size--;
            next_ci++;
        }
        if (will_flip_spins)
            for (int r = 0; r < op->nv; r++)
                if (R->v2s[op->vars[r]] != NONE) R->substate[R->v2s[op->vars[r]]] = op->out[r];
        if (completely_in_cluster) { /* ising_ratio, qmc_ising.rs:722-735; rvb_update (:60-79) uses 1.0 */
            const int long_bond = fabs(g->longitudinal) > DBL_EPSILON && op->bond >= g->nedges + g->nvars;
            mult *= long_bond ? 0.0 : 1.0;
            if (mult < DBL_EPSILON) break;
        }
        if (will_change_bonds) {
            mult *= rvb_calculate_mult(&R->bonds_before, &R->bonds_after, n_bonds);
            n_bonds = 0;
            if (mult < DBL_EPSILON) break;
            for (int r = 0; r < op->nv; r++) { /* :901-934 */
                const uint32_t v = op->vars[r];
                const int64_t s = R->v2s[v];
                if (s == NONE) continue;
                for (uint32_t k = R->vb_start[v]; k < R->vb_start[v + 1]; k++) {
                    const uint32_t b = R->vb_list[k];
                    const int64_t os = R->v2s[rvb_other_var(g, v, b)];
                    if (os == NONE) continue;
                    if (R->cstate[s] == R->cstate[os]) {
                        if (bc_contains(&R->bonds_before, b, NONE)) bc_remove(&R->bonds_before, b, NONE), bc_remove(&R->bonds_after, b, NONE);
                    } else {
                        double wb, wa;
                        rvb_ws_for_flip(R, b, R->cstate[s] ? s : os, &wb, &wa);
                        bc_insert(&R->bonds_before, b, NONE, wb), bc_insert(&R->bonds_after, b, NONE, wa);
                    }
                }
            }
        }
    }
    mult *= rvb_calculate_mult(&R->bonds_before, &R->bonds_after, n_bonds);
    return mult;
}
/* the "Now update bonds" block of mutate_graph, :561-593 */
static void rvb_mutate_update_bonds(Rvb *R, const Node *op) {
    const OrcSse *g = R->g;
    for (int r = 0; r < op->nv; r++) {
        const uint32_t v = op->vars[r];
        const int64_t s = R->v2s[v];
        if (s == NONE) continue;
        for (uint32_t k = R->vb_start[v]; k < R->vb_start[v + 1]; k++) {
            const uint32_t b = R->vb_list[k];
            const int64_t os = R->v2s[rvb_other_var(g, v, b)];
            if (os == NONE) continue;
            if (R->cstate[s] == R->cstate[os]) {
                if (bc_contains(&R->bonds, b, NONE)) bc_remove(&R->bonds, b, NONE);
            } else {
                bc_insert(&R->bonds, b, NONE, rvb_edge_weight(g, b, R->substate[R->v2s[g->ea[b]]], R->substate[R->v2s[g->eb[b]]]));
            }
        }
    }
}
/* mutate_graph, :294-616 */
static void rvb_mutate_graph(Rvb *R) {
    OrcSse *g = R->g;
    uint64_t *jump_to = (uint64_t *)malloc((R->ntog + 2) * sizeof(uint64_t));
    uint64_t *cont_until = (uint64_t *)malloc((R->ntog + 2) * sizeof(uint64_t));
    uint64_t nj = 0, nc = 0, count = 0;
    for (uint32_t s = 0; s < R->nsub; s++) count += R->cstate[s];
    if (count != 0) {
        jump_to[nj++] = 0;
        for (uint32_t s = 0; s < R->nsub; s++) R->substate[s] ^= R->cstate[s];
    }
    for (uint64_t t = 0; t < R->ntog; t++) {
        const uint64_t p = R->toggles[t];
        if (count == 0) jump_to[nj++] = p;
        const Node *op = &g->ops[p];
        for (int r = 0; r < op->nv; r++) {
            const int64_t s = R->v2s[op->vars[r]];
            if (s == NONE) continue;
            R->cstate[s] = !R->cstate[s];
            if (R->cstate[s]) count++;
            else count--;
        }
        if (count == 0) cont_until[nc++] = p;
    }
    if (count != 0) cont_until[nc++] = g->ops_len; /* rvb.get_cutoff() = ops.len(), fast_ops.rs:1255-1257 */
    if (nj != nc) g->error |= 64;
    bc_clear(&R->bonds);
    uint64_t next_ci = 0;
    for (uint32_t s = 0; s < R->nsub && !(g->error & 64); s++) { /* :366-381 */
        if (!R->cstate[s]) continue;
        const uint32_t v = R->subvars[s];
        for (uint32_t k = R->vb_start[v]; k < R->vb_start[v + 1]; k++) {
            const uint32_t b = R->vb_list[k];
            const int64_t os = R->v2s[rvb_other_var(g, v, b)];
            if (os == NONE) { g->error |= 64; break; }
            if (!R->cstate[os])
                bc_insert(&R->bonds, b, NONE, rvb_edge_weight(g, b, R->substate[R->v2s[g->ea[b]]], R->substate[R->v2s[g->eb[b]]]));
        }
    }
    for (uint64_t k = 0; k < nj && k < nc && !(g->error & 64); k++) {
        const uint64_t from = jump_to[k], until = cont_until[k];
        /* get_propagated_substate_with_hint (fast_ops.rs:1027-1172): the state just before `from` */
        for (uint32_t s = 0; s < R->nsub; s++) {
            const uint32_t v = R->subvars[s];
            int val = g->state[v];
            for (uint64_t x = from; x-- > 0;) {
                const Node *nd = &g->ops[x];
                if (!nd->present) continue;
                int hit = 0;
                for (int r = 0; r < nd->nv; r++)
                    if (nd->vars[r] == v) val = nd->out[r], hit = 1;
                if (hit) break;
            }
            R->substate[s] = (uint8_t)(val ^ R->cstate[s]); /* :396-399 */
        }
        /* mutate_subsection_ops (fast_ops.rs:639-775): ops with p in [from, until] on the sub-variables' world lines */
        for (uint64_t p = from; p <= until && p < g->ops_len && !(g->error & 64); p++) {
            Node *op = &g->ops[p];
            if (!op->present || !rvb_near(R, op)) continue;
            const int in_bonds = bc_contains(&R->bonds, op->bond, NONE);
            const int at_next = next_ci < R->ntog && p == R->toggles[next_ci];
            if (in_bonds) { /* :411-432: rotate the diagonal op onto a bond drawn from the border */
                const int64_t i = bc_get_random(&R->bonds, g);
                if (i == NONE) break;
                const uint32_t nb = (uint32_t)R->bonds.keys[i].v;
                const int64_t sa = R->v2s[g->ea[nb]], sb = R->v2s[g->eb[nb]];
                if (sa == NONE || sb == NONE) { g->error |= 64; break; }
                op->nv = 2, op->vars[0] = g->ea[nb], op->vars[1] = g->eb[nb], op->bond = nb;
                op->in[0] = op->out[0] = R->substate[sa], op->in[1] = op->out[1] = R->substate[sb];
                continue; /* constant stays op.is_constant() */
            }
            const Node old = *op; /* the bonds are updated with the vars of the op as it was (same vars either way) */
            if (at_next) { /* :434-469 */
                for (int r = 0; r < op->nv; r++) {
                    const int64_t s = R->v2s[op->vars[r]];
                    if (s == NONE) { g->error |= 64; break; }
                    op->in[r] = op->in[r] != R->cstate[s];
                    op->out[r] = op->out[r] != !R->cstate[s];
                }
                for (int r = 0; r < op->nv && !(g->error & 64); r++) {
                    const int64_t s = R->v2s[op->vars[r]];
                    R->cstate[s] = !R->cstate[s];
                    R->substate[s] = op->out[r];
                }
                next_ci++;
            } else { /* :470-557 */
                int any_in_cluster = 0;
                for (int r = 0; r < op->nv; r++) {
                    const int64_t s = R->v2s[op->vars[r]];
                    if (s != NONE && R->cstate[s]) any_in_cluster = 1;
                }
                if (!any_in_cluster && node_is_diagonal(op)) {
                } else if (any_in_cluster) {
                    for (int r = 0; r < op->nv; r++) {
                        if (R->v2s[op->vars[r]] == NONE) { g->error |= 64; break; }
                        op->in[r] = !op->in[r], op->out[r] = !op->out[r];
                    }
                    if (!node_is_diagonal(op))
                        for (int r = 0; r < op->nv && !(g->error & 64); r++) R->substate[R->v2s[op->vars[r]]] = op->out[r];
                } else {
                    int k2 = 0; /* filter_map(var_to_subvar).zip(outputs): the k-th sub-variable takes output k */
                    for (int r = 0; r < op->nv; r++)
                        if (R->v2s[op->vars[r]] != NONE) R->substate[R->v2s[op->vars[r]]] = op->out[k2++];
                }
            }
            rvb_mutate_update_bonds(R, &old);
        }
    }
    free(jump_to), free(cont_until);
}
/* contiguous_bits, rvb.rs:1190-1192: next_u64().trailing_ones() */
static uint64_t rvb_contiguous_bits(Stream *s) {
    uint64_t x = next_u64(s);
    return ~x == 0 ? 64u : (uint64_t)__builtin_ctzll(~x);
}
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}
/* RvbUpdater::rvb_update_with_ising_weight as QmcIsingGraph calls it (qmc_ising.rs:705-752): returns the successes */
uint64_t orc_sse_rvb_update(OrcSse *g, uint64_t updates) {
    if (g->is_qmc) { g->error |= 64; return 0; }
    Rvb R;
    memset(&R, 0, sizeof(R));
    R.g = g;
    const uint32_t N = g->nvars, E = g->nedges;
    /* make_classical_bonds */
    R.vb_start = (uint32_t *)calloc(N + 2, sizeof(uint32_t));
    R.vb_list = (uint32_t *)malloc((2 * (size_t)E + 1) * sizeof(uint32_t));
    for (uint32_t b = 0; b < E; b++) R.vb_start[g->ea[b] + 1]++, R.vb_start[g->eb[b] + 1]++;
    for (uint32_t v = 0; v < N; v++) R.vb_start[v + 1] += R.vb_start[v];
    {
        uint32_t *fill = (uint32_t *)calloc(N + 1, sizeof(uint32_t));
        for (uint32_t b = 0; b < E; b++) {
            R.vb_list[R.vb_start[g->ea[b]] + fill[g->ea[b]]++] = b;
            R.vb_list[R.vb_start[g->eb[b]] + fill[g->eb[b]]++] = b;
        }
        free(fill);
    }
    /* find_constants, rvb.rs:1162-1188 (constant_ops_on_var, fast_ops.rs:1466-1480: p order on each variable) */
    R.var_starts = (uint64_t *)calloc(N + 1, sizeof(uint64_t));
    R.var_lengths = (uint64_t *)calloc(N + 1, sizeof(uint64_t));
    R.zero_vars = (uint32_t *)malloc((N + 1) * sizeof(uint32_t));
    for (uint64_t p = 0; p < g->ops_len; p++) {
        const Node *nd = &g->ops[p];
        if (nd->present && nd->constant) {
            if (nd->nv != 1) g->error |= 64;
            R.var_lengths[nd->vars[0]]++, R.ncp++;
        }
    }
    R.constant_ps = (uint64_t *)malloc((R.ncp + 1) * sizeof(uint64_t));
    {
        uint64_t acc = 0;
        for (uint32_t v = 0; v < N; v++) {
            R.var_starts[v] = acc, acc += R.var_lengths[v];
            if (R.var_lengths[v] == 0) R.zero_vars[R.nzero++] = v;
        }
        uint64_t *fill = (uint64_t *)calloc(N + 1, sizeof(uint64_t));
        for (uint64_t p = 0; p < g->ops_len; p++) {
            const Node *nd = &g->ops[p];
            if (nd->present && nd->constant && nd->nv == 1) R.constant_ps[R.var_starts[nd->vars[0]] + fill[nd->vars[0]]++] = p;
        }
        free(fill);
    }
    R.pos_popped = (uint8_t *)calloc(R.ncp + N + 1, 1);
    R.nopos_popped = (uint8_t *)calloc(N + 1, 1);
    R.cl_vars = (int64_t *)malloc(80 * sizeof(int64_t)), R.cl_flips = (int64_t *)malloc(80 * sizeof(int64_t));
    R.subvars = (uint32_t *)malloc((N + 1) * sizeof(uint32_t));
    R.v2s = (int64_t *)malloc((N + 1) * sizeof(int64_t));
    R.cstate = (uint8_t *)malloc(N + 1), R.substate = (uint8_t *)malloc(N + 1);
    R.toggles = (uint64_t *)malloc(170 * sizeof(uint64_t));
    uint8_t *mark = (uint8_t *)calloc(N + 1, 1);
    uint64_t num_succ = 0;
    for (uint64_t u = 0; u < updates && !g->error && !g->rng.error; u++) {
        const uint64_t choice = gen_range_usize(&g->rng, R.ncp + R.nzero); /* rvb.rs:125 */
        uint32_t v;
        int64_t flip;
        if (choice < R.ncp) { /* :126-139: the last variable whose start is <= choice */
            uint32_t lo = 0;
            for (uint32_t i = 0; i < N; i++)
                if (R.var_starts[i] <= choice) lo = i;
            v = lo, flip = (int64_t)choice;
        } else {
            v = R.zero_vars[choice - R.ncp], flip = NONE;
        }
        const uint64_t cluster_size = rvb_contiguous_bits(&g->rng) + 1;
        /* fresh boundary manager (:152) */
        bc_clear(&R.b_flips), bc_clear(&R.b_noflips);
        memset(R.pos_popped, 0, R.ncp + N + 1), memset(R.nopos_popped, 0, N + 1);
        R.ncl = 0;
        rvb_build_cluster(&R, cluster_size, v, flip, g->ops_len);
        if (g->error || g->rng.error) break;
        /* dissolve_into (:987-1007) + sub-variables (:168-180) */
        R.nsub = 0;
        for (uint64_t i = 0; i < R.ncl; i++) mark[R.cl_vars[i]] = 1;
        for (uint64_t i = 0; i < R.b_flips.len; i++) mark[R.b_flips.keys[i].v] = 1;
        for (uint64_t i = 0; i < R.b_noflips.len; i++) mark[R.b_noflips.keys[i].v] = 1;
        for (uint32_t w = 0; w < N; w++) {
            R.v2s[w] = NONE;
            if (mark[w]) R.v2s[w] = R.nsub, R.subvars[R.nsub++] = w, mark[w] = 0;
        }
        memset(R.cstate, 0, R.nsub);
        R.ntog = 0;
        for (uint64_t i = 0; i < R.ncl; i++) { /* :182-203 */
            const int64_t s = R.v2s[R.cl_vars[i]], fi = R.cl_flips[i];
            if (fi != NONE) {
                const uint64_t vstart = R.var_starts[R.cl_vars[i]], fi_rel = (uint64_t)fi - vstart;
                if (fi_rel + 1 >= R.var_lengths[R.cl_vars[i]]) {
                    R.cstate[s] = 1;
                    R.toggles[R.ntog++] = R.constant_ps[fi], R.toggles[R.ntog++] = R.constant_ps[vstart];
                } else {
                    R.toggles[R.ntog++] = R.constant_ps[fi], R.toggles[R.ntog++] = R.constant_ps[fi + 1];
                }
            } else {
                R.cstate[s] = 1;
            }
        }
        for (uint32_t s = 0; s < R.nsub; s++) R.substate[s] = g->state[R.subvars[s]];
        qsort(R.toggles, R.ntog, sizeof(uint64_t), cmp_u64);
        { /* remove_doubles, vec_help.rs:2-23 */
            uint64_t ii = 0, jj = 0;
            while (jj + 1 < R.ntog) {
                if (R.toggles[jj] == R.toggles[jj + 1]) jj += 2;
                else R.toggles[ii++] = R.toggles[jj++];
            }
            if (jj < R.ntog) R.toggles[ii++] = R.toggles[jj++];
            R.ntog = ii;
        }
        const double p_to_flip = rvb_calculate_flip_prob(&R);
        if (g->error) break;
        const int should_mutate = p_to_flip >= 1.0 ? 1 : gen_bool(&g->rng, p_to_flip);
        if (should_mutate) {
            rvb_mutate_graph(&R);
            int starting = 0;
            for (uint32_t s = 0; s < R.nsub; s++) starting |= R.cstate[s];
            if (starting)
                for (uint32_t s = 0; s < R.nsub; s++) g->state[R.subvars[s]] ^= R.cstate[s];
            num_succ++;
        }
    }
    rebuild_links(g);
    free(R.vb_start), free(R.vb_list), free(R.var_starts), free(R.var_lengths), free(R.constant_ps), free(R.zero_vars);
    free(R.pos_popped), free(R.nopos_popped), free(R.cl_vars), free(R.cl_flips), free(R.subvars), free(R.v2s);
    free(R.cstate), free(R.substate), free(R.toggles), free(mark);
    bc_free(&R.b_flips), bc_free(&R.b_noflips), bc_free(&R.bonds), bc_free(&R.bonds_before), bc_free(&R.bonds_after);
    return num_succ;
}

/* set_run_rvb (qmc_ising.rs:434-441), rvb_success_rate (:604-607), single_rvb_sweep (:322-420) */
void orc_sse_set_run_rvb(OrcSse *g, int run_rvb) { g->run_rvb = run_rvb != 0; }
double orc_sse_rvb_success_rate(const OrcSse *g) { return (double)g->total_rvb_successes / (double)g->rvb_clusters_counted; }
uint64_t orc_sse_single_rvb_sweep(OrcSse *g, int64_t updates_in_sweep, uint64_t *steps_out) {
    uint64_t steps = updates_in_sweep >= 0 ? (uint64_t)updates_in_sweep : ((uint64_t)g->nvars + 1) / 2;
    if (steps_out) *steps_out = steps;
    return orc_sse_rvb_update(g, steps);
}

/* QmcIsingGraph::timestep, qmc_ising.rs:644-795 (rvb off by default, :122; when on it runs between the diagonal and the
 * cluster update, :705-752) */
void orc_sse_timestep(OrcSse *g, double beta, int mode) {
    if (g->is_qmc) {
        orc_qmc_timestep(g, beta, mode);
        return;
    }
    diagonal_step(g, beta, mode);
    if (g->run_rvb) {
        const uint64_t steps = ((uint64_t)g->nvars + 1) / 2; /* :711 */
        g->total_rvb_successes += orc_sse_rvb_update(g, steps);
        g->rvb_clusters_counted += steps;
    }
    cluster_and_free_spins(g, mode);
    uint64_t grown = g->n + g->n / 2; /* :786 */
    if (grown > g->cutoff) g->cutoff = grown;
}

/* qmc_ising.rs:208-270 */
void orc_sse_single_diagonal_step(OrcSse *g, double beta) { orc_sse_single_diagonal_step_mode(g, beta, ORC_MODE_STRICT); }
void orc_sse_single_diagonal_step_mode(OrcSse *g, double beta, int mode) {
    diagonal_step(g, beta, mode);
    uint64_t grown = g->n + g->n / 2;
    if (grown > g->cutoff) g->cutoff = grown;
}
uint64_t orc_sse_single_cluster_step(OrcSse *g, int mode) { return cluster_and_free_spins(g, mode); }

/* QmcStepper::timesteps_measure_with_self, qmc_stepper.rs:133-162; energy qmc_ising.rs:805-809 */
double orc_sse_timesteps(OrcSse *g, uint64_t t, double beta, uint64_t sampling_freq, int mode,
                         uint8_t *samples_or_null) {
    uint64_t steps_measured = 0, total_n = 0;
    if (sampling_freq == 0) sampling_freq = 1;
    for (uint64_t i = 0; i < t; i++) {
        orc_sse_timestep(g, beta, mode);
        if ((i + 1) % sampling_freq == 0) {
            if (samples_or_null) memcpy(samples_or_null + steps_measured * g->nvars, g->state, g->nvars);
            steps_measured++;
            total_n += g->n;
        }
    }
    double average_n = (double)total_n / (double)steps_measured;
    return -(average_n / beta) + g->offset;
}

uint32_t orc_sse_nvars(const OrcSse *g) { return g->nvars; }
uint64_t orc_sse_get_n(const OrcSse *g) { return g->n; }
uint64_t orc_sse_get_cutoff(const OrcSse *g) { return g->cutoff; }
void orc_sse_set_cutoff(OrcSse *g, uint64_t cutoff) { /* qmc_ising.rs:537-540 */
    g->cutoff = cutoff;
    ops_resize(g, cutoff);
}
void orc_sse_use_small_rng(OrcSse *g) { small_rng_seed(&g->rng, g->rng.key); } /* timing only, see Stream */
uint64_t orc_sse_get_cursor(const OrcSse *g) { return g->rng.cursor; }
void orc_sse_set_cursor(OrcSse *g, uint64_t cursor) { g->rng.cursor = cursor; }
void orc_sse_set_key(OrcSse *g, uint64_t key) { g->rng.key = key; }
uint64_t orc_sse_get_key(const OrcSse *g) { return g->rng.key; }
double orc_sse_get_offset(const OrcSse *g) { return g->offset; }
void orc_sse_get_state(const OrcSse *g, uint8_t *out) { memcpy(out, g->state, g->nvars); }
void orc_sse_set_state(OrcSse *g, const uint8_t *in) { memcpy(g->state, in, g->nvars); }

uint64_t orc_sse_get_bond_count(const OrcSse *g, uint32_t bond) { /* fast_ops.rs:1281-1294 */
    uint64_t c = 0;
    for (uint64_t p = 0; p < g->ops_len; p++) c += g->ops[p].present && g->ops[p].bond == bond;
    return c;
}

/* OpContainer::itime_fold (fast_ops.rs:1296-1315) through QmcIsingGraph::imaginary_time_fold
 * (qmc_ising.rs:815-821) with the fold "accumulate the magnetisation of the propagated state":
 * for p in 0..cutoff the fold sees the state BEFORE the op at p.  sums[0] = sum_p m_p, sums[1] = sum_p m_p^2,
 * sums[2] = sum_p |m_p| with the integer m_p = sum_v (2 s_v - 1); returns the number of slots folded. */
uint64_t orc_sse_itime_magnetization(const OrcSse *g, int64_t sums[3]) {
    uint8_t *state = (uint8_t *)malloc(g->nvars);
    memcpy(state, g->state, g->nvars);
    sums[0] = sums[1] = sums[2] = 0;
    for (uint64_t p = 0; p < g->cutoff; p++) {
        int64_t m = 0;
        for (uint32_t v = 0; v < g->nvars; v++) m += state[v] ? 1 : -1;
        sums[0] += m, sums[1] += m * m, sums[2] += m < 0 ? -m : m;
        if (p < g->ops_len && g->ops[p].present) {
            const Node *nd = &g->ops[p];
            for (int r = 0; r < nd->nv; r++) state[nd->vars[r]] = nd->out[r];
        }
    }
    free(state);
    return g->cutoff;
}
/* the propagated state before slot p (what the fold closure is handed at step p) */
void orc_sse_itime_state(const OrcSse *g, uint64_t p_at, uint8_t *out) {
    memcpy(out, g->state, g->nvars);
    for (uint64_t p = 0; p < p_at && p < g->ops_len; p++)
        if (g->ops[p].present)
            for (int r = 0; r < g->ops[p].nv; r++) out[g->ops[p].vars[r]] = g->ops[p].out[r];
}

void orc_sse_dump_ops(const OrcSse *g, uint32_t *words) {
    for (uint64_t p = 0; p < g->cutoff; p++) {
        const Node *nd = p < g->ops_len ? &g->ops[p] : NULL;
        if (!nd || !nd->present) {
            words[p] = ORC_OP_EMPTY;
            continue;
        }
        uint32_t w = nd->bond;
        for (int r = 0; r < nd->nv; r++) w |= (uint32_t)nd->in[r] << (24 + r), w |= (uint32_t)nd->out[r] << (26 + r);
        words[p] = w;
    }
}

int orc_sse_load_ops(OrcSse *g, const uint32_t *words, uint64_t nwords, const uint8_t *state) {
    /* FastOps::new_from_ops, fast_ops.rs:80-87 */
    if (nwords > g->cutoff) orc_sse_set_cutoff(g, nwords);
    memset(g->ops, 0, g->ops_len * sizeof(Node));
    for (uint64_t p = 0; p < nwords; p++) {
        uint32_t w = words[p];
        if (w == ORC_OP_EMPTY) continue;
        uint32_t b = w & 0xFFFFFFu;
        if (b >= (g->is_qmc ? g->ninter : g->nedges + 2 * g->nvars)) return -1;
        Node *nd = &g->ops[p];
        int nv, constant;
        edge_fn(g, b, nd->vars, &nv, &constant);
        nd->present = 1, nd->nv = (uint8_t)nv, nd->constant = (uint8_t)constant, nd->bond = b;
        for (int r = 0; r < nv; r++) nd->in[r] = (w >> (24 + r)) & 1u, nd->out[r] = (w >> (26 + r)) & 1u;
    }
    if (state) memcpy(g->state, state, g->nvars);
    rebuild_links(g);
    return 0;
}

/* OpContainer::verify op_container.rs:137-159 + QmcIsingGraph::verify qmc_ising.rs:829-860 */
int orc_sse_verify(const OrcSse *g) {
    uint8_t *rolling = (uint8_t *)malloc(g->nvars);
    memcpy(rolling, g->state, g->nvars);
    int ok = 1;
    uint64_t n = 0;
    for (uint64_t p = 0; p < g->ops_len && ok; p++) {
        const Node *nd = &g->ops[p];
        if (!nd->present) continue;
        n++;
        if (p >= g->cutoff) ok = 0;
        if (!(fabs(hamiltonian(g, nd->bond, nd->in, nd->out)) > DBL_EPSILON)) ok = 0;
        for (int r = 0; r < nd->nv; r++)
            if (rolling[nd->vars[r]] != nd->in[r]) ok = 0;
        for (int r = 0; r < nd->nv; r++) rolling[nd->vars[r]] = nd->out[r];
    }
    if (memcmp(rolling, g->state, g->nvars) != 0) ok = 0;
    if (n != g->n) ok = 0;
    free(rolling);
    return ok;
}

void orc_sse_get_boundaries(const OrcSse *g, int64_t *b_in, int64_t *b_out, uint64_t nslots) {
    for (uint64_t p = 0; p < nslots; p++) {
        b_in[p] = p < g->b_len ? g->b_in[p] : NONE;
        b_out[p] = p < g->b_len ? g->b_out[p] : NONE;
    }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* graphs.par_iter_mut().for_each(|(g, beta)| g.timesteps(t, beta)), tempering_container.rs:367-371 */
uint64_t orc_sse_batch_timesteps(OrcSse **reps, uint32_t nreps, uint64_t t, const double *betas,
                                 int mode, double *energies_or_null, int nthreads) {
    uint64_t total = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : total)
#endif
    for (uint32_t r = 0; r < nreps; r++) {
        uint64_t tn = 0;
        for (uint64_t i = 0; i < t; i++) {
            orc_sse_timestep(reps[r], betas[r], mode);
            tn += reps[r]->n;
        }
        if (energies_or_null) energies_or_null[r] = -(((double)tn / (double)t) / betas[r]) + reps[r]->offset;
        total += tn;
    }
    (void)nthreads;
    return total;
}

/* ===================================================================================
 * Parallel tempering: tempering_container.rs:83-149, :241-302; swap qmc_ising.rs:593-602
 * =================================================================================== */
static void swap_manager_and_state(OrcSse *a, OrcSse *b) {
#define SWAPF(T, f) do { T t_ = a->f; a->f = b->f; b->f = t_; } while (0)
    SWAPF(Node *, ops);
    SWAPF(uint64_t, ops_len);
    SWAPF(uint64_t, n);
    SWAPF(int64_t, first_p);
    SWAPF(int64_t, last_p);
    SWAPF(int64_t *, vfirst_p);
    SWAPF(int64_t *, vlast_p);
    SWAPF(int8_t *, vfirst_r);
    SWAPF(int8_t *, vlast_r);
    SWAPF(uint8_t *, state);
#undef SWAPF
}

/* GraphWeights::ham_eq (tempering_traits.rs:122-124) -> HamInfo::eq (qmc_ising.rs:899-903): edges and
 * transverse field only; the longitudinal field is NOT compared */
static int ham_eq(const OrcSse *a, const OrcSse *b) {
    if (a->nedges != b->nedges || a->transverse != b->transverse) return 0;
    for (uint32_t e = 0; e < a->nedges; e++)
        if (a->ea[e] != b->ea[e] || a->eb[e] != b->eb[e] || a->J[e] != b->J[e]) return 0;
    return 1;
}

/* GraphWeights::relative_weight (tempering_traits.rs:126-154): weight of self's operators under h's
 * couplings relative to self's own */
static double relative_weight(const OrcSse *self, const OrcSse *h) {
    const uint32_t nb = self->nedges + 2 * self->nvars;
    uint64_t *count = (uint64_t *)calloc(nb, sizeof(uint64_t)); /* get_count, fast_ops.rs:1281-1294 */
    for (uint64_t p = 0; p < self->ops_len; p++)
        if (self->ops[p].present) count[self->ops[p].bond]++;
    double bond_ratio = 1.0; /* Iterator::product */
    for (uint32_t b = 0; b < self->nedges; b++) {
        double ja = h->J[b], jb = self->J[b]; /* h.get_edges().zip(self.get_edges()) */
        bond_ratio = bond_ratio * orc_powi(ja / jb, (int32_t)count[b]);
    }
    uint64_t t_count = 0;
    for (uint32_t v = 0; v < self->nvars; v++) t_count += count[v + self->nedges];
    double transverse_ratio = orc_powi(h->transverse / self->transverse, (int32_t)t_count);
    double res;
    if (fabs(self->longitudinal) > DBL_EPSILON) {
        uint64_t l_count = 0;
        for (uint32_t v = 0; v < self->nvars; v++) l_count += count[v + self->nvars + self->nedges];
        double longitudinal_ratio = orc_powi(h->longitudinal / self->longitudinal, (int32_t)l_count);
        res = bond_ratio * transverse_ratio * longitudinal_ratio;
    } else {
        res = bond_ratio * transverse_ratio;
    }
    free(count);
    return res;
}

/* SwapManagers::can_swap_graphs -> can_swap_managers (qmc_ising.rs:563-590); 0 = ok */
int orc_sse_can_swap(const OrcSse *a, const OrcSse *b) {
    if (a->nedges != b->nedges) return 1;
    for (uint32_t e = 0; e < a->nedges; e++) {
        if (a->ea[e] != b->ea[e] || a->eb[e] != b->eb[e]) return 1;
        if (signbit(a->J[e]) != signbit(b->J[e])) return 2;
    }
    if (signbit(a->longitudinal) != signbit(b->longitudinal)) return 3;
    return 0;
}

static uint64_t perform_swaps(Stream *rng, OrcSse **graphs, const double *betas, uint32_t len) {
    /* tempering_container.rs:241-260, swap_on_chunks :274-302 */
    uint64_t swaps = 0;
    for (uint32_t i = 0; i + 1 < len; i += 2) {
        double p = gen_range_f64_01(rng); /* :255 */
        OrcSse *ga = graphs[i], *gb = graphs[i + 1];
        double rel_h_weight = 1.0; /* :286-292; hameqs :109-119 */
        if (!ham_eq(ga, gb)) {
            double rel_bstate = relative_weight(ga, gb);
            double rel_astate = relative_weight(gb, ga);
            rel_h_weight = rel_bstate * rel_astate;
        }
        double temp_swap = orc_powi(betas[i] / betas[i + 1], (int32_t)gb->n - (int32_t)ga->n); /* :294 */
        double p_swap = temp_swap * rel_h_weight;
        if (p_swap > p) { /* :296-301 */
            swap_manager_and_state(ga, gb);
            swaps++;
        }
    }
    return swaps;
}

uint64_t orc_pt_step(OrcSse **slots, uint32_t nslots, const double *betas, uint64_t pt_key,
                     uint64_t *pt_cursor) {
    if (nslots <= 1) return 0; /* :122-124 */
    uint64_t max_cutoff = 0; /* :129-137 */
    for (uint32_t i = 0; i < nslots; i++)
        if (slots[i]->cutoff > max_cutoff) max_cutoff = slots[i]->cutoff;
    for (uint32_t i = 0; i < nslots; i++) orc_sse_set_cutoff(slots[i], max_cutoff);
    Stream rng = {pt_key, *pt_cursor, NULL, 0, 0, 0, 0, {0, 0}, 0, 0, {0, 0, 0, 0}};
    /* make_first_subgraphs / make_second_subgraphs :83-99 */
    uint32_t a_len = (nslots % 2 == 0) ? nslots : nslots - 1;
    uint32_t b_len = (nslots % 2 == 1) ? nslots - 1 : nslots - 2;
    uint64_t swaps = 0;
    if (gen_bool(&rng, 0.5)) { /* :140-146 */
        swaps += perform_swaps(&rng, slots, betas, a_len);
        swaps += perform_swaps(&rng, slots + 1, betas + 1, b_len);
    } else {
        swaps += perform_swaps(&rng, slots + 1, betas + 1, b_len);
        swaps += perform_swaps(&rng, slots, betas, a_len);
    }
    *pt_cursor = rng.cursor;
    return swaps;
}

/* ===================================================================================
 * Classical graph: classical/graph.rs:56-119, :339-347, :430-453
 * =================================================================================== */
#define TAG_CB 0x43420000u
#define TAG_CB2 0x43430000u

struct OrcCls {
    uint32_t nvars;
    uint32_t *adj_start; /* binding_mat as CSR, neighbours sorted by index (graph.rs:69-78) */
    uint32_t *adj_idx;
    double *adj_j;
    double *biases;
    uint8_t *state;
    Stream rng;
    uint64_t sweep; /* checkerboard sweep counter */
    uint32_t nedges; /* edges in construction order (graph.rs:9, :81) */
    uint32_t *ea, *eb;
    double *ej;
    double *cum_w; /* enable_edge_importance_sampling :321-336, NULL when off */
    double total_w;
    int error; /* the reference returns Err / panics: 1 empty range */
};

OrcCls *orc_cls_create(uint32_t nvars, uint32_t nedges, const uint32_t *ea, const uint32_t *eb,
                       const double *J, const double *biases, uint64_t rng_key,
                       const uint8_t *state_or_null) {
    OrcCls *g = (OrcCls *)calloc(1, sizeof(OrcCls));
    g->nvars = nvars;
    g->rng.key = rng_key;
    g->state = (uint8_t *)malloc(nvars);
    if (state_or_null) memcpy(g->state, state_or_null, nvars);
    else
        for (uint32_t v = 0; v < nvars; v++) g->state[v] = (uint8_t)gen_std_bool(&g->rng); /* :57, :451-453 */
    g->biases = (double *)malloc(sizeof(double) * nvars);
    memcpy(g->biases, biases, sizeof(double) * nvars);
    g->nedges = nedges;
    g->ea = (uint32_t *)malloc(sizeof(uint32_t) * (nedges ? nedges : 1));
    g->eb = (uint32_t *)malloc(sizeof(uint32_t) * (nedges ? nedges : 1));
    g->ej = (double *)malloc(sizeof(double) * (nedges ? nedges : 1));
    memcpy(g->ea, ea, sizeof(uint32_t) * nedges), memcpy(g->eb, eb, sizeof(uint32_t) * nedges);
    memcpy(g->ej, J, sizeof(double) * nedges);
    g->adj_start = (uint32_t *)calloc((size_t)nvars + 1, sizeof(uint32_t));
    for (uint32_t e = 0; e < nedges; e++) g->adj_start[ea[e] + 1]++, g->adj_start[eb[e] + 1]++;
    for (uint32_t v = 0; v < nvars; v++) g->adj_start[v + 1] += g->adj_start[v];
    uint32_t tot = g->adj_start[nvars];
    g->adj_idx = (uint32_t *)malloc(sizeof(uint32_t) * (tot ? tot : 1));
    g->adj_j = (double *)malloc(sizeof(double) * (tot ? tot : 1));
    uint32_t *fill = (uint32_t *)calloc(nvars, sizeof(uint32_t));
    for (uint32_t e = 0; e < nedges; e++) { /* push order :71-74 */
        uint32_t a = ea[e], b = eb[e];
        g->adj_idx[g->adj_start[a] + fill[a]] = b, g->adj_j[g->adj_start[a] + fill[a]++] = J[e];
        g->adj_idx[g->adj_start[b] + fill[b]] = a, g->adj_j[g->adj_start[b] + fill[b]++] = J[e];
    }
    free(fill);
    for (uint32_t v = 0; v < nvars; v++) { /* stable sort_by_key :76-78 */
        uint32_t s = g->adj_start[v], e = g->adj_start[v + 1];
        for (uint32_t i = s + 1; i < e; i++) {
            uint32_t ki = g->adj_idx[i];
            double kj = g->adj_j[i];
            uint32_t j = i;
            while (j > s && g->adj_idx[j - 1] > ki) {
                g->adj_idx[j] = g->adj_idx[j - 1], g->adj_j[j] = g->adj_j[j - 1];
                j--;
            }
            g->adj_idx[j] = ki, g->adj_j[j] = kj;
        }
    }
    return g;
}

void orc_cls_destroy(OrcCls *g) {
    if (!g) return;
    free(g->adj_start), free(g->adj_idx), free(g->adj_j), free(g->biases), free(g->state);
    free(g->ea), free(g->eb), free(g->ej), free(g->cum_w);
    free(g);
}

/* delta_e of do_spin_flip, graph.rs:98-115 */
static double cls_delta_e(const OrcCls *g, uint32_t i) {
    int curr = g->state[i];
    double delta_e = 0.0;
    for (uint32_t k = g->adj_start[i]; k < g->adj_start[i + 1]; k++) {
        double old_coupling = !(curr ^ g->state[g->adj_idx[k]]) ? 1.0 : -1.0;
        delta_e += -2.0 * g->adj_j[k] * old_coupling;
    }
    return delta_e + (2.0 * g->biases[i] * (curr ? 1.0 : -1.0));
}

void orc_cls_spin_flips(OrcCls *g, double beta, uint64_t count) {
    for (uint64_t t = 0; t < count; t++) {
        uint32_t i = (uint32_t)gen_range_usize(&g->rng, g->nvars); /* :98 */
        double delta_e = cls_delta_e(g, i);
        int flip; /* should_flip :339-347 */
        if (delta_e > 0.0) {
            double chance = exp(-beta * delta_e);
            flip = gen_f64(&g->rng) < chance;
        } else {
            flip = 1;
        }
        if (flip) g->state[i] = !g->state[i];
    }
}

/* ---- the reference's own schedule: do_time_step with spin, edge and worm moves (graph.rs:121-406) ---- */

/* GraphState::delta_e, graph.rs:155-176: no bias term; the edge to `omit` is left out (omit < 0: none) */
static double cls_delta_e_omit(const OrcCls *g, uint32_t v, int64_t omit) {
    int curr = g->state[v];
    double delta_e = 0.0;
    for (uint32_t k = g->adj_start[v]; k < g->adj_start[v + 1]; k++) {
        if ((int64_t)g->adj_idx[k] == omit) continue;
        double old_coupling = !(curr ^ g->state[g->adj_idx[k]]) ? 1.0 : -1.0;
        delta_e += -2.0 * g->adj_j[k] * old_coupling;
    }
    return delta_e;
}

/* should_flip, graph.rs:339-347 */
static int cls_should_flip(OrcCls *g, double beta, double delta_e) {
    if (delta_e > 0.0) {
        double chance = exp(-beta * delta_e);
        return gen_f64(&g->rng) < chance;
    }
    return 1;
}

/* enable_edge_importance_sampling, graph.rs:321-336: running sums of the edge weights in edge order */
void orc_cls_enable_edge_importance_sampling(OrcCls *g, int enable) {
    free(g->cum_w);
    g->cum_w = NULL, g->total_w = 0.0;
    if (!enable) return;
    g->cum_w = (double *)malloc(sizeof(double) * (g->nedges ? g->nedges : 1));
    double acc = 0.0;
    for (uint32_t e = 0; e < g->nedges; e++) acc = acc + g->ej[e], g->cum_w[e] = acc;
    g->total_w = acc;
}

/* slice::binary_search_by (core 1.5x..): any index with an equal element, else the insertion point */
static uint32_t cls_binary_search(const double *a, uint32_t n, double p) {
    uint32_t size = n, left = 0, right = n;
    while (left < right) {
        uint32_t mid = left + size / 2;
        if (a[mid] < p) left = mid + 1;
        else if (a[mid] > p) right = mid;
        else return mid;
        size = right - left;
    }
    return left;
}

/* do_edge_flip, graph.rs:122-153 */
void orc_cls_edge_flips(OrcCls *g, double beta, uint64_t count) {
    for (uint64_t t = 0; t < count; t++) {
        uint32_t indx_edge;
        if (g->cum_w) {
            if (!(0.0 < g->total_w)) { g->error = 1; return; } /* gen_range panics on an empty range */
            double p = gen_range_f64(&g->rng, 0.0, g->total_w);
            indx_edge = cls_binary_search(g->cum_w, g->nedges, p);
        } else {
            if (g->nedges == 0) { g->error = 1; return; }
            indx_edge = (uint32_t)gen_range_usize(&g->rng, g->nedges);
        }
        if (indx_edge >= g->nedges) { g->error = 1; return; } /* index out of bounds panic */
        uint32_t va = g->ea[indx_edge], vb = g->eb[indx_edge];
        double da = cls_delta_e_omit(g, va, vb) + (2.0 * g->biases[va] * (g->state[va] ? 1.0 : -1.0));
        double db = cls_delta_e_omit(g, vb, va) + (2.0 * g->biases[vb] * (g->state[vb] ? 1.0 : -1.0));
        double delta_e = da + db;
        if (cls_should_flip(g, beta, delta_e)) g->state[va] = !g->state[va], g->state[vb] = !g->state[vb];
    }
}

/* WormMove (graph.rs:46-50): Single(a) is (a, NONE), Double(a, b) is (a, b) */
#define WM_NONE 0xFFFFFFFFu
typedef struct { uint32_t a, b; } Worm;
static double worm_delta_e(const OrcCls *g, Worm m) { /* :191-199 */
    if (m.b == WM_NONE) return cls_delta_e_omit(g, m.a, -1);
    double de = cls_delta_e_omit(g, m.a, m.b);
    return de + cls_delta_e_omit(g, m.b, m.a);
}
static int cmp_u32(const void *x, const void *y) {
    uint32_t a = *(const uint32_t *)x, b = *(const uint32_t *)y;
    return a < b ? -1 : a > b;
}
#define EPS_F64 2.220446049250313e-16

/* do_worm_flip, graph.rs:179-318 */
void orc_cls_worm_flips(OrcCls *g, double beta, uint64_t count, int allow_doubles) {
    const uint32_t N = g->nvars;
    uint32_t maxdeg = 0;
    for (uint32_t v = 0; v < N; v++)
        if (g->adj_start[v + 1] - g->adj_start[v] > maxdeg) maxdeg = g->adj_start[v + 1] - g->adj_start[v];
    Worm *path = (Worm *)malloc(sizeof(Worm) * ((size_t)N + 2));
    Worm *sm = (Worm *)malloc(sizeof(Worm) * ((size_t)maxdeg * (maxdeg + 1) + 1));
    double *sde = (double *)malloc(sizeof(double) * ((size_t)maxdeg * (maxdeg + 1) + 1));
    uint32_t *flat = (uint32_t *)malloc(sizeof(uint32_t) * 2 * ((size_t)N + 2));
    for (uint64_t t = 0; t < count; t++) {
        uint32_t start_index = (uint32_t)gen_range_usize(&g->rng, N); /* :187 */
        size_t plen = 0;
        path[plen++] = (Worm){start_index, WM_NONE};
        uint32_t last_index = start_index;
        double starting_e = worm_delta_e(g, (Worm){start_index, WM_NONE});
        g->state[start_index] = !g->state[start_index];
        int update_failed = 0;
        for (;;) {
            uint32_t ns = 0;
            Worm sel_move = path[plen - 1];
            uint32_t sel_var = sel_move.b == WM_NONE ? sel_move.a : sel_move.b;
            int any_resolve = 0;
            for (uint32_t k = g->adj_start[sel_var]; k < g->adj_start[sel_var + 1]; k++) { /* :214-242 */
                uint32_t ov = g->adj_idx[k];
                if (ov == last_index) continue;
                double de = worm_delta_e(g, (Worm){ov, WM_NONE});
                if (fabs(de) < EPS_F64) {
                    sm[ns] = (Worm){ov, WM_NONE}, sde[ns++] = de;
                } else if (fabs(de + starting_e) < EPS_F64) {
                    sm[ns] = (Worm){ov, WM_NONE}, sde[ns++] = de;
                    any_resolve = 1;
                }
                if (allow_doubles) { /* jumps from ov as if ov were flipped :225-240 */
                    g->state[ov] = !g->state[ov];
                    for (uint32_t kk = g->adj_start[ov]; kk < g->adj_start[ov + 1]; kk++) {
                        uint32_t oov = g->adj_idx[kk];
                        if (oov != ov && oov != sel_var) {
                            double de2 = worm_delta_e(g, (Worm){oov, WM_NONE}) + de;
                            if (fabs(de2) < EPS_F64) {
                                sm[ns] = (Worm){ov, oov}, sde[ns++] = de2;
                            } else if (fabs(de2 + starting_e) < EPS_F64) {
                                sm[ns] = (Worm){ov, oov}, sde[ns++] = de2;
                                any_resolve = 1;
                            }
                        }
                    }
                    g->state[ov] = !g->state[ov];
                }
            }
            if (any_resolve) { /* retain :243-245 */
                uint32_t w = 0;
                for (uint32_t i = 0; i < ns; i++)
                    if (fabs(sde[i] + starting_e) < EPS_F64) sm[w] = sm[i], sde[w++] = sde[i];
                ns = w;
            }
            Worm ov;
            double de;
            if (ns) { /* :247-251 */
                uint32_t choice = (uint32_t)gen_range_usize(&g->rng, ns);
                ov = sm[choice], de = sde[choice];
                path[plen++] = ov;
            } else { /* turn around, undo the last move :252-262 */
                ov = sel_move.b == WM_NONE ? sel_move : (Worm){sel_move.b, sel_move.a};
                path[plen++] = ov;
                de = worm_delta_e(g, ov);
            }
            g->state[ov.a] = !g->state[ov.a]; /* :263-271 */
            if (ov.b != WM_NONE) g->state[ov.b] = !g->state[ov.b];
            if (ov.b != WM_NONE) last_index = ov.a; /* :272-276, against the move selected at the top of the round */
            else last_index = sel_move.b == WM_NONE ? sel_move.a : sel_move.b;
            if (fabs(de + starting_e) < EPS_F64) break; /* back at the initial energy :277-280 */
            if (plen > N) { /* :282-285 */
                update_failed = 1;
                break;
            }
        }
        size_t nf = 0; /* :287-298: flatten, sort, drop pairs (a variable visited twice is back where it was) */
        for (size_t i = 0; i < plen; i++) {
            flat[nf++] = path[i].a;
            if (path[i].b != WM_NONE) flat[nf++] = path[i].b;
        }
        qsort(flat, nf, sizeof(uint32_t), cmp_u32);
        size_t ii = 0, jj = 0; /* util/vec_help.rs:2-24 */
        while (jj + 1 < nf) {
            if (flat[jj] == flat[jj + 1]) jj += 2;
            else flat[ii++] = flat[jj++];
        }
        if (jj < nf) flat[ii++] = flat[jj++];
        nf = ii;
        int undo = update_failed;
        if (!update_failed) { /* :300-311 */
            double total_he = 0.0;
            for (size_t i = 0; i < nf; i++) total_he += 2.0 * g->biases[flat[i]] * (g->state[flat[i]] ? 1.0 : -1.0);
            undo = !cls_should_flip(g, beta, total_he);
        }
        if (undo)
            for (size_t i = 0; i < nf; i++) g->state[flat[i]] = !g->state[flat[i]];
    }
    free(path), free(sm), free(sde), free(flat);
}

/* do_time_step, graph.rs:350-406.  A count of UINT64_MAX stands for None. */
int orc_cls_do_time_step(OrcCls *g, double beta, uint64_t nspinupdates, uint64_t nedgeupdates,
                         uint64_t nwormupdates, int only_basic_moves) {
    if (nspinupdates == UINT64_MAX) nspinupdates = g->nvars / 2 > 1 ? g->nvars / 2 : 1;
    if (nedgeupdates == UINT64_MAX) nedgeupdates = g->nedges / 2 > 1 ? g->nedges / 2 : 1;
    if (nwormupdates == UINT64_MAX) nwormupdates = 1;
    uint32_t t = only_basic_moves ? 2 : 3;
    uint32_t choice = gen_range_u8(&g->rng, t); /* :365 */
    if (choice == 0) orc_cls_spin_flips(g, beta, nspinupdates);
    else if (choice == 1) orc_cls_edge_flips(g, beta, nedgeupdates);
    else orc_cls_worm_flips(g, beta, nwormupdates, 1);
    return (int)choice;
}
int orc_cls_get_error(const OrcCls *g) { return g->error; }
void orc_cls_set_cursor(OrcCls *g, uint64_t cursor) { g->rng.cursor = cursor, g->rng.blk_valid = 0; }

uint64_t orc_cls_threshold(double beta, double delta_e) {
    if (!(delta_e > 0.0)) return 4294967296ull;
    double chance = exp(-beta * delta_e);
    double scaled = ceil(chance * 4294967296.0); /* #{d : d * 2^-32 < chance} */
    if (scaled >= 4294967296.0) return 4294967296ull;
    return (uint64_t)scaled;
}

/* Checkerboard schedule (builder-defined; the reference has none).  Per site the rule is
 * graph.rs:98-118 + :339-347 with a fixed-width 32-bit draw d.  Draws are defined bit-sliced so that
 * a SIMD implementation can compare 32 sites at once: the sites of colour c are grouped by rank,
 * 32 consecutive ranks per group; bit-plane k (k = 0 is the most significant bit of d) of group g is
 * output word (k & 3) of Philox4x32-10(key, ctr = (4 g + ((k & 15) >> 2), sweep_lo, sweep_hi, tag | c)),
 * tag = 'CB' << 16 for planes 0..15 and 'CC' << 16 for planes 16..31; the draw of rank r is
 * d = sum_k bit (r & 31) of plane_k << (31 - k).  The site flips iff delta_e <= 0 or
 * d * 2^-32 < exp(-beta * delta_e).  (The low 16 planes only matter when the high 16 tie.) */
static void cb_planes(const uint32_t k[2], uint32_t group, uint64_t sweep, uint32_t c, uint32_t planes[32]) {
    for (uint32_t q = 0; q < 8; q++) {
        uint32_t ctr[4] = {4 * group + (q & 3u), (uint32_t)sweep, (uint32_t)(sweep >> 32), (q < 4 ? TAG_CB : TAG_CB2) | c};
        orc_philox4x32_10(ctr, k, planes + 4 * q);
    }
}
static uint32_t cb_draw_from_planes(const uint32_t planes[32], uint32_t j) {
    uint32_t d = 0;
    for (uint32_t kk = 0; kk < 32; kk++) d |= ((planes[kk] >> j) & 1u) << (31 - kk);
    return d;
}

void orc_cls_checkerboard_sweeps(OrcCls *g, double beta, const uint32_t *colours,
                                 uint32_t ncolours, uint64_t nsweeps) {
    uint32_t k[2] = {(uint32_t)g->rng.key, (uint32_t)(g->rng.key >> 32)};
    for (uint64_t s = 0; s < nsweeps; s++, g->sweep++) {
        for (uint32_t c = 0; c < ncolours; c++) {
            uint32_t rank = 0;
            uint32_t planes[32];
            for (uint32_t i = 0; i < g->nvars; i++) {
                if (colours[i] != c) continue;
                if ((rank & 31u) == 0) cb_planes(k, rank >> 5, g->sweep, c, planes);
                uint32_t d = cb_draw_from_planes(planes, rank & 31u);
                rank++;
                double delta_e = cls_delta_e(g, i);
                int flip;
                if (delta_e > 0.0) {
                    double chance = exp(-beta * delta_e);
                    flip = (double)d * (1.0 / 4294967296.0) < chance;
                } else {
                    flip = 1;
                }
                if (flip) g->state[i] = !g->state[i];
            }
        }
    }
}

/* get_energy, graph.rs:430-447 */
double orc_cls_energy(const OrcCls *g) {
    double acc = 0.0;
    for (uint32_t i = 0; i < g->nvars; i++) {
        int si = g->state[i];
        double total_e = 0.0;
        for (uint32_t k = g->adj_start[i]; k < g->adj_start[i + 1]; k++) {
            double old_coupling = !(si ^ g->state[g->adj_idx[k]]) ? 1.0 : -1.0;
            total_e += g->adj_j[k] * old_coupling / 2.0;
        }
        double bias_e = si ? -g->biases[i] : g->biases[i];
        acc = acc + total_e + bias_e;
    }
    return acc;
}

double orc_cls_magnetization(const OrcCls *g) {
    int64_t m = 0;
    for (uint32_t i = 0; i < g->nvars; i++) m += g->state[i] ? 1 : -1;
    return (double)m / (double)g->nvars;
}
void orc_cls_get_state(const OrcCls *g, uint8_t *out) { memcpy(out, g->state, g->nvars); }
void orc_cls_set_state(OrcCls *g, const uint8_t *in) { memcpy(g->state, in, g->nvars); }
uint64_t orc_cls_get_cursor(const OrcCls *g) { return g->rng.cursor; }
uint64_t orc_cls_get_sweep(const OrcCls *g) { return g->sweep; }
void orc_cls_set_sweep(OrcCls *g, uint64_t sweep) { g->sweep = sweep; }

void orc_cls_batch_checkerboard(OrcCls **reps, uint32_t nreps, const double *betas,
                                const uint32_t *colours, uint32_t ncolours, uint64_t nsweeps,
                                int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
    for (uint32_t r = 0; r < nreps; r++)
        orc_cls_checkerboard_sweeps(reps[r], betas[r], colours, ncolours, nsweeps);
    (void)nthreads;
}

void orc_cls_batch_spin_flips(OrcCls **reps, uint32_t nreps, const double *betas, uint64_t count,
                              int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
    for (uint32_t r = 0; r < nreps; r++) orc_cls_spin_flips(reps[r], betas[r], count);
    (void)nthreads;
}
