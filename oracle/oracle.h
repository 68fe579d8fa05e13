/*
 * oracle.h -- CPU restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the checker for the CUDA path, not a product path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product library (libqmcb.so) never links or calls anything here.
 *
 * It restates, in plain C, the algorithms of Renmusxd/IsingMonteCarlo (crate `qmc`
 * 2.20.0); every function in oracle.c cites the reference file:line it follows.
 *
 * PARITY STATUS: "parity unpinned" at the rand-0.8 boundary.  The reference cannot be
 * built here (no Rust toolchain) and its tests hold no golden vectors for this path
 * (SURVEY.md section 8c).  The oracle is pinned instead by (a) Philox4x32-10 Random123
 * known answers, (b) the hand-derived known answers of SURVEY.md Appendix B driven by a
 * scripted word stream, (c) exact-diagonalisation energies of small TFIM systems and
 * (d) the reference's own structural invariant `verify()`.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- RNG contract ------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint64_t orc_stream_word(uint64_t key, uint64_t cursor);
/* rand-0.8 helpers, exposed for unit tests; each consumes words W[*cursor], ... */
int orc_gen_bool(uint64_t key, uint64_t *cursor, double p);
uint64_t orc_gen_range_usize(uint64_t key, uint64_t *cursor, uint64_t n);
uint32_t orc_gen_range_u8(uint64_t key, uint64_t *cursor, uint32_t n);
double orc_gen_range_f64_01(uint64_t key, uint64_t *cursor);
double orc_gen_range_f64(uint64_t key, uint64_t *cursor, double low, double high);
double orc_gen_f64(uint64_t key, uint64_t *cursor);
int orc_gen_std_bool(uint64_t key, uint64_t *cursor);
double orc_powi(double a, int b);
uint64_t orc_bool_threshold(double p); /* (p * 2^64) as u64, the Bernoulli p_int */

/* ---- SSE transverse-field Ising replica (QmcIsingGraph<R, FastOps>) ------------ */
typedef struct OrcSse OrcSse;

#define ORC_MODE_STRICT 0 /* reference order: LIFO DFS cluster numbering, sequential draws */
#define ORC_MODE_FAST 1   /* canonical order: min-id cluster roots, counter-based flip bits */
#define ORC_MODE_COUNTER 2 /* FAST cluster order + one Philox block per SLOT for the diagonal update (oracle.c) */

#define ORC_OP_EMPTY 0xFFFFFFFFu /* op word of the identity */

OrcSse *orc_sse_create(uint32_t nvars, uint32_t nedges, const uint32_t *ea, const uint32_t *eb,
                       const double *J, double transverse, double longitudinal, uint64_t cutoff,
                       uint64_t rng_key, const uint8_t *state_or_null);
void orc_sse_destroy(OrcSse *g);
/* replace the Philox stream by a scripted list of 64-bit words (known-answer tests). */
void orc_sse_set_script(OrcSse *g, const uint64_t *words, uint64_t nwords);
int orc_sse_error(const OrcSse *g);

/* QmcIsingGraph::set_enable_heatbath (qmc_ising.rs:444-486): heat-bath diagonal update
 * (heatbath.rs:149-209) instead of the Metropolis rule */
void orc_sse_set_enable_heatbath(OrcSse *g, int enable);
int orc_sse_get_enable_heatbath(const OrcSse *g);
void orc_sse_timestep(OrcSse *g, double beta, int mode);
/* generic Qmc (qmc_runner.rs): interactions of one or two variables; handles are OrcSse handles and every orc_sse_*
 * accessor works on them */
OrcSse *orc_qmc_create(uint32_t nvars, uint64_t rng_key, const uint8_t *state_or_null);
int orc_qmc_make_interaction(OrcSse *g, const double *mat, uint32_t len, const uint32_t *vars, uint32_t nvars_given, int diagonal, int and_offset);
int orc_qmc_flags(const OrcSse *g); /* bit 0 has_cluster_edges, bit 1 breaks_ising_symmetry */
void orc_qmc_timestep(OrcSse *g, double beta, int mode);
void orc_sse_use_small_rng(OrcSse *g); /* TIMING ONLY: xoshiro256++ words (the generator the reference's benches use) instead of Philox */
/* RVB update (rvb.rs:60-291 as QmcIsingGraph::timestep calls it, qmc_ising.rs:705-752): `updates` cluster proposals,
 * returns the number accepted; set_run_rvb (:434-441), single_rvb_sweep (:322-420; updates_in_sweep < 0 = None),
 * rvb_success_rate (:604-607).  orc_sse_error() bit 64 = the reference would have panicked inside the update. */
uint64_t orc_sse_rvb_update(OrcSse *g, uint64_t updates);
void orc_sse_set_run_rvb(OrcSse *g, int run_rvb);
uint64_t orc_sse_single_rvb_sweep(OrcSse *g, int64_t updates_in_sweep, uint64_t *steps_out);
double orc_sse_rvb_success_rate(const OrcSse *g);
void orc_qmc_loop_update(OrcSse *g);                      /* Qmc::loop_update, qmc_runner.rs:205-220 -> directed_loop.rs:103-301 */
void orc_qmc_set_do_loop_updates(OrcSse *g, int enable);  /* qmc_runner.rs:268-270 */
void orc_sse_single_diagonal_step(OrcSse *g, double beta);
void orc_sse_single_diagonal_step_mode(OrcSse *g, double beta, int mode);
uint64_t orc_sse_single_cluster_step(OrcSse *g, int mode);
/* QmcStepper::timesteps_measure_with_self; samples_or_null is [t/freq][nvars] bytes. */
double orc_sse_timesteps(OrcSse *g, uint64_t t, double beta, uint64_t sampling_freq, int mode,
                         uint8_t *samples_or_null);

uint32_t orc_sse_nvars(const OrcSse *g);
uint64_t orc_sse_get_n(const OrcSse *g);
uint64_t orc_sse_get_cutoff(const OrcSse *g);
void orc_sse_set_cutoff(OrcSse *g, uint64_t cutoff);
uint64_t orc_sse_get_cursor(const OrcSse *g);
void orc_sse_set_cursor(OrcSse *g, uint64_t cursor);
void orc_sse_set_key(OrcSse *g, uint64_t key);
uint64_t orc_sse_get_key(const OrcSse *g);
double orc_sse_get_offset(const OrcSse *g);
void orc_sse_get_state(const OrcSse *g, uint8_t *out);
void orc_sse_set_state(OrcSse *g, const uint8_t *in);
uint64_t orc_sse_get_bond_count(const OrcSse *g, uint32_t bond);
/* imaginary_time_fold (qmc_ising.rs:815-821, fast_ops.rs:1296-1315) with the magnetisation fold */
uint64_t orc_sse_itime_magnetization(const OrcSse *g, int64_t sums[3]);
void orc_sse_itime_state(const OrcSse *g, uint64_t p_at, uint8_t *out);
/* op words in p order, [cutoff] entries (format: SURVEY.md Appendix D). */
void orc_sse_dump_ops(const OrcSse *g, uint32_t *words);
/* FastOps::new_from_ops equivalent: install a string given as op words + state. */
int orc_sse_load_ops(OrcSse *g, const uint32_t *words, uint64_t nwords, const uint8_t *state);
int orc_sse_verify(const OrcSse *g);
/* cluster ids (in, out) per slot from the last cluster step, -1 where no op. */
void orc_sse_get_boundaries(const OrcSse *g, int64_t *b_in, int64_t *b_out, uint64_t nslots);

/* OpenMP batch driver: one replica per thread (rayon par_iter_mut equivalent). Returns
 * the sum over replicas and sweeps of n after each sweep (the vertex-update count). */
uint64_t orc_sse_batch_timesteps(OrcSse **reps, uint32_t nreps, uint64_t t, const double *betas,
                                 int mode, double *energies_or_null, int nthreads);

/* ---- parallel tempering (TemperingContainer) ---------------------------------- */
/* one tempering_step over slots[0..nslots); swaps op strings + states between slots.
 * Returns the number of swaps performed; *pt_cursor advances over the PT stream. */
/* can_swap_managers (qmc_ising.rs:563-590): 0 ok, 1 edges differ, 2 bond signs differ, 3 field signs differ */
int orc_sse_can_swap(const OrcSse *a, const OrcSse *b);
uint64_t orc_pt_step(OrcSse **slots, uint32_t nslots, const double *betas, uint64_t pt_key,
                     uint64_t *pt_cursor);

/* ---- classical Ising graph (GraphState) --------------------------------------- */
typedef struct OrcCls OrcCls;
OrcCls *orc_cls_create(uint32_t nvars, uint32_t nedges, const uint32_t *ea, const uint32_t *eb,
                       const double *J, const double *biases, uint64_t rng_key,
                       const uint8_t *state_or_null);
void orc_cls_destroy(OrcCls *g);
/* `count` calls of GraphState::do_spin_flip (random-site Metropolis, reference schedule) */
void orc_cls_spin_flips(OrcCls *g, double beta, uint64_t count);
/* the reference's other moves and its schedule, graph.rs:121-406 (counts of UINT64_MAX = None) */
void orc_cls_enable_edge_importance_sampling(OrcCls *g, int enable);
void orc_cls_edge_flips(OrcCls *g, double beta, uint64_t count);
void orc_cls_worm_flips(OrcCls *g, double beta, uint64_t count, int allow_doubles);
int orc_cls_do_time_step(OrcCls *g, double beta, uint64_t nspinupdates, uint64_t nedgeupdates,
                         uint64_t nwormupdates, int only_basic_moves); /* returns the move type drawn */
int orc_cls_get_error(const OrcCls *g);
void orc_cls_set_cursor(OrcCls *g, uint64_t cursor);
/* checkerboard sweeps (builder-defined schedule, reference per-site rule); colours[nvars]. */
void orc_cls_checkerboard_sweeps(OrcCls *g, double beta, const uint32_t *colours,
                                 uint32_t ncolours, uint64_t nsweeps);
double orc_cls_energy(const OrcCls *g);
double orc_cls_magnetization(const OrcCls *g);
void orc_cls_get_state(const OrcCls *g, uint8_t *out);
void orc_cls_set_state(OrcCls *g, const uint8_t *in);
uint64_t orc_cls_get_cursor(const OrcCls *g);
uint64_t orc_cls_get_sweep(const OrcCls *g);
void orc_cls_set_sweep(OrcCls *g, uint64_t sweep);
/* acceptance threshold of the checkerboard contract: #{d in [0,2^32): d*2^-32 < exp(-beta*dE)} */
uint64_t orc_cls_threshold(double beta, double delta_e);
void orc_cls_batch_checkerboard(OrcCls **reps, uint32_t nreps, const double *betas,
                                const uint32_t *colours, uint32_t ncolours, uint64_t nsweeps,
                                int nthreads);
void orc_cls_batch_spin_flips(OrcCls **reps, uint32_t nreps, const double *betas, uint64_t count,
                              int nthreads);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
