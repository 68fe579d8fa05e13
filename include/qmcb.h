/*
 * qmcb.h -- C ABI of the B200-native replica engine (libqmcb.so).
 *
 * This is the drop-in boundary for the data-parallel hot path of Renmusxd/IsingMonteCarlo
 * (crate `qmc` 2.20.0).  The reference is pure safe Rust with no FFI of its own, so each entry
 * point below names the reference interface it replaces (file:line into the reference tree);
 * INTEGRATION.md shows the `-sys` crate binding a maintainer would add.  One handle is a BATCH:
 * R independent replicas (chains / tempering temperatures) of one lattice, resident on one GPU.
 *
 * Conventions: plain pointers and sizes only; every function returns QMCB_OK (0) or a negative
 * QmcbStatus and never aborts or throws; `qmcb_last_error()` gives the message of the last
 * failure on the calling thread.  Host pointers unless a name ends in `_dev`.  A handle is not
 * thread-safe; distinct handles may be used from distinct threads.  There is no CPU fallback:
 * creating a handle without a CUDA device fails with QMCB_ERR_CUDA.
 */
#ifndef QMCB_H
#define QMCB_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    QMCB_OK = 0,
    QMCB_ERR_BAD_ARG = -1,
    QMCB_ERR_CAPACITY = -2, /* operator string would outgrow `capacity` (see qmcb_create) */
    QMCB_ERR_CUDA = -3,
    QMCB_ERR_UNSUPPORTED = -4,
    QMCB_ERR_INTERNAL = -5, /* a device-side invariant failed (reference: unreachable!/panic) */
    QMCB_ERR_NCCL = -6      /* libnccl missing or a collective failed (multi-GPU tempering only) */
} QmcbStatus;

/* Cluster-update order (DESIGN.md "Two orders").
 * STRICT: clusters are numbered in the reference's LIFO discovery order and draw k of the block
 *         goes to cluster k (cluster.rs:57-97,129-137) -- bit-exact with the reference's update
 *         path under the same injected stream.
 * FAST:   clusters are keyed by their smallest segment id and flip on a counter-based Philox bit;
 *         same Markov kernel, labelled by a parallel union-find.  Bit-exact with the oracle's
 *         FAST mode; statistically equivalent to STRICT.
 * COUNTER: the FAST cluster order, plus a diagonal update whose uniform words are one Philox block per SLOT,
 *         Philox(key, (p, c_lo, c_hi, 'DIAG')) with c the stream cursor at the start of the step, instead of
 *         positions in the sequential stream: the acceptance arithmetic is the reference's (diagonal.rs:142-191),
 *         only n couples the slots.  Bit-exact with the oracle's COUNTER mode; statistically equivalent to STRICT
 *         (exact diagonalisation tests).  Metropolis rule only (no heat-bath contract). */
#define QMCB_MODE_STRICT 0
#define QMCB_MODE_FAST 1
#define QMCB_MODE_COUNTER 2

/* Operator word (one uint32 per slot p of the operator string; replaces BasicOp,
 * op_container.rs:224-237): bits 0..23 bond index, bits 24,25 input spins of leg 0,1,
 * bits 26,27 output spins, bits 28..31 zero.  The identity is QMCB_OP_EMPTY.  vars / nvars /
 * `constant` are functions of the bond index exactly as in qmc_ising.rs:671-681:
 * [0,E) two-site bonds, [E,E+N) transverse ops, [E+N,E+2N) longitudinal ops. */
#define QMCB_OP_EMPTY 0xFFFFFFFFu

typedef struct {
    uint32_t nvars;      /* N; qmc_ising.rs:92 (max index + 1) */
    uint32_t nedges;     /* E */
    const uint32_t *va;  /* [E] first variable of each edge */
    const uint32_t *vb;  /* [E] second variable */
    const double *J;     /* [E] coupling; H = sum J sz sz - Gamma sum sx - h sum sz */
    double transverse;   /* Gamma >= 0 */
    double longitudinal; /* h */
} QmcbLattice;

typedef struct QmcbHandle QmcbHandle;

/* ---- construction: QmcIsingGraph::new_with_rng (qmc_ising.rs:131-148, :80-128) ---------- */
/* betas[R]; rng_keys[R]: Philox4x32-10 key of each replica's injected stream (word c of the
 * stream is what a Rust `RngCore::next_u64` shim returns on its c-th call).
 * cutoff0: initial cutoff M (reference `cutoff` argument).  capacity: slots allocated per replica
 * (>= cutoff0; 0 = choose automatically and grow on demand).  init_state: [R*N] bytes (0/1) or
 * NULL to draw it from each stream as make_random_spin_state does (classical/graph.rs:451-453).
 * device: CUDA ordinal. */
int qmcb_create(const QmcbLattice *lattice, uint32_t n_replicas, const double *betas,
                const uint64_t *rng_keys, uint64_t cutoff0, uint64_t capacity,
                const uint8_t *init_state, int device, QmcbHandle **out);
/* ---- generic interactions: Qmc (qmc_runner.rs:22-403), IntoQmc::into_qmc (qmc_ising.rs:943-976) -------------------
 * Interactions as the reference's make_interaction / make_diagonal_interaction take them: matrices in Interaction::at
 * indexing (qmc_runner.rs:560-664: first variable most significant, outputs more significant than inputs), 2^nv entries
 * for a diagonal interaction, 4^nv for a full one.  The weights of the diagonal update come from these tables (a
 * different code path from the (J, Gamma, h) arithmetic of qmcb_create); Qmc::timestep = diagonal update, loop update
 * when do_loop_updates, cluster update with Ising symmetry, free-spin flips (:363-377).  Supported shape (anything else:
 * QMCB_ERR_UNSUPPORTED with the reason): E interactions of two variables, each symmetric under the global flip, followed
 * by either exactly one CONSTANT one-variable interaction per variable, in variable order -- what into_qmc builds, with
 * arbitrary two-variable weights and per-variable transverse weights -- or by no one-variable interaction at all (no
 * cluster edges: the model moves by loop updates only).
 * Loop updates (directed_loop.rs:103-301, Qmc::new(.., do_loop_updates) / set_do_loop_updates / loop_update): the
 * two-variable matrices may then carry off-diagonal elements (exchange terms).  A handle with loop updates on, with
 * such elements, or without cluster edges runs in QMCB_MODE_STRICT only, one lane per replica on the reference's link
 * structure (the walk of a directed loop is one chain; replicas are the parallel axis).
 * `offset`: what Qmc::get_offset returns (the *_and_offset constructors accumulate it on the host side). */
typedef struct {
    uint32_t nvars;
    uint32_t n_interactions;
    const uint32_t *nv;      /* [n] variables of each interaction (1 or 2) */
    const uint32_t *vars;    /* [2 n] */
    const uint32_t *mat_len; /* [n] 2^nv (diagonal) or 4^nv (full) */
    const double *mats;      /* the matrices, concatenated */
    double offset;
    int do_loop_updates;     /* Qmc::new's flag (qmc_runner.rs:48) */
} QmcbInteractions;
int qmcb_create_qmc(const QmcbInteractions *interactions, uint32_t n_replicas, const double *betas, const uint64_t *rng_keys,
                    uint64_t cutoff0, uint64_t capacity, const uint8_t *init_state, int device, QmcbHandle **out);
int qmcb_destroy(QmcbHandle *h);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = handle's own stream */
int qmcb_set_stream(QmcbHandle *h, void *cuda_stream);
int qmcb_set_mode(QmcbHandle *h, int mode);
int qmcb_get_mode(const QmcbHandle *h, int *mode);
/* QmcIsingGraph::set_enable_heatbath (qmc_ising.rs:444-486): use the heat-bath diagonal update
 * (heatbath.rs:149-209; bond chosen by cumulative maximum weight, BondWeights :10-61) instead of
 * the Metropolis rule (diagonal.rs:142-191).  Same injected stream, reference draw order:
 * gen_bool, then gen_range(0. ..1.0) and gen_range(0. ..total) when an insertion is attempted. */
int qmcb_set_enable_heatbath(QmcbHandle *h, int enable);
int qmcb_get_enable_heatbath(const QmcbHandle *h, int *enabled);
/* Replicas with unequal Hamiltonians in one batch (what the reference expresses as graphs with different
 * couplings in one TemperingContainer, tempering_traits.rs:122-154).  J_tab [n_ham][E], transverse[n_ham],
 * longitudinal[n_ham]; ham_of_replica[R] picks each replica's row.  Rows must satisfy can_swap_managers
 * (qmc_ising.rs:563-590): couplings of the same sign per edge, longitudinal fields of the same sign. */
int qmcb_set_hamiltonians(QmcbHandle *h, uint32_t n_ham, const double *J_tab, const double *transverse,
                          const double *longitudinal, const uint32_t *ham_of_replica);
int qmcb_num_hamiltonians(const QmcbHandle *h, uint32_t *n_ham);
int qmcb_get_hamiltonian_index(QmcbHandle *h, uint32_t *ham_of_replica /* [R] */);
int qmcb_get_offsets(QmcbHandle *h, double *offsets /* [R]: get_offset of each replica's Hamiltonian */);
/* tuning knobs.  Per handle: "impl" 0 = warp-parallel kernels where available (default), 1 = serial-order kernels only;
 * "auto_capacity" 1 = grow the strings on demand even if a capacity was given; "debug_counters".  Kernel selection
 * (also per handle; for measurements and tests): "minblocks" 0 = choose the register budget from the batch shape
 * (default), 4/6/7/8 = force that build; "shared_edge_table" 0/1; "pipeline" 0/1 (two warps per replica when few
 * replicas are resident); "smem_pad", "smem_carveout" (experiments). */
int qmcb_set_option(QmcbHandle *h, const char *name, int64_t value);
/* event counters and phase timers of the SSE kernels (after qmcb_set_option(h, "debug_counters", 1)); diagnostics
 * only, and only counted by a library built with -DQMCB_PHASE_TIMERS (the production build compiles them out) */
int qmcb_get_debug_counters(QmcbHandle *h, uint64_t *out64 /* [64] */);
int qmcb_set_betas(QmcbHandle *h, const double *betas);
int qmcb_get_betas(const QmcbHandle *h, double *betas);
int qmcb_num_replicas(const QmcbHandle *h, uint32_t *r);
int qmcb_num_vars(const QmcbHandle *h, uint32_t *n);
int qmcb_num_bonds(const QmcbHandle *h, uint32_t *nb); /* qmc_ising.rs:664-670 */
int qmcb_num_edges(const QmcbHandle *h, uint32_t *ne);
/* get_edges / get_transverse_field / get_longitudinal_field (qmc_ising.rs:496-535) */
int qmcb_get_edges(const QmcbHandle *h, uint32_t *va /* [E] */, uint32_t *vb, double *J);
int qmcb_get_fields(const QmcbHandle *h, double *transverse, double *longitudinal);

/* ---- stepping: QmcStepper (qmc_stepper.rs:2-168) ---------------------------------------- */
/* timesteps_measure_with_self (qmc_stepper.rs:133-162): t sweeps (QmcIsingGraph::timestep,
 * qmc_ising.rs:644-795) of every replica at its beta; after sweep i (1-based) with
 * i % sampling_freq == 0 the state is sampled and n accumulated.  energy_out[R] (or NULL) gets
 * -(mean n)/beta + offset (qmc_ising.rs:805-809; NaN if nothing was sampled, as the reference).
 * samples_out (or NULL): [R][t / sampling_freq][N] bytes, the `Vec<Vec<bool>>` of
 * timesteps_sample (qmc_stepper.rs:23-40).  sampling_freq 0 means 1. */
int qmcb_timesteps(QmcbHandle *h, uint64_t t, uint64_t sampling_freq, double *energy_out,
                   uint8_t *samples_out);
/* launch-only variant: enqueue t sweeps on the handle's stream and return without synchronising
 * (for device-side timing); energy of these sweeps is folded into the next qmcb_timesteps call
 * only through the vertex-update counter below. */
int qmcb_enqueue_sweeps(QmcbHandle *h, uint64_t t);
int qmcb_synchronize(QmcbHandle *h);
/* single_diagonal_step / single_cluster_step (qmc_ising.rs:208-270, :273-320) */
int qmcb_single_diagonal_step(QmcbHandle *h);
int qmcb_single_cluster_step(QmcbHandle *h, uint64_t *n_clusters_out /* [R] or NULL */);
/* Qmc::loop_update (qmc_runner.rs:205-220 -> LoopUpdater::make_loop_update_with_rng, directed_loop.rs:103-171): one
 * directed-loop update of every replica; Qmc::set_do_loop_updates / should_do_loop_update (:268-275).  Handles made by
 * qmcb_create_qmc only. */
int qmcb_loop_update(QmcbHandle *h);
int qmcb_set_do_loop_updates(QmcbHandle *h, int enable);
int qmcb_get_do_loop_updates(const QmcbHandle *h, int *enabled);
/* RVB update (RvbUpdater::rvb_update_with_ising_weight, rvb.rs:60-291, with build_cluster :1054-1122, calculate_flip_prob
 * :649-946, mutate_graph :294-616 and BondContainer util/bondcontainer.rs).  qmcb_set_run_rvb = QmcIsingGraph::set_run_rvb
 * (qmc_ising.rs:434-441): every sweep then runs (nvars + 1) / 2 RVB updates between the diagonal and the cluster update
 * (:705-752), in every mode (the update draws from the replica's sequential stream).  qmcb_single_rvb_sweep =
 * single_rvb_sweep (:322-420; updates_in_sweep < 0 = None); qmcb_rvb_success_rate = rvb_success_rate (:604-607), any
 * output may be NULL.  Handles made by qmcb_create only (Qmc has no RVB step).  The flag travels in the checkpoint blob
 * (the two counters restart at 0); under tempering the counters follow the configuration, not the slot. */
int qmcb_set_run_rvb(QmcbHandle *h, int run_rvb);
int qmcb_get_run_rvb(const QmcbHandle *h, int *run_rvb);
int qmcb_single_rvb_sweep(QmcbHandle *h, int64_t updates_in_sweep, uint64_t *successes_out /* [R] or NULL */, uint64_t *attempts_out /* or NULL */);
int qmcb_rvb_success_rate(QmcbHandle *h, double *rate_out /* [R] */, uint64_t *successes_out /* [R] */, uint64_t *counted_out /* [R] */);
/* sum over replicas and sweeps so far of n after each sweep (the metric's "vertex updates") */
int qmcb_total_vertex_updates(QmcbHandle *h, uint64_t *total);
/* number of kernels this handle has launched so far */
int qmcb_launch_count(const QmcbHandle *h, uint64_t *launches);

/* ---- accessors (qmc_ising.rs:496-560, qmc_stepper.rs:5-14) ------------------------------- */
int qmcb_get_state(QmcbHandle *h, uint32_t r, uint8_t *state /* [N] */);   /* state_ref */
int qmcb_get_states(QmcbHandle *h, uint8_t *states /* [R*N] */);
int qmcb_set_state(QmcbHandle *h, uint32_t r, const uint8_t *state);
int qmcb_get_n(QmcbHandle *h, uint64_t *n /* [R] */);                       /* get_n */
int qmcb_get_cutoffs(QmcbHandle *h, uint64_t *cutoffs /* [R] */);          /* get_cutoff */
int qmcb_set_cutoff(QmcbHandle *h, uint32_t r, uint64_t cutoff);           /* set_cutoff :537-540 */
int qmcb_get_capacity(const QmcbHandle *h, uint64_t *capacity);
int qmcb_get_offset(const QmcbHandle *h, double *offset);                  /* get_offset */
int qmcb_get_bond_counts(QmcbHandle *h, uint32_t r, uint64_t *counts /* [num_bonds] */);
/* imaginary_time_fold (qmc_stepper.rs:165-168, qmc_ising.rs:815-821, OpContainer::itime_fold fast_ops.rs:1296-1315).
 * A C ABI cannot take the reference's closure, so the fold is offered two ways: (1) on the device with the
 * magnetisation fold, for every replica at once: per-site m = mean_v (2 s_v - 1) of the propagated state before each
 * of the M slots, averaged over the slots as <m>, <m^2>, <|m|> (any output may be NULL); (2) the propagated state
 * before slot p of one replica, for folds evaluated by the caller. */
int qmcb_itime_magnetization(QmcbHandle *h, double *m_mean /* [R] */, double *m_sq, double *m_abs);
int qmcb_itime_state(QmcbHandle *h, uint32_t r, uint64_t p, uint8_t *state /* [N] */);
/* QmcAutoCorrelations::calculate_variable_autocorrelation (autocorrelations.rs:48-61; fft_autocorrelation :99-133):
 * runs t sweeps sampling every sampling_freq (T = t / sampling_freq samples), returns per replica the circular
 * autocorrelation of the mean-removed, normalised +-1 series of every variable, averaged over the variables
 * (autocorr_out [R][T], entry 0 is 1).  Computed exactly on the device from bit-packed time series instead of an
 * FFT.  samples_out ([R][T][N] bytes) and energy_out ([R]) are optional, as in qmcb_timesteps. */
int qmcb_variable_autocorrelation(QmcbHandle *h, uint64_t t, uint64_t sampling_freq, double *autocorr_out,
                                  uint8_t *samples_out, double *energy_out);
/* calculate_spin_product_autocorrelation (autocorrelations.rs:53-71): the series are products of spins; product k is over
 * product_vars[product_offsets[k] .. product_offsets[k+1]).  calculate_bond_autocorrelation (:80-97, value_for_bond
 * qmc_ising.rs:988-997): one series per edge (whether the bond is satisfied).  autocorr_out [R][T] as above. */
int qmcb_spin_product_autocorrelation(QmcbHandle *h, uint64_t t, uint64_t sampling_freq, uint32_t n_products,
                                      const uint32_t *product_offsets /* [n_products + 1] */, const uint32_t *product_vars,
                                      double *autocorr_out, uint8_t *samples_out, double *energy_out);
int qmcb_bond_autocorrelation(QmcbHandle *h, uint64_t t, uint64_t sampling_freq, double *autocorr_out, uint8_t *samples_out,
                              double *energy_out);
/* ParallelTemperingAutocorrelations::calculate_variable_autocorrelation / calculate_spin_product_autocorrelation and
 * ParallelTemperingBondAutoCorrelations::calculate_bond_autocorrelation for a TemperingContainer
 * (tempering_container.rs:484-630): parallel_timesteps_sample (:411-453) with a tempering step every replica_swap_freq
 * sweeps (0 = None = 1), then one autocorrelation per ladder SLOT over the samples taken at that slot.  autocorr_out
 * [S][T], samples_out [S][T][nvars] in slot order or NULL, energy_out [S] = the reference's energy_acc or NULL.  The
 * whole ladder must live in this handle (QMCB_ERR_UNSUPPORTED otherwise). */
int qmcb_pt_variable_autocorrelation(QmcbHandle *h, uint64_t timesteps, uint64_t replica_swap_freq, uint64_t sampling_freq,
                                     double *autocorr_out, uint8_t *samples_out, double *energy_out);
int qmcb_pt_spin_product_autocorrelation(QmcbHandle *h, uint64_t timesteps, uint64_t replica_swap_freq, uint64_t sampling_freq,
                                         uint32_t n_products, const uint32_t *product_offsets, const uint32_t *product_vars,
                                         double *autocorr_out, uint8_t *samples_out, double *energy_out);
int qmcb_pt_bond_autocorrelation(QmcbHandle *h, uint64_t timesteps, uint64_t replica_swap_freq, uint64_t sampling_freq,
                                 double *autocorr_out, uint8_t *samples_out, double *energy_out);
int qmcb_get_rng_cursors(QmcbHandle *h, uint64_t *cursors /* [R] */);
int qmcb_set_rng_cursor(QmcbHandle *h, uint32_t r, uint64_t cursor);
int qmcb_get_rng_keys(QmcbHandle *h, uint64_t *keys /* [R] */);
/* operator string of replica r in p order, nwords >= cutoff(r) entries are written */
int qmcb_dump_ops(QmcbHandle *h, uint32_t r, uint32_t *opwords, uint64_t nwords);
/* FastOps::new_from_ops (fast_ops.rs:80-87): install a string (and state if non-NULL) */
int qmcb_load_ops(QmcbHandle *h, uint32_t r, const uint32_t *opwords, uint64_t nwords,
                  const uint8_t *state);
/* OpContainer::verify + QmcIsingGraph::verify (op_container.rs:137-159, qmc_ising.rs:829-860),
 * replayed on the device; *ok = 1 iff the invariant holds for replica r. */
int qmcb_verify(QmcbHandle *h, uint32_t r, int *ok);
/* cluster ids per slot of the last STRICT cluster step (tests of cluster numbering) */
int qmcb_get_boundaries(QmcbHandle *h, uint32_t r, uint32_t *b_in, uint32_t *b_out, uint64_t nslots);

/* ---- parallel tempering: TemperingContainer (tempering_container.rs:19-302) -------------- */
/* The handle's R replicas are `n_chains` independent ladders of `n_betas` slots each
 * (slot s = chain * n_betas + k).  On several GPUs each rank holds a contiguous block of the
 * global slot range [slot_begin, slot_begin + R); ranks exchange only (n, cursor, cutoff) per slot.
 * A swap exchanges the slot LABELS (beta, rng key, rng cursor) of two configurations instead of
 * moving operator strings (equivalent to swap_manager_and_state, qmc_ising.rs:593-602). */
int qmcb_pt_configure(QmcbHandle *h, uint32_t n_chains_global, uint32_t n_betas,
                      uint32_t slot_begin, const double *betas_global /* [n_chains*n_betas] */,
                      const uint64_t *keys_global, uint64_t pt_key);
/* Hamiltonian row (qmcb_set_hamiltonians) of every slot of the ladder: a label of the slot, like beta
 * (swap_manager_and_state leaves couplings where they are).  Pairs of slots that are not ham_eq
 * (tempering_traits.rs:122-124) swap with the relative_weight factors of tempering_container.rs:286-292. */
int qmcb_pt_set_slot_hamiltonians(QmcbHandle *h, const uint32_t *ham_of_slot /* [n_chains*n_betas] */);
/* uint64 words per record of qmcb_pt_export: 4 {slot, n, cursor, cutoff}; 8 with slot Hamiltonians
 * (+ three relative weights as f64 bits: first-pass partner, second-pass partner if the configuration
 * stays, second-pass partner if it moves; + pad) */
int qmcb_pt_record_words(const QmcbHandle *h, uint32_t *words);
/* step 1 of tempering_step: write this rank's per-configuration record for the all-gather;
 * rec_dev is DEVICE memory, [R] records of qmcb_pt_record_words() x uint64. */
int qmcb_pt_export(QmcbHandle *h, uint64_t *rec_dev);
/* step 2: given ALL ranks' records (device, [n_chains*n_betas] records in any order), set every
 * cutoff to the global max (tempering_container.rs:129-137), evaluate the swaps of every ladder
 * from the shared PT stream (:140-146, :241-302) and relabel the local configurations. */
int qmcb_pt_apply(QmcbHandle *h, const uint64_t *all_rec_dev, uint64_t n_records);
/* tempering_step (tempering_container.rs:121-149) when the whole container lives on this handle (slot_begin = 0 and
 * R = n_chains * n_betas): steps 1 and 2 back to back on the device */
int qmcb_pt_step_local(QmcbHandle *h);
/* ---- the same step with the exchange inside the library (qmcb_pt_create(..., ncclComm_t) / qmcb_pt_step /
 * qmcb_pt_timesteps_sample).  NCCL is bound at run time (dlopen of libnccl.so.2, or $QMCB_NCCL_LIB); without it these
 * return QMCB_ERR_NCCL and the single-handle paths are unaffected.
 * qmcb_pt_comm_unique_id: ncclGetUniqueId -- rank 0 calls it and hands the 128 bytes to the other ranks by any means
 * (MPI, files, torch.distributed broadcast).  qmcb_pt_comm_init: ncclCommInitRank on this handle's device, collective
 * over the ranks; the handle owns the communicator.  qmcb_pt_comm_attach: use the caller's ncclComm_t instead. */
int qmcb_pt_comm_unique_id(uint8_t *id128 /* [128] */);
int qmcb_pt_comm_init(QmcbHandle *h, const uint8_t *id128, int nranks, int rank);
int qmcb_pt_comm_attach(QmcbHandle *h, void *nccl_comm /* ncclComm_t or NULL to detach */, int nranks, int rank);
/* TemperingContainer::tempering_step / parallel_tempering_step (tempering_container.rs:121-149, :373-402) wherever the
 * ladder lives: on this handle alone, or block-partitioned over the ranks of the attached communicator -- then export,
 * ONE ncclAllGather of the records (32 or 64 B per slot) and apply are enqueued on the handle's stream with no host
 * synchronisation in between. */
int qmcb_pt_step(QmcbHandle *h);
/* bytes the all-gather of one qmcb_pt_step moves per rank (0 when the ladder is local) */
int qmcb_pt_collective_bytes(const QmcbHandle *h, uint64_t *bytes_per_step);
/* TemperingContainer::timesteps_sample / parallel_timesteps_sample (tempering_container.rs:166-208, :411-453).
 * energy_acc [n_chains*n_betas], by global SLOT: sum over the segments of (mean energy of the segment x its length), the
 * reference's energy_acc (identical on every rank).  samples_out (or NULL): [R][timesteps / sampling_freq][N] bytes by
 * LOCAL configuration; sample_slots_out (or NULL) [R][timesteps / sampling_freq]: the slot each of those samples
 * belongs to (the reference files samples under the graph = slot). */
int qmcb_pt_timesteps_sample(QmcbHandle *h, uint64_t timesteps, uint64_t replica_swap_freq, uint64_t sampling_freq,
                             double *energy_acc, uint8_t *samples_out, uint32_t *sample_slots_out);
int qmcb_pt_total_swaps(QmcbHandle *h, uint64_t *swaps); /* get_total_swaps :231-233 */
int qmcb_pt_get_config(const QmcbHandle *h, uint32_t *n_chains, uint32_t *n_betas, uint32_t *slot_begin);
int qmcb_pt_get_slots(QmcbHandle *h, uint32_t *slots /* [R] current slot of each configuration */);

/* ---- checkpoints (replaces the serde path: SerializeQmcGraph qmc_ising.rs:1001-1087,
 * SerializeTemperingContainer tempering_container.rs:671-793) -------------------------------------
 * One self-describing little-endian blob per handle: lattice, mode, heat-bath flag, and per replica
 * beta, stream (key, cursor), cutoff, n, sweep counters, spins and the operator string up to its
 * cutoff; the tempering ladder, slot labels, PT stream position and swap count when configured;
 * FNV-1a checksum.  A loaded handle continues bit-identically to the one that was saved. */
int qmcb_checkpoint_size(QmcbHandle *h, uint64_t *bytes);
int qmcb_checkpoint_save(QmcbHandle *h, void *buf, uint64_t bytes);
int qmcb_checkpoint_load(const void *buf, uint64_t bytes, int device, QmcbHandle **out);

/* ---- classical graph: GraphState (classical/graph.rs:56-88, :350-447) --------------------- */
typedef struct CmcbHandle CmcbHandle;
/* edges + biases as GraphState::new (graph.rs:56-88); init_state [R*N] or NULL (stream-drawn). */
int cmcb_create(const QmcbLattice *lattice /* transverse/longitudinal ignored */,
                const double *biases /* [N] */, uint32_t n_replicas, const double *betas,
                const uint64_t *rng_keys, const uint8_t *init_state, int device, CmcbHandle **out);
int cmcb_destroy(CmcbHandle *h);
int cmcb_set_stream(CmcbHandle *h, void *cuda_stream);
/* "fused" 1 (default): the bit-packed square layout runs both colours of many sweeps in ONE launch (the blocks of a
 * replica meet at a per-replica barrier between colour passes); 0: one launch per colour pass.  Same results. */
int cmcb_set_option(CmcbHandle *h, const char *name, int64_t value);
/* nsweeps checkerboard sweeps: every site once per sweep, colour by colour, with the reference's
 * per-site rule (do_spin_flip delta_e graph.rs:98-115, should_flip :339-347). */
int cmcb_sweeps(CmcbHandle *h, uint64_t nsweeps);
int cmcb_enqueue_sweeps(CmcbHandle *h, uint64_t nsweeps);
int cmcb_synchronize(CmcbHandle *h);
int cmcb_get_state(CmcbHandle *h, uint32_t r, uint8_t *state);
int cmcb_get_states(CmcbHandle *h, uint8_t *states /* [R*N] */);
int cmcb_set_states(CmcbHandle *h, const uint8_t *states /* [R*N] */);
int cmcb_energy(CmcbHandle *h, double *energy /* [R] */);               /* get_energy :430-447 */
int cmcb_magnetization(CmcbHandle *h, double *m /* [R], mean of +-1 */);
/* The reference's OWN schedule, replica-parallel (one thread per replica, each under its sequential stream; serial
 * by construction, so this is the reference-exact path, not the throughput path):
 * GraphState::do_time_step (graph.rs:350-406): one u8 draw picks spin flips (do_spin_flip :91-119), edge flips
 * (do_edge_flip :122-153) or -- unless only_basic_moves -- worm flips with doubles (do_worm_flip :179-318); counts of
 * UINT64_MAX stand for None (max(1, N/2), max(1, |E|/2), 1).  beta is the replica's.  choices_out [R] or NULL
 * receives the move each replica drew.  The three moves are also callable directly, as the reference's tests do
 * (graph.rs:481-647).  BAD_ARG where the reference panics (edge move with no edges / non-positive total weight). */
int cmcb_do_time_step(CmcbHandle *h, uint64_t nspinupdates, uint64_t nedgeupdates, uint64_t nwormupdates,
                      int only_basic_moves, uint8_t *choices_out);
int cmcb_spin_flips(CmcbHandle *h, uint64_t count);
int cmcb_edge_flips(CmcbHandle *h, uint64_t count);
int cmcb_worm_flips(CmcbHandle *h, uint64_t count, int allow_doubles);
int cmcb_enable_edge_importance_sampling(CmcbHandle *h, int enable);    /* graph.rs:321-336 */
/* position of every replica in its sequential stream (N after a stream-drawn initial state, graph.rs:451-453) */
int cmcb_get_rng_cursors(CmcbHandle *h, uint64_t *cursors /* [R] */);
int cmcb_set_rng_cursors(CmcbHandle *h, const uint64_t *cursors /* [R] */);
int cmcb_get_colours(const CmcbHandle *h, uint32_t *colours /* [N] */, uint32_t *ncolours);
int cmcb_get_sweep_count(const CmcbHandle *h, uint64_t *sweeps);
int cmcb_set_sweep_count(CmcbHandle *h, uint64_t sweeps);
int cmcb_layout(const CmcbHandle *h, int *is_bitpacked_square);
int cmcb_launch_count(const CmcbHandle *h, uint64_t *launches);

const char *qmcb_last_error(void);
const char *qmcb_version(void);

#ifdef __cplusplus
}
#endif
#endif
