//! Raw bindings to `include/qmcb.h`.  SOURCE ONLY in this repository: the build image has no
//! Rust toolchain, so this crate has not been compiled here.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct QmcbLattice {
    pub nvars: u32,
    pub nedges: u32,
    pub va: *const u32,
    pub vb: *const u32,
    pub j: *const f64,
    pub transverse: f64,
    pub longitudinal: f64,
}
#[repr(C)]
pub struct QmcbHandle { _private: [u8; 0] }
#[repr(C)]
pub struct CmcbHandle { _private: [u8; 0] }

pub const QMCB_OK: c_int = 0;
pub const QMCB_MODE_STRICT: c_int = 0;
pub const QMCB_MODE_FAST: c_int = 1;
pub const QMCB_OP_EMPTY: u32 = 0xFFFF_FFFF;

extern "C" {
    pub fn qmcb_create(lattice: *const QmcbLattice, n_replicas: u32, betas: *const f64, rng_keys: *const u64,
                       cutoff0: u64, capacity: u64, init_state: *const u8, device: c_int, out: *mut *mut QmcbHandle) -> c_int;
    pub fn qmcb_destroy(h: *mut QmcbHandle) -> c_int;
    pub fn qmcb_set_stream(h: *mut QmcbHandle, cuda_stream: *mut c_void) -> c_int;
    pub fn qmcb_set_mode(h: *mut QmcbHandle, mode: c_int) -> c_int;
    pub fn qmcb_set_enable_heatbath(h: *mut QmcbHandle, enable: c_int) -> c_int;
    pub fn qmcb_get_enable_heatbath(h: *const QmcbHandle, enabled: *mut c_int) -> c_int;
    pub fn qmcb_set_hamiltonians(h: *mut QmcbHandle, n_ham: u32, j_tab: *const f64, transverse: *const f64,
                                 longitudinal: *const f64, ham_of_replica: *const u32) -> c_int;
    pub fn qmcb_get_hamiltonian_index(h: *mut QmcbHandle, ham_of_replica: *mut u32) -> c_int;
    pub fn qmcb_get_offsets(h: *mut QmcbHandle, offsets: *mut f64) -> c_int;
    pub fn qmcb_set_betas(h: *mut QmcbHandle, betas: *const f64) -> c_int;
    pub fn qmcb_timesteps(h: *mut QmcbHandle, t: u64, sampling_freq: u64, energy_out: *mut f64, samples_out: *mut u8) -> c_int;
    pub fn qmcb_single_diagonal_step(h: *mut QmcbHandle) -> c_int;
    pub fn qmcb_single_cluster_step(h: *mut QmcbHandle, n_clusters_out: *mut u64) -> c_int;
    pub fn qmcb_get_states(h: *mut QmcbHandle, states: *mut u8) -> c_int;
    pub fn qmcb_get_state(h: *mut QmcbHandle, r: u32, state: *mut u8) -> c_int;
    pub fn qmcb_get_n(h: *mut QmcbHandle, n: *mut u64) -> c_int;
    pub fn qmcb_get_cutoffs(h: *mut QmcbHandle, cutoffs: *mut u64) -> c_int;
    pub fn qmcb_set_cutoff(h: *mut QmcbHandle, r: u32, cutoff: u64) -> c_int;
    pub fn qmcb_get_offset(h: *const QmcbHandle, offset: *mut f64) -> c_int;
    pub fn qmcb_get_bond_counts(h: *mut QmcbHandle, r: u32, counts: *mut u64) -> c_int;
    pub fn qmcb_dump_ops(h: *mut QmcbHandle, r: u32, opwords: *mut u32, nwords: u64) -> c_int;
    pub fn qmcb_load_ops(h: *mut QmcbHandle, r: u32, opwords: *const u32, nwords: u64, state: *const u8) -> c_int;
    pub fn qmcb_verify(h: *mut QmcbHandle, r: u32, ok: *mut c_int) -> c_int;
    pub fn qmcb_pt_configure(h: *mut QmcbHandle, n_chains_global: u32, n_betas: u32, slot_begin: u32,
                             betas_global: *const f64, keys_global: *const u64, pt_key: u64) -> c_int;
    pub fn qmcb_pt_set_slot_hamiltonians(h: *mut QmcbHandle, ham_of_slot: *const u32) -> c_int;
    pub fn qmcb_pt_record_words(h: *const QmcbHandle, words: *mut u32) -> c_int;
    pub fn qmcb_pt_export(h: *mut QmcbHandle, rec_dev: *mut u64) -> c_int;
    pub fn qmcb_pt_apply(h: *mut QmcbHandle, all_rec_dev: *const u64, n_records: u64) -> c_int;
    pub fn qmcb_pt_step_local(h: *mut QmcbHandle) -> c_int;
    pub fn qmcb_pt_total_swaps(h: *mut QmcbHandle, swaps: *mut u64) -> c_int;
    pub fn qmcb_itime_magnetization(h: *mut QmcbHandle, m_mean: *mut f64, m_sq: *mut f64, m_abs: *mut f64) -> c_int;
    pub fn qmcb_itime_state(h: *mut QmcbHandle, r: u32, p: u64, state: *mut u8) -> c_int;
    pub fn qmcb_variable_autocorrelation(h: *mut QmcbHandle, t: u64, sampling_freq: u64, autocorr_out: *mut f64,
                                         samples_out: *mut u8, energy_out: *mut f64) -> c_int;
    pub fn qmcb_checkpoint_size(h: *mut QmcbHandle, bytes: *mut u64) -> c_int;
    pub fn qmcb_checkpoint_save(h: *mut QmcbHandle, buf: *mut c_void, bytes: u64) -> c_int;
    pub fn qmcb_checkpoint_load(buf: *const c_void, bytes: u64, device: c_int, out: *mut *mut QmcbHandle) -> c_int;
    pub fn cmcb_create(lattice: *const QmcbLattice, biases: *const f64, n_replicas: u32, betas: *const f64,
                       rng_keys: *const u64, init_state: *const u8, device: c_int, out: *mut *mut CmcbHandle) -> c_int;
    pub fn cmcb_destroy(h: *mut CmcbHandle) -> c_int;
    pub fn cmcb_sweeps(h: *mut CmcbHandle, nsweeps: u64) -> c_int;
    pub fn cmcb_get_states(h: *mut CmcbHandle, states: *mut u8) -> c_int;
    pub fn cmcb_energy(h: *mut CmcbHandle, energy: *mut f64) -> c_int;
    pub fn qmcb_last_error() -> *const c_char;
}
