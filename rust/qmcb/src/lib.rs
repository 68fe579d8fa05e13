//! Safe wrappers with the shape of the reference's host API for the hot path, over batched GPU handles:
//! `BatchedQmcIsingGraph` + `BatchedQmcStepper` (`qmc::sse::QmcIsingGraph` / `QmcStepper`, qmc_ising.rs:131-148,
//! qmc_stepper.rs:2-168), `TemperingContainer` (tempering_container.rs:19-302, :316-478; `SwapManagers` /
//! `GraphWeights` / `StateGetter`, tempering_traits.rs:9-46, are what the handle does internally) and `GraphState`
//! (classical/graph.rs:56-88, :350-447).  One handle = R replicas, so every per-graph return value becomes a Vec.
//! SOURCE ONLY: not compiled in this repository (no Rust toolchain in the build image).
use qmcb_sys as sys;
use rand_core::{impls, Error, RngCore};
use std::ffi::CStr;

/// The injected generator of SURVEY Appendix A.3: Philox4x32-10, one 64-bit word per call.
/// Feeding this RngCore to the reference's `QmcIsingGraph::new_with_rng` reproduces, draw for
/// draw, the stream the GPU engine consumes for the replica with the same key.
pub struct PhiloxStream { pub key: u64, pub cursor: u64 }

fn philox4x32_10(mut c: [u32; 4], key: u64) -> [u32; 4] {
    let (mut k0, mut k1) = (key as u32, (key >> 32) as u32);
    for _ in 0..10 {
        let p0 = 0xD2511F53u64 * c[0] as u64;
        let p1 = 0xCD9E8D57u64 * c[2] as u64;
        c = [(p1 >> 32) as u32 ^ c[1] ^ k0, p1 as u32, (p0 >> 32) as u32 ^ c[3] ^ k1, p0 as u32];
        k0 = k0.wrapping_add(0x9E3779B9);
        k1 = k1.wrapping_add(0xBB67AE85);
    }
    c
}

impl RngCore for PhiloxStream {
    fn next_u64(&mut self) -> u64 {
        let blk = self.cursor >> 1;
        let x = philox4x32_10([blk as u32, (blk >> 32) as u32, 0, 0], self.key);
        let h = ((self.cursor & 1) * 2) as usize;
        self.cursor += 1;
        ((x[h + 1] as u64) << 32) | x[h] as u64
    }
    fn next_u32(&mut self) -> u32 { (self.next_u64() >> 32) as u32 }
    fn fill_bytes(&mut self, dest: &mut [u8]) { impls::fill_bytes_via_next(self, dest) }
    fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), Error> { self.fill_bytes(dest); Ok(()) }
}

fn check(rc: i32) -> Result<(), String> {
    if rc == sys::QMCB_OK { Ok(()) } else {
        Err(unsafe { CStr::from_ptr(sys::qmcb_last_error()) }.to_string_lossy().into_owned())
    }
}

/// Cluster / draw contract of a handle (include/qmcb.h).
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub enum Mode { Strict = 0, Fast = 1, Counter = 2 }

/// `QmcStepper` (qmc_stepper.rs:2-168) for a batch: the same method names, one value per replica.
pub trait BatchedQmcStepper {
    /// timestep(beta) -> state_ref of every replica
    fn timestep(&mut self) -> Result<Vec<Vec<bool>>, String>;
    fn timesteps(&mut self, t: usize) -> Result<Vec<f64>, String>;
    fn timesteps_sample(&mut self, t: usize, sampling_freq: Option<usize>) -> Result<(Vec<Vec<Vec<bool>>>, Vec<f64>), String>;
    fn get_n(&mut self) -> Result<Vec<u64>, String>;
    fn state_ref(&mut self) -> Result<Vec<Vec<bool>>, String>;
    fn get_bond_count(&mut self, r: usize) -> Result<Vec<u64>, String>;
    fn get_energy_for_average_n(&self, average_n: f64, beta: f64) -> f64;
}

/// R replicas of one lattice; method names follow `QmcIsingGraph` / `QmcStepper`.
pub struct BatchedQmcIsingGraph { h: *mut sys::QmcbHandle, nvars: usize, replicas: usize, offset: f64 }

impl BatchedQmcIsingGraph {
    /// qmc_ising.rs:131-148 with one rng key and beta per replica.
    pub fn new_with_rng(edges: Vec<((usize, usize), f64)>, transverse: f64, longitudinal: f64, cutoff: usize,
                        rng_keys: &[u64], betas: &[f64], state: Option<Vec<bool>>, mode: Mode, device: i32) -> Result<Self, String> {
        let nvars = edges.iter().map(|((a, b), _)| *a.max(b)).max().unwrap() + 1;
        let va: Vec<u32> = edges.iter().map(|((a, _), _)| *a as u32).collect();
        let vb: Vec<u32> = edges.iter().map(|((_, b), _)| *b as u32).collect();
        let j: Vec<f64> = edges.iter().map(|(_, j)| *j).collect();
        let lat = sys::QmcbLattice { nvars: nvars as u32, nedges: edges.len() as u32, va: va.as_ptr(), vb: vb.as_ptr(),
                                     j: j.as_ptr(), transverse, longitudinal };
        let init: Option<Vec<u8>> = state.map(|s| (0..rng_keys.len()).flat_map(|_| s.iter().map(|b| *b as u8)).collect());
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::qmcb_create(&lat, rng_keys.len() as u32, betas.as_ptr(), rng_keys.as_ptr(), cutoff as u64, 0,
                                        init.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()), device, &mut h) })?;
        let mut g = Self { h, nvars, replicas: rng_keys.len(), offset: 0.0 };
        check(unsafe { sys::qmcb_set_mode(g.h, mode as i32) })?;
        check(unsafe { sys::qmcb_get_offset(g.h, &mut g.offset) })?;
        Ok(g)
    }
    pub fn raw(&self) -> *mut sys::QmcbHandle { self.h }
    pub fn num_replicas(&self) -> usize { self.replicas }
    pub fn get_nvars(&self) -> usize { self.nvars }
    pub fn get_offset(&self) -> f64 { self.offset }
    /// get_transverse_field / get_longitudinal_field (qmc_ising.rs:522-529)
    pub fn get_fields(&self) -> Result<(f64, f64), String> {
        let (mut t, mut l) = (0f64, 0f64);
        check(unsafe { sys::qmcb_get_fields(self.h, &mut t, &mut l) })?;
        Ok((t, l))
    }
    pub fn get_transverse_field(&self) -> Result<f64, String> { Ok(self.get_fields()?.0) }
    pub fn get_longitudinal_field(&self) -> Result<f64, String> { Ok(self.get_fields()?.1) }
    pub fn set_mode(&mut self, mode: Mode) -> Result<(), String> { check(unsafe { sys::qmcb_set_mode(self.h, mode as i32) }) }
    pub fn set_betas(&mut self, betas: &[f64]) -> Result<(), String> {
        assert_eq!(betas.len(), self.replicas);
        check(unsafe { sys::qmcb_set_betas(self.h, betas.as_ptr()) })
    }
    /// QmcIsingGraph::set_enable_heatbath (qmc_ising.rs:444-486)
    pub fn set_enable_heatbath(&mut self, enable_heatbath: bool) -> Result<(), String> {
        check(unsafe { sys::qmcb_set_enable_heatbath(self.h, enable_heatbath as i32) })
    }
    /// QmcIsingGraph::set_run_rvb (qmc_ising.rs:434-441): RVB updates (rvb.rs:60-291) in every sweep
    pub fn set_run_rvb(&mut self, run_rvb: bool) -> Result<(), String> {
        check(unsafe { sys::qmcb_set_run_rvb(self.h, run_rvb as i32) })
    }
    /// QmcIsingGraph::single_rvb_sweep (qmc_ising.rs:322-420) of every replica: (successes, attempts)
    pub fn single_rvb_sweep(&mut self, updates_in_sweep: Option<usize>) -> Result<(Vec<u64>, usize), String> {
        let mut succ = vec![0u64; self.replicas];
        let mut attempts = 0u64;
        let updates = updates_in_sweep.map(|u| u as i64).unwrap_or(-1);
        check(unsafe { sys::qmcb_single_rvb_sweep(self.h, updates, succ.as_mut_ptr(), &mut attempts) })?;
        Ok((succ, attempts as usize))
    }
    /// QmcIsingGraph::rvb_success_rate (qmc_ising.rs:604-607) of every replica
    pub fn rvb_success_rate(&mut self) -> Result<Vec<f64>, String> {
        let mut rate = vec![0f64; self.replicas];
        check(unsafe { sys::qmcb_rvb_success_rate(self.h, rate.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut()) })?;
        Ok(rate)
    }
    /// single_diagonal_step / single_cluster_step (qmc_ising.rs:208-320)
    pub fn single_diagonal_step(&mut self) -> Result<(), String> { check(unsafe { sys::qmcb_single_diagonal_step(self.h) }) }
    pub fn single_cluster_step(&mut self) -> Result<Vec<u64>, String> {
        let mut n = vec![0u64; self.replicas];
        check(unsafe { sys::qmcb_single_cluster_step(self.h, n.as_mut_ptr()) })?;
        Ok(n)
    }
    pub fn get_cutoff(&mut self) -> Result<Vec<u64>, String> {
        let mut c = vec![0u64; self.replicas];
        check(unsafe { sys::qmcb_get_cutoffs(self.h, c.as_mut_ptr()) })?;
        Ok(c)
    }
    pub fn set_cutoff(&mut self, r: usize, cutoff: usize) -> Result<(), String> {
        check(unsafe { sys::qmcb_set_cutoff(self.h, r as u32, cutoff as u64) })
    }
    /// the operator string of replica r as op words (include/qmcb.h), what `get_manager_ref().get_pth(p)` walks
    pub fn dump_ops(&mut self, r: usize) -> Result<Vec<u32>, String> {
        let m = self.get_cutoff()?[r] as usize;
        let mut w = vec![0u32; m];
        check(unsafe { sys::qmcb_dump_ops(self.h, r as u32, w.as_mut_ptr(), m as u64) })?;
        Ok(w)
    }
    /// serde replacement (SerializeQmcGraph, qmc_ising.rs:1001-1087): the whole batch, stream position included
    pub fn to_bytes(&mut self) -> Result<Vec<u8>, String> {
        let mut n = 0u64;
        check(unsafe { sys::qmcb_checkpoint_size(self.h, &mut n) })?;
        let mut buf = vec![0u8; n as usize];
        check(unsafe { sys::qmcb_checkpoint_save(self.h, buf.as_mut_ptr() as *mut _, n) })?;
        Ok(buf)
    }
    pub fn from_bytes(blob: &[u8], device: i32) -> Result<Self, String> {
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::qmcb_checkpoint_load(blob.as_ptr() as *const _, blob.len() as u64, device, &mut h) })?;
        let (mut r, mut n, mut off) = (0u32, 0u32, 0.0);
        check(unsafe { sys::qmcb_num_replicas(h, &mut r) })?;
        check(unsafe { sys::qmcb_num_vars(h, &mut n) })?;
        check(unsafe { sys::qmcb_get_offset(h, &mut off) })?;
        Ok(Self { h, nvars: n as usize, replicas: r as usize, offset: off })
    }
    /// QmcStepper::imaginary_time_fold (qmc_stepper.rs:165-168) with the closure evaluated on the host
    pub fn imaginary_time_fold<F, T>(&mut self, r: usize, fold_fn: F, init: T) -> Result<T, String>
    where F: Fn(T, &[bool]) -> T {
        let cut = self.get_cutoff()?;
        let (mut acc, mut raw) = (init, vec![0u8; self.nvars]);
        for p in 0..cut[r] {
            check(unsafe { sys::qmcb_itime_state(self.h, r as u32, p, raw.as_mut_ptr()) })?;
            let st: Vec<bool> = raw.iter().map(|b| *b != 0).collect();
            acc = fold_fn(acc, &st);
        }
        Ok(acc)
    }
    /// Verify::verify (op_container.rs:137-159, qmc_ising.rs:829-860), replayed on the device
    pub fn verify(&mut self, r: usize) -> Result<bool, String> {
        let mut ok = 0;
        check(unsafe { sys::qmcb_verify(self.h, r as u32, &mut ok) })?;
        Ok(ok != 0)
    }
}

impl BatchedQmcStepper for BatchedQmcIsingGraph {
    fn timestep(&mut self) -> Result<Vec<Vec<bool>>, String> {
        check(unsafe { sys::qmcb_timesteps(self.h, 1, 1, std::ptr::null_mut(), std::ptr::null_mut()) })?;
        self.state_ref()
    }
    /// QmcStepper::timesteps (qmc_stepper.rs:17-20): average energy per replica.
    fn timesteps(&mut self, t: usize) -> Result<Vec<f64>, String> {
        let mut e = vec![0.0; self.replicas];
        check(unsafe { sys::qmcb_timesteps(self.h, t as u64, 1, e.as_mut_ptr(), std::ptr::null_mut()) })?;
        Ok(e)
    }
    /// QmcStepper::timesteps_sample (qmc_stepper.rs:23-40): samples[replica][k][var].
    fn timesteps_sample(&mut self, t: usize, sampling_freq: Option<usize>) -> Result<(Vec<Vec<Vec<bool>>>, Vec<f64>), String> {
        let f = sampling_freq.unwrap_or(1);
        let k = t / f;
        let mut e = vec![0.0; self.replicas];
        let mut raw = vec![0u8; self.replicas * k * self.nvars];
        check(unsafe { sys::qmcb_timesteps(self.h, t as u64, f as u64, e.as_mut_ptr(), raw.as_mut_ptr()) })?;
        let s = raw.chunks(k * self.nvars).map(|r| r.chunks(self.nvars).map(|c| c.iter().map(|b| *b != 0).collect()).collect()).collect();
        Ok((s, e))
    }
    fn get_n(&mut self) -> Result<Vec<u64>, String> {
        let mut n = vec![0u64; self.replicas];
        check(unsafe { sys::qmcb_get_n(self.h, n.as_mut_ptr()) })?;
        Ok(n)
    }
    fn state_ref(&mut self) -> Result<Vec<Vec<bool>>, String> {
        let mut raw = vec![0u8; self.replicas * self.nvars];
        check(unsafe { sys::qmcb_get_states(self.h, raw.as_mut_ptr()) })?;
        Ok(raw.chunks(self.nvars).map(|c| c.iter().map(|b| *b != 0).collect()).collect())
    }
    fn get_bond_count(&mut self, r: usize) -> Result<Vec<u64>, String> {
        let mut nb = 0u32;
        check(unsafe { sys::qmcb_num_bonds(self.h, &mut nb) })?;
        let mut c = vec![0u64; nb as usize];
        check(unsafe { sys::qmcb_get_bond_counts(self.h, r as u32, c.as_mut_ptr()) })?;
        Ok(c)
    }
    /// qmc_ising.rs:805-809
    fn get_energy_for_average_n(&self, average_n: f64, beta: f64) -> f64 { -(average_n / beta) + self.offset }
}
impl Drop for BatchedQmcIsingGraph { fn drop(&mut self) { unsafe { sys::qmcb_destroy(self.h); } } }

/// `qmc::sse::Qmc` (qmc_runner.rs:22-403) for a batch of replicas: interactions are collected with the reference's
/// `make_interaction*` calls (matrices in `Interaction::at` indexing) and the batch is built by `build()`; every
/// `BatchedQmcStepper` method then works on `graph_mut()`.  `loop_update` / `set_do_loop_updates` are the directed-loop
/// update of directed_loop.rs:103-301 (STRICT mode).
pub struct BatchedQmc { nvars: usize, keys: Vec<u64>, state: Option<Vec<bool>>, do_loop_updates: bool, offset: f64, cutoff: usize,
                        nv: Vec<u32>, vars: Vec<u32>, mat_len: Vec<u32>, mats: Vec<f64>, graph: Option<BatchedQmcIsingGraph> }

impl BatchedQmc {
    /// Qmc::new / new_with_state (qmc_runner.rs:48-87): cutoff = nvars, one Philox key per replica
    pub fn new(nvars: usize, rng_keys: &[u64], do_loop_updates: bool, state: Option<Vec<bool>>) -> Self {
        Self { nvars, keys: rng_keys.to_vec(), state, do_loop_updates, offset: 0.0, cutoff: nvars, nv: vec![], vars: vec![], mat_len: vec![],
               mats: vec![], graph: None }
    }
    fn add(&mut self, mut mat: Vec<f64>, vars: &[usize], diagonal: bool, and_offset: bool) -> Result<(), String> {
        if self.graph.is_some() { return Err("interactions are fixed once the batch has been built".into()); }
        let (n, tn) = (vars.len(), 1usize << vars.len());
        if n == 0 || n > 2 || mat.len() != if diagonal { tn } else { tn * tn } { return Err(format!("Given {} vars, matrix of {} entries", n, mat.len())); }
        if and_offset {  // Interaction::new_offset :508-521 / new_diagonal_offset :424-436
            let stride = if diagonal { 1 } else { tn + 1 };
            let lo = (0..tn).map(|k| mat[k * stride]).fold(f64::MAX, f64::min);
            (0..tn).for_each(|k| mat[k * stride] -= lo);
            self.offset -= lo;
        }
        if !diagonal && mat.iter().any(|x| *x < 0.0) { return Err("Interaction contains negative weights".into()); }
        self.nv.push(n as u32);
        self.vars.push(vars[0] as u32);
        self.vars.push(if n > 1 { vars[1] as u32 } else { 0 });
        self.mat_len.push(mat.len() as u32);
        self.mats.extend(mat);
        Ok(())
    }
    pub fn make_interaction(&mut self, mat: Vec<f64>, vars: Vec<usize>) -> Result<(), String> { self.add(mat, &vars, false, false) }            // :113-122
    pub fn make_interaction_and_offset(&mut self, mat: Vec<f64>, vars: Vec<usize>) -> Result<(), String> { self.add(mat, &vars, false, true) }   // :125-135
    pub fn make_diagonal_interaction(&mut self, mat: Vec<f64>, vars: Vec<usize>) -> Result<(), String> { self.add(mat, &vars, true, false) }     // :138-146
    pub fn make_diagonal_interaction_and_offset(&mut self, mat: Vec<f64>, vars: Vec<usize>) -> Result<(), String> { self.add(mat, &vars, true, true) }  // :149-156
    pub fn increase_cutoff_to(&mut self, cutoff: usize) { self.cutoff = self.cutoff.max(cutoff); }  // :307-309, before build()
    pub fn get_offset(&self) -> f64 { self.offset }
    /// qmcb_create_qmc: from here on the interactions are fixed
    pub fn build(&mut self, betas: &[f64], mode: Mode, device: i32) -> Result<&mut BatchedQmcIsingGraph, String> {
        if self.graph.is_none() {
            let ints = sys::QmcbInteractions { nvars: self.nvars as u32, n_interactions: self.nv.len() as u32, nv: self.nv.as_ptr(), vars: self.vars.as_ptr(),
                                               mat_len: self.mat_len.as_ptr(), mats: self.mats.as_ptr(), offset: self.offset,
                                               do_loop_updates: self.do_loop_updates as i32 };
            let init: Option<Vec<u8>> = self.state.as_ref().map(|s| (0..self.keys.len()).flat_map(|_| s.iter().map(|b| *b as u8)).collect());
            let mut h = std::ptr::null_mut();
            check(unsafe { sys::qmcb_create_qmc(&ints, self.keys.len() as u32, betas.as_ptr(), self.keys.as_ptr(), self.cutoff as u64, 0,
                                                init.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()), device, &mut h) })?;
            let g = BatchedQmcIsingGraph { h, nvars: self.nvars, replicas: self.keys.len(), offset: self.offset };
            check(unsafe { sys::qmcb_set_mode(g.h, mode as i32) })?;
            self.graph = Some(g);
        }
        Ok(self.graph.as_mut().unwrap())
    }
    pub fn graph_mut(&mut self) -> Option<&mut BatchedQmcIsingGraph> { self.graph.as_mut() }
    /// Qmc::loop_update (qmc_runner.rs:205-220): one directed-loop update of every replica
    pub fn loop_update(&mut self) -> Result<(), String> {
        let g = self.graph.as_ref().ok_or("build() first")?;
        check(unsafe { sys::qmcb_loop_update(g.h) })
    }
    /// Qmc::set_do_loop_updates / should_do_loop_update (qmc_runner.rs:268-275)
    pub fn set_do_loop_updates(&mut self, do_loop_updates: bool) -> Result<(), String> {
        self.do_loop_updates = do_loop_updates;
        match self.graph.as_ref() { Some(g) => check(unsafe { sys::qmcb_set_do_loop_updates(g.h, do_loop_updates as i32) }), None => Ok(()) }
    }
    pub fn should_do_loop_update(&self) -> bool { self.do_loop_updates }
}

/// `TemperingContainer` (tempering_container.rs:19-302; parallel variants :316-478) over one handle per rank:
/// `n_chains` independent ladders of `betas.len()` slots.  `add_qmc_stepper` is replaced by giving the ladder at
/// construction (all slots share one lattice, so `can_swap_graphs` holds; unequal Hamiltonians per ladder position
/// go through `qmcb_set_hamiltonians` / `qmcb_pt_set_slot_hamiltonians`).  On several GPUs every rank owns a
/// contiguous block of slots and `init_comm` makes `tempering_step` exchange the slot records with one
/// ncclAllGather inside the library.
pub struct TemperingContainer { graph: BatchedQmcIsingGraph, n_slots: usize }

impl TemperingContainer {
    #[allow(clippy::too_many_arguments)]
    pub fn new(edges: Vec<((usize, usize), f64)>, transverse: f64, longitudinal: f64, cutoff: usize, betas: &[f64], n_chains: usize,
               rng_keys: &[u64], pt_key: u64, mode: Mode, device: i32, rank: usize, world: usize) -> Result<Self, String> {
        let n_slots = betas.len() * n_chains;
        assert_eq!(rng_keys.len(), n_slots);
        assert_eq!(n_slots % world, 0, "slots must divide evenly over the ranks");
        let per = n_slots / world;
        let betas_global: Vec<f64> = (0..n_chains).flat_map(|_| betas.iter().copied()).collect();
        let (lo, hi) = (rank * per, (rank + 1) * per);
        let graph = BatchedQmcIsingGraph::new_with_rng(edges, transverse, longitudinal, cutoff, &rng_keys[lo..hi], &betas_global[lo..hi],
                                                       None, mode, device)?;
        check(unsafe { sys::qmcb_pt_configure(graph.h, n_chains as u32, betas.len() as u32, lo as u32, betas_global.as_ptr(),
                                              rng_keys.as_ptr(), pt_key) })?;
        Ok(Self { graph, n_slots })
    }
    /// rank 0: the 128-byte ncclUniqueId to hand to the other ranks (MPI, files, ...)
    pub fn comm_unique_id() -> Result<[u8; 128], String> {
        let mut id = [0u8; 128];
        check(unsafe { sys::qmcb_pt_comm_unique_id(id.as_mut_ptr()) })?;
        Ok(id)
    }
    /// collective over the ranks: ncclCommInitRank on this handle's device
    pub fn init_comm(&mut self, id: &[u8; 128], world: usize, rank: usize) -> Result<(), String> {
        check(unsafe { sys::qmcb_pt_comm_init(self.graph.h, id.as_ptr(), world as i32, rank as i32) })
    }
    pub fn num_graphs(&self) -> usize { self.n_slots }
    pub fn graph_mut(&mut self) -> &mut BatchedQmcIsingGraph { &mut self.graph }
    /// tempering_container.rs:76-81 / parallel_timesteps :366-371
    pub fn timesteps(&mut self, t: usize) -> Result<Vec<f64>, String> { self.graph.timesteps(t) }
    /// tempering_step / parallel_tempering_step (tempering_container.rs:121-149, :373-402)
    pub fn tempering_step(&mut self) -> Result<(), String> { check(unsafe { sys::qmcb_pt_step(self.graph.h) }) }
    /// timesteps_sample / parallel_timesteps_sample (tempering_container.rs:166-208, :411-453): (states, energy_acc)
    /// per SLOT; states[slot] holds the samples taken while a configuration of this rank sat in that slot.
    pub fn timesteps_sample(&mut self, timesteps: usize, replica_swap_freq: usize, sampling_freq: usize)
                            -> Result<Vec<(Vec<Vec<bool>>, f64)>, String> {
        let (r, n, t) = (self.graph.replicas, self.graph.nvars, timesteps / sampling_freq);
        let mut energy = vec![0.0; self.n_slots];
        let mut raw = vec![0u8; r * t * n];
        let mut slots = vec![0u32; r * t];
        check(unsafe { sys::qmcb_pt_timesteps_sample(self.graph.h, timesteps as u64, replica_swap_freq as u64, sampling_freq as u64,
                                                     energy.as_mut_ptr(), raw.as_mut_ptr(), slots.as_mut_ptr()) })?;
        let mut out: Vec<(Vec<Vec<bool>>, f64)> = energy.into_iter().map(|e| (Vec::new(), e)).collect();
        for k in 0..t {
            for c in 0..r {
                let s = &raw[(c * t + k) * n..(c * t + k + 1) * n];
                out[slots[c * t + k] as usize].0.push(s.iter().map(|b| *b != 0).collect());
            }
        }
        Ok(out)
    }
    /// ParallelTemperingAutocorrelations::calculate_variable_autocorrelation (tempering_container.rs:536-554): one
    /// autocorrelation per ladder slot
    pub fn calculate_variable_autocorrelation(&mut self, timesteps: usize, replica_swap_freq: Option<usize>, sampling_freq: Option<usize>)
                                              -> Result<Vec<Vec<f64>>, String> {
        let (swap, freq) = (replica_swap_freq.unwrap_or(1), sampling_freq.unwrap_or(1));
        let t = timesteps / freq;
        let mut ac = vec![0.0; self.n_slots * t];
        check(unsafe { sys::qmcb_pt_variable_autocorrelation(self.graph.h, timesteps as u64, swap as u64, freq as u64, ac.as_mut_ptr(),
                                                             std::ptr::null_mut(), std::ptr::null_mut()) })?;
        Ok(ac.chunks(t.max(1)).map(|c| c.to_vec()).collect())
    }
    /// calculate_spin_product_autocorrelation (:556-578)
    pub fn calculate_spin_product_autocorrelation(&mut self, timesteps: usize, replica_swap_freq: Option<usize>, var_products: &[&[usize]],
                                                  sampling_freq: Option<usize>) -> Result<Vec<Vec<f64>>, String> {
        let (swap, freq) = (replica_swap_freq.unwrap_or(1), sampling_freq.unwrap_or(1));
        let t = timesteps / freq;
        let mut off = vec![0u32];
        let mut flat: Vec<u32> = Vec::new();
        for p in var_products { flat.extend(p.iter().map(|v| *v as u32)); off.push(flat.len() as u32); }
        let mut ac = vec![0.0; self.n_slots * t];
        check(unsafe { sys::qmcb_pt_spin_product_autocorrelation(self.graph.h, timesteps as u64, swap as u64, freq as u64, var_products.len() as u32,
                                                                 off.as_ptr(), flat.as_ptr(), ac.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut()) })?;
        Ok(ac.chunks(t.max(1)).map(|c| c.to_vec()).collect())
    }
    /// ParallelTemperingBondAutoCorrelations::calculate_bond_autocorrelation (:608-630)
    pub fn calculate_bond_autocorrelation(&mut self, timesteps: usize, replica_swap_freq: Option<usize>, sampling_freq: Option<usize>)
                                          -> Result<Vec<Vec<f64>>, String> {
        let (swap, freq) = (replica_swap_freq.unwrap_or(1), sampling_freq.unwrap_or(1));
        let t = timesteps / freq;
        let mut ac = vec![0.0; self.n_slots * t];
        check(unsafe { sys::qmcb_pt_bond_autocorrelation(self.graph.h, timesteps as u64, swap as u64, freq as u64, ac.as_mut_ptr(),
                                                         std::ptr::null_mut(), std::ptr::null_mut()) })?;
        Ok(ac.chunks(t.max(1)).map(|c| c.to_vec()).collect())
    }
    /// get_total_swaps (tempering_container.rs:231-233)
    pub fn get_total_swaps(&mut self) -> Result<u64, String> {
        let mut s = 0u64;
        check(unsafe { sys::qmcb_pt_total_swaps(self.graph.h, &mut s) })?;
        Ok(s)
    }
    /// the slot each local configuration currently holds (StateGetter / SwapManagers seen from the configuration's side)
    pub fn slots(&mut self) -> Result<Vec<u32>, String> {
        let mut s = vec![0u32; self.graph.replicas];
        check(unsafe { sys::qmcb_pt_get_slots(self.graph.h, s.as_mut_ptr()) })?;
        Ok(s)
    }
}

/// `GraphState` (classical/graph.rs:56-88, :350-447) for R replicas; the schedule is a checkerboard sweep (every site
/// once per sweep) instead of random sites, the per-site rule is the reference's (`do_spin_flip`, `should_flip`).
pub struct GraphState { h: *mut sys::CmcbHandle, nvars: usize, replicas: usize }

impl GraphState {
    /// GraphState::new / new_with_state_and_rng (graph.rs:56-88)
    pub fn new(edges: &[((usize, usize), f64)], biases: &[f64], rng_keys: &[u64], betas: &[f64], state: Option<&[bool]>, device: i32)
               -> Result<Self, String> {
        let va: Vec<u32> = edges.iter().map(|((a, _), _)| *a as u32).collect();
        let vb: Vec<u32> = edges.iter().map(|((_, b), _)| *b as u32).collect();
        let j: Vec<f64> = edges.iter().map(|(_, j)| *j).collect();
        let lat = sys::QmcbLattice { nvars: biases.len() as u32, nedges: edges.len() as u32, va: va.as_ptr(), vb: vb.as_ptr(), j: j.as_ptr(),
                                     transverse: 0.0, longitudinal: 0.0 };
        let init: Option<Vec<u8>> = state.map(|s| (0..rng_keys.len()).flat_map(|_| s.iter().map(|b| *b as u8)).collect());
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::cmcb_create(&lat, biases.as_ptr(), rng_keys.len() as u32, betas.as_ptr(), rng_keys.as_ptr(),
                                        init.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()), device, &mut h) })?;
        Ok(Self { h, nvars: biases.len(), replicas: rng_keys.len() })
    }
    /// `sweeps` checkerboard sweeps of every replica (the throughput path; builder-defined schedule, reference per-site rule)
    pub fn sweeps(&mut self, sweeps: usize) -> Result<(), String> { check(unsafe { sys::cmcb_sweeps(self.h, sweeps as u64) }) }
    /// do_time_step (graph.rs:350-406), the reference's own schedule under each replica's sequential stream: one draw
    /// picks spin flips, edge flips or worm flips.  Returns the move every replica drew.
    pub fn do_time_step(&mut self, nspinupdates: Option<usize>, nedgeupdates: Option<usize>, nwormupdates: Option<usize>,
                        only_basic_moves: Option<bool>) -> Result<Vec<u8>, String> {
        let f = |x: Option<usize>| x.map_or(u64::MAX, |v| v as u64);
        let mut choice = vec![0u8; self.replicas];
        check(unsafe { sys::cmcb_do_time_step(self.h, f(nspinupdates), f(nedgeupdates), f(nwormupdates),
                                              only_basic_moves.unwrap_or(false) as i32, choice.as_mut_ptr()) })?;
        Ok(choice)
    }
    /// do_worm_flip (graph.rs:179-318)
    pub fn do_worm_flip(&mut self, count: usize, allow_doubles: bool) -> Result<(), String> {
        check(unsafe { sys::cmcb_worm_flips(self.h, count as u64, allow_doubles as i32) })
    }
    /// enable_edge_importance_sampling (graph.rs:321-336)
    pub fn enable_edge_importance_sampling(&mut self, enable: bool) -> Result<(), String> {
        check(unsafe { sys::cmcb_enable_edge_importance_sampling(self.h, enable as i32) })
    }
    /// get_energy (graph.rs:430-447)
    pub fn get_energy(&mut self) -> Result<Vec<f64>, String> {
        let mut e = vec![0.0; self.replicas];
        check(unsafe { sys::cmcb_energy(self.h, e.as_mut_ptr()) })?;
        Ok(e)
    }
    pub fn state_ref(&mut self) -> Result<Vec<Vec<bool>>, String> {
        let mut raw = vec![0u8; self.replicas * self.nvars];
        check(unsafe { sys::cmcb_get_states(self.h, raw.as_mut_ptr()) })?;
        Ok(raw.chunks(self.nvars).map(|c| c.iter().map(|b| *b != 0).collect()).collect())
    }
    pub fn set_state(&mut self, states: &[Vec<bool>]) -> Result<(), String> {
        let raw: Vec<u8> = states.iter().flat_map(|s| s.iter().map(|b| *b as u8)).collect();
        assert_eq!(raw.len(), self.replicas * self.nvars);
        check(unsafe { sys::cmcb_set_states(self.h, raw.as_ptr()) })
    }
}
impl Drop for GraphState { fn drop(&mut self) { unsafe { sys::cmcb_destroy(self.h); } } }
