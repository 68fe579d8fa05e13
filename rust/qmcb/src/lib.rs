//! Safe wrappers with the shape of `qmc::sse::QmcIsingGraph` / `QmcStepper`
//! (qmc_ising.rs:131-148, qmc_stepper.rs:2-168) over one batched GPU handle.
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

/// R replicas of one lattice; method names follow `QmcStepper`.
pub struct BatchedQmcIsingGraph { h: *mut sys::QmcbHandle, nvars: usize, replicas: usize }

impl BatchedQmcIsingGraph {
    /// qmc_ising.rs:131-148 with one rng key and beta per replica.
    pub fn new_with_rng(edges: Vec<((usize, usize), f64)>, transverse: f64, longitudinal: f64, cutoff: usize,
                        rng_keys: &[u64], betas: &[f64], state: Option<Vec<bool>>) -> Result<Self, String> {
        let nvars = edges.iter().map(|((a, b), _)| *a.max(b)).max().unwrap() + 1;
        let va: Vec<u32> = edges.iter().map(|((a, _), _)| *a as u32).collect();
        let vb: Vec<u32> = edges.iter().map(|((_, b), _)| *b as u32).collect();
        let j: Vec<f64> = edges.iter().map(|(_, j)| *j).collect();
        let lat = sys::QmcbLattice { nvars: nvars as u32, nedges: edges.len() as u32, va: va.as_ptr(), vb: vb.as_ptr(),
                                     j: j.as_ptr(), transverse, longitudinal };
        let init: Option<Vec<u8>> = state.map(|s| (0..rng_keys.len()).flat_map(|_| s.iter().map(|b| *b as u8)).collect());
        let mut h = std::ptr::null_mut();
        check(unsafe { sys::qmcb_create(&lat, rng_keys.len() as u32, betas.as_ptr(), rng_keys.as_ptr(), cutoff as u64, 0,
                                        init.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()), 0, &mut h) })?;
        Ok(Self { h, nvars, replicas: rng_keys.len() })
    }
    /// QmcStepper::timesteps (qmc_stepper.rs:17-20): average energy per replica.
    pub fn timesteps(&mut self, t: usize) -> Result<Vec<f64>, String> {
        let mut e = vec![0.0; self.replicas];
        check(unsafe { sys::qmcb_timesteps(self.h, t as u64, 1, e.as_mut_ptr(), std::ptr::null_mut()) })?;
        Ok(e)
    }
    /// QmcStepper::timesteps_sample (qmc_stepper.rs:23-40): samples[replica][k][var].
    pub fn timesteps_sample(&mut self, t: usize, sampling_freq: Option<usize>) -> Result<(Vec<Vec<Vec<bool>>>, Vec<f64>), String> {
        let f = sampling_freq.unwrap_or(1);
        let k = t / f;
        let mut e = vec![0.0; self.replicas];
        let mut raw = vec![0u8; self.replicas * k * self.nvars];
        check(unsafe { sys::qmcb_timesteps(self.h, t as u64, f as u64, e.as_mut_ptr(), raw.as_mut_ptr()) })?;
        let s = raw.chunks(k * self.nvars).map(|r| r.chunks(self.nvars).map(|c| c.iter().map(|b| *b != 0).collect()).collect()).collect();
        Ok((s, e))
    }
    /// QmcIsingGraph::set_enable_heatbath (qmc_ising.rs:444-486)
    pub fn set_enable_heatbath(&mut self, enable_heatbath: bool) -> Result<(), String> {
        check(unsafe { sys::qmcb_set_enable_heatbath(self.h, enable_heatbath as i32) })
    }
    /// serde replacement (SerializeQmcGraph, qmc_ising.rs:1001-1087): the whole batch, stream position included
    pub fn to_bytes(&mut self) -> Result<Vec<u8>, String> {
        let mut n = 0u64;
        check(unsafe { sys::qmcb_checkpoint_size(self.h, &mut n) })?;
        let mut buf = vec![0u8; n as usize];
        check(unsafe { sys::qmcb_checkpoint_save(self.h, buf.as_mut_ptr() as *mut _, n) })?;
        Ok(buf)
    }
    /// QmcStepper::imaginary_time_fold (qmc_stepper.rs:165-168) with the closure evaluated on the host
    pub fn imaginary_time_fold<F, T>(&mut self, r: usize, fold_fn: F, init: T) -> Result<T, String>
    where F: Fn(T, &[bool]) -> T {
        let mut cut = vec![0u64; self.replicas];
        check(unsafe { sys::qmcb_get_cutoffs(self.h, cut.as_mut_ptr()) })?;
        let (mut acc, mut raw) = (init, vec![0u8; self.nvars]);
        for p in 0..cut[r] {
            check(unsafe { sys::qmcb_itime_state(self.h, r as u32, p, raw.as_mut_ptr()) })?;
            let st: Vec<bool> = raw.iter().map(|b| *b != 0).collect();
            acc = fold_fn(acc, &st);
        }
        Ok(acc)
    }
    pub fn get_n(&mut self) -> Result<Vec<u64>, String> {
        let mut n = vec![0u64; self.replicas];
        check(unsafe { sys::qmcb_get_n(self.h, n.as_mut_ptr()) })?;
        Ok(n)
    }
    pub fn verify(&mut self, r: usize) -> Result<bool, String> {
        let mut ok = 0;
        check(unsafe { sys::qmcb_verify(self.h, r as u32, &mut ok) })?;
        Ok(ok != 0)
    }
}
impl Drop for BatchedQmcIsingGraph { fn drop(&mut self) { unsafe { sys::qmcb_destroy(self.h); } } }
