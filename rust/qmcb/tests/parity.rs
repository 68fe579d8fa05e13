//! The parity pin a maintainer can run in one command on a machine with cargo:
//!     QMCB_LIB_DIR=$PWD/isingmontecarlo_b200/_build cargo test -p qmcb --test parity
//! It runs the REAL reference (`qmc::sse::DefaultQmcIsingGraph`) with the injected `PhiloxStream` and checks n, cutoff,
//! spins and the operator string of every STRICT case of tests/golden/sse_golden.json -- the same fixtures
//! `tests/test_golden.py` checks against the CUDA path.  Green here + green there = the CUDA path is bit-exact with the
//! reference's own update path, and the "parity unpinned" label in DESIGN.md 6 can be dropped.
//! A second test does the same for the generic `Qmc` runner with directed-loop updates (tests/golden/qmc_golden.json).
//! SOURCE ONLY in this repository (no Rust toolchain in the build image); needs no GPU.
use qmc::sse::*;
use qmcb::PhiloxStream;
use serde_json::Value;

fn fnv1a64(bytes: impl Iterator<Item = u8>) -> u64 {
    bytes.fold(0xCBF29CE484222325u64, |h, b| (h ^ b as u64).wrapping_mul(0x100000001B3))
}

/// include/qmcb.h operator word: bond | inputs << 24 | outputs << 26, identity = 0xFFFFFFFF
fn op_word<O: Op>(op: Option<&O>) -> u32 {
    match op {
        None => 0xFFFF_FFFF,
        Some(op) => {
            let bits = |s: &[bool]| s.iter().enumerate().fold(0u32, |acc, (k, b)| acc | ((*b as u32) << k));
            op.get_bond() as u32 | (bits(op.get_inputs()) << 24) | (bits(op.get_outputs()) << 26)
        }
    }
}

#[test]
fn reference_reproduces_the_strict_golden_cases() {
    let path = concat!(env!("CARGO_MANIFEST_DIR"), "/../../tests/golden/sse_golden.json");
    let doc: Value = serde_json::from_str(&std::fs::read_to_string(path).unwrap()).unwrap();
    let mut checked = 0;
    for case in doc["cases"].as_array().unwrap() {
        if case["mode"] != "strict" { continue; }  // FAST / COUNTER are builder-defined contracts, not the reference's
        let edges: Vec<((usize, usize), f64)> = case["edges"].as_array().unwrap().iter()
            .map(|e| ((e[0].as_u64().unwrap() as usize, e[1].as_u64().unwrap() as usize), e[2].as_f64().unwrap())).collect();
        let (gamma, h, beta) = (case["gamma"].as_f64().unwrap(), case["h"].as_f64().unwrap(), case["beta"].as_f64().unwrap());
        let (cutoff, sweeps) = (case["cutoff"].as_u64().unwrap() as usize, case["sweeps"].as_u64().unwrap() as usize);
        for rep in case["replicas"].as_array().unwrap() {
            let rng = PhiloxStream { key: rep["key"].as_u64().unwrap(), cursor: 0 };
            // state = None: the spins are drawn from the stream (classical/graph.rs:451-453), as qmcb_create does
            let mut g = DefaultQmcIsingGraph::<PhiloxStream>::new_with_rng(edges.clone(), gamma, h, cutoff, rng, None);
            g.set_enable_heatbath(case["heatbath"].as_bool().unwrap());
            g.set_run_rvb(case["rvb"].as_bool().unwrap_or(false));  // RVB updates (rvb.rs:60-291) in every sweep, qmc_ising.rs:705-752
            let e = g.timesteps(sweeps, beta);
            assert_eq!(g.get_n() as u64, rep["n"].as_u64().unwrap(), "{} n", case["name"]);
            assert_eq!(g.get_cutoff() as u64, rep["cutoff"].as_u64().unwrap(), "{} cutoff", case["name"]);
            let state: String = g.state_ref().iter().map(|b| if *b { '1' } else { '0' }).collect();
            assert_eq!(state, rep["state"].as_str().unwrap(), "{} state", case["name"]);
            let m = g.get_manager_ref();
            let words: Vec<u32> = (0..g.get_cutoff()).map(|p| op_word(m.get_pth(p))).collect();
            let head: Vec<u64> = words.iter().take(8).map(|w| *w as u64).collect();
            let want_head: Vec<u64> = rep["ops_head"].as_array().unwrap().iter().map(|w| w.as_u64().unwrap()).collect();
            assert_eq!(head, want_head, "{} first op words", case["name"]);
            let hash = fnv1a64(words.iter().flat_map(|w| w.to_le_bytes()));
            assert_eq!(format!("{:016x}", hash), rep["ops_fnv1a64"].as_str().unwrap(), "{} operator string", case["name"]);
            assert_eq!(e.to_bits(), python_hex_bits(rep["energy_hex"].as_str().unwrap()), "{} energy", case["name"]);
            assert!(g.verify());
            checked += 1;
        }
    }
    assert!(checked >= 18);  // 8 cases x 3 replicas, two of them with RVB steps
}

/// The generic runner with directed-loop updates: the real `qmc::sse::Qmc` (qmc_runner.rs:22-403, loop update
/// directed_loop.rs:103-301) against tests/golden/qmc_golden.json -- the fixtures `tests/test_golden.py` checks against
/// `qmcb_create_qmc` / `qmcb_loop_update` on the GPU.
#[test]
fn reference_reproduces_the_qmc_golden_cases() {
    let path = concat!(env!("CARGO_MANIFEST_DIR"), "/../../tests/golden/qmc_golden.json");
    let doc: Value = serde_json::from_str(&std::fs::read_to_string(path).unwrap()).unwrap();
    let mut checked = 0;
    for case in doc["cases"].as_array().unwrap() {
        let nvars = case["nvars"].as_u64().unwrap() as usize;
        let (beta, sweeps) = (case["beta"].as_f64().unwrap(), case["sweeps"].as_u64().unwrap() as usize);
        for rep in case["replicas"].as_array().unwrap() {
            let rng = PhiloxStream { key: rep["key"].as_u64().unwrap(), cursor: 0 };
            // Qmc::new draws the state from the stream (qmc_runner.rs:48-51), as qmcb_create_qmc does with init_state = NULL
            let mut q = DefaultQmc::<PhiloxStream>::new(nvars, rng, case["do_loop_updates"].as_bool().unwrap());
            for it in case["interactions"].as_array().unwrap() {
                let mat: Vec<f64> = it["mat"].as_array().unwrap().iter().map(|x| x.as_f64().unwrap()).collect();
                let vars: Vec<usize> = it["vars"].as_array().unwrap().iter().map(|v| v.as_u64().unwrap() as usize).collect();
                if it["diagonal"].as_bool().unwrap() { q.make_diagonal_interaction(mat, vars).unwrap(); } else { q.make_interaction(mat, vars).unwrap(); }
            }
            let e = q.timesteps(sweeps, beta);
            assert_eq!(q.get_n() as u64, rep["n"].as_u64().unwrap(), "{} n", case["name"]);
            assert_eq!(q.get_cutoff() as u64, rep["cutoff"].as_u64().unwrap(), "{} cutoff", case["name"]);
            let state: String = q.state_ref().iter().map(|b| if *b { '1' } else { '0' }).collect();
            assert_eq!(state, rep["state"].as_str().unwrap(), "{} state", case["name"]);
            let m = q.get_manager_ref();
            let words: Vec<u32> = (0..q.get_cutoff()).map(|p| op_word(m.get_pth(p))).collect();
            let hash = fnv1a64(words.iter().flat_map(|w| w.to_le_bytes()));
            assert_eq!(format!("{:016x}", hash), rep["ops_fnv1a64"].as_str().unwrap(), "{} operator string", case["name"]);
            assert_eq!(e.to_bits(), python_hex_bits(rep["energy_hex"].as_str().unwrap()), "{} energy", case["name"]);
            checked += 1;
        }
    }
    assert!(checked >= 12);
}

/// bits of a Python float.hex() literal
fn python_hex_bits(s: &str) -> u64 {
    let (neg, s) = if let Some(r) = s.strip_prefix('-') { (true, r) } else { (false, s) };
    let s = s.strip_prefix("0x").unwrap();
    let (mant, exp) = s.split_once('p').unwrap();
    let (int, frac) = mant.split_once('.').unwrap_or((mant, ""));
    let mut m = u64::from_str_radix(int, 16).unwrap() as f64;
    let mut scale = 1.0 / 16.0;
    for c in frac.chars() { m += c.to_digit(16).unwrap() as f64 * scale; scale /= 16.0; }
    let v = m * 2f64.powi(exp.parse::<i32>().unwrap());
    (if neg { -v } else { v }).to_bits()
}
