//! Pins the oracle's restatement of rand 0.8 (oracle/oracle.c:89-162) against rand itself:
//! the 64 words of tests/golden/rand_kat.json are fed to rand's own `Rng` methods through a scripted `RngCore`
//! and every value / consumed-word count must equal the fixture the oracle produced.
//! SOURCE ONLY here (no Rust toolchain in the build image).  Run: `cargo test -p qmcb --test rand_kat`.
use rand::Rng;
use rand_core::{impls, Error, RngCore};
use serde_json::Value;

/// One 64-bit word per call, a u32 is the word's high half: the contract of `qmcb::PhiloxStream`.
struct Scripted { words: Vec<u64>, cursor: usize }
impl RngCore for Scripted {
    fn next_u64(&mut self) -> u64 { let w = self.words[self.cursor]; self.cursor += 1; w }
    fn next_u32(&mut self) -> u32 { (self.next_u64() >> 32) as u32 }
    fn fill_bytes(&mut self, dest: &mut [u8]) { impls::fill_bytes_via_next(self, dest) }
    fn try_fill_bytes(&mut self, dest: &mut [u8]) -> Result<(), Error> { self.fill_bytes(dest); Ok(()) }
}

fn fixture() -> (Value, Vec<u64>) {
    let path = concat!(env!("CARGO_MANIFEST_DIR"), "/../../tests/golden/rand_kat.json");
    let doc: Value = serde_json::from_str(&std::fs::read_to_string(path).unwrap()).unwrap();
    let words = doc["words_hex"].as_array().unwrap().iter().map(|w| u64::from_str_radix(w.as_str().unwrap(), 16).unwrap()).collect();
    (doc, words)
}
fn hexf(v: &Value) -> f64 { parse_hex_f64(v.as_str().unwrap()) }
/// Python's float.hex(): [-]0x1.<13 hex digits>p<exp>, or 0x0.0p+0
fn parse_hex_f64(s: &str) -> f64 {
    let (neg, s) = if let Some(r) = s.strip_prefix('-') { (true, r) } else { (false, s) };
    let s = s.strip_prefix("0x").unwrap();
    let (mant, exp) = s.split_once('p').unwrap();
    let (int, frac) = mant.split_once('.').unwrap_or((mant, ""));
    let mut m = u64::from_str_radix(int, 16).unwrap() as f64;
    let mut scale = 1.0 / 16.0;
    for c in frac.chars() { m += c.to_digit(16).unwrap() as f64 * scale; scale /= 16.0; }
    let v = m * 2f64.powi(exp.parse::<i32>().unwrap());
    if neg { -v } else { v }
}

#[test]
fn gen_bool_matches() {
    let (doc, words) = fixture();
    for case in doc["gen_bool"].as_array().unwrap() {
        let p = hexf(&case["p_hex"]);
        for (c, want) in case["results"].as_array().unwrap().iter().enumerate() {
            let mut rng = Scripted { words: words.clone(), cursor: c };
            assert_eq!(rng.gen_bool(p) as u64, want.as_u64().unwrap(), "gen_bool({}) at word {}", p, c);
            assert_eq!(rng.cursor, c + 1);
        }
    }
    let mut rng = Scripted { words: words.clone(), cursor: 0 };
    assert!(rng.gen_bool(1.0));
    assert_eq!(rng.cursor, 0, "gen_bool(1.0) must not consume a word");
}

#[test]
fn gen_range_matches() {
    let (doc, words) = fixture();
    for case in doc["gen_range_usize"].as_array().unwrap() {
        let n = case["n"].as_u64().unwrap() as usize;
        let mut rng = Scripted { words: words.clone(), cursor: 0 };
        for call in case["calls"].as_array().unwrap() {
            assert_eq!(rng.gen_range(0..n) as u64, call[0].as_u64().unwrap(), "gen_range(0..{})", n);
            assert_eq!(rng.cursor as u64, call[1].as_u64().unwrap(), "words consumed by gen_range(0..{})", n);
        }
    }
    for case in doc["gen_range_u8"].as_array().unwrap() {
        let n = case["n"].as_u64().unwrap() as u8;
        let mut rng = Scripted { words: words.clone(), cursor: 0 };
        for call in case["calls"].as_array().unwrap() {
            assert_eq!(rng.gen_range(0..n) as u64, call[0].as_u64().unwrap());
            assert_eq!(rng.cursor as u64, call[1].as_u64().unwrap());
        }
    }
    let mut rng = Scripted { words: words.clone(), cursor: 0 };
    for call in doc["gen_range_f64_unit"].as_array().unwrap() {
        let v: f64 = rng.gen_range(0. ..1.0);
        assert_eq!(v.to_bits(), hexf(&call[0]).to_bits());
        assert_eq!(rng.cursor as u64, call[1].as_u64().unwrap());
    }
    for case in doc["gen_range_f64"].as_array().unwrap() {
        let (lo, hi) = (hexf(&case["low_hex"]), hexf(&case["high_hex"]));
        let mut rng = Scripted { words: words.clone(), cursor: 0 };
        for call in case["calls"].as_array().unwrap() {
            let v: f64 = rng.gen_range(lo..hi);
            assert_eq!(v.to_bits(), hexf(&call[0]).to_bits());
            assert_eq!(rng.cursor as u64, call[1].as_u64().unwrap());
        }
    }
}

#[test]
fn standard_draws_match() {
    let (doc, words) = fixture();
    for (c, want) in doc["gen_f64"].as_array().unwrap().iter().enumerate() {
        let mut rng = Scripted { words: words.clone(), cursor: c };
        assert_eq!(rng.gen::<f64>().to_bits(), hexf(want).to_bits());
    }
    for (c, want) in doc["gen_std_bool"].as_array().unwrap().iter().enumerate() {
        let mut rng = Scripted { words: words.clone(), cursor: c };
        assert_eq!(rng.gen::<bool>() as u64, want.as_u64().unwrap());
    }
    for case in doc["powi"].as_array().unwrap() {  // f64::powi = compiler-rt __powidf2 (tempering_container.rs:294)
        let v = hexf(&case["a_hex"]).powi(case["b"].as_i64().unwrap() as i32);
        assert_eq!(v.to_bits(), hexf(&case["value_hex"]).to_bits());
    }
}

#[test]
fn philox_stream_is_the_fixture_stream() {
    let (doc, words) = fixture();
    let mut s = qmcb::PhiloxStream { key: doc["key"].as_u64().unwrap(), cursor: 0 };
    for w in words { assert_eq!(s.next_u64(), w); }
}
