"""Generates tests/golden/rand_kat.json: known answers for the rand-0.8 draw -> value mappings the SSE / classical /
tempering path uses, evaluated on the first 64 words of the injected Philox stream with key
0x4B41_5431.  PROVENANCE: produced by the oracle's restatement of rand 0.8 (oracle/oracle.c:89-162), NOT by rand itself
(no Rust toolchain in this image).  The fixture exists so that ONE `cargo test` on a machine with Rust
(rust/qmcb/tests/rand_kat.rs feeds the same words to rand's own Rng methods) pins that restatement and lifts the
"parity unpinned at the rand boundary" label; here it only freezes the oracle.
Run from the repo root:  python tests/golden/make_rand_kat.py"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

KEY, NWORDS = 0x4B415431, 64


def sequential(fn):
    """call fn(cursor_ref) from cursor 0 until the 64 words are used up; returns [(value, cursor_after)]"""
    cur, out = C.c_uint64(0), []
    while True:
        v = fn(cur)
        if cur.value > NWORDS:
            break
        out.append((v, int(cur.value)))
        if cur.value == NWORDS:
            break
    return out


def main():
    L = po.lib()
    words = [int(L.orc_stream_word(KEY, c)) for c in range(NWORDS)]
    doc = {"_provenance": __doc__.split("Run from")[0].strip(), "key": KEY, "words_hex": ["%016x" % w for w in words]}
    doc["gen_bool"] = [{"p_hex": float(p).hex(), "results": [int(L.orc_gen_bool(KEY, C.byref(C.c_uint64(c)), p)) for c in range(NWORDS)]}
                       for p in (0.5, 1.0 / 3.0, 0.0, 1.0 - 2.0 ** -53, 1e-9, 0.9403)]
    doc["gen_bool_one_consumes_no_word"] = True
    doc["gen_range_usize"] = [{"n": n, "calls": [[int(v), c] for v, c in sequential(lambda cur: L.orc_gen_range_usize(KEY, C.byref(cur), n))]}
                              for n in (3, 7, 3072, 11520, 2 ** 33 + 5)]
    doc["gen_range_u8"] = [{"n": n, "calls": [[int(v), c] for v, c in sequential(lambda cur: L.orc_gen_range_u8(KEY, C.byref(cur), n))]}
                           for n in (3, 200)]
    doc["gen_range_f64_unit"] = [[float(v).hex(), c] for v, c in sequential(lambda cur: L.orc_gen_range_f64_01(KEY, C.byref(cur)))]
    doc["gen_range_f64"] = [{"low_hex": float(lo).hex(), "high_hex": float(hi).hex(),
                             "calls": [[float(v).hex(), c] for v, c in sequential(lambda cur: L.orc_gen_range_f64(KEY, C.byref(cur), lo, hi))]}
                            for lo, hi in ((0.0, 3.7), (0.0, 12288.0 * 2.0))]
    doc["gen_f64"] = [float(L.orc_gen_f64(KEY, C.byref(C.c_uint64(c)))).hex() for c in range(NWORDS)]
    doc["gen_std_bool"] = [int(L.orc_gen_std_bool(KEY, C.byref(C.c_uint64(c)))) for c in range(NWORDS)]
    doc["powi"] = [{"a_hex": float(a).hex(), "b": b, "value_hex": float(L.orc_powi(a, b)).hex()}
                   for a, b in ((0.97, 13), (1.0309278350515463, -7), (0.5, 0), (1.25, 1), (0.999, -1000), (2.0, 31))]
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rand_kat.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote rand_kat.json")


if __name__ == "__main__":
    main()
