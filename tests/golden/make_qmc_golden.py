"""Generates tests/golden/qmc_golden.json: end states of fixed small runs of the generic `Qmc` runner (qmc_runner.rs) with and
without directed-loop updates (directed_loop.rs:103-301), produced by the CPU oracle in STRICT (reference) order.
PROVENANCE: as tests/golden/make_golden.py -- these freeze the oracle, they are not outputs of the Rust reference; the
pin kit (rust/qmcb/tests/parity.rs) checks them against the real `qmc::sse::Qmc` on a machine with cargo.
Run from the repo root:  python tests/golden/make_qmc_golden.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import digest, fnv1a64  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def xxz(d0, d1, x):
    """two-variable matrix in Interaction::at order (out0 out1 in0 in1): diag(00, 01, 10, 11) = d0, d1, d1, d0; 01 <-> 10 = x"""
    m = np.zeros((4, 4))
    m[0, 0] = m[3, 3] = d0
    m[1, 1] = m[2, 2] = d1
    m[1, 2] = m[2, 1] = x
    return [float(v) for v in m.reshape(-1)]


RING = [(xxz(0.4, 1.1, 0.8), [0, 1]), (xxz(0.9, 0.5, 0.6), [1, 2]), (xxz(0.3, 1.0, 1.0), [2, 3]), (xxz(0.7, 0.7, 0.5), [3, 0])]
CHAIN6 = [(xxz(0.2 + 0.1 * k, 1.0, 0.5 + 0.1 * k), [k, k + 1]) for k in range(5)]
CASES = [
    # name, nvars, interactions [(matrix, vars, diagonal)], do_loop_updates, beta, sweeps
    ("exchange_ring_loops_only", 4, [(m, v, False) for m, v in RING], True, 1.2, 40),
    ("exchange_ring_with_sites", 4, [(m, v, False) for m, v in RING] + [([g] * 4, [v], False) for v, g in enumerate([0.5, 1.0, 1.5, 0.8])], True, 1.2, 40),
    ("exchange_chain6_with_sites", 6, [(m, v, False) for m, v in CHAIN6] + [([0.6 + 0.1 * v] * 4, [v], False) for v in range(6)], True, 2.0, 30),
    ("diagonal_ring_with_sites_no_loops", 4, [([0.3, 1.0, 1.0, 0.3], [0, 1], True), ([2.0, 0.5, 0.5, 2.0], [1, 2], True), ([0.2, 0.9, 0.9, 0.2], [2, 3], True),
                                            ([0.6, 1.4, 1.4, 0.6], [3, 0], True)] + [([g] * 4, [v], False) for v, g in enumerate([0.5, 1.0, 1.5, 0.8])], False, 1.5, 30),
]
KEYS = [0x100D0000 + r for r in range(3)]


def build(case, key):
    name, nvars, inters, loops, beta, sweeps = case
    q = po.QmcOracle(nvars, key=key)
    for mat, vs, diagonal in inters:
        (q.make_diagonal_interaction if diagonal else q.make_interaction)(mat, vs)
    q.set_do_loop_updates(loops)
    return q


def main():
    doc = {"_provenance": __doc__.split("Run from")[0].strip(), "cases": []}
    for case in CASES:
        name, nvars, inters, loops, beta, sweeps = case
        reps = []
        for k in KEYS:
            q = build(case, k)
            e = q.timesteps(sweeps, beta, po.MODE_STRICT)
            assert q.error == 0 and q.verify()
            ops = q.dump_ops()
            reps.append({"key": k, "n": int(q.n), "cutoff": int(q.cutoff), "cursor": int(q.cursor), "energy_hex": float(e).hex(),
                         "state": "".join(str(int(b)) for b in q.state()), "ops_sha256": digest(ops.astype("<u4")),
                         "ops_fnv1a64": "%016x" % fnv1a64(ops), "ops_head": [int(w) for w in ops[:8]]})
        doc["cases"].append({"name": name, "nvars": nvars, "do_loop_updates": loops, "beta": beta, "sweeps": sweeps, "mode": "strict",
                             "interactions": [{"mat": m, "vars": v, "diagonal": d} for m, v, d in inters], "replicas": reps})
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "qmc_golden.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", len(doc["cases"]), "cases")


if __name__ == "__main__":
    main()
