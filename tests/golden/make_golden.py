"""Generates tests/golden/sse_golden.json: end states of fixed small runs, produced by the CPU oracle
(oracle/oracle.c).  PROVENANCE: these are NOT outputs of the Rust reference (it cannot be built in this image and
its own tests hold no golden vectors for this path); they freeze the oracle -- which is pinned by the
hand-derived known answers in tests/test_oracle_known_answers.py and by exact diagonalisation -- so that (a) a
change of compiler flags or a later edit cannot silently move it (e.g. fused multiply-adds would shift the
Bernoulli thresholds) and (b) the CUDA path can be checked without the oracle in the loop.
Run from the repo root:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from isingmontecarlo_b200 import lattices  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

CASES = [
    # name, lattice builder (by name, args), gamma, h, cutoff, beta, sweeps, heatbath[, rvb steps (set_run_rvb, qmc_ising.rs:434-441)]
    ("small_qmc_ring", ("small_qmc_ring", []), 1.0, 0.0, 3, 1.0, 25, False),
    ("mixed4x4_h", ("two_d_periodic_mixed", [4]), 1.0, 1.0, 16, 1.0, 20, False),
    ("square8_crit", ("square_periodic", [8, -1.0]), 3.04, 0.0, 64, 4.0, 12, False),
    ("tri6_frustrated_h", ("triangular_periodic", [6, 1.0]), 1.0, 0.2, 36, 2.0, 10, False),
    ("square8_heatbath", ("square_periodic", [8, -1.0]), 3.04, 0.0, 64, 4.0, 12, True),
    ("two_unit_cell_heatbath_h", ("two_unit_cell", []), 1.0, -0.4, 8, 2.0, 20, True),
    ("two_unit_cell_rvb", ("two_unit_cell", []), 1.0, 0.0, 8, 1.0, 25, False, True),          # check_rvb_crash.rs:340-359
    ("tri6_frustrated_rvb_h", ("triangular_periodic", [6, 1.0]), 1.0, 0.2, 36, 2.0, 10, False, True),
]
KEYS = [0x601D0000 + r for r in range(3)]


def digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def fnv1a64(arr):
    """FNV-1a over the little-endian bytes of the op words: a hash the Rust parity kit can restate in three lines"""
    h = 0xCBF29CE484222325
    for b in np.ascontiguousarray(arr).astype("<u4").tobytes():
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def run_case(case, mode):
    name, (builder, args), gamma, h, cutoff, beta, sweeps, hb = case[:8]
    edges = getattr(lattices, builder)(*args)
    out = []
    for k in KEYS:
        g = po.SseOracle(edges, gamma, h, cutoff, key=k)
        g.set_enable_heatbath(hb)
        g.set_run_rvb(len(case) > 8 and case[8])
        e = g.timesteps(sweeps, beta, mode)
        assert g.error == 0 and g.verify()
        ops = g.dump_ops()
        out.append({"key": k, "n": int(g.n), "cutoff": int(g.cutoff), "cursor": int(g.cursor), "energy_hex": float(e).hex(),
                    "state": "".join(str(int(b)) for b in g.state()), "ops_sha256": digest(ops.astype("<u4")),
                    "ops_fnv1a64": "%016x" % fnv1a64(ops), "ops_head": [int(w) for w in ops[:8]]})
    return out


def main():
    doc = {"_provenance": __doc__.split("Run from")[0].strip(), "cases": []}
    for case in CASES:
        for mode, mname in ((po.MODE_STRICT, "strict"), (po.MODE_FAST, "fast"), (po.MODE_COUNTER, "counter")):
            name, (builder, args), gamma, h, cutoff, beta, sweeps, hb = case[:8]
            if hb and mode == po.MODE_COUNTER:
                continue  # the heat-bath rule has no COUNTER-mode contract
            edges = getattr(lattices, builder)(*args)
            doc["cases"].append({"name": name, "builder": builder, "args": args, "gamma": gamma, "h": h, "cutoff": cutoff, "beta": beta,
                                 "sweeps": sweeps, "heatbath": hb, "rvb": bool(len(case) > 8 and case[8]), "mode": mname, "edges": [[a, b, j] for (a, b), j in edges],
                                 "replicas": run_case(case, mode)})
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sse_golden.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", len(doc["cases"]), "cases")


if __name__ == "__main__":
    main()
